"""Where the per-chunk class-API call (SignalProcessor.process + TetraDecoder.decode, host numpy in and out) spends its time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tetraear_b200 import synth, _lib
from tetraear_b200.signal import SignalProcessor
from tetraear_b200.core import TetraDecoder

CH = 131072
x = synth.make_capture(CH * 8, 130, seed=1234).reshape(8, CH)
for mode in ("fast", "fast32"):
    sp, dec = SignalProcessor(2.4e6, mode=mode), TetraDecoder(auto_decrypt=False)
    for c in range(8):
        dec.decode(sp.process(x[c], 0.0))
    N = 200
    t0 = time.perf_counter()
    for i in range(N):
        d = sp.process(x[i % 8], 0.0)
    t1 = time.perf_counter()
    for i in range(N):
        f = dec.decode(d)
    t2 = time.perf_counter()
    h = sp._h
    h.profile(True) if hasattr(h, "profile") else None
    for i in range(20):
        sp.process(x[i % 8], 0.0)
    prof = h.profile_read() if hasattr(h, "profile_read") else {}
    # bare copies of the same sizes
    dx = torch.empty(CH * 2, dtype=torch.float32, device="cuda")
    hx = torch.from_numpy(x[0].view(np.float32))
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    for i in range(N):
        dx.copy_(hx); torch.cuda.synchronize()
    t4 = time.perf_counter()
    print(f"{mode}: process {1e3 * (t1 - t0) / N:.4f} ms, decode {1e3 * (t2 - t1) / N:.4f} ms ({len(f)} frames), "
          f"bare pageable 1 MB H2D + sync {1e3 * (t4 - t3) / N:.4f} ms, device stages {prof}")
