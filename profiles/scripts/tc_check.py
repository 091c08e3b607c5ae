"""GPU check + timing of the tensor-core decimator (dsp_tc.cu) against the FFMA2 one (dsp_fast.cu) and scipy.
   python profiles/scripts/tc_check.py [n_chunks]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from scipy import signal
import __graft_entry__ as entry
entry.build()
from tetraear_b200 import _lib

q, CH = 10, 131072
n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 6
rng = np.random.default_rng(3)
x = ((rng.standard_normal((n_chunks, CH)) + 1j * rng.standard_normal((n_chunks, CH))) * 0.25
     + 0.5 * np.exp(2j * np.pi * 0.004 * np.arange(CH))[None, :]).astype(np.complex64)
sos32 = np.zeros(24, np.float32); d64 = np.zeros(24); z64 = np.zeros(8); z32 = np.zeros(8, np.float32)
_lib.check(_lib.load().tb200_design_decimator(q, d64.ctypes.data, z64.ctypes.data, sos32.ctypes.data, z32.ctypes.data))
ref = np.stack([signal.sosfiltfilt(sos32.astype(np.float64).reshape(4, 6), c.astype(np.complex128))[::q] for c in x])
sci = np.stack([signal.decimate(c, q) for c in x])
scale = np.abs(ref).max()
dev = torch.device("cuda:0")
xd = torch.from_numpy(x.view(np.float32).reshape(n_chunks, CH, 2)).to(dev)
n_dec = (CH + q - 1) // q
stride = n_dec + 8
res = {}
for tc in (0, 1, 2):
    h = _lib.Handle(sample_rate=2.4e6, max_chunks=n_chunks, max_chunk_len=CH, mode=_lib.MODE_FAST, device=0)
    h.set_option(_lib.OPT_DECIM_TC, tc)
    dib = torch.zeros((n_chunks, stride), dtype=torch.uint8, device=dev)
    cnt = torch.zeros((n_chunks,), dtype=torch.int32, device=dev)
    ph = torch.zeros((n_chunks,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def run():
        _lib.check(h.lib.tb200_process_batch(h.h, xd.data_ptr(), _lib.DTYPE_C64, n_chunks, CH, CH, None, dib.data_ptr(), None,
                                             stride, cnt.data_ptr(), ph.data_ptr(), st))
    run(); torch.cuda.synchronize()
    out = np.zeros((n_chunks, n_dec), np.complex64)
    _lib.check(h.lib.tb200_debug_read(h.h, 0, out.ctypes.data, out.nbytes))
    err = np.abs(out - ref)
    print(f"tc={tc}: max |y - exact| / full scale = {err.max() / scale:.3e} (body {err[:, :-200].max() / scale:.3e}, tail {err[:, -200:].max() / scale:.3e}, "
          f"head {err[:, :200].max() / scale:.3e}); vs scipy float32 {np.abs(out - sci).max() / scale:.3e}; worst chunk {err.max(axis=1).argmax()} at k={err.max(axis=0).argmax()}")
    res[tc] = (out, dib.cpu().numpy().copy(), cnt.cpu().numpy().copy())
    import ctypes
    tm = (ctypes.c_int64 * 16)()
    if tc:
        _lib.check(h.lib.tb200_debug_tc_timing(h.h, 1, None))
        for _ in range(5):
            run()
        _lib.check(h.lib.tb200_debug_tc_timing(h.h, 0, tm))
        t = list(tm)
        g = max(t[15], 1)
        if tc == 1:
            names = ["top", "input wait", "slot read", "relayout+split", "mma issue", "mma wait", "tmem ld", "scans", "carries+store"]
            print("   TC phases, cycles per group (CTA 0, thread 0):", {n: round(t[i] / g) for i, n in enumerate(names)}, "groups", t[15] // 5)
        else:
            names = {0: "loader: wait raw free", 1: "relayout: wait raw full", 2: "relayout: wait A free", 3: "relayout: work", 4: "mma: wait A full / D free",
                     5: "mma: issue", 7: "scan team 0: wait D full", 8: "scan: tmem ld", 9: "scan: KS + publish", 10: "scan: wait neighbours",
                     11: "scan: carries + stores"}
            print("   tc4 stages, cycles per group of CTA 0 (scan team 0 sees every TC4_NS-th group):", {n: round(t[i] / g) for i, n in names.items()}, "groups", t[15] // 5)
    h.profile(True)
    for _ in range(20):
        run()
    torch.cuda.synchronize()
    prof = h.profile_read()
    print("   stage ms:", {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]})
print("tc vs fast decimated: max diff / scale", np.abs(res[0][0] - res[1][0]).max() / scale)
print("dibits equal:", np.array_equal(res[0][1], res[1][1]), "counts", res[0][2][:4], res[1][2][:4])
