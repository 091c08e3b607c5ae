"""Device-resident C1 step with and without a per-chunk NCO offset (frequency_shift active), per mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tetraear_b200 import pipeline, synth

n_chunks, CH = 183, 131072
x = synth.make_capture(CH * 8, 130, seed=1234)
x = np.tile(x.reshape(8, CH), (23, 1))[:n_chunks]
dev = torch.from_numpy(x.view(np.float32).reshape(n_chunks, CH, 2)).cuda()
f = torch.full((n_chunks,), 1234.5, dtype=torch.float64, device="cuda")
for mode in ("fast", "fast32"):
    pipe = pipeline.ReceivePipeline(2.4e6, CH, max_chunks=n_chunks, mode=mode, n_lanes=2)
    for name, fo in (("no offset", None), ("offset 1234.5 Hz", f)):
        for _ in range(5):
            pipe.run_device(dev, fo)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            pipe.run_device(dev, fo, join=False)
        pipe.join()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 50
        print(f"{mode:7s} {name:18s} {ms:.4f} ms/step  {n_chunks * CH / ms / 1e3:.0f} Msamples/s")
