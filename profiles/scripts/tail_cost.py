"""How much of the C1 step do the latency-bound tail kernels cost once the lanes overlap them?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tetraear_b200 import pipeline, synth

n_chunks, CH = 183, 131072
x = synth.make_capture(CH * 8, 130, seed=1234)
x = np.tile(x.reshape(8, CH), (23, 1))[:n_chunks]
dev = torch.from_numpy(x.view(np.float32).reshape(n_chunks, CH, 2)).cuda()
for lanes in (1, 2, 3, 4, 6):
    pipe = pipeline.ReceivePipeline(2.4e6, CH, max_chunks=n_chunks, mode="fast32", n_lanes=lanes)
    for dec in (True, False):
        for _ in range(5):
            pipe.run_device(dev, None, decode=dec)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100):
            pipe.run_device(dev, None, decode=dec, join=False)
        pipe.join()
        b.record()
        torch.cuda.synchronize()
        print(f"lanes {lanes} decode={dec}: {a.elapsed_time(b) / 100:.4f} ms/step")
