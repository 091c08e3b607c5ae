"""Host-side enqueue cost of one C1 step (how far ahead of the GPU the Python driver runs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tetraear_b200 import pipeline, synth

n_chunks, CH = 183, 131072
x = synth.make_capture(CH * 8, 130, seed=1234)
x = np.tile(x.reshape(8, CH), (23, 1))[:n_chunks]
dev = torch.from_numpy(x.view(np.float32).reshape(n_chunks, CH, 2)).cuda()
for lanes in (1, 2, 3, 4):
    pipe = pipeline.ReceivePipeline(2.4e6, CH, max_chunks=n_chunks, mode="fast32", n_lanes=lanes)
    for _ in range(5):
        pipe.run_device(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        pipe.run_device(dev, join=False)
    t1 = time.perf_counter()
    pipe.join()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"lanes {lanes}: host enqueue {1e3 * (t1 - t0) / 200:.4f} ms/step, wall {1e3 * (t2 - t0) / 200:.4f} ms/step")
