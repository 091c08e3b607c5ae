"""One C1 step (183 x 131072, single lane) per mode, for ncu captures:  python profiles/scripts/one_step.py fast fast32"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import __graft_entry__ as entry
entry.build()
from tetraear_b200 import pipeline, synth
CH, NC = 131072, 183
x = synth.make_capture(NC * CH, 130, seed=1234).reshape(NC, CH)
dev = torch.from_numpy(x.view(np.float32).reshape(NC, CH, 2)).to("cuda:0")
for mode in sys.argv[1:] or ["fast"]:
    pipe = pipeline.ReceivePipeline(2.4e6, CH, max_chunks=NC, mode=mode, device=0, n_lanes=1)
    for _ in range(2):
        pipe.run_device(dev)
    torch.cuda.synchronize()
    print(mode, "crc_pass", int(pipe.stats[1].item()) // 2)
