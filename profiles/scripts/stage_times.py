"""Per-stage device time of one C1 step (183 x 131072) for the fast modes with the FFMA2 and the tensor-core decimator.
   python profiles/scripts/stage_times.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import __graft_entry__ as entry
entry.build()
from tetraear_b200 import _lib, pipeline, synth
CH, NC = 131072, 183
x = synth.make_capture(NC * CH, 130, seed=1234).reshape(NC, CH)
dev = torch.from_numpy(x.view(np.float32).reshape(NC, CH, 2)).to("cuda:0")
for mode in ("fast", "fast32"):
    for tc in (0,):
        pipe = pipeline.ReceivePipeline(2.4e6, CH, max_chunks=NC, mode=mode, device=0, n_lanes=1)
        for h in pipe.handles:
            h.set_option(_lib.OPT_DECIM_TC, tc)
        for _ in range(3):
            pipe.run_device(dev)
        torch.cuda.synchronize()
        pipe.profile(True)
        for _ in range(20):
            pipe.run_device(dev)
        torch.cuda.synchronize()
        prof = pipe.profile_read()
        pipe.profile(False)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            pipe.run_device(dev)
        b.record(); torch.cuda.synchronize()
        print(mode, "tc" if tc else "ffma2", {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]},
              "step(1 lane, python-driven) ms:", round(a.elapsed_time(b) / 50, 4), "crc_pass", int(pipe.stats[1].item()) // 73)
