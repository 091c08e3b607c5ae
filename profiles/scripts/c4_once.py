"""One reference-mode wideband chunk (2^20 samples @20 MS/s -> 800 channels -> dibits -> frames), for ncu.
   python profiles/scripts/c4_once.py [repeats]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import __graft_entry__ as entry
entry.build()
from tetraear_b200 import synth
from tetraear_b200.multichannel import WidebandReceiver
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
x = synth.make_wideband(1 << 20, [0, 7, -37, 399], 20e6, 800, 1079, seed=5)
rx = WidebandReceiver(20e6, 800, 1 << 20, mode="fast32", keep_symbols=False)
d = torch.from_numpy(x.view(np.float32).reshape(-1, 2)).cuda()
for _ in range(reps):
    rx.run_device(d)
torch.cuda.synchronize()
print("frames", int(rx.n_frames.sum().item()), "crc_pass", int(rx.stats[1].item()))
