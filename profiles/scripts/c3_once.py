"""C3 reference mode once: 1M bursts x 255 dibits -> type / data bits / CRC (tb200_burst_batch), for ncu.
   python profiles/scripts/c3_once.py [repeats]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as entry
entry.build()
from tetraear_b200 import _lib, synth
from tetraear_b200.core.protocol import TetraProtocolParser
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nb = 1_000_000
rows = torch.from_numpy(synth.make_burst_batch(nb, seed=99)).cuda()
rec = torch.empty((nb, 32), dtype=torch.uint8, device="cuda")
st = torch.zeros((3,), dtype=torch.int64, device="cuda")
p = TetraProtocolParser(device=0)
h = p._handle()
stream = torch.cuda.current_stream().cuda_stream
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    if i == reps - 1: a.record()
    _lib.check(h.lib.tb200_burst_batch(h.h, rows.data_ptr(), nb, rec.data_ptr(), None, st.data_ptr(), stream))
b.record(); torch.cuda.synchronize()
print("burst_batch ms", round(a.elapsed_time(b), 4), "crc_pass", int(rec[:, 28].sum().item()))
