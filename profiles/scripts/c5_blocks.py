import sys, numpy as np, torch
sys.path.insert(0, '.')  # run from the repo root
from tetraear_b200 import etsi
from tetraear_b200.replay import ArchiveReplay
M, n = 800, 4_000_000
dev = torch.device("cuda:0")
wide = 0.02 * torch.randn((n, 2), device=dev, dtype=torch.float32)
bits_s, _, init = etsi.make_etsi_stream(14, seed=3)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for nb in (1, 4, 8, 16, 32):
    rp = ArchiveReplay(M, n, n_blocks=nb)
    rp.chan.run_device(wide, n, rp.rows)
    print("blocks", nb, rp.n_blocks, "receive ms", round(t(lambda: rp.receive_rows(rp.nf, init)), 3), "chan ms", round(t(lambda: rp.chan.run_device(wide, n, rp.rows)), 3))
