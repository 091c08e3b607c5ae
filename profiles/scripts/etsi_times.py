import sys, numpy as np, torch
sys.path.insert(0, '.')  # run from the repo root
from tetraear_b200 import etsi, _lib, _dev
from tetraear_b200.replay import ArchiveReplay
M, n = 800, 4_000_000
dev = torch.device("cuda:0")
wide = 0.02 * torch.randn((n, 2), device=dev, dtype=torch.float32)
rp = ArchiveReplay(M, n)
bits_s, _, init = etsi.make_etsi_stream(14, seed=3)
rx = rp.rx
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print("channelizer ms", t(lambda: rp.chan.run_device(wide, n, rp.rows)))
h, lib = rx._h, rx._h.lib
st = _dev.stream_ptr(0)
print("demod ms", t(lambda: _lib.check(lib.tb200_etsi_demod_batch(h.h, rp.rows.data_ptr(), M, rp.nf, rp.nf, 50000.0, None, rx.soft.data_ptr(), rx.soft_stride, rx.n_soft.data_ptr(), st))))
print("find ms", t(lambda: _lib.check(lib.tb200_etsi_find_bursts(h.h, rx.soft.data_ptr(), M, rx.soft_stride, rx.n_soft.data_ptr(), 0.8, rx.metric.data_ptr(), rx.blocks.data_ptr(), rx.burst_pos.data_ptr(), rx.n_bursts.data_ptr(), rx.max_bursts, st))))
nblk = M * rx.max_bursts * 2
print("decode ms", nblk, t(lambda: _lib.check(lib.tb200_etsi_decode_batch(h.h, rx.blocks.data_ptr(), nblk, 0, init, rx.bits.data_ptr(), rx.crc_ok.data_ptr(), rx.path_metric.data_ptr(), st))))
print("metric zero ms", t(lambda: rx.metric.zero_()))
