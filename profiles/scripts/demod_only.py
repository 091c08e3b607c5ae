"""etsi demodulator alone: 800 streams x 10000 samples at 50 kHz (what the channelizer hands over for 0.2 s of C4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tetraear_b200 import etsi

M, nf = 800, 10000
bits_s, _, init = etsi.make_etsi_stream(14, seed=3)
base = etsi.modulate_dqpsk(bits_s, 50000.0, sigma=0.02)
L = min(len(base), nf)
y = torch.zeros((M, nf, 2), dtype=torch.float32, device="cuda")
y[:, :L] = torch.from_numpy(base.view(np.float32).reshape(-1, 2))[:L].cuda()
rx = etsi.EtsiReceiver(50000.0, max_streams=M, max_samples=nf)
for _ in range(2):
    rx.run_device(y, nf, init)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    rx.run_device(y, nf, init)
b.record(); torch.cuda.synchronize()
print(f"receiver: {a.elapsed_time(b) / 5:.3f} ms per 800 x {nf} samples; bursts {int(rx.n_bursts.sum())} crc_ok {int(rx.crc_ok.sum())}")
