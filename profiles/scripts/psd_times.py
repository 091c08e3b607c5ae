import sys, numpy as np, torch
sys.path.insert(0, '.')  # run from the repo root
from tetraear_b200.signal import SpectrumAnalyzer
dev = torch.device("cuda:0")
n = 144_000_000; nfr = n // 4096
iq = 0.1 * torch.randn((n, 2), device=dev, dtype=torch.float32)
power = torch.empty((nfr, 4096), dtype=torch.float32, device=dev); smooth = torch.empty_like(power)
stats = torch.empty((nfr, 8), dtype=torch.float64, device=dev)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for prec in ("f64", "f32"):
    an = SpectrumAnalyzer(2.4e6, 4096, precision=prec)
    print(prec, "psd only", t(lambda: an.run_device(iq, nfr, 4096, denoise=False, power=power, stats=stats)))
    print(prec, "psd no stats", t(lambda: an.run_device(iq, nfr, 4096, denoise=False, want_stats=False, power=power)))
    print(prec, "psd+ema", t(lambda: an.run_device(iq, nfr, 4096, denoise=True, power=power, smoothed=smooth, stats=stats)))
# EMA exactness on a long run: compare with sequential numpy on a few bins
an = SpectrumAnalyzer(2.4e6, 4096)
p, s2, st = an.run_device(iq, 3000, 4096, denoise=True)
P = p.cpu().numpy().astype(np.float64); S = s2.cpu().numpy().astype(np.float64)
ref = np.empty_like(P[:, :64]); acc = P[0, :64].copy(); ref[0] = acc
for i in range(1, 3000):
    acc = acc * (1 - 0.15) + P[i, :64] * 0.15; ref[i] = acc
print("ema max err", np.abs(S[:, :64] - ref).max())
p, s3, st = an.run_device(iq[4096*3000:], 500, 4096, denoise=True)     # continuation uses the carried state
P2 = p.cpu().numpy().astype(np.float64); S3 = s3.cpu().numpy().astype(np.float64)
for i in range(500):
    acc = acc * (1 - 0.15) + P2[i, :64] * 0.15
    if i in (0, 1, 499): print("cont err", i, np.abs(S3[i, :64] - acc).max())
