"""Device-resident timings of tb200_detect_batch (64 dwells x 262144) and tb200_voice_blocks (1M frames)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tetraear_b200 import _lib
from tetraear_b200.signal import TetraSignalDetector
from tetraear_b200.voice import VoiceBlockBuilder

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev); g.manual_seed(1)
nd, n = 64, 262144
dw = 0.05 * torch.randn((nd, n, 2), generator=g, device=dev, dtype=torch.float32)
det = TetraSignalDetector(2.4e6)
det.stats_batch(dw, 5)
h = det._handle(nd * n)
rec = torch.empty((nd, _lib.DETECT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
ms = timed(lambda: _lib.check(h.lib.tb200_detect_batch(h.h, dw.data_ptr(), _lib.DTYPE_C64, nd, n, n, 2.4e6, 5, rec.data_ptr(), stream)))
print(f"detect 64x262144: {ms:.4f} ms  {nd * n / ms / 1e3:.0f} Msamples/s  {nd * n * 16 / ms / 1e6:.0f} GB/s")
nv = 1_000_000
vb = VoiceBlockBuilder()
rows = torch.randint(0, 2, (nv, 432), generator=g, device=dev, dtype=torch.uint8)
blocks = torch.empty((nv, 690), dtype=torch.int16, device=dev)
vld = torch.empty((nv,), dtype=torch.uint8, device=dev)
ms = timed(lambda: _lib.check(vb._h.lib.tb200_voice_blocks(vb._h.h, _lib.VOICE_BITS, rows.data_ptr(), 432, None, None, None, nv, 127.0,
                                                           blocks.data_ptr(), vld.data_ptr(), stream)))
print(f"voice 1M frames: {ms:.4f} ms  {nv * (432 + 1380) / ms / 1e6:.0f} GB/s")
