"""Device-resident timing of the ETSI channel decoder (tb200_etsi_decode_batch) per logical channel, 1M blocks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tetraear_b200 import etsi

nb, pool = 1_000_000, 4096
rng = np.random.default_rng(7)
for ch in ("SCH/HD", "SCH/F", "BSCH", "SCH/HU", "TCH/4.8", "TCH/2.4", (124, 144, 432, 103, "1/3", True), (124, 144, 216, 101, "2/3", True)):
    K1 = etsi.CHANNELS[ch][0] if isinstance(ch, str) else ch[0]
    init = 3 if ch == "BSCH" else etsi.scrambling_init(262, 42, 17)
    t1 = rng.integers(0, 2, (pool, K1), dtype=np.uint8)
    soft = etsi.to_soft(etsi.encode(t1, ch, init) if isinstance(ch, str) else etsi.encode_custom(t1, *ch, scramb_init=init), 0.5, rng)
    d = torch.from_numpy(soft).cuda().repeat(nb // pool + 1, 1)[:nb].contiguous()
    dec = etsi.ChannelDecoder(ch)
    bits = torch.empty((nb, 40), dtype=torch.uint8, device="cuda")
    ok = torch.empty((nb,), dtype=torch.uint8, device="cuda")
    met = torch.empty((nb,), dtype=torch.int32, device="cuda")
    for _ in range(2):
        dec.decode_device(d, init, bits, ok, met)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(5):
        dec.decode_device(d, init, bits, ok, met)
    a1.record(); torch.cuda.synchronize()
    ms = a0.elapsed_time(a1) / 5
    print(f"{str(ch):42s} {ms:.3f} ms per 1M blocks  {nb / ms / 1e3:.0f} M blocks/s  crc_ok {int(ok.sum())}")

# AACH (30, 14) soft ML decoder
b = rng.integers(0, 2, (pool, 14), dtype=np.uint8)
soft = etsi.to_soft(etsi.aach_encode(b, 7), 0.5, rng)
d = torch.from_numpy(soft).cuda().repeat(nb // pool + 1, 1)[:nb].contiguous()
dec = etsi.AachDecoder()
for _ in range(2):
    dec.decode_device(d, 7)
torch.cuda.synchronize()
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a0.record()
for _ in range(3):
    w, m, g = dec.decode_device(d, 7)
a1.record(); torch.cuda.synchronize()
ms = a0.elapsed_time(a1) / 3
print(f"AACH (30,14) ML {ms:.3f} ms per 1M blocks  {nb / ms / 1e3:.0f} M blocks/s")
