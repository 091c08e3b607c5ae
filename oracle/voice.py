"""CPU restatement of the codec-frame builders on the output side of the path —
TEST INFRASTRUCTURE ONLY.

A codec frame is 690 little-endian int16: [0] = 0x6B21, 432 soft bits spread over four runs
(1..114, 116..229, 231..344, 346..435), zeros elsewhere.  Three builders exist in the reference:
    from_symbols  /root/reference/tetraear/ui/modern.py:2309-2423  (_extract_voice_slot_from_symbols):
                  slot = dibits[pos//3 : pos//3 + 255] (None if it does not fit); symbols 0..107 and
                  119..226, MSB first, bit -> +-16384 -> int(sb/16384*127) = +-127
    from_bits     /root/reference/continuous_capture.py:100-124: frame["bits"][:432] -> +-127
    from_params   /root/reference/tetraear/ui/modern.py:2143-2200 (fallback when the slot is unavailable):
                  four 108-bit sub-frames at 1 + 172 k: energy (6 bits LSB first) -> int(e/63*1000) - 500,
                  pitch (7) -> int(p/127*200) + 20, gain (3) -> int(g/7*150), then 92 bits -> +-300
Only tests/ may import this module.  from_symbols is pinned against the reference's own function text
(tests/golden/voice.npz); the other two are inline blocks of larger functions — restated, "parity unpinned".
"""
import numpy as np

RUNS = ((1, 115), (116, 230), (231, 345), (346, 436))
HEADER = 0x6B21


def _place(soft):
    block = np.zeros(690, dtype=np.int16)
    block[0] = HEADER
    idx = 0
    for lo, hi in RUNS:
        n = min(hi - lo, max(0, len(soft) - idx))
        block[lo:lo + n] = soft[idx:idx + n]
        idx += n
    return block


def from_symbols(position, dibits):
    dibits = np.asarray(dibits)
    sp = int(position) // 3
    if sp + 255 > len(dibits):
        return None
    slot = dibits[sp:sp + 255].astype(np.int64)
    sel = np.concatenate([slot[0:108], slot[119:227]])
    bits = np.stack([(sel >> 1) & 1, sel & 1], axis=1).reshape(-1)
    return _place(np.where(bits != 0, 127, -127).astype(np.int16))


def from_bits(bits):
    bits = np.asarray(bits)
    if len(bits) < 432:
        return None
    return _place(np.where(bits[:432].astype(np.int64) != 0, 127, -127).astype(np.int16))


def from_params(bits):
    bits = np.asarray(bits).astype(np.int64)
    block = np.zeros(690, dtype=np.int16)
    block[0] = HEADER
    if len(bits) < 432:
        return block
    for k in range(4):
        sub = bits[108 * k:108 * k + 108]
        base = 1 + 172 * k
        e = int(sum(int(b) << i for i, b in enumerate(sub[0:6])))
        p = int(sum(int(b) << i for i, b in enumerate(sub[6:13])))
        g = int(sum(int(b) << i for i, b in enumerate(sub[13:16])))
        block[base] = int((e / 63.0) * 1000) - 500
        block[base + 1] = int((p / 127.0) * 200) + 20
        block[base + 2] = int((g / 7.0) * 150)
        for i, b in enumerate(sub[16:108]):
            if base + 3 + i < 690:
                block[base + 3 + i] = 300 if b else -300
    return block
