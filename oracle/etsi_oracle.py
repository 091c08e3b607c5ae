"""CPU restatement of the ETSI EN 300 392-2 clause-8 channel DEcoder (TEST INFRASTRUCTURE).

PARITY UNPINNED: the reference implements none of this (core/protocol.py:300); the algorithm is
the published one (descrambling LFSR, block de-interleaver (K, a), RCPC rate-2/3 de-puncturing
of the 16-state rate-1/4 mother code, Viterbi, CRC-16 with complemented parity).  This file pins
the CUDA decoder bit for bit (integer metrics, same tie rule: keep the lower predecessor unless
the upper one is strictly better) and is itself validated by encode -> decode round trips against
the encoder in tetraear_b200.etsi.
"""
import numpy as np

CHANNELS = {"SCH/HD": (124, 144, 216, 101), "SCH/F": (268, 288, 432, 103), "BSCH": (60, 80, 120, 11)}
GENERATORS = ((1, 1, 0, 0, 1), (1, 0, 1, 1, 1), (1, 1, 1, 0, 1), (1, 1, 0, 1, 1))


def scrambling_sequence(init, n):
    taps = (32, 26, 23, 22, 16, 12, 11, 10, 8, 7, 5, 4, 2, 1)
    lfsr = init & 0xFFFFFFFF
    out = np.empty(n, np.uint8)
    for i in range(n):
        bit = 0
        for t in taps:
            bit ^= (lfsr >> (32 - t)) & 1
        lfsr = (lfsr >> 1) | (bit << 31)
        out[i] = bit
    return out


def _branch_outputs():
    """out[state, b, g] for state = last four inputs (bit 0 = most recent), input b."""
    out = np.zeros((16, 2, 4), np.int32)
    for s in range(16):
        for b in range(2):
            reg = (b, s & 1, (s >> 1) & 1, (s >> 2) & 1, (s >> 3) & 1)        # x[n], x[n-1], ..., x[n-4]
            for g, taps in enumerate(GENERATORS):
                out[s, b, g] = sum(c & r for c, r in zip(taps, reg)) & 1
    return out


def decode(soft, channel="SCH/HD", scramb_init=3):
    """int8 soft type-5 values [n, K3] (+ = 1) -> (type-1 bits [n, K1], crc_ok [n], metric [n])."""
    K1, K2, K3, a = CHANNELS[channel]
    s = np.asarray(soft, dtype=np.int32).reshape(-1, K3)
    n = s.shape[0]
    sign = 1 - 2 * scrambling_sequence(scramb_init, K3).astype(np.int32)
    t4 = s * sign[None, :]
    k_of_i = (a * np.arange(1, K3 + 1)) % K3
    t3 = t4[:, k_of_i]
    # de-puncture: step 2m gets (g1, g2) = t3[3m], t3[3m+1]; step 2m+1 gets g1 = t3[3m+2]
    mother = np.zeros((n, K2, 4), np.int32)
    mother[:, 0::2, 0] = t3[:, 0::3]
    mother[:, 0::2, 1] = t3[:, 1::3]
    mother[:, 1::2, 0] = t3[:, 2::3]
    outs = 2 * _branch_outputs() - 1                                          # +-1
    NEG = -(1 << 28)
    pm = np.full((n, 16), NEG, np.int64)
    pm[:, 0] = 0
    dec = np.zeros((n, K2, 16), np.uint8)
    ns = np.arange(16)
    b = ns & 1
    p0 = ns >> 1
    p1 = p0 | 8
    for t in range(K2):
        bm0 = np.einsum("ng,sg->ns", mother[:, t, :], outs[p0, b, :])
        bm1 = np.einsum("ng,sg->ns", mother[:, t, :], outs[p1, b, :])
        m0 = pm[:, p0] + bm0
        m1 = pm[:, p1] + bm1
        take1 = m1 > m0
        dec[:, t, :] = take1
        pm = np.where(take1, m1, m0)
    state = np.zeros(n, np.int64)                                             # tail bits force state 0
    metric = pm[np.arange(n), state].astype(np.int32)
    bits = np.zeros((n, K2), np.uint8)
    rows = np.arange(n)
    for t in range(K2 - 1, -1, -1):
        bits[:, t] = state & 1
        d = dec[rows, t, state]
        state = (state >> 1) | (d.astype(np.int64) << 3)
    t1 = bits[:, :K1]
    crc = np.full(n, 0xFFFF, np.uint32)
    for i in range(K1):
        fb = (t1[:, i].astype(np.uint32) ^ (crc >> 15)) & 1
        crc = ((crc << 1) & 0xFFFF) ^ (fb * 0x1021)
    crc = (~crc) & 0xFFFF
    rx = np.zeros(n, np.uint32)
    for i in range(16):
        rx = (rx << 1) | bits[:, K1 + i]
    return t1, (crc == rx), metric
