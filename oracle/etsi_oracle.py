"""CPU restatement of the ETSI EN 300 392-2 clause-8 channel DEcoder (TEST INFRASTRUCTURE).

PARITY UNPINNED: the reference implements none of this (core/protocol.py:300); the algorithm is
the published one (descrambling LFSR, block de-interleaver (K, a), RCPC de-puncturing (rates 2/3, 1/3, 292/432, 148/432)
of the 16-state rate-1/4 mother code, Viterbi, CRC-16 with complemented parity).  This file pins
the CUDA decoder bit for bit (integer metrics, same tie rule: keep the lower predecessor unless
the upper one is strictly better) and is itself validated by encode -> decode round trips against
the encoder in tetraear_b200.etsi.
"""
import numpy as np

# logical channel -> (K1 type-1 bits, K2 type-2 bits, K3 type-3 bits, interleaver a, RCPC rate, 16-bit CRC present)
CHANNELS = {"SCH/HD": (124, 144, 216, 101, "2/3", True), "SCH/F": (268, 288, 432, 103, "2/3", True),
            "BSCH": (60, 80, 120, 11, "2/3", True), "SCH/HU": (92, 112, 168, 13, "2/3", True),
            "TCH/4.8": (288, 292, 432, 103, "292/432", False), "TCH/2.4": (144, 148, 432, 103, "148/432", False)}
# RCPC puncturing (EN 300 392-2 cl. 8.2.3.1.3-8.2.3.1.6): rate -> (P(1..t), t, divisor of the i(j) = j + (j-1) div d rule)
PUNCTURERS = {"2/3": ((1, 2, 5), 3, 0), "1/3": ((1, 2, 3, 5, 6, 7), 6, 0),
              "292/432": ((1, 2, 5), 3, 65), "148/432": ((1, 2, 3, 5, 6, 7), 6, 35)}
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


def puncture_map(rate, K3):
    """0-based index into the mother-code output V(1..4 K2) of the type-3 bit j = 1..K3 (cl. 8.2.3.1.2):
    k = 8 ((i-1) div t) + P(i - t ((i-1) div t)), i = j for rates 2/3 and 1/3, i = j + (j-1) div 65 for
    rate 292/432 and i = j + (j-1) div 35 for rate 148/432."""
    P, t, div = PUNCTURERS[rate]
    j = np.arange(1, K3 + 1)
    i = j + (j - 1) // div if div else j
    blk = (i - 1) // t
    return 8 * blk + np.asarray(P)[i - t * blk - 1] - 1


def decode(soft, channel="SCH/HD", scramb_init=3):
    """int8 soft type-5 values [n, K3] (+ = 1) -> (type-1 bits [n, K1], crc_ok [n], metric [n])."""
    return decode_custom(soft, *CHANNELS[channel], scramb_init=scramb_init)


def decode_custom(soft, K1, K2, K3, a, rate="2/3", crc=True, scramb_init=3):
    """The same for any block geometry: K2 = K1 (+ 16 CRC bits) + 4 tail bits, K3 type-3 bits of the given rate."""
    s = np.asarray(soft, dtype=np.int32).reshape(-1, K3)
    n = s.shape[0]
    sign = 1 - 2 * scrambling_sequence(scramb_init, K3).astype(np.int32)
    t4 = s * sign[None, :]
    k_of_i = (a * np.arange(1, K3 + 1)) % K3
    t3 = t4[:, k_of_i]
    # de-puncture: punctured mother bits count as erasures (0)
    kidx = puncture_map(rate, K3)
    assert kidx.max() < 4 * K2 and len(set(kidx.tolist())) == K3
    flat = np.zeros((n, K2 * 4), np.int32)
    flat[:, kidx] = t3
    mother = flat.reshape(n, K2, 4)
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
    if not crc:
        return t1, np.ones(n, bool), metric
    crc = np.full(n, 0xFFFF, np.uint32)
    for i in range(K1):
        fb = (t1[:, i].astype(np.uint32) ^ (crc >> 15)) & 1
        crc = ((crc << 1) & 0xFFFF) ^ (fb * 0x1021)
    crc = (~crc) & 0xFFFF
    rx = np.zeros(n, np.uint32)
    for i in range(16):
        rx = (rx << 1) | bits[:, K1 + i]
    return t1, (crc == rx), metric


def aach_decode(soft, generator, scramb_init=3):
    """Brute-force soft ML decoding of a (30, 14) block code (AACH, cl. 8.2.3.2): int8 [n, 30] -> (bits [n, 14],
    correlation [n], margin to the runner-up [n]).  Ties go to the smaller information word (bit 1 = MSB).  The
    generator [14, 30] is a parameter — see tetraear_b200.etsi.rm3014_generator for what the tests use."""
    G = np.asarray(generator, np.int64).reshape(14, 30)
    s = np.asarray(soft, np.int64).reshape(-1, 30) * (1 - 2 * scrambling_sequence(scramb_init, 30).astype(np.int64))[None, :]
    c = np.arange(1 << 14)
    info = (c[:, None] >> np.arange(13, -1, -1)[None, :]) & 1
    pm = 2 * ((info @ G) & 1) - 1                                              # [16384, 30] +-1
    corr = s @ pm.T
    best = corr.argmax(axis=1)                                                 # first maximum = smallest word
    top2 = np.partition(corr, -2, axis=1)[:, -2:]
    return info[best].astype(np.uint8), top2[:, 1].astype(np.int32), (top2[:, 1] - top2[:, 0]).astype(np.int32)
