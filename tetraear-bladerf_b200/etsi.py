"""ETSI EN 300 392-2 channel coding for the `etsi` mode of the receive path.

The reference does NOT implement descrambling / de-interleaving / RCPC decoding ("we do not do
full channel decoding", /root/reference/tetraear/core/protocol.py:300) — parity for this stage is
UNPINNED by the reference.  It follows the published ETSI algorithm (clause 8, as commonly
implemented, e.g. osmo-tetra) and is validated by encode -> channel ->
decode round trips and against the numpy restatement in oracle/etsi_oracle.py.

This module holds (a) the ENCODER, a numpy workload generator used by the tests and bench, and
(b) ``ChannelDecoder``, the host API of the CUDA decoder (tb200_etsi_decode_batch).
"""
import numpy as np

from . import _lib

# logical channels: (K1 type-1 bits, K2 = K1 + 16 CRC + 4 tail, K3 = 3/2 K2, interleaver a)
CHANNELS = {"SCH/HD": (124, 144, 216, 101), "SCH/F": (268, 288, 432, 103), "BSCH": (60, 80, 120, 11)}
CHANNEL_IDS = {"SCH/HD": 0, "SCH/F": 1, "BSCH": 2}
# rate-1/4 mother code, constraint length 5: taps on (x[n], x[n-1], x[n-2], x[n-3], x[n-4])
GENERATORS = ((1, 1, 0, 0, 1), (1, 0, 1, 1, 1), (1, 1, 1, 0, 1), (1, 1, 0, 1, 1))
PUNCT_2_3 = (1, 2, 5)          # kept positions (1-based) of every 8 mother bits, t = 3


def scrambling_init(mcc=0, mnc=0, colour=0):
    """30-bit extended colour code followed by '11' (BSCH uses the all-zero code -> 3)."""
    return (((colour & 0x3F) | ((mnc & 0x3FFF) << 6) | ((mcc & 0x3FF) << 20)) << 2) | 3


def scrambling_sequence(init, n):
    """Fibonacci LFSR, c(x) = 1+x+x2+x4+x5+x7+x8+x10+x11+x12+x16+x22+x23+x26+x32."""
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


def crc16_block(bits):
    """(K1+16, K1) block code: CRC-16-CCITT (X16+X12+X5+1), register preset to ones, parity
    complemented.  bits: [n, K1] -> [n, 16]."""
    bits = np.asarray(bits, dtype=np.uint8)
    crc = np.full(bits.shape[0], 0xFFFF, dtype=np.uint32)
    for i in range(bits.shape[1]):
        fb = (bits[:, i].astype(np.uint32) ^ (crc >> 15)) & 1
        crc = ((crc << 1) & 0xFFFF) ^ (fb * 0x1021)
    crc = (~crc) & 0xFFFF
    return ((crc[:, None] >> np.arange(15, -1, -1)[None, :]) & 1).astype(np.uint8)


def interleave_index(K, a):
    """type4[k] = type3[i] with k = 1 + (a * i mod K), i = 1..K  ->  0-based array k_of_i."""
    i = np.arange(1, K + 1)
    return (a * i) % K          # = k - 1


def encode(type1, channel="SCH/HD", scramb_init=3):
    """[n, K1] type-1 bits -> [n, K3] type-5 bits."""
    K1, K2, K3, a = CHANNELS[channel]
    t1 = np.asarray(type1, dtype=np.uint8).reshape(-1, K1)
    n = t1.shape[0]
    t2 = np.concatenate([t1, crc16_block(t1), np.zeros((n, 4), np.uint8)], axis=1)
    assert t2.shape[1] == K2
    pad = np.concatenate([np.zeros((n, 4), np.uint8), t2], axis=1)         # x[n-4..n]
    mother = np.zeros((n, K2, 4), np.uint8)
    for g, taps in enumerate(GENERATORS):
        acc = np.zeros((n, K2), np.uint8)
        for d, c in enumerate(taps):
            if c:
                acc ^= pad[:, 4 - d:4 - d + K2]
        mother[:, :, g] = acc
    flat = mother.reshape(n, K2 * 4)                                      # V(1..4 K2)
    j = np.arange(K3)
    idx = 8 * (j // 3) + np.asarray(PUNCT_2_3)[j % 3] - 1
    t3 = flat[:, idx]
    t4 = np.empty_like(t3)
    t4[:, interleave_index(K3, a)] = t3
    return t4 ^ scrambling_sequence(scramb_init, K3)[None, :]


def to_soft(bits, sigma=0.0, rng=None, scale=32.0):
    """Hard bits -> int8 soft values (+ means 1) through a BPSK/AWGN channel."""
    x = (2.0 * np.asarray(bits, dtype=np.float64) - 1.0)
    if sigma > 0:
        rng = np.random.default_rng(0) if rng is None else rng
        x = x + sigma * rng.standard_normal(x.shape)
    return np.clip(np.rint(x * scale), -127, 127).astype(np.int8)


class ChannelDecoder:
    """Batched descramble + de-interleave + RCPC (rate 2/3) Viterbi + CRC on the GPU."""

    def __init__(self, channel="SCH/HD", device=0):
        self.channel = channel
        self.K1, self.K2, self.K3, self.a = CHANNELS[channel]
        self.device = device
        self._h = _lib.Handle(max_chunks=0, max_chunk_len=0, device=device)

    def decode_device(self, soft, scramb_init=3, bits=None, crc_ok=None, metric=None):
        """soft: CUDA int8 tensor [n, K3].  Returns (bits uint8 [n, 40] packed MSB-first type-1 bits,
        crc_ok uint8 [n], metric int32 [n]) CUDA tensors.  Asynchronous on the current stream."""
        import torch
        from . import _dev
        n = soft.shape[0]
        if bits is None:
            bits = _dev.empty((n, 40), torch.uint8, self.device)
        if crc_ok is None:
            crc_ok = _dev.empty((n,), torch.uint8, self.device)
        if metric is None:
            metric = _dev.empty((n,), torch.int32, self.device)
        h = self._h
        _lib.check(h.lib.tb200_etsi_decode_batch(h.h, soft.data_ptr(), n, CHANNEL_IDS[self.channel],
                                                 int(scramb_init) & 0xFFFFFFFF, bits.data_ptr(), crc_ok.data_ptr(),
                                                 metric.data_ptr(), _dev.stream_ptr(self.device)))
        return bits, crc_ok, metric

    def decode(self, soft, scramb_init=3):
        """numpy int8 [n, K3] -> (type-1 bits uint8 [n, K1], crc_ok bool [n], metric int32 [n])."""
        from . import _dev
        s = np.ascontiguousarray(soft, dtype=np.int8).reshape(-1, self.K3)
        d = _dev.to_dev(s, self.device)
        bits, ok, met = self.decode_device(d, scramb_init)
        b = np.unpackbits(bits.cpu().numpy(), axis=1)[:, :self.K1]
        return b, ok.cpu().numpy().astype(bool), met.cpu().numpy()
