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

# logical channels: (K1 type-1 bits, K2 = K1 (+ 16 CRC) + 4 tail, K3 type-3 bits, interleaver a)
CHANNELS = {"SCH/HD": (124, 144, 216, 101), "SCH/F": (268, 288, 432, 103), "BSCH": (60, 80, 120, 11),
            "SCH/HU": (92, 112, 168, 13), "TCH/4.8": (288, 292, 432, 103), "TCH/2.4": (144, 148, 432, 103)}
CHANNEL_IDS = {"SCH/HD": 0, "SCH/F": 1, "BSCH": 2, "SCH/HU": 3, "TCH/4.8": 4, "TCH/2.4": 5}
# (RCPC rate, 16-bit CRC present): the traffic channels carry no block code (cl. 8.3.1), only the 4 tail bits;
# their interleaving over N blocks is taken with N = 1 (block interleaver (432, 103) only)
CHANNEL_CODING = {"SCH/HD": ("2/3", True), "SCH/F": ("2/3", True), "BSCH": ("2/3", True), "SCH/HU": ("2/3", True),
                  "TCH/4.8": ("292/432", False), "TCH/2.4": ("148/432", False)}
RATE_IDS = {"2/3": 0, "1/3": 1, "292/432": 2, "148/432": 3}
# RCPC puncturing, cl. 8.2.3.1.3-8.2.3.1.6: rate -> (P(1..t), t, d of i(j) = j + (j-1) div d; 0: i = j)
PUNCTURERS = {"2/3": ((1, 2, 5), 3, 0), "1/3": ((1, 2, 3, 5, 6, 7), 6, 0),
              "292/432": ((1, 2, 5), 3, 65), "148/432": ((1, 2, 3, 5, 6, 7), 6, 35)}
# rate-1/4 mother code, constraint length 5: taps on (x[n], x[n-1], x[n-2], x[n-3], x[n-4])
GENERATORS = ((1, 1, 0, 0, 1), (1, 0, 1, 1, 1), (1, 1, 1, 0, 1), (1, 1, 0, 1, 1))


def puncture_index(rate, K3):
    """0-based position in the mother-code output V(1..4 K2) of type-3 bit j = 1..K3 (cl. 8.2.3.1.2)."""
    P, t, div = PUNCTURERS[rate]
    j = np.arange(1, K3 + 1)
    i = j + (j - 1) // div if div else j
    blk = (i - 1) // t
    return 8 * blk + np.asarray(P)[i - t * blk - 1] - 1


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
    return encode_custom(type1, *CHANNELS[channel], *CHANNEL_CODING[channel], scramb_init=scramb_init)


def encode_custom(type1, K1, K2, K3, a, rate="2/3", crc=True, scramb_init=3):
    """The same for any block geometry (K2 = K1 + 16 crc + 4, K3 the number of type-3 bits the rate leaves)."""
    t1 = np.asarray(type1, dtype=np.uint8).reshape(-1, K1)
    n = t1.shape[0]
    t2 = np.concatenate([t1] + ([crc16_block(t1)] if crc else []) + [np.zeros((n, 4), np.uint8)], axis=1)
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
    t3 = flat[:, puncture_index(rate, K3)]
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
    """Batched descramble + de-interleave + RCPC Viterbi (+ CRC) on the GPU.  `channel` is a logical channel name
    or a tuple (K1, K2, K3, a, rate, crc) for any other block geometry (K2 <= 320; tb200_etsi_decode_custom)."""

    def __init__(self, channel="SCH/HD", device=0):
        self.channel = channel
        if isinstance(channel, str):
            self.K1, self.K2, self.K3, self.a = CHANNELS[channel]
            self.rate, self.crc = CHANNEL_CODING[channel]
        else:
            self.K1, self.K2, self.K3, self.a, self.rate, self.crc = channel
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
        if isinstance(self.channel, str):
            _lib.check(h.lib.tb200_etsi_decode_batch(h.h, soft.data_ptr(), n, CHANNEL_IDS[self.channel],
                                                     int(scramb_init) & 0xFFFFFFFF, bits.data_ptr(), crc_ok.data_ptr(),
                                                     metric.data_ptr(), _dev.stream_ptr(self.device)))
        else:
            _lib.check(h.lib.tb200_etsi_decode_custom(h.h, soft.data_ptr(), n, self.K1, self.K2, self.K3, self.a,
                                                      RATE_IDS[self.rate], int(bool(self.crc)),
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


# ---- AACH: (30, 14) shortened Reed-Muller code (cl. 8.2.3.2)
def rm3014_generator():
    """A systematic generator [I14 | P] of a (30, 14, d = 8) shortened second-order Reed-Muller code: RM(2, 5) (the 16
    monomials of degree <= 2 evaluated at the 32 points of GF(2)^5), restricted to the codewords that vanish at the
    points 00000 and 11111, those two positions deleted, reduced to systematic form.  This is the CODE FAMILY of
    cl. 8.2.3.2; the standard's own matrix (a particular basis and bit order) is not reproduced here — pass it as
    `generator` to aach_encode / AachDecoder for on-air signals.  uint8 [14, 30]."""
    import itertools
    pts = np.array([[(v >> (4 - i)) & 1 for i in range(5)] for v in range(32)], np.uint8)
    rows = [np.ones(32, np.uint8)] + [pts[:, i] for i in range(5)]
    rows += [pts[:, i] & pts[:, j] for i, j in itertools.combinations(range(5), 2)]
    G = np.array(rows, np.uint8)

    def rref(M):
        M = M.copy()
        r, piv = 0, []
        for c in range(M.shape[1]):
            p = [i for i in range(r, M.shape[0]) if M[i, c]]
            if not p:
                continue
            M[[r, p[0]]] = M[[p[0], r]]
            for i in range(M.shape[0]):
                if i != r and M[i, c]:
                    M[i] ^= M[r]
            piv.append(c)
            r += 1
            if r == M.shape[0]:
                break
        return M[:r], piv

    R, _ = rref(G[:, [0, 31] + list(range(1, 31))])
    S, piv = rref(R[2:, 2:])
    return S[:, piv + [c for c in range(30) if c not in piv]]


def generator_words(generator):
    """[14, 30] 0/1 matrix -> the 14 uint32 row words of tb200_etsi_aach_decode (code bit j in word bit 29 - j)."""
    g = np.asarray(generator, dtype=np.uint64).reshape(14, 30)
    return (g << np.arange(29, -1, -1, dtype=np.uint64)[None, :]).sum(axis=1).astype(np.uint32)


def aach_encode(info, scramb_init=3, generator=None):
    """[n, 14] information bits -> [n, 30] scrambled type-5 bits."""
    G = rm3014_generator() if generator is None else np.asarray(generator, np.uint8)
    b = np.asarray(info, dtype=np.uint8).reshape(-1, 14)
    return ((b.astype(np.int64) @ G.astype(np.int64)) & 1).astype(np.uint8) ^ scrambling_sequence(scramb_init, 30)[None, :]


class AachDecoder:
    """Batched soft ML decoder of the AACH block code on the GPU (tb200_etsi_aach_decode)."""

    def __init__(self, generator=None, device=0):
        self.generator = rm3014_generator() if generator is None else np.asarray(generator, np.uint8).reshape(14, 30)
        self._rows = np.ascontiguousarray(generator_words(self.generator))
        self.device = device
        self._h = _lib.Handle(max_chunks=0, max_chunk_len=0, device=device)

    def decode_device(self, soft, scramb_init=3):
        """soft: CUDA int8 [n, 30] -> (uint16 [n] information word, bit 1 = MSB of 14; int32 [n] correlation; int32 [n]
        margin to the runner-up) CUDA tensors (int16 view for the words)."""
        import torch
        from . import _dev
        n = soft.shape[0]
        bits = _dev.empty((n,), torch.int16, self.device)
        metric = _dev.empty((n,), torch.int32, self.device)
        margin = _dev.empty((n,), torch.int32, self.device)
        h = self._h
        _lib.check(h.lib.tb200_etsi_aach_decode(h.h, soft.data_ptr(), n, self._rows.ctypes.data, int(scramb_init) & 0xFFFFFFFF,
                                                bits.data_ptr(), metric.data_ptr(), margin.data_ptr(),
                                                _dev.stream_ptr(self.device)))
        return bits, metric, margin

    def decode(self, soft, scramb_init=3):
        """numpy int8 [n, 30] -> (bits uint8 [n, 14], metric int32 [n], margin int32 [n])."""
        from . import _dev
        s = np.ascontiguousarray(soft, dtype=np.int8).reshape(-1, 30)
        w, m, g = self.decode_device(_dev.to_dev(s, self.device), scramb_init)
        w = w.cpu().numpy().view(np.uint16).astype(np.int64)
        return ((w[:, None] >> np.arange(13, -1, -1)[None, :]) & 1).astype(np.uint8), m.cpu().numpy(), g.cpu().numpy()


# ======================================================================================
# ETSI physical layer: burst building + pi/4-DQPSK RRC modulator (workload generator) and the
# host API of the CUDA demodulator (RRC matched filter, closed-loop timing, soft bits, TS search)
# ======================================================================================
TS_N = np.array([1, 1, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0, 1, 0, 0], dtype=np.uint8)   # normal 1
TS_P = np.array([0, 1, 1, 1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 1, 1, 1, 0], dtype=np.uint8)   # normal 2
TS_Q = np.array([1, 0, 1, 1, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1, 1, 0, 1, 0, 1, 1, 0, 1], dtype=np.uint8)   # tail/extended
TS_X = np.array([int(c) for c in "100111010000111010011101000011"], dtype=np.uint8)                    # extended (uplink control)
SEQUENCE_IDS = {"n": 0, "p": 1, "q": 2, "x": 3, "y": 4}
# normal continuous downlink burst, bit offsets (EN 300 392-2 cl. 9.4.4.3.2)
NDB = {"q_tail": (0, 12), "pa": (12, 14), "bkn1": (14, 230), "bb1": (230, 244), "ts": (244, 266),
       "bb2": (266, 282), "bkn2": (282, 498), "pb": (498, 500), "q_head": (500, 510)}
# synchronisation continuous downlink burst (cl. 9.4.4.2.6): SB1 = the BSCH block, y = 38-bit sync training sequence
TS_Y = np.array([int(c) for c in "11000001100111001110100111000001100111"], dtype=np.uint8)
SDB = {"q_tail": (0, 12), "pa": (12, 14), "fc": (14, 94), "sb1": (94, 214), "ts": (214, 252), "bb": (252, 282),
       "bkn2": (282, 498), "pb": (498, 500), "q_head": (500, 510)}
FREQ_CORRECTION = np.concatenate([np.ones(8, np.uint8), np.zeros(64, np.uint8), np.ones(8, np.uint8)])
# BSCH type-1 block: SYNC PDU (cl. 21.4.4.2) + D-MLE-SYNC (cl. 18.4.2.1), 60 bits
BSCH_FIELDS = (("system_code", 4), ("colour_code", 6), ("timeslot", 2), ("frame", 5), ("multiframe", 6), ("sharing_mode", 2),
               ("ts_reserved_frames", 3), ("u_plane_dtx", 1), ("frame18_extension", 1), ("reserved", 1),
               ("mcc", 10), ("mnc", 14), ("neighbour_cell_broadcast", 2), ("cell_service_level", 2), ("late_entry", 1))
# dibit (b1, b2) -> differential phase
DQPSK_DPHI = {(0, 0): np.pi / 4, (0, 1): 3 * np.pi / 4, (1, 1): -3 * np.pi / 4, (1, 0): -np.pi / 4}


def rrc_pulse(t, alpha=0.35):
    """Root-raised-cosine impulse response at times t (in symbols), unit energy per symbol."""
    t = np.asarray(t, dtype=np.float64)
    out = np.empty_like(t)
    eps = 1e-9
    z = np.abs(t) < eps
    s = np.abs(np.abs(t) - 1 / (4 * alpha)) < eps
    r = ~(z | s)
    out[z] = 1 - alpha + 4 * alpha / np.pi
    out[s] = alpha / np.sqrt(2) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha)) + (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
    tr = t[r]
    out[r] = (np.sin(np.pi * tr * (1 - alpha)) + 4 * alpha * tr * np.cos(np.pi * tr * (1 + alpha))) / \
             (np.pi * tr * (1 - (4 * alpha * tr) ** 2))
    return out


def build_normal_burst(bkn1, bkn2, rng=None, ts=TS_N, bb=None):
    """510-bit normal continuous downlink burst around two 216-bit (type-5) blocks; bb = the 30 broadcast-block bits
    (the AACH block, 14 before and 16 behind the training sequence) or None for random bits."""
    rng = np.random.default_rng(0) if rng is None else rng
    b = rng.integers(0, 2, 510, dtype=np.uint8)
    if bb is not None:
        b[230:244] = bb[:14]
        b[266:282] = bb[14:30]
    b[0:12] = TS_Q[10:22]
    b[14:230] = bkn1
    b[244:266] = ts
    b[282:498] = bkn2
    b[500:510] = TS_Q[0:10]
    return b


def bsch_pack(**fields):
    """60 type-1 bits of a BSCH block from named fields (MSB first; missing fields = 0)."""
    out = []
    for name, width in BSCH_FIELDS:
        v = int(fields.get(name, 0))
        out += [(v >> (width - 1 - i)) & 1 for i in range(width)]
    return np.array(out, dtype=np.uint8)


def bsch_unpack(bits):
    bits = np.asarray(bits).astype(np.int64)
    out, pos = {}, 0
    for name, width in BSCH_FIELDS:
        out[name] = int(sum(int(b) << (width - 1 - i) for i, b in enumerate(bits[pos:pos + width])))
        pos += width
    return out


def build_sync_burst(sb1, bkn2, rng=None):
    """510-bit synchronisation continuous downlink burst around the 120-bit BSCH block and a 216-bit block."""
    rng = np.random.default_rng(0) if rng is None else rng
    b = rng.integers(0, 2, 510, dtype=np.uint8)
    b[0:12] = TS_Q[10:22]
    b[14:94] = FREQ_CORRECTION
    b[94:214] = sb1
    b[214:252] = TS_Y
    b[282:498] = bkn2
    b[500:510] = TS_Q[0:10]
    return b


def make_etsi_stream(n_bursts, seed=0, scramb_init=None, ts=TS_N, sync_every=0, network=(262, 42, 17), aach=False):
    """Back-to-back downlink bursts carrying random SCH/HD blocks.  With sync_every = k > 0 every k-th burst is a
    synchronisation burst whose BSCH block announces `network` = (MCC, MNC, colour code) and whose BKN2 is an SCH/HD
    block (its first type-1 row of that burst is then unused).  Returns (bits, type1 [n,2,124], scrambling init);
    with aach=True the broadcast block of every normal burst carries a random 14-bit access-assignment word through
    the (30, 14) code of rm3014_generator(), and the words [n, 14] are returned as a fourth item."""
    rng = np.random.default_rng(seed)
    init = scrambling_init(*network) if scramb_init is None else scramb_init
    aach_bits = np.random.default_rng(seed + 100003).integers(0, 2, (n_bursts, 14), dtype=np.uint8) if aach else None
    bb = aach_encode(aach_bits, init) if aach else None
    t1 = rng.integers(0, 2, (n_bursts, 2, 124), dtype=np.uint8)
    t5 = encode(t1.reshape(-1, 124), "SCH/HD", init).reshape(n_bursts, 2, 216)
    out = []
    for i in range(n_bursts):
        if sync_every and i % sync_every == 0:
            sb = bsch_pack(mcc=network[0], mnc=network[1], colour_code=network[2], timeslot=i % 4, frame=1 + (i // 4) % 18,
                           multiframe=1 + (i // 72) % 60)
            out.append(build_sync_burst(encode(sb[None, :], "BSCH", 3)[0], t5[i, 1], rng))
        else:
            out.append(build_normal_burst(t5[i, 0], t5[i, 1], rng, ts, None if bb is None else bb[i]))
    if aach:
        return np.concatenate(out), t1, init, aach_bits
    return np.concatenate(out), t1, init


def modulate_dqpsk(bits, fs, symbol_rate=18000.0, sigma=0.0, freq=0.0, t_offset=0.0, amplitude=1.0, seed=0,
                   n_samples=None, span=6):
    """pi/4-DQPSK with RRC(0.35) pulse shaping at an arbitrary (non-integer) samples-per-symbol ratio,
    mixed to `freq` Hz; complex64."""
    bits = np.asarray(bits, dtype=np.uint8)
    dib = bits.reshape(-1, 2)
    dphi = np.array([DQPSK_DPHI[(int(a), int(b))] for a, b in dib])
    sym = amplitude * np.exp(1j * np.concatenate([[0.0], np.cumsum(dphi)]))
    sps = fs / symbol_rate
    n = int(np.ceil((len(sym) + 2 * span) * sps)) if n_samples is None else n_samples
    t = (np.arange(n) / sps) - span - t_offset                   # time in symbols, symbol k at t = k
    k0 = np.floor(t).astype(np.int64)
    x = np.zeros(n, dtype=np.complex128)
    for d in range(-span + 1, span + 1):
        k = k0 + d
        ok = (k >= 0) & (k < len(sym))
        x[ok] += sym[k[ok]] * rrc_pulse(t[ok] - k[ok])
    if freq:
        x *= np.exp(2j * np.pi * freq * np.arange(n) / fs)
    if sigma > 0:
        rng = np.random.default_rng(seed + 1000)
        x += sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(np.complex64)


class EtsiReceiver:
    """`etsi` mode receive chain for a batch of 25 kHz channels, all on the GPU:
    RRC matched filter + closed-loop timing + soft pi/4-DQPSK bits (tb200_etsi_demod_batch) ->
    training-sequence search + block extraction (tb200_etsi_find_bursts) ->
    descramble / de-interleave / Viterbi / CRC of both SCH/HD blocks (tb200_etsi_decode_batch)."""

    def __init__(self, fs, max_streams, max_samples, device=0, max_bursts_per_stream=None, afc=True, segmented=True, seg_lanes=0):
        import torch
        from . import _dev
        self.fs = float(fs)
        self.device = device
        self.max_streams, self.max_samples = int(max_streams), int(max_samples)
        sps = self.fs / 18000.0
        self.max_sym = int(max_samples / sps) + 8
        self.soft_stride = 2 * self.max_sym
        self.max_bursts = max_bursts_per_stream or (self.max_sym // 255 + 2)
        self._h = _lib.Handle(max_chunks=0, max_chunk_len=0, device=device)
        S = self.max_streams
        self.soft = _dev.zeros((S, self.soft_stride), torch.float32, device)
        self.n_soft = _dev.zeros((S,), torch.int32, device)
        self.metric = _dev.zeros((S, self.max_sym), torch.float32, device)
        self.blocks = _dev.zeros((S, self.max_bursts, 432), torch.int8, device)
        self.burst_pos = _dev.zeros((S, self.max_bursts), torch.int32, device)
        self.n_bursts = _dev.zeros((S,), torch.int32, device)
        self.bits = _dev.zeros((S * self.max_bursts * 2, 40), torch.uint8, device)
        self.crc_ok = _dev.zeros((S * self.max_bursts * 2,), torch.uint8, device)
        self.path_metric = _dev.zeros((S * self.max_bursts * 2,), torch.int32, device)
        self.symbols = None
        # time-parallel demodulator (tb200_etsi_demod_segmented): streams of two segments (1280 symbols) or more are cut
        # into concurrently running segments; shorter ones, or segmented=False, run the sequential loop
        self.segmented = bool(segmented)
        # form of the time-parallel demodulator: 0 = two stages (matched filter once per sample, then one thread per
        # segment interpolating its output), 8 = eight lanes per segment with the matched filter inside the loop
        import os
        seg_lanes = int(seg_lanes or os.environ.get("TB200_ETSI_SEG_LANES", "0"))     # the environment variable is for A/B timing
        if seg_lanes:
            self._h.set_option(_lib.OPT_ETSI_SEG_LANES, seg_lanes)
        self._scratch = None
        self.afc = bool(afc)                      # feed-forward carrier-offset correction (tb200_etsi_afc), |df| < 2250 Hz
        self.theta = _dev.zeros((S,), torch.float32, device)
        self.aach = None                          # AachDecoder, created by run(..., aach=True)
        self.bb = _dev.zeros((S, self.max_bursts, 30), torch.int8, device)
        self.aach_words = self.aach_metric = self.aach_margin = None
        self.max_sync = self.max_bursts
        self.sb1 = _dev.zeros((S, self.max_sync, 120), torch.int8, device)
        self.sync_pos = _dev.zeros((S, self.max_sync), torch.int32, device)
        self.n_sync = _dev.zeros((S,), torch.int32, device)
        self.sync_bits = _dev.zeros((S * self.max_sync, 40), torch.uint8, device)
        self.sync_crc = _dev.zeros((S * self.max_sync,), torch.uint8, device)
        self.sync_metric = _dev.zeros((S * self.max_sync,), torch.int32, device)

    def _demod(self, iq_ptr, n_streams, n_samples, stride, sym_ptr, st):
        """Matched filter + timing loop + soft bits of every stream (segmented when that pays, see __init__)."""
        import torch
        from . import _dev
        h, lib = self._h, self._h.lib
        need = 0
        if self.segmented:
            need = int(lib.tb200_etsi_demod_scratch_bytes(int(n_streams), int(n_samples), self.fs, 1 if sym_ptr else 0))
        if need <= 0:
            _lib.check(lib.tb200_etsi_demod_batch(h.h, int(iq_ptr), int(n_streams), int(n_samples), int(stride), self.fs, sym_ptr,
                                                  self.soft.data_ptr(), self.soft_stride, self.n_soft.data_ptr(), st))
            return
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = _dev.empty((need,), torch.uint8, self.device)
        _lib.check(lib.tb200_etsi_demod_segmented(h.h, int(iq_ptr), int(n_streams), int(n_samples), int(stride), self.fs, sym_ptr,
                                                  self.soft.data_ptr(), self.soft_stride, self.n_soft.data_ptr(),
                                                  self._scratch.data_ptr(), int(self._scratch.numel()), st))

    def acquire_device(self, n_streams, threshold=0.8):
        """After a demodulation (run_device*): search the soft bits of every stream for synchronisation bursts (38-bit
        sequence y) and decode their BSCH blocks (scrambling init 3).  Asynchronous; see acquire() for the host view."""
        from . import _dev
        h, lib = self._h, self._h.lib
        st = _dev.stream_ptr(self.device)
        self.metric.zero_()
        _lib.check(lib.tb200_etsi_find_sync_bursts(h.h, self.soft.data_ptr(), n_streams, self.soft_stride, self.n_soft.data_ptr(),
                                                   float(threshold), self.metric.data_ptr(), self.sb1.data_ptr(), None,
                                                   self.sync_pos.data_ptr(), self.n_sync.data_ptr(), self.max_sync, st))
        _lib.check(lib.tb200_etsi_decode_batch(h.h, self.sb1.data_ptr(), n_streams * self.max_sync, CHANNEL_IDS["BSCH"], 3,
                                               self.sync_bits.data_ptr(), self.sync_crc.data_ptr(), self.sync_metric.data_ptr(), st))

    def acquire(self, iq, threshold=0.8):
        """numpy complex64 [n_streams, n] -> per stream the list of CRC-clean BSCH blocks found:
        dicts(start, fields (bsch_unpack), scrambling_init).  The bootstrap of `etsi` mode: no network code needed."""
        from . import _dev
        x = np.ascontiguousarray(iq, dtype=np.complex64)
        if x.ndim == 1:
            x = x[None, :]
        d = _dev.to_dev(x.view(np.float32).reshape(x.shape[0], x.shape[1], 2), self.device)
        h, lib = self._h, self._h.lib
        self._demod(d.data_ptr(), x.shape[0], x.shape[1], x.shape[1], None, _dev.stream_ptr(self.device))
        if self.afc:
            _lib.check(lib.tb200_etsi_afc(h.h, self.soft.data_ptr(), x.shape[0], self.soft_stride, self.n_soft.data_ptr(),
                                          self.theta.data_ptr(), _dev.stream_ptr(self.device)))
        self.acquire_device(x.shape[0], threshold)
        return self.sync_results(x.shape[0])

    def find_sequence(self, n_streams, sequence="q", threshold=0.8, min_gap_symbols=200, max_hits=None):
        """Positions (bit offsets into the soft bits of the last demodulation) where training sequence `sequence` —
        'n', 'p', 'q' (tail: slot boundary = position + 10), 'x' (extended, 30 bits) or 'y' (sync, 38 bits) — correlates
        at or above `threshold` (tb200_etsi_find_sequence).  Returns a list per stream of (position, correlation)."""
        import torch
        from . import _dev
        h, lib = self._h, self._h.lib
        mh = int(max_hits or self.max_bursts)
        pos = _dev.zeros((self.max_streams, mh), torch.int32, self.device)
        corr = _dev.zeros((self.max_streams, mh), torch.float32, self.device)
        nh = _dev.zeros((self.max_streams,), torch.int32, self.device)
        self.metric.zero_()
        _lib.check(lib.tb200_etsi_find_sequence(h.h, self.soft.data_ptr(), int(n_streams), self.soft_stride,
                                                self.n_soft.data_ptr(), SEQUENCE_IDS[sequence], float(threshold),
                                                int(min_gap_symbols), self.metric.data_ptr(), pos.data_ptr(), corr.data_ptr(),
                                                nh.data_ptr(), mh, _dev.stream_ptr(self.device)))
        pos, corr, nh = pos.cpu().numpy(), corr.cpu().numpy(), nh.cpu().numpy()
        return [[(int(pos[s, i]), float(corr[s, i])) for i in range(int(nh[s]))] for s in range(int(n_streams))]

    def carrier_offsets_hz(self, n_streams):
        """Carrier offset estimated by the last run per stream (Hz)."""
        return self.theta.cpu().numpy()[:n_streams].astype(np.float64) * 18000.0 / (2 * np.pi)

    def sync_results(self, n_streams):
        ns = self.n_sync.cpu().numpy()[:n_streams]
        pos = self.sync_pos.cpu().numpy()[:n_streams]
        bits = np.unpackbits(self.sync_bits.cpu().numpy(), axis=1)[:, :60].reshape(self.max_streams, self.max_sync, 60)
        ok = self.sync_crc.cpu().numpy().reshape(self.max_streams, self.max_sync).astype(bool)
        out = []
        for s in range(n_streams):
            rows = []
            for i in range(int(ns[s])):
                if ok[s, i]:
                    f = bsch_unpack(bits[s, i])
                    rows.append({"start": int(pos[s, i]), "fields": f,
                                 "scrambling_init": scrambling_init(f["mcc"], f["mnc"], f["colour_code"])})
            out.append(rows)
        return out

    def run_device(self, iq, n_samples, scramb_init, threshold=0.8, keep_symbols=False):
        """iq: CUDA float32 tensor [n_streams, stride, 2].  Asynchronous; results stay on the device."""
        return self.run_device_strided(iq.data_ptr(), iq.shape[0], n_samples, iq.shape[1], scramb_init, threshold,
                                       keep_symbols)

    def run_device_strided(self, iq_ptr, n_streams, n_samples, stride, scramb_init, threshold=0.8, keep_symbols=False):
        """Same on a raw device pointer: stream s starts `stride` complex samples after stream s-1 (the streams may
        be windows into longer rows)."""
        import torch
        from . import _dev
        h, lib = self._h, self._h.lib
        st = _dev.stream_ptr(self.device)
        if keep_symbols and self.symbols is None:
            self.symbols = _dev.zeros((self.max_streams, self.max_sym, 2), torch.float32, self.device)
        self.metric.zero_()
        self._demod(iq_ptr, n_streams, n_samples, stride, self.symbols.data_ptr() if keep_symbols else None, st)
        if self.afc:
            _lib.check(lib.tb200_etsi_afc(h.h, self.soft.data_ptr(), n_streams, self.soft_stride, self.n_soft.data_ptr(),
                                          self.theta.data_ptr(), st))
        _lib.check(lib.tb200_etsi_find_bursts(h.h, self.soft.data_ptr(), n_streams, self.soft_stride, self.n_soft.data_ptr(),
                                              float(threshold), self.metric.data_ptr(), self.blocks.data_ptr(),
                                              self.burst_pos.data_ptr(), self.n_bursts.data_ptr(), self.max_bursts, st))
        nblk = n_streams * self.max_bursts * 2
        _lib.check(lib.tb200_etsi_decode_batch(h.h, self.blocks.data_ptr(), nblk, 0, int(scramb_init) & 0xFFFFFFFF,
                                               self.bits.data_ptr(), self.crc_ok.data_ptr(), self.path_metric.data_ptr(), st))
        return n_streams

    def decode_aach_device(self, n_streams, scramb_init, generator=None):
        """After run_device: the broadcast block (bits 230..243 + 266..281) of every found burst through the (30, 14)
        soft ML decoder (tb200_etsi_extract_bits + tb200_etsi_aach_decode).  Results stay on the device in
        aach_words (int16 [S * max_bursts]), aach_metric, aach_margin."""
        from . import _dev
        h, lib = self._h, self._h.lib
        if self.aach is None or generator is not None:
            self.aach = AachDecoder(generator, self.device)
        _lib.check(lib.tb200_etsi_extract_bits(h.h, self.soft.data_ptr(), int(n_streams), self.soft_stride, self.n_soft.data_ptr(),
                                               self.burst_pos.data_ptr(), self.n_bursts.data_ptr(), self.max_bursts,
                                               230, 14, 266, 16, self.bb.data_ptr(), _dev.stream_ptr(self.device)))
        self.aach_words, self.aach_metric, self.aach_margin = self.aach.decode_device(
            self.bb.view(-1, 30)[:int(n_streams) * self.max_bursts], scramb_init)

    def run(self, iq, scramb_init, threshold=0.8, aach=False):
        """numpy complex64 [n_streams, n] -> list per stream of dicts(start, blocks [2,124] bits, crc_ok [2]); with
        aach=True also "aach" (the 14 access-assignment bits) and "aach_margin" per burst."""
        from . import _dev
        x = np.ascontiguousarray(iq, dtype=np.complex64)
        if x.ndim == 1:
            x = x[None, :]
        d = _dev.to_dev(x.view(np.float32).reshape(x.shape[0], x.shape[1], 2), self.device)
        S = self.run_device(d, x.shape[1], scramb_init, threshold)
        if aach:
            self.decode_aach_device(S, scramb_init)
            w = self.aach_words.cpu().numpy().view(np.uint16).astype(np.int64).reshape(S, self.max_bursts)
            ab = ((w[:, :, None] >> np.arange(13, -1, -1)[None, None, :]) & 1).astype(np.uint8)
            am = self.aach_margin.cpu().numpy().reshape(S, self.max_bursts)
        nb = self.n_bursts.cpu().numpy()[:S]
        pos = self.burst_pos.cpu().numpy()[:S]
        bits = np.unpackbits(self.bits.cpu().numpy(), axis=1)[:, :124].reshape(self.max_streams, self.max_bursts, 2, 124)
        ok = self.crc_ok.cpu().numpy().reshape(self.max_streams, self.max_bursts, 2).astype(bool)
        out = []
        for s in range(S):
            rows = [{"start": int(pos[s, i]), "blocks": bits[s, i], "crc_ok": ok[s, i]} for i in range(int(nb[s]))]
            if aach:
                for i, r in enumerate(rows):
                    r["aach"], r["aach_margin"] = ab[s, i], int(am[s, i])
            out.append(rows)
        return out


CHANNEL_SPACING = 25000.0
CHANNELIZER_TAPS_PER_BRANCH = 6


def channelizer_prototype(M, beta=8.0):
    """The prototype low-pass of tb200_channelize (Kaiser-windowed sinc, cutoff 25 kHz, 6 M taps, unit DC gain)."""
    L = CHANNELIZER_TAPS_PER_BRANCH * M
    t = np.arange(L) - (L - 1) / 2.0
    h = 2.0 / M * np.sinc(2.0 / M * t) * np.kaiser(L, beta)
    return h / h.sum()


class Channelizer:
    """Polyphase FIR channelizer: wideband complex64 at fs = M x 25 kHz -> [M, n_frames] at 50 kHz."""

    def __init__(self, n_channels, device=0):
        self.M = int(n_channels)
        self.device = device
        self.fs_in = self.M * CHANNEL_SPACING
        self.fs_out = 2 * CHANNEL_SPACING
        self._h = _lib.Handle(max_chunks=0, max_chunk_len=0, device=device, max_channels=self.M)

    def n_frames(self, n_samples):
        return (int(n_samples) - 1) // (self.M // 2) + 1 if n_samples > 0 else 0

    def run_device(self, iq, n_samples, out=None):
        """iq: CUDA float32 tensor [n_samples, 2]; returns CUDA float32 [M, n_frames, 2]."""
        import torch
        from . import _dev
        nf = self.n_frames(n_samples)
        if out is None:
            out = _dev.empty((self.M, nf, 2), torch.float32, self.device)
        h = self._h
        _lib.check(h.lib.tb200_channelize(h.h, iq.data_ptr(), int(n_samples), self.M, out.data_ptr(), int(out.shape[1]), nf,
                                          _dev.stream_ptr(self.device)))
        return out

    def run(self, x):
        from . import _dev
        x = np.ascontiguousarray(x, dtype=np.complex64)
        d = _dev.to_dev(x.view(np.float32).reshape(-1, 2), self.device)
        return self.run_device(d, len(x)).cpu().numpy().view(np.complex64)[..., 0]
