"""Reference-matched synthetic TETRA-like modem.

Workload generator for the tests and ``bench.py`` — not part of the receive path.  It
produces IQ that the *reference* receiver (open-loop 13-phase timing pick, slicer with
decision regions centred on 0, +-pi/2, pi, 22-bit training sequence expected at bit 216,
CRC-16 field at data bits 200..215) demodulates and CRC-passes, so that "CRC-pass
bursts/s" is non-zero on both sides of the comparison.
"""
import numpy as np

TS1 = np.array([1, 1, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0, 1, 0, 0], dtype=np.uint8)
TS2 = np.array([0, 1, 1, 1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 1, 1, 0, 0], dtype=np.uint8)
SYNC_CONT = TS1
SYNC_DISC = np.array([0, 0, 1, 1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 0, 0, 1, 1], dtype=np.uint8)

# dibit -> differential phase at the centre of the reference slicer's decision regions
DPHI = np.array([0.0, np.pi / 2, -np.pi / 2, np.pi])


def crc16_bits(bits):
    """CRC-16-CCITT (0x1021, init 0xFFFF, MSB first, no xorout) over a 0/1 sequence."""
    crc = 0xFFFF
    for b in bits:
        crc ^= (int(b) & 1) << 15
        crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc


def _payload_index():
    """burst-bit index of each of the 200 CRC payload bits (data = b[0:108] ++ b[122:230])."""
    return np.concatenate([np.arange(0, 108), np.arange(122, 214)])


_SOLVE_CACHE = {}


def _crc_solver(free_burst_bits):
    """Precompute the GF(2) system mapping 16 free payload bits to the CRC value."""
    key = tuple(free_burst_bits)
    if key in _SOLVE_CACHE:
        return _SOLVE_CACHE[key]
    pidx = _payload_index()
    where = {int(b): i for i, b in enumerate(pidx)}
    cols = [where[int(b)] for b in free_burst_bits]
    zero = crc16_bits(np.zeros(200, np.uint8))
    basis = []
    for c in cols:
        e = np.zeros(200, np.uint8)
        e[c] = 1
        basis.append(crc16_bits(e) ^ zero)
    # Gaussian elimination over GF(2) on a 16x16 matrix stored as ints (row r = CRC bit r)
    A = [[(basis[c] >> (15 - r)) & 1 for c in range(16)] for r in range(16)]
    inv = _gf2_inverse(A)
    _SOLVE_CACHE[key] = (cols, inv)
    return cols, inv


def _gf2_inverse(A):
    n = len(A)
    M = [row[:] + [1 if i == j else 0 for j in range(n)] for i, row in enumerate(A)]
    for col in range(n):
        piv = next((r for r in range(col, n) if M[r][col]), None)
        if piv is None:
            raise ValueError("singular CRC system; pick other free bits")
        M[col], M[piv] = M[piv], M[col]
        for r in range(n):
            if r != col and M[r][col]:
                M[r] = [a ^ b for a, b in zip(M[r], M[col])]
    return [row[n:] for row in M]


FREE_BITS = tuple(range(196, 212))  # burst bits reserved to force the CRC


def make_burst(rng, kind="pass", sync_type=False):
    """One 510-bit burst.

    kind: "pass"  CRC matches exactly;  "near" 1-2 flipped CRC bits (still passes, the
    reference tolerates <=2 errors);  "random" unconstrained payload (CRC almost surely fails).
    sync_type: put SYNC_CONT at bits 255..276 so the burst is typed Synchronization.
    """
    b = rng.integers(0, 2, 510, dtype=np.uint8)
    b[0:4] = 0                      # MAC-RESOURCE, clear -> no decrypt attempts
    b[216:238] = TS1                # where TetraDecoder.decode expects the TS
    if sync_type:
        b[255:277] = SYNC_CONT
    else:
        # keep bits 255..276 away from both 22-bit sync patterns (> 0.8 would retype it)
        for _ in range(64):
            w = b[255:277]
            if max(np.sum(w == SYNC_CONT), np.sum(w == SYNC_DISC)) / 22 <= 0.7:
                break
            b[255:277] = rng.integers(0, 2, 22, dtype=np.uint8)
    if kind == "random" or sync_type:
        return b
    # keep MAC-RESOURCE length field (data bits 29..34) small so parse_mac_pdu accepts it
    b[29:35] = [0, 0, 0, 1, 0, 0]
    cols, inv = _crc_solver(FREE_BITS)
    pidx = _payload_index()
    b[list(FREE_BITS)] = 0
    payload = b[pidx]
    target = 0
    for v in b[214:230]:
        target = (target << 1) | int(v)
    delta = crc16_bits(payload) ^ target
    dvec = [(delta >> (15 - r)) & 1 for r in range(16)]
    sol = [sum(inv[i][j] & dvec[j] for j in range(16)) & 1 for i in range(16)]
    for i, fb in enumerate(FREE_BITS):
        b[fb] = sol[i]
    assert crc16_bits(b[pidx]) == target
    if kind == "near":
        # flip 1-2 of the two CRC bits that are not training-sequence bits (214, 215)
        nflip = int(rng.integers(1, 3))
        for fb in (214, 215)[:nflip]:
            b[fb] ^= 1
    return b


def bits_to_dibits(bits):
    bits = np.asarray(bits, dtype=np.uint8)
    return (bits[0::2] << 1) | bits[1::2]


def modulate(dibits, samples_per_symbol, sigma=0.02, amplitude=1.0, rng=None, n_samples=None,
             start_phase=0.0):
    """Hold-per-symbol differential modem: phase accumulates DPHI[dibit], each symbol held
    for ``samples_per_symbol`` samples (q * 13: 130 @2.4 MS/s, 1079 @20 MS/s); + AWGN; complex64."""
    rng = np.random.default_rng(0) if rng is None else rng
    phase = start_phase + np.concatenate([[0.0], np.cumsum(DPHI[np.asarray(dibits, dtype=np.int64)])])
    sym = amplitude * np.exp(1j * phase)
    x = np.repeat(sym, samples_per_symbol)
    if n_samples is not None:
        if len(x) < n_samples:
            x = np.concatenate([x, np.zeros(n_samples - len(x), dtype=x.dtype)])
        x = x[:n_samples]
    if sigma > 0:
        x = x + sigma * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))
    return x.astype(np.complex64)


def make_capture(n_samples, samples_per_symbol=130, seed=1234, sigma=0.02, amplitude=1.0,
                 mix=(0.94, 0.03, 0.03), return_bits=False):
    """Continuous back-to-back 510-bit bursts filling ``n_samples`` samples.

    mix = probabilities of (pass, near, random) bursts.  Returns complex64 IQ (and the
    transmitted bit stream when return_bits)."""
    rng = np.random.default_rng(seed)
    n_sym = n_samples // samples_per_symbol + 2
    n_bursts = n_sym // 255 + 1
    kinds = rng.choice(3, size=n_bursts, p=np.asarray(mix) / np.sum(mix))
    bits = np.concatenate([make_burst(rng, ("pass", "near", "random")[k]) for k in kinds])
    dibits = bits_to_dibits(bits)[:n_sym]
    iq = modulate(dibits, samples_per_symbol, sigma=sigma, amplitude=amplitude, rng=rng,
                  n_samples=n_samples)
    return (iq, bits) if return_bits else iq


def make_burst_batch(n_bursts, seed=99, fractions=(0.5, 0.25, 0.25), sync_fraction=0.01):
    """C3 input: [n_bursts, 255] uint8 dibits — 50 % CRC-pass, 25 % 1-2 flipped CRC bits,
    25 % random; ~1 % typed Synchronization."""
    rng = np.random.default_rng(seed)
    # build a pool of distinct bursts, then tile: python-level burst crafting is slow
    pool = min(n_bursts, 4096)
    kinds = rng.choice(3, size=pool, p=np.asarray(fractions) / np.sum(fractions))
    is_sync = rng.random(pool) < sync_fraction
    rows = np.stack([bits_to_dibits(make_burst(rng, ("pass", "near", "random")[k], bool(s)))
                     for k, s in zip(kinds, is_sync)])
    if pool == n_bursts:
        return rows
    idx = rng.integers(0, pool, n_bursts)
    out = rows[idx]
    # decorrelate the tiled copies a little: random bits in the tail (bits 230..509 are
    # outside data/CRC for normal bursts) — keeps CRC verdicts of normal bursts unchanged
    tail = rng.integers(0, 4, (n_bursts, 255 - 139), dtype=np.uint8)
    keep = out[:, 127:139].copy()  # bits 254..277 region (sync-pattern window) untouched
    out[:, 139:] = tail
    out[:, 127:139] = keep
    return out


def make_wideband(n_samples, carriers, sample_rate=20e6, n_channels=800, samples_per_symbol=1079, seed=5,
                  sigma=0.02, amplitude=0.05):
    """C4 input: a wideband capture with independent reference-matched carriers.

    ``carriers``: channel numbers k (centre k * sample_rate / n_channels, negative allowed), each modulated with its own
    back-to-back bursts at ``samples_per_symbol`` = q * sps of the wideband rate (83 * 13 = 1079 at 20 MS/s), summed,
    plus complex AWGN of ``sigma`` per component over the whole band.  complex64."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64)
    x = sigma * (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples))
    for i, k in enumerate(carriers):
        base = make_capture(n_samples, samples_per_symbol, seed=seed * 1000 + i, sigma=0.0, amplitude=amplitude).astype(np.complex128)
        x += base * np.exp(2j * np.pi * (k / n_channels) * t + 1j * rng.uniform(0, 2 * np.pi))
    return x.astype(np.complex64)
