"""Seeded inputs shared by tests/golden/make_golden.py and the parity tests.

No dependency on /root/reference: the GPU box regenerates the same arrays from the same
seeds and checks their SHA against the one stored next to the golden outputs.
"""
import hashlib

import numpy as np

from tetraear_b200 import synth


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def process_cases():
    """(name, sample_rate, freq_offset, iq) — every input is reproducible from its seed."""
    cases = []
    cap = synth.make_capture(131072 * 2, 130, seed=1234)
    cases.append(("c1_chunk0", 2.4e6, 0.0, cap[:131072]))
    cases.append(("c1_chunk1", 2.4e6, 0.0, cap[131072:]))
    n = 131072
    sh = (cap[:n] * np.exp(2j * np.pi * 1234.5 * np.arange(n) / 2.4e6)).astype(np.complex64)
    cases.append(("c1_afc", 2.4e6, 1234.5, sh))
    rng = np.random.default_rng(5)
    noise = (0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    cases.append(("noise", 2.4e6, 0.0, noise))
    cases.append(("noise_afc", 2.4e6, -3210.0, noise))
    wide = synth.make_capture(1 << 20, 83 * 13, seed=21)
    cases.append(("wide_q83", 20e6, 0.0, wide))
    cases.append(("wide_q83_afc", 20e6, 700.0, wide))
    cases.append(("rate_1p8M", 1.8e6, 0.0, synth.make_capture(98000, 98, seed=3)))
    cases.append(("rate_240k_nodecim", 240000.0, 0.0, synth.make_capture(20000, 13, seed=4)))
    cases.append(("rate_2p0M_q8", 2.0e6, 0.0, synth.make_capture(90000, 111, seed=6)))
    cases.append(("rate_2p0M_q8_afc", 2.0e6, -812.5, synth.make_capture(90000, 111, seed=6)))
    cases.append(("c128_input", 2.4e6, 0.0, cap[:40000].astype(np.complex128)))
    cases.append(("ragged_40001", 2.4e6, 250.0, cap[:40001]))
    for m in (1, 2, 16, 27, 28, 29, 100, 150, 160, 161, 290, 500):
        cases.append((f"short_{m}", 2.4e6, 0.0, cap[1000:1000 + m]))
    cases.append(("zeros", 2.4e6, 0.0, np.zeros(5000, np.complex64)))
    return cases




def psd_chunks():
    """8 chunks of 131072 samples @2.4 MS/s: modem at -20 dBFS + CW spur at +7.3 kHz (C2),
    a near-silent chunk and a strong tone."""
    fs = 2.4e6
    n = 131072
    cap = synth.make_capture(n * 6, 130, seed=1234, amplitude=0.1, sigma=0.002)
    spur = (0.03 * np.exp(2j * np.pi * 7300.0 * np.arange(len(cap)) / fs)).astype(np.complex64)
    x = (cap + spur).astype(np.complex64)
    rng = np.random.default_rng(8)
    quiet = (1e-4 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    tone = (0.5 * np.exp(2j * np.pi * 3000.0 * np.arange(n) / fs)).astype(np.complex64) + quiet
    chunks = [x[i * n:(i + 1) * n] for i in range(6)] + [quiet, tone]
    return chunks


def burst_batch(n=3000, seed=99):
    d = synth.make_burst_batch(n, seed=seed)
    d[0] = 0   # degenerate rows the reference special-cases (all-zero / all-one data)
    d[1] = 3
    return d


def scanner_cases():
    """(name, sample_rate, samples) dwells for the scanner heuristics."""
    cap = synth.make_capture(262144, 130, seed=1234)
    rng = np.random.default_rng(17)
    n = 60000
    noise = (0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    tone = (0.4 * np.exp(2j * np.pi * 4500.0 * np.arange(n) / 2.4e6)).astype(np.complex64) + noise * 0.01
    cases = [("capture_256k", 2.4e6, cap), ("capture_c128", 2.4e6, cap[:60000].astype(np.complex128)),
             ("noise", 2.4e6, noise), ("tone", 2.4e6, tone.astype(np.complex64)),
             ("gap", 2.4e6, np.concatenate([cap[:30000], np.zeros(30000, np.complex64)])),
             ("burst_then_quiet", 2.4e6, np.concatenate([cap[:20000], (1e-4 * noise[:40000]).astype(np.complex64)])),
             ("short_900", 2.4e6, cap[:900]), ("short_1300", 2.4e6, cap[:1300]), ("short_4999", 2.4e6, cap[:4999]),
             ("rate_1p8M", 1.8e6, synth.make_capture(98000, 98, seed=3)),
             ("zeros", 2.4e6, np.zeros(12000, np.complex64))]
    return cases


def voice_cases():
    """(dibit stream, bit positions) for the codec-frame builder."""
    rng = np.random.default_rng(23)
    dibits = rng.integers(0, 4, 1007, dtype=np.uint8)
    pos = [0, 1, 2, 3, 300, 510, 1019, 2255, 2256, 2258, 2259, 3020, 5000]
    bits = [rng.integers(0, 2, m, dtype=np.uint8) for m in (510, 432, 431, 600, 0, 510)]
    return dibits, pos, bits


def afc_cases(n=40, seed=12):
    """(freqs_mhz, powers_db, tuned_mhz) triples for the auto-follow rule (ui/modern.py:5223-5257)."""
    rng = np.random.default_rng(seed)
    out = []
    for case in range(n):
        fc = 390.0 + rng.uniform(-1, 1)
        freqs = fc + (np.arange(2048) - 1024) * (2.4 / 2048)
        powers = rng.uniform(-110, -70, 2048)
        if case % 4:
            powers[1024 + int(rng.integers(-12, 12))] = rng.uniform(-65, -20)
        out.append((freqs, powers, fc + rng.uniform(-0.004, 0.004)))
    return out


def peak_cases(n=60, seed=21):
    """(freqs_mhz, powers_db, click_mhz, search_range_hz) for WaterfallWidget.find_peak_near (ui/modern.py:727-772):
    narrow spikes between bins, flat tops (zero denominator), peaks on the first / last bin, empty search windows,
    unequal array lengths and a non-uniform frequency axis."""
    rng = np.random.default_rng(seed)
    out = []
    for case in range(n):
        nb = 2048 if case % 5 else 257
        fc = 390.0 + rng.uniform(-1, 1)
        freqs = fc + (np.arange(nb) - nb // 2) * (2.4 / nb)
        powers = rng.uniform(-110, -70, nb)
        kind = case % 6
        pk = int(rng.integers(3, nb - 3))
        if kind == 0:
            powers[pk] = -30.0; powers[pk - 1] = -36.0; powers[pk + 1] = -33.0
        elif kind == 1:
            powers[pk - 1:pk + 2] = -40.0                       # flat top: denominator 0
        elif kind == 2:
            pk = 0 if case % 2 else nb - 1
            powers[pk] = -20.0                                   # edge bin: no interpolation
        elif kind == 3:
            freqs = freqs + np.cumsum(rng.uniform(0, 2e-4, nb))  # non-uniform axis
            powers[pk] = -25.0; powers[pk + 1] = -26.0
        elif kind == 4:
            powers = powers[: nb - 7]                            # lengths differ: truncated to the shorter
            pk = min(pk, nb - 12)
            powers[pk] = -22.0
        click = float(freqs[pk] + rng.uniform(-0.002, 0.002))
        rng_hz = float(rng.choice([500.0, 5000.0, 25000.0, 200000.0]))
        if kind == 5:
            click, rng_hz = float(freqs[-1] + 1.0), 1000.0      # nothing in range: the click comes back
        out.append((freqs, powers, click, rng_hz))
    return out


def stage_edge_cases():
    """Stage-level calls with awkward shapes / rates / dtypes: (method, kwargs, input).  Shared by the live
    oracle-vs-reference check (CPU, build container) and the CUDA-vs-oracle check (GPU box)."""
    rng = np.random.default_rng(77)

    def c(n, dtype=np.complex128, scale=1.0):
        return (scale * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(dtype)

    cases = []
    for n in (16, 17, 31, 100, 4999, 70001):
        for fs in (240000.0, 1.0e6, 2.4e6):
            cases.append(("filter_signal", {"bandwidth": 25000, "sample_rate": fs}, c(n)))
    cases.append(("filter_signal", {"bandwidth": 50000, "sample_rate": 240000.0}, c(3000, np.complex64)))
    cases.append(("filter_signal", {"bandwidth": 1.0, "sample_rate": 2.4e6}, c(2000)))             # cutoff clipped to 0.01
    cases.append(("filter_signal", {"bandwidth": 1e9, "sample_rate": 2.4e6}, c(2000)))             # cutoff clipped to 0.99
    cases.append(("filter_signal", {"bandwidth": 25000, "sample_rate": 240000.0}, c(15)))          # too short: returned as is
    for n, f, fs in ((1, 100.0, 2.4e6), (1000, -12345.678, 2.4e6), (70001, 0.5, 240000.0), (257, 1e6, 20e6)):
        cases.append(("frequency_shift", {"freq_offset": f, "sample_rate": fs}, c(n)))
    cases.append(("frequency_shift", {"freq_offset": 1000.0, "sample_rate": 2.4e6}, c(500, np.complex64)))
    cases.append(("frequency_shift", {"freq_offset": 1000.0, "sample_rate": 2.4e6}, rng.standard_normal(300)))   # real input
    for n, fs in ((13, 240000.0), (14, 240000.0), (1000, 240000.0), (1000, 18000.0), (1000, 9000.0), (50, 2.4e6),
                  (5000, 300000.0), (40000, 2.4e6)):
        cases.append(("extract_symbols", {"sample_rate": fs}, c(n)))
    for n in (2, 3, 1000):
        cases.append(("demodulate_dqpsk", {}, c(n)))
    cases.append(("demodulate_dqpsk", {}, np.zeros(10, complex)))
    cases.append(("demodulate_dqpsk", {}, c(500, np.complex64, 1e-6)))
    cases.append(("demodulate_dqpsk", {}, np.exp(1j * np.pi / 8 * np.arange(64) * 3)))              # steps exactly on 3 pi / 8
    return cases


def process_edge_cases():
    """process(samples, freq_offset) at unusual sample rates / lengths / dtypes: (fs, freq_offset, iq)."""
    out = []
    rates = [480001.0, 719999.0, 0.96e6, 1.2e6, 1.92e6, 2.16e6, 3.0e6, 200000.0, 100000.0, 30000.0, 15000.0, 2.4e6]
    for i, fs in enumerate(rates):
        sps = max(2, int(round(fs / 18000.0)))
        n = [9000, 20011, 33333][i % 3] if fs >= 100000.0 else 4000
        iq = synth.make_capture(n, sps, seed=700 + i, sigma=0.03)
        f = [0.0, 321.0, -1500.5][i % 3]
        if f:
            iq = (iq * np.exp(2j * np.pi * f * np.arange(n) / fs)).astype(np.complex64)
        out.append((fs, f, iq))
    cap = synth.make_capture(30000, 130, seed=777, sigma=0.03)
    out.append((2.4e6, 0.0, cap[:27]))
    out.append((2.4e6, 0.0, cap[:28]))
    out.append((2.4e6, 100.0, cap[:29]))
    out.append((2.4e6, 0.0, cap[:159]))           # 16 decimated samples: the shortest input the channel filter accepts
    out.append((2.4e6, 0.0, cap[:151]))           # 16 -> 15 decimated samples: filter skipped (returns input)
    out.append((2.4e6, -700.0, cap.astype(np.complex128)))
    out.append((1.8e6, 0.0, synth.make_capture(25000, 98, seed=778).astype(np.complex128)))
    return out


def spike_cases(n=48, seed=33):
    """(freqs_mhz, powers_db, click_mhz, search_range_hz, noise_floor_db) for WaterfallWidget.find_spike_band_near
    (ui/modern.py:774-857): carriers of 1 ... 40 bins on a noise floor, spikes touching the window edge, two spikes in
    one window, flat windows (threshold above everything but the peak), empty windows, unequal array lengths."""
    rng = np.random.default_rng(seed)
    out = []
    for case in range(n):
        nb = 2048 if case % 4 else 512
        fc = 392.0 + rng.uniform(-1, 1)
        freqs = fc + (np.arange(nb) - nb // 2) * (2.4 / nb)
        floor = float(rng.uniform(-105, -85))
        powers = floor + rng.uniform(-3, 3, nb)
        kind = case % 6
        pk = int(rng.integers(60, nb - 60))
        width = int(rng.choice([1, 2, 5, 11, 21, 40]))
        if kind in (0, 1, 2):
            powers[pk - width // 2: pk - width // 2 + width] = -40.0 + rng.uniform(-2, 2, width)
        if kind == 1:
            powers[pk + width + 3: pk + width + 6] = -38.0      # a second, stronger spike nearby
        if kind == 2:
            powers = powers[: nb - 9]
            pk = min(pk, nb - 80)
        rng_hz = float(rng.choice([3000.0, 12500.0, 25000.0, 100000.0]))
        click = float(freqs[pk] + rng.uniform(-0.003, 0.003))
        if kind == 3:                                           # spike cut by the window edge
            powers[pk: pk + 30] = -45.0
            click, rng_hz = float(freqs[pk + 29]), 12500.0
        if kind == 4:
            click, rng_hz = float(freqs[-1] + 1.0), 1000.0      # empty window
        noise = floor if case % 3 else -100.0
        out.append((freqs, powers, click, rng_hz, noise))
    return out
