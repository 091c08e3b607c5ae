"""GPU parity for the SURVEY.md 8(f) rows N3 (scanner heuristics) and N4 (codec-frame builder): the CUDA path
through the C-ABI against the golden outputs of the unmodified reference and against the oracle."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN
from tests.golden.cases import scanner_cases, voice_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import __graft_entry__ as entry
    entry.build()
    from tetraear_b200 import signal, voice, _lib
    return signal, voice, _lib


def _close(name, key, got, want):
    """Integers/bools exact; the float fields within 1e-4 of full scale (dB: 1e-4 dB; ratios: matches may move by
    a couple of samples whose float32 phase sits within an ulp of a decision boundary)."""
    if key in ("is_tetra", "sync_detected", "frames_validated", "power_stable", "signal_present"):
        assert float(got) == want, (name, key, got, want)
    elif key == "power_db":
        assert abs(float(got) - want) <= 1e-4, (name, key, got, want)
    else:
        assert abs(float(got) - want) <= 1e-4, (name, key, got, want)


def test_detector_matches_reference_golden(mods):
    signal, _, _ = mods
    g = np.load(os.path.join(GOLDEN, "scanner.npz"))
    keys = [str(k) for k in g["keys"]]
    for name, fs, x in scanner_cases():
        det = signal.TetraSignalDetector(sample_rate=fs)
        r = det.analyze_signal(x)
        for k, w in zip(keys, g[name + "/analysis"]):
            _close(name, k, r[k], w)
        assert type(r["power_db"]).__name__ == str(g[name + "/power_dtype"]), name
        assert abs(float(det.calculate_power(x)) - float(g[name + "/power"])) <= 1e-4, name
        m = det.detect_tetra_modulation(x)
        assert float(m[0]) == g[name + "/mod"][0] and abs(m[1] - g[name + "/mod"][1]) <= 1e-4, (name, m)
        s = det.detect_sync_pattern(x)
        assert float(s[0]) == g[name + "/sync"][0] and s[1] == g[name + "/sync"][1], (name, s)
        st = [float(det.check_power_stability(x)), float(det.check_power_stability(x, 3))]
        assert st == list(g[name + "/stable"]), name


def test_detector_counts_match_oracle(mods):
    """The raw integers behind the ratios: sync stream best-match exact; modulation matches equal to the reference's
    count except for phase steps that sit within 4 float32 ulps of a decision boundary.  The reference's own count is
    not reproducible there: numpy evaluates float32 arctan2 with its AVX-512 SVML kernel on this build machine (38 % of
    its results differ from the correctly rounded value, which the CUDA kernel and any libm-based numpy produce), so
    the bound below is the number of such borderline steps, counted per dwell — zero borderline steps => exact."""
    from oracle import scanner as osc
    signal, _, _lib = mods
    cases = list(scanner_cases())
    rng = np.random.default_rng(3)
    for n in (1000, 1001, 4095, 50000):
        cases.append((f"rand_{n}", 2.4e6, (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)))
    for name, fs, x in cases:
        det = signal.TetraSignalDetector(sample_rate=fs)
        rec, _ = det.stats_batch(x, 5)
        rec = rec[0]
        m, ns = osc.modulation_matches(x)
        assert rec["n_steps"] == ns
        border = 0
        if ns:
            xn = x / (np.abs(x).max() + 1e-10)
            d = osc._wrapped_steps(xn).astype(np.float64)
            edges = np.arange(-7, 9, 2) * np.pi / 8                 # the open ends of the eight +-pi/8 windows
            gap = np.abs(d[:, None] - edges[None, :]).min(axis=1)
            border = int(np.count_nonzero(gap <= 4 * np.spacing(np.float32(np.pi))))
        assert abs(int(rec["matches"]) - m) <= border, (name, rec["matches"], m, border)
        bits = osc.sync_bits(x, fs)
        if bits is None:
            assert rec["n_sync_bits"] == 0, name
        else:
            assert rec["n_sync_bits"] == len(bits) and rec["sync_best"] == osc.sync_best(bits), name
        assert rec["max_abs"] == pytest.approx(float(np.abs(x).max()) if len(x) else 0.0, rel=1e-6)


def test_analyze_batch_equals_single_dwells(mods):
    signal, _, _ = mods
    from tetraear_b200 import synth
    cap = synth.make_capture(131072 * 4, 130, seed=77)
    rng = np.random.default_rng(9)
    dw = np.stack([cap[:131072], cap[131072:262144],
                   (0.05 * (rng.standard_normal(131072) + 1j * rng.standard_normal(131072))).astype(np.complex64),
                   np.concatenate([cap[262144:262144 + 65536], np.zeros(65536, np.complex64)])])
    det = signal.TetraSignalDetector(sample_rate=2.4e6)
    batch = det.analyze_batch(dw)
    assert len(batch) == 4
    for i in range(4):
        one = det.analyze_signal(dw[i])
        assert one == batch[i], (i, one, batch[i])
    assert batch[0]["frames_validated"] and batch[0]["is_tetra"] and not batch[2]["frames_validated"]


def test_voice_blocks_match_reference_golden(mods):
    from oracle import voice as ov
    _, voice, _lib = mods
    g = np.load(os.path.join(GOLDEN, "voice.npz"))
    dibits, pos, bit_rows = voice_cases()
    vb = voice.VoiceBlockBuilder()
    got = vb.from_symbols([{"position": p} for p in pos] + [{}], dibits)
    assert got[-1] is None
    for i, p in enumerate(pos):
        assert (got[i] is not None) == bool(g["valid"][i]), p
        if got[i] is not None:
            assert len(got[i]) == 1380 and np.array_equal(np.frombuffer(got[i], "<i2"), g["blocks"][i]), p
    fb, fp = vb.from_bits(bit_rows), vb.from_params(bit_rows)
    for i, (r, b, p) in enumerate(zip(bit_rows, fb, fp)):
        want = ov.from_bits(r)
        assert (b is None) == (want is None) == (not g["bit_valid"][i])
        if want is not None:
            assert np.array_equal(np.frombuffer(b, "<i2"), want)
            # pinned: the reference's own lines (continuous_capture.py:103-124) exec'd by tests/golden/make_golden.py
            assert np.array_equal(np.frombuffer(b, "<i2"), g["bit_blocks"][i]), i
            assert np.array_equal(np.frombuffer(p, "<i2"), ov.from_params(r))
    # every 6/7/3-bit field value of the parameter packing (int() truncation of float64 products)
    rows = []
    for v in range(128):
        r = np.zeros(432, np.uint8)
        for k in range(4):
            for b in range(6):
                r[108 * k + b] = (v >> b) & 1
            for b in range(7):
                r[108 * k + 6 + b] = ((v + k) >> b) & 1
            for b in range(3):
                r[108 * k + 13 + b] = ((v + 2 * k) >> b) & 1
        rows.append(r)
    for r, p in zip(rows, vb.from_params(rows)):
        assert np.array_equal(np.frombuffer(p, "<i2"), ov.from_params(r))
    # soft variant: sign and clamp
    soft = np.linspace(-2, 2, 432, dtype=np.float32)[None, :].repeat(2, 0)
    sb = vb.from_soft(soft)
    want = ov._place(np.clip(np.rint(soft[0] * np.float32(127.0)), -127, 127).astype(np.int16))
    assert np.array_equal(sb[0], want) and np.array_equal(sb[1], want)


def test_voice_blocks_from_decoded_frames(mods):
    """End to end: frames decoded on the GPU feed the builder; equals the oracle applied to the same positions."""
    from oracle import voice as ov
    signal, voice, _ = mods
    from tetraear_b200 import synth, core
    cap = synth.make_capture(131072, 130, seed=1234)
    sp = signal.SignalProcessor(2.4e6)
    d = sp.process(cap)
    frames = core.TetraDecoder(auto_decrypt=False).decode(d)
    assert frames
    got = voice.VoiceBlockBuilder().from_symbols(frames, d)
    for f, b in zip(frames, got):
        want = ov.from_symbols(f["position"], d)
        assert (b is None) == (want is None)
        if want is not None:
            assert np.array_equal(np.frombuffer(b, "<i2"), want)
