"""The scenarios of the reference's own unit / integration tests (tests/unit/test_signal_processor.py,
test_tetra_decoder.py, test_tetra_protocol.py, tests/integration/test_frequency_scanner.py), run against the
B200 classes — and, beyond the reference's shape/dtype assertions, compared with the oracle wherever the call
produces numbers.  Written from the behaviours listed in SURVEY.md §4; no reference code is used."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as entry
    entry.build()
    from tetraear_b200 import signal, core
    return signal, core


@pytest.fixture(scope="module")
def iq():
    """10 ms of a noisy baseband carrier at 2.4 MS/s, float64-based complex (like the reference's fixture)."""
    rng = np.random.default_rng(2024)
    n = 24000
    return np.ones(n, complex) + 0.1 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))


TS1 = np.array([1, 1, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0, 1, 0, 0])


# ------------------------------------------------------------------------------------ SignalProcessor
def test_processor_attributes(api):
    sp = api[0].SignalProcessor()
    assert sp.sample_rate == 2.4e6 and sp.symbol_rate == 18000 and sp.samples_per_symbol == 133 and sp.symbols is None
    sp = api[0].SignalProcessor(sample_rate=1.0e6)
    assert sp.sample_rate == 1.0e6 and sp.symbol_rate == 18000 and sp.samples_per_symbol == 55


def test_processor_empty_and_tiny_inputs(api):
    sp = api[0].SignalProcessor()
    e = np.array([])
    assert len(sp.filter_signal(e)) == 0 and len(sp.extract_symbols(e)) == 0
    d = sp.demodulate_dqpsk(e)
    assert isinstance(d, np.ndarray) and len(d) == 0
    assert len(sp.demodulate_dqpsk(np.array([1.0 + 1.0j]))) == 0          # needs two samples
    assert len(sp.process(e)) == 0


def test_processor_stage_methods_match_oracle(api, iq):
    sp = api[0].SignalProcessor()
    r = sp.resample(iq, 1.2e6)
    assert isinstance(r, np.ndarray) and np.iscomplexobj(r) and len(r) == 12000
    for bw in (25000, 50000):
        y = sp.filter_signal(iq, bandwidth=bw)
        assert isinstance(y, np.ndarray) and len(y) == len(iq)
        assert np.abs(y - oracle.filter_signal(iq, bw, 2.4e6)).max() <= 1e-9
    s = sp.frequency_shift(iq, 1000)
    assert np.iscomplexobj(s) and len(s) == len(iq)
    assert np.abs(s - oracle.frequency_shift(iq, 1000, 2.4e6)).max() <= 1e-12
    assert len(sp.frequency_shift(iq, 0)) == len(iq)
    d = sp.demodulate_dqpsk(iq)
    assert d.dtype == np.uint8 and len(d) == len(iq) - 1 and d.max() <= 3
    assert np.array_equal(d, oracle.demodulate_dqpsk(iq))
    for fs in (None, 1.0e6):
        sym = sp.extract_symbols(iq) if fs is None else sp.extract_symbols(iq, sample_rate=fs)
        want, _ = oracle.extract_symbols(iq, 2.4e6 if fs is None else fs)
        assert np.iscomplexobj(sym) and len(sym) == len(want) > 0 and np.abs(sym - want).max() <= 1e-12


# ------------------------------------------------------------------------------------ TetraDecoder
def test_decoder_attributes_and_keys(api):
    core = api[1]
    dec = core.TetraDecoder()
    assert dec.auto_decrypt is True and core.TetraDecoder(auto_decrypt=False).auto_decrypt is False
    assert dec.FRAME_LENGTH == 510 and len(dec.SYNC_PATTERN) > 0
    assert "TEA1" in dec.common_keys and "TEA2" in dec.common_keys and len(dec.common_keys["TEA1"]) > 0
    dec.set_keys(["00112233445566778899", "00112233445566778899AABBCCDDEEFF"])
    assert len(dec.user_keys) > 0
    dec.set_keys(["invalid_key_format"])                                    # logs, never raises


def test_decoder_bits_and_sync(api):
    dec = api[1].TetraDecoder()
    bits, mapped = dec.symbols_to_bits(np.array([0, 1, 2, 3]))
    assert list(bits) == [0, 0, 0, 1, 1, 0, 1, 1] and list(mapped) == [0, 1, 2, 3]
    s8 = np.array([0, 1, 2, 3, 4, 5, 6, 7])
    bits, mapped = dec.symbols_to_bits(s8)
    ob, om = oracle.symbols_to_bits(s8)
    assert np.array_equal(bits, ob) and np.array_equal(mapped, om) and len(bits) == 2 * len(mapped)
    assert len(dec.find_sync(np.array([0] * 10))) == 0                      # fewer than 22 bits
    b = np.zeros(100, dtype=np.int64)
    b[20:42] = TS1
    pos, mc = dec.find_sync(b, threshold=0.8, return_max_corr=True)
    assert isinstance(pos, list) and isinstance(mc, float)
    opos, omc = oracle.find_sync(b.astype(np.uint8), 0.8)
    assert pos == opos == [20] and mc == omc == 1.0
    rnd = np.random.default_rng(3).integers(0, 2, 100)
    assert dec.find_sync(rnd, threshold=0.9) == oracle.find_sync(rnd.astype(np.uint8), 0.9)[0]


def test_decoder_decode_frame(api):
    dec = api[1].TetraDecoder(auto_decrypt=False)
    assert dec.decode_frame(np.array([0] * 100), start_pos=0) is None       # < 510 bits
    r = dec.decode_frame(np.array([0, 1] * 300), start_pos=0)
    assert r is None or isinstance(r, dict)
    if r is not None:
        assert {"type", "number", "bits", "position", "encrypted"} <= set(r)


# ------------------------------------------------------------------------------------ TetraProtocolParser
def test_parser_burst_paths(api):
    core = api[1]
    p = core.TetraProtocolParser()
    assert p.current_frame_number == 0 and p.current_multiframe == 0 and p.stats["total_bursts"] == 0
    assert p.parse_burst(np.array([0] * 100), slot_number=0) is None        # < 255 symbols
    sym = np.zeros(255, dtype=np.int64)
    sym[108:130] = TS1
    b = p.parse_burst(sym, slot_number=0)
    assert isinstance(b, core.TetraBurst) and p.stats["total_bursts"] == 1
    want = oracle.parse_bursts((sym & 3).astype(np.uint8)[None, :])
    assert b.burst_type.value == int(want["burst_type"][0]) and bool(b.crc_ok) == bool(want["crc_ok"][0])
    assert np.array_equal(np.asarray(b.data_bits), want["data_bits"][0, :int(want["n_data"][0])])
    bits = np.zeros(255, dtype=np.int64)
    bits[108:130] = TS1
    assert isinstance(p._detect_burst_type(bits), core.BurstType)
    assert bool(p._check_sync_pattern(np.array(p.SYNC_CONTINUOUS_DOWNLINK))) is True
    assert bool(p._check_sync_pattern(np.array([0] * 22))) is False
    alt = np.array([0, 1] * 130)
    assert len(p._extract_training_sequence(alt, core.BurstType.NormalDownlink)) == 14
    assert len(p._extract_data_bits(alt, core.BurstType.NormalDownlink)) == 216


def test_parser_crc(api):
    p = api[1].TetraProtocolParser()
    bits = np.array([0, 1] * 100)
    crc = p._calculate_crc16(bits[:-16])
    assert len(crc) == 16 and all(int(v) in (0, 1) for v in crc)
    assert int("".join(str(int(v)) for v in crc), 2) == oracle.crc16(bits[:-16].astype(np.uint8))
    bits[-16:] = crc
    assert bool(p._check_crc(bits)) is True
    bits[3] ^= 1                                                             # one payload error: CRC differs by > 2 bits
    flipped = bits.copy()
    assert isinstance(bool(p._check_crc(flipped)), bool)
    short = p._calculate_crc16(np.array([0, 1, 0, 1] * 10))
    assert len(short) == 16


def test_parser_mac_pdu(api):
    core = api[1]
    p = core.TetraProtocolParser()
    assert p.parse_mac_pdu(np.array([0] * 4)) is None                       # < 8 bits
    r = p.parse_mac_pdu(np.array([0, 0] + [0, 1] * 50))
    assert r is None or isinstance(r, core.MacPDU)
    assert (r is not None) == oracle.mac_pdu_valid(np.array([0, 0] + [0, 1] * 50, dtype=np.uint8))


# ------------------------------------------------------------------------------------ TetraSignalDetector
def test_detector_scenarios(api, iq):
    from oracle import scanner as osc
    det = api[0].TetraSignalDetector()
    assert det.sample_rate == 2.4e6 and det.noise_floor == -45 and det.bottom_threshold == -85
    d2 = api[0].TetraSignalDetector(sample_rate=1.8e6, noise_floor=-50, bottom_threshold=-90)
    assert (d2.sample_rate, d2.noise_floor, d2.bottom_threshold) == (1.8e6, -50, -90)
    assert det.calculate_power(np.array([])) == -85.0
    assert abs(det.calculate_power(iq) - osc.calculate_power(iq)) <= 1e-9
    ok, conf = det.detect_tetra_modulation(iq)
    assert (ok, conf) == osc.detect_tetra_modulation(iq)
    assert det.detect_tetra_modulation(np.ones(100, complex)) == (False, 0.0)
    assert det.detect_sync_pattern(iq) == osc.detect_sync_pattern(iq, 2.4e6)
    rng = np.random.default_rng(8)
    noise = rng.standard_normal(24000) + 1j * rng.standard_normal(24000)
    assert det.detect_tetra_modulation(noise)[1] == osc.detect_tetra_modulation(noise)[1]
    a = det.analyze_signal(iq)
    want = osc.analyze_signal(iq, 2.4e6)
    assert set(a) == set(want)
    for k in want:
        assert a[k] == want[k] or abs(float(a[k]) - float(want[k])) <= 1e-9, k


# ------------------------------------------------------------------------------------ end-to-end scenarios
def test_stage_chain_and_full_pipeline(api, iq):
    """tests/integration/test_end_to_end.py: filter -> demodulate / extract -> bits -> burst, each link vs the oracle."""
    signal, core = api
    sp = signal.SignalProcessor(sample_rate=2.4e6)
    filtered = sp.filter_signal(iq)
    ofil = oracle.filter_signal(iq, 25000, 2.4e6)
    assert len(filtered) == len(iq) and np.abs(filtered - ofil).max() <= 1e-9
    symbols = sp.demodulate_dqpsk(filtered)
    assert len(symbols) == len(iq) - 1
    osym = oracle.demodulate_dqpsk(ofil)
    assert np.count_nonzero(symbols != osym) <= 2          # the two filter outputs differ by ~1e-13: slicer ties only
    assert len(sp.extract_symbols(filtered)) > 0
    dec = core.TetraDecoder()
    bits, mapped = dec.symbols_to_bits(symbols)
    ob, om = oracle.symbols_to_bits(symbols)
    assert np.array_equal(bits, ob) and np.array_equal(mapped, om)
    p = core.TetraProtocolParser()
    burst = p.parse_burst(symbols[:255], slot_number=0)
    assert burst is not None and hasattr(burst, "data_bits")
    want = oracle.parse_bursts(np.asarray(symbols[:255], dtype=np.uint8)[None, :])
    assert bool(burst.crc_ok) == bool(want["crc_ok"][0])


def test_decoding_and_statistics_scenarios(api):
    core = api[1]
    dec = core.TetraDecoder()
    pattern31 = [0, 1, 0, 1, 1, 0, 0, 1, 1, 1, 0, 0, 0, 1, 0, 0, 1, 0, 1, 1, 0, 0, 1, 1, 1, 0, 0, 0, 1, 0, 0]
    bits = np.array(pattern31 + [i % 2 for i in range(510 - 31)])
    pos = dec.find_sync(bits, threshold=0.7)
    assert pos == oracle.find_sync(bits.astype(np.uint8), 0.7)[0]
    if pos:
        r = dec.decode_frame(bits, start_pos=pos[0])
        assert r is None or isinstance(r, dict)
    p = core.TetraProtocolParser()
    sym = np.array([0, 1, 2, 3] * 64)
    b = p.parse_burst(sym, slot_number=0)
    assert b is None or hasattr(b, "burst_type")
    for _ in range(3):
        p.parse_burst(sym, slot_number=0)
    assert p.stats["total_bursts"] == 4 and p.stats["crc_pass"] + p.stats["crc_fail"] == 4
    assert p.get_statistics()["total_bursts"] == 4
    p.fragment_buffer = bytearray(b"Hello " + b"World")
    assert len(p.fragment_buffer) == 11


# ------------------------------------------------------------------------------------ awkward stage-level inputs
@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_stage_edge_cases(api, mode):
    """Shapes / rates / dtypes off the beaten path (tests/golden/cases.py::stage_edge_cases); the oracle is pinned to
    the live reference on the same list by tests/test_oracle_vs_reference.py::test_stage_edge_cases_live."""
    from tests.golden.cases import stage_edge_cases
    sp = api[0].SignalProcessor(2.4e6, mode=mode)
    for i, (method, kw, x) in enumerate(stage_edge_cases()):
        fs = kw.get("sample_rate", 2.4e6)
        got = getattr(sp, method)(x, **kw)
        if method == "filter_signal":
            want = oracle.filter_signal(x, kw["bandwidth"], fs)
        elif method == "frequency_shift":
            want = oracle.frequency_shift(x, kw["freq_offset"], fs)
        elif method == "extract_symbols":
            want = oracle.extract_symbols(x, fs)[0]
        else:
            want = oracle.demodulate_dqpsk(x)
        assert len(got) == len(want), (i, method, kw, len(got), len(want))
        if not len(want):
            continue
        if method == "demodulate_dqpsk":
            bad = np.nonzero(np.asarray(got) != np.asarray(want))[0]
            # identical except where a phase step sits on a slicer boundary to the last ulp (the crafted 3 pi / 8 case)
            xs = np.asarray(x, dtype=np.complex128)
            for k in bad:
                ph = np.angle(xs[k + 1] * np.conj(xs[k]))
                assert np.min(np.abs(np.abs(ph) - np.array([3, 5]) * np.pi / 8)) < 1e-9, (i, k)
        else:
            tol = 1e-9 if (mode == "exact" or method != "filter_signal") else 1e-9
            assert np.abs(np.asarray(got) - np.asarray(want)).max() <= tol * max(1.0, np.abs(want).max()), (i, method, kw)


@pytest.mark.parametrize("mode", ["exact", "fast", "fast32"])
def test_process_edge_cases(api, mode):
    """process() at unusual sample rates (q = 2, 3, 4, 5, 8, 9, 12, none; sps down to 1 and 0), lengths around the
    decimator / filter pad limits and complex128 input; oracle pinned by test_process_edge_cases_live."""
    from tests.golden.cases import process_edge_cases
    for i, (fs, f, iq) in enumerate(process_edge_cases()):
        sp = api[0].SignalProcessor(fs, mode=mode)
        got = sp.process(iq, f)
        want, sym, _ = oracle.process(iq, fs, f)
        assert got.dtype == np.uint8 and len(got) == len(want), (i, fs, len(got), len(want))
        assert len(sp.symbols) == len(sym), (i, fs)
        if mode == "exact":
            assert np.array_equal(got, want), (i, fs, f)
            if len(sym):
                assert np.abs(sp.symbols - sym).max() <= 1e-9 * max(1.0, np.abs(sym).max()), (i, fs)
        else:
            if len(sym):
                assert np.abs(sp.symbols - sym).max() <= 1e-4 * max(1.0, np.abs(sym).max()), (i, fs, f)
            bad = np.nonzero(got != want)[0]
            ph = np.angle(sym[1:] * np.conj(sym[:-1])) if len(sym) > 1 else np.zeros(0)
            for k in bad:                    # only decisions lying on a slicer boundary may differ
                assert np.min(np.abs(np.abs(ph[k]) - np.array([3, 5]) * np.pi / 8)) < 1e-3, (i, fs, k)
            assert len(bad) <= 0.01 * max(len(want), 1) + 1, (i, fs, len(bad))


@pytest.mark.parametrize("mode,chunk_len,n_chunks,lanes", [("exact", 50001, 5, 2), ("fast32", 50001, 5, 2), ("fast", 33000, 7, 3),
                                                         ("fast32", 131072, 1, 2), ("fast32", 7000, 9, 4)])
def test_pipeline_odd_geometries(api, mode, chunk_len, n_chunks, lanes):
    """The batched pipeline with chunk lengths that are not multiples of q, fewer chunks than lanes, short chunks."""
    from tetraear_b200 import pipeline, synth
    cap = synth.make_capture(chunk_len * n_chunks, 130, seed=31 + n_chunks, mix=(0.7, 0.15, 0.15)).reshape(n_chunks, chunk_len)
    pipe = pipeline.ReceivePipeline(2.4e6, chunk_len, max_chunks=n_chunks + 2, mode=mode, n_lanes=lanes)
    nfr, frames, stats = pipe.run_host(cap)
    ndib = pipe.n_dibits.cpu().numpy()
    dib = pipe.dibits.cpu().numpy()
    tot = [0, 0, 0]
    for c in range(n_chunks):
        od, _, _ = oracle.process(cap[c], 2.4e6, 0.0)
        assert ndib[c] == len(od) and np.array_equal(dib[c, :ndib[c]], od), (mode, c)
        ofr, ost = oracle.decode(od)
        kept = [r for r in frames[c, :nfr[c]] if r["kept"]]
        assert [(int(r["number"]), int(r["type"]), bool(r["crc_ok"])) for r in kept] == \
               [(f["number"], f["type"], f["burst_crc"]) for f in ofr], (mode, c)
        for k, key in enumerate(("total_bursts", "crc_pass", "crc_fail")):
            tot[k] += ost[key]
    assert list(stats) == tot
    # device-resident entry point, twice (double-buffered dibit rows), then joined
    import torch
    dev = torch.from_numpy(cap.view(np.float32).reshape(n_chunks, chunk_len, 2)).cuda()
    pipe.stats.zero_()
    pipe.run_device(dev, join=False)
    pipe.run_device(dev, join=False)
    pipe.join()
    torch.cuda.synchronize()
    assert list(pipe.stats.cpu().numpy()) == [2 * t for t in tot]
    assert np.array_equal(pipe.n_dibits.cpu().numpy()[:n_chunks], ndib[:n_chunks])
    assert np.array_equal(pipe.dibits.cpu().numpy()[:n_chunks], dib[:n_chunks])


def test_integration_md_ctypes_binding(api):
    """The reference-side ctypes stub printed in INTEGRATION.md (section B) is executed as written (only the library
    path is substituted) and its two functions are compared with the oracle."""
    import os
    import re
    from tests.conftest import ROOT
    from tetraear_b200 import _lib, synth
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## B. Bind the C-ABI directly"):]
    code = re.search(r"```python\n(.*?)```", sec, re.S).group(1)
    assert 'C.CDLL("libtetra_b200.so")' in code
    ns = {}
    exec(code.replace('C.CDLL("libtetra_b200.so")', f'C.CDLL({_lib.LIB_PATH!r})'), ns)
    iq = synth.make_capture(131072, 130, seed=1234)
    d = ns["process"](iq)
    want, _, _ = oracle.process(iq, 2.4e6, 0.0)
    assert np.array_equal(d, want)                 # FAST mode handle: identical hard decisions on a signal chunk
    shifted = (iq * np.exp(2j * np.pi * 800.0 * np.arange(len(iq)) / 2.4e6)).astype(np.complex64)
    assert np.array_equal(ns["process"](shifted, 800.0), oracle.process(shifted, 2.4e6, 800.0)[0])
    rows = synth.make_burst_batch(32, seed=5)
    for row in rows:
        r = ns["parse_burst"](row)
        o = oracle.parse_bursts(row[None, :])
        assert r["burst_type"] == int(o["burst_type"][0]) and r["crc_ok"] == bool(o["crc_ok"][0])
        nd = int(o["n_data"][0])
        if nd == 216:
            assert np.array_equal(r["data_bits"][:216], o["data_bits"][0, :216])


@pytest.mark.parametrize("alpha,n_fft", [(0.15, 512), (0.05, 512), (0.01, 256), (1.0, 256), (0.5, 1024)])
def test_ema_smoothing_factors(api, alpha, n_fft):
    """The dB-domain EMA (ui/modern.py:929-935) for smoothing factors other than the GUI's 0.15: the grouped kernel
    sizes its warm-up from alpha (single exact group when the history would be too long), incl. a second call
    that continues from the carried state."""
    rng = np.random.default_rng(4)
    nfr = 3000
    x = (0.1 * (rng.standard_normal(nfr * n_fft) + 1j * rng.standard_normal(nfr * n_fft))).astype(np.complex64)
    an = api[0].SpectrumAnalyzer(2.4e6, n_fft, ema_alpha=alpha)
    r1 = an.run(x[: 2000 * n_fft], denoise=True)
    r2 = an.run(x[2000 * n_fft:], denoise=True)
    power = np.concatenate([r1["power"], r2["power"]]).astype(np.float64)
    got = np.concatenate([r1["smoothed"], r2["smoothed"]]).astype(np.float64)
    want = oracle.ema_denoise(power, alpha)
    assert np.abs(got - want).max() <= 2e-4, (alpha, np.abs(got - want).max())


@pytest.mark.parametrize("fs,n_fft", [(2.4e6, 2048), (1.8e6, 2048), (240000.0, 2048), (50000.0, 1024), (20e6, 4096),
                                      (30000.0, 256), (2.4e6, 64)])
def test_spectrum_band_statistics_at_other_rates(api, fs, n_fft):
    """The band / noise window arithmetic of the spectrum block (ui/modern.py:1956-2010) away from 2.4 MS/s:
    band wider than the FFT, no noise bins left (noise floor -100), band of zero bins."""
    rng = np.random.default_rng(int(fs) % 1000 + n_fft)
    n = n_fft * 3
    x = (0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    x += (0.4 * np.exp(2j * np.pi * (fs / n_fft * 3.3) * np.arange(n) / fs)).astype(np.complex64)
    an = api[0].SpectrumAnalyzer(fs, n_fft)
    r = an.run(x, denoise=False)
    for i in range(3):
        _, want_p = oracle.psd_frame(x[i * n_fft:(i + 1) * n_fft], fs, n_fft)
        m = want_p > -120
        assert np.abs(r["power"][i].astype(np.float64) - want_p)[m].max() <= 1e-3
        o = oracle.band_stats(want_p, fs)
        st = r["stats"][i]
        if o is None:                       # zero-width band: the reference block would divide by zero / skip
            continue
        assert st["peak_idx"] == o["peak_idx"] and st["signal_present"] == o["signal_present"], (fs, i)
        for k in ("signal_power", "peak_power", "peak_freq_offset", "noise_floor", "snr", "afc_offset"):
            assert abs(st[k] - o[k]) <= 2e-3, (fs, i, k, st[k], o[k])


def test_handles_on_two_devices_in_one_process(api):
    """Several handles on different GPUs in one process: every entry point runs on its handle's device and leaves the
    caller's current device alone (skipped on a single-GPU box)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from tetraear_b200 import synth
    iq = synth.make_capture(60000, 130, seed=5)
    want, _, _ = oracle.process(iq, 2.4e6, 0.0)
    before = torch.cuda.current_device()
    for mode in ("exact", "fast32"):
        sps = [api[0].SignalProcessor(2.4e6, mode=mode, device=d) for d in (0, 1)]
        for _ in range(2):
            for sp in sps:
                assert np.array_equal(sp.process(iq), want)
                assert torch.cuda.current_device() == before
    dec = api[1].TetraDecoder(auto_decrypt=False, device=1)
    assert len(dec.decode(want)) == len(oracle.decode(want)[0])


@pytest.mark.parametrize("mode,lanes", [("fast32", 2), ("exact", 3)])
def test_cuda_graph_replay_equals_direct_launch(api, mode, lanes):
    """ReceivePipeline.capture(): one captured step replayed as a CUDA graph gives the same rows, frames and counters
    as the direct launches, replay after replay, and follows new data written into the captured input tensor."""
    import torch
    from tetraear_b200 import pipeline, synth
    n_chunks, L = 6, 40000
    cap = synth.make_capture(L * n_chunks * 2, 130, seed=91, mix=(0.7, 0.15, 0.15)).reshape(2, n_chunks, L)
    pipe = pipeline.ReceivePipeline(2.4e6, L, max_chunks=n_chunks, mode=mode, n_lanes=lanes)
    buf = torch.from_numpy(cap[0].view(np.float32).reshape(n_chunks, L, 2)).cuda()
    g = pipe.capture(buf)
    for rep in range(3):
        src = cap[rep % 2]
        buf.copy_(torch.from_numpy(src.view(np.float32).reshape(n_chunks, L, 2)))      # refill the captured buffer
        pipe.stats.zero_()
        g.replay()
        torch.cuda.synchronize()
        ndib, dib = pipe.n_dibits.cpu().numpy(), pipe.dibits.cpu().numpy()
        nfr = pipe.n_frames.cpu().numpy()
        tot = [0, 0, 0]
        for c in range(n_chunks):
            od, _, _ = oracle.process(src[c], 2.4e6, 0.0)
            assert ndib[c] == len(od) and np.array_equal(dib[c, :ndib[c]], od), (rep, c)
            ofr, ost = oracle.decode(od)
            assert nfr[c] >= len(ofr)
            for k, key in enumerate(("total_bursts", "crc_pass", "crc_fail")):
                tot[k] += ost[key]
        assert list(pipe.stats.cpu().numpy()) == tot, rep


def test_pipeline_records_to_frame_dicts(api):
    """ReceivePipeline output -> TetraDecoder.frames_from_records == TetraDecoder.decode(dibits) chunk by chunk
    (the streaming form documented in INTEGRATION.md section C)."""
    from tetraear_b200 import pipeline, synth
    core = api[1]
    n_chunks, L = 4, 131072
    cap = synth.make_capture(L * n_chunks, 130, seed=17, mix=(0.7, 0.15, 0.15)).reshape(n_chunks, L)
    pipe = pipeline.ReceivePipeline(2.4e6, L, max_chunks=n_chunks, mode="fast32", n_lanes=2)
    nfr, frames, _ = pipe.run_host(cap)
    ndib, dib = pipe.n_dibits.cpu().numpy(), pipe.dibits.cpu().numpy()
    total = 0
    for c in range(n_chunks):
        a = core.TetraDecoder(auto_decrypt=False).frames_from_records(frames[c, :nfr[c]])
        b = core.TetraDecoder(auto_decrypt=False).decode(dib[c, :ndib[c]])
        assert len(a) == len(b) > 0
        for fa, fb in zip(a, b):
            assert set(fa) == set(fb)
            for k in fa:
                if isinstance(fa[k], np.ndarray):
                    assert np.array_equal(fa[k], fb[k]), k
                else:
                    assert fa[k] == fb[k], k
        total += len(a)
    assert total >= 8


@pytest.mark.parametrize("mode", ["exact", "fast", "fast32"])
def test_process_to_decode_hand_over_changes_nothing(api, mode):
    """process() runs the decoder's device part in its own launch sequence (tb200_process_decode_host) and decode() of
    the dibits it returned takes the records from there (_spec): frames, statistics, dibits and symbols must equal the
    two-call path, a modified copy of the dibits must NOT hit the hand-over, and the result equals the oracle's."""
    from tetraear_b200 import synth, _spec
    signal, core = api
    L = 131072
    cap = synth.make_capture(L * 3, 130, seed=23, mix=(0.6, 0.2, 0.2)).reshape(3, L)
    assert _spec.enabled

    def run(spec):
        _spec.enabled = spec
        try:
            sp, dec = signal.SignalProcessor(2.4e6, mode=mode), core.TetraDecoder(auto_decrypt=False)
            out = []
            for c in range(3):
                d = sp.process(cap[c])
                out.append((d, sp.symbols.copy(), dec.decode(d)))
            return out, dict(dec.protocol_parser.stats)
        finally:
            _spec.enabled = True

    a, sa = run(True)
    b, sb = run(False)
    assert sa == sb and sa["total_bursts"] >= 3
    n_frames = 0
    for (da, ya, fa), (db, yb, fb) in zip(a, b):
        assert np.array_equal(da, db) and np.array_equal(ya, yb) and len(fa) == len(fb)
        for x, y in zip(fa, fb):
            assert set(x) == set(y)
            for k in x:
                assert np.array_equal(x[k], y[k]) if isinstance(x[k], np.ndarray) else x[k] == y[k], k
        n_frames += len(fa)
    assert n_frames >= 6
    # oracle: same dibits -> same frame positions / CRC verdicts
    ofr, ostats = oracle.decode(a[0][0])
    assert [(f["number"], f["burst_crc"]) for f in a[0][2]] == [(f["number"], bool(f["burst_crc"])) for f in ofr]
    # a different array never hits the hand-over: flip one dibit and compare with a fresh two-call decode
    d2 = a[0][0].copy()
    d2[300] ^= 1
    assert _spec.lookup(d2) is None and _spec.lookup(a[2][0]) is not None
    f_spec = core.TetraDecoder(auto_decrypt=False).decode(d2)
    _spec.enabled = False
    try:
        f_plain = core.TetraDecoder(auto_decrypt=False).decode(d2)
    finally:
        _spec.enabled = True
    assert len(f_spec) == len(f_plain) and all(x["number"] == y["number"] and x.get("burst_crc") == y.get("burst_crc")
                                               for x, y in zip(f_spec, f_plain))


def test_decode_host_entry_point_equals_device_entry_point(api):
    """tb200_decode_batch_host (one call) vs tb200_decode_batch on device tensors, through decode_records."""
    import torch
    from tetraear_b200 import synth, _lib, _spec, _dev
    signal, core = api
    x = synth.make_capture(131072, 130, seed=31)
    _spec.enabled = False
    try:
        d = signal.SignalProcessor(2.4e6, mode="fast").process(x)
        dec = core.TetraDecoder(auto_decrypt=False)
        rec = dec.decode_records(d)                       # host entry point
    finally:
        _spec.enabled = True
    h = dec._handle(len(d))
    dd = _dev.to_dev(d, 0)
    cnt = _dev.to_dev(np.array([len(d)], np.int32), 0)
    mf = (2 * len(d)) // 250 + 2
    frames = _dev.zeros((mf * _lib.FRAME_DTYPE.itemsize,), torch.uint8, 0)
    nfr, stats = _dev.zeros((1,), torch.int32, 0), _dev.zeros((3,), torch.int64, 0)
    _lib.check(h.lib.tb200_decode_batch(h.h, dd.data_ptr(), 1, len(d), cnt.data_ptr(), frames.data_ptr(), mf, nfr.data_ptr(),
                                        stats.data_ptr(), _dev.stream_ptr(0)))
    k = int(nfr.cpu()[0])
    ref = frames.cpu().numpy().view(_lib.FRAME_DTYPE)[:k]
    assert k == len(rec) > 0 and rec.tobytes() == ref.tobytes()


def test_symbols_to_bits_out_of_range_values_follow_the_reference(api):
    """decoder.py:149-166: with max <= 3 every symbol goes through int(s) & 3 (-1 -> 3, -6 -> 2); with max > 3 the 8-PSK
    table applies and anything outside 0..7 (negative, 9, 300) maps to 0."""
    dec = api[1].TetraDecoder(auto_decrypt=False)
    for sym in (np.array([-1, 0, 3, -6, 2, 1, -4]), np.array([-1, 9, 5, 300, 7, 3, 0, -20, 4]), np.array([-3, -2, -1])):
        bits, mapped = dec.symbols_to_bits(sym)
        ob, om = oracle.symbols_to_bits(sym)
        assert np.array_equal(bits, ob) and np.array_equal(mapped, om), sym
    rng = np.random.default_rng(4)
    s = rng.integers(-4, 4, 1200)
    a = dec.decode(s)
    b = dec.decode((s & 3).astype(np.uint8))
    assert len(a) == len(b)


def test_pipeline_as_many_chunks_as_lanes(api):
    """max_chunks == n_lanes == n_chunks: every lane's handle holds ONE chunk, so the batch must be split across the
    lanes (round-1 advice: the single-lane shortcut overran handle 0); varying n_chunks between unjoined calls."""
    import torch
    from tetraear_b200 import pipeline, synth
    L = 20000
    cap = synth.make_capture(L * 4, 130, seed=5).reshape(4, L)
    pipe = pipeline.ReceivePipeline(2.4e6, L, max_chunks=4, mode="fast32", n_lanes=4)
    nfr, frames, stats = pipe.run_host(cap)
    ndib, dib = pipe.n_dibits.cpu().numpy(), pipe.dibits.cpu().numpy()
    for c in range(4):
        od, _, _ = oracle.process(cap[c], 2.4e6, 0.0)
        assert ndib[c] == len(od) and np.array_equal(dib[c, :ndib[c]], od)
    dev = torch.from_numpy(cap.view(np.float32).reshape(4, L, 2)).cuda()
    for n in (4, 3, 2, 4, 1, 4):
        pipe.run_device(dev[:n].contiguous(), join=False)
    pipe.join()
    torch.cuda.synchronize()
    assert np.array_equal(pipe.n_dibits.cpu().numpy()[:4], ndib[:4])
