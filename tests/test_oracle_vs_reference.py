"""Live cross-check of the oracle against the UNMODIFIED reference on fresh seeds (beyond the committed
golden vectors).  Runs only where /root/reference exists (the build container); skipped on the GPU box."""
import logging

import numpy as np
import pytest

from oracle import oracle, ref_bridge
from tetraear_b200 import synth

pytestmark = pytest.mark.skipif(not ref_bridge.available(), reason="reference tree not present")
# tests that read ui/modern.py / signal/scanner.py need the full tree (build container); oracle/_ref stages only the path's modules
needs_full = pytest.mark.skipif(not ref_bridge.FULL_TREE, reason="full reference tree not present")


@pytest.fixture(scope="module")
def ref():
    logging.disable(logging.CRITICAL)
    yield ref_bridge.load()
    logging.disable(logging.NOTSET)


def test_process_and_decode_live(ref):
    SP, TD, TP = ref
    rng = np.random.default_rng(20260101)
    for case in range(6):
        fs, sps_total = [(2.4e6, 130), (1.8e6, 98), (2.4e6, 130)][case % 3]
        n = int(rng.integers(20000, 140000))
        f = float(rng.choice([0.0, rng.uniform(-2000, 2000)]))
        iq = synth.make_capture(n, sps_total, seed=900 + case, sigma=0.04, mix=(0.6, 0.2, 0.2))
        if f:
            iq = (iq * np.exp(2j * np.pi * f * np.arange(n) / fs)).astype(np.complex64)
        sp = SP(fs)
        d_ref = sp.process(iq, f)
        d, sym, _ = oracle.process(iq, fs, f)
        assert np.array_equal(d, d_ref), case
        assert np.abs(sym - sp.symbols).max() <= 1e-12 * max(1.0, np.abs(sp.symbols).max()), case
        dec = TD(auto_decrypt=False)
        fr_ref = dec.decode(d_ref)
        fr, st = oracle.decode(d)
        assert [(f_["number"], f_["type"], bool(f_.get("burst_crc"))) for f_ in fr_ref] == \
               [(f_["number"], f_["type"], f_["burst_crc"]) for f_ in fr], case
        rs = dec.protocol_parser.stats
        assert (rs["total_bursts"], rs["crc_pass"], rs["crc_fail"]) == (st["total_bursts"], st["crc_pass"], st["crc_fail"])


def test_parse_burst_and_mac_validity_live(ref):
    SP, TD, TP = ref
    rows = synth.make_burst_batch(400, seed=31337)
    o = oracle.parse_bursts(rows)
    for i, row in enumerate(rows):
        b = TP().parse_burst(row)
        assert b.burst_type.value == o["burst_type"][i] and bool(b.crc_ok) == bool(o["crc_ok"][i])
        nd = int(o["n_data"][i])
        assert np.array_equal(np.asarray(b.data_bits), o["data_bits"][i, :nd])
        assert (TP().parse_mac_pdu(b.data_bits) is not None) == oracle.mac_pdu_valid(o["data_bits"][i, :nd])


def test_host_mac_header_walk_matches_reference(ref):
    """The host-side parse_mac_pdu kept in tetraear_b200.core (frame drop rule, fragment state) against the
    reference's, PDU by PDU in sequence (fragment buffer state carried)."""
    SP, TD, TP = ref
    from tetraear_b200.core.protocol import TetraProtocolParser as Mine
    rng = np.random.default_rng(5)
    r, m = TP(), Mine.__new__(Mine)
    Mine.__init__(m)
    for i in range(600):
        n = int(rng.choice([216, 510, 40, 7, 12]))
        bits = rng.integers(0, 2, n)
        if rng.random() < 0.5:
            bits[0:2] = rng.choice([[0, 0], [0, 1], [1, 0], [1, 1]])
            if rng.random() < 0.5 and n > 40:
                bits[29:35] = [0, 0, 0, 1, 0, 1]
        a, b = r.parse_mac_pdu(bits), m.parse_mac_pdu(bits)
        assert (a is None) == (b is None), i
        if a is not None:
            assert (a.pdu_type.name, a.encrypted, a.address, a.length, a.data, a.fill_bits, a.encryption_mode,
                    a.reassembled_data) == (b.pdu_type.name, b.encrypted, b.address, b.length, b.data, b.fill_bits,
                                            b.encryption_mode, b.reassembled_data), i
    assert {k: r.stats[k] for k in ("clear_mode_frames", "encrypted_frames")} == \
           {k: m.stats[k] for k in ("clear_mode_frames", "encrypted_frames")}


@needs_full
def test_scanner_heuristics_live(ref):
    """oracle/scanner.py against the live TetraSignalDetector on fresh seeds (scanner.py:42-289)."""
    from oracle import scanner as osc
    from tetraear.signal.scanner import TetraSignalDetector
    rng = np.random.default_rng(4242)
    for case in range(5):
        n = int(rng.integers(12000, 50000))
        if case % 2:
            x = (rng.uniform(0.01, 0.5) * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
        else:
            x = synth.make_capture(n, 130, seed=300 + case, sigma=0.05)
        if case == 3:
            x = x.astype(np.complex128)
        want = TetraSignalDetector(2.4e6).analyze_signal(x)
        got = osc.analyze_signal(x, 2.4e6)
        for k in want:
            assert got[k] == want[k], (case, k, got[k], want[k])


@needs_full
def test_afc_rule_live(ref):
    """oracle.afc_step and the shipped host helper against the reference's own lines (ui/modern.py:5227-5251,
    exec'd from the source text: the module needs PyQt6 and cannot be imported)."""
    import os, textwrap
    from tetraear_b200.signal import afc_step
    path = os.path.join(ref_bridge.REFERENCE_ROOT, "tetraear", "ui", "modern.py")
    with open(path) as f:
        src = textwrap.dedent("".join(f.readlines()[5226:5251]))
    assert src.startswith("center_idx = np.abs(freqs - freq_mhz).argmin()") and "new_freq" in src
    rng = np.random.default_rng(12)
    fired = 0
    for case in range(40):
        fc = 390.0 + rng.uniform(-1, 1)
        freqs = fc + (np.arange(2048) - 1024) * (2.4 / 2048)
        powers = rng.uniform(-110, -70, 2048)
        if case % 4:
            powers[1024 + int(rng.integers(-12, 12))] = rng.uniform(-65, -20)
        tuned = fc + rng.uniform(-0.004, 0.004)
        ns = {"np": np, "freqs": freqs, "powers": powers, "freq_mhz": tuned}
        exec(src, ns)
        want = ns.get("new_freq", tuned)
        fired += "new_freq" in ns
        assert oracle.afc_step(freqs, powers, tuned) == want, case
        assert afc_step(freqs, powers, tuned) == want, case
    assert 5 < fired < 40


@needs_full
def test_find_peak_near_live(ref):
    """oracle.find_peak_near and the shipped helper against ui/modern.py:727-772 exec'd from the source text."""
    import os, textwrap, types
    from tests.golden.cases import peak_cases
    from tetraear_b200.signal import find_peak_near
    path = os.path.join(ref_bridge.REFERENCE_ROOT, "tetraear", "ui", "modern.py")
    with open(path) as f:
        src = textwrap.dedent("".join(f.readlines()[726:772]))
    assert src.startswith("def find_peak_near(self, freq, search_range_hz):")
    ns = {"np": np}
    exec(src, ns)
    for freqs, powers, click, rng_hz in peak_cases(n=90, seed=4):
        want = ns["find_peak_near"](types.SimpleNamespace(current_fft=powers, current_freqs=freqs), click, rng_hz)
        assert oracle.find_peak_near(freqs, powers, click, rng_hz) == want
        assert find_peak_near(freqs, powers, click, rng_hz) == want


@needs_full
def test_find_spike_band_near_live(ref):
    """The shipped helper against ui/modern.py:774-857 exec'd from the source text, on fresh cases."""
    import os, textwrap, types
    from tests.golden.cases import spike_cases
    from tetraear_b200.signal import find_spike_band_near
    path = os.path.join(ref_bridge.REFERENCE_ROOT, "tetraear", "ui", "modern.py")
    with open(path) as f:
        src = textwrap.dedent("".join(f.readlines()[773:857]))
    assert src.startswith("def find_spike_band_near(self, freq_mhz: float, search_range_hz: float)")
    ns = {"np": np}
    exec(src, ns)
    for freqs, powers, click, rng_hz, noise in spike_cases(n=72, seed=9):
        stub = types.SimpleNamespace(current_fft=powers, current_freqs=freqs, noise_floor=noise, bandwidth=20000.0)
        assert find_spike_band_near(freqs, powers, click, rng_hz, noise_floor=noise, bandwidth=20000.0) == \
               ns["find_spike_band_near"](stub, click, rng_hz)


def _oracle_stage(method, kw, x):
    """The oracle's counterpart of a SignalProcessor stage call (reference defaults: sample_rate 2.4e6)."""
    fs = kw.get("sample_rate", 2.4e6)
    if method == "filter_signal":
        return oracle.filter_signal(x, kw["bandwidth"], fs)
    if method == "frequency_shift":
        return oracle.frequency_shift(x, kw["freq_offset"], fs)
    if method == "extract_symbols":
        return oracle.extract_symbols(x, fs)[0]
    return oracle.demodulate_dqpsk(x)


def test_stage_edge_cases_live(ref):
    """Pins the oracle on the awkward stage-level inputs the GPU test (tests/test_gpu_dropin_api.py) relies on."""
    from tests.golden.cases import stage_edge_cases
    SP = ref[0]
    for i, (method, kw, x) in enumerate(stage_edge_cases()):
        want = getattr(SP(2.4e6), method)(x, **kw)
        got = _oracle_stage(method, kw, x)
        assert len(got) == len(want), (i, method, kw)
        if method == "demodulate_dqpsk":
            assert np.array_equal(got, want), (i, method)
        elif len(want):
            assert np.abs(np.asarray(got) - np.asarray(want)).max() <= 1e-12 * max(1.0, np.abs(want).max()), (i, method, kw)


def test_process_edge_cases_live(ref):
    """Pins oracle.process on the unusual rates / lengths / dtypes of tests/golden/cases.py::process_edge_cases."""
    from tests.golden.cases import process_edge_cases
    SP = ref[0]
    for i, (fs, f, iq) in enumerate(process_edge_cases()):
        sp = SP(fs)
        want = sp.process(iq, f)
        got, sym, _ = oracle.process(iq, fs, f)
        assert np.array_equal(got, want), (i, fs, f, len(iq))
        rs = np.asarray(sp.symbols)
        assert len(sym) == len(rs), (i, fs)
        if len(rs):
            assert np.abs(sym - rs).max() <= 1e-12 * max(1.0, np.abs(rs).max()), (i, fs)


@needs_full
def test_band_statistics_other_rates_live(ref):
    """oracle.band_stats against the reference's inline block (ui/modern.py:1953-2010, exec'd from the source text)
    at sample rates where the 25 kHz band is wider than the FFT or leaves no noise bins."""
    import os, textwrap, types
    path = os.path.join(ref_bridge.REFERENCE_ROOT, "tetraear", "ui", "modern.py")
    with open(path) as f:
        lines = f.readlines()
    spec_src = textwrap.dedent("".join(lines[1927:1944]))
    stat_src = textwrap.dedent("".join(lines[1952:2010]))
    assert "np.hanning" in spec_src and "is_signal_strong" in stat_src
    rng = np.random.default_rng(5)
    for fs in (2.4e6, 1.8e6, 240000.0, 50000.0, 30000.0, 20e6):
        x = (0.05 * (rng.standard_normal(2048) + 1j * rng.standard_normal(2048))).astype(np.complex64)
        x += (0.4 * np.exp(2j * np.pi * (fs / 2048 * 3.3) * np.arange(2048) / fs)).astype(np.complex64)
        ns = {"np": np, "samples": x, "current_time": 0.0,
              "self": types.SimpleNamespace(sample_rate=fs, frequency=390e6, last_signal_time=0.0)}
        exec(spec_src, ns)
        exec(stat_src, ns)
        _, p = oracle.psd_frame(x, fs, 2048)
        assert np.array_equal(p, ns["power"])
        o = oracle.band_stats(p, fs)
        assert o["peak_idx"] == ns["peak_idx"] and o["signal_present"] == bool(ns["signal_present"]), fs
        for k, ref_k in (("signal_power", "signal_power"), ("peak_power", "peak_power"), ("noise_floor", "noise_floor"),
                         ("snr", "snr"), ("peak_freq_offset", "peak_freq_offset")):
            assert o[k] == float(ns[ref_k]), (fs, k)
