"""Pin the oracle (oracle/) against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle
from tests.golden.cases import sha, psd_chunks, burst_batch


def test_inputs_reproduce(golden_process, process_inputs):
    for name in golden_process["names"]:
        fs, f, iq = process_inputs[str(name)]
        assert sha(iq) == str(golden_process[f"{name}/in_sha"]), name


def test_process_matches_reference(golden_process, process_inputs):
    for name in golden_process["names"]:
        name = str(name)
        fs, f, iq = process_inputs[name]
        d, sym, _ = oracle.process(iq, fs, f)
        gd, gs = golden_process[name + "/dibits"], golden_process[name + "/symbols"]
        assert d.dtype == np.uint8 and np.array_equal(d, gd), name
        assert len(sym) == len(gs), name
        if len(gs):
            scale = max(np.abs(gs).max(), 1e-30)
            assert np.abs(np.asarray(sym) - gs).max() <= 1e-12 * scale, name


def test_bits_sync_decode_match_reference(golden_decode):
    g = golden_decode
    for name in g["names"]:
        name = str(name)
        d = g[name + "/dibits"]
        bits, mapped = oracle.symbols_to_bits(d)
        assert np.array_equal(bits, g[name + "/bits"]), name
        assert np.array_equal(mapped, g[name + "/mapped"]), name
        for thr in (0.90, 0.85, 0.80, 0.7727):
            pos, mc = oracle.find_sync(bits, thr)
            assert pos == list(g[name + f"/sync_{thr}"]), (name, thr)
            assert mc == float(g[name + f"/maxcorr_{thr}"]), (name, thr)
        frames, stats = oracle.decode(d)
        gf = g[name + "/frames"]
        assert len(frames) == len(gf), name
        for fr, row, fb in zip(frames, gf, g[name + "/frame_bits"]):
            assert [fr["number"], fr["type"], fr["enc_mode"], int(fr["burst_crc"]), int(fr["mac_ok"])] == list(row[:5]), name
            assert np.array_equal(fr["bits"], fb), name
        assert [stats["total_bursts"], stats["crc_pass"], stats["crc_fail"]] == list(g[name + "/stats"]), name


def test_parse_bursts_match_reference(golden_bursts):
    g = golden_bursts
    d = burst_batch()
    assert sha(d) == str(g["in_sha"])
    r = oracle.parse_bursts(d)
    assert np.array_equal(r["burst_type"], g["burst_type"])
    assert np.array_equal(r["crc_ok"], g["crc_ok"])
    assert np.array_equal(r["n_data"], g["n_data"])
    gbits = np.unpackbits(g["data_bits"], axis=1)[:, :510]
    assert np.array_equal(r["data_bits"], gbits)
    mac = np.array([oracle.mac_pdu_valid(r["data_bits"][i, :r["n_data"][i]]) for i in range(len(d))], np.uint8)
    assert np.array_equal(mac, g["mac_ok"])
    crcs = [oracle.crc16(gbits[i, :200]) for i in range(64)]
    assert crcs == list(g["crc16_first64"])
    assert [len(d), int(r["crc_ok"].sum()), int(len(d) - r["crc_ok"].sum())] == list(g["stats"])
    # the batch must exercise every verdict
    assert 0 < r["crc_ok"].sum() < len(d) and (r["burst_type"] == 5).any() and (g["mac_ok"] == 0).any()


def test_psd_matches_reference(golden_psd):
    g = golden_psd
    chunks = psd_chunks()
    assert sha(np.concatenate(chunks)) == str(g["in_sha"])
    powers = []
    for i, c in enumerate(chunks):
        freqs, p = oracle.psd_frame(c, 2.4e6, 2048)
        powers.append(p)
        assert np.array_equal(freqs, g["freqs"])
        assert np.abs(p - g["power"][i]).max() <= 1e-9
        st = oracle.band_stats(p, 2.4e6)
        row = g["stats"][i]
        got = [st["signal_power"], st["peak_power"], st["peak_idx"], st["peak_freq_offset"], st["noise_floor"],
               st["snr"], float(st["signal_present"])]
        assert np.allclose(got, row, rtol=0, atol=1e-9), i
    sm = oracle.ema_denoise(np.stack(powers))
    assert np.abs(sm - g["smoothed"]).max() <= 1e-9
    assert g["stats"][:, 6].any() and not g["stats"][:, 6].all()  # both gate outcomes covered


def test_known_answers_from_reference_tests():
    """The few known-answer facts the reference's own tests hold."""
    # tests/unit/test_tetra_protocol.py:53-65: continuous-downlink pattern is a sync, zeros are not
    row = np.zeros(255, np.uint8)
    assert oracle.parse_bursts(row)["burst_type"][0] == 2
    bits = np.zeros(510, np.uint8)
    bits[255:277] = [1, 1, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0, 1, 0, 0]
    assert oracle.parse_bursts((bits[0::2] << 1) | bits[1::2])["burst_type"][0] == 5
    # tests/unit/test_tetra_decoder.py:49-54: < 22 bits -> no sync
    assert oracle.find_sync(np.zeros(10, np.uint8)) == ([], 0.0)
    # tests/unit/test_signal_processor.py:36-41, 72-86, 97-102: empty / single sample
    assert len(oracle.process(np.array([], np.complex64), 2.4e6)[0]) == 0
    assert len(oracle.demodulate_dqpsk(np.array([1 + 1j]))) == 0
    assert oracle.FRAME_LENGTH == 510


# ---------------------------------------------------------------- SURVEY.md 8(f) N3 / N4
def test_scanner_heuristics_match_reference():
    import os
    from oracle import scanner as osc
    from tests.conftest import GOLDEN
    from tests.golden.cases import scanner_cases
    g = np.load(os.path.join(GOLDEN, "scanner.npz"))
    keys = [str(k) for k in g["keys"]]
    for name, fs, x in scanner_cases():
        assert sha(x) == str(g[name + "/in_sha"]), name
        r = osc.analyze_signal(x, fs)
        want = g[name + "/analysis"]
        for k, w in zip(keys, want):
            assert float(r[k]) == w, (name, k, r[k], w)
        assert type(r["power_db"]).__name__ == str(g[name + "/power_dtype"]), name
        assert float(osc.calculate_power(x)) == float(g[name + "/power"]), name
        assert [float(v) for v in osc.detect_tetra_modulation(x)] == list(g[name + "/mod"]), name
        assert [float(v) for v in osc.detect_sync_pattern(x, fs)] == list(g[name + "/sync"]), name
        assert [float(osc.check_power_stability(x)), float(osc.check_power_stability(x, 3))] == list(g[name + "/stable"]), name


def test_voice_blocks_match_reference():
    import os
    from oracle import voice as ov
    from tests.conftest import GOLDEN
    from tests.golden.cases import voice_cases
    g = np.load(os.path.join(GOLDEN, "voice.npz"))
    dibits, pos, bit_rows = voice_cases()
    assert sha(dibits) == str(g["in_sha"])
    for i, p in enumerate(pos):
        b = ov.from_symbols(p, dibits)
        assert (b is not None) == bool(g["valid"][i]), p
        if b is not None:
            assert np.array_equal(b, g["blocks"][i]), p
    # the two inline builders have no callable counterpart in the reference: internal consistency only
    for r in bit_rows:
        b = ov.from_bits(r)
        assert (b is None) == (len(r) < 432)
        if b is not None:
            assert b[0] == 0x6B21 and np.count_nonzero(b) == 433 and set(np.unique(b[1:115])) <= {-127, 127}
            p = ov.from_params(r)
            assert p[0] == 0x6B21 and -500 <= p[1] <= 500 and 20 <= p[2] <= 220 and 0 <= p[3] <= 150


def test_find_peak_near_matches_reference():
    """WaterfallWidget.find_peak_near (ui/modern.py:727-772): golden values from the reference's own source lines."""
    import os
    from tests.conftest import GOLDEN
    from tests.golden.cases import peak_cases
    from tetraear_b200.signal import find_peak_near
    want = np.load(os.path.join(GOLDEN, "peak.npz"))["peak"]
    cases = peak_cases()
    assert len(want) == len(cases)
    interpolated = 0
    for (freqs, powers, click, rng_hz), w in zip(cases, want):
        assert oracle.find_peak_near(freqs, powers, click, rng_hz) == w
        assert find_peak_near(freqs, powers, click, rng_hz) == w      # host helper (scalar GUI logic, no device work)
        interpolated += not np.any(np.asarray(freqs) == w) and w != click
    assert interpolated >= 10


def test_afc_rule_matches_reference():
    import os
    from tests.conftest import GOLDEN
    from tests.golden.cases import afc_cases
    from tetraear_b200.signal import afc_step
    want = np.load(os.path.join(GOLDEN, "afc.npz"))["new_freq"]
    moved = 0
    for (freqs, powers, tuned), w in zip(afc_cases(), want):
        assert oracle.afc_step(freqs, powers, tuned) == w
        assert afc_step(freqs, powers, tuned) == w          # host helper (scalar GUI logic, no device work)
        moved += w != tuned
    assert 5 < moved < len(want)


def test_find_spike_band_near_matches_reference():
    """WaterfallWidget.find_spike_band_near (ui/modern.py:774-857): centre, edges and clamped bandwidth equal to the golden
    values produced by the reference's own source lines (tests/golden/make_golden.py::gen_spike)."""
    import os
    from tests.conftest import GOLDEN
    from tests.golden.cases import spike_cases
    from tetraear_b200.signal import find_spike_band_near
    want = np.load(os.path.join(GOLDEN, "spike.npz"))["spike"]
    cases = spike_cases()
    assert len(want) == len(cases)
    for (freqs, powers, click, rng_hz, noise), w in zip(cases, want):
        d = find_spike_band_near(freqs, powers, click, rng_hz, noise_floor=noise, bandwidth=25000.0)
        assert [d["center_mhz"], d["left_mhz"], d["right_mhz"], d["bandwidth_hz"]] == list(w)
    assert len(set(want[:, 3])) >= 6 and (want[:, 3] == 1000.0).any() and (want[:, 3] == 25000.0).any()
    none = find_spike_band_near(None, None, 390.0, 1000.0, bandwidth=12500.0)
    assert none == {"center_mhz": 390.0, "left_mhz": 390.0, "right_mhz": 390.0, "bandwidth_hz": 12500.0}
