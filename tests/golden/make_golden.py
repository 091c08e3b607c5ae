"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

Usage (from the repo root, in the container that has /root/reference):
    python tests/golden/make_golden.py

Inputs are regenerated from seeds by ``tetraear_b200.synth`` (committed), so only the
reference's OUTPUTS (and a checksum of each input) are stored.  The reference classes are
imported read-only through ``oracle/ref_bridge.py`` (bitstring stand-in, no bytecode).
The PSD/AFC block is inline code inside a QThread method of ``tetraear/ui/modern.py``
(PyQt6 is absent, the module cannot be imported), so its source LINES are read from the
file, dedented and exec'd against a stub ``self`` — the arithmetic executed is still the
reference's own text, not a restatement.
"""
import hashlib
import logging
import os
import sys
import textwrap
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
logging.disable(logging.CRITICAL)

from oracle import ref_bridge  # noqa: E402
from tetraear_b200 import synth  # noqa: E402
from tests.golden.cases import process_cases, psd_chunks, burst_batch, scanner_cases, voice_cases, afc_cases, peak_cases, spike_cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SP, TD, TP = ref_bridge.load()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name), **kw)
    print("wrote", name, len(kw), "arrays")


# ----------------------------------------------------------------------------- inputs
def gen_process():
    out = {}
    names = []
    for name, fs, f, iq in process_cases():
        sp = SP(fs)
        d = sp.process(iq, f)
        names.append(name)
        out[name + "/fs"] = np.float64(fs)
        out[name + "/f"] = np.float64(f)
        out[name + "/in_sha"] = np.array(sha(iq))
        out[name + "/dibits"] = np.asarray(d, dtype=np.uint8)
        out[name + "/symbols"] = np.asarray(sp.symbols)
    out["names"] = np.array(names)
    save("process.npz", **out)
    return out


def frame_summary(frames):
    """[sync-independent] per-frame tuple: number, type, encrypted(final), burst_crc, header enc bits."""
    rows = []
    for fr in frames:
        bits = np.asarray(fr["bits"])
        rows.append([fr["number"], fr["type"], int(bits[2]) * 2 + int(bits[3]),
                     int(bool(fr.get("burst_crc", False))), int("mac_pdu" in fr), int(bool(fr["encrypted"]))])
    return np.array(rows, dtype=np.int64).reshape(-1, 6)


def gen_decode(proc):
    out = {}
    names = []
    rng = np.random.default_rng(77)
    streams = {}
    for n in ("c1_chunk0", "c1_chunk1", "c1_afc", "noise", "wide_q83", "rate_1p8M", "ragged_40001"):
        streams[n] = proc[n + "/dibits"]
    # long clean stream with every burst kind, transmitted dibits directly (no modem)
    rb = np.random.default_rng(11)
    bits = np.concatenate([synth.make_burst(rb, k) for k in
                           ["pass", "near", "random", "pass", "pass", "random", "near", "pass"] * 3])
    for shift in (0, 1, 2, 3, 215, 216, 217):
        streams[f"tx_shift{shift}"] = synth.bits_to_dibits(
            np.concatenate([rb.integers(0, 2, shift, dtype=np.uint8), bits])[: (len(bits) + shift) // 2 * 2])
    # corrupted training sequences: exercises the 0.85 / 0.80 / adaptive cascade
    for nerr in (2, 3, 4, 5, 6):
        b = bits[: 510 * 4].copy()
        for k in range(4):
            idx = rb.choice(22, nerr, replace=False)
            b[510 * k + 216 + idx] ^= 1
        streams[f"ts_err{nerr}"] = synth.bits_to_dibits(b)
    for i in range(6):
        streams[f"random_{i}"] = rng.integers(0, 4, 1007 + i, dtype=np.uint8)
    streams["psk8"] = rng.integers(0, 8, 1500, dtype=np.uint8)
    streams["tiny_10"] = rng.integers(0, 4, 10, dtype=np.uint8)
    streams["tiny_254"] = synth.bits_to_dibits(bits[:508])
    streams["exact_255"] = synth.bits_to_dibits(bits[:510])
    streams["exact_256"] = synth.bits_to_dibits(bits[:512])
    for name, d in streams.items():
        dec = TD(auto_decrypt=False)
        b, mapped = dec.symbols_to_bits(d)
        frames = dec.decode(d)
        names.append(name)
        out[name + "/dibits"] = np.asarray(d, dtype=np.uint8)
        out[name + "/bits"] = np.asarray(b, dtype=np.uint8)
        out[name + "/mapped"] = np.asarray(mapped, dtype=np.uint8)
        for thr in (0.90, 0.85, 0.80, 0.7727):
            pos, mc = dec.find_sync(b, threshold=thr, return_max_corr=True)
            out[name + f"/sync_{thr}"] = np.asarray(pos, dtype=np.int64)
            out[name + f"/maxcorr_{thr}"] = np.float64(mc)
        out[name + "/frames"] = frame_summary(frames)
        out[name + "/frame_bits"] = (np.stack([np.asarray(f["bits"], dtype=np.uint8) for f in frames])
                                     if frames else np.zeros((0, 510), np.uint8))
        st = dec.protocol_parser.stats
        out[name + "/stats"] = np.array([st["total_bursts"], st["crc_pass"], st["crc_fail"]], dtype=np.int64)
    out["names"] = np.array(names)
    save("decode.npz", **out)


def gen_bursts():
    d = burst_batch()
    p = TP()
    bt, ok, nd, mac = [], [], [], []
    data = np.zeros((len(d), 510), np.uint8)
    ts = np.zeros((len(d), 22), np.uint8)
    for i, row in enumerate(d):
        b = p.parse_burst(row, slot_number=i % 4)
        bt.append(b.burst_type.value)
        ok.append(int(bool(b.crc_ok)))
        nd.append(len(b.data_bits))
        data[i, :len(b.data_bits)] = b.data_bits
        ts[i, :len(b.training_sequence)] = b.training_sequence
        mac.append(int(TP().parse_mac_pdu(b.data_bits) is not None))
    st = p.stats
    crcs = np.array([int("".join(str(int(x)) for x in p._calculate_crc16(row_bits)), 2)
                     for row_bits in data[:64, :200]], dtype=np.int64)
    save("bursts.npz", in_sha=np.array(sha(d)), burst_type=np.array(bt, np.uint8), crc_ok=np.array(ok, np.uint8),
         n_data=np.array(nd, np.int32), data_bits=np.packbits(data, axis=1), training=ts,
         mac_ok=np.array(mac, np.uint8), crc16_first64=crcs,
         stats=np.array([st["total_bursts"], st["crc_pass"], st["crc_fail"]], np.int64))


# ------------------------------------------------------------------------------- PSD
def _lines(path, lo, hi):
    with open(path) as f:
        src = f.readlines()[lo - 1:hi]
    return textwrap.dedent("".join(src))


def gen_psd():
    modern = os.path.join(ref_bridge.REFERENCE_ROOT, "tetraear", "ui", "modern.py")
    spec_src = _lines(modern, 1928, 1944)          # n_fft .. freqs_actual
    stat_src = _lines(modern, 1953, 2010)          # signal_present .. is_signal_strong
    ema_src = _lines(modern, 929, 937)             # denoiser EMA
    assert "np.hanning" in spec_src and "is_signal_strong" in stat_src and "avg_factor" in ema_src
    fs = 2.4e6
    chunks = psd_chunks()
    powers, stats = [], []
    widget = types.SimpleNamespace(denoiser_enabled=True, avg_factor=0.15, smoothed_fft=None, current_fft=None)
    smoothed = []
    for samples in chunks:
        ns = {"np": np, "samples": samples, "current_time": 0.0,
              "self": types.SimpleNamespace(sample_rate=fs, frequency=390e6, last_signal_time=0.0)}
        exec(spec_src, ns)
        exec(stat_src, ns)
        powers.append(ns["power"])
        stats.append([ns["signal_power"], ns["peak_power"], ns["peak_idx"], ns["peak_freq_offset"],
                      ns["noise_floor"], ns["snr"], float(ns["signal_present"])])
        ens = {"np": np, "self": widget, "powers": ns["power"]}
        exec(ema_src, ens)
        smoothed.append(np.array(widget.current_fft))
    save("psd.npz", in_sha=np.array(sha(np.concatenate(chunks))), power=np.stack(powers), stats=np.array(stats),
         smoothed=np.stack(smoothed), freqs=ns["freqs"])


# --------------------------------------------------------------- scanner heuristics (N3)
def gen_scanner():
    from tetraear.signal.scanner import TetraSignalDetector
    out, names = {}, []
    keys = ["power_db", "is_tetra", "confidence", "modulation_confidence", "sync_detected", "sync_correlation",
            "frames_validated", "crc_pass_rate", "power_stable", "signal_present"]
    for name, fs, x in scanner_cases():
        det = TetraSignalDetector(sample_rate=fs)
        r = det.analyze_signal(x)
        names.append(name)
        out[name + "/in_sha"] = np.array(sha(x))
        out[name + "/fs"] = np.float64(fs)
        out[name + "/analysis"] = np.array([float(r[k]) for k in keys], dtype=np.float64)
        out[name + "/power_dtype"] = np.array(type(r["power_db"]).__name__)
        out[name + "/power"] = np.float64(det.calculate_power(x))
        out[name + "/mod"] = np.array(det.detect_tetra_modulation(x), dtype=np.float64)
        out[name + "/sync"] = np.array(det.detect_sync_pattern(x), dtype=np.float64)
        out[name + "/stable"] = np.array([det.check_power_stability(x), det.check_power_stability(x, 3)], dtype=np.float64)
    out["names"] = np.array(names)
    out["keys"] = np.array(keys)
    save("scanner.npz", **out)


# --------------------------------------------------------------- codec-frame builder (N4)
def gen_voice():
    """_extract_voice_slot_from_symbols is a method of a PyQt6 class (module not importable here): its source
    LINES are read from the file, dedented and exec'd as a plain function — the reference's own text."""
    modern = os.path.join(ref_bridge.REFERENCE_ROOT, "tetraear", "ui", "modern.py")
    src = _lines(modern, 2309, 2423)
    assert src.startswith("def _extract_voice_slot_from_symbols") and "0x6B21" in src
    ns = {"logger": logging.getLogger("golden")}
    exec(src, ns)
    fn = ns["_extract_voice_slot_from_symbols"]
    dibits, pos, _ = voice_cases()
    blocks = np.zeros((len(pos), 690), np.int16)
    valid = np.zeros(len(pos), np.uint8)
    for i, p in enumerate(pos):
        b = fn(None, {"position": p}, dibits, 13)
        if b is not None:
            blocks[i] = np.frombuffer(b, dtype="<i2")
            valid[i] = 1
    none_pos = fn(None, {}, dibits, 13)
    assert none_pos is None
    # the inline bit-row builder of continuous_capture.py:103-124, exec'd from its source lines per frame
    cc = os.path.join(ref_bridge.REFERENCE_ROOT, "continuous_capture.py")
    bsrc = _lines(cc, 103, 124)
    assert bsrc.startswith("soft_bits = [127 if int(b) else -127") and bsrc.rstrip().endswith("dtype=np.int16).tobytes()")
    _, _, bit_rows = voice_cases()
    bit_blocks = np.zeros((len(bit_rows), 690), np.int16)
    bit_valid = np.zeros(len(bit_rows), np.uint8)
    for i, bits in enumerate(bit_rows):
        if bits is not None and len(bits) >= 432:              # the guard of continuous_capture.py:102
            ns = {"np": np, "bits": bits}
            exec(bsrc, ns)
            bit_blocks[i] = np.frombuffer(ns["codec_input"], dtype="<i2")
            bit_valid[i] = 1
    save("voice.npz", in_sha=np.array(sha(dibits)), blocks=blocks, valid=valid, bit_blocks=bit_blocks, bit_valid=bit_valid)


def gen_afc():
    """Auto-follow rule: lines 5227-5251 of ui/modern.py exec'd from the source text."""
    modern = os.path.join(ref_bridge.REFERENCE_ROOT, "tetraear", "ui", "modern.py")
    src = _lines(modern, 5227, 5251)
    assert src.startswith("center_idx = np.abs(freqs - freq_mhz).argmin()") and "new_freq" in src
    out = []
    for freqs, powers, tuned in afc_cases():
        ns = {"np": np, "freqs": freqs, "powers": powers, "freq_mhz": tuned}
        exec(src, ns)
        out.append(ns.get("new_freq", tuned))
    save("afc.npz", new_freq=np.array(out, dtype=np.float64))


def gen_peak():
    """WaterfallWidget.find_peak_near: lines 727-772 of ui/modern.py exec'd as a plain function against a stub self."""
    modern = os.path.join(ref_bridge.REFERENCE_ROOT, "tetraear", "ui", "modern.py")
    src = _lines(modern, 727, 772)
    assert src.startswith("def find_peak_near(self, freq, search_range_hz):") and "Sub-bin quadratic" in src
    ns = {"np": np}
    exec(src, ns)
    out = []
    for freqs, powers, click, rng_hz in peak_cases():
        stub = types.SimpleNamespace(current_fft=powers, current_freqs=freqs)
        out.append(ns["find_peak_near"](stub, click, rng_hz))
    save("peak.npz", peak=np.array(out, dtype=np.float64))


def gen_spike():
    """WaterfallWidget.find_spike_band_near: lines 774-857 of ui/modern.py exec'd as a plain function against a stub self."""
    modern = os.path.join(ref_bridge.REFERENCE_ROOT, "tetraear", "ui", "modern.py")
    src = _lines(modern, 774, 857)
    assert src.startswith("def find_spike_band_near(self, freq_mhz: float, search_range_hz: float)") and "bandwidth_hz" in src
    ns = {"np": np}
    exec(src, ns)
    out = []
    for freqs, powers, click, rng_hz, noise in spike_cases():
        stub = types.SimpleNamespace(current_fft=powers, current_freqs=freqs, noise_floor=noise, bandwidth=25000.0)
        r = ns["find_spike_band_near"](stub, click, rng_hz)
        out.append([r["center_mhz"], r["left_mhz"], r["right_mhz"], r["bandwidth_hz"]])
    save("spike.npz", spike=np.array(out, dtype=np.float64))


if __name__ == "__main__":
    proc = gen_process()
    gen_decode(proc)
    gen_bursts()
    gen_psd()
    gen_scanner()
    gen_voice()
    gen_afc()
    gen_peak()
    gen_spike()
