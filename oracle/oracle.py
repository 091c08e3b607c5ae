"""CPU oracle for the IQ->bursts path of TetraEar (TEST INFRASTRUCTURE ONLY).

A restatement — numpy for the vector arithmetic, plain C (oracle/tetra_oracle.c) for the
sequential loops — of what the reference computes on this path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg
may import it, and only as the checker / the CPU arm.  The product package never does.

Pinning: the reference's own tests hold no golden vectors for this path, so
the oracle is pinned against the reference itself, executed in the build container on
seeded synthetic IQ: ``tests/golden/make_golden.py`` (committed) ran the UNMODIFIED
``tetraear.signal.processor.SignalProcessor`` / ``tetraear.core.decoder.TetraDecoder`` /
``tetraear.core.protocol.TetraProtocolParser`` from /root/reference and stored their
outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
here against those files (bit-exact dibits / bits / sync positions / CRC flags / stats,
bit-exact float32 decimator output, <=1e-12 relative on the float64 stages), and
``tests/test_oracle_vs_reference.py`` re-runs the live comparison whenever
/root/reference is present.  The float half of the path lives in scipy 1.18.1 / numpy
2.3.5 (un-vendored, unpinned by the reference): restated in ``oracle/design.py`` and the C
file, and pinned against scipy in ``tests/test_oracle_design.py``.

Reference citations are relative to /root/reference/.
"""
import ctypes
import math

import numpy as np

from . import build as _build
from . import design

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_build.build())
        c = ctypes
        vp = c.c_void_p
        L.orc_decimate_lane_f32.argtypes = [vp, c.c_int, vp, vp, c.c_long, c.c_long, c.c_int, vp, c.c_long]
        L.orc_decimate_lane_f32.restype = c.c_int
        L.orc_decimate_lane_f64.argtypes = L.orc_decimate_lane_f32.argtypes
        L.orc_decimate_lane_f64.restype = c.c_int
        L.orc_filtfilt_lane_f64.argtypes = [vp, vp, c.c_int, vp, vp, c.c_long, c.c_long, c.c_int, vp, c.c_long]
        L.orc_filtfilt_lane_f64.restype = c.c_int
        L.orc_slicer.argtypes = [vp, c.c_long, vp]
        L.orc_slicer.restype = None
        L.orc_find_sync.argtypes = [vp, c.c_long, c.c_double, vp, vp]
        L.orc_find_sync.restype = c.c_long
        L.orc_crc16.argtypes = [vp, c.c_long, c.c_long]
        L.orc_crc16.restype = c.c_uint16
        L.orc_parse_bursts.argtypes = [vp, c.c_long, vp, vp, vp, vp]
        L.orc_parse_bursts.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- DSP stages
def decimation_factor(sample_rate):
    """tetraear/signal/processor.py:245-250: q, or 1 when no decimation happens."""
    target = 240000
    if sample_rate > target * 2:
        q = int(sample_rate / target)
        if q > 1:
            return q
    return 1


def decimate(samples, q):
    """scipy.signal.decimate(samples, q) as called at processor.py:254.

    complex64 in -> complex64 out through the float32 cascade; complex128 in -> complex128.
    Raises ValueError when len(samples) <= 27 like scipy (processor.py:256 catches it).
    """
    x = np.ascontiguousarray(samples)
    if x.dtype == np.complex64:
        sos, zi = design.decimator_design(q)
        real_t, fn = np.float32, lib().orc_decimate_lane_f32
    else:
        x = x.astype(np.complex128)
        sos = design.cheby1_sos(q)
        zi = design.sosfilt_zi(sos)
        real_t, fn = np.float64, lib().orc_decimate_lane_f64
    n = len(x)
    n_out = (n + q - 1) // q
    out = np.empty(n_out, dtype=x.dtype)
    xv = x.view(real_t)
    ov = out.view(real_t)
    sos = np.ascontiguousarray(sos, dtype=real_t)
    zi = np.ascontiguousarray(zi, dtype=real_t)
    item = xv.itemsize
    for comp in (0, 1):
        rc = fn(_p(sos), sos.shape[0], _p(zi), ctypes.c_void_p(xv.ctypes.data + comp * item), n, 2, q,
                ctypes.c_void_p(ov.ctypes.data + comp * item), 2)
        if rc != 0:
            raise ValueError("The length of the input vector x must be greater than padlen, which is 27.")
    return out


def frequency_shift(samples, freq_offset, sample_rate):
    """processor.py:97-100 (numpy evaluates -1j*2*pi*f first, then * t)."""
    t = np.arange(len(samples)) / sample_rate
    shift = np.exp(-1j * 2 * np.pi * freq_offset * t)
    return samples * shift


def filter_signal(samples, bandwidth, sample_rate):
    """processor.py:66-83: butter(4, cutoff) + filtfilt, complex128 out; input returned
    unchanged when filtfilt would raise (len <= 15)."""
    if len(samples) == 0:
        return samples
    b, a, zi = design.channel_filter_design(bandwidth, sample_rate)
    x = np.ascontiguousarray(samples)
    ext_f32 = int(x.dtype == np.complex64)
    xd = x.astype(np.complex128)
    n = len(xd)
    out = np.empty(n, dtype=np.complex128)
    xv = xd.view(np.float64)
    ov = out.view(np.float64)
    for comp in (0, 1):
        rc = lib().orc_filtfilt_lane_f64(_p(b), _p(a), len(b), _p(zi),
                                         ctypes.c_void_p(xv.ctypes.data + 8 * comp), n, 2, ext_f32,
                                         ctypes.c_void_p(ov.ctypes.data + 8 * comp), 2)
        if rc != 0:
            return samples  # processor.py:81-83
    return out


def extract_symbols(samples, sample_rate, symbol_rate=18000):
    """processor.py:179-219.  Returns (symbols, best_phase)."""
    if len(samples) == 0:
        return np.array([], dtype=complex), 0
    sps = int(sample_rate / symbol_rate)
    if sps <= 1:
        return samples, 0
    best_phase, max_power = 0, -1
    step = max(1, sps // 8)
    for phase in range(0, sps, step):
        num = (len(samples) - phase) // sps
        if num <= 0:
            continue
        power = np.mean(np.abs(samples[phase + np.arange(num) * sps]) ** 2)
        if power > max_power:
            max_power, best_phase = power, phase
    num = (len(samples) - best_phase) // sps
    return samples[best_phase + np.arange(num) * sps], best_phase


def demodulate_dqpsk(samples):
    """processor.py:120-166."""
    if len(samples) < 2:
        return np.array([], dtype=np.uint8)
    s = np.asarray(samples)
    mx = np.max(np.abs(s))
    if mx > 0:
        s = s / mx
    s = np.ascontiguousarray(s, dtype=np.complex128)
    out = np.empty(len(s) - 1, dtype=np.uint8)
    lib().orc_slicer(_p(s), len(s), _p(out))
    return out


def process(samples, sample_rate, freq_offset=0.0):
    """SignalProcessor.process, processor.py:221-273.  Returns (dibits, symbols, info)."""
    info = {"q": 1, "best_phase": 0}
    if len(samples) == 0:
        return np.array([], dtype=np.uint8), np.array([], dtype=complex), info
    rate = sample_rate
    q = decimation_factor(sample_rate)
    if q > 1:
        try:
            samples = decimate(samples, q)
            rate = rate / q
            info["q"] = q
        except ValueError:
            pass
    info["decimated"] = samples
    if freq_offset != 0:
        samples = frequency_shift(samples, freq_offset, rate)
    filtered = filter_signal(samples, 25000, rate)
    info["filtered"] = filtered
    symbols, best = extract_symbols(filtered, rate)
    info["best_phase"] = best
    return demodulate_dqpsk(symbols), symbols, info


# ------------------------------------------------------------------- integer half (L3)
FRAME_LENGTH = 510


def symbols_to_bits(symbols):
    """TetraDecoder.symbols_to_bits, core/decoder.py:140-169 -> (bits, mapped) uint8."""
    s = np.asarray(symbols).astype(np.int64)
    if len(s) == 0:
        return np.zeros(0, np.uint8), np.zeros(0, np.uint8)
    if s.max() <= 3:
        mapped = s & 3
    else:
        table = np.array([0, 0, 0, 1, 1, 3, 2, 2], dtype=np.int64)  # decoder.py:160-164
        mapped = np.where((s >= 0) & (s <= 7), table[np.clip(s, 0, 7)], 0)
    bits = np.empty(2 * len(s), dtype=np.uint8)
    bits[0::2] = mapped >> 1
    bits[1::2] = mapped & 1
    return bits, mapped.astype(np.uint8)


def find_sync(bits, threshold=0.85):
    """TetraDecoder.find_sync, core/decoder.py:171-295 -> (positions list, max_corr)."""
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    pos = np.empty(max(len(b), 1), dtype=np.int32)
    mc = ctypes.c_double(0.0)
    n = lib().orc_find_sync(_p(b), len(b), float(threshold), _p(pos), ctypes.byref(mc))
    return [int(v) for v in pos[:n]], float(mc.value)


def find_sync_cascade(bits):
    """The threshold cascade of TetraDecoder.decode, core/decoder.py:845-856."""
    pos, mc = find_sync(bits, 0.90)
    if not pos:
        pos, mc = find_sync(bits, 0.85)
        if not pos:
            pos, mc = find_sync(bits, 0.80)
            if not pos and mc >= 0.75:
                pos, _ = find_sync(bits, max(0.75, mc - 0.02))
    return pos


def crc16(bits):
    """protocol.py:331-347 -> int."""
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    return int(lib().orc_crc16(_p(b), len(b), 1))


def parse_bursts(dibits):
    """parse_burst integer path over a [n, 255] batch, protocol.py:192-329.

    Returns dict(burst_type u8 [n] (BurstType values: 2 normal downlink, 5 sync),
    crc_ok u8 [n], data_bits u8 [n, 510] (first n_data valid), n_data i32 [n])."""
    d = np.ascontiguousarray(dibits, dtype=np.uint8).reshape(-1, 255)
    n = d.shape[0]
    bt = np.empty(n, np.uint8)
    ok = np.empty(n, np.uint8)
    data = np.zeros((n, 510), np.uint8)
    nd = np.empty(n, np.int32)
    lib().orc_parse_bursts(_p(d), n, _p(bt), _p(ok), _p(data), _p(nd))
    return {"burst_type": bt, "crc_ok": ok, "data_bits": data, "n_data": nd}


def mac_pdu_valid(data_bits):
    """True when TetraProtocolParser.parse_mac_pdu (protocol.py:349-596) returns a PDU,
    False when it returns None — the only part of the MAC stage the frame drop rule
    (core/decoder.py:1092-1100) depends on.  Pure function of the first 35 data bits and
    the data length."""
    b = [int(v) for v in data_bits]
    n = len(b)
    if n < 8:
        return False

    def val(lo, hi):
        v = 0
        for x in b[lo:hi]:
            v = (v << 1) | x
        return v

    t = (b[0] << 1) | b[1]
    if t == 0:  # MAC-RESOURCE protocol.py:399-434
        if n < 5 + 24 or n < 5 + 24 + 6:
            return False
        length = val(29, 35)
        return not (length * 8 > n - 35 + 16)
    if t == 1:  # MAC-FRAG
        return True
    if t == 2:  # MAC-BROADCAST protocol.py:471-498
        if ((b[2] << 1) | b[3]) == 0:
            if n < 4 + 30:
                return False
            mcc, mnc = val(4, 14), val(14, 28)
            if mcc < 200 or mcc > 799 or mnc > 999:
                return False
        return True
    # MAC-END protocol.py:506-526
    if n < 5 + 6:
        return False
    length = val(5, 11)
    return not (length * 8 > n - 11 + 16)


def decode(dibits):
    """TetraDecoder.decode (core/decoder.py:835-888) + the integer head of decode_frame
    (:890-992, :1092-1100).  Returns (frames, stats): one dict per kept frame with
    position (= sync pos), start_pos, number, timeslot, type, enc_mode, encrypted (header
    bit test only), burst_type, burst_crc, mac_ok, bits (510 uint8); stats counts
    total_bursts / crc_pass / crc_fail like TetraProtocolParser.stats (protocol.py:228-232)."""
    stats = {"total_bursts": 0, "crc_pass": 0, "crc_fail": 0}
    bits, mapped = symbols_to_bits(dibits)
    frames = []
    for pos in find_sync_cascade(bits):
        start_pos = pos - 216
        if start_pos < 0:
            continue
        start_sym = start_pos // 2
        if start_sym + 255 > len(mapped):
            continue
        fbits = bits[start_pos:start_pos + 510]
        if len(fbits) < FRAME_LENGTH:  # decode_frame :894 — before parse_burst, no stats
            continue
        pb = parse_bursts(mapped[start_sym:start_sym + 255])
        crc_ok = bool(pb["crc_ok"][0])
        stats["total_bursts"] += 1
        stats["crc_pass" if crc_ok else "crc_fail"] += 1
        nd = int(pb["n_data"][0])
        mac_ok = mac_pdu_valid(pb["data_bits"][0, :nd])
        if not mac_ok and not crc_ok:
            continue
        number = start_pos // 510
        frames.append({
            "sync_pos": pos, "start_pos": start_pos, "number": number, "timeslot": number % 4,
            "type": int(fbits[0]) * 2 + int(fbits[1]),
            "enc_mode": int(fbits[2]) * 2 + int(fbits[3]),
            "encrypted": (int(fbits[2]) * 2 + int(fbits[3])) > 0,
            "burst_type": int(pb["burst_type"][0]), "burst_crc": crc_ok, "mac_ok": mac_ok,
            "bits": fbits.copy(),
        })
    return frames, stats


# ----------------------------------------------------------------------- PSD / AFC (L4)
def psd_frame(samples, sample_rate, n_fft=2048):
    """The inline spectrum block of CaptureThread.run, tetraear/ui/modern.py:1928-1941.
    Returns (freqs, power_db) float64 [n_fft]."""
    x = np.asarray(samples)[:n_fft]
    window = np.hanning(n_fft)
    fft = np.fft.fftshift(np.fft.fft(x * window))
    freqs = np.fft.fftshift(np.fft.fftfreq(n_fft, 1 / sample_rate))
    power = 20 * np.log10(np.abs(fft) / n_fft + 1e-20)
    return freqs, power


def psd_welch(samples, sample_rate, n_fft=4096, average=8):
    """BASELINE configs[1] "4096-pt windowed FFT w/ averaging": the reference's per-frame amplitude spectrum
    (ui/modern.py:1932-1941: Hann window, fftshift(fft), |X|/N) averaged LINEARLY in power over `average` consecutive
    dense frames, then the reference's dBFS formula on the RMS amplitude, 20 log10(rms + 1e-20).  average = 1 is
    psd_frame.  The reference itself has no linear averaging (dB-domain EMA only) — parity unpinned, this is
    the restatement the CUDA kernel is checked against.  Returns float64 [n_out, n_fft]."""
    x = np.asarray(samples)
    nfr = len(x) // n_fft
    n_out = nfr // average
    w = np.hanning(n_fft)
    out = np.empty((n_out, n_fft))
    for o in range(n_out):
        acc = np.zeros(n_fft)
        for k in range(average):
            seg = x[(o * average + k) * n_fft:(o * average + k + 1) * n_fft]
            acc += (np.abs(np.fft.fftshift(np.fft.fft(seg * w))) / n_fft) ** 2
        out[o] = 20 * np.log10(np.sqrt(acc / average) + 1e-20)
    return out


def band_stats(power, sample_rate, n_fft=None):
    """ui/modern.py:1956-2010, :2028 -> dict(signal_power, peak_power, peak_idx,
    peak_freq_offset, noise_floor, snr, signal_present, afc_offset)."""
    n_fft = len(power) if n_fft is None else n_fft
    freqs = np.fft.fftshift(np.fft.fftfreq(n_fft, 1 / sample_rate))
    center = len(power) // 2
    bw_bins = int(25000 / (sample_rate / n_fft))
    start = max(0, center - bw_bins // 2)
    end = min(len(power), center + bw_bins // 2)
    if end <= start:
        return None
    sig = np.mean(power[start:end])
    peak = np.max(power[start:end])
    peak_idx = start + int(np.argmax(power[start:end]))
    noise = []
    if max(0, start - 10) > 0:
        noise.extend(power[0:max(0, start - 10)])
    if len(power) > min(len(power), end + 10):
        noise.extend(power[min(len(power), end + 10):])
    noise_floor = np.mean(noise) if noise else -100
    snr = sig - noise_floor
    present = bool(snr > 15 and peak > -70 and (peak - sig) > 3)
    return {
        "signal_power": float(sig), "peak_power": float(peak), "peak_idx": peak_idx,
        "peak_freq_offset": float(freqs[peak_idx]), "noise_floor": float(noise_floor),
        "snr": float(snr), "signal_present": present,
        "afc_offset": float(freqs[peak_idx]) if present and peak > -70 else 0.0,
    }


def ema_denoise(frames_db, alpha=0.15):
    """WaterfallWidget.update_spectrum EMA, ui/modern.py:929-935, over a [n, N] stack."""
    out = np.empty_like(frames_db)
    s = None
    for i, p in enumerate(frames_db):
        s = p if s is None else s * (1 - alpha) + p * alpha
        out[i] = s
    return out


def find_peak_near(freqs_mhz, powers, freq_mhz, search_range_hz):
    """WaterfallWidget.find_peak_near, ui/modern.py:727-772 (parabolic sub-bin peak)."""
    freqs = np.asarray(freqs_mhz, dtype=np.float64)
    powers = np.asarray(powers, dtype=np.float64)
    if len(freqs) == 0 or len(powers) == 0:                    # :729-735
        return freq_mhz
    if freqs.shape != powers.shape:                            # :745-748
        n = min(len(freqs), len(powers))
        freqs, powers = freqs[:n], powers[:n]
    half = search_range_hz / 1e6
    mask = (freqs >= float(freq_mhz) - half) & (freqs <= float(freq_mhz) + half)
    if not np.any(mask):
        return freq_mhz
    idxs = np.where(mask)[0]
    pk = int(idxs[np.argmax(powers[idxs])])
    if 0 < pk < len(powers) - 1:
        y1, y2, y3 = powers[pk - 1], powers[pk], powers[pk + 1]
        den = y1 - 2 * y2 + y3
        if den != 0:
            delta = float(max(-1.0, min(1.0, 0.5 * (y1 - y3) / den)))
            x1, x2, x3 = freqs[pk - 1], freqs[pk], freqs[pk + 1]
            step = float((x3 - x1) / 2.0) if x3 != x1 else float(x3 - x2)
            return float(x2 + delta * step)
    return float(freqs[pk])


def afc_step(freqs_mhz, powers, freq_mhz):
    """Auto-follow rule of ModernTetraGUI.on_spectrum, ui/modern.py:5223-5257."""
    freqs = np.asarray(freqs_mhz)
    powers = np.asarray(powers)
    mask = np.abs(freqs - freq_mhz) < 0.010
    if not np.any(mask):
        return freq_mhz
    win = np.where(mask)[0]
    pk = win[np.argmax(powers[win])]
    offset = freqs[pk] - freq_mhz
    if powers[pk] > -60 and abs(offset) > 0.00005:
        return freq_mhz + (offset * 0.1)
    return freq_mhz
