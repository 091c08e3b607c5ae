"""CPU restatement of the scanner-side signal heuristics — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/tetraear/signal/scanner.py:
    calculate_power          :42-55     mean |x|^2 in dB (+1e-10), `bottom_threshold` for an empty input
    detect_tetra_modulation  :57-98     fraction of phase steps within pi/8 of a multiple of pi/4  (> 0.4)
    detect_sync_pattern      :100-147   every int(fs/18000/10)-th sample, phase step -> 1 bit (|round(4 phi/pi)| == 0),
                                        best sliding match against a 31-bit pattern (> 0.75)
    validate_frames          :149-202   process() + decode(); CRC pass rate (burst_crc True = 1, missing = 0.5)
    check_power_stability    :204-233   std of 5 window powers (dB) < 10
    analyze_signal           :235-289   combination rules
The reference loops sample by sample in Python; here the same arithmetic is vectorised (numpy keeps the
reference's dtypes: complex64 input => float32 angles and powers, complex128 => float64).
Only tests/, __graft_entry__.smoke() and bench.py's cpu baseline may import this module.
Parity pinned by tests/golden/scanner.npz (outputs of the live reference) and test_oracle_vs_reference.py.
"""
import numpy as np

from . import oracle as _o

SYNC31 = np.array([0, 1, 0, 1, 1, 0, 0, 1, 1, 1, 0, 0, 0, 1, 0, 0,
                   1, 0, 1, 1, 0, 0, 1, 1, 1, 0, 0, 0, 1, 0, 0], dtype=np.int64)        # scanner.py:133-134
EXPECTED = np.array([-np.pi, -3 * np.pi / 4, -np.pi / 2, -np.pi / 4, 0, np.pi / 4, np.pi / 2, 3 * np.pi / 4])   # :77-78


def calculate_power(samples, bottom_threshold=-85):
    samples = np.asarray(samples)
    if samples.size == 0:
        return float(bottom_threshold)
    return 10 * np.log10(np.mean(np.abs(samples) ** 2) + 1e-10)


def _wrapped_steps(x):
    d = np.diff(np.angle(x))
    return (d + np.pi) % (2 * np.pi) - np.pi


def modulation_matches(samples):
    """(matches, n_steps) of detect_tetra_modulation; (0, 0) below 1000 samples (:66-67)."""
    samples = np.asarray(samples)
    if len(samples) < 1000:
        return 0, 0
    x = samples / (np.abs(samples).max() + 1e-10)
    d = _wrapped_steps(x)
    matches = 0
    for lo in range(0, len(d), 1 << 16):              # bounded memory: [block, 8] float64
        blk = d[lo:lo + (1 << 16)]
        dist = np.abs(EXPECTED[None, :] - blk[:, None])
        matches += int(np.count_nonzero(dist.min(axis=1) < np.pi / 8))
    return matches, len(d)


def detect_tetra_modulation(samples):
    m, n = modulation_matches(samples)
    if n == 0:
        return False, 0.0
    conf = m / n
    return conf > 0.4, conf


def sync_bits(samples, sample_rate):
    """The 0/1 stream detect_sync_pattern correlates (:108-128); None when fewer than 100 'symbols'."""
    samples = np.asarray(samples)
    step = max(1, int(sample_rate / 18000 / 10))
    sym = samples[::step]
    if len(sym) < 100:
        return None
    d = _wrapped_steps(sym)
    q = np.rint(d / (np.pi / 4))                       # round(): half to even, like np.rint
    return (np.abs(q * (np.pi / 4)) < np.pi / 8).astype(np.int64)


def sync_best(bits):
    """max over i in range(len-31) of the number of positions equal to the 31-bit pattern (:137-141)."""
    n = len(bits) - 31
    if n <= 0:
        return 0
    win = np.lib.stride_tricks.sliding_window_view(bits, 31)[:n]
    return int((win == SYNC31[None, :]).sum(axis=1).max())


def detect_sync_pattern(samples, sample_rate):
    bits = sync_bits(samples, sample_rate)
    if bits is None or len(bits) < 31:
        return False, 0.0
    corr = max(0.0, sync_best(bits) / 31)
    return corr > 0.75, corr


def window_powers(samples, num_windows=5, bottom_threshold=-85):
    samples = np.asarray(samples)
    w = len(samples) // num_windows
    return [calculate_power(samples[i * w:(i + 1) * w], bottom_threshold) for i in range(num_windows)]


def check_power_stability(samples, num_windows=5, bottom_threshold=-85):
    if len(samples) < num_windows * 1000:
        return False
    p = window_powers(samples, num_windows, bottom_threshold)
    if len(p) > 1:
        return bool(np.std(p) < 10.0)
    return True


def validate_frames(samples, sample_rate):
    samples = np.asarray(samples)
    if len(samples) < 10000:
        return False, 0.0
    dibits, _, _ = _o.process(samples, sample_rate, 0.0)
    if len(dibits) < 255:
        return False, 0.0
    frames, _ = _o.decode(dibits)
    if not frames:
        return False, 0.0
    good = sum(1.0 if f["burst_crc"] else 0.0 for f in frames)       # burst_crc is always set by decode_frame (:992)
    rate = good / max(len(frames), 1)
    return len(frames) >= 2 and rate > 0.5, rate


def analyze_signal(samples, sample_rate=2.4e6, bottom_threshold=-85):
    power = calculate_power(samples, bottom_threshold)
    is_mod, mod_conf = detect_tetra_modulation(samples)
    has_sync, sync_corr = detect_sync_pattern(samples, sample_rate)
    frames_valid, crc_rate = validate_frames(samples, sample_rate)
    stable = check_power_stability(samples, 5, bottom_threshold)
    if has_sync and is_mod:
        conf = mod_conf * 0.4 + sync_corr * 0.4 + crc_rate * 0.2
    elif has_sync:
        conf = sync_corr * 0.6
    elif is_mod:
        conf = mod_conf * 0.5
    else:
        conf = 0.0
    is_tetra = bool(is_mod and has_sync and stable)
    if frames_valid:
        is_tetra = True
        conf = max(conf, 0.7)
    return {"power_db": power, "is_tetra": is_tetra, "confidence": conf, "modulation_confidence": mod_conf,
            "sync_detected": has_sync, "sync_correlation": sync_corr, "frames_validated": frames_valid,
            "crc_pass_rate": crc_rate, "power_stable": stable, "signal_present": bool(power > bottom_threshold)}
