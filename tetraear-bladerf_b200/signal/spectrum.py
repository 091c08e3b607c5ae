"""Waterfall PSD + AFC statistics — the spectrum entry point the reference lacks as a function.

In the reference this arithmetic is inline in ``CaptureThread.run``
(/root/reference/tetraear/ui/modern.py:1926-2029) and in ``WaterfallWidget.update_spectrum``
(:923-947, the dB-domain EMA "denoiser").  ``compute_spectrum`` is the function
``CaptureThread.run`` calls instead; its ``(freqs_hz, power_db)`` pair is
exactly the payload of the ``spectrum_update`` signal (modern.py:1794).
"""
import numpy as np

from .. import _lib

STAT_KEYS = ("signal_power", "peak_power", "peak_idx", "peak_freq_offset", "noise_floor", "snr",
             "signal_present", "afc_offset")


def _freq_axis(n_fft, sample_rate):
    # np.fft.fftshift(np.fft.fftfreq(n_fft, 1 / sample_rate)) written out (modern.py:1936)
    val = 1.0 / (n_fft * (1 / sample_rate))
    return (np.arange(n_fft, dtype=np.int64) - n_fft // 2) * val


class SpectrumAnalyzer:
    """Batched Hann-window FFT PSD with optional dB-domain EMA, B200 resident.

    reference mode : n_fft=2048, one frame from the first n_fft samples of every chunk
                     (frame_stride = chunk length), EMA alpha 0.15 if denoise.
    config mode    : any power-of-two n_fft (e.g. 4096), frame_stride = n_fft (dense).
    """

    def __init__(self, sample_rate=2.4e6, n_fft=2048, ema_alpha=0.15, device=0, precision="f64", average=1):
        self.sample_rate = float(sample_rate)
        self.n_fft = int(n_fft)
        self.ema_alpha = float(ema_alpha)
        self.device = device
        self._h = _lib.Handle(sample_rate=self.sample_rate, max_chunks=0, max_chunk_len=0, device=device,
                              max_fft=max(self.n_fft, 4096))
        if precision == "f32":
            self._h.set_option(_lib.OPT_PSD_F32, 1)
        self.precision = precision
        # linear (Welch) averaging: every output frame = power average of `average` consecutive input frames (then dB);
        # 1 = the reference's per-frame spectrum.  The EMA ("denoise") then runs over the averaged frames.
        self.average = max(1, int(average))
        if self.average > 1:
            self._h.set_option(_lib.OPT_PSD_AVG, self.average)
        self._state = None
        self._first = True
        self.freqs = _freq_axis(self.n_fft, self.sample_rate)

    def reset(self):
        self._first = True

    def run_device(self, iq, n_frames, frame_stride, denoise=False, want_stats=True, power=None, smoothed=None,
                   stats=None):
        """iq: CUDA float32 tensor viewing interleaved complex64.  Returns (power, smoothed, stats)
        CUDA tensors (float32 [n_frames, n_fft] x2, float64 [n_frames, 8])."""
        import torch
        from .. import _dev
        dev = self.device
        n_out = int(n_frames) // self.average            # rows written (== n_frames without averaging)
        if power is None:
            power = _dev.empty((n_out, self.n_fft), torch.float32, dev)
        if denoise and smoothed is None:
            smoothed = _dev.empty((n_out, self.n_fft), torch.float32, dev)
        if want_stats and stats is None:
            stats = _dev.empty((n_out, 8), torch.float64, dev)
        if self._state is None:
            self._state = _dev.zeros((self.n_fft,), torch.float64, dev)
        h = self._h
        _lib.check(h.lib.tb200_psd_batch(h.h, iq.data_ptr(), int(n_frames), int(frame_stride), self.n_fft,
                                         self.sample_rate, self.ema_alpha, int(self._first), power.data_ptr(),
                                         smoothed.data_ptr() if denoise else None, self._state.data_ptr(),
                                         stats.data_ptr() if want_stats else None, _dev.stream_ptr(dev)))
        if denoise:
            self._first = False
        return power, (smoothed if denoise else None), (stats if want_stats else None)

    def run(self, samples, frame_stride=None, denoise=False):
        """Host convenience: complex64 numpy in -> dict(freqs, power [n, N] float64, smoothed, stats list)."""
        from .. import _dev
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        stride = self.n_fft if frame_stride is None else int(frame_stride)
        if len(x) < self.n_fft:
            return {"freqs": self.freqs, "power": np.zeros((0, self.n_fft)), "smoothed": None, "stats": []}
        n_frames = (len(x) - self.n_fft) // stride + 1
        dx = _dev.to_dev(x.view(np.float32), self.device)
        p, s, st = self.run_device(dx, n_frames, stride, denoise=denoise)
        st = st.cpu().numpy()
        stats = []
        for row in st:
            d = dict(zip(STAT_KEYS, (float(v) for v in row)))
            d["peak_idx"] = int(d["peak_idx"])
            d["signal_present"] = bool(d["signal_present"])
            stats.append(d)
        return {"freqs": self.freqs, "power": p.cpu().numpy().astype(np.float64),
                "smoothed": None if s is None else s.cpu().numpy().astype(np.float64), "stats": stats}


_ANALYZERS = {}


def compute_spectrum(samples, sample_rate, center_freq=0.0, n_fft=2048, device=0):
    """One spectrum frame exactly as CaptureThread.run computes it (modern.py:1928-2029).

    Returns (freqs_hz + center_freq, power_db float64 [n_fft], stats dict) or (None, None, None)
    when fewer than n_fft samples are available (the reference skips the block, :1929)."""
    if len(samples) < n_fft:
        return None, None, None
    key = (float(sample_rate), int(n_fft), device)
    an = _ANALYZERS.get(key)
    if an is None:
        an = _ANALYZERS[key] = SpectrumAnalyzer(sample_rate, n_fft, device=device)
    r = an.run(np.asarray(samples)[:n_fft], frame_stride=n_fft, denoise=False)
    return r["freqs"] + center_freq, r["power"][0], r["stats"][0]


def find_peak_near(freqs_mhz, powers, freq_mhz, search_range_hz):
    """WaterfallWidget.find_peak_near (modern.py:727-772): strongest bin within the range,
    refined by a 3-point parabolic fit clamped to +-1 bin.  Scalar GUI helper (a click handler),
    host side by design — it touches 3 numbers."""
    freqs = np.asarray(freqs_mhz, dtype=np.float64)
    powers = np.asarray(powers, dtype=np.float64)
    if len(freqs) == 0 or len(powers) == 0:
        return freq_mhz
    n = min(len(freqs), len(powers))
    freqs, powers = freqs[:n], powers[:n]
    half = search_range_hz / 1e6
    idx = np.nonzero((freqs >= float(freq_mhz) - half) & (freqs <= float(freq_mhz) + half))[0]
    if len(idx) == 0:
        return freq_mhz
    pk = int(idx[np.argmax(powers[idx])])
    if 0 < pk < n - 1:
        y1, y2, y3 = powers[pk - 1], powers[pk], powers[pk + 1]
        den = y1 - 2 * y2 + y3
        if den != 0:
            delta = float(max(-1.0, min(1.0, 0.5 * (y1 - y3) / den)))
            x1, x2, x3 = freqs[pk - 1], freqs[pk], freqs[pk + 1]
            step = float((x3 - x1) / 2.0) if x3 != x1 else float(x3 - x2)
            return float(x2 + delta * step)
    return float(freqs[pk])


def find_spike_band_near(freqs_mhz, powers, freq_mhz, search_range_hz, noise_floor=-100.0, bandwidth=25000.0):
    """WaterfallWidget.find_spike_band_near (modern.py:774-857): the strongest bin within the range and the contiguous
    run around it that stays above max(peak - 6 dB, 20th percentile of the window + 6 dB, noise_floor + 3 dB); returns
    centre / edges (MHz) and the bandwidth clamped to 1 kHz ... 250 kHz.  `noise_floor` and `bandwidth` are the widget
    attributes the reference reads.  Scalar GUI helper, host side like find_peak_near."""
    nothing = {"center_mhz": float(freq_mhz), "left_mhz": float(freq_mhz), "right_mhz": float(freq_mhz),
               "bandwidth_hz": float(bandwidth)}
    if freqs_mhz is None or powers is None or len(freqs_mhz) == 0 or len(powers) == 0:
        return nothing
    freqs = np.asarray(freqs_mhz, dtype=np.float64)
    pw = np.asarray(powers, dtype=np.float64)
    n = min(len(freqs), len(pw))
    freqs, pw = freqs[:n], pw[:n]
    half = float(search_range_hz) / 1e6
    idx = np.nonzero((freqs >= float(freq_mhz) - half) & (freqs <= float(freq_mhz) + half))[0]
    if len(idx) == 0:
        return nothing
    pk = int(idx[np.argmax(pw[idx])])
    thr = max(float(pw[pk]) - 6.0, float(np.percentile(pw[idx], 20)) + 6.0, float(noise_floor) + 3.0)
    left = pk
    while left > idx[0] and pw[left] >= thr:
        left -= 1
    if pw[left] < thr and left < pk:
        left += 1
    right = pk
    while right < idx[-1] and pw[right] >= thr:
        right += 1
    if pw[right] < thr and right > pk:
        right -= 1
    lo, hi = float(freqs[left]), float(freqs[right])
    if hi < lo:
        lo, hi = hi, lo
    return {"center_mhz": (lo + hi) / 2.0, "left_mhz": lo, "right_mhz": hi,
            "bandwidth_hz": float(max(1000.0, min(250000.0, (hi - lo) * 1e6)))}


def afc_step(freqs_mhz, powers, freq_mhz, window_mhz=0.010, min_power_db=-60.0, min_offset_mhz=0.00005, gain=0.1):
    """The GUI's auto-follow rule (ModernTetraGUI.on_spectrum, modern.py:5223-5257): strongest bin within
    +-10 kHz of the tuned frequency; if it is above -60 dB and more than 50 Hz away, move 10 % of the way.
    Returns the new tuned frequency in MHz (unchanged when the rule does not fire).  Scalar host logic."""
    freqs = np.asarray(freqs_mhz)
    powers = np.asarray(powers)
    idx = np.nonzero(np.abs(freqs - freq_mhz) < window_mhz)[0]
    if len(idx) == 0:
        return freq_mhz
    pk = idx[np.argmax(powers[idx])]
    offset = freqs[pk] - freq_mhz
    if powers[pk] > min_power_db and abs(offset) > min_offset_mhz:
        return freq_mhz + (offset * gain)
    return freq_mhz
