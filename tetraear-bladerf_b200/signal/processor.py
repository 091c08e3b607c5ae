"""``SignalProcessor`` — drop-in for ``tetraear.signal.processor.SignalProcessor``
(/root/reference/tetraear/signal/processor.py:18-273), computed on a B200.

Same constructor, attributes (``sample_rate`` may be changed live, ``symbol_rate``,
``samples_per_symbol``, ``symbols``), method names, argument meaning, return dtypes and
"never raise on a bad signal" behaviour.  Every method marshals numpy -> device, makes one
C-ABI call into libtetra_b200 and unmarshals; nothing is computed on the CPU.
"""
import ctypes as C
import logging

import numpy as np

from .. import _lib, _spec

logger = logging.getLogger(__name__)

# "fast" / "fast32": time-parallel float32 decimator + float64 / float32 channel filter; the "-tc" variants run the
# decimator's linear part on the tensor cores (tcgen05, csrc/dsp_tc.cu) instead of packed FFMA2 (csrc/dsp_fast.cu)
_MODES = {"exact": _lib.MODE_EXACT, "fast": _lib.MODE_FAST, "fast32": _lib.MODE_FAST, "fast-tc": _lib.MODE_FAST,
          "fast32-tc": _lib.MODE_FAST, "fast-tc4": _lib.MODE_FAST, "fast32-tc4": _lib.MODE_FAST}


def _as_complex(samples):
    """complex64 stays complex64 (float32 decimator, like scipy); anything else -> complex128."""
    x = np.asarray(samples)
    if x.dtype == np.complex64:
        return np.ascontiguousarray(x), _lib.DTYPE_C64
    return np.ascontiguousarray(x, dtype=np.complex128), _lib.DTYPE_C128


class SignalProcessor:
    """Processes raw IQ samples for TETRA demodulation (B200 implementation)."""

    def __init__(self, sample_rate=2.4e6, mode="exact", device=0):
        self.sample_rate = sample_rate
        self.symbol_rate = 18000
        self.samples_per_symbol = int(sample_rate / self.symbol_rate)
        self.symbols = None
        self.mode = mode
        self.device = device
        self._h = None
        self._cap = (0, 0.0)

    # ------------------------------------------------------------------ handle management
    def _handle(self, n, fs=None):
        """A handle whose workspace fits one chunk of n samples at the current sample rate."""
        fs = float(self.sample_rate if fs is None else fs)
        cap, rate = self._cap
        if self._h is None or n > cap or rate != fs:
            if self._h is not None:
                self._h.close()
            # complex128 input needs twice the float32 scratch: size for it
            cap = max(2 * int(n) + 128, 1 << 16)
            self._h = _lib.Handle(sample_rate=fs, max_chunks=1, max_chunk_len=cap, mode=_MODES[self.mode],
                                  device=self.device)
            if self.mode.startswith("fast32"):
                self._h.set_option(_lib.OPT_CHAN_F32, 1)
            if self.mode.endswith("-tc") or self.mode.endswith("-tc4"):
                self._h.set_option(_lib.OPT_DECIM_TC, 2 if self.mode.endswith("4") else 1)
            self._cap = ((cap - 128) // 2, fs)
        return self._h

    # ------------------------------------------------------------------ public API
    def resample(self, samples, target_rate):
        """processor.py:35-49: scipy.signal.resample(samples, int(n * target_rate / sample_rate)) on the GPU
        (direct DFT; no caller on the receive path uses it).  Real input comes back real, like scipy's."""
        x_in = np.asarray(samples)
        n = len(x_in)
        num = int(n * target_rate / self.sample_rate)
        if n == 0 or num <= 0:
            return np.zeros(max(num, 0), dtype=x_in.dtype if x_in.dtype.kind in "fc" else np.float64)
        x, dt = _as_complex(x_in)
        import torch
        from .. import _dev
        # scratch = the decimator workspace (~32 bytes per sample of capacity): room for the two twiddle tables + bins
        h = self._handle(n + 2 * num + 64)
        real = np.float32 if dt == _lib.DTYPE_C64 else np.float64
        dx = _dev.to_dev(x.view(real), self.device)
        out = _dev.empty((2 * num,), torch.float32 if dt == _lib.DTYPE_C64 else torch.float64, self.device)
        _lib.check(h.lib.tb200_resample(h.h, dx.data_ptr(), dt, n, num, out.data_ptr(), _dev.stream_ptr(self.device)))
        y = out.cpu().numpy().view(np.complex64 if dt == _lib.DTYPE_C64 else np.complex128)
        return y if np.iscomplexobj(x_in) else y.real.astype(x_in.dtype if x_in.dtype.kind == "f" else np.float64)

    def process(self, samples, freq_offset=0):
        """processor.py:221-273: decimate -> NCO -> Butterworth -> timing pick -> DQPSK slicer.
        Like the reference (its stages log and carry on, processor.py:81-83, :256-257; callers rely on it,
        ui/modern.py:2087-2089), an input the kernels refuse is logged and answered with an empty result instead of
        an exception; a missing library / device still raises."""
        if len(samples) == 0:
            self.symbols = np.array([], dtype=complex)
            return np.array([], dtype=np.uint8)
        try:
            return self._process(samples, freq_offset)
        except _lib.ComputeError as e:
            logger.error(f"process failed: {e}")
            self.symbols = np.array([], dtype=complex)
            return np.array([], dtype=np.uint8)

    def _process(self, samples, freq_offset):
        x, dt = _as_complex(samples)
        n = len(x)
        h = self._handle(n)
        p = _lib.plan(self.sample_rate, n)
        stride = max(int(p.max_symbols), 1)
        dibits = np.empty(stride, dtype=np.uint8)
        symbols = np.empty(stride, dtype=np.complex128)
        cnt = np.zeros(1, dtype=np.int32)
        phase = np.zeros(1, dtype=np.int32)
        f = np.array([float(freq_offset)], dtype=np.float64)
        max_frames = (2 * stride) // 250 + 2
        if _spec.enabled and stride <= 6000 and h.lib.tb200_host_oneshot_fits(h.h, 1, stride, max_frames, 1):
            # process + the decoder's device part in one call; decode() of these dibits finds its records in _spec
            rec = np.zeros(max_frames, dtype=_lib.FRAME_DTYPE)
            nfr = np.zeros(1, dtype=np.int32)
            stats = np.zeros(3, dtype=np.int64)
            _lib.check(h.lib.tb200_process_decode_host(h.h, x.ctypes.data, dt, 1, n, f.ctypes.data, dibits.ctypes.data,
                                                       symbols.ctypes.data, stride, cnt.ctypes.data, phase.ctypes.data,
                                                       rec.ctypes.data, max_frames, nfr.ctypes.data, stats.ctypes.data, None))
            out = dibits[:int(cnt[0])].copy()
            _spec.publish(out.copy(), rec[:int(nfr[0])].copy(), stats)
        else:
            _lib.check(h.lib.tb200_process_batch_host(h.h, x.ctypes.data, dt, 1, n, f.ctypes.data, dibits.ctypes.data,
                                                      symbols.ctypes.data, stride, cnt.ctypes.data, phase.ctypes.data, None))
            out = dibits[:int(cnt[0])].copy()
        nsym = self._n_symbols(p, int(phase[0]))
        self.symbols = symbols[:nsym].copy()
        self.best_phase = int(phase[0])
        return out

    def process_sc16(self, iq_int16, freq_offset=0):
        """process() for the bladeRF's native SC16Q11 buffer (interleaved int16 I, Q): the conversion
        ``complex64(I, Q) / 32768`` of BladeRFCapture.read_samples (signal/capture.py:255-269) runs on the
        device, so only 4 bytes per sample cross PCIe.  Bit-identical to process(complex64(I, Q) / 32768)."""
        raw = np.ascontiguousarray(iq_int16, dtype=np.int16).reshape(-1, 2)
        n = raw.shape[0]
        if n == 0:
            self.symbols = np.array([], dtype=complex)
            return np.array([], dtype=np.uint8)
        h = self._handle(n)
        p = _lib.plan(self.sample_rate, n)
        stride = max(int(p.max_symbols), 1)
        dibits = np.empty(stride, dtype=np.uint8)
        symbols = np.empty(stride, dtype=np.complex128)
        cnt = np.zeros(1, dtype=np.int32)
        phase = np.zeros(1, dtype=np.int32)
        f = np.array([float(freq_offset)], dtype=np.float64)
        _lib.check(h.lib.tb200_process_batch_host(h.h, raw.ctypes.data, _lib.DTYPE_SC16, 1, n, f.ctypes.data,
                                                  dibits.ctypes.data, symbols.ctypes.data, stride, cnt.ctypes.data,
                                                  phase.ctypes.data, None))
        self.symbols = symbols[:self._n_symbols(p, int(phase[0]))].copy()
        self.best_phase = int(phase[0])
        return dibits[:int(cnt[0])].copy()

    @staticmethod
    def _n_symbols(p, best):
        if p.n_dec <= 0:
            return 0
        if p.sps <= 1:
            return int(p.n_dec)
        return max((int(p.n_dec) - best) // int(p.sps), 0)

    def frequency_shift(self, samples, freq_offset, sample_rate=None):
        """processor.py:85-100."""
        fs = sample_rate if sample_rate is not None else self.sample_rate
        x, dt = _as_complex(samples)
        n = len(x)
        if n == 0:
            return np.asarray(samples) * np.complex128(1)
        import torch
        from .. import _dev
        h = self._handle(n)
        dx = _dev.to_dev(x.view(np.float32 if dt == _lib.DTYPE_C64 else np.float64), self.device)
        out = _dev.empty((2 * n,), torch.float64, self.device)
        _lib.check(h.lib.tb200_freq_shift(h.h, dx.data_ptr(), dt, n, float(freq_offset), float(fs), out.data_ptr(),
                                          _dev.stream_ptr(self.device)))
        return out.cpu().numpy().view(np.complex128)

    def filter_signal(self, samples, bandwidth=25000, sample_rate=None):
        """processor.py:51-83; like the reference, returns the input when it is too short to filter."""
        if len(samples) == 0:
            return samples
        fs = sample_rate if sample_rate is not None else self.sample_rate
        x, dt = _as_complex(samples)
        n = len(x)
        if n <= 15:
            logger.warning("Filter design failed, using unfiltered samples: input too short")
            return samples
        import torch
        from .. import _dev
        h = self._handle(n)
        dx = _dev.to_dev(x.view(np.float32 if dt == _lib.DTYPE_C64 else np.float64), self.device)
        out = _dev.empty((2 * n,), torch.float64, self.device)
        _lib.check(h.lib.tb200_filter(h.h, dx.data_ptr(), dt, n, float(bandwidth), float(fs), out.data_ptr(),
                                      _dev.stream_ptr(self.device)))
        return out.cpu().numpy().view(np.complex128)

    def extract_symbols(self, samples, sample_rate=None):
        """processor.py:168-219."""
        if len(samples) == 0:
            return np.array([], dtype=complex)
        fs = sample_rate if sample_rate is not None else self.sample_rate
        x = np.ascontiguousarray(samples, dtype=np.complex128)
        n = len(x)
        import torch
        from .. import _dev
        h = self._handle(n)
        dx = _dev.to_dev(x.view(np.float64), self.device)
        out = _dev.empty((2 * (n + 1),), torch.float64, self.device)
        meta = _dev.zeros((2,), torch.int32, self.device)
        _lib.check(h.lib.tb200_extract_symbols(h.h, dx.data_ptr(), n, float(fs), out.data_ptr(), meta.data_ptr(),
                                               meta.data_ptr() + 4, _dev.stream_ptr(self.device)))
        m = meta.cpu().numpy()
        return out.cpu().numpy().view(np.complex128)[: int(m[0])].copy()

    def demodulate_dqpsk(self, samples):
        """processor.py:102-166."""
        if len(samples) < 2:
            return np.array([], dtype=np.uint8)
        x = np.ascontiguousarray(samples, dtype=np.complex128)
        n = len(x)
        import torch
        from .. import _dev
        h = self._handle(n)
        dx = _dev.to_dev(x.view(np.float64), self.device)
        out = _dev.empty((n - 1,), torch.uint8, self.device)
        _lib.check(h.lib.tb200_demod_dqpsk(h.h, dx.data_ptr(), n, out.data_ptr(), _dev.stream_ptr(self.device)))
        return out.cpu().numpy()

    def decimate(self, samples, q):
        """scipy.signal.decimate(samples, q) as called at processor.py:254 (stage-level access)."""
        x, dt = _as_complex(samples)
        n = len(x)
        import torch
        from .. import _dev
        h = self._handle(n)
        real = np.float32 if dt == _lib.DTYPE_C64 else np.float64
        dx = _dev.to_dev(x.view(real), self.device)
        n_out = (n + q - 1) // q
        out = _dev.empty((2 * n_out,), torch.float32 if dt == _lib.DTYPE_C64 else torch.float64, self.device)
        _lib.check(h.lib.tb200_decimate(h.h, dx.data_ptr(), dt, n, int(q), out.data_ptr(), _dev.stream_ptr(self.device)))
        return out.cpu().numpy().view(np.complex64 if dt == _lib.DTYPE_C64 else np.complex128)
