"""Reference-mode multichannel receive: every 25 kHz channel of a wideband chunk in one pass.

The reference has no channelizer: channel k of a wideband capture is reached only by calling
``SignalProcessor.frequency_shift(x, f_k, fs)`` and then ``SignalProcessor.process`` on the result
(/root/reference/tetraear/signal/processor.py:85-100, :221-273) — M passes over the input, each a
complex128 NCO, a zero-phase cheby1 decimation by q = int(fs / 240000) in float64, the Butterworth channel filter, the
13-phase timing pick and the DQPSK slicer.  ``WidebandReceiver.process(iq)`` returns what those M calls return
(dibits bit-exact, soft symbols within 1e-4 of full scale), computed as

    tb200_refchan_batch     polyphase fold of the decimator's zero-phase response + size-M DFT + scipy's chunk-end terms
                            -> [M, ceil(n/q)] complex64  == decimate(frequency_shift(x, f_k, fs), q) for every k
    tb200_process_batch     the reference path from the decimated rate on (channel filter, timing pick, slicer), M "chunks"
    tb200_decode_batch      TetraDecoder.decode per channel

so that the wideband input is read once.  PyTorch tensors only carry the device buffers.
"""
import numpy as np
import torch

from . import _lib


class WidebandReceiver:
    def __init__(self, sample_rate=20e6, n_channels=800, chunk_len=1 << 20, mode="fast32", device=0, keep_symbols=True):
        self.sample_rate = float(sample_rate)
        self.M = int(n_channels)
        self.chunk_len = int(chunk_len)
        self.device = device
        self.dev = torch.device(f"cuda:{device}")
        self.q = int(self.sample_rate / 240000)                       # processor.py:245-250
        if not (self.sample_rate > 480000 and self.q > 1):
            raise ValueError("WidebandReceiver is for rates the reference decimates (> 480 kHz)")
        self.rate = self.sample_rate / self.q                         # current_rate after decimation, processor.py:255
        self.n_dec = (self.chunk_len + self.q - 1) // self.q
        self.bank = _lib.Handle(sample_rate=self.sample_rate, max_chunks=0, max_chunk_len=0, device=device, max_channels=self.M)
        m = {"exact": _lib.MODE_EXACT, "fast": _lib.MODE_FAST, "fast32": _lib.MODE_FAST}[mode]
        self.rx = _lib.Handle(sample_rate=self.rate, max_chunks=self.M, max_chunk_len=self.n_dec, mode=m, device=device)
        if mode == "fast32":
            self.rx.set_option(_lib.OPT_CHAN_F32, 1)
        self.plan = _lib.plan(self.rate, self.n_dec)
        self.stride = int(self.plan.max_symbols) + 8
        self.max_frames = (2 * self.stride) // 250 + 2
        M = self.M
        self.decimated = torch.zeros((M, self.n_dec, 2), dtype=torch.float32, device=self.dev)
        self.dibits = torch.zeros((M, self.stride), dtype=torch.uint8, device=self.dev)
        self.symbols = torch.zeros((M, self.stride, 2), dtype=torch.float64, device=self.dev) if keep_symbols else None
        self.n_dibits = torch.zeros((M,), dtype=torch.int32, device=self.dev)
        self.phase = torch.zeros((M,), dtype=torch.int32, device=self.dev)
        self.frames = torch.zeros((M, self.max_frames, _lib.FRAME_DTYPE.itemsize), dtype=torch.uint8, device=self.dev)
        self.n_frames = torch.zeros((M,), dtype=torch.int32, device=self.dev)
        self.stats = torch.zeros((3,), dtype=torch.int64, device=self.dev)

    def channel_frequency(self, k):
        """Centre frequency (Hz, relative to the capture centre) of output row k: DFT bin order."""
        k = int(k) % self.M
        return (k if k < self.M // 2 else k - self.M) * (self.sample_rate / self.M)

    def row_of(self, channel):
        """Output row of channel number `channel` (centre = channel * fs / M, negative allowed)."""
        return int(channel) % self.M

    def run_device(self, iq, decode=True):
        """iq: CUDA float32 tensor [chunk_len, 2] (interleaved complex64).  Asynchronous on the current stream."""
        assert iq.is_cuda and iq.dtype == torch.float32 and iq.shape[0] == self.chunk_len and iq.is_contiguous()
        st = torch.cuda.current_stream(self.dev).cuda_stream
        _lib.check(self.bank.lib.tb200_refchan_batch(self.bank.h, iq.data_ptr(), self.chunk_len, self.M, self.q, self.sample_rate,
                                                     self.decimated.data_ptr(), self.n_dec, st))
        _lib.check(self.rx.lib.tb200_process_batch(
            self.rx.h, self.decimated.data_ptr(), _lib.DTYPE_C64, self.M, self.n_dec, self.n_dec, None, self.dibits.data_ptr(),
            self.symbols.data_ptr() if self.symbols is not None else None, self.stride, self.n_dibits.data_ptr(),
            self.phase.data_ptr(), st))
        if decode:
            _lib.check(self.rx.lib.tb200_decode_batch(self.rx.h, self.dibits.data_ptr(), self.M, self.stride, self.n_dibits.data_ptr(),
                                                      self.frames.data_ptr(), self.max_frames, self.n_frames.data_ptr(),
                                                      self.stats.data_ptr(), st))

    def process(self, iq, decode=True):
        """iq: numpy complex64 [chunk_len].  Returns (dibits per row, symbols per row, frame records per row)."""
        x = np.ascontiguousarray(iq, dtype=np.complex64)
        d = torch.from_numpy(x.view(np.float32).reshape(-1, 2)).to(self.dev)
        self.stats.zero_()
        self.run_device(d, decode)
        torch.cuda.synchronize(self.dev)
        nd = self.n_dibits.cpu().numpy()
        ph = self.phase.cpu().numpy()
        dib = self.dibits.cpu().numpy()
        sym = self.symbols.cpu().numpy().view(np.complex128)[..., 0] if self.symbols is not None else None
        sps = int(self.plan.sps)
        out_d = [dib[k, :nd[k]] for k in range(self.M)]
        out_s = [sym[k, :max((self.n_dec - int(ph[k])) // sps, 0)] for k in range(self.M)] if sym is not None else None
        fr = None
        if decode:
            nfr = self.n_frames.cpu().numpy()
            rec = self.frames.cpu().numpy().view(_lib.FRAME_DTYPE).reshape(self.M, self.max_frames)
            fr = [rec[k, :nfr[k]] for k in range(self.M)]
        return out_d, out_s, fr

    @property
    def launches(self):
        return self.bank.launches + self.rx.launches
