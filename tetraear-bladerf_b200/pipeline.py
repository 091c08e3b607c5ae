"""Batched IQ->bursts pipeline: many equal-length chunks per call, everything device resident.

This is the shape the reference's receive loop takes once the per-chunk Python loop
(``CaptureThread.run``, /root/reference/tetraear/ui/modern.py:1901-2089: read 131072 samples ->
``processor.process`` -> ``decoder.decode``) is batched: chunks are stateless in the reference
(SURVEY.md §5), so a capture is an independent [n_chunks, chunk_len] array.  One
``tb200_process_batch`` + one ``tb200_decode_batch`` per batch; PyTorch tensors only carry the
device buffers and the stream.
"""
import numpy as np
import torch

from . import _lib


class ReceivePipeline:
    def __init__(self, sample_rate=2.4e6, chunk_len=131072, max_chunks=184, mode="exact", device=0,
                 keep_symbols=False):
        self.sample_rate = float(sample_rate)
        self.chunk_len = int(chunk_len)
        self.max_chunks = int(max_chunks)
        self.device = device
        self.dev = torch.device(f"cuda:{device}")
        self.mode = mode
        self.plan = _lib.plan(self.sample_rate, self.chunk_len)
        self.h = _lib.Handle(sample_rate=self.sample_rate, max_chunks=self.max_chunks, max_chunk_len=self.chunk_len,
                             mode={"exact": _lib.MODE_EXACT, "fast": _lib.MODE_FAST}[mode], device=device)
        C = self.max_chunks
        self.stride = int(self.plan.max_symbols) + 8
        self.max_frames = (2 * self.stride) // 250 + 2
        self.dibits = torch.zeros((C, self.stride), dtype=torch.uint8, device=self.dev)
        self.symbols = torch.zeros((C, self.stride, 2), dtype=torch.float64, device=self.dev) if keep_symbols else None
        self.n_dibits = torch.zeros((C,), dtype=torch.int32, device=self.dev)
        self.phase = torch.zeros((C,), dtype=torch.int32, device=self.dev)
        self.frames = torch.zeros((C, self.max_frames, _lib.FRAME_DTYPE.itemsize), dtype=torch.uint8, device=self.dev)
        self.n_frames = torch.zeros((C,), dtype=torch.int32, device=self.dev)
        self.stats = torch.zeros((3,), dtype=torch.int64, device=self.dev)
        self._staging = None

    # ---------------------------------------------------------------- device-resident path
    def run_device(self, iq, freq_off=None, decode=True):
        """iq: CUDA tensor, float32 [n_chunks, chunk_len, 2] (interleaved complex64).  Asynchronous."""
        n_chunks = iq.shape[0]
        assert iq.is_cuda and iq.dtype == torch.float32 and iq.shape[1] == self.chunk_len and iq.is_contiguous()
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        lib, h = self.h.lib, self.h.h
        _lib.check(lib.tb200_process_batch(
            h, iq.data_ptr(), _lib.DTYPE_C64, n_chunks, self.chunk_len, self.chunk_len,
            freq_off.data_ptr() if freq_off is not None else None, self.dibits.data_ptr(),
            self.symbols.data_ptr() if self.symbols is not None else None, self.stride,
            self.n_dibits.data_ptr(), self.phase.data_ptr(), stream))
        if decode:
            _lib.check(lib.tb200_decode_batch(h, self.dibits.data_ptr(), n_chunks, self.stride, self.n_dibits.data_ptr(),
                                              self.frames.data_ptr(), self.max_frames, self.n_frames.data_ptr(),
                                              self.stats.data_ptr(), stream))
        return n_chunks

    # ---------------------------------------------------------------- host-buffer path (end to end)
    def run_host(self, iq_host, freq_off=None, decode=True):
        """iq_host: pinned (or pageable) CPU tensor / numpy complex64 [n_chunks, chunk_len].
        H2D -> kernels -> D2H of (n_dibits, n_frames, stats, frame records).  Synchronous.
        Returns (n_frames per chunk, frame records, stats) as numpy."""
        if isinstance(iq_host, np.ndarray):
            iq_host = torch.from_numpy(iq_host.view(np.float32).reshape(iq_host.shape[0], -1, 2))
        n_chunks = iq_host.shape[0]
        if self._staging is None or self._staging.shape[0] < n_chunks:
            self._staging = torch.empty((self.max_chunks, self.chunk_len, 2), dtype=torch.float32, device=self.dev)
            self._host_nfr = torch.empty((self.max_chunks,), dtype=torch.int32).pin_memory()
            self._host_stats = torch.empty((3,), dtype=torch.int64).pin_memory()
            self._host_frames = torch.empty(self.frames.shape, dtype=torch.uint8).pin_memory()
            self._host_ndib = torch.empty((self.max_chunks,), dtype=torch.int32).pin_memory()
        dev_iq = self._staging[:n_chunks]
        dev_iq.copy_(iq_host, non_blocking=True)
        f = None
        if freq_off is not None:
            f = torch.as_tensor(np.asarray(freq_off, dtype=np.float64)).to(self.dev, non_blocking=True)
        self.stats.zero_()
        self.run_device(dev_iq, f, decode=decode)
        self._host_ndib[:n_chunks].copy_(self.n_dibits[:n_chunks], non_blocking=True)
        if decode:
            self._host_nfr[:n_chunks].copy_(self.n_frames[:n_chunks], non_blocking=True)
            self._host_frames[:n_chunks].copy_(self.frames[:n_chunks], non_blocking=True)
            self._host_stats.copy_(self.stats, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        if not decode:
            return self._host_ndib[:n_chunks].numpy(), None, None
        return (self._host_nfr[:n_chunks].numpy(),
                self._host_frames[:n_chunks].numpy().view(_lib.FRAME_DTYPE).reshape(n_chunks, self.max_frames),
                self._host_stats.numpy())

    def h2d_bytes(self, n_chunks):
        return n_chunks * self.chunk_len * 8

    def d2h_bytes(self, n_chunks):
        return n_chunks * (4 + 4 + self.max_frames * _lib.FRAME_DTYPE.itemsize) + 24

    def kept_frames(self, n_frames, frames):
        """Flatten the kept records of a run_host() result: list of (chunk, record)."""
        out = []
        for c in range(len(n_frames)):
            for r in frames[c, : int(n_frames[c])]:
                if r["kept"]:
                    out.append((c, r))
        return out
