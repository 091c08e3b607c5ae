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
                 keep_symbols=False, n_lanes=1):
        self.sample_rate = float(sample_rate)
        self.chunk_len = int(chunk_len)
        self.max_chunks = int(max_chunks)
        self.device = device
        self.dev = torch.device(f"cuda:{device}")
        self.mode = mode
        self.plan = _lib.plan(self.sample_rate, self.chunk_len)
        # n_lanes > 1: the batch is cut into n_lanes contiguous sub-batches, each with its own handle
        # (workspace) and CUDA stream, so the latency-bound tail kernels of one sub-batch (edge fix,
        # timing pick, sync search) overlap the bandwidth/FMA-bound kernels of the next.
        self.n_lanes = max(1, int(n_lanes))
        per = (self.max_chunks + self.n_lanes - 1) // self.n_lanes
        self.lane_chunks = per
        m = {"exact": _lib.MODE_EXACT, "fast": _lib.MODE_FAST, "fast32": _lib.MODE_FAST, "fast-tc": _lib.MODE_FAST,
             "fast32-tc": _lib.MODE_FAST, "fast-tc4": _lib.MODE_FAST, "fast32-tc4": _lib.MODE_FAST}[mode]
        # "-tc" / "-tc4": tensor-core decimator (csrc/dsp_tc.cu), streaming teams / warp-specialised pipeline
        self.handles = [_lib.Handle(sample_rate=self.sample_rate, max_chunks=per, max_chunk_len=self.chunk_len, mode=m,
                                    device=device) for _ in range(self.n_lanes)]
        for hh in self.handles:
            if mode.startswith("fast32"):
                hh.set_option(_lib.OPT_CHAN_F32, 1)
            if mode.endswith("-tc") or mode.endswith("-tc4"):
                hh.set_option(_lib.OPT_DECIM_TC, 2 if mode.endswith("4") else 1)
        self.h = self.handles[0]
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(self.n_lanes)]
        # the sync search / burst parse of a sub-batch is latency-bound (one small CTA per chunk): it runs on a
        # HIGH-priority stream so that its CTAs take the SM slots the bulk kernels of the other lanes free up,
        # instead of queueing behind them; the dibit rows are double-buffered so that the next batch's
        # demodulation need not wait for this batch's decode
        self.tail_streams = [torch.cuda.Stream(device=self.dev, priority=-1) for _ in range(self.n_lanes)]
        self._pending = []
        self._dec_done = [[None, None] for _ in range(self.n_lanes)]
        self._flip = 0
        C = self.max_chunks
        self.stride = int(self.plan.max_symbols) + 8
        self.max_frames = (2 * self.stride) // 250 + 2
        self._dibits2 = torch.zeros((2, C, self.stride), dtype=torch.uint8, device=self.dev)
        self.dibits = self._dibits2[0]
        self.symbols = torch.zeros((C, self.stride, 2), dtype=torch.float64, device=self.dev) if keep_symbols else None
        self._n_dibits2 = torch.zeros((2, C), dtype=torch.int32, device=self.dev)
        self.n_dibits = self._n_dibits2[0]
        self.phase = torch.zeros((C,), dtype=torch.int32, device=self.dev)
        # frame records and their per-chunk counts live in ONE allocation (`records`): a multi-GPU job gathers a step's
        # results with a single collective on it
        fbytes = C * self.max_frames * _lib.FRAME_DTYPE.itemsize
        self.records = torch.zeros((fbytes + 4 * C,), dtype=torch.uint8, device=self.dev)
        self.frames = self.records[:fbytes].view(C, self.max_frames, _lib.FRAME_DTYPE.itemsize)
        self.n_frames = self.records[fbytes:].view(torch.int32)
        self.stats = torch.zeros((3,), dtype=torch.int64, device=self.dev)
        self._staging = None

    # ---------------------------------------------------------------- device-resident path
    def run_device(self, iq, freq_off=None, decode=True, join=True):
        """iq: CUDA tensor, float32 [n_chunks, chunk_len, 2] (interleaved complex64) or int16 (SC16Q11).
        Asynchronous.  With join=False the current stream does NOT wait for the lanes: consecutive batches then
        pipeline through the lane streams (each lane processes its sub-batches in order); call join() before
        reading the outputs or timing."""
        n_chunks = iq.shape[0]
        assert iq.is_cuda and iq.dtype in (torch.float32, torch.int16) and iq.shape[1] == self.chunk_len and iq.is_contiguous()
        main = torch.cuda.current_stream(self.dev)
        if self.n_lanes == 1 or n_chunks <= min(self.n_lanes, self.lane_chunks):
            self.join()                         # handle 0 and the output rows may still be in use by unjoined lane work
            self._run_slice(self.handles[0], main.cuda_stream, iq, freq_off, 0, n_chunks, decode)
            return n_chunks
        per = (n_chunks + self.n_lanes - 1) // self.n_lanes
        fork = torch.cuda.Event()
        fork.record(main)                       # the batch (input, previous consumers) is ready at this point of `main`
        buf = self._flip
        self._flip ^= 1
        self.dibits, self.n_dibits = self._dibits2[buf], self._n_dibits2[buf]     # where this batch's rows go
        prev_done = [d[buf] for d in self._dec_done]            # decodes of EARLIER calls (not the ones launched below)
        for li in range(self.n_lanes):
            lo, hi = li * per, min(n_chunks, (li + 1) * per)
            if lo >= hi:
                break
            st = self.streams[li]
            st.wait_event(fork)
            if not decode:
                self._run_slice(self.handles[li], st.cuda_stream, iq, freq_off, lo, hi, False)
                ev = torch.cuda.Event()
                ev.record(st)
                self._pending.append(ev)
                continue
            for done_prev in prev_done:                         # every decode that last read this dibit buffer (the
                if done_prev is not None:                       # rows of a lane move when n_chunks changes between calls)
                    st.wait_event(done_prev)
            self._run_slice(self.handles[li], st.cuda_stream, iq, freq_off, lo, hi, False)
            ev = torch.cuda.Event()
            ev.record(st)
            tl = self.tail_streams[li]
            tl.wait_event(ev)
            self._decode_slice(self.handles[li], tl.cuda_stream, lo, hi)
            done = torch.cuda.Event()
            done.record(tl)
            self._dec_done[li][buf] = done
            self._pending.append(done)
        if join:
            self.join()
        return n_chunks

    # ---------------------------------------------------------------- CUDA graph of one device-resident step
    def capture(self, iq, freq_off=None, decode=True):
        """Capture one whole step (every lane's kernels on its streams, the decode on the high-priority streams,
        the fork / join dependencies) into a CUDA graph for THIS input tensor (fixed device pointer — e.g. the
        staging buffer a capture thread refills).  Returns the graph; ``graph.replay()`` re-runs the step with one
        launch from the host (~10 us instead of ~40 us per lane of Python and C-ABI calls), on the current stream.
        Two pipelines captured on the same input and replayed on two streams overlap consecutive batches."""
        self.run_device(iq, freq_off, decode)            # warm-up outside the capture: attributes, caches, workspaces
        torch.cuda.synchronize(self.dev)
        self._dec_done = [[None, None] for _ in range(self.n_lanes)]    # no dependency on events from outside the graph
        self._flip = 0
        g = torch.cuda.CUDAGraph()
        cs = torch.cuda.Stream(device=self.dev)
        with torch.cuda.graph(g, stream=cs):
            self.run_device(iq, freq_off, decode, join=True)
        self._dec_done = [[None, None] for _ in range(self.n_lanes)]
        self._captured_input = iq                        # keep the tensor (and its pointer) alive
        return g

    def join(self):
        """Make the current stream wait for every lane's outstanding work."""
        main = torch.cuda.current_stream(self.dev)
        for ev in self._pending:
            main.wait_event(ev)
        self._pending = []

    def _run_slice(self, handle, stream, iq, freq_off, lo, hi, decode):
        lib, h = handle.lib, handle.h
        n = hi - lo
        sc16 = iq.dtype == torch.int16
        esz = self.chunk_len * (4 if sc16 else 8)
        sym_ptr = self.symbols.data_ptr() + lo * self.stride * 16 if self.symbols is not None else None
        if decode:
            # process + decode as one call: the tail (timing pick, slicer, sync search, frame records) is one launch
            fsz = self.max_frames * _lib.FRAME_DTYPE.itemsize
            _lib.check(lib.tb200_process_decode_batch(
                h, iq.data_ptr() + lo * esz, _lib.DTYPE_SC16 if sc16 else _lib.DTYPE_C64, n, self.chunk_len, self.chunk_len,
                freq_off.data_ptr() + lo * 8 if freq_off is not None else None, self.dibits.data_ptr() + lo * self.stride,
                sym_ptr, self.stride, self.n_dibits.data_ptr() + lo * 4, self.phase.data_ptr() + lo * 4,
                self.frames.data_ptr() + lo * fsz, self.max_frames, self.n_frames.data_ptr() + lo * 4,
                self.stats.data_ptr(), stream))
            return
        _lib.check(lib.tb200_process_batch(
            h, iq.data_ptr() + lo * esz, _lib.DTYPE_SC16 if sc16 else _lib.DTYPE_C64, n, self.chunk_len, self.chunk_len,
            freq_off.data_ptr() + lo * 8 if freq_off is not None else None, self.dibits.data_ptr() + lo * self.stride,
            sym_ptr, self.stride, self.n_dibits.data_ptr() + lo * 4, self.phase.data_ptr() + lo * 4, stream))

    def _decode_slice(self, handle, stream, lo, hi):
        fsz = self.max_frames * _lib.FRAME_DTYPE.itemsize
        _lib.check(handle.lib.tb200_decode_batch(handle.h, self.dibits.data_ptr() + lo * self.stride, hi - lo, self.stride,
                                                 self.n_dibits.data_ptr() + lo * 4, self.frames.data_ptr() + lo * fsz,
                                                 self.max_frames, self.n_frames.data_ptr() + lo * 4,
                                                 self.stats.data_ptr(), stream))

    # ---------------------------------------------------------------- host-buffer path (end to end)
    def run_host(self, iq_host, freq_off=None, decode=True):
        """iq_host: pinned (or pageable) CPU tensor / numpy complex64 [n_chunks, chunk_len].
        H2D -> kernels -> D2H of (n_dibits, n_frames, stats, frame records).  Synchronous.
        Returns (n_frames per chunk, frame records, stats) as numpy."""
        if isinstance(iq_host, np.ndarray):
            if iq_host.dtype == np.int16:              # SC16Q11: [n_chunks, chunk_len, 2] int16, converted on the device
                iq_host = torch.from_numpy(np.ascontiguousarray(iq_host).reshape(iq_host.shape[0], -1, 2))
            else:
                iq_host = torch.from_numpy(iq_host.view(np.float32).reshape(iq_host.shape[0], -1, 2))
        n_chunks = iq_host.shape[0]
        if self._staging is None or self._staging.shape[0] < n_chunks or self._staging.dtype != iq_host.dtype:
            self._staging = torch.empty((self.max_chunks, self.chunk_len, 2), dtype=iq_host.dtype, device=self.dev)
            self._host_nfr = torch.empty((self.max_chunks,), dtype=torch.int32).pin_memory()
            self._host_stats = torch.empty((3,), dtype=torch.int64).pin_memory()
            self._host_frames = torch.empty(self.frames.shape, dtype=torch.uint8).pin_memory()
            self._host_ndib = torch.empty((self.max_chunks,), dtype=torch.int32).pin_memory()
        dev_iq = self._staging[:n_chunks]
        f = None
        if freq_off is not None:
            f = torch.as_tensor(np.asarray(freq_off, dtype=np.float64)).to(self.dev, non_blocking=True)
        self.stats.zero_()
        main = torch.cuda.current_stream(self.dev)
        lanes = 1 if n_chunks <= min(self.n_lanes, self.lane_chunks) else self.n_lanes
        per = (n_chunks + lanes - 1) // lanes
        fork = torch.cuda.Event()
        fork.record(main)
        for li in range(lanes):
            lo, hi = li * per, min(n_chunks, (li + 1) * per)
            if lo >= hi:
                break
            st = self.streams[li] if lanes > 1 else main
            with torch.cuda.stream(st):
                if lanes > 1:
                    st.wait_event(fork)
                # H2D of this lane's chunks, its kernels, D2H of its results: lane i+1's copy overlaps lane i's kernels
                dev_iq[lo:hi].copy_(iq_host[lo:hi], non_blocking=True)
                self._run_slice(self.handles[li], st.cuda_stream, dev_iq, f, lo, hi, decode)
                self._host_ndib[lo:hi].copy_(self.n_dibits[lo:hi], non_blocking=True)
                if decode:
                    self._host_nfr[lo:hi].copy_(self.n_frames[lo:hi], non_blocking=True)
                    self._host_frames[lo:hi].copy_(self.frames[lo:hi], non_blocking=True)
                if lanes > 1:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    main.wait_event(ev)
        if decode:
            self._host_stats.copy_(self.stats, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        if not decode:
            return self._host_ndib[:n_chunks].numpy(), None, None
        return (self._host_nfr[:n_chunks].numpy(),
                self._host_frames[:n_chunks].numpy().view(_lib.FRAME_DTYPE).reshape(n_chunks, self.max_frames),
                self._host_stats.numpy())

    @property
    def launches(self):
        return sum(h.launches for h in self.handles)

    def profile(self, on=True):
        for h in self.handles:
            h.profile(on)

    def profile_read(self):
        tot = {}
        for h in self.handles:
            for k, (ms, calls) in h.profile_read().items():
                a = tot.get(k, (0.0, 0))
                tot[k] = (a[0] + ms, a[1] + calls)
        return tot

    def h2d_bytes(self, n_chunks, sc16=False):
        return n_chunks * self.chunk_len * (4 if sc16 else 8)

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
