"""Multichannel archive replay (BASELINE config 5): a long wideband capture cut into overlapping
segments, every segment channelized to M x 25 kHz and received in `etsi` mode, bursts de-duplicated
by absolute position and gathered on rank 0.

Sharding: the unit of work is a capture SEGMENT (all channels); rank r takes
segments r, r + world, ...  No collective on the data path — the only communication is the final
host gather (``shard.gather_frames``).  Segments overlap by two bursts so a burst straddling a
segment edge is whole (and after timing lock-in) in the next segment; both copies decode to the same absolute position and
the second is dropped.  The reference has no equivalent (it drops straddling bursts,
/root/reference/tetraear/ui/modern.py:1915-1919, and can only reach one channel at a time).
"""
import numpy as np
import torch

from . import etsi, shard


class ArchiveReplay:
    def __init__(self, n_channels, segment_len, device=0, threshold=0.8, n_blocks=4):
        self.M = int(n_channels)
        self.fs = self.M * etsi.CHANNEL_SPACING
        self.segment_len = int(segment_len)
        self.device = device
        self.threshold = threshold
        self.D = self.M // 2
        self.chan = etsi.Channelizer(self.M, device=device)
        self.nf = self.chan.n_frames(self.segment_len)
        self.rows = torch.empty((self.M, self.nf, 2), dtype=torch.float32, device=f"cuda:{device}")
        # The per-channel timing loop is sequential, so one segment offers only M warps of work.  The frame
        # axis of a segment is therefore cut into n_blocks time blocks (each re-running two bursts of the
        # previous block so its loop is locked and straddling bursts are whole) that are received
        # concurrently on their own CUDA streams: n_blocks x M warps in flight.
        self.ovl_frames = int(np.ceil(2 * 255 / 18000.0 * 2 * etsi.CHANNEL_SPACING))
        self.n_blocks = max(1, int(n_blocks))
        self.block_step = -(-self.nf // self.n_blocks)
        if self.block_step <= self.ovl_frames:
            self.n_blocks, self.block_step = 1, self.nf
        blen = self.block_step + (self.ovl_frames if self.n_blocks > 1 else 0)
        self.rxs = [etsi.EtsiReceiver(2 * etsi.CHANNEL_SPACING, max_streams=self.M, max_samples=blen, device=device)
                    for _ in range(self.n_blocks)]
        self.rx = self.rxs[0]
        self.streams = [torch.cuda.Stream(device=f"cuda:{device}") for _ in range(self.n_blocks - 1)]
        self.block_first = [0] * self.n_blocks
        # overlap = two bursts (255 symbols each at 18 ksym/s) in wideband samples: one for the burst that
        # straddles the edge, one for the timing loop of the next segment to lock before it
        self.burst_samples = int(np.ceil(2 * 255 / 18000.0 * self.fs))

    def segments(self, n_samples):
        return shard.segment_bounds(n_samples, self.segment_len, self.burst_samples)

    def process_segment_device(self, wide, n_samples, scramb_init):
        """wide: CUDA float32 [n_samples, 2] (one segment).  Kernels only, results stay on the device."""
        self.chan.run_device(wide, n_samples, self.rows)
        nf = self.chan.n_frames(n_samples)
        self.receive_rows(nf, scramb_init)
        return nf

    def receive_rows(self, nf, scramb_init):
        """etsi receive of self.rows[:, :nf]: the time blocks run concurrently on side streams."""
        main = torch.cuda.current_stream(self.rows.device)
        fork = torch.cuda.Event()
        fork.record(main)
        row_stride = self.rows.shape[1]
        for b in range(self.n_blocks):
            first = max(0, b * self.block_step - self.ovl_frames)
            last = min(nf, (b + 1) * self.block_step)
            self.block_first[b] = first
            st = main if b == 0 else self.streams[b - 1]
            if last <= first:
                self.rxs[b].n_bursts.zero_()
                continue
            with torch.cuda.stream(st):
                if b:
                    st.wait_event(fork)
                view = self.rows[:, first:last]                      # [M, len, 2] view, row stride = whole row
                self.rxs[b].run_device_strided(view.data_ptr(), self.M, last - first, row_stride, scramb_init, self.threshold)
                if b:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    main.wait_event(ev)

    def collect(self, seg_start):
        """Burst records of the last segment: (channel, absolute symbol position, crc_ok[2], blocks[2,124])."""
        out = []
        sym_per_sample = 18000.0 / self.fs
        sym_per_frame = 18000.0 / (2 * etsi.CHANNEL_SPACING)
        for b, rx in enumerate(self.rxs):
            nb = rx.n_bursts.cpu().numpy()
            if not nb.any():
                continue
            pos = rx.burst_pos.cpu().numpy()
            ok = rx.crc_ok.cpu().numpy().reshape(self.M, rx.max_bursts, 2).astype(bool)
            bits = np.unpackbits(rx.bits.cpu().numpy(), axis=1)[:, :124].reshape(self.M, rx.max_bursts, 2, 124)
            t0 = seg_start * sym_per_sample + self.block_first[b] * sym_per_frame
            for k in np.nonzero(nb)[0]:
                for i in range(int(nb[k])):
                    if not ok[k, i].any():
                        continue
                    # burst start in symbols since the start of the capture; +-2 symbols of slack are absorbed by
                    # rounding to the 255-symbol slot grid of the channel
                    abs_sym = t0 + pos[k, i] / 2.0
                    out.append((int(k), int(round(abs_sym / 255.0)), (ok[k, i].copy(), bits[k, i].copy())))
        return out

    def replay(self, wide_host, scramb_init, rank=0, world_size=1):
        """wide_host: complex64 numpy (or pinned tensor) of the whole capture.  Processes this rank's segments
        and returns its burst records (use shard.gather_frames for the final gather)."""
        n = len(wide_host)
        segs = self.segments(n)
        recs = []
        buf = torch.empty((self.segment_len, 2), dtype=torch.float32, device=f"cuda:{self.device}")
        host = torch.from_numpy(np.ascontiguousarray(wide_host).view(np.float32).reshape(-1, 2)) \
            if isinstance(wide_host, np.ndarray) else wide_host
        for si in range(rank, len(segs), world_size):
            a, b = segs[si]
            buf[: b - a].copy_(host[a:b], non_blocking=True)
            self.process_segment_device(buf, b - a, scramb_init)
            recs += self.collect(a)
        return shard.dedup_frames(sorted(recs, key=lambda r: (r[0], r[1])))
