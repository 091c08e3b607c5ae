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
    def __init__(self, n_channels, segment_len, device=0, threshold=0.8):
        self.M = int(n_channels)
        self.fs = self.M * etsi.CHANNEL_SPACING
        self.segment_len = int(segment_len)
        self.device = device
        self.threshold = threshold
        self.D = self.M // 2
        self.chan = etsi.Channelizer(self.M, device=device)
        self.nf = self.chan.n_frames(self.segment_len)
        self.rx = etsi.EtsiReceiver(2 * etsi.CHANNEL_SPACING, max_streams=self.M, max_samples=self.nf, device=device)
        self.rows = torch.empty((self.M, self.nf, 2), dtype=torch.float32, device=f"cuda:{device}")
        # overlap = two bursts (255 symbols each at 18 ksym/s) in wideband samples: one for the burst that
        # straddles the edge, one for the timing loop of the next segment to lock before it
        self.burst_samples = int(np.ceil(2 * 255 / 18000.0 * self.fs))

    def segments(self, n_samples):
        return shard.segment_bounds(n_samples, self.segment_len, self.burst_samples)

    def process_segment_device(self, wide, n_samples, scramb_init):
        """wide: CUDA float32 [n_samples, 2] (one segment).  Kernels only, results stay on the device."""
        self.chan.run_device(wide, n_samples, self.rows)
        nf = self.chan.n_frames(n_samples)
        self.rx.run_device(self.rows, nf, scramb_init, self.threshold)
        return nf

    def collect(self, seg_start):
        """Burst records of the last segment: (channel, absolute symbol position, crc_ok[2], blocks[2,124])."""
        nb = self.rx.n_bursts.cpu().numpy()
        pos = self.rx.burst_pos.cpu().numpy()
        ok = self.rx.crc_ok.cpu().numpy().reshape(self.M, self.rx.max_bursts, 2).astype(bool)
        bits = np.unpackbits(self.rx.bits.cpu().numpy(), axis=1)[:, :124].reshape(self.M, self.rx.max_bursts, 2, 124)
        out = []
        sym_per_sample = 18000.0 / self.fs
        for k in np.nonzero(nb)[0]:
            for i in range(int(nb[k])):
                if not ok[k, i].any():
                    continue
                # burst start in symbols since the start of the capture (the timing loop starts sps+16 samples in;
                # +-2 symbols of slack are absorbed by rounding to the 255-symbol slot grid of the channel)
                abs_sym = seg_start * sym_per_sample + pos[k, i] / 2.0
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
