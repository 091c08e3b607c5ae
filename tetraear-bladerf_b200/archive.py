"""Capture archives and their replay (SURVEY.md 8(f) N2, §5 checkpoint row; BASELINE configs[4]).

File format = the buffer ``BladeRFCapture.read_samples`` gets from ``device.sync_rx``
(/root/reference/tetraear/signal/capture.py:240-257): SC16Q11, interleaved little-endian int16 I, Q, four bytes per
sample, no header.  ``Sc16Archive`` memory-maps such a file; the conversion ``complex64(I, Q) / 32768`` (capture.py:255-269)
runs on the device, so a replay moves 4 bytes per sample over PCIe.

``CaptureReplay`` replays a single-carrier archive through the reference-mode path in work units the reference's own
receive loop defines — stateless chunks of ``chunk_len`` samples (ui/modern.py:1915-1919) — with two additions the
reference lacks: consecutive chunks may OVERLAP by one burst so that a burst straddling a chunk edge is whole in the
next chunk (its two copies land on the same absolute slot and the second is dropped), and the replay is RESUMABLE:
after every batch the index of the next batch and the frames found so far go to a JSON-lines checkpoint, and a
restarted replay continues behind the last completed batch.  Chunks are independent, so ranks take contiguous ranges
of the chunk list (``shard.partition``) and meet only in the final gather (``shard.gather_frames``).
"""
import json
import os

import numpy as np

from . import _lib, shard


class Sc16Archive:
    """Raw SC16Q11 capture file (int16 I, Q interleaved), memory-mapped read-only."""

    def __init__(self, path):
        self.path = path
        size = os.path.getsize(path)
        if size % 4:
            raise ValueError(f"{path}: size {size} is not a whole number of SC16 samples")
        self.n_samples = size // 4
        self._mm = np.memmap(path, dtype="<i2", mode="r", shape=(self.n_samples, 2)) if self.n_samples else \
            np.zeros((0, 2), "<i2")

    def read(self, start, count):
        """int16 [count, 2] view (no copy); short at the end of the file."""
        return self._mm[start:start + count]

    def to_complex64(self, start=0, count=None):
        """The host conversion of capture.py:255-269 (for tests; the replay converts on the device)."""
        raw = self.read(start, self.n_samples - start if count is None else count)
        return (raw[:, 0].astype(np.float32) + 1j * raw[:, 1].astype(np.float32)).astype(np.complex64) / np.float32(32768.0)

    @staticmethod
    def write(path, iq):
        """Quantise complex IQ in [-1, 1) to SC16Q11 and write it (test / bench helper)."""
        x = np.asarray(iq)
        raw = np.empty((len(x), 2), "<i2")
        raw[:, 0] = np.clip(np.rint(x.real * 32768.0), -32768, 32767)
        raw[:, 1] = np.clip(np.rint(x.imag * 32768.0), -32768, 32767)
        raw.tofile(path)
        return raw


def chunk_starts(n_samples, chunk_len, overlap=0):
    """Start sample of every WHOLE chunk; consecutive starts are chunk_len - overlap apart (the ragged tail is dropped,
    like a capture loop that stops on a chunk boundary)."""
    if chunk_len <= overlap:
        raise ValueError("chunk_len must exceed overlap")
    if n_samples < chunk_len:
        return []
    return list(range(0, n_samples - chunk_len + 1, chunk_len - overlap))


def slot_of(chunk_start, start_pos_bits, samples_per_symbol):
    """Absolute 255-symbol slot index of a frame: the key two copies of one burst share.  chunk_start in samples,
    start_pos in bits from the chunk's first dibit; the timing phase (< one symbol) and the +-1 bit of the sync search
    are far below half a slot."""
    sym = chunk_start / float(samples_per_symbol) + start_pos_bits / 2.0
    return int(round(sym / 255.0))


class CaptureReplay:
    def __init__(self, path, sample_rate=2.4e6, chunk_len=131072, overlap=0, batch_chunks=64, mode="fast32", device=0,
                 rank=0, world_size=1, checkpoint=None):
        from . import pipeline
        self.arc = Sc16Archive(path)
        self.sample_rate, self.chunk_len, self.overlap = float(sample_rate), int(chunk_len), int(overlap)
        self.starts = chunk_starts(self.arc.n_samples, self.chunk_len, self.overlap)
        items = shard.partition(1, len(self.starts), world_size, rank)        # contiguous range of the chunk list
        self.mine = [s for _, s in items]
        self.batch_chunks = int(batch_chunks)
        self.rank, self.world = rank, world_size
        self.checkpoint = checkpoint
        self.pipe = pipeline.ReceivePipeline(self.sample_rate, self.chunk_len, max_chunks=self.batch_chunks, mode=mode,
                                             device=device, n_lanes=2 if self.batch_chunks >= 4 else 1)
        self.sps_in = self.pipe.plan.q * self.pipe.plan.sps if self.pipe.plan.q > 1 else max(self.pipe.plan.sps, 1)
        self.records = []          # (channel 0, slot, (chunk index, number, type, crc_ok, 510 frame bits packed))
        self.next_batch = 0
        self.stats = np.zeros(3, np.int64)
        if checkpoint and os.path.exists(checkpoint):
            self._load_checkpoint()

    # ------------------------------------------------------------------ checkpoint: JSON lines, one per completed batch
    def _load_checkpoint(self):
        with open(self.checkpoint) as f:
            for line in f:
                try:
                    rec = json.loads(line)
                except ValueError:
                    break                                  # torn last line of an interrupted run: redo that batch
                if rec.get("rank") != self.rank or rec.get("world") != self.world or rec.get("batch") != self.next_batch:
                    continue
                self.next_batch += 1
                self.stats += np.asarray(rec["stats"], np.int64)
                for slot, ci, num, typ, crc, bits in rec["frames"]:
                    self.records.append((0, slot, (ci, num, typ, bool(crc), bytes.fromhex(bits))))

    def _append_checkpoint(self, batch, stats, new):
        if not self.checkpoint:
            return
        with open(self.checkpoint, "a") as f:
            f.write(json.dumps({"rank": self.rank, "world": self.world, "batch": batch, "stats": [int(v) for v in stats],
                                "frames": [[slot, p[0], p[1], p[2], int(p[3]), p[4].hex()] for _, slot, p in new]}) + "\n")
            f.flush()
            os.fsync(f.fileno())

    # ------------------------------------------------------------------ replay
    def n_batches(self):
        return (len(self.mine) + self.batch_chunks - 1) // self.batch_chunks

    def run(self, max_batches=None):
        """Process this rank's remaining batches (at most max_batches of them); returns the de-duplicated records of
        the rank so far.  Use shard.gather_frames(replay.records) for the merged list on rank 0."""
        done = 0
        while self.next_batch < self.n_batches() and (max_batches is None or done < max_batches):
            b = self.next_batch
            idx = self.mine[b * self.batch_chunks:(b + 1) * self.batch_chunks]
            raw = np.stack([self.arc.read(self.starts[i], self.chunk_len) for i in idx])      # [n, chunk_len, 2] int16
            nfr, frames, st = self.pipe.run_host(np.ascontiguousarray(raw))
            new = []
            for c, ci in enumerate(idx):
                for r in frames[c, :int(nfr[c])]:
                    if not r["kept"]:
                        continue
                    slot = slot_of(self.starts[ci], int(r["start_pos"]), self.sps_in)
                    new.append((0, slot, (int(ci), int(r["number"]), int(r["type"]), bool(r["crc_ok"]), bytes(r["bits"]))))
            self.records += new
            self.stats += np.asarray(st, np.int64)
            self._append_checkpoint(b, st, new)
            self.next_batch += 1
            done += 1
        return shard.dedup_frames(sorted(self.records, key=lambda r: (r[0], r[1])))
