"""Sharding of the receive path across the GPUs of one box.

The path has no exchange step: the reference treats every chunk it reads as stateless
(/root/reference/tetraear/ui/modern.py:1915-1919), so a work item is one (channel, segment)
pair and ranks never talk during the data path.  The only communication is the final host
gather of burst records (a few KB per chunk).  Ordering constraint: MAC fragment reassembly
(``TetraProtocolParser.fragment_buffer``) is sequential per channel, so a channel's segments
are kept on ONE rank and in order.
"""
import numpy as np


def work_items(n_channels, n_segments):
    """All (channel, segment) pairs, channel-major."""
    return [(c, s) for c in range(n_channels) for s in range(n_segments)]


def partition(n_channels, n_segments, world_size, rank):
    """Items of `rank`: whole channels round-robin when there are at least as many channels as
    ranks (keeps every channel's segments ordered on one rank); otherwise contiguous segment
    ranges of the few channels."""
    if n_channels >= world_size:
        return [(c, s) for c in range(rank, n_channels, world_size) for s in range(n_segments)]
    items = work_items(n_channels, n_segments)
    per = (len(items) + world_size - 1) // world_size
    return items[rank * per:(rank + 1) * per]


def segment_bounds(n_samples, segment_len, overlap):
    """Overlapping capture segments [start, stop): consecutive starts are segment_len - overlap apart,
    so a burst straddling one segment edge lies whole in the next (overlap >= one burst)."""
    if segment_len <= overlap:
        raise ValueError("segment_len must exceed overlap")
    step = segment_len - overlap
    out = []
    start = 0
    while True:
        stop = min(start + segment_len, n_samples)
        out.append((start, stop))
        if stop >= n_samples:
            break
        start += step
    return out


def dedup_frames(records):
    """records: iterable of (channel, absolute_bit_position, payload).  Frames found twice in the
    overlap of two segments share channel and absolute position: keep the first."""
    seen = set()
    out = []
    for ch, pos, payload in records:
        key = (ch, pos)
        if key in seen:
            continue
        seen.add(key)
        out.append((ch, pos, payload))
    return out


def gather_frames(local_records, group=None):
    """Final host gather over torch.distributed (any backend): rank 0 gets the concatenation
    ordered by (channel, position); other ranks get None."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return sorted(local_records, key=lambda r: (r[0], r[1]))
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(list(local_records), bucket, dst=0, group=group)
    if rank != 0:
        return None
    merged = [r for part in bucket for r in part]
    return sorted(dedup_frames(sorted(merged, key=lambda r: (r[0], r[1]))), key=lambda r: (r[0], r[1]))
