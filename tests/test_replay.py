"""Archive replay (config 5) on one GPU: overlapping segments + de-duplication recover every burst once."""
import numpy as np
import pytest

from tetraear_b200 import etsi

pytestmark = pytest.mark.gpu


def test_overlapping_segments_find_every_burst_once():
    import __graft_entry__ as entry
    entry.build()
    from tetraear_b200.replay import ArchiveReplay
    M, fs = 16, 16 * 25000.0
    n_bursts = 30
    truth = {}
    wide = None
    for k, amp in ((2, 1.0), (9, 0.5)):
        bits, t1, init = etsi.make_etsi_stream(n_bursts, seed=40 + k)
        f = k * 25000.0 if k < M // 2 else (k - M) * 25000.0
        x = etsi.modulate_dqpsk(bits, fs, freq=f, amplitude=amp, t_offset=0.2 * k)
        wide = x if wide is None else wide[:min(len(wide), len(x))] + x[:min(len(wide), len(x))]
        truth[k] = t1
    wide = wide.astype(np.complex64)
    seg_len = len(wide) // 3 + 4000
    rp = ArchiveReplay(M, seg_len)
    segs = rp.segments(len(wide))
    assert len(segs) >= 3 and all(b0 < a1 for (a0, a1), (b0, b1) in zip(segs, segs[1:]))      # really overlapping
    # two "ranks" on one GPU, then the same merge the gather does
    from tetraear_b200 import shard
    recs = rp.replay(wide, init, rank=0, world_size=2) + rp.replay(wide, init, rank=1, world_size=2)
    merged = shard.dedup_frames(sorted(recs, key=lambda r: (r[0], r[1])))
    for k in truth:
        mine = [r for r in merged if r[0] == k]
        slots = [r[1] for r in mine]
        assert len(slots) == len(set(slots))                      # de-duplicated
        assert len(mine) >= n_bursts - 4, (k, len(mine))           # every burst (minus loop lock-in at segment starts)
        for _, slot, (ok, blocks) in mine:
            if ok.all() and 0 <= slot < n_bursts:
                assert np.array_equal(blocks, truth[k][slot]), (k, slot)
    assert not [r for r in merged if r[0] not in truth]           # idle channels stay silent
    assert len(recs) > len(merged)                                # the overlap produced duplicates that were dropped
