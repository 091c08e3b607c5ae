"""Archive replay on the GPU (SURVEY.md 8(f) N2 + the checkpoint/resume row of §5): an SC16Q11 file replayed chunk by
chunk equals the reference path applied to the host conversion of the same chunks; an interrupted replay resumes
behind its last completed batch; overlapping chunks recover the bursts that straddle chunk edges; ranks that take
ranges of the chunk list merge to the single-rank result."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu

CH, NCH = 131072, 14


@pytest.fixture(scope="module")
def capture_file(tmp_path_factory):
    import __graft_entry__ as entry
    entry.build()
    from tetraear_b200 import synth, archive
    x = synth.make_capture(NCH * CH + 5000, 130, seed=21, amplitude=0.5)
    p = tmp_path_factory.mktemp("arc") / "capture.sc16"
    archive.Sc16Archive.write(str(p), x)
    return str(p)


def _key(recs):
    return [(c, slot, p[1], p[2], p[3], p[4]) for c, slot, p in recs]


def test_replay_equals_reference_per_chunk_and_resumes(capture_file, tmp_path):
    from tetraear_b200 import archive
    arc = archive.Sc16Archive(capture_file)
    rp = archive.CaptureReplay(capture_file, 2.4e6, CH, overlap=0, batch_chunks=5, mode="exact")
    recs = rp.run()
    assert rp.next_batch == rp.n_batches() == 3 and len(rp.starts) == NCH
    want = []
    tot = [0, 0, 0]
    for ci, s in enumerate(rp.starts):
        d, _, _ = oracle.process(arc.to_complex64(s, CH), 2.4e6, 0.0)      # capture.py:255-269 conversion on the host
        fr, st = oracle.decode(d)
        for f in fr:
            want.append((ci, f["number"], f["type"], f["burst_crc"], np.packbits(np.asarray(f["bits"], np.uint8)).tobytes()))
        for k, key in enumerate(("total_bursts", "crc_pass", "crc_fail")):
            tot[k] += st[key]
    got = sorted((p[0], p[1], p[2], p[3], p[4][:64]) for _, _, p in rp.records)
    assert got == sorted(want) and list(rp.stats) == tot and tot[1] >= 20
    # interrupted after two batches, then a NEW object continues from the checkpoint file
    ck = str(tmp_path / "replay.ckpt")
    a = archive.CaptureReplay(capture_file, 2.4e6, CH, overlap=0, batch_chunks=5, mode="exact", checkpoint=ck)
    a.run(max_batches=2)
    assert a.next_batch == 2
    with open(ck, "a") as f:
        f.write('{"rank": 0, "world": 1, "batch": 2, "sta')                 # torn line of a killed run: ignored
    b = archive.CaptureReplay(capture_file, 2.4e6, CH, overlap=0, batch_chunks=5, mode="exact", checkpoint=ck)
    assert b.next_batch == 2 and len(b.records) == len(a.records)
    assert _key(b.run()) == _key(recs) and list(b.stats) == tot


def test_overlap_recovers_straddling_bursts_and_ranks_merge(capture_file):
    from tetraear_b200 import archive, shard
    burst = 255 * 130
    plain = archive.CaptureReplay(capture_file, 2.4e6, CH, overlap=0, batch_chunks=8, mode="fast32").run()
    ovl = archive.CaptureReplay(capture_file, 2.4e6, CH, overlap=burst, batch_chunks=8, mode="fast32").run()
    slots_plain = {s: p[3] for _, s, p in plain}
    slots_ovl = {s: p[3] for _, s, p in ovl}
    # (almost) nothing is lost: a frame hung on a training-sequence look-alike inside a payload sits at an arbitrary
    # offset, and which look-alikes win the 250-bit skip-ahead of find_sync depends on where the chunk starts
    assert len(set(slots_plain) - set(slots_ovl)) <= max(1, len(slots_plain) // 30)
    total_slots = (NCH * CH) // burst
    assert len(slots_ovl) >= len(slots_plain) + NCH // 2                    # ... and the straddlers are recovered
    assert len(slots_ovl) >= int(0.9 * total_slots)                         # (find_sync itself misses a few: look-alikes)
    assert len(ovl) == len(slots_ovl)                                       # de-duplicated
    # two ranks: contiguous halves of the chunk list, merged and de-duplicated on "rank 0"
    parts = [archive.CaptureReplay(capture_file, 2.4e6, CH, overlap=burst, batch_chunks=8, mode="fast32", rank=r, world_size=2)
             for r in range(2)]
    merged = shard.dedup_frames(sorted(parts[0].run() + parts[1].run(), key=lambda r: (r[0], r[1])))
    assert _key(merged) == _key(ovl)
