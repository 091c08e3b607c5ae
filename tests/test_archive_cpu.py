"""Host logic of the archive replay (no GPU): SC16 file format, overlapping chunk list, slot keys, checkpoint lines."""
import json

import numpy as np

from tetraear_b200 import archive, shard, synth


def test_sc16_archive_round_trip_matches_capture_conversion(tmp_path):
    x = synth.make_capture(5000, 130, seed=3, amplitude=0.4)
    p = tmp_path / "cap.sc16"
    raw = archive.Sc16Archive.write(str(p), x)
    arc = archive.Sc16Archive(str(p))
    assert arc.n_samples == 5000 and np.array_equal(arc.read(0, 5000), raw)
    # BladeRFCapture.read_samples (signal/capture.py:255-269): complex(I, Q) as complex64, then / 32768.0
    want = np.array([complex(int(i), int(q)) for i, q in raw[:100]], dtype=np.complex64) / 32768.0
    assert np.array_equal(arc.to_complex64(0, 100), want.astype(np.complex64))
    assert np.array_equal(arc.read(4990, 100), raw[4990:])          # short read at the end of the file


def test_chunk_list_and_slots():
    burst = 255 * 130
    st = archive.chunk_starts(24_000_000, 131072, 0)
    assert len(st) == 183 and st[1] - st[0] == 131072
    so = archive.chunk_starts(24_000_000, 131072, burst)
    assert so[1] - so[0] == 131072 - burst and so[-1] + 131072 <= 24_000_000
    assert archive.chunk_starts(1000, 131072) == []
    # a burst found in two overlapping chunks lands on the same slot although its in-chunk position differs
    abs_bits = 2 * 255 * 7 + 3                                        # burst 7 of the capture, found 3 bits late
    c0, c1 = so[0], so[1]
    pos0 = abs_bits - round(c0 / 130 * 2)
    pos1 = abs_bits - round(c1 / 130 * 2)
    assert archive.slot_of(c0, pos0, 130) == archive.slot_of(c1, pos1 + 1, 130) == 7
    recs = [(0, archive.slot_of(c0, pos0, 130), "a"), (0, archive.slot_of(c1, pos1, 130), "b")]
    assert shard.dedup_frames(recs) == [(0, 7, "a")]


def test_ranks_take_contiguous_ranges_of_the_chunk_list():
    n = len(archive.chunk_starts(24_000_000, 131072, 255 * 130))
    seen = []
    for r in range(8):
        mine = [s for _, s in shard.partition(1, n, 8, r)]
        assert mine == list(range(mine[0], mine[0] + len(mine))) if mine else True
        seen += mine
    assert seen == list(range(n))
