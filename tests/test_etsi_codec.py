"""`etsi` mode channel coding: encoder/oracle round trips on CPU, CUDA decoder vs oracle on GPU.
Parity is unpinned by the reference (it has no channel decoder): truth = the encoder."""
import numpy as np
import pytest

from oracle import etsi_oracle
from tetraear_b200 import etsi


def _batch(ch, n, seed, sigma):
    rng = np.random.default_rng(seed)
    K1 = etsi.CHANNELS[ch][0]
    t1 = rng.integers(0, 2, (n, K1), dtype=np.uint8)
    init = etsi.scrambling_init(262, 42, 17) if ch != "BSCH" else 3
    t5 = etsi.encode(t1, ch, init)
    return t1, t5, etsi.to_soft(t5, sigma, rng), init


@pytest.mark.parametrize("ch", ["SCH/HD", "SCH/F", "BSCH"])
def test_oracle_round_trip(ch):
    t1, t5, soft, init = _batch(ch, 64, 1, 0.0)
    K1, K2, K3, a = etsi.CHANNELS[ch]
    assert t5.shape == (64, K3) and K3 * 2 == K2 * 3 and K2 == K1 + 20
    b, ok, metric = etsi_oracle.decode(soft, ch, init)
    assert np.array_equal(b, t1) and ok.all() and (metric == 32 * K3).all()
    t1, t5, soft, init = _batch(ch, 64, 2, 0.45)          # ~1.3 % raw bit errors: still error free
    b, ok, _ = etsi_oracle.decode(soft, ch, init)
    assert np.array_equal(b, t1) and ok.all() and ((soft > 0) != (t5 > 0)).any()
    # wrong scrambling code -> CRC must fail
    _, ok_bad, _ = etsi_oracle.decode(soft, ch, init ^ 0x40)
    assert not ok_bad.any()


def test_scrambler_and_interleaver_properties():
    s = etsi.scrambling_sequence(3, 432)
    assert np.array_equal(s, etsi_oracle.scrambling_sequence(3, 432)) and 150 < s.sum() < 280
    for K, a in ((216, 101), (432, 103), (120, 11)):
        assert sorted(etsi.interleave_index(K, a)) == list(range(K))           # a permutation


@pytest.mark.gpu
@pytest.mark.parametrize("ch", ["SCH/HD", "SCH/F", "BSCH"])
def test_cuda_decoder_matches_oracle(ch):
    import __graft_entry__ as entry
    entry.build()
    dec = etsi.ChannelDecoder(ch)
    for n, sigma, seed in ((1, 0.0, 3), (37, 0.0, 4), (500, 0.5, 5), (300, 0.9, 6), (64, 3.0, 7)):
        t1, t5, soft, init = _batch(ch, n, seed, sigma)
        ob, ook, om = etsi_oracle.decode(soft, ch, init)
        b, ok, m = dec.decode(soft, init)
        assert np.array_equal(b, ob), (ch, n, sigma)
        assert np.array_equal(ok, ook) and np.array_equal(m, om)
        if sigma == 0.0:
            assert np.array_equal(b, t1) and ok.all()
        elif sigma <= 0.5:
            assert ok.mean() > 0.9 and np.array_equal(b[ok], t1[ok])       # CRC-pass frames are the sent ones
    # heavy noise must produce CRC failures (the flag is not stuck)
    assert not ook.all()
