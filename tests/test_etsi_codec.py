"""`etsi` mode channel coding: encoder/oracle round trips on CPU, CUDA decoder vs oracle on GPU.
Parity is unpinned by the reference (it has no channel decoder): truth = the encoder."""
import numpy as np
import pytest

from oracle import etsi_oracle
from tetraear_b200 import etsi


def _batch(ch, n, seed, sigma):
    rng = np.random.default_rng(seed)
    K1 = etsi.CHANNELS[ch][0] if isinstance(ch, str) else ch[0]
    t1 = rng.integers(0, 2, (n, K1), dtype=np.uint8)
    init = etsi.scrambling_init(262, 42, 17) if ch != "BSCH" else 3
    t5 = etsi.encode(t1, ch, init) if isinstance(ch, str) else etsi.encode_custom(t1, *ch, scramb_init=init)
    return t1, t5, etsi.to_soft(t5, sigma, rng), init


def _oracle(soft, ch, init):
    return etsi_oracle.decode(soft, ch, init) if isinstance(ch, str) else etsi_oracle.decode_custom(soft, *ch, scramb_init=init)


# block geometries beyond the named logical channels: (K1, K2, K3, a, rate, crc)
CUSTOM = [(124, 144, 432, 103, "1/3", True), (140, 144, 432, 103, "1/3", False), (92, 112, 168, 13, "2/3", True),
          (268, 288, 432, 103, "2/3", True), (272, 292, 432, 103, "292/432", True), (128, 148, 432, 103, "148/432", True)]


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


def test_puncturing_tables():
    """Facts of cl. 8.2.3.1 that do not depend on any implementation: every rate consumes its K3 type-3 bits from distinct
    mother bits inside 4 K2, never the fourth generator; 2/3 keeps {1, 2, 5} of 8, 1/3 keeps {1, 2, 3, 5, 6, 7};
    292/432 (148/432) is 2/3 (1/3) with one further kept position skipped after every 65 (35) bits."""
    for rate, K2, K3 in (("2/3", 144, 216), ("1/3", 144, 432), ("292/432", 292, 432), ("148/432", 148, 432)):
        k = etsi.puncture_index(rate, K3)
        assert np.array_equal(k, etsi_oracle.puncture_map(rate, K3))
        assert len(set(k.tolist())) == K3 and k.max() < 4 * K2 and (np.diff(k) > 0).all() and (k % 4 != 3).all()
    assert etsi.puncture_index("2/3", 6).tolist() == [0, 1, 4, 8, 9, 12]
    assert etsi.puncture_index("1/3", 6).tolist() == [0, 1, 2, 4, 5, 6]
    k = etsi.puncture_index("292/432", 432)
    base = etsi.puncture_index("2/3", 438)
    assert np.array_equal(k, np.delete(base, [65 + 66 * m for m in range(6)]))        # i = j + (j - 1) div 65
    k = etsi.puncture_index("148/432", 432)
    base = etsi.puncture_index("1/3", 444)
    assert np.array_equal(k, np.delete(base, [35 + 36 * m for m in range(12)]))


@pytest.mark.parametrize("ch", ["SCH/HU", "TCH/4.8", "TCH/2.4"] + CUSTOM)
def test_oracle_round_trip_other_rates(ch):
    """Encoder -> oracle for the rates / channels added in round 2 (rate 1/3, 292/432, 148/432, SCH/HU)."""
    rate = etsi.CHANNEL_CODING[ch][0] if isinstance(ch, str) else ch[4]
    t1, t5, soft, init = _batch(ch, 48, 11, 0.0)
    b, ok, metric = _oracle(soft, ch, init)
    assert np.array_equal(b, t1) and ok.all() and (metric == 32 * t5.shape[1]).all()
    sigma = {"2/3": 0.45, "292/432": 0.45, "1/3": 0.7, "148/432": 0.7}[rate]        # the stronger codes take more noise
    t1, t5, soft, init = _batch(ch, 48, 12, sigma)
    b, ok, _ = _oracle(soft, ch, init)
    assert ((soft > 0) != (t5 > 0)).mean() > (0.01 if sigma < 0.5 else 0.06)
    assert (b != t1).mean() < 1e-3


def test_scrambler_and_interleaver_properties():
    s = etsi.scrambling_sequence(3, 432)
    assert np.array_equal(s, etsi_oracle.scrambling_sequence(3, 432)) and 150 < s.sum() < 280
    for K, a in ((216, 101), (432, 103), (120, 11), (168, 13)):
        assert sorted(etsi.interleave_index(K, a)) == list(range(K))           # a permutation


@pytest.mark.gpu
@pytest.mark.parametrize("ch", ["SCH/HU", "TCH/4.8", "TCH/2.4"] + CUSTOM)
def test_cuda_decoder_other_rates_match_oracle(ch):
    """tb200_etsi_decode_batch channels 3-5 and tb200_etsi_decode_custom against the oracle: bits, CRC flag and final
    path metric identical; the general kernel at rate 2/3 equals the specialised one."""
    import __graft_entry__ as entry
    entry.build()
    dec = etsi.ChannelDecoder(ch)
    for n, sigma, seed in ((1, 0.0, 3), (37, 0.0, 4), (130, 0.5, 5), (200, 0.9, 6), (65, 3.0, 7)):
        t1, t5, soft, init = _batch(ch, n, seed, sigma)
        ob, ook, om = _oracle(soft, ch, init)
        b, ok, m = dec.decode(soft, init)
        assert np.array_equal(b, ob), (ch, n, sigma)
        assert np.array_equal(ok, ook) and np.array_equal(m, om)
        if sigma == 0.0:
            assert np.array_equal(b, t1) and ok.all()
    if not isinstance(ch, str) and ch[4] == "2/3" and ch[:4] in ((92, 112, 168, 13), (268, 288, 432, 103)):
        name = "SCH/HU" if ch[0] == 92 else "SCH/F"
        b2, ok2, m2 = etsi.ChannelDecoder(name).decode(soft, init)
        assert np.array_equal(b2, b) and np.array_equal(ok2, ok) and np.array_equal(m2, m)
    with pytest.raises(RuntimeError):
        etsi.ChannelDecoder((124, 144, 216, 100, "2/3", True)).decode(np.zeros((1, 216), np.int8))     # a = 100: not a permutation
    with pytest.raises(RuntimeError):
        etsi.ChannelDecoder((124, 140, 216, 101, "2/3", True)).decode(np.zeros((1, 216), np.int8))     # K2 != K1 + 20


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


def test_aach_code_properties_and_oracle_cpu():
    """The (30, 14) code used for AACH: systematic, minimum distance 8 (so 3 hard errors are always corrected), and the
    brute-force oracle recovers the information word from noisy soft values."""
    G = etsi.rm3014_generator()
    assert G.shape == (14, 30) and np.array_equal(G[:, :14], np.eye(14, dtype=np.uint8))
    c = np.arange(1 << 14)
    info = ((c[:, None] >> np.arange(13, -1, -1)[None, :]) & 1).astype(np.int64)
    wt = ((info @ G.astype(np.int64)) & 1).sum(axis=1)
    assert wt[1:].min() == 8 and np.bincount(wt)[8] == 345
    rng = np.random.default_rng(2)
    b = rng.integers(0, 2, (200, 14), dtype=np.uint8)
    init = etsi.scrambling_init(262, 42, 17)
    t5 = etsi.aach_encode(b, init)
    hard = t5.copy()
    for r in range(200):                                                      # three bit errors per block
        hard[r, rng.choice(30, 3, replace=False)] ^= 1
    d, corr, margin = etsi_oracle.aach_decode(etsi.to_soft(hard), G, init)
    assert np.array_equal(d, b) and (corr == 32 * 24).all() and (margin > 0).all()
    d, _, _ = etsi_oracle.aach_decode(etsi.to_soft(t5, 0.7, rng), G, init)
    assert (d != b).any(axis=1).mean() < 0.05
    w = etsi.generator_words(G)
    assert w.dtype == np.uint32 and [(int(w[k]) >> (29 - j)) & 1 for k in (0, 13) for j in range(30)] == list(G[0]) + list(G[13])


@pytest.mark.gpu
def test_cuda_aach_decoder_matches_oracle():
    """tb200_etsi_aach_decode == the brute-force oracle: information word, correlation and margin, also for a second
    (random, non-systematic) generator — the decoder is generic in the matrix."""
    import __graft_entry__ as entry
    entry.build()
    rng = np.random.default_rng(8)
    gens = [etsi.rm3014_generator()]
    while len(gens) < 2:
        Gr = rng.integers(0, 2, (14, 30), dtype=np.uint8)
        if len({tuple(r) for r in ((np.array([[(c >> k) & 1 for k in range(14)] for c in range(1, 1 << 14, 37)]) @ Gr) & 1)}) > 400:
            gens.append(Gr)
    for G in gens:
        dec = etsi.AachDecoder(G)
        for n, sigma, seed in ((1, 0.0, 1), (5, 0.0, 2), (301, 0.6, 3), (250, 1.5, 4), (64, 4.0, 5)):
            r2 = np.random.default_rng(seed)
            b = r2.integers(0, 2, (n, 14), dtype=np.uint8)
            init = etsi.scrambling_init(1, 2, 3)
            soft = etsi.to_soft(etsi.aach_encode(b, init, G), sigma, r2)
            ob, oc, om = etsi_oracle.aach_decode(soft, G, init)
            d, c, m = dec.decode(soft, init)
            assert np.array_equal(c, oc) and np.array_equal(m, om), (n, sigma)
            unique = om > 0                                                   # a tie has no unique winner to compare
            assert np.array_equal(d[unique], ob[unique])
            tie = ~unique
            if tie.any():                                                     # ties: the smaller word, as in the oracle
                assert np.array_equal(d[tie], ob[tie])
            if sigma == 0.0:
                assert np.array_equal(d, b)
