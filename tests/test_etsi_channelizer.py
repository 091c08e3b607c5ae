"""`etsi` mode polyphase channelizer (no reference implementation: parity unpinned).  Checked
against the defining sum evaluated in numpy, then end to end through the etsi receiver."""
import numpy as np
import pytest

from tetraear_b200 import etsi

pytestmark = pytest.mark.gpu


def _direct(x, M, frames, channels):
    """y_k[m] = sum_n x[n] h[m D - n] exp(-j 2 pi k n / M) straight from the definition (float64)."""
    h = etsi.channelizer_prototype(M)
    D = M // 2
    out = np.zeros((len(channels), len(frames)), np.complex128)
    n = np.arange(len(x))
    for ci, k in enumerate(channels):
        mixed = x.astype(np.complex128) * np.exp(-2j * np.pi * k * n / M)
        for fi, m in enumerate(frames):
            idx = m * D - np.arange(len(h))
            ok = (idx >= 0) & (idx < len(x))
            out[ci, fi] = np.sum(mixed[idx[ok]] * h[ok])
    return out


@pytest.mark.parametrize("M", [16, 50, 800])
def test_channelizer_matches_definition(M):
    import __graft_entry__ as entry
    entry.build()
    rng = np.random.default_rng(M)
    n = 40 * M + 123
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    ch = etsi.Channelizer(M)
    y = ch.run(x)
    assert y.shape == (M, ch.n_frames(n))
    frames = [0, 1, 2, 13, 14, 37, y.shape[1] - 1]
    channels = [0, 1, 3, M // 2 - 1, M // 2, M - 1]
    ref = _direct(x, M, frames, channels)
    got = y[np.ix_(channels, frames)]
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 1e-4 * scale, np.abs(got - ref).max() / scale


def test_wideband_to_bursts_end_to_end():
    """16 x 25 kHz: three carriers of different power -> channelizer -> etsi receiver -> decoded blocks."""
    import __graft_entry__ as entry
    entry.build()
    M, fs = 16, 16 * 25000.0
    n_bursts = 16
    carriers = {1: 1.0, 5: 0.3, 13: 0.6}                  # channel index -> amplitude (13 = -3 x 25 kHz)
    truth, init = {}, None
    wide = None
    for k, amp in carriers.items():
        bits, t1, init = etsi.make_etsi_stream(n_bursts, seed=100 + k)
        f = k * 25000.0 if k < M // 2 else (k - M) * 25000.0
        x = etsi.modulate_dqpsk(bits, fs, freq=f, amplitude=amp, t_offset=0.13 * k)
        wide = x if wide is None else wide[:min(len(wide), len(x))] + x[:min(len(wide), len(x))]
        truth[k] = t1
    rng = np.random.default_rng(5)
    wide = (wide + 0.01 * (rng.standard_normal(len(wide)) + 1j * rng.standard_normal(len(wide)))).astype(np.complex64)
    y = etsi.Channelizer(M).run(wide)
    rx = etsi.EtsiReceiver(50000.0, max_streams=M, max_samples=y.shape[1])
    res = rx.run(y, init)
    for k in range(M):
        good = [b for b in res[k] if b["crc_ok"].all()]
        if k in carriers:
            assert len(good) >= n_bursts - 5, (k, len(good))
            for b in good:
                # burst index from its position; the channelizer delays by (6 M - 1)/2 input samples
                matches = [i for i in range(n_bursts) if np.array_equal(b["blocks"], truth[k][i])]
                assert len(matches) == 1, k
        else:
            assert len(good) == 0, (k, len(good))           # idle channels decode nothing
