"""C4 in reference mode (BASELINE configs[3]): every channel of a wideband chunk must equal what the
reference computes for it, process(frequency_shift(x, f_k, fs)) then decode — dibits / frames / CRC bit-exact, the
decimated samples and soft symbols within 1e-4 of full scale."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import __graft_entry__ as entry
    entry.build()
    from tetraear_b200 import multichannel, synth
    return multichannel, synth


CARRIERS = [0, 1, -1, 7, -37, 120, 399, -400, 250, -333, 64, -199, 3, 311, -272, 158]


@pytest.mark.parametrize("mode", ["fast32", "fast"])
def test_wideband_channels_match_reference_per_channel(mods, mode):
    multichannel, synth = mods
    fs, M, n = 20e6, 800, 1 << 20
    x = synth.make_wideband(n, CARRIERS, fs, M, 1079, seed=5)
    rx = multichannel.WidebandReceiver(fs, M, n, mode=mode)
    dibits, symbols, frames = rx.process(x)
    dec = rx.decimated.cpu().numpy().view(np.complex64)[..., 0]
    total_crc = 0
    for ch in CARRIERS + [17, -90]:                      # two empty channels as well (noise only)
        k = rx.row_of(ch)
        f_k = ch * fs / M
        assert rx.channel_frequency(k) == f_k
        shifted = oracle.frequency_shift(x, f_k, fs)      # complex128, like the reference
        want_dec = oracle.decimate(shifted, rx.q)
        scale = max(np.abs(want_dec).max(), 0.05)
        err = np.abs(dec[k] - want_dec)
        assert err.max() <= 1e-4 * scale, (ch, err.max() / scale, int(err.argmax()))
        od, osym, _ = oracle.process(shifted, fs, 0.0)
        if ch in CARRIERS:
            assert np.array_equal(dibits[k], od), ch
            assert len(symbols[k]) == len(osym)
            assert np.abs(symbols[k] - osym).max() <= 1e-4, ch                  # of full scale (IQ normalised to [-1, 1))
            ofr, ost = oracle.decode(od)
            kept = [r for r in frames[k] if r["kept"]]
            assert [(int(r["number"]), int(r["type"]), bool(r["crc_ok"])) for r in kept] == \
                   [(f["number"], f["type"], f["burst_crc"]) for f in ofr], ch
            total_crc += ost["crc_pass"]
        else:
            assert len(dibits[k]) == len(od)
            assert np.mean(dibits[k] != od) <= 0.02, ch      # noise: only symbols on a slicer boundary may differ
    assert total_crc >= 10                                   # the carriers really decode
