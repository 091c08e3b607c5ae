"""Parity at BASELINE.json's FULL sizes: C1 (10 s @2.4 MS/s = 183 chunks), C3 (1 000 000 bursts), C2 (60 s PSD).
The C oracle port is fast enough for C1/C3 in full; C2 is checked on a frame subset plus the EMA recurrence."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import __graft_entry__ as entry
    entry.build()
    from tetraear_b200 import signal, core, pipeline, _lib
    return signal, core, pipeline, _lib


@pytest.fixture(scope="module")
def c1_capture():
    from tetraear_b200 import synth
    n_chunks, L = 183, 131072
    return synth.make_capture(n_chunks * L, 130, seed=1234).reshape(n_chunks, L)


@pytest.fixture(scope="module")
def c1_oracle(c1_capture):
    out = []
    for c in range(c1_capture.shape[0]):
        d, _, _ = oracle.process(c1_capture[c], 2.4e6, 0.0)
        fr, st = oracle.decode(d)
        out.append((d, fr, st))
    return out


@pytest.mark.parametrize("mode", ["exact", "fast", "fast32", "fast-tc", "fast32-tc"])
def test_c1_full_capture(mods, c1_capture, c1_oracle, mode):
    """Every chunk of the 10 s capture: dibits bit-exact, kept frames (number, type, CRC flag, 510 bits) and the
    parser counters identical to the oracle — in all three arithmetic modes."""
    pipeline = mods[2]
    x = c1_capture
    pipe = pipeline.ReceivePipeline(2.4e6, x.shape[1], max_chunks=x.shape[0], mode=mode, n_lanes=2)
    nfr, frames, stats = pipe.run_host(x)
    ndib = pipe.n_dibits.cpu().numpy()
    dib = pipe.dibits.cpu().numpy()
    tot = [0, 0, 0]
    for c, (od, ofr, ost) in enumerate(c1_oracle):
        assert ndib[c] == len(od) and np.array_equal(dib[c, :ndib[c]], od), (mode, c)
        kept = [r for r in frames[c, :nfr[c]] if r["kept"]]
        assert [(int(r["number"]), int(r["type"]), bool(r["crc_ok"])) for r in kept] == \
               [(f["number"], f["type"], f["burst_crc"]) for f in ofr], (mode, c)
        for r, f in zip(kept, ofr):
            assert np.array_equal(np.unpackbits(r["bits"])[:510], f["bits"]), (mode, c)
        for k, key in enumerate(("total_bursts", "crc_pass", "crc_fail")):
            tot[k] += ost[key]
    assert list(stats) == tot and tot[1] >= 500


def test_c3_one_million_bursts(mods):
    core = mods[1]
    from tetraear_b200 import synth
    rows = synth.make_burst_batch(1_000_000, seed=99)
    want = oracle.parse_bursts(rows)
    got = core.TetraProtocolParser().parse_burst_batch(rows)
    assert np.array_equal(got["burst_type"], want["burst_type"])
    assert np.array_equal(got["crc_ok"].astype(np.uint8), np.asarray(want["crc_ok"]).astype(np.uint8))
    # data bits of every burst: compare a checksum per row (sum of bit * (index + 1)) and the first 4096 rows in full
    nd = np.asarray(want["n_data"])
    assert np.array_equal(got["n_data"], nd)
    w = np.arange(1, 511, dtype=np.int32)
    gb, wb = np.asarray(got["data_bits"]), np.asarray(want["data_bits"])
    assert np.array_equal(gb[:4096], wb[:4096])
    for lo in range(0, len(gb), 65536):
        a = (gb[lo:lo + 65536].astype(np.int32) * w).sum(axis=1)
        b = (wb[lo:lo + 65536].astype(np.int32) * w).sum(axis=1)
        assert np.array_equal(a, b), lo


def test_c2_sixty_seconds_psd(mods):
    """60 s @2.4 MS/s, N = 4096 dense frames with the dB-domain EMA: a random subset of frames against the oracle
    (abs <= 1e-3 dB where P > -120 dB, peak bin exact) and the EMA recurrence S_k = 0.85 S_(k-1) + 0.15 P_k on all."""
    import torch
    from tetraear_b200.signal import SpectrumAnalyzer
    fs, N = 2.4e6, 4096
    n = 144_000_000
    nfr = n // N
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    iq = 0.1 * torch.randn((n, 2), generator=g, device="cuda", dtype=torch.float32)
    t = torch.arange(n, device="cuda", dtype=torch.float64) * (2 * np.pi * 7300.0 / fs)
    iq[:, 0] += (0.03 * torch.cos(t)).float(); iq[:, 1] += (0.03 * torch.sin(t)).float()
    del t
    an = SpectrumAnalyzer(fs, N, device=0, precision="f64")
    power = torch.empty((nfr, N), dtype=torch.float32, device="cuda")
    smooth = torch.empty((nfr, N), dtype=torch.float32, device="cuda")
    stats = torch.empty((nfr, 8), dtype=torch.float64, device="cuda")
    an.run_device(iq, nfr, N, denoise=True, power=power, smoothed=smooth, stats=stats)
    torch.cuda.synchronize()
    rng = np.random.default_rng(1)
    for k in sorted(rng.choice(nfr, 48, replace=False)) + [0, nfr - 1]:
        x = iq[k * N:(k + 1) * N].cpu().numpy().view(np.complex64).reshape(-1)
        want = oracle.psd_frame(x, fs, N)
        want = want[1] if isinstance(want, tuple) else want
        got = power[k].cpu().numpy().astype(np.float64)
        m = want > -120
        assert np.abs(got[m] - want[m]).max() <= 1e-3, k
        assert int(np.argmax(got)) == int(np.argmax(want)), k
    # EMA on all frames, evaluated on the device in float64 from the float32 frames the kernel stored
    p64, s64 = power.double(), smooth.double()
    assert torch.equal(smooth[0], power[0])
    err = (s64[1:] - (0.85 * s64[:-1] + 0.15 * p64[1:])).abs().max().item()
    assert err <= 2e-4, err          # float32 storage of S: 120 dB * 2^-24 * a few steps


@pytest.mark.parametrize("precision,n_fft,avg", [("f64", 4096, 8), ("f32", 4096, 8), ("f32", 4096, 3), ("f64", 2048, 4), ("f32", 1024, 5)])
def test_psd_linear_averaging_matches_oracle(mods, precision, n_fft, avg):
    """BASELINE configs[1] "w/ averaging": K-frame linear (Welch) power average before the dB conversion, against
    oracle.psd_welch (the reference's per-frame formula, ui/modern.py:1932-1941, averaged in power; parity unpinned — the
    reference only has the dB-domain EMA) — abs <= 1e-3 dB where P > -120 dB, statistics from the averaged frame, and
    K = 1 unchanged (bit-identical to the un-averaged call)."""
    import torch
    from tetraear_b200.signal import SpectrumAnalyzer
    from tetraear_b200 import synth
    fs = 2.4e6
    nfr = 4 * avg + 2                                           # two input frames are left over and must be ignored
    x = synth.make_capture(nfr * n_fft, 130, seed=9, amplitude=0.1)
    x = (x + 0.03 * np.exp(2j * np.pi * 7300.0 / fs * np.arange(len(x)))).astype(np.complex64)
    d = torch.from_numpy(x.view(np.float32).reshape(-1, 2)).cuda()
    an = SpectrumAnalyzer(fs, n_fft, device=0, precision=precision, average=avg)
    p, s, st = an.run_device(d, nfr, n_fft, denoise=True)
    torch.cuda.synchronize()
    assert p.shape[0] == nfr // avg == 4
    want = oracle.psd_welch(x, fs, n_fft, avg)
    got = p.cpu().numpy().astype(np.float64)
    m = want > -120
    assert np.abs(got[m] - want[m]).max() <= 1e-3
    for k in range(4):
        assert int(np.argmax(got[k])) == int(np.argmax(want[k]))
        ws = oracle.band_stats(want[k], fs, n_fft)
        row = st[k].cpu().numpy()
        assert abs(row[0] - ws["signal_power"]) <= 1e-3 and int(row[2]) == ws["peak_idx"]
    # EMA over the averaged frames
    sm = s.cpu().numpy().astype(np.float64)
    assert np.array_equal(sm[0], got[0].astype(np.float32).astype(np.float64))
    assert np.abs(sm[1:] - (0.85 * sm[:-1] + 0.15 * got[1:])).max() <= 2e-4
    # K = 1 is the plain spectrum
    a1 = SpectrumAnalyzer(fs, n_fft, device=0, precision=precision, average=1)
    a0 = SpectrumAnalyzer(fs, n_fft, device=0, precision=precision)
    p1, _, _ = a1.run_device(d, nfr, n_fft)
    p0, _, _ = a0.run_device(d, nfr, n_fft)
    assert torch.equal(p1, p0)
