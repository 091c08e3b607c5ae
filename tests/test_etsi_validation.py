"""The validations SURVEY.md 8(a') lists for the `etsi` stages that have no reference arithmetic (parity unpinned):
RRC matched filter vs numpy.convolve with the same taps, BER vs timing offset, soft bits (sign == hard decision,
reliability monotone in SNR) and training-sequence detect rate vs SNR.  Truth = what was transmitted."""
import ctypes as C

import numpy as np
import pytest

from tetraear_b200 import etsi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as entry
    entry.build()


def _soft_bits(rx, x):
    """demodulate [n_streams, n] complex64 -> (soft [n_streams, 2 * n_sym] float32, counts)."""
    import torch
    from tetraear_b200 import _dev, _lib
    d = _dev.to_dev(np.ascontiguousarray(x, np.complex64).view(np.float32).reshape(x.shape[0], x.shape[1], 2), rx.device)
    rx.symbols = _dev.zeros((rx.max_streams, rx.max_sym, 2), torch.float32, rx.device)
    _lib.check(rx._h.lib.tb200_etsi_demod_batch(rx._h.h, d.data_ptr(), x.shape[0], x.shape[1], x.shape[1], rx.fs,
                                                rx.symbols.data_ptr(), rx.soft.data_ptr(), rx.soft_stride,
                                                rx.n_soft.data_ptr(), _dev.stream_ptr(rx.device)))
    torch.cuda.synchronize()
    return rx.soft.cpu().numpy(), rx.n_soft.cpu().numpy(), rx.symbols.cpu().numpy().view(np.complex64)[..., 0]


def _align(soft, bits, max_lag=40):
    """lag (in bits, even) at which sign(soft) best matches the transmitted bits; returns (lag, bit errors, compared)."""
    hard = (soft > 0).astype(np.uint8)
    best = (0, 1 << 30, 0)
    for lag in range(0, max_lag, 2):
        n = min(len(hard) - lag, len(bits))
        if n < 1000:
            continue
        e = int(np.count_nonzero(hard[lag:lag + n] != bits[:n]))
        if e < best[1]:
            best = (lag, e, n)
    return best


@pytest.mark.parametrize("fs", [72000.0, 50000.0, 36000.0])
def test_matched_filter_equals_numpy_convolve_with_the_same_taps(fs):
    """Open loop (timing gains 0): symbol k is the interpolating RRC filter applied at t = 16 + (k+1) sps; the same
    32 taps of the same fractional phase through numpy.convolve must give the same value."""
    from tetraear_b200 import _lib
    rng = np.random.default_rng(3)
    n = 6000
    x = (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))).astype(np.complex64)
    rx = etsi.EtsiReceiver(fs, max_streams=2, max_samples=n)
    rx._h.set_option(_lib.OPT_ETSI_OPEN_LOOP, 1)
    soft, cnt, sym = _soft_bits(rx, x)
    taps = np.zeros((32, 32), np.float32)
    _lib.check(rx._h.lib.tb200_debug_read(rx._h.h, 4, taps.ctypes.data, taps.nbytes))
    sps = np.float32(fs / 18000.0)
    tau = np.float32(16.0) + sps
    nsym = int(cnt[0]) // 2 + 1
    conv = [[np.convolve(x[s].astype(np.complex128), taps[p][::-1].astype(np.float64)) for p in range(32)] for s in range(2)]
    worst = 0.0
    for k in range(nsym):
        i_on = int(np.floor(tau))
        ph = min(max(int((tau - np.float32(i_on)) * 32), 0), 31)
        for s in range(2):
            # sum_j taps[ph][j] x[i_on - 15 + j]  ==  (x * reversed taps)[i_on + 16]
            want = conv[s][ph][i_on + 16]
            worst = max(worst, abs(sym[s, k] - want))
        tau = np.float32(tau + sps)
    assert nsym > 1000 and worst <= 2e-5 * np.abs(sym[:, :nsym]).max()
    # the table itself: every phase is the RRC(0.35) pulse sampled at the fractional delay, unit energy ordering
    assert np.all(np.argmax(np.abs(taps), axis=1) >= 14) and np.all(np.argmax(np.abs(taps), axis=1) <= 17)


def test_ber_is_flat_over_the_timing_offset_sweep():
    """Closed-loop timing: after lock-in the bit error count must not depend on where the sampling grid falls."""
    fs, sigma = 72000.0, 0.12                      # ~15 dB Es/N0 after the matched filter: a few errors per 10^4 bits
    offs = np.linspace(0.0, 0.875, 8)
    bits, _, _ = etsi.make_etsi_stream(40, seed=5)
    xs = [etsi.modulate_dqpsk(bits, fs, sigma=sigma, t_offset=float(o), seed=int(100 * o)) for o in offs]
    n = min(len(v) for v in xs)
    rx = etsi.EtsiReceiver(fs, max_streams=len(offs), max_samples=n)
    soft, cnt, _ = _soft_bits(rx, np.stack([v[:n] for v in xs]))
    bers = []
    for s in range(len(offs)):
        lock = 2 * 255 * 2                                           # skip two bursts of lock-in
        lag, err, ncmp = _align(soft[s, lock:cnt[s]], bits[lock:])
        bers.append(err / ncmp)
    assert max(bers) <= 5e-3, bers
    assert max(bers) - min(bers) <= 3e-3, bers                       # flat: the loop removes the offset
    # and at high SNR there are no errors at any offset
    xs = [etsi.modulate_dqpsk(bits, fs, sigma=0.01, t_offset=float(o), seed=3) for o in offs]
    soft, cnt, _ = _soft_bits(rx, np.stack([v[:n] for v in xs]))
    for s in range(len(offs)):
        lag, err, ncmp = _align(soft[s, 1020:cnt[s]], bits[1020:])
        assert err == 0 and ncmp > 15000, (offs[s], err)


def test_soft_bits_sign_is_the_hard_decision_and_reliability_grows_with_snr():
    fs = 50000.0
    bits, _, _ = etsi.make_etsi_stream(30, seed=9)
    sigmas = [0.4, 0.28, 0.2, 0.14, 0.1, 0.05]
    xs = [etsi.modulate_dqpsk(bits, fs, sigma=s, t_offset=0.3, seed=7) for s in sigmas]
    n = min(len(v) for v in xs)
    rx = etsi.EtsiReceiver(fs, max_streams=len(sigmas), max_samples=n)
    soft, cnt, sym = _soft_bits(rx, np.stack([v[:n] for v in xs]))
    rel, ber = [], []
    for s in range(len(sigmas)):
        # sign(soft) IS the differential hard decision of the kernel's own symbols: b1 = [Im d < 0], b2 = [Re d < 0]
        k = cnt[s] // 2
        d = sym[s, 1:k + 1] * np.conj(sym[s, :k])
        hard = np.stack([d.imag < 0, d.real < 0], 1).reshape(-1)
        sv = soft[s, :2 * k]
        clear = np.abs(sv) > 1e-3
        assert np.array_equal((sv > 0)[clear], hard[clear])
        lag, err, ncmp = _align(sv[1020:], bits[1020:])
        tx = 2.0 * bits[1020:1020 + ncmp] - 1.0
        llr = sv[1020 + lag:1020 + lag + ncmp] * tx                  # > 0 where the soft bit agrees with what was sent
        rel.append(float(llr.mean() / llr.std()))
        ber.append(err / ncmp)
    assert all(b >= a for a, b in zip(rel, rel[1:])), rel            # reliability monotone in SNR
    assert all(b <= a for a, b in zip(ber, ber[1:])) and ber[-1] == 0 and ber[0] > 0.01, ber


def test_training_sequence_detect_rate_vs_snr():
    fs, n_bursts = 50000.0, 40
    sigmas = [0.7, 0.5, 0.35, 0.25, 0.15, 0.05]
    streams = []
    for i, s in enumerate(sigmas):
        bits, t1, init = etsi.make_etsi_stream(n_bursts, seed=11)
        streams.append(etsi.modulate_dqpsk(bits, fs, sigma=s, t_offset=0.2, seed=20 + i))
    n = min(len(v) for v in streams)
    rx = etsi.EtsiReceiver(fs, max_streams=len(sigmas), max_samples=n)
    res = rx.run(np.stack([v[:n] for v in streams]), init)
    rate, false_pos = [], []
    for s in range(len(sigmas)):
        hits = {int(round(b["start"] / 510.0)) for b in res[s] if abs(b["start"] - 510 * round(b["start"] / 510.0)) <= 4}
        rate.append(len(hits) / n_bursts)
        false_pos.append(sum(1 for b in res[s] if abs(b["start"] - 510 * round(b["start"] / 510.0)) > 4))
    assert rate[-1] >= 0.95 and rate[-2] >= 0.95 and rate[-3] >= 0.9, rate       # >= 10 dB: (nearly) every burst
    assert all(b >= a - 0.05 for a, b in zip(rate, rate[1:])), rate              # monotone (within one burst of noise)
    assert rate[0] < rate[-1] and max(false_pos[2:]) <= 2, (rate, false_pos)
