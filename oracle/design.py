"""Filter design restated from scipy.signal 1.18.1 (TEST INFRASTRUCTURE — the oracle).

The reference never designs a filter itself: it calls ``scipy.signal.decimate(x, q)``
(/root/reference/tetraear/signal/processor.py:254) and ``scipy.signal.butter(4, wn)`` +
``filtfilt`` (processor.py:78-79).  scipy is a third-party, un-vendored dependency of the
reference (requirements.txt: ``scipy>=1.10.0``, no lock); the oracle of record is the
version in this image, scipy 1.18.1 / numpy 2.3.5.  This module restates, in plain numpy,
the published algorithm those calls execute:

* ``cheby1(8, 0.05, 0.8/q, output='sos')``  = cheb1ap -> lp2lp_zpk -> bilinear_zpk -> zpk2sos
* ``butter(4, wn)``                         = buttap  -> lp2lp_zpk -> bilinear_zpk -> zpk2tf
* ``sosfilt_zi`` / ``lfilter_zi``           (steady-state initial conditions for filtfilt)

``tests/test_oracle_design.py`` pins every function here against scipy itself (bit-equal
float32 SOS/zi, <=1 ulp float64 b/a) for q in {2..100} and the Butterworth cut-offs the
path can produce.
"""
import numpy as np


def _cheb1ap(n, rp):
    eps = np.sqrt(10 ** (0.1 * rp) - 1.0)
    mu = 1.0 / n * np.arcsinh(1 / eps)
    m = np.arange(-n + 1, n, 2)
    theta = np.pi * m / (2 * n)
    p = -np.sinh(mu + 1j * theta)
    k = np.prod(-p, axis=0).real
    if n % 2 == 0:
        k = k / np.sqrt(1 + eps * eps)
    return p, k


def _buttap(n):
    m = np.arange(-n + 1, n, 2)
    p = -np.exp(1j * np.pi * m / (2 * n))
    return p, 1.0


def _analog_lowpass_to_digital(p, k, wn):
    """lp2lp_zpk(wo = 4 tan(pi wn / 2)) followed by bilinear_zpk(fs = 2); no analog zeros."""
    fs = 2.0
    warped = 2 * fs * np.tan(np.pi * wn / fs)
    degree = len(p)
    p = warped * p
    k = k * warped ** degree
    fs2 = 2.0 * fs
    pz = (fs2 + p) / (fs2 - p)
    kz = k * np.real(1.0 / np.prod(fs2 - p))
    zz = -np.ones(degree)
    return zz, pz, kz


def cheby1_sos(q, order=8, rp=0.05):
    """float64 SOS rows [b0 b1 b2 1 a1 a2] of ``cheby1(order, rp, 0.8/q, output='sos')``.

    zpk2sos('nearest') for an all-pole-pairs low-pass with every zero at -1 orders the
    sections by pole radius, the pair nearest the unit circle LAST, and folds the gain
    into the first section (SURVEY.md Appendix A.1 lists the q=10 rows).
    """
    p, k = _cheb1ap(order, rp)
    _, pz, kz = _analog_lowpass_to_digital(p, k, 0.8 / q)
    pos = pz[pz.imag > 0]
    assert 2 * len(pos) == order, "even order with complex pairs only"
    # distance from the unit circle decides the section order (worst pole last)
    pos = pos[np.argsort(np.abs(1 - np.abs(pos)))[::-1]]
    sos = np.zeros((len(pos), 6))
    for i, pp in enumerate(pos):
        # np.poly([p, conj p]) = [1, -(p + p*), p p*]
        a = np.convolve(np.array([1.0, -pp]), np.array([1.0, -np.conj(pp)]))
        sos[i, :3] = [1.0, 2.0, 1.0]
        sos[i, 3:] = a.real
    sos[0, :3] *= kz
    return sos


def butter_ba(wn, order=4):
    """float64 (b, a) of ``butter(order, wn)`` (TF form through np.poly, like zpk2tf)."""
    p, k = _buttap(order)
    zz, pz, kz = _analog_lowpass_to_digital(p, k, wn)
    b = kz * np.poly(zz)
    a = np.real(np.poly(pz))
    return np.ascontiguousarray(np.atleast_1d(b), dtype=np.float64), np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)


def lfilter_zi(b, a):
    """scipy.signal.lfilter_zi, same dtype in as out (float32 stays float32)."""
    b = np.asarray(b)
    a = np.asarray(a)
    if a[0] != 1:
        b, a = b / a[0], a / a[0]
    y_inf = np.sum(b) / np.sum(a)
    return np.flip(np.cumsum(np.flip(b - y_inf * a)))[1:]


def sosfilt_zi(sos):
    """scipy.signal.sosfilt_zi in the dtype of ``sos`` (decimate passes float32/complex64)."""
    sos = np.asarray(sos)
    zi = np.empty((sos.shape[0], 2), dtype=sos.dtype)
    scale = 1.0
    for s in range(sos.shape[0]):
        b = sos[s, :3]
        a = sos[s, 3:]
        zi[s] = scale * lfilter_zi(b, a)
        scale *= np.sum(b) / np.sum(a)
    return zi


def decimator_design(q):
    """float32 (sos, zi) exactly as ``scipy.signal.decimate`` uses them for complex64 IQ.

    decimate casts the SOS to the input dtype, complex64 (imaginary parts zero), and
    sosfiltfilt calls sosfilt_zi on THAT array, so zi is evaluated in complex64 arithmetic.
    numpy's complex64 quotient is ``a * (1/b)`` for a real divisor (two roundings), which
    is why the zi differ in the last bit from a float32 evaluation for ~30 % of the q's.
    """
    sos_c = np.asarray(cheby1_sos(q), dtype=np.complex64)
    zi_c = sosfilt_zi(sos_c)
    assert not np.any(sos_c.imag) and not np.any(zi_c.imag)
    return np.ascontiguousarray(sos_c.real), np.ascontiguousarray(zi_c.real)


def channel_filter_design(bandwidth, fs):
    """(b, a, zi) of filter_signal (processor.py:70-79): butter(4, clip((bw/2)/(fs/2)))."""
    cutoff = (bandwidth / 2) / (fs / 2)
    cutoff = min(0.99, max(0.01, cutoff))
    b, a = butter_ba(cutoff, 4)
    return b, a, np.ascontiguousarray(lfilter_zi(b, a))
