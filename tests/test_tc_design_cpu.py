"""CPU check of the tensor-core decimator's constants (tetraear-bladerf_b200/csrc/dsp_tc.cu).

The weight matrix B and the carry constants the kernel uses are produced on the host; here the kernel's algorithm
(one [slots x K] @ [K x 48] contraction, Kogge-Stone-equivalent lane recurrences, neighbour-pass carries, analytic
chunk-end states) is emulated in float64 numpy with those very constants and compared with the exact-arithmetic
zero-phase filter scipy.signal.decimate stands for (processor.py:254).  No GPU needed: this pins the math, the
`-m gpu` parity tests pin the CUDA implementation.
"""
import ctypes as C

import numpy as np
import pytest
from scipy import signal

from tetraear_b200 import _lib

R, PASS, GROUP, N = 3, 96, 384, 48


def design(q):
    lib = _lib.load()
    K = C.c_int(0)
    _lib.check(lib.tb200_debug_tc_design(q, None, 0, None, None, 0, None, 0, C.byref(K)))
    K = K.value
    B = np.zeros((K, N), np.float64)
    bhi = np.zeros((K // 4, N, 4), np.float32)
    blo = np.zeros((K // 4, N, 4), np.float32)
    prm = np.zeros(5 * 8 + 32 * 8 + 3 * 8 + 3 * 8 + 8 + 8, np.float32)
    _lib.check(lib.tb200_debug_tc_design(q, B.ctypes.data, B.size, bhi.ctypes.data, blo.ctypes.data, bhi.size,
                                         prm.ctypes.data, prm.size, C.byref(K := C.c_int(0))))
    o = 0

    def take(shape):
        nonlocal o
        n = int(np.prod(shape)) * 2
        v = prm[o:o + n].astype(np.float64).reshape(*shape, 2)
        o += n
        return v[..., 0] + 1j * v[..., 1]
    P = dict(lampow=take((5, 4)), lanepow=take((32, 4)), cF=take((3, 4)), cB=take((3, 4)), initF=take((4,)), initB=take((4,)))
    return B, bhi, blo, P


def padded(x, i):
    """scipy's odd extension (edge 27) continued by constants, at chunk sample indices i (may be out of range)."""
    n = len(x)
    i = np.asarray(i)
    t = i + 27
    out = np.empty(i.shape, np.complex128)
    inside = (i >= 0) & (i < n)
    out[inside] = x[i[inside]]
    left = i < 0
    tl = t[left]
    out[left] = 2 * x[0] - x[np.where(tl < 0, 27, 27 - tl)]
    right = i >= n
    u = i[right] - n
    out[right] = 2 * x[n - 1] - x[np.where(u < 27, n - 2 - u, n - 1 - 27)]
    return out


def emulate(x, q, B, P):
    n = len(x)
    n_dec = (n + q - 1) // q
    lead = 27 // q
    gpc = (n_dec + lead + 4 + GROUP - 1) // GROUP
    n_slots = gpc * GROUP // R
    K = B.shape[0]
    k0 = -lead + R * np.arange(n_slots)
    i_first = q * (k0 - 1)                                  # chunk sample under slot float 0 (the front pad)
    idx = i_first[:, None] + np.arange(K // 2)[None, :]
    s = padded(x, idx)
    A = np.empty((n_slots, K))
    A[:, 0::2] = s.real
    A[:, 1::2] = s.imag
    D = A @ B
    Y0 = D[:, 0:6:2] + 1j * D[:, 1:6:2]                     # [slot][r], I + jQ packed as a complex number
    def agg(c0):
        cols = D[:, c0:c0 + 16].reshape(n_slots, 4, 2, 2)   # [slot][mode][re/im][I/Q]
        return cols[:, :, 0, :] + 1j * cols[:, :, 1, :]      # complex state per (mode, component)
    AF, AB = agg(6), agg(22)
    lam3 = P["lampow"][0]
    y = np.zeros((n_slots, R), np.complex128)
    n_pass = n_slots // 32
    SF = np.zeros_like(AF)
    SB = np.zeros_like(AB)
    for p in range(n_pass):                                  # inclusive scans inside each pass, zero state at its ends
        a = slice(32 * p, 32 * p + 32)
        accF = np.zeros((4, 2), np.complex128)
        for l in range(32):
            accF = AF[32 * p + l] + lam3[:, None] * accF
            SF[32 * p + l] = accF
        accB = np.zeros((4, 2), np.complex128)
        for l in range(31, -1, -1):
            accB = AB[32 * p + l] + lam3[:, None] * accB
            SB[32 * p + l] = accB
    c0 = padded(x, np.array([-27]))[0]
    cM = padded(x, np.array([n + 26]))[0]
    comp = lambda c: np.array([c.real, c.imag])
    for p in range(n_pass):
        inF = SF[32 * p - 1] if p > 0 else P["initF"][:, None] * comp(c0)[None, :]
        inB = SB[32 * p + 32] if p < n_pass - 1 else P["initB"][:, None] * comp(cM)[None, :]
        for l in range(32):
            sl = 32 * p + l
            tF = (SF[sl - 1] if l > 0 else 0) + P["lanepow"][l][:, None] * inF
            tB = (SB[sl + 1] if l < 31 else 0) + P["lanepow"][31 - l][:, None] * inB
            for r in range(R):
                v = (P["cF"][r][:, None] * tF + P["cB"][r][:, None] * tB).real.sum(axis=0)    # [I, Q]
                y[sl, r] = Y0[sl, r] + v[0] + 1j * v[1]
    k = (k0[:, None] + np.arange(R)[None, :]).ravel()
    yy = y.ravel()
    keep = (k >= 0) & (k < n_dec)
    return yy[keep]


@pytest.mark.parametrize("q", [10, 8])
def test_tc_constants_reproduce_the_zero_phase_decimator(q):
    B, bhi, blo, P = design(q)
    K = B.shape[0]
    # the TF32 split of the weights: hi and lo are TF32 numbers (13 low mantissa bits clear) and hi + lo = B to 2^-21
    for m in (bhi, blo):
        assert not np.any(m.view(np.uint32) & 0x1FFF)
    canon = (bhi.astype(np.float64) + blo.astype(np.float64)).transpose(0, 2, 1).reshape(K, N)
    assert np.abs(canon - B).max() <= 2.0 ** -21 * np.abs(B).max()
    # I weights never touch Q samples
    assert not np.any(B[0::2, 1:6:2]) and not np.any(B[1::2, 0:6:2])
    assert not np.any(B[:2]) and not np.any(B[2 + 2 * R * q:]) and not np.any(B[:, 38:])
    rng = np.random.default_rng(5)
    n = 2 * GROUP * q * 2 + 777
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.3 + 0.2 * np.exp(2j * np.pi * 0.013 * np.arange(n))
    sos32 = np.zeros(24, np.float32)
    d64, z64, z32 = np.zeros(24), np.zeros(8), np.zeros(8, np.float32)
    _lib.check(_lib.load().tb200_design_decimator(q, d64.ctypes.data, z64.ctypes.data, sos32.ctypes.data, z32.ctypes.data))
    ref = signal.sosfiltfilt(sos32.astype(np.float64).reshape(4, 6), x)[::q]
    got = emulate(x, q, B, P)
    assert got.shape == ref.shape
    # the last 160 outputs also carry scipy's right-edge transient, which a separate kernel applies (dsp_tc.cu)
    body = slice(0, len(ref) - 200)
    assert np.abs(got[body] - ref[body]).max() <= 5e-7 * np.abs(ref).max()
