"""CPU-side checks of the C-ABI library: it loads, exports every declared symbol, designs the
reference's filters bit-for-bit, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry
from oracle import design

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    from tetraear_b200 import _lib
    return _lib.load()


def test_exports_match_header(lib):
    from tetraear_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tetra_b200.h")).read()
    declared = set(re.findall(r"\b(tb200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/tetra_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.tb200_version() >= 100


def test_struct_sizes():
    from tetraear_b200 import _lib
    assert _lib.FRAME_DTYPE.itemsize == 156
    assert ctypes.sizeof(_lib.Config) == 40       # 8 + 8 x int32
    assert ctypes.sizeof(_lib.Plan) == 32
    assert _lib.DETECT_DTYPE.itemsize == 112     # tb200_detect_stats


def test_decimator_design_is_scipys(lib):
    for q in list(range(2, 101)):
        s64 = (ctypes.c_double * 24)(); z64 = (ctypes.c_double * 8)()
        s32 = (ctypes.c_float * 24)(); z32 = (ctypes.c_float * 8)()
        assert lib.tb200_design_decimator(q, s64, z64, s32, z32) == 0
        sos32, zi32 = design.decimator_design(q)
        assert np.array_equal(np.array(s32, dtype=np.float32).reshape(4, 6), sos32), q
        assert np.array_equal(np.array(z32, dtype=np.float32).reshape(4, 2), zi32), q
        sos = design.cheby1_sos(q)
        assert np.allclose(np.array(s64).reshape(4, 6), sos, rtol=1e-14, atol=0), q


def test_channel_filter_design_is_scipys(lib):
    for fs in (240000.0, 20e6 / 83, 1.8e6 / 7, 2.4e6, 1e6, 48000.0, 300000.0):
        for bw in (25000, 50000):
            b = (ctypes.c_double * 5)(); a = (ctypes.c_double * 5)(); z = (ctypes.c_double * 4)()
            assert lib.tb200_design_channel_filter(float(bw), float(fs), b, a, z) == 0
            ob, oa, oz = design.channel_filter_design(bw, fs)
            assert np.allclose(np.array(b), ob, rtol=1e-13, atol=0)
            assert np.allclose(np.array(a), oa, rtol=1e-13, atol=0)
            assert np.allclose(np.array(z), oz, rtol=1e-10, atol=0)


def test_plan_mirrors_process_control_flow(lib):
    from tetraear_b200 import _lib
    p = _lib.plan(2.4e6, 131072)
    assert (p.q, p.n_dec, p.filtered, p.sps, p.phase_step, p.max_symbols) == (10, 13108, 1, 13, 1, 1008)
    p = _lib.plan(20e6, 1 << 20)
    assert (p.q, p.sps) == (83, 13)
    p = _lib.plan(1.8e6, 98000)
    assert (p.q, p.sps) == (7, 14)
    assert _lib.plan(2.4e6, 27).q == 1 and _lib.plan(2.4e6, 28).q == 10    # scipy padlen rule
    assert _lib.plan(2.4e6, 150).filtered == 0 and _lib.plan(2.4e6, 160).filtered == 1
    assert _lib.plan(240000.0, 20000).q == 1


def test_no_cpu_fallback(lib):
    import torch
    from tetraear_b200 import _lib
    from tetraear_b200.signal import SignalProcessor
    if torch.cuda.is_available():
        pytest.skip("device present")
    assert lib.tb200_device_count() == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.Handle()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SignalProcessor().process(np.ones(1000, np.complex64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SignalProcessor().resample(np.ones(100, np.complex64), 1.2e6)
    from tetraear_b200.signal import TetraSignalDetector
    from tetraear_b200.voice import VoiceBlockBuilder
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TetraSignalDetector().detect_tetra_modulation(np.ones(2000, np.complex64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VoiceBlockBuilder()
    # below the reference's own length gates nothing is computed (scanner.py:66-67, :217-218)
    assert TetraSignalDetector().detect_tetra_modulation(np.ones(10, np.complex64)) == (False, 0.0)
    assert TetraSignalDetector().check_power_stability(np.ones(10, np.complex64)) is False
    # empty input needs no device, exactly like the reference (processor.py:239-241)
    sp = SignalProcessor()
    assert len(sp.process(np.array([], np.complex64))) == 0 and len(sp.symbols) == 0
