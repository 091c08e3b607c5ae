"""ctypes binding of libtetra_b200.so (include/tetra_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) into
``tetraear-bladerf_b200/lib/``.  There is no CPU fallback: a missing library or a missing
CUDA device raises ``RuntimeError`` from the first compute call.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtetra_b200.so")

DTYPE_C64, DTYPE_C128, DTYPE_SC16 = 0, 1, 2
MODE_EXACT, MODE_FAST = 0, 1
OPT_CHAN_F32 = 1
OPT_PSD_F32 = 2
OPT_DECIM_TC = 3
OPT_PSD_AVG = 4
OPT_ETSI_OPEN_LOOP = 5
OPT_ETSI_SEG_LANES = 6


class Config(C.Structure):
    _fields_ = [("sample_rate", C.c_double), ("max_chunks", C.c_int32), ("max_chunk_len", C.c_int32),
                ("mode", C.c_int32), ("device", C.c_int32), ("max_fft", C.c_int32), ("max_bursts", C.c_int32),
                ("max_stream_len", C.c_int32), ("max_channels", C.c_int32)]


class Plan(C.Structure):
    _fields_ = [("q", C.c_int32), ("n_dec", C.c_int32), ("filtered", C.c_int32), ("sps", C.c_int32),
                ("phase_step", C.c_int32), ("max_symbols", C.c_int32), ("rate", C.c_double)]


FRAME_DTYPE = np.dtype([("sync_pos", "<i4"), ("start_pos", "<i4"), ("number", "<i4"), ("type", "u1"),
                        ("enc_mode", "u1"), ("burst_type", "u1"), ("crc_ok", "u1"), ("mac_ok", "u1"),
                        ("kept", "u1"), ("mac_fill", "u1"), ("mac_length", "u1"), ("mac_address", "<u4"),
                        ("mac_data_bytes", "u1"), ("mac_unique", "u1"), ("mac_b0", "u1"), ("mac_b1", "u1"),
                        ("bits", "u1", (64,)), ("data", "u1", (64,))])
assert FRAME_DTYPE.itemsize == 156
DETECT_DTYPE = np.dtype([("power_sum", "<f8"), ("window_sum", "<f8", (8,)), ("max_abs", "<f8"), ("n", "<i8"),
                         ("window_size", "<i4"), ("n_steps", "<i4"), ("matches", "<i4"), ("n_sync_bits", "<i4"),
                         ("sync_best", "<i4"), ("reserved", "<i4")])
assert DETECT_DTYPE.itemsize == 112
VOICE_SYMBOLS, VOICE_BITS, VOICE_PARAMS, VOICE_SOFT = 0, 1, 2, 3

_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_PROTOS = {
    "tb200_last_error": (C.c_char_p, []),
    "tb200_version": (_i, []),
    "tb200_device_count": (_i, []),
    "tb200_design_decimator": (_i, [_i, _vp, _vp, _vp, _vp]),
    "tb200_design_channel_filter": (_i, [_d, _d, _vp, _vp, _vp]),
    "tb200_plan_process": (_i, [_d, _i, C.POINTER(Plan)]),
    "tb200_create": (_i, [C.POINTER(Config), C.POINTER(_vp)]),
    "tb200_destroy": (_i, [_vp]),
    "tb200_set_sample_rate": (_i, [_vp, _d]),
    "tb200_set_option": (_i, [_vp, _i, _i]),
    "tb200_workspace_bytes": (C.c_size_t, [_vp]),
    "tb200_launch_count": (_i64, [_vp]),
    "tb200_process_batch": (_i, [_vp, _vp, _i, _i, _i, _i64, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "tb200_process_decode_batch": (_i, [_vp, _vp, _i, _i, _i, _i64, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "tb200_process_decode_host": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "tb200_host_oneshot_fits": (_i, [_vp, _i, _i, _i, _i]),
    "tb200_decode_batch_host": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "tb200_process_batch_host": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "tb200_resample": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "tb200_decimate": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "tb200_freq_shift": (_i, [_vp, _vp, _i, _i, _d, _d, _vp, _vp]),
    "tb200_filter": (_i, [_vp, _vp, _i, _i, _d, _d, _vp, _vp]),
    "tb200_extract_symbols": (_i, [_vp, _vp, _i, _d, _vp, _vp, _vp, _vp]),
    "tb200_demod_dqpsk": (_i, [_vp, _vp, _i, _vp, _vp]),
    "tb200_symbols_to_bits": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "tb200_find_sync": (_i, [_vp, _vp, _i, _d, _vp, _i, _vp, _vp, _vp]),
    "tb200_decode_batch": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "tb200_burst_batch": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "tb200_burst_batch_host": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "tb200_crc16_rows": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "tb200_etsi_decode_batch": (_i, [_vp, _vp, _i64, _i, C.c_uint32, _vp, _vp, _vp, _vp]),
    "tb200_etsi_extract_bits": (_i, [_vp, _vp, _i, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "tb200_etsi_find_sequence": (_i, [_vp, _vp, _i, _i64, _vp, _i, C.c_double, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "tb200_etsi_aach_decode": (_i, [_vp, _vp, _i64, _vp, C.c_uint32, _vp, _vp, _vp, _vp]),
    "tb200_etsi_decode_custom": (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _i, _i, C.c_uint32, _vp, _vp, _vp, _vp]),
    "tb200_etsi_demod_batch": (_i, [_vp, _vp, _i, _i64, _i64, _d, _vp, _vp, _i64, _vp, _vp]),
    "tb200_etsi_demod_scratch_bytes": (_i64, [_i, _i64, _d, _i]),
    "tb200_etsi_demod_segmented": (_i, [_vp, _vp, _i, _i64, _i64, _d, _vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "tb200_etsi_find_bursts": (_i, [_vp, _vp, _i, _i64, _vp, _d, _vp, _vp, _vp, _vp, _i, _vp]),
    "tb200_etsi_afc": (_i, [_vp, _vp, _i, _i64, _vp, _vp, _vp]),
    "tb200_etsi_find_sync_bursts": (_i, [_vp, _vp, _i, _i64, _vp, _d, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "tb200_channelize": (_i, [_vp, _vp, _i64, _i, _vp, _i64, _i64, _vp]),
    "tb200_refchan_batch": (_i, [_vp, _vp, _i64, _i, _i, _d, _vp, _i64, _vp]),
    "tb200_profile_enable": (_i, [_vp, _i]),
    "tb200_profile_read": (_i, [_vp, _vp, _vp]),
    "tb200_debug_read": (_i, [_vp, _i, _vp, C.c_size_t]),
    "tb200_debug_tc_timing": (_i, [_vp, _i, _vp]),
    "tb200_debug_tc_design": (_i, [_i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp]),
    "tb200_psd_batch": (_i, [_vp, _vp, _i64, _i64, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _vp]),
    "tb200_detect_batch": (_i, [_vp, _vp, _i, _i, _i64, _i64, _d, _i, _vp, _vp]),
    "tb200_voice_blocks": (_i, [_vp, _i, _vp, _i64, _vp, _vp, _vp, _i64, _d, _vp, _vp, _vp]),
}
EXPORTS = tuple(_PROTOS)

_lib = None


def load():
    """Load the shared library (no device needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). tetraear_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def last_error():
    return load().tb200_last_error().decode("utf-8", "replace")


class ComputeError(RuntimeError):
    """A compute entry point refused its input (status code < 0).  Distinct from the RuntimeError raised when the
    library or a CUDA device is missing: the class API turns THIS one into the reference's log-and-return-empty
    behaviour (processor.py:81-83, :256-257), never the missing-backend one."""

    def __init__(self, rc, msg):
        super().__init__(f"libtetra_b200 error {rc}: {msg}")
        self.rc = rc


def check(rc):
    if rc != 0:
        raise ComputeError(rc, last_error())


def require_device():
    lib = load()
    if lib.tb200_device_count() <= 0:
        raise RuntimeError("tetraear_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return lib


def plan(sample_rate, chunk_len):
    p = Plan()
    check(load().tb200_plan_process(float(sample_rate), int(chunk_len), C.byref(p)))
    return p


class Handle:
    """RAII wrapper of tb200_handle."""

    def __init__(self, sample_rate=2.4e6, max_chunks=1, max_chunk_len=131072, mode=MODE_EXACT, device=0,
                 max_fft=4096, max_bursts=0, max_stream_len=0, max_channels=0):
        self.lib = require_device()
        self.cfg = Config(float(sample_rate), int(max_chunks), int(max_chunk_len), int(mode), int(device),
                          int(max_fft), int(max_bursts), int(max_stream_len), int(max_channels))
        self.h = _vp()
        check(self.lib.tb200_create(C.byref(self.cfg), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.tb200_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, option, value):
        check(self.lib.tb200_set_option(self.h, int(option), int(value)))

    def profile(self, on=True):
        check(self.lib.tb200_profile_enable(self.h, int(on)))

    def profile_read(self):
        """{stage name: (total ms, calls)} since the last read."""
        ms = (C.c_double * 8)()
        calls = (C.c_int64 * 8)()
        check(self.lib.tb200_profile_read(self.h, ms, calls))
        names = ("decimator", "channel_filter", "symbols", "bits_sync", "decode", "burst_batch", "psd", "etsi")
        return {n: (float(ms[i]), int(calls[i])) for i, n in enumerate(names)}

    @property
    def launches(self):
        return int(self.lib.tb200_launch_count(self.h))

    @property
    def workspace_bytes(self):
        return int(self.lib.tb200_workspace_bytes(self.h))
