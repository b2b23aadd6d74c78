"""ctypes mirror of include/bwts_b200.h (libbwts_b200.so).

The reference (NealB/Bijective-BWT) is three C command-line tools with no Python
surface; this module exists for the test-suite and bench.py.  It is a thin pass-through:
every function forwards to the C ABI and raises BwtsError on a non-zero status.  There is
no fallback of any kind -- if the shared library is missing, or no CUDA device is
present, calls fail loudly.
"""
import ctypes
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libbwts_b200.so"

NCLASS = 16
NPHASE = 8
MAX_LEN = (1 << 31) - 1

EXPORTS = [
    "bwts_b200_forward", "bwts_b200_inverse", "bwts_b200_forward_blocks", "bwts_b200_inverse_blocks",
    "bwts_b200_device_count", "bwts_b200_create", "bwts_b200_destroy", "bwts_b200_reserve",
    "bwts_b200_forward_host", "bwts_b200_inverse_host", "bwts_b200_forward_device",
    "bwts_b200_inverse_device", "bwts_b200_get_stats", "bwts_b200_class_name", "bwts_b200_set_profile",
    "bwts_b200_strerror", "bwts_b200_last_cuda_error", "bwts_b200_version", "bwts_b200_tune",
    "bwts_b200_divsufsort", "bwts_b200_phase_name",
]


class BwtsError(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        msg = lib().bwts_b200_strerror(code).decode()
        super().__init__(f"{what}: {msg} (code {code})" if what else f"{msg} (code {code})")


class Stats(ctypes.Structure):
    _fields_ = [
        ("len", ctypes.c_long), ("direction", ctypes.c_int), ("factors", ctypes.c_long),
        ("longest_factor", ctypes.c_long), ("alphabet_bits", ctypes.c_int), ("initial_depth", ctypes.c_int),
        ("rounds", ctypes.c_int), ("radix_passes", ctypes.c_int), ("local_rounds", ctypes.c_int),
        ("cta_rounds", ctypes.c_int), ("live_sum", ctypes.c_long),
        ("splitters", ctypes.c_long), ("unreached", ctypes.c_long), ("launches", ctypes.c_long),
        ("total_ms", ctypes.c_double),
        ("class_launches", ctypes.c_long * NCLASS), ("class_ms", ctypes.c_double * NCLASS),
        ("class_bytes", ctypes.c_double * NCLASS),
        ("h2d_ms", ctypes.c_double), ("d2h_ms", ctypes.c_double), ("lyndon_fallback", ctypes.c_int),
        ("phase_ms", ctypes.c_double * NPHASE), ("arena_bytes", ctypes.c_long), ("first_live", ctypes.c_long),
        ("tuple_rounds", ctypes.c_int), ("tuple_live_sum", ctypes.c_long), ("inverse_attempts", ctypes.c_int),
        ("binned_rounds", ctypes.c_int),
    ]


_lib = None


def lib():
    """Load libbwts_b200.so (built by `make -C bijective-bwt_b200` / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = ctypes.CDLL(os.environ.get("BWTS_B200_LIB", str(LIB_PATH)))  # override: diagnostic builds (make profile-lib)
    vp, cl, ci = ctypes.c_void_p, ctypes.c_long, ctypes.c_int
    for name in ("bwts_b200_forward", "bwts_b200_inverse"):
        getattr(L, name).argtypes = [vp, cl, vp, ci]
    for name in ("bwts_b200_forward_blocks", "bwts_b200_inverse_blocks"):
        getattr(L, name).argtypes = [vp, cl, cl, vp, ctypes.POINTER(ci), ci]
    L.bwts_b200_create.argtypes = [ci]
    L.bwts_b200_create.restype = vp
    L.bwts_b200_destroy.argtypes = [vp]
    L.bwts_b200_destroy.restype = None
    L.bwts_b200_reserve.argtypes = [vp, cl]
    for name in ("bwts_b200_forward_host", "bwts_b200_inverse_host"):
        getattr(L, name).argtypes = [vp, vp, cl, vp]
    for name in ("bwts_b200_forward_device", "bwts_b200_inverse_device"):
        getattr(L, name).argtypes = [vp, vp, cl, vp, vp]
    L.bwts_b200_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.bwts_b200_class_name.argtypes = [ci]
    L.bwts_b200_class_name.restype = ctypes.c_char_p
    L.bwts_b200_phase_name.argtypes = [ci, ci]
    L.bwts_b200_phase_name.restype = ctypes.c_char_p
    L.bwts_b200_set_profile.argtypes = [vp, ci]
    L.bwts_b200_strerror.argtypes = [ci]
    L.bwts_b200_strerror.restype = ctypes.c_char_p
    L.bwts_b200_last_cuda_error.argtypes = [vp]
    L.bwts_b200_version.restype = ctypes.c_char_p
    L.bwts_b200_tune.argtypes = [ci, cl]
    L.bwts_b200_divsufsort.argtypes = [vp, vp, ci, ci]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise BwtsError(rc, what)


def _as_u8(data):
    if isinstance(data, np.ndarray):
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        return data
    return np.frombuffer(bytes(data), dtype=np.uint8)


def device_count():
    return lib().bwts_b200_device_count()


def version():
    return lib().bwts_b200_version().decode()


def tune(key, value):
    _check(lib().bwts_b200_tune(key, value), "bwts_b200_tune")


def _one_call(fn, what, data, device):
    src = _as_u8(data)
    out = np.empty(len(src), dtype=np.uint8)
    _check(fn(src.ctypes.data if len(src) else None, len(src), out.ctypes.data if len(src) else None, device), what)
    return out.tobytes()


def forward(data, device=0):
    """bwts_b200_forward: BWTS of `data` (bytes-like) -> bytes, host buffers."""
    return _one_call(lib().bwts_b200_forward, "bwts_b200_forward", data, device)


def inverse(data, device=0):
    """bwts_b200_inverse: inverse BWTS of `data` -> bytes, host buffers."""
    return _one_call(lib().bwts_b200_inverse, "bwts_b200_inverse", data, device)


def _blocks(fn, what, data, block_len, devices):
    src = _as_u8(data)
    out = np.empty(len(src), dtype=np.uint8)
    devs = (ctypes.c_int * len(devices))(*devices)
    _check(fn(src.ctypes.data, len(src), block_len, out.ctypes.data, devs, len(devices)), what)
    return out.tobytes()


def forward_blocks(data, block_len, devices=(0,)):
    return _blocks(lib().bwts_b200_forward_blocks, "bwts_b200_forward_blocks", data, block_len, list(devices))


def inverse_blocks(data, block_len, devices=(0,)):
    return _blocks(lib().bwts_b200_inverse_blocks, "bwts_b200_inverse_blocks", data, block_len, list(devices))


def blocks_ptr(direction, in_ptr, length, block_len, out_ptr, devices=(0,)):
    """bwts_b200_{forward,inverse}_blocks on raw host pointers (pinned buffers take the direct-copy path)."""
    fn = lib().bwts_b200_inverse_blocks if direction else lib().bwts_b200_forward_blocks
    devs = (ctypes.c_int * len(devices))(*devices)
    _check(fn(in_ptr, length, block_len, out_ptr, devs, len(devices)), "bwts_b200_*_blocks")


def suffix_array(data, device=0):
    """bwts_b200_divsufsort: suffix array (int32) of `data`."""
    src = _as_u8(data)
    sa = np.empty(len(src), dtype=np.int32)
    if len(src) == 0:
        return sa
    _check(lib().bwts_b200_divsufsort(src.ctypes.data, sa.ctypes.data, len(src), device), "bwts_b200_divsufsort")
    return sa


class Context:
    """bwts_b200_ctx: reusable device workspace (one per device per host thread)."""

    def __init__(self, device=0):
        self._h = lib().bwts_b200_create(device)
        if not self._h:
            raise BwtsError(-3, f"bwts_b200_create({device})")
        self.device = device

    def close(self):
        if self._h:
            lib().bwts_b200_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reserve(self, max_len):
        _check(lib().bwts_b200_reserve(self._h, max_len), "bwts_b200_reserve")

    def set_profile(self, on):
        _check(lib().bwts_b200_set_profile(self._h, int(on)), "bwts_b200_set_profile")

    # host buffers (addresses of `length` bytes; e.g. pinned torch tensors' data_ptr())
    def forward_host_ptr(self, in_ptr, length, out_ptr):
        _check(lib().bwts_b200_forward_host(self._h, in_ptr, length, out_ptr), "bwts_b200_forward_host")

    def inverse_host_ptr(self, in_ptr, length, out_ptr):
        _check(lib().bwts_b200_inverse_host(self._h, in_ptr, length, out_ptr), "bwts_b200_inverse_host")

    def forward_host(self, data):
        src = _as_u8(data)
        out = np.empty(len(src), dtype=np.uint8)
        self.forward_host_ptr(src.ctypes.data, len(src), out.ctypes.data)
        return out.tobytes()

    def inverse_host(self, data):
        src = _as_u8(data)
        out = np.empty(len(src), dtype=np.uint8)
        self.inverse_host_ptr(src.ctypes.data, len(src), out.ctypes.data)
        return out.tobytes()

    # device-resident buffers (device addresses, e.g. torch cuda tensors' data_ptr())
    def forward_device(self, d_in, length, d_out, stream=None):
        _check(lib().bwts_b200_forward_device(self._h, d_in, length, d_out, stream), "bwts_b200_forward_device")

    def inverse_device(self, d_in, length, d_out, stream=None):
        _check(lib().bwts_b200_inverse_device(self._h, d_in, length, d_out, stream), "bwts_b200_inverse_device")

    def stats(self):
        s = Stats()
        _check(lib().bwts_b200_get_stats(self._h, ctypes.byref(s)), "bwts_b200_get_stats")
        out = {f: getattr(s, f) for f, _ in Stats._fields_ if not f.startswith("class_") and f != "phase_ms"}
        out["arena_bytes_per_byte"] = s.arena_bytes / s.len if s.len else None
        phases = {}
        for ph in range(NPHASE):
            name = lib().bwts_b200_phase_name(s.direction, ph)
            if name:
                phases[name.decode()] = s.phase_ms[ph]
        out["phases"] = phases
        classes = {}
        for c in range(NCLASS):
            if s.class_launches[c]:
                classes[lib().bwts_b200_class_name(c).decode()] = {
                    "launches": s.class_launches[c], "ms": s.class_ms[c], "bytes": s.class_bytes[c]}
        out["classes"] = classes
        return out

    def last_cuda_error(self):
        return lib().bwts_b200_last_cuda_error(self._h)
