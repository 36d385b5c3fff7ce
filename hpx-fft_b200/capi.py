"""ctypes binding of the C ABI declared in include/hpxfft_b200.h (one prototype per exported symbol)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HPXFFT_B200_LIB") or os.path.join(HERE, "libhpxfft_b200.so")  # override: A/B builds

OK, EINVAL, EPLANFLAG, ECOMMFLAG, ECUDA, ENCCL, ESTATE = 0, -1, -2, -3, -4, -5, -6
UNIQUE_ID_BYTES = 128
IPC_HANDLE_BYTES = 64
PATTERN_RAMP, PATTERN_UNIFORM, PATTERN_SEPARABLE = 0, 1, 2

# symbol -> (restype, argtypes); must list every function of include/hpxfft_b200.h
PROTOTYPES = {
    "hpxfft_b200_version": (C.c_int, []),
    "hpxfft_b200_last_error": (C.c_char_p, []),
    "hpxfft_b200_device_count": (C.c_int, []),
    "hpxfft_b200_partition": (C.c_int, [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "hpxfft_b200_get_unique_id": (C.c_int, [C.c_void_p]),
    "hpxfft_b200_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                     C.c_char_p, C.c_char_p, C.c_void_p]),
    "hpxfft_b200_ipc_count": (C.c_int, [C.c_void_p]),
    "hpxfft_b200_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpxfft_b200_ipc_import": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpxfft_b200_transport": (C.c_char_p, [C.c_void_p]),
    "hpxfft_b200_bind_host_to_device": (C.c_int, [C.c_int]),
    "hpxfft_b200_download_tile": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]),
    "hpxfft_b200_bench_exchange": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "hpxfft_b200_c2c_cols_variant": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int]),
    "hpxfft_b200_upload_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpxfft_b200_download_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpxfft_b200_on_complete": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hpxfft_b200_upload": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpxfft_b200_download": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpxfft_b200_fill": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64]),
    "hpxfft_b200_execute": (C.c_int, [C.c_void_p]),
    "hpxfft_b200_execute_async": (C.c_int, [C.c_void_p]),
    "hpxfft_b200_synchronize": (C.c_int, [C.c_void_p]),
    "hpxfft_b200_reset_timers": (C.c_int, [C.c_void_p]),
    "hpxfft_b200_transform": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpxfft_b200_transform_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpxfft_b200_measurement": (C.c_double, [C.c_void_p, C.c_char_p]),
    "hpxfft_b200_write_plans": (C.c_int, [C.c_void_p, C.c_char_p]),
    "hpxfft_b200_device_ptr": (C.c_void_p, [C.c_void_p]),
    "hpxfft_b200_stream": (C.c_void_p, [C.c_void_p]),
    "hpxfft_b200_launches_per_execute": (C.c_int, [C.c_void_p]),
    "hpxfft_b200_destroy": (None, [C.c_void_p]),
    "hpxfft_b200_host_alloc": (C.c_void_p, [C.c_size_t]),
    "hpxfft_b200_host_free": (None, [C.c_void_p]),
    "hpxfft_b200_r2c_rows": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]),
    "hpxfft_b200_c2c_cols": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]),
}

_lib = None
CALLBACK = C.CFUNCTYPE(None, C.c_void_p)   # hpxfft_b200_callback


class Hpxfft_b200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"hpxfft_b200 error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    """Loads libhpxfft_b200.so.  There is no fallback: a missing library is a hard error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python hpx-fft_b200/build.py` "
                          "(nvcc, sm_100a); hpxfft_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != OK:
        raise Hpxfft_b200Error(rc, load().hpxfft_b200_last_error().decode())
