"""hpxfft::shared::loop mirror (core/include/hpxfft/shared/loop.hpp:15-31) over the C ABI."""
from __future__ import annotations

import ctypes as C

from . import capi
from .util import vector_2d

_PLAN_FLAGS = ("estimate", "measure", "patient", "exhaustive")  # util/adapter_fftw.hpp:22-44


def check_plan_flag(flag: str) -> None:
    if flag not in _PLAN_FLAGS:
        raise ValueError("Invalid FFTW plan flag string")  # std::invalid_argument in the reference


class loop:
    """Single-locality 2-D r2c FFT.  Same call sequence as the reference:
        fft = loop(); fft.initialize(values_vec, "estimate"); out = fft.fft_2d_r2c_par()
    The object is single-use per initialize (the reference returns std::move(values_vec_))."""

    def __init__(self, device: int = -1):
        self._lib = capi.load()
        self._plan = C.c_void_p()
        self._values = None
        self._device = device
        self._extra = {}

    def initialize(self, values_vec: vector_2d, PLAN_FLAG: str) -> None:
        check_plan_flag(PLAN_FLAG)
        self._destroy()
        self._values = values_vec
        capi.check(self._lib.hpxfft_b200_create(C.byref(self._plan), values_vec.n_row(), values_vec.n_col(), 0, 1,
                                                self._device, None, PLAN_FLAG.encode(), None))
        capi.check(self._lib.hpxfft_b200_upload(self._plan, values_vec.data().ctypes.data))

    def _run(self) -> vector_2d:
        if not self._plan or self._values is None:
            raise RuntimeError("loop: initialize() must be called before fft_2d_r2c")
        capi.check(self._lib.hpxfft_b200_execute(self._plan))
        out, self._values = self._values, None
        capi.check(self._lib.hpxfft_b200_download(self._plan, out.data().ctypes.data))
        return out

    def fft_2d_r2c_par(self) -> vector_2d:
        return self._run()

    def fft_2d_r2c_seq(self) -> vector_2d:
        # the reference's _seq differs only in CPU scheduling (and has an out-of-bounds first
        # transpose, shared/loop.cpp:126); on the GPU both names run the same kernels
        return self._run()

    fft_2d_r2c = fft_2d_r2c_par  # BASELINE wording

    def get_measurement(self, name: str) -> float:
        if not self._plan:
            return 0.0
        return float(self._lib.hpxfft_b200_measurement(self._plan, name.encode()))

    def write_plans_to_file(self, file_path: str) -> None:
        rc = self._lib.hpxfft_b200_write_plans(self._plan, file_path.encode())
        if rc != capi.OK:
            raise RuntimeError("Failed to open file: " + file_path)  # shared/loop.cpp:198-201

    # extensions used by the benchmark (device-resident operation)
    def plan_handle(self) -> C.c_void_p:
        return self._plan

    def _destroy(self) -> None:
        if self._plan:
            self._lib.hpxfft_b200_destroy(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass
