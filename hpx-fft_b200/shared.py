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
        self._lib.hpxfft_b200_synchronize(self._plan)   # refreshes the timers of transforms that were only enqueued
        return float(self._lib.hpxfft_b200_measurement(self._plan, name.encode()))

    def write_plans_to_file(self, file_path: str) -> None:
        rc = self._lib.hpxfft_b200_write_plans(self._plan, file_path.encode())
        if rc != capi.OK:
            raise RuntimeError("Failed to open file: " + file_path)  # shared/loop.cpp:198-201

    def initialize_device(self, device_ptr: int, n_row: int, n_col: int, PLAN_FLAG: str) -> None:
        """Extension: the slab already lives on the GPU (e.g. `tensor.data_ptr()` of a float64 CUDA tensor in the
        vector_2d layout); nothing crosses PCIe.  Pair with fft_2d_r2c_device()."""
        check_plan_flag(PLAN_FLAG)
        self._destroy()
        self._values = None
        capi.check(self._lib.hpxfft_b200_create(C.byref(self._plan), n_row, n_col, 0, 1, self._device, None, PLAN_FLAG.encode(), None))
        capi.check(self._lib.hpxfft_b200_upload(self._plan, device_ptr))

    def fft_2d_r2c_device(self, device_out_ptr: int) -> None:
        if not self._plan:
            raise RuntimeError("loop: initialize() must be called before fft_2d_r2c")
        capi.check(self._lib.hpxfft_b200_execute(self._plan))
        capi.check(self._lib.hpxfft_b200_download(self._plan, device_out_ptr))

    def fft_2d_r2c_async(self):
        """Enqueue the transform and the copy back; returns a concurrent.futures.Future that a CUDA stream callback
        (hpxfft_b200_on_complete) fulfils with the vector_2d -- the agas client surface, no thread waits on the GPU."""
        from concurrent.futures import Future
        if not self._plan or self._values is None:
            raise RuntimeError("loop: initialize() must be called before fft_2d_r2c")
        fut = Future()
        out, self._values = self._values, None
        capi.check(self._lib.hpxfft_b200_execute_async(self._plan))
        capi.check(self._lib.hpxfft_b200_download_async(self._plan, out.data().ctypes.data))

        def done(_user):
            fut.set_result(out)
        cb = capi.CALLBACK(done)
        self._extra["cb"] = cb          # keep the trampoline alive until it has run
        capi.check(self._lib.hpxfft_b200_on_complete(self._plan, cb, None))
        return fut

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


class agas:
    """Client surface of hpxfft::shared::agas (core/include/hpxfft/shared/agas.hpp:13-27): the same two calls returning
    futures.  initialize() (plan creation, host work) runs on a worker thread; fft_2d_r2c() only enqueues and its future
    is fulfilled by a CUDA stream callback."""

    def __init__(self, device: int = -1):
        from concurrent.futures import ThreadPoolExecutor
        self._loop = loop(device)
        self._pool = ThreadPoolExecutor(max_workers=1)

    def initialize(self, values_vec: vector_2d, PLAN_FLAG: str):
        check_plan_flag(PLAN_FLAG)
        return self._pool.submit(self._loop.initialize, values_vec, PLAN_FLAG)

    def fft_2d_r2c(self):
        return self._loop.fft_2d_r2c_async()

    def get_measurement(self, name: str) -> float:
        return self._loop.get_measurement(name)
