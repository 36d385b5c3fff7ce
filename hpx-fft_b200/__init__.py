"""hpxfft_b200 -- B200-native (sm_100a) replacement for HPX-FFT's 2-D r2c hot path.

The directory is called `hpx-fft_b200` (not a valid Python identifier); import it through
`__graft_entry__.load_package()` or tests/conftest.py, which register it as module `hpxfft_b200`.

Layout: csrc/ (CUDA kernels + C ABI, built into libhpxfft_b200.so), capi.py (ctypes prototypes of
include/hpxfft_b200.h) and the host-side mirrors of the reference classes:
    hpxfft_b200.util.vector_2d          <- hpxfft::util::vector_2d<double>
    hpxfft_b200.shared.loop             <- hpxfft::shared::loop
    hpxfft_b200.distributed.loop / agas <- hpxfft::distributed::loop / agas
There is no CPU fallback: importing works without a GPU, every compute call needs the CUDA library."""
from . import capi, distributed, shared, util  # noqa: F401
from .util import vector_2d  # noqa: F401

__all__ = ["capi", "shared", "distributed", "util", "vector_2d"]
