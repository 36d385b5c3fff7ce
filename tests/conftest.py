import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no built artefacts (they are git-ignored): build them once, in-tree
    lib = os.path.join(ROOT, "hpx-fft_b200", "libhpxfft_b200.so")
    if not os.path.exists(lib):
        entry.build()


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def oracle():
    return entry.load_oracle()


@pytest.fixture(scope="session")
def lib(pkg):
    return pkg.capi.load()


@pytest.fixture(scope="session")
def oracle_c():
    """ctypes handle on the plain-C oracle (oracle/libhpxfft_oracle.so); built on demand."""
    import ctypes
    import subprocess
    path = os.path.join(ROOT, "oracle", "libhpxfft_oracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    lib = ctypes.CDLL(path)
    lib.hpxfft_oracle_shared_loop.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                              ctypes.c_void_p]
    lib.hpxfft_oracle_shared_loop.restype = ctypes.c_int
    lib.hpxfft_oracle_c2c_rows.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int]
    lib.hpxfft_oracle_c2c_rows.restype = ctypes.c_int
    return lib


def gpu_count():
    try:
        return entry.load_package().capi.load().hpxfft_b200_device_count()
    except Exception:
        return 0
