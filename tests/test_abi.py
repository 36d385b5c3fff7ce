"""The C-ABI library loads and exports every symbol include/hpxfft_b200.h declares; error behaviour
that needs no GPU.  CPU only."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, gpu_count


def header_functions():
    text = open(os.path.join(ROOT, "include", "hpxfft_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hpxfft_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg, lib):
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hpxfft_b200.h but not exported"
    assert sorted(pkg.capi.PROTOTYPES) == names  # the ctypes table covers the header exactly


def test_version_and_constants(pkg, lib):
    assert lib.hpxfft_b200_version() == 200
    hdr = open(os.path.join(ROOT, "include", "hpxfft_b200.h")).read()
    assert f"HPXFFT_B200_UNIQUE_ID_BYTES {pkg.capi.UNIQUE_ID_BYTES}" in hdr
    assert f"HPXFFT_B200_IPC_HANDLE_BYTES {pkg.capi.IPC_HANDLE_BYTES}" in hdr


def test_flag_errors_precede_device_checks(pkg, lib):
    plan = C.c_void_p()
    rc = lib.hpxfft_b200_create(C.byref(plan), 4, 6, 0, 1, -1, None, b"fastest", None)
    assert rc == pkg.capi.EPLANFLAG and not plan
    assert b"Invalid FFTW plan flag string" in lib.hpxfft_b200_last_error()  # util/adapter_fftw.hpp:42
    rc = lib.hpxfft_b200_create(C.byref(plan), 4, 6, 0, 1, -1, b"gather", b"estimate", None)
    assert rc == pkg.capi.ECOMMFLAG
    assert b"scatter or all_to_all" in lib.hpxfft_b200_last_error()  # distributed/loop.cpp:344
    rc = lib.hpxfft_b200_create(C.byref(plan), 0, 6, 0, 1, -1, None, b"estimate", None)
    assert rc == pkg.capi.EINVAL
    rc = lib.hpxfft_b200_create(C.byref(plan), 4, 6, 0, 2, -1, None, b"estimate", None)
    assert rc == pkg.capi.EINVAL  # shared::loop aborts unless 1 locality (examples/hpxfft/shared_loop_2d.cpp:12-17)


@pytest.mark.skipif(gpu_count() > 0, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(pkg, lib):
    plan = C.c_void_p()
    rc = lib.hpxfft_b200_create(C.byref(plan), 4, 6, 0, 1, -1, None, b"estimate", None)
    assert rc == pkg.capi.ECUDA and not plan
    assert b"no CPU fallback" in lib.hpxfft_b200_last_error()
    import numpy as np
    a = np.zeros((4, 6))
    assert lib.hpxfft_b200_r2c_rows(a.ctypes.data, 4, 6, -1) == pkg.capi.ECUDA
    assert lib.hpxfft_b200_c2c_cols(a.ctypes.data, 4, 3, -1) == pkg.capi.ECUDA
    with pytest.raises(pkg.capi.Hpxfft_b200Error):
        fft = pkg.shared.loop()
        fft.initialize(pkg.vector_2d(4, 6, 0.0), "estimate")


def test_partition(lib):
    c0, w = C.c_size_t(), C.c_size_t()
    for cy, P in [(3, 1), (3, 2), (8193, 8), (16385, 8), (16385, 4), (65537, 8), (9, 4), (8, 8)]:
        cover = []
        for r in range(P):
            assert lib.hpxfft_b200_partition(cy, P, r, C.byref(c0), C.byref(w)) == 0
            cover.append((c0.value, w.value))
        assert cover[0][0] == 0
        for (a, wa), (b, _) in zip(cover, cover[1:]):
            assert a + wa == b and wa == cy // P
        assert cover[-1][0] + cover[-1][1] == cy  # last rank absorbs cy mod P: nothing is dropped
    assert lib.hpxfft_b200_partition(3, 4, 0, C.byref(c0), C.byref(w)) != 0


def test_no_kernel_uses_the_miscompiled_cp_async_form(pkg, lib):
    """build.py rejects a library whose SASS contains cp.async (LDGSTS) with an unset uniform descriptor register -- the form
    ptxas 12.9 emitted once for a hinted cp.async and that faulted on a B200 with "illegal instruction"."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("hpxfft_b200_build", os.path.join(ROOT, "hpx-fft_b200", "build.py"))
    build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build)
    assert build.suspicious_sass() == []
