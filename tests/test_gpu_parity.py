"""Parity of the CUDA path (through the C ABI / the loop classes) against the oracle.  Needs a GPU.

Tolerances: the golden 4x4 case is exact (`==`, test/src/test_shared_loop.cpp:53); everything else is
relative L2 error <= 1e-12 over all nx*(ny/2+1) complex outputs (BASELINE.json north_star), measured
against the FP64 pocketfft restatement (itself ~2e-16 from the long-double oracle)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def shared_fft(pkg, a, method="fft_2d_r2c_par"):
    fft = pkg.shared.loop(device=0)
    fft.initialize(pkg.vector_2d.from_array(a.copy()), "estimate")
    out = getattr(fft, method)()
    return out.data(), fft


# ---------------------------------------------------------------- golden vector (reference's own test)
@pytest.mark.parametrize("method", ["fft_2d_r2c_par", "fft_2d_r2c_seq", "fft_2d_r2c"])
def test_golden_shared_loop(pkg, oracle, method):
    out, fft = shared_fft(pkg, oracle.GOLDEN_4x4_IN, method)
    assert np.array_equal(out, oracle.GOLDEN_4x4_OUT)          # REQUIRE(out2 == expected_output)
    assert fft.get_measurement("total") >= 0.0                  # REQUIRE(total >= 0.0)
    assert fft.get_measurement("plan_flops") > 0.0
    assert fft.get_measurement("no_such_key") == 0.0


@pytest.mark.parametrize("comm", ["scatter", "all_to_all", "p2p"])
def test_golden_distributed_loop_one_locality(pkg, oracle, comm):
    # test/src/test_distributed_loop.cpp:17-48 with num_localities == 1
    fft = pkg.distributed.loop(device=0)
    fft.initialize(pkg.vector_2d.from_array(oracle.GOLDEN_4x4_IN.copy()), comm, "estimate")
    out = fft.fft_2d_r2c()
    assert np.array_equal(out.data(), oracle.GOLDEN_4x4_OUT)
    assert fft.get_measurement("total") >= 0.0


def test_golden_distributed_agas(pkg, oracle):
    # test/src/test_distributed_agas.cpp:17-48 (plan flag "measure")
    fft = pkg.distributed.agas(device=0)
    fft.initialize(pkg.vector_2d.from_array(oracle.GOLDEN_4x4_IN.copy()), "scatter", "measure").result()
    out = fft.fft_2d_r2c().result()
    assert np.array_equal(out.data(), oracle.GOLDEN_4x4_OUT)


def test_golden_shared_agas(pkg, oracle):
    # test/src/test_shared_agas.cpp (plan flag "measure"); the transform future is fulfilled by a stream callback
    fft = pkg.shared.agas(device=0)
    fft.initialize(pkg.vector_2d.from_array(oracle.GOLDEN_4x4_IN.copy()), "measure").result()
    fut = fft.fft_2d_r2c()
    out = fut.result(timeout=60)
    assert np.array_equal(out.data(), oracle.GOLDEN_4x4_OUT)
    assert fft.get_measurement("total") >= 0.0
    # a larger transform whose future is pending while the host goes on
    a = oracle.make_input(1024, 4096, oracle.PATTERN_UNIFORM, seed=9)
    fft2 = pkg.shared.agas(device=0)
    fft2.initialize(pkg.vector_2d.from_array(a.copy()), "estimate").result()
    fut2 = fft2.fft_2d_r2c()
    ref = oracle.fft_2d_r2c_shared(a, workers=4)          # host work overlapping the GPU round trip
    assert oracle.rel_l2(fut2.result(timeout=60).data(), ref) <= TOL


# ---------------------------------------------------------------- 1-D kernels (the fftw_adapter seam)
@pytest.mark.parametrize("ny", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072])
def test_r2c_rows(pkg, lib, oracle, ny):
    batch = 37 if ny < 8192 else 5
    rng = np.random.default_rng(ny)
    a = np.zeros((batch, ny + 2))
    a[:, :ny] = rng.uniform(-1, 1, (batch, ny))
    got = a.copy()
    pkg.capi.check(lib.hpxfft_b200_r2c_rows(got.ctypes.data, batch, ny + 2, 0))
    import scipy.fft as sfft
    ref = sfft.rfft(a[:, :ny].astype(np.longdouble), axis=1)
    assert oracle.rel_l2(got, np.ascontiguousarray(ref).view(np.longdouble).reshape(batch, -1)) <= 1e-13


@pytest.mark.parametrize("env,ny", [({"HPXFFT_B200_ROWS_V1": "1"}, 16384), ({"HPXFFT_B200_ROWS_PF": "1"}, 16384),
                                    ({"HPXFFT_B200_ROWS_GENERAL": "1"}, 16384), ({"HPXFFT_B200_ROWS_ILV": "1"}, 16384),
                                    ({"HPXFFT_B200_ROWS_ILV": "1", "HPXFFT_B200_ROWS_GENERAL": "1"}, 16384),
                                    ({"HPXFFT_B200_ROWS_ILV": "0"}, 16384),
                                    ({"HPXFFT_B200_ROWS_LONG": "1"}, 32768), ({"HPXFFT_B200_ROWS_LONG": "2"}, 32768),
                                    ({"HPXFFT_B200_ROWS_LONG": "2", "HPXFFT_B200_ROWS_GENERAL": "1"}, 32768),
                                    ({"HPXFFT_B200_ROWS_PF": "0"}, 32768),
                                    ({"HPXFFT_B200_ROWS_LONG": "1"}, 65536), ({"HPXFFT_B200_ROWS_LONG": "1"}, 131072)])
def test_r2c_rows_selectable_variants(pkg, lib, oracle, monkeypatch, env, ny):
    """The row kernels that are not the default for their length stay selectable for A/B runs and stay correct (launch_rows.cu):
    ROWS_V1: the Stockham kernel rows_r2c_kernel<8192> for ny = 16384 (default: rows_r2c_v2_kernel); ROWS_PF: bulk L2 prefetch of
    the next row on / off; ROWS_GENERAL: the distributed slabs' output addressing on one GPU;
    ROWS_LONG=1: rows_long_kernel<C>, 2: rows_long2_kernel (ny = 32768) -- the defaults are the decimation-in-time kernels."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    batch = 5
    a = np.zeros((batch, ny + 2))
    a[:, :ny] = np.random.default_rng(ny + 1).uniform(-1, 1, (batch, ny))
    got = a.copy()
    pkg.capi.check(lib.hpxfft_b200_r2c_rows(got.ctypes.data, batch, ny + 2, 0))
    import scipy.fft as sfft
    ref = sfft.rfft(a[:, :ny].astype(np.longdouble), axis=1)
    assert oracle.rel_l2(got, np.ascontiguousarray(ref).view(np.longdouble).reshape(batch, -1)) <= 1e-13


@pytest.mark.parametrize("general", ["0", "1"])
@pytest.mark.parametrize("ny,batch,variant", [(32768, 5, "0"), (32768, 449, "0"), (32768, 5, "5"), (32768, 449, "5"),
                                               (65536, 5, "0"), (65536, 301, "0"), (131072, 3, "0"), (131072, 160, "0")])
def test_r2c_rows_decimation_in_time(pkg, lib, oracle, monkeypatch, general, ny, batch, variant):
    """The default long-row kernels: rows_dit2_kernel (ny = 32768, kernels_rows_dit2.cuh) and rows_ditc_kernel<C> (C = 4, 8;
    C = 2 with ROWS_LONG=5; kernels_rows_ditc.cuh), with the one-GPU output addressing and with the general one that the
    distributed slabs use.  The larger batches make every persistent CTA walk several rows (scratch reuse, next-row gather in
    flight, ragged last wave)."""
    monkeypatch.setenv("HPXFFT_B200_ROWS_LONG", variant)
    monkeypatch.setenv("HPXFFT_B200_ROWS_GENERAL", general)
    a = np.zeros((batch, ny + 2))
    a[:, :ny] = np.random.default_rng(ny + batch).uniform(-1, 1, (batch, ny))
    got = a.copy()
    pkg.capi.check(lib.hpxfft_b200_r2c_rows(got.ctypes.data, batch, ny + 2, 0))
    import scipy.fft as sfft
    ref = sfft.rfft(a[:, :ny], axis=1, workers=8)
    assert oracle.rel_l2(got, np.ascontiguousarray(ref).view(np.float64).reshape(batch, -1)) <= 1e-13
    if batch <= 5:   # bin by bin against extended precision: no output may be missing or misplaced
        refl = sfft.rfft(a[:, :ny].astype(np.longdouble), axis=1)
        err = np.abs(got.view(np.complex128).reshape(batch, -1) - refl.astype(np.complex128))
        assert err.max() <= 1e-9


@pytest.mark.parametrize("general", ["0", "1"])
def test_2d_32768_rows_decimation_in_time(pkg, oracle, monkeypatch, general):
    monkeypatch.setenv("HPXFFT_B200_ROWS_GENERAL", general)
    a = oracle.make_input(320, 32768, oracle.PATTERN_UNIFORM, seed=11)
    got, _ = shared_fft(pkg, a)
    assert oracle.rel_l2(got, oracle.fft_2d_r2c_shared(a, workers=8)) <= TOL


# variant 0 = what a plan launches: the persistent fused four-step kernel for n > 256, i.e. every (N1, N2) pair of
# launch_fused.cu -- 512 (32,16), 1024 (32,32), 2048 (64,32), 4096 (64,64), 8192 (128,64), 16384 (128,128),
# 32768 (256,128), 65536 (256,256), 131072 (512,256), 262144 (512,512); variant 1 = the unfused launches.
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072, 262144])
def test_c2c_cols(pkg, lib, oracle, n, variant):
    if variant == 1 and n <= 256:
        pytest.skip("single-level lengths have one kernel")
    width = 19 if n <= 16384 else 3          # ragged: one full 16-column tile + a partial one
    rng = np.random.default_rng(n)
    a = rng.uniform(-1, 1, (n, width)) + 1j * rng.uniform(-1, 1, (n, width))
    got = a.copy()
    pkg.capi.check(lib.hpxfft_b200_c2c_cols_variant(got.ctypes.data, n, width, 0, variant))
    import scipy.fft as sfft
    ref = sfft.fft(a.astype(np.clongdouble), axis=0)
    assert oracle.rel_l2(got.view(np.float64), np.ascontiguousarray(ref).view(np.longdouble)) <= 1e-13


def test_c2c_cols_32768_with_prestage(pkg, lib, oracle, monkeypatch):
    """nx = 32768 runs on 256 x 128 tiles by default; HPXFFT_B200_COLSPLIT=1 selects the radix-2 pre-stage + 128 x 128 variant."""
    monkeypatch.setenv("HPXFFT_B200_COLSPLIT", "1")
    n, width = 32768, 16 * 9 + 3
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, (n, width)) + 1j * rng.uniform(-1, 1, (n, width))
    got = a.copy()
    pkg.capi.check(lib.hpxfft_b200_c2c_cols_variant(got.ctypes.data, n, width, 0, 0))
    import scipy.fft as sfft
    ref = sfft.fft(a, axis=0, workers=8)
    assert oracle.rel_l2(got.view(np.float64), np.ascontiguousarray(ref).view(np.float64)) <= 1e-13


@pytest.mark.parametrize("n,width", [(512, 16 * 90 + 5), (4096, 16 * 45), (16384, 16 * 50 + 1), (32768, 16 * 24 + 7), (65536, 16 * 12)])
def test_c2c_cols_fused_ring_wraps(pkg, lib, oracle, n, width):
    """More strips than scratch-ring slots: the level-A tiles reuse slots that level-B tiles have drained
    (the dependency counters of cols_fused_kernel), with a ragged last strip."""
    rng = np.random.default_rng(n + width)
    a = rng.uniform(-1, 1, (n, width)) + 1j * rng.uniform(-1, 1, (n, width))
    got = a.copy()
    pkg.capi.check(lib.hpxfft_b200_c2c_cols_variant(got.ctypes.data, n, width, 0, 0))
    import scipy.fft as sfft
    ref = sfft.fft(a, axis=0, workers=8)
    assert oracle.rel_l2(got.view(np.float64), np.ascontiguousarray(ref).view(np.float64)) <= 1e-13


# ---------------------------------------------------------------- 2-D sweeps
SWEEP = [(2, 2), (2, 4), (4, 2), (8, 8), (16, 64), (64, 16), (1, 32), (32, 2), (128, 128), (256, 512), (512, 256),
         (1024, 64), (64, 2048), (512, 512), (2048, 1024), (1024, 4096), (64, 32768), (16, 65536), (8, 131072),
         # long columns through the 2-D path: every fused pair up to (512,256), short rows
         (4096, 64), (8192, 64), (16384, 32), (32768, 64), (65536, 32), (131072, 16),
         # both dimensions long: C=2 rows (ny = 32768) x (256,128) columns, the kernels of BASELINE configs 3/4
         (32768, 32768 // 64), (512, 32768), (1024, 65536)]


@pytest.mark.parametrize("nx,ny", SWEEP)
def test_small_pow2_sweep(pkg, oracle, nx, ny):
    a = oracle.make_input(nx, ny, oracle.PATTERN_UNIFORM, seed=42)
    got, _ = shared_fft(pkg, a)
    ref = oracle.fft_2d_r2c_longdouble(a) if nx * ny <= 1 << 18 else oracle.fft_2d_r2c_shared(a, workers=4)
    assert oracle.rel_l2(got, ref) <= TOL


@pytest.mark.parametrize("nx,ny", [(8, 14), (6, 10), (12, 30), (10, 16), (16, 18), (100, 200), (3, 2), (7, 1022), (1000, 64),
                                   # mixed radix (odd factor x power of two): sizes of the reference's weak-scaling sweep
                                   # 512 * t (benchmark/shared_benchmark.sh:100-102) and their factors
                                   (1536, 1536), (2560, 2560), (3584, 3584), (96, 24), (24, 96), (4608, 512), (512, 4608),
                                   (5632, 5632), (6656, 256), (256, 6656), (7680, 7680), (15872, 64), (64, 15872),
                                   (12288, 128), (128, 12288), (8704, 9728), (14336, 32), (7, 14), (62, 62)])
def test_generic_lengths(pkg, oracle, nx, ny):
    """Lengths that are not powers of two (FFTW accepts any n; 8 x 14 is the reference example's default,
    examples/hpxfft/shared_loop_2d.cpp:142-143): mixed radix when the odd factor is small, direct DFT otherwise."""
    a = oracle.make_input(nx, ny, oracle.PATTERN_UNIFORM, seed=7)
    got, _ = shared_fft(pkg, a)
    ref = oracle.fft_2d_r2c_longdouble(a) if nx * ny <= 1 << 20 else oracle.fft_2d_r2c_shared(a, workers=8)
    assert oracle.rel_l2(got, ref) <= TOL
    if (nx, ny) == (8, 14):
        r = oracle.make_input(8, 14, oracle.PATTERN_RAMP)
        z, _ = shared_fft(pkg, r)
        assert z[0, 0] == 728.0 and abs(z[0, 2] + 56) < 1e-10 and abs(z[0, 3] - 245.352031) < 1e-5


@pytest.mark.parametrize("t", list(range(1, 33)))
def test_reference_weak_sweep_sizes(pkg, lib, oracle, t):
    """Every size of the reference's shared weak-scaling sweep, nx = ny = 512 * t, t = 1..32 (benchmark/shared_benchmark.sh:100-102,
    sbatch_scripts/run_hpxfft_weak_shared.sh:34-42): separable input generated on the device, tile-sampled against the closed form."""
    import sampled
    n = 512 * t
    plan = C.c_void_p()
    pkg.capi.check(lib.hpxfft_b200_create(C.byref(plan), n, n + 2, 0, 1, 0, None, b"estimate", None))
    try:
        r = sampled.check_plan(lib, plan, n, n, 0, 1, seed=11)
        assert (r["num"] / r["den"]) ** 0.5 <= TOL, (t, r)
    finally:
        lib.hpxfft_b200_destroy(plan)


def test_x_dependence_is_real(pkg, oracle):
    # the reference's tests only ever feed x-constant rows; a delta in x must produce the right phase ramp
    nx, ny = 64, 32
    a = np.zeros((nx, ny + 2))
    a[5, 3] = 1.0
    got, _ = shared_fft(pkg, a)
    kx = np.arange(nx)[:, None]
    ky = np.arange(ny // 2 + 1)[None, :]
    ref = np.exp(-2j * np.pi * (5 * kx / nx + 3 * ky / ny))
    assert np.abs(oracle.to_complex(got) - ref).max() < 1e-13


@pytest.mark.parametrize("pattern", ["ramp", "uniform", "separable"])
def test_c1_anchor_256x16384(pkg, oracle, oracle_c, pattern):
    # BASELINE config 1: hpxfft_shared_loop --nx=256 --ny=16384
    nx, ny = 256, 16384
    pat = {"ramp": oracle.PATTERN_RAMP, "uniform": oracle.PATTERN_UNIFORM, "separable": oracle.PATTERN_SEPARABLE}[pattern]
    a = oracle.make_input(nx, ny, pat, seed=42)
    got, fft = shared_fft(pkg, a)
    ref = a.copy()
    assert oracle_c.hpxfft_oracle_shared_loop(ref.ctypes.data, nx, ny + 2, 0, None) == 0
    assert oracle.rel_l2(got, ref) <= TOL
    if pattern == "ramp":
        assert oracle.rel_l2(got, oracle.ramp_analytic(nx, ny)) <= TOL
    if pattern == "separable":
        s = oracle.separable_spectrum(nx, ny, 42)
        assert oracle.rel_l2(oracle.to_complex(got).real, s.real) <= TOL
        assert oracle.rel_l2(oracle.to_complex(got).imag, s.imag) <= TOL
    for key in ("total", "first_fftw", "second_fftw"):
        assert fft.get_measurement(key) > 0.0


def test_fill_matches_host_generator(pkg, lib, oracle):
    nx, ny = 64, 256
    for pat in (oracle.PATTERN_RAMP, oracle.PATTERN_UNIFORM, oracle.PATTERN_SEPARABLE):
        plan = C.c_void_p()
        pkg.capi.check(lib.hpxfft_b200_create(C.byref(plan), nx, ny + 2, 0, 1, 0, None, b"estimate", None))
        pkg.capi.check(lib.hpxfft_b200_fill(plan, pat, 42))
        got = np.empty((nx, ny + 2))
        pkg.capi.check(lib.hpxfft_b200_download(plan, got.ctypes.data))
        lib.hpxfft_b200_destroy(plan)
        ref = oracle.make_input(nx, ny, pat, seed=42)
        if pat == oracle.PATTERN_SEPARABLE:
            assert np.abs(got - ref).max() < 1e-14
        else:
            assert np.array_equal(got, ref)


def test_transform_end_to_end_and_plan_reuse(pkg, lib, oracle):
    nx, ny = 128, 512
    plan = C.c_void_p()
    pkg.capi.check(lib.hpxfft_b200_create(C.byref(plan), nx, ny + 2, 0, 1, 0, None, b"estimate", None))
    for seed in (1, 2, 3):   # the same plan object serves several transforms
        a = oracle.make_input(nx, ny, oracle.PATTERN_UNIFORM, seed=seed)
        got = a.copy()
        pkg.capi.check(lib.hpxfft_b200_transform(plan, got.ctypes.data))
        assert oracle.rel_l2(got, oracle.fft_2d_r2c_shared(a)) <= TOL
    assert lib.hpxfft_b200_launches_per_execute(plan) == 2   # rows + single-level columns (nx <= 256)
    lib.hpxfft_b200_destroy(plan)


def test_transform_async_double_buffered(pkg, lib, oracle):
    """Two plans driven alternately (what bench.py's e2e leg does): every round trip must still be exact."""
    nx, ny = 256, 1024
    plans = []
    for _ in range(2):
        plan = C.c_void_p()
        pkg.capi.check(lib.hpxfft_b200_create(C.byref(plan), nx, ny + 2, 0, 1, 0, None, b"estimate", None))
        plans.append(plan)
    hosts = [pkg.vector_2d(nx, ny + 2, 0.0, pinned=True) for _ in plans]
    refs = [None, None]
    for step in range(6):
        i = step % 2
        if step >= 2:
            pkg.capi.check(lib.hpxfft_b200_synchronize(plans[i]))
            assert oracle.rel_l2(hosts[i].data(), refs[i]) <= TOL
        a = oracle.make_input(nx, ny, oracle.PATTERN_UNIFORM, seed=100 + step)
        refs[i] = oracle.fft_2d_r2c_shared(a)
        hosts[i].data()[...] = a
        pkg.capi.check(lib.hpxfft_b200_transform_async(plans[i], hosts[i].data().ctypes.data))
    for i in range(2):
        pkg.capi.check(lib.hpxfft_b200_synchronize(plans[i]))
        assert oracle.rel_l2(hosts[i].data(), refs[i]) <= TOL
        lib.hpxfft_b200_destroy(plans[i])


def test_write_plans_to_file(pkg, oracle, tmp_path):
    _, fft = shared_fft(pkg, oracle.GOLDEN_4x4_IN)
    path = tmp_path / "plans" / "plan.txt"
    with pytest.raises(RuntimeError):                      # shared/loop.cpp:198-201
        fft.write_plans_to_file(str(path))
    path.parent.mkdir()
    fft.write_plans_to_file(str(path))
    text = path.read_text()
    assert "FFTW r2c 1D plan:" in text and "FFTW c2c 1D plan:" in text
    assert "single Stockham tile FFT" in text
    # a two-level size must describe the four-step column plan
    _, fft2 = shared_fft(pkg, np.zeros((1024, 66)))
    fft2.write_plans_to_file(str(path))
    assert "four-step 32 x 32" in path.read_text() and "fused persistent launch" in path.read_text()


# ---------------------------------------------------------------- full size (BASELINE config 2)
def test_c2_16384_separable_and_properties(pkg, lib, oracle):
    """16384 x 16384 on one GPU: input generated on the device, checked against the rank-4 separable
    closed form (1-D long-double FFTs) on every output, plus size-independent properties."""
    nx = ny = 16384
    plan = C.c_void_p()
    pkg.capi.check(lib.hpxfft_b200_create(C.byref(plan), nx, ny + 2, 0, 1, 0, None, b"estimate", None))
    try:
        pkg.capi.check(lib.hpxfft_b200_fill(plan, oracle.PATTERN_SEPARABLE, 42))
        pkg.capi.check(lib.hpxfft_b200_execute(plan))
        got = np.empty((nx, ny + 2))
        pkg.capi.check(lib.hpxfft_b200_download(plan, got.ctypes.data))
        z = oracle.to_complex(got)
        a, b = oracle.sep_vectors(nx, ny, 42)
        import scipy.fft as sfft
        fa = sfft.fft(a.astype(np.longdouble), axis=1).astype(np.complex128)
        fb = sfft.rfft(b.astype(np.longdouble), axis=1).astype(np.complex128)
        num = den = 0.0
        for r0 in range(0, nx, 1024):                       # blockwise to bound host memory
            ref = np.einsum("rx,ry->xy", fa[:, r0:r0 + 1024], fb)
            d = z[r0:r0 + 1024] - ref
            num += float(np.vdot(d, d).real)
            den += float(np.vdot(ref, ref).real)
        assert (num / den) ** 0.5 <= TOL
        # DC bin = sum of all inputs; Nyquist/DC columns of a real input have Hermitian symmetry in kx
        assert abs(z[0, 0].imag) <= 1e-6 * abs(z[0, 0].real) + 1e-6
        assert np.abs(z[1:, 0] - np.conj(z[:0:-1, 0])).max() <= 1e-9 * np.abs(z[:, 0]).max()
        assert np.abs(z[1:, -1] - np.conj(z[:0:-1, -1])).max() <= 1e-9 * np.abs(z[:, -1]).max() + 1e-9
        # ramp at full size against the closed form (the reference example's own input)
        pkg.capi.check(lib.hpxfft_b200_fill(plan, oracle.PATTERN_RAMP, 0))
        pkg.capi.check(lib.hpxfft_b200_execute(plan))
        pkg.capi.check(lib.hpxfft_b200_download(plan, got.ctypes.data))
        ref0 = oracle.ramp_analytic(1, ny)[0] * nx
        assert oracle.rel_l2(got[0], ref0) <= TOL
        assert np.abs(got[1:]).max() <= 1e-12 * np.abs(got[0]).max() * 16384
    finally:
        lib.hpxfft_b200_destroy(plan)


# ---------------------------------------------------------------- BASELINE config 3/4 kernels on one GPU
def test_c3_kernels_32768_sampled(pkg, lib, oracle):
    """32768 x 32768 on ONE GPU: rows_r2c_kernel<8192,2> and cols_fused_kernel<256,128> -- the kernels the
    multi-GPU configurations launch -- on device-generated separable input, tile-sampled against the closed form."""
    import sampled
    nx = ny = 32768
    plan = C.c_void_p()
    pkg.capi.check(lib.hpxfft_b200_create(C.byref(plan), nx, ny + 2, 0, 1, 0, None, b"estimate", None))
    try:
        r = sampled.check_plan(lib, plan, nx, ny, 0, 1, seed=42)
        assert (r["num"] / r["den"]) ** 0.5 <= TOL and r["max_tile_rel"] <= 10 * TOL and r["tiles"] == 16
    finally:
        lib.hpxfft_b200_destroy(plan)


def test_download_tile_and_device_pointer_upload(pkg, lib, oracle):
    nx, ny = 64, 256
    plan, plan2 = C.c_void_p(), C.c_void_p()
    pkg.capi.check(lib.hpxfft_b200_create(C.byref(plan), nx, ny + 2, 0, 1, 0, None, b"estimate", None))
    pkg.capi.check(lib.hpxfft_b200_create(C.byref(plan2), nx, ny + 2, 0, 1, 0, None, b"estimate", None))
    try:
        a = oracle.make_input(nx, ny, oracle.PATTERN_UNIFORM, seed=3)
        pkg.capi.check(lib.hpxfft_b200_upload(plan, a.ctypes.data))
        t = np.empty((5, 12))
        pkg.capi.check(lib.hpxfft_b200_download_tile(plan, 7, 5, 20, 12, t.ctypes.data))
        assert np.array_equal(t, a[7:12, 20:32])
        assert lib.hpxfft_b200_download_tile(plan, 60, 5, 0, 4, t.ctypes.data) == pkg.capi.EINVAL
        # device-pointer hand-over: plan2 initialised from plan's device slab, result back into it
        pkg.capi.check(lib.hpxfft_b200_upload(plan2, lib.hpxfft_b200_device_ptr(plan)))
        pkg.capi.check(lib.hpxfft_b200_execute(plan2))
        pkg.capi.check(lib.hpxfft_b200_download(plan2, lib.hpxfft_b200_device_ptr(plan)))
        got = np.empty_like(a)
        pkg.capi.check(lib.hpxfft_b200_download(plan, got.ctypes.data))
        assert oracle.rel_l2(got, oracle.fft_2d_r2c_shared(a)) <= TOL
    finally:
        lib.hpxfft_b200_destroy(plan)
        lib.hpxfft_b200_destroy(plan2)


def test_initialize_from_device_memory(pkg, oracle):
    """initialize from / return to device memory (a torch CUDA tensor in the vector_2d layout), no PCIe."""
    import torch
    nx, ny = 256, 2048
    a = oracle.make_input(nx, ny, oracle.PATTERN_UNIFORM, seed=21)
    t = torch.from_numpy(a).cuda()
    out = torch.empty_like(t)
    fft = pkg.shared.loop(device=0)
    fft.initialize_device(t.data_ptr(), nx, ny + 2, "estimate")
    fft.fft_2d_r2c_device(out.data_ptr())
    torch.cuda.synchronize()
    assert oracle.rel_l2(out.cpu().numpy(), oracle.fft_2d_r2c_shared(a)) <= TOL
    assert torch.equal(t.cpu(), torch.from_numpy(a))       # the caller's input tensor is untouched
