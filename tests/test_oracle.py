"""Pins the oracle (oracle/oracle.py, oracle/hpxfft_oracle.c) on the reference's golden vector and on
independent closed forms.  CPU only."""
import numpy as np
import pytest


def run_c(oracle_c, a, threads=0):
    v = np.ascontiguousarray(a, dtype=np.float64).copy()
    t = np.zeros(5)
    rc = oracle_c.hpxfft_oracle_shared_loop(v.ctypes.data, v.shape[0], v.shape[1], threads, t.ctypes.data)
    assert rc == 0
    return v, t


def test_golden_4x4_python_oracle(oracle):
    # test/src/test_shared_loop.cpp:15-34,53 -- exact equality
    out = oracle.fft_2d_r2c_shared(oracle.GOLDEN_4x4_IN)
    assert np.array_equal(out, oracle.GOLDEN_4x4_OUT)


def test_golden_4x4_c_oracle(oracle, oracle_c):
    out, t = run_c(oracle_c, oracle.GOLDEN_4x4_IN)
    assert np.array_equal(out, oracle.GOLDEN_4x4_OUT)
    assert t[0] >= 0.0  # REQUIRE(total >= 0.0), test_shared_loop.cpp:50-52


def test_golden_fixture_file(oracle):
    import json
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "golden_4x4.json")
    g = json.load(open(path))
    assert np.array_equal(np.array(g["input"]), oracle.GOLDEN_4x4_IN)
    assert np.array_equal(np.array(g["expected"]), oracle.GOLDEN_4x4_OUT)
    out = oracle.fft_2d_r2c_shared(np.array(g["input"]))
    assert np.array_equal(out, np.array(g["expected"]))


def test_default_example_8x14(oracle, oracle_c):
    # examples/hpxfft/shared_loop_2d.cpp:142-143 defaults; SURVEY appendix A quotes row 0
    v = oracle.make_input(8, 14, oracle.PATTERN_RAMP)
    z = oracle.fft_2d_r2c_shared(v)
    assert z[0, 0] == 728.0 and abs(z[0, 1]) < 1e-12
    assert abs(z[0, 2] + 56) < 1e-10 and abs(z[0, 3] - 245.352031) < 1e-5
    assert abs(z[0, 5] - 116.285198) < 1e-5
    assert np.abs(z[1:]).max() < 1e-9
    zc, _ = run_c(oracle_c, v)
    assert oracle.rel_l2(zc, z) < 1e-14


@pytest.mark.parametrize("nx,ny", [(2, 2), (4, 4), (8, 14), (16, 64), (64, 16), (32, 1024), (128, 128), (6, 10), (256, 2048)])
def test_c_oracle_matches_pocketfft_and_longdouble(oracle, oracle_c, nx, ny):
    v = oracle.make_input(nx, ny, oracle.PATTERN_UNIFORM, seed=42)
    zp = oracle.fft_2d_r2c_shared(v)
    zl = oracle.fft_2d_r2c_longdouble(v)
    zc, _ = run_c(oracle_c, v)
    assert oracle.rel_l2(zp, zl) < 5e-16 * max(4, np.log2(nx * ny))
    assert oracle.rel_l2(zc, zl) < 5e-16 * max(4, np.log2(nx * ny))
    # and against numpy's rfft2 (an implementation sharing no code with either)
    ref = np.fft.rfft2(v[:, :ny])
    assert oracle.rel_l2(oracle.to_complex(zp).view(np.float64), ref.view(np.float64)) < 1e-14


@pytest.mark.parametrize("nx,ny", [(8, 16), (256, 16384)])
def test_ramp_analytic(oracle, nx, ny):
    v = oracle.make_input(nx, ny, oracle.PATTERN_RAMP)
    z = oracle.fft_2d_r2c_shared(v, workers=4)
    assert oracle.rel_l2(z, oracle.ramp_analytic(nx, ny)) < 1e-13


def test_separable_spectrum(oracle):
    nx, ny = 64, 256
    v = oracle.make_input(nx, ny, oracle.PATTERN_SEPARABLE, seed=7)
    z = oracle.to_complex(oracle.fft_2d_r2c_shared(v))
    s = oracle.separable_spectrum(nx, ny, 7)
    assert oracle.rel_l2(z.real, s.real) < 1e-14 and oracle.rel_l2(z.imag, s.imag) < 1e-14
    blk = oracle.separable_spectrum(nx, ny, 7, rows=slice(3, 9), cols=slice(100, 129))
    assert np.allclose(np.asarray(blk, dtype=np.complex128), z[3:9, 100:129], rtol=0, atol=1e-11)


def test_distributed_restatement_equals_shared(oracle):
    nx, ny, L = 32, 64, 4
    slabs = [oracle.make_input(nx // L, ny, oracle.PATTERN_UNIFORM, seed=5, row0=r * (nx // L)) for r in range(L)]
    full = oracle.make_input(nx, ny, oracle.PATTERN_UNIFORM, seed=5)
    assert np.array_equal(np.concatenate(slabs), full)  # counter-based generator is slab-consistent
    out = oracle.fft_2d_r2c_distributed(slabs)
    assert np.array_equal(np.concatenate(out), oracle.fft_2d_r2c_shared(full))


def test_input_generator_known_values(oracle):
    # pins the splitmix64 stream shared with the CUDA fill kernel
    v = oracle.make_input(1, 8, oracle.PATTERN_UNIFORM, seed=42)[0, :4]
    assert np.allclose(v, [0.48312976, 0.45635755, -0.57328248, -0.86283822], atol=1e-8)
    assert np.all(oracle.make_input(3, 6, oracle.PATTERN_UNIFORM)[:, 6:] == 0.0)
