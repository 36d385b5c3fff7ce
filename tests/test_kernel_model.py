"""The index algebra the CUDA kernels implement, restated in NumPy (tools/model_kernels.py) and checked
against numpy.fft: Stockham passes with bit-reversed DIF butterflies, the paired last pass + Hermitian split
of the row kernel, and the four-step column FFT.  Guards the documented algorithm on CPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import model_kernels as mk  # noqa: E402


@pytest.mark.parametrize("n,plan", [(16, [16]), (32, [8, 4]), (64, [8, 8]), (128, [16, 8]), (256, [16, 16]), (512, [8, 8, 8]), (32, [32])])
def test_stockham_plans_of_the_column_tiles(n, plan):      # kernels_cols.cuh: col_radix()
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert np.abs(mk.fft_stockham(x, plan) - np.fft.fft(x)).max() < 1e-12 * n


@pytest.mark.parametrize("m,prefix", [(32, [2]), (64, [4]), (128, [8]), (256, [16]), (512, [32]), (1024, [8, 8])])
def test_row_kernel_paired_last_pass_and_hermitian_split(m, prefix):   # kernels_rows.cuh: RowPlan<M>
    xr = np.random.default_rng(m).standard_normal(2 * m)
    assert np.abs(mk.r2c_row_model(xr, prefix) - np.fft.rfft(xr)).max() < 1e-11 * m


@pytest.mark.parametrize("m,prefix", [(256, [4, 8]), (1024, [16, 8]), (2048, [16, 16])])
def test_paired_radix8_last_pass(m, prefix):
    """The 16-points-per-thread row plan (paired radix-8 last pass) planned for the 512-thread row kernel."""
    xr = np.random.default_rng(m + 1).standard_normal(2 * m)
    assert np.abs(mk.r2c_row_model(xr, prefix, RL=8) - np.fft.rfft(xr)).max() < 1e-11 * m


@pytest.mark.parametrize("n1,n2,p1,p2", [(16, 16, [16], [16]), (32, 16, [8, 4], [16]), (32, 32, [8, 4], [8, 4])])
def test_four_step_column_fft(n1, n2, p1, p2):             # kernels_cols.cuh: level A / level B
    rng = np.random.default_rng(n1 * n2)
    x = rng.standard_normal(n1 * n2) + 1j * rng.standard_normal(n1 * n2)
    assert np.abs(mk.col_two_level(x, n1, n2, p1, p2) - np.fft.fft(x)).max() < 1e-11 * n1 * n2


# ---------------------------------------------------------------- round-2 kernels
def test_swizzle_is_a_permutation_and_conflict_free():        # kernels_rows_v2.cuh: rv2::pad
    pos = [mk.swizzle(p) for p in range(8192)]
    assert sorted(pos) == list(range(8192))
    # 8 consecutive lanes (one quarter-warp of a 128-bit access) must hit 8 different 16-byte bank groups in each access pattern
    for j1 in (0, 5, 15):
        for r in (0, 3, 15):
            assert len({mk.swizzle(j1 + 16 * u + 512 * r) & 7 for u in range(8)}) == 8            # pass A: lanes = u
            assert len({mk.swizzle(j1 + 16 * r + 512 * s) & 7 for s in range(8)}) == 8            # pass B / final pass: lanes = s
    assert all(mk.swizzle(p) >> 3 == p >> 3 for p in range(8192))                               # stays inside its 128-byte row


@pytest.mark.parametrize("m", [512, 1024])
def test_rows_v2_core(m):                                      # kernels_rows_v2.cuh with 16 x (m/16) instead of 16 x 512
    rng = np.random.default_rng(m)
    z = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    assert np.abs(mk.rows_v2_model(z) - np.fft.fft(z)).max() < 1e-11 * m


@pytest.mark.parametrize("n", [64, 256, 1024])
def test_rows_long2_parked_even_bins_and_pairs(n):             # kernels_rows_long2.cuh
    x = np.random.default_rng(n).standard_normal(n)
    assert np.abs(mk.rows_long2_model(x) - np.fft.rfft(x)).max() < 1e-11 * n
    z = x[0::2] + 1j * x[1::2]
    assert np.abs(mk.herm_split(np.fft.fft(z), n) - np.fft.rfft(x)).max() < 1e-11 * n


@pytest.mark.parametrize("PP", [4, 16, 64])
def test_rows_dit2_combine_covers_every_bin_once(PP):          # kernels_rows_dit2.cuh
    n = 64 * PP
    x = np.random.default_rng(PP).standard_normal(n)
    X, cnt = mk.rows_dit2_model(x, PP)
    assert (cnt == 1).all()
    assert np.abs(X - np.fft.rfft(x)).max() < 1e-11 * n


@pytest.mark.parametrize("C,PP", [(2, 4), (4, 4), (8, 4), (4, 16), (8, 16)])
def test_rows_ditc_combine_covers_every_bin_once(C, PP):      # kernels_rows_ditc.cuh
    n = 32 * C * PP
    x = np.random.default_rng(C * PP).standard_normal(n)
    X, cnt = mk.rows_ditc_model(x, C, PP)
    assert (cnt == 1).all()
    assert np.abs(X - np.fft.rfft(x)).max() < 1e-11 * n


@pytest.mark.parametrize("t,q", [(7, 1), (3, 2), (5, 8), (3, 64), (13, 16), (31, 32), (9, 128)])
def test_mixed_radix_rows(t, q):                               # kernels_generic.cuh: rows_mixed_kernel
    m = t * q
    rng = np.random.default_rng(m)
    z = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    assert np.abs(mk.rows_mixed_model(z, t) - np.fft.fft(z)).max() < 1e-11 * m
    lg = q.bit_length() - 1
    assert sorted(mk.gen_rev(k, q, lg) for k in range(q)) == list(range(q))


@pytest.mark.parametrize("t,q", [(7, 1), (3, 4), (31, 16), (5, 64)])
def test_mixed_radix_columns(t, q):                            # kernels_generic.cuh: cols_odd_kernel + virtual strips
    nx = t * q
    rng = np.random.default_rng(nx + 1)
    y = rng.standard_normal(nx) + 1j * rng.standard_normal(nx)
    assert np.abs(mk.cols_mixed_model(y, t) - np.fft.fft(y)).max() < 1e-11 * nx


def test_column_prestage_split2():                             # kernels_cols.cuh: SPLIT = 2
    rng = np.random.default_rng(3)
    y = rng.standard_normal(512) + 1j * rng.standard_normal(512)
    assert np.abs(mk.cols_split2_model(y, 16, 16) - np.fft.fft(y)).max() < 1e-10


@pytest.mark.parametrize("n", [7, 97, 1021, 8209])
def test_bluestein_chirp_z(n):                                 # kernels_bluestein.cuh
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert np.abs(mk.bluestein_model(x) - np.fft.fft(x)).max() < 1e-10 * np.sqrt(n)
    z = rng.standard_normal(2 * n)                             # r2c of an even length 2n through the half-length complex transform
    zc = z[0::2] + 1j * z[1::2]
    assert np.abs(mk.herm_split(mk.bluestein_model(zc), 2 * n) - np.fft.rfft(z)).max() < 1e-10 * np.sqrt(n)
