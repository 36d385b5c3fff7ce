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
