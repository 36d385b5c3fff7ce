"""Lengths with a large prime factor above the direct-DFT bound: Bluestein's chirp-z on top of the power-of-two column kernels
(kernels_bluestein.cuh).  FFTW accepts every length (core/src/util/adapter_fftw.cpp:6-10,24-30).  Needs a GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("nx,ny", [(8209, 64),        # prime column length (convolution length 32768), power-of-two rows
                                   (64, 16418),       # rows: ny/2 = 8209 prime
                                   (8209, 16418),     # both
                                   (10007, 96),       # prime columns x mixed-radix rows
                                   (33, 2 * 4099)])   # 33 rows: a ragged 16-row strip; ny/2 = 4099 prime but ny <= 8192 -> direct DFT
def test_large_prime_lengths(pkg, oracle, nx, ny):
    a = oracle.make_input(nx, ny, oracle.PATTERN_UNIFORM, seed=5)
    fft = pkg.shared.loop(device=0)
    fft.initialize(pkg.vector_2d.from_array(a.copy()), "estimate")
    got = fft.fft_2d_r2c_par().data()
    ref = oracle.fft_2d_r2c_shared(a, workers=8)
    assert oracle.rel_l2(got, ref) <= TOL
