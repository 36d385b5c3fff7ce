"""Tile-sampled parity check for slabs that are too large to bring back whole -- TEST INFRASTRUCTURE ONLY
(the checker; used by tests/ and by the parity leg of bench.py, never by the product path).

The device fills its slab with the rank-4 separable pattern (HPXFFT_B200_PATTERN_SEPARABLE, same
definition as oracle.make_input), transforms it, and a handful of tiles per rank are compared with the
closed form  Z[kx, ky] = sum_r FFT(a_r)[kx] * rFFT(b_r)[ky]  evaluated from 1-D long-double transforms
(SURVEY.md 8c item 4).  The pattern is x-dependent and dense, so a wrong exchange order, a dropped column
or a mis-placed tile shows up in every sampled block."""
from __future__ import annotations

import ctypes as C

import numpy as np

import oracle


class SeparableReference:
    def __init__(self, nx: int, ny: int, seed: int = 42):
        import scipy.fft as sfft
        self.nx, self.ny, self.cy = nx, ny, ny // 2 + 1
        a, b = oracle.sep_vectors(nx, ny, seed)
        self.fa = sfft.fft(a.astype(np.longdouble), axis=1)     # (4, nx)
        self.fb = sfft.rfft(b.astype(np.longdouble), axis=1)    # (4, cy)

    def block(self, kx0: int, nkx: int, ky0: int, nky: int) -> np.ndarray:
        return np.einsum("rx,ry->xy", self.fa[:, kx0:kx0 + nkx], self.fb[:, ky0:ky0 + nky])


def tile_list(nxl: int, cy: int, th: int = 32, tw: int = 64):
    """(row0, nrows, col0, ncols) in local rows / complex columns: corners, middles and odd offsets,
    always including column 0 and the Nyquist column cy-1."""
    th, tw = min(th, nxl), min(tw, cy)

    def clip(v, lim):
        return max(0, min(v, lim))
    rows = sorted({0, clip(nxl // 2 - 3, nxl - th), clip((2 * nxl) // 3 + 1, nxl - th), nxl - th})
    cols = sorted({0, clip(cy // 3 + 5, cy - tw), clip((5 * cy) // 8 - 7, cy - tw), cy - tw})
    return [(r, th, c, tw) for r in rows for c in cols]


def check_plan(lib, plan, nx: int, ny: int, rank: int, world: int, seed: int = 42, ref: SeparableReference | None = None) -> dict:
    """Fill (separable) -> execute -> compare sampled tiles.  Collective when world > 1 (every rank calls).
    Returns {"num": sum |got-ref|^2, "den": sum |ref|^2, "tiles": n, "max_tile_rel": worst single tile}."""
    def ck(rc):
        if rc != 0:
            raise RuntimeError(lib.hpxfft_b200_last_error().decode())
    ref = ref or SeparableReference(nx, ny, seed)
    nxl, cy = nx // world, ny // 2 + 1
    ck(lib.hpxfft_b200_fill(plan, oracle.PATTERN_SEPARABLE, seed))
    ck(lib.hpxfft_b200_execute(plan))
    num = den = 0.0
    worst = 0.0
    tiles = tile_list(nxl, cy)
    for (r0, nr, c0, nc) in tiles:
        got = np.empty((nr, 2 * nc))
        ck(lib.hpxfft_b200_download_tile(plan, r0, nr, 2 * c0, 2 * nc, got.ctypes.data_as(C.c_void_p)))
        z = got.view(np.complex128).astype(np.clongdouble)
        want = ref.block(rank * nxl + r0, nr, c0, nc)
        d = z - want
        n_, d_ = float(np.vdot(d, d).real), float(np.vdot(want, want).real)
        num += n_
        den += d_
        worst = max(worst, (n_ / d_) ** 0.5 if d_ > 0 else n_ ** 0.5)
    return {"num": num, "den": den, "tiles": len(tiles), "max_tile_rel": worst}
