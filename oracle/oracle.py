"""CPU oracle for the HPX-FFT 2-D r2c hot path  --  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.  The product path (hpx-fft_b200/) never does.

What is restated (all citations relative to /root/reference):
  * hpxfft::shared::loop::fft_2d_r2c_par     core/src/shared/loop.cpp:56-113
      phase 1  per-row 1-D r2c of length ny, in place      loop.cpp:6-9, 61-68
      phase 2  transpose  trans(ky, x) = vals(x, ky)         loop.cpp:18-25, 72-80
      phase 3  per-row forward c2c of length nx on trans     loop.cpp:11-15, 83-91
      phase 4  transpose back vals(kx, ky) = trans(ky, kx)   loop.cpp:46-53, 94-102
  * dimension inference of loop::initialize                  loop.cpp:163-165
      dim_c_x = n_row, dim_c_y = n_col/2, dim_r_y = 2*dim_c_y-2
  * hpxfft::distributed::loop::fft_2d_r2c   core/src/distributed/loop.cpp:130-272
      restated in *natural order* (what fftw_mpi_plan_dft_r2c_2d returns,
      examples/fftw/fftw_mpi_omp_2d.cpp:119-120): gather slabs in locality order,
      transform, re-slice rows.  The reference's own L>1 index algebra
      (distributed/loop.cpp:87-127) is stride-permuted and not a 2-D DFT
      (SURVEY.md section 8e); its tests only pin L=1.

The arithmetic itself lives in a third-party dependency that is ABSENT from
/root/reference:  FFTW 3.3.10  (spack-repo/environments/hpxfft_ci.yaml:4), call sites
core/src/util/adapter_fftw.cpp:9 (fftw_plan_dft_r2c_1d), :14 (fftw_execute_dft_r2c),
:28-29 (fftw_plan_dft_1d, FFTW_FORWARD), :34 (fftw_execute_dft).  FFTW's published
definition is restated here: forward, unnormalised
    Y[k] = sum_j x[j] * exp(-2 pi i j k / n),   r2c keeps k = 0..n/2.
pocketfft (scipy.fft) stands in for FFTW.

PINNING: the oracle is checked (tests/test_oracle.py) against the only known-answer
vector the reference's tests hold for this path: 4x6 rows [1,2,3,4,0,0] ->
row0 [40,0,-8,8,-8,0], rest 0, exact equality (test/src/test_shared_loop.cpp:15-34,53;
test_distributed_loop.cpp:17-48).  Beyond that single vector the reference's own tests
leave parity unpinned; we add the analytic ramp closed form (the reference example's input,
examples/hpxfft/shared_loop_2d.cpp:33-40) and a long-double cross-check.
"""
from __future__ import annotations

import numpy as np

try:  # scipy's pocketfft supports long double and worker threads
    import scipy.fft as _sfft
except Exception:  # pragma: no cover
    _sfft = None

MASK64 = (1 << 64) - 1


# --------------------------------------------------------------------------------------
# deterministic synthetic inputs (shared definition with the CUDA fill kernel,
# hpx-fft_b200/csrc/hpxfft_b200.cu : fill_kernel)
# --------------------------------------------------------------------------------------
def splitmix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wraps mod 2^64)."""
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def u64_to_unit(x: np.ndarray) -> np.ndarray:
    """top 53 bits -> double in [-1, 1)."""
    return (x >> np.uint64(11)).astype(np.float64) * (2.0 ** -52) - 1.0


PATTERN_RAMP = 0       # v(i, j) = j           (examples/hpxfft/shared_loop_2d.cpp:33-40)
PATTERN_UNIFORM = 1    # v(i, j) = u(splitmix64(seed ^ (i_global*ny + j)))
PATTERN_SEPARABLE = 2  # v(i, j) = sum_r a_r(i) b_r(j), rank 4


def sep_vectors(nx: int, ny: int, seed: int, rank: int = 4):
    a = np.empty((rank, nx))
    b = np.empty((rank, ny))
    for r in range(rank):
        sa = np.uint64((seed + 1000 + r) & MASK64)
        sb = np.uint64((seed + 2000 + r) & MASK64)
        a[r] = u64_to_unit(splitmix64(np.arange(nx, dtype=np.uint64) ^ sa))
        b[r] = u64_to_unit(splitmix64(np.arange(ny, dtype=np.uint64) ^ sb))
    return a, b


def make_input(nx_local: int, ny: int, pattern: int, seed: int = 42, row0: int = 0,
               nx_global: int | None = None) -> np.ndarray:
    """Padded slab  nx_local x (ny+2)  (pads = 0), rows row0..row0+nx_local of the global array."""
    n_col = 2 * (ny // 2 + 1)
    v = np.zeros((nx_local, n_col), dtype=np.float64)
    if pattern == PATTERN_RAMP:
        v[:, :ny] = np.arange(ny, dtype=np.float64)[None, :]
    elif pattern == PATTERN_UNIFORM:
        i = np.arange(row0, row0 + nx_local, dtype=np.uint64)[:, None]
        j = np.arange(ny, dtype=np.uint64)[None, :]
        with np.errstate(over="ignore"):
            ctr = i * np.uint64(ny) + j
        v[:, :ny] = u64_to_unit(splitmix64(ctr ^ np.uint64(seed & MASK64)))
    elif pattern == PATTERN_SEPARABLE:
        nxg = nx_global if nx_global is not None else row0 + nx_local
        a, b = sep_vectors(nxg, ny, seed)
        v[:, :ny] = a[:, row0:row0 + nx_local].T @ b
    else:
        raise ValueError("unknown pattern")
    return v


# --------------------------------------------------------------------------------------
# the reference algorithm, phase by phase
# --------------------------------------------------------------------------------------
def infer_dims(n_row: int, n_col: int, n_localities: int = 1):
    """core/src/shared/loop.cpp:163-165, core/src/distributed/loop.cpp:284-287."""
    dim_c_x = n_row * n_localities
    dim_c_y = n_col // 2
    dim_r_y = 2 * dim_c_y - 2
    return dim_c_x, dim_c_y, dim_r_y


def fft_2d_r2c_shared(values: np.ndarray, workers: int = 1, dtype=np.float64,
                      timings: dict | None = None) -> np.ndarray:
    """4-phase restatement of shared::loop::fft_2d_r2c_par (loop.cpp:56-113).

    values: (n_row, n_col) padded real array.  Returns a new padded array holding the
    interleaved (re, im) Hermitian half, same layout as the reference's in-place result.
    """
    import time
    cdt = np.complex128 if dtype == np.float64 else np.clongdouble
    vals = np.array(values, dtype=dtype, copy=True)
    n_row, n_col = vals.shape
    dim_c_x, dim_c_y, dim_r_y = infer_dims(n_row, n_col)
    kw = {"workers": workers} if dtype == np.float64 else {}
    t0 = time.perf_counter()
    # phase 1: per-row r2c, in place (loop.cpp:6-9)
    y = _sfft.rfft(vals[:, :dim_r_y], axis=1, **kw)           # (nx, cy)
    vals_c = vals.view(cdt)                                    # (nx, cy) view on the padded rows
    vals_c[:, :] = y
    t1 = time.perf_counter()
    # phase 2: trans(ky, x) = vals(x, ky)  (loop.cpp:18-25)
    trans = np.ascontiguousarray(vals_c.T)                     # (cy, nx)
    t2 = time.perf_counter()
    # phase 3: forward c2c of length nx on every trans row (loop.cpp:11-15)
    trans = _sfft.fft(trans, axis=1, **kw)
    t3 = time.perf_counter()
    # phase 4: vals(kx, ky) = trans(ky, kx)  (loop.cpp:46-53)
    vals_c[:, :] = trans.T
    t4 = time.perf_counter()
    if timings is not None:
        timings.update(total=t4 - t0, first_fftw=t1 - t0, first_trans=t2 - t1,
                       second_fftw=t3 - t2, second_trans=t4 - t3)
    return vals


def fft_2d_r2c_longdouble(values: np.ndarray) -> np.ndarray:
    """Same algorithm in 80-bit long double; returned as longdouble padded array."""
    return fft_2d_r2c_shared(values, dtype=np.longdouble)


def fft_2d_r2c_distributed(slabs: list[np.ndarray], workers: int = 1) -> list[np.ndarray]:
    """Natural-order distributed result: gather slabs (locality order), transform, re-slice.
    Equals distributed::loop for L=1 bit-for-bit in layout (distributed/loop.cpp:130-272)."""
    full = np.concatenate(slabs, axis=0)
    out = fft_2d_r2c_shared(full, workers=workers)
    nxl = slabs[0].shape[0]
    return [out[r * nxl:(r + 1) * nxl] for r in range(len(slabs))]


# --------------------------------------------------------------------------------------
# size-independent oracles
# --------------------------------------------------------------------------------------
def ramp_analytic(nx: int, ny: int) -> np.ndarray:
    """Closed form for v(i,j)=j: row 0 = nx*(-ny/2 + i (ny/2) cot(pi k/ny)), Z[0,0]=nx ny (ny-1)/2."""
    cy = ny // 2 + 1
    out = np.zeros((nx, 2 * cy), dtype=np.longdouble)
    k = np.arange(1, cy, dtype=np.longdouble)
    pi = np.longdouble(np.pi) if np.finfo(np.longdouble).eps > 1e-17 else \
        np.longdouble("3.14159265358979323846264338327950288")
    out[0, 0] = np.longdouble(nx) * ny * (ny - 1) / 2
    out[0, 2::2] = np.longdouble(nx) * (-np.longdouble(ny) / 2)
    with np.errstate(divide="ignore"):
        cot = np.cos(pi * k / ny) / np.sin(pi * k / ny)
    out[0, 3::2] = np.longdouble(nx) * (np.longdouble(ny) / 2) * cot
    if ny % 2 == 0:
        out[0, 2 * (cy - 1) + 1] = 0.0  # Nyquist bin is real
    return out


def separable_spectrum(nx: int, ny: int, seed: int, rows: slice | None = None,
                       cols: slice | None = None) -> np.ndarray:
    """Z[kx, ky] = sum_r FFT(a_r)[kx] * rFFT(b_r)[ky]  from 1-D long-double transforms.
    Returns complex longdouble block for the requested (kx, ky) window."""
    a, b = sep_vectors(nx, ny, seed)
    fa = _sfft.fft(a.astype(np.longdouble), axis=1)            # (rank, nx)
    fb = _sfft.rfft(b.astype(np.longdouble), axis=1)           # (rank, cy)
    rows = rows or slice(None)
    cols = cols or slice(None)
    return np.einsum("rx,ry->xy", fa[:, rows], fb[:, cols])


def to_complex(padded: np.ndarray) -> np.ndarray:
    """(n_row, 2*cy) interleaved reals -> (n_row, cy) complex (same precision family)."""
    cdt = np.complex128 if padded.dtype == np.float64 else np.clongdouble
    return np.ascontiguousarray(padded).view(cdt)


def rel_l2(got: np.ndarray, ref: np.ndarray) -> float:
    g = np.asarray(got, dtype=np.longdouble).ravel()
    r = np.asarray(ref, dtype=np.longdouble).ravel()
    den = np.sqrt(np.sum(r * r))
    num = np.sqrt(np.sum((g - r) ** 2))
    return float(num / den) if den > 0 else float(num)


GOLDEN_4x4_IN = np.array([[1.0, 2.0, 3.0, 4.0, 0.0, 0.0]] * 4)
GOLDEN_4x4_OUT = np.zeros((4, 6))
GOLDEN_4x4_OUT[0] = [40.0, 0.0, -8.0, 8.0, -8.0, 0.0]
