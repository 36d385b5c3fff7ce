"""hpxfft::util::vector_2d<double> mirror (core/include/hpxfft/util/vector_2d.hpp).

Row-major owning 2-D array with the reference's accessor names.  Storage is a NumPy float64 array
(`values_`), optionally page-locked so that the host<->device hand-over runs at PCIe speed."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class vector_2d:
    """vector_2d(n_row, n_col[, v])  --  vector_2d.hpp:97-124"""

    def __init__(self, n_row: int = 0, n_col: int = 0, v: float = 0.0, pinned: bool = False, _data=None):
        self.n_row_ = int(n_row)
        self.n_col_ = int(n_col)
        self.size_ = self.n_row_ * self.n_col_
        self._pinned_ptr = None
        if _data is not None:
            arr = np.ascontiguousarray(_data, dtype=np.float64)
            assert arr.shape == (self.n_row_, self.n_col_)
            self.values_ = arr
        elif pinned and self.size_ > 0:
            lib = capi.load()
            ptr = lib.hpxfft_b200_host_alloc(self.size_ * 8)
            if not ptr:
                raise capi.Hpxfft_b200Error(capi.ECUDA, lib.hpxfft_b200_last_error().decode())
            self._pinned_ptr = ptr
            buf = (C.c_double * self.size_).from_address(ptr)
            self.values_ = np.frombuffer(buf, dtype=np.float64).reshape(self.n_row_, self.n_col_)
            self.values_[...] = v
        else:
            self.values_ = np.full((self.n_row_, self.n_col_), float(v), dtype=np.float64)

    @classmethod
    def from_array(cls, a) -> "vector_2d":
        a = np.asarray(a, dtype=np.float64)
        return cls(a.shape[0], a.shape[1], _data=a)

    def __del__(self):
        ptr, self._pinned_ptr = getattr(self, "_pinned_ptr", None), None
        if ptr:
            self.values_ = None
            try:
                capi.load().hpxfft_b200_host_free(ptr)
            except Exception:
                pass

    # accessors (vector_2d.hpp:198-275)
    def __call__(self, i: int, j: int) -> float:
        return float(self.values_[i, j])

    def at(self, i: int, j: int) -> float:
        if i * self.n_col_ + j >= self.size_ or i < 0 or j < 0:
            raise RuntimeError("out of range exception")  # vector_2d.hpp:216-221
        return float(self.values_.reshape(-1)[i * self.n_col_ + j])

    def set(self, i: int, j: int, v: float) -> None:
        self.values_[i, j] = v

    def row(self, i: int) -> np.ndarray:
        return self.values_[i]

    def data(self) -> np.ndarray:
        return self.values_

    def size(self) -> int:
        return self.size_

    def n_row(self) -> int:
        return self.n_row_

    def n_col(self) -> int:
        return self.n_col_

    def __eq__(self, other) -> bool:  # exact comparison, vector_2d.hpp:277-294
        if not isinstance(other, vector_2d):
            return NotImplemented
        if self.n_row_ != other.n_row_ or self.n_col_ != other.n_col_:
            return False
        return bool(np.array_equal(self.values_, other.values_))

    def __repr__(self) -> str:
        return f"vector_2d({self.n_row_}, {self.n_col_})"


def print_vector_2d(v: vector_2d) -> None:
    """core/include/hpxfft/util/print_vector.hpp:10-38: rows of '(re im) ' pairs."""
    for i in range(v.n_row()):
        row = v.row(i)
        print("".join(f"({row[j]:g} {row[j + 1]:g}) " for j in range(0, v.n_col() - 1, 2)))
    print()
