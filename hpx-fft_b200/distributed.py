"""hpxfft::distributed::loop / agas mirrors (core/include/hpxfft/distributed/loop.hpp:17-31,
agas.hpp:13-29) over the C ABI.  SPMD: one process per GPU; the few hundred bytes of bootstrap
(NCCL unique id, IPC handles) travel through torch.distributed, which stands in for HPX's AGAS."""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import Future, ThreadPoolExecutor

import numpy as np

from . import capi
from .shared import check_plan_flag
from .util import vector_2d


class Bootstrap:
    """this_locality / num_localities + tiny host collectives (hpx::get_locality_id,
    hpx::get_num_localities, core/src/distributed/loop.cpp:281-282)."""

    def __init__(self):
        self.rank, self.size, self._dist = 0, 1, None
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._dist = dist
                self.rank, self.size = dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass

    def broadcast_bytes(self, data: bytes | None, root: int = 0) -> bytes:
        if self.size == 1:
            return data
        box = [data if self.rank == root else None]
        self._dist.broadcast_object_list(box, src=root)
        return box[0]

    def all_gather_bytes(self, data: bytes) -> list[bytes]:
        if self.size == 1:
            return [data]
        out = [None] * self.size
        self._dist.all_gather_object(out, data)
        return out


def default_device(rank: int) -> int:
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    n = capi.load().hpxfft_b200_device_count()
    return rank % n if n else -1


class loop:
    """Slab-decomposed 2-D r2c FFT: every locality passes its n_x_local x n_col slab."""

    def __init__(self, device: int | None = None, bootstrap: Bootstrap | None = None):
        self._lib = capi.load()
        self._plan = C.c_void_p()
        self._values = None
        self._boot = bootstrap or Bootstrap()
        self._device = default_device(self._boot.rank) if device is None else device
        self._comm_ok = False

    def initialize(self, values_vec: vector_2d, COMM_FLAG: str, PLAN_FLAG: str) -> None:
        check_plan_flag(PLAN_FLAG)
        self._destroy()
        self._values = values_vec
        self._comm_ok = COMM_FLAG in ("scatter", "all_to_all", "p2p")
        if not self._comm_ok:
            # core/src/distributed/loop.cpp:342-346: message, no exception
            print("Specify communication scheme: scatter or all_to_all")
            return
        b = self._boot
        uid = None
        if b.size > 1:
            raw = None
            if b.rank == 0:
                buf = C.create_string_buffer(capi.UNIQUE_ID_BYTES)
                capi.check(self._lib.hpxfft_b200_get_unique_id(buf))
                raw = buf.raw
            uid = b.broadcast_bytes(raw, 0)
        capi.check(self._lib.hpxfft_b200_create(C.byref(self._plan), values_vec.n_row(), values_vec.n_col(), b.rank, b.size,
                                                self._device, COMM_FLAG.encode(), PLAN_FLAG.encode(), uid))
        cnt = self._lib.hpxfft_b200_ipc_count(self._plan)  # peer windows (copy-engine / fused transports)
        if cnt > 0:
            buf = C.create_string_buffer(cnt * capi.IPC_HANDLE_BYTES)
            capi.check(self._lib.hpxfft_b200_ipc_export(self._plan, buf))
            allh = b"".join(b.all_gather_bytes(buf.raw))
            capi.check(self._lib.hpxfft_b200_ipc_import(self._plan, allh))
        capi.check(self._lib.hpxfft_b200_upload(self._plan, values_vec.data().ctypes.data))

    def fft_2d_r2c(self) -> vector_2d:
        if self._values is None:
            raise RuntimeError("loop: initialize() must be called before fft_2d_r2c")
        out, self._values = self._values, None
        if not self._comm_ok:
            print("Communication scheme not specified during initialization")  # distributed/loop.cpp:175-179
            return out
        capi.check(self._lib.hpxfft_b200_execute(self._plan))
        capi.check(self._lib.hpxfft_b200_download(self._plan, out.data().ctypes.data))
        return out

    def get_measurement(self, name: str) -> float:
        if not self._plan:
            return 0.0
        return float(self._lib.hpxfft_b200_measurement(self._plan, name.encode()))

    def plan_handle(self) -> C.c_void_p:
        return self._plan

    def _destroy(self) -> None:
        if self._plan:
            self._lib.hpxfft_b200_destroy(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass


class agas:
    """Client surface of hpxfft::distributed::agas (agas.hpp:21-26): the same two calls returning
    futures.  The per-row HPX action DAG behind it is a CPU task-granularity experiment and is out
    of scope; the future wraps the GPU path of `loop` on a worker thread."""

    def __init__(self, device: int | None = None, bootstrap: Bootstrap | None = None):
        self._loop = loop(device, bootstrap)
        self._pool = ThreadPoolExecutor(max_workers=1)

    def initialize(self, values_vec: vector_2d, COMM_FLAG: str, PLAN_FLAG: str) -> Future:
        check_plan_flag(PLAN_FLAG)  # invalid flag throws at the call site like the by-value action would
        return self._pool.submit(self._loop.initialize, values_vec, COMM_FLAG, PLAN_FLAG)

    def fft_2d_r2c(self) -> Future:
        return self._pool.submit(self._loop.fft_2d_r2c)

    def get_measurement(self, name: str) -> float:
        return self._loop.get_measurement(name)
