"""world_size-2 gloo test of the host-side N>1 plumbing (Bootstrap: unique-id broadcast and
IPC-handle all-gather) and of the slab bookkeeping the multi-GPU parity tests rely on.  CPU only."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    import __graft_entry__ as entry
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pkg = entry.load_package()
        oracle = entry.load_oracle()
        b = pkg.distributed.Bootstrap()
        assert (b.rank, b.size) == (rank, world)
        uid = b.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 0)
        assert uid == bytes(range(128))
        handles = b.all_gather_bytes(bytes([rank]) * 128)
        assert handles == [bytes([r]) * 128 for r in range(world)]
        # slab-consistent synthetic input + natural-order restatement
        nx, ny = 16, 32
        nxl = nx // world
        slab = oracle.make_input(nxl, ny, oracle.PATTERN_UNIFORM, seed=3, row0=rank * nxl)
        gathered = [None] * world
        dist.all_gather_object(gathered, slab)
        full = np.concatenate(gathered)
        assert np.array_equal(full, oracle.make_input(nx, ny, oracle.PATTERN_UNIFORM, seed=3))
        mine = oracle.fft_2d_r2c_distributed(gathered)[rank]
        assert np.array_equal(mine, oracle.fft_2d_r2c_shared(full)[rank * nxl:(rank + 1) * nxl])
        # bad comm flag is collective-free and must not hang
        fft = pkg.distributed.loop(device=-1, bootstrap=b)
        fft.initialize(pkg.vector_2d.from_array(slab.copy()), "gather", "estimate")
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_bootstrap_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(30)
    assert res == {0: "ok", 1: "ok"}, res
