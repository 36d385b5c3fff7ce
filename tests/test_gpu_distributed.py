"""Multi-GPU parity: gathered distributed::loop result == shared::loop / oracle result of the gathered
input, natural order, all ny/2+1 columns, for scatter / all_to_all / p2p.
One process per GPU (spawned here), gloo for the bootstrap bytes, NCCL / peer stores for the data."""
import os
import socket

import numpy as np
import pytest

from conftest import gpu_count

pytestmark = pytest.mark.gpu
TOL = 1e-12
CASES = [(64, 64), (16, 256), (256, 512), (1024, 2048), (2048, 64),
         # long rows with P > 1: C = 2 / C = 4 row kernels (non-FAST addressing, several destination ranks),
         # long columns with P > 1: the (256,128) fused pair that BASELINE configs 3/4 launch
         (64, 32768), (32, 65536), (32768, 64), (16, 131072),
         (296, 16384)]   # ny = 16384 with P > 1: rows_r2c_v2_kernel<false>, two full waves of persistent CTAs
PIPELINED_CASES = [(1024, 2048), (2048, 512)]   # HPXFFT_B200_CHUNKS=4: sub-slab pipelined exchange
TRANSPORTS = ["ce", "nccl", "fused"]            # HPXFFT_B200_A2A: transports behind the all_to_all run mode


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
        results = []
        cases = CASES + [(world, 32)]                       # n_x_local == 1
        if os.environ.get("HPXFFT_B200_DIST_CASES") == "fast":   # 8-GPU box time is expensive: one case per kernel family
            cases = [(64, 64), (1024, 2048), (64, 32768), (32768, 64), (296, 16384), (world, 32)]
        # the decimation-in-time long-row kernels are the one-GPU default; with several destination ranks they are opt-in
        # (launch_rows.cu) and stay covered here: ROWS_LONG=3 (ny = 32768), 5 (ny = 65536, 131072)
        forced = [(64, 32768, "3"), (64, 32768, "5"), (32, 65536, "5"), (16, 131072, "5")]
        if os.environ.get("HPXFFT_B200_DIST_CASES") == "rows":   # the long-row kernels with several destination ranks
            cases = [(64, 32768), (32, 65536), (16, 131072), (296, 16384)]
        elif os.environ.get("HPXFFT_B200_DIST_CASES") == "fast":
            forced = forced[:1]
        elif world > 2:                      # keep the larger worlds inside the test's time limit
            forced = [forced[0], forced[2]]
        for case in cases + forced:
            nx, ny = case[0], case[1]
            if len(case) > 2:
                os.environ["HPXFFT_B200_ROWS_LONG"] = case[2]
            else:
                os.environ.pop("HPXFFT_B200_ROWS_LONG", None)
            nxl = nx // world
            full = oracle.make_input(nx, ny, oracle.PATTERN_UNIFORM, seed=11)   # x-dependent input
            ref = oracle.fft_2d_r2c_shared(full, workers=2)
            for comm in ("all_to_all:ce", "all_to_all:nccl", "scatter", "p2p"):
                slab = full[rank * nxl:(rank + 1) * nxl].copy()
                fft = pkg.distributed.loop(device=rank)
                if ":" in comm:
                    os.environ["HPXFFT_B200_A2A"] = comm.split(":")[1]
                fft.initialize(pkg.vector_2d.from_array(slab), comm.split(":")[0], "estimate")
                os.environ.pop("HPXFFT_B200_A2A", None)
                out = fft.fft_2d_r2c().data()
                err = oracle.rel_l2(out, ref[rank * nxl:(rank + 1) * nxl])
                # relative to the whole array's norm so that near-empty slabs do not inflate the figure
                scale = np.linalg.norm(ref[rank * nxl:(rank + 1) * nxl]) / (np.linalg.norm(ref) / np.sqrt(world))
                results.append((nx, ny, comm, err * scale, fft.get_measurement("total")))
                del fft
                dist.barrier()
        os.environ.pop("HPXFFT_B200_ROWS_LONG", None)
        # opt-in pipelined exchange (row chunks / strip chunks on a second stream)
        os.environ["HPXFFT_B200_CHUNKS"] = "4"
        os.environ["HPXFFT_B200_A2A"] = "nccl"
        for (nx, ny) in ([] if os.environ.get("HPXFFT_B200_DIST_CASES") == "rows" else PIPELINED_CASES):
            nxl = nx // world
            full = oracle.make_input(nx, ny, oracle.PATTERN_UNIFORM, seed=12)
            ref = oracle.fft_2d_r2c_shared(full, workers=2)
            for comm in ("all_to_all",):
                fft = pkg.distributed.loop(device=rank)
                fft.initialize(pkg.vector_2d.from_array(full[rank * nxl:(rank + 1) * nxl].copy()), comm, "estimate")
                out = fft.fft_2d_r2c().data()
                err = oracle.rel_l2(out, ref[rank * nxl:(rank + 1) * nxl])
                results.append((nx, ny, comm + "+chunks4", err, fft.get_measurement("total")))
                del fft
                dist.barrier()
        os.environ.pop("HPXFFT_B200_CHUNKS")
        os.environ.pop("HPXFFT_B200_A2A")
        # BASELINE configs 3/4 at full size: 32768 x 32768, device-generated separable input, sampled tiles
        import ctypes as C

        import sampled
        lib = pkg.capi.load()
        nx = ny = 32768
        ref = sampled.SeparableReference(nx, ny, 42)
        for comm in ("all_to_all:ce", "all_to_all:nccl", "scatter", "p2p"):
            if ":" in comm:
                os.environ["HPXFFT_B200_A2A"] = comm.split(":")[1]
            fft = pkg.distributed.loop(device=rank)
            fft.initialize(pkg.vector_2d(nx // world, ny + 2, 0.0), comm.split(":")[0], "estimate")
            os.environ.pop("HPXFFT_B200_A2A", None)
            r = sampled.check_plan(lib, fft.plan_handle(), nx, ny, rank, world, seed=42, ref=ref)
            results.append((nx, ny, comm + "+sampled", (r["num"] / r["den"]) ** 0.5, fft.get_measurement("total")))
            del fft
            dist.barrier()
        q.put((rank, results))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "ERR " + repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.timeout(900)
def test_distributed_equals_shared(world):
    if gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=800) for _ in procs)
    for p in procs:
        p.join(60)
    for rank, rr in res.items():
        assert not isinstance(rr, str), f"rank {rank}: {rr}"
        for (nx, ny, comm, err, total) in rr:
            assert err <= TOL, (rank, nx, ny, comm, err)
            assert total >= 0.0
