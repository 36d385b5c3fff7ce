#!/usr/bin/env python
"""bench.py -- headline benchmark of the HPX-FFT 2-D r2c hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host cores

A "step" is one 2-D r2c transform of one synthetic array.
  N = 1 : BASELINE.json configs[1]  16384 x 16384 FP64 through the shared::loop path.
  N > 1 : BASELINE.json configs[2]  32768 x 32768 FP64 slab-decomposed, distributed::loop all_to_all
          (strong scaling over N = 2, 4, 8); `anchor_n1` in the line is the same workload on ONE GPU.
  --config c5 : BASELINE.json configs[4]  131072 x 131072 on 8 GPUs (device-generated input, no host copy).
metric  = GFLOP/s with the BASELINE flop convention 2.5 * N * log2(N), N = nx * ny;  ms_per_step beside it.
value   = device-resident throughput: K transforms enqueued back to back, CUDA events on the plan's
          own stream, barrier + synchronize on both sides, max over ranks.
e2e     = the same metric through the reference-facing call with HOST buffers: every step copies the
          slab host->device from pinned memory, transforms, copies the result back.
roofline= dominant kernel: algorithmic bytes per launch / its average CUDA-event duration over the
          same timed region, against MEASURED_PEAKS.json hbm_gbs; at N > 1 also the NVLink roofline of the
          two exchanges (algorithmic bytes per GPU and direction / measured time, against 770 GB/s).
parity  = after the timed loop: separable input generated on the device, one transform, sampled tiles per
          rank against the closed form (oracle/sampled.py -- the checker, not the product); rc != 0 above 1e-12.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
NVLINK_PEAK_GBS = 770.0    # measured peer copy per direction per GPU (B200_PROFILING.md); 900 nominal
PARITY_TOL = 1e-12
METRIC = "2D r2c FFT GFLOP/s (2.5*N*log2N)"


def flops(nx: int, ny: int) -> float:
    n = float(nx) * float(ny)
    return 2.5 * n * math.log2(n)


def algorithmic_bytes(nx: int, ny: int) -> dict:
    """one read + one write per dimension pass.  Per kernel:
    rows  : read 8*nx*ny (reals)        + write 16*nx*cy
    cols  : read 16*nx*cy               + write 16*nx*cy   (whole column pass, however many launches)"""
    cy = ny // 2 + 1
    return {"rows": 8.0 * nx * ny + 16.0 * nx * cy, "cols": 32.0 * nx * cy, "total": 8.0 * nx * ny + 48.0 * nx * cy}


def exchange_bytes(nx: int, ny: int, world: int, rank: int = 0) -> float:
    """bytes one GPU sends (= receives) per exchange: 16 * nxl * (cy - w_rank)."""
    cy = ny // 2 + 1
    wq0 = cy // world
    w = cy - (world - 1) * wq0 if rank == world - 1 else wq0
    return 16.0 * (nx // world) * (cy - w)


def ncu_traffic(kernel: str, nx: int, ny: int, world: int):
    """dram__bytes_read + dram__bytes_write per launch of `kernel` from the committed ncu --set full captures
    (profiles/ncu_traffic.json), or None when no capture exists for this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        entries = t.get("workloads", [t])
        for e in entries:
            if e.get("workload") == f"{nx}x{ny}" and int(e.get("n_gpus", 1)) == world:
                k = e["kernels"][kernel]
                return k["dram_bytes_read"] + k["dram_bytes_write"]
    except Exception:
        pass
    return None


def measured_peaks() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def workload(args, world: int):
    if args.nx and args.ny:
        return args.nx, args.ny
    if args.config == "c5":
        return 131072, 131072
    if args.config == "c1":
        return 256, 16384
    return (16384, 16384) if world == 1 else (32768, 32768)


def config_dict(nx: int, ny: int, world: int, comm: str | None) -> dict:
    """The SAME dict for both arms (the driver compares them)."""
    nxl = nx // world
    slab_mib = nxl * (ny + 2) * 8 / 2**20
    return {"workload": f"{nx}x{ny} FP64 r2c 2-D FFT, " + ("shared::loop path, 1 locality" if world == 1 else
                        f"distributed::loop {comm}, {world} localities, slab {nxl}x{ny + 2}"),
            "nx": nx, "ny": ny, "run": comm or "par", "plan": "estimate",
            "input": "uniform(-1,1) splitmix64 seed 42",
            "l2_hygiene": f"inputs larger than L2 ({slab_mib:.0f} MiB slab per locality vs 126 MiB L2)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.gpu, self.proc, self.lines, self.thr = gpu_index, None, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
                power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup():
    """torchrun env -> (rank, world, local_rank, dist or None).  gloo group for the bootstrap bytes
    and the max-over-ranks reduction; the data path is NCCL / peer copies inside the library."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world == 1:
        return 0, 1, 0, None
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # keep stdout to the one JSON line: NCCL's "NCCL version ..." banner goes to a file instead
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/hpxfft_b200_nccl_%h_%p.log")
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    return rank, world, local, dist


def run_ours(args) -> dict | None:
    rank, world, local, dist = dist_setup()
    import numpy as np
    import torch
    pkg = entry.load_package()
    lib = pkg.capi.load()
    if lib.hpxfft_b200_device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device; hpxfft_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    # host staging buffers (e2e leg) live on the NUMA node of this rank's GPU
    try:
        orig_affinity = os.sched_getaffinity(0)
    except Exception:
        orig_affinity = None
    numa_bound = lib.hpxfft_b200_bind_host_to_device(local) == 0

    nx, ny = workload(args, world)
    comm = None if world == 1 else args.run
    nxl = nx // world
    boot = pkg.distributed.Bootstrap()
    big = nxl * (ny + 2) * 8 > 6 * 2**30      # C5-class slabs: no host staging, fewer plans

    def make_plan(nx_, ny_, world_=world, rank_=rank, comm_=comm):
        plan = C.c_void_p()
        uid = None
        if world_ > 1:
            raw = None
            if rank_ == 0:
                buf = C.create_string_buffer(pkg.capi.UNIQUE_ID_BYTES)
                pkg.capi.check(lib.hpxfft_b200_get_unique_id(buf))
                raw = buf.raw
            uid = boot.broadcast_bytes(raw, 0)
        rc = lib.hpxfft_b200_create(C.byref(plan), nx_ // world_, ny_ + 2, rank_, world_, local, comm_.encode() if comm_ else None,
                                    b"estimate", uid)
        if rc != 0:
            return None, lib.hpxfft_b200_last_error().decode()
        cnt = lib.hpxfft_b200_ipc_count(plan)
        if cnt > 0:
            buf = C.create_string_buffer(cnt * pkg.capi.IPC_HANDLE_BYTES)
            pkg.capi.check(lib.hpxfft_b200_ipc_export(plan, buf))
            pkg.capi.check(lib.hpxfft_b200_ipc_import(plan, b"".join(boot.all_gather_bytes(buf.raw))))
        return plan, ""

    plan, err = make_plan(nx, ny)
    if plan is None:
        raise RuntimeError(err)
    transport = lib.hpxfft_b200_transport(plan).decode()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    stream = torch.cuda.ExternalStream(lib.hpxfft_b200_stream(plan), device=local)
    pkg.capi.check(lib.hpxfft_b200_fill(plan, pkg.capi.PATTERN_UNIFORM, 42))

    # ---- device-resident timing ------------------------------------------------------------
    for _ in range(args.warmup):
        pkg.capi.check(lib.hpxfft_b200_execute_async(plan))
    pkg.capi.check(lib.hpxfft_b200_synchronize(plan))
    pkg.capi.check(lib.hpxfft_b200_fill(plan, pkg.capi.PATTERN_UNIFORM, 42))  # keep magnitudes finite
    pkg.capi.check(lib.hpxfft_b200_reset_timers(plan))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        pkg.capi.check(lib.hpxfft_b200_execute_async(plan))
    e1.record(stream)
    pkg.capi.check(lib.hpxfft_b200_synchronize(plan))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    keys = ("total", "first_fftw", "first_comm", "second_fftw", "second_comm", "second_trans", "rows_kernel",
            "cols_levelA_kernel", "cols_levelB_kernel", "cols_kernel", "first_comm_span", "second_comm_span", "timer_samples")
    meas = {k: lib.hpxfft_b200_measurement(plan, k.encode()) for k in keys}
    launches = lib.hpxfft_b200_launches_per_execute(plan) * args.steps

    # ---- parity of THIS plan (the kernels and the exchange that were just timed) -----------------
    oracle = entry.load_oracle()   # the checker
    import sampled                 # oracle/sampled.py
    chk = sampled.check_plan(lib, plan, nx, ny, rank, world, seed=42) if not args.no_parity else {"num": 0.0, "den": 1.0, "tiles": 0, "max_tile_rel": 0.0}
    par = torch.tensor([chk["num"], chk["den"], float(chk["tiles"])], dtype=torch.float64)
    worst = torch.tensor([chk["max_tile_rel"]], dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(par, op=dist.ReduceOp.SUM)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    parity = {"rel_l2": float((par[0] / par[1]) ** 0.5), "max_tile_rel_l2": float(worst[0]), "tiles": int(par[2]), "tol": PARITY_TOL,
              "input": "separable rank-4, generated on device, seed 42", "reference": "closed form from 1-D long-double FFTs (oracle/sampled.py)"}

    # ---- exchanges alone: the NVLink roofline of the transport --------------------------------------
    xalone = {}
    if world > 1 and transport in ("nccl", "nccl-rooted", "nccl-pipelined", "copy-engine"):
        for which in (1, 2):
            ms = C.c_double(0.0)
            barrier()
            pkg.capi.check(lib.hpxfft_b200_bench_exchange(plan, which, 5, C.byref(ms)))
            xalone[which] = ms.value

    # ---- end to end: host buffers through the reference-facing call -----------------------------
    # Every step = H2D of that step's slab from pinned host memory + transform + D2H of the result.
    # Two plans / two pinned slabs are driven alternately with hpxfft_b200_transform_async, so the D2H of
    # step i overlaps the H2D of step i+1 (PCIe is full duplex); every step still moves its own bytes.
    e2e = None
    slab_bytes = nxl * (ny + 2) * 8
    if not big and not args.no_e2e:
        e2e_steps = max(2, min(args.steps, args.e2e_steps))
        plan2, err2 = make_plan(nx, ny)
        if plan2 is None:
            raise RuntimeError(err2)
        plans = [plan, plan2]
        hosts = [pkg.vector_2d(nxl, ny + 2, 0.0, pinned=True) for _ in plans]
        for i, h in enumerate(hosts):
            h.data()[:, :ny] = np.random.default_rng(1234 + rank + 100 * i).uniform(-1, 1, (nxl, ny))
        for pl, h in zip(plans, hosts):  # warm-up (page-touch, clocks)
            pkg.capi.check(lib.hpxfft_b200_transform(pl, h.data().ctypes.data))
        for h in hosts:                  # keep magnitudes finite over the timed steps
            h.data()[:, ny:] = 0.0
            h.data()[:, :ny] = np.random.default_rng(99).uniform(-1, 1, (nxl, ny))
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            pl, h = plans[i % 2], hosts[i % 2]
            if i >= 2:
                pkg.capi.check(lib.hpxfft_b200_synchronize(pl))   # this slab's previous round trip is complete
            pkg.capi.check(lib.hpxfft_b200_transform_async(pl, h.data().ctypes.data))
        for pl in plans:
            pkg.capi.check(lib.hpxfft_b200_synchronize(pl))
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        lib.hpxfft_b200_destroy(plan2)
        e2e = (e2e_ms, e2e_steps)

    # ---- the same workload on ONE GPU (anchor of the strong-scaling curve), rank 0 only -------------
    anchor = None
    if world > 1 and rank == 0 and not big and not args.no_anchor:
        p1, err1 = make_plan(nx, ny, world_=1, rank_=0, comm_=None)
        if p1 is not None:
            pkg.capi.check(lib.hpxfft_b200_fill(p1, pkg.capi.PATTERN_UNIFORM, 42))
            for _ in range(2):
                pkg.capi.check(lib.hpxfft_b200_execute_async(p1))
            pkg.capi.check(lib.hpxfft_b200_synchronize(p1))
            pkg.capi.check(lib.hpxfft_b200_fill(p1, pkg.capi.PATTERN_UNIFORM, 42))
            pkg.capi.check(lib.hpxfft_b200_reset_timers(p1))
            s1 = torch.cuda.ExternalStream(lib.hpxfft_b200_stream(p1), device=local)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(s1)
            for _ in range(5):
                pkg.capi.check(lib.hpxfft_b200_execute_async(p1))
            a1.record(s1)
            pkg.capi.check(lib.hpxfft_b200_synchronize(p1))
            ams = a0.elapsed_time(a1) / 5
            anchor = {"ms": ams, "value": flops(nx, ny) / (ams * 1e-3) / 1e9, "unit": "GFLOP/s", "steps": 5,
                      "rows_ms": lib.hpxfft_b200_measurement(p1, b"rows_kernel") * 1e3,
                      "cols_ms": lib.hpxfft_b200_measurement(p1, b"cols_kernel") * 1e3,
                      "what": f"{nx}x{ny} shared::loop on one GPU of the same box"}
            lib.hpxfft_b200_destroy(p1)
        else:
            anchor = {"unavailable": err1}

    # ---- reductions: max over ranks -----------------------------------------------------------------
    ms_step = ms_total / args.steps
    e2e_ms = e2e[0] if e2e else 0.0
    if dist is not None:
        t = torch.tensor([ms_step, e2e_ms] + [xalone.get(1, 0.0), xalone.get(2, 0.0)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms = float(t[0]), float(t[1])
        if xalone:
            xalone = {1: float(t[2]), 2: float(t[3])}
        km = torch.tensor([meas[k] for k in sorted(meas)], dtype=torch.float64)
        dist.all_reduce(km, op=dist.ReduceOp.MAX)
        meas = dict(zip(sorted(meas), [float(x) for x in km]))

    result = None
    if rank == 0:
        gf = flops(nx, ny)
        ab = algorithmic_bytes(nx, ny)
        peak, how = measured_peaks()
        kernels = {"rows_r2c": (meas["rows_kernel"], ab["rows"] / world),
                   "cols_c2c": (meas["cols_kernel"], ab["cols"] / world)}
        per_kernel = {k: {"ms": v[0] * 1e3, "algorithmic_bytes": v[1], "achieved_gbs": (v[1] / v[0] / 1e9 if v[0] > 0 else 0.0),
                          "frac": (v[1] / v[0] / 1e9 / peak if v[0] > 0 else 0.0),
                          "traffic": ncu_traffic(k, nx, ny, world)} for k, v in kernels.items()}
        dom = max(kernels, key=lambda k: kernels[k][0])
        dsec, dbytes = kernels[dom]
        achieved = dbytes / dsec / 1e9 if dsec > 0 else 0.0
        compute_sec = meas["rows_kernel"] + meas["cols_kernel"]
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({how})",
                "traffic": ncu_traffic(dom, nx, ny, world), "traffic_source": "profiles/ncu_traffic.json (ncu --set full, per launch)",
                "algorithmic_bytes_per_launch": dbytes, "avg_kernel_ms": dsec * 1e3, "per_kernel": per_kernel,
                "whole_transform": {"algorithmic_bytes": ab["total"] / world, "compute_ms": compute_sec * 1e3,
                                    "achieved": ab["total"] / world / compute_sec / 1e9 if compute_sec > 0 else 0.0,
                                    "frac": ab["total"] / world / compute_sec / 1e9 / peak if compute_sec > 0 else 0.0}}
        if world > 1:
            bx = exchange_bytes(nx, ny, world, 0)
            nv = {"bytes_per_gpu_per_direction": bx, "peak": NVLINK_PEAK_GBS, "unit": "GB/s",
                  "peak_source": "B200_PROFILING.md measured peer copy per direction (900 nominal)", "transport": transport}
            for name, key, which in (("first_comm", "first_comm_span", 1), ("second_comm", "second_comm_span", 2)):
                span = meas[key]
                ent = {"in_step_span_ms": span * 1e3, "in_step_gbs": bx / span / 1e9 if span > 0 else None,
                       "exposed_ms": meas[name] * 1e3}
                if which in xalone and xalone[which] > 0:
                    ent.update(alone_ms=xalone[which], achieved=bx / (xalone[which] * 1e-3) / 1e9,
                               frac=bx / (xalone[which] * 1e-3) / 1e9 / NVLINK_PEAK_GBS)
                elif transport == "fused-peer-store":
                    # the exchange IS the producing kernel's store stream: its bytes leave over NVLink while the kernel runs
                    ksec = meas["rows_kernel"] if which == 1 else meas["cols_kernel"]
                    ent.update(fused_into="rows_r2c" if which == 1 else "cols_c2c", kernel_ms=ksec * 1e3,
                               achieved=bx / ksec / 1e9 if ksec > 0 else None,
                               frac=bx / ksec / 1e9 / NVLINK_PEAK_GBS if ksec > 0 else None)
                nv[name] = ent
            roof["nvlink"] = nv
            # < 1 when communication hides behind the kernels: total / (kernels + both exchanges run alone)
            if xalone:
                serial = compute_sec * 1e3 + xalone[1] + xalone[2] + meas["second_trans"] * 1e3 * (transport.startswith("nccl"))
                roof["overlap"] = {"total_ms": meas["total"] * 1e3, "sum_of_phases_alone_ms": serial, "ratio": meas["total"] * 1e3 / serial}
            elif transport == "fused-peer-store" and anchor and "rows_ms" in anchor:
                # phases alone = the kernels without remote stores (1-GPU anchor / N) + both exchanges at the NVLink peak
                t_x = bx / NVLINK_PEAK_GBS / 1e6
                serial = (anchor["rows_ms"] + anchor["cols_ms"]) / world + 2 * t_x
                roof["overlap"] = {"total_ms": meas["total"] * 1e3, "sum_of_phases_alone_ms": serial, "ratio": meas["total"] * 1e3 / serial,
                                   "how": "kernels alone = 1-GPU anchor kernel times / N; exchanges alone = bytes / 770 GB/s"}
        result = {
            "metric": METRIC, "value": gf / (ms_step * 1e-3) / 1e9, "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(nx, ny, world, comm),
            "transport": transport, "host_numa_bound": bool(numa_bound),
            "gpu_launches": launches,
            "phases_ms": {k: meas[k] * 1e3 for k in meas if k != "timer_samples"},
            "roofline": roof,
            "parity": parity if not args.no_parity else {"skipped": "--no-parity (diagnostic run, not a bench value)"},
            "clocks": clocks,
        }
        if e2e:
            result["e2e"] = {"value": gf / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_ms, "steps": e2e[1],
                             "h2d_bytes_per_step": slab_bytes * world, "d2h_bytes_per_step": slab_bytes * world,
                             "api": "hpxfft_b200_transform_async, 2 plans double-buffered (pinned host vector_2d -> device -> host, every step)"}
        else:
            result["e2e"] = None
            result["e2e_note"] = "skipped: slab too large for host staging (input generated on the device)" if big else "skipped (--no-e2e)"
        if anchor:
            result["anchor_n1"] = anchor
        if world == 1 and not args.no_cpu_baseline:
            if orig_affinity is not None:   # the CPU leg uses ALL host cores, not only those of the GPU's NUMA node
                try:
                    os.sched_setaffinity(0, orig_affinity)
                except Exception:
                    pass
            result["cpu_baseline"] = cpu_baseline(nx, ny, budget_s=args.cpu_budget)
    lib.hpxfft_b200_destroy(plan)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and not (parity["rel_l2"] <= PARITY_TOL):
        print(json.dumps(result), flush=True)
        print(f"bench.py: PARITY FAILURE rel_l2={parity['rel_l2']:.3e} > {PARITY_TOL}", file=sys.stderr)
        sys.exit(3)
    return result


# ---------------------------------------------------------------------------------------------------
# CPU legs: the oracle port is EXECUTED here only as the timed baseline / reference arm
# ---------------------------------------------------------------------------------------------------
def oracle_c_lib():
    path = os.path.join(ROOT, "oracle", "libhpxfft_oracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    lib = C.CDLL(path)
    lib.hpxfft_oracle_shared_loop.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
    lib.hpxfft_oracle_shared_loop.restype = C.c_int
    return lib


def host_cores() -> int:
    """All the host threads this process may use -- NOT omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1, and the port takes its thread count as an explicit argument."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def mem_available_bytes() -> int:
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) * 1024
    except Exception:
        pass
    return 8 << 30


def cpu_sample_shape(nx: int, ny: int, budget_s: float, runs: int, gflops_guess: float) -> tuple[int, int]:
    """Bounded sample of the workload, the same rule at every N: keep ny (the row length), keep the whole nx if
    `runs` transforms fit the time budget and 2.2x the array fits in host memory, otherwise halve nx."""
    sx = nx
    while sx > 256 and (runs * flops(sx, ny) / (gflops_guess * 1e9) > budget_s or 3.2 * sx * (ny + 2) * 8 > 0.7 * mem_available_bytes()):
        sx //= 2
    return sx, ny


def cpu_calibrate(lib, cores: int) -> float:
    """GFLOP/s of the port on a small case, to size the sample."""
    import numpy as np
    sx, sy = 1024, 4096
    a = np.zeros((sx, sy + 2))
    a[:, :sy] = 1.0
    tm = np.zeros(5)
    best = None
    for _ in range(3):
        v = a.copy()
        assert lib.hpxfft_oracle_shared_loop(v.ctypes.data, sx, sy + 2, cores, tm.ctypes.data) == 0
        best = tm[0] if best is None else min(best, tm[0])
    return flops(sx, sy) / best / 1e9


def cpu_baseline(nx: int, ny: int, budget_s: float = 20.0) -> dict:
    import numpy as np
    lib = oracle_c_lib()
    cores = host_cores()
    sx, sy = cpu_sample_shape(nx, ny, budget_s, 3, 0.6 * cpu_calibrate(lib, cores))
    a = np.zeros((sx, sy + 2))
    a[:, :sy] = np.random.default_rng(0).uniform(-1, 1, (sx, sy))
    best, spent, runs, tm = None, 0.0, 0, np.zeros(5)
    while runs < 2 or (spent < budget_s and runs < 5):
        v = a.copy()
        assert lib.hpxfft_oracle_shared_loop(v.ctypes.data, sx, sy + 2, cores, tm.ctypes.data) == 0
        spent += tm[0]
        runs += 1
        best = tm[0] if best is None else min(best, tm[0])
    return {"value": flops(sx, sy) / best / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "port",
            "sample": f"{sx}x{sy} FP64 r2c" + ("" if sx == nx else f" (nx reduced from {nx})") +
                      f", best of {runs} runs of the 4-phase C restatement of shared::loop (oracle/hpxfft_oracle.c, "
                      f"OpenMP, {cores} threads); FFTW/HPX unavailable in this image", "ms": best * 1e3}


def run_reference(args) -> dict | None:
    """Reference arm: the reference's own CPU algorithm (4-phase loop, core/src/shared/loop.cpp:56-113) on ALL of
    the box's host cores.  HPX + FFTW cannot be built in this image, so the plain-C port in oracle/ is timed; for
    the distributed configurations it is the single-host shared::loop restatement at the same GLOBAL size
     -- no multi-process HPX exists here.  Under torchrun rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return None
    import numpy as np
    nx, ny = workload(args, world)
    lib = oracle_c_lib()
    cores = host_cores()
    runs = args.warmup + args.steps
    sx, sy = cpu_sample_shape(nx, ny, args.ref_budget, runs, 0.6 * cpu_calibrate(lib, cores))
    a = np.zeros((sx, sy + 2))
    a[:, :sy] = np.random.default_rng(0).uniform(-1, 1, (sx, sy))
    tm = np.zeros(5)
    times = []
    v = np.empty_like(a)
    for i in range(runs):
        np.copyto(v, a)
        assert lib.hpxfft_oracle_shared_loop(v.ctypes.data, sx, sy + 2, cores, tm.ctypes.data) == 0
        if i >= args.warmup:
            times.append(tm[0])
    ms = 1e3 * sum(times) / len(times)
    val = flops(sx, sy) / (ms * 1e-3) / 1e9
    comm = None if world == 1 else args.run
    sample = (f"{sx}x{sy} FP64 r2c per step" + ("" if sx == nx else f" (nx reduced from {nx} to bound the run; GFLOP/s is size-normalised)") +
              f"; single-host 4-phase C restatement of shared::loop at the global size (oracle/hpxfft_oracle.c, OpenMP "
              f"{cores} threads on {os.cpu_count()} CPUs); FFTW/HPX unavailable in this image, no multi-process HPX")
    return {"impl": "reference", "metric": METRIC, "value": val, "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(nx, ny, world, comm),
            "sample_is_full_config": sx == nx,
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="", choices=["", "c1", "c5"])
    ap.add_argument("--nx", type=int, default=0)
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--comm", "--run", dest="run", default="all_to_all", choices=["all_to_all", "scatter", "p2p"])
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--ref-budget", type=float, default=100.0, help="seconds of CPU work the reference arm may spend")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-anchor", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="diagnostic builds only (their results are garbage by construction)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    res = run_reference(args) if args.impl == "reference" else run_ours(args)
    if res is not None:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
