#!/usr/bin/env python
"""GPU counterparts of HPX-FFT's benchmark sweeps (no SLURM: the sweeps run on the GPUs of this box).

Same shapes as the reference's drivers, same CSV files as its example programs write
(`runtimes/runtimes_hpx_shared_loop.txt`, `runtimes/runtimes_hpx_distributed_loop.txt`):

  size      nx = ny = 2^5 .. 2^12 on one GPU          benchmark/shared_benchmark.sh:132-135,
                                                      sbatch_scripts/run_hpxfft_size_shared.sh:27-41
  strong    nx = ny = 16384, GPUs 1, 2, 4, 8          benchmark/distributed_benchmark.sh:29-31,71-73
  weak      nx = ny = base * GPUs (base 8192)         benchmark/distributed_benchmark.sh:104-106
  message   2 GPUs, nx = ny = 128 * 2^0 .. 2^5        benchmark/message_benchmark.sh:59-61,68

usage:  python benchmark/sweeps.py size|strong|weak|message [--plan estimate] [--run all_to_all] [--loop 2]
The executables are tests/cpp/hpxfft_shared_loop and hpxfft_distributed_loop (built by __graft_entry__.build()).
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp")


def shared(nx: int, ny: int, plan: str, header: bool, cwd: str) -> None:
    cmd = [os.path.join(BIN, "hpxfft_shared_loop"), f"--nx={nx}", f"--ny={ny}", f"--plan={plan}", f"--header={int(header)}"]
    subprocess.run(cmd, check=True, cwd=cwd, stdout=subprocess.DEVNULL)


def distributed(n_gpus: int, nx: int, ny: int, plan: str, run: str, header: bool, cwd: str) -> None:
    procs = []
    for rank in range(n_gpus):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(n_gpus), LOCAL_RANK=str(rank), MASTER_PORT="29555",
                   HPXFFT_B200_RENDEZVOUS=cwd)
        cmd = [os.path.join(BIN, "hpxfft_distributed_loop"), f"--nx={nx}", f"--ny={ny}", f"--plan={plan}", f"--run={run}",
               f"--header={int(header)}"]
        procs.append(subprocess.Popen(cmd, env=env, cwd=cwd, stdout=subprocess.DEVNULL))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("hpxfft_distributed_loop failed")
    for f in os.listdir(cwd):  # rendezvous files of this launch
        if f.startswith("hpxfft_b200_29555_"):
            os.remove(os.path.join(cwd, f))


def gpu_count() -> int:
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except OSError:
        return 0


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("sweep", choices=["size", "strong", "weak", "message"])
    ap.add_argument("--plan", default="estimate")
    ap.add_argument("--run", default="all_to_all", choices=["all_to_all", "scatter", "p2p"])
    ap.add_argument("--loop", type=int, default=2)
    ap.add_argument("--out", default=".")
    a = ap.parse_args()
    out = os.path.abspath(os.path.join(a.out, f"{a.sweep}_scaling"))
    os.makedirs(out, exist_ok=True)
    subprocess.run(["make", "-s", "-C", BIN], check=True)
    ngpu = gpu_count()
    if ngpu == 0:
        sys.exit("no GPU visible: hpxfft_b200 has no CPU fallback")
    first = True
    if a.sweep == "size":
        for p in range(5, 13):
            for _ in range(a.loop):
                shared(2 ** p, 2 ** p, a.plan, first, out)
                first = False
    elif a.sweep == "strong":
        n = 1
        while n <= ngpu:
            for _ in range(a.loop):
                distributed(n, 16384, 16384, a.plan, a.run, first, out)
                first = False
            n *= 2
    elif a.sweep == "weak":
        n = 1
        while n <= ngpu:
            for _ in range(a.loop):
                distributed(n, 8192 * n, 8192 * n, a.plan, a.run, first, out)
                first = False
            n *= 2
    else:
        if ngpu < 2:
            sys.exit("the message-size sweep needs 2 GPUs")
        for k in range(0, 6):
            for _ in range(a.loop):
                distributed(2, 128 * 2 ** k, 128 * 2 ** k, a.plan, a.run, first, out)
                first = False
    print(f"runtimes written under {out}/runtimes/")


if __name__ == "__main__":
    main()
