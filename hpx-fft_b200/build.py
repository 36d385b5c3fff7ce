"""Builds libhpxfft_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).
The translation units are compiled in parallel and linked into one shared library."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhpxfft_b200.so")
# (source, extra defines, object suffix)
UNITS = [("common.cu", (), ""), ("plan.cu", (), ""), ("launch_rows.cu", (), ""), ("launch_rows_long.cu", (), ""), ("launch_rows_ditc.cu", (), ""), ("launch_cols.cu", (), ""), ("launch_misc.cu", (), ""),
         ("launch_generic.cu", (), ""), ("launch_bluestein.cu", (), "")] + \
        [("launch_fused.cu", (f"HPXFFT_B200_FUSED_GROUP={g}",), f"_g{g}") for g in range(4)]
HEADERS = ["fft_device.cuh", "layout.cuh", "kernels_rows.cuh", "kernels_rows_long.cuh", "kernels_rows_long2.cuh", "kernels_rows_dit2.cuh", "kernels_rows_ditc.cuh", "kernels_rows_v2.cuh", "kernels_cols.cuh", "kernels_misc.cuh", "kernels_generic.cuh", "kernels_bluestein.cuh",
           "internal.h", "launch_util.h", os.path.join("..", "..", "include", "hpxfft_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _units():
    return [u for u in UNITS if os.path.exists(os.path.join(CSRC, u[0]))]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in [u[0] for u in UNITS] + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, defines: tuple = (), out: str = LIB) -> str:
    """`defines` / `out` produce variant libraries for A/B runs (selected with HPXFFT_B200_LIB)."""
    if not force and out == LIB and not needs_build():
        return LIB
    tag = "" if out == LIB else "_" + os.path.splitext(os.path.basename(out))[0]
    objdir = os.path.join(HERE, "build" + tag)
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS if os.path.exists(os.path.join(CSRC, h)))

    def compile_one(unit):
        src, defs, suffix = unit
        obj = os.path.join(objdir, os.path.splitext(src)[0] + suffix + ".o")
        srcp = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(srcp)):
            return obj
        cmd = [nvcc] + ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-c", "-o", obj] + \
              [f"-D{d}" for d in tuple(defines) + tuple(defs)] + [srcp]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True, cwd=CSRC)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _units()))
    subprocess.run([nvcc] + ARCH + ["-shared", "-o", out] + objs + ["-ldl"], check=True, cwd=CSRC)
    bad = suspicious_sass(out)
    if bad:
        os.replace(out, out + ".rejected")
        raise RuntimeError("ptxas emitted LDGSTS with an unset uniform descriptor register (see suspicious_sass) in: " + ", ".join(bad))
    return out


def suspicious_sass(lib: str = LIB) -> list:
    """Kernels whose cp.async instructions use the `[R+URx], desc[URy]` form.  ptxas 12.9 produced it for one instantiation of
    rows_ditc_kernel (hinted cp.async under uniform-register pressure) with URx / URy never written: the launch died on a B200
    with "illegal instruction".  No kernel of this library needs that form, so its presence rejects the build."""
    cuobjdump = os.path.join(os.path.dirname(_nvcc()), "cuobjdump")
    if not os.path.exists(cuobjdump):
        return []
    sass = subprocess.run([cuobjdump, "-sass", lib], check=True, capture_output=True, text=True).stdout
    bad, fn = set(), "?"
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        elif "LDGSTS" in line and "+UR" in line.split("desc[")[0]:
            bad.add(fn)
    return sorted(bad)


if __name__ == "__main__":
    defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    outs = [a[len("--out="):] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else LIB))
