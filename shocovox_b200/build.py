"""Builds libshocovox_b200.so (CUDA kernels + C ABI) for sm_100a, in-tree.

nvcc cross-compiles without a GPU. Flags that matter for parity with the reference's f32 arithmetic:
  -fmad=false                      Rust never contracts a*b+c into an FMA
  -prec-div=true -prec-sqrt=true   IEEE-rounded division and sqrt (nvcc defaults, stated explicitly)
  -ftz=false                       denormals kept (default)
  -Xcompiler -ffp-contract=off     the host code that derives the per-frame camera constants follows the same rule
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libshocovox_b200.so"
SOURCES = ["host_octree.cpp", "host_octree_mip.cpp", "host_octree_io.cpp", "vox_import.cpp", "gpu_tree.cpp", "kernels.cu", "capi.cu", "multi_gpu.cu"]
HEADERS = ["host_octree.hpp", "gpu_tree.hpp", "kernels.cuh", "traverse.cuh", "traverse_refill.cuh", "capi_internal.hpp", "vox_import.hpp", "../../include/shocovox_b200.h"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES + HEADERS] + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: Path = LIB, defines=()) -> Path:
    """Builds the library. `out` / `defines` produce tuning variants (e.g. -DSVX_MIN_BLOCKS=4) next to the default."""
    if not force and out == LIB and not is_stale():
        return LIB
    cmd = [
        nvcc_path(), "-shared", "-o", str(out), *[f"-D{d}" for d in defines],
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-std=c++17", "-lineinfo",
        "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
        "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden,-Wall",
        "-Xptxas", "-v",
        "-cudart", "static",
        "-x", "cu",
    ] + [str(CSRC / s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stderr[-4000:])
    (PKG / ("build_ptxas.log" if out == LIB else out.name + ".ptxas.log")).write_text(res.stderr)
    check_no_packed_fma(out)
    return out


def check_no_packed_fma(lib: Path) -> None:
    """The kernels use Blackwell's packed f32 add / mul (FADD2 / FMUL2, traverse.cuh). ptxas 12.9 contracts a packed multiply
    that feeds a packed add into FFMA2 even for `.rn` operands and with --fmad=false - one rounding instead of the reference's
    two. The sources avoid that pattern; a library that contains FFMA2 anyway is refused rather than shipped."""
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([tool, "-sass", str(lib)], capture_output=True, text=True)
    if sass.returncode != 0:
        raise RuntimeError("cuobjdump failed on " + str(lib) + ":\n" + sass.stderr[-2000:])
    fused = [ln.strip() for ln in sass.stdout.splitlines() if "FFMA2" in ln]
    if fused:
        lib.unlink()
        raise RuntimeError(f"{len(fused)} packed fused multiply-adds in the SASS (results would differ from the reference): {fused[0]}")


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
