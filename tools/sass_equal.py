#!/usr/bin/env python3
"""Compares the SASS of every kernel of two builds of the library (labels and internal subroutine numbers normalised).

    python tools/sass_equal.py <lib A> [<lib B>, default shocovox_b200/libshocovox_b200.so]

For refactors that must not change device code (host-only #ifdefs, new experimental options that are off by default):
if every kernel is identical to a build that passed the GPU parity suite, the refactor needs no GPU to be trusted."""
import re
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import sass_loop as S  # noqa: E402

KERNELS = ["render_kernel_brick32E", "render_kernel_brick8E", "render_kernelE", "render_lod_kernel_brick8E", "render_lod_kernel_brick32E",
           "render_lod_kernelE", "render_shaded_kernelE", "render_lod_shaded_kernelE", "render_kernel_persistentE",
           "render_lod_kernel_persistentE", "rays_kernelE", "rays_lod_kernelE", "occupancy_bits_kernelE", "occupancy_bits_small_kernelE",
           "lut_selftest_kernelE"]


def normalised(lib: Path, kernel: str):
    S.LIB = lib
    return [re.sub(r"__internal_\d+_", "__internal_N_", re.sub(r"\.L_x_\d+", "L", ln)) for ln in S.kernel_sass(kernel)]


def main():
    a = Path(sys.argv[1])
    b = Path(sys.argv[2]) if len(sys.argv) > 2 else S.ROOT / "shocovox_b200" / "libshocovox_b200.so"
    same = True
    for k in KERNELS:
        x, y = normalised(a, k), normalised(b, k)
        verdict = "identical" if x == y else ("missing in one" if not x or not y else f"DIFFERENT ({len(x)} vs {len(y)} lines)")
        same &= x == y
        print(f"{k:40s} {verdict}")
    sys.exit(0 if same else 1)


if __name__ == "__main__":
    main()
