#!/usr/bin/env python3
"""Compares the SASS of EVERY kernel of two builds of the library (labels and internal subroutine numbers normalised).

    python tools/sass_equal.py <lib A> [<lib B>, default shocovox_b200/libshocovox_b200.so]

For refactors that must not change device code (host-only changes, new options that are off by default): if every kernel
is identical to a build that passed the GPU parity suite, the refactor needs no GPU to be trusted. Exit code 1 on any
difference (a kernel present in only one of the two counts as one)."""
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def kernels_of(lib: Path) -> dict:
    """kernel name -> normalised SASS lines"""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", str(lib.resolve())], cwd=d, check=True, capture_output=True)
        out = ""
        for cubin in sorted(Path(d).glob("*.cubin")):
            out += subprocess.run(["nvdisasm", "-c", str(cubin)], capture_output=True, text=True, check=True).stdout
    found, name = {}, None
    for ln in out.splitlines():
        if ln.startswith(".text."):
            name = ln[len(".text."):].rstrip(":")
            found[name] = []
            continue
        if ln.startswith(".section") or ln.startswith("\t.section"):
            name = None
        if name is None:
            continue
        ln = re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip()
        if not ln or ln.startswith("//"):
            continue
        ln = re.sub(r"__internal_\d+_", "__internal_N_", re.sub(r"\.L_x_\d+", "L", ln))
        found[name].append(ln)
    return found


def short(mangled: str) -> str:
    m = re.search(r"svx\d+(\w+?)E", mangled)
    return m.group(1) if m else mangled


def main():
    a = Path(sys.argv[1])
    b = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "shocovox_b200" / "libshocovox_b200.so"
    ka, kb = kernels_of(a), kernels_of(b)
    same = True
    for k in sorted(set(ka) | set(kb)):
        x, y = ka.get(k), kb.get(k)
        verdict = "identical" if x == y else ("missing in one" if x is None or y is None else f"DIFFERENT ({len(x)} vs {len(y)} lines)")
        same &= x == y
        print(f"{short(k):56s} {verdict}")
    print(f"{len(set(ka) | set(kb))} kernels, {'all identical' if same else 'NOT all identical'}")
    sys.exit(0 if same else 1)


if __name__ == "__main__":
    main()
