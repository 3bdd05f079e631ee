#!/usr/bin/env python3
"""Prints the SASS of the brick DDA loop(s) of one kernel of the built library, with per-loop instruction counts.

    python tools/sass_loop.py [kernel substring, default render_kernelE] [--full]

A loop = the code between a label and the last backward branch to it that contains the bit-brick load's reload test
(SHF.R.U32.HI ..., 0x5). Used to check what a source change did to the hottest loop before spending GPU time."""
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "shocovox_b200" / "libshocovox_b200.so"


def kernel_sass(name: str) -> list[str]:
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", str(LIB)], cwd=d, check=True, capture_output=True)
        cubin = next(p for p in Path(d).glob("*.cubin") if p.stat().st_size > 100000)
        out = subprocess.run(["nvdisasm", "-c", str(cubin)], capture_output=True, text=True, check=True).stdout
    lines, keep = [], False
    for ln in out.splitlines():
        if ln.startswith(".text."):
            keep = name in ln
            continue
        if keep:
            ln = re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip()
            if ln and not ln.startswith("//") and not ln.startswith(".") or re.match(r"\.L_x_\d+:", ln):
                lines.append(ln)
    return lines


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    name = args[0] if args else "render_kernelE"
    lines = kernel_sass(name)
    n_inst = sum(1 for l in lines if not l.endswith(":"))
    print(f"{name}: {n_inst} instructions")
    labels = {l[:-1]: i for i, l in enumerate(lines) if l.endswith(":")}
    for lab, start in labels.items():
        back = [i for i, l in enumerate(lines) if i > start and re.search(r"BRA.*`\(" + re.escape(lab) + r"\)", l)]
        if not back:
            continue
        body = lines[start:back[-1] + 1]
        if not any("SHF.R.U32.HI" in l and "0x5" in l for l in body) or len(body) > 90:
            continue
        count = sum(1 for l in body if not l.endswith(":"))
        print(f"--- loop {lab}: {count} instructions")
        if "--full" in sys.argv:
            print("\n".join("    " + l for l in body))


if __name__ == "__main__":
    main()
