#!/usr/bin/env python3
"""Where a viewport kernel's instructions go: executed warp instructions of an .ncu-rep (captured with
`--import-source on`, read here without a GPU) summed per function of csrc/traverse.cuh and per phase of traverse().

usage: tools/ncu_breakdown.py gpurun_out/x.ncu-rep

The line -> function map is taken from the CURRENT csrc/traverse.cuh, so use it on captures of the current source.
"""
import csv
import io
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
# phases of traverse<LOD, BS>() by the comments that open them
PHASES = [("crawl: root popped right away, fast-forward", "if (root_can_crawl) {"),
          ("node loop head: record load, LOD test", "count = min(count + 1u, 4u);  // node_stack.push(root)"),
          ("leaf probe call sites", "#if SVX_SINGLE_PROBE_SITE"),
          ("position in the node's 4x4x4 bitmap, occupancy test, POP", "// position inside the node in 4x4x4 bitmap cells"),
          ("PUSH", "const float hs = bsize * 0.5f;\n            float tbx"),
          ("ADVANCE (sibling walk)", "// ADVANCE (:497-544)"),
          ("restart nudge", "// restart from the root after a 0.1 nudge")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    per, cur_file, hdr = {}, None, None
    for r in csv.reader(io.StringIO(out)):
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) > 5 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0] != "":
            try:
                per[(cur_file, int(r[0]))] = (int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("Thread Instructions Executed")]))
            except ValueError:
                pass
    total = sum(v[0] for v in per.values()) or 1
    threads = sum(v[1] for v in per.values())
    print(f"{rep}: {total} executed warp instructions, {threads / total:.2f} threads per instruction")
    text = (ROOT / "shocovox_b200" / "csrc" / "traverse.cuh").read_text()
    src = text.split("\n")
    starts = [(i, m.group(1)) for i, l in enumerate(src, 1) if (m := re.match(r"^__device__ __forceinline__ .*?(\w+)\(", l))]

    def func_of(line):
        name = "(top)"
        for s, n in starts:
            if s <= line:
                name = n
        return name

    agg = {}
    for (f, l), (n, tn) in per.items():
        a = agg.setdefault((f, func_of(l) if f == "traverse.cuh" else ""), [0, 0])
        a[0] += n
        a[1] += tn
    print("\n share  threads/inst  where")
    for k, (n, tn) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if n * 1000 >= total:
            print(f"{100 * n / total:5.1f}%  {tn / max(n, 1):5.1f}  {k[0]}{':' + k[1] if k[1] else ''}")
    # phases of traverse(): from each marker to the next
    t0 = next(s for s, n in starts if n == "traverse")
    t1 = min(s for s, n in starts if s > t0)
    marks = []
    for name, marker in PHASES:
        pos = text.find(marker, sum(len(l) + 1 for l in src[:t0 - 1]))
        marks.append((text.count("\n", 0, pos) + 1, name))
    marks.append((t1, None))
    print("\n share  threads/inst  phase of traverse() (own lines only; inlined helpers are listed above)")
    for (a, name), (b, _) in zip(marks, marks[1:]):
        n = sum(v[0] for (f, l), v in per.items() if f == "traverse.cuh" and a <= l < b)
        tn = sum(v[1] for (f, l), v in per.items() if f == "traverse.cuh" and a <= l < b)
        print(f"{100 * n / total:5.1f}%  {tn / max(n, 1):5.1f}  L{a}-{b - 1}  {name}")


if __name__ == "__main__":
    main()
