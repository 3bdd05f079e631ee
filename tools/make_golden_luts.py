#!/usr/bin/env python3
"""Extract the literal look-up tables of the reference into tests/golden/luts.json.

Runs only in the build container (reads /root/reference/src/spatial/lut.rs:154-896); the JSON it writes is
the committed fixture the oracle's *regenerated* tables are checked against (tests/test_oracle_spatial.py).
The product and the oracle never read this file to obtain their tables - they rebuild them from the
generator logic (lut.rs:12-152).
"""
import json
import re
import sys
from pathlib import Path

SRC = Path("/root/reference/src/spatial/lut.rs")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "luts.json"


def block(text: str, name: str) -> str:
    start = text.index(f"const {name}")
    eq = text.index("=", start)
    end = text.index("];\n", eq)
    return text[eq + 1 : end + 1]


def ints(s: str):
    return [int(tok, 0) for tok in re.findall(r"0x[0-9A-Fa-f]+|\d+", s)]


def main() -> int:
    text = SRC.read_text()
    offsets = [float(v) for v in re.findall(r"[xyz]:\s*([0-9.]+)", block(text, "OCTANT_OFFSET_REGION_LUT"))]
    out = {
        "source": "src/spatial/lut.rs:154-896",
        "OOB_OCTANT": ints(re.search(r"OOB_OCTANT: u8 = (\d+)", text).group(1))[0],
        "OCTANT_OFFSET_REGION_LUT": [offsets[i : i + 3] for i in range(0, 24, 3)],
        "BITMAP_MASK_FOR_OCTANT_LUT": ints(block(text, "BITMAP_MASK_FOR_OCTANT_LUT")),
        "BITMAP_INDEX_LUT": ints(block(text, "BITMAP_INDEX_LUT")),  # [x][y][z] flattened
        "OCTANT_STEP_RESULT_LUT": ints(block(text, "OCTANT_STEP_RESULT_LUT")),  # [x][y][z] flattened
        "RAY_TO_NODE_OCCUPANCY_BITMASK_LUT": ints(block(text, "RAY_TO_NODE_OCCUPANCY_BITMASK_LUT")),  # [pos][dir]
    }
    assert len(out["BITMAP_MASK_FOR_OCTANT_LUT"]) == 8
    assert len(out["BITMAP_INDEX_LUT"]) == 64
    assert len(out["OCTANT_STEP_RESULT_LUT"]) == 27
    assert len(out["RAY_TO_NODE_OCCUPANCY_BITMASK_LUT"]) == 512
    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_text(json.dumps(out, indent=1) + "\n")
    print(f"wrote {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
