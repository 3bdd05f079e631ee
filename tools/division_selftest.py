#!/usr/bin/env python3
"""Runs svx_selftest_division on cuda:0: the shared-reciprocal form of the per-ray divisions (traverse.cuh: Reciprocal /
div_by, kernel option SVX_SHARED_RCP, off by default) against the IEEE `a / b`.

    python tools/division_selftest.py [pairs, default 2^32] [seeds, default 4]

Prints one JSON line per seed; every `mismatches` must be 0 before SVX_SHARED_RCP=1 may become the default."""
import ctypes as C
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import shocovox_b200 as S  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 32
    seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    L = S.lib()
    for seed in range(seeds):
        bad, done = C.c_uint64(0), C.c_uint64(0)
        rc = L.svx_selftest_division(0, n, 0x5EED0000 + seed, C.byref(bad), C.byref(done))
        print(json.dumps({"seed": seed, "status": rc, "pairs": n, "tested": done.value, "mismatches": bad.value}))
        if rc != 0:
            print(L.svx_last_error_message().decode(), file=sys.stderr)
            sys.exit(1)


if __name__ == "__main__":
    main()
