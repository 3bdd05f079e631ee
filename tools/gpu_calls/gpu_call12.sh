#!/bin/bash
# round 2, call 12: the palette test, then the default bench line of the final code (e2e headline = 8 B/pixel)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "albedo_plane or pipelined" > gpurun_out/c12_tests.log 2>&1; tail -3 gpurun_out/c12_tests.log
timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/c12_bench.err; tail -c 600 gpurun_out/c12_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, json.dumps(d["e2e"])[:1800])
PY
