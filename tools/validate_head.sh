#!/bin/bash
# Quick GPU validation of the current HEAD on one B200 (~1.5 GPU-minutes): parity tests, smoke, default bench.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 200 python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 1800 gpurun_out/bench_default.log
