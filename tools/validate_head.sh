#!/bin/bash
# Quick GPU validation of the current HEAD on one B200: parity tests, smoke, default bench, reference arm, kernel-time probe.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 1800 gpurun_out/bench_default.log
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -c 500 gpurun_out/bench_reference.log
SCENES="dot_cube_1080p cpu_render_4k colonnade_4k terrain_512_8_4k minecraft_256_32_4k sponza_2048_32_4k"
timeout 300 python tools/perf_probe.py $SCENES > gpurun_out/perf_probe_plain.log 2>&1; cat gpurun_out/perf_probe_plain.log
