#!/bin/bash
# one B200, final code: whole gpu suite, smoke, default bench, launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_gputest_1gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
( time timeout 600 python bench.py > gpurun_out/r02_bench_default_1gpu.json 2> gpurun_out/r02_bench_default_1gpu.err ) 2>&1 | tail -3; head -c 400 gpurun_out/r02_bench_default_1gpu.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_sponza_4k.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/c9_ncu_launch.log 2>&1; tail -c 300 gpurun_out/c9_ncu_launch.log
