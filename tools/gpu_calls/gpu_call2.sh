#!/bin/bash
# Round-2 GPU call 2 (one B200): whole gpu suite, default bench with extras + reference arm, ncu launch list + full captures.
mkdir -p gpurun_out
echo "== gather tests with CUDA tracing"
SVX_DEBUG_CUDA=1 timeout 300 python -m pytest tests/test_gpu_gather.py -x -q -s 2>&1 | grep -v "last error cudaSuccess" | tail -40 | tee gpurun_out/c2_gather_debug.log
echo "== whole gpu suite"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/c2_pytest_gpu.log
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/c2_smoke.log
echo "== default bench"
( time timeout 800 python bench.py > gpurun_out/c2_bench_default.json 2> gpurun_out/c2_bench_default.err ) 2>&1 | tail -4; tail -c 6000 gpurun_out/c2_bench_default.json; tail -5 gpurun_out/c2_bench_default.err
echo "== reference arm"
( time timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c2_bench_reference.json 2>&1 ) 2>&1 | tail -4; tail -c 600 gpurun_out/c2_bench_reference.json
echo "== ncu launch list of the bench command"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_sponza_4k.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/c2_ncu_launch.log 2>&1; tail -3 gpurun_out/c2_ncu_launch.log
echo "== ncu full: sponza, minecraft, terrain pose"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r02_sponza_2048_32_4k python tools/perf_probe.py sponza_2048_32_4k > gpurun_out/c2_ncu_sponza.log 2>&1; tail -2 gpurun_out/c2_ncu_sponza.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r02_minecraft_1024_32_4k python tools/perf_probe.py minecraft_1024_32_4k > gpurun_out/c2_ncu_minecraft.log 2>&1; tail -2 gpurun_out/c2_ncu_minecraft.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r02_terrain_1024_8_1080p python tools/perf_probe.py terrain_1024_8_1080p > gpurun_out/c2_ncu_terrain.log 2>&1; tail -2 gpurun_out/c2_ncu_terrain.log
