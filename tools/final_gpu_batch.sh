#!/bin/bash
# Round-end evidence on one B200: parity tests, benches, launch list, full ncu captures, sanitizer. Outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python tools/reload_probe.py minecraft 2>&1 | tail -5
python tools/reload_probe.py terrain 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 3000 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -c 600 gpurun_out/bench_reference.log
timeout 600 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1; cat gpurun_out/perf_probe.log
# launch list of the default bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_final2.csv python bench.py --steps 10 --warmup 3 > gpurun_out/ncu_launch_bench.log 2>&1
# full captures: the render kernel on the default workload and on C3, the occupancy-bit kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r01_render_dotcube_final2 python tools/perf_probe.py dot_cube_1080p > gpurun_out/ncu_full_final2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:occupancy_bits_kernel --launch-count 1 -f -o gpurun_out/r01_occupancy_bits_minecraft python tools/perf_probe.py minecraft_1024_32_4k > gpurun_out/ncu_bits.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck_smoke.log
