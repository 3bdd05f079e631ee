#!/bin/bash
# Round-end evidence on one B200: parity tests, smoke, benches (plain, reference arm, LOD), launch list, full ncu captures,
# sanitizer. Outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 2500 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -c 600 gpurun_out/bench_reference.log
timeout 600 python bench.py --mips frustum --steps 50 > gpurun_out/bench_dot_cube_mips_frustum.log 2>&1; tail -c 1800 gpurun_out/bench_dot_cube_mips_frustum.log
timeout 900 python bench.py --workload minecraft_4k --steps 20 --cpu-budget 6 > gpurun_out/bench_minecraft_4k.log 2>&1; tail -c 1500 gpurun_out/bench_minecraft_4k.log
timeout 900 python bench.py --workload minecraft_4k --mips 2000 --steps 20 --cpu-budget 6 > gpurun_out/bench_minecraft_4k_mips_2000.log 2>&1; tail -c 1500 gpurun_out/bench_minecraft_4k_mips_2000.log
# launch list of the default bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_final3.csv python bench.py --steps 10 --warmup 3 > gpurun_out/ncu_launch_bench.log 2>&1
# full capture: the render kernel on the default workload
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r01_render_dotcube_final3 python tools/perf_probe.py dot_cube_1080p > gpurun_out/ncu_full_final3.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck_smoke.log
