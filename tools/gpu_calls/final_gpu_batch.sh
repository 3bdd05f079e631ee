#!/bin/bash
# Round-end evidence on one B200: parity tests, smoke, benches (plain, reference arm, LOD), kernel-time probes with and without
# MIP maps, launch list, full ncu captures, sanitizer. Outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 1500 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -c 400 gpurun_out/bench_reference.log
timeout 600 python bench.py --mips frustum --steps 50 > gpurun_out/bench_dot_cube_mips_frustum.log 2>&1; tail -c 600 gpurun_out/bench_dot_cube_mips_frustum.log
timeout 900 python bench.py --workload minecraft_4k --steps 20 --cpu-budget 6 > gpurun_out/bench_minecraft_4k.log 2>&1; tail -c 600 gpurun_out/bench_minecraft_4k.log
timeout 900 python bench.py --workload sponza_4k --steps 20 --cpu-budget 6 > gpurun_out/bench_sponza_4k.log 2>&1; tail -c 600 gpurun_out/bench_sponza_4k.log
timeout 900 python bench.py --workload terrain_poses_1080p --steps 32 --cpu-budget 6 > gpurun_out/bench_terrain_poses_1080p.log 2>&1; tail -c 600 gpurun_out/bench_terrain_poses_1080p.log
timeout 600 python bench.py --extra --no-cpu-baseline --steps 30 > gpurun_out/bench_extra.log 2>&1; tail -c 500 gpurun_out/bench_extra.log
SCENES="dot_cube_1080p cpu_render_4k colonnade_4k terrain_512_8_4k minecraft_256_32_4k"
timeout 300 python tools/perf_probe.py $SCENES > gpurun_out/perf_probe_plain.log 2>&1; cat gpurun_out/perf_probe_plain.log
timeout 300 python tools/perf_probe.py --mips 3.4e38 $SCENES > gpurun_out/perf_probe_mips_max.log 2>&1; grep -v "MIP maps" gpurun_out/perf_probe_mips_max.log
timeout 300 python tools/perf_probe.py --mips frustum $SCENES > gpurun_out/perf_probe_mips_frustum.log 2>&1; grep -v "MIP maps" gpurun_out/perf_probe_mips_frustum.log
# launch list of the default bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_final5.csv python bench.py --steps 10 --warmup 3 > gpurun_out/ncu_launch_bench.log 2>&1
# full captures: the render kernel on the default workload and on the node-heavy sponza scene
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r01_render_dotcube_final5 python tools/perf_probe.py dot_cube_1080p > gpurun_out/ncu_full_final5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r01_sponza_2048_32_4k_v8 python tools/perf_probe.py sponza_2048_32_4k > gpurun_out/ncu_full_sponza.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck_smoke.log
