#!/bin/bash
# Round-end evidence on one B200, short form (~5 GPU-minutes): parity tests, smoke, default bench + reference arm, kernel-time
# probe, launch list of the bench command, two full ncu captures, sanitizer. Outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 200 python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 1500 gpurun_out/bench_default.log
timeout 120 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -c 300 gpurun_out/bench_reference.log
SCENES="dot_cube_1080p dot_cube_4k cpu_render_1080p cpu_render_4k colonnade_4k terrain_512_8_4k terrain_1024_8_1080p minecraft_256_32_4k minecraft_1024_32_4k sponza_2048_32_4k"
timeout 200 python tools/perf_probe.py $SCENES > gpurun_out/perf_probe_plain.log 2>&1; cat gpurun_out/perf_probe_plain.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_final9.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r01_render_dotcube_final9 python tools/perf_probe.py dot_cube_1080p > gpurun_out/ncu_full_final9.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r01_minecraft_1024_32_4k_v9 python tools/perf_probe.py minecraft_1024_32_4k > gpurun_out/ncu_full_minecraft.log 2>&1
timeout 150 python bench.py --mips frustum --steps 50 --no-cpu-baseline > gpurun_out/bench_dot_cube_mips_frustum.log 2>&1; tail -c 400 gpurun_out/bench_dot_cube_mips_frustum.log
timeout 200 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck_smoke.log
# occupancy variants (9 / 10 CTAs per SM) against the default, if the libraries were built next to it
if [ -f shocovox_b200/libmb9.so ]; then
  timeout 150 bash tools/probe_variants.sh "- mb9 mb10" dot_cube_1080p cpu_render_4k colonnade_4k terrain_512_8_4k minecraft_1024_32_4k sponza_2048_32_4k > gpurun_out/r01_experiment_register_caps_v3.log 2>&1; cat gpurun_out/r01_experiment_register_caps_v3.log
fi
