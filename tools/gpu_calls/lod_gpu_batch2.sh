set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lod.py -x -q 2>&1 | tail -15
timeout 300 python tools/fuzz_parity.py --mips --seconds 120 --seed 21 2>&1 | tail -5
SCENES="dot_cube_1080p cpu_render_4k colonnade_4k terrain_512_8_4k minecraft_256_32_4k"
timeout 300 python tools/perf_probe.py --mips 3.4e38 $SCENES > gpurun_out/perf_probe_mips_max.log 2>&1; cat gpurun_out/perf_probe_mips_max.log
timeout 300 python tools/perf_probe.py --mips frustum $SCENES > gpurun_out/perf_probe_mips_frustum.log 2>&1; cat gpurun_out/perf_probe_mips_frustum.log
timeout 300 python tools/perf_probe.py --mips 1000 $SCENES > gpurun_out/perf_probe_mips_1000.log 2>&1; cat gpurun_out/perf_probe_mips_1000.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_lod_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r01_render_lod_terrain_512_8_4k_v2 python tools/perf_probe.py --mips 1000 terrain_512_8_4k > gpurun_out/ncu_lod.log 2>&1; tail -3 gpurun_out/ncu_lod.log
