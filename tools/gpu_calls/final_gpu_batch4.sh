#!/bin/bash
# Round 2: (1) ncu --set full of the viewport kernel on BASELINE config 2 (dot_cube 1080p, glass at frustum.z) with the final
# kernels, caches flushed before the captured launch -> roofline.traffic of that workload; (2) randomised GPU-vs-oracle soak
# of the final library, plain and with MIP maps.
mkdir -p gpurun_out
timeout 80 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r02_dot_cube_1080p python tools/perf_probe.py dot_cube_1080p > gpurun_out/r02_ncu_dot_cube.log 2>&1; tail -2 gpurun_out/r02_ncu_dot_cube.log
timeout 45 python tools/fuzz_parity.py --seconds 30 --seed 20261018 > gpurun_out/r02_fuzz_gpu.log 2>&1; tail -2 gpurun_out/r02_fuzz_gpu.log; cp gpurun_out/fuzz_parity.json gpurun_out/r02_fuzz_parity.json 2>/dev/null
timeout 45 python tools/fuzz_parity.py --seconds 30 --seed 20261019 --mips > gpurun_out/r02_fuzz_gpu_lod.log 2>&1; tail -2 gpurun_out/r02_fuzz_gpu_lod.log; cp gpurun_out/fuzz_parity.json gpurun_out/r02_fuzz_parity_lod.json 2>/dev/null
