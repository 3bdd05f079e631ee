#!/bin/bash
# round 2, call 11: ncu --set full of the lane-refill kernel on sponza 4K and terrain 1080p (the static kernel's captures
# are profiles/r02_sponza_2048_32_4k.summary.txt and r02_terrain_1024_8_1080p.summary.txt)
mkdir -p gpurun_out
export SVX_SCHEDULE=refill SVX_REFILL_STEPS=64 SVX_REFILL_MIN_IDLE=16 SVX_REFILL_UNIT=1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:render_kernel_refill --launch-skip 5 --launch-count 1 -f -o gpurun_out/r02_refill_sponza_2048_32_4k python tools/perf_probe.py sponza_2048_32_4k > gpurun_out/c11_ncu_sponza.log 2>&1; tail -2 gpurun_out/c11_ncu_sponza.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:render_kernel_refill --launch-skip 5 --launch-count 1 -f -o gpurun_out/r02_refill_terrain_1024_8_1080p python tools/perf_probe.py terrain_1024_8_1080p > gpurun_out/c11_ncu_terrain.log 2>&1; tail -2 gpurun_out/c11_ncu_terrain.log
ls -la gpurun_out/*.ncu-rep
