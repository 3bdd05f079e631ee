#!/bin/bash
# one B200: the whole gpu suite after the gather / block-order work, schedule probe of the final build, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/c5_pytest_gpu.log
timeout 300 python tools/schedule_probe.py sponza_4k minecraft_4k 2>&1 | tail -3 | head -2 | tee gpurun_out/r02_schedule_probe_v5.json | cut -c1-900
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
