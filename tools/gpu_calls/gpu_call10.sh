#!/bin/bash
# round 2, call 10 (and 12, with the pool unit): lane-refill schedule - parity tests, then the A/B (tools/refill_probe.py)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_refill.py -x -q > gpurun_out/r02_refill_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02_refill_tests.log
tail -5 gpurun_out/r02_refill_tests.log
timeout 280 python tools/refill_probe.py > gpurun_out/r02_refill_probe.log 2>&1
echo "probe exit $?" >> gpurun_out/r02_refill_probe.log
tail -8 gpurun_out/r02_refill_probe.log
