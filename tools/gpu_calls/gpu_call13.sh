#!/bin/bash
# round 2, call 13: the root-goes-first gather test, then what a CUDA graph would buy (tools/graph_probe.py)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_gather.py -x -q -k "root_that_goes_first or any_order or on_one_device" > gpurun_out/c13_tests.log 2>&1; tail -4 gpurun_out/c13_tests.log
timeout 200 python tools/graph_probe.py > gpurun_out/r02_graph_probe.log 2>&1; tail -5 gpurun_out/r02_graph_probe.log
