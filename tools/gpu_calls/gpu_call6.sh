#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/c6_pytest_gpu_$i.log; done
