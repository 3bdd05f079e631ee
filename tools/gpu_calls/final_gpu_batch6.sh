#!/bin/bash
# Round 2: the rest of the GPU suite (everything final_gpu_batch5.sh did not run) on the final library.
mkdir -p gpurun_out
timeout 95 python -m pytest tests -x -q -m gpu --ignore=tests/test_gpu_fullsize.py --ignore=tests/test_gpu_vox.py --ignore=tests/test_gpu_parity.py 2>&1 | tail -5 > gpurun_out/r02_gputest_1gpu_final3_rest.log; cat gpurun_out/r02_gputest_1gpu_final3_rest.log
