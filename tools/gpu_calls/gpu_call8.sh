#!/bin/bash
# Round-2 8-GPU call: where the time of a tile-sharded frame goes at N=8 / N=4, the multi-GPU tests on all devices, one default bench run at N=8.
mkdir -p gpurun_out
nvidia-smi -L | wc -l | tee gpurun_out/c8_gpus.txt
tr() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 "$@" 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$"; }
echo "== probe N=8"
tr 8 tools/gather_probe.py sponza_4k 20 8 0,16,4,20,8 | tail -2 | tee gpurun_out/r02_gather_probe_n8.json
echo "== probe N=4"
tr 4 tools/gather_probe.py sponza_4k 20 8 0,16,20 | tail -2 | tee gpurun_out/r02_gather_probe_n4.json
echo "== bench N=8 default"
tr 8 bench.py --gpus 8 > gpurun_out/c8_bench_n8.json; tail -c 1500 gpurun_out/c8_bench_n8.json
echo "== bench N=8 wire 12"
tr 8 bench.py --gpus 8 --wire 12 --no-extra --no-cpu-baseline --steps 30 > gpurun_out/c8_bench_n8_wire12.json; head -c 400 gpurun_out/c8_bench_n8_wire12.json; echo
echo "== multi tests on 8 devices"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_cpp_mirror.py -q -m gpu 2>&1 | tail -6 | tee gpurun_out/c8_pytest_multi.log
