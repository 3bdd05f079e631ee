#!/bin/bash
# 8 GPUs: the multi-GPU tests on all devices, then the scaling table of the default bench with the final gather defaults
mkdir -p gpurun_out
tr() { n=$1; shift; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29601 "$@" 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version"; }
echo "== multi tests"; timeout 700 python -m pytest tests/test_gpu_multi.py tests/test_cpp_mirror.py tests/test_gpu_gather.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_gputest_multi_8gpu.log
echo "== probe N=8"; tr 8 tools/gather_probe.py sponza_4k 20 8,4 0,32 | tail -1 > gpurun_out/r02_gather_probe_n8_rotated.json; head -c 1800 gpurun_out/r02_gather_probe_n8_rotated.json; echo
for n in 8 4 2; do echo "== N=$n"; tr $n bench.py --gpus $n > gpurun_out/r02_bench_n$n.json; head -c 330 gpurun_out/r02_bench_n$n.json; echo; done
echo "== N=1"; timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2>gpurun_out/r02_bench_n1.err; head -c 300 gpurun_out/r02_bench_n1.json; echo
echo "== reference N=8"; tr 8 bench.py --gpus 8 --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_n8.json; head -c 300 gpurun_out/r02_bench_reference_n8.json; echo
