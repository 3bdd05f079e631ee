#!/bin/bash
# 8 GPUs: scaling table of the default bench (N = 1, 2, 4, 8) with the final gather / block-order defaults, and what the order buys at N = 8
mkdir -p gpurun_out
tr() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29571 "$@" 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version"; }
echo "== N=1"; timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/c8b_n1.json 2>gpurun_out/c8b_n1.err; head -c 300 gpurun_out/c8b_n1.json; echo
for n in 2 4 8; do echo "== N=$n"; tr $n bench.py --gpus $n --no-cpu-baseline > gpurun_out/c8b_n$n.json; head -c 420 gpurun_out/c8b_n$n.json; echo; done
echo "== N=8 without the block order"; SVX_CTA_ORDER=0 tr 8 bench.py --gpus 8 --no-cpu-baseline --no-extra --steps 30 > gpurun_out/c8b_n8_raster.json; head -c 420 gpurun_out/c8b_n8_raster.json; echo
echo "== N=8 wire 12"; tr 8 bench.py --gpus 8 --no-cpu-baseline --no-extra --steps 30 --wire 12 > gpurun_out/c8b_n8_wire12.json; head -c 420 gpurun_out/c8b_n8_wire12.json; echo
echo "== N=8 minecraft"; tr 8 bench.py --gpus 8 --no-cpu-baseline --no-extra --steps 30 --workload minecraft_4k > gpurun_out/c8b_n8_minecraft.json; head -c 420 gpurun_out/c8b_n8_minecraft.json; echo
echo "== probe N=8"; tr 8 tools/gather_probe.py sponza_4k 20 8 0,4 | tail -1 > gpurun_out/r02_gather_probe_n8_final.json; head -c 1500 gpurun_out/r02_gather_probe_n8_final.json; echo
