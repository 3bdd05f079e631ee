#!/bin/bash
# 8 GPUs: staged (128-byte row) peer stores against direct stores, both wire formats, with and without the block order
mkdir -p gpurun_out
tr() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29591 "$@" 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version"; }
echo "== probe N=8 staged(0) direct(32) local(4)"; tr 8 tools/gather_probe.py sponza_4k 20 8 0,32,4 | tail -1 > gpurun_out/r02_gather_probe_n8_staged.json; head -c 2500 gpurun_out/r02_gather_probe_n8_staged.json; echo
echo "== probe N=4"; tr 4 tools/gather_probe.py sponza_4k 20 8 0,32 | tail -1 > gpurun_out/r02_gather_probe_n4_staged.json; head -c 1500 gpurun_out/r02_gather_probe_n4_staged.json; echo
echo "== D2H ceiling"; tr 8 tools/d2h_ceiling.py | tail -1 | tee gpurun_out/r02_d2h_ceiling.json
