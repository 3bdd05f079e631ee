#!/bin/bash
# Round-2 GPU call 3 (TWO B200s): the fused gather across real devices - tests, bench at N=2 in both wire formats and the NCCL comparison.
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/c3_gpus.txt
echo "== multi tests"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_cpp_mirror.py tests/test_gpu_gather.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/c3_pytest_multi.log
run() { # name, nproc, args...
  name=$1; n=$2; shift 2
  echo "== bench N=$n $*"
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n "$@" > gpurun_out/c3_bench_$name.json 2> gpurun_out/c3_bench_$name.err
  tail -c 2500 gpurun_out/c3_bench_$name.json; grep -v "^W\|^\*\*\*\|Setting OMP" gpurun_out/c3_bench_$name.err | tail -5
}
run n2_wire8 2 --steps 30 --no-extra --wire 8
run n2_wire12 2 --steps 30 --no-extra --wire 12 --no-cpu-baseline
run n2_nccl 2 --steps 30 --no-extra --mode tiles_nccl --no-cpu-baseline
run n2_default 2
echo "== N=1 on the same box"
timeout 300 python bench.py --steps 30 --no-extra --no-cpu-baseline > gpurun_out/c3_bench_n1.json 2>gpurun_out/c3_bench_n1.err; tail -c 1200 gpurun_out/c3_bench_n1.json
