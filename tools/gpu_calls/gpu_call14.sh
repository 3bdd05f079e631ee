#!/bin/bash
# round 2, call 14 (2 GPUs): graph probe again now that frames queued without a frame out carry no timing events, the multi-GPU
# tests at world 2, and the 2-GPU bench line with untimed peers
mkdir -p gpurun_out
timeout 120 python tools/graph_probe.py > gpurun_out/r02_graph_probe_after.log 2>&1; tail -1 gpurun_out/r02_graph_probe_after.log
timeout 200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_gather.py -x -q > gpurun_out/c14_tests.log 2>&1; tail -3 gpurun_out/c14_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-extra > gpurun_out/r02_bench_n2_untimed_peers.json 2> gpurun_out/c14_bench.err; tail -c 400 gpurun_out/c14_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n2_untimed_peers.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "ms_per_step_per_rank", "gpu_launches")}, d["config"].get("gathered_frame_equals_single_gpu_frame"), d["e2e"]["value"], d["e2e"].get("host_assembled_frame_equals_single_gpu_frame"))
PY
