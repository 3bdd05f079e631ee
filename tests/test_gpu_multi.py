"""N-GPU framebuffer == 1-GPU framebuffer, byte for byte. Needs >= 2 GPUs: skipped on a 1-GPU box."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

import shocovox_b200 as S

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("workload", ["cpu_render_1080p", "dot_cube_1080p"])
def test_tile_sharded_frame_equals_single_gpu(workload):
    n = S.cuda_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tools" / "multi_gpu_check.py"), "--workload", workload, "--steps", "5"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    report = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert report["all_ranks_ok"]
    assert report["nccl_gather"]["equal_to_single_gpu"] and report["fused_peer_stores"]["equal_to_single_gpu"]
