"""bench.py's JSON contract on the legs that run without a GPU: the --impl reference arm (CPU oracle) and its
behaviour under a multi-rank launch (rank 0 prints, the others exit 0 without work)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    res = run(["--impl", "reference", "--workload", "cpu_render_150", "--steps", "3", "--warmup", "3"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "primary_mrays_per_s" and line["unit"] == "Mrays/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f32"
    assert line["steps"] == 3 and line["warmup"] >= 3 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"] == "cpu_render_150" and "model" not in line["config"]
    # the shared keys carry what our arm prints for the same launch: one whole frame per step, labelled like ours
    assert line["config"]["rays_per_step"] == 150 * 150 and line["config"]["resolution"] == [150, 150] and line["scaling"] == "strong"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    res = run(["--impl", "reference", "--gpus", "2", "--workload", "cpu_render_150", "--steps", "1"],
              env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_gpu():
    import shocovox_b200 as S

    if S.cuda_device_count() > 0:
        return
    res = run(["--steps", "1"])
    assert res.returncode != 0 and "no CUDA device" in res.stdout
