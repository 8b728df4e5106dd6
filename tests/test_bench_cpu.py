"""bench.py's reference arm (`--impl reference`: the CPU oracle port timed on the host cores) runs without a
GPU, so its JSON contract is checked here: same metric / unit / config as the GPU arm, `impl`, a `cpu_baseline`
describing the run, a zero-copy `e2e`, and rank != 0 exits silently under torchrun."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(env_extra):
    env = dict(os.environ, DSEP_REF_BUDGET_S="15", **env_extra)     # a short sample: the contract is what is checked
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["unit"] == "utt/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["metric"].startswith("separated utterances/sec")
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None
    assert line["data"] == "synthetic" and "configs[1]" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
