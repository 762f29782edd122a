"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU, prints exactly one JSON line with the contract's keys,
uses every host core even when launched like torchrun does (OMP_NUM_THREADS=1), and only rank 0 prints under a multi-rank launch."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CMD = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--rows", "20000", "--cpu-sample-rows", "20000", "--steps", "2",
       "--warmup", "1", "--no-pipeline"]


def _run(extra_env):
    env = dict(os.environ, **extra_env)
    r = subprocess.run(CMD, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_json_line_and_threads():
    lines = _run({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "1"})
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1), "the CPU baseline must lift torchrun's OMP_NUM_THREADS=1"
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
