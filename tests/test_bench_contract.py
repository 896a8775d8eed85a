"""CPU: the reference arm of bench.py (`--impl reference`) must print one JSON line with the contract's keys; under a
multi-rank launch only rank 0 prints.  (The GPU arm's line is checked on the GPU box by the driver.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, NRT_BENCH_REF_STEP_S="0.5", **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_the_contract_line():
    from oracle import orc

    if not (orc.Reference.available(True) or os.path.exists(os.path.join(ROOT, "oracle", "liborc.so"))):
        pytest.skip("no oracle library built")
    lines = _run({})
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []


def test_every_tool_script_and_the_bench_parse():
    """tools/*.py, bench.py and __graft_entry__.py at least compile (they only run on the GPU box)."""
    import glob
    import py_compile

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "tools", "*.py"))) + [os.path.join(root, "bench.py"),
                                                                       os.path.join(root, "__graft_entry__.py")]
    assert len(files) > 20
    for f in files:
        py_compile.compile(f, doraise=True)
