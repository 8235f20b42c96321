"""bench.py's reference arm (`--impl reference`): the one leg of the bench that runs without a GPU. Checks the JSON contract the
driver reads — exactly one line on stdout, the keys of the base contract plus `impl`, `cpu_baseline` and a zero-copy `e2e`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--scale", "0.0625", *args], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = [l for l in _run().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "proofs/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "no size extrapolation" in cb["sample"]
    # the test runs a 1/16-size layer to stay fast: the line itself must say that it is not the headline configuration
    assert d["same_config"] is False and d["same_steps"] is True and d["extrapolated"] is False
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_without_work():
    """Under torchrun only rank 0 runs the CPU arm; the other ranks print nothing and exit 0."""
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ("--gpus", "2")).strip() == ""
