"""CPU: the parts of bench.py's contract that do not need a GPU — the reference arm (`--impl reference`: the CPU
restatement timed on the host cores) prints ONE JSON line with the keys the driver compares against our arm, both arms
describe the workload with the same `config` dict, and under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_ref(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--submap-points", "20000"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = run_ref()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "aligns/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "aligns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the workload description is the one our arm prints (bench.config_dict is the single source of both)
    sys.path.insert(0, ROOT)
    import bench
    cfg = bench.config_dict(d["config"]["n_source"], d["config"]["n_target"], d["config"]["pairs_cycled"], 1)
    assert cfg == d["config"] and d["config"]["n_target"] == 20000 and d["metric"] == bench.METRIC


def test_reference_arm_is_silent_on_other_ranks():
    assert run_ref({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
