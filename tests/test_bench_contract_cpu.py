"""bench.py's reference arm (`--impl reference`: the CPU port of the reference step on the host cores) prints ONE JSON line
with the keys the driver reads.  Runs on a few seconds of CPU work (PCSEG_REF_BUDGET_S)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, extra_env=None):
    env = dict(os.environ, PCSEG_REF_BUDGET_S="3")
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_json_line():
    lines = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["metric"] == "segmentation points/sec (fwd+bwd train step)" and d["value"] > 0 and d["steps"] == 1
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    # a bounded sample (3 s budget here) must be SAID in the workload string, never labelled as the full batch
    assert d["config"]["same_config"] is False and "bounded sample" in d["config"]["workload"]
    assert d["config"]["sample_batch"][0] * d["config"]["sample_batch"][1] < 8 * 16384
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    """under torchrun only rank 0 runs and prints the reference arm"""
    assert _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], {"RANK": "1", "WORLD_SIZE": "2"}) == []


def test_roofline_traffic_comes_from_the_committed_ncu_capture():
    """bench.py's roofline.traffic (dram read + write bytes per launch of the dominant kernels) is parsed from the `# traffic`
    lines of the committed `ncu --set full` summary, not hard-coded: the file must be there and carry the four global_feat
    kernels (forward, data gradient, Gram matrix, inference forward) with plausible sizes (>= the 268 MB of one 1024-channel
    bf16 tensor at 8 x 16 384 points)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_bench_mod", os.path.join(root, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert os.path.exists(os.path.join(root, mod.NCU_TRAFFIC_SOURCE)), mod.NCU_TRAFFIC_SOURCE
    t = mod.load_ncu_traffic()
    for tag in (5, 21, 53, 69):
        assert tag in t and 260.0 < t[tag] < 700.0, (tag, t)
    assert t[21] > t[5]          # the data gradient also writes its 268 MB result
