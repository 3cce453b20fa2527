"""bench.py contract pieces that need no GPU: the reference arm (CPU oracle port) prints ONE JSON line with the keys the
driver reads, under torchrun only rank 0 works, and the roofline accounting helpers agree with the spec."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.strip().splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_json_line():
    lines = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-samples", "2"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "M=200" in d["metric"] and "c2" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    lines = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-samples", "2", "--gpus", "2"],
                 env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert lines == []


def test_flop_accounting_matches_survey():
    """2.930 GFLOP per sample per cell step at 64x64 with A = 8 (conv-LSTM 2.268), recomputed from the layer
    table; c2 = 8.20 TFLOP per CEM iteration."""
    sys.path.insert(0, ROOT)
    from visual_foresight_b200 import spec as S
    sp = S.spec_64(height=64, width=64, seq_len=15)
    fl = S.flops_per_sample_step(sp)
    assert abs(fl["total"] / 1e9 - 2.930) < 0.01 and abs(fl["conv_lstm"] / 1e9 - 2.268) < 0.005
    assert abs(S.flops_per_plan(sp, 200, 1) / 1e12 - 8.20) < 0.03
