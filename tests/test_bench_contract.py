"""The JSON-line contract of bench.py: the keys the driver reads, on the last committed GPU line (profiles/) and on a
live run of the reference arm (CPU: the oracle port on a bounded sample)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _check_common(line):
    assert BASE <= set(line), sorted(BASE - set(line))
    assert line["metric"] == "a2gnn_train_epochs_per_sec" and line["unit"] == "epochs/s"
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["dtype"] == "f32" and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])


def test_committed_gpu_line_has_the_contract_keys():
    raw = open(os.path.join(ROOT, "profiles", "r2_m_final", "bench.json")).read()
    line = json.loads([l for l in raw.splitlines() if l.startswith("{")][-1])
    _check_common(line)
    g = line["roofline"]["gemm"]["forward"]                  # round 2: the tile-packed first layer reports its own bound
    assert g["bound"] == "tensor" and g["unit"] == "TFLOP/s" and abs(g["frac"] - g["achieved"] / g["peak"]) < 1e-9
    assert line["value"] > 0 and line["gpu_launches"] > 0 and line["warmup"] >= 3
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["e2e"]["value"] != line["value"]
    r = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = line["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--ref-scale", "50"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    _check_common(line)
    assert line["impl"] == "reference" and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    c = line["cpu_baseline"]
    assert c["value"] == line["value"] and c["kind"] == "port" and c["cores"] >= 1 and c["sample"]
    # measured, and says what was run: the requested counts, honestly (the full-scale run is what the driver times;
    # this test uses the reduced-size knob and the line must say so)
    assert line["steps"] == 2 and line["warmup"] == 1 and c["estimated"] is False
    assert "REDUCED 1/50 scale" in c["sample"] and c["small_scale_estimate"]["estimated"] is True
    assert abs(line["ms_per_step"] * 1e-3 * line["value"] - 1.0) < 1e-9


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], cwd=ROOT,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stdout + out.stderr)
