"""The parts of bench.py's contract that can be checked without a GPU: the reference arm prints exactly one
JSON line with the agreed keys, and the synthetic workload is the one BASELINE.json names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "voxels/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("voxels/sec monoexp T2 fit") and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    # the unmodified reference from baseline/_ref when it is installed there, else the oracle port
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "dosma", "core"))
    assert cb["kind"] == ("reference" if have_ref else "port"), cb
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "voxels per step" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_workload_matches_baseline_json():
    sys.path.insert(0, ROOT)
    import bench

    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert "8 echoes" in base["metric"] and "384^3" in base["metric"]
    assert bench.ECHOES == 8 and bench.SHAPE == (384, 384, 384) and bench.BYTES_PER_VOXEL == 44
    x, y = bench.synth_numpy(1000, 0)
    assert y.shape == (8, 1000) and y.dtype.name == "float32" and list(x) == [10.0 * k for k in range(1, 9)]
