"""The JSON line contract of bench.py: the committed GPU line of the final tree carries every key the driver and the judge
read, and the reference arm (runnable here: CPU port on the host cores) prints a line with the same metric / config."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_committed_gpu_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2k_pipeline_bench_A.json")))
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert d["metric"].startswith("images/sec") and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"] == "pipeline" and d["config"]["batch_per_gpu"] == 8 and "l2_policy" in d["config"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 8 * 1024 * 1024 * 3 and e["d2h_bytes_per_step"] == 8 * (100 * 6 + 100 * 784) * 4
    assert e["value"] != d["value"]                                         # measured separately, not a copy of `value`
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 0 and {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    a = d["e2e_agreement"]                                                  # same images through the CPU port and the GPU path
    assert a["images"] >= 4 and a["cpu"] > 0 and a["matched"] >= 0.95 * a["cpu"] and a["max_box_delta"] < 1e-3
    assert "1e-4" in d["config"]["mask_head"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_prints_the_same_metric_and_config():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                                  # ONE JSON line on stdout
    d = json.loads(lines[0])
    gpu = json.load(open(os.path.join(ROOT, "profiles", "r2k_pipeline_bench_A.json")))
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert (d["metric"], d["unit"], d["higher_is_better"]) == (gpu["metric"], gpu["unit"], gpu["higher_is_better"])
    for k in ("workload", "name", "image", "batch_per_gpu", "pre_nms", "rois", "detections", "model"):
        assert d["config"][k] == gpu["config"][k]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_roialign_workload_line_and_reference_arm():
    """configs[4] through bench.py: the committed GPU line carries the three byte models side by side, and the reference arm
    (oracle crop_and_resize on the host cores) prints the same metric."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r2j_bench_roialign.json")))
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline", "sweep"} <= set(d)
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    assert r["frac_dram"] is not None and r["frac_survey_model"] >= r["frac"]          # the survey model charges untouched pixels too
    assert len(d["sweep"]) == 30 and all(x["bytes_touched"] <= x["bytes_survey"] for x in d["sweep"])
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "roialign", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    ref = json.loads([l for l in p.stdout.splitlines() if l.strip()][0])
    assert ref["impl"] == "reference" and (ref["metric"], ref["unit"]) == (d["metric"], d["unit"]) and ref["cpu_baseline"]["cores"] >= 1
