"""Post-processing tools that need no GPU keep working on the committed evidence."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layer_breakdown_on_the_committed_launch_list():
    """tools/layer_breakdown.py: the 130 conv launches of the ncu list line up with the layer plan (it asserts the
    count), and the totals agree with what bench.py measured live for the same tree (conv class 0.59 of the tensor peak)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "layer_breakdown.py"),
                        os.path.join(ROOT, "profiles", "r1w_pipeline_launches_warm.csv"), "--one-term-masks", "--unfused"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    total = [l for l in r.stdout.splitlines() if l.startswith("all conv launches")]
    assert len(total) == 1
    f = total[0].split()
    assert f[3] == "130" and 780.0 < float(f[6]) < 860.0            # TFLOP/s of the conv class
    pure = [l for l in r.stdout.splitlines() if l.startswith("# pure tensor roofline")][0]
    assert 0.55 < float(pure.split("=")[1].split()[0]) < 0.65
    assert "rpn shared 3x3" in r.stdout and "roialign_nhwc_kernel" in r.stdout


def test_layer_breakdown_on_the_fused_launch_list():
    """The final tree's launch list: 103 conv launches (27 of them the fused expansion + reduction kernel) line up with the plan."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "layer_breakdown.py"),
                        os.path.join(ROOT, "profiles", "r2k_pipeline_launches_warm.csv")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    total = [l for l in r.stdout.splitlines() if l.startswith("all conv launches")][0].split()
    assert total[3] == "103" and 850.0 < float(total[6]) < 950.0
    fused = [l for l in r.stdout.splitlines() if "(fused)" in l]
    assert len(fused) == 3 and sum(int(l.split("(fused)")[1].split()[0]) for l in fused) == 27
