"""bench.py's host logic without a GPU: the algorithmic FLOP / byte figures the roofline uses are the ones SURVEY.md section 8(d)
states for the BASELINE configs, the workload table names those configs, and the reference arm (the oracle port on the host cores)
prints the JSON line the measurement contract asks for."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def bench():
    import bench as b
    return b


def test_algorithmic_flops_match_the_survey(bench):
    """C2 0.1374 TFLOP; C3 fwd 0.5498, fwd+bwd 1.9241 TFLOP; C4 5.498 TFLOP per GPU; C5 8.796 TFLOP."""
    tf = lambda name: tuple(x / 1e12 for x in bench.algorithmic_flops(bench.WORKLOADS[name]))
    assert tf("c2")[0] == pytest.approx(0.1374, rel=1e-3) and tf("c2")[1] == 0.0
    f, b = tf("c3")
    assert f == pytest.approx(0.5498, rel=1e-3) and f + b == pytest.approx(1.9241, rel=1e-3) and b == pytest.approx(2.5 * f)
    assert tf("c4")[0] == pytest.approx(5.498, rel=1e-3)
    assert tf("c5")[0] == pytest.approx(8.796, rel=1e-3)


def test_algorithmic_bytes_match_the_survey(bench):
    """C2 0.135 GB; C3 fwd+bwd 1.62 GB; C4 2.7 GB per GPU; C5 0.541 GB."""
    gb = lambda name: sum(bench.algorithmic_bytes(bench.WORKLOADS[name])) / 1e9
    assert gb("c2") == pytest.approx(0.135, rel=1e-2)
    assert gb("c3") == pytest.approx(1.62, rel=1e-2)
    assert gb("c4") == pytest.approx(2.7, rel=1e-2)
    assert gb("c5") == pytest.approx(0.541, rel=1e-2)


def test_workloads_are_the_baseline_configs(bench):
    cfg = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    w = bench.WORKLOADS
    assert (w["c3"]["B"], w["c3"]["H"], w["c3"]["S"], w["c3"]["D"], w["c3"]["n"], w["c3"]["causal"], w["c3"]["dropout"]) == (4, 32, 4096, 128, 0.5, True, 0.1)
    assert "B=4 H=32 S=4096 D=128, n=0.5, causal + dropout=0.1" in cfg[2] and w["c3"]["desc"].startswith("BASELINE.json configs[2]")
    assert (w["c2"]["B"], w["c2"]["H"], w["c2"]["S"], w["c2"]["D"], w["c2"]["causal"], w["c2"]["bwd"]) == (8, 16, 2048, 64, False, False)
    assert "B=8 H=16 S=2048 D=64" in cfg[1]
    assert w["c4"]["B"] * w["c4"]["H"] * 8 == 64 * 40 and (w["c4"]["S"], w["c4"]["D"]) == (8192, 128) and "B=64 H=40 S=8192 D=128" in cfg[3]
    assert (w["c5"]["B"], w["c5"]["H"], w["c5"]["S"], w["c5"]["D"]) == (1, 16, 65536, 64) and "B=1 H=16 S=65536 D=64" in cfg[4]
    assert bench.METRIC == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]


def test_peaks_come_from_the_driver_file_or_the_stated_fallback(bench):
    p = bench.measured_peaks()
    assert p["burst"] >= p["sustained"] > 0 and p["hbm"] > 0
    assert ("MEASURED_PEAKS.json" in p["source"]) == os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json"))


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU only: the oracle port on the host cores) on a tiny budget."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-step-seconds", "0.5", "--workload", "c2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "reference" and d["unit"] == "TFLOP/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[1]" in d["config"]["workload"]
