"""bench.py on a CPU-only box: the reference arm runs (it times the compiled reference CPU path),
the product arm fails loudly (there is no CPU fallback), and the traffic model matches SURVEY §8d."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from fluidsimulator_b200 import capi  # noqa: E402


def test_reference_arm_prints_the_contract_line(built):
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--scene", "fluid_large",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "particle-substeps/s" and line["value"] > 0
    assert line["config"]["workload"] == "fluid_large" and line["config"]["particles"] == 19683
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0 and line["higher_is_better"] is True


def test_reference_arm_other_ranks_exit_quietly(built):
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback(built):
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--scene", "fluid_large", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_algorithmic_bytes_model():
    """B_alg = 254 + 44 I + 44 [xsph] + 92 [vorticity]: 430 / 474 / 566 B at I = 4."""
    assert bench.b_alg(4, bench.FLAGSETS["none"]) == 430
    assert bench.b_alg(4, bench.FLAGSETS["stable"]) == 474
    assert bench.b_alg(4, bench.FLAGSETS["all"]) == 566
    assert bench.b_alg(2, bench.FLAGSETS["all"]) == 478 and bench.b_alg(8, bench.FLAGSETS["all"]) == 742
    assert bench.ALG_BYTES["lambda"] == 16 and bench.ALG_BYTES["delta"] == 28


def test_roofline_traffic_is_keyed_by_regime_and_kernel_family():
    """bench.py's roofline.traffic comes from profiles/ncu_traffic.json, which
    tools/ncu_traffic_from_summaries.py regenerates from the committed ncu summaries of the kernels
    that ship: one figure per (scene, regime, stage, kernel family), none for what was never captured."""
    import bench
    t0 = bench.ncu_traffic("fluid_million", 0, "lambda", False)
    settled = bench.ncu_traffic("fluid_million", 200, "lambda", False)
    impact = bench.ncu_traffic("fluid_million", 100, "lambda", False)      # the default timed regime
    assert 1.8e8 < impact < 2.2e8
    brick = bench.ncu_traffic("fluid_million", 0, "lambda", True)
    assert 1.0e8 < t0 < 1.5e8 and 1.4e8 < settled < 1.8e8     # bytes per launch, 1 M particles
    assert brick < 0.7 * t0                                   # 16-bit lists: what the brick family does save
    assert bench.ncu_traffic("fluid_million", 100, "lambda", True) is None
    assert bench.ncu_traffic("fluid_large", 0, "lambda", False) is None
