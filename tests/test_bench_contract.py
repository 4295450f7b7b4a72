"""bench.py's JSON contract: the reference arm on the CPU (bounded sample), our arm on a tiny workload on the GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    env = dict(os.environ)
    env.pop("RANK", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_precision_bookkeeping_of_the_roofline_note():
    """executed MMAs per algorithmic product of the chain kernels, per precision (what executed_frac multiplies by)."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.chain_product_units("f16x3") == 3 and bench.chain_product_units("f16f8") == 2 and bench.chain_product_units("bf16") == 1
    assert bench.mix_layers("f16x3") is None and bench.mix_layers("f16mix") == ["qstn_max", "stn_max"]
    assert bench.mix_layers("f16mix:enc_max+qstn_small") == ["enc_max", "qstn_small"]
    u = bench.chain_product_units("f16mix")            # two of the three 128 -> 1024 layers at one product
    assert abs(u - (3 * (131072 + 8192 + 12288 + 16384) + 2 * 131072) / (3 * 131072 + 8192 + 12288 + 16384)) < 1e-12
    assert 1.7 < u < 1.9 and bench.chain_product_units("f16mix:qstn_max+stn_max+enc_max") < 1.2


def test_reference_arm_json_contract():
    line = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--points", "4096", "--ref-sample", "64"], 600)
    assert COMMON <= set(line)
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["higher_is_better"] is True
    assert line["metric"] == "normals_per_sec_deepfit_jet3_k256" and line["unit"] == "points/s"
    assert line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


@pytest.mark.gpu
def test_own_arm_json_contract_small():
    line = _run(["--steps", "1", "--warmup", "1", "--points", "8192", "--chunk", "4096", "--ref-sample", "64"], 900)
    assert COMMON | {"gpu_launches", "clocks", "roofline"} <= set(line)
    assert "impl" not in line or line["impl"] != "reference"
    assert line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 8192 * 12 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0
    rf = line["roofline"]
    assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and 0 < rf["frac"] < 0.6 and rf["peak"] > 0
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
