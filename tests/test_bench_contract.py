"""CPU test of the bench.py output contract on the committed final line of the round (profiles/bench_r01_final.json is the
unmodified stdout of `python bench.py` on a B200): every key the driver and the judge read is present and consistent."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_final_bench_line_carries_the_contract():
    d = json.load(open(os.path.join(ROOT, "profiles", "bench_r01_final.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "fvp_samples_per_sec" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["data"] == "synthetic" and d["dtype"] == "f64"
    assert d["vs_baseline"] is None                       # BASELINE.md has no published number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = samples of 10 FVPs over the measured solve time
    assert abs(d["value"] - 10 * 1_000_000 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] >= 1_000_000 * 17 * 8 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                        # host copies inside the timed region
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 0.5 < r["frac"] < 1.0 and r["traffic"] is not None
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]
    assert d["gpu_launches"] > 0
    ck = d["clocks"]
    assert not set(ck["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert ck["sm_mhz"] > 0.9 * ck["sm_max_mhz"]
    also = d["also"]
    assert {"arm_50k", "loop_body_1m", "humanoid256_200k"} <= set(also)
    assert also["humanoid256_200k"]["kernel_path"] == "gemm_chain" and also["humanoid256_200k"]["roofline_frac"] > 0.5
