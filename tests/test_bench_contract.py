"""CPU test of the bench.py output contract on the committed final line of the round (profiles/bench_r02_final.json is the
unmodified stdout of `python bench.py` on a B200): every key the driver and the judge read is present and consistent."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_final_bench_line_carries_the_contract():
    d = json.load(open(os.path.join(ROOT, "profiles", "bench_r02_final.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity"):
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
    # the roofline is that of the kernel that ran: 10 FVP passes per launch of the persistent solve kernel
    assert r["fvp_passes_per_launch"] == 10 and r["kernel_launches_timed"] == d["steps"]
    assert abs(r["achieved"] - 10 * 1_000_000 * r["flops_per_sample"] / (r["kernel_avg_ms"] * 1e-3) / 1e12) < 1e-9 * r["achieved"]
    assert r["traffic"] < 1.1 * r["hbm_context"]["algorithmic_bytes_per_launch"] * (1 + 6.8 / 136)
    assert d["parity"]["prefix_vs_reference_max_rel"] < 1e-8 and d["parity"]["prefix_states"] >= 10_000
    also = d["also"]
    assert {"arm_50k", "arm_1m", "loop_body_1m", "humanoid256_1m", "update_sweep", "file_dropins_50k"} <= set(also)
    h = also["humanoid256_1m"]
    assert h["fp64"]["roofline_frac"] > 0.5 and h["fp32_vs_fp64_speedup_fvp"] > 2.0
    assert h["fp32_fvp_error_vs_fp64_gpu_1m_states"]["rel_l2"] < h["fp32_fvp_error_vs_fp64_gpu_1m_states"]["stated_tolerance"]
    assert [r_["states"] for r_ in also["update_sweep"]["mlp64"]["rows"]] == [10_000, 100_000, 1_000_000, 4_000_000]
    assert also["file_dropins_50k"]["text_vs_binary_bitwise_equal"] is True
