"""CPU checks of bench.py's line contract on the committed lines under profiles/ (the driver parses the same keys at round end):
metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better / scaling / dtype / data / config.workload,
the roofline and cpu_baseline objects, e2e with its copy sizes, clocks sampled inside the timed region, gpu_launches, and the
self-verification keys -- and that the numbers in a line agree with each other."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n", [("bench_r2_n1.json", 1), ("bench_r2_n2.json", 2), ("bench_r2_n4.json", 4), ("bench_r2_n8.json", 8)])
def test_default_line_contract(name, n):
    d = line(name)
    assert d["metric"] == "iq_msamples_per_s" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["steps"] >= 1 and d["warmup"] >= 3 and "workload" in d["config"] and "model" not in d["config"]
    # value = units all ranks processed / the timed region
    per_step = n * 1024 * 2_590_000
    assert d["value"] == pytest.approx(per_step / (d["ms_per_step"] * 1e-3) / 1e6, rel=1e-6)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert 0.5 < r["frac"] <= 1.05 and r["kernel_ms"] < d["ms_per_step"]           # the kernel is part of the step
    assert r["kernel_ms_min"] <= r["kernel_ms_median"] <= r["kernel_ms_max"]
    assert r["traffic"] is None or r["traffic"] >= 0.99 * r["achieved"] * r["kernel_ms"] * 1e6     # DRAM bytes >= algorithmic bytes
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert 0.9 <= e["frac_of_ceiling"] <= 1.02 and e["h2d_ceiling_gbs"] > 0
    c = d["clocks"]
    assert c["sampled"] == "inside the timed region" and c["sm_mhz"] <= c["sm_max_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] > 0
    k = d["check"]
    assert k["decoded_exact_all_ranks"] == k["bulletins_expected_all_ranks"] == n * 1024
    assert k["gathered_multiset_equals_union_of_expected"] is True and k["ranks_with_exact_multiset"] == n
    p = k["oracle_parity"]
    assert p["pass"] is True and p["streams"] >= 64 and p["bit_mismatches_occupied"] == 0 and p["y3_max_rel_pair_peak"] <= 1e-5
    for key in ("config4", "config5", "int16_input", "channels"):
        assert key in d, key
    assert d["config4"]["check"]["decoded_exact_all_ranks"] == d["config4"]["streams_total"] == n * 8192
    assert d["config5"]["check"]["decoded_exact_all_ranks"] == n * 1024 and d["config5"]["long_tc_fallbacks"] == 0
    if n == 1:
        b = d["cpu_baseline"]
        assert b["kind"] == "reference" and b["cores"] >= 1 and b["value"] > 0 and b["unit"] == d["unit"] and b["streams_in_sample"] >= 64


@pytest.mark.parametrize("taps", [65, 127, 255, 383, 511])
def test_config5_line_contract(taps):
    d = line("bench_r2_config5_%dtaps.json" % taps)
    assert d["config"]["taps"] == [taps] * 3 and d["check"]["decoded_exact_all_ranks"] == 1024 and d["long_tc_fallbacks"] == 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and r["frac"] < r["executed_frac"] < 1
    h = r["hbm"]
    assert h["y1_rows_per_stream"] == 1 and h["frac"] == pytest.approx(h["achieved"] / h["peak"], rel=1e-9) and 0.3 < h["frac"] < 1
    assert h["algorithmic_bytes_per_sample"] == pytest.approx(8 + 2 * 8 / 4 + 4 * 8 / 28 + 2 * 8 / 280)
    assert r["kernel_ms"] < d["ms_per_step"] and d["clocks"]["sampled"] == "inside the timed region"


def test_reference_arm_line_contract():
    d = line("bench_r2_reference_n1box.json")
    assert d["impl"] == "reference" and d["metric"] == "iq_msamples_per_s" and d["higher_is_better"] is True
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
