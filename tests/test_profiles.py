"""The committed measurement records keep the shape the bench contract asks for (CPU-only consistency check)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


def test_bench_records_carry_the_contract_keys():
    for name, gpus in (("bench_r01_s4.json", 1), ("bench_r01_s3.json", 1), ("bench_r01_s2.json", 1), ("bench_r01_s2_2gpu.json", 2),
                       ("bench_r01_s2_8gpu.json", 8)):
        d = _load(name)
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
            assert key in d, (name, key)
        assert d["n_gpus"] == gpus and d["scaling"] == "weak" and d["higher_is_better"] is True and d["vs_baseline"] is None
        assert d["unit"] == "proofs/s" and "workload" in d["config"] and d["gpu_launches"] > 0
        assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
        assert d["e2e"]["h2d_bytes_per_step"] == d["config"]["msm_terms_per_gpu"] * 64
        assert 0 < d["e2e"]["value"] <= d["value"] * 1.001
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] >= 0.9 * d["clocks"]["sm_max_mhz"]
        assert d["accept_bits"] == [1] * gpus
        assert d["value"] > 0.9 * gpus * 15.5e6          # weak scaling holds


def test_roofline_traffic_record_matches_the_bench_size():
    p = _load("r01_ingest_ncu.json")
    d = _load("bench_r01_s2.json")
    assert p["n_terms"] == d["config"]["msm_terms_per_gpu"]
    assert abs(p["traffic_bytes_per_launch"] - sum(p["traffic_bytes_each_launch"]) / 2) < 1
    assert d["roofline"]["traffic"] == p["traffic_bytes_per_launch"]
    assert p["algorithmic_bytes_per_launch"] == d["roofline"]["algorithmic_bytes_per_launch"]


def test_cpu_baseline_of_the_latest_record_names_its_backend():
    c = _load("bench_r01_s4.json")["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and "backend" in c
    assert c["value"] >= c["u64_serial_backend_value"] > c["single_thread_value"] > 0


def test_end_to_end_step_is_within_two_percent_of_the_device_resident_one():
    d = _load("bench_r01_s4.json")
    assert d["e2e"]["ms_per_step"] <= 1.02 * d["ms_per_step"]
