"""CPU checks of bench.py's bookkeeping (the numbers the JSON line is built from); the timed paths run on the GPU box."""
import json
import os

import bench


def test_algorithmic_flops_match_the_survey():
    """308.2 TFLOP per CFG step at config 2 (L = 11,648), 2210 TFLOP at the long-clip config 5."""
    cfg = bench.real_cfg()
    assert abs(bench.step_flops(cfg) / 1e12 - 308.2) < 0.1
    saved = bench.GRID
    try:
        bench.GRID = (49, 44, 80)
        assert abs(bench.step_flops(cfg) / 1e12 - 2210.0) < 1.0
        bench.GRID = (5, 16, 28)
        assert abs(bench.step_flops(cfg) / 1e12 - 13.5) < 0.1          # config 1
    finally:
        bench.GRID = saved


def test_traffic_record_and_peaks():
    t = bench.ncu_traffic()
    assert t is not None and t["dram_bytes_per_launch"] > t["algorithmic_bytes_per_launch"] > 0
    assert len(t["launches"]) == 4 and all(v["ratio"] >= 1.0 for v in t["launches"].values())
    M, N, K = 23296, 14336, 3072
    assert t["algorithmic_bytes_per_launch"] == 2 * (M * K + N * K + M * N + N)
    sustained, burst, hbm, src = bench.measured_peaks()
    assert 1000 < sustained <= burst < 2500 and 4000 < hbm < 9000 and src in ("measured", "fallback")


def test_committed_bench_lines_follow_the_contract():
    """The bench lines committed under profiles/ carry every key the driver and the judge read."""
    root = os.path.join(os.path.dirname(os.path.abspath(bench.__file__)), "profiles")
    for name, n in (("bench_r1l.json", 1), ("bench_r1t_n2.json", 2), ("bench_r1o_n4.json", 4), ("bench_r1p_n8.json", 8),
                    ("bench_r2b.json", 1), ("bench_r2c_cfg2.json", 2), ("bench_r2c_sp2.json", 2), ("bench_r2d_n8.json", 8),
                    ("bench_r2s.json", 1), ("bench_r2g_sp2.json", 2), ("bench_r2x_n2.json", 2), ("bench_r2m_n4.json", 4), ("bench_r2l_n8.json", 8)):
        text = open(os.path.join(root, name)).read().strip().splitlines()[-1]
        d = json.loads(text)
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
            assert key in d, (name, key)
        assert d["n_gpus"] == n and d["metric"] == bench.METRIC and d["config"]["workload"] == bench.WORKLOAD
        assert abs(d["value"] * d["ms_per_step"] - 1e3) < 1e-6 and d["gpu_launches"] > 0
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
        if n == 1:
            assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
        if name.startswith("bench_r2"):      # round 2: parity at every N, the same output checksum on every layout
            assert d["parity"]["ok"] and d["parity"]["rel_l2_vs_oracle"] <= 1e-2
            # the control fuser's implicit-GEMM convolution (run G onwards) sums its K range tap-major: new checksum, again
            # the same at every N
            final = name in ("bench_r2s.json", "bench_r2g_sp2.json", "bench_r2x_n2.json", "bench_r2m_n4.json", "bench_r2l_n8.json")
            assert d["parity"]["checksum_sha256_16"] == ("66064cea93a14880" if final else "219599d7e176cae3")
            if n == 1:
                assert d["library_baseline"]["ms_per_step"] > d["ms_per_step"] and d["cpu_baseline"]["config1_forward"]["finite"]
