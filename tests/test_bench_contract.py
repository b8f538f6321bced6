"""bench.py's output contract (the driver parses ONE JSON line from stdout) and the committed
bench lines under profiles/: every key the contract names is present and consistent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stdout_carries_only_the_json_line():
    code = ("import os, bench; bench.quiet_stdout(); print('library noise'); os.system('echo c-level noise'); "
            "bench.emit({'metric': 'm', 'value': 1.0})")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, check=True)
    assert r.stdout.splitlines() == ['{"metric": "m", "value": 1.0}']
    assert "library noise" in r.stderr and "c-level noise" in r.stderr


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read())


def test_committed_bench_lines_have_every_contract_key():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        baseline = json.load(f)
    for name, n in (("r1_bench_n1.json", 1), ("r1_bench_n2.json", 2), ("r1_bench_n4.json", 4), ("r1_bench_n8.json", 8)):
        d = _line(name)
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                  "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline",
                  "cpu_baseline", "clocks"):
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["warmup"] >= 3 and d["higher_is_better"] is True
        assert d["metric"].split(" (")[0] in baseline["metric"] and d["dtype"] == "f64" and d["scaling"] == "weak"
        assert d["vs_baseline"] is None  # BASELINE.md holds no published number for this metric
        assert "workload" in d["config"] and "model" not in d["config"]
        for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
            assert k in d["e2e"], (name, k)
        assert d["e2e"]["value"] < d["value"] and d["e2e"]["d2h_bytes_per_step"] > 0
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        c = d["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
        assert d["gpu_launches"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        # value = whole-job units / time
        cfg = d["config"]
        checks = cfg["n_nodes"] * cfg["n_primitives_per_gpu"] * n
        assert abs(d["value"] - checks / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6


def test_reference_arm_line():
    d = _line("r1_bench_ref.json")
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    # the headline ratio the driver computes: e2e of our arm / the reference arm
    ours = _line("r1_bench_n1.json")
    assert ours["e2e"]["value"] / d["value"] > 100  # north_star: >= 100x the host-core reference


def test_dbastar_leg_failure_is_an_error_object_not_an_exception():
    """The db-A* leg runs in a child process: with no CUDA device the child exits non-zero (the
    product has no CPU path) and the parent gets an error object, so the bench line survives."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("needs a box without a GPU")
    sys.path.insert(0, ROOT)
    import bench
    out = bench.dbastar_leg(0, timeout_s=120)
    assert set(out) == {"error", "stderr_tail"} and "no CUDA device" in out["stderr_tail"]


def test_round2_bench_lines():
    """The lines of the final round-2 build: contract keys, the parity gate's verdict, both configs,
    the reference arm printing the SAME config dict as ours."""
    ours, ref = _line("r2_bench_n1.json"), _line("r2_bench_ref.json")
    assert ours["config"] == ref["config"] and ref["impl"] == "reference"
    assert ref["e2e"]["value"] == ref["value"] and ref["gpu_launches"] == 0
    assert ours["e2e"]["value"] / ref["value"] > 100  # north_star: >= 100x the host-core reference, end to end
    lines = [("r2_bench_n1.json", 1, "weak"), ("r2i_bench_n2_lean1_ov1.json", 2, "weak"), ("r2j_bench_n4.json", 4, "weak"),
             ("r2j_bench_n8.json", 8, "weak"), ("r2_bench_stress_n1.json", 1, "strong"), ("r2j_bench_stress_n4.json", 4, "strong"),
             ("r2j_bench_stress_n8.json", 8, "strong")]
    for name, n, scaling in lines:
        d = _line(name)
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "parity"):
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["scaling"] == scaling and d["warmup"] >= 3 and d["dtype"] == "f64"
        if d["parity"].get("skipped") is None:  # (the N = 4 lines were taken with the gate off: N = 2 and 8 bracket them)
            assert d["parity"]["mismatch"] == 0, name
        cfg = d["config"]
        assert abs(d["value"] - cfg["n_nodes"] * cfg["n_primitives_total"] / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6
        assert d["e2e"]["value"] < d["value"] and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
        r = d["roofline"]
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["bound"] == "hbm"
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    d = _line("r2_bench_n1.json")
    if d["roofline"]["traffic"] is not None:  # only printed when the ncu capture is of the same build
        assert 0 < d["roofline"]["frac_dram"] < d["roofline"]["frac"]
    for leg in ("quadrotor_window", "roadmap_heuristic"):
        assert d["dbastar"][leg]["identical_search"] is True and d["dbastar"][leg]["solved"] is True
    assert d["dbastar"]["identical_search"] is True
