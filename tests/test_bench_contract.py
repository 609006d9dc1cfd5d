"""bench.py's reference arm runs on CPU only: one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-limit", "2", "--ref-instances", "2"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "placement instances/sec" and d["unit"] == "instances/s"
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=60, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_bench_line_and_traffic_record_are_consistent():
    """the bench line committed under profiles/ carries the contract's objects, and the DRAM-traffic record bench.py reads for
    `roofline.traffic` is the ncu capture of the kernel the line names (host logic only: nothing is run here)"""
    line = json.loads(open(os.path.join(ROOT, "profiles", "r02c_bench_default.json")).read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "quality"):
        assert key in line, key
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    assert abs(roof["achieved"] - roof["bytes_per_iteration"] / roof["us_per_iteration"] / 1e3) < 1e-6 * roof["achieved"]
    assert "k_mf_iter_bulk" in roof["kernel"] and line["gpu_launches"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0
    traffic = json.load(open(os.path.join(ROOT, "profiles", "r02c_pdhg_traffic.json")))
    assert "k_mf_iter_bulk" in traffic["kernel_build"]
    assert traffic["per_instance_dram_bytes"] * 256 == traffic["bulk_dram_read_bytes"] + traffic["bulk_dram_write_bytes"]
    # measured DRAM traffic stays below the algorithmic bytes (no re-read waste), and the bench scales it by its batch
    assert traffic["bulk_dram_read_bytes"] + traffic["bulk_dram_write_bytes"] < traffic["algorithmic_bytes_per_iteration_B256"]
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "r02c_pdhg_traffic.json" in src
