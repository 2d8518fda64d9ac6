"""bench.py's reference arm runs on CPU alone: the JSON line it prints must carry the keys of the bench contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--games", "512"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "rollout_plies_per_s" and line["unit"] == "plies/s"
    assert line["steps"] == 1 and line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["gpu_launches"] == 0 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_measured_capture_file_feeds_the_roofline_block():
    """roofline.traffic and the 'as issued' ALU counts are read from the committed ncu capture (profiles/ncu_traffic.json, written by
    profiles/ncu_traffic.py), not typed into bench.py: the file carries both launches, their source reports and plausible counts."""
    sys.path.insert(0, ROOT)
    import bench
    t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    for key in ("rollout_pair_kernel", "rollout_pair_kernel_forced"):
        e = t[key]
        assert e["games_per_launch"] == bench.GAMES_PER_STEP and ".ncu-rep" in e["source"]
        assert e["alu_pipe_warp_inst_per_launch"] > 0 and e["warp_inst_per_launch"] >= e["alu_pipe_warp_inst_per_launch"]
        assert 0 < e["alu_pipe_pct"] <= 100 and e["kernel_us"] > 0 and e["captured"] and e["command"].startswith("python bench.py")
    full = t["rollout_pair_kernel"]
    assert full["dram_bytes_per_launch"] == full["dram_bytes_read"] + full["dram_bytes_write"]
    assert bench.ncu_traffic()["dram_bytes_per_launch"] == full["dram_bytes_per_launch"]
    plies = 59.84 * bench.GAMES_PER_STEP            # one launch of the bench workload (59.8 stones per game)
    ops = bench.alu_lane_ops_per_ply("rollout_pair_kernel", plies)
    ops_rules = bench.alu_lane_ops_per_ply("rollout_pair_kernel_forced", plies)
    assert 200 < ops_rules < ops < 1000              # rules alone issue fewer ALU operations than rules + policy + sampling
    assert bench.alu_lane_ops_per_ply("no_such_kernel", plies) is None


def test_committed_gpu_bench_line_has_the_contract_keys():
    """The newest committed bench record of the GPU arm (profiles/r02_bench_v*.json, written by `python bench.py` on a B200) carries
    every key the bench contract names, the end-to-end block counts its transfers, and the roofline block is self-consistent."""
    import glob
    import json
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [f for f in glob.glob(os.path.join(root, "profiles", "r02_bench_v*.json")) if re.search(r"_v\d+\.json$", f)]
    assert files
    newest = max(files, key=lambda f: int(re.search(r"_v(\d+)\.json$", f).group(1)))
    line = json.loads(open(newest).read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, (newest, key)
    assert line["metric"] == "rollout_plies_per_s" and line["unit"] == "plies/s" and line["n_gpus"] == 1 and line["warmup"] >= 3
    assert "workload" in line["config"] and "model" not in line["config"] and line["gpu_launches"] == line["steps"]
    e2e = line["e2e"]
    assert e2e["unit"] == line["unit"] and e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert abs(e2e["value"] - line["value"]) > 1e-6 * line["value"]          # measured on its own, not a copy of the device-timed value
    rf = line["roofline"]
    assert rf["bound"] in ("alu", "hbm", "tensor") and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and 0 < rf["frac"] <= 1
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
