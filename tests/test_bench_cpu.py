"""bench.py host logic that needs no GPU: which rows make up the job, how the default slab is sharded over 1/2/4/8
ranks (strong scaling), the workload name the JSON line carries, and the contract of the reference arm's line."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from v2r_b200.distributed import shard_rows  # noqa: E402


def _args(**kw):
    d = dict(workload="c4", slab="", scaling="strong", exps="", band_rows=0)
    d.update(kw)
    return argparse.Namespace(**d)


def test_default_job_is_the_c4_slab_split_into_equal_row_bands():
    a = _args()
    assert bench.job_rows(a, 16384) == (7168, 2048)
    name = bench._workload_name(a, 8)
    assert "c4" in name and "slab rows [7168,9216)" in name and "p=2.0" in name
    for world in (1, 2, 4, 8):
        bands = [shard_rows(2048, world, r) for r in range(world)]
        assert [b[1] for b in bands] == [2048 // world] * world          # equal shares, whole CTA tiles
        assert bands[0][0] == 0 and all(bands[i][0] + bands[i][1] == bands[i + 1][0] for i in range(world - 1))


def test_other_workloads_and_overrides():
    assert bench.job_rows(_args(workload="c2"), 4096) == (0, 4096)
    assert bench.job_rows(_args(slab="100:50"), 16384) == (100, 50)
    assert bench.job_rows(_args(slab="16380:50"), 16384) == (16380, 4)          # clipped to the grid
    assert bench.job_rows(_args(workload="c4", scaling="weak"), 16384 * 2) == (0, 32768)
    assert "[exps overridden: 1.7]" in bench._workload_name(_args(workload="c2", exps="1.7"), 1)
    assert bench.ALGO_SLOTS["c4"] == 8.0 and bench.ALGO_SLOTS["c3"] == 38.0


def test_traffic_lookup_is_keyed_by_kernel_and_launch_shape():
    doc = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    cap = [c for c in doc["captures"] if c["rows"] == 2048][0]
    assert bench.lookup_traffic(cap["kernel"], 2048, 16384, 967142) == cap["dram_bytes"]
    assert bench.lookup_traffic(cap["kernel"], 2047, 16384, 967142) is None          # another launch shape: no number
    assert bench.lookup_traffic("fp64/other kernel", 2048, 16384, 967142) is None


def test_reference_arm_prints_the_contract_line():
    """--impl reference: one JSON line with impl, metric, unit, cpu_baseline{kind, cores, sample, value} and an e2e block
    that repeats the value with zero copies; non-zero ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"], capture_output=True, text=True, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "1",
                        "--warmup", "1", "--cpu-seconds", "0.5"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "points" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # `config` is the job, built by the same function as the GPU arm's (the sample is described in cpu_baseline.sample)
    a = _args(workload="c2")
    assert line["config"] == bench.make_config(a, 1, 99699, 1, 0, 4096, 4096, 4096)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--gpus", "8", "--steps", "1",
                        "--warmup", "0", "--cpu-seconds", "0.3"], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, RANK="0", WORLD_SIZE="8"))
    assert p.returncode == 0, p.stderr[-2000:]
    cfg = json.loads(p.stdout.strip().splitlines()[-1])["config"]
    assert cfg == bench.make_config(a, 8, 99699, 1, 0, 4096, 512, 4096) and cfg["parallelism"] == "row bands x8"


def test_config_of_the_default_job():
    """what the driver's run prints as `config` at N = 1 and N = 8 (both arms call make_config with these values)"""
    a = _args()
    c1 = bench.make_config(a, 1, 967142, 1, 7168, 2048, 2048, 16384)
    c8 = bench.make_config(a, 8, 967142, 1, 7168, 2048, 256, 16384)
    assert c1["workload"] == c8["workload"] and "slab rows [7168,9216)" in c1["workload"]
    assert c1["job_rows"] == [7168, 9216] and (c1["rows_per_rank"], c8["rows_per_rank"]) == (2048, 256)
    assert c1["points_after_dedupe"] == 967142 and c1["cols"] == 16384 and "flushed" in c1["l2"]
    assert json.loads(json.dumps(c8)) == c8
