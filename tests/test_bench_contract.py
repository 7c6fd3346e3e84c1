"""CPU: the parts of bench.py the driver depends on that need no GPU -- the reference arm prints one JSON line with the
contract's keys, the same `config` object as the GPU arm, and times only the hashing call."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--log2-batch", "12"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                       # exactly one JSON line on stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "hades_permutations_per_sec" and d["unit"] == "perm/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    # value = digests hashed / time spent inside the hashing call
    assert abs(d["value"] - (1 << 12) / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(12, 1)           # identical to the GPU arm's config object


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--gpus", "2", "--log2-batch", "10"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_gpu_arm_line_small():
    """the GPU arm end to end on a reduced batch: one JSON line, contract keys, roofline / imad / e2e / tree blocks,
    tree parity ok"""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--log2-batch", "14",
                          "--log4-leaves", "7", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "gpu_launches", "roofline", "e2e", "tree"):
        assert k in d, k
    assert d["gpu_launches"] == 3 and d["n_gpus"] == 1 and d["value"] > 1e6
    assert d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] < 1
    assert d["roofline"]["imad"]["bound"] == "imad" and 0 < d["roofline"]["imad"]["frac"] < 1
    assert d["e2e"]["h2d_bytes_per_step"] == (1 << 14) * 128 and d["e2e"]["d2h_bytes_per_step"] == (1 << 14) * 32
    assert d["tree"]["parity"] == "ok" and d["tree"]["leaves_log4"] == 7 and len(d["tree"]["per_level_rank0"]) == 7
