"""CPU, world_size 2 over gloo: the N > 1 host logic (level partition + per-level all-gather of the
multi-GPU tree build; rank handling of bench.py's reference arm)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(args, nproc=2, timeout=240):
    port = 29500 + (os.getpid() % 500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_shard_plan_properties():
    from poseidon252_b200 import merkle
    for k in (2, 3, 5, 14):
        n = 4 ** k
        for world in (1, 2, 4, 8):
            if (n // world) % 4:
                continue
            plans = [merkle.shard_plan(n, world, r) for r in range(world)]
            assert len(plans[0]) == k
            for lv in range(k):
                size = plans[0][lv]["level_size"]
                assert size == n // 4 ** (lv + 1)
                if plans[0][lv]["sharded"]:
                    # slices tile the level exactly, in rank order, and every slice's children are the
                    # owner's slice of the level below
                    assert [p[lv]["my_offset"] for p in plans] == [r * size // world for r in range(world)]
                    assert sum(p[lv]["my_count"] for p in plans) == size
                    if lv > 0:
                        assert all(p[lv - 1]["sharded"] and p[lv - 1]["my_offset"] == 4 * p[lv]["my_offset"] for p in plans)
                else:
                    assert all(p[lv]["my_count"] == size and p[lv]["my_offset"] == 0 for p in plans)
            assert plans[0][0]["sharded"] == 1
    # config 4 of BASELINE.json: 2^28 leaves over 8 GPUs
    p = merkle.shard_plan(4 ** 14, 8, 3)
    assert p[0]["my_count"] == 4 ** 13 // 8 and sum(x["level_size"] for x in p) == 89478485
    assert [x["sharded"] for x in p] == [1] * 12 + [0, 0]


@pytest.mark.parametrize("k", [3, 5])
def test_tree_build_logic_world2_gloo(k):
    res = _torchrun([os.path.join("tests", "dist_worker.py"), str(k)])
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "DIST_OK" in res.stdout


def test_bench_reference_arm_world2():
    """Under torchrun rank 0 alone runs and prints the reference arm; the other rank exits 0."""
    res = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    import json
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
