"""Worker for tests/test_dist_cpu.py (world_size 2, gloo, CPU).  Executes the multi-GPU tree build's
host logic -- the level partition of p252_merkle4_shard_plan and one all-gather per sharded level -- with
the C oracle standing in for the CUDA node hash, and compares with a single-process oracle tree."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import c_oracle  # noqa: E402
import hades_oracle as o  # noqa: E402
from poseidon252_b200 import merkle  # noqa: E402
from poseidon252_b200.scalar import random_limbs_fast, to_mont  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n_leaves = 4 ** k
    leaves = random_limbs_fast(np.random.default_rng(99), n_leaves)          # same on every rank
    tag = to_mont(o.hash_to_scalar(o.tag_input([o.Absorb(4), o.Squeeze(1)], o.Domain.Merkle4)))
    plan = merkle.shard_plan(n_leaves, world, rank)
    n_internal = (n_leaves - 1) // 3
    nodes = np.zeros((n_internal, 4), dtype=np.uint64)
    shard = n_leaves // world
    below_mine = leaves[rank * shard:(rank + 1) * shard]
    below_full = None
    gathers = 0
    for lv in plan:
        off, m = lv["level_offset"], lv["level_size"]
        if lv["sharded"]:
            mine = c_oracle.digest(tag, below_mine.reshape(-1, 4, 4), 4, 1).reshape(-1, 4)
            assert mine.shape[0] == lv["my_count"]
            parts = [torch.zeros((lv["my_count"], 4), dtype=torch.int64) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(mine.view(np.int64)))      # the only collective
            gathers += 1
            nodes[off:off + m] = torch.cat(parts).numpy().view(np.uint64)
            below_mine = mine
        else:
            nodes[off:off + m] = c_oracle.digest(tag, below_full.reshape(-1, 4, 4), 4, 1).reshape(-1, 4)
        below_full = nodes[off:off + m]
    # single-process reference tree
    want = np.zeros_like(nodes)
    level, off = leaves, 0
    while level.shape[0] > 1:
        nxt = c_oracle.digest(tag, level.reshape(-1, 4, 4), 4, 1).reshape(-1, 4)
        want[off:off + nxt.shape[0]] = nxt
        off += nxt.shape[0]
        level = nxt
    ok = np.array_equal(nodes, want)
    flags = [None] * world
    dist.all_gather_object(flags, (ok, gathers))
    if rank == 0:
        print("DIST_OK" if all(f[0] for f in flags) else "DIST_MISMATCH", flags, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
