"""Worker for tests/test_gpu_dist.py: one process per GPU (torchrun), NCCL inside the C library.
Builds an arity-4 tree with p252_merkle4_build_dist (leaf shards, one all-gather per level) and checks every
rank's complete node array against the single-GPU p252_merkle4_build of the same leaves."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import poseidon252_b200 as pb  # noqa: E402
from poseidon252_b200.scalar import random_limbs_fast  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n_leaves = 4 ** k
    leaves = random_limbs_fast(np.random.default_rng(2024), n_leaves)          # identical on every rank
    eng = pb.Engine(local)
    box = [eng.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    eng.dist_init(box[0], rank, world)
    shard = n_leaves // world
    d_shard = torch.from_numpy(leaves[rank * shard:(rank + 1) * shard].view(np.int64)).cuda()
    torch.cuda.synchronize()
    nodes = eng.merkle4_build_dist(d_shard, n_leaves)
    got = nodes.cpu().numpy().view(np.uint64)
    want = eng.merkle4_build(leaves)                                            # single GPU, host buffers
    ok = bool(np.array_equal(got, want))
    # run it a second time (communicator reuse) and asynchronously
    nodes2 = eng.merkle4_build_dist(d_shard, n_leaves, async_=True)
    eng.sync()
    ok = ok and bool(np.array_equal(nodes2.cpu().numpy().view(np.uint64), want))
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    eng.dist_finalize()
    eng.close()
    if rank == 0:
        print("GPU_DIST_OK" if all(flags) else "GPU_DIST_MISMATCH", flags, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
