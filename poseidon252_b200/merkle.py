"""Arity-4 Merkle trees of Domain::Merkle4 digests (node = Hash::digest(Domain::Merkle4, 4 children),
/root/reference/src/hash.rs:22-26).  Tree logic itself left the reference crate in 0.29.0
(CHANGELOG.md:164-168); only the node hash is defined there."""
from .engine import default_engine


def merkle4_level(children, engine=None, out=None, async_=False):
    eng = engine or default_engine(children.device.index if hasattr(children, "is_cuda") else 0)
    return eng.merkle4_level(children, out=out, async_=async_)


def merkle4_build(leaves, engine=None, out=None, async_=False):
    """leaves (4^k, 4) -> internal nodes bottom-up, root last."""
    eng = engine or default_engine(leaves.device.index if hasattr(leaves, "is_cuda") else 0)
    return eng.merkle4_build(leaves, out=out, async_=async_)


def merkle2_build(leaves, engine=None, out=None, async_=False):
    """Binary tree of Domain::Merkle2 digests (src/hash.rs:27-31): leaves (2^k, 4) -> internal nodes, root last."""
    eng = engine or default_engine(leaves.device.index if hasattr(leaves, "is_cuda") else 0)
    return eng.merkle_build(leaves, arity=2, out=out, async_=async_)


def level_offsets(n_leaves, arity=4):
    """[(offset, size)] of each internal level inside the node array, bottom-up."""
    out, off, m = [], 0, n_leaves // arity
    while m >= 1:
        out.append((off, m))
        off += m
        if m == 1:
            break
        m //= arity
    return out


def shard_plan(n_leaves_total, nranks, rank):
    """The per-level partition of the multi-GPU build (p252_merkle4_shard_plan): list of dicts with
    level_offset, level_size, my_offset, my_count, sharded -- bottom-up."""
    import ctypes

    from . import _native
    from .errors import raise_for_status
    lib = _native.lib()
    n = ctypes.c_int(0)
    raise_for_status(lib.p252_merkle4_shard_plan(int(n_leaves_total), int(nranks), int(rank), None, 0, ctypes.byref(n)), lib)
    arr = (_native.LevelPlan * n.value)()
    raise_for_status(lib.p252_merkle4_shard_plan(int(n_leaves_total), int(nranks), int(rank), arr, n.value, ctypes.byref(n)), lib)
    return [{f: int(getattr(a, f)) for f in ("level_offset", "level_size", "my_offset", "my_count", "sharded")} for a in arr]
