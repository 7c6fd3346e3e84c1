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


def level_offsets(n_leaves):
    """[(offset, size)] of each internal level inside the node array, bottom-up."""
    out, off, m = [], 0, n_leaves // 4
    while m >= 1:
        out.append((off, m))
        off += m
        if m == 1:
            break
        m //= 4
    return out
