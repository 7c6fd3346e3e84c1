"""Engine: one CUDA device + stream behind the C ABI (include/poseidon252_b200.h).

Buffers are either numpy uint64 arrays (HOST: the library stages H2D/D2H in overlapped chunks) or
torch CUDA tensors of dtype int64/uint64 (DEVICE: zero-copy, enqueued on the engine's stream).
There is no CPU fallback: constructing an Engine without a B200-class GPU raises EngineError."""
import ctypes

import numpy as np

from . import _native
from .errors import EngineError, raise_for_status

_DEFAULT = {}


def _is_torch(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


class Engine:
    def __init__(self, device=0, stream=None):
        """stream: None -> the engine creates its own non-blocking stream; an int -> an existing
        cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream; 0 means the legacy
        default stream)."""
        self._lib = _native.lib()
        self._ctx = ctypes.c_void_p()
        self.device = int(device)
        if stream is None:
            rc = self._lib.p252_create(self.device, ctypes.byref(self._ctx))
        else:
            handle = int(stream) or 1          # 0 -> cudaStreamLegacy
            rc = self._lib.p252_create_on_stream(self.device, ctypes.c_void_p(handle), ctypes.byref(self._ctx))
        if rc != 0:
            self._ctx = ctypes.c_void_p()
            raise EngineError(rc, self._lib.p252_strerror(rc).decode())
        self._dist = False
        self._stream_handle = None if stream is None else int(stream)

    def _fence_torch(self):
        """Device tensors are produced on torch's current stream; unless the engine was bound to that
        very stream, wait for it before enqueueing on ours (cross-stream ordering)."""
        import torch
        cur = torch.cuda.current_stream(self.device)
        if self._stream_handle is None or int(cur.cuda_stream) != self._stream_handle:
            cur.synchronize()

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.p252_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def sync(self):
        self._check(self._lib.p252_sync(self._ctx))

    @property
    def launch_count(self):
        return int(self._lib.p252_launch_count(self._ctx))

    def _check(self, rc):
        raise_for_status(rc, self._lib, self._ctx)

    # -- buffer plumbing --------------------------------------------------------------------------
    def _in(self, x, shape_tail):
        """-> (pointer, leading-shape, flags, keepalive)"""
        if _is_torch(x):
            if not x.is_cuda or x.device.index != self.device:
                raise EngineError(-1, "tensor is not on cuda:%d" % self.device)
            if x.element_size() != 8 or not x.is_contiguous():
                raise EngineError(-1, "device buffers must be contiguous 64-bit integer tensors")
            if tuple(x.shape[-len(shape_tail):]) != tuple(shape_tail):
                raise EngineError(-1, "expected trailing shape %s, got %s" % (shape_tail, tuple(x.shape)))
            self._fence_torch()
            return x.data_ptr(), tuple(x.shape[:-len(shape_tail)]), _native.MEM_DEVICE, x
        a = np.ascontiguousarray(x, dtype=np.uint64)
        if tuple(a.shape[-len(shape_tail):]) != tuple(shape_tail):
            raise EngineError(-1, "expected trailing shape %s, got %s" % (shape_tail, a.shape))
        return a.ctypes.data, tuple(a.shape[:-len(shape_tail)]), _native.MEM_HOST, a

    def _out_like(self, ref, shape, dtype=None):
        if _is_torch(ref):
            import torch
            return torch.empty(shape, dtype=ref.dtype if dtype is None else dtype, device=ref.device)
        return np.empty(shape, dtype=np.uint64 if dtype is None else dtype)

    @staticmethod
    def _ptr(x):
        return x.data_ptr() if _is_torch(x) else x.ctypes.data

    # -- hades::permute_batch ---------------------------------------------------------------------
    def permute_batch(self, states, dense=False, out=None, async_=False):
        """n x Safe::permute (src/hades/permutation/scalar.rs:25-27).  states: (n, 5, 4)."""
        ptr, lead, flags, keep = self._in(states, (5, 4))
        n = int(np.prod(lead)) if lead else 1
        if _is_torch(keep):
            res = keep.clone() if out is None else out
            if out is not None and out is not keep:
                out.copy_(keep)
            self._fence_torch()
        else:
            res = keep.copy() if (out is None and keep is states) else keep
        fn = self._lib.p252_permute_batch_dense if dense else self._lib.p252_permute_batch
        self._check(fn(self._ctx, self._ptr(res), n, flags | (_native.ASYNC if async_ and flags else 0)))
        return res

    def permute_batch_inplace(self, states, async_=False):
        ptr, lead, flags, keep = self._in(states, (5, 4))
        n = int(np.prod(lead)) if lead else 1
        if not _is_torch(keep) and keep is not states:
            raise EngineError(-1, "in-place permute needs a contiguous uint64 array")
        self._check(self._lib.p252_permute_batch(self._ctx, ptr, n,
                                                 flags | (_native.ASYNC if async_ and flags else 0)))
        return states

    # -- sponges ------------------------------------------------------------------------------------
    def digest_batch_with_tag(self, tag, inputs, out_len=1, out=None, async_=False):
        """start(tag) -> absorb(in_len) -> squeeze(out_len) for every item.  inputs: (n, in_len, 4)."""
        tag = np.ascontiguousarray(tag, dtype=np.uint64).reshape(4)
        if inputs.ndim != 3:
            raise EngineError(-1, "inputs must have shape (n, in_len, 4)")
        in_len = int(inputs.shape[1])
        ptr, lead, flags, keep = self._in(inputs, (in_len, 4))
        n = lead[0]
        res = self._out_like(keep, (n, int(out_len), 4)) if out is None else out
        self._check(self._lib.p252_digest_batch(self._ctx, tag.ctypes.data, ptr, n, in_len, self._ptr(res),
                                                int(out_len), flags | (_native.ASYNC if async_ and flags else 0)))
        return res

    def hash_batch(self, domain, inputs, out_len=1, out=None, async_=False):
        """n x Hash::digest(domain, inputs[i]) with output_len(out_len)."""
        if inputs.ndim != 3:
            raise EngineError(-1, "inputs must have shape (n, in_len, 4)")
        in_len = int(inputs.shape[1])
        ptr, lead, flags, keep = self._in(inputs, (in_len, 4))
        n = lead[0]
        res = self._out_like(keep, (n, int(out_len), 4)) if out is None else out
        self._check(self._lib.p252_hash_batch(self._ctx, int(domain), ptr, n, in_len, self._ptr(res), int(out_len),
                                              flags | (_native.ASYNC if async_ and flags else 0)))
        return res

    def hash_batch_truncated(self, domain, inputs, out_len=1, out=None, async_=False):
        """n x Hash::digest_truncated: raw (canonical, 250-bit masked) limbs for JubJubScalar::from_raw."""
        if inputs.ndim != 3:
            raise EngineError(-1, "inputs must have shape (n, in_len, 4)")
        in_len = int(inputs.shape[1])
        ptr, lead, flags, keep = self._in(inputs, (in_len, 4))
        n = lead[0]
        res = self._out_like(keep, (n, int(out_len), 4)) if out is None else out
        self._check(self._lib.p252_hash_batch_truncated(self._ctx, int(domain), ptr, n, in_len, self._ptr(res),
                                                        int(out_len), flags | (_native.ASYNC if async_ and flags else 0)))
        return res

    def scalars_from_bytes(self, data, async_=False):
        """(n, 32) uint8 canonical little-endian (host) or (n, 4) 64-bit device tensor of the same bytes
        -> (scalars (n, 4), ok (n,) uint8); ok == 0 where the value is >= p."""
        if _is_torch(data):
            import torch
            ptr, lead, flags, keep = self._in(data, (4,))
            n = lead[0]
            out = torch.empty((n, 4), dtype=keep.dtype, device=keep.device)
            ok = torch.empty((n,), dtype=torch.uint8, device=keep.device)
        else:
            keep = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 32)
            ptr, n, flags = keep.ctypes.data, keep.shape[0], _native.MEM_HOST
            out = np.empty((n, 4), dtype=np.uint64)
            ok = np.empty((n,), dtype=np.uint8)
        self._check(self._lib.p252_scalars_from_bytes(self._ctx, ptr, n, self._ptr(out), self._ptr(ok),
                                                      flags | (_native.ASYNC if async_ and flags else 0)))
        return out, ok

    def scalars_to_bytes(self, scalars, async_=False):
        """(n, 4) scalars -> canonical little-endian bytes: (n, 32) uint8 (host) or (n, 4) device tensor."""
        ptr, lead, flags, keep = self._in(scalars, (4,))
        n = lead[0]
        if _is_torch(keep):
            import torch
            out = torch.empty((n, 4), dtype=keep.dtype, device=keep.device)
        else:
            out = np.empty((n, 32), dtype=np.uint8)
        self._check(self._lib.p252_scalars_to_bytes(self._ctx, ptr, n, self._ptr(out),
                                                    flags | (_native.ASYNC if async_ and flags else 0)))
        return out

    def encrypt_batch(self, messages, secrets_uv, nonces, out=None, async_=False):
        """messages (n, L, 4), secrets_uv (n, 2, 4), nonces (n, 4) -> ciphers (n, L+1, 4)."""
        L = int(messages.shape[1])
        mp, lead, flags, mk = self._in(messages, (L, 4))
        n = lead[0]
        sp, _, f2, sk = self._in(secrets_uv, (2, 4))
        np_, _, f3, nk = self._in(nonces, (4,))
        if not (flags == f2 == f3):
            raise EngineError(-1, "all buffers must live in the same memory space")
        res = self._out_like(mk, (n, L + 1, 4)) if out is None else out
        self._check(self._lib.p252_encrypt_batch(self._ctx, mp, n, L, sp, np_, self._ptr(res),
                                                 flags | (_native.ASYNC if async_ and flags else 0)))
        return res

    def decrypt_batch(self, ciphers, secrets_uv, nonces, async_=False):
        """ciphers (n, L+1, 4) -> (messages (n, L, 4), ok (n,) uint8)."""
        L = int(ciphers.shape[1]) - 1
        cp, lead, flags, ck = self._in(ciphers, (L + 1, 4))
        n = lead[0]
        sp, _, f2, sk = self._in(secrets_uv, (2, 4))
        np_, _, f3, nk = self._in(nonces, (4,))
        if not (flags == f2 == f3):
            raise EngineError(-1, "all buffers must live in the same memory space")
        if _is_torch(ck):
            import torch
            msg = torch.empty((n, max(L, 0), 4), dtype=ck.dtype, device=ck.device)
            ok = torch.empty((n,), dtype=torch.uint8, device=ck.device)
        else:
            msg = np.empty((n, max(L, 0), 4), dtype=np.uint64)
            ok = np.empty((n,), dtype=np.uint8)
        nfail = ctypes.c_size_t(0)
        self._check(self._lib.p252_decrypt_batch(self._ctx, cp, n, max(L, 0), sp, np_, self._ptr(msg), self._ptr(ok),
                                                 ctypes.byref(nfail), flags | (_native.ASYNC if async_ and flags else 0)))
        return msg, ok

    # -- arity-4 Merkle tree ----------------------------------------------------------------------
    def merkle4_level(self, children, out=None, async_=False):
        """children (4*m, 4) -> parents (m, 4)."""
        cp, lead, flags, ck = self._in(children, (4,))
        m = lead[0] // 4
        if lead[0] % 4:
            from .errors import IOPatternViolation
            raise IOPatternViolation()
        res = self._out_like(ck, (m, 4)) if out is None else out
        self._check(self._lib.p252_merkle4_level(self._ctx, cp, m, self._ptr(res),
                                                 flags | (_native.ASYNC if async_ and flags else 0)))
        return res

    def tree_nodes(self, n_leaves):
        ni, nl = ctypes.c_size_t(0), ctypes.c_int(0)
        self._check(self._lib.p252_merkle4_tree_nodes(int(n_leaves), ctypes.byref(ni), ctypes.byref(nl)))
        return int(ni.value), int(nl.value)

    def merkle4_build(self, leaves, out=None, async_=False):
        """leaves (4^k, 4) -> all internal nodes bottom-up ((4^k-1)/3, 4); root = last row."""
        lp, lead, flags, lk = self._in(leaves, (4,))
        n_internal, _ = self.tree_nodes(lead[0])
        res = self._out_like(lk, (n_internal, 4)) if out is None else out
        self._check(self._lib.p252_merkle4_build(self._ctx, lp, lead[0], self._ptr(res),
                                                 flags | (_native.ASYNC if async_ and flags else 0)))
        return res

    def merkle_build(self, leaves, arity=4, out=None, async_=False):
        """leaves (arity^k, 4) -> all internal nodes bottom-up, root last; arity 2 (Domain::Merkle2) or 4."""
        lp, lead, flags, lk = self._in(leaves, (4,))
        ni = ctypes.c_size_t(0)
        self._check(self._lib.p252_merkle_tree_nodes(int(arity), lead[0], ctypes.byref(ni), None))
        res = self._out_like(lk, (int(ni.value), 4)) if out is None else out
        self._check(self._lib.p252_merkle_build(self._ctx, int(arity), lp, lead[0], self._ptr(res),
                                                flags | (_native.ASYNC if async_ and flags else 0)))
        return res

    # -- multi-GPU (one process per GPU) ----------------------------------------------------------
    def dist_unique_id(self):
        buf = (ctypes.c_uint8 * _native.NCCL_UNIQUE_ID_BYTES)()
        self._check(self._lib.p252_dist_unique_id(buf))
        return bytes(buf)

    def dist_init(self, unique_id, rank, nranks):
        buf = (ctypes.c_uint8 * _native.NCCL_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        self._check(self._lib.p252_dist_init(self._ctx, buf, int(rank), int(nranks)))
        self._dist = True

    def dist_finalize(self):
        if self._dist:
            self._check(self._lib.p252_dist_finalize(self._ctx))
            self._dist = False

    def merkle4_build_dist(self, leaves_shard, n_leaves_total, out=None, async_=False):
        """This rank's contiguous shard of the leaves (device tensor) -> complete internal levels
        on every rank (one NCCL all-gather per level)."""
        lp, lead, flags, lk = self._in(leaves_shard, (4,))
        if flags != _native.MEM_DEVICE:
            raise EngineError(-1, "merkle4_build_dist takes device tensors")
        n_internal, _ = self.tree_nodes(n_leaves_total)
        res = self._out_like(lk, (n_internal, 4)) if out is None else out
        self._check(self._lib.p252_merkle4_build_dist(self._ctx, lp, int(n_leaves_total), self._ptr(res),
                                                      flags | (_native.ASYNC if async_ else 0)))
        return res


def default_engine(device=0):
    """Process-wide engine per device (created on first use)."""
    if device not in _DEFAULT:
        _DEFAULT[device] = Engine(device)
    return _DEFAULT[device]
