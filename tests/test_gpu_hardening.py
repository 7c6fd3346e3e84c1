"""GPU (-m gpu): argument validation of the Python engine, device-side failure counts, and the host
pipeline's behaviour when a chunk fails in the middle."""
import ctypes

import numpy as np
import pytest

import poseidon252_b200 as pb
from poseidon252_b200.scalar import random_scalars

pytestmark = pytest.mark.gpu


def test_permute_batch_numpy_out(engine, coracle):
    rng = np.random.default_rng(3)
    states = random_scalars(rng, (50, 5))
    keep = states.copy()
    out = np.zeros_like(states)
    res = engine.permute_batch(states, out=out)
    assert res is out and np.array_equal(states, keep)            # the input is not permuted in place
    assert np.array_equal(out, coracle.permute(keep))
    with pytest.raises(pb.EngineError):
        engine.permute_batch(states, out=np.zeros((49, 5, 4), dtype=np.uint64))
    with pytest.raises(pb.EngineError):
        engine.permute_batch(states, out=np.zeros((50, 5, 4), dtype=np.int64))


def test_out_and_aux_buffer_validation(engine):
    import torch
    rng = np.random.default_rng(4)
    msg, sec, non = random_scalars(rng, (10, 2)), random_scalars(rng, (10, 2)), random_scalars(rng, 10)
    with pytest.raises(pb.EngineError):
        engine.encrypt_batch(msg, sec[:9], non)                  # short secrets
    with pytest.raises(pb.EngineError):
        engine.encrypt_batch(msg, sec, non[:5])                  # short nonces
    with pytest.raises(pb.EngineError):
        engine.encrypt_batch(msg, sec, non, out=np.zeros((10, 2, 4), dtype=np.uint64))     # undersized out
    cip = engine.encrypt_batch(msg, sec, non)
    with pytest.raises(pb.EngineError):
        engine.decrypt_batch(cip, sec[:3], non)
    with pytest.raises(pb.EngineError):
        engine.hash_batch(pb.Domain.Other, msg, out=np.zeros((10, 1, 4), dtype=np.float64))
    with pytest.raises(pb.EngineError):
        engine.hash_batch(pb.Domain.Other, msg, out=np.zeros((20, 1, 4), dtype=np.uint64)[::2])   # non-contiguous
    d = torch.from_numpy(msg.view(np.int64)).cuda()
    with pytest.raises(pb.EngineError):
        engine.hash_batch(pb.Domain.Other, d.to(torch.float64))  # 8-byte but not an integer tensor
    with pytest.raises(pb.EngineError):
        engine.hash_batch(pb.Domain.Other, d, out=torch.zeros((10, 1, 4), dtype=torch.int64))      # out on the CPU
    with pytest.raises(pb.EngineError):
        engine.hash_batch(pb.Domain.Other, d, out=torch.zeros((9, 1, 4), dtype=torch.int64, device="cuda"))


def test_decrypt_device_failure_count(engine):
    import torch
    rng = np.random.default_rng(6)
    n = 5000
    msg, sec, non = random_scalars(rng, (n, 3)), random_scalars(rng, (n, 2)), random_scalars(rng, n)
    cip = engine.encrypt_batch(msg, sec, non)
    bad = rng.choice(n, size=137, replace=False)
    cip[bad, 3, 0] ^= np.uint64(1)                               # tamper the authentication scalar
    d = [torch.from_numpy(a.view(np.int64)).cuda() for a in (cip, sec, non)]
    m, ok = engine.decrypt_batch(*d)
    assert engine.last_decrypt_failures() == 137 and int((ok == 0).sum()) == 137
    m2, ok2 = engine.decrypt_batch(*d, async_=True)
    engine.sync()
    assert engine.last_decrypt_failures() == 137
    mh, okh = engine.decrypt_batch(cip, sec, non)                # host buffers: counted from ok[]
    assert engine.last_decrypt_failures() == 137 and np.array_equal(okh, ok.cpu().numpy())
    good = np.setdiff1d(np.arange(n), bad)
    assert np.array_equal(mh[good], msg[good]) and not mh[bad].any()


def test_host_pipeline_failure_mid_batch_wipes_and_recovers():
    """A chunk that fails after earlier chunks were staged: the call reports the error, nothing stays in flight,
    the staging arenas that held secrets are zero, and the context works again afterwards."""
    eng = pb.Engine(0)
    lib, ctx = eng._lib, eng._ctx
    rng = np.random.default_rng(9)
    n = 200_000                                                   # several staged chunks
    msg, sec, non = random_scalars(rng, (n, 2)), random_scalars(rng, (n, 2)), random_scalars(rng, n)
    want = eng.encrypt_batch(msg, sec, non)
    nz = ctypes.c_size_t(1)
    assert lib.p252_debug_staging_nonzero(ctx, ctypes.byref(nz)) == 0 and nz.value == 0   # wiped after success too
    assert lib.p252_debug_fail_chunk(ctx, 2) == 0
    with pytest.raises(pb.EngineError) as ei:
        eng.encrypt_batch(msg, sec, non)
    assert "injected" in str(ei.value)
    assert lib.p252_debug_staging_nonzero(ctx, ctypes.byref(nz)) == 0 and nz.value == 0   # secrets wiped on the error path
    assert np.array_equal(eng.encrypt_batch(msg, sec, non), want)                           # context reusable
    # a digest call (no wipe) failing at its first chunk leaves the context usable as well
    assert lib.p252_debug_fail_chunk(ctx, 0) == 0
    with pytest.raises(pb.EngineError):
        eng.hash_batch(pb.Domain.Other, msg)
    assert eng.hash_batch(pb.Domain.Other, msg).shape == (n, 1, 4)
    eng.close()


def test_context_serialises_concurrent_callers(engine, coracle):
    """Two host threads on ONE context: calls block on the context mutex instead of racing."""
    import threading
    rng = np.random.default_rng(12)
    states = [random_scalars(rng, (3000, 5)) for _ in range(4)]
    out = [None] * 4

    def work(i):
        for _ in range(3):
            out[i] = engine.permute_batch(states[i])
    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    before = engine.launch_count
    [t.start() for t in th]
    [t.join() for t in th]
    assert engine.launch_count - before == 12
    for i in range(4):
        assert np.array_equal(out[i], coracle.permute(states[i]))


def test_kernel_info(engine):
    info = engine.kernel_info()
    assert info["montmul_per_permutation"] == 365 and info["dfma_per_permutation"] == 68 * 200
    assert 36_000 < info["wide_mul_per_permutation"] < 45_000
