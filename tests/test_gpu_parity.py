"""GPU (-m gpu): the CUDA path, called through the C ABI, is bit-exact against the oracle, the
reference's KATs and the committed golden vectors; edge cases; full-size properties."""
import numpy as np
import pytest

import poseidon252_b200 as pb
from conftest import edge_and_random_scalars, hx, mont, unmont
from poseidon252_b200 import hash as H
from poseidon252_b200.scalar import random_limbs_fast, random_scalars

pytestmark = pytest.mark.gpu


def _tag(oracle, pattern, dom):
    return mont(oracle.hash_to_scalar(oracle.tag_input(pattern, dom)))


# ---- permutation -------------------------------------------------------------------------------
def test_permute_golden(engine, golden):
    ins = mont([[hx(s) for s in e["in"]] for e in golden["perm"]])
    out = engine.permute_batch(ins)
    for row, e in zip(out, golden["perm"]):
        assert ["0x%064x" % v for v in unmont(row)] == e["out"]
    dense = engine.permute_batch(ins, dense=True)
    assert np.array_equal(out, dense)


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 127, 128, 129, 1000, 4099])
def test_permute_vs_oracle_ragged(engine, coracle, n):
    rng = np.random.default_rng(n)
    states = edge_and_random_scalars(rng, n * 5).reshape(n, 5, 4)
    want = coracle.permute(states)
    got = engine.permute_batch(states)
    assert np.array_equal(got, want)
    assert np.array_equal(engine.permute_batch(states, dense=True), want)


def test_permute_empty_and_inplace(engine):
    z = np.zeros((0, 5, 4), dtype=np.uint64)
    assert engine.permute_batch(z).shape == (0, 5, 4)
    s = random_scalars(np.random.default_rng(1), (10, 5))
    t = s.copy()
    engine.permute_batch_inplace(t)
    assert np.array_equal(t, engine.permute_batch(s))
    # hades_det, src/hades/permutation/scalar.rs:86-98
    x = pb.hades.permute(mont([17] * 5), engine)
    y = pb.hades.permute(mont([17] * 5), engine)
    z = pb.hades.permute(mont([19] * 5), engine)
    assert np.array_equal(x, y) and not np.array_equal(x, z)


def test_permute_device_tensors(engine, coracle):
    import torch
    rng = np.random.default_rng(11)
    states = random_scalars(rng, (777, 5))
    d = torch.from_numpy(states.view(np.int64)).cuda()
    out = engine.permute_batch(d)
    assert out.is_cuda and np.array_equal(out.cpu().numpy().view(np.uint64), coracle.permute(states))


# ---- reference KATs through the GPU sponge -----------------------------------------------------------
def test_reference_kats_on_gpu(engine, oracle):
    """src/hades.rs:128-162: tag 0, Absorb(n) + Absorb(1) of one, Squeeze(1).  Two absorbs aggregate
    into one run of n+1 elements in the sponge schedule."""
    kin = oracle.kat_inputs()
    for n, want in oracle.KAT_EXPECTED.items():
        data = mont(kin[:n] + [1]).reshape(1, n + 1, 4)
        out = engine.digest_batch_with_tag(mont(0), data, 1)
        assert oracle.debug_hex(unmont(out)[0, 0]) == want


# ---- Hash --------------------------------------------------------------------------------------
def test_digest_golden(engine, golden):
    for e in golden["digest"]:
        dom = {0xF: pb.Domain.Merkle4, 0x3: pb.Domain.Merkle2, 0: pb.Domain.Other}[e["domain"]]
        h = pb.Hash(dom, engine)
        h.output_len(e["out_len"])
        h.update(mont([hx(s) for s in e["in"]]))
        assert ["0x%064x" % v for v in unmont(h.finalize())] == e["out"]


def test_readme_example(engine, oracle):
    """README.md:22-51 (config 1 of BASELINE.json): 42 scalars, Domain::Other."""
    rng = np.random.default_rng(0xBEEF)
    x = random_scalars(rng, 42)
    one = pb.Hash.digest(pb.Domain.Other, x, engine)
    h = pb.Hash(pb.Domain.Other, engine)
    h.update(x[:3])
    h.update(x[3:])
    assert np.array_equal(h.finalize(), one)
    m4 = pb.Hash.digest(pb.Domain.Merkle4, x[:4], engine)
    assert not np.array_equal(m4, pb.Hash.digest(pb.Domain.Other, x[:4], engine))
    want = oracle.Hash.digest(oracle.Domain.Other, [int(v) for v in unmont(x)])
    assert [int(v) for v in unmont(one)] == want


def test_hash_errors(engine):
    x = mont([1, 2, 3, 4, 5])
    with pytest.raises(pb.IOPatternViolation):
        pb.Hash.digest(pb.Domain.Merkle4, x, engine)
    with pytest.raises(pb.IOPatternViolation):
        pb.Hash.digest(pb.Domain.Merkle2, x[:3], engine)
    with pytest.raises(pb.InvalidIOPattern):
        pb.Hash.digest(pb.Domain.Other, x[:0], engine)
    with pytest.raises(pb.IOPatternViolation):
        pb.Hash.digest_batch(pb.Domain.Merkle4, np.zeros((3, 5, 4), dtype=np.uint64), engine=engine)


@pytest.mark.parametrize("in_len,out_len", [(1, 1), (2, 1), (3, 3), (4, 1), (5, 2), (4, 7), (8, 4), (9, 5),
                                            (15, 1), (16, 8), (17, 1), (42, 1), (64, 9)])
def test_digest_batch_vs_oracle(engine, oracle, coracle, in_len, out_len):
    rng = np.random.default_rng(in_len * 100 + out_len)
    n = 257
    data = edge_and_random_scalars(rng, n * in_len).reshape(n, in_len, 4)
    got = pb.Hash.digest_batch(pb.Domain.Other, data, out_len, engine=engine)
    tag = _tag(oracle, [oracle.Absorb(in_len), oracle.Squeeze(out_len)], oracle.Domain.Other)
    assert np.array_equal(got, coracle.digest(tag, data, in_len, out_len))


def test_merkle_domains_batch(engine, oracle, coracle):
    rng = np.random.default_rng(3)
    for dom, od, k in ((pb.Domain.Merkle4, oracle.Domain.Merkle4, 4), (pb.Domain.Merkle2, oracle.Domain.Merkle2, 2)):
        data = edge_and_random_scalars(rng, 1000 * k).reshape(1000, k, 4)
        got = pb.Hash.digest_batch(dom, data, engine=engine)
        tag = _tag(oracle, [oracle.Absorb(k), oracle.Squeeze(1)], od)
        assert np.array_equal(got, coracle.digest(tag, data, k, 1))
        # output_len is ignored for Merkle domains (src/hash.rs:111-115)
        assert pb.Hash.digest_batch(dom, data[:5], 3, engine=engine).shape == (5, 1, 4)


def test_sweep_lengths(engine, oracle, coracle):
    """config 5 shape at reduced batch: Domain::Other, EVERY in_len 1..256 (ragged 40-item batches: partial warps
    and, for in_len % 4 != 0, partial tile loads)."""
    import os
    rng = np.random.default_rng(5)
    th = min(8, os.cpu_count() or 1)
    for in_len in range(1, 257):
        n = 40
        data = random_limbs_fast(rng, (n, in_len))
        got = pb.Hash.digest_batch(pb.Domain.Other, data, engine=engine)
        tag = _tag(oracle, [oracle.Absorb(in_len), oracle.Squeeze(1)], oracle.Domain.Other)
        assert np.array_equal(got, coracle.digest(tag, data, in_len, 1, threads=th)), in_len


@pytest.mark.parametrize("in_len", [1, 4, 5, 255, 256])
def test_sweep_lengths_4096_items(engine, oracle, coracle, in_len):
    """the sweep's corner lengths on a 2^12-item batch (several blocks per SM, full and partial last chunks)"""
    import os
    rng = np.random.default_rng(50 + in_len)
    data = random_limbs_fast(rng, (1 << 12, in_len))
    got = pb.Hash.digest_batch(pb.Domain.Other, data, engine=engine)
    tag = _tag(oracle, [oracle.Absorb(in_len), oracle.Squeeze(1)], oracle.Domain.Other)
    assert np.array_equal(got, coracle.digest(tag, data, in_len, 1, threads=min(8, os.cpu_count() or 1)))


# ---- encryption --------------------------------------------------------------------------------------
def test_encrypt_golden_and_errors(engine, golden):
    for e in golden["encrypt"]:
        msg, sec, non = mont([hx(s) for s in e["msg"]]), mont([hx(s) for s in e["secret"]]), mont(hx(e["nonce"]))
        c = pb.encrypt(msg, sec, non, engine)
        assert ["0x%064x" % v for v in unmont(c)] == e["cipher"]
        assert c.shape[0] == e["L"] + 1                               # src/encryption.rs:61
        assert np.array_equal(pb.decrypt(c, sec, non, engine), msg)
        # tests/encryption.rs:48-115
        wrong = sec.copy()
        wrong[1] = mont(12345)
        with pytest.raises(pb.DecryptionFailed):
            pb.decrypt(c, wrong, non, engine)
        with pytest.raises(pb.DecryptionFailed):
            pb.decrypt(c, sec, mont(99), engine)
        for idx in (e["L"], 0):
            bad = mont([(int(v) + (42 if k == idx else 0)) % pb.scalar.P for k, v in enumerate(unmont(c))])
            with pytest.raises(pb.DecryptionFailed):
                pb.decrypt(bad, sec, non, engine)
    # doc example src/encryption.rs:29-42
    msg = mont([10, 20, 30])
    c = pb.encrypt(msg, mont([5, 6]), mont(7), engine)
    assert np.array_equal(pb.decrypt(c, mont([5, 6]), mont(7), engine), msg)


@pytest.mark.parametrize("L", [1, 2, 3, 4, 5, 7, 8, 9, 21, 42])
def test_encrypt_batch_vs_oracle(engine, oracle, coracle, L):
    rng = np.random.default_rng(L)
    n = 300
    msg = edge_and_random_scalars(rng, n * L).reshape(n, L, 4)
    sec = random_scalars(rng, (n, 2))
    non = random_scalars(rng, n)
    tag = _tag(oracle, [oracle.Absorb(2), oracle.Absorb(1), oracle.Squeeze(L), oracle.Absorb(L), oracle.Squeeze(1)],
               oracle.Domain.Encryption)
    c = pb.encrypt_batch(msg, sec, non, engine=engine)
    assert np.array_equal(c, coracle.encrypt(tag, msg, L, sec, non))
    m, ok = pb.decrypt_batch(c, sec, non, engine=engine)
    assert ok.all() and np.array_equal(m, msg)
    # tamper every third item in a different place
    bad = c.copy()
    bad[0::3, 0, 0] ^= np.uint64(1)
    bad[1::3, L, 3] ^= np.uint64(1 << 40)
    m2, ok2 = pb.decrypt_batch(bad, sec, non, engine=engine)
    want_ok = np.ones(n, dtype=np.uint8)
    want_ok[0::3] = 0
    want_ok[1::3] = 0
    assert np.array_equal(ok2, want_ok)
    assert np.array_equal(m2[2::3], msg[2::3]) and not m2[0::3].any()        # failed items are zeroed
    _, ok_o = coracle.decrypt(tag, bad, L, sec, non)
    assert np.array_equal(ok_o, want_ok)


# ---- Merkle tree -----------------------------------------------------------------------------------------
def test_merkle4_tree_vs_oracle(engine, oracle, coracle):
    rng = np.random.default_rng(8)
    n_leaves = 4 ** 5
    leaves = random_limbs_fast(rng, n_leaves)
    leaves[5:9] = 0                                            # empty slots are zero (src/hash.rs:24-26)
    nodes = pb.merkle4_build(leaves, engine)
    tag = _tag(oracle, [oracle.Absorb(4), oracle.Squeeze(1)], oracle.Domain.Merkle4)
    level = leaves
    for off, m in pb.merkle.level_offsets(n_leaves):
        want = coracle.digest(tag, level.reshape(m, 4, 4), 4, 1).reshape(m, 4)
        assert np.array_equal(nodes[off:off + m], want)
        level = want
    assert nodes.shape[0] == (n_leaves - 1) // 3
    # one level through the public helper, and the device-tensor path
    assert np.array_equal(pb.merkle4_level(leaves, engine), nodes[:n_leaves // 4])
    import torch
    d = torch.from_numpy(leaves.view(np.int64)).cuda()
    assert np.array_equal(pb.merkle4_build(d, engine).cpu().numpy().view(np.uint64), nodes)
    with pytest.raises(pb.IOPatternViolation):
        pb.merkle4_build(leaves[:48], engine)


# ---- BASELINE.json full sizes: size-independent properties + multi-thread oracle --------------------------
def test_full_size_merkle4_batch(engine, oracle, coracle):
    """config 2: 2^20 Merkle4 digests.  (a) fast kernel == dense kernel (the reference's formulation
    executed on the device) on 2^17 raw states; (b) the whole 2^20 batch against the multi-threaded C
    oracle on a 2^16-item strided sample plus an XOR checksum over all items computed both ways on a
    contiguous 2^17 prefix; (c) permuting the batch order permutes the digests (no cross-item state)."""
    import os
    rng = np.random.default_rng(20)
    n = 1 << 20
    data = random_limbs_fast(rng, (n, 4))
    got = pb.Hash.digest_batch(pb.Domain.Merkle4, data, engine=engine)
    tag = _tag(oracle, [oracle.Absorb(4), oracle.Squeeze(1)], oracle.Domain.Merkle4)
    thr = max(1, min(32, os.cpu_count() or 1))
    idx = np.arange(0, n, 16)
    assert np.array_equal(got[idx], coracle.digest(tag, data[idx], 4, 1, threads=thr))
    pre = 1 << 17
    want = coracle.digest(tag, data[:pre], 4, 1, threads=thr)
    assert np.array_equal(np.bitwise_xor.reduce(got[:pre].reshape(-1, 4), axis=0),
                          np.bitwise_xor.reduce(want.reshape(-1, 4), axis=0))
    perm = rng.permutation(n)
    assert np.array_equal(pb.Hash.digest_batch(pb.Domain.Merkle4, data[perm], engine=engine), got[perm])
    states = random_limbs_fast(rng, (pre, 5))
    assert np.array_equal(engine.permute_batch(states), engine.permute_batch(states, dense=True))


def test_full_size_encrypt_roundtrip(engine):
    """config 3: 2^20 messages, L = 2 (benches/encrypt.rs:17): bit-exact round trip."""
    rng = np.random.default_rng(21)
    n = 1 << 20
    msg, sec, non = random_limbs_fast(rng, (n, 2)), random_limbs_fast(rng, (n, 2)), random_limbs_fast(rng, n)
    c = pb.encrypt_batch(msg, sec, non, engine=engine)
    m, ok = pb.decrypt_batch(c, sec, non, engine=engine)
    assert ok.all() and np.array_equal(m, msg)
    non2 = non.copy()
    non2[::2, 0] ^= np.uint64(1)
    _, ok2 = pb.decrypt_batch(c, sec, non2, engine=engine)
    assert not ok2[::2].any() and ok2[1::2].all()


# ---- "next" rows of SURVEY 8(f): truncated digests and the wire format ------------------------------------
def test_digest_truncated(engine, oracle):
    """Hash::digest_truncated / finalize_truncated, src/hash.rs:164-183,203-210 (shapes of tests/hash.rs:188-203)."""
    rng = np.random.default_rng(31)
    for in_len, out_len in ((3, 1), (5, 1), (15, 1), (4, 7)):
        data = random_scalars(rng, (50, in_len))
        got = pb.Hash.digest_truncated_batch(pb.Domain.Other, data, out_len, engine=engine)
        assert got.shape == (50, out_len, 4)
        for i in (0, 17, 49):
            h = oracle.Hash(oracle.Domain.Other)
            h.output_len(out_len)
            h.update([int(v) for v in unmont(data[i])])
            want = h.finalize_truncated()
            have = [sum(int(got[i, o, k]) << (64 * k) for k in range(4)) for o in range(out_len)]
            assert have == want and all(v < (1 << 250) for v in have)
    # single-item API and the Merkle4 domain
    x = random_scalars(rng, 4)
    one = pb.Hash.digest_truncated(pb.Domain.Merkle4, x, engine)
    want = oracle.Hash.digest_truncated(oracle.Domain.Merkle4, [int(v) for v in unmont(x)])
    assert [sum(int(one[0, k]) << (64 * k) for k in range(4))] == want
    h = pb.Hash(pb.Domain.Other, engine)
    h.update(x[:1])
    h.update(x[1:])
    want = oracle.Hash.digest_truncated(oracle.Domain.Other, [int(v) for v in unmont(x)])
    assert [sum(int(h.finalize_truncated()[0, k]) << (64 * k) for k in range(4))] == want


def test_wire_format_roundtrip(engine, oracle):
    """BlsScalar::from_bytes / to_bytes (src/hades.rs:94-105 parses its KAT inputs this way;
    src/hades/round_constants.rs:64-68 round-trips every constant)."""
    kin = oracle.kat_inputs()
    raw = np.frombuffer(b"".join(bytes.fromhex(s) for s in oracle.KAT_INPUTS_LE_HEX), dtype=np.uint8).reshape(-1, 32)
    sc, ok = engine.scalars_from_bytes(raw)
    assert ok.all() and [int(v) for v in unmont(sc)] == kin
    assert np.array_equal(engine.scalars_to_bytes(sc), raw)
    # all 340 round constants round-trip
    arc = mont(oracle._ARC_FLAT)
    b = engine.scalars_to_bytes(arc)
    assert b.tobytes() == oracle.arc_bin_bytes()
    back, ok = engine.scalars_from_bytes(b)
    assert ok.all() and np.array_equal(back, arc)
    # non-canonical encodings are rejected like from_bytes -> None
    bad = np.stack([np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
                    for v in (oracle.P, oracle.P + 1, (1 << 256) - 1, oracle.P - 1, 0)])
    sc, ok = engine.scalars_from_bytes(bad)
    assert list(ok) == [0, 0, 0, 1, 1] and not sc[:3].any()
    assert [int(v) for v in unmont(sc[3:])] == [oracle.P - 1, 0]
    # large ragged batch, device tensors
    import torch
    rng = np.random.default_rng(32)
    big = random_scalars(rng, 100003 // 100)          # ~1000 scalars
    d = torch.from_numpy(big.view(np.int64)).cuda()
    db = engine.scalars_to_bytes(d)
    dback, dok = engine.scalars_from_bytes(db)
    assert bool(dok.all()) and np.array_equal(dback.cpu().numpy().view(np.uint64), big)
    assert [int.from_bytes(r.tobytes(), "little") for r in db.cpu().numpy().view(np.uint8).reshape(-1, 32)[:5]] == \
        [int(v) for v in unmont(big[:5])]


def test_merkle2_tree_vs_oracle(engine, oracle, coracle):
    """Binary tree of Domain::Merkle2 nodes (src/hash.rs:27-31,49), host and device buffers."""
    rng = np.random.default_rng(9)
    n_leaves = 2 ** 9
    leaves = random_limbs_fast(rng, n_leaves)
    nodes = pb.merkle.merkle2_build(leaves, engine)
    assert nodes.shape[0] == n_leaves - 1
    tag = _tag(oracle, [oracle.Absorb(2), oracle.Squeeze(1)], oracle.Domain.Merkle2)
    level = leaves
    for off, m in pb.merkle.level_offsets(n_leaves, 2):
        want = coracle.digest(tag, level.reshape(m, 2, 4), 2, 1).reshape(m, 4)
        assert np.array_equal(nodes[off:off + m], want)
        level = want
    import torch
    d = torch.from_numpy(leaves.view(np.int64)).cuda()
    assert np.array_equal(engine.merkle_build(d, arity=2).cpu().numpy().view(np.uint64), nodes)
    with pytest.raises(pb.IOPatternViolation):
        engine.merkle_build(leaves[:48], arity=2)
    with pytest.raises(pb.EngineError):
        engine.merkle_build(leaves[:27], arity=3)


def test_randomized_shapes_vs_oracle(engine, oracle, coracle):
    """Randomised sweep over (n, in_len, out_len) and (n, L): ragged batch sizes, multi-permutation absorbs and
    squeezes, every result bit-exact against the C oracle."""
    rng = np.random.default_rng(20260924)
    for _ in range(24):
        n = int(rng.choice([1, 2, 5, 31, 32, 33, 63, 100, 129, 300, 1025]))
        in_len = int(rng.integers(1, 40))
        out_len = int(rng.integers(1, 11))
        data = random_limbs_fast(rng, (n, in_len))
        got = pb.Hash.digest_batch(pb.Domain.Other, data, out_len, engine=engine)
        tag = _tag(oracle, [oracle.Absorb(in_len), oracle.Squeeze(out_len)], oracle.Domain.Other)
        assert np.array_equal(got, coracle.digest(tag, data, in_len, out_len)), (n, in_len, out_len)
    for _ in range(12):
        n = int(rng.choice([1, 3, 32, 65, 127, 500]))
        L = int(rng.integers(1, 30))
        msg, sec, non = random_limbs_fast(rng, (n, L)), random_limbs_fast(rng, (n, 2)), random_limbs_fast(rng, n)
        tag = _tag(oracle, [oracle.Absorb(2), oracle.Absorb(1), oracle.Squeeze(L), oracle.Absorb(L), oracle.Squeeze(1)],
                   oracle.Domain.Encryption)
        c = pb.encrypt_batch(msg, sec, non, engine=engine)
        assert np.array_equal(c, coracle.encrypt(tag, msg, L, sec, non)), (n, L)
        m, ok = pb.decrypt_batch(c, sec, non, engine=engine)
        assert ok.all() and np.array_equal(m, msg)
    # worst-case style inputs: every limb all-ones below p, p-1, and values just below 2^255
    special = mont([oracle.P - 1, oracle.P - 2, (1 << 254) + 12345, (1 << 255) % oracle.P, 1, 0])
    states = np.stack([np.stack([special[(i + j) % 6] for j in range(5)]) for i in range(64)])
    assert np.array_equal(engine.permute_batch(states), coracle.permute(states))
