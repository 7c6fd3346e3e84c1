"""CPU: the generated PTX carry chains, emulated instruction by instruction, equal the integer
definitions of the scaled-lazy model; the model equals the oracle; the generated headers are
up to date with their generators."""
import os
import random

import gen_field_ptx as g
import hades_model as hm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PINV = pow(hm.P, -1, hm.TWO256)


def mm(x, y):
    t = x * y
    m = (-t * PINV) % hm.TWO256
    return (t + m * hm.P) >> 256


def test_montmul_rows():
    rnd = random.Random(1)
    P, T = hm.P, hm.TWO256
    xs = [0, 1, P - 1, P, T - P, T - P - 1, (1 << 255) - 1]
    ys = [0, 1, P - 1, T - 1, 1 << 255, hm.M32, T - (1 << 32)]
    for x in xs:
        for y in ys:
            if x + P <= T and mm(x, y) < T:
                assert g.emu_montmul(x, y) == mm(x, y)
    for _ in range(600):
        x = rnd.randrange(T - P + 1)
        y = rnd.randrange(int(2.2 * P))
        if rnd.random() < 0.25:
            x = T - P - rnd.randrange(1 << rnd.randrange(1, 200))
        if rnd.random() < 0.25:
            y = int(2.2 * P) - rnd.randrange(1 << rnd.randrange(1, 200))
        if mm(x, y) >= T:
            continue
        assert g.emu_montmul(x, y) == mm(x, y)


def test_montsqr_and_wide_reduction():
    rnd = random.Random(4)
    P, T = hm.P, hm.TWO256
    for a in [0, 1, P - 1, P, (1 << 255) - 1, 1 << 255, T - P, hm.M32, T - 1]:
        if mm(a, a) < T:
            assert g.emu_montsqr(a) == mm(a, a)
    for _ in range(400):
        a = rnd.randrange(int(1.6 * P))
        if rnd.random() < 0.3:
            a = int(1.6 * P) - rnd.randrange(1 << rnd.randrange(1, 250))
        if rnd.random() < 0.1:
            a = rnd.randrange(1 << rnd.randrange(1, 256))
        assert g.emu_montsqr(a) == mm(a, a)
        x, y = rnd.randrange(int(2.2 * P)), rnd.randrange(int(1.2 * P))
        if mm(x, y) < T:
            assert g.emu_redc_wide(x * y) == mm(x, y)


def test_mix_tail():
    """FP64 column sums (exact, < 2^52) folded with their exponent words + wrap-around constant + one
    Montgomery row == redc1(C z + A)."""
    rnd = random.Random(2)
    for _ in range(500):
        z = [rnd.randrange(int(1.9 * hm.P)) for _ in range(5)]
        if rnd.random() < 0.3:
            z = [int(1.9 * hm.P) - rnd.randrange(1 << 40) for _ in range(5)]
        if rnd.random() < 0.1:
            z = [hm.TWO256 - 1 - rnd.randrange(1 << 20) for _ in range(5)]      # all-ones limbs: largest columns
        i = rnd.randrange(5)
        arc = rnd.choice([None, rnd.randrange(hm.P), hm.P - 1])
        zl = [hm.limbs32(v) for v in z]
        cols = [sum(hm.CMAT[i][j] * zl[j][k] for j in range(5)) for k in range(8)]
        for c in cols:                                   # exact in an IEEE double next to the 2^52 bias
            assert c < (1 << 52) and int(float((1 << 52) + c)) == (1 << 52) + c
        t = sum(hm.CMAT[i][j] * z[j] for j in range(5)) + (arc or 0)
        assert g.emu_mix_lane(cols, arc) == hm.redc1(t)


def test_conditional_subtractions_and_addsub():
    rnd = random.Random(3)
    P, T = hm.P, hm.TWO256
    for _ in range(500):
        a = rnd.choice([(1 << 255) - 1, 1 << 255, P, P - 1, 0, T - 1, rnd.randrange(T)])
        assert g.emu_unary("fr_condsub255", a) == (a - P if a >> 255 else a)
        a = rnd.choice([P, P - 1, 0, 2 * P - 1, P + 1, rnd.randrange(2 * P)])
        assert g.emu_unary("fr_condsub", a) == a % P
        a, b = rnd.choice([(0, 0), (0, P - 1), (P - 1, 0), (P - 1, P - 1), (5, 5), (0, 1),
                           (rnd.randrange(P), rnd.randrange(P))])
        assert g.emu_binary("fr_sub_mod", a, b) == (a - b) % P
        assert g.emu_binary("fr_add_lazy", a, b) == a + b


def test_scaled_lazy_model_equals_oracle(oracle):
    rnd = random.Random(7)
    cases = [[0] * 5, [1] * 5, [hm.P - 1] * 5, [17] * 5, list(range(5))]
    cases += [[rnd.randrange(hm.P) for _ in range(5)] for _ in range(25)]
    for c in cases:
        got = hm.permute_model([x * hm.R % hm.P for x in c])
        assert got == [x * hm.R % hm.P for x in oracle.perm(c)]
    # operand bounds the CUDA code relies on (DESIGN.md "Operand bounds")
    assert hm.Bounds.seen["sqr1"] < 1.4534 and hm.Bounds.seen["sqr2"] < 1.9565
    assert hm.Bounds.seen["x5"] < 1.8862 and hm.Bounds.seen["gmul"] < 1.855


def test_model_constants_equal_oracle(oracle):
    assert hm.ARC == oracle._ARC_FLAT and hm.MDS == oracle.MDS_MATRIX


def test_generated_sources_are_current():
    import gen_tables  # noqa: F401
    cur = open(os.path.join(ROOT, "poseidon252_b200", "csrc", "fr_ptx.cuh")).read()
    assert cur == g.emit_header(), "run python tools/gen_field_ptx.py"
