// Device-side Hades permutation for BLS12-381 Fr on sm_100a -- one width-5 state per thread, all
// 40 state words in registers.
//
// Replaces, on the batch path, the reference's scalar loop
//   Hades::perm                      /root/reference/src/hades/permutation.rs:105-123
//   add_round_constants / quintic_s_box / mul_matrix
//                                    /root/reference/src/hades/permutation/scalar.rs:39-64
// with the bit-exact "scaled lazy" formulation derived in tools/hades_model.py:
//   * 365 unreduced Montgomery products per permutation (IMAD.WIDE carry chains, fr_ptx.cuh)
//     instead of the reference's 2000 (dense 25-multiply MDS every round);
//   * the MDS layer is 25 small-integer (<= 17 bit) multiply-adds per round with plain
//     mad.wide.u32 (no carries: 5 * 2^17 * 2^32 < 2^64 per 64-bit column) followed by ONE
//     Montgomery row per lane; round constants ride inside that same accumulation;
//   * values stay in [0, 2^256) without modular correction except one predicated subtraction per
//     S-box (bound analysis: DESIGN.md "Operand bounds").
// Input and output are BlsScalar.0 (4 x u64 LE limbs, Montgomery form, < p), bit-exact.
#pragma once
#include <stdint.h>

#include "fr_ptx.cuh"

namespace p252 {

#include "hades_tables.inc"

// hades_tables.inc defines, in the constant bank (statically initialised at module load):
//   kA[68][5][8]  per-round additive constants (scaled)      kG[60][8]  lane-4 correction
//   kF[8]         final multiplier                           kDenseArc / kDenseMds  dense tables
// Every thread of a warp reads the same word in the same instruction (the round index is
// warp-uniform), which the constant cache serves as a broadcast operand.

constexpr int kRounds = 68;
constexpr int kHalfFull = 4;
constexpr int kPartial = 60;

// r = (x*y + m p) / 2^256.  Row operand x must satisfy x + p <= 2^256; y < 2^256.
__device__ __forceinline__ void montmul(uint32_t (&r)[8], const uint32_t (&x)[8],
                                        const uint32_t (&y)[8]) {
    uint32_t a[8], b[8];
    fr_row_first(a, b, x, y[0]);
    fr_row(b, a, x, y[1]);
    fr_row(a, b, x, y[2]);
    fr_row(b, a, x, y[3]);
    fr_row(a, b, x, y[4]);
    fr_row(b, a, x, y[5]);
    fr_row(a, b, x, y[6]);
    fr_row(b, a, x, y[7]);
    fr_merge(r, b, a);
}

// r = (a*a + m p) / 2^256: 36-product square, then eight Montgomery rows on the low half + high half.
__device__ __forceinline__ void montsqr(uint32_t (&r)[8], const uint32_t (&a)[8]) {
    uint32_t t[16];
    fr_sqr_wide(t, a);
    fr_redc_wide(r, t);
}

// z = u^5 / R^4 (unreduced): two squarings and one product like quintic_s_box
// (/root/reference/src/hades/permutation/scalar.rs:50-52).  u < 1.0003 p  =>  z < 1.89 p.
__device__ __forceinline__ void sbox(uint32_t (&z)[8], const uint32_t (&u)[8]) {
    uint32_t a[8], b[8];
    montsqr(a, u);             // < 1.4533 p
    montsqr(b, a);             // < 1.9564 p
    montmul(z, u, b);          // row operand u (u + p <= 2^256), result < 1.8861 p
}

__device__ __forceinline__ void load_const(uint32_t (&d)[8], const uint32_t* c) {
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = c[k];
}

// One lane of the MDS layer: t = sum_j C[i][j] * z[j] as even/odd 64-bit column sums.
template <int I>
__device__ __forceinline__ void mix_lane(uint32_t (&t)[9], const uint32_t (&z)[5][8]) {
    uint64_t e[4], o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        e[k] = (uint64_t)hades_cmat(I, 0) * z[0][2 * k];
        o[k] = (uint64_t)hades_cmat(I, 0) * z[0][2 * k + 1];
    }
#pragma unroll
    for (int j = 1; j < 5; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            e[k] += (uint64_t)hades_cmat(I, j) * z[j][2 * k];
            o[k] += (uint64_t)hades_cmat(I, j) * z[j][2 * k + 1];
        }
    }
    uint32_t e32[8], o32[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        e32[2 * k] = (uint32_t)e[k];
        e32[2 * k + 1] = (uint32_t)(e[k] >> 32);
        o32[2 * k] = (uint32_t)o[k];
        o32[2 * k + 1] = (uint32_t)(o[k] >> 32);
    }
    fr_mix_sum(t, e32, o32);
}

// u = redc1(C z + A[next_round])  -- mul_matrix (+ the next add_round_constants) of the reference,
// /root/reference/src/hades/permutation/scalar.rs:39-48,54-64.  next_round < 0: no constants.
__device__ __forceinline__ void mix(uint32_t (&u)[5][8], const uint32_t (&z)[5][8], int next_round) {
    uint32_t t[9], c[8];
#define P252_MIX_LANE(I)                              \
    mix_lane<I>(t, z);                                \
    if (next_round >= 0) {                            \
        load_const(c, kA[next_round][I]);             \
        fr_arc_redc1(u[I], t, c);                     \
    } else {                                          \
        fr_redc1(u[I], t);                            \
    }
    P252_MIX_LANE(0)
    P252_MIX_LANE(1)
    P252_MIX_LANE(2)
    P252_MIX_LANE(3)
    P252_MIX_LANE(4)
#undef P252_MIX_LANE
}

// In-register Hades permutation, standard Montgomery form in and out (both < p).
__device__ __forceinline__ void hades_permute(uint32_t (&s)[5][8]) {
    uint32_t c[8];
    // first add_round_constants: explicit, then one full conditional subtraction
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        load_const(c, kA[0][i]);
        uint32_t t[8];
        fr_add_lazy(t, s[i], c);
        fr_condsub(t);
#pragma unroll
        for (int k = 0; k < 8; ++k) s[i][k] = t[k];
    }
    uint32_t z[5][8];
#pragma unroll 1
    for (int r = 0; r < kRounds; ++r) {
        const bool full = (r < kHalfFull) || (r >= kHalfFull + kPartial);
        if (full) {
            // S-box on every lane: process slot 4 and rotate, so that one code instance serves all
#pragma unroll 1
            for (int it = 0; it < 5; ++it) {
                uint32_t w[8];
                sbox(w, s[4]);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    s[4][k] = s[3][k];
                    s[3][k] = s[2][k];
                    s[2][k] = s[1][k];
                    s[1][k] = s[0][k];
                    s[0][k] = w[k];
                }
            }
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int k = 0; k < 8; ++k) z[i][k] = s[i][k];
        } else {
            uint32_t w[8];
            sbox(w, s[4]);
            load_const(c, kG[r - kHalfFull]);
            montmul(z[4], c, w);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 8; ++k) z[i][k] = s[i][k];
        }
        mix(s, z, (r + 1 < kRounds) ? (r + 1) : -1);
    }
    // leave the scaled domain: out = montmul(F, v) fully reduced
    load_const(c, kF);
#pragma unroll 1
    for (int it = 0; it < 5; ++it) {
        uint32_t w[8];
        montmul(w, c, s[4]);
        fr_condsub(w);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s[4][k] = s[3][k];
            s[3][k] = s[2][k];
            s[2][k] = s[1][k];
            s[1][k] = s[0][k];
            s[0][k] = w[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Modular add / sub on fully reduced operands (Safe::add, Encryption::subtract,
// /root/reference/src/hades/permutation/scalar.rs:33-35,69-75)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fr_add_mod(uint32_t (&r)[8], const uint32_t (&a)[8],
                                           const uint32_t (&b)[8]) {
    fr_add_lazy(r, a, b);
    fr_condsub(r);
}

// ---------------------------------------------------------------------------------------------
// Dense formulation: the reference's algorithm verbatim on the device (25 full products per MDS,
// full reduction after every operation).  Cross-check / "what a straight port would cost".
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dense_mul(uint32_t (&r)[8], const uint32_t (&x)[8],
                                          const uint32_t (&y)[8]) {
    montmul(r, x, y);          // x < p  => x + p < 2^256
    fr_condsub(r);
}

__device__ __forceinline__ void dense_permute(uint32_t (&s)[5][8]) {
#pragma unroll 1
    for (int r = 0; r < kRounds; ++r) {
        const bool full = (r < kHalfFull) || (r >= kHalfFull + kPartial);
        uint32_t c[8], t[8];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            load_const(c, kDenseArc[r * 5 + i]);
            fr_add_mod(t, s[i], c);
#pragma unroll
            for (int k = 0; k < 8; ++k) s[i][k] = t[k];
        }
#pragma unroll 1
        for (int it = 0; it < 5; ++it) {
            uint32_t w[8];
            if (full || it == 0) {
                uint32_t a[8], b[8];
                dense_mul(a, s[4], s[4]);
                dense_mul(b, a, a);
                dense_mul(w, s[4], b);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) w[k] = s[4][k];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                s[4][k] = s[3][k];
                s[3][k] = s[2][k];
                s[2][k] = s[1][k];
                s[1][k] = s[0][k];
                s[0][k] = w[k];
            }
        }
        uint32_t acc[5][8];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[i][k] = 0;
#pragma unroll 1
        for (int j = 0; j < 5; ++j) {
            // column j of the matrix times lane j; rotate the lanes so indexing stays static
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                load_const(c, kDenseMds[i * 5 + j]);
                uint32_t pr[8], sum[8];
                dense_mul(pr, c, s[0]);
                fr_add_mod(sum, acc[i], pr);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[i][k] = sum[k];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t t0 = s[0][k];
                s[0][k] = s[1][k];
                s[1][k] = s[2][k];
                s[2][k] = s[3][k];
                s[3][k] = s[4][k];
                s[4][k] = t0;
            }
        }
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) s[i][k] = acc[i][k];
    }
}

}  // namespace p252
