// Prototype: BLS12-381 Fr in radix 2^29, 9 unsaturated limbs, Montgomery radix R29 = 2^261.
// All column sums are formed with plain 64-bit multiply-accumulate (IMAD.WIDE without carry chains).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

namespace fr29 {
constexpr int N = 9;
constexpr uint32_t MASK = (1u << 29) - 1;
// p in radix 2^29 (p0 = 1)
HD constexpr uint32_t P(int i) {
    constexpr uint32_t t[9] = {0x00000001, 0x1ffffff8, 0x1f96ffbf, 0x1b4805ff, 0x1d80553b, 0x0c0404d0,
                               0x1520cce7, 0x0a6533af, 0x0073eda7};
    return t[i];
}

// r = a*b/2^261 mod-ish (unreduced Montgomery): inputs limbs < 2^29 (a few may be slightly larger), output
// limbs < 2^29 except the top one.
HD void montmul(uint32_t (&r)[N], const uint32_t (&a)[N], const uint32_t (&b)[N]) {
    uint64_t c[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; ++k) c[k] = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) c[i + j] += (uint64_t)a[i] * b[j];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint32_t m = (0u - (uint32_t)c[i]) & MASK;       // p0 = 1  =>  -p^-1 = -1 mod 2^29
#pragma unroll
        for (int j = 1; j < N; ++j) c[i + j] += (uint64_t)m * P(j);
        c[i + 1] += (c[i] + m) >> 29;                           // low 29 bits cancel
    }
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const uint64_t v = c[N + k] + carry;
        r[k] = (k < N - 1) ? ((uint32_t)v & MASK) : (uint32_t)v;
        carry = v >> 29;
    }
}

HD void montsqr(uint32_t (&r)[N], const uint32_t (&a)[N]) {
    uint64_t c[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; ++k) c[k] = 0;
    uint32_t a2[N];
#pragma unroll
    for (int i = 0; i < N; ++i) a2[i] = a[i] << 1;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        c[2 * i] += (uint64_t)a[i] * a[i];
#pragma unroll
        for (int j = i + 1; j < N; ++j) c[i + j] += (uint64_t)a2[i] * a[j];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint32_t m = (0u - (uint32_t)c[i]) & MASK;
#pragma unroll
        for (int j = 1; j < N; ++j) c[i + j] += (uint64_t)m * P(j);
        c[i + 1] += (c[i] + m) >> 29;
    }
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const uint64_t v = c[N + k] + carry;
        r[k] = (k < N - 1) ? ((uint32_t)v & MASK) : (uint32_t)v;
        carry = v >> 29;
    }
}
}  // namespace fr29
