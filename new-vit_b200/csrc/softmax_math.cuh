// Softmax arithmetic shared by the tcgen05 attention kernels (attention_tc16.cu: N = 257; attention_tcg.cu: any N): the exponentials of
// 16 score columns as packed bf16 pairs, part of them on the FMA / ALU pipes instead of the MUFU unit.
#pragma once
#include "ptx.cuh"

namespace mst {

namespace smx {
constexpr float LOG2E = 1.4426950408889634f;
}

// 2^t for a packed pair on the FMA/ALU pipes (no MUFU): t = n + f with n = round(t) (magic-number add), f in [-0.5, 0.5],
// 2^f by a degree-3 minimax polynomial (max relative error 7.5e-5, far below the bf16 rounding of P), 2^n by adding n
// to the exponent field.  t is clamped at -126 (the result underflows to ~1e-38 there).
__device__ __forceinline__ void ex2_poly_pair_16(f32x2 t, float& p0, float& p1) {
    float t0, t1;
    f2_unpack(t, t0, t1);
    t = f2_pack(fmaxf(t0, -126.0f), fmaxf(t1, -126.0f));
    const f32x2 magic = f2_pack(12582912.0f, 12582912.0f), nmagic = f2_pack(-12582912.0f, -12582912.0f);
    const f32x2 r = f2_add(t, magic);                 // low mantissa bits of r = round(t)
    const f32x2 n = f2_add(r, nmagic);
    const f32x2 f = f2_fma(n, f2_pack(-1.0f, -1.0f), t);
    f32x2 q = f2_fma(f2_pack(0.05517145f, 0.05517145f), f, f2_pack(0.24261084f, 0.24261084f));
    q = f2_fma(q, f, f2_pack(0.69326097f, 0.69326097f));
    q = f2_fma(q, f, f2_pack(0.9999281f, 0.9999281f));
    float q0, q1, r0, r1;
    f2_unpack(q, q0, q1);
    f2_unpack(r, r0, r1);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(r0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(r1) << 23));
}

// pass-2 math for 32 score columns already in registers: p = 2^(s*log2e - mb) -> 16 packed bf16 pairs; returns sum(p).
// The exponentials are what bounds this kernel (16 MUFU results per clock and SM): kPolyPairs of every 16 column pairs
// take the polynomial path instead, which balances the MUFU pipe against the issue slots of the sub-partition.
template <int kPolyPairs>
__device__ __forceinline__ float softmax_math32_16(const uint32_t (&r)[32], uint32_t (&o)[16], float mb) {
    const f32x2 l2 = f2_pack(smx::LOG2E, smx::LOG2E), nmb = f2_pack(-mb, -mb);
    f32x2 acc = f2_pack(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const f32x2 t = f2_fma(f2_pack(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), l2, nmb);
        float p0, p1;
        // interleave the two kinds so that MUFU latency hides behind polynomial work
        if ((i * kPolyPairs) % 16 < kPolyPairs) {
            ex2_poly_pair_16(t, p0, p1);
        } else {
            float t0, t1;
            f2_unpack(t, t0, t1);
            p0 = ex2_approx(t0);
            p1 = ex2_approx(t1);
        }
        o[i] = pack_bf16x2(p0, p1);
        acc = f2_add(acc, f2_pack(p0, p1));
    }
    float s0, s1;
    f2_unpack(acc, s0, s1);
    return s0 + s1;
}
// pass-2 math for 16 score columns (8 pairs).  kOff = 0 / 8 selects which half of the 16-pair polynomial/MUFU pattern the
// chunk uses, so that over two consecutive chunks kPolyPairs of every 16 pairs take the polynomial path.
template <int kPolyPairs, int kOff>
__device__ __forceinline__ float softmax_math16(const uint32_t (&r)[16], uint32_t (&o)[8], float mb) {
    const f32x2 l2 = f2_pack(smx::LOG2E, smx::LOG2E), nmb = f2_pack(-mb, -mb);
    f32x2 acc = f2_pack(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const f32x2 t = f2_fma(f2_pack(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), l2, nmb);
        float p0, p1;
        if (((i + kOff) * kPolyPairs) % 16 < kPolyPairs) {
            ex2_poly_pair_16(t, p0, p1);
        } else {
            float t0, t1;
            f2_unpack(t, t0, t1);
            p0 = ex2_approx(t0);
            p1 = ex2_approx(t1);
        }
        o[i] = pack_bf16x2(p0, p1);
        acc = f2_add(acc, f2_pack(p0, p1));
    }
    float s0, s1;
    f2_unpack(acc, s0, s1);
    return s0 + s1;
}
}  // namespace mst
