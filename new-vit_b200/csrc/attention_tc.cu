// Encoder attention on tcgen05 for N = 257 tokens, head_dim 64 (ViT-S/B @ 224x224):
//     out[s, i, h*64:(h+1)*64] = softmax(q_i . K^T) V      per (slice s, head h), q pre-scaled
// (reference layers/attention.py:56-69).  Persistent, one CTA per SM, 512 threads; items = (slice, head).
//
// The 257 = 1 + 2*128 tokens are split as: two query tiles of 128 PATCH tokens (tokens 1..128, 129..256) on the
// tensor cores, and the single CLS query on CUDA cores from the same shared-memory K/V tiles.
//
//   warp 0      TMA producer: per item Q0,Q1 (128x64), K, V (272x64: 128+128+16 rows, rows >= 257 zero-filled by
//               the 3D tensor map) into a 2-stage ring (100 KB per stage, 128B swizzle)
//   warp 1      MMA issuer:  S[128x272] = Q K^T   (SS: N=256 + N=16, K=64 -> 8 tcgen05.mma)
//                            O[128x64]  = P V     (TS: A = P from TMEM, B = V MN-major from smem, 17 x K=16)
//   warp 2      TMEM allocator (512 columns: S 272 fp32 | P 136 (bf16 pairs) | O 64 fp32)
//   warps 4-11  softmax: warp (quadrant q = w%4, half hf) owns 32 rows x 136 score columns; pass 1 row max
//               (exchanged between the two halves through smem), pass 2 p = 2^(s*log2e - m*log2e) -> bf16 pairs
//               -> tcgen05.st into P; partial row sums to smem
//   warps 12-15 CLS query on CUDA cores (scores, softmax, P.V from the smem tiles) and the O epilogue
//               (tcgen05.ld O, 1/l, bf16, swizzled staging tile, TMA store)
//
// Per tile the MMA issuer queues QK(g+1) ahead of PV(g), so the tensor pipe works while the softmax warps are in
// pass 1 of the next tile; the kernel is bound by the exponentials (MUFU), not by the tensor pipe.
#include <math_constants.h>
#include "common.cuh"
#include "ptx.cuh"

namespace mst {

namespace atc {
constexpr int N_TOK = 257;
constexpr int KEYS_PAD = 272;               // 17 x 16
constexpr int Q_TILE_BYTES = 128 * 128;     // 16 KB
constexpr int KV_BYTES = KEYS_PAD * 128;    // 34816
constexpr int STAGE_BYTES = 2 * Q_TILE_BYTES + 2 * KV_BYTES;  // 102400
constexpr int NUM_STAGES = 2;
constexpr int OSTG_OFF = NUM_STAGES * STAGE_BYTES;            // 4 x 4 KB O staging tiles
constexpr int STATS_OFF = OSTG_OFF + 8 * 2048;                // [2 parity][max h0, max h1, sum h0, sum h1][128] floats
constexpr int CLS_OFF = STATS_OFF + 2 * 2 * 2 * 128 * 4;      // pbuf[272] + red[16] + part[128] floats
constexpr int BAR_OFF = CLS_OFF + (272 + 16 + 256) * 4;
constexpr int NUM_BARS = 2 * NUM_STAGES + 1 + 2 + 1 + 1;      // kv_full[2], kv_empty[2], s_full, sp_done[2], o_full, o_free
constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16;
constexpr int DYN_BYTES = TOTAL + 1024;
static_assert(DYN_BYTES <= 232448, "shared memory budget");
constexpr int THREADS = 512;
constexpr uint32_t TM_S = 0, TM_P = 272, TM_O = 408;
constexpr float LOG2E = 1.4426950408889634f;
}  // namespace atc

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* desc, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(desc)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// pass-2 math for 32 score columns already in registers: p = 2^(s*log2e - mb) -> 16 packed bf16 pairs; returns sum(p)
template <bool kMasked>
__device__ __forceinline__ float softmax_math32(const uint32_t (&r)[32], uint32_t (&o)[16], float mb, int n_valid) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), atc::LOG2E, -mb));
        float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), atc::LOG2E, -mb));
        if (kMasked) {
            if (i >= n_valid) p0 = 0.f;
            if (i + 1 >= n_valid) p1 = 0.f;
        }
        o[i >> 1] = pack_bf16x2(p0, p1);
        s0 += p0;
        s1 += p1;
    }
    return s0 + s1;
}
template <bool kMasked>
__device__ __forceinline__ float max32(const uint32_t (&r)[32], float m, int n_valid) {
    float a = m, b = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        if (!kMasked || i < n_valid) a = fmaxf(a, __uint_as_float(r[i]));
        if (!kMasked || i + 1 < n_valid) b = fmaxf(b, __uint_as_float(r[i + 1]));
    }
    return fmaxf(a, b);
}

__global__ void __launch_bounds__(atc::THREADS, 1)
attention_tc257_kernel(const __grid_constant__ TmaDesc map128, const __grid_constant__ TmaDesc map16,
                       const __grid_constant__ TmaDesc mapO, const bf16* __restrict__ qkv, bf16* __restrict__ out,
                       int num_items, int heads) {
    using namespace atc;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    float* stats = reinterpret_cast<float*>(smem + STATS_OFF);  // [parity][max h0, max h1, sum h0, sum h1][128]
    float* clsbuf = reinterpret_cast<float*>(smem + CLS_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* kv_full = bars;
    uint64_t* kv_empty = bars + NUM_STAGES;
    uint64_t* s_full = bars + 2 * NUM_STAGES;
    uint64_t* sp_done = bars + 2 * NUM_STAGES + 1;  // [2]
    uint64_t* o_full = bars + 2 * NUM_STAGES + 3;
    uint64_t* o_free = bars + 2 * NUM_STAGES + 4;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int E = heads * 64;
    const int my_items = blockIdx.x < num_items ? (num_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map128); tma_prefetch_desc(&map16); tma_prefetch_desc(&mapO);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NUM_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 5); }
        mbar_init(s_full, 1);
        mbar_init(&sp_done[0], 8); mbar_init(&sp_done[1], 8);
        mbar_init(o_full, 1);
        mbar_init(o_free, 8);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_ptr_smem);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            for (int it = 0; it < my_items; ++it) {
                const int item = blockIdx.x + it * gridDim.x;
                const int s = item / heads, h = item % heads;
                const int st = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&kv_empty[st], ph ^ 1);
                uint8_t* base = smem + st * STAGE_BYTES;
                uint8_t* sK = base + 2 * Q_TILE_BYTES;
                uint8_t* sV = sK + KV_BYTES;
                mbar_arrive_expect_tx(&kv_full[st], STAGE_BYTES);
                tma_load_3d(base, &map128, &kv_full[st], h * 64, 1, s);                   // Q tile 0: tokens 1..128
                tma_load_3d(base + Q_TILE_BYTES, &map128, &kv_full[st], h * 64, 129, s);  // Q tile 1: tokens 129..256
                tma_load_3d(sK, &map128, &kv_full[st], E + h * 64, 0, s);
                tma_load_3d(sK + 16384, &map128, &kv_full[st], E + h * 64, 128, s);
                tma_load_3d(sK + 32768, &map16, &kv_full[st], E + h * 64, 256, s);        // token 256 + 15 zero rows
                tma_load_3d(sV, &map128, &kv_full[st], 2 * E + h * 64, 0, s);
                tma_load_3d(sV + 16384, &map128, &kv_full[st], 2 * E + h * 64, 128, s);
                tma_load_3d(sV + 32768, &map16, &kv_full[st], 2 * E + h * 64, 256, s);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc_s256 = umma_idesc_bf16_f32(128, 256);
            constexpr uint32_t idesc_s16 = umma_idesc_bf16_f32(128, 16);
            constexpr uint32_t idesc_pv = umma_idesc_bf16_f32(128, 64) | (1u << 16);  // B (V) is MN-major
            const int n_tiles = 2 * my_items;
            auto issue_pv = [&](int g) {  // O = P(g) . V(item of g)
                const int it = g >> 1;
                const uint32_t v_addr = smem_u32(smem + (it & 1) * STAGE_BYTES + 2 * Q_TILE_BYTES + KV_BYTES);
#pragma unroll 1
                for (int j = 0; j < KEYS_PAD / 16; ++j)
                    umma_bf16_ts(tmem_base + TM_O, tmem_base + TM_P + 8 * j,
                                 umma_desc_sw128_mnmajor(v_addr + j * 2048, KV_BYTES), idesc_pv, j != 0 ? 1u : 0u);
                umma_commit(o_full);
                if (g & 1) umma_commit(&kv_empty[it & 1]);  // last tensor-core read of this stage
            };
            for (int g = 0; g < n_tiles; ++g) {
                const int it = g >> 1, t = g & 1, st = it & 1;
                if (t == 0) { mbar_wait(&kv_full[st], (it >> 1) & 1); tc_fence_after_sync(); }
                if (g > 0) {  // S free and P(g-1) complete
                    mbar_wait(&sp_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
                    tc_fence_after_sync();
                }
                const uint32_t q_addr = smem_u32(smem + st * STAGE_BYTES + t * Q_TILE_BYTES);
                const uint32_t k_addr = smem_u32(smem + st * STAGE_BYTES + 2 * Q_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t da = umma_desc_sw128_kmajor(q_addr + k * 32);
                    umma_bf16_ss(tmem_base + TM_S, da, umma_desc_sw128_kmajor(k_addr + k * 32), idesc_s256, k != 0 ? 1u : 0u);
                    umma_bf16_ss(tmem_base + TM_S + 256, da, umma_desc_sw128_kmajor(k_addr + 256 * 128 + k * 32), idesc_s16,
                                 k != 0 ? 1u : 0u);
                }
                umma_commit(s_full);
                if (g > 0) {
                    if (g > 1) { mbar_wait(o_free, (g - 2) & 1); tc_fence_after_sync(); }
                    issue_pv(g - 1);
                }
            }
            if (n_tiles > 0) {
                const int g = n_tiles;  // drain: PV of the last tile
                mbar_wait(&sp_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
                tc_fence_after_sync();
                if (g > 1) { mbar_wait(o_free, (g - 2) & 1); tc_fence_after_sync(); }
                issue_pv(g - 1);
            }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 12) {
        // ===================== softmax + O epilogue =====================
        const int e = warp - 4, q = e & 3, hf = e >> 2;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t ts = tmem_base + lane_base + TM_S + hf * 136;
        const uint32_t tp = tmem_base + lane_base + TM_P + hf * 68;
        const uint32_t to = tmem_base + lane_base + TM_O + hf * 32;
        const int row = q * 32 + lane;
        const int n_tiles = 2 * my_items;
        uint8_t* ostg = smem + OSTG_OFF + e * 2048;  // 32 rows x 32 dims bf16
        uint4* ostg_row = reinterpret_cast<uint4*>(ostg + lane * 64);
        const int last_valid = N_TOK - (136 + 96);  // valid columns in the last 32-chunk of half 1 (25)

        // O(g) -> 1/l -> bf16 -> staging tile -> TMA store   (runs once PV(g) has completed)
        auto epilogue = [&](int g) {
            mbar_wait(o_full, g & 1);
            mbar_wait(&sp_done[g & 1], (g >> 1) & 1);  // acquire the partner warp's row sums of tile g
            tc_fence_after_sync();
            uint32_t r[32];
            tmem_ld_32x32b_x32(to, r);
            tmem_ld_wait();
            tc_fence_before_sync();
            const float* st_sum = stats + ((g & 1) * 4 + 2) * 128;
            const float inv = 1.0f / (st_sum[row] + st_sum[128 + row]);
            __syncwarp();
            if (lane == 0) { mbar_arrive(o_free); tma_store_wait_read<0>(); }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(r[8 * c + 0]) * inv, __uint_as_float(r[8 * c + 1]) * inv);
                u.y = pack_bf16x2(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv);
                u.z = pack_bf16x2(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv);
                u.w = pack_bf16x2(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv);
                ostg_row[c] = u;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                const int item = blockIdx.x + (g >> 1) * gridDim.x;
                tma_store_3d(&mapO, ostg, (item % heads) * 64 + hf * 32, 1 + (g & 1) * 128 + q * 32, item / heads);
                tma_store_commit();
            }
        };

        for (int g = 0; g < n_tiles; ++g) {
            float* st_max = stats + ((g & 1) * 4 + 0) * 128;
            float* st_sum = stats + ((g & 1) * 4 + 2) * 128;
            mbar_wait(s_full, g & 1);
            tc_fence_after_sync();
            uint32_t ra[32], rb[32];
            // ---- pass 1: row max over this warp's 136 columns; the next TMEM load is in flight while reducing ----
            float m = -CUDART_INF_F;
            tmem_ld_32x32b_x32(ts, ra); tmem_ld_wait();
            tmem_ld_32x32b_x32(ts + 32, rb); m = max32<false>(ra, m, 32); tmem_ld_wait();
            tmem_ld_32x32b_x32(ts + 64, ra); m = max32<false>(rb, m, 32); tmem_ld_wait();
            tmem_ld_32x32b_x32(ts + 96, rb); m = max32<false>(ra, m, 32); tmem_ld_wait();
            if (hf == 0) {
                uint32_t r8[8];
                tmem_ld_32x32b_x8(ts + 128, r8);
                m = max32<false>(rb, m, 32);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) m = fmaxf(m, __uint_as_float(r8[i]));
            } else {
                m = max32<true>(rb, m, last_valid);
            }
            st_max[hf * 128 + row] = m;
            named_bar_sync(2 + q, 64);  // the two column halves of this lane quadrant
            m = fmaxf(m, st_max[(hf ^ 1) * 128 + row]);
            const float mb = m * LOG2E;
            // ---- previous tile: PV(g-1) has finished -> its P columns are free and its O can be stored ----
            if (g > 0) epilogue(g - 1);
            // ---- pass 2 ----
            float sum = 0.f;
            uint32_t o[16];
            tmem_ld_32x32b_x32(ts, ra); tmem_ld_wait();
            tmem_ld_32x32b_x32(ts + 32, rb); sum += softmax_math32<false>(ra, o, mb, 32); tmem_st_32x32b_x16(tp, o); tmem_ld_wait();
            tmem_ld_32x32b_x32(ts + 64, ra); sum += softmax_math32<false>(rb, o, mb, 32); tmem_st_32x32b_x16(tp + 16, o); tmem_ld_wait();
            tmem_ld_32x32b_x32(ts + 96, rb); sum += softmax_math32<false>(ra, o, mb, 32); tmem_st_32x32b_x16(tp + 32, o); tmem_ld_wait();
            {
                uint32_t o4[4] = {0u, 0u, 0u, 0u};
                if (hf == 0) {
                    uint32_t r8[8];
                    tmem_ld_32x32b_x8(ts + 128, r8);
                    sum += softmax_math32<false>(rb, o, mb, 32);
                    tmem_st_32x32b_x16(tp + 48, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(r8[i]), LOG2E, -mb));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(r8[i + 1]), LOG2E, -mb));
                        o4[i >> 1] = pack_bf16x2(p0, p1);
                        sum += p0 + p1;
                    }
                } else {
                    sum += softmax_math32<true>(rb, o, mb, last_valid);
                    tmem_st_32x32b_x16(tp + 48, o);
                }
                tmem_st_32x32b_x4(tp + 64, o4);  // half 1: keys 264..271 are padding -> P = 0
            }
            st_sum[hf * 128 + row] = sum;
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sp_done[g & 1]);
        }
        if (n_tiles > 0) epilogue(n_tiles - 1);
        if (lane == 0) tma_store_wait_all<0>();
        __syncwarp();
    } else if (warp >= 12) {
        // ===================== CLS query (token 0) on CUDA cores, decoupled from the tensor pipeline =====================
        const int q = warp - 12;
        const int te = threadIdx.x - 384;   // 0..127
        float* pbuf = clsbuf;               // [272]
        float* red = clsbuf + 272;          // [16]
        float* part = clsbuf + 288;         // [4 key quarters][64 dims]
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int s = item / heads, h = item % heads;
            const int st = it & 1;
            const uint8_t* sK = smem + st * STAGE_BYTES + 2 * Q_TILE_BYTES;
            const uint8_t* sV = sK + KV_BYTES;
            float qv[64];
            {
                const uint4* qp = reinterpret_cast<const uint4*>(qkv + (static_cast<int64_t>(s) * N_TOK) * 3 * E + h * 64);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 u = __ldg(qp + i);
                    float2 f;
                    f = unpack_bf16x2(u.x); qv[8 * i + 0] = f.x; qv[8 * i + 1] = f.y;
                    f = unpack_bf16x2(u.y); qv[8 * i + 2] = f.x; qv[8 * i + 3] = f.y;
                    f = unpack_bf16x2(u.z); qv[8 * i + 4] = f.x; qv[8 * i + 5] = f.y;
                    f = unpack_bf16x2(u.w); qv[8 * i + 6] = f.x; qv[8 * i + 7] = f.y;
                }
            }
            mbar_wait(&kv_full[st], (it >> 1) & 1);
            // scores: keys te and te+128 together (independent chains), key 256 by thread 0
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
            {
                const int j0 = te, j1 = te + 128;
                const uint8_t* k0 = sK + j0 * 128;
                const uint8_t* k1 = sK + j1 * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 u = *reinterpret_cast<const uint4*>(k0 + ((c ^ (j0 & 7)) << 4));
                    const uint4 w = *reinterpret_cast<const uint4*>(k1 + ((c ^ (j1 & 7)) << 4));
                    float2 f, g2;
                    f = unpack_bf16x2(u.x); g2 = unpack_bf16x2(w.x);
                    a0 = fmaf(qv[8 * c + 0], f.x, a0); a1 = fmaf(qv[8 * c + 1], f.y, a1);
                    b0 = fmaf(qv[8 * c + 0], g2.x, b0); b1 = fmaf(qv[8 * c + 1], g2.y, b1);
                    f = unpack_bf16x2(u.y); g2 = unpack_bf16x2(w.y);
                    a0 = fmaf(qv[8 * c + 2], f.x, a0); a1 = fmaf(qv[8 * c + 3], f.y, a1);
                    b0 = fmaf(qv[8 * c + 2], g2.x, b0); b1 = fmaf(qv[8 * c + 3], g2.y, b1);
                    f = unpack_bf16x2(u.z); g2 = unpack_bf16x2(w.z);
                    a0 = fmaf(qv[8 * c + 4], f.x, a0); a1 = fmaf(qv[8 * c + 5], f.y, a1);
                    b0 = fmaf(qv[8 * c + 4], g2.x, b0); b1 = fmaf(qv[8 * c + 5], g2.y, b1);
                    f = unpack_bf16x2(u.w); g2 = unpack_bf16x2(w.w);
                    a0 = fmaf(qv[8 * c + 6], f.x, a0); a1 = fmaf(qv[8 * c + 7], f.y, a1);
                    b0 = fmaf(qv[8 * c + 6], g2.x, b0); b1 = fmaf(qv[8 * c + 7], g2.y, b1);
                }
            }
            const float sc0 = a0 + a1, sc1 = b0 + b1;
            float sc2 = -CUDART_INF_F;
            if (te == 0) {
                float c0 = 0.f, c1 = 0.f;
                const uint8_t* k2 = sK + 256 * 128;  // 256 & 7 == 0: chunks unswizzled
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 u = *reinterpret_cast<const uint4*>(k2 + (c << 4));
                    float2 f;
                    f = unpack_bf16x2(u.x); c0 = fmaf(qv[8 * c + 0], f.x, c0); c1 = fmaf(qv[8 * c + 1], f.y, c1);
                    f = unpack_bf16x2(u.y); c0 = fmaf(qv[8 * c + 2], f.x, c0); c1 = fmaf(qv[8 * c + 3], f.y, c1);
                    f = unpack_bf16x2(u.z); c0 = fmaf(qv[8 * c + 4], f.x, c0); c1 = fmaf(qv[8 * c + 5], f.y, c1);
                    f = unpack_bf16x2(u.w); c0 = fmaf(qv[8 * c + 6], f.x, c0); c1 = fmaf(qv[8 * c + 7], f.y, c1);
                }
                sc2 = c0 + c1;
            }
            float lmax = fmaxf(fmaxf(sc0, sc1), sc2);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
            if (lane == 0) red[q] = lmax;
            named_bar_sync(1, 128);
            const float mb = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])) * LOG2E;
            const float p0 = ex2_approx(fmaf(sc0, LOG2E, -mb));
            const float p1 = ex2_approx(fmaf(sc1, LOG2E, -mb));
            float lsum = p0 + p1;
            pbuf[te] = p0;
            pbuf[te + 128] = p1;
            if (te == 0) { const float p2 = ex2_approx(fmaf(sc2, LOG2E, -mb)); pbuf[256] = p2; lsum += p2; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
            if (lane == 0) red[4 + q] = lsum;
            named_bar_sync(1, 128);
            const float inv_cls = 1.0f / ((red[4] + red[5]) + (red[6] + red[7]));
            {   // o[d] = sum_j p_j v[j][d]: lane -> dims 2*lane, 2*lane+1; warp q -> keys j = q (mod 4)
                const uint8_t* vbase = sV + (lane & 3) * 4;
                const int dc = lane >> 2;
                float x0 = 0.f, y0 = 0.f, x1 = 0.f, y1 = 0.f;
                int j = q;
#pragma unroll 4
                for (; j + 4 < N_TOK; j += 8) {
                    const float2 va = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vbase + j * 128 + ((dc ^ (j & 7)) << 4)));
                    const float2 vb = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vbase + (j + 4) * 128 + ((dc ^ ((j + 4) & 7)) << 4)));
                    const float pa = pbuf[j], pb = pbuf[j + 4];
                    x0 = fmaf(pa, va.x, x0); y0 = fmaf(pa, va.y, y0);
                    x1 = fmaf(pb, vb.x, x1); y1 = fmaf(pb, vb.y, y1);
                }
                for (; j < N_TOK; j += 4) {
                    const float2 va = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vbase + j * 128 + ((dc ^ (j & 7)) << 4)));
                    const float pa = pbuf[j];
                    x0 = fmaf(pa, va.x, x0); y0 = fmaf(pa, va.y, y0);
                }
                part[q * 64 + 2 * lane] = x0 + x1;
                part[q * 64 + 2 * lane + 1] = y0 + y1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&kv_empty[st]);  // this warp no longer reads the stage
            named_bar_sync(1, 128);
            if (te < 64)
                out[(static_cast<int64_t>(s) * N_TOK) * E + h * 64 + te] =
                    __float2bfloat16_rn(((part[te] + part[64 + te]) + (part[128 + te] + part[192 + te])) * inv_cls);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}

int launch_attention_tc257(const bf16* qkv, bf16* out, int BD, int heads, int num_sms, cudaStream_t stream) {
    using namespace atc;
    const int E = heads * 64;
    TmaDesc m128, m16, mO;
    MST_PROPAGATE(make_tma_3d_bf16(&m128, qkv, 3 * E, N_TOK, BD, 3 * E, static_cast<uint64_t>(N_TOK) * 3 * E, 64, 128, true));
    MST_PROPAGATE(make_tma_3d_bf16(&m16, qkv, 3 * E, N_TOK, BD, 3 * E, static_cast<uint64_t>(N_TOK) * 3 * E, 64, 16, true));
    MST_PROPAGATE(make_tma_3d_bf16(&mO, out, E, N_TOK, BD, E, static_cast<uint64_t>(N_TOK) * E, 32, 32, false));
    static bool attr = false;
    if (!attr) {
        MST_CHECK_CUDA(cudaFuncSetAttribute(attention_tc257_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_BYTES));
        attr = true;
    }
    const int items = BD * heads;
    const int grid = items < num_sms ? items : num_sms;
    attention_tc257_kernel<<<grid, THREADS, DYN_BYTES, stream>>>(m128, m16, mO, qkv, out, items, heads);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mst
