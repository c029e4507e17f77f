// Encoder attention on tcgen05 for N = 257 tokens, head_dim 64 -- SIXTEEN softmax warps (a TMEM lane quadrant is readable by warps with warp % 4 == q only).
//     out[s, i, h*64:(h+1)*64] = softmax(q_i . K^T) V      per (slice, head), q pre-scaled
// (reference layers/attention.py:56-69).  Persistent, one CTA per SM, 768 threads; items = (slice, head).
//
// Data flow: (two 128-query tiles per item against the 256 patch keys on the tensor cores, two
// 256-column TMEM buffers = two tiles in flight, the CLS token as a key on CUDA cores and as a query on warp-level MMA),
// but each tile is now worked on by EIGHT warps instead of four: a TMEM lane quadrant (32 query rows) can only be read
// by warps with warp % 4 == quadrant, so the two warps of a quadrant split the 256 KEY columns, exchange row max and row
// sum through shared memory, and split the 64 output dims in the epilogue.  The eight-warp version was latency-bound on
// each team's serial chain  S -> max -> exp -> P -> (P.V) -> O  (issue slots 48 %, MUFU 26 %, tensor pipe 26 % busy, ncu);
// halving the per-warp work of every phase shortens that chain and puts four softmax warps on every SM sub-partition.
//
//   warp 0        TMA producer
//   warp 1        MMA issuer:  S = Q K^T (SS, M128 N256 K16 x 4);  O = P V (TS, N64 K16 x 16), P of key half 0 in columns
//                 [0,64), of key half 1 in [128,192) (each in place at the start of its own half of S), O in [64,128)
//   warps 4-19    softmax + epilogue: warp e -> quadrant e%4, team (e/4)%2 = tile parity = TMEM buffer, key half e/8
//   warps 2,20,21 / 3,22,23   CLS query (warp-level MMA), two groups of three on alternate items
#include <math_constants.h>
#include "common.cuh"
#include "ptx.cuh"
#include "softmax_math.cuh"

namespace mst {

namespace atc16 {
constexpr int N_TOK = 257;
constexpr int Q_TILE_BYTES = 128 * 128;     // 16 KB
constexpr int KV_BYTES = 256 * 128;         // 32 KB: tokens 1..256
constexpr int KV0_BYTES = 16 * 128;         // 2 KB: tokens 0..15, only row 0 (CLS) is used
constexpr int STAGE_BYTES = 2 * Q_TILE_BYTES + 2 * KV_BYTES + 2 * KV0_BYTES;  // 102400
constexpr int OFF_K = 2 * Q_TILE_BYTES, OFF_V = OFF_K + KV_BYTES, OFF_K0 = OFF_V + KV_BYTES, OFF_V0 = OFF_K0 + KV0_BYTES;
constexpr int NUM_STAGES = 2;
constexpr int OSTG_OFF = NUM_STAGES * STAGE_BYTES;            // 8 x 2 KB O staging tiles
constexpr int STATS_OFF = OSTG_OFF + 8 * 2048;                // [2 parity][6 kinds][128] floats
constexpr int STATS_KINDS = 7;                                // per team: max[2 halves], sum[2 halves], p0 (CLS key), 2 spare
constexpr int CLS_OFF = STATS_OFF + 2 * STATS_KINDS * 128 * 4;  // pbuf[272] + red[16] + part[256] floats
constexpr int BAR_OFF = CLS_OFF + (272 + 16 + 256) * 4;
constexpr int NUM_BARS = 2 * NUM_STAGES + 2 + 2 + 2 + 2;      // kv_full[2], kv_empty[2], s_full[2], sp_done[2], o_full[2], o_free[2]
constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16;
constexpr int DYN_BYTES = TOTAL + 1024;
static_assert(DYN_BYTES <= 232448, "shared memory budget");
constexpr int THREADS = 768;
constexpr float LOG2E = 1.4426950408889634f;
}  // namespace atc16

__device__ __forceinline__ void tma_load_3d_16(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d_16(const void* desc, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(desc)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x8_16(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ float max32_16(const uint32_t (&r)[32], float m) {
    float a = m, b = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        a = fmaxf(a, __uint_as_float(r[i]));
        b = fmaxf(b, __uint_as_float(r[i + 1]));
    }
    return fmaxf(a, b);
}
// dot of 8 bf16 (one 16-byte chunk) with 8 floats
__device__ __forceinline__ float dot8_16(const uint4& u, const float* q, float a) {
    float2 f;
    f = unpack_bf16x2(u.x); a = fmaf(q[0], f.x, a); a = fmaf(q[1], f.y, a);
    f = unpack_bf16x2(u.y); a = fmaf(q[2], f.x, a); a = fmaf(q[3], f.y, a);
    f = unpack_bf16x2(u.z); a = fmaf(q[4], f.x, a); a = fmaf(q[5], f.y, a);
    f = unpack_bf16x2(u.w); a = fmaf(q[6], f.x, a); a = fmaf(q[7], f.y, a);
    return a;
}

// phase timing of one softmax warp of block 0 (dbg != nullptr only in profiles/attn_timing.py)
#undef ATT_T
#define ATT_T(i) do { if (dbg_on) { const long long _t = clock64(); dbg_acc[i] += _t - dbg_t; dbg_t = _t; } } while (0)

// kLse (training forward): the log2-domain row log-sum-exp  lse[item][token] = max * log2e + log2(sum)  is written for the backward
// pass (train_enc.cu), which then skips its own recomputation; the inference instantiations do not carry the code.
template <int kPolyPairs, bool kLse = false>
__global__ void __launch_bounds__(atc16::THREADS, 1)
attention_tc257x16_kernel(const __grid_constant__ TmaDesc map128, const __grid_constant__ TmaDesc map16,
                       const __grid_constant__ TmaDesc mapO, const bf16* __restrict__ qkv, bf16* __restrict__ out,
                       int num_items, int heads, long long* __restrict__ dbg, float* __restrict__ lse_out) {
    using namespace atc16;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* kv_full = bars;        // [2] TMA -> everyone
    uint64_t* kv_empty = bars + 2;   // [2] 1 (MMA commit) + 4 (CLS-query warps) + 8 (softmax warps, after their last epilogue)
    uint64_t* s_full = bars + 4;     // [2] MMA -> softmax: S(g) complete in buffer g&1
    uint64_t* sp_done = bars + 6;    // [2] softmax -> MMA: P(g) written (and S consumed)
    uint64_t* o_full = bars + 8;     // [2] MMA -> softmax: O(g) complete
    uint64_t* o_free = bars + 10;    // [2] softmax -> MMA: O(g) read out, buffer g&1 reusable
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int E = heads * 64;
    const int my_items = blockIdx.x < num_items ? (num_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int n_tiles = 2 * my_items;

    griddep_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map128); tma_prefetch_desc(&map16); tma_prefetch_desc(&mapO);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 20);  // MMA commit + 16 softmax warps + 3 CLS warps
            mbar_init(&s_full[i], 1); mbar_init(&sp_done[i], 8);
            mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 8);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_ptr_smem);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_smem;
    griddep_wait();   // qkv comes from the kernel in front

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            for (int it = 0; it < my_items; ++it) {
                const int item = blockIdx.x + it * gridDim.x;
                const int s = item / heads, h = item % heads;
                const int st = it & 1;
                mbar_wait(&kv_empty[st], ((it >> 1) & 1) ^ 1);
                uint8_t* base = smem + st * STAGE_BYTES;
                mbar_arrive_expect_tx(&kv_full[st], STAGE_BYTES);
                tma_load_3d_16(base, &map128, &kv_full[st], h * 64, 1, s);                      // Q tile 0: tokens 1..128
                tma_load_3d_16(base + Q_TILE_BYTES, &map128, &kv_full[st], h * 64, 129, s);     // Q tile 1: tokens 129..256
                tma_load_3d_16(base + OFF_K, &map128, &kv_full[st], E + h * 64, 1, s);          // K tokens 1..128
                tma_load_3d_16(base + OFF_K + 16384, &map128, &kv_full[st], E + h * 64, 129, s);
                tma_load_3d_16(base + OFF_V, &map128, &kv_full[st], 2 * E + h * 64, 1, s);
                tma_load_3d_16(base + OFF_V + 16384, &map128, &kv_full[st], 2 * E + h * 64, 129, s);
                tma_load_3d_16(base + OFF_K0, &map16, &kv_full[st], E + h * 64, 0, s);          // row 0 = K of the CLS token
                tma_load_3d_16(base + OFF_V0, &map16, &kv_full[st], 2 * E + h * 64, 0, s);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp convergent, one elected lane issues) =====================
        constexpr uint32_t idesc_s = umma_idesc_bf16_f32(128, 256);
        constexpr uint32_t idesc_pv = umma_idesc_bf16_f32(128, 64) | (1u << 16);              // B (V) is MN-major
        constexpr uint32_t kDescHiK = (1024u >> 4) | (1u << 14) | (2u << 29);                  // K-major SW128, SBO 1024
        constexpr uint32_t kDescHiV = (1024u >> 4) | (1u << 14) | (2u << 29);                  // MN-major SW128: same fields
        constexpr uint32_t kLboV = (static_cast<uint32_t>(KV_BYTES) >> 4) << 16;               // LBO (unused: N == 64)
        const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFF) >> 4;
        auto issue_pv = [&](int g) {  // O(g) = P(g) . V(item of g): P in columns [0,64) + [128,192), O into [64,128) of buffer g&1
            const int it = g >> 1;
            const uint32_t buf = tmem_base + static_cast<uint32_t>((g & 1) * 256);
            const uint32_t v_lo = smem_lo + (((it & 1) * STAGE_BYTES + OFF_V) >> 4);
            if (elect_one_sync()) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    umma_bf16_ts(buf + 64, buf + (j < 8 ? 8 * j : 128 + 8 * (j - 8)),
                                 make_desc((v_lo + j * (2048 >> 4)) | kLboV, kDescHiV), idesc_pv, j != 0 ? 1u : 0u);
                umma_commit(&o_full[g & 1]);
                if (g & 1) umma_commit(&kv_empty[it & 1]);  // last tensor-core read of this stage
            }
            __syncwarp();
        };
        for (int g = 0; g < n_tiles; ++g) {
            const int it = g >> 1, t = g & 1, st = it & 1, b = g & 1;
            if (t == 0) { mbar_wait(&kv_full[st], (it >> 1) & 1); tc_fence_after_sync(); }
            if (g >= 2) { mbar_wait(&o_free[b], ((g - 2) >> 1) & 1); tc_fence_after_sync(); }   // O(g-2) read out
            {
                const uint32_t q_lo = smem_lo + ((st * STAGE_BYTES + t * Q_TILE_BYTES) >> 4);
                const uint32_t k_lo = smem_lo + ((st * STAGE_BYTES + OFF_K) >> 4);
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base + static_cast<uint32_t>(b * 256), make_desc(q_lo + 2 * k, kDescHiK),
                                     make_desc(k_lo + 2 * k, kDescHiK), idesc_s, k != 0 ? 1u : 0u);
                    umma_commit(&s_full[b]);
                }
                __syncwarp();
            }
            if (g > 0) {  // P(g-1) complete
                mbar_wait(&sp_done[b ^ 1], ((g - 1) >> 1) & 1);
                tc_fence_after_sync();
                issue_pv(g - 1);
            }
        }
        if (n_tiles > 0) {
            const int g = n_tiles;  // drain: PV of the last tile
            mbar_wait(&sp_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
            tc_fence_after_sync();
            issue_pv(g - 1);
        }
    } else if (warp >= 4 && warp < 20) {
        // ===================== softmax + O epilogue: two teams of eight warps =====================
        // Team tm owns the tiles g = tm (mod 2), i.e. TMEM buffer tm; warp (q, tm, ch) owns rows q*32.. of its tile and the
        // 128 score columns of key half ch.  Row max / row sum / p0 cross the two halves through shared memory.
        const int e = warp - 4, q = e & 3, tm = (e >> 2) & 1, ch = e >> 3;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t buf = tmem_base + lane_base + static_cast<uint32_t>(tm * 256);
        const uint32_t sbase = buf + static_cast<uint32_t>(ch * 128);     // this warp's S columns; its P starts here too
        const int row = q * 32 + lane;
        float* st = reinterpret_cast<float*>(smem + STATS_OFF) + tm * STATS_KINDS * 128;   // max[2][128] sum[2][128] p0[128]
        const uint32_t bar_id = 1 + tm * 4 + q;                           // named barrier of the (team, quadrant) warp pair
        for (int g = tm; g < n_tiles; g += 2) {
            const int it = g >> 1, stg = it & 1;
            const uint32_t par = static_cast<uint32_t>(it & 1);
            mbar_wait(&s_full[tm], par);
            tc_fence_after_sync();
            uint32_t ra[16], rb[16];
            tmem_ld_32x32b_x16(sbase, ra);
            // ---- CLS key: s0 = q_row . k_cls (CUDA cores, from the smem tiles); key half 0 owns it ----
            float s0 = -CUDART_INF_F;
            if (ch == 0) {
                s0 = 0.f;
                const uint8_t* qrow = smem + stg * STAGE_BYTES + tm * Q_TILE_BYTES + row * 128;
                const uint8_t* k0 = smem + stg * STAGE_BYTES + OFF_K0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 kk = *reinterpret_cast<const uint4*>(k0 + c * 16);
                    const uint4 qq = *reinterpret_cast<const uint4*>(qrow + ((c ^ (row & 7)) << 4));
                    float kf[8];
                    float2 f;
                    f = unpack_bf16x2(kk.x); kf[0] = f.x; kf[1] = f.y;
                    f = unpack_bf16x2(kk.y); kf[2] = f.x; kf[3] = f.y;
                    f = unpack_bf16x2(kk.z); kf[4] = f.x; kf[5] = f.y;
                    f = unpack_bf16x2(kk.w); kf[6] = f.x; kf[7] = f.y;
                    s0 = dot8_16(qq, kf, s0);
                }
            }
            // ---- pass 1: row max over this warp's 128 columns; the next TMEM load is in flight while reducing ----
            float m = s0;
            auto max16 = [&](const uint32_t (&r)[16]) {
                float a = m, b2 = -CUDART_INF_F;
#pragma unroll
                for (int i = 0; i < 16; i += 2) { a = fmaxf(a, __uint_as_float(r[i])); b2 = fmaxf(b2, __uint_as_float(r[i + 1])); }
                m = fmaxf(a, b2);
            };
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                tmem_ld_wait();
                tmem_ld_32x32b_x16(sbase + (c + 1) * 16, rb);
                max16(ra);
                tmem_ld_wait();
                if (c + 2 < 8) tmem_ld_32x32b_x16(sbase + (c + 2) * 16, ra);
                max16(rb);
            }
            st[ch * 128 + row] = m;
            tmem_ld_32x32b_x16(sbase, ra);            // first chunk of pass 2 rides over the exchange
            named_bar_sync(bar_id, 64);
            m = fmaxf(m, st[(ch ^ 1) * 128 + row]);
            const float mb = m * LOG2E;
            // ---- pass 2: P (bf16 pairs) in place at the start of this warp's half: chunk c (16 columns) -> 8 columns ----
            float sum = 0.f;
            if (ch == 0) {
                sum = ex2_approx(fmaf(s0, LOG2E, -mb));   // probability of the CLS key: rank-1 term of the epilogue
                st[4 * 128 + row] = sum;
            }
            uint32_t o[8];
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                tmem_ld_wait();
                tmem_ld_32x32b_x16(sbase + (c + 1) * 16, rb);
                sum += softmax_math16<kPolyPairs, 0>(ra, o, mb); tmem_st_32x32b_x8_16(sbase + c * 8, o);
                tmem_ld_wait();
                if (c + 2 < 8) tmem_ld_32x32b_x16(sbase + (c + 2) * 16, ra);
                sum += softmax_math16<kPolyPairs, 8>(rb, o, mb); tmem_st_32x32b_x8_16(sbase + (c + 1) * 8, o);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            st[(2 + ch) * 128 + row] = sum;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sp_done[tm]);
            named_bar_sync(bar_id, 64);
            const float tot = sum + st[(2 + (ch ^ 1)) * 128 + row];
            const float inv = 1.0f / tot;
            const float p0 = st[4 * 128 + row];
            if (kLse && ch == 0)
                lse_out[static_cast<int64_t>(blockIdx.x + it * gridDim.x) * N_TOK + 1 + tm * 128 + row] = mb + log2f(tot);
            // ---- epilogue: O columns [64 + 32 ch, +32) of this warp's rows + p0 * v_cls, 1/l, bf16, 64-byte row pieces ----
            mbar_wait(&o_full[tm], par);
            tc_fence_after_sync();
            uint32_t ro[32];
            tmem_ld_32x32b_x32(buf + 64 + ch * 32, ro);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[tm]);
            {
                const int item = blockIdx.x + it * gridDim.x;
                const int s = item / heads, h = item % heads;
                const uint8_t* v0 = smem + stg * STAGE_BYTES + OFF_V0 + ch * 64;  // V of the CLS token: row 0, chunks unswizzled
                uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(s) * N_TOK + 1 + tm * 128 + row) * E + h * 64 + ch * 32);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t* r = &ro[8 * c];
                    const uint4 vv = *reinterpret_cast<const uint4*>(v0 + c * 16);
                    float2 f;
                    uint4 u;
                    f = unpack_bf16x2(vv.x);
                    u.x = pack_bf16x2(fmaf(p0, f.x, __uint_as_float(r[0])) * inv, fmaf(p0, f.y, __uint_as_float(r[1])) * inv);
                    f = unpack_bf16x2(vv.y);
                    u.y = pack_bf16x2(fmaf(p0, f.x, __uint_as_float(r[2])) * inv, fmaf(p0, f.y, __uint_as_float(r[3])) * inv);
                    f = unpack_bf16x2(vv.z);
                    u.z = pack_bf16x2(fmaf(p0, f.x, __uint_as_float(r[4])) * inv, fmaf(p0, f.y, __uint_as_float(r[5])) * inv);
                    f = unpack_bf16x2(vv.w);
                    u.w = pack_bf16x2(fmaf(p0, f.x, __uint_as_float(r[6])) * inv, fmaf(p0, f.y, __uint_as_float(r[7])) * inv);
                    dst[c] = u;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&kv_empty[stg]);  // Q rows / K_cls / V_cls of this stage are no longer needed
        }
    } else {
        // ===================== CLS query (token 0): warp-level MMA from the same smem tiles =====================
        // Six warps in two groups of three (group A = warps 2,12,13 takes even items, group B = 3,14,15 odd items).  The
        // CLS query is the only valid row of a 16-row mma.sync block; keys = 8 chunks of 32 patch keys + one chunk holding
        // the CLS key (row 0 of the K0/V0 boxes).  Member i of a group runs the chunks i, i+3, i+6 with an online softmax
        // in registers; the three partial (max, sum, O) triples are merged through shared memory.  A single warp needs
        // ~21k cycles per item (long mma.sync dependency chains on a busy sub-partition) and was the kernel's bottleneck.
        const int grp = (warp == 2 || warp == 20 || warp == 21) ? 0 : 1;
        const int mem = (warp == 2 || warp == 3) ? 0 : ((warp == 20 || warp == 22) ? 1 : 2);
        const int gq = lane >> 2, tq = lane & 3;
        float* merge = reinterpret_cast<float*>(smem + CLS_OFF) + grp * 3 * 68;  // per member: m, l, pad, pad, o[64]
        for (int it = grp; it < my_items; it += 2) {
            const int item = blockIdx.x + it * gridDim.x;
            const int s = item / heads, h = item % heads;
            const int st = it & 1;
            const uint32_t sK_u = smem_u32(smem + st * STAGE_BYTES + OFF_K), sV_u = smem_u32(smem + st * STAGE_BYTES + OFF_V);
            const uint32_t sK0_u = smem_u32(smem + st * STAGE_BYTES + OFF_K0), sV0_u = smem_u32(smem + st * STAGE_BYTES + OFF_V0);
            uint32_t qa[4][4];
            {
                const uint32_t* q0 = reinterpret_cast<const uint32_t*>(qkv + (static_cast<int64_t>(s) * N_TOK) * 3 * E + h * 64);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    qa[ks][0] = gq == 0 ? __ldg(q0 + ks * 8 + tq) : 0u;
                    qa[ks][1] = 0u;
                    qa[ks][2] = gq == 0 ? __ldg(q0 + ks * 8 + 4 + tq) : 0u;
                    qa[ks][3] = 0u;
                }
            }
            float oacc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) { oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f; }
            float m0 = -CUDART_INF_F, l0 = 0.f;
            mbar_wait(&kv_full[st], (it >> 1) & 1);
#pragma unroll 1
            for (int ci = mem; ci < 9; ci += 3) {  // chunks 0..7: 32 patch keys each; chunk 8: the CLS key block
                const bool cls_chunk = ci == 8;
                const uint32_t kb = cls_chunk ? sK0_u : sK_u + ci * 32 * 128;
                const uint32_t vb = cls_chunk ? sV0_u : sV_u + ci * 32 * 128;
                const int nblk = cls_chunk ? 1 : 4;
                float sacc[4][4];
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    sacc[nb][0] = sacc[nb][1] = sacc[nb][2] = sacc[nb][3] = 0.f;
                    if (nb < nblk) {
                        const int key = nb * 8 + (lane & 7);  // row inside the chunk's tile; (row & 7) == (key & 7)
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const int c = half * 4 + (lane >> 3);
                            uint32_t bfr[4];
                            ldmatrix_x4(bfr, kb + key * 128 + ((c ^ (key & 7)) << 4));
                            mma_bf16_16816(sacc[nb], qa[2 * half], bfr[0], bfr[1]);
                            mma_bf16_16816(sacc[nb], qa[2 * half + 1], bfr[2], bfr[3]);
                        }
                    }
                }
                if (cls_chunk) {  // only key 0 of the box is the CLS token; rows 1..7 are other tokens
                    if (tq != 0) sacc[0][0] = -CUDART_INF_F;
                    sacc[0][1] = -CUDART_INF_F;
                }
                float cm = -CUDART_INF_F;
#pragma unroll
                for (int nb = 0; nb < 4; ++nb)
                    if (nb < nblk) cm = fmaxf(cm, fmaxf(sacc[nb][0], sacc[nb][1]));
                cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
                cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
                const float mn = fmaxf(m0, cm);
                const float sc = ex2_approx((m0 - mn) * LOG2E);
                m0 = mn;
                const float ms = mn * LOG2E;
                float rs = 0.f;
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    if (nb < nblk) {
                        sacc[nb][0] = ex2_approx(fmaf(sacc[nb][0], LOG2E, -ms));
                        sacc[nb][1] = ex2_approx(fmaf(sacc[nb][1], LOG2E, -ms));
                        rs += sacc[nb][0] + sacc[nb][1];
                    }
                    sacc[nb][2] = 0.f; sacc[nb][3] = 0.f;  // rows 8..15 of the block do not exist
                }
                l0 = fmaf(l0, sc, rs);
#pragma unroll
                for (int i = 0; i < 8; ++i) { oacc[i][0] *= sc; oacc[i][1] *= sc; }
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    if (2 * kk < nblk) {
                        uint32_t pa[4];
                        pa[0] = pack_bf16x2(sacc[2 * kk][0], sacc[2 * kk][1]);
                        pa[1] = 0u;
                        pa[2] = pack_bf16x2(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
                        pa[3] = 0u;
                        const int key = kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
                        for (int dn = 0; dn < 8; dn += 2) {
                            const int c = dn + (lane >> 4);
                            uint32_t bfr[4];
                            ldmatrix_x4_trans(bfr, vb + key * 128 + ((c ^ (key & 7)) << 4));
                            mma_bf16_16816(oacc[dn], pa, bfr[0], bfr[1]);
                            mma_bf16_16816(oacc[dn + 1], pa, bfr[2], bfr[3]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&kv_empty[st]);  // this warp no longer reads the stage
            // ---- merge the three members' partials (row 0 lives in lanes 0..3: dims dn*8 + 2*tq, +1) ----
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            float* mine = merge + mem * 68;
            if (gq == 0) {
                if (tq == 0) { mine[0] = m0; mine[1] = l0; }
#pragma unroll
                for (int dn = 0; dn < 8; ++dn) { mine[4 + dn * 8 + 2 * tq] = oacc[dn][0]; mine[4 + dn * 8 + 2 * tq + 1] = oacc[dn][1]; }
            }
            named_bar_sync(9 + grp, 96);
            if (mem == 0) {
                const float ma = merge[0], mb2 = merge[68], mc = merge[136];
                const float mm = fmaxf(ma, fmaxf(mb2, mc));
                const float wa = ex2_approx((ma - mm) * LOG2E), wb = ex2_approx((mb2 - mm) * LOG2E), wc = ex2_approx((mc - mm) * LOG2E);
                const float tot = merge[1] * wa + merge[69] * wb + merge[137] * wc;
                const float inv = 1.0f / tot;
                if (kLse && lane == 0) lse_out[static_cast<int64_t>(item) * N_TOK] = mm * LOG2E + log2f(tot);
                const float o0 = (merge[4 + 2 * lane] * wa + merge[72 + 2 * lane] * wb + merge[140 + 2 * lane] * wc) * inv;
                const float o1 = (merge[5 + 2 * lane] * wa + merge[73 + 2 * lane] * wb + merge[141 + 2 * lane] * wc) * inv;
                reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(s) * N_TOK) * E + h * 64)[lane] = pack_bf16x2(o0, o1);
            }
            named_bar_sync(9 + grp, 96);  // the merge buffer may be overwritten by the next item
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}

int launch_attention_tc257x16(const bf16* qkv, bf16* out, int BD, int heads, int num_sms, cudaStream_t stream, long long* dbg, float* lse_out) {
    using namespace atc16;
    const int E = heads * 64;
    TmaDesc m128, m16, mO;
    MST_PROPAGATE(make_tma_3d_bf16(&m128, qkv, 3 * E, N_TOK, BD, 3 * E, static_cast<uint64_t>(N_TOK) * 3 * E, 64, 128, true));
    MST_PROPAGATE(make_tma_3d_bf16(&m16, qkv, 3 * E, N_TOK, BD, 3 * E, static_cast<uint64_t>(N_TOK) * 3 * E, 64, 16, true));
    MST_PROPAGATE(make_tma_3d_bf16(&mO, out, E, N_TOK, BD, E, static_cast<uint64_t>(N_TOK) * E, 32, 32, false));
    static const int poly = exp_env("MST_ATTN_POLY", 7);  // experiments: 0 = all exponentials on MUFU
    // 7 of 16 pairs on the polynomial: measured optimum (kernel 0.547 ms with 0, 0.495 / 0.469 / 0.510 / 0.499 / 0.569 ms with
    // 6 / 7 / 8 / 10 / 12 of 16, config-2 shape, profiles/attn_timing.py)
    auto kern = lse_out != nullptr ? attention_tc257x16_kernel<7, true> : (poly == 0 ? attention_tc257x16_kernel<0> : attention_tc257x16_kernel<7>);
    MST_SET_DYN_SMEM(attention_tc257x16_kernel<0>, DYN_BYTES);
    MST_SET_DYN_SMEM(attention_tc257x16_kernel<7>, DYN_BYTES);
    {
        auto kl = attention_tc257x16_kernel<7, true>;
        MST_SET_DYN_SMEM(kl, DYN_BYTES);
    }
    const int items = BD * heads;
    const int grid = items < num_sms ? items : num_sms;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = DYN_BYTES; cfg.stream = stream;
    cudaLaunchAttribute pattr[1];
    pattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pattr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = pattr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    MST_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, m128, m16, mO, qkv, out, items, heads, dbg, lse_out));
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mst
