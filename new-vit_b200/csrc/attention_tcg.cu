// Encoder attention on tcgen05 for ANY token count 17 <= N <= 352 (head_dim 64):  config 4 of BASELINE.json runs ViT-B/14 at
// 252x252 = 325 tokens, registers add 4, other input sizes give other N; N == 257 has its own specialised kernel
// (attention_tc16.cu).        out[s, i, h*64:(h+1)*64] = softmax(q_i . K^T) V        (reference layers/attention.py:56-69)
//
// Persistent, one CTA per SM, 640 threads; items = (slice, head).  K and V of the item (all N tokens, zero-padded by TMA
// to KPAD = N rounded up to 16) are loaded once into a 2-stage shared-memory ring; the queries run in tiles of 128 rows:
//     S = Q K^T    SS MMAs, M128 x KPAD (one or two N chunks of <= 256) x K16 x 4, fp32 in TMEM columns [0, KPAD)
//     softmax      SIXTEEN warps: warp (q, cq) owns TMEM lanes 32q.. (query rows; a TMEM lane quadrant can only be read by
//                  warps with warp % 4 == q) and one of FOUR contiguous groups of 16-column chunks, so four warps share the rows
//                  and split the keys; row max and row sum are combined through shared memory; P (bf16 pairs) is written IN
//                  PLACE at the start of the warp's own column group, so no warp's P lands on scores another warp has yet to
//                  read.  Part of the exponentials run on the FMA / ALU pipes (softmax_math.cuh), as in the N = 257 kernel.
//     O = P V      TS MMAs (A = P from TMEM, V MN-major from shared memory), M128 N64 K16 x KPAD/16, TMEM columns [448, 512)
//     epilogue     the same 16 warps: O / l -> bf16 -> 32-byte row pieces
// One score buffer (S of 325 tokens takes 336 of the 512 TMEM columns, two do not fit): the tensor pipe idles while the softmax
// runs, so the softmax phase is what is made short -- the first version ran it on eight warps with every exponential on the MUFU
// unit and took 1.8x the time per score of the N = 257 kernel (config 4: 17.7 of 62 ms).  The N x N probabilities never touch HBM.
#include <math_constants.h>
#include "common.cuh"
#include "ptx.cuh"
#include "softmax_math.cuh"

namespace mst {

namespace atg {
constexpr int THREADS = 640;
constexpr int SW = 16;                    // softmax / epilogue warps
constexpr int kPoly = 7;                  // of every 16 column pairs, exponentials on the FMA / ALU pipes (the N = 257 kernel's optimum)
constexpr int Q_TILE_BYTES = 128 * 128;   // 128 queries x 64 dims bf16
constexpr int O_COL = 448;                // O accumulator columns [448, 512)
constexpr float LOG2E = 1.4426950408889634f;
}  // namespace atg

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void atg_tma_load_3d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// shared-memory layout (all offsets multiples of 1024): [K stage 0][V stage 0][K stage 1][V stage 1][Q 0][Q 1][stats][bars]
__global__ void __launch_bounds__(atg::THREADS, 1)
attention_tcg_kernel(const __grid_constant__ TmaDesc mapKV, const __grid_constant__ TmaDesc mapQ, bf16* __restrict__ out,
                     int num_items, int heads, int N, int KPAD, int kv_box, int kv_boxes) {
    using namespace atg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int kv_bytes = kv_box * kv_boxes * 128;                 // one of K / V for one item (multiple of 1024)
    uint8_t* sQ = smem + 4 * kv_bytes;
    float* stats = reinterpret_cast<float*>(sQ + 2 * Q_TILE_BYTES);   // [2 tile parities][max, sum][4 column groups][128 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stats + 2 * 2 * 4 * 128);
    uint64_t* kv_full = bars;          // [2]
    uint64_t* kv_empty = bars + 2;     // [2]
    uint64_t* q_full = bars + 4;       // [2]
    uint64_t* q_empty = bars + 6;      // [2]
    uint64_t* s_full = bars + 8;       // MMA -> softmax
    uint64_t* p_ready = bars + 9;      // softmax (16 warps) -> MMA
    uint64_t* o_full = bars + 10;      // MMA -> epilogue
    uint64_t* o_free = bars + 11;      // epilogue (16 warps) -> MMA
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int E = heads * 64;
    const int tiles_per_item = (N + 127) >> 7;
    const int my_items = static_cast<int>(blockIdx.x) < num_items ? (num_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&mapKV); tma_prefetch_desc(&mapQ); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
        }
        mbar_init(s_full, 1); mbar_init(p_ready, SW); mbar_init(o_full, 1); mbar_init(o_free, SW);
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
            int g = 0;
            for (int it = 0; it < my_items; ++it) {
                const int item = blockIdx.x + it * gridDim.x;
                const int s = item / heads, h = item % heads;
                const int st = it & 1;
                mbar_wait(&kv_empty[st], ((it >> 1) & 1) ^ 1);
                uint8_t* sK = smem + st * 2 * kv_bytes;
                uint8_t* sV = sK + kv_bytes;
                mbar_arrive_expect_tx(&kv_full[st], 2 * kv_bytes);
                for (int b = 0; b < kv_boxes; ++b) {   // rows >= N are zero-filled by the TMA unit
                    atg_tma_load_3d(sK + b * kv_box * 128, &mapKV, &kv_full[st], E + h * 64, b * kv_box, s);
                    atg_tma_load_3d(sV + b * kv_box * 128, &mapKV, &kv_full[st], 2 * E + h * 64, b * kv_box, s);
                }
                for (int t = 0; t < tiles_per_item; ++t, ++g) {
                    const int qs = g & 1;
                    mbar_wait(&q_empty[qs], ((g >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&q_full[qs], Q_TILE_BYTES);
                    atg_tma_load_3d(sQ + qs * Q_TILE_BYTES, &mapQ, &q_full[qs], h * 64, t * 128, s);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp convergent, one elected lane issues) =====================
        constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B | version 1 | SWIZZLE_128B
        const uint32_t idesc_pv = umma_idesc_bf16_f32(128, 64) | (1u << 16);   // B (V) is MN-major
        const int n0 = KPAD < 256 ? KPAD : 256, n1 = KPAD - n0;                 // S in one or two N chunks
        const uint32_t idesc_s0 = umma_idesc_bf16_f32(128, n0), idesc_s1 = umma_idesc_bf16_f32(128, n1 > 0 ? n1 : 16);
        const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFF) >> 4;
        const int nct = KPAD >> 4;                                              // 16-key steps
        const int cg_base = nct >> 2, cg_rem = nct & 3;                         // column group i: cg_base + (i < cg_rem) chunks
        int g = 0;
        for (int it = 0; it < my_items; ++it) {
            const int st = it & 1;
            const uint32_t k_lo = smem_lo + ((st * 2 * kv_bytes) >> 4);
            const uint32_t v_lo = k_lo + (kv_bytes >> 4);
            mbar_wait(&kv_full[st], (it >> 1) & 1);
            for (int t = 0; t < tiles_per_item; ++t, ++g) {
                const int qs = g & 1;
                mbar_wait(&q_full[qs], (g >> 1) & 1);
                tc_fence_after_sync();
                const uint32_t q_lo = smem_lo + ((4 * kv_bytes + qs * Q_TILE_BYTES) >> 4);
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_bf16_ss(tmem_base, make_desc(q_lo + 2 * k, kDescHi), make_desc(k_lo + 2 * k, kDescHi), idesc_s0, k != 0 ? 1u : 0u);
                        if (n1 > 0)
                            umma_bf16_ss(tmem_base + 256, make_desc(q_lo + 2 * k, kDescHi), make_desc(k_lo + (256 * 128 >> 4) + 2 * k, kDescHi),
                                         idesc_s1, k != 0 ? 1u : 0u);
                    }
                    umma_commit(s_full);
                    umma_commit(&q_empty[qs]);
                }
                __syncwarp();
                mbar_wait(p_ready, g & 1);
                if (g > 0) mbar_wait(o_free, (g - 1) & 1);
                tc_fence_after_sync();
                if (elect_one_sync()) {
                    int j = 0;
                    for (int i = 0; i < 4; ++i) {                                // P of column group i starts at its first S column
                        const int cs = i * cg_base + (i < cg_rem ? i : cg_rem), cn = cg_base + (i < cg_rem ? 1 : 0);
                        for (int k = 0; k < cn; ++k, ++j)
                            umma_bf16_ts(tmem_base + O_COL, tmem_base + 16 * cs + 8 * k, make_desc(v_lo + j * (2048 >> 4), kDescHi), idesc_pv,
                                         j != 0 ? 1u : 0u);
                    }
                    umma_commit(o_full);
                    if (t == tiles_per_item - 1) umma_commit(&kv_empty[st]);   // last tensor-core read of this K/V stage
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================== softmax + epilogue =====================
        const int e = warp - 4, q = e & 3, cq = e >> 2;
        const int nct = KPAD >> 4;
        const int cg_base = nct >> 2, cg_rem = nct & 3;
        const int c_begin = cq * cg_base + (cq < cg_rem ? cq : cg_rem), c_end = c_begin + cg_base + (cq < cg_rem ? 1 : 0);   // this warp's chunks
        const bool has_cols = c_end > c_begin;                                    // (tiny N: the last groups are empty)
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int row = q * 32 + lane;
        const uint32_t pbase = static_cast<uint32_t>(16 * c_begin);              // P of this group starts here (in place)
        int g = 0;
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int s = item / heads, h = item % heads;
            for (int t = 0; t < tiles_per_item; ++t, ++g) {
                float* st_max = stats + (g & 1) * 1024;         // [4 groups][128]
                float* st_sum = st_max + 512;
                mbar_wait(s_full, g & 1);
                tc_fence_after_sync();
                // ---- pass 1: row max over this warp's columns (keys >= N are padding); the next chunk's TMEM load is in
                //      flight while the current one is reduced ----
                float m = -CUDART_INF_F;
                uint32_t ra[16], rb[16];
                auto max16 = [&](const uint32_t (&r)[16], int c) {
                    if (16 * c + 16 <= N) {
                        float a = m, b2 = -CUDART_INF_F;
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            a = fmax3(a, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                            b2 = fmax3(b2, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                        }
                        m = fmaxf(a, b2);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (16 * c + i < N) m = fmaxf(m, __uint_as_float(r[i]));
                    }
                };
                if (has_cols) tmem_ld_32x32b_x16(lane_base + 16 * c_begin, ra);
                for (int c = c_begin; c < c_end; c += 2) {
                    tmem_ld_wait();
                    if (c + 1 < c_end) tmem_ld_32x32b_x16(lane_base + 16 * (c + 1), rb);
                    max16(ra, c);
                    if (c + 1 < c_end) {
                        tmem_ld_wait();
                        if (c + 2 < c_end) tmem_ld_32x32b_x16(lane_base + 16 * (c + 2), ra);
                        max16(rb, c + 1);
                    }
                }
                st_max[cq * 128 + row] = m;
                if (has_cols) tmem_ld_32x32b_x16(lane_base + 16 * c_begin, ra);   // first chunk of pass 2 rides over the exchange
                named_bar_sync(1 + q, 128);
                m = fmaxf(fmaxf(st_max[row], st_max[128 + row]), fmaxf(st_max[256 + row], st_max[384 + row]));   // finite: group 0 holds key 0
                const float mb = m * LOG2E;
                // ---- pass 2: p = 2^(s*log2e - mb) as bf16 pairs, in place at the start of this warp's group ----
                float sum = 0.f;
                auto exp16 = [&](const uint32_t (&r)[16], int c, int par) {
                    uint32_t o[8];
                    if (16 * c + 16 <= N) {   // whole chunk of real keys: kPoly of 16 pairs on the polynomial (pattern alternates per chunk)
                        sum += par ? softmax_math16<kPoly, 8>(r, o, mb) : softmax_math16<kPoly, 0>(r, o, mb);
                    } else {                  // the chunk that holds the padding keys: their probabilities are exactly 0
                        const f32x2 l2 = f2_pack(LOG2E, LOG2E), nmb = f2_pack(-mb, -mb);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float t0, t1;
                            f2_unpack(f2_fma(f2_pack(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), l2, nmb), t0, t1);
                            const float p0 = 16 * c + 2 * i < N ? ex2_approx(t0) : 0.f;
                            const float p1 = 16 * c + 2 * i + 1 < N ? ex2_approx(t1) : 0.f;
                            o[i] = pack_bf16x2(p0, p1);
                            sum += p0 + p1;
                        }
                    }
                    tmem_st_32x32b_x8(lane_base + pbase + 8 * (c - c_begin), o);
                };
                for (int c = c_begin; c < c_end; c += 2) {
                    tmem_ld_wait();
                    if (c + 1 < c_end) tmem_ld_32x32b_x16(lane_base + 16 * (c + 1), rb);
                    exp16(ra, c, 0);
                    if (c + 1 < c_end) {
                        tmem_ld_wait();
                        if (c + 2 < c_end) tmem_ld_32x32b_x16(lane_base + 16 * (c + 2), ra);
                        exp16(rb, c + 1, 1);
                    }
                }
                tmem_st_wait();
                tc_fence_before_sync();
                st_sum[cq * 128 + row] = sum;
                __syncwarp();
                if (lane == 0) mbar_arrive(p_ready);
                named_bar_sync(1 + q, 128);
                const float inv = 1.0f / ((st_sum[row] + st_sum[128 + row]) + (st_sum[256 + row] + st_sum[384 + row]));
                // ---- epilogue: O columns [448 + 16 cq, +16) of this warp's rows ----
                mbar_wait(o_full, g & 1);
                tc_fence_after_sync();
                uint32_t ro[16];
                tmem_ld_32x32b_x16(lane_base + O_COL + cq * 16, ro);
                tmem_ld_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_free);
                const int tok = t * 128 + row;
                if (tok < N) {
                    uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(s) * N + tok) * E + h * 64 + cq * 16);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        dst[i] = make_uint4(pack_bf16x2(__uint_as_float(ro[8 * i]) * inv, __uint_as_float(ro[8 * i + 1]) * inv),
                                            pack_bf16x2(__uint_as_float(ro[8 * i + 2]) * inv, __uint_as_float(ro[8 * i + 3]) * inv),
                                            pack_bf16x2(__uint_as_float(ro[8 * i + 4]) * inv, __uint_as_float(ro[8 * i + 5]) * inv),
                                            pack_bf16x2(__uint_as_float(ro[8 * i + 6]) * inv, __uint_as_float(ro[8 * i + 7]) * inv));
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}

struct AtgGeometry {
    int KPAD, kv_boxes, kv_box, kv_bytes;
    size_t smem;
};
static AtgGeometry atg_geometry(int N) {
    AtgGeometry g;
    g.KPAD = (N + 15) & ~15;
    g.kv_boxes = (g.KPAD + 127) / 128;
    g.kv_box = (((g.KPAD + g.kv_boxes - 1) / g.kv_boxes) + 7) & ~7;   // rows per TMA box: whole 8-row swizzle atoms, <= 128
    g.kv_bytes = g.kv_box * g.kv_boxes * 128;
    g.smem = static_cast<size_t>(4) * g.kv_bytes + 2 * atg::Q_TILE_BYTES + 2 * 2 * 4 * 128 * 4 + 12 * 8 + 16 + 1024;
    return g;
}
// two K/V stages must fit in shared memory (N <= 352) and S + O in the 512 TMEM columns
bool attention_tcg_supported(int N) {
    if (N < 17 || ((N + 15) & ~15) > atg::O_COL) return false;
    return atg_geometry(N).smem <= 232448;
}

int launch_attention_tcg(const bf16* qkv, bf16* out, int BD, int N, int heads, int num_sms, cudaStream_t stream) {
    using namespace atg;
    MST_REQUIRE(attention_tcg_supported(N), "attention_tcg: N=%d tokens unsupported", N);
    const int E = heads * 64;
    const AtgGeometry geo = atg_geometry(N);
    const int KPAD = geo.KPAD, kv_boxes = geo.kv_boxes, kv_box = geo.kv_box;
    const size_t smem = geo.smem;
    MST_REQUIRE(geo.kv_bytes % 1024 == 0, "attention_tcg: K/V stage must be a multiple of 1024 bytes");
    TmaDesc mKV, mQ;
    MST_PROPAGATE(make_tma_3d_bf16(&mKV, qkv, 3 * E, N, BD, 3 * E, static_cast<uint64_t>(N) * 3 * E, 64, kv_box, true));
    MST_PROPAGATE(make_tma_3d_bf16(&mQ, qkv, 3 * E, N, BD, 3 * E, static_cast<uint64_t>(N) * 3 * E, 64, 128, true));
    MST_SET_DYN_SMEM(attention_tcg_kernel, 232448);
    const int items = BD * heads;
    const int grid = items < num_sms ? items : num_sms;
    attention_tcg_kernel<<<grid, THREADS, smem, stream>>>(mKV, mQ, out, items, heads, N, KPAD, kv_box, kv_boxes);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mst
