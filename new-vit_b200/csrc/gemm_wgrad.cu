// Weight gradient of a linear layer on tcgen05, straight from the row-major activations:
//     dW[n, k] = sum_t dY[t, n] X[t, k]          db[n] = sum_t dY[t, n]
// (autograd of nn.Linear: attention.py:57/68, mlp.py:35-39 of the reference run under Lightning's backward).
//
// The reduction runs over TOKENS, the dimension both operands are strided in, so a K-major GEMM needs dY^T and X^T: the
// first version of the backward pass wrote both transposes to HBM for every linear (4.8 ms of a 33 ms step).  The tensor
// core can read the operands the way they lie instead: a TMA box of 64 tokens x 64 features (128 B per token row, 128B
// swizzle) IS the canonical MN-major SWIZZLE_128B operand atom, rows = reduction index.  Both descriptors are MN-major
// (instruction descriptor bits 15 and 16), the leading byte offset is the distance between 64-feature boxes, the stride
// byte offset the 1024 B between 8-token groups, and one K = 16 step advances the start address by two groups.
//
// One CTA owns a 128 x 192 block of dW and a range of 64-token chunks (split over tokens so that about one CTA per SM
// runs); the fp32 accumulator (192 TMEM columns) is added to dW with 16-byte vector reductions at the end.  The bias
// gradient rides along: in the CTAs of the first k-block the four epilogue warps, idle during the main loop, add up
// the dY tile of every stage from shared memory before the stage is released.
//
//   warp 0 (1 lane)  TMA producer: 2 + 3 boxes per stage (40 KB), 5 stages
//   warp 1           TMEM allocator; 1 lane issues 4 x tcgen05.mma (M128 N192 K16, both operands MN-major) per stage
//   warps 2..5       column sums of dY while the main loop runs (first k-block only), then the epilogue
#include "common.cuh"
#include "ptx.cuh"

namespace mst {

namespace wg {
constexpr int BM = 128;                       // output features (rows of dW) per CTA = MMA M
constexpr int BN = 192;                       // input features (columns of dW) per CTA = MMA N
constexpr int TCH = 64;                       // tokens per ring stage
constexpr int BOX_BYTES = TCH * 128;          // 8 KB: 64 tokens x 64 features
constexpr int A_BYTES = (BM / 64) * BOX_BYTES;
constexpr int B_BYTES = (BN / 64) * BOX_BYTES;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 40 KB
constexpr int STAGES = 5;
constexpr int PART_OFF = STAGES * STAGE_BYTES;   // [4 warps][128 features] partial column sums
constexpr int BAR_OFF = PART_OFF + 4 * BM * 4;
constexpr int NUM_BARS = 2 * STAGES + 1;         // full[S], empty[S], acc_full
constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16;
constexpr int DYN_BYTES = TOTAL + 1024;
constexpr int THREADS = 192;
constexpr uint32_t TMEM_COLS = 256;
static_assert(DYN_BYTES <= 232448, "shared memory budget");
}  // namespace wg

__global__ void __launch_bounds__(wg::THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ TmaDesc tmA, const __grid_constant__ TmaDesc tmB, float* __restrict__ dW, int ldw,
                float* __restrict__ db, int chunks_total, int chunks_per_split, int k_tiles) {
    using namespace wg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    float* part = reinterpret_cast<float*>(smem + PART_OFF);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x / k_tiles, k_tile = blockIdx.x % k_tiles;
    const int n0 = n_tile * BM, k0 = k_tile * BN;
    const int c0 = blockIdx.y * chunks_per_split;
    const int c1 = min(chunks_total, c0 + chunks_per_split);
    const bool colsum = db != nullptr && k_tile == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], colsum ? 5 : 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int c = c0; c < c1; ++c) {
                const int i = c - c0, s = i % STAGES;
                if (i >= STAGES) mbar_wait(&empty[s], ((i / STAGES) - 1) & 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
#pragma unroll
                for (int b = 0; b < BM / 64; ++b) tma_load_2d(st + b * BOX_BYTES, &tmA, &full[s], n0 + 64 * b, c * TCH);
#pragma unroll
                for (int b = 0; b < BN / 64; ++b) tma_load_2d(st + A_BYTES + b * BOX_BYTES, &tmB, &full[s], k0 + 64 * b, c * TCH);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32(BM, BN) | (1u << 15) | (1u << 16);   // A and B MN-major
            for (int c = c0; c < c1; ++c) {
                const int i = c - c0, s = i % STAGES;
                mbar_wait(&full[s], (i / STAGES) & 1);
                tc_fence_after_sync();
                const uint32_t a_u = smem_u32(smem + s * STAGE_BYTES), b_u = a_u + A_BYTES;
#pragma unroll
                for (int ks = 0; ks < TCH / 16; ++ks)
                    umma_bf16_ss(tmem, umma_desc_sw128_mnmajor(a_u + ks * 2048, BOX_BYTES), umma_desc_sw128_mnmajor(b_u + ks * 2048, BOX_BYTES),
                                 idesc, (i | ks) != 0 ? 1u : 0u);
                umma_commit(&empty[s]);
            }
            umma_commit(acc_full);
        }
    } else {
        const int ew = warp - 2;             // 0..3
        if (colsum) {
            // warp ew adds token rows 16 ew .. 16 ew + 15 of every stage; lane -> features 4 lane .. 4 lane + 3
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            const int blk = lane >> 4, ch = (lane & 15) >> 1, sub = (lane & 1) * 8;
            for (int c = c0; c < c1; ++c) {
                const int i = c - c0, s = i % STAGES;
                mbar_wait(&full[s], (i / STAGES) & 1);
                const uint8_t* st = smem + s * STAGE_BYTES + blk * BOX_BYTES;
#pragma unroll
                for (int rr = 0; rr < 16; ++rr) {
                    const int r = ew * 16 + rr;
                    const uint2 u = *reinterpret_cast<const uint2*>(st + r * 128 + ((ch ^ (r & 7)) << 4) + sub);
                    const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
                    a0 += lo.x; a1 += lo.y; a2 += hi.x; a3 += hi.y;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
            *reinterpret_cast<float4*>(part + ew * BM + 4 * lane) = make_float4(a0, a1, a2, a3);
            named_bar_sync(1, 128);
            const int f = threadIdx.x - 64;
            atomicAdd(db + n0 + f, part[f] + part[BM + f] + part[2 * BM + f] + part[3 * BM + f]);
        }
        // epilogue: lane quadrant warp % 4 -> rows n0 + 32 q + lane of dW, 192 columns
        mbar_wait(acc_full, 0);
        tc_fence_after_sync();
        const int q = warp & 3;
        float* row = dW + static_cast<int64_t>(n0 + 32 * q + lane) * ldw + k0;
#pragma unroll 1
        for (int cb = 0; cb < BN / 32; ++cb) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(32 * q) << 16) + cb * 32, r);
            tmem_ld_wait();
            float* pf = row + cb * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(pf + 4 * j), "f"(__uint_as_float(r[4 * j])),
                             "f"(__uint_as_float(r[4 * j + 1])), "f"(__uint_as_float(r[4 * j + 2])), "f"(__uint_as_float(r[4 * j + 3]))
                             : "memory");
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem);
}

bool wgrad_tc_supported(int Nout, int Kin) { return Nout % wg::BM == 0 && Kin % wg::BN == 0; }

// dW [Nout, Kin] fp32 and db [Nout] fp32 (nullable) are OVERWRITTEN.  dY [M, Nout], X [M, Kin] bf16 row-major.
int launch_wgrad_tc(const bf16* dY, int Nout, const bf16* X, int Kin, int M, float* dW, float* db, int num_sms, cudaStream_t stream) {
    using namespace wg;
    MST_REQUIRE(wgrad_tc_supported(Nout, Kin), "wgrad_tc: Nout %d must be a multiple of %d and Kin %d of %d", Nout, BM, Kin, BN);
    MST_REQUIRE(M >= 1, "wgrad_tc: no tokens");
    TmaDesc tmA, tmB;
    MST_PROPAGATE(make_tma_2d_bf16(&tmA, dY, Nout, M, Nout, 64, TCH, true, false));
    MST_PROPAGATE(make_tma_2d_bf16(&tmB, X, Kin, M, Kin, 64, TCH, true, false));
    const int chunks = (M + TCH - 1) / TCH, k_tiles = Kin / BN, tiles = (Nout / BM) * k_tiles;
    int splits = num_sms / tiles;
    splits = splits < 1 ? 1 : (splits > chunks ? chunks : splits);
    const int cps = (chunks + splits - 1) / splits;
    splits = (chunks + cps - 1) / cps;          // no empty token range
    MST_CHECK_CUDA(cudaMemsetAsync(dW, 0, static_cast<size_t>(Nout) * Kin * 4, stream));
    if (db) MST_CHECK_CUDA(cudaMemsetAsync(db, 0, static_cast<size_t>(Nout) * 4, stream));
    MST_SET_DYN_SMEM(wgrad_tc_kernel, DYN_BYTES);
    wgrad_tc_kernel<<<dim3(tiles, splits), THREADS, DYN_BYTES, stream>>>(tmA, tmB, dW, Kin, db, chunks, cps, k_tiles);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mst
