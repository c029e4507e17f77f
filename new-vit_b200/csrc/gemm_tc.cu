// bf16 GEMM on the 5th-gen tensor cores:  acc[M,N] = A[M,K] . W[N,K]^T,  fused epilogues (common.cuh).
//
// One persistent CTA per SM, warp-specialised (384 threads):
//   warp 0 (1 lane)  TMA producer   : A k-chunks (128 x 64 bf16, 128B swizzle) through a STAGES-deep ring
//   warp 1 (1 lane)  MMA issuer     : tcgen05.mma cta_group::1, M=128, N=BN, K=16 per instruction
//   warp 2           TMEM allocator : 2 accumulator stages of BN fp32 columns (epilogue overlaps next tile)
//   warps 4..11      epilogue       : warp e reads TMEM lane quadrant e%4, column half e/4 (tcgen05.ld 32x32b, the next
//                                     chunk's load in flight while the current one is processed), adds bias (smem
//                                     cache) / GELU / residual, packs bf16, transposes each 32x32 block through a
//                                     private swizzled smem tile and writes 8 rows x 64 B per store instruction; the
//                                     in-place residual update is a 16-byte vector reduction (REDG.ADD.BF16x8), so
//                                     the SM never reads the residual.
//
// Weight-resident mode (KCH > 0): K <= 384, so the whole [BN x K] weight slab (<= 144 KB) is loaded into
// shared memory ONCE per CTA and only A streams; each CTA owns one n-block and walks m-blocks.  This cuts
// the L2->SM traffic per 128x192x384 tile from 240 KB to 96 KB, which is what bounds a K=384 GEMM here
// (DESIGN.md "GEMM roofline").  Streaming mode (KCH == 0) rings both operands (fc2, K = 1536).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"

namespace mst {

// ---------------------------------------------------------------------------------------------------
// TMA descriptor creation (driver entry point resolved at run time; the library does not link libcuda)
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

int tma_init() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MST_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    MST_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available (driver too old?)");
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    return 0;
}

int make_tma_2d_bf16(TmaDesc* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                     uint32_t box_inner, uint32_t box_rows, bool swizzle128, bool swizzle64) {
    MST_PROPAGATE(tma_init());
    static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "CUtensorMap size");
    MST_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    MST_REQUIRE((row_stride_elems * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
    MST_REQUIRE((!swizzle128 || box_inner * 2 == 128) && (box_inner * 2) % 16 == 0 && box_rows <= 256, "bad TMA box");
    cuuint64_t gdim[2] = {inner, rows};
    cuuint64_t gstride[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                          const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MST_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu rows=%llu)", (int)r,
                (unsigned long long)inner, (unsigned long long)rows);
    return 0;
}

// 3D view [d2][d1][d0] (d0 contiguous), box {box0, box1, 1}; out-of-range rows read as zero / are not written
int make_tma_3d_bf16(TmaDesc* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                     uint64_t stride2_elems, uint32_t box0, uint32_t box1, bool swizzle128) {
    MST_PROPAGATE(tma_init());
    MST_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    MST_REQUIRE((stride1_elems * 2) % 16 == 0 && (stride2_elems * 2) % 16 == 0, "TMA strides must be multiples of 16 bytes");
    MST_REQUIRE((!swizzle128 || box0 * 2 == 128) && box1 <= 256, "bad TMA box");
    cuuint64_t gdim[3] = {d0, d1, d2};
    cuuint64_t gstride[2] = {stride1_elems * 2, stride2_elems * 2};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                          const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MST_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3D) failed with CUresult %d", (int)r);
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
// epilogue warps EW (8 or 12): 4 lane quadrants x EW/4 column groups of BN/(EW/4) columns each
constexpr int gemm_threads(int EW) { return 128 + EW * 32; }
constexpr int STG_TILE = 32 * 32 * 2;   // staging tile: 32 rows x 32 bf16


// PAIR: 0 = one CTA per tile; 1 = CTA pair sharing A by TMA multicast (adjacent n-blocks); 2 = CTA pair driving ONE
// 256 x BN tcgen05.mma.cta_group::2 (each CTA holds its 128 rows of A and BN/2 rows of the weight tile)
template <int BN, int KCH, int STAGES, int NSTG, int EW, int PAIR = 0>
struct GemmSmem {
    static constexpr int BNL = PAIR == 2 ? BN / 2 : BN;  // weight rows held in THIS CTA's shared memory
    // per-warp staging: NSTG 1/2 = that many 32x32 tiles (one TMA store per 32-column chunk; 2: one store in flight while
    // the next tile fills); NSTG 3 = the warp's whole 32 x BN/2 region (ONE async-proxy fence + ONE TMA store per tile)
    static constexpr int EPI_WARPS = EW;
    static_assert(NSTG != 3 || BN / (EW / 4) == 64, "whole-region staging needs 64-column groups");
    static constexpr int STG_BYTES = NSTG == 3 ? 32 * 64 * 2 : NSTG * STG_TILE;
    static constexpr int B_TILE_BYTES = BNL * BK * 2;
    static constexpr int B_BUFS = KCH > 0 ? KCH : STAGES;
    static constexpr int BIAS_FLOATS = KCH > 0 ? BN : 2048;  // resident: this CTA's n-block; streaming: the whole vector
    static constexpr int A_OFF = 0;
    static constexpr int B_OFF = STAGES * A_STAGE_BYTES;
    static constexpr int STG_OFF = B_OFF + B_BUFS * B_TILE_BYTES;
    static constexpr int BIAS_OFF = STG_OFF + EPI_WARPS * STG_BYTES;
    // row-statistics exchange of the fused residual + LayerNorm-statistics epilogue (streaming pair mode only):
    // [2 parities][2 column groups][128 rows] x (sum, sum of squares)
    static constexpr int STAT_BYTES = (KCH == 0 && PAIR == 2 && EW == 8) ? 2 * 2 * 128 * 2 * 4 : 0;
    static constexpr int STAT_OFF = BIAS_OFF + BIAS_FLOATS * 4;
    static constexpr int BAR_OFF = STAT_OFF + STAT_BYTES;
    static constexpr int NUM_BARS = 2 * STAGES + 1 + 4;
    static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-byte alignment
    static_assert(DYN_BYTES <= 232448, "shared memory budget (227 KB)");
};

// acc + bias (activation, residual) for 32 consecutive columns of one row -> 16 packed bf16x2.
// `bias` points to the 32 bias values of this chunk (shared memory cache or global).
template <int MODE>
__device__ __forceinline__ void epilogue_math(const uint32_t (&r)[32], uint32_t (&o)[16], const EpiParams& ep,
                                              const float* bias, bool bias_smem, int64_t row, bool row_ok, int N, int n0,
                                              float rstat) {
    constexpr bool LNF = MODE == EPI_LN_BIAS || MODE == EPI_LN_BIAS_GELU;
    // packed fp32 pairs throughout: one FADD2 / FFMA2 / FMUL2 per two columns (the GELU epilogue was issue-bound)
    f32x2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = f2_pack(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
    if (MODE == EPI_PATCH) {
        const int p = static_cast<int>(row % ep.P);
        const float4* pb = reinterpret_cast<const float4*>(ep.posb + static_cast<int64_t>(p) * N + n0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(pb + i);
            v[2 * i] = f2_add(v[2 * i], f2_pack(b.x, b.y));
            v[2 * i + 1] = f2_add(v[2 * i + 1], f2_pack(b.z, b.w));
        }
    } else if (LNF) {
        // LN(x) W^T + b = rstd * (x . Wc^T) + bias'[n]: the weight rows are centred (sum_k Wc[n,k] = 0), which subtracts the
        // row mean inside the MMA; one FFMA2 per column pair, same issue count as the plain bias add
        const f32x2 rs = f2_pack(rstat, rstat);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            f32x2 b0, b1;
            if (bias_smem) {
                // `asm volatile`: an explicit LDS that stays behind the barrier after the bias-cache fill (a NON-volatile asm load
                // is a pure function to the compiler and may be hoisted above it; a plain C++ load through the generic pointer
                // cost qkv 12 % in-step)
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(b0), "=l"(b1) : "r"(smem_u32(bias) + i * 16));
            } else {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + i);
                b0 = f2_pack(b.x, b.y); b1 = f2_pack(b.z, b.w);
            }
            v[2 * i] = f2_fma(v[2 * i], rs, b0);
            v[2 * i + 1] = f2_fma(v[2 * i + 1], rs, b1);
        }
    } else {
        if (bias_smem) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                f32x2 b0, b1;
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(b0), "=l"(b1) : "r"(smem_u32(bias) + i * 16));
                v[2 * i] = f2_add(v[2 * i], b0);
                v[2 * i + 1] = f2_add(v[2 * i + 1], b1);
            }
        } else {
            const float4* pb = reinterpret_cast<const float4*>(bias);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 b = __ldg(pb + i);
                v[2 * i] = f2_add(v[2 * i], f2_pack(b.x, b.y));
                v[2 * i + 1] = f2_add(v[2 * i + 1], f2_pack(b.z, b.w));
            }
        }
    }
    if (MODE == EPI_BIAS_GELU || MODE == EPI_LN_BIAS_GELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = gelu_tanh_fit2(v[i]);
    }
    if (MODE == EPI_BIAS_RES) {
        if (row_ok) {
            const uint4* pr = reinterpret_cast<const uint4*>(static_cast<const bf16*>(ep.res) + row * ep.ldr + n0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 u = pr[i];   // plain load: with the fused statistics the same buffer is written later in this kernel
                float2 f;
                f = unpack_bf16x2(u.x); v[4 * i + 0] = f2_add(v[4 * i + 0], f2_pack(f.x, f.y));
                f = unpack_bf16x2(u.y); v[4 * i + 1] = f2_add(v[4 * i + 1], f2_pack(f.x, f.y));
                f = unpack_bf16x2(u.z); v[4 * i + 2] = f2_add(v[4 * i + 2], f2_pack(f.x, f.y));
                f = unpack_bf16x2(u.w); v[4 * i + 3] = f2_add(v[4 * i + 3], f2_pack(f.x, f.y));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float lo, hi;
        f2_unpack(v[i], lo, hi);
        o[i] = pack_bf16x2(lo, hi);
    }
}

__device__ __forceinline__ void red_add_bf16x8(void* gptr, const uint4& v) {
    asm volatile("red.global.v4.bf16x2.add.noftz [%0], {%1, %2, %3, %4};" ::"l"(gptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// One 32x32 bf16 block (lane = row, o = its 32 columns) -> global, coalesced: the block is transposed through a
// per-warp, XOR-swizzled shared-memory tile so that every store instruction writes 8 rows x 64 contiguous bytes
// (8 LSU wavefronts instead of 32).  ACCUM: 16-byte vector reductions (REDG.ADD.BF16x8) instead of stores.
__device__ __forceinline__ void store_block_32x32(uint8_t* stg, const uint32_t (&o)[16], int lane, int mode, const EpiParams& ep,
                                                  int row0, int M, int n0) {
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint4*>(stg + lane * 64 + ((i ^ sw) << 4)) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    __syncwarp();
    const int pc = lane & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = k * 8 + (lane >> 2);
        const uint4 u = *reinterpret_cast<const uint4*>(stg + r * 64 + ((pc ^ ((r >> 1) & 3)) << 4));
        const int64_t row = static_cast<int64_t>(row0) + r;
        if (row < M) {
            int64_t orow = row;
            if (mode == EPI_PATCH) orow = (row / ep.P) * (ep.P + 1 + ep.R) + 1 + ep.R + (row % ep.P);
            bf16* dst = static_cast<bf16*>(ep.out) + orow * ep.ldo + n0 + pc * 8;
            if (mode == EPI_BIAS_ACCUM) red_add_bf16x8(dst, u);
            else *reinterpret_cast<uint4*>(dst) = u;
        }
    }
    __syncwarp();
}

template <int BN, int KCH, int STAGES, int NSTG, int EW, int PAIR>
__global__ void __launch_bounds__(gemm_threads(EW), 1)
gemm_tc_kernel(const __grid_constant__ TmaDesc tmA, const __grid_constant__ TmaDesc tmB, const __grid_constant__ TmaDesc tmC,
               int M, int N, int K, int mode_flags, EpiParams ep) {
    using L = GemmSmem<BN, KCH, STAGES, NSTG, EW, PAIR>;
    constexpr int MC = PAIR == 1 ? 2 : (PAIR == 4 ? 4 : 1);  // CTAs sharing one A stage by TMA multicast
    constexpr bool MCAST = MC > 1;
    constexpr bool TWO = PAIR == 2;
    constexpr int BMT = TWO ? 2 * BM : BM;  // rows of one tile (the pair's 256-row tile in cta_group::2 mode)
    const int mode = mode_flags & 0xff;
    const int xp_flags = kDbgTiming ? mode_flags : 0;   // experiment bits 0x100 / 0x200 (mainloop / epilogue ceilings): compiled out of the product
    constexpr bool kResident = KCH > 0;
    constexpr uint32_t kTmemCols = (2 * BN <= 256) ? 256 : 512;
    constexpr int EPI_WARPS = L::EPI_WARPS;
    constexpr int COLS_PER_WARP = BN / (EW / 4);
    constexpr int CHUNKS_PER_WARP = COLS_PER_WARP / 32;  // 32-column chunks per epilogue warp and tile
    static_assert(2 * BN <= 512 && (CHUNKS_PER_WARP == 2 || CHUNKS_PER_WARP == 3), "tile shape");
    static_assert(!MCAST || KCH > 0, "multicast is wired for the weight-resident schedule");
    // MCAST: the CTA pair (2j, 2j+1) of a cluster walks the same m-blocks with adjacent n-blocks; each CTA fetches half
    // of every A stage and TMA-multicasts it to both, halving the bytes each SM must keep in flight per tile.
    const uint32_t cta_rank = PAIR != 0 ? cluster_ctarank() : 0;
    const bool leader = cta_rank == 0;
    constexpr int kPrefetchTiles = 2;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sA = smem + L::A_OFF;
    uint8_t* sB = smem + L::B_OFF;
    float* sBias = reinterpret_cast<float*>(smem + L::BIAS_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* bfull_bar = bars + 2 * STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES + 1;
    uint64_t* tempty_bar = bars + 2 * STAGES + 3;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tiles = N / BN;
    const int m_tiles = (M + BMT - 1) / BMT;
    const int kchunks = K / BK;
    // split-K (weight gradients: few output tiles, a contraction over all tokens): mode_flags >> 16 units share one output tile, each
    // takes a contiguous range of k-chunks and adds its partial product to the (zero-initialised) fp32 output
    const int ksplit = kResident ? 1 : max(1, mode_flags >> 16);
    const int kper = (kchunks + ksplit - 1) / ksplit;

    // tile sequence of this CTA (cta_group::2: of this CTA pair -- both CTAs walk the same tiles, 128 rows each)
    const int unit = TWO ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int units = TWO ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int row_off = TWO ? static_cast<int>(cta_rank) * BM : 0;  // this CTA's rows inside a tile
    int t_first, t_step, t_count;
    if (kResident) {  // one n-block per CTA; m-blocks strided by the group size
        const int gs = units / n_tiles;
        const int m0 = unit / n_tiles;
        t_first = m0; t_step = gs;
        t_count = m0 < m_tiles ? (m_tiles - m0 + gs - 1) / gs : 0;
    } else if (mode_flags & 0x400) {
        // fused row statistics: a unit takes WHOLE m-blocks, its n-blocks back to back, so one CTA sees every column of its rows
        t_first = unit; t_step = units;
        t_count = unit < m_tiles ? ((m_tiles - unit + units - 1) / units) * n_tiles : 0;
    } else {
        const int total = m_tiles * n_tiles * ksplit;
        t_first = unit; t_step = units;
        t_count = t_first < total ? (total - t_first + t_step - 1) / t_step : 0;
    }
    const bool seqn = !kResident && (mode_flags & 0x400) != 0;
    const int n_fixed = unit % n_tiles;

    griddep_launch_dependents();   // the next kernel's prologue may overlap this one's tail (it waits before touching data)
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], MC); }
        mbar_init(bfull_bar, 1);
        // cta_group::2: the leader's accumulator-free barrier collects the epilogue warps of BOTH CTAs
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], TWO ? 2 * EPI_WARPS : EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (TWO) tmem_alloc_pair<kTmemCols>(tmem_ptr_smem); else tmem_alloc<kTmemCols>(tmem_ptr_smem);
    }
    // bias cache (EPI_PATCH has no bias vector)
    const bool ln_fold = mode == EPI_LN_BIAS || mode == EPI_LN_BIAS_GELU;
    const bool bias_cached = mode != EPI_PATCH && mode != EPI_RAW_F32 && (kResident || N <= L::BIAS_FLOATS);
    if (bias_cached && warp >= 4) {
        const int cnt = kResident ? BN : N;
        const float* src = ep.bias + (kResident ? n_fixed * BN : 0);
        for (int i = threadIdx.x - 128; i < cnt; i += EPI_WARPS * 32) sBias[i] = __ldg(src + i);
    }
    tc_fence_before_sync();
    if (PAIR != 0) cluster_sync_all(); else __syncthreads();  // peers' barriers must be initialised before any remote arrive
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // Everything above reads weights only.  The producer lane also fetches its resident weight block before it waits for
    // the kernel in front (whose output is A, the row statistics and, in place, the residual).
    if (warp != 0 || lane != 0) griddep_wait();

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            // cta_group::2: every operand byte of BOTH CTAs completes on the LEADER's barriers (its MMA warp consumes them)
            const uint32_t bfull_leader = TWO ? mapa_u32(bfull_bar, 0) : 0;
            if (kResident && t_count > 0) {
                if (!TWO || leader) mbar_arrive_expect_tx(bfull_bar, (TWO ? 2 : 1) * KCH * L::B_TILE_BYTES);
                for (int kc = 0; kc < KCH; ++kc) {
                    if (TWO)
                        tma_load_2d_pair(sB + kc * L::B_TILE_BYTES, &tmB, bfull_leader, kc * BK,
                                         n_fixed * BN + static_cast<int>(cta_rank) * L::BNL);
                    else
                        tma_load_2d(sB + kc * L::B_TILE_BYTES, &tmB, bfull_bar, kc * BK, n_fixed * BN);
                }
            }
            griddep_wait();
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < t_count; ++it) {
                const int t = t_first + it * t_step;
                const int tt = t / ksplit, kpart = t - tt * ksplit;
                const int m_blk = kResident ? t : (seqn ? t_first + (it / n_tiles) * t_step : tt / n_tiles);
                const int n_blk = kResident ? n_fixed : (seqn ? it % n_tiles : tt % n_tiles);
                const int kc0 = kpart * kper, kc1 = min(kchunks, kc0 + kper);
                // (streaming mode: prefetching there made fc2 13 % slower -- every CTA of an m-block row would issue the same
                //  prefetches -- so it is limited to the resident schedule, where one cluster per m-group issues them)
                if (kResident && n_fixed < MC && it + kPrefetchTiles < t_count) {
                    // pull the A rows of a later tile into L2 now: under load a TMA load that misses L2 takes ~3000 clk,
                    // far more than the 4 x 384 clk of MMA work the operand ring can cover (profiles/gemm_timing.py)
                    const int m_pf = t + kPrefetchTiles * t_step;
                    for (int kc = 0; kc < kchunks; ++kc)
                        tma_prefetch_l2_2d(&tmA, kc * BK, m_pf * BMT + row_off + (MCAST ? static_cast<int>(cta_rank) * (BM / MC) : 0));
                }
                for (int kc = kc0; kc < kc1; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (TWO) {
                        const uint32_t full_leader = mapa_u32(&full_bar[stage], 0);
                        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + (kResident ? 0 : L::B_TILE_BYTES)));
                        tma_load_2d_pair(sA + stage * A_STAGE_BYTES, &tmA, full_leader, kc * BK, m_blk * BMT + row_off);
                        if (!kResident)
                            tma_load_2d_pair(sB + stage * L::B_TILE_BYTES, &tmB, full_leader, kc * BK,
                                             n_blk * BN + static_cast<int>(cta_rank) * L::BNL);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full_bar[stage], A_STAGE_BYTES + (kResident ? 0 : L::B_TILE_BYTES));
                    if (MCAST)
                        tma_load_2d_multicast(sA + stage * A_STAGE_BYTES + cta_rank * (A_STAGE_BYTES / MC), &tmA, &full_bar[stage],
                                              kc * BK, m_blk * BM + static_cast<int>(cta_rank) * (BM / MC), (1u << MC) - 1);
                    else
                        tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], kc * BK, m_blk * BM);
                    if (!kResident)
                        tma_load_2d(sB + stage * L::B_TILE_BYTES, &tmB, &full_bar[stage], kc * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1 && (!TWO || leader)) {
        // ===================== MMA issuer (cta_group::2: the leader CTA's warp drives both SMs) =====================
        // The whole warp runs this loop convergently (operands stay in uniform registers; a single-lane loop made
        // ptxas route every descriptor through ELECT/R2UR and the issue loop, not the tensor pipe, set the pace);
        // one elected lane issues.  Descriptors are a constant high word + (smem address >> 4) in the low word.
        constexpr uint32_t idesc = umma_idesc_bf16_f32(BMT, BN);
        constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B | version 1 | SWIZZLE_128B
        const uint32_t a_lo0 = (smem_u32(sA) & 0x3FFFF) >> 4;
        const uint32_t b_lo0 = (smem_u32(sB) & 0x3FFFF) >> 4;
        if (kResident && t_count > 0) { mbar_wait(bfull_bar, 0); tc_fence_after_sync(); }
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        long long gd[2] = {0, 0};
        const long long gstart = MST_DBG_CLOCK();
        for (int it = 0; it < t_count; ++it) {
            const long long g0 = MST_DBG_CLOCK();
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            gd[0] += MST_DBG_CLOCK() - g0;
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
            const int kpart_m = (t_first + it * t_step) % ksplit;
            const int kc0 = kpart_m * kper, kc1 = min(kchunks, kc0 + kper);
            for (int kc = kc0; kc < kc1; ++kc) {
                const long long g1 = MST_DBG_CLOCK();
                mbar_wait(&full_bar[stage], phase);
                gd[1] += MST_DBG_CLOCK() - g1;
                tc_fence_after_sync();
                const uint32_t a_lo = a_lo0 + stage * (A_STAGE_BYTES >> 4);
                const uint32_t b_lo = b_lo0 + (kResident ? kc : stage) * (L::B_TILE_BYTES >> 4);
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        if (TWO)
                            umma_bf16_ss_pair(d_tmem, make_desc(a_lo + 2 * k, kDescHi), make_desc(b_lo + 2 * k, kDescHi), idesc,
                                              (kc != kc0 || k != 0) ? 1u : 0u);
                        else
                            umma_bf16_ss(d_tmem, make_desc(a_lo + 2 * k, kDescHi), make_desc(b_lo + 2 * k, kDescHi), idesc,
                                         (kc != kc0 || k != 0) ? 1u : 0u);
                    }
                    if (TWO) {
                        umma_commit_pair(&empty_bar[stage], 0x3);                          // both CTAs refill their slot
                        if (kc == kc1 - 1) umma_commit_pair(&tfull_bar[acc], 0x3);          // both epilogues read their half
                    } else {
                        if (MCAST) umma_commit_multicast(&empty_bar[stage], (1u << MC) - 1);  // every CTA of the cluster refills this slot
                        else umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
                        if (kc == kc1 - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            acc ^= 1; if (acc == 0) acc_phase ^= 1;
        }
        if (kDbgTiming && ep.dbg != nullptr && blockIdx.x == 0 && lane == 0) {
            ep.dbg[0] = gd[0]; ep.dbg[1] = gd[1]; ep.dbg[2] = MST_DBG_CLOCK() - gstart; ep.dbg[3] = t_count;
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int e = warp - 4;
        const int q = e & 3;    // == warp % 4: the TMEM lane quadrant this warp may access
        const int hf = e >> 2;  // column group (64 columns)
        uint8_t* stg = smem + L::STG_OFF + e * L::STG_BYTES;
        int stg_sel = 0;
        int acc = 0; uint32_t acc_phase = 0;
        float st1 = 0.f, st2 = 0.f;   // fused row statistics: sum and sum of squares of this lane's row over this warp's columns
        const bool stats_on = (L::STAT_BYTES > 0 && seqn) || (kResident && ep.rowpart_out != nullptr);
        for (int it = 0; it < t_count; ++it) {
            const int t = t_first + it * t_step;
            const int tt = t / ksplit;
            const int m_blk = kResident ? t : (seqn ? t_first + (it / n_tiles) * t_step : tt / n_tiles);
            const int n_blk = kResident ? n_fixed : (seqn ? it % n_tiles : tt % n_tiles);
            const int row0 = m_blk * BMT + row_off + q * 32;
            const int64_t row = static_cast<int64_t>(row0) + lane;
            const bool row_ok = row < M;
            const int col0 = hf * COLS_PER_WARP;                  // first column of this warp inside the tile
            const int nbase = n_blk * BN + col0;                  // ... and in the output
            const float* bias0 = bias_cached ? sBias + (kResident ? col0 : nbase)
                                             : ((mode == EPI_PATCH || mode == EPI_RAW_F32) ? nullptr : ep.bias + nbase);
            float rstat = 0.f;
            if (ln_fold && row_ok) rstat = __ldg(ep.rowstat + row);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + col0);

            const bool dbg_w = kDbgTiming && ep.dbg != nullptr && blockIdx.x == 0 && e == 0 && lane == 0;
            auto process = [&](const uint32_t (&r)[32], int c) {
                if (mode == EPI_RAW_F32) {   // weight gradients: the fp32 accumulators as they are (small outputs: per-lane row accesses)
                    if (row_ok) {
                        float* pf = static_cast<float*>(ep.out) + row * ep.ldo + nbase + c * 32;
                        if (ksplit > 1) {        // split-K partial: fp32 vector reductions into the zero-initialised output
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(pf + 4 * i), "f"(__uint_as_float(r[4 * i])),
                                             "f"(__uint_as_float(r[4 * i + 1])), "f"(__uint_as_float(r[4 * i + 2])), "f"(__uint_as_float(r[4 * i + 3]))
                                             : "memory");
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                reinterpret_cast<float4*>(pf)[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                                               __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
                        }
                    }
                    return;
                }
                uint32_t o[16];
                const long long p0 = MST_DBG_CLOCK();
                const int n0 = nbase + c * 32;
                const float* b = bias0 + c * 32;
                switch (mode) {
                    case EPI_BIAS:
                    case EPI_BIAS_ACCUM: epilogue_math<EPI_BIAS>(r, o, ep, b, bias_cached, row, row_ok, N, n0, rstat); break;
                    case EPI_BIAS_GELU: epilogue_math<EPI_BIAS_GELU>(r, o, ep, b, bias_cached, row, row_ok, N, n0, rstat); break;
                    case EPI_BIAS_RES: epilogue_math<EPI_BIAS_RES>(r, o, ep, b, bias_cached, row, row_ok, N, n0, rstat); break;
                    case EPI_LN_BIAS: epilogue_math<EPI_LN_BIAS>(r, o, ep, b, bias_cached, row, row_ok, N, n0, rstat); break;
                    case EPI_LN_BIAS_GELU: epilogue_math<EPI_LN_BIAS_GELU>(r, o, ep, b, bias_cached, row, row_ok, N, n0, rstat); break;
                    default: if (NSTG != 2) epilogue_math<EPI_PATCH>(r, o, ep, b, false, row, row_ok, N, n0, rstat); break;
                }
                if (stats_on && row_ok) {   // statistics of the bf16 values that are written
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float2 f = unpack_bf16x2(o[i]);
                        st1 += f.x + f.y;
                        st2 = fmaf(f.x, f.x, st2);
                        st2 = fmaf(f.y, f.y, st2);
                    }
                }
                if (mode == EPI_PATCH) {
                    if (NSTG == 2) {
                        // The fp32 accumulators are transposed through this warp's 4 KB staging area so that BOTH the position
                        // table read (fp32 [P, N]) and the re-mapped row store (one CLS row inserted per slice) are coalesced:
                        // 4 lanes cover 32 consecutive columns of one row, 8 rows per instruction.  (Per-lane row access -- 32
                        // scattered 16-byte pieces per instruction -- made this GEMM epilogue-bound at 12 % tensor activity.)
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            *reinterpret_cast<uint4*>(stg + lane * 128 + ((i ^ (lane & 7)) << 4)) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
                        __syncwarp();
                        const int cc = lane & 3;
                        // slice / patch index of the block's first row, once (32-bit; M < 2^31): rows of a 32-row block span at
                        // most two slices, so the per-row indices follow without a division
                        const int s_blk = row0 / ep.P, p_blk = row0 - s_blk * ep.P;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int rr = k * 8 + (lane >> 2);
                            const int grow = row0 + rr;
                            const float4 a0 = *reinterpret_cast<const float4*>(stg + rr * 128 + (((2 * cc) ^ (rr & 7)) << 4));
                            const float4 a1 = *reinterpret_cast<const float4*>(stg + rr * 128 + (((2 * cc + 1) ^ (rr & 7)) << 4));
                            if (grow < M) {
                                int p = p_blk + rr, sl = s_blk;
                                while (p >= ep.P) { p -= ep.P; ++sl; }
                                const float4* pb = reinterpret_cast<const float4*>(ep.posb + static_cast<int64_t>(p) * N + n0 + cc * 8);
                                const float4 b0 = __ldg(pb), b1 = __ldg(pb + 1);
                                const int64_t orow = static_cast<int64_t>(sl) * (ep.P + 1 + ep.R) + 1 + ep.R + p;
                                *reinterpret_cast<uint4*>(static_cast<bf16*>(ep.out) + orow * ep.ldo + n0 + cc * 8) =
                                    make_uint4(pack_bf16x2(a0.x + b0.x, a0.y + b0.y), pack_bf16x2(a0.z + b0.z, a0.w + b0.w),
                                               pack_bf16x2(a1.x + b1.x, a1.y + b1.y), pack_bf16x2(a1.z + b1.z, a1.w + b1.w));
                            }
                        }
                        __syncwarp();
                    } else if (row_ok) {
                        // rows are re-mapped (one CLS row inserted per slice): direct 64-byte row stores
                        const int64_t orow = (row / ep.P) * (ep.P + 1 + ep.R) + 1 + ep.R + (row % ep.P);
                        uint4* po = reinterpret_cast<uint4*>(static_cast<bf16*>(ep.out) + orow * ep.ldo + n0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) po[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                    }
                } else if (NSTG == 3) {
                    // whole-region staging: rows of BN bytes; the previous tile's store must have read the buffer
                    if (c == 0) {
                        if (lane == 0) tma_store_wait_read<0>();
                        __syncwarp();
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<uint4*>(stg + lane * 128 + (((c * 4 + i) ^ (lane & 7)) << 4)) =
                            make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                    if (c == CHUNKS_PER_WARP - 1) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0 && !(xp_flags & 0x200)) {
                            if (mode == EPI_BIAS_ACCUM) tma_reduce_add_2d(&tmC, stg, nbase, row0);
                            else tma_store_2d(&tmC, stg, nbase, row0);
                            tma_store_commit();
                        }
                    }
                } else {
                    // the tile written two stores ago must no longer be read by its TMA store
                    uint8_t* tile = stg + stg_sel * STG_TILE;
                    if (NSTG == 2) stg_sel ^= 1;
                    const long long p1 = MST_DBG_CLOCK();
                    if (lane == 0) tma_store_wait_read<(NSTG == 2 ? 1 : 0)>();
                    __syncwarp();
                    const long long p2 = MST_DBG_CLOCK();
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<uint4*>(tile + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) =
                            make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    const long long p3 = MST_DBG_CLOCK();
                    if (lane == 0) {
                        if (mode == EPI_BIAS_ACCUM) tma_reduce_add_2d(&tmC, tile, n0, row0);
                        else tma_store_2d(&tmC, tile, n0, row0);
                        tma_store_commit();
                    }
                    if (dbg_w) { const long long p4 = MST_DBG_CLOCK(); ep.dbg[8] += p1 - p0; ep.dbg[9] += p2 - p1; ep.dbg[10] += p3 - p2; ep.dbg[11] += p4 - p3; }
                }
            };
            auto release_tmem = [&]() {  // every tcgen05.ld of this tile has completed
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    if (TWO && !leader) mbar_arrive_cluster_relaxed(mapa_u32(&tempty_bar[acc], 0));
                    else mbar_arrive(&tempty_bar[acc]);
                }
            };

            const long long e0 = MST_DBG_CLOCK();
            mbar_wait(&tfull_bar[acc], acc_phase);
            const long long e1 = MST_DBG_CLOCK();
            tc_fence_after_sync();
            if (xp_flags & 0x100) {  // experiment: mainloop ceiling (no epilogue work at all)
                release_tmem();
                acc ^= 1; if (acc == 0) acc_phase ^= 1;
                continue;
            }
            uint32_t ra[32], rb[32], rc[CHUNKS_PER_WARP == 3 ? 32 : 1];
            tmem_ld_32x32b_x32(taddr, ra);
            tmem_ld_32x32b_x32(taddr + 32, rb);
            if constexpr (CHUNKS_PER_WARP == 3) tmem_ld_32x32b_x32(taddr + 64, rc);
            tmem_ld_wait();
            release_tmem();  // the accumulator stage goes back to the MMA warp before any epilogue math
            const long long e2 = MST_DBG_CLOCK();
            process(ra, 0);
            process(rb, 1);
            if constexpr (CHUNKS_PER_WARP == 3) process(rc, 2);
            if (kResident && ep.rowpart_out != nullptr) {   // this warp's partial for its rows: slot 2 * n_block + column group
                if (row_ok) *reinterpret_cast<float2*>(ep.rowpart_out + (row * 4 + n_blk * 2 + hf) * 2) = make_float2(st1, st2);
                st1 = 0.f; st2 = 0.f;
            }
            if constexpr (L::STAT_BYTES > 0) {
                if (seqn && n_blk == n_tiles - 1) {
                    // both column groups of a row quadrant meet in shared memory; group 0 turns the totals into rstd
                    float* sst = reinterpret_cast<float*>(smem + L::STAT_OFF) + ((it / n_tiles) & 1) * 512;
                    const int r = q * 32 + lane;
                    sst[(hf * 128 + r) * 2] = st1;
                    sst[(hf * 128 + r) * 2 + 1] = st2;
                    named_bar_sync(1 + q, 64);
                    if (hf == 0 && row_ok) {
                        const float t1 = st1 + sst[(128 + r) * 2], t2 = st2 + sst[(128 + r) * 2 + 1];
                        const float mean = t1 / N;
                        ep.rowstat_out[row] = rsqrtf(fmaxf(t2 / N - mean * mean, 0.f) + ep.stat_eps);
                    }
                    st1 = 0.f; st2 = 0.f;
                }
            }
            if (kDbgTiming && ep.dbg != nullptr && blockIdx.x == 0 && e == 0 && lane == 0) {
                const long long e3 = MST_DBG_CLOCK();
                ep.dbg[4] += e1 - e0; ep.dbg[5] += e2 - e1; ep.dbg[6] += e3 - e2;
            }
            acc ^= 1; if (acc == 0) acc_phase ^= 1;
        }
        if (lane == 0) tma_store_wait_all<0>();  // all output bytes are in global memory before the CTA retires
        __syncwarp();
    }

    tc_fence_before_sync();
    if (PAIR != 0) cluster_sync_all(); else __syncthreads();  // the peer may still multicast into / arrive on this CTA's smem
    if (warp == 2) {
        tc_fence_after_sync();
        if (TWO) tmem_dealloc_pair<kTmemCols>(tmem_base); else tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------------
template <int BN, int KCH, int STAGES, int NSTG, int EW, int PAIR = 0>
static int launch_cfg(const TmaDesc& tmA, const TmaDesc& tmB, const TmaDesc& tmC, int M, int N, int K, int mode,
                      const EpiParams& ep, int num_sms, cudaStream_t stream) {
    using L = GemmSmem<BN, KCH, STAGES, NSTG, EW, PAIR>;
    constexpr bool MCAST = PAIR != 0;  // launched as clusters
    constexpr int CLUSTER = PAIR == 4 ? 4 : 2;
    auto kern = gemm_tc_kernel<BN, KCH, STAGES, NSTG, EW, PAIR>;
    MST_SET_DYN_SMEM(kern, L::DYN_BYTES);
    constexpr int CPU_ = PAIR == 2 ? 2 : 1;               // CTAs per scheduling unit (a cta_group::2 pair shares its tiles)
    const int n_tiles = N / BN;
    const int m_tiles = (M + BM * CPU_ - 1) / (BM * CPU_);
    int units = num_sms / CPU_;
    if (CLUSTER == 4 && MCAST) {  // clusters of 4 do not tile every GPC: size the grid to what is co-resident
        static int max_clusters_dev[kMaxDevices];
        static bool max_clusters_set[kMaxDevices] = {};
        int dev = 0;
        MST_CHECK_CUDA(cudaGetDevice(&dev));
        MST_REQUIRE(dev >= 0 && dev < kMaxDevices, "device ordinal %d out of range", dev);
        int& max_clusters = max_clusters_dev[dev];
        if (!max_clusters_set[dev]) {
            max_clusters_set[dev] = true;
            cudaLaunchConfig_t q{};
            q.gridDim = dim3(num_sms / 4 * 4); q.blockDim = dim3(gemm_threads(EW)); q.dynamicSmemBytes = L::DYN_BYTES;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 4; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.attrs = qa; q.numAttrs = 1;
            MST_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &q));
        }
        if (max_clusters * 4 < units) units = max_clusters * 4;
    }
    int grid;
    if (KCH > 0) {
        int gs = units / n_tiles;
        if (gs < 1) gs = 1;
        if (gs > m_tiles) gs = m_tiles;
        grid = gs * n_tiles * CPU_;
    } else {
        const int ksplit = (mode >> 16) > 1 ? (mode >> 16) : 1;
        const long long total = static_cast<long long>(m_tiles) * n_tiles * ksplit;
        grid = static_cast<int>(total < units ? total : units) * CPU_;
    }
    if (MCAST) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(gemm_threads(EW)); cfg.dynamicSmemBytes = L::DYN_BYTES; cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
        MST_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, M, N, K, mode, ep));
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(gemm_threads(EW)); cfg.dynamicSmemBytes = L::DYN_BYTES; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
        MST_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, M, N, K, mode, ep));
    }
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Programmatic dependent launch of the tcgen05 kernels (ptx.cuh: every kernel of the encoder calls griddepcontrol.wait before it
// touches data): on for launch-bound small batches, where a kernel's prologue (barrier init, TMEM allocation, descriptor
// prefetch, the resident weight block) is a visible share of its ~10 us and can overlap the tail of the kernel in front; off for
// large batches (measured flat under the power cap).  Set per forward by mst_forward.
static thread_local bool g_pdl = false;
void set_pdl(bool on) { g_pdl = on; }
bool pdl_enabled() {
    static const int pdl = exp_env("MST_PDL", -1);   // experiments: force on (1) / off (0)
    return pdl >= 0 ? pdl != 0 : g_pdl;
}

bool gemm_wt_enabled() {
    static const int use_wt = exp_env("MST_GEMM_WT", 1);  // 0: experiments / A-B comparisons
    return use_wt != 0;
}

int gemm_bf16_tc(const bf16* A, const bf16* W, int M, int N, int K, int mode, const EpiParams& ep, int num_sms,
                 cudaStream_t stream) {
    MST_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    MST_REQUIRE(K % BK == 0, "gemm: K=%d must be a multiple of %d", K, BK);
    MST_REQUIRE(N % 192 == 0 || N % 128 == 0, "gemm: N=%d must be a multiple of 192 or 128", N);
    MST_REQUIRE(ep.ldo % 8 == 0, "gemm: output row stride must be a multiple of 8 elements");
    MST_REQUIRE(mode != EPI_RAW_F32 || (K > 384 && N % 192 == 0), "gemm: fp32 output is wired for the streaming schedule (K > 384, N %% 192 == 0)");
    if (mode == EPI_RAW_F32) {
        // few output tiles, a long contraction: split K so that every SM pair has work (partials meet by fp32 reductions, so the
        // output is zeroed first; the summation order across the splits is not fixed)
        const int tiles = ((M + 255) / 256) * (N / 192);
        int ks = (num_sms / 2) / (tiles > 0 ? tiles : 1);
        const int kchunks = K / BK;
        if (ks > kchunks / 8) ks = kchunks / 8;    // at least 8 k-chunks per unit
        if (ks > 64) ks = 64;
        while (ks > 1 && (ks - 1) * ((kchunks + ks - 1) / ks) >= kchunks) --ks;   // every unit gets a non-empty range of k-chunks
        if (ks > 1) {
            MST_CHECK_CUDA(cudaMemsetAsync(ep.out, 0, static_cast<size_t>(M) * ep.ldo * sizeof(float), stream));
            mode |= ks << 16;
        }
    }
    // in-place residual update: `out += acc + bias` as 16-byte vector reductions (no residual read by the SM)
    const bool want_stats = ep.rowstat_out != nullptr;
    if (ep.rowpart_out != nullptr) {
        MST_REQUIRE(mode == EPI_BIAS_RES && N == 384 && K == 384 && !want_stats,
                    "gemm: row partials need EPI_BIAS_RES on the weight-resident N = K = 384 GEMM; got N=%d K=%d mode=%d", N, K, mode);
    } else if (want_stats) {
        MST_REQUIRE(mode == EPI_BIAS_RES && N % 192 == 0 && N / 192 == 2 && K > 384,
                    "gemm: fused row statistics need EPI_BIAS_RES, N = 384 and the streaming pair schedule (K > 384); got N=%d K=%d mode=%d",
                    N, K, mode);
    } else if (mode == EPI_BIAS_RES && ep.res == ep.out && ep.ldr == ep.ldo) {
        mode = EPI_BIAS_ACCUM;
    }
    const int use_wt = gemm_wt_enabled();
    if (use_wt && gemm_wt_supported(M, N, K, mode, ep)) {
        static const int wt_skip = exp_env("MST_GEMM_SKIP_EPI", 0);  // experiments only
        EpiParams e2 = ep;
        e2.P = wt_skip;
        return gemm_bf16_wt(A, W, M, N, K, mode, e2, num_sms, stream);  // weights in TMEM, 16 epilogue warps (gemm_wt.cu)
    }
    MST_REQUIRE(ep.rowpart == nullptr, "gemm: row partials (rowpart) are consumed by gemm_wt only");
    TmaDesc tmA, tmB, tmC;
    MST_PROPAGATE(make_tma_2d_bf16(&tmA, A, K, M, K, BK, BM));
    // output maps (unused by EPI_PATCH, whose rows are re-mapped)
    const uint64_t out_rows = mode == EPI_PATCH ? static_cast<uint64_t>(M / ep.P) * (ep.P + 1 + ep.R) : static_cast<uint64_t>(M);
    static const int force_bn = exp_env("MST_GEMM_BN", 0);  // experiments only
    static const int skip_epi = exp_env("MST_GEMM_SKIP_EPI", 0);
    if (skip_epi) mode |= (skip_epi & 3) << 8;  // 1: no epilogue at all, 2: epilogue without the final store
    static const int no_mcast = exp_env("MST_GEMM_NO_MCAST", 0);  // experiments only
    static const int use_two = exp_env("MST_GEMM_TWO", 1);             // experiments only
    if (N % 192 == 0 && force_bn != 128) {
        MST_PROPAGATE(make_tma_2d_bf16(&tmB, W, K, N, K, BK, 192));
        if (K == 256) {  // patch embedding: weight-resident, chunk staging (its epilogue stores rows directly)
            MST_PROPAGATE(make_tma_2d_bf16(&tmC, ep.out, N, out_rows, ep.ldo, 32, 32, false, true));
            // 12 epilogue warps (two 32-column chunks each instead of three): the epilogue, not the 16 MMAs, paces this GEMM
            // (0.47 -> 0.38 ms per launch in-step); MST_PATCH_EW12=0 keeps the 8-warp configuration for A-B runs
            static const int ew12 = exp_env("MST_PATCH_EW12", 1);
            if (ew12) return launch_cfg<192, 4, 4, 2, 12>(tmA, tmB, tmC, M, N, K, mode, ep, num_sms, stream);
            return launch_cfg<192, 4, 4, 2, 8>(tmA, tmB, tmC, M, N, K, mode, ep, num_sms, stream);
        }
        if (K == 384 && use_two == 2) {  // measured slower than the multicast pair below (0.512 vs 0.428 ms for qkv): every CTA
                                         // streams all of its own A again, and A supply is what bounds the resident schedule
            // cta_group::2: one 256 x 192 MMA per CTA pair; each CTA keeps HALF of the resident weight slab (72 KB) and
            // streams its own 128 rows of A, so the tensor core reads 7 KB instead of 10 KB of shared memory per k-step
            // (the shared-memory port, shared with the TMA fills and the epilogue staging, is what bounds this kernel)
            TmaDesc tmBh;
            MST_PROPAGATE(make_tma_2d_bf16(&tmBh, W, K, N, K, BK, 96));
            MST_PROPAGATE(make_tma_2d_bf16(&tmC, ep.out, N, out_rows, ep.ldo, 32, 32, false, true));
            if ((mode & 0xff) == EPI_BIAS_GELU || (mode & 0xff) == EPI_LN_BIAS_GELU)
                return launch_cfg<192, 6, 6, 2, 12, 2>(tmA, tmBh, tmC, M, N, K, mode, ep, num_sms, stream);
            return launch_cfg<192, 6, 6, 2, 8, 2>(tmA, tmBh, tmC, M, N, K, mode, ep, num_sms, stream);
        }
        if (K == 384 && (N / 192) % 2 == 0 && !no_mcast) {
            // weight slab resident (144 KB), A stages fetched half-and-half by a CTA pair and TMA-multicast to both
            TmaDesc tmAh;
            MST_PROPAGATE(make_tma_2d_bf16(&tmAh, A, K, M, K, BK, BM / 2));
            MST_PROPAGATE(make_tma_2d_bf16(&tmC, ep.out, N, out_rows, ep.ldo, 32, 32, false, true));
            // GELU epilogue (fc1) is the longest: two staging tiles per warp (a TMA store stays in flight while the next
            // chunk is computed) paid for with a 3-deep A ring; the L2 prefetch keeps the shorter ring fed
            static const int mc4 = exp_env("MST_GEMM_MC4", 0);  // experiments only
            if (mc4 && (N / 192) % 4 == 0) {  // four adjacent n-blocks share every A stage: each CTA fetches 32 rows
                TmaDesc tmAq;
                MST_PROPAGATE(make_tma_2d_bf16(&tmAq, A, K, M, K, BK, BM / 4));
                if ((mode & 0xff) == EPI_BIAS_GELU || (mode & 0xff) == EPI_LN_BIAS_GELU)
                    return launch_cfg<192, 6, 3, 1, 12, 4>(tmAq, tmB, tmC, M, N, K, mode, ep, num_sms, stream);
                return launch_cfg<192, 6, 4, 1, 8, 4>(tmAq, tmB, tmC, M, N, K, mode, ep, num_sms, stream);
            }
            if ((mode & 0xff) == EPI_BIAS_GELU || (mode & 0xff) == EPI_LN_BIAS_GELU)
                return launch_cfg<192, 6, 3, 1, 12, 1>(tmAh, tmB, tmC, M, N, K, mode, ep, num_sms, stream);
            static const int alt = exp_env("MST_GEMM_ALT", 0);  // experiments only
            if (alt == 1)   // two staging tiles per epilogue warp (no wait for the previous chunk's TMA store), 3-stage A ring
                return launch_cfg<192, 6, 3, 2, 8, 1>(tmAh, tmB, tmC, M, N, K, mode, ep, num_sms, stream);
            if (alt == 2)   // 12 epilogue warps (2 chunks each), 3-stage A ring
                return launch_cfg<192, 6, 3, 1, 12, 1>(tmAh, tmB, tmC, M, N, K, mode, ep, num_sms, stream);
            return launch_cfg<192, 6, 4, 1, 8, 1>(tmAh, tmB, tmC, M, N, K, mode, ep, num_sms, stream);
        }
        // K too large for a resident weight slab (fc2): both operands stream through the ring
        MST_PROPAGATE(make_tma_2d_bf16(&tmC, ep.out, N, out_rows, ep.ldo, 32, 32, false, true));
        if (want_stats) mode |= 0x400;
        MST_REQUIRE(!want_stats || use_two, "gemm: fused row statistics need the cta_group::2 streaming schedule");
        if (use_two) {  // cta_group::2: 16 KB of A + 12 KB (half) of the weight tile per CTA and stage
            TmaDesc tmBh;
            MST_PROPAGATE(make_tma_2d_bf16(&tmBh, W, K, N, K, BK, 96));
            return launch_cfg<192, 0, 7, 1, 8, 2>(tmA, tmBh, tmC, M, N, K, mode, ep, num_sms, stream);
        }
        return launch_cfg<192, 0, 5, 1, 8>(tmA, tmB, tmC, M, N, K, mode, ep, num_sms, stream);
    }
    MST_PROPAGATE(make_tma_2d_bf16(&tmB, W, K, N, K, BK, 128));
    MST_PROPAGATE(make_tma_2d_bf16(&tmC, ep.out, N, out_rows, ep.ldo, 64, 32, true, false));
    return launch_cfg<128, 0, 5, 3, 8>(tmA, tmB, tmC, M, N, K, mode, ep, num_sms, stream);
}

}  // namespace mst
