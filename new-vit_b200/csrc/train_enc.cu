// Backward pass of the DINOv2 encoder (BASELINE.json config 5 with the default, un-frozen construction: main_train.py:110-126 trains
// every parameter of dino.py:56-103).  bf16 activations and activation gradients, fp32 weight gradients, fp32 statistics.
//
// The heavy contractions are tcgen05 GEMMs:
//   dgrad   dX = dY . W          ->  gemm_tc.cu, gemm(A = dY [M, N_out], weight = W^T [K_in, N_out])          (bf16 out)
//   wgrad   dW = dY^T . X        ->  gemm_wgrad.cu, MN-major operands straight from the row-major activations  (fp32 out, bias gradient too);
//                                    shapes that kernel does not take go through the transposes below and gemm_tc.cu's EPI_RAW_F32
// This file holds what sits between them: LayerNorm backward, GELU forward / backward, the attention backward (warp-level tensor cores:
// recomputes the 257 x 257 probabilities per (slice, head) from q, k and the row log-sum-exp; nothing N x N touches HBM), the fallback
// transposes (token-major -> K-major operands, with the bias gradient = column sums on the way), and the small gradients of the patch
// embedding / position table / class token.
#include <math_constants.h>
#include "common.cuh"
#include "ptx.cuh"

namespace mst {

namespace {
__device__ __forceinline__ float wsum32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
}  // namespace

// ---------------------------------------------------------------------------------------------------
// LayerNorm backward, one warp per row (E = 32 * EPL, EPL <= 32), eps as in the forward (1e-6 encoder):
//   xhat = (x - mean) rstd;  g = dy * gamma;  dx = rstd (g - mean(g) - xhat mean(g xhat)) (+ dres)
//   dgamma += dy * xhat, dbeta += dy: per-CTA partial sums [gridDim.x][2][E], reduced by ln_bwd_reduce_kernel
// row_map (nullable): dy / dres / dx rows are 0..rows-1, the x row of row r is row_map_stride * r (the CLS rows of the final norm)
// ---------------------------------------------------------------------------------------------------
template <int EPL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const bf16* __restrict__ x, int64_t x_row_stride, const bf16* __restrict__ dy,
                                                      const float* __restrict__ dy_f32, const bf16* __restrict__ dres,
                                                      const float* __restrict__ gamma, bf16* __restrict__ dx, float* __restrict__ partial,
                                                      int rows, float eps) {
    // lane -> elements 4 lane + 128 g + {0..3}, g < EPL / 4: every load and store of a warp is 256 contiguous bytes
    constexpr int E = EPL * 32, G = EPL / 4;
    static_assert(EPL % 4 == 0, "row length must be a multiple of 128");
    __shared__ float red[2][E];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 2 * E; i += blockDim.x) (&red[0][0])[i] = 0.f;
    __syncthreads();
    float gam[EPL], dg[EPL], db[EPL];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(gamma + 128 * g + 4 * lane);
        gam[4 * g] = v.x; gam[4 * g + 1] = v.y; gam[4 * g + 2] = v.z; gam[4 * g + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < EPL; ++i) { dg[i] = 0.f; db[i] = 0.f; }
    auto load4 = [&](const bf16* p, float* o) {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    };
    // E = 384 with bf16 dy (every LayerNorm of the blocks): the NEXT row's x / dy / dres are requested (raw, 8 bytes per piece) before
    // this row's four warp reductions, so a warp always has loads in flight -- one row at a time left the kernel at 3.6 TB/s.
    constexpr bool kPrefetch = EPL == 12;
    const bool pf = kPrefetch && dy != nullptr;
    const int r_first = blockIdx.x * nw + warp, r_step = gridDim.x * nw;
    uint2 nx[G], ndy[G], nrs[G];
    auto request = [&](int r) {
        const bf16* xr = x + static_cast<int64_t>(r) * x_row_stride;
        const int64_t ro = static_cast<int64_t>(r) * E;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            nx[g] = *reinterpret_cast<const uint2*>(xr + 128 * g + 4 * lane);
            ndy[g] = *reinterpret_cast<const uint2*>(dy + ro + 128 * g + 4 * lane);
            nrs[g] = dres ? *reinterpret_cast<const uint2*>(dres + ro + 128 * g + 4 * lane) : make_uint2(0u, 0u);
        }
    };
    auto unpack4 = [&](const uint2& u, float* o) {
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    };
    if (pf && r_first < rows) request(r_first);
    for (int r = r_first; r < rows; r += r_step) {
        const bf16* xr = x + static_cast<int64_t>(r) * x_row_stride;
        const int64_t ro = static_cast<int64_t>(r) * E;
        float xv[EPL], dyv[EPL];
        uint2 crs[G];
        float s = 0.f;
        if (pf) {
#pragma unroll
            for (int g = 0; g < G; ++g) { unpack4(nx[g], xv + 4 * g); unpack4(ndy[g], dyv + 4 * g); crs[g] = nrs[g]; }
            if (r + r_step < rows) request(r + r_step);
        } else {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                load4(xr + 128 * g + 4 * lane, xv + 4 * g);
                if (dy) load4(dy + ro + 128 * g + 4 * lane, dyv + 4 * g);
                else {
                    const float4 v = *reinterpret_cast<const float4*>(dy_f32 + ro + 128 * g + 4 * lane);
                    dyv[4 * g] = v.x; dyv[4 * g + 1] = v.y; dyv[4 * g + 2] = v.z; dyv[4 * g + 3] = v.w;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < EPL; ++i) s += xv[i];
        const float mean = wsum32(s) * (1.0f / E);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < EPL; ++i) { const float d = xv[i] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(wsum32(q) * (1.0f / E) + eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            xv[i] = (xv[i] - mean) * rstd;            // xhat
            const float g = dyv[i] * gam[i];
            s1 += g; s2 = fmaf(g, xv[i], s2);
            dg[i] = fmaf(dyv[i], xv[i], dg[i]);
            db[i] += dyv[i];
        }
        const float m1 = wsum32(s1) * (1.0f / E), m2 = wsum32(s2) * (1.0f / E);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float v[4], rs[4] = {0.f, 0.f, 0.f, 0.f};
            if (pf) unpack4(crs[g], rs);
            else if (dres) load4(dres + ro + 128 * g + 4 * lane, rs);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = rstd * (dyv[4 * g + j] * gam[4 * g + j] - m1 - xv[4 * g + j] * m2) + rs[j];
            *reinterpret_cast<uint2*>(dx + ro + 128 * g + 4 * lane) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&red[0][128 * g + 4 * lane + j], dg[4 * g + j]);
            atomicAdd(&red[1][128 * g + 4 * lane + j], db[4 * g + j]);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * E; i += blockDim.x) partial[static_cast<int64_t>(blockIdx.x) * 2 * E + i] = (&red[0][0])[i];
}
// dgamma[e] = sum over CTAs of partial[.][0][e], dbeta likewise: one CTA per 32 columns, 8 partial rows in flight per column,
// summed in a fixed order (the result does not depend on scheduling)
__global__ void __launch_bounds__(256) ln_bwd_reduce_kernel(const float* __restrict__ partial, int nparts, int E, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta) {
    __shared__ float acc[8][33];
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5, i = blockIdx.x * 32 + c;
    float a = 0.f;
    for (int p = g; p < nparts; p += 8) a += partial[static_cast<int64_t>(p) * 2 * E + i];
    acc[g][c] = a;
    __syncthreads();
    if (g == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += acc[k][c];
        if (i < E) dgamma[i] = t; else dbeta[i - E] = t;
    }
}
constexpr int LN_BWD_GRID = 296;   // 2 CTAs per SM are resident (128 registers x 256 threads); fewer partials for the reduce
size_t ln_bwd_workspace_bytes(int E) { return static_cast<size_t>(LN_BWD_GRID) * 2 * E * sizeof(float); }

int launch_ln_bwd(const bf16* x, int64_t x_row_stride, const bf16* dy, const float* dy_f32, const bf16* dres, const float* gamma, bf16* dx,
                  float* dgamma, float* dbeta, int rows, int E, float eps, float* workspace, cudaStream_t stream) {
    MST_REQUIRE(E == 384 || E == 768, "LayerNorm backward: E=%d unsupported (384 / 768)", E);
    MST_REQUIRE((dy != nullptr) != (dy_f32 != nullptr), "LayerNorm backward: exactly one of dy (bf16) / dy_f32");
    int grid = (rows + 7) / 8;
    grid = grid < 1 ? 1 : (grid > LN_BWD_GRID ? LN_BWD_GRID : grid);
    if (E == 384) ln_bwd_kernel<12><<<grid, 256, 0, stream>>>(x, x_row_stride, dy, dy_f32, dres, gamma, dx, workspace, rows, eps);
    else ln_bwd_kernel<24><<<grid, 256, 0, stream>>>(x, x_row_stride, dy, dy_f32, dres, gamma, dx, workspace, rows, eps);
    MST_CHECK_CUDA(cudaGetLastError());
    ln_bwd_reduce_kernel<<<2 * E / 32, 256, 0, stream>>>(workspace, grid, E, dgamma, dbeta);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// GELU (mlp.py:36, exact erf form in the reference).  Forward in training = the inference epilogue's function (common.cuh
// gelu_tanh_fit, |deviation from erf-GELU| <= 2.5e-5) so that a training forward equals an inference forward bit for bit;
// backward = Phi(u) + u phi(u), the derivative of the exact form, evaluated with the forward's fit of Phi.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const bf16* __restrict__ u, bf16* __restrict__ y, int64_t n8) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint4 a = reinterpret_cast<const uint4*>(u)[i];
        uint4 o;
        const uint32_t* ai = reinterpret_cast<const uint32_t*>(&a);
        uint32_t* oi = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(ai[j]);
            oi[j] = pack_bf16x2(gelu_tanh_fit(f.x), gelu_tanh_fit(f.y));
        }
        reinterpret_cast<uint4*>(y)[i] = o;
    }
}
// d/du of u Phi(u) = Phi(u) + u phi(u): Phi from the forward's tanh fit, phi by one ex2 (|error| < 3e-4 including tanh.approx,
// an eighth of the bf16 step of the result; erff + expf made this kernel instruction-bound at half the HBM rate)
__device__ __forceinline__ float gelu_grad(float u) {
    float t;
    const float uu = u * u, x2 = fminf(uu, 64.0f);
    float p = fmaf(-3.51516788e-04f, x2, 3.70056460e-02f);
    p = fmaf(p, x2, 7.97507884e-01f);
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u * p));
    const float pdf = 0.3989422804014327f * ex2_approx(-0.72134752044448f * uu);
    return fmaf(u, pdf, fmaf(0.5f, t, 0.5f));
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const bf16* __restrict__ u, const bf16* dy, bf16* du, int64_t n8) {   // du may alias dy
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint4 a = reinterpret_cast<const uint4*>(u)[i], d = reinterpret_cast<const uint4*>(dy)[i];
        uint4 o;
        const uint32_t* ai = reinterpret_cast<const uint32_t*>(&a);
        const uint32_t* di = reinterpret_cast<const uint32_t*>(&d);
        uint32_t* oi = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(ai[j]), g = unpack_bf16x2(di[j]);
            oi[j] = pack_bf16x2(g.x * gelu_grad(f.x), g.y * gelu_grad(f.y));
        }
        reinterpret_cast<uint4*>(du)[i] = o;
    }
}
int launch_gelu_fwd(const bf16* u, bf16* y, int64_t n, int num_sms, cudaStream_t stream) {
    MST_REQUIRE(n % 8 == 0, "gelu: element count must be a multiple of 8");
    gelu_fwd_kernel<<<num_sms * 8, 256, 0, stream>>>(u, y, n / 8);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int launch_gelu_bwd(const bf16* u, const bf16* dy, bf16* du, int64_t n, int num_sms, cudaStream_t stream) {
    MST_REQUIRE(n % 8 == 0, "gelu backward: element count must be a multiple of 8");
    gelu_bwd_kernel<<<num_sms * 8, 256, 0, stream>>>(u, dy, du, n / 8);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// in [M, C] bf16 (row stride ld) -> out [C, Mpad] bf16 (columns M..Mpad-1 zero), 64 x 64 tiles through shared memory;
// colsum (nullable) [C] fp32 += column sums (the bias gradient of the Linear whose dY this is): one atomicAdd per column and CTA.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_colsum_kernel(const bf16* __restrict__ in, int64_t ld, bf16* __restrict__ out,
                                                                float* __restrict__ colsum, int M, int C, int Mpad) {
    // 64 (tokens) x 64 (features) tile held as 32-bit words (two adjacent features per word), rows padded to 33 words.
    // Load: 8 lanes read one token row's 128 bytes (16-byte loads).  Store: 8 lanes write 64 consecutive tokens of one feature
    // row (128 contiguous bytes), each lane gathering its 8 tokens from 8 tile rows (one PRMT per output word).
    __shared__ uint32_t tile[64][33];
    __shared__ float csum[64];
    const int m0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    if (threadIdx.x < 64) csum[threadIdx.x] = 0.f;
    __syncthreads();
    {
        const int q = threadIdx.x & 7;
        float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int r = pass * 32 + (threadIdx.x >> 3), m = m0 + r;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (m < M) v = __ldg(reinterpret_cast<const uint4*>(in + static_cast<int64_t>(m) * ld + c0 + 8 * q));
            tile[r][4 * q + 0] = v.x; tile[r][4 * q + 1] = v.y; tile[r][4 * q + 2] = v.z; tile[r][4 * q + 3] = v.w;
            if (colsum) {
                float2 f;
                f = unpack_bf16x2(v.x); cs[0] += f.x; cs[1] += f.y;
                f = unpack_bf16x2(v.y); cs[2] += f.x; cs[3] += f.y;
                f = unpack_bf16x2(v.z); cs[4] += f.x; cs[5] += f.y;
                f = unpack_bf16x2(v.w); cs[6] += f.x; cs[7] += f.y;
            }
        }
        if (colsum) {
            // lanes q, q+8, q+16, q+24 of a warp hold the same 8 columns
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 8);
                cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
            }
            if ((threadIdx.x & 31) < 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) atomicAdd(&csum[8 * q + j], cs[j]);
            }
        }
    }
    __syncthreads();
    {
        const int mg = threadIdx.x & 7, cw = threadIdx.x >> 3;   // 8 tokens mg*8.., feature pair cw (features 2cw, 2cw+1)
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = tile[mg * 8 + j][cw];
        uint4 lo, hi;   // feature 2cw: low halves; feature 2cw+1: high halves
        lo.x = __byte_perm(w[0], w[1], 0x5410); lo.y = __byte_perm(w[2], w[3], 0x5410);
        lo.z = __byte_perm(w[4], w[5], 0x5410); lo.w = __byte_perm(w[6], w[7], 0x5410);
        hi.x = __byte_perm(w[0], w[1], 0x7632); hi.y = __byte_perm(w[2], w[3], 0x7632);
        hi.z = __byte_perm(w[4], w[5], 0x7632); hi.w = __byte_perm(w[6], w[7], 0x7632);
        bf16* o0 = out + static_cast<int64_t>(c0 + 2 * cw) * Mpad + m0 + mg * 8;
        *reinterpret_cast<uint4*>(o0) = lo;
        *reinterpret_cast<uint4*>(o0 + Mpad) = hi;
    }
    if (colsum && threadIdx.x < 64) atomicAdd(colsum + c0 + threadIdx.x, csum[threadIdx.x]);
}
int launch_transpose_colsum(const bf16* in, int64_t ld, bf16* out, float* colsum, int M, int C, int Mpad, cudaStream_t stream) {
    MST_REQUIRE(C % 64 == 0 && Mpad % 64 == 0 && Mpad >= M && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "transpose: C=%d Mpad=%d M=%d ld=%lld unsupported", C, Mpad, M, (long long)ld);
    dim3 grid(Mpad / 64, C / 64);
    transpose_colsum_kernel<<<grid, 256, 0, stream>>>(in, ld, out, colsum, M, C, Mpad);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// W [N, K] fp32 (nn.Linear layout) -> Wt [K, N] bf16 (the dgrad GEMM's weight operand), optional per-row scale of W
__global__ void __launch_bounds__(256) transpose_f32_to_bf16_kernel(const float* __restrict__ W, bf16* __restrict__ Wt, int N, int K) {
    __shared__ float tile[32][33];
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) tile[r][tx] = (n0 + r < N && k0 + tx < K) ? W[static_cast<int64_t>(n0 + r) * K + k0 + tx] : 0.f;
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
        if (k0 + r < K && n0 + tx < N) Wt[static_cast<int64_t>(k0 + r) * N + n0 + tx] = __float2bfloat16_rn(tile[tx][r]);
}
int launch_transpose_f32_to_bf16(const float* W, bf16* Wt, int N, int K, cudaStream_t stream) {
    dim3 grid((N + 31) / 32, (K + 31) / 32);
    transpose_f32_to_bf16_kernel<<<grid, 256, 0, stream>>>(W, Wt, N, K);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mst

namespace mst {

// ---------------------------------------------------------------------------------------------------
// Attention backward (attention.py:56-69), head_dim 64, one CTA per (slice, head), warp-level tensor cores (mma.sync m16n8k16).
//   qkv [BD*N, 3E] bf16 with q pre-scaled by 1/8 (as the forward's qkv GEMM writes it), o [BD*N, E] the forward output,
//   dO  [BD*N, E]  ->  dqkv [BD*N, 3E]: d/d(q_raw) = 1/8 dS K,  d/dk = dS^T q,  d/dv = P^T dO
// Q, K, V and dO of the item sit in shared memory (XOR-swizzled 16-byte chunks, as attention_mma.cu).  Three phases:
//   1  D_i = dO_i . O_i  and  lse_i = log2 sum_j 2^(s_ij log2e)   (one query block of 16 rows per warp, online over key chunks;
//      skipped when the forward kernel saved lse: N = 257)
//   2  dQ:  per query block, per 32-key chunk:  P = 2^(S log2e - lse), dP = dO V^T, dS = P (dP - D), dQ += dS K
//   3  dK, dV: per 16-key block, per 32-query chunk, the transposed products:  P^T from K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q
// The N x N probabilities exist only as register fragments.
// ---------------------------------------------------------------------------------------------------
namespace attb {
constexpr int MAX_WARPS = 12;     // 384 threads x 168 registers fill the register file
constexpr float LOG2E = 1.4426950408889634f;

// A-operand fragments of 16 rows (g, g+8) x 64 dims from a swizzled [rows][128 B] shared-memory tile
__device__ __forceinline__ void load_a(uint32_t (&a)[4][4], const uint8_t* tile, int r0, int r1, int t) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        a[ks][0] = *reinterpret_cast<const uint32_t*>(tile + r0 * 128 + (((2 * ks) ^ (r0 & 7)) << 4) + 4 * t);
        a[ks][1] = *reinterpret_cast<const uint32_t*>(tile + r1 * 128 + (((2 * ks) ^ (r1 & 7)) << 4) + 4 * t);
        a[ks][2] = *reinterpret_cast<const uint32_t*>(tile + r0 * 128 + (((2 * ks + 1) ^ (r0 & 7)) << 4) + 4 * t);
        a[ks][3] = *reinterpret_cast<const uint32_t*>(tile + r1 * 128 + (((2 * ks + 1) ^ (r1 & 7)) << 4) + 4 * t);
    }
}
// acc[nb] (16 x 8) = A (16 x 64) . B[row0 + nb*8 .. +8]^T for nb < 4: B rows are the "n" index, K-major (ldmatrix, no transpose)
__device__ __forceinline__ void mma_rows(float (&acc)[4][4], const uint32_t (&a)[4][4], uint32_t tile_u32, int row0, int lane) {
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
    // issue order: consecutive MMAs write different accumulators (a dependent pair back to back waits out the MMA latency)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int c = half * 4 + (lane >> 3);
        uint32_t b[4][4];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            const int row = row0 + nb * 8 + (lane & 7);
            ldmatrix_x4(b[nb], tile_u32 + row * 128 + ((c ^ (row & 7)) << 4));
        }
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) mma_bf16_16816(acc[nb], a[2 * half], b[nb][0], b[nb][1]);
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) mma_bf16_16816(acc[nb], a[2 * half + 1], b[nb][2], b[nb][3]);
    }
}
// out (16 x 64) += A (16 x 32, from the C fragments c[4][4] of a 16 x 32 block) . B[row0 .. row0+32] (32 x 64, ldmatrix transposed)
__device__ __forceinline__ void mma_cols(float (&out)[8][4], const float (&c)[4][4], uint32_t tile_u32, int row0, int lane) {
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(c[2 * kk][0], c[2 * kk][1]);
        pa[1] = pack_bf16x2(c[2 * kk][2], c[2 * kk][3]);
        pa[2] = pack_bf16x2(c[2 * kk + 1][0], c[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(c[2 * kk + 1][2], c[2 * kk + 1][3]);
        const int row = row0 + kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
        for (int dn = 0; dn < 8; dn += 2) {
            const int ch = dn + (lane >> 4);
            uint32_t b[4];
            ldmatrix_x4_trans(b, tile_u32 + row * 128 + ((ch ^ (row & 7)) << 4));
            mma_bf16_16816(out[dn], pa, b[0], b[1]);
            mma_bf16_16816(out[dn + 1], pa, b[2], b[3]);
        }
    }
}
}  // namespace attb

__global__ void __launch_bounds__(attb::MAX_WARPS * 32, 1)
attention_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o, const bf16* __restrict__ dO, bf16* __restrict__ dqkv,
                     const float* __restrict__ lse_in, int N, int heads, int NKP) {
    using namespace attb;
    extern __shared__ __align__(128) uint8_t smem_ab[];
    uint8_t* Qs = smem_ab;
    uint8_t* Ks = Qs + NKP * 128;
    uint8_t* Vs = Ks + NKP * 128;
    uint8_t* Gs = Vs + NKP * 128;                                   // dO
    float* lse = reinterpret_cast<float*>(Gs + NKP * 128);          // [NKP] log2-domain log-sum-exp (+inf for padded rows)
    float* Dv = lse + NKP;                                          // [NKP] dO_i . O_i
    const int s = blockIdx.x / heads, h = blockIdx.x % heads;
    const int E = heads * 64;
    const int64_t ld = 3 * E;
    const bf16* qbase = qkv + static_cast<int64_t>(s) * N * ld + h * 64;
    const bf16* obase = o + static_cast<int64_t>(s) * N * E + h * 64;
    const bf16* gbase = dO + static_cast<int64_t>(s) * N * E + h * 64;
    bf16* dbase = dqkv + static_cast<int64_t>(s) * N * ld + h * 64;
    const uint32_t qs_u = smem_u32(Qs), ks_u = smem_u32(Ks), vs_u = smem_u32(Vs), gs_u = smem_u32(Gs);

    const int THREADS = blockDim.x, WARPS = blockDim.x >> 5;
    for (int idx = threadIdx.x; idx < NKP * 8; idx += THREADS) {
        const int r = idx >> 3, c = idx & 7;
        const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
        if (r < N) {
            const bf16* g = qbase + r * ld + c * 8;
            cp_async_16(qs_u + off, g);
            cp_async_16(ks_u + off, g + E);
            cp_async_16(vs_u + off, g + 2 * E);
            cp_async_16(gs_u + off, gbase + static_cast<int64_t>(r) * E + c * 8);
        } else {
            const uint4 z = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(Qs + off) = z; *reinterpret_cast<uint4*>(Ks + off) = z;
            *reinterpret_cast<uint4*>(Vs + off) = z; *reinterpret_cast<uint4*>(Gs + off) = z;
        }
    }
    cp_async_commit_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int n_rb = (N + 15) >> 4;       // 16-row blocks
    const int n_ch = NKP >> 5;            // 32-row chunks (NKP is a multiple of 32)

    // ---- phase 1: D_i and lse_i ----
    for (int i = threadIdx.x; i < NKP; i += THREADS) {
        float d = 0.f;
        if (i < N) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 a = *reinterpret_cast<const uint4*>(Gs + i * 128 + ((c ^ (i & 7)) << 4));
                const uint4 b = __ldg(reinterpret_cast<const uint4*>(obase + static_cast<int64_t>(i) * E + c * 8));
                const uint32_t* ai = reinterpret_cast<const uint32_t*>(&a);
                const uint32_t* bi = reinterpret_cast<const uint32_t*>(&b);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 x = unpack_bf16x2(ai[j]), y = unpack_bf16x2(bi[j]);
                    d = fmaf(x.x, y.x, d); d = fmaf(x.y, y.y, d);
                }
            }
        }
        Dv[i] = d;
    }
    if (lse_in != nullptr) {   // the forward kernel kept the row log-sum-exp (attention_tc16.cu, kLse): nothing to recompute
        for (int i = threadIdx.x; i < NKP; i += THREADS) lse[i] = i < N ? lse_in[static_cast<int64_t>(blockIdx.x) * N + i] : CUDART_INF_F;
    }
    for (int rb = lse_in != nullptr ? (NKP >> 4) : warp; rb < (NKP >> 4); rb += WARPS) {
        if (rb >= n_rb) {   // rows beyond N: probabilities are defined as 0
            if (lane < 16) lse[rb * 16 + lane] = CUDART_INF_F;
            continue;
        }
        uint32_t qa[4][4];
        load_a(qa, Qs, rb * 16 + g, rb * 16 + g + 8, t);
        float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;
        for (int ch = 0; ch < n_ch; ++ch) {
            float sacc[4][4];
            mma_rows(sacc, qa, ks_u, ch * 32, lane);
            float cm0 = -CUDART_INF_F, cm1 = -CUDART_INF_F;
            if (ch * 32 + 32 > N) {   // only the last chunk(s) hold padding keys
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const int key = ch * 32 + nb * 8 + 2 * t;
                    if (key >= N) { sacc[nb][0] = -CUDART_INF_F; sacc[nb][2] = -CUDART_INF_F; }
                    if (key + 1 >= N) { sacc[nb][1] = -CUDART_INF_F; sacc[nb][3] = -CUDART_INF_F; }
                }
            }
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                cm0 = fmaxf(cm0, fmaxf(sacc[nb][0], sacc[nb][1]));
                cm1 = fmaxf(cm1, fmaxf(sacc[nb][2], sacc[nb][3]));
            }
            cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1)); cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
            cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1)); cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
            if (cm0 == -CUDART_INF_F && cm1 == -CUDART_INF_F) continue;   // a chunk of padding keys only (warp-uniform per quad pair)
            const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                rs0 += ex2_approx((sacc[nb][0] - mn0) * LOG2E) + ex2_approx((sacc[nb][1] - mn0) * LOG2E);
                rs1 += ex2_approx((sacc[nb][2] - mn1) * LOG2E) + ex2_approx((sacc[nb][3] - mn1) * LOG2E);
            }
            l0 = l0 * ex2_approx((m0 - mn0) * LOG2E) + rs0;
            l1 = l1 * ex2_approx((m1 - mn1) * LOG2E) + rs1;
            m0 = mn0; m1 = mn1;
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        if (t == 0) {
            const int r0 = rb * 16 + g, r1 = r0 + 8;
            lse[r0] = r0 < N ? m0 * LOG2E + log2f(l0) : CUDART_INF_F;
            lse[r1] = r1 < N ? m1 * LOG2E + log2f(l1) : CUDART_INF_F;
        }
    }
    __syncthreads();

    // ---- phases 2 and 3 read the same tiles and write disjoint outputs: ONE pool of 2 n_rb work units over all warps, the longer
    //      dK/dV units first (N = 257: 34 units on 12 warps = 3 rounds instead of 2 + 2 on 9 warps) ----
    for (int u = warp; u < 2 * n_rb; u += WARPS) {
        if (u >= n_rb) {   // dQ of one 16-query block (queries as rows)
            const int rb = u - n_rb;
            const int r0 = rb * 16 + g, r1 = r0 + 8;
            uint32_t qa[4][4], ga[4][4];
            load_a(qa, Qs, r0, r1, t);
            load_a(ga, Gs, r0, r1, t);
            const float ls0 = lse[r0], ls1 = lse[r1], d0 = Dv[r0], d1 = Dv[r1];
            float dq[8][4];
    #pragma unroll
            for (int i = 0; i < 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
            for (int ch = 0; ch < n_ch; ++ch) {
                float sacc[4][4], dp[4][4];
                mma_rows(sacc, qa, ks_u, ch * 32, lane);
                mma_rows(dp, ga, vs_u, ch * 32, lane);
                if (ch * 32 + 32 > N) {   // padding keys (zero K rows give S = 0, not -inf): only in the last chunk(s)
    #pragma unroll
                    for (int nb = 0; nb < 4; ++nb) {
                        const int key = ch * 32 + nb * 8 + 2 * t;
                        if (key >= N) { sacc[nb][0] = -CUDART_INF_F; sacc[nb][2] = -CUDART_INF_F; }
                        if (key + 1 >= N) { sacc[nb][1] = -CUDART_INF_F; sacc[nb][3] = -CUDART_INF_F; }
                    }
                }
    #pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const float p00 = ex2_approx(fmaf(sacc[nb][0], LOG2E, -ls0)), p01 = ex2_approx(fmaf(sacc[nb][1], LOG2E, -ls0));
                    const float p10 = ex2_approx(fmaf(sacc[nb][2], LOG2E, -ls1)), p11 = ex2_approx(fmaf(sacc[nb][3], LOG2E, -ls1));
                    sacc[nb][0] = p00 * (dp[nb][0] - d0); sacc[nb][1] = p01 * (dp[nb][1] - d0);     // dS
                    sacc[nb][2] = p10 * (dp[nb][2] - d1); sacc[nb][3] = p11 * (dp[nb][3] - d1);
                }
                mma_cols(dq, sacc, ks_u, ch * 32, lane);
            }
            if (r0 < N) {
                uint32_t* dst = reinterpret_cast<uint32_t*>(dbase + static_cast<int64_t>(r0) * ld);
    #pragma unroll
                for (int dn = 0; dn < 8; ++dn) dst[dn * 4 + t] = pack_bf16x2(dq[dn][0] * 0.125f, dq[dn][1] * 0.125f);
            }
            if (r1 < N) {
                uint32_t* dst = reinterpret_cast<uint32_t*>(dbase + static_cast<int64_t>(r1) * ld);
    #pragma unroll
                for (int dn = 0; dn < 8; ++dn) dst[dn * 4 + t] = pack_bf16x2(dq[dn][2] * 0.125f, dq[dn][3] * 0.125f);
            }
    
        } else {           // dK, dV of one 16-key block (keys as rows)
            const int kb = u;
            const int r0 = kb * 16 + g, r1 = r0 + 8;
            uint32_t ka[4][4], va[4][4];
            load_a(ka, Ks, r0, r1, t);
            load_a(va, Vs, r0, r1, t);
            float dk[8][4], dv[8][4];
    #pragma unroll
            for (int i = 0; i < 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
            for (int ch = 0; ch < n_ch; ++ch) {
                float st[4][4], dpt[4][4];
                mma_rows(st, ka, qs_u, ch * 32, lane);       // S^T[key, query] = k . q'
                mma_rows(dpt, va, gs_u, ch * 32, lane);      // dP^T[key, query] = v . dO
    #pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const int qi = ch * 32 + nb * 8 + 2 * t;
                    const float la = lse[qi], lb = lse[qi + 1], da = Dv[qi], dbq = Dv[qi + 1];
                    const float p00 = ex2_approx(fmaf(st[nb][0], LOG2E, -la)), p01 = ex2_approx(fmaf(st[nb][1], LOG2E, -lb));   // lse = +inf for padding
                    const float p10 = ex2_approx(fmaf(st[nb][2], LOG2E, -la)), p11 = ex2_approx(fmaf(st[nb][3], LOG2E, -lb));
                    st[nb][0] = p00; st[nb][1] = p01; st[nb][2] = p10; st[nb][3] = p11;
                    dpt[nb][0] = p00 * (dpt[nb][0] - da); dpt[nb][1] = p01 * (dpt[nb][1] - dbq);
                    dpt[nb][2] = p10 * (dpt[nb][2] - da); dpt[nb][3] = p11 * (dpt[nb][3] - dbq);
                }
                mma_cols(dv, st, gs_u, ch * 32, lane);       // dV += P^T dO
                mma_cols(dk, dpt, qs_u, ch * 32, lane);      // dK += dS^T q'
            }
            if (r0 < N) {
                uint32_t* dK = reinterpret_cast<uint32_t*>(dbase + static_cast<int64_t>(r0) * ld + E);
                uint32_t* dV = reinterpret_cast<uint32_t*>(dbase + static_cast<int64_t>(r0) * ld + 2 * E);
    #pragma unroll
                for (int dn = 0; dn < 8; ++dn) { dK[dn * 4 + t] = pack_bf16x2(dk[dn][0], dk[dn][1]); dV[dn * 4 + t] = pack_bf16x2(dv[dn][0], dv[dn][1]); }
            }
            if (r1 < N) {
                uint32_t* dK = reinterpret_cast<uint32_t*>(dbase + static_cast<int64_t>(r1) * ld + E);
                uint32_t* dV = reinterpret_cast<uint32_t*>(dbase + static_cast<int64_t>(r1) * ld + 2 * E);
    #pragma unroll
                for (int dn = 0; dn < 8; ++dn) { dK[dn * 4 + t] = pack_bf16x2(dk[dn][2], dk[dn][3]); dV[dn * 4 + t] = pack_bf16x2(dv[dn][2], dv[dn][3]); }
            }
    
        }
    }
}

int launch_attention_bwd(const bf16* qkv, const bf16* o, const bf16* dO, bf16* dqkv, int BD, int N, int heads, cudaStream_t stream,
                         const float* lse) {
    const int NKP = ((N + 31) / 32) * 32;
    const size_t smem = static_cast<size_t>(NKP) * (4 * 128 + 8);
    MST_REQUIRE(smem <= 227 * 1024, "attention backward: N=%d tokens do not fit shared memory", N);
    MST_SET_DYN_SMEM(attention_bwd_kernel, 227 * 1024);
    int warps = 2 * ((N + 15) / 16);      // one pool of 2 x (row blocks) units for the dQ and dK/dV phases
    warps = warps < 4 ? 4 : (warps > attb::MAX_WARPS ? attb::MAX_WARPS : warps);
    attention_bwd_kernel<<<BD * heads, warps * 32, smem, stream>>>(qkv, o, dO, dqkv, lse, N, heads, NKP);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mst
