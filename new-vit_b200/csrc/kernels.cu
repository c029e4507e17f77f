// HBM-bound and small kernels of the MST-DINOv2 path: im2col(+CLS row), LayerNorm, CLS-row attention,
// fp32 CUDA-core attention and GEMM (fp32 parity mode), the slice transformer, and the saliency
// combiner + upsampler.  All arithmetic that the reference does in fp32 statistics (LN mean/var,
// softmax) is fp32 here too.
#include <math_constants.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "ptx.cuh"

namespace mst {

// ---------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum; `red` must hold >= 33 floats; all threads get the result
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nw ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nw ? red[lane] : -CUDART_INF_F;
        t = warp_max(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4<bf16> {
    static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
        uint2 t;
        t.x = pack_bf16x2(v[0], v[1]);
        t.y = pack_bf16x2(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = t;
    }
};

// ---------------------------------------------------------------------------------------------------
// im2col: src [BD,H,W] (fp32, bf16 or fp16) -> A0 [BD*P, KP] (column ky*14+kx, zero padded to KP), plus the CLS token
// row of every slice: x[s*NT] = cls_token + pos_embed[0], followed by the R register tokens (no position
// embedding, vision_transformer.py:222-230); NT = 1 + R + P.  Replaces the rearrange + 3x repeat +
// conv unfold of reference dino.py:125-127 / patch_embed.py:75-77 (the RGB copies are never
// materialised: the conv weight is channel-summed at pack time).
// Test-time augmentation (scripts/main_predict.py:147-149): with tta_BD = B*D > 0 the grid covers 8*B*D VIRTUAL slices,
// variant v = s / (B*D) of torch.flip(source, dims) for dims in [(), (2,), (3,), (4,), (2,3), (2,4), (3,4), (2,3,4)];
// the flips are index arithmetic on the load, the flipped volumes are never materialised.
// ---------------------------------------------------------------------------------------------------
template <typename TS> __device__ __forceinline__ float src_to_f(TS v);
template <> __device__ __forceinline__ float src_to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float src_to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float src_to_f<__half>(__half v) { return __half2float(v); }

template <typename TS, typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const TS* __restrict__ src, T* __restrict__ A0, T* __restrict__ x,
                                                      const float* __restrict__ cls_pos0, const float* __restrict__ regs, int R,
                                                      int H, int W, int KP, int E, int tta_BD, int D) {
    extern __shared__ float tile[];  // [14][W]
    griddep_launch_dependents();
    const int gh = H / 14, gw = W / 14, P = gh * gw;
    const int s = blockIdx.x / gh, py = blockIdx.x % gh;
    int ss = s;
    bool flip_h = false, flip_w = false;
    if (tta_BD > 0) {
        const int v = s / tta_BD, rem = s - v * tta_BD, b = rem / D, d = rem - b * D;
        const bool flip_d = (0xB2 >> v) & 1;   // variants 1, 4, 5, 7 flip dim 2 (depth)
        flip_h = (0xD4 >> v) & 1;              // variants 2, 4, 6, 7 flip dim 3 (height)
        flip_w = (0xE8 >> v) & 1;              // variants 3, 5, 6, 7 flip dim 4 (width)
        ss = b * D + (flip_d ? D - 1 - d : d);
    }
    if (flip_h || flip_w) {
        const TS* sl = src + static_cast<int64_t>(ss) * H * W;
        for (int i = threadIdx.x; i < 14 * W; i += blockDim.x) {
            const int r = i / W, xc = i - r * W;
            const int yy = flip_h ? H - 1 - (py * 14 + r) : py * 14 + r;
            const int xx = flip_w ? W - 1 - xc : xc;
            tile[i] = src_to_f<TS>(sl[static_cast<int64_t>(yy) * W + xx]);
        }
    } else {
        const TS* base = src + (static_cast<int64_t>(ss) * H + py * 14) * W;
        if (sizeof(TS) == 4 && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            // 16-byte loads (the 14-row band starts on a 16-byte boundary when W % 4 == 0)
            const float4* b4 = reinterpret_cast<const float4*>(base);
            float4* t4 = reinterpret_cast<float4*>(tile);
            for (int i = threadIdx.x; i < 14 * W / 4; i += blockDim.x) t4[i] = __ldg(b4 + i);
        } else if (sizeof(TS) == 2 && (W & 7) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const uint4* b8 = reinterpret_cast<const uint4*>(base);   // 8 two-byte voxels per load
            for (int i = threadIdx.x; i < 14 * W / 8; i += blockDim.x) {
                const uint4 u = __ldg(b8 + i);
                const TS* e8 = reinterpret_cast<const TS*>(&u);
                float4 lo = make_float4(src_to_f<TS>(e8[0]), src_to_f<TS>(e8[1]), src_to_f<TS>(e8[2]), src_to_f<TS>(e8[3]));
                float4 hi = make_float4(src_to_f<TS>(e8[4]), src_to_f<TS>(e8[5]), src_to_f<TS>(e8[6]), src_to_f<TS>(e8[7]));
                reinterpret_cast<float4*>(tile)[2 * i] = lo;
                reinterpret_cast<float4*>(tile)[2 * i + 1] = hi;
            }
        } else {
            for (int i = threadIdx.x; i < 14 * W; i += blockDim.x) tile[i] = src_to_f<TS>(base[i]);
        }
    }
    __syncthreads();
    // thread -> 8 consecutive columns (taps) of a patch row, written as ONE 16-byte (bf16) / two 16-byte (fp32) stores;
    // the tap -> pixel offsets depend only on the thread's column chunk and are computed once
    const int chunks = KP >> 3;                       // 8-column chunks per patch row (KP % 8 == 0)
    const int ch = threadIdx.x % chunks, pp = threadIdx.x / chunks, pstep = blockDim.x / chunks;
    int offs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = ch * 8 + j;
        offs[j] = c < 196 ? (c / 14) * W + (c % 14) : -1;
    }
    if (pp < pstep) {
        for (int px = pp; px < gw; px += pstep) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = offs[j] >= 0 ? tile[offs[j] + px * 14] : 0.f;
            T* o = A0 + (static_cast<int64_t>(s) * P + py * gw + px) * KP + ch * 8;
            if constexpr (sizeof(T) == 2) {
                *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                                          pack_bf16x2(v[6], v[7]));
            } else {
                float lo[4] = {v[0], v[1], v[2], v[3]}, hi[4] = {v[4], v[5], v[6], v[7]};
                Vec4<T>::store(o, lo);
                Vec4<T>::store(o + 4, hi);
            }
        }
    }
    if (py == 0) {
        T* xrow = x + static_cast<int64_t>(s) * (P + 1 + R) * E;
        for (int e = threadIdx.x; e < E; e += blockDim.x) xrow[e] = from_f<T>(cls_pos0[e]);
        for (int e = threadIdx.x; e < R * E; e += blockDim.x) xrow[E + e] = from_f<T>(regs[e]);
    }
}

// src_dtype: 0 fp32, 1 bf16, 2 fp16 (MST_SRC_*).  BD counts the slices the grid covers (8x the real count with TTA).
template <typename T>
int launch_im2col(const void* src, int src_dtype, T* A0, T* x, const float* cls_pos0, const float* regs, int R, int BD, int H, int W,
                  int KP, int E, int tta_BD, int D, cudaStream_t stream) {
    const int gh = H / 14;
    const size_t smem = 14 * W * sizeof(float);
    const dim3 grid(BD * gh);
    if (src_dtype == 0)
        im2col_kernel<float, T><<<grid, 256, smem, stream>>>(static_cast<const float*>(src), A0, x, cls_pos0, regs, R, H, W, KP, E, tta_BD, D);
    else if (src_dtype == 1)
        im2col_kernel<bf16, T><<<grid, 256, smem, stream>>>(static_cast<const bf16*>(src), A0, x, cls_pos0, regs, R, H, W, KP, E, tta_BD, D);
    else if (src_dtype == 2)
        im2col_kernel<__half, T><<<grid, 256, smem, stream>>>(static_cast<const __half*>(src), A0, x, cls_pos0, regs, R, H, W, KP, E, tta_BD, D);
    else
        MST_REQUIRE(false, "im2col: unknown source dtype %d", src_dtype);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}
template int launch_im2col<float>(const void*, int, float*, float*, const float*, const float*, int, int, int, int, int, int, int, int, cudaStream_t);
template int launch_im2col<bf16>(const void*, int, bf16*, bf16*, const float*, const float*, int, int, int, int, int, int, int, int, cudaStream_t);

// ---------------------------------------------------------------------------------------------------
// LayerNorm over rows of E (E % 128 == 0): one warp per row, two-pass in registers, fp32 statistics.
// (reference block.py:91,94 eps 1e-6; vision_transformer.py:263)
// ---------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut, int E>
__global__ void __launch_bounds__(256) layernorm_kernel(const TIn* __restrict__ x, int64_t ldx, TOut* __restrict__ y,
                                                         int64_t ldy, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int rows, float eps) {
    constexpr int V = E / 128;
    griddep_launch_dependents();   // a GEMM launched as a programmatic dependent loads its weights while this kernel drains
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const TIn* xr = x + static_cast<int64_t>(row) * ldx;
    float v[V][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        Vec4<TIn>::load(xr + i * 128 + lane * 4, v[i]);
        s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
    const float mean = warp_sum(s) * (1.0f / E);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float d = v[i][j] - mean;
            q = fmaf(d, d, q);
        }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / E) + eps);
    TOut* yr = y + static_cast<int64_t>(row) * ldy;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c = i * 128 + lane * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
        float o[4];
        o[0] = fmaf((v[i][0] - mean) * rstd, g.x, b.x);
        o[1] = fmaf((v[i][1] - mean) * rstd, g.y, b.y);
        o[2] = fmaf((v[i][2] - mean) * rstd, g.z, b.z);
        o[3] = fmaf((v[i][3] - mean) * rstd, g.w, b.w);
        Vec4<TOut>::store(yr + c, o);
    }
}

template <typename TIn, typename TOut>
int launch_layernorm(const TIn* x, int64_t ldx, TOut* y, int64_t ldy, const float* gamma, const float* beta, int rows,
                     int E, float eps, cudaStream_t stream) {
    if (rows <= 0) return 0;
    const int grid = (rows + 7) / 8;
    if (E == 384) layernorm_kernel<TIn, TOut, 384><<<grid, 256, 0, stream>>>(x, ldx, y, ldy, gamma, beta, rows, eps);
    else if (E == 768) layernorm_kernel<TIn, TOut, 768><<<grid, 256, 0, stream>>>(x, ldx, y, ldy, gamma, beta, rows, eps);
    else if (E == 1024) layernorm_kernel<TIn, TOut, 1024><<<grid, 256, 0, stream>>>(x, ldx, y, ldy, gamma, beta, rows, eps);
    else MST_REQUIRE(false, "layernorm: unsupported embed dim %d", E);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}
template int launch_layernorm<float, float>(const float*, int64_t, float*, int64_t, const float*, const float*, int, int, float, cudaStream_t);
template int launch_layernorm<bf16, bf16>(const bf16*, int64_t, bf16*, int64_t, const float*, const float*, int, int, float, cudaStream_t);
template int launch_layernorm<bf16, float>(const bf16*, int64_t, float*, int64_t, const float*, const float*, int, int, float, cudaStream_t);

// Row statistics for the LayerNorm that is folded into the consuming GEMM (bf16 path): reads x once, writes 4 bytes
// per row instead of a normalised copy.  Same two-pass fp32 arithmetic as layernorm_kernel.
template <int E>
__global__ void __launch_bounds__(256) row_stats_kernel(const bf16* __restrict__ x, float* __restrict__ rowstat, int rows,
                                                         float eps) {
    // half a warp per row, 16-byte loads: lane l of the half-warp owns the chunks l, l+16, ... of the row
    constexpr int V = E / 128;
    griddep_launch_dependents();
    const int row = blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4);
    const int hl = threadIdx.x & 15;
    const bool ok = row < rows;
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(ok ? row : 0) * E);
    float v[V][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const uint4 u = __ldg(xr + i * 16 + hl);
        float2 f;
        f = unpack_bf16x2(u.x); v[i][0] = f.x; v[i][1] = f.y;
        f = unpack_bf16x2(u.y); v[i][2] = f.x; v[i][3] = f.y;
        f = unpack_bf16x2(u.z); v[i][4] = f.x; v[i][5] = f.y;
        f = unpack_bf16x2(u.w); v[i][6] = f.x; v[i][7] = f.y;
        s += ((v[i][0] + v[i][1]) + (v[i][2] + v[i][3])) + ((v[i][4] + v[i][5]) + (v[i][6] + v[i][7]));
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / E);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = v[i][j] - mean;
            q = fmaf(d, d, q);
        }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (ok && hl == 0) rowstat[row] = rsqrtf(q * (1.0f / E) + eps);
}

int launch_row_stats(const bf16* x, float* rowstat, int rows, int E, float eps, cudaStream_t stream) {
    if (rows <= 0) return 0;
    const int grid = (rows + 15) / 16;
    if (E == 384) row_stats_kernel<384><<<grid, 256, 0, stream>>>(x, rowstat, rows, eps);
    else if (E == 768) row_stats_kernel<768><<<grid, 256, 0, stream>>>(x, rowstat, rows, eps);
    else if (E == 1024) row_stats_kernel<1024><<<grid, 256, 0, stream>>>(x, rowstat, rows, eps);
    else MST_REQUIRE(false, "row stats: unsupported embed dim %d", E);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// CLS-row attention of one encoder block (head_dim 64): for (slice, head) the CLS query against all N
// keys.  Used for the LAST block, whose other query rows are dead (only x[:,0] is consumed,
// vision_transformer.py:265,329), and it is the side output the saliency path needs: row 0 of the
// last block's attention (reference dino.py:190-192 reads attention_maps[-1][:, :, 0, :]).
// q is pre-scaled (0.125 folded into Wq,bq at pack time; attention.py:60).
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) cls_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out_cls,
                                                             float* __restrict__ plane_cls, int N, int heads) {
    extern __shared__ float sm[];  // p[N] | q[64] | red[40] | part[128]
    float* p = sm;
    float* q = sm + N;
    float* red = q + 64;
    float* part = red + 40;
    const int s = blockIdx.x / heads, h = blockIdx.x % heads;
    const int E = heads * 64;
    const int64_t ld = 3 * E;
    const T* base = qkv + static_cast<int64_t>(s) * N * ld;
    if (threadIdx.x < 64) q[threadIdx.x] = to_f<T>(base[h * 64 + threadIdx.x]);
    __syncthreads();
    float lmax = -CUDART_INF_F;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const T* kr = base + j * ld + E + h * 64;
        float acc = 0.f, acc2 = 0.f;
        if constexpr (sizeof(T) == 2) {   // 16-byte loads, two accumulation chains
#pragma unroll
            for (int d = 0; d < 64; d += 8) {
                const uint4 u = *reinterpret_cast<const uint4*>(kr + d);
                float2 f;
                f = unpack_bf16x2(u.x); acc = fmaf(q[d], f.x, acc); acc2 = fmaf(q[d + 1], f.y, acc2);
                f = unpack_bf16x2(u.y); acc = fmaf(q[d + 2], f.x, acc); acc2 = fmaf(q[d + 3], f.y, acc2);
                f = unpack_bf16x2(u.z); acc = fmaf(q[d + 4], f.x, acc); acc2 = fmaf(q[d + 5], f.y, acc2);
                f = unpack_bf16x2(u.w); acc = fmaf(q[d + 6], f.x, acc); acc2 = fmaf(q[d + 7], f.y, acc2);
            }
            acc += acc2;
        } else {
#pragma unroll
            for (int d = 0; d < 64; d += 4) {
                float kv[4];
                Vec4<T>::load(kr + d, kv);
                acc = fmaf(q[d], kv[0], acc); acc = fmaf(q[d + 1], kv[1], acc);
                acc = fmaf(q[d + 2], kv[2], acc); acc = fmaf(q[d + 3], kv[3], acc);
            }
        }
        p[j] = acc;
        lmax = fmaxf(lmax, acc);
    }
    const float m = block_max(lmax, red);
    float lsum = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float e = expf(p[j] - m);
        p[j] = e;
        lsum += e;
    }
    const float inv = 1.0f / block_sum(lsum, red);
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float pj = p[j] * inv;
        p[j] = pj;
        if (plane_cls) plane_cls[(static_cast<int64_t>(s) * heads + h) * N + j] = pj;
    }
    __syncthreads();
    const int d = threadIdx.x & 63, half = threadIdx.x >> 6;
    float acc = 0.f;
    {
        const T* vp = base + 2 * E + h * 64 + d;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // four independent chains: the loop is latency-bound otherwise
        int j = half;
        for (; j + 6 < N; j += 8) {
            a0 = fmaf(p[j], to_f<T>(vp[static_cast<int64_t>(j) * ld]), a0);
            a1 = fmaf(p[j + 2], to_f<T>(vp[static_cast<int64_t>(j + 2) * ld]), a1);
            a2 = fmaf(p[j + 4], to_f<T>(vp[static_cast<int64_t>(j + 4) * ld]), a2);
            a3 = fmaf(p[j + 6], to_f<T>(vp[static_cast<int64_t>(j + 6) * ld]), a3);
        }
        for (; j < N; j += 2) a0 = fmaf(p[j], to_f<T>(vp[static_cast<int64_t>(j) * ld]), a0);
        acc = (a0 + a1) + (a2 + a3);
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < 64) out_cls[static_cast<int64_t>(s) * E + h * 64 + d] = from_f<T>(part[d] + part[64 + d]);
}

template <typename T>
int launch_cls_attention(const T* qkv, T* out_cls, float* plane_cls, int BD, int N, int heads, cudaStream_t stream) {
    const size_t smem = (N + 64 + 40 + 128) * sizeof(float);
    cls_attention_kernel<T><<<BD * heads, 128, smem, stream>>>(qkv, out_cls, plane_cls, N, heads);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}
template int launch_cls_attention<float>(const float*, float*, float*, int, int, int, cudaStream_t);
template int launch_cls_attention<bf16>(const bf16*, bf16*, float*, int, int, int, cudaStream_t);

// ---------------------------------------------------------------------------------------------------
// fp32 CUDA-core attention (fp32 parity mode only): one CTA per (slice, head), K and V in shared memory.
// softmax(q k^T) v with q pre-scaled (attention.py:56-69).  N <= 431 (K and V of one head in 227 KB).
// ---------------------------------------------------------------------------------------------------
constexpr int ATT32_WARPS = 8;
constexpr int ATT32_MAXJ = 14;   // 32-key register groups per query row: N <= 448 by registers, <= 431 by shared memory
__global__ void __launch_bounds__(ATT32_WARPS * 32) attention_f32_kernel(const float* __restrict__ qkv,
                                                                          float* __restrict__ out, int N, int heads) {
    extern __shared__ float sm[];
    float* Ks = sm;                 // [N][65]
    float* Vs = Ks + N * 65;        // [N][64]
    float* qs = Vs + N * 64;        // [warps][64]
    float* ps = qs + ATT32_WARPS * 64;  // [warps][N]
    const int s = blockIdx.x / heads, h = blockIdx.x % heads;
    const int E = heads * 64;
    const int64_t ld = 3 * E;
    const float* base = qkv + static_cast<int64_t>(s) * N * ld;
    for (int i = threadIdx.x; i < N * 64; i += blockDim.x) {
        const int j = i >> 6, d = i & 63;
        Ks[j * 65 + d] = base[j * ld + E + h * 64 + d];
        Vs[j * 64 + d] = base[j * ld + 2 * E + h * 64 + d];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* q = qs + warp * 64;
    float* p = ps + warp * N;
    for (int r = warp; r < N; r += ATT32_WARPS) {
        q[lane] = base[r * ld + h * 64 + lane];
        q[lane + 32] = base[r * ld + h * 64 + lane + 32];
        __syncwarp();
        float sc[ATT32_MAXJ];
        float lmax = -CUDART_INF_F;
#pragma unroll
        for (int jj = 0; jj < ATT32_MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            float acc = -CUDART_INF_F;
            if (j < N) {
                acc = 0.f;
                const float* kr = Ks + j * 65;
#pragma unroll 16
                for (int d = 0; d < 64; ++d) acc = fmaf(q[d], kr[d], acc);
            }
            sc[jj] = acc;
            lmax = fmaxf(lmax, acc);
        }
        const float m = warp_max(lmax);
        float lsum = 0.f;
#pragma unroll
        for (int jj = 0; jj < ATT32_MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            const float e = j < N ? expf(sc[jj] - m) : 0.f;
            sc[jj] = e;
            lsum += e;
        }
        const float inv = 1.0f / warp_sum(lsum);
#pragma unroll
        for (int jj = 0; jj < ATT32_MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            if (j < N) p[j] = sc[jj] * inv;
        }
        __syncwarp();
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j < N; ++j) {
            const float pj = p[j];
            a0 = fmaf(pj, Vs[j * 64 + lane], a0);
            a1 = fmaf(pj, Vs[j * 64 + lane + 32], a1);
        }
        float* o = out + (static_cast<int64_t>(s) * N + r) * E + h * 64;
        o[lane] = a0;
        o[lane + 32] = a1;
        __syncwarp();
    }
}

int launch_attention_f32(const float* qkv, float* out, int BD, int N, int heads, cudaStream_t stream) {
    MST_REQUIRE(N <= ATT32_MAXJ * 32, "fp32 attention supports at most %d tokens per slice (got %d)", ATT32_MAXJ * 32, N);
    const size_t smem = (static_cast<size_t>(N) * 65 + N * 64 + ATT32_WARPS * 64 + ATT32_WARPS * N) * sizeof(float);
    MST_REQUIRE(smem <= 227 * 1024, "fp32 attention: %zu bytes of shared memory needed", smem);
    MST_SET_DYN_SMEM(attention_f32_kernel, 227 * 1024);
    attention_f32_kernel<<<BD * heads, ATT32_WARPS * 32, smem, stream>>>(qkv, out, N, heads);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// fp32 CUDA-core GEMM (fp32 parity mode only): 64x64 tile, BK 16, 4x4 micro-tile, same epilogues.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W,
                                                        int M, int N, int K, int mode, EpiParams ep) {
    __shared__ float As[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + lr < M) a = *reinterpret_cast<const float4*>(A + static_cast<int64_t>(m0 + lr) * lda + k0 + lk);
        const float4 w = *reinterpret_cast<const float4*>(W + static_cast<int64_t>(n0 + lr) * K + k0 + lk);
        As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
        Ws[lk + 0][lr] = w.x; Ws[lk + 1][lr] = w.y; Ws[lk + 2][lr] = w.z; Ws[lk + 3][lr] = w.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 wv = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            const float a4[4] = {av.x, av.y, av.z, av.w};
            const float w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], w4[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = m0 + ty * 4 + i;
        if (row >= M) continue;
        const int n = n0 + tx * 4;
        float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
        int64_t orow = row;
        if (mode == EPI_PATCH) {
            const int p = static_cast<int>(row % ep.P);
            orow = (row / ep.P) * (ep.P + 1 + ep.R) + 1 + ep.R + p;
            const float4 b = *reinterpret_cast<const float4*>(ep.posb + static_cast<int64_t>(p) * N + n);
            v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        } else {
            const float4 b = *reinterpret_cast<const float4*>(ep.bias + n);
            v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        }
        if (mode == EPI_BIAS_GELU) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = gelu_erf<true>(v[j]);
        }
        if (mode == EPI_BIAS_RES) {
            const float4 r = *reinterpret_cast<const float4*>(static_cast<const float*>(ep.res) + row * ep.ldr + n);
            v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        *reinterpret_cast<float4*>(static_cast<float*>(ep.out) + orow * ep.ldo + n) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

int gemm_f32_simt(const float* A, int64_t lda, const float* W, int M, int N, int K, int mode, const EpiParams& ep,
                  cudaStream_t stream) {
    MST_REQUIRE(M > 0 && N % 64 == 0 && K % 16 == 0 && lda % 4 == 0, "fp32 gemm: bad shape M=%d N=%d K=%d", M, N, K);
    dim3 grid(N / 64, (M + 63) / 64);
    MST_REQUIRE(grid.y <= 65535, "fp32 gemm: M=%d too large for the parity path", M);
    gemm_f32_kernel<<<grid, 256, 0, stream>>>(A, lda, W, M, N, K, mode, ep);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Slice transformer + head, one CTA per volume (reference dino.py:145-166,
// utils/transformer_blocks.py:524-587 norm_first branch, :29-318 attention).
//
// Only token 0 (the slice-CLS) of the single layer's output is consumed (dino.py:153), so only the CLS
// query is evaluated, and K/V are never materialised:   s[h,j] = q_h.(Wk_h n_j + bk_h) = (Wk_h^T q_h).n_j + q_h.bk_h
// and   o_h = sum_j p[h,j] (Wv_h n_j + bv_h) = Wv_h (sum_j p[h,j] n_j) + bv_h   (softmax rows sum to 1).
// Exact re-association of the reference's arithmetic, all fp32.  Emits p[h,:] = row 0 of the slice attention
// (what get_slice_attention reads, dino.py:174-175).  Weights are fp32, pre-transposed to [in][out].
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_layernorm(const float* in, float* out, const float* g, const float* b, int E,
                                                float eps, float* red) {
    float s = 0.f;
    for (int i = threadIdx.x; i < E; i += blockDim.x) s += in[i];
    const float mean = block_sum(s, red) / E;
    float q = 0.f;
    for (int i = threadIdx.x; i < E; i += blockDim.x) { const float d = in[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(block_sum(q, red) / E + eps);
    for (int i = threadIdx.x; i < E; i += blockDim.x) out[i] = fmaf((in[i] - mean) * rstd, g[i], b[i]);
    __syncthreads();
}
// out[n] = act((sum_k in_n[k] * Wt[k*ldw + n] + bias[n]) * scale) (+ res[n]);  in / out / scr in shared memory.
// in_n = in + (n / in_group) * in_stride  (in_group = Nout: one input vector; in_group = head_dim: one input row per head).
// A matrix-vector product on ONE SM is bound by load latency over loads in flight, so the K range is split MV_PARTS ways across
// the CTA and every thread streams float4 weight columns with eight independent 16-byte loads in flight; the MV_PARTS partials
// are summed in a fixed order (results do not depend on the launch shape).  Nout % 4 == 0, ldw % 4 == 0, K % MV_PARTS == 0.
constexpr int MV_PARTS = 8;
__device__ __forceinline__ void block_matvec(const float* in, int in_group, int in_stride, const float* __restrict__ Wt, int ldw,
                                             const float* __restrict__ bias, const float* res, float* out, float* scr, int K, int Nout,
                                             bool relu, float scale) {
    const int groups = Nout >> 2, kper = K / MV_PARTS;
    for (int task = threadIdx.x; task < groups * MV_PARTS; task += blockDim.x) {
        const int g = task % groups, part = task / groups;
        const int n = g << 2;
        const float* x = in + (n / in_group) * in_stride + part * kper;
        const float4* wp = reinterpret_cast<const float4*>(Wt + static_cast<int64_t>(part) * kper * ldw + n);
        const int64_t step = ldw >> 2;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int k = 0;
        for (; k + 7 < kper; k += 8) {
            float4 wv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) wv[u] = __ldg(wp + (k + u) * step);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float xv = x[k + u];
                acc.x = fmaf(xv, wv[u].x, acc.x); acc.y = fmaf(xv, wv[u].y, acc.y);
                acc.z = fmaf(xv, wv[u].z, acc.z); acc.w = fmaf(xv, wv[u].w, acc.w);
            }
        }
        for (; k < kper; ++k) {
            const float4 w1 = __ldg(wp + k * step);
            const float xv = x[k];
            acc.x = fmaf(xv, w1.x, acc.x); acc.y = fmaf(xv, w1.y, acc.y); acc.z = fmaf(xv, w1.z, acc.z); acc.w = fmaf(xv, w1.w, acc.w);
        }
        *reinterpret_cast<float4*>(scr + part * Nout + n) = acc;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < Nout; n += blockDim.x) {
        float a = 0.f;
#pragma unroll
        for (int p = 0; p < MV_PARTS; ++p) a += scr[p * Nout + n];
        float v = (a + bias[n]) * scale;
        if (relu) v = fmaxf(v, 0.f);
        if (res) v += res[n];
        out[n] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(768) slice_fusion_kernel(const float* __restrict__ enc_cls, const uint8_t* __restrict__ pad_mask,
                                                            SliceWeights w, float* __restrict__ hs_all, float* __restrict__ logits,
                                                            float* __restrict__ feat, float* __restrict__ slice_cls, int D, int Eenc,
                                                            int E, int heads, int out_ch, int mode, int mask_period) {
    extern __shared__ float sm[];
    const int L = mode == SLICE_FUSION_TRANSFORMER ? D + 1 : D, l0 = L - D, hd = E / heads;
    float* x0 = sm;                 // [E]  raw CLS token (residual)
    float* n0 = x0 + E;             // [E]  LN1(token 0)
    float* q = n0 + E;              // [E]
    float* qk = q + E;              // [heads][E]
    float* hbar = qk + heads * E;   // [heads][E]
    float* t0 = hbar + heads * E;   // [E]
    float* t1 = t0 + E;             // [E]
    float* t2 = t1 + E;             // [E]
    float* cterm = t2 + E;          // [heads] (padded to 32)
    float* red = cterm + 32;        // [40]
    float* scr = red + 40;          // [MV_PARTS][E] partial sums of the split-K matrix-vector products
    float* p = scr + MV_PARTS * E;  // [heads][L]
    float* kmat = p + ((heads * L + 3) & ~3);   // [L][E] projected keys, RoPE only
    const int b = blockIdx.x;
    // TTA: the 8 flipped variants of a volume all get the volume's UN-flipped padding mask (main_predict.py:149)
    const int bm = mask_period > 0 ? b % mask_period : b;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float* hs = hs_all + static_cast<int64_t>(b) * (D + 1) * E;

    // 0. slice tokens: optional bottleneck Linear(Eenc -> E) (dino.py:134-135), optional slice position embedding
    //    (dino.py:140-142), the slice-CLS token in front for the transformer (dino.py:145)
    for (int l = warp; l < L; l += nwarps) {
        float* dst = hs + static_cast<int64_t>(l) * E;
        if (l < l0) {
            for (int i = lane; i < E; i += 32) dst[i] = w.cls_token[i];
            continue;
        }
        const int d = l - l0;
        const float* src = enc_cls + (static_cast<int64_t>(b) * D + d) * Eenc;
        for (int i = lane; i < E; i += 32) {
            float v;
            if (w.bott_wt) {
                float a0 = 0.f, a1 = 0.f;
                for (int k = 0; k < Eenc; k += 2) {
                    a0 = fmaf(src[k], __ldg(w.bott_wt + static_cast<int64_t>(k) * E + i), a0);
                    a1 = fmaf(src[k + 1], __ldg(w.bott_wt + static_cast<int64_t>(k + 1) * E + i), a1);
                }
                v = (a0 + a1) + w.bott_b[i];
            } else {
                v = src[i];
            }
            if (w.pos_emb) v += w.pos_emb[static_cast<int64_t>(d) * E + i];
            dst[i] = v;
        }
    }
    __syncthreads();

    if (mode != SLICE_FUSION_TRANSFORMER) {
        // 'linear': flatten [D*E] (dino.py:154-155); 'average': mean over slices (dino.py:156-157); the padding mask is
        // not consulted by either (same as the reference)
        const int F = mode == SLICE_FUSION_LINEAR ? D * E : E;
        if (mode == SLICE_FUSION_AVERAGE) {
            for (int i = threadIdx.x; i < E; i += blockDim.x) {
                float a = 0.f;
                for (int d = 0; d < D; ++d) a += hs[static_cast<int64_t>(d) * E + i];
                hs[i] = a / D;  // row 0 is only read by this thread at column i
            }
            __syncthreads();
        }
        if (feat)
            for (int i = threadIdx.x; i < F; i += blockDim.x) feat[static_cast<int64_t>(b) * F + i] = hs[i];
        if (logits)
            for (int c = warp; c < out_ch; c += nwarps) {
                float a = 0.f;
                for (int k = lane; k < F; k += 32) a = fmaf(hs[k], w.head_wt[static_cast<int64_t>(k) * out_ch + c], a);
                a = warp_sum(a);
                if (lane == 0) logits[static_cast<int64_t>(b) * out_ch + c] = a + w.head_b[c];
            }
        return;
    }

    // 1. hs[l] = LN1(token l), in place                                 (transformer_blocks.py:566)
    for (int l = warp; l < L; l += nwarps) {
        float* src = hs + static_cast<int64_t>(l) * E;
        float s = 0.f;
        for (int i = lane; i < E; i += 32) s += src[i];
        const float mean = warp_sum(s) / E;
        float v = 0.f;
        for (int i = lane; i < E; i += 32) { const float d = src[i] - mean; v = fmaf(d, d, v); }
        const float rstd = rsqrtf(warp_sum(v) / E + 1e-5f);
        for (int i = lane; i < E; i += 32) {
            const float raw = src[i];
            const float o = fmaf((raw - mean) * rstd, w.n1w[i], w.n1b[i]);
            src[i] = o;
            if (l == 0) { n0[i] = o; x0[i] = raw; }
        }
    }
    __syncthreads();
    // 2. q = (Wq n0 + bq) / sqrt(hd)                                    (transformer_blocks.py:166,268)
    block_matvec(n0, E, 0, w.in_wt, 3 * E, w.in_b, nullptr, q, scr, E, E, false, rsqrtf(static_cast<float>(hd)));
    if (w.rope_freqs || w.liere) {
        // RoPE (transformer_blocks.py:262-264; rotary_embedding_torch.py:159-173,45-62): the key of sequence position j is
        // rotated by j*freqs[i] in each feature pair (2i, 2i+1) of its head before the dot product, so the keys are
        // materialised here (the W_k^T q re-association below needs position-independent keys).  The slice-CLS query sits
        // at position 0, where the rotation is the identity.
        // 3'. K[j][n] = hs[j] . Wk[n] + bk[n]
        for (int n = threadIdx.x; n < E; n += blockDim.x) {
            for (int j0 = 0; j0 < L; j0 += 11) {
                float acc[11];
#pragma unroll
                for (int jj = 0; jj < 11; ++jj) acc[jj] = 0.f;
                for (int k = 0; k < E; ++k) {
                    const float wv = __ldg(w.in_wt + static_cast<int64_t>(k) * 3 * E + E + n);
#pragma unroll
                    for (int jj = 0; jj < 11; ++jj)
                        if (j0 + jj < L) acc[jj] = fmaf(hs[static_cast<int64_t>(j0 + jj) * E + k], wv, acc[jj]);
                }
#pragma unroll
                for (int jj = 0; jj < 11; ++jj)
                    if (j0 + jj < L) kmat[(j0 + jj) * E + n] = acc[jj] + w.in_b[E + n];
            }
        }
        __syncthreads();
        if (w.liere) {
            // LiRE as the reference evaluates it (rotary_embedding_torch.py:346-396 + the caller's view, transformer_blocks.py:263;
            // batch 1 and L = 33 only, everything else raises there): q and k of every (position, head) are multiplied by ONE
            // orthogonal matrix -- which cancels in q . k and is therefore not applied -- and the [hd, L, heads] result is re-read
            // as [heads, L, hd]: slot (head i, position j) holds the vector of position (L i + j) / heads, head (L i + j) % heads.
            // The slice-CLS query of head i is thus the query of token (L i) / heads in head (L i) % heads.
            const float scale = rsqrtf(static_cast<float>(hd));
            for (int n = threadIdx.x; n < E; n += blockDim.x) {
                const int i = n / hd, d = n - i * hd, m = L * i, l1 = m / heads, h1 = m - l1 * heads, col = h1 * hd + d;
                const float* x = hs + static_cast<int64_t>(l1) * E;
                float a0 = 0.f, a1 = 0.f;
                for (int k = 0; k < E; k += 2) {
                    a0 = fmaf(x[k], __ldg(w.in_wt + static_cast<int64_t>(k) * 3 * E + col), a0);
                    a1 = fmaf(x[k + 1], __ldg(w.in_wt + static_cast<int64_t>(k + 1) * 3 * E + col), a1);
                }
                q[n] = ((a0 + a1) + w.in_b[col]) * scale;
            }
            __syncthreads();
            for (int task = warp; task < heads * L; task += nwarps) {
                const int i = task / L, j = task - i * L, m = L * i + j, l2 = m / heads, h2 = m - l2 * heads;
                const float* kr = kmat + l2 * E + h2 * hd;
                float a = 0.f;
                for (int d = lane; d < hd; d += 32) a = fmaf(q[i * hd + d], kr[d], a);
                a = warp_sum(a);
                if (j > 0 && pad_mask && pad_mask[static_cast<int64_t>(bm) * D + j - 1]) a = -CUDART_INF_F;
                if (lane == 0) p[i * L + j] = a;
            }
            __syncthreads();
        } else {
        // 4'. s[h][j] = q_h . R_j k_{h,j}; key-padding mask -> -inf
        for (int task = warp; task < heads * L; task += nwarps) {
            const int h = task / L, j = task % L;
            const float* kr = kmat + j * E + h * hd;
            float a = 0.f;
            for (int d = lane; d < hd; d += 32) {
                const float kd = kr[d], kp = kr[d ^ 1];
                float sn, cs;
                sincosf(static_cast<float>(j) * w.rope_freqs[d >> 1], &sn, &cs);
                const float rot = (d & 1) ? kp : -kp;                // rotate_half: (x1, x2) -> (-x2, x1)
                a = fmaf(q[h * hd + d], kd * cs + rot * sn, a);
            }
            a = warp_sum(a);
            if (j > 0 && pad_mask && pad_mask[static_cast<int64_t>(bm) * D + j - 1]) a = -CUDART_INF_F;
            if (lane == 0) p[h * L + j] = a;
        }
        __syncthreads();
        }
    } else {
    // 3. qk[h][k] = sum_d q[h,d] Wk[h*hd+d][k];  cterm[h] = q_h . bk_h
    //    (reads the UN-transposed in_proj_weight [3E][E]: consecutive threads -> consecutive k -> coalesced)
    for (int idx = threadIdx.x; idx < heads * E; idx += blockDim.x) {
        const int h = idx / E, k = idx % E;
        const float* wr = w.in_w + (static_cast<int64_t>(E) + h * hd) * E + k;
        float a = 0.f;
#pragma unroll 8
        for (int d = 0; d < hd; ++d) a = fmaf(q[h * hd + d], __ldg(wr + static_cast<int64_t>(d) * E), a);   // hd % 8 == 0
        qk[idx] = a;
    }
    if (threadIdx.x < heads) {
        float a = 0.f;
        for (int d = 0; d < hd; ++d) a = fmaf(q[threadIdx.x * hd + d], w.in_b[E + threadIdx.x * hd + d], a);
        cterm[threadIdx.x] = a;
    }
    __syncthreads();
    // 4. scores, key-padding mask -> -inf (transformer_blocks.py:244-252; CLS column never masked, dino.py:149-150)
    for (int task = warp; task < heads * L; task += nwarps) {
        const int h = task / L, j = task % L;
        const float* hr = hs + static_cast<int64_t>(j) * E;
        float a = 0.f;
        for (int k = lane; k < E; k += 32) a = fmaf(qk[h * E + k], hr[k], a);
        a = warp_sum(a) + cterm[h];
        if (j > 0 && pad_mask && pad_mask[static_cast<int64_t>(bm) * D + j - 1]) a = -CUDART_INF_F;
        if (lane == 0) p[h * L + j] = a;
    }
    __syncthreads();
    }
    // 5. softmax per head
    for (int h = warp; h < heads; h += nwarps) {
        float m = -CUDART_INF_F;
        for (int j = lane; j < L; j += 32) m = fmaxf(m, p[h * L + j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < L; j += 32) { const float e = expf(p[h * L + j] - m); p[h * L + j] = e; s += e; }
        const float inv = 1.0f / warp_sum(s);
        for (int j = lane; j < L; j += 32) {
            const float pj = p[h * L + j] * inv;
            p[h * L + j] = pj;
            if (slice_cls) slice_cls[(static_cast<int64_t>(b) * heads + h) * L + j] = pj;
        }
    }
    __syncthreads();
    // 6. hbar[h][k] = sum_j p[h][j] hs[j][k]
    for (int k = threadIdx.x; k < E; k += blockDim.x) {
        float acc[16];
#pragma unroll
        for (int h = 0; h < 16; ++h) acc[h] = 0.f;
        for (int j = 0; j < L; ++j) {
            const float hv = hs[static_cast<int64_t>(j) * E + k];
#pragma unroll
            for (int h = 0; h < 16; ++h)
                if (h < heads) acc[h] = fmaf(p[h * L + j], hv, acc[h]);
        }
#pragma unroll
        for (int h = 0; h < 16; ++h)
            if (h < heads) hbar[h * E + k] = acc[h];
    }
    __syncthreads();
    // 7. o[n] = Wv[n,:] . hbar[head(n)] + bv[n]
    block_matvec(hbar, hd, E, w.in_wt + 2 * E, 3 * E, w.in_b + 2 * E, nullptr, t0, scr, E, E, false, 1.0f);
    // 8. x1 = x0 + out_proj(o)                                           (transformer_blocks.py:566)
    block_matvec(t0, E, 0, w.out_wt, E, w.out_b, x0, t1, scr, E, E, false, 1.0f);
    // 9. x2 = x1 + W2 relu(W1 LN2(x1) + b1) + b2                         (transformer_blocks.py:567,585)
    block_layernorm(t1, t0, w.n2w, w.n2b, E, 1e-5f, red);
    block_matvec(t0, E, 0, w.l1_wt, E, w.l1_b, nullptr, t2, scr, E, E, true, 1.0f);
    block_matvec(t2, E, 0, w.l2_wt, E, w.l2_b, t1, t0, scr, E, E, false, 1.0f);
    // 10. final LayerNorm (dino.py:95), feature = row 0 (dino.py:153), logits (dino.py:166; nn.Identity if !enable_linear)
    block_layernorm(t0, t1, w.nfw, w.nfb, E, 1e-5f, red);
    if (feat)
        for (int i = threadIdx.x; i < E; i += blockDim.x) feat[static_cast<int64_t>(b) * E + i] = t1[i];
    if (logits)
        for (int c = warp; c < out_ch; c += nwarps) {
            float a = 0.f;
            for (int k = lane; k < E; k += 32) a = fmaf(t1[k], w.head_wt[static_cast<int64_t>(k) * out_ch + c], a);
            a = warp_sum(a);
            if (lane == 0) logits[static_cast<int64_t>(b) * out_ch + c] = a + w.head_b[c];
        }
}

int launch_slice_fusion(const float* enc_cls, const uint8_t* pad_mask, const SliceWeights& w, float* hs_scratch,
                        float* logits, float* feat, float* slice_cls, int B, int D, int Eenc, int E, int heads, int out_ch,
                        int mode, int mask_period, cudaStream_t stream) {
    MST_REQUIRE(heads <= 16 && E % heads == 0 && Eenc % 2 == 0, "slice fusion: heads=%d E=%d unsupported", heads, E);
    const int L = D + 1;
    MST_REQUIRE(E % (4 * MV_PARTS) == 0, "slice fusion: slice embedding %d must be a multiple of %d", E, 4 * MV_PARTS);
    const size_t smem = (static_cast<size_t>(6 + MV_PARTS) * E + 2 * heads * E + 32 + 40 + ((heads * L + 3) & ~3) +
                         ((w.rope_freqs || w.liere) ? static_cast<size_t>(L) * E : 0)) * sizeof(float);
    MST_REQUIRE(!w.liere || (B == 1 && L == 33), "rotary_positional_encoding='LiRE' exists for batch 1 and 32 slices only (the reference raises otherwise)");
    MST_REQUIRE(smem <= 227 * 1024, "slice transformer: %zu bytes of shared memory needed (D=%d too large)", smem, D);
    MST_SET_DYN_SMEM(slice_fusion_kernel, 227 * 1024);
    slice_fusion_kernel<<<B, 768, smem, stream>>>(enc_cls, pad_mask, w, hs_scratch, logits, feat, slice_cls, D, Eenc, E, heads,
                                                  out_ch, mode, mask_period);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Saliency combiner (reference dino.py:173-202 + scripts/main_predict.py:73-74,100) and the x14 upsampler
// (main_predict.py:161-162: trilinear with depth scale 1 == per-slice bilinear, align_corners=False).
// ---------------------------------------------------------------------------------------------------
// One WARP per output slice (8 slices per CTA, no block-wide synchronisation: the kernel is a few hundred loads per slice and was
// latency-bound as one CTA per slice).  nvar = 8 (test-time augmentation, main_predict.py:147-158): plane_cls / slice_cls hold the
// 8 flipped variants of every volume, variant-major ([8*B*D, ...] / [8*B, ...]); the warp of output slice (b, d) walks the variants
// in the script's order, un-flips each variant's coarse map (flip of dims 2/3/4 = depth / grid rows / grid columns) through its
// shared-memory row and averages, so the x14 upsample runs ONCE on the averaged coarse map as the script does (:161-162).
constexpr int SAL_WARPS = 8;
__global__ void __launch_bounds__(SAL_WARPS * 32) saliency_combine_kernel(const float* __restrict__ plane_cls, const float* __restrict__ slice_cls,
                                                                            int B, int D, int heads, int sheads, int gh, int gw, int skip,
                                                                            int nvar, float* __restrict__ attn_maps, float* __restrict__ plane_attn,
                                                                            float* __restrict__ slice_attn, float* __restrict__ coarse) {
    extern __shared__ float sm[];  // per warp: acc[P] | tot[P]
    const int P = gh * gw, L = D + 1, N = P + skip;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * SAL_WARPS + warp;
    if (s >= B * D) return;
    float* acc = sm + static_cast<int64_t>(warp) * 2 * P;
    float* tot = acc + P;
    const int b = s / D, d = s - b * D;
    float wtot = 0.f;
    for (int v = 0; v < nvar; ++v) {
        const bool flip_d = (0xB2 >> v) & 1, flip_h = (0xD4 >> v) & 1, flip_w = (0xE8 >> v) & 1;
        const int bv = v * B + b, dv = flip_d ? D - 1 - d : d;
        const int64_t sv = static_cast<int64_t>(bv) * D + dv;
        // slice weight: mean over the slice heads of S[b,h,1+d] / sum_j S[b,h,1+j]                    (dino.py:174-181)
        float wslice = 0.f;
        for (int h = 0; h < sheads; ++h) {
            const float* sr = slice_cls + (static_cast<int64_t>(bv) * sheads + h) * L + 1;
            float t = 0.f;
            for (int j = lane; j < D; j += 32) t += sr[j];
            t = warp_sum(t);
            wslice += sr[dv] / t;
        }
        wslice /= sheads;
        __syncwarp();
        for (int i = lane; i < P; i += 32) acc[i] = 0.f;
        for (int h = 0; plane_cls && h < heads; ++h) {
            const float* pr = plane_cls + (sv * heads + h) * N + skip;   // drop CLS (+ registers)          (dino.py:191-192)
            float t = 0.f;
            for (int i = 1 + lane; i < P; i += 32) t += pr[i];          // patch 0 := 0                      (dino.py:193)
            t = warp_sum(t);
            for (int i = lane; i < P; i += 32) {
                const float a = (i == 0) ? 0.f : pr[i] / t;             // dino.py:194
                const float m = wslice * a;                             // dino.py:201
                if (attn_maps) attn_maps[(sv * heads + h) * P + i] = m;
                if (plane_attn) plane_attn[(sv * heads + h) * P + i] = a;
                acc[i] += m;
            }
        }
        __syncwarp();
        // head mean (main_predict.py:73-74), un-flipped into the output orientation, summed in the script's order (:157)
        for (int i = lane; i < P; i += 32) {
            const int y = i / gw, x = i - y * gw;
            const float m = acc[(flip_h ? gh - 1 - y : y) * gw + (flip_w ? gw - 1 - x : x)] / heads;
            tot[i] = v == 0 ? m : tot[i] + m;
        }
        wtot = v == 0 ? wslice : wtot + wslice;
    }
    const float inv = 1.0f / nvar;   // exact (1 or 1/8)
    if (coarse)
        for (int i = lane; i < P; i += 32) coarse[static_cast<int64_t>(s) * P + i] = tot[i] * inv;
    if (slice_attn && lane == 0) slice_attn[s] = wtot * inv;
}

// x(H/gh) bilinear upsample, align_corners=False (== F.interpolate 'trilinear' with depth scale 1, main_predict.py:161-162):
//   out[y, x] = hy * (hx * c[y0, x0] + lx * c[y0, x1]) + ly * (hx * c[y1, x0] + lx * c[y1, x1])
// One CTA = one slice (or one of `chunks` contiguous pieces of it when there are few slices).  The horizontal blends
// hrow[r][x] = hx * c[r, x0] + lx * c[r, x1] of all coarse rows are built once in shared memory; every output is then two
// shared-memory values, one FMUL and one FMA, and the slice is written as ONE flat run of 16-byte streaming stores (thread i ->
// float4 i, i + 256, ...: every warp instruction covers 512 contiguous bytes).  This is the bandwidth-bound step of the saliency
// path (6.4 MB per volume): 6.8 TB/s on a 256-volume batch, against 7.0-7.3 TB/s for a plain fill / cudaMemset
// (profiles/micro/upsample_bw.cu; a row-blocked layout with 56 of 64 lanes active reached 4.4 TB/s).
__global__ void __launch_bounds__(256) saliency_upsample_kernel(const float* __restrict__ coarse, float* __restrict__ full,
                                                                 int gh, int gw, int H, int W, int chunks) {
    extern __shared__ float hrow[];  // [gh][W]
    const int s = blockIdx.x;
    const float sy = static_cast<float>(gh) / static_cast<float>(H), sx = static_cast<float>(gw) / static_cast<float>(W);
    const float* c = coarse + static_cast<int64_t>(s) * gh * gw;
    for (int idx = threadIdx.x; idx < gh * W; idx += blockDim.x) {
        const int r = idx / W, x = idx - r * W;
        const float fx = fmaxf(sx * (x + 0.5f) - 0.5f, 0.f);
        const int x0 = static_cast<int>(fx), x1 = min(x0 + 1, gw - 1);
        const float lx = fx - x0, hx = 1.f - lx;
        hrow[idx] = hx * __ldg(c + r * gw + x0) + lx * __ldg(c + r * gw + x1);
    }
    __syncthreads();
    float* out = full + static_cast<int64_t>(s) * H * W;
    if ((W & 3) == 0) {
        const int W4 = W >> 2, total = H * W4;
        const int per = (total + chunks - 1) / chunks;
        const int i_end = min(total, (static_cast<int>(blockIdx.y) + 1) * per);
        float4* out4 = reinterpret_cast<float4*>(out);
        for (int i = blockIdx.y * per + threadIdx.x; i < i_end; i += blockDim.x) {
            const int y = i / W4, cg = i - y * W4;
            const float fy = fmaxf(sy * (y + 0.5f) - 0.5f, 0.f);
            const int y0 = static_cast<int>(fy), y1 = min(y0 + 1, gh - 1);
            const float ly = fy - y0, hy = 1.f - ly;
            const float4 a = reinterpret_cast<const float4*>(hrow + y0 * W)[cg], bq = reinterpret_cast<const float4*>(hrow + y1 * W)[cg];
            __stcs(out4 + i, make_float4(hy * a.x + ly * bq.x, hy * a.y + ly * bq.y, hy * a.z + ly * bq.z, hy * a.w + ly * bq.w));
        }
    } else {
        const int total = H * W, per = (total + chunks - 1) / chunks;
        const int i_end = min(total, (static_cast<int>(blockIdx.y) + 1) * per);
        for (int i = blockIdx.y * per + threadIdx.x; i < i_end; i += blockDim.x) {
            const int y = i / W, x = i - y * W;
            const float fy = fmaxf(sy * (y + 0.5f) - 0.5f, 0.f);
            const int y0 = static_cast<int>(fy), y1 = min(y0 + 1, gh - 1);
            const float ly = fy - y0, hy = 1.f - ly;
            out[i] = hy * hrow[y0 * W + x] + ly * hrow[y1 * W + x];
        }
    }
}

int launch_saliency_combine(const float* plane_cls, const float* slice_cls, int B, int D, int heads, int slice_heads, int skip, int gh,
                            int gw, int tta, float* attn_maps, float* plane_attn, float* slice_attn, float* coarse, cudaStream_t stream) {
    const int P = gh * gw, BD = B * D;
    MST_REQUIRE(!tta || (attn_maps == nullptr && plane_attn == nullptr), "saliency: the per-head maps are per variant; with tta only coarse / slice_attn");
    MST_REQUIRE(heads <= 32 && slice_heads <= 32, "saliency: at most 32 heads");
    const size_t smem = static_cast<size_t>(SAL_WARPS) * 2 * P * sizeof(float);
    MST_REQUIRE(smem <= 200 * 1024, "saliency: a %d x %d patch grid needs %zu bytes of shared memory", gh, gw, smem);
    MST_SET_DYN_SMEM(saliency_combine_kernel, 200 * 1024);
    saliency_combine_kernel<<<(BD + SAL_WARPS - 1) / SAL_WARPS, SAL_WARPS * 32, smem, stream>>>(plane_cls, slice_cls, B, D, heads, slice_heads, gh,
                                                                                                gw, skip, tta ? 8 : 1, attn_maps, plane_attn,
                                                                                                slice_attn, coarse);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int launch_saliency_upsample(const float* coarse, float* full, int B, int D, int gh, int gw, int H, int W, cudaStream_t stream) {
    const int BD = B * D;
    const size_t smem = static_cast<size_t>(gh) * W * sizeof(float);
    MST_REQUIRE(smem <= 200 * 1024, "saliency upsample: a %d x %d row table needs %zu bytes of shared memory", gh, W, smem);
    MST_SET_DYN_SMEM(saliency_upsample_kernel, 200 * 1024);
    int chunks = 1;   // few slices (one volume): split every slice so that the grid still covers the GPU twice
    if (BD < 296) chunks = (296 + BD - 1) / BD;
    if (chunks > 8) chunks = 8;
    saliency_upsample_kernel<<<dim3(BD, chunks), 256, smem, stream>>>(coarse, full, gh, gw, H, W, chunks);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mst
