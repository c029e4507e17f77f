// Kernels either side of the main path: position-embedding interpolation for inputs whose
// patch grid differs from the checkpoint's, full attention probabilities + attention rollout
// (get_attention_cls), and the exact quantiles the predict script takes of the saliency volume.
#include <math_constants.h>
#include "common.cuh"
#include "ptx.cuh"

namespace mst {

// ---------------------------------------------------------------------------------------------------
// Bicubic resampling of the patch position embedding (reference vision_transformer.py:179-211:
// F.interpolate(mode="bicubic", antialias=False, scale_factor=((gh+0.1)/M, (gw+0.1)/M)), align_corners=False).
// ATen semantics: source coordinate = scale*(dst+0.5)-0.5 with scale = 1/scale_factor (not in/out), NOT clamped for
// cubic; taps at floor-1..floor+2 clamped to the border; Keys kernel with A = -0.75.
//   pos [1+M*M, E] (row 0 = class position)  ->  posb[p, n] = interp(p, n) + conv_bias[n],  cls_pos0[n] = cls[n] + pos[0, n]
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
    const float A = -0.75f;
    const float x0 = t + 1.0f, x1 = t, x2 = 1.0f - t, x3 = 2.0f - t;
    c[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
    c[1] = ((A + 2.0f) * x1 - (A + 3.0f)) * x1 * x1 + 1.0f;
    c[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
    c[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

__global__ void __launch_bounds__(128) pos_bicubic_kernel(const float* __restrict__ pos, const float* __restrict__ cbias,
                                                           float* __restrict__ posb, int M, int gh, int gw, int E,
                                                           float scale_y, float scale_x) {
    const int p = blockIdx.x, oy = p / gw, ox = p % gw;
    const float ry = scale_y * (oy + 0.5f) - 0.5f, rx = scale_x * (ox + 0.5f) - 0.5f;
    const float fy = floorf(ry), fx = floorf(rx);
    const int iy = static_cast<int>(fy), ix = static_cast<int>(fx);
    float cy[4], cx[4];
    cubic_coeffs(ry - fy, cy);
    cubic_coeffs(rx - fx, cx);
    const float* grid = pos + E;  // skip the class row
    for (int n = threadIdx.x; n < E; n += blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int y = min(max(iy - 1 + i, 0), M - 1);
            float row = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = min(max(ix - 1 + j, 0), M - 1);
                row = fmaf(cx[j], grid[(static_cast<int64_t>(y) * M + x) * E + n], row);
            }
            acc = fmaf(cy[i], row, acc);
        }
        posb[static_cast<int64_t>(p) * E + n] = acc + (cbias ? cbias[n] : 0.f);
    }
}

int launch_pos_bicubic(const float* pos, const float* cbias, float* posb, int M, int gh, int gw, int E, float scale_y,
                       float scale_x, cudaStream_t stream) {
    pos_bicubic_kernel<<<gh * gw, 128, 0, stream>>>(pos, cbias, posb, M, gh, gw, E, scale_y, scale_x);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// The same table for the torch.hub "_reg" encoders (dinov2_vit*14_reg, dino.py:60-61), which are built with
// interpolate_antialias=True, interpolate_offset=0.0 (vision_transformer.py:66-67,198-210):
// F.interpolate(size=(gh, gw), mode="bicubic", antialias=True).  ATen's separable anti-aliased resampling
// (UpSampleKernel.cpp, _compute_indices_min_size_weights_aa): per output index i, scale = in/out,
// support = 2*max(scale,1), center = scale*(i+0.5), taps xmin = max(int(center-support+0.5), 0) ..
// min(int(center+support+0.5), in), weight = keys_{a=-0.5}((j+xmin-center+0.5)/max(scale,1)), normalised to sum 1;
// width pass first, then height, fp32 throughout.
// ---------------------------------------------------------------------------------------------------
constexpr int AA_MAX_TAPS = 64;
__device__ __forceinline__ float aa_keys(float x) {
    const float a = -0.5f;
    x = fabsf(x);
    if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
    if (x < 2.0f) return (((x - 5.0f) * x + 8.0f) * x - 4.0f) * a;
    return 0.0f;
}
__device__ __forceinline__ void aa_taps(int i, int in_size, int out_size, int& xmin, int& xsize, float* w) {
    const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
    const float support = scale >= 1.0f ? 2.0f * scale : 2.0f;
    const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
    const float center = scale * (i + 0.5f);
    xmin = max(static_cast<int>(center - support + 0.5f), 0);
    xsize = min(static_cast<int>(center + support + 0.5f), in_size) - xmin;
    xsize = min(max(xsize, 0), AA_MAX_TAPS);
    float total = 0.f;
    for (int j = 0; j < xsize; ++j) { w[j] = aa_keys((j + xmin - center + 0.5f) * invscale); total += w[j]; }
    if (total != 0.f)
        for (int j = 0; j < xsize; ++j) w[j] /= total;
}
__global__ void __launch_bounds__(128) pos_bicubic_aa_kernel(const float* __restrict__ pos, const float* __restrict__ cbias,
                                                              float* __restrict__ posb, int M, int gh, int gw, int E) {
    __shared__ float wy[AA_MAX_TAPS], wx[AA_MAX_TAPS];
    __shared__ int lim[4];
    const int p = blockIdx.x, oy = p / gw, ox = p % gw;
    if (threadIdx.x == 0) aa_taps(oy, M, gh, lim[0], lim[1], wy);
    if (threadIdx.x == 32) aa_taps(ox, M, gw, lim[2], lim[3], wx);
    __syncthreads();
    const int y0 = lim[0], ny = lim[1], x0 = lim[2], nx = lim[3];
    const float* grid = pos + E;  // skip the class row
    for (int n = threadIdx.x; n < E; n += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < ny; ++i) {
            float row = 0.f;   // the width pass of ATen's separable scheme for source row y0 + i
            for (int j = 0; j < nx; ++j) row += wx[j] * grid[(static_cast<int64_t>(y0 + i) * M + x0 + j) * E + n];
            acc += wy[i] * row;
        }
        posb[static_cast<int64_t>(p) * E + n] = acc + (cbias ? cbias[n] : 0.f);
    }
}
int launch_pos_bicubic_aa(const float* pos, const float* cbias, float* posb, int M, int gh, int gw, int E, cudaStream_t stream) {
    const float smax = fmaxf(static_cast<float>(M) / gh, static_cast<float>(M) / gw);
    MST_REQUIRE(2.0f * fmaxf(smax, 1.0f) * 2.0f + 2.0f <= AA_MAX_TAPS, "anti-aliased position resampling %d -> %dx%d needs more than %d taps",
                M, gh, gw, AA_MAX_TAPS);
    pos_bicubic_aa_kernel<<<gh * gw, 128, 0, stream>>>(pos, cbias, posb, M, gh, gw, E);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Full attention probabilities of one encoder block: probs[s, h, i, j] = softmax_j(q_i . k_j), q pre-scaled.
// This is what the reference's hook stores for every block when save_attn=True (dino.py:229-241); only
// get_attention_cls (dino.py:204-212) reads more than row 0 of the last one, so it is produced on request.
// One CTA per (slice, head), K in shared memory, one warp per query row, fp32.
// ---------------------------------------------------------------------------------------------------
constexpr int PROBS_WARPS = 8;
constexpr int PROBS_MAXJ = 12;
template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<bf16>(const bf16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(PROBS_WARPS * 32) attention_probs_kernel(const T* __restrict__ qkv, float* __restrict__ probs,
                                                                            int N, int heads) {
    extern __shared__ float sm[];
    float* Ks = sm;                  // [N][65]
    float* qs = Ks + N * 65;         // [warps][64]
    const int s = blockIdx.x / heads, h = blockIdx.x % heads;
    const int E = heads * 64;
    const int64_t ld = 3 * E;
    const T* base = qkv + static_cast<int64_t>(s) * N * ld;
    for (int i = threadIdx.x; i < N * 64; i += blockDim.x) {
        const int j = i >> 6, d = i & 63;
        Ks[j * 65 + d] = ld_f<T>(base + j * ld + E + h * 64 + d);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* q = qs + warp * 64;
    float* out = probs + (static_cast<int64_t>(s) * heads + h) * N * N;
    for (int r = warp; r < N; r += PROBS_WARPS) {
        q[lane] = ld_f<T>(base + r * ld + h * 64 + lane);
        q[lane + 32] = ld_f<T>(base + r * ld + h * 64 + lane + 32);
        __syncwarp();
        float sc[PROBS_MAXJ];
        float lmax = -CUDART_INF_F;
#pragma unroll
        for (int jj = 0; jj < PROBS_MAXJ; ++jj) {
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
        for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        float lsum = 0.f;
#pragma unroll
        for (int jj = 0; jj < PROBS_MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            const float e = j < N ? expf(sc[jj] - lmax) : 0.f;
            sc[jj] = e;
            lsum += e;
        }
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        const float inv = 1.0f / lsum;
#pragma unroll
        for (int jj = 0; jj < PROBS_MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            if (j < N) out[static_cast<int64_t>(r) * N + j] = sc[jj] * inv;
        }
        __syncwarp();
    }
}

template <typename T>
int launch_attention_probs(const T* qkv, float* probs, int BD, int N, int heads, cudaStream_t stream) {
    MST_REQUIRE(N <= PROBS_MAXJ * 32, "full attention maps support at most %d tokens per slice (got %d)", PROBS_MAXJ * 32, N);
    const size_t smem = (static_cast<size_t>(N) * 65 + PROBS_WARPS * 64) * sizeof(float);
    MST_SET_DYN_SMEM(attention_probs_kernel<T>, 227 * 1024);
    attention_probs_kernel<T><<<BD * heads, PROBS_WARPS * 32, smem, stream>>>(qkv, probs, N, heads);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}
template int launch_attention_probs<float>(const float*, float*, int, int, int, cudaStream_t);
template int launch_attention_probs<bf16>(const bf16*, float*, int, int, int, cudaStream_t);

// ---------------------------------------------------------------------------------------------------
// Attention rollout (dino.py:204-212):  R = maps[-1];  for attn in reversed(maps[:-1]): R = attn @ R.
// Batched fp32 N x N products (N = 257 is not a tile multiple: guarded loads), 64x64 tile, 4x4 micro-tile.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bmm_nn_f32_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                          float* __restrict__ C, int N) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int64_t mat = static_cast<int64_t>(blockIdx.z) * N * N;
    const float* a = A + mat;
    const float* b = Bm + mat;
    float* c = C + mat;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < N; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i >> 4, k = i & 15;       // A tile: rows m0+r, cols k0+k
            As[k][r] = (m0 + r < N && k0 + k < N) ? a[static_cast<int64_t>(m0 + r) * N + k0 + k] : 0.f;
            const int kk = i >> 6, n = i & 63;      // B tile: rows k0+kk, cols n0+n
            Bs[kk][n] = (k0 + kk < N && n0 + n < N) ? b[static_cast<int64_t>(k0 + kk) * N + n0 + n] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a4[4], b4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a4[i] = As[k][ty * 4 + i]; b4[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (r < N && n < N) c[static_cast<int64_t>(r) * N + n] = acc[i][j];
        }
}

int launch_rollout(const float* maps, int depth, int nmat, int N, float* out, float* scratch, cudaStream_t stream) {
    const int64_t per_layer = static_cast<int64_t>(nmat) * N * N;
    if (depth == 1) {
        MST_CHECK_CUDA(cudaMemcpyAsync(out, maps, per_layer * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        return 0;
    }
    MST_REQUIRE(nmat <= 65535, "rollout: %d matrices per launch exceed the grid limit", nmat);
    // depth-1 products ping-pong between `scratch` and `out`, arranged so that the last one lands in `out`
    const float* cur = maps + static_cast<int64_t>(depth - 1) * per_layer;
    const dim3 grid((N + 63) / 64, (N + 63) / 64, nmat);
    for (int l = depth - 2, step = 0; l >= 0; --l, ++step) {
        float* dst = ((depth - 2 - step) % 2 == 0) ? out : scratch;
        bmm_nn_f32_kernel<<<grid, 256, 0, stream>>>(maps + static_cast<int64_t>(l) * per_layer, cur, dst, N);
        MST_CHECK_CUDA(cudaGetLastError());
        cur = dst;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Exact quantiles of each item's n values (scripts/main_predict.py:243-245,296: np.quantile(weight, 0.999),
// np.quantile(weight, [0.995, 0.999]) on the upsampled saliency volume; numpy's default 'linear' method).
// Order statistics by 4-pass MSB radix select on order-preserving uint32 keys; one selection per requested
// rank (two adjacent ranks per quantile), then numpy's _lerp in its own arithmetic (fp32 difference, fp64 blend).
// ---------------------------------------------------------------------------------------------------
constexpr int QSEL_MAX = 16;  // selections per item (2 per quantile)
__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
struct QSelState {  // per (item, selection)
    uint32_t prefix;       // key bits decided so far (high bits)
    unsigned long long rank;  // remaining rank inside the prefix bucket
};

__global__ void __launch_bounds__(256) qsel_hist_kernel(const float* __restrict__ data, int64_t n, int nsel, int pass,
                                                         const QSelState* __restrict__ st, unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[QSEL_MAX * 256];
    __shared__ uint32_t spre[QSEL_MAX];
    const int item = blockIdx.y;
    for (int i = threadIdx.x; i < nsel * 256; i += blockDim.x) sh[i] = 0;
    if (threadIdx.x < nsel) spre[threadIdx.x] = st[item * nsel + threadIdx.x].prefix;
    __syncthreads();
    const int shift = 24 - 8 * pass;
    const uint32_t himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    const float* d = data + static_cast<int64_t>(item) * n;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint32_t k = f2key(d[i]);
        const uint32_t bin = (k >> shift) & 0xffu;
        for (int s = 0; s < nsel; ++s)
            if ((k & himask) == spre[s]) atomicAdd(&sh[s * 256 + bin], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nsel * 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[(static_cast<int64_t>(item) * nsel) * 256 + i], sh[i]);
}

__global__ void qsel_pick_kernel(int nsel, int pass, QSelState* __restrict__ st, unsigned int* __restrict__ hist) {
    const int item = blockIdx.x, s = threadIdx.x;
    if (s >= nsel) return;
    QSelState& q = st[item * nsel + s];
    unsigned int* h = hist + (static_cast<int64_t>(item) * nsel + s) * 256;
    unsigned long long r = q.rank;
    int bin = 0;
    for (; bin < 255; ++bin) {
        if (r < h[bin]) break;
        r -= h[bin];
    }
    q.rank = r;
    q.prefix |= static_cast<uint32_t>(bin) << (24 - 8 * pass);
    for (int i = 0; i < 256; ++i) h[i] = 0;
}

__global__ void qsel_init_kernel(int64_t n, int nq, const double* __restrict__ q, QSelState* __restrict__ st) {
    const int item = blockIdx.x, s = threadIdx.x;
    if (s >= 2 * nq) return;
    const double vi = q[s >> 1] * static_cast<double>(n - 1);  // numpy 'linear': virtual index q*(n-1)
    long long lo = static_cast<long long>(floor(vi));
    lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
    const long long hi = lo + 1 > n - 1 ? n - 1 : lo + 1;
    st[item * 2 * nq + s].prefix = 0;
    st[item * 2 * nq + s].rank = static_cast<unsigned long long>((s & 1) ? hi : lo);
}

__global__ void qsel_final_kernel(int64_t n, int nq, const double* __restrict__ q, const QSelState* __restrict__ st,
                                  double* __restrict__ out) {
    const int item = blockIdx.x, j = threadIdx.x;
    if (j >= nq) return;
    const float a = key2f(st[item * 2 * nq + 2 * j].prefix), b = key2f(st[item * 2 * nq + 2 * j + 1].prefix);
    const double vi = q[j] * static_cast<double>(n - 1);
    double lo = floor(vi);
    lo = lo < 0 ? 0 : (lo > static_cast<double>(n - 1) ? static_cast<double>(n - 1) : lo);
    const double t = vi - lo;
    const float diff = b - a;  // numpy: subtract(b, a) in the array dtype
    double r = static_cast<double>(a) + static_cast<double>(diff) * t;
    if (t >= 0.5) r = static_cast<double>(b) - static_cast<double>(diff) * (1.0 - t);
    out[item * nq + j] = r;
}

size_t quantile_workspace_bytes(int items, int nq) {
    const size_t nsel = 2 * static_cast<size_t>(nq);
    return items * nsel * (sizeof(QSelState) + 256 * sizeof(unsigned int)) + 256;
}

int launch_quantile(const float* data, int64_t n, int items, const double* q_dev, int nq, double* out, void* workspace,
                    int num_sms, cudaStream_t stream) {
    MST_REQUIRE(nq >= 1 && 2 * nq <= QSEL_MAX, "quantile: at most %d quantiles per call", QSEL_MAX / 2);
    MST_REQUIRE(n >= 1 && items >= 1 && items <= 65535, "quantile: bad sizes n=%lld items=%d", (long long)n, items);
    const int nsel = 2 * nq;
    QSelState* st = static_cast<QSelState*>(workspace);
    unsigned int* hist = reinterpret_cast<unsigned int*>(st + static_cast<size_t>(items) * nsel);
    MST_CHECK_CUDA(cudaMemsetAsync(hist, 0, static_cast<size_t>(items) * nsel * 256 * sizeof(unsigned int), stream));
    qsel_init_kernel<<<items, 32, 0, stream>>>(n, nq, q_dev, st);
    MST_CHECK_CUDA(cudaGetLastError());
    int gx = static_cast<int>((n + 256 * 16 - 1) / (256 * 16));
    const int cap = (4 * num_sms + items - 1) / items;
    gx = gx < 1 ? 1 : (gx > cap ? cap : gx);
    for (int pass = 0; pass < 4; ++pass) {
        qsel_hist_kernel<<<dim3(gx, items), 256, 0, stream>>>(data, n, nsel, pass, st, hist);
        MST_CHECK_CUDA(cudaGetLastError());
        qsel_pick_kernel<<<items, 32, 0, stream>>>(nsel, pass, st, hist);
        MST_CHECK_CUDA(cudaGetLastError());
    }
    qsel_final_kernel<<<items, 32, 0, stream>>>(n, nq, q_dev, st, out);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mst
