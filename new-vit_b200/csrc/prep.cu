// Input pipeline in front of the forward: what DUKE_Dataset3D's default transform chain does to one
// volume between the HDF5 read and the model's `source` tensor (reference mst/data/datasets/dataset_3d_duke.py:36-47, with
// image_resize / resample None and the random augmentations off, i.e. the predict configuration):
//   tio.Flip(1)                                              flip the H axis of the torchio tensor [C=1, W, H, D]
//   CropOrPad(image_crop, padding_mode='minimum')            augmentations_3d.py:144-195 (ini = ceil(n/2), fin = n - ini;
//                                                            pad first, then crop; tio.Pad -> np.pad(mode='minimum'))
//   ZNormalization(percentiles=(0.5, 99.5), masking_method = (x > x.min()) & (x < x.max()))   augmentations_3d.py:41-86
//   ImageOrSubjectToTensor                                   swapaxes(1, -1): [C, W, H, D] -> [C, D, H, W]  (:23-29)
// All of it is HBM-bound element work: one gather pass (flip, crop, pad and the axis swap through a shared-memory tile, so
// that both the reads along D and the writes along W are coalesced), a masked 4-pass radix select for the two cutoffs, one
// moments pass and one normalise pass; the passes after the gather read the 6.4 MB output volume, which stays in L2.
#include <math_constants.h>
#include "common.cuh"

namespace mst {

namespace {

__device__ __forceinline__ uint32_t pf2key(float f) {  // order-preserving float -> uint32
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float pkey2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

constexpr int PSEL = 4;  // order statistics per item: below / above rank of the low and of the high percentile

struct PrepState {  // per item, in the workspace
    uint32_t kmin, kmax;            // keys of the volume's min / max (after crop / pad)
    uint32_t prefix[PSEL];          // radix-select state
    unsigned long long rank[PSEL];
    unsigned long long count;       // masked voxels
    float w[2];                     // interpolation weights of the two percentiles
    float lo, hi;                   // cutoffs
    float mean, stdv;
    double sum, sumsq;              // of (clamped - lo) over the mask
    unsigned int hist[PSEL * 256];
};

// np.pad(mode='minimum') with the default stat_length pads axis by axis with the minimum of each 1-D line over the ORIGINAL
// extent of that axis, lines running through the pads of the axes handled before.  In closed form: a padded voxel whose axes in
// the set S are out of range holds the minimum over all source voxels with the in-range coordinates fixed and the axes in S
// free.  Those minima are seven marginal tables: m_w [H0,D0], m_h [W0,D0], m_d [W0,H0], m_wh [D0], m_wd [H0], m_hd [W0], m_whd.
// dst = min over `axis` of src viewed as [n0, n1, n2] (row-major); dst keeps the other two axes in order.
__global__ void __launch_bounds__(256) axis_min_kernel(const float* __restrict__ src, int64_t src_item, int n0, int n1, int n2,
                                                        int axis, float* __restrict__ dst, int64_t dst_item) {
    const float* s = src + blockIdx.y * src_item;
    float* d = dst + blockIdx.y * dst_item;
    const int na = axis == 0 ? n0 : (axis == 1 ? n1 : n2);
    const int o1 = axis == 2 ? n1 : n2;                    // fastest kept axis
    const int o0 = axis == 0 ? n1 : n0;                    // slowest kept axis
    const int64_t total = static_cast<int64_t>(o0) * o1;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int a = static_cast<int>(i / o1), b = static_cast<int>(i % o1);
        int64_t base, step;
        if (axis == 0) { base = static_cast<int64_t>(a) * n2 + b; step = static_cast<int64_t>(n1) * n2; }
        else if (axis == 1) { base = static_cast<int64_t>(a) * n1 * n2 + b; step = n2; }
        else { base = (static_cast<int64_t>(a) * n1 + b) * n2; step = 1; }
        float m = CUDART_INF_F;
        for (int k = 0; k < na; ++k) m = fminf(m, s[base + k * step]);
        d[i] = m;
    }
}

struct PrepGeom {
    int W0, H0, D0;     // source extents (torchio order: W, H, D; D fastest in memory)
    int W, H, D;        // target extents
    int ow, oh, od;     // source coordinate = target coordinate + offset (crop_ini - pad_ini)
    int flip_h;
};
struct PrepMargins {    // per-item tables, nullptr when no axis is padded
    const float *m_w, *m_h, *m_d, *m_wh, *m_wd, *m_hd, *m_whd;
    int64_t item;       // floats per item in the margin block
};

__device__ __forceinline__ float prep_fetch(const float* __restrict__ s, const PrepGeom& g, const PrepMargins& pm, int64_t moff,
                                            int w, int h, int d) {
    const int sw = w + g.ow, sd = d + g.od;
    int sh = h + g.oh;
    const bool iw = sw >= 0 && sw < g.W0, ih = sh >= 0 && sh < g.H0, id = sd >= 0 && sd < g.D0;
    if (g.flip_h) sh = g.H0 - 1 - sh;   // tio.Flip(1) runs before the pad; minima along H do not see it
    if (iw && ih && id) return s[(static_cast<int64_t>(sw) * g.H0 + sh) * g.D0 + sd];
    if (!iw && ih && id) return pm.m_w[moff + static_cast<int64_t>(sh) * g.D0 + sd];
    if (iw && !ih && id) return pm.m_h[moff + static_cast<int64_t>(sw) * g.D0 + sd];
    if (iw && ih && !id) return pm.m_d[moff + static_cast<int64_t>(sw) * g.H0 + sh];
    if (!iw && !ih && id) return pm.m_wh[moff + sd];
    if (!iw && ih && !id) return pm.m_wd[moff + sh];
    if (iw && !ih && !id) return pm.m_hd[moff + sw];
    return pm.m_whd[moff];
}

// out[item, d, h, w] = padded / cropped / flipped src[item, w, h, d]; per-item min / max keys by atomics.
// grid (ceil(W/32), H, items), block (32, 8): 32(w) x 32(d) tiles through shared memory.
__global__ void __launch_bounds__(256) prep_gather_kernel(const float* __restrict__ src, PrepGeom g, PrepMargins pm,
                                                           float* __restrict__ out, PrepState* __restrict__ st) {
    __shared__ float tile[32][33];
    __shared__ uint32_t smin[8], smax[8];
    const int item = blockIdx.z, h = blockIdx.y, w0 = blockIdx.x * 32;
    const float* s = src + static_cast<int64_t>(item) * g.W0 * g.H0 * g.D0;
    float* o = out + static_cast<int64_t>(item) * g.W * g.H * g.D;
    const int64_t moff = static_cast<int64_t>(item) * pm.item;
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    for (int d0 = 0; d0 < g.D; d0 += 32) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int wl = threadIdx.y + 8 * k, w = w0 + wl, d = d0 + threadIdx.x;
            if (w < g.W && d < g.D) {
                const float v = prep_fetch(s, g, pm, moff, w, h, d);
                tile[wl][threadIdx.x] = v;
                const uint32_t key = pf2key(v);
                kmin = min(kmin, key);
                kmax = max(kmax, key);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int dl = threadIdx.y + 8 * k, d = d0 + dl, w = w0 + threadIdx.x;
            if (w < g.W && d < g.D) o[(static_cast<int64_t>(d) * g.H + h) * g.W + w] = tile[threadIdx.x][dl];
        }
        __syncthreads();
    }
    for (int off = 16; off > 0; off >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, off));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, off));
    }
    if (threadIdx.x == 0) { smin[threadIdx.y] = kmin; smax[threadIdx.y] = kmax; }
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        for (int i = 1; i < 8; ++i) { kmin = min(kmin, smin[i]); kmax = max(kmax, smax[i]); }
        atomicMin(&st[item].kmin, kmin);
        atomicMax(&st[item].kmax, kmax);
    }
}

__global__ void prep_init_kernel(PrepState* __restrict__ st) {
    PrepState& q = st[blockIdx.x];
    for (int i = threadIdx.x; i < PSEL * 256; i += blockDim.x) q.hist[i] = 0;
    if (threadIdx.x == 0) {
        q.kmin = 0xffffffffu; q.kmax = 0u; q.count = 0; q.sum = 0.0; q.sumsq = 0.0;
        for (int s = 0; s < PSEL; ++s) { q.prefix[s] = 0; q.rank[s] = 0; }
    }
}

// One radix pass over the masked voxels (min < x < max, augmentations_3d.py:75 masked_select with dataset_3d_duke.py:43's mask).
__global__ void __launch_bounds__(256) prep_hist_kernel(const float* __restrict__ vol, int64_t n, int pass,
                                                         PrepState* __restrict__ st) {
    __shared__ unsigned int sh[PSEL * 256];
    __shared__ uint32_t spre[PSEL];
    const int item = blockIdx.y;
    PrepState& q = st[item];
    for (int i = threadIdx.x; i < PSEL * 256; i += blockDim.x) sh[i] = 0;
    if (threadIdx.x < PSEL) spre[threadIdx.x] = q.prefix[threadIdx.x];
    const uint32_t kmin = q.kmin, kmax = q.kmax;
    __syncthreads();
    const int shift = 24 - 8 * pass;
    const uint32_t himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    const float4* v4 = reinterpret_cast<const float4*>(vol + static_cast<int64_t>(item) * n);  // n % 4 == 0 (checked by the host)
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n / 4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float4 v = v4[i];
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t k = pf2key(e[j]);
            if (k > kmin && k < kmax) {
                const uint32_t bin = (k >> shift) & 0xffu;
                if (pass == 0) atomicAdd(&sh[bin], 1u);  // all four selections share the first histogram
                else
#pragma unroll
                    for (int s = 0; s < PSEL; ++s)
                        if ((k & himask) == spre[s]) atomicAdd(&sh[s * 256 + bin], 1u);
            }
        }
    }
    __syncthreads();
    const int lim = pass == 0 ? 256 : PSEL * 256;
    for (int i = threadIdx.x; i < lim; i += blockDim.x)
        if (sh[i]) atomicAdd(&q.hist[i], sh[i]);
}

// torch.quantile(values, tensor(percentiles) / 100) (augmentations_3d.py:75), 'linear': ranks = q * (n - 1) in fp32,
// below = trunc, above = ceil, weight = ranks - below.
__global__ void prep_pick_kernel(int pass, float q_lo, float q_hi, PrepState* __restrict__ st) {
    PrepState& q = st[blockIdx.x];
    const int s = threadIdx.x;
    if (s < PSEL) {
        unsigned int* h = q.hist + (pass == 0 ? 0 : s * 256);
        if (pass == 0) {
            unsigned long long total = 0;
            for (int i = 0; i < 256; ++i) total += h[i];
            const float last = static_cast<float>(total > 0 ? total - 1 : 0);
            const float r = (s < 2 ? q_lo : q_hi) * last;
            const long long below = static_cast<long long>(r);
            const long long above = static_cast<long long>(ceilf(r));
            q.rank[s] = static_cast<unsigned long long>((s & 1) ? above : below);
            if (!(s & 1)) q.w[s >> 1] = r - static_cast<float>(below);
            if (s == 0) q.count = total;
        }
        unsigned long long r = q.rank[s];
        int bin = 0;
        for (; bin < 255; ++bin) {
            if (r < h[bin]) break;
            r -= h[bin];
        }
        q.rank[s] = r;
        q.prefix[s] |= static_cast<uint32_t>(bin) << (24 - 8 * pass);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PSEL * 256; i += blockDim.x) q.hist[i] = 0;
}

// ATen lerp (values_below.lerp_(values_above, weights)): w < 0.5 ? a + w (b - a) : b - (b - a)(1 - w), fused multiply-add
__device__ __forceinline__ float aten_lerp(float a, float b, float w) {
    const float diff = b - a;
    return fabsf(w) < 0.5f ? fmaf(w, diff, a) : fmaf(w - 1.0f, diff, b);
}
__global__ void prep_cutoff_kernel(PrepState* __restrict__ st) {
    PrepState& q = st[blockIdx.x];
    if (threadIdx.x == 0) {
        q.lo = aten_lerp(pkey2f(q.prefix[0]), pkey2f(q.prefix[1]), q.w[0]);
        q.hi = aten_lerp(pkey2f(q.prefix[2]), pkey2f(q.prefix[3]), q.w[1]);
    }
}

// mean / unbiased std of the CLAMPED values over the mask (tio.ZNormalization.znorm: values = tensor[mask]; mean(), std());
// accumulated in fp64 on (value - lo) >= 0.
__global__ void __launch_bounds__(256) prep_moments_kernel(const float* __restrict__ vol, int64_t n, PrepState* __restrict__ st) {
    __shared__ double ssum[8], ssq[8];
    const int item = blockIdx.y;
    PrepState& q = st[item];
    const uint32_t kmin = q.kmin, kmax = q.kmax;
    const float lo = q.lo, hi = q.hi;
    double sum = 0.0, sq = 0.0;
    const float4* v4 = reinterpret_cast<const float4*>(vol + static_cast<int64_t>(item) * n);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n / 4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float4 v = v4[i];
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t k = pf2key(e[j]);
            if (k > kmin && k < kmax) {
                const float c = fminf(fmaxf(e[j], lo), hi) - lo;
                sum += static_cast<double>(c);
                sq += static_cast<double>(c) * static_cast<double>(c);
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
        sq += __shfl_xor_sync(0xffffffffu, sq, off);
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { ssum[warp] = sum; ssq[warp] = sq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { sum += ssum[i]; sq += ssq[i]; }
        atomicAdd(&q.sum, sum);
        atomicAdd(&q.sumsq, sq);
    }
}

// stats [items, 8] fp64: min, max, cutoff_lo, cutoff_hi, mean, std, masked count, status (0 ok, 1 std == 0, 2 mask empty)
__global__ void prep_finalize_kernel(PrepState* __restrict__ st, double* __restrict__ stats) {
    PrepState& q = st[blockIdx.x];
    if (threadIdx.x != 0) return;
    const double n = static_cast<double>(q.count);
    double mean = 0.0, var = 0.0;
    int status = 0;
    if (q.count == 0) status = 2;
    else {
        mean = q.sum / n;
        var = q.count > 1 ? (q.sumsq - q.sum * mean) / (n - 1.0) : CUDART_NAN;
        if (var < 0.0) var = 0.0;
        mean += static_cast<double>(q.lo);
    }
    const double sd = sqrt(var);
    q.mean = static_cast<float>(mean);
    q.stdv = static_cast<float>(sd);
    if (status == 0 && q.stdv == 0.f) status = 1;   // the reference raises RuntimeError (augmentations_3d.py:79-84)
    if (stats) {
        double* o = stats + blockIdx.x * 8;
        o[0] = pkey2f(q.kmin); o[1] = pkey2f(q.kmax); o[2] = q.lo; o[3] = q.hi; o[4] = q.mean; o[5] = q.stdv; o[6] = n;
        o[7] = status;
    }
}

// torch.clamp(image, lo, hi); tensor -= mean; tensor /= std  (all voxels, fp32, two roundings like the reference)
__global__ void __launch_bounds__(256) prep_normalize_kernel(float* __restrict__ vol, int64_t n, const PrepState* __restrict__ st) {
    const int item = blockIdx.y;
    const float lo = st[item].lo, hi = st[item].hi, mean = st[item].mean, sd = st[item].stdv;
    float4* v4 = reinterpret_cast<float4*>(vol + static_cast<int64_t>(item) * n);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n / 4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float4 v = v4[i];
        v.x = __fdiv_rn(fminf(fmaxf(v.x, lo), hi) - mean, sd);
        v.y = __fdiv_rn(fminf(fmaxf(v.y, lo), hi) - mean, sd);
        v.z = __fdiv_rn(fminf(fmaxf(v.z, lo), hi) - mean, sd);
        v.w = __fdiv_rn(fminf(fmaxf(v.w, lo), hi) - mean, sd);
        v4[i] = v;
    }
}

inline int64_t margin_floats(int W0, int H0, int D0) {
    return static_cast<int64_t>(H0) * D0 + static_cast<int64_t>(W0) * D0 + static_cast<int64_t>(W0) * H0 + D0 + H0 + W0 + 4;
}
inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

}  // namespace

size_t prepare_volume_workspace_bytes(int items, int W0, int H0, int D0) {
    return align256(static_cast<size_t>(items) * sizeof(PrepState)) +
           align256(static_cast<size_t>(items) * margin_floats(W0, H0, D0) * sizeof(float));
}

int launch_prepare_volume(const float* src, int items, int W0, int H0, int D0, int W, int H, int D, int flip_h, float q_lo,
                          float q_hi, float* out, double* stats, void* workspace, int num_sms, cudaStream_t stream) {
    const int64_t n = static_cast<int64_t>(W) * H * D;
    MST_REQUIRE(n % 4 == 0, "prepare_volume: target volume W*H*D = %lld must be a multiple of 4", (long long)n);
    MST_REQUIRE(items >= 1 && items <= 65535, "prepare_volume: bad item count %d", items);
    PrepState* st = static_cast<PrepState*>(workspace);
    float* mg = reinterpret_cast<float*>(static_cast<char*>(workspace) + align256(static_cast<size_t>(items) * sizeof(PrepState)));

    // CropOrPad._get_six_bounds_parameters (augmentations_3d.py:166-175): ini = ceil(n / 2), fin = n - ini, for the padding and
    // for the cropping alike; tio.Pad then tio.Crop (:190-194).  An axis is padded or cropped, never both.
    PrepGeom g{W0, H0, D0, W, H, D, 0, 0, 0, flip_h};
    const int src_n[3] = {W0, H0, D0}, dst_n[3] = {W, H, D};
    int off[3];
    bool any_pad = false;
    for (int a = 0; a < 3; ++a) {
        const int diff = dst_n[a] - src_n[a];
        if (diff > 0) { off[a] = -((diff + 1) / 2); any_pad = true; }   // pad_ini in front
        else off[a] = ((-diff) + 1) / 2;                                 // crop_ini removed in front
    }
    g.ow = off[0]; g.oh = off[1]; g.od = off[2];

    prep_init_kernel<<<items, 256, 0, stream>>>(st);
    MST_CHECK_CUDA(cudaGetLastError());

    PrepMargins pm{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    if (any_pad) {
        const int64_t per = margin_floats(W0, H0, D0);
        float* m_w = mg;
        float* m_h = m_w + static_cast<int64_t>(H0) * D0;
        float* m_d = m_h + static_cast<int64_t>(W0) * D0;
        float* m_wh = m_d + static_cast<int64_t>(W0) * H0;
        float* m_wd = m_wh + D0;
        float* m_hd = m_wd + H0;
        float* m_whd = m_hd + W0;
        const int64_t vol = static_cast<int64_t>(W0) * H0 * D0;
        auto amin = [&](const float* s, int64_t s_item, int n0, int n1, int n2, int axis, float* d) -> int {
            const int64_t total = static_cast<int64_t>(n0) * n1 * n2 / (axis == 0 ? n0 : (axis == 1 ? n1 : n2));
            int gx = static_cast<int>((total + 255) / 256);
            gx = gx < 1 ? 1 : (gx > 8 * num_sms ? 8 * num_sms : gx);
            axis_min_kernel<<<dim3(gx, items), 256, 0, stream>>>(s, s_item, n0, n1, n2, axis, d, per);
            MST_CHECK_CUDA(cudaGetLastError());
            return 0;
        };
        MST_PROPAGATE(amin(src, vol, W0, H0, D0, 0, m_w));     // [H0, D0]
        MST_PROPAGATE(amin(src, vol, W0, H0, D0, 1, m_h));     // [W0, D0]
        MST_PROPAGATE(amin(src, vol, W0, H0, D0, 2, m_d));     // [W0, H0]
        MST_PROPAGATE(amin(m_w, per, 1, H0, D0, 1, m_wh));     // [D0]
        MST_PROPAGATE(amin(m_w, per, 1, H0, D0, 2, m_wd));     // [H0]
        MST_PROPAGATE(amin(m_h, per, W0, 1, D0, 2, m_hd));     // [W0]
        MST_PROPAGATE(amin(m_wh, per, 1, 1, D0, 2, m_whd));    // [1]
        pm = PrepMargins{m_w, m_h, m_d, m_wh, m_wd, m_hd, m_whd, per};
    }

    prep_gather_kernel<<<dim3((W + 31) / 32, H, items), dim3(32, 8), 0, stream>>>(src, g, pm, out, st);
    MST_CHECK_CUDA(cudaGetLastError());

    int gx = static_cast<int>((n / 4 + 256 * 8 - 1) / (256 * 8));
    const int cap = (4 * num_sms + items - 1) / items;
    gx = gx < 1 ? 1 : (gx > cap ? cap : gx);
    for (int pass = 0; pass < 4; ++pass) {
        prep_hist_kernel<<<dim3(gx, items), 256, 0, stream>>>(out, n, pass, st);
        MST_CHECK_CUDA(cudaGetLastError());
        prep_pick_kernel<<<items, 256, 0, stream>>>(pass, q_lo, q_hi, st);
        MST_CHECK_CUDA(cudaGetLastError());
    }
    prep_cutoff_kernel<<<items, 32, 0, stream>>>(st);
    MST_CHECK_CUDA(cudaGetLastError());
    prep_moments_kernel<<<dim3(gx, items), 256, 0, stream>>>(out, n, st);
    MST_CHECK_CUDA(cudaGetLastError());
    prep_finalize_kernel<<<items, 32, 0, stream>>>(st, stats);
    MST_CHECK_CUDA(cudaGetLastError());
    prep_normalize_kernel<<<dim3(gx, items), 256, 0, stream>>>(out, n, st);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mst
