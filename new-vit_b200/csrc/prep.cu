// Input pipeline in front of the forward: what DUKE_Dataset3D's default transform chain does to one
// volume between the HDF5 read and the model's `source` tensor (reference mst/data/datasets/dataset_3d_duke.py:36-47, with
// image_resize / resample None and the random augmentations off, i.e. the predict configuration):
//   tio.Flip(1)                                              flip the H axis of the torchio tensor [C=1, W, H, D]
//   CropOrPad(image_crop, padding_mode='minimum')            augmentations_3d.py:144-195 (ini = ceil(n/2), fin = n - ini;
//                                                            pad first, then crop; tio.Pad -> np.pad(mode='minimum'))
//   ZNormalization(percentiles=(0.5, 99.5), masking_method = (x > x.min()) & (x < x.max()))   augmentations_3d.py:41-86
//   ImageOrSubjectToTensor                                   swapaxes(1, -1): [C, W, H, D] -> [C, D, H, W]  (:23-29)
// All of it is HBM-bound element work: one gather pass (flip, crop, pad and the axis swap through a shared-memory tile, so
// that both the reads along D and the writes along W are coalesced), a masked 3-pass radix select (12 + 10 + 10 key bits) for the two cutoffs, one
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

// f(float4) over a volume, four independent 16-byte loads in flight per thread (a loop with shared-memory atomics or branches
// in its body is not unrolled by the compiler: one load per round trip ran at 2.4 TB/s, this form at the copy rate).
template <typename F>
__device__ __forceinline__ void for_each_vec4(const float4* __restrict__ v4, int64_t n4, F f) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 a = __ldg(v4 + i), b = __ldg(v4 + i + stride), c = __ldg(v4 + i + 2 * stride), d = __ldg(v4 + i + 3 * stride);
        f(a); f(b); f(c); f(d);
    }
    for (; i < n4; i += stride) f(__ldg(v4 + i));
}

constexpr int PSEL = 4;  // order statistics per item: below / above rank of the low and of the high percentile
// Radix digits of the 32-bit key, most significant first: 12 + 10 + 10 bits = three passes.  (8-bit digits took four, and the first
// one discriminates badly: sign + seven exponent bits put 68 % of an MRI-like volume into one bin, so the second pass still
// counted most voxels with atomics.)
__device__ __forceinline__ int radix_shift(int pass) { return pass == 0 ? 20 : (pass == 1 ? 10 : 0); }
__device__ __forceinline__ int radix_bins(int pass) { return pass == 0 ? 4096 : 1024; }
constexpr int RADIX_PASSES = 3;

struct PrepState {  // per item, in the workspace
    uint32_t kmin, kmax;            // keys of the volume's min / max (after crop / pad)
    uint32_t prefix[PSEL];          // radix-select state
    unsigned long long rank[PSEL];
    unsigned long long count;       // masked voxels
    float w[2];                     // interpolation weights of the two percentiles
    float lo, hi;                   // cutoffs
    float mean, stdv;
    double sum, sumsq;              // of (clamped - lo) over the mask
    unsigned int hist[4096];        // first pass: 4096 shared bins; later passes: 4 x 1024
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
    int vec_ok;         // 16-byte loads along D and stores along W are aligned (pointers included)
};
struct PrepMargins {    // per-item tables, nullptr when no axis is padded
    const float *m_w, *m_h, *m_d, *m_wh, *m_wd, *m_hd, *m_whd;
    int64_t item;       // floats per item in the margin block
};

// Address of the voxel (or of the marginal minimum) that target voxel (w, h, d) takes: address arithmetic only, so that the
// caller can issue all its loads back to back.
__device__ __forceinline__ const float* prep_addr(const float* __restrict__ s, const PrepGeom& g, const PrepMargins& pm,
                                                  int64_t moff, int w, int h, int d) {
    const int sw = w + g.ow, sd = d + g.od;
    int sh = h + g.oh;
    const bool iw = sw >= 0 && sw < g.W0, ih = sh >= 0 && sh < g.H0, id = sd >= 0 && sd < g.D0;
    if (g.flip_h) sh = g.H0 - 1 - sh;   // tio.Flip(1) runs before the pad; minima along H do not see it
    if (iw && ih && id) return s + (static_cast<int64_t>(sw) * g.H0 + sh) * g.D0 + sd;
    if (!iw && ih && id) return pm.m_w + moff + static_cast<int64_t>(sh) * g.D0 + sd;
    if (iw && !ih && id) return pm.m_h + moff + static_cast<int64_t>(sw) * g.D0 + sd;
    if (iw && ih && !id) return pm.m_d + moff + static_cast<int64_t>(sw) * g.H0 + sh;
    if (!iw && !ih && id) return pm.m_wh + moff + sd;
    if (!iw && ih && !id) return pm.m_wd + moff + sh;
    if (iw && !ih && !id) return pm.m_hd + moff + sw;
    return pm.m_whd + moff;
}

// out[item, d, h, w] = padded / cropped / flipped src[item, w, h, d]; per-item min / max keys by atomics.
// grid (ceil(W/32), ceil(H/GATHER_HB), items), block (32, 8): GATHER_HB tiles of 32(w) x 32(d) through shared memory per step,
// so that every thread has 4 * GATHER_HB independent loads in flight (one tile per block and step: 1.0 TB/s, latency-bound).
constexpr int GATHER_HB = 4;
__global__ void __launch_bounds__(256, 4) prep_gather_kernel(const float* __restrict__ src, PrepGeom g, PrepMargins pm,
                                                           float* __restrict__ out, PrepState* __restrict__ st) {
    __shared__ float tile[GATHER_HB][32][33];
    __shared__ uint32_t smin[8], smax[8];
    const int item = blockIdx.z, h0 = blockIdx.y * GATHER_HB, w0 = blockIdx.x * 32;
    const float* s = src + static_cast<int64_t>(item) * g.W0 * g.H0 * g.D0;
    float* o = out + static_cast<int64_t>(item) * g.W * g.H * g.D;
    const int64_t moff = static_cast<int64_t>(item) * pm.item;
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    const int w1 = min(w0 + 32, g.W) - 1, h1 = min(h0 + GATHER_HB, g.H) - 1;
    const bool in_wh = w0 + g.ow >= 0 && w1 + g.ow < g.W0 && h0 + g.oh >= 0 && h1 + g.oh < g.H0;
    const bool vec_ok = g.vec_ok && w0 + 32 <= g.W;
    for (int d0 = 0; d0 < g.D; d0 += 32) {
        const bool interior = in_wh && d0 + g.od >= 0 && min(d0 + 32, g.D) - 1 + g.od < g.D0;
        const bool vec = vec_ok && d0 + 32 <= g.D;
        float v[GATHER_HB][4];
        if (interior && vec) {
            // 16-byte path (D0, the D offset and W multiples of 4, full 32 x 32 tiles): one LDG.128 along D and one STG.128
            // along W per thread and tile, the 4 x 4 transposition through scalar shared-memory accesses (row stride 33:
            // conflict-free both ways).  The scalar path was issue-bound (ncu: issue slots 76 %, DRAM 50 %).
            const int t = threadIdx.y * 32 + threadIdx.x;
            const int lw = t >> 3, ld4 = (t & 7) * 4;          // load: tile row (w) and first of four d
            const int sw4 = (t & 7) * 4, sd = t >> 3;           // store: first of four w, d
            float4 r[GATHER_HB];
#pragma unroll
            for (int hb = 0; hb < GATHER_HB; ++hb) {
                const int sh = g.flip_h ? g.H0 - 1 - (h0 + hb + g.oh) : h0 + hb + g.oh;
                r[hb] = h0 + hb < g.H ? __ldg(reinterpret_cast<const float4*>(
                                            s + (static_cast<int64_t>(w0 + lw + g.ow) * g.H0 + sh) * g.D0 + (d0 + ld4 + g.od)))
                                      : make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
            }
#pragma unroll
            for (int hb = 0; hb < GATHER_HB; ++hb) {
                tile[hb][lw][ld4 + 0] = r[hb].x; tile[hb][lw][ld4 + 1] = r[hb].y;
                tile[hb][lw][ld4 + 2] = r[hb].z; tile[hb][lw][ld4 + 3] = r[hb].w;
                if (h0 + hb < g.H) {
                    const uint32_t k0 = pf2key(r[hb].x), k1 = pf2key(r[hb].y), k2 = pf2key(r[hb].z), k3 = pf2key(r[hb].w);
                    kmin = min(min(kmin, min(k0, k1)), min(k2, k3));
                    kmax = max(max(kmax, max(k0, k1)), max(k2, k3));
                }
            }
            __syncthreads();
#pragma unroll
            for (int hb = 0; hb < GATHER_HB; ++hb) {
                if (h0 + hb < g.H) {
                    const float4 w4 = make_float4(tile[hb][sw4 + 0][sd], tile[hb][sw4 + 1][sd], tile[hb][sw4 + 2][sd], tile[hb][sw4 + 3][sd]);
                    *reinterpret_cast<float4*>(o + (static_cast<int64_t>(d0 + sd) * g.H + (h0 + hb)) * g.W + (w0 + sw4)) = w4;
                }
            }
            __syncthreads();
            continue;
        }
        if (interior) {
            // every voxel of the tile lies inside the source: the address is affine in (k, hb), no table lookups, few registers
            const int sh0 = g.flip_h ? g.H0 - 1 - (h0 + g.oh) : h0 + g.oh;
            const float* base = s + (static_cast<int64_t>(w0 + threadIdx.y + g.ow) * g.H0 + sh0) * g.D0 + (d0 + threadIdx.x + g.od);
            const int64_t sk = static_cast<int64_t>(8) * g.H0 * g.D0;
            const int shs = g.flip_h ? -g.D0 : g.D0;
#pragma unroll
            for (int hb = 0; hb < GATHER_HB; ++hb)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool ok = w0 + threadIdx.y + 8 * k < g.W && d0 + threadIdx.x < g.D && h0 + hb < g.H;
                    v[hb][k] = ok ? __ldg(base + k * sk + hb * shs) : CUDART_NAN_F;
                }
        } else {
#pragma unroll
            for (int hb = 0; hb < GATHER_HB; ++hb)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int w = w0 + threadIdx.y + 8 * k, d = d0 + threadIdx.x, h = h0 + hb;
                    v[hb][k] = (w < g.W && d < g.D && h < g.H) ? __ldg(prep_addr(s, g, pm, moff, w, h, d)) : CUDART_NAN_F;
                }
        }
#pragma unroll
        for (int hb = 0; hb < GATHER_HB; ++hb)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int wl = threadIdx.y + 8 * k;
                tile[hb][wl][threadIdx.x] = v[hb][k];
                if (w0 + wl < g.W && d0 + threadIdx.x < g.D && h0 + hb < g.H) {
                    const uint32_t key = pf2key(v[hb][k]);
                    kmin = min(kmin, key);
                    kmax = max(kmax, key);
                }
            }
        __syncthreads();
#pragma unroll
        for (int hb = 0; hb < GATHER_HB; ++hb)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int dl = threadIdx.y + 8 * k, d = d0 + dl, w = w0 + threadIdx.x, h = h0 + hb;
                if (w < g.W && d < g.D && h < g.H) o[(static_cast<int64_t>(d) * g.H + h) * g.W + w] = tile[hb][threadIdx.x][dl];
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
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) q.hist[i] = 0;
    if (threadIdx.x == 0) {
        q.kmin = 0xffffffffu; q.kmax = 0u; q.count = 0; q.sum = 0.0; q.sumsq = 0.0;
        for (int s = 0; s < PSEL; ++s) { q.prefix[s] = 0; q.rank[s] = 0; }
    }
}

// One radix pass over the masked voxels (min < x < max, augmentations_3d.py:75 masked_select with dataset_3d_duke.py:43's mask).
__global__ void __launch_bounds__(256) prep_hist_kernel(const float* __restrict__ vol, int64_t n, int pass,
                                                         PrepState* __restrict__ st) {
    __shared__ unsigned int sh[4096];
    const int item = blockIdx.y;
    PrepState& q = st[item];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 0;
    const uint32_t kmin = q.kmin, kmax = q.kmax;
    // in registers: read from shared memory inside the loop they would be re-loaded after every atomic (possible alias)
    const uint32_t p0 = q.prefix[0], p1 = q.prefix[1], p2 = q.prefix[2], p3 = q.prefix[3];
    // a selection counts only under a prefix no earlier selection has (prep_pick_kernel reads the first equal one's histogram)
    const bool u1 = p1 != p0, u2 = p2 != p0 && p2 != p1, u3 = p3 != p0 && p3 != p1 && p3 != p2;
    __syncthreads();
    const int shift = radix_shift(pass);
    const uint32_t himask = pass == 0 ? 0u : (0xffffffffu << radix_shift(pass - 1));
    const float4* v4 = reinterpret_cast<const float4*>(vol + static_cast<int64_t>(item) * n);  // n % 4 == 0 (checked by the host)
    if (pass == 0) {
        for_each_vec4(v4, n / 4, [&](const float4 v) {   // all four selections share the first histogram
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t k = pf2key(e[j]);
                if (k > kmin && k < kmax) atomicAdd(&sh[k >> 20], 1u);
            }
        });
    } else {
        // Later passes: the four voxels of a 16-byte load share ONE branch (a branch per voxel and selection made these passes
        // instruction-bound).  The below / above ranks of a percentile are adjacent and nearly always share their prefix (u1 and
        // u3 false): then a voxel counts under p0 or under p2, one predicated atomic.  Some lane of nearly every warp takes the
        // counting path in the second pass (about 4 % of an MRI-like volume shares the upper cutoff's 12-bit prefix), so its
        // instruction count is what that pass costs.
        for_each_vec4(v4, n / 4, [&](const float4 v) {
            const float e[4] = {v.x, v.y, v.z, v.w};
            uint32_t k[4];
            bool any = false;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                k[j] = pf2key(e[j]);
                const uint32_t kp = k[j] & himask;
                any |= (kp == p0) | (kp == p2) | (u1 & (kp == p1)) | (u3 & (kp == p3));
            }
            if (any) {
                if (!u1 && !u3) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t kp = k[j] & himask, bin = (k[j] >> shift) & 0x3ffu;
                        const bool m0 = kp == p0, m2 = u2 && kp == p2;
                        if ((m0 || m2) && k[j] > kmin && k[j] < kmax) atomicAdd(&sh[(m0 ? 0u : 2048u) + bin], 1u);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (k[j] > kmin && k[j] < kmax) {
                            const uint32_t kp = k[j] & himask, bin = (k[j] >> shift) & 0x3ffu;
                            if (kp == p0) atomicAdd(&sh[bin], 1u);
                            if (u1 && kp == p1) atomicAdd(&sh[1024 + bin], 1u);
                            if (u2 && kp == p2) atomicAdd(&sh[2048 + bin], 1u);
                            if (u3 && kp == p3) atomicAdd(&sh[3072 + bin], 1u);
                        }
                    }
                }
            }
        });
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x)
        if (sh[i]) atomicAdd(&q.hist[i], sh[i]);
}

// torch.quantile(values, tensor(percentiles) / 100) (augmentations_3d.py:75), 'linear': ranks = q * (n - 1) in fp32,
// below = trunc, above = ceil, weight = ranks - below.
__global__ void __launch_bounds__(256) prep_pick_kernel(int pass, float q_lo, float q_hi, PrepState* __restrict__ st) {
    __shared__ unsigned long long excl[PSEL][256];   // voxels in the bins before thread t's group, per selection
    __shared__ unsigned long long wsum[PSEL][8];
    __shared__ unsigned long long total0;
    PrepState& q = st[blockIdx.x];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int nb = radix_bins(pass), per = nb / 256;   // consecutive bins per thread: 16, then 4
    int src[PSEL];
#pragma unroll
    for (int s = 0; s < PSEL; ++s) {
        // histogram of selection s: shared in the first pass; later the one of the FIRST selection with the same prefix
        src[s] = pass == 0 ? 0 : s * 1024;
        if (pass != 0)
            for (int e = s - 1; e >= 0; --e)
                if (q.prefix[e] == q.prefix[s]) src[s] = e * 1024;
        unsigned long long mine = 0;
        for (int i = 0; i < per; ++i) mine += q.hist[src[s] + t * per + i];
        unsigned long long v = mine;
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, v, off);
            if (lane >= off) v += u;
        }
        excl[s][t] = v - mine;
        if (lane == 31) wsum[s][warp] = v;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < PSEL; ++s) {
        unsigned long long base = 0;
        for (int i = 0; i < warp; ++i) base += wsum[s][i];
        excl[s][t] += base;
    }
    if (t == 0) {
        unsigned long long tot = 0;
        for (int i = 0; i < 8; ++i) tot += wsum[0][i];
        total0 = tot;
    }
    __syncthreads();
    if (pass == 0 && t < PSEL) {
        const unsigned long long total = total0;
        const float last = static_cast<float>(total > 0 ? total - 1 : 0);
        const float r = (t < 2 ? q_lo : q_hi) * last;
        const long long below = static_cast<long long>(r);
        const long long above = static_cast<long long>(ceilf(r));
        q.rank[t] = static_cast<unsigned long long>((t & 1) ? above : below);
        if (!(t & 1)) q.w[t >> 1] = r - static_cast<float>(below);
        if (t == 0) q.count = total;
    }
    __syncthreads();
    // thread t owns the rank when excl[t] <= rank < excl[t+1]; a rank past the end (empty mask) falls into the last bin
#pragma unroll
    for (int s = 0; s < PSEL; ++s) {
        const unsigned long long r = q.rank[s];
        const unsigned long long lo = excl[s][t];
        const bool last_t = t == 255;
        const unsigned long long hi = last_t ? ~0ull : excl[s][t + 1];
        __syncthreads();
        if (r >= lo && r < hi) {
            unsigned long long rem = r - lo;
            int bin = t * per;
            for (int i = 0; i < per - 1; ++i) {
                const unsigned long long c = q.hist[src[s] + bin];
                if (rem < c) break;
                rem -= c;
                ++bin;
            }
            const unsigned long long c = q.hist[src[s] + bin];
            q.rank[s] = rem < c ? rem : 0ull;
            q.prefix[s] |= static_cast<uint32_t>(bin) << radix_shift(pass);
        }
    }
    __syncthreads();
    for (int i = t; i < 4096; i += blockDim.x) q.hist[i] = 0;
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
    for_each_vec4(v4, n / 4, [&](const float4 v) {
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
    });
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
// raw scanner voxels (int16 / uint16, as the HDF5 / DICOM files hold them) -> fp32, so that the host-to-device copy of the raw-data
// route carries 2 bytes per voxel; the fp32 copy lives in the caller's workspace behind the state of prepare_volume
template <typename TI>
__global__ void __launch_bounds__(256) raw_to_f32_kernel(const TI* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = static_cast<float>(in[i]);
}
size_t prepare_volume_raw_bytes(int items, int W0, int H0, int D0) { return align256(static_cast<size_t>(items) * W0 * H0 * D0 * sizeof(float)); }
int launch_raw_to_f32(const void* in, int dtype, float* out, int64_t n, int num_sms, cudaStream_t stream) {
    if (dtype == 3) raw_to_f32_kernel<int16_t><<<num_sms * 8, 256, 0, stream>>>(static_cast<const int16_t*>(in), out, n);
    else if (dtype == 4) raw_to_f32_kernel<uint16_t><<<num_sms * 8, 256, 0, stream>>>(static_cast<const uint16_t*>(in), out, n);
    else MST_REQUIRE(false, "prepare_volume: source dtype %d is not int16 / uint16", dtype);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// kernels one launch_prepare_volume call issues: init, (7 marginal-minimum tables when an axis is padded), gather,
// RADIX_PASSES x (histogram, pick), cutoff, moments, finalize, normalize
int launch_prepare_volume_count(int W0, int H0, int D0, int W, int H, int D) {
    const bool any_pad = W > W0 || H > H0 || D > D0;
    return 1 + (any_pad ? 7 : 0) + 1 + 2 * RADIX_PASSES + 4;
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
    PrepGeom g{W0, H0, D0, W, H, D, 0, 0, 0, flip_h, 0};
    const int src_n[3] = {W0, H0, D0}, dst_n[3] = {W, H, D};
    int off[3];
    bool any_pad = false;
    for (int a = 0; a < 3; ++a) {
        const int diff = dst_n[a] - src_n[a];
        if (diff > 0) { off[a] = -((diff + 1) / 2); any_pad = true; }   // pad_ini in front
        else off[a] = ((-diff) + 1) / 2;                                 // crop_ini removed in front
    }
    g.ow = off[0]; g.oh = off[1]; g.od = off[2];
    g.vec_ok = D0 % 4 == 0 && ((g.od % 4) + 4) % 4 == 0 && W % 4 == 0 && reinterpret_cast<uintptr_t>(src) % 16 == 0 &&
               reinterpret_cast<uintptr_t>(out) % 16 == 0;

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

    prep_gather_kernel<<<dim3((W + 31) / 32, (H + GATHER_HB - 1) / GATHER_HB, items), dim3(32, 8), 0, stream>>>(src, g, pm, out, st);
    MST_CHECK_CUDA(cudaGetLastError());

    int gx = static_cast<int>((n / 4 + 256 * 8 - 1) / (256 * 8));
    const int cap = (4 * num_sms + items - 1) / items;
    gx = gx < 1 ? 1 : (gx > cap ? cap : gx);
    for (int pass = 0; pass < RADIX_PASSES; ++pass) {
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
