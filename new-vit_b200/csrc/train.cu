// Training step of the slice transformer + head on top of a FROZEN encoder (BASELINE.json config 5, the `freeze=True`
// construction of reference dino.py:69-71): forward with saved activations, backward, AdamW.
//
// The encoder output per slice (enc [B, D, E] fp32, no gradient) comes out of mst_forward; what trains is reference
// dino.py:84-103: cls_token, slice_fusion.layers.0.{norm1, self_attn.in_proj, self_attn.out_proj, norm2, linear1, linear2},
// slice_fusion.norm and the linear head (1.48 M parameters for ViT-S).  Loss and its gradient w.r.t. the logits stay with the
// caller (base_model.py:159,180-181: CrossEntropyLoss on [B, out_ch]).
//
// One layer, and only token 0 of its output is consumed (dino.py:153), so -- as in the inference kernel (kernels.cu) -- only the
// slice-CLS query is evaluated and K / V are never materialised:
//   s[h,j] = (Wk_h^T q_h) . n_j + q_h . bk_h          o_h = Wv_h (sum_j p[h,j] n_j) + bv_h
// The backward pass re-associates the same way; every weight gradient of one volume is an outer product of two vectors
//   dW2 = dx2 (x) f   dW1 = dpre (x) n2   dWo = dx1 (x) o   dWq = dq (x) n_0   dWk_h = q_h (x) g_h   dWv_h = do_h (x) hbar_h
// (g_h = sum_j ds[h,j] n_j), so the kernels are: (A) one CTA per volume producing those vectors, (B) one CTA per weight row
// summing the B rank-1 terms in a fixed order -- no atomics, deterministic.  All fp32; weights in nn.Linear layout [out][in].
#include <math_constants.h>
#include "common.cuh"

namespace mst {

namespace {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float bsum(float v, float* red) {   // all threads get the block-wide sum; red: >= 33 floats
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = wsum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nw ? red[lane] : 0.f;
        t = wsum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
// out[n] = (W[n,:] . x + bias[n]) * scale, n in [0, N): one warp per row, lanes stride the (contiguous) input dimension
__device__ __forceinline__ void rows_dot(const float* __restrict__ W, int ld, const float* x, const float* __restrict__ bias,
                                         float* out, int N, int K, float scale) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int n = warp; n < N; n += nw) {
        const float* w = W + static_cast<int64_t>(n) * ld;
        float a = 0.f;
        for (int k = lane; k < K; k += 32) a = fmaf(__ldg(w + k), x[k], a);
        a = wsum(a);
        if (lane == 0) out[n] = (a + (bias ? bias[n] : 0.f)) * scale;
    }
    __syncthreads();
}
// out[k] = sum_n W[n,k] * u[n], k in [0, K): thread per column (coalesced across k), rows split over blockDim / K thread groups
__device__ __forceinline__ void cols_dot(const float* __restrict__ W, int ld, const float* u, float* out, float* scr, int N, int K) {
    const int parts = max(1, static_cast<int>(blockDim.x) / K);
    const int part = threadIdx.x / K, k = threadIdx.x - part * K;
    if (part < parts) {
        const int nper = (N + parts - 1) / parts, n0 = part * nper, n1 = min(N, n0 + nper);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int n = n0;
        for (; n + 3 < n1; n += 4) {
            a0 = fmaf(__ldg(W + static_cast<int64_t>(n) * ld + k), u[n], a0);
            a1 = fmaf(__ldg(W + static_cast<int64_t>(n + 1) * ld + k), u[n + 1], a1);
            a2 = fmaf(__ldg(W + static_cast<int64_t>(n + 2) * ld + k), u[n + 2], a2);
            a3 = fmaf(__ldg(W + static_cast<int64_t>(n + 3) * ld + k), u[n + 3], a3);
        }
        for (; n < n1; ++n) a0 = fmaf(__ldg(W + static_cast<int64_t>(n) * ld + k), u[n], a0);
        scr[part * K + k] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        float a = 0.f;
        for (int p = 0; p < parts; ++p) a += scr[p * K + i];
        out[i] = a;
    }
    __syncthreads();
}
// y = LayerNorm(x) (eps 1e-5); st = {mean, rstd}
__device__ __forceinline__ void ln_fwd(const float* x, const float* __restrict__ g, const float* __restrict__ b, float* y, float* st,
                                       int E, float* red) {
    float s = 0.f;
    for (int i = threadIdx.x; i < E; i += blockDim.x) s += x[i];
    const float mean = bsum(s, red) / E;
    float q = 0.f;
    for (int i = threadIdx.x; i < E; i += blockDim.x) { const float d = x[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(bsum(q, red) / E + 1e-5f);
    for (int i = threadIdx.x; i < E; i += blockDim.x) y[i] = fmaf((x[i] - mean) * rstd, g[i], b[i]);
    if (threadIdx.x == 0) { st[0] = mean; st[1] = rstd; }
    __syncthreads();
}
// dx = LayerNorm backward of dy at x; dgam[i] (+)= dy*xhat, dbet[i] (+)= dy.  dx may alias dy.
__device__ __forceinline__ void ln_bwd(const float* x, float mean, float rstd, const float* __restrict__ g, const float* dy, float* dx,
                                       float* dgam, float* dbet, bool accumulate, int E, float* red) {
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
        const float xh = (x[i] - mean) * rstd, dxh = dy[i] * g[i];
        s1 += dxh; s2 = fmaf(dxh, xh, s2);
    }
    const float m1 = bsum(s1, red) / E;
    const float m2 = bsum(s2, red) / E;
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
        const float xh = (x[i] - mean) * rstd, d = dy[i];
        const float gam = d * xh;
        dgam[i] = accumulate ? dgam[i] + gam : gam;
        dbet[i] = accumulate ? dbet[i] + d : d;
        dx[i] = rstd * (d * g[i] - m1 - xh * m2);
    }
    __syncthreads();
}

}  // namespace

// floats of per-volume state kept between forward and backward
__host__ __device__ inline int64_t slice_train_saved_floats(int L, int E, int heads) {
    return static_cast<int64_t>(L) * E      // N1 = LN1 of every token
           + 2 * L                          // mean, rstd of every token
           + static_cast<int64_t>(heads) * L  // p
           + 2 * static_cast<int64_t>(heads) * E  // qk, hbar
           + 8 * static_cast<int64_t>(E)    // q, o, x1, n2, f, x2, y, (spare)
           + 8;                             // LN2 / final LN statistics
}
// floats of per-volume backward factors handed from kernel A to kernel B
__host__ __device__ inline int64_t slice_train_factor_floats(int E, int heads, int C) {
    return 12 * static_cast<int64_t>(E)     // dx2, dpre, dx1, dq, do, dcls, dg1, db1, dg2, db2, dgf, dbf
           + static_cast<int64_t>(heads) * E  // g
           + heads + C + 8;                 // sum_j ds[h,j], dlogits
}

struct SliceTrainParams {   // nn.Linear layout, fp32 (reference dino.py:84-103 state_dict tensors)
    const float *cls_token, *n1w, *n1b, *in_w, *in_b, *out_w, *out_b, *n2w, *n2b, *l1_w, *l1_b, *l2_w, *l2_b, *nfw, *nfb, *head_w, *head_b;
};
struct SliceTrainGrads {    // same shapes as the parameters
    float *cls_token, *n1w, *n1b, *in_w, *in_b, *out_w, *out_b, *n2w, *n2b, *l1_w, *l1_b, *l2_w, *l2_b, *nfw, *nfb, *head_w, *head_b;
};

// saved-state layout helpers (offsets in floats inside one volume's block)
struct SavedLayout {
    int64_t n1, st1, p, qk, hbar, q, o, x1, n2, f, x2, y, st2;
    __host__ __device__ SavedLayout(int L, int E, int heads) {
        int64_t off = 0;
        n1 = off; off += static_cast<int64_t>(L) * E;
        st1 = off; off += 2 * L;
        p = off; off += static_cast<int64_t>(heads) * L;
        qk = off; off += static_cast<int64_t>(heads) * E;
        hbar = off; off += static_cast<int64_t>(heads) * E;
        q = off; off += E; o = off; off += E; x1 = off; off += E; n2 = off; off += E;
        f = off; off += E; x2 = off; off += E; y = off; off += 2 * E;
        st2 = off;
    }
};
struct FactorLayout {
    int64_t dx2, dpre, dx1, dq, dov, dcls, dg1, db1, dg2, db2, dgf, dbf, g, sds, dl;
    __host__ __device__ FactorLayout(int E, int heads) {
        int64_t off = 0;
        dx2 = off; off += E; dpre = off; off += E; dx1 = off; off += E; dq = off; off += E; dov = off; off += E; dcls = off; off += E;
        dg1 = off; off += E; db1 = off; off += E; dg2 = off; off += E; db2 = off; off += E; dgf = off; off += E; dbf = off; off += E;
        g = off; off += static_cast<int64_t>(heads) * E;
        sds = off; off += heads;
        dl = off;
    }
};

// ---------------------------------------------------------------------------------------------------
// forward with saved activations: one CTA per volume
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(768) slice_train_forward_kernel(const float* __restrict__ enc, const uint8_t* __restrict__ pad_mask,
                                                                   SliceTrainParams w, float* __restrict__ saved_all,
                                                                   float* __restrict__ logits, int D, int E, int heads, int C) {
    extern __shared__ float sm[];
    const int L = D + 1, hd = E / heads, b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const SavedLayout S(L, E, heads);
    float* sv = saved_all + static_cast<int64_t>(b) * slice_train_saved_floats(L, E, heads);
    float* x0 = sm;            // [E] raw token 0
    float* v0 = x0 + E;        // [E] work
    float* v1 = v0 + E;        // [E] work
    float* v2 = v1 + E;        // [E] work
    float* cterm = v2 + E;     // [32]
    float* red = cterm + 32;   // [40]
    float* ps = red + 40;      // [heads][L]
    float* N1 = sv + S.n1;

    // tokens: slice-CLS in front of the D encoder outputs (dino.py:145); LN1 of every token (transformer_blocks.py:566)
    for (int l = warp; l < L; l += nw) {
        const float* src = l == 0 ? w.cls_token : enc + (static_cast<int64_t>(b) * D + l - 1) * E;
        float s = 0.f;
        for (int i = lane; i < E; i += 32) s += src[i];
        const float mean = wsum(s) / E;
        float q = 0.f;
        for (int i = lane; i < E; i += 32) { const float d = src[i] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(wsum(q) / E + 1e-5f);
        for (int i = lane; i < E; i += 32) {
            const float o = fmaf((src[i] - mean) * rstd, w.n1w[i], w.n1b[i]);
            N1[static_cast<int64_t>(l) * E + i] = o;
            if (l == 0) { x0[i] = src[i]; v0[i] = o; }
        }
        if (lane == 0) { sv[S.st1 + 2 * l] = mean; sv[S.st1 + 2 * l + 1] = rstd; }
    }
    __syncthreads();
    // q = (Wq n_0 + bq) / sqrt(hd)                                       (transformer_blocks.py:166,268)
    rows_dot(w.in_w, E, v0, w.in_b, v1, E, E, rsqrtf(static_cast<float>(hd)));
    for (int i = threadIdx.x; i < E; i += blockDim.x) sv[S.q + i] = v1[i];
    // qk[h][k] = sum_d q[h,d] Wk[h*hd+d][k];  cterm[h] = q_h . bk_h
    for (int idx = threadIdx.x; idx < heads * E; idx += blockDim.x) {
        const int h = idx / E, k = idx - h * E;
        const float* wr = w.in_w + (static_cast<int64_t>(E) + h * hd) * E + k;
        float a = 0.f;
#pragma unroll 8
        for (int d = 0; d < hd; ++d) a = fmaf(v1[h * hd + d], __ldg(wr + static_cast<int64_t>(d) * E), a);
        sv[S.qk + idx] = a;
    }
    if (threadIdx.x < heads) {
        float a = 0.f;
        for (int d = 0; d < hd; ++d) a = fmaf(v1[threadIdx.x * hd + d], w.in_b[E + threadIdx.x * hd + d], a);
        cterm[threadIdx.x] = a;
    }
    __syncthreads();
    // scores with the key-padding mask as -inf (transformer_blocks.py:244-252; CLS column never masked, dino.py:149-150)
    for (int task = warp; task < heads * L; task += nw) {
        const int h = task / L, j = task - h * L;
        const float* nr = N1 + static_cast<int64_t>(j) * E;
        const float* qr = sv + S.qk + static_cast<int64_t>(h) * E;
        float a = 0.f;
        for (int k = lane; k < E; k += 32) a = fmaf(qr[k], nr[k], a);
        a = wsum(a) + cterm[h];
        if (j > 0 && pad_mask && pad_mask[static_cast<int64_t>(b) * D + j - 1]) a = -CUDART_INF_F;
        if (lane == 0) ps[task] = a;
    }
    __syncthreads();
    for (int h = warp; h < heads; h += nw) {
        float m = -CUDART_INF_F;
        for (int j = lane; j < L; j += 32) m = fmaxf(m, ps[h * L + j]);
        m = wmax(m);
        float s = 0.f;
        for (int j = lane; j < L; j += 32) { const float e = expf(ps[h * L + j] - m); ps[h * L + j] = e; s += e; }
        const float inv = 1.0f / wsum(s);
        for (int j = lane; j < L; j += 32) { const float pj = ps[h * L + j] * inv; ps[h * L + j] = pj; sv[S.p + h * L + j] = pj; }
    }
    __syncthreads();
    // hbar[h][k] = sum_j p[h][j] n_j[k]
    for (int idx = threadIdx.x; idx < heads * E; idx += blockDim.x) {
        const int h = idx / E, k = idx - h * E;
        float a = 0.f;
        for (int j = 0; j < L; ++j) a = fmaf(ps[h * L + j], N1[static_cast<int64_t>(j) * E + k], a);
        sv[S.hbar + idx] = a;
    }
    __syncthreads();
    // o[n] = Wv[n,:] . hbar[head(n)] + bv[n]
    for (int n = warp; n < E; n += nw) {
        const float* wr = w.in_w + (static_cast<int64_t>(2) * E + n) * E;
        const float* hb = sv + S.hbar + static_cast<int64_t>(n / hd) * E;
        float a = 0.f;
        for (int k = lane; k < E; k += 32) a = fmaf(__ldg(wr + k), hb[k], a);
        a = wsum(a);
        if (lane == 0) { const float o = a + w.in_b[2 * E + n]; v0[n] = o; sv[S.o + n] = o; }
    }
    __syncthreads();
    // x1 = x0 + out_proj(o)                                               (transformer_blocks.py:566)
    rows_dot(w.out_w, E, v0, w.out_b, v1, E, E, 1.0f);
    for (int i = threadIdx.x; i < E; i += blockDim.x) { v1[i] += x0[i]; sv[S.x1 + i] = v1[i]; }
    __syncthreads();
    // x2 = x1 + W2 relu(W1 LN2(x1) + b1) + b2                             (transformer_blocks.py:567,585)
    ln_fwd(v1, w.n2w, w.n2b, v0, sv + S.st2, E, red);
    for (int i = threadIdx.x; i < E; i += blockDim.x) sv[S.n2 + i] = v0[i];
    rows_dot(w.l1_w, E, v0, w.l1_b, v2, E, E, 1.0f);
    for (int i = threadIdx.x; i < E; i += blockDim.x) { v2[i] = fmaxf(v2[i], 0.f); sv[S.f + i] = v2[i]; }
    __syncthreads();
    rows_dot(w.l2_w, E, v2, w.l2_b, v0, E, E, 1.0f);
    for (int i = threadIdx.x; i < E; i += blockDim.x) { v0[i] += v1[i]; sv[S.x2 + i] = v0[i]; }
    __syncthreads();
    // final LayerNorm (dino.py:95), token 0 (dino.py:153), head (dino.py:166)
    ln_fwd(v0, w.nfw, w.nfb, v1, sv + S.st2 + 2, E, red);
    for (int i = threadIdx.x; i < E; i += blockDim.x) sv[S.y + i] = v1[i];
    for (int c = warp; c < C; c += nw) {
        float a = 0.f;
        for (int k = lane; k < E; k += 32) a = fmaf(v1[k], w.head_w[static_cast<int64_t>(c) * E + k], a);
        a = wsum(a);
        if (lane == 0) logits[static_cast<int64_t>(b) * C + c] = a + w.head_b[c];
    }
}

// ---------------------------------------------------------------------------------------------------
// backward, kernel A: one CTA per volume -> the per-volume factor vectors
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(768) slice_train_backward_kernel(const float* __restrict__ enc, const float* __restrict__ dlogits,
                                                                    SliceTrainParams w, const float* __restrict__ saved_all,
                                                                    float* __restrict__ fact_all, float* __restrict__ denc, int D,
                                                                    int E, int heads, int C) {
    extern __shared__ float sm[];
    const int L = D + 1, hd = E / heads, b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const SavedLayout S(L, E, heads);
    const FactorLayout Fq(E, heads);
    const float* sv = saved_all + static_cast<int64_t>(b) * slice_train_saved_floats(L, E, heads);
    float* fc = fact_all + static_cast<int64_t>(b) * slice_train_factor_floats(E, heads, C);
    float* v0 = sm;             // [E]
    float* v1 = v0 + E;         // [E]
    float* v2 = v1 + E;         // [E]
    float* scr = v2 + E;        // [2 E] cols_dot partials
    float* red = scr + 2 * E;   // [40]
    float* ds = red + 40;       // [heads][L]
    float* vo = ds + ((heads * L + 3) & ~3);   // [heads][E]
    const float* N1 = sv + S.n1;
    const float* p = sv + S.p;

    // dy = head_w^T dlogits; final LayerNorm backward -> dx2
    for (int c = threadIdx.x; c < C; c += blockDim.x) fc[Fq.dl + c] = dlogits[static_cast<int64_t>(b) * C + c];
    for (int k = threadIdx.x; k < E; k += blockDim.x) {
        float a = 0.f;
        for (int c = 0; c < C; ++c) a = fmaf(w.head_w[static_cast<int64_t>(c) * E + k], dlogits[static_cast<int64_t>(b) * C + c], a);
        v0[k] = a;
    }
    __syncthreads();
    ln_bwd(sv + S.x2, sv[S.st2 + 2], sv[S.st2 + 3], w.nfw, v0, v0, fc + Fq.dgf, fc + Fq.dbf, false, E, red);   // v0 = dx2
    for (int i = threadIdx.x; i < E; i += blockDim.x) fc[Fq.dx2 + i] = v0[i];
    // FFN backward: df = W2^T dx2; dpre = df * (f > 0); dn2 = W1^T dpre; dx1 = dx2 + LN2_bwd(dn2)
    cols_dot(w.l2_w, E, v0, v1, scr, E, E);
    for (int i = threadIdx.x; i < E; i += blockDim.x) { v1[i] = sv[S.f + i] > 0.f ? v1[i] : 0.f; fc[Fq.dpre + i] = v1[i]; }
    __syncthreads();
    cols_dot(w.l1_w, E, v1, v2, scr, E, E);
    ln_bwd(sv + S.x1, sv[S.st2], sv[S.st2 + 1], w.n2w, v2, v2, fc + Fq.dg2, fc + Fq.db2, false, E, red);
    for (int i = threadIdx.x; i < E; i += blockDim.x) { v0[i] += v2[i]; fc[Fq.dx1 + i] = v0[i]; }             // v0 = dx1
    __syncthreads();
    // attention backward: do = Wo^T dx1
    cols_dot(w.out_w, E, v0, v1, scr, E, E);                                                                   // v1 = do
    for (int i = threadIdx.x; i < E; i += blockDim.x) fc[Fq.dov + i] = v1[i];
    // vo[h][k] = sum_d do[h,d] Wv[h*hd+d][k]   (so that dp[h,j] = vo_h . n_j + const_h)
    for (int idx = threadIdx.x; idx < heads * E; idx += blockDim.x) {
        const int h = idx / E, k = idx - h * E;
        const float* wr = w.in_w + (static_cast<int64_t>(2) * E + h * hd) * E + k;
        float a = 0.f;
#pragma unroll 8
        for (int d = 0; d < hd; ++d) a = fmaf(v1[h * hd + d], __ldg(wr + static_cast<int64_t>(d) * E), a);
        vo[idx] = a;
    }
    __syncthreads();
    for (int task = warp; task < heads * L; task += nw) {
        const int h = task / L, j = task - h * L;
        const float* nr = N1 + static_cast<int64_t>(j) * E;
        float a = 0.f;
        for (int k = lane; k < E; k += 32) a = fmaf(vo[h * E + k], nr[k], a);
        a = wsum(a);
        if (lane == 0) ds[task] = a;   // dp (up to a per-head constant that the softmax backward cancels)
    }
    __syncthreads();
    for (int h = warp; h < heads; h += nw) {   // ds = p * (dp - sum_j p dp)
        float s = 0.f;
        for (int j = lane; j < L; j += 32) s = fmaf(p[h * L + j], ds[h * L + j], s);
        s = wsum(s);
        float t = 0.f;
        for (int j = lane; j < L; j += 32) { const float v = p[h * L + j] * (ds[h * L + j] - s); ds[h * L + j] = v; t += v; }
        t = wsum(t);
        if (lane == 0) fc[Fq.sds + h] = t;
    }
    __syncthreads();
    // g[h][k] = sum_j ds[h,j] n_j[k];  dq[n] = (Wk[n,:] . g[head(n)] + bk[n] * sum_j ds) * scale
    for (int idx = threadIdx.x; idx < heads * E; idx += blockDim.x) {
        const int h = idx / E, k = idx - h * E;
        float a = 0.f;
        for (int j = 0; j < L; ++j) a = fmaf(ds[h * L + j], N1[static_cast<int64_t>(j) * E + k], a);
        fc[Fq.g + idx] = a;
    }
    __syncthreads();
    const float scale = rsqrtf(static_cast<float>(hd));
    for (int n = warp; n < E; n += nw) {
        const float* wr = w.in_w + (static_cast<int64_t>(E) + n) * E;
        const float* gr = fc + Fq.g + static_cast<int64_t>(n / hd) * E;
        float a = 0.f;
        for (int k = lane; k < E; k += 32) a = fmaf(__ldg(wr + k), gr[k], a);
        a = wsum(a);
        if (lane == 0) { const float dq = (a + w.in_b[E + n] * fc[Fq.sds + n / hd]) * scale; v2[n] = dq; fc[Fq.dq + n] = dq; }
    }
    __syncthreads();
    // dn1[0] gets Wq^T dq (token 0 is the only query)
    cols_dot(w.in_w, E, v2, v1, scr, E, E);                                                                    // v1 = Wq^T dq
    // LN1 backward token by token: dn1[j] = sum_h ds[h,j] qk_h + p[h,j] vo_h (+ v1 for j = 0); parameter gradients summed over j
    for (int i = threadIdx.x; i < E; i += blockDim.x) { fc[Fq.dg1 + i] = 0.f; fc[Fq.db1 + i] = 0.f; }
    __syncthreads();
    for (int j = 0; j < L; ++j) {
        for (int k = threadIdx.x; k < E; k += blockDim.x) {
            float a = j == 0 ? v1[k] : 0.f;
            for (int h = 0; h < heads; ++h) a += ds[h * L + j] * sv[S.qk + static_cast<int64_t>(h) * E + k] + p[h * L + j] * vo[h * E + k];
            v2[k] = a;
        }
        __syncthreads();
        const float* xr = j == 0 ? w.cls_token : enc + (static_cast<int64_t>(b) * D + j - 1) * E;
        ln_bwd(xr, sv[S.st1 + 2 * j], sv[S.st1 + 2 * j + 1], w.n1w, v2, v2, fc + Fq.dg1, fc + Fq.db1, true, E, red);
        if (j == 0) {
            for (int i = threadIdx.x; i < E; i += blockDim.x) fc[Fq.dcls + i] = v2[i] + v0[i];   // + the residual path (dx1)
        } else if (denc) {
            for (int i = threadIdx.x; i < E; i += blockDim.x) denc[(static_cast<int64_t>(b) * D + j - 1) * E + i] = v2[i];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// backward, kernel B: gradients = fixed-order sums over the batch of the per-volume rank-1 terms.
//   blockIdx.x in [0, 6E): one row of in_proj (3E rows), out_proj, linear1, linear2;  then C head rows;  then one "vector" CTA
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) slice_train_wgrad_kernel(const float* __restrict__ saved_all, const float* __restrict__ fact_all,
                                                                 SliceTrainGrads g, int B, int D, int E, int heads, int C) {
    const int L = D + 1, hd = E / heads;
    const SavedLayout S(L, E, heads);
    const FactorLayout Fq(E, heads);
    const int64_t sstride = slice_train_saved_floats(L, E, heads), fstride = slice_train_factor_floats(E, heads, C);
    const int r = blockIdx.x;
    if (r < 6 * E + C) {
        // row `n` of matrix `m`: dW[n, k] = sum_b u_b[n] * v_b[k]
        int m, n;
        if (r < 3 * E) { m = r / E; n = r - m * E; }            // 0 q rows, 1 k rows, 2 v rows of in_proj_weight
        else if (r < 6 * E) { m = r / E; n = r - m * E; }        // 3 out_proj, 4 linear1, 5 linear2
        else { m = 6; n = r - 6 * E; }                           // head
        float* out;
        float bias_acc = 0.f;
        switch (m) {
            case 0: case 1: case 2: out = g.in_w + (static_cast<int64_t>(m) * E + n) * E; break;
            case 3: out = g.out_w + static_cast<int64_t>(n) * E; break;
            case 4: out = g.l1_w + static_cast<int64_t>(n) * E; break;
            case 5: out = g.l2_w + static_cast<int64_t>(n) * E; break;
            default: out = g.head_w + static_cast<int64_t>(n) * E; break;
        }
        for (int k = threadIdx.x; k < E; k += blockDim.x) {
            float a = 0.f;
            for (int b = 0; b < B; ++b) {
                const float* sv = saved_all + b * sstride;
                const float* fc = fact_all + b * fstride;
                float u, v;
                switch (m) {
                    case 0: u = fc[Fq.dq + n]; v = sv[S.n1 + k]; break;                                         // dq (x) n_0
                    case 1: u = sv[S.q + n]; v = fc[Fq.g + static_cast<int64_t>(n / hd) * E + k]; break;          // q_h (x) g_h
                    case 2: u = fc[Fq.dov + n]; v = sv[S.hbar + static_cast<int64_t>(n / hd) * E + k]; break;     // do_h (x) hbar_h
                    case 3: u = fc[Fq.dx1 + n]; v = sv[S.o + k]; break;
                    case 4: u = fc[Fq.dpre + n]; v = sv[S.n2 + k]; break;
                    case 5: u = fc[Fq.dx2 + n]; v = sv[S.f + k]; break;
                    default: u = fc[Fq.dl + n]; v = sv[S.y + k]; break;
                }
                a = fmaf(u, v, a);
            }
            out[k] = a;
        }
        if (threadIdx.x == 0) {   // the bias entry of this row
            for (int b = 0; b < B; ++b) {
                const float* sv = saved_all + b * sstride;
                const float* fc = fact_all + b * fstride;
                switch (m) {
                    case 0: bias_acc += fc[Fq.dq + n]; break;
                    case 1: bias_acc += sv[S.q + n] * fc[Fq.sds + n / hd]; break;   // sum_j ds = 0 up to rounding: as autograd has it
                    case 2: bias_acc += fc[Fq.dov + n]; break;                       // sum_j p = 1
                    case 3: bias_acc += fc[Fq.dx1 + n]; break;
                    case 4: bias_acc += fc[Fq.dpre + n]; break;
                    case 5: bias_acc += fc[Fq.dx2 + n]; break;
                    default: bias_acc += fc[Fq.dl + n]; break;
                }
            }
            switch (m) {
                case 0: case 1: case 2: g.in_b[m * E + n] = bias_acc; break;
                case 3: g.out_b[n] = bias_acc; break;
                case 4: g.l1_b[n] = bias_acc; break;
                case 5: g.l2_b[n] = bias_acc; break;
                default: g.head_b[n] = bias_acc; break;
            }
        }
        return;
    }
    // vector parameters: LayerNorm affine pairs and the slice-CLS token
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
        float a[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int b = 0; b < B; ++b) {
            const float* fc = fact_all + b * fstride;
            a[0] += fc[Fq.dg1 + i]; a[1] += fc[Fq.db1 + i]; a[2] += fc[Fq.dg2 + i]; a[3] += fc[Fq.db2 + i];
            a[4] += fc[Fq.dgf + i]; a[5] += fc[Fq.dbf + i]; a[6] += fc[Fq.dcls + i];
        }
        g.n1w[i] = a[0]; g.n1b[i] = a[1]; g.n2w[i] = a[2]; g.n2b[i] = a[3]; g.nfw[i] = a[4]; g.nfb[i] = a[5]; g.cls_token[i] = a[6];
    }
}

// ---------------------------------------------------------------------------------------------------
// AdamW (torch.optim.AdamW semantics, base_model.py:103-110 with optimizer = AdamW, dino.py:41): decoupled weight decay, bias
// correction, eps added after the sqrt.  One flat parameter / gradient / moment buffer, 16-byte accesses.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                     float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                     float wd, float bc1, float bc2_sqrt, float gscale) {
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float* pa = reinterpret_cast<float*>(&pp); float* ma = reinterpret_cast<float*>(&mm); float* va = reinterpret_cast<float*>(&vv);
        const float* ga = reinterpret_cast<const float*>(&gg);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = ga[j] * gscale;
            pa[j] *= 1.0f - lr * wd;
            ma[j] = b1 * ma[j] + (1.0f - b1) * gr;
            va[j] = b2 * va[j] + (1.0f - b2) * gr * gr;
            const float denom = sqrtf(va[j]) / bc2_sqrt + eps;
            pa[j] -= (lr / bc1) * (ma[j] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = n4 << 2; i < n; ++i) {
            const float gr = g[i] * gscale;
            p[i] *= 1.0f - lr * wd;
            m[i] = b1 * m[i] + (1.0f - b1) * gr;
            v[i] = b2 * v[i] + (1.0f - b2) * gr * gr;
            p[i] -= (lr / bc1) * (m[i] / (sqrtf(v[i]) / bc2_sqrt + eps));
        }
    }
}

size_t slice_train_saved_bytes(int B, int D, int E, int heads) { return static_cast<size_t>(B) * slice_train_saved_floats(D + 1, E, heads) * 4; }
size_t slice_train_factor_bytes(int B, int E, int heads, int C) { return static_cast<size_t>(B) * slice_train_factor_floats(E, heads, C) * 4; }

static SliceTrainParams as_params(const float* const* p) {
    return SliceTrainParams{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[12], p[13], p[14], p[15], p[16]};
}

int launch_slice_train_forward(const float* enc, const uint8_t* pad_mask, const float* const* params, float* saved, float* logits,
                               int B, int D, int E, int heads, int C, cudaStream_t stream) {
    MST_REQUIRE(heads >= 1 && heads <= 32 && E % heads == 0 && (E / heads) % 8 == 0 && E <= 768, "slice training: E=%d heads=%d unsupported", E, heads);
    const int L = D + 1;
    const size_t smem = (static_cast<size_t>(4) * E + 32 + 40 + static_cast<size_t>(heads) * L + 8) * sizeof(float);
    MST_REQUIRE(smem <= 200 * 1024, "slice training: D=%d too large", D);
    MST_SET_DYN_SMEM(slice_train_forward_kernel, 200 * 1024);
    slice_train_forward_kernel<<<B, 768, smem, stream>>>(enc, pad_mask, as_params(params), saved, logits, D, E, heads, C);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_slice_train_backward(const float* enc, const float* dlogits, const float* const* params, const float* saved, float* factors,
                                float* const* grads, float* denc, int B, int D, int E, int heads, int C, cudaStream_t stream) {
    MST_REQUIRE(heads >= 1 && heads <= 32 && E % heads == 0 && (E / heads) % 8 == 0 && E <= 768, "slice training: E=%d heads=%d unsupported", E, heads);
    const int L = D + 1;
    const size_t smem = (static_cast<size_t>(5) * E + 40 + ((heads * L + 3) & ~3) + static_cast<size_t>(heads) * E + 8) * sizeof(float);
    MST_REQUIRE(smem <= 200 * 1024, "slice training: D=%d too large", D);
    MST_SET_DYN_SMEM(slice_train_backward_kernel, 200 * 1024);
    slice_train_backward_kernel<<<B, 768, smem, stream>>>(enc, dlogits, as_params(params), saved, factors, denc, D, E, heads, C);
    MST_CHECK_CUDA(cudaGetLastError());
    SliceTrainGrads g{grads[0], grads[1], grads[2], grads[3], grads[4], grads[5], grads[6], grads[7], grads[8], grads[9], grads[10], grads[11],
                      grads[12], grads[13], grads[14], grads[15], grads[16]};
    slice_train_wgrad_kernel<<<6 * E + C + 1, 128, 0, stream>>>(saved, factors, g, B, D, E, heads, C);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float wd, int step,
                 float grad_scale, int num_sms, cudaStream_t stream) {
    MST_REQUIRE(step >= 1 && n >= 1, "adamw: bad step / size");
    MST_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
                "adamw: buffers must be 16-byte aligned");
    const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(b1), step));          // torch evaluates these in double
    const float bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(b2), step)));
    int grid = static_cast<int>((n / 4 + 255) / 256);
    grid = grid < 1 ? 1 : (grid > 8 * num_sms ? 8 * num_sms : grid);
    adamw_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2_sqrt, grad_scale);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mst
