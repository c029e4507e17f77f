// C ABI (include/mst_b200.h): handle, weight store + packing, forward orchestration, saliency.
#include <stdarg.h>
#include <algorithm>
#include <map>
#include <type_traits>
#include <string>
#include <vector>

#include "../../include/mst_b200.h"
#include <stdlib.h>
#include "common.cuh"

namespace mst {

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

constexpr int KP = 256;  // im2col K (196) padded to a multiple of the 64-wide TMA box

// ---------------------------------------------------------------------------------------------------
// weight packing kernels
// ---------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T cvt(float v);
template <> __device__ __forceinline__ float cvt<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 cvt<bf16>(float v) { return __float2bfloat16_rn(v); }

// dst[n,k] = W[n,k] * (rowscale ? rowscale[n] : 1) * (n < nscaled ? s : 1);  bias likewise
template <typename T>
__global__ void pack_linear_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ rowscale,
                                   int nscaled, float s, T* __restrict__ Wd, float* __restrict__ bd, int N, int K) {
    const int64_t total = static_cast<int64_t>(N) * K;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int n = static_cast<int>(i / K);
        float f = (rowscale ? rowscale[n] : 1.0f) * (n < nscaled ? s : 1.0f);
        Wd[i] = cvt<T>(W[i] * f);
        if (i % K == 0) bd[n] = b[n] * f;
    }
}
// LayerNorm folded into the consuming Linear (EPI_LN_*): one CTA per output row n.
//   g[k] = W[n,k] * f_n * gamma[k];  Wd[n,k] = bf16(g[k] - mean_k g)  (centred row: x . Wd[n,:] = (x - mean x) . g);
//   bd[n] = b[n]*f_n + sum_k beta[k] * W[n,k] * f_n
// The centring must survive the bf16 rounding: x . Wd[n,:] carries mean(x) * sum_k Wd[n,k], and independent roundings leave
// sum_k Wd[n,k] ~ sqrt(K) * ulp/sqrt(12) (7e-4 for K = 384, |W| ~ 0.02), which a row with mean(x) * rstd ~ 20 (real DINOv2
// residual streams have such rows) turns into a 1.4e-2 error on every output.  So after rounding, single elements are moved
// by ONE bf16 step each -- greedily the move that brings the row sum closest to zero -- until no move helps: |sum_k Wd[n,k]|
// ends below half the smallest step available in the row (~1e-7), at the cost of about ten elements per row carrying up to
// 1.5 ulp of rounding error instead of half an ulp.
constexpr int kPackLnMaxK = 1024;
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(static_cast<uint32_t>(b) << 16); }
__global__ void __launch_bounds__(128) pack_linear_ln_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             int nscaled, float s, bf16* __restrict__ Wd, float* __restrict__ bd,
                                                             int K) {
    __shared__ float red[2][4];
    __shared__ float tk[kPackLnMaxK];       // exact centred values
    __shared__ uint16_t qb[kPackLnMaxK];    // their bf16 roundings (bits)
    const int n = blockIdx.x;
    const float f = n < nscaled ? s : 1.0f;
    float cs = 0.f, ds = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float w = W[static_cast<int64_t>(n) * K + k] * f;
        cs = fmaf(w, gamma[k], cs);
        ds = fmaf(beta[k], w, ds);
    }
    for (int o = 16; o > 0; o >>= 1) { cs += __shfl_xor_sync(0xffffffffu, cs, o); ds += __shfl_xor_sync(0xffffffffu, ds, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = cs; red[1][threadIdx.x >> 5] = ds; }
    __syncthreads();
    const float mean = ((red[0][0] + red[0][1]) + (red[0][2] + red[0][3])) / K;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float t = W[static_cast<int64_t>(n) * K + k] * f * gamma[k] - mean;
        tk[k] = t;
        qb[k] = __bfloat16_as_ushort(__float2bfloat16_rn(t));
    }
    __syncthreads();
    // row sum of the rounded values (exact in double), then the greedy single-step moves, searched by the whole CTA
    __shared__ double rsum[4];
    __shared__ double cand_abs[4];
    __shared__ int cand_k[4];
    __shared__ double r_now;
    {
        double part = 0.0;
        for (int k = threadIdx.x; k < K; k += blockDim.x) part += static_cast<double>(bf16_bits_to_float(qb[k]));
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((threadIdx.x & 31) == 0) rsum[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) r_now = (rsum[0] + rsum[1]) + (rsum[2] + rsum[3]);
        __syncthreads();
    }
    for (int iter = 0; iter < 64; ++iter) {
        const double r = r_now;
        // the single one-step move (of an element not moved yet) that brings the row sum closest to zero
        double best_abs = fabs(r);
        int best = -1;
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const uint16_t bits = qb[k];
            if ((bits & 0x7f80) == 0 || tk[k] == 3.0e38f) continue;     // zero / subnormal, or already moved
            const float qv = bf16_bits_to_float(bits);
            const bool up = r < 0.0;                                   // the row sum must grow
            const uint16_t nb = (up == (qv > 0.f)) ? bits + 1 : bits - 1;
            if ((nb & 0x7f80) == 0x7f80 || (nb & 0x7f80) == 0) continue;
            const double d = static_cast<double>(bf16_bits_to_float(nb)) - static_cast<double>(qv);
            if (fabs(r + d) < best_abs) { best_abs = fabs(r + d); best = k; }
        }
        for (int o = 16; o > 0; o >>= 1) {   // (ties: the smaller index, so the result does not depend on the reduction shape)
            const double oa = __shfl_xor_sync(0xffffffffu, best_abs, o);
            const int ok = __shfl_xor_sync(0xffffffffu, best, o);
            if (ok >= 0 && (oa < best_abs || (oa == best_abs && (best < 0 || ok < best)))) { best_abs = oa; best = ok; }
        }
        if ((threadIdx.x & 31) == 0) { cand_abs[threadIdx.x >> 5] = best_abs; cand_k[threadIdx.x >> 5] = best; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double ba = fabs(r);
            int bk = -1;
            for (int w = 0; w < 4; ++w)
                if (cand_k[w] >= 0 && (cand_abs[w] < ba || (cand_abs[w] == ba && bk >= 0 && cand_k[w] < bk))) { ba = cand_abs[w]; bk = cand_k[w]; }
            if (bk >= 0) {
                const uint16_t bits = qb[bk];
                const float qv = bf16_bits_to_float(bits);
                const uint16_t nb = ((r < 0.0) == (qv > 0.f)) ? bits + 1 : bits - 1;
                r_now = r + (static_cast<double>(bf16_bits_to_float(nb)) - static_cast<double>(qv));
                qb[bk] = nb;
                tk[bk] = 3.0e38f;   // marks the element as moved
            }
            cand_k[0] = bk;
        }
        __syncthreads();
        if (cand_k[0] < 0) break;
        __syncthreads();
    }
    if (threadIdx.x == 0) bd[n] = b[n] * f + ((red[1][0] + red[1][1]) + (red[1][2] + red[1][3]));
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) Wd[static_cast<int64_t>(n) * K + k] = __ushort_as_bfloat16(qb[k]);
}
// conv weight [E,3,14,14] -> [E,KP]: sum over the 3 identical input channels (dino.py:127 repeats gray -> RGB)
template <typename T>
__global__ void pack_patch_kernel(const float* __restrict__ W, T* __restrict__ Wd, int E) {
    const int total = E * KP;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / KP, c = i % KP;
        float v = 0.f;
        if (c < 196) v = (W[(n * 3 + 0) * 196 + c] + W[(n * 3 + 1) * 196 + c]) + W[(n * 3 + 2) * 196 + c];
        Wd[i] = cvt<T>(v);
    }
}
// posb[p,n] = pos[1+p,n] + conv_bias[n];  cls_pos0[n] = cls[n] + pos[0,n]   (vision_transformer.py:219-220)
__global__ void pack_pos_kernel(const float* __restrict__ pos, const float* __restrict__ cls, const float* __restrict__ cbias,
                                float* __restrict__ posb, float* __restrict__ cls_pos0, int P, int E) {
    const int total = (P + 1) * E;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int t = i / E, n = i % E;
        if (t == 0) cls_pos0[n] = cls[n] + pos[n];
        else posb[(t - 1) * E + n] = pos[i] + (cbias ? cbias[n] : 0.f);
    }
}
__global__ void transpose_kernel(const float* __restrict__ W, float* __restrict__ Wt, int N, int K) {  // W[N,K] -> Wt[K,N]
    const int64_t total = static_cast<int64_t>(N) * K;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i / N), n = static_cast<int>(i % N);
        Wt[i] = W[static_cast<int64_t>(n) * K + k];
    }
}

// ---------------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------------
struct Layer {
    void *wqkv = nullptr, *wproj = nullptr, *wfc1 = nullptr, *wfc2 = nullptr;
    float *bqkv = nullptr, *bproj = nullptr, *bfc1 = nullptr, *bfc2 = nullptr;
    bool fold_qkv = false, fold_fc1 = false;  // bf16 path: norm1 / norm2 folded into the qkv / fc1 weights (EPI_LN_*)
    const float *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr;
};

}  // namespace mst

namespace mst {
enum Cat { CAT_IM2COL = 0, CAT_GEMM_PATCH, CAT_LAYERNORM, CAT_GEMM_QKV, CAT_ATTENTION, CAT_GEMM_PROJ, CAT_GEMM_FC1,
           CAT_GEMM_FC2, CAT_CLS_ATTENTION, CAT_GEMM_CLS_ROWS, CAT_SLICE_FUSION, CAT_FULL_MAPS, CAT_SALIENCY_COMBINE,
           CAT_SALIENCY_UPSAMPLE, CAT_PREPARE_VOLUME, CAT_TRAIN_FORWARD, CAT_TRAIN_BACKWARD, CAT_ADAMW,
           CAT_BWD_ATTENTION, CAT_BWD_WGRAD, CAT_BWD_DGRAD, CAT_BWD_TRANSPOSE, CAT_BWD_POINTWISE, CAT_WEIGHT_PACK, NUM_CAT };
static const char* kCatNames = "im2col,gemm_patch,layernorm,gemm_qkv,attention,gemm_proj,gemm_fc1,gemm_fc2,cls_attention,gemm_cls_rows,slice_fusion,full_maps,saliency_combine,saliency_upsample,prepare_volume,train_slice_forward,train_slice_backward,adamw,bwd_attention,bwd_wgrad_gemm,bwd_dgrad_gemm,bwd_transpose,bwd_layernorm_gelu,weight_pack";
struct Profiler {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Rec { int cat; size_t e0, e1; };
    std::vector<Rec> recs;
    cudaEvent_t next() {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[used++];
    }
};
}  // namespace mst

namespace mst {
// One captured forward: every kernel of mst_forward for one exact argument tuple (pointers included -- they are baked into the
// kernel parameters and TMA descriptors), replayed with a single cudaGraphLaunch.  What it removes is host time: ~150
// cuTensorMapEncodeTiled calls and ~77 launches per forward, which is what a one-volume forward costs (main_predict.py:208).
struct GraphKey {
    const void* src; int src_dtype, B, D, H, W; const void* mask; int tta;
    void *logits, *feat, *enc, *plane, *slc, *full, *ws; size_t ws_bytes; void* stream;
    bool operator==(const GraphKey& o) const {
        return src == o.src && src_dtype == o.src_dtype && B == o.B && D == o.D && H == o.H && W == o.W && mask == o.mask && tta == o.tta &&
               logits == o.logits && feat == o.feat && enc == o.enc && plane == o.plane && slc == o.slc && full == o.full && ws == o.ws &&
               ws_bytes == o.ws_bytes && stream == o.stream;
    }
};
struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec = nullptr;   // null until the key has been seen twice (one-off argument tuples are not captured)
    unsigned long long launches = 0;
    unsigned long long last_use = 0;
};
constexpr int kMaxGraphs = 16;
}  // namespace mst

struct mst_handle_s {
    mst::Profiler prof;
    unsigned long long launches = 0;
    std::vector<mst::GraphEntry> graphs;
    unsigned long long graph_clock = 0, graph_replays = 0;
    long long graph_max_tokens = 0;             // 0 = CUDA graphs off (mst_set_graph_threshold)
    cudaStream_t cap_stream = nullptr;          // capture happens here (the caller's stream may be the legacy default stream, which
                                                // cannot be captured); the instantiated graph is launched into the caller's stream
    mst_config cfg;
    int num_sms = 0;
    bool finalized = false;
    std::map<std::string, float*> master;       // canonical name -> device fp32 (owned)
    std::map<std::string, int64_t> expected;    // canonical name -> numel
    std::map<std::string, bool> have;
    std::vector<void*> owned;                   // packed buffers, in allocation order
    std::vector<size_t> owned_bytes;            // ... and their sizes: a re-finalize (every optimizer step when training) reuses them
    size_t owned_cursor = 0;
    std::map<std::string, float*> grad_ptr;     // canonical name -> caller-owned fp32 gradient buffer (mst_set_grad)
    std::vector<void*> wT;                      // per layer x {qkv, proj, fc1, fc2}: bf16 W^T, the weight operand of the dgrad GEMMs
    float* zero_bias = nullptr;                 // 4 * embed_dim zeros (the dgrad GEMMs carry no bias)
    std::vector<mst::Layer> layers;
    void* wpatch = nullptr;
    float *posb = nullptr, *cls_pos0 = nullptr;
    std::map<int, float*> pos_cache;            // (gh << 16 | gw) -> interpolated patch position table (+ conv bias), owned
    mst::SliceWeights sw{};
    int slice_emb() const { return cfg.use_bottleneck ? cfg.embed_dim / 4 : cfg.embed_dim; }
    int feat_dim(int D) const { return cfg.slice_fusion == mst::SLICE_FUSION_LINEAR ? slice_emb() * D : slice_emb(); }
};

namespace mst {

static size_t elem_size(const mst_config& c) { return c.precision == MST_PRECISION_BF16 ? 2 : 4; }

static void expected_names(mst_handle h) {
    const int E = h->cfg.embed_dim, C = h->cfg.out_ch, Es = h->slice_emb();
    auto& x = h->expected;
    if (h->cfg.depth > 0) {
    x["encoder.cls_token"] = E;
    if (h->cfg.num_registers > 0) x["encoder.register_tokens"] = static_cast<int64_t>(h->cfg.num_registers) * E;
    x["encoder.pos_embed"] = static_cast<int64_t>(h->cfg.pos_tokens) * E;
    x["encoder.patch_embed.proj.weight"] = static_cast<int64_t>(E) * 3 * 196;
    x["encoder.patch_embed.proj.bias"] = E;
    for (int i = 0; i < h->cfg.depth; ++i) {
        const std::string p = "encoder.blocks." + std::to_string(i) + ".";
        x[p + "norm1.weight"] = E; x[p + "norm1.bias"] = E;
        x[p + "attn.qkv.weight"] = 3LL * E * E; x[p + "attn.qkv.bias"] = 3 * E;
        x[p + "attn.proj.weight"] = 1LL * E * E; x[p + "attn.proj.bias"] = E;
        x[p + "norm2.weight"] = E; x[p + "norm2.bias"] = E;
        x[p + "mlp.fc1.weight"] = 4LL * E * E; x[p + "mlp.fc1.bias"] = 4 * E;
        x[p + "mlp.fc2.weight"] = 4LL * E * E; x[p + "mlp.fc2.bias"] = E;
    }
    x["encoder.norm.weight"] = E; x["encoder.norm.bias"] = E;
    }   // depth == 0: a slice-transformer head on its own (MST-ResNet, resnet.py:127-198): no encoder tensors
    if (h->cfg.use_bottleneck) { x["bottleneck.weight"] = 1LL * Es * E; x["bottleneck.bias"] = Es; }  // dino.py:75-77
    if (h->cfg.slice_fusion == SLICE_FUSION_TRANSFORMER) {                                             // dino.py:80-97
        if (h->cfg.use_slice_pos_emb) x["slice_pos_emb.weight"] = 256LL * Es;
        x["cls_token"] = Es;
        const std::string q = "slice_fusion.layers.0.";
        x[q + "self_attn.in_proj_weight"] = 3LL * Es * Es; x[q + "self_attn.in_proj_bias"] = 3 * Es;
        x[q + "self_attn.out_proj.weight"] = 1LL * Es * Es; x[q + "self_attn.out_proj.bias"] = Es;
        x[q + "linear1.weight"] = 1LL * Es * Es; x[q + "linear1.bias"] = Es;
        x[q + "linear2.weight"] = 1LL * Es * Es; x[q + "linear2.bias"] = Es;
        x[q + "norm1.weight"] = Es; x[q + "norm1.bias"] = Es;
        x[q + "norm2.weight"] = Es; x[q + "norm2.bias"] = Es;
        x["slice_fusion.norm.weight"] = Es; x["slice_fusion.norm.bias"] = Es;
        if (h->cfg.rotary == MST_ROTARY_ROPE)   // RotaryEmbedding(dim = head_dim).freqs (rotary_embedding_torch.py:104,117)
            x[q + "self_attn.rotary_positional_encoding.freqs"] = Es / h->cfg.slice_heads / 2;
        if (h->cfg.rotary == MST_ROTARY_LIRE) {  // AttentionLiereRotator.vars (rotary_embedding_torch.py:342-344): accepted for the
            const int blk = Es / h->cfg.slice_heads / 2;   // state_dict round trip; the rotation they define cancels in q . k
            for (int i = 0; i < 2; ++i)
                x[q + "self_attn.rotary_positional_encoding.vars." + std::to_string(i)] = static_cast<int64_t>((blk * blk - blk) / 2) * 33;
        }
    }
    if (h->cfg.enable_linear) {                                                                        // dino.py:98-103
        const int64_t in = h->cfg.slice_fusion == SLICE_FUSION_LINEAR ? 32LL * Es : Es;
        x["linear.weight"] = C * in; x["linear.bias"] = C;
    }
}

// "encoder.blocks.0.7.x" (BlockChunk naming, vision_transformer.py:153-160) -> "encoder.blocks.7.x"
static std::string canonical_name(const std::string& name) {
    const std::string pre = "encoder.blocks.";
    if (name.compare(0, pre.size(), pre) != 0) return name;
    size_t p = pre.size(), e = name.find('.', p);
    if (e == std::string::npos) return name;
    const size_t e2 = name.find('.', e + 1);
    if (e2 != std::string::npos) {
        bool digits = e2 > e + 1;
        for (size_t i = e + 1; i < e2; ++i) digits = digits && isdigit(static_cast<unsigned char>(name[i]));
        if (digits) return pre + name.substr(e + 1);
    }
    return name;
}

template <typename T>
static int alloc_dev(mst_handle h, T** p, size_t count) {
    const size_t bytes = count * sizeof(T) ? count * sizeof(T) : 16;
    if (h->owned_cursor < h->owned.size() && h->owned_bytes[h->owned_cursor] == bytes) {   // re-finalize: same layout, same buffer
        *p = static_cast<T*>(h->owned[h->owned_cursor++]);
        return 0;
    }
    void* q = nullptr;
    MST_CHECK_CUDA(cudaMalloc(&q, bytes));
    if (h->owned_cursor < h->owned.size()) {
        cudaFree(h->owned[h->owned_cursor]);
        h->owned[h->owned_cursor] = q;
        h->owned_bytes[h->owned_cursor] = bytes;
    } else {
        h->owned.push_back(q);
        h->owned_bytes.push_back(bytes);
    }
    h->owned_cursor++;
    *p = static_cast<T*>(q);
    return 0;
}

template <typename T>
static int pack_linear(mst_handle h, const std::string& wname, const std::string& bname, const float* rowscale, int nscaled,
                       float s, int N, int K, void** Wd, float** bd, cudaStream_t st) {
    T* w;
    MST_PROPAGATE(alloc_dev<T>(h, &w, static_cast<size_t>(N) * K));
    MST_PROPAGATE(alloc_dev<float>(h, bd, N));
    pack_linear_kernel<T><<<1024, 256, 0, st>>>(h->master[wname], h->master[bname], rowscale, nscaled, s, w, *bd, N, K);
    MST_CHECK_CUDA(cudaGetLastError());
    *Wd = w;
    return 0;
}

// bf16 path: Linear with the preceding LayerNorm (gamma, beta) folded in (EPI_LN_*)
static int pack_linear_ln(mst_handle h, const std::string& wname, const std::string& bname, const float* gamma, const float* beta,
                          int nscaled, float s, int N, int K, void** Wd, float** bd, cudaStream_t st) {
    bf16* w;
    MST_REQUIRE(K <= kPackLnMaxK, "pack_linear_ln: K=%d exceeds %d", K, kPackLnMaxK);
    MST_PROPAGATE(alloc_dev<bf16>(h, &w, static_cast<size_t>(N) * K));
    MST_PROPAGATE(alloc_dev<float>(h, bd, N));
    pack_linear_ln_kernel<<<N, 128, 0, st>>>(h->master[wname], h->master[bname], gamma, beta, nscaled, s, w, *bd, K);
    MST_CHECK_CUDA(cudaGetLastError());
    *Wd = w;
    return 0;
}

static int transposed(mst_handle h, const std::string& name, int N, int K, const float** out, cudaStream_t st) {
    float* t;
    MST_PROPAGATE(alloc_dev<float>(h, &t, static_cast<size_t>(N) * K));
    transpose_kernel<<<512, 256, 0, st>>>(h->master[name], t, N, K);
    MST_CHECK_CUDA(cudaGetLastError());
    *out = t;
    return 0;
}

template <typename T>
static int finalize_t(mst_handle h, cudaStream_t st) {
    const int E = h->cfg.embed_dim, P = h->cfg.depth > 0 ? h->cfg.pos_tokens - 1 : 0;
    for (auto& kv : h->pos_cache) cudaFree(kv.second);
    h->pos_cache.clear();
    h->layers.assign(h->cfg.depth, Layer());
    for (int i = 0; i < h->cfg.depth; ++i) {
        const std::string p = "encoder.blocks." + std::to_string(i) + ".";
        Layer& L = h->layers[i];
        const float* g1 = h->have.count(p + "ls1.gamma") ? h->master[p + "ls1.gamma"] : nullptr;
        const float* g2 = h->have.count(p + "ls2.gamma") ? h->master[p + "ls2.gamma"] : nullptr;
        L.n1w = h->master[p + "norm1.weight"]; L.n1b = h->master[p + "norm1.bias"];
        L.n2w = h->master[p + "norm2.weight"]; L.n2b = h->master[p + "norm2.bias"];
        // q rows (first E outputs) carry the 1/sqrt(64) attention scale (attention.py:60): exact power of two.
        // bf16 path: norm1 is folded into qkv and norm2 into fc1 (the last block's fc1 runs on CLS rows behind a real
        // LayerNorm kernel, so it stays unfolded).
        static const bool no_fold = exp_env("MST_NO_LN_FOLD", 0) != 0;  // experiments only: separate LayerNorm kernels
        const bool kFold = std::is_same<T, bf16>::value && !no_fold;
        L.fold_qkv = kFold;
        L.fold_fc1 = kFold && i != h->cfg.depth - 1;
        if (L.fold_qkv)
            MST_PROPAGATE(pack_linear_ln(h, p + "attn.qkv.weight", p + "attn.qkv.bias", L.n1w, L.n1b, E, 0.125f, 3 * E, E, &L.wqkv,
                                         &L.bqkv, st));
        else
            MST_PROPAGATE(pack_linear<T>(h, p + "attn.qkv.weight", p + "attn.qkv.bias", nullptr, E, 0.125f, 3 * E, E, &L.wqkv, &L.bqkv, st));
        MST_PROPAGATE(pack_linear<T>(h, p + "attn.proj.weight", p + "attn.proj.bias", g1, 0, 1.f, E, E, &L.wproj, &L.bproj, st));
        if (L.fold_fc1)
            MST_PROPAGATE(pack_linear_ln(h, p + "mlp.fc1.weight", p + "mlp.fc1.bias", L.n2w, L.n2b, 0, 1.f, 4 * E, E, &L.wfc1, &L.bfc1, st));
        else
            MST_PROPAGATE(pack_linear<T>(h, p + "mlp.fc1.weight", p + "mlp.fc1.bias", nullptr, 0, 1.f, 4 * E, E, &L.wfc1, &L.bfc1, st));
        MST_PROPAGATE(pack_linear<T>(h, p + "mlp.fc2.weight", p + "mlp.fc2.bias", g2, 0, 1.f, E, 4 * E, &L.wfc2, &L.bfc2, st));
    }
    if (h->cfg.depth > 0) {
        T* wp;
        MST_PROPAGATE(alloc_dev<T>(h, &wp, static_cast<size_t>(E) * KP));
        pack_patch_kernel<T><<<256, 256, 0, st>>>(h->master["encoder.patch_embed.proj.weight"], wp, E);
        MST_CHECK_CUDA(cudaGetLastError());
        h->wpatch = wp;
        MST_PROPAGATE(alloc_dev<float>(h, &h->posb, static_cast<size_t>(P) * E));
        MST_PROPAGATE(alloc_dev<float>(h, &h->cls_pos0, E));
        pack_pos_kernel<<<256, 256, 0, st>>>(h->master["encoder.pos_embed"], h->master["encoder.cls_token"],
                                             h->master["encoder.patch_embed.proj.bias"], h->posb, h->cls_pos0, P, E);
        MST_CHECK_CUDA(cudaGetLastError());
    }
    const std::string q = "slice_fusion.layers.0.";
    SliceWeights& s = h->sw;
    s = SliceWeights{};
    const int Es = h->slice_emb();
    if (h->cfg.use_bottleneck) {
        MST_PROPAGATE(transposed(h, "bottleneck.weight", Es, E, &s.bott_wt, st));
        s.bott_b = h->master["bottleneck.bias"];
    }
    if (h->cfg.slice_fusion == SLICE_FUSION_TRANSFORMER) {
        if (h->cfg.use_slice_pos_emb) s.pos_emb = h->master["slice_pos_emb.weight"];
        s.cls_token = h->master["cls_token"];
        s.n1w = h->master[q + "norm1.weight"]; s.n1b = h->master[q + "norm1.bias"];
        s.n2w = h->master[q + "norm2.weight"]; s.n2b = h->master[q + "norm2.bias"];
        s.nfw = h->master["slice_fusion.norm.weight"]; s.nfb = h->master["slice_fusion.norm.bias"];
        s.in_w = h->master[q + "self_attn.in_proj_weight"];
        if (h->cfg.rotary == MST_ROTARY_ROPE) s.rope_freqs = h->master[q + "self_attn.rotary_positional_encoding.freqs"];
        s.liere = h->cfg.rotary == MST_ROTARY_LIRE ? 1 : 0;
        s.in_b = h->master[q + "self_attn.in_proj_bias"]; s.out_b = h->master[q + "self_attn.out_proj.bias"];
        s.l1_b = h->master[q + "linear1.bias"]; s.l2_b = h->master[q + "linear2.bias"];
        MST_PROPAGATE(transposed(h, q + "self_attn.in_proj_weight", 3 * Es, Es, &s.in_wt, st));
        MST_PROPAGATE(transposed(h, q + "self_attn.out_proj.weight", Es, Es, &s.out_wt, st));
        MST_PROPAGATE(transposed(h, q + "linear1.weight", Es, Es, &s.l1_wt, st));
        MST_PROPAGATE(transposed(h, q + "linear2.weight", Es, Es, &s.l2_wt, st));
    }
    if (h->cfg.enable_linear) {
        s.head_b = h->master["linear.bias"];
        const int in = h->cfg.slice_fusion == SLICE_FUSION_LINEAR ? 32 * Es : Es;
        MST_PROPAGATE(transposed(h, "linear.weight", h->cfg.out_ch, in, &s.head_wt, st));
    }
    return 0;
}

// interpolate_pos_encoding's resampling step (vision_transformer.py:194-208) for the handle's interpolate_* configuration
static int resample_pos(mst_handle h, const float* cbias, float* dst, int M, int gh, int gw, cudaStream_t st) {
    const int E = h->cfg.embed_dim;
    const float* pos = h->master["encoder.pos_embed"];
    if (h->cfg.interpolate_antialias) {
        MST_REQUIRE(h->cfg.interpolate_offset == 0.0f, "interpolate_antialias is built for interpolate_offset = 0 (the hub _reg configuration)");
        return launch_pos_bicubic_aa(pos, cbias, dst, M, gh, gw, E, st);
    }
    // scale_factor = (g + offset) / M (:199-200); ATen maps coordinates with 1/scale_factor.  offset == 0: size=(gh, gw) and the
    // coordinate scale is in/out (:203-204)
    const double off = h->cfg.interpolate_offset;
    const float sy = static_cast<float>(1.0 / ((gh + off) / static_cast<double>(M)));
    const float sx = static_cast<float>(1.0 / ((gw + off) / static_cast<double>(M)));
    return launch_pos_bicubic(pos, cbias, dst, M, gh, gw, E, sy, sx, st);
}

// Patch position table (+ conv bias) for a gh x gw patch grid: the checkpoint's own rows when the grid matches
// (vision_transformer.py:183-184), else bicubic resampling (:185-211), cached per grid for the handle's lifetime.
static int pos_for_grid(mst_handle h, int gh, int gw, const float** out, cudaStream_t st) {
    const int E = h->cfg.embed_dim, Pn = h->cfg.pos_tokens - 1;
    int M = 1;
    while ((M + 1) * (M + 1) <= Pn) ++M;
    if (gh == M && gw == M && M * M == Pn) { *out = h->posb; return 0; }
    MST_REQUIRE(M * M == Pn, "pos_embed has %d patch rows, not a square grid: cannot interpolate (vision_transformer.py:193)", Pn);
    const int key = (gh << 16) | gw;
    auto it = h->pos_cache.find(key);
    if (it == h->pos_cache.end()) {
        float* t = nullptr;
        MST_CHECK_CUDA(cudaMalloc(&t, static_cast<size_t>(gh) * gw * E * sizeof(float)));
        MST_PROPAGATE(resample_pos(h, h->master["encoder.patch_embed.proj.bias"], t, M, gh, gw, st));
        it = h->pos_cache.emplace(key, t).first;
    }
    *out = it->second;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
struct Workspace {
    void *A0, *x, *xn, *qkv, *hid, *ao_cls, *xc, *xcn, *hc;
    float *enc_cls, *hs;
    float* rowstat;
    float* rowpart;   // [M][4][2] row partials written by the proj epilogue, read by fc1 (bf16 ViT-S path)
    size_t total;
};
static Workspace carve(const mst_config& c, int B, int D, int H, int W, uint8_t* base) {
    const size_t es = elem_size(c);
    const int64_t E = c.embed_dim, BD = static_cast<int64_t>(B) * D, P = static_cast<int64_t>(H / 14) * (W / 14), N = P + 1 + c.num_registers, M = BD * N;
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base ? base + off : nullptr; off += (bytes + 255) & ~static_cast<size_t>(255); return p; };
    Workspace w;
    w.A0 = take(BD * P * KP * es);
    w.x = take(M * E * es);
    w.xn = take(M * E * es);
    w.qkv = take(M * 3 * E * es);
    w.hid = take(M * 4 * E * es);
    w.ao_cls = take(BD * E * es);
    w.xc = take(BD * E * es);
    w.xcn = take(BD * E * es);
    w.hc = take(BD * 4 * E * es);
    w.enc_cls = static_cast<float*>(take(BD * E * 4));
    w.hs = static_cast<float*>(take(static_cast<size_t>(B) * (D + 1) * E * 4));
    w.rowstat = static_cast<float*>(take(M * sizeof(float)));
    w.rowpart = static_cast<float*>(take(M * 8 * sizeof(float)));
    w.total = off;
    return w;
}

template <typename T> struct Ops;
template <> struct Ops<bf16> {
    static int gemm(mst_handle h, const void* A, int64_t lda, const void* W, int M, int N, int K, int mode, const EpiParams& ep, cudaStream_t st) {
        MST_REQUIRE(lda == K, "bf16 gemm expects a dense A (lda == K)");
        return gemm_bf16_tc(static_cast<const bf16*>(A), static_cast<const bf16*>(W), M, N, K, mode, ep, h->num_sms, st);
    }
    static int attention(mst_handle h, const void* qkv, void* out, int BD, int N, int heads, cudaStream_t st) {
        if (N == 257)   // ViT @224: the specialised tcgen05 kernel (16 softmax warps in two teams)
            return launch_attention_tc257x16(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), BD, heads, h->num_sms, st);
        static const int use_tcg = exp_env("MST_ATTN_TCG", 1);   // 0: A-B comparisons
        if (use_tcg && attention_tcg_supported(N))
            return launch_attention_tcg(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), BD, N, heads, h->num_sms, st);
        return launch_attention_bf16(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), BD, N, heads, st);
    }
};
template <> struct Ops<float> {
    static int gemm(mst_handle, const void* A, int64_t lda, const void* W, int M, int N, int K, int mode, const EpiParams& ep, cudaStream_t st) {
        return gemm_f32_simt(static_cast<const float*>(A), lda, static_cast<const float*>(W), M, N, K, mode, ep, st);
    }
    static int attention(mst_handle, const void* qkv, void* out, int BD, int N, int heads, cudaStream_t st) {
        return launch_attention_f32(static_cast<const float*>(qkv), static_cast<float*>(out), BD, N, heads, st);
    }
};

// one kernel launch, counted, and bracketed by events when profiling is on
#define MST_LAUNCH(cat, call)                                                             \
    do {                                                                                  \
        size_t _e0 = 0;                                                                   \
        if (h->prof.on) { _e0 = h->prof.used; cudaEventRecord(h->prof.next(), st); }      \
        MST_PROPAGATE(call);                                                              \
        h->launches++;                                                                    \
        if (h->prof.on) { cudaEventRecord(h->prof.next(), st); h->prof.recs.push_back({cat, _e0, _e0 + 1}); } \
    } while (0)

// qkv = LN1(x) W^T + b with the LayerNorm folded into the GEMM (rowstat = rstd of every row), inference and training forward.
// ViT-S: 1152 features = 4.5 pairs of 128.  The first 1024 run on the weights-in-TMEM kernel (whole cta_group::2 pairs, sixteen
// epilogue warps), the last 128 (the tail of V) on the streaming 128-wide tile kernel: two launches, the second one HBM-bound (it
// re-reads the rows), together faster than one launch of the weights-in-smem kernel.  (Small batches keep the single launch: below
// ~150 m-tiles both halves are latency, not throughput.)
template <typename T>
static int qkv_ln_folded(mst_handle h, const void* x, const Layer& L, const float* rowstat, void* qkv, int M, int E, cudaStream_t st) {
    EpiParams ep{};
    ep.bias = L.bqkv; ep.rowstat = rowstat; ep.out = qkv; ep.ldo = 3 * E;
    static const int qkv_split = exp_env("MST_QKV_SPLIT", 1);
    if (qkv_split && sizeof(T) == 2 && E == 384 && gemm_wt_enabled() && M >= 32768) {
        MST_LAUNCH(CAT_GEMM_QKV, Ops<T>::gemm(h, x, E, L.wqkv, M, 1024, E, EPI_LN_BIAS, ep, st));
        EpiParams ep2 = ep;
        ep2.bias = L.bqkv + 1024;
        ep2.out = static_cast<T*>(qkv) + 1024;
        MST_LAUNCH(CAT_GEMM_QKV, Ops<T>::gemm(h, x, E, static_cast<const T*>(L.wqkv) + static_cast<size_t>(1024) * E, M, 128, E,
                                              EPI_LN_BIAS, ep2, st));
    } else {
        MST_LAUNCH(CAT_GEMM_QKV, Ops<T>::gemm(h, x, E, L.wqkv, M, 3 * E, E, EPI_LN_BIAS, ep, st));
    }
    return 0;
}

template <typename T>
static int forward_t(mst_handle h, const void* src, int src_dtype, int B, int D, int H, int W, const uint8_t* pad_mask, int tta,
                     float* logits, float* feat, float* enc_cls_out, float* plane_cls, float* slice_cls, float* full_maps,
                     const Workspace& ws, cudaStream_t st) {
    // B counts the volumes that run through the encoder: with tta, 8 flipped variants of each of the B / 8 source volumes
    const mst_config& c = h->cfg;
    const int E = c.embed_dim, BD = B * D, P = (H / 14) * (W / 14), R = c.num_registers, N = P + 1 + R;
    const float* posb = nullptr;
    MST_PROPAGATE(pos_for_grid(h, H / 14, W / 14, &posb, st));
    const int64_t M64 = static_cast<int64_t>(BD) * N;
    MST_REQUIRE(M64 * 4 * E < (1LL << 40) && M64 < (1LL << 31) - 256, "batch too large: %lld tokens", (long long)M64);
    const int M = static_cast<int>(M64);
    T* x = static_cast<T*>(ws.x);
    T* xn = static_cast<T*>(ws.xn);

    // patch embedding + CLS/pos (K1-K3)
    MST_LAUNCH(CAT_IM2COL, launch_im2col<T>(src, src_dtype, static_cast<T*>(ws.A0), x, h->cls_pos0,
                                            R > 0 ? h->master["encoder.register_tokens"] : nullptr, R, BD, H, W, KP, E,
                                            tta ? BD / 8 : 0, D, st));
    {
        EpiParams ep{};
        ep.posb = posb; ep.P = P; ep.R = R; ep.out = x; ep.ldo = E;
        MST_LAUNCH(CAT_GEMM_PATCH, Ops<T>::gemm(h, ws.A0, KP, h->wpatch, BD * P, E, KP, EPI_PATCH, ep, st));
    }
    bool stats_from_fc2 = false;
    for (int l = 0; l < c.depth; ++l) {
        const Layer& L = h->layers[l];
        const bool last = (l == c.depth - 1);
        // x = x + ls1(attn(norm1(x)))                                   (block.py:112)
        if (L.fold_qkv) {  // bf16: only the row statistics are materialised; norm1 itself rides in the qkv epilogue
            if (!stats_from_fc2)   // (blocks >= 1: the previous block's fc2 epilogue has already written them)
                MST_LAUNCH(CAT_LAYERNORM, launch_row_stats(reinterpret_cast<const bf16*>(x), ws.rowstat, M, E, 1e-6f, st));
            stats_from_fc2 = false;
            MST_PROPAGATE((qkv_ln_folded<T>(h, x, L, ws.rowstat, ws.qkv, M, E, st)));
        } else {
            MST_LAUNCH(CAT_LAYERNORM, (launch_layernorm<T, T>(x, E, xn, E, L.n1w, L.n1b, M, E, 1e-6f, st)));
            EpiParams ep{};
            ep.bias = L.bqkv; ep.out = ws.qkv; ep.ldo = 3 * E;
            MST_LAUNCH(CAT_GEMM_QKV, Ops<T>::gemm(h, xn, E, L.wqkv, M, 3 * E, E, EPI_BIAS, ep, st));
        }
        if (full_maps)  // what the reference's hook appends for every block (dino.py:241): [BD, heads, N, N] fp32
            MST_LAUNCH(CAT_FULL_MAPS, launch_attention_probs<T>(static_cast<const T*>(ws.qkv),
                                                                full_maps + static_cast<int64_t>(l) * BD * c.enc_heads * N * N, BD, N,
                                                                c.enc_heads, st));
        if (!last) {
            MST_LAUNCH(CAT_ATTENTION, Ops<T>::attention(h, ws.qkv, xn, BD, N, c.enc_heads, st));  // xn is dead: reuse as attention output
            // norm2 statistics as row partials out of proj's epilogue, consumed by fc1: measured SLOWER (proj 0.21 -> 0.40 ms with
            // the explicit per-lane residual loads instead of the TMA reduce-add, fc1 0.68 -> 0.78 ms with 32 bytes of partials
            // per token instead of 4 bytes of rstd: step 31.3 -> 32.1 ms), so it is off unless MST_PROJ_STAT_FUSE=1
            static const int fuse_stats = exp_env("MST_PROJ_STAT_FUSE", 0);
            const bool part_fc1 = fuse_stats && sizeof(T) == 2 && E == 384 && L.fold_fc1 && gemm_wt_enabled();
            {
                EpiParams ep{};
                ep.bias = L.bproj; ep.res = x; ep.ldr = E; ep.out = x; ep.ldo = E;
                if (part_fc1) ep.rowpart_out = ws.rowpart;
                MST_LAUNCH(CAT_GEMM_PROJ, Ops<T>::gemm(h, xn, E, L.wproj, M, E, E, EPI_BIAS_RES, ep, st));
            }
            // x = x + ls2(mlp(norm2(x)))                                  (block.py:113)
            if (L.fold_fc1) {
                if (!part_fc1)
                    MST_LAUNCH(CAT_LAYERNORM, launch_row_stats(reinterpret_cast<const bf16*>(x), ws.rowstat, M, E, 1e-6f, st));
                EpiParams ep{};
                ep.bias = L.bfc1; ep.rowstat = ws.rowstat; ep.out = ws.hid; ep.ldo = 4 * E;
                if (part_fc1) { ep.rowpart = ws.rowpart; ep.stat_eps = 1e-6f; }
                MST_LAUNCH(CAT_GEMM_FC1, Ops<T>::gemm(h, x, E, L.wfc1, M, 4 * E, E, EPI_LN_BIAS_GELU, ep, st));
            } else {
                MST_LAUNCH(CAT_LAYERNORM, (launch_layernorm<T, T>(x, E, xn, E, L.n2w, L.n2b, M, E, 1e-6f, st)));
                EpiParams ep{};
                ep.bias = L.bfc1; ep.out = ws.hid; ep.ldo = 4 * E;
                MST_LAUNCH(CAT_GEMM_FC1, Ops<T>::gemm(h, xn, E, L.wfc1, M, 4 * E, E, EPI_BIAS_GELU, ep, st));
            }
            {
                EpiParams ep{};
                ep.bias = L.bfc2; ep.res = x; ep.ldr = E; ep.out = x; ep.ldo = E;
                // The next block's norm1 statistics come out of this GEMM's epilogue (it then adds the residual itself instead
                // of a TMA reduce-add, and one CTA pair walks both n-blocks of its rows): one pass over x less per block.
                static const int fuse = exp_env("MST_NO_STAT_FUSE", 0) ? 0 : 1;
                if (fuse && sizeof(T) == 2 && E == 384 && h->layers[l + 1].fold_qkv) {
                    ep.rowstat_out = ws.rowstat; ep.stat_eps = 1e-6f;
                    stats_from_fc2 = true;
                }
                MST_LAUNCH(CAT_GEMM_FC2, Ops<T>::gemm(h, ws.hid, 4 * E, L.wfc2, M, E, 4 * E, EPI_BIAS_RES, ep, st));
            }
        } else {
            // Last block: only token 0 of each slice is consumed downstream (vision_transformer.py:265,329), so
            // the query side, proj and the MLP run on the BD CLS rows only; K/V still come from every token.
            T* ao = static_cast<T*>(ws.ao_cls);
            T* xc = static_cast<T*>(ws.xc);
            T* xcn = static_cast<T*>(ws.xcn);
            MST_LAUNCH(CAT_CLS_ATTENTION, launch_cls_attention<T>(static_cast<const T*>(ws.qkv), ao, plane_cls, BD, N, c.enc_heads, st));
            {
                EpiParams ep{};
                ep.bias = L.bproj; ep.res = x; ep.ldr = static_cast<int64_t>(N) * E; ep.out = xc; ep.ldo = E;
                MST_LAUNCH(CAT_GEMM_CLS_ROWS, Ops<T>::gemm(h, ao, E, L.wproj, BD, E, E, EPI_BIAS_RES, ep, st));
            }
            MST_LAUNCH(CAT_LAYERNORM, (launch_layernorm<T, T>(xc, E, xcn, E, L.n2w, L.n2b, BD, E, 1e-6f, st)));
            {
                EpiParams ep{};
                ep.bias = L.bfc1; ep.out = ws.hc; ep.ldo = 4 * E;
                MST_LAUNCH(CAT_GEMM_CLS_ROWS, Ops<T>::gemm(h, xcn, E, L.wfc1, BD, 4 * E, E, EPI_BIAS_GELU, ep, st));
            }
            {
                EpiParams ep{};
                ep.bias = L.bfc2; ep.res = xc; ep.ldr = E; ep.out = xc; ep.ldo = E;
                MST_LAUNCH(CAT_GEMM_CLS_ROWS, Ops<T>::gemm(h, ws.hc, 4 * E, L.wfc2, BD, E, 4 * E, EPI_BIAS_RES, ep, st));
            }
            // final encoder LayerNorm on the CLS rows (vision_transformer.py:263-265), fp32 out
            float* enc = enc_cls_out ? enc_cls_out : ws.enc_cls;
            MST_LAUNCH(CAT_LAYERNORM, (launch_layernorm<T, float>(xc, E, enc, E, h->master["encoder.norm.weight"],
                                                                 h->master["encoder.norm.bias"], BD, E, 1e-6f, st)));
            MST_LAUNCH(CAT_SLICE_FUSION, launch_slice_fusion(enc, pad_mask, h->sw, ws.hs, c.enable_linear ? logits : nullptr, feat,
                                                         slice_cls, B, D, E, h->slice_emb(), c.slice_heads, c.out_ch,
                                                         c.slice_fusion, tta ? B / 8 : 0, st));
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// training step of the encoder (bf16): forward that keeps every block's activations, and the backward pass over them
// (kernels: train_enc.cu; contractions: the tcgen05 GEMM).  Reference: Lightning's step, base_model.py:148-170, on the
// default un-frozen construction (dino.py:56-103); dropout / drop-path are 0 there, so train == eval arithmetic.
// ---------------------------------------------------------------------------------------------------
struct TrainLayer { bf16 *x_in, *qkv, *ao, *x_mid, *u, *hid; float* lse; };
struct TrainWs {
    bf16* A0;                       // im2col of the input [BD*P, KP]
    std::vector<TrainLayer> L;
    bf16* x_out;                    // residual stream after the last block [M, E]
    float* rowstat;                 // [M]
    // backward scratch
    bf16 *dX, *dXm, *dln, *dao, *dqkv, *dhid, *ln, *Ta, *Tb;
    float* lnws;                    // LayerNorm-backward partials
    float* wg;                      // [KP, E] fp32: patch weight gradient, transposed
    int Mpad;
    size_t total;
};
static TrainWs carve_train(const mst_config& c, int B, int D, int H, int W, uint8_t* base) {
    const int64_t E = c.embed_dim, BD = static_cast<int64_t>(B) * D, P = static_cast<int64_t>(H / 14) * (W / 14), N = P + 1, M = BD * N;
    // (>= 448: the fp32-output GEMM that contracts over the padded token dimension runs on the streaming schedule, K > 384)
    const int64_t Mpad = std::max<int64_t>(448, (std::max(M, BD * P) + 63) / 64 * 64);
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base ? base + off : nullptr; off += (bytes + 255) & ~static_cast<size_t>(255); return p; };
    TrainWs w;
    w.Mpad = static_cast<int>(Mpad);
    w.A0 = static_cast<bf16*>(take(BD * P * KP * 2));
    w.L.resize(c.depth);
    for (int l = 0; l < c.depth; ++l) {
        w.L[l].x_in = static_cast<bf16*>(take(M * E * 2));
        w.L[l].qkv = static_cast<bf16*>(take(M * 3 * E * 2));
        w.L[l].ao = static_cast<bf16*>(take(M * E * 2));
        w.L[l].x_mid = static_cast<bf16*>(take(M * E * 2));
        w.L[l].u = static_cast<bf16*>(take(M * 4 * E * 2));
        w.L[l].hid = static_cast<bf16*>(take(M * 4 * E * 2));
        w.L[l].lse = static_cast<float*>(take(static_cast<size_t>(M) * c.enc_heads * 4));   // row log-sum-exp per (slice, head, token)
    }
    w.x_out = static_cast<bf16*>(take(M * E * 2));
    w.rowstat = static_cast<float*>(take(M * 4));
    w.dX = static_cast<bf16*>(take(M * E * 2));
    w.dXm = static_cast<bf16*>(take(M * E * 2));
    w.dln = static_cast<bf16*>(take(M * E * 2));
    w.dao = static_cast<bf16*>(take(M * E * 2));
    w.dqkv = static_cast<bf16*>(take(M * 3 * E * 2));
    w.dhid = static_cast<bf16*>(take(M * 4 * E * 2));
    w.ln = static_cast<bf16*>(take(M * E * 2));
    w.Ta = static_cast<bf16*>(take(4 * E * Mpad * 2));
    w.Tb = static_cast<bf16*>(take(4 * E * Mpad * 2));
    w.lnws = static_cast<float*>(take(ln_bwd_workspace_bytes(static_cast<int>(E))));
    w.wg = static_cast<float*>(take(static_cast<size_t>(KP) * E * 4));
    w.total = off;
    return w;
}

static int check_trainable(mst_handle h, int H, int W) {
    const mst_config& c = h->cfg;
    MST_REQUIRE(c.precision == MST_PRECISION_BF16, "encoder training runs in bf16 (construct with precision='bf16')");
    MST_REQUIRE(c.embed_dim == 384 || c.embed_dim == 768, "encoder training: embed_dim %d unsupported (384 / 768)", c.embed_dim);
    MST_REQUIRE(c.num_registers == 0, "encoder training: register tokens are not supported");
    int Mg = 1;
    while ((Mg + 1) * (Mg + 1) <= c.pos_tokens - 1) ++Mg;
    MST_REQUIRE(H / 14 == Mg && W / 14 == Mg && Mg * Mg == c.pos_tokens - 1,
                "encoder training needs the position table's own grid (%d x %d patches); resampled tables have no backward here", Mg, Mg);
    for (int i = 0; i < c.depth; ++i)
        MST_REQUIRE(!h->have.count("encoder.blocks." + std::to_string(i) + ".ls1.gamma"), "encoder training: LayerScale is not supported");
    return 0;
}

// W^T in bf16 for the four Linears of every block, from the fp32 master weights (refreshed by mst_train_forward: the weights
// change every optimizer step)
static int refresh_dgrad_weights(mst_handle h, cudaStream_t st) {
    const int E = h->cfg.embed_dim, depth = h->cfg.depth;
    if (h->wT.empty()) {
        const size_t sz[4] = {static_cast<size_t>(3) * E * E, static_cast<size_t>(E) * E, static_cast<size_t>(4) * E * E, static_cast<size_t>(4) * E * E};
        for (int l = 0; l < depth; ++l)
            for (int k = 0; k < 4; ++k) {
                void* q = nullptr;
                MST_CHECK_CUDA(cudaMalloc(&q, sz[k] * 2));
                h->wT.push_back(q);
            }
        MST_CHECK_CUDA(cudaMalloc(&h->zero_bias, static_cast<size_t>(4) * E * 4));
        MST_CHECK_CUDA(cudaMemsetAsync(h->zero_bias, 0, static_cast<size_t>(4) * E * 4, st));
    }
    for (int l = 0; l < depth; ++l) {
        const std::string p = "encoder.blocks." + std::to_string(l) + ".";
        MST_PROPAGATE(launch_transpose_f32_to_bf16(h->master[p + "attn.qkv.weight"], static_cast<bf16*>(h->wT[4 * l + 0]), 3 * E, E, st));
        MST_PROPAGATE(launch_transpose_f32_to_bf16(h->master[p + "attn.proj.weight"], static_cast<bf16*>(h->wT[4 * l + 1]), E, E, st));
        MST_PROPAGATE(launch_transpose_f32_to_bf16(h->master[p + "mlp.fc1.weight"], static_cast<bf16*>(h->wT[4 * l + 2]), 4 * E, E, st));
        MST_PROPAGATE(launch_transpose_f32_to_bf16(h->master[p + "mlp.fc2.weight"], static_cast<bf16*>(h->wT[4 * l + 3]), E, 4 * E, st));
    }
    h->launches += 4ull * depth;
    return 0;
}

static int train_forward(mst_handle h, const void* src, int src_dtype, int B, int D, int H, int W, float* enc_out, const TrainWs& ws,
                         cudaStream_t st) {
    const mst_config& c = h->cfg;
    const int E = c.embed_dim, BD = B * D, P = (H / 14) * (W / 14), N = P + 1, M = BD * N;
    const float* posb = nullptr;
    MST_PROPAGATE(pos_for_grid(h, H / 14, W / 14, &posb, st));
    MST_LAUNCH(CAT_IM2COL, launch_im2col<bf16>(src, src_dtype, ws.A0, ws.L[0].x_in, h->cls_pos0, nullptr, 0, BD, H, W, KP, E, 0, D, st));
    {
        EpiParams ep{};
        ep.posb = posb; ep.P = P; ep.R = 0; ep.out = ws.L[0].x_in; ep.ldo = E;
        MST_LAUNCH(CAT_GEMM_PATCH, Ops<bf16>::gemm(h, ws.A0, KP, h->wpatch, BD * P, E, KP, EPI_PATCH, ep, st));
    }
    for (int l = 0; l < c.depth; ++l) {
        const Layer& L = h->layers[l];
        const TrainLayer& T = ws.L[l];
        bf16* x_next = l + 1 < c.depth ? ws.L[l + 1].x_in : ws.x_out;
        MST_REQUIRE(L.fold_qkv, "encoder training needs the LayerNorm-folded weight packing");
        MST_LAUNCH(CAT_LAYERNORM, launch_row_stats(T.x_in, ws.rowstat, M, E, 1e-6f, st));
        MST_PROPAGATE((qkv_ln_folded<bf16>(h, T.x_in, L, ws.rowstat, T.qkv, M, E, st)));
        if (N == 257)   // the specialised kernel keeps the row log-sum-exp for the backward pass
            MST_LAUNCH(CAT_ATTENTION, launch_attention_tc257x16(T.qkv, T.ao, BD, c.enc_heads, h->num_sms, st, nullptr, T.lse));
        else
            MST_LAUNCH(CAT_ATTENTION, Ops<bf16>::attention(h, T.qkv, T.ao, BD, N, c.enc_heads, st));
        {
            EpiParams ep{};
            ep.bias = L.bproj; ep.res = T.x_in; ep.ldr = E; ep.out = T.x_mid; ep.ldo = E;
            MST_LAUNCH(CAT_GEMM_PROJ, Ops<bf16>::gemm(h, T.ao, E, L.wproj, M, E, E, EPI_BIAS_RES, ep, st));
        }
        MST_LAUNCH(CAT_LAYERNORM, launch_row_stats(T.x_mid, ws.rowstat, M, E, 1e-6f, st));
        if (L.fold_fc1) {
            EpiParams ep{};
            ep.bias = L.bfc1; ep.rowstat = ws.rowstat; ep.out = T.u; ep.ldo = 4 * E;
            MST_LAUNCH(CAT_GEMM_FC1, Ops<bf16>::gemm(h, T.x_mid, E, L.wfc1, M, 4 * E, E, EPI_LN_BIAS, ep, st));
        } else {   // the last block's fc1 is packed un-folded (inference runs it behind a real LayerNorm on the CLS rows)
            MST_LAUNCH(CAT_LAYERNORM, (launch_layernorm<bf16, bf16>(T.x_mid, E, ws.ln, E, L.n2w, L.n2b, M, E, 1e-6f, st)));
            EpiParams ep{};
            ep.bias = L.bfc1; ep.out = T.u; ep.ldo = 4 * E;
            MST_LAUNCH(CAT_GEMM_FC1, Ops<bf16>::gemm(h, ws.ln, E, L.wfc1, M, 4 * E, E, EPI_BIAS, ep, st));
        }
        MST_LAUNCH(CAT_GEMM_FC1, launch_gelu_fwd(T.u, T.hid, static_cast<int64_t>(M) * 4 * E, h->num_sms, st));
        {
            EpiParams ep{};
            ep.bias = L.bfc2; ep.res = T.x_mid; ep.ldr = E; ep.out = x_next; ep.ldo = E;
            MST_LAUNCH(CAT_GEMM_FC2, Ops<bf16>::gemm(h, T.hid, 4 * E, L.wfc2, M, E, 4 * E, EPI_BIAS_RES, ep, st));
        }
    }
    // final encoder LayerNorm on the CLS rows (vision_transformer.py:263-265,329), fp32 out
    MST_LAUNCH(CAT_LAYERNORM, (launch_layernorm<bf16, float>(ws.x_out, static_cast<int64_t>(N) * E, enc_out, E, h->master["encoder.norm.weight"],
                                                         h->master["encoder.norm.bias"], BD, E, 1e-6f, st)));
    return 0;
}

// ---- small gradient kernels of the token assembly (vision_transformer.py:219-220, patch_embed.py:75-77) ----
// dX rows of token t summed over the slices: dpos[t] (token 0 also = dcls); conv-bias gradient = sum over the patch tokens
__global__ void __launch_bounds__(768) token_grads_kernel(const bf16* __restrict__ dX, float* __restrict__ dpos, float* __restrict__ dcls,
                                                           int BD, int N, int E) {
    // one CTA per token: E / 8 lanes of eight features x (768 / lanes) slice groups, 16-byte loads, groups summed in a fixed order
    extern __shared__ float tg_part[];          // [groups][E]
    const int t = blockIdx.x, lanes = E / 8, groups = blockDim.x / lanes;
    const int l = threadIdx.x % lanes, g = threadIdx.x / lanes;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = g; s < BD; s += groups) {
        const uint4 v = *reinterpret_cast<const uint4*>(dX + (static_cast<int64_t>(s) * N + t) * E + 8 * l);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // bf16 -> fp32: the bits are the upper half of the float
            a[2 * j] += __uint_as_float(w[j] << 16);
            a[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) tg_part[g * E + 8 * l + j] = a[j];
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float r = 0.f;
        for (int k = 0; k < groups; ++k) r += tg_part[k * E + e];
        dpos[static_cast<int64_t>(t) * E + e] = r;
        if (t == 0 && dcls) dcls[e] = r;
    }
}
__global__ void __launch_bounds__(128) conv_bias_grad_kernel(const float* __restrict__ dpos, float* __restrict__ dbias, int N, int E) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float a = 0.f;
    for (int t = 1; t < N; ++t) a += dpos[static_cast<int64_t>(t) * E + e];
    dbias[e] = a;
}
// patch rows of dX (token rows 1.. of every slice) gathered contiguously: [BD*P, E]
__global__ void __launch_bounds__(256) gather_patch_rows_kernel(const bf16* __restrict__ dX, bf16* __restrict__ out, int64_t rows, int P, int E) {
    const int64_t n8 = rows * (E / 8);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / (E / 8);
        const int c = static_cast<int>(i - r * (E / 8));
        const int64_t srow = (r / P) * (P + 1) + 1 + r % P;
        reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(dX + srow * E)[c];
    }
}
// wg [KP, E] fp32 (= dWsum^T) -> conv weight gradient [E, 3, 14, 14]: the three input channels carry the same gray image
// (dino.py:127), so each gets the gradient of the channel-summed weight
__global__ void conv_weight_grad_kernel(const float* __restrict__ wg, float* __restrict__ dW, int E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * 3 * 196) return;
    const int n = i / (3 * 196), c = i % 196;
    dW[i] = wg[static_cast<int64_t>(c) * E + n];
}

static float* grad_of(mst_handle h, const std::string& name) {
    auto it = h->grad_ptr.find(name);
    return it == h->grad_ptr.end() ? nullptr : it->second;
}

// dW = dY^T X (fp32) and db = column sums of dY, for Y = X W^T + b with dY [M, Nout], X [M, Kin]
static int linear_wgrad(mst_handle h, const bf16* dY, int Nout, const bf16* X, int Kin, int M, const TrainWs& ws, float* dW, float* db,
                        cudaStream_t st) {
    if (wgrad_tc_supported(Nout, Kin)) {   // operands read as they lie (MN-major descriptors), bias gradient in the same kernel
        MST_LAUNCH(CAT_BWD_WGRAD, launch_wgrad_tc(dY, Nout, X, Kin, M, dW, db, h->num_sms, st));
        return 0;
    }
    if (db) MST_CHECK_CUDA(cudaMemsetAsync(db, 0, static_cast<size_t>(Nout) * 4, st));
    MST_LAUNCH(CAT_BWD_TRANSPOSE, launch_transpose_colsum(dY, Nout, ws.Ta, db, M, Nout, ws.Mpad, st));
    MST_LAUNCH(CAT_BWD_TRANSPOSE, launch_transpose_colsum(X, Kin, ws.Tb, nullptr, M, Kin, ws.Mpad, st));
    EpiParams ep{};
    ep.out = dW; ep.ldo = Kin;
    MST_LAUNCH(CAT_BWD_WGRAD, gemm_bf16_tc(ws.Ta, ws.Tb, Nout, Kin, ws.Mpad, EPI_RAW_F32, ep, h->num_sms, st));
    return 0;
}
// dX = dY W (bf16): the dgrad GEMM against W^T
static int linear_dgrad(mst_handle h, const bf16* dY, int Nout, const void* WT, int Kin, int M, bf16* dX, cudaStream_t st) {
    EpiParams ep{};
    ep.bias = h->zero_bias; ep.out = dX; ep.ldo = Kin;
    MST_LAUNCH(CAT_BWD_DGRAD, gemm_bf16_tc(dY, static_cast<const bf16*>(WT), M, Kin, Nout, EPI_BIAS, ep, h->num_sms, st));
    return 0;
}

static int train_backward(mst_handle h, const float* denc, int B, int D, int H, int W, const TrainWs& ws, cudaStream_t st) {
    const mst_config& c = h->cfg;
    const int E = c.embed_dim, BD = B * D, P = (H / 14) * (W / 14), N = P + 1, M = BD * N;
    const int64_t ME = static_cast<int64_t>(M) * E;
    float scratch_needed = 0.f; (void)scratch_needed;
    // every gradient buffer must have been registered
    auto need = [&](const std::string& n) -> float* { return grad_of(h, n); };
    // final norm on the CLS rows: dX = 0 except the CLS rows
    MST_CHECK_CUDA(cudaMemsetAsync(ws.dX, 0, ME * 2, st));
    {
        float* dg = need("encoder.norm.weight"); float* db = need("encoder.norm.bias");
        MST_REQUIRE(dg && db, "mst_train_backward: gradient buffers of encoder.norm.* were not set");
        // rows of dy / dx are the BD CLS rows; x rows are N*E apart; dx is written compactly into ws.dln and scattered below
        MST_LAUNCH(CAT_BWD_POINTWISE, launch_ln_bwd(ws.x_out, static_cast<int64_t>(N) * E, nullptr, denc, nullptr, h->master["encoder.norm.weight"],
                                                     ws.dln, dg, db, BD, E, 1e-6f, ws.lnws, st));
        MST_CHECK_CUDA(cudaMemcpy2DAsync(ws.dX, static_cast<size_t>(N) * E * 2, ws.dln, static_cast<size_t>(E) * 2, static_cast<size_t>(E) * 2, BD,
                                         cudaMemcpyDeviceToDevice, st));
    }
    for (int l = c.depth - 1; l >= 0; --l) {
        const std::string p = "encoder.blocks." + std::to_string(l) + ".";
        const TrainLayer& T = ws.L[l];
        float *gqw = need(p + "attn.qkv.weight"), *gqb = need(p + "attn.qkv.bias"), *gpw = need(p + "attn.proj.weight"), *gpb = need(p + "attn.proj.bias");
        float *g1w = need(p + "mlp.fc1.weight"), *g1b = need(p + "mlp.fc1.bias"), *g2w = need(p + "mlp.fc2.weight"), *g2b = need(p + "mlp.fc2.bias");
        float *n1w = need(p + "norm1.weight"), *n1b = need(p + "norm1.bias"), *n2w = need(p + "norm2.weight"), *n2b = need(p + "norm2.bias");
        MST_REQUIRE(gqw && gqb && gpw && gpb && g1w && g1b && g2w && g2b && n1w && n1b && n2w && n2b,
                    "mst_train_backward: gradient buffers of block %d were not set", l);
        // ---- x_out = x_mid + fc2(gelu(fc1(LN2(x_mid))))                                   (block.py:113, mlp.py:34-40) ----
        MST_PROPAGATE(linear_wgrad(h, ws.dX, E, T.hid, 4 * E, M, ws, g2w, g2b, st));
        MST_PROPAGATE(linear_dgrad(h, ws.dX, E, h->wT[4 * l + 3], 4 * E, M, ws.dhid, st));
        MST_LAUNCH(CAT_BWD_POINTWISE, launch_gelu_bwd(T.u, ws.dhid, ws.dhid, static_cast<int64_t>(M) * 4 * E, h->num_sms, st));   // du, in place
        MST_LAUNCH(CAT_BWD_POINTWISE, (launch_layernorm<bf16, bf16>(T.x_mid, E, ws.ln, E, h->master[p + "norm2.weight"], h->master[p + "norm2.bias"],
                                                                     M, E, 1e-6f, st)));
        MST_PROPAGATE(linear_wgrad(h, ws.dhid, 4 * E, ws.ln, E, M, ws, g1w, g1b, st));
        MST_PROPAGATE(linear_dgrad(h, ws.dhid, 4 * E, h->wT[4 * l + 2], E, M, ws.dln, st));
        MST_LAUNCH(CAT_BWD_POINTWISE, launch_ln_bwd(T.x_mid, E, ws.dln, nullptr, ws.dX, h->master[p + "norm2.weight"], ws.dXm, n2w, n2b, M, E, 1e-6f,
                                                     ws.lnws, st));
        // ---- x_mid = x_in + proj(attn(LN1(x_in)))                                         (block.py:112, attention.py:56-69) ----
        MST_PROPAGATE(linear_wgrad(h, ws.dXm, E, T.ao, E, M, ws, gpw, gpb, st));
        MST_PROPAGATE(linear_dgrad(h, ws.dXm, E, h->wT[4 * l + 1], E, M, ws.dao, st));
        MST_LAUNCH(CAT_BWD_ATTENTION, launch_attention_bwd(T.qkv, T.ao, ws.dao, ws.dqkv, BD, N, c.enc_heads, st, N == 257 ? T.lse : nullptr));
        MST_LAUNCH(CAT_BWD_POINTWISE, (launch_layernorm<bf16, bf16>(T.x_in, E, ws.ln, E, h->master[p + "norm1.weight"], h->master[p + "norm1.bias"],
                                                                     M, E, 1e-6f, st)));
        MST_PROPAGATE(linear_wgrad(h, ws.dqkv, 3 * E, ws.ln, E, M, ws, gqw, gqb, st));
        MST_PROPAGATE(linear_dgrad(h, ws.dqkv, 3 * E, h->wT[4 * l + 0], E, M, ws.dln, st));
        MST_LAUNCH(CAT_BWD_POINTWISE, launch_ln_bwd(T.x_in, E, ws.dln, nullptr, ws.dXm, h->master[p + "norm1.weight"], ws.dX, n1w, n1b, M, E, 1e-6f,
                                                     ws.lnws, st));
    }
    // ---- token assembly: x[s, 0] = cls + pos[0]; x[s, 1 + p] = conv(patch p) + bias + pos[1 + p] ----
    float *gpos = need("encoder.pos_embed"), *gcls = need("encoder.cls_token"), *gcw = need("encoder.patch_embed.proj.weight"),
          *gcb = need("encoder.patch_embed.proj.bias");
    MST_REQUIRE(gpos && gcls && gcw && gcb, "mst_train_backward: gradient buffers of the patch embedding / position table were not set");
    {
        const int lanes = E / 8, groups = 768 / lanes;     // E = 384 / 768 / 1024: 48 x 16, 96 x 8, 128 x 6
        token_grads_kernel<<<N, lanes * groups, static_cast<size_t>(groups) * E * 4, st>>>(ws.dX, gpos, gcls, BD, N, E);
    }
    MST_CHECK_CUDA(cudaGetLastError());
    conv_bias_grad_kernel<<<(E + 127) / 128, 128, 0, st>>>(gpos, gcb, N, E);
    MST_CHECK_CUDA(cudaGetLastError());
    {
        const int64_t rows = static_cast<int64_t>(BD) * P;
        gather_patch_rows_kernel<<<h->num_sms * 4, 256, 0, st>>>(ws.dX, ws.dXm, rows, P, E);
        MST_CHECK_CUDA(cudaGetLastError());
        // dWsum^T [KP, E] = A0^T dXpatch: A = A0^T [KP rows, rows], weight = dXpatch^T [E rows, rows]
        MST_LAUNCH(CAT_BWD_TRANSPOSE, launch_transpose_colsum(ws.A0, KP, ws.Ta, nullptr, static_cast<int>(rows), KP, ws.Mpad, st));
        MST_LAUNCH(CAT_BWD_TRANSPOSE, launch_transpose_colsum(ws.dXm, E, ws.Tb, nullptr, static_cast<int>(rows), E, ws.Mpad, st));
        EpiParams ep{};
        ep.out = ws.wg; ep.ldo = E;
        MST_LAUNCH(CAT_BWD_WGRAD, gemm_bf16_tc(ws.Ta, ws.Tb, KP, E, ws.Mpad, EPI_RAW_F32, ep, h->num_sms, st));
        conv_weight_grad_kernel<<<(E * 3 * 196 + 255) / 256, 256, 0, st>>>(ws.wg, gcw, E);
        MST_CHECK_CUDA(cudaGetLastError());
        h->launches += 4;
    }
    return 0;
}

}  // namespace mst

// ---------------------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------------------
using namespace mst;

extern "C" {

int mst_abi_version(void) { return MST_ABI_VERSION; }
const char* mst_last_error(void) { return g_err.c_str(); }

int mst_create(const mst_config* cfg, mst_handle* out) {
    MST_REQUIRE(cfg && out, "mst_create: null argument");
    if (cfg->depth == 0) {   // slice-transformer head only (features come from the caller's backbone: MST-ResNet, resnet.py:127-198)
        MST_REQUIRE(cfg->embed_dim >= 32 && cfg->embed_dim <= 1024 && cfg->embed_dim % 32 == 0,
                    "mst_create: a slice-transformer head takes 32 <= embed_dim <= 1024, a multiple of 32 (got %d)", cfg->embed_dim);
        MST_REQUIRE(cfg->slice_fusion == MST_FUSION_TRANSFORMER && !cfg->use_bottleneck && !cfg->use_slice_pos_emb && cfg->rotary == MST_ROTARY_NONE,
                    "mst_create: the head-only handle is the plain slice transformer (resnet.py:155-170)");
    } else {
        MST_REQUIRE(cfg->embed_dim == 384 || cfg->embed_dim == 768 || cfg->embed_dim == 1024,
                    "mst_create: embed_dim %d unsupported (384/768/1024)", cfg->embed_dim);
        MST_REQUIRE(cfg->enc_heads * 64 == cfg->embed_dim, "mst_create: enc_heads*64 must equal embed_dim");
        MST_REQUIRE(cfg->pos_tokens >= 2, "mst_create: bad pos_tokens");
    }
    MST_REQUIRE(cfg->depth >= 0 && cfg->out_ch >= 1, "mst_create: bad depth/out_ch");
    MST_REQUIRE(cfg->slice_heads >= 1 && cfg->slice_heads <= 16 && cfg->embed_dim % cfg->slice_heads == 0 &&
                    (cfg->embed_dim / (cfg->use_bottleneck ? 4 : 1) / cfg->slice_heads) % 8 == 0,
                "mst_create: bad slice_heads (head dimension must be a multiple of 8)");
    MST_REQUIRE(cfg->precision == MST_PRECISION_FP32 || cfg->precision == MST_PRECISION_BF16, "mst_create: bad precision");
    MST_REQUIRE(cfg->num_registers >= 0 && cfg->num_registers <= 16, "mst_create: bad num_registers %d", cfg->num_registers);
    MST_REQUIRE(cfg->slice_fusion >= MST_FUSION_TRANSFORMER && cfg->slice_fusion <= MST_FUSION_AVERAGE, "mst_create: bad slice_fusion %d",
                cfg->slice_fusion);
    MST_REQUIRE(cfg->rotary == MST_ROTARY_NONE ||
                    ((cfg->rotary == MST_ROTARY_ROPE || cfg->rotary == MST_ROTARY_LIRE) && cfg->slice_fusion == MST_FUSION_TRANSFORMER),
                "mst_create: rotary=%d unsupported (RoPE / LiRE need slice_fusion='transformer')", cfg->rotary);
    MST_REQUIRE(cfg->rotary != MST_ROTARY_LIRE || (cfg->embed_dim / (cfg->use_bottleneck ? 4 : 1) / cfg->slice_heads) % 2 == 0,
                "mst_create: LiRE needs an even slice head dimension");
    MST_REQUIRE(cfg->interpolate_offset >= 0.0f && cfg->interpolate_offset < 1.0f, "mst_create: bad interpolate_offset %f", (double)cfg->interpolate_offset);
    MST_REQUIRE(!cfg->use_bottleneck || (cfg->embed_dim / 4) % cfg->slice_heads == 0, "mst_create: bottleneck width %d not divisible by %d heads",
                cfg->embed_dim / 4, cfg->slice_heads);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    MST_REQUIRE(e == cudaSuccess && ndev > 0, "mst_create: no CUDA device (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    MST_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "mst_create: device %d out of range", cfg->device);
    MST_CHECK_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    MST_CHECK_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    MST_REQUIRE(prop.major == 10, "mst_create: device is sm_%d%d; this library is built for sm_100a (B200) only", prop.major, prop.minor);
    mst_handle h = new mst_handle_s();
    h->cfg = *cfg;
    h->num_sms = prop.multiProcessorCount;
    expected_names(h);
    *out = h;
    return 0;
}

static void drop_graphs(mst_handle h) {
    for (auto& g : h->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
}

int mst_destroy(mst_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    drop_graphs(h);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    for (auto& kv : h->master) cudaFree(kv.second);
    for (void* p : h->owned) cudaFree(p);
    for (void* p : h->wT) cudaFree(p);
    if (h->zero_bias) cudaFree(h->zero_bias);
    for (auto& kv : h->pos_cache) cudaFree(kv.second);
    for (cudaEvent_t e : h->prof.pool) cudaEventDestroy(e);
    delete h;
    return 0;
}

int mst_set_weight(mst_handle h, const char* name, const float* dev_fp32, int64_t numel, void* stream) {
    MST_REQUIRE(h && name && dev_fp32, "mst_set_weight: null argument");
    const std::string key = canonical_name(name);
    if (key == "encoder.mask_token") return 0;
    int64_t want = -1;
    auto it = h->expected.find(key);
    if (it != h->expected.end()) want = it->second;
    else if (key.size() > 10 && (key.rfind(".ls1.gamma") == key.size() - 10 || key.rfind(".ls2.gamma") == key.size() - 10)) want = h->cfg.embed_dim;
    MST_REQUIRE(want >= 0, "mst_set_weight: unexpected tensor '%s'", name);
    MST_REQUIRE(want == numel, "mst_set_weight: '%s' has %lld elements, expected %lld", name, (long long)numel, (long long)want);
    MST_CHECK_CUDA(cudaSetDevice(h->cfg.device));
    float*& dst = h->master[key];
    if (!dst) MST_CHECK_CUDA(cudaMalloc(&dst, numel * sizeof(float)));
    MST_CHECK_CUDA(cudaMemcpyAsync(dst, dev_fp32, numel * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    h->have[key] = true;
    h->finalized = false;
    return 0;
}

int mst_set_weights(mst_handle h, int32_t count, const char* const* names, const float* const* dev_fp32, const int64_t* numel, void* stream) {
    MST_REQUIRE(h && names && dev_fp32 && numel && count >= 0, "mst_set_weights: null argument");
    for (int i = 0; i < count; ++i) MST_PROPAGATE(mst_set_weight(h, names[i], dev_fp32[i], numel[i], stream));
    return 0;
}

int mst_finalize_weights(mst_handle h, void* stream) {
    MST_REQUIRE(h, "mst_finalize_weights: null handle");
    for (auto& kv : h->expected) MST_REQUIRE(h->have.count(kv.first), "mst_finalize_weights: tensor '%s' was never set", kv.first.c_str());
    MST_CHECK_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MST_CHECK_CUDA(cudaStreamSynchronize(st));
    drop_graphs(h);   // captured forwards were taken with the old weights' values in some kernels' constant parameters
    h->owned_cursor = 0;   // packed buffers are reused in allocation order (same architecture => same sizes)
    if (h->cfg.precision == MST_PRECISION_BF16) {
        MST_PROPAGATE(tma_init());
        MST_PROPAGATE(finalize_t<bf16>(h, st));
    } else {
        MST_PROPAGATE(finalize_t<float>(h, st));
    }
    MST_CHECK_CUDA(cudaStreamSynchronize(st));
    h->finalized = true;
    return 0;
}

static int check_shape(mst_handle h, int B, int D, int H, int W) {
    MST_REQUIRE(B >= 1 && D >= 1, "empty batch: B=%d D=%d", B, D);
    MST_REQUIRE(H > 0 && W > 0 && H % 14 == 0 && W % 14 == 0,
                "Input image height/width (%d, %d) is not a multiple of patch size 14", H, W);  // patch_embed.py:72-73
    MST_REQUIRE(h->cfg.slice_fusion != MST_FUSION_LINEAR || D == 32,
                "slice_fusion='linear' is built for 32 slices (dino.py:99), got D=%d", D);
    MST_REQUIRE(!h->cfg.use_slice_pos_emb || D <= 256, "slice position embedding holds 256 slices (dino.py:82), got D=%d", D);
    MST_REQUIRE(h->cfg.rotary != MST_ROTARY_LIRE || (B == 1 && D == 32),
                "rotary_positional_encoding='LiRE' runs for batch 1 and 32 slices only (rotary_embedding_torch.py:350, transformer_blocks.py:263 "
                "raise otherwise), got B=%d D=%d", B, D);
    return 0;
}

int mst_workspace_bytes(mst_handle h, int32_t B, int32_t D, int32_t H, int32_t W, size_t* bytes) {
    MST_REQUIRE(h && bytes, "mst_workspace_bytes: null argument");
    MST_PROPAGATE(check_shape(h, B, D, H, W));
    *bytes = carve(h->cfg, B, D, H, W, nullptr).total;
    return 0;
}

int mst_forward(mst_handle h, const void* src, int32_t src_dtype, int32_t B, int32_t D, int32_t H, int32_t W, const uint8_t* pad_mask,
                int32_t tta, float* logits, float* feat, float* enc_cls, float* plane_cls, float* slice_cls, float* full_maps,
                void* workspace, size_t workspace_bytes, void* stream) {
    MST_REQUIRE(h && src && workspace && (logits || !h->cfg.enable_linear), "mst_forward: null argument");
    MST_REQUIRE(src_dtype == MST_SRC_F32 || src_dtype == MST_SRC_BF16 || src_dtype == MST_SRC_F16, "mst_forward: bad src_dtype %d", src_dtype);
    MST_REQUIRE(src_dtype == MST_SRC_F32 || h->cfg.precision == MST_PRECISION_BF16,
                "mst_forward: the fp32 parity mode takes fp32 volumes (a 16-bit source would be the only rounding on the path)");
    MST_REQUIRE(!tta || full_maps == nullptr, "mst_forward: full_maps and tta are exclusive");
    if (tta) B *= 8;   // the encoder sees 8 flipped variants of every volume (variant-major); outputs are sized for them
    MST_REQUIRE(h->cfg.slice_fusion == MST_FUSION_TRANSFORMER || slice_cls == nullptr,
                "mst_forward: slice attention exists only for slice_fusion='transformer' (dino.py:257-260)");
    MST_REQUIRE(h->finalized, "mst_forward: weights not finalized");
    MST_PROPAGATE(check_shape(h, B, D, H, W));
    MST_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "mst_forward: workspace must be 256-byte aligned");
    Workspace ws = carve(h->cfg, B, D, H, W, static_cast<uint8_t*>(workspace));
    MST_REQUIRE(ws.total <= workspace_bytes, "mst_forward: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    MST_CHECK_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto run_on = [&](cudaStream_t s_) -> int {
        if (h->cfg.precision == MST_PRECISION_BF16)
            return forward_t<bf16>(h, src, src_dtype, B, D, H, W, pad_mask, tta, logits, feat, enc_cls, plane_cls, slice_cls, full_maps, ws, s_);
        return forward_t<float>(h, src, src_dtype, B, D, H, W, pad_mask, tta, logits, feat, enc_cls, plane_cls, slice_cls, full_maps, ws, s_);
    };
    auto run = [&]() -> int { return run_on(st); };
    set_pdl(h->graph_max_tokens > 0 && static_cast<long long>(B) * D * ((H / 14) * (W / 14) + 1 + h->cfg.num_registers) <= h->graph_max_tokens);
    // ---- CUDA graph replay for small batches (launch-bound: the kernels of a one-volume forward take less time than issuing them) ----
    const long long tokens = static_cast<long long>(B) * D * ((H / 14) * (W / 14) + 1 + h->cfg.num_registers);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (h->graph_max_tokens <= 0 || tokens > h->graph_max_tokens || h->prof.on || full_maps != nullptr ||
        cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)
        return run();
    const GraphKey key{src, src_dtype, B, D, H, W, pad_mask, tta, logits, feat, enc_cls, plane_cls, slice_cls, full_maps, workspace,
                       workspace_bytes, stream};
    GraphEntry* e = nullptr;
    for (auto& g : h->graphs)
        if (g.key == key) { e = &g; break; }
    if (e && e->exec) {
        MST_CHECK_CUDA(cudaGraphLaunch(e->exec, st));
        e->last_use = ++h->graph_clock;
        h->launches += e->launches;
        h->graph_replays++;
        return 0;
    }
    if (!e) {   // first sight of this argument tuple: run it eagerly, remember it (least recently used entry makes room)
        if (static_cast<int>(h->graphs.size()) >= kMaxGraphs) {
            size_t lru = 0;
            for (size_t i = 1; i < h->graphs.size(); ++i)
                if (h->graphs[i].last_use < h->graphs[lru].last_use) lru = i;
            if (h->graphs[lru].exec) cudaGraphExecDestroy(h->graphs[lru].exec);
            h->graphs.erase(h->graphs.begin() + lru);
        }
        GraphEntry ne;
        ne.key = key;
        ne.last_use = ++h->graph_clock;
        h->graphs.push_back(ne);
        return run();
    }
    // second sight: capture.  (The position table of this grid and every per-device kernel attribute were set up by the eager run.)
    const unsigned long long l0 = h->launches;
    if (!h->cap_stream) MST_CHECK_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    MST_CHECK_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeRelaxed));
    const int rc = run_on(h->cap_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &graph);
    if (rc != 0 || ce != cudaSuccess || graph == nullptr) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        h->launches = l0;
        h->graph_max_tokens = 0;   // capture is not possible in this context: stay on the eager path
        return run();
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess || exec == nullptr) {
        cudaGetLastError();
        h->launches = l0;
        h->graph_max_tokens = 0;
        return run();
    }
    e->exec = exec;
    e->launches = h->launches - l0;
    e->last_use = ++h->graph_clock;
    MST_CHECK_CUDA(cudaGraphLaunch(exec, st));
    h->graph_replays++;
    return 0;
}

int mst_set_graph_threshold(mst_handle h, int64_t max_tokens) {
    MST_REQUIRE(h, "mst_set_graph_threshold: null handle");
    h->graph_max_tokens = max_tokens;
    if (max_tokens <= 0) drop_graphs(h);
    return 0;
}
unsigned long long mst_graph_replays(mst_handle h) { return h ? h->graph_replays : 0; }

int mst_slice_head_forward(mst_handle h, const float* feats, int32_t B, int32_t D, const uint8_t* pad_mask, float* logits, float* feat,
                           float* slice_cls, float* scratch, void* stream) {
    MST_REQUIRE(h && feats && scratch && (logits || !h->cfg.enable_linear), "mst_slice_head_forward: null argument");
    MST_REQUIRE(h->finalized, "mst_slice_head_forward: weights not finalized");
    MST_REQUIRE(B >= 1 && D >= 1, "empty batch: B=%d D=%d", B, D);
    MST_REQUIRE(h->cfg.slice_fusion == MST_FUSION_TRANSFORMER && !h->cfg.use_bottleneck, "mst_slice_head_forward: transformer fusion without bottleneck");
    MST_CHECK_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const mst_config& c = h->cfg;
    MST_LAUNCH(CAT_SLICE_FUSION, launch_slice_fusion(feats, pad_mask, h->sw, scratch, c.enable_linear ? logits : nullptr, feat, slice_cls, B, D,
                                                 c.embed_dim, c.embed_dim, c.slice_heads, c.out_ch, c.slice_fusion, 0, st));
    return 0;
}

int mst_saliency(mst_handle h, const float* plane_cls, const float* slice_cls, int32_t B, int32_t D, int32_t enc_heads,
                 int32_t slice_heads, int32_t skip_tokens, int32_t gh, int32_t gw, int32_t H, int32_t W, int32_t tta, float* attn_maps,
                 float* plane_attn, float* slice_attn, float* coarse, float* full, void* stream) {
    MST_REQUIRE(slice_cls && (plane_cls || (!attn_maps && !plane_attn && !coarse && !full)), "mst_saliency: null argument");
    MST_REQUIRE(B >= 1 && D >= 1 && gh >= 1 && gw >= 1 && skip_tokens >= 1, "mst_saliency: empty input");
    MST_REQUIRE(coarse != nullptr || full == nullptr, "mst_saliency: the full-resolution map needs the coarse buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    mst_handle_s none;   // instrumentation is optional: without a handle the launches are not counted
    if (!h) h = &none;
    MST_LAUNCH(CAT_SALIENCY_COMBINE, launch_saliency_combine(plane_cls, slice_cls, B, D, enc_heads, slice_heads, skip_tokens, gh, gw, tta,
                                                             attn_maps, plane_attn, slice_attn, coarse, st));
    if (full) MST_LAUNCH(CAT_SALIENCY_UPSAMPLE, launch_saliency_upsample(coarse, full, B, D, gh, gw, H, W, st));
    return 0;
}

int mst_rollout(const float* maps, int32_t depth, int32_t nmat, int32_t N, float* out, float* scratch, void* stream) {
    MST_REQUIRE(maps && out && scratch && depth >= 1 && nmat >= 1 && N >= 1, "mst_rollout: bad argument");
    return launch_rollout(maps, depth, nmat, N, out, scratch, static_cast<cudaStream_t>(stream));
}

int mst_pos_embed(mst_handle h, int32_t H, int32_t W, float* out, void* stream) {
    MST_REQUIRE(h && out && h->finalized, "mst_pos_embed: null argument or weights not finalized");
    MST_REQUIRE(H > 0 && W > 0 && H % 14 == 0 && W % 14 == 0, "mst_pos_embed: %dx%d is not a multiple of 14", H, W);
    MST_CHECK_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int E = h->cfg.embed_dim, gh = H / 14, gw = W / 14;
    int M = 1;
    while ((M + 1) * (M + 1) <= h->cfg.pos_tokens - 1) ++M;
    const float* pos = h->master["encoder.pos_embed"];
    MST_CHECK_CUDA(cudaMemcpyAsync(out, pos, E * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (gh == M && gw == M && M * M == h->cfg.pos_tokens - 1) {
        MST_CHECK_CUDA(cudaMemcpyAsync(out + E, pos + E, static_cast<size_t>(M) * M * E * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    MST_REQUIRE(M * M == h->cfg.pos_tokens - 1, "mst_pos_embed: pos_embed is not a square grid");
    return resample_pos(h, nullptr, out + E, M, gh, gw, st);
}

int mst_quantile_workspace_bytes(int32_t items, int32_t nq, size_t* bytes) {
    MST_REQUIRE(bytes && items >= 1 && nq >= 1, "mst_quantile_workspace_bytes: bad argument");
    *bytes = quantile_workspace_bytes(items, nq);
    return 0;
}
int mst_quantile(const float* data, int64_t n, int32_t items, const double* q_dev, int32_t nq, double* out, void* workspace,
                 size_t workspace_bytes, void* stream) {
    MST_REQUIRE(data && q_dev && out && workspace, "mst_quantile: null argument");
    MST_REQUIRE(workspace_bytes >= quantile_workspace_bytes(items, nq), "mst_quantile: workspace too small");
    int dev = 0, sms = 0;
    MST_CHECK_CUDA(cudaGetDevice(&dev));
    MST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return launch_quantile(data, n, items, q_dev, nq, out, workspace, sms, static_cast<cudaStream_t>(stream));
}

int mst_prepare_volume_workspace_bytes(int32_t items, int32_t W0, int32_t H0, int32_t D0, size_t* bytes) {
    MST_REQUIRE(bytes && items >= 1 && W0 >= 1 && H0 >= 1 && D0 >= 1, "mst_prepare_volume_workspace_bytes: bad argument");
    *bytes = prepare_volume_workspace_bytes(items, W0, H0, D0) + prepare_volume_raw_bytes(items, W0, H0, D0);
    return 0;
}
int mst_prepare_volume(mst_handle h, const void* src_any, int32_t src_dtype, int32_t items, int32_t W0, int32_t H0, int32_t D0, int32_t W, int32_t H, int32_t D,
                       int32_t flip_h, float q_lo, float q_hi, float* out, double* stats, void* workspace,
                       size_t workspace_bytes, void* stream) {
    MST_REQUIRE(src_any && out && workspace, "mst_prepare_volume: null argument");
    MST_REQUIRE(src_dtype == MST_SRC_F32 || src_dtype == MST_SRC_I16 || src_dtype == MST_SRC_U16, "mst_prepare_volume: bad src_dtype %d", src_dtype);
    MST_REQUIRE(items >= 1 && W0 >= 1 && H0 >= 1 && D0 >= 1 && W >= 1 && H >= 1 && D >= 1, "mst_prepare_volume: bad shape");
    MST_REQUIRE(q_lo >= 0.f && q_lo <= q_hi && q_hi <= 1.f, "mst_prepare_volume: quantiles must satisfy 0 <= q_lo <= q_hi <= 1");
    const size_t base_bytes = prepare_volume_workspace_bytes(items, W0, H0, D0);
    MST_REQUIRE(workspace_bytes >= base_bytes + (src_dtype == MST_SRC_F32 ? 0 : prepare_volume_raw_bytes(items, W0, H0, D0)),
                "mst_prepare_volume: workspace too small");
    int dev = 0, sms = 0;
    MST_CHECK_CUDA(cudaGetDevice(&dev));
    MST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    mst_handle_s none;
    if (!h) h = &none;
    const float* src = static_cast<const float*>(src_any);
    if (src_dtype != MST_SRC_F32) {   // raw int16 / uint16 voxels: widen on the device (the host-to-device copy carried 2 bytes per voxel)
        float* raw = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + base_bytes);
        MST_LAUNCH(CAT_PREPARE_VOLUME, launch_raw_to_f32(src_any, src_dtype, raw, static_cast<int64_t>(items) * W0 * H0 * D0, sms, st));
        src = raw;
    }
    MST_LAUNCH(CAT_PREPARE_VOLUME, launch_prepare_volume(src, items, W0, H0, D0, W, H, D, flip_h, q_lo, q_hi, out, stats, workspace, sms, st));
    h->launches += launch_prepare_volume_count(W0, H0, D0, W, H, D) - 1;   // the chain is several kernels, timed as one category
    return 0;
}

int mst_slice_train_bytes(int32_t B, int32_t D, int32_t E, int32_t heads, int32_t C, size_t* saved_bytes, size_t* factor_bytes) {
    MST_REQUIRE(saved_bytes && factor_bytes && B >= 1 && D >= 1 && E >= 1 && heads >= 1 && C >= 1, "mst_slice_train_bytes: bad argument");
    *saved_bytes = slice_train_saved_bytes(B, D, E, heads);
    *factor_bytes = slice_train_factor_bytes(B, E, heads, C);
    return 0;
}
int mst_slice_train_forward(mst_handle h, const float* enc, const uint8_t* pad_mask, const float* const* params, int32_t B, int32_t D,
                            int32_t E, int32_t heads, int32_t C, float* saved, float* logits, void* stream) {
    MST_REQUIRE(enc && params && saved && logits && B >= 1 && D >= 1, "mst_slice_train_forward: bad argument");
    for (int i = 0; i < 17; ++i) MST_REQUIRE(params[i] != nullptr, "mst_slice_train_forward: parameter %d is null", i);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    mst_handle_s none;
    if (!h) h = &none;
    MST_LAUNCH(CAT_TRAIN_FORWARD, launch_slice_train_forward(enc, pad_mask, params, saved, logits, B, D, E, heads, C, st));
    return 0;
}
int mst_slice_train_backward(mst_handle h, const float* enc, const float* dlogits, const float* const* params, const float* saved,
                             float* factors, float* const* grads, float* denc, int32_t B, int32_t D, int32_t E, int32_t heads, int32_t C,
                             void* stream) {
    MST_REQUIRE(enc && dlogits && params && saved && factors && grads && B >= 1 && D >= 1, "mst_slice_train_backward: bad argument");
    for (int i = 0; i < 17; ++i) MST_REQUIRE(params[i] != nullptr && grads[i] != nullptr, "mst_slice_train_backward: tensor %d is null", i);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    mst_handle_s none;
    if (!h) h = &none;
    MST_LAUNCH(CAT_TRAIN_BACKWARD, launch_slice_train_backward(enc, dlogits, params, saved, factors, grads, denc, B, D, E, heads, C, st));
    h->launches += 1;   // kernel A (per volume) + kernel B (per weight row)
    return 0;
}
int mst_adamw(mst_handle h, float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
              float weight_decay, int32_t step, float grad_scale, void* stream) {
    MST_REQUIRE(p && g && m && v, "mst_adamw: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    mst_handle_s none;
    if (!h) h = &none;
    int dev = 0, sms = 0;
    MST_CHECK_CUDA(cudaGetDevice(&dev));
    MST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MST_LAUNCH(CAT_ADAMW, launch_adamw(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, sms, st));
    return 0;
}

int mst_train_workspace_bytes(mst_handle h, int32_t B, int32_t D, int32_t H, int32_t W, size_t* bytes) {
    MST_REQUIRE(h && bytes, "mst_train_workspace_bytes: null argument");
    MST_PROPAGATE(check_shape(h, B, D, H, W));
    MST_PROPAGATE(check_trainable(h, H, W));
    *bytes = carve_train(h->cfg, B, D, H, W, nullptr).total;
    return 0;
}
int mst_set_grad(mst_handle h, const char* name, float* dev_fp32, int64_t numel) {
    MST_REQUIRE(h && name, "mst_set_grad: null argument");
    const std::string key = canonical_name(name);
    auto it = h->expected.find(key);
    MST_REQUIRE(it != h->expected.end(), "mst_set_grad: unknown tensor '%s'", name);
    MST_REQUIRE(dev_fp32 == nullptr || it->second == numel, "mst_set_grad: '%s' has %lld elements, expected %lld", name, (long long)numel,
                (long long)it->second);
    if (dev_fp32) h->grad_ptr[key] = dev_fp32; else h->grad_ptr.erase(key);
    return 0;
}
int mst_train_forward(mst_handle h, const void* src, int32_t src_dtype, int32_t B, int32_t D, int32_t H, int32_t W, float* enc_cls,
                      void* workspace, size_t workspace_bytes, void* stream) {
    MST_REQUIRE(h && src && enc_cls && workspace, "mst_train_forward: null argument");
    MST_REQUIRE(h->finalized, "mst_train_forward: weights not finalized");
    MST_REQUIRE(src_dtype >= MST_SRC_F32 && src_dtype <= MST_SRC_F16, "mst_train_forward: bad src_dtype %d", src_dtype);
    MST_PROPAGATE(check_shape(h, B, D, H, W));
    MST_PROPAGATE(check_trainable(h, H, W));
    MST_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "mst_train_forward: workspace must be 256-byte aligned");
    TrainWs ws = carve_train(h->cfg, B, D, H, W, static_cast<uint8_t*>(workspace));
    MST_REQUIRE(ws.total <= workspace_bytes, "mst_train_forward: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    MST_CHECK_CUDA(cudaSetDevice(h->cfg.device));
    set_pdl(false);
    return train_forward(h, src, src_dtype, B, D, H, W, enc_cls, ws, static_cast<cudaStream_t>(stream));
}
int mst_train_backward(mst_handle h, const float* denc, int32_t B, int32_t D, int32_t H, int32_t W, void* workspace, size_t workspace_bytes,
                       void* stream) {
    MST_REQUIRE(h && denc && workspace, "mst_train_backward: null argument");
    MST_PROPAGATE(check_shape(h, B, D, H, W));
    MST_PROPAGATE(check_trainable(h, H, W));
    TrainWs ws = carve_train(h->cfg, B, D, H, W, static_cast<uint8_t*>(workspace));
    MST_REQUIRE(ws.total <= workspace_bytes, "mst_train_backward: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    MST_CHECK_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    set_pdl(false);
    MST_PROPAGATE(refresh_dgrad_weights(h, st));
    return train_backward(h, denc, B, D, H, W, ws, st);
}

const char* mst_profile_categories(void) { return kCatNames; }
unsigned long long mst_launch_count(mst_handle h) { return h ? h->launches : 0; }
int mst_profile_begin(mst_handle h) {
    MST_REQUIRE(h, "mst_profile_begin: null handle");
    h->prof.on = true; h->prof.used = 0; h->prof.recs.clear();
    return 0;
}
int mst_profile_end(mst_handle h, double* ms, int64_t* launches, int32_t n) {
    MST_REQUIRE(h && ms && launches && n >= NUM_CAT, "mst_profile_end: need room for %d categories", (int)NUM_CAT);
    MST_CHECK_CUDA(cudaSetDevice(h->cfg.device));
    MST_CHECK_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < n; ++i) { ms[i] = 0.0; launches[i] = 0; }
    for (auto& r : h->prof.recs) {
        float t = 0.f;
        MST_CHECK_CUDA(cudaEventElapsedTime(&t, h->prof.pool[r.e0], h->prof.pool[r.e1]));
        ms[r.cat] += t; launches[r.cat] += 1;
    }
    h->prof.on = false; h->prof.used = 0; h->prof.recs.clear();
    return 0;
}

static int num_sms_current() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

int mst_kernel_gemm_bf16(const void* A, const void* W, int32_t M, int32_t N, int32_t K, int32_t mode, const float* bias,
                         const void* res, void* out, void* stream) {
    MST_REQUIRE(A && W && out && bias && mode >= 0 && mode <= 2, "mst_kernel_gemm_bf16: bad argument");
    EpiParams ep{};
    ep.bias = bias; ep.res = res; ep.ldr = N; ep.out = out; ep.ldo = N;
    return gemm_bf16_tc(static_cast<const bf16*>(A), static_cast<const bf16*>(W), M, N, K, mode, ep, num_sms_current(),
                        static_cast<cudaStream_t>(stream));
}
int mst_kernel_gemm_bf16_ln(const void* A, const void* W, int32_t M, int32_t N, int32_t K, int32_t gelu, const float* bias,
                            const float* rowstat, void* out, void* stream) {
    MST_REQUIRE(A && W && out && bias && rowstat, "mst_kernel_gemm_bf16_ln: null argument");
    EpiParams ep{};
    ep.bias = bias; ep.rowstat = rowstat; ep.out = out; ep.ldo = N;
    return gemm_bf16_tc(static_cast<const bf16*>(A), static_cast<const bf16*>(W), M, N, K, gelu ? EPI_LN_BIAS_GELU : EPI_LN_BIAS, ep,
                        num_sms_current(), static_cast<cudaStream_t>(stream));
}
int mst_kernel_pack_linear_ln(const float* W, const float* b, const float* gamma, const float* beta, int32_t N, int32_t K,
                              void* Wd_bf16, float* bd, void* stream) {
    MST_REQUIRE(W && b && gamma && beta && Wd_bf16 && bd && N >= 1 && K >= 1 && K <= kPackLnMaxK, "mst_kernel_pack_linear_ln: bad argument");
    pack_linear_ln_kernel<<<N, 128, 0, static_cast<cudaStream_t>(stream)>>>(W, b, gamma, beta, 0, 1.f, static_cast<bf16*>(Wd_bf16), bd, K);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int mst_kernel_gemm_bf16_res_stats(const void* A, const void* W, int32_t M, int32_t N, int32_t K, const float* bias, void* x,
                                   float* rowstat_out, float eps, void* stream) {
    MST_REQUIRE(A && W && bias && x && rowstat_out, "mst_kernel_gemm_bf16_res_stats: null argument");
    EpiParams ep{};
    ep.bias = bias; ep.res = x; ep.ldr = N; ep.out = x; ep.ldo = N; ep.rowstat_out = rowstat_out; ep.stat_eps = eps;
    return gemm_bf16_tc(static_cast<const bf16*>(A), static_cast<const bf16*>(W), M, N, K, EPI_BIAS_RES, ep, num_sms_current(),
                        static_cast<cudaStream_t>(stream));
}
int mst_kernel_gemm_bf16_f32out(const void* A, const void* W, int32_t M, int32_t N, int32_t K, float* out, void* stream) {
    MST_REQUIRE(A && W && out, "mst_kernel_gemm_bf16_f32out: null argument");
    EpiParams ep{};
    ep.out = out; ep.ldo = N;
    return gemm_bf16_tc(static_cast<const bf16*>(A), static_cast<const bf16*>(W), M, N, K, EPI_RAW_F32, ep, num_sms_current(),
                        static_cast<cudaStream_t>(stream));
}
int mst_kernel_wgrad_bf16(const void* dY, const void* X, int32_t M, int32_t Nout, int32_t Kin, float* dW, float* db, void* stream) {
    MST_REQUIRE(dY && X && dW, "mst_kernel_wgrad_bf16: null argument");
    MST_REQUIRE(wgrad_tc_supported(Nout, Kin), "mst_kernel_wgrad_bf16: Nout must be a multiple of 128 and Kin of 192 (got %d, %d)", Nout, Kin);
    return launch_wgrad_tc(static_cast<const bf16*>(dY), Nout, static_cast<const bf16*>(X), Kin, M, dW, db, num_sms_current(),
                           static_cast<cudaStream_t>(stream));
}
int mst_kernel_ln_bwd_bf16(const void* x, const void* dy, const void* dres, const float* gamma, void* dx, float* dgamma, float* dbeta,
                           int32_t rows, int32_t E, float eps, void* stream) {
    MST_REQUIRE(x && dy && gamma && dx && dgamma && dbeta, "mst_kernel_ln_bwd_bf16: null argument");
    float* ws = nullptr;
    MST_CHECK_CUDA(cudaMalloc(&ws, ln_bwd_workspace_bytes(E)));
    const int rc = launch_ln_bwd(static_cast<const bf16*>(x), E, static_cast<const bf16*>(dy), nullptr, static_cast<const bf16*>(dres), gamma,
                                 static_cast<bf16*>(dx), dgamma, dbeta, rows, E, eps, ws, static_cast<cudaStream_t>(stream));
    cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    cudaFree(ws);
    return rc;
}
int mst_kernel_gelu_bf16(const void* u, void* y, const void* dy, void* du, int64_t n, void* stream) {
    MST_REQUIRE(u && (y || (dy && du)), "mst_kernel_gelu_bf16: null argument");
    if (y) MST_PROPAGATE(launch_gelu_fwd(static_cast<const bf16*>(u), static_cast<bf16*>(y), n, num_sms_current(), static_cast<cudaStream_t>(stream)));
    if (dy) MST_PROPAGATE(launch_gelu_bwd(static_cast<const bf16*>(u), static_cast<const bf16*>(dy), static_cast<bf16*>(du), n, num_sms_current(),
                                          static_cast<cudaStream_t>(stream)));
    return 0;
}
int mst_kernel_transpose_bf16(const void* in, void* out, float* colsum, int32_t M, int32_t C, int32_t Mpad, void* stream) {
    MST_REQUIRE(in && out, "mst_kernel_transpose_bf16: null argument");
    return launch_transpose_colsum(static_cast<const bf16*>(in), C, static_cast<bf16*>(out), colsum, M, C, Mpad, static_cast<cudaStream_t>(stream));
}
int mst_kernel_attention_bwd_bf16(const void* qkv, const void* o, const void* dO, void* dqkv, int32_t BD, int32_t N, int32_t heads, void* stream) {
    MST_REQUIRE(qkv && o && dO && dqkv, "mst_kernel_attention_bwd_bf16: null argument");
    return launch_attention_bwd(static_cast<const bf16*>(qkv), static_cast<const bf16*>(o), static_cast<const bf16*>(dO), static_cast<bf16*>(dqkv),
                                BD, N, heads, static_cast<cudaStream_t>(stream));
}
int mst_kernel_attention_lse_bf16(const void* qkv, void* out, float* lse, int32_t BD, int32_t heads, void* stream) {
    MST_REQUIRE(qkv && out && lse, "mst_kernel_attention_lse_bf16: null argument");
    return launch_attention_tc257x16(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), BD, heads, num_sms_current(),
                                     static_cast<cudaStream_t>(stream), nullptr, lse);
}
int mst_kernel_attention_bwd_lse_bf16(const void* qkv, const void* o, const void* dO, const float* lse, void* dqkv, int32_t BD, int32_t N,
                                      int32_t heads, void* stream) {
    MST_REQUIRE(qkv && o && dO && lse && dqkv, "mst_kernel_attention_bwd_lse_bf16: null argument");
    return launch_attention_bwd(static_cast<const bf16*>(qkv), static_cast<const bf16*>(o), static_cast<const bf16*>(dO), static_cast<bf16*>(dqkv),
                                BD, N, heads, static_cast<cudaStream_t>(stream), lse);
}
int mst_kernel_row_stats_bf16(const void* x, float* rowstat, int32_t rows, int32_t E, float eps, void* stream) {
    MST_REQUIRE(x && rowstat, "mst_kernel_row_stats_bf16: null argument");
    return launch_row_stats(static_cast<const bf16*>(x), rowstat, rows, E, eps, static_cast<cudaStream_t>(stream));
}
int mst_debug_gemm_timing(const void* A, const void* W, int32_t M, int32_t N, int32_t K, int32_t mode, const float* bias,
                          const void* res, void* out, long long* dbg_dev, void* stream) {
    MST_REQUIRE(A && W && out && bias && dbg_dev && mode >= 0 && mode <= 2, "mst_debug_gemm_timing: bad argument");
    MST_REQUIRE(kDbgTiming, "mst_debug_gemm_timing: the phase counters are compiled in only with -DMST_EXPERIMENTS (build.py --experiments)");
    EpiParams ep{};
    ep.bias = bias; ep.res = res; ep.ldr = N; ep.out = out; ep.ldo = N; ep.dbg = dbg_dev;
    return gemm_bf16_tc(static_cast<const bf16*>(A), static_cast<const bf16*>(W), M, N, K, mode, ep, num_sms_current(),
                        static_cast<cudaStream_t>(stream));
}
int mst_kernel_gemm_f32(const float* A, const float* W, int32_t M, int32_t N, int32_t K, int32_t mode, const float* bias,
                        const float* res, float* out, void* stream) {
    MST_REQUIRE(A && W && out && bias && mode >= 0 && mode <= 2, "mst_kernel_gemm_f32: bad argument");
    EpiParams ep{};
    ep.bias = bias; ep.res = res; ep.ldr = N; ep.out = out; ep.ldo = N;
    return gemm_f32_simt(A, K, W, M, N, K, mode, ep, static_cast<cudaStream_t>(stream));
}
int mst_kernel_attention_bf16(const void* qkv, void* out, int32_t BD, int32_t N, int32_t heads, void* stream) {
    MST_REQUIRE(qkv && out, "mst_kernel_attention_bf16: null argument");
    if (N == 257)
        return launch_attention_tc257x16(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), BD, heads, num_sms_current(),
                                         static_cast<cudaStream_t>(stream));
    if (attention_tcg_supported(N))
        return launch_attention_tcg(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), BD, N, heads, num_sms_current(),
                                    static_cast<cudaStream_t>(stream));
    return launch_attention_bf16(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), BD, N, heads, static_cast<cudaStream_t>(stream));
}
int mst_kernel_attention_bf16_warp_mma(const void* qkv, void* out, int32_t BD, int32_t N, int32_t heads, void* stream) {
    MST_REQUIRE(qkv && out, "mst_kernel_attention_bf16_warp_mma: null argument");
    return launch_attention_bf16(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), BD, N, heads, static_cast<cudaStream_t>(stream));
}
int mst_kernel_attention_f32(const float* qkv, float* out, int32_t BD, int32_t N, int32_t heads, void* stream) {
    MST_REQUIRE(qkv && out, "mst_kernel_attention_f32: null argument");
    return launch_attention_f32(qkv, out, BD, N, heads, static_cast<cudaStream_t>(stream));
}
int mst_kernel_layernorm_bf16(const void* x, void* y, const float* gamma, const float* beta, int32_t rows, int32_t E,
                              float eps, void* stream) {
    MST_REQUIRE(x && y && gamma && beta, "mst_kernel_layernorm_bf16: null argument");
    return launch_layernorm<bf16, bf16>(static_cast<const bf16*>(x), E, static_cast<bf16*>(y), E, gamma, beta, rows, E, eps,
                                        static_cast<cudaStream_t>(stream));
}

}  // extern "C"
