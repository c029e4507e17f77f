// Shared declarations for the MST-DINOv2 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>

namespace mst {

typedef __nv_bfloat16 bf16;

// ---- error plumbing (C-ABI returns int status; message via mst_last_error) -----------------------
void set_error(const char* fmt, ...);
#define MST_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            mst::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)
#define MST_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            mst::set_error(__VA_ARGS__);       \
            return 1;                          \
        }                                      \
    } while (0)
#define MST_PROPAGATE(expr)        \
    do {                           \
        int _s = (expr);           \
        if (_s != 0) return _s;    \
    } while (0)

// ---- experiment switches and profiling counters: compiled in only with -DMST_EXPERIMENTS (new-vit_b200/build.py --experiments);
// the product library reads no environment variable and carries no clock64() on its hot paths.
#ifdef MST_EXPERIMENTS
inline int exp_env(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
constexpr bool kDbgTiming = true;
#else
inline int exp_env(const char*, int dflt) { return dflt; }
constexpr bool kDbgTiming = false;
#endif
#define MST_DBG_CLOCK() (mst::kDbgTiming ? clock64() : 0LL)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: set it once per (kernel, device), so a
// second GPU used by the same process (a model on cuda:1 after one on cuda:0, one thread per GPU) gets it too.
constexpr int kMaxDevices = 64;
#define MST_SET_DYN_SMEM(kern, bytes)                                                                                   \
    do {                                                                                                                \
        static bool _done[mst::kMaxDevices] = {};                                                                       \
        int _dev = 0;                                                                                                   \
        MST_CHECK_CUDA(cudaGetDevice(&_dev));                                                                           \
        if (_dev < 0 || _dev >= mst::kMaxDevices || !_done[_dev]) {                                                     \
            MST_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));             \
            if (_dev >= 0 && _dev < mst::kMaxDevices) _done[_dev] = true;                                               \
        }                                                                                                               \
    } while (0)

// ---- GEMM epilogues --------------------------------------------------------------------------------
// C[row, n] = epi(acc[row, n]) for  acc = A[M,K] . W[N,K]^T  (nn.Linear convention: W is [out, in]).
enum EpiMode : int {
    EPI_BIAS = 0,       // out = acc + bias[n]
    EPI_BIAS_GELU = 1,  // out = gelu_erf(acc + bias[n])                (reference mlp.py:35-36)
    EPI_BIAS_RES = 2,   // out = res[row*ldr + n] + acc + bias[n]       (block.py:112-113; LayerScale folded in W,b)
    EPI_PATCH = 3,      // out[(row/P)*(P+1+R) + 1 + R + row%P, n] = acc + posb[(row%P)*N + n]   (patch_embed.py:75-77 +
                        //                                      vision_transformer.py:219-220; bias folded in posb)
    EPI_BIAS_ACCUM = 4, // out[row, n] += acc + bias[n]   (in-place residual update; bf16 path: TMA reduce-add)
    // LayerNorm folded into the GEMM that consumes it (bf16 path): A holds the RAW residual rows; the weight is scaled by the
    // LayerNorm weight and every row is CENTRED, Wc[n,k] = gamma[k] W[n,k] - mean_k(gamma W[n,:]), so that x . Wc[n,:] =
    // (x - mean(x)) . (gamma W[n,:]) -- the mean subtraction happens inside the MMA; bias[n] = b[n] + sum_k beta[k] W[n,k];
    // rowstat[row] = rstd:   LN(x) W^T + b  =  rstd * acc + bias[n]
    EPI_LN_BIAS = 5,
    EPI_LN_BIAS_GELU = 6,
    EPI_RAW_F32 = 7     // out[row, n] = acc as fp32 (weight gradients: dW = dY^T X, train_enc.cu); no bias
};

struct EpiParams {
    const float* bias;   // [N] fp32 (EPI_BIAS*, EPI_BIAS_RES)
    const void* res;     // residual rows, same dtype as out (EPI_BIAS_RES)
    int64_t ldr;         // residual row stride in elements
    const float* posb;   // [P, N] fp32 (EPI_PATCH)
    int P;               // patches per slice (EPI_PATCH)
    int R;               // register tokens between the CLS row and the patch rows (EPI_PATCH)
    void* out;           // output, dtype T
    int64_t ldo;         // output row stride in elements
    long long* dbg;      // nullable: phase cycle counters of CTA 0 (profiles/gemm_timing.py)
    const float* rowstat;   // [M] rstd of every A row (EPI_LN_*)
    float* rowstat_out;     // nullable, EPI_BIAS_RES on the streaming pair GEMM: rstd of every OUTPUT row (the LayerNorm statistics
                            // the next GEMM needs), computed in the epilogue from the bf16-rounded rows it writes
    float stat_eps;         // LayerNorm eps for rowstat_out / rowpart
    float* rowpart_out;     // nullable, EPI_BIAS_RES on the weight-resident GEMM with N = 384: [M][4][2] partial (sum, sum of squares)
                            // of every output row, slot = 2 * n_block + column group (plain stores, fixed slots: deterministic)
    const float* rowpart;   // gemm_wt, EPI_LN_*: the same partials as INPUT (instead of rowstat); rstd is formed on the fly
};

// erf via the rational minimax on [-4,4] (max abs error 3.8e-7 in fp32; checked against math.erf).
__device__ __forceinline__ float erf_fast(float x) {
    x = fminf(fmaxf(x, -4.0f), 4.0f);
    const float x2 = x * x;
    float p = -2.72614225801306e-10f;
    p = fmaf(p, x2, 2.77068142495902e-08f);
    p = fmaf(p, x2, -2.10102402082508e-06f);
    p = fmaf(p, x2, -5.69250639462346e-05f);
    p = fmaf(p, x2, -7.34990630326855e-04f);
    p = fmaf(p, x2, -2.95459980854025e-03f);
    p = fmaf(p, x2, -1.60960333262415e-02f);
    p *= x;
    float q = -1.45660718464996e-05f;
    q = fmaf(q, x2, -2.13374055278905e-04f);
    q = fmaf(q, x2, -1.68282697438203e-03f);
    q = fmaf(q, x2, -7.37332916720468e-03f);
    q = fmaf(q, x2, -1.42647390514189e-02f);
    return __fdividef(p, q);
}
// bf16-path GELU: x*Phi(x) with Phi(x) ~ 0.5*(1 + tanh(x*(c0 + c1 x^2 + c2 x^4))), coefficients fitted to the exact
// erf form (max abs deviation 2.5e-5 over all x; the bf16 output ulp is >= 2.4e-4 for |y| >= 0.06), one MUFU op.
// x^2 is clamped so the odd polynomial keeps its sign for outliers (|x| > 8 -> tanh = +-1 -> y = x or 0).
__device__ __forceinline__ float gelu_tanh_fit(float x) {
    float t, x2 = fminf(x * x, 64.0f);
    float p = fmaf(-3.51516788e-04f, x2, 3.70056460e-02f);
    p = fmaf(p, x2, 7.97507884e-01f);
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}
// the same function on a packed pair: 4 packed FP32 ops + 2 FMNMX + 2 MUFU per two elements
__device__ __forceinline__ unsigned long long gelu_tanh_fit2(unsigned long long x) {
    unsigned long long x2, p, u, hx, t, y;
    float a, b;
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(x2) : "l"(x));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x2));
    a = fminf(a, 64.0f); b = fminf(b, 64.0f);
    asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "f"(a), "f"(b));
    constexpr unsigned long long C2 = 0xB9B84BC9B9B84BC9ULL;  // -3.51516788e-04f twice
    constexpr unsigned long long C1 = 0x3D17933B3D17933BULL;  //  3.70056460e-02f
    constexpr unsigned long long C0 = 0x3F4C297A3F4C297AULL;  //  7.97507884e-01f
    constexpr unsigned long long HALF = 0x3F0000003F000000ULL;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(C2), "l"(x2), "l"(C1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(p), "l"(x2), "l"(C0));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(u) : "l"(x), "l"(p));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(u));
    asm("tanh.approx.f32 %0, %0;" : "+f"(a));
    asm("tanh.approx.f32 %0, %0;" : "+f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(t) : "f"(a), "f"(b));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(hx) : "l"(x), "l"(HALF));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(y) : "l"(hx), "l"(t), "l"(hx));
    return y;
}
template <bool kExact>
__device__ __forceinline__ float gelu_erf(float x) {
    const float e = kExact ? erff(x * 0.70710678118654752f) : erf_fast(x * 0.70710678118654752f);
    const float hx = 0.5f * x;
    return fmaf(hx, e, hx);
}

// ---- kernel launchers (each returns 0 or sets the error and returns non-zero) ----------------------
struct TmaDesc {  // opaque 128-byte CUtensorMap
    alignas(64) uint8_t bytes[128];
};
int tma_init();  // resolves cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency)
int make_tma_2d_bf16(TmaDesc* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                     uint32_t box_inner, uint32_t box_rows, bool swizzle128 = true, bool swizzle64 = false);

int make_tma_3d_bf16(TmaDesc* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                     uint64_t stride2_elems, uint32_t box0, uint32_t box1, bool swizzle128);
// weights-in-TMEM transposed-accumulator GEMM for K = 384, N >= 1024 (qkv, fc1): gemm_wt.cu
struct EpiParams;
bool gemm_wt_supported(int M, int N, int K, int mode, const EpiParams& ep);
bool gemm_wt_enabled();   // MST_GEMM_WT != 0
bool pdl_enabled();       // GEMM / attention kernels are launched as programmatic dependents (ptx.cuh): small batches only
void set_pdl(bool on);
int gemm_bf16_wt(const bf16* A, const bf16* W, int M, int N, int K, int mode, const EpiParams& ep, int num_sms,
                 cudaStream_t stream);

// bf16 tensor-core GEMM (tcgen05 + TMA + TMEM). A: [M,K] bf16 row-major (lda=K), W: [N,K] bf16 row-major.
int gemm_bf16_tc(const bf16* A, const bf16* W, int M, int N, int K, int mode, const EpiParams& ep, int num_sms,
                 cudaStream_t stream);
// fp32 CUDA-core GEMM (fp32 parity mode). A: [M,K] fp32 (lda), W: [N,K] fp32.
int gemm_f32_simt(const float* A, int64_t lda, const float* W, int M, int N, int K, int mode, const EpiParams& ep,
                  cudaStream_t stream);

template <typename T>
int launch_im2col(const void* src, int src_dtype, T* A0, T* x, const float* cls_pos0, const float* regs, int R, int BD, int H, int W,
                  int KP, int E, int tta_BD, int D, cudaStream_t stream);
template <typename TIn, typename TOut>
int launch_layernorm(const TIn* x, int64_t ldx, TOut* y, int64_t ldy, const float* gamma, const float* beta, int rows,
                     int E, float eps, cudaStream_t stream);
// rowstat[row] = rstd of x[row, 0..E) (fp32 statistics, two-pass), the per-row part of a folded LayerNorm
int launch_row_stats(const bf16* x, float* rowstat, int rows, int E, float eps, cudaStream_t stream);
int launch_attention_bf16(const bf16* qkv, bf16* out, int BD, int N, int heads, cudaStream_t stream);
// tcgen05 attention for N == 257 tokens (ViT @224), sixteen softmax warps (two per TMEM lane quadrant and tile, splitting the key
// columns): attention_tc16.cu
int launch_attention_tc257x16(const bf16* qkv, bf16* out, int BD, int heads, int num_sms, cudaStream_t stream,
                              long long* dbg = nullptr, float* lse_out = nullptr);   // lse_out [BD*heads, 257]: training forward

// tcgen05 attention for any token count 17 <= N <= 352 (attention_tcg.cu); N == 257 keeps its specialised kernels
bool attention_tcg_supported(int N);
int launch_attention_tcg(const bf16* qkv, bf16* out, int BD, int N, int heads, int num_sms, cudaStream_t stream);
int launch_attention_f32(const float* qkv, float* out, int BD, int N, int heads, cudaStream_t stream);
template <typename T>
int launch_cls_attention(const T* qkv, T* out_cls, float* plane_cls, int BD, int N, int heads, cudaStream_t stream);

enum SliceFusionMode : int { SLICE_FUSION_TRANSFORMER = 0, SLICE_FUSION_LINEAR = 1, SLICE_FUSION_AVERAGE = 2 };  // dino.py:80-101
struct SliceWeights {  // fp32, linear weights pre-transposed to [in][out]; bott_* / pos_emb nullable (dino.py:75-82)
    const float *cls_token, *n1w, *n1b, *in_wt, *in_b, *out_wt, *out_b, *n2w, *n2b, *l1_wt, *l1_b, *l2_wt, *l2_b, *nfw,
        *nfb, *head_wt, *head_b, *bott_wt, *bott_b, *pos_emb, *in_w /* in_proj_weight as stored, [3E][E] */,
        *rope_freqs /* nullable [head_dim/2]: RoPE on the slice tokens (transformer_blocks.py:262-264) */;
    int liere;  /* rotary_positional_encoding='LiRE' (batch 1, 33 tokens): q / k slots re-read as the reference's view does */
};
// enc_cls [B*D, Eenc]; E = slice embedding (Eenc, or Eenc/4 behind the bottleneck); logits/feat nullable
int launch_slice_fusion(const float* enc_cls, const uint8_t* pad_mask, const SliceWeights& w, float* hs_scratch,
                        float* logits, float* feat, float* slice_cls, int B, int D, int Eenc, int E, int heads, int out_ch,
                        int mode, int mask_period, cudaStream_t stream);
// tta: plane_cls / slice_cls hold 8 flipped variants per volume (variant-major); coarse / slice_attn are the un-flipped averages
int launch_saliency_combine(const float* plane_cls, const float* slice_cls, int B, int D, int heads, int slice_heads, int skip, int gh,
                            int gw, int tta, float* attn_maps, float* plane_attn, float* slice_attn, float* coarse, cudaStream_t stream);
int launch_saliency_upsample(const float* coarse, float* full, int B, int D, int gh, int gw, int H, int W, cudaStream_t stream);


// ---- kernels either side of the main path (extras.cu) ---------------------------------------------
int launch_pos_bicubic(const float* pos, const float* cbias, float* posb, int M, int gh, int gw, int E, float scale_y,
                       float scale_x, cudaStream_t stream);
// anti-aliased variant with an exact output size (hub "_reg" encoders: interpolate_antialias=True, interpolate_offset=0.0)
int launch_pos_bicubic_aa(const float* pos, const float* cbias, float* posb, int M, int gh, int gw, int E, cudaStream_t stream);
template <typename T>
int launch_attention_probs(const T* qkv, float* probs, int BD, int N, int heads, cudaStream_t stream);
int launch_rollout(const float* maps, int depth, int nmat, int N, float* out, float* scratch, cudaStream_t stream);
size_t quantile_workspace_bytes(int items, int nq);
int launch_quantile(const float* data, int64_t n, int items, const double* q_dev, int nq, double* out, void* workspace,
                    int num_sms, cudaStream_t stream);

// ---- input pipeline in front of the forward (prep.cu) -----------------------
size_t prepare_volume_workspace_bytes(int items, int W0, int H0, int D0);
int launch_prepare_volume_count(int W0, int H0, int D0, int W, int H, int D);
size_t prepare_volume_raw_bytes(int items, int W0, int H0, int D0);
int launch_raw_to_f32(const void* in, int dtype, float* out, int64_t n, int num_sms, cudaStream_t stream);
int launch_prepare_volume(const float* src, int items, int W0, int H0, int D0, int W, int H, int D, int flip_h, float q_lo,
                          float q_hi, float* out, double* stats, void* workspace, int num_sms, cudaStream_t stream);

// ---- training step of the slice transformer + head on a frozen encoder (train.cu; BASELINE config 5, frozen-encoder slice) ----
// params / grads: 17 device pointers in this order: cls_token, norm1.weight, norm1.bias, in_proj_weight, in_proj_bias,
// out_proj.weight, out_proj.bias, norm2.weight, norm2.bias, linear1.weight, linear1.bias, linear2.weight, linear2.bias,
// slice_fusion.norm.weight, slice_fusion.norm.bias, linear.weight, linear.bias (nn.Linear layout, fp32)
size_t slice_train_saved_bytes(int B, int D, int E, int heads);
size_t slice_train_factor_bytes(int B, int E, int heads, int C);
int launch_slice_train_forward(const float* enc, const uint8_t* pad_mask, const float* const* params, float* saved, float* logits,
                               int B, int D, int E, int heads, int C, cudaStream_t stream);
int launch_slice_train_backward(const float* enc, const float* dlogits, const float* const* params, const float* saved, float* factors,
                                float* const* grads, float* denc, int B, int D, int E, int heads, int C, cudaStream_t stream);
int launch_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float wd, int step,
                 float grad_scale, int num_sms, cudaStream_t stream);

// ---- encoder backward pieces (train_enc.cu) ----
size_t ln_bwd_workspace_bytes(int E);
int launch_ln_bwd(const bf16* x, int64_t x_row_stride, const bf16* dy, const float* dy_f32, const bf16* dres, const float* gamma, bf16* dx,
                  float* dgamma, float* dbeta, int rows, int E, float eps, float* workspace, cudaStream_t stream);
int launch_gelu_fwd(const bf16* u, bf16* y, int64_t n, int num_sms, cudaStream_t stream);
int launch_gelu_bwd(const bf16* u, const bf16* dy, bf16* du, int64_t n, int num_sms, cudaStream_t stream);
int launch_transpose_colsum(const bf16* in, int64_t ld, bf16* out, float* colsum, int M, int C, int Mpad, cudaStream_t stream);
int launch_transpose_f32_to_bf16(const float* W, bf16* Wt, int N, int K, cudaStream_t stream);
bool wgrad_tc_supported(int Nout, int Kin);
int launch_wgrad_tc(const bf16* dY, int Nout, const bf16* X, int Kin, int M, float* dW, float* db, int num_sms, cudaStream_t stream);
int launch_attention_bwd(const bf16* qkv, const bf16* o, const bf16* dO, bf16* dqkv, int BD, int N, int heads, cudaStream_t stream,
                         const float* lse = nullptr);   // lse [BD*heads, N] from the forward (nullable: recomputed)

}  // namespace mst
