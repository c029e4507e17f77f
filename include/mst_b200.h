/* mst_b200.h -- C ABI of the B200-native MST-DINOv2 hot path (libmst_b200.so).
 *
 * The reference (gabrielfnayres/new-vit) has no FFI: its seam for this path is the Python class
 * `mst.models.DinoV2ClassifierSlice` (reference mst/models/dino.py:32-275).  These entry points are what a
 * binding for that class calls (ctypes stub: new-vit_b200/_cabi.py; INTEGRATION.md shows the reference-side
 * patch).  Plain pointers and sizes only; every pointer named "dev" / every tensor argument is a DEVICE
 * pointer owned by the caller; `stream` is a cudaStream_t passed as void*.  All functions return 0 on
 * success, non-zero on failure with a message available from mst_last_error().  No function synchronises
 * the device except mst_finalize_weights and mst_destroy.  One handle per GPU; a handle is not thread-safe.
 * There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef MST_B200_H_
#define MST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MST_ABI_VERSION 4
#if defined(__GNUC__)
#define MST_API __attribute__((visibility("default")))
#else
#define MST_API
#endif

enum { MST_PRECISION_FP32 = 0, MST_PRECISION_BF16 = 1 };
/* slice_fusion constructor argument (reference dino.py:80-101,144-157) */
enum { MST_FUSION_TRANSFORMER = 0, MST_FUSION_LINEAR = 1, MST_FUSION_AVERAGE = 2 };
enum { MST_ROTARY_NONE = 0, MST_ROTARY_ROPE = 1, MST_ROTARY_LIRE = 2 };
/* element type of the `source` volume handed to mst_forward.  The bf16 path rounds every voxel to bf16 before the patch GEMM
 * anyway, so a bf16 upload is bit-identical to an fp32 one at half the host-to-device bytes. */
enum { MST_SRC_F32 = 0, MST_SRC_BF16 = 1, MST_SRC_F16 = 2,
       MST_SRC_I16 = 3, MST_SRC_U16 = 4 /* raw scanner voxels: mst_prepare_volume only */ };

/* Architecture of one DinoV2ClassifierSlice instance.  Replaces the constructor arguments of
 * reference dino.py:33-103 (model_size -> embed_dim/depth/enc_heads per vision_transformer.py:340-396). */
typedef struct mst_config {
    int32_t embed_dim;   /* 384 (ViT-S/14), 768 (ViT-B/14), 1024 (ViT-L/14); head_dim is always 64 */
    int32_t depth;       /* encoder blocks, 12 for S/B; 0 = no encoder: a slice-transformer head for another backbone's features
                            (mst_slice_head_forward), embed_dim then is that backbone's feature width */
    int32_t enc_heads;   /* embed_dim / 64 */
    int32_t slice_heads; /* 12 (dino.py:87) */
    int32_t out_ch;      /* classes (dino.py:103) */
    int32_t pos_tokens;  /* rows of encoder.pos_embed = 1 + M*M the weights were built for (257 for the local factory
                            @224, 1370 for the hub checkpoints @518); other patch grids are bicubically resampled
                            (vision_transformer.py:179-211) */
    int32_t precision;   /* MST_PRECISION_FP32: CUDA-core parity mode; MST_PRECISION_BF16: tcgen05 tensor cores */
    int32_t device;      /* CUDA device ordinal */
    int32_t num_registers;     /* use_registers: 4 for the dinov2_vit*14_reg hub checkpoints (dino.py:60-61), else 0 */
    int32_t use_bottleneck;    /* dino.py:75-77: Linear(embed_dim -> embed_dim/4) on the per-slice features */
    int32_t use_slice_pos_emb; /* dino.py:81-82,140-142: nn.Embedding(256, emb) added to the slice tokens */
    int32_t slice_fusion;      /* MST_FUSION_* (dino.py:80-101) */
    int32_t enable_linear;     /* dino.py:103: 0 = nn.Identity head (forward returns the feature) */
    int32_t rotary;            /* rotary_positional_encoding (dino.py:40,92; utils/transformer_blocks.py:335-351):
                                  MST_ROTARY_NONE, or MST_ROTARY_ROPE = RoPE on the slice-token queries and keys; the
                                  checkpoint then carries slice_fusion.layers.0.self_attn.rotary_positional_encoding.freqs;
                                  MST_ROTARY_LIRE = 'LiRE' as the reference evaluates it (utils/rotary_embedding_torch.py:328-396):
                                  batch 1 and 32 slices only, carries ...rotary_positional_encoding.vars.{0,1} */
    int32_t interpolate_antialias; /* DinoVisionTransformer(interpolate_antialias=, interpolate_offset=) (vision_transformer.py:66-67,
                                      198-210): 0 / 0.1 for the vendored factory and the plain hub checkpoints; 1 / 0.0 for the hub
                                      "_reg" checkpoints (use_registers, dino.py:60-61) */
    float interpolate_offset;
} mst_config;

typedef struct mst_handle_s* mst_handle;

MST_API int mst_abi_version(void);
MST_API const char* mst_last_error(void);

/* dino.py:33 (__init__): allocate the weight store for `cfg` on cfg->device. */
MST_API int mst_create(const mst_config* cfg, mst_handle* out);
MST_API int mst_destroy(mst_handle h);

/* state_dict()/load_state_dict() (SURVEY.md section 5 key layout; base_model.py:67-81).  `name` is the
 * reference's state_dict key ("encoder.blocks.0.3.attn.qkv.weight", "slice_fusion.layers.0.linear1.bias",
 * "cls_token", ...; both the chunked "blocks.0.<i>" and the hub "blocks.<i>" spellings; optional
 * "ls1.gamma"/"ls2.gamma").  dev_fp32 holds `numel` fp32 values on the device.  "encoder.mask_token" is
 * accepted and ignored (unused by the path). */
MST_API int mst_set_weight(mst_handle h, const char* name, const float* dev_fp32, int64_t numel, void* stream);
/* The same for `count` tensors in one call (names / dev_fp32 / numel are HOST arrays): what a training loop does every step. */
MST_API int mst_set_weights(mst_handle h, int32_t count, const char* const* names, const float* const* dev_fp32, const int64_t* numel,
                    void* stream);
/* Pack the weights for the selected precision (bf16 conversion, 1/8 attention scale folded into Wq/bq,
 * LayerScale folded into proj/fc2, conv weight summed over the 3 identical RGB channels, slice-transformer
 * matrices transposed).  Fails if a required tensor was never set.  Synchronises `stream`. */
MST_API int mst_finalize_weights(mst_handle h, void* stream);

/* Bytes of scratch mst_forward needs for a [B,1,D,H,W] batch. */
MST_API int mst_workspace_bytes(mst_handle h, int32_t B, int32_t D, int32_t H, int32_t W, size_t* bytes);

/* DinoV2ClassifierSlice.forward (dino.py:110-167).
 *   src        [B,1,D,H,W] of src_dtype (MST_SRC_*; H, W multiples of 14; the fp32 parity mode takes fp32 only)
 *   pad_mask   nullable [B,D] uint8, non-zero = ignore slice (dino.py:147-150)
 *   tta        0, or 1 = run_pred's test-time augmentation (scripts/main_predict.py:147-149) as ONE batch: the encoder sees the 8
 *              flipped variants torch.flip(source, dims), dims in [(), (2,), (3,), (4,), (2,3), (2,4), (3,4), (2,3,4)], of every
 *              volume (flips are index arithmetic on the load; every variant gets the volume's un-flipped pad_mask, as the script
 *              passes it).  Every output below then holds 8*B volumes, variant-major (volume v*B + b), and workspace must be
 *              sized with mst_workspace_bytes(h, 8*B, ...).  mst_saliency(tta=1) un-flips and averages.
 *   logits     [B,out_ch] fp32 (nullable iff enable_linear == 0)
 *   feat       nullable [B,F] (without_linear, dino.py:164): F = emb for 'transformer'/'average', emb*D for 'linear',
 *              emb = embed_dim or embed_dim/4 behind the bottleneck
 *   enc_cls    nullable [B*D,embed_dim]: encoder output per slice (dino.py:131)
 *   plane_cls  nullable [B*D,enc_heads,NT], NT = 1 + num_registers + (H/14)*(W/14): row 0 of the LAST encoder block's
 *              attention -- the only part of attention_maps the getters read (dino.py:190-192)
 *   slice_cls  nullable [B,slice_heads,D+1]: row 0 of the slice attention (dino.py:174-175); 'transformer' only
 *   full_maps  nullable [depth,B*D,enc_heads,NT,NT] fp32: every block's full attention, as the reference's hook stores
 *              them (dino.py:241); only get_attention_cls needs them (mst_rollout) */
MST_API int mst_forward(mst_handle h, const void* src, int32_t src_dtype, int32_t B, int32_t D, int32_t H, int32_t W,
                const uint8_t* pad_mask, int32_t tta, float* logits, float* feat, float* enc_cls, float* plane_cls, float* slice_cls,
                float* full_maps, void* workspace, size_t workspace_bytes, void* stream);

/* Small batches are launch-bound (one volume: 77 kernels of ~10 us each): forwards of at most `max_tokens` token rows
 * (B * D * tokens per slice) are captured into a CUDA graph the second time mst_forward sees the exact same argument tuple
 * (pointers, shapes, stream) and replayed afterwards.  0 switches graphs off (the default).  The caller keeps the buffers of a
 * captured call alive and unchanged in address; results are identical to the eager path.  mst_graph_replays counts replays. */
MST_API int mst_set_graph_threshold(mst_handle h, int64_t max_tokens);
MST_API unsigned long long mst_graph_replays(mst_handle h);

/* The slice transformer + head on its own (MST-ResNet, reference mst/models/resnet.py:127-198, shares it): a handle created
 * with depth = 0 holds only cls_token, slice_fusion.* and linear.* (embed_dim = the backbone's feature width, e.g. 512 with 16 heads
 * for ResNet-34, resnet.py:152-170).  feats [B, D, embed_dim] fp32 = the backbone's per-slice features; scratch holds
 * B * (D + 1) * embed_dim floats; outputs as mst_forward's. */
MST_API int mst_slice_head_forward(mst_handle h, const float* feats, int32_t B, int32_t D, const uint8_t* pad_mask, float* logits, float* feat,
                           float* slice_cls, float* scratch, void* stream);

/* get_plane_attention / get_slice_attention / get_attention_maps (dino.py:173-202) and the caller's
 * head-mean + reshape + trilinear upsample (scripts/main_predict.py:73-74,100,161-162), batched.
 *   attn_maps  nullable [B*D,enc_heads,P] (get_attention_maps)   plane_attn nullable [B*D,enc_heads,P] (get_plane_attention)
 *   slice_attn nullable [B*D] (get_slice_attention)
 *   coarse     nullable [B,1,D,gh,gw] (required when full != NULL)   full nullable [B,1,D,H,W]
 *   skip_tokens  tokens in front of the patches in plane_cls: 1, or 5 with registers (dino.py:191)
 *   tta        1: plane_cls / slice_cls come from mst_forward(tta=1) (8*B volumes, variant-major); coarse and slice_attn are
 *              the un-flipped averages over the 8 variants in the script's summation order (main_predict.py:147-158), upsampled
 *              ONCE (:161-162); attn_maps / plane_attn must be NULL (they are per variant)
 *   h          nullable: the handle whose launch counter / profiler categories record the two kernels */
MST_API int mst_saliency(mst_handle h, const float* plane_cls, const float* slice_cls, int32_t B, int32_t D, int32_t enc_heads,
                 int32_t slice_heads, int32_t skip_tokens, int32_t gh, int32_t gw, int32_t H, int32_t W, int32_t tta,
                 float* attn_maps, float* plane_attn, float* slice_attn, float* coarse, float* full, void* stream);

/* get_attention_cls (dino.py:204-212), attention rollout: out = maps[0] @ maps[1] @ ... @ maps[depth-1] evaluated right to
 * left.  maps [depth,nmat,N,N] fp32 (mst_forward's full_maps with nmat = B*D*enc_heads); out, scratch [nmat,N,N]. */
MST_API int mst_rollout(const float* maps, int32_t depth, int32_t nmat, int32_t N, float* out, float* scratch, void* stream);

/* interpolate_pos_encoding (vision_transformer.py:179-211) for an H x W input: out [1 + (H/14)*(W/14), embed_dim] fp32,
 * row 0 the class position.  (mst_forward applies the same table internally; exposed for parity tests.) */
MST_API int mst_pos_embed(mst_handle h, int32_t H, int32_t W, float* out, void* stream);

/* np.quantile(x, q) per item, numpy's default 'linear' method (scripts/main_predict.py:243-245,296 on the upsampled saliency
 * volume).  data [items,n] fp32, q_dev [nq] fp64 on the device (nq <= 8), out [items,nq] fp64. */
MST_API int mst_quantile_workspace_bytes(int32_t items, int32_t nq, size_t* bytes);
MST_API int mst_quantile(const float* data, int64_t n, int32_t items, const double* q_dev, int32_t nq, double* out,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Input pipeline in front of mst_forward: DUKE_Dataset3D's default transform chain for one batch of equally shaped volumes
 * (mst/data/datasets/dataset_3d_duke.py:36-47 with image_resize / resample None and the random augmentations off):
 * tio.Flip(1) -> CropOrPad((W,H,D), padding_mode='minimum') (augmentations/augmentations_3d.py:144-195) ->
 * ZNormalization(percentiles, mask = (x > min) & (x < max)) (:41-86) -> ImageOrSubjectToTensor swapaxes(1,-1) (:23-29).
 *   src   [items, W0, H0, D0] fp32, or the raw int16 / uint16 voxels (2 bytes per voxel over PCIe; widened on the device into
 *         the workspace, which mst_prepare_volume_workspace_bytes sizes for it) (torchio's [C=1, W, H, D] per item)      out [items, 1, D, H, W] fp32 (the model's `source`)
 *   q_lo, q_hi  percentiles / 100 (0.005, 0.995), clamped before the statistics; flip_h = 1 for tio.Flip(1)
 *   stats nullable [items, 8] fp64: min, max, cutoff_lo, cutoff_hi, mean, std, masked voxels, status
 *         (status 0 ok; 1 std == 0 and 2 empty mask: the reference raises RuntimeError, augmentations_3d.py:75-84)
 * W*H*D must be a multiple of 4. */
MST_API int mst_prepare_volume_workspace_bytes(int32_t items, int32_t W0, int32_t H0, int32_t D0, size_t* bytes);
MST_API int mst_prepare_volume(mst_handle h /* nullable: instrumentation only */, const void* src, int32_t src_dtype /* MST_SRC_F32 / _I16 / _U16 */, int32_t items, int32_t W0, int32_t H0, int32_t D0, int32_t W, int32_t H,
                       int32_t D, int32_t flip_h, float q_lo, float q_hi, float* out, double* stats, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Training step on a FROZEN encoder (BASELINE.json config 5, `freeze=True`, reference dino.py:69-71; Lightning's step is
 * base_model.py:148-170: forward -> CrossEntropyLoss -> backward -> AdamW, base_model.py:103-110 / dino.py:41).  What trains is the
 * slice transformer + head of dino.py:84-103 (default construction: no bottleneck, no slice position embedding, no rotary).
 *   enc     [B, D, E] fp32: the encoder's per-slice output (mst_forward's enc_cls), no gradient
 *   params  HOST array of 17 DEVICE pointers, fp32, nn.Linear layout, in this order: cls_token, slice_fusion.layers.0.norm1.weight,
 *           .norm1.bias, .self_attn.in_proj_weight, .self_attn.in_proj_bias, .self_attn.out_proj.weight, .self_attn.out_proj.bias,
 *           .norm2.weight, .norm2.bias, .linear1.weight, .linear1.bias, .linear2.weight, .linear2.bias, slice_fusion.norm.weight,
 *           slice_fusion.norm.bias, linear.weight, linear.bias
 *   saved / factors  scratch of mst_slice_train_bytes: activations kept from forward to backward / per-volume backward factors
 *   logits  [B, C];  dlogits [B, C] = d loss / d logits (the caller's loss);  grads: 17 device pointers shaped like params, overwritten
 *   denc    nullable [B, D, E]: gradient w.r.t. the encoder outputs (what an un-frozen encoder's backward would consume)
 * mst_adamw: torch.optim.AdamW's update on one flat buffer (n floats): g is multiplied by grad_scale first (1/world after a summing
 * all-reduce), step counts from 1. */
MST_API int mst_slice_train_bytes(int32_t B, int32_t D, int32_t E, int32_t heads, int32_t C, size_t* saved_bytes, size_t* factor_bytes);
MST_API int mst_slice_train_forward(mst_handle h /* nullable */, const float* enc, const uint8_t* pad_mask, const float* const* params,
                            int32_t B, int32_t D, int32_t E, int32_t heads, int32_t C, float* saved, float* logits, void* stream);
MST_API int mst_slice_train_backward(mst_handle h /* nullable */, const float* enc, const float* dlogits, const float* const* params,
                             const float* saved, float* factors, float* const* grads, float* denc, int32_t B, int32_t D, int32_t E,
                             int32_t heads, int32_t C, void* stream);
MST_API int mst_adamw(mst_handle h /* nullable */, float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
              float beta2, float eps, float weight_decay, int32_t step, float grad_scale, void* stream);

/* Training step of the whole model (BASELINE.json config 5 as main_train.py:110-126 runs it: every parameter trainable): the encoder's
 * forward with every block's activations kept, and its backward pass.  bf16 activations and activation gradients, fp32 weight
 * gradients.  Supported: the vendored-factory construction (no LayerScale, no registers) at the position table's own patch grid.
 *   mst_train_forward   src as mst_forward; enc_cls [B*D, E] fp32 = the encoder output per slice (feeds mst_slice_train_forward)
 *   mst_set_grad        register (or clear, with NULL) the caller-owned fp32 gradient buffer of one encoder tensor by state_dict name
 *   mst_train_backward  denc [B*D, E] fp32 = d loss / d enc_cls (mst_slice_train_backward's denc); overwrites every registered gradient
 *                       buffer (all encoder.* tensors except mask_token must be registered); same workspace as the forward, untouched
 *                       in between.  The weights must not change between forward and backward. */
MST_API int mst_train_workspace_bytes(mst_handle h, int32_t B, int32_t D, int32_t H, int32_t W, size_t* bytes);
MST_API int mst_train_forward(mst_handle h, const void* src, int32_t src_dtype, int32_t B, int32_t D, int32_t H, int32_t W, float* enc_cls,
                      void* workspace, size_t workspace_bytes, void* stream);
MST_API int mst_set_grad(mst_handle h, const char* name, float* dev_fp32, int64_t numel);
MST_API int mst_train_backward(mst_handle h, const float* denc, int32_t B, int32_t D, int32_t H, int32_t W, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Instrumentation: kernels launched by this handle so far; per-category device time (CUDA events recorded on the
 * caller's stream around every launch between begin and end; end synchronises the device).  `ms`/`launches` must
 * hold at least 32 entries; mst_profile_categories() names them, comma separated, in order. */
MST_API unsigned long long mst_launch_count(mst_handle h);
MST_API const char* mst_profile_categories(void);
MST_API int mst_profile_begin(mst_handle h);
MST_API int mst_profile_end(mst_handle h, double* ms, int64_t* launches, int32_t n);

/* Kernel-level entry points (unit parity tests and micro-benchmarks call the same kernels the forward uses).
 * mode: 0 bias, 1 bias+GELU(erf), 2 bias+residual.  A [M,K], W [N,K] (nn.Linear layout), out/res [M,N]. */
MST_API int mst_kernel_gemm_bf16(const void* A, const void* W, int32_t M, int32_t N, int32_t K, int32_t mode,
                         const float* bias, const void* res, void* out, void* stream);
/* LayerNorm folded into the GEMM that consumes it (bf16 path; norm1 -> qkv, norm2 -> fc1; block.py:112-113):
 * A holds RAW rows; W rows are gamma-scaled and CENTRED, W[n,k] = bf16(gamma[k] W0[n,k] - mean_k(gamma W0[n,:])), so the
 * row mean cancels inside the MMA; bias[n] = b[n] + sum_k beta[k] W0[n,k]; rowstat [M] = rstd from
 * mst_kernel_row_stats_bf16:  out = act(rstd * acc + bias[n]). */
MST_API int mst_kernel_gemm_bf16_ln(const void* A, const void* W, int32_t M, int32_t N, int32_t K, int32_t gelu,
                            const float* bias, const float* rowstat, void* out, void* stream);
/* The weight packing behind mst_kernel_gemm_bf16_ln, as mst_finalize_weights applies it to qkv / fc1: W [N,K], b [N], gamma / beta [K]
 * fp32 -> Wd [N,K] bf16 (gamma-scaled, centred, rounded so that every row still sums to ~0) and bd [N] fp32. */
MST_API int mst_kernel_pack_linear_ln(const float* W, const float* b, const float* gamma, const float* beta, int32_t N, int32_t K,
                              void* Wd_bf16, float* bd, void* stream);
/* fc2 as the forward runs it between two blocks (block.py:113 -> :112 of the next block): x[M,N] += A[M,K] W[N,K]^T + bias in place,
 * and rowstat_out[M] = rstd of every UPDATED row (the next block's norm1 statistics) out of the same epilogue.  N = 384, K > 384. */
MST_API int mst_kernel_gemm_bf16_res_stats(const void* A, const void* W, int32_t M, int32_t N, int32_t K, const float* bias, void* x,
                                   float* rowstat_out, float eps, void* stream);
/* Encoder backward pieces (csrc/train_enc.cu), kernel level:
 *   gemm f32out   out[M,N] fp32 = A[M,K] W[N,K]^T (weight gradients dW = dY^T X on transposed activations; K > 384, N % 192 == 0)
 *   wgrad         dW [Nout, Kin] fp32 = dY[M, Nout]^T X[M, Kin] and db [Nout] fp32 (nullable) = column sums of dY, from the ROW-MAJOR
 *                 activations (MN-major tensor-core operands, no transposes); Nout % 128 == 0, Kin % 192 == 0; both outputs overwritten
 *   ln_bwd        dx = LayerNorm backward of dy at x (+ dres), dgamma / dbeta [E] fp32 (E = 384 / 768); synchronises the stream
 *   gelu          y = GELU(u) (y != NULL) and / or du = dy * GELU'(u) (dy, du != NULL); n % 8 == 0
 *   transpose     out [C, Mpad] = in [M, C]^T zero-padded to Mpad columns; colsum [C] fp32 (nullable) += column sums of `in`
 *   attention_bwd dqkv [BD*N, 3E] from qkv (q pre-scaled by 1/8), the forward output o and dO, both [BD*N, E]; d/dq is w.r.t. the
 *                 UN-scaled q projection (attention.py:58-60) */
MST_API int mst_kernel_gemm_bf16_f32out(const void* A, const void* W, int32_t M, int32_t N, int32_t K, float* out, void* stream);
MST_API int mst_kernel_wgrad_bf16(const void* dY, const void* X, int32_t M, int32_t Nout, int32_t Kin, float* dW, float* db, void* stream);
MST_API int mst_kernel_ln_bwd_bf16(const void* x, const void* dy, const void* dres, const float* gamma, void* dx, float* dgamma,
                           float* dbeta, int32_t rows, int32_t E, float eps, void* stream);
MST_API int mst_kernel_gelu_bf16(const void* u, void* y, const void* dy, void* du, int64_t n, void* stream);
MST_API int mst_kernel_transpose_bf16(const void* in, void* out, float* colsum, int32_t M, int32_t C, int32_t Mpad, void* stream);
MST_API int mst_kernel_attention_bwd_bf16(const void* qkv, const void* o, const void* dO, void* dqkv, int32_t BD, int32_t N, int32_t heads,
                                  void* stream);
/* the N = 257 forward kernel with the log2-domain row log-sum-exp lse [BD*heads, 257] kept (what the training forward runs), and the
 * backward pass that takes it instead of recomputing it */
MST_API int mst_kernel_attention_lse_bf16(const void* qkv, void* out, float* lse, int32_t BD, int32_t heads, void* stream);
MST_API int mst_kernel_attention_bwd_lse_bf16(const void* qkv, const void* o, const void* dO, const float* lse, void* dqkv, int32_t BD,
                                              int32_t N, int32_t heads, void* stream);
MST_API int mst_kernel_row_stats_bf16(const void* x, float* rowstat, int32_t rows, int32_t E, float eps, void* stream);
/* profiling aid: same as mst_kernel_gemm_bf16, plus cycle counters of CTA 0 (8 x int64: MMA warp wait-for-accumulator,
 * wait-for-operands, total, tiles; epilogue warp 0 wait-for-MMA, TMEM read, math+store) */
MST_API int mst_debug_gemm_timing(const void* A, const void* W, int32_t M, int32_t N, int32_t K, int32_t mode,
                          const float* bias, const void* res, void* out, long long* dbg_dev, void* stream);
MST_API int mst_kernel_gemm_f32(const float* A, const float* W, int32_t M, int32_t N, int32_t K, int32_t mode,
                        const float* bias, const float* res, float* out, void* stream);
/* qkv [BD*N, 3*heads*64] (q pre-scaled) -> out [BD*N, heads*64].  _bf16 picks the tcgen05 kernel for N == 257 and the
 * warp-MMA kernel otherwise; _bf16_warp_mma always runs the latter. */
MST_API int mst_kernel_attention_bf16(const void* qkv, void* out, int32_t BD, int32_t N, int32_t heads, void* stream);
MST_API int mst_kernel_attention_bf16_warp_mma(const void* qkv, void* out, int32_t BD, int32_t N, int32_t heads, void* stream);
MST_API int mst_kernel_attention_f32(const float* qkv, float* out, int32_t BD, int32_t N, int32_t heads, void* stream);
MST_API int mst_kernel_layernorm_bf16(const void* x, void* y, const float* gamma, const float* beta, int32_t rows, int32_t E,
                              float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MST_B200_H_ */
