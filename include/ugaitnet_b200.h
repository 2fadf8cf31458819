/*
 * ugaitnet_b200 -- C ABI of the B200 (sm_100a) hot path of UGaitNet.
 *
 * The reference (avagait/ugaitnet) is pure Python/Keras and has NO FFI / operator
 * boundary of its own; every entry point below therefore replaces a Keras layer call or
 * Lambda body of the reference graph, cited as  file:line  relative to the reference
 * root.  The reference-side binding (ctypes) a maintainer would add is shown in
 * INTEGRATION.md; the in-tree binding is ugaitnet_b200/_ffi.py.
 *
 * Conventions
 *  - extern "C", plain structs/pointers only.  Tensors cross as `ugn_tensor`, which is
 *    layout-identical to DLPack's `DLTensor` (dlpack.h, v0.8/v1.0): Python passes
 *    `&DLManagedTensor.dl_tensor` of a `tensor.__dlpack__()` capsule.  The library never
 *    takes ownership and never calls the DLPack deleter; the caller keeps the capsule
 *    alive until the stream has been synchronised.
 *  - All tensors must live on the ctx's CUDA device (device_type == kDLCUDA == 2), be
 *    dense row-major (strides NULL or compact) and have byte_offset folded by the caller
 *    or left in the struct (both honoured).
 *  - Every call returns 0 on success or a negative ugn_status; `ugn_last_error()` gives
 *    the thread-local message.  Nothing throws across the boundary.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*).
 *  - There is NO CPU fallback: a call on a non-CUDA tensor fails with UGN_ERR_DEVICE.
 *
 * Storage modes (selected by the dtype of the activation / weight operands):
 *  - f32 : "fp32 validation mode" -- SIMT FFMA kernels, NHWC f32 activations.
 *  - bf16 | f16: tensor-core mode -- tcgen05.mma (kind::f16, 16-bit in / FP32 accumulate
 *          in TMEM) fed by TMA.  A 16-bit operand carries a leading plane dimension P:
 *          P == 1 plain;  P == 2 "split" (plane 0 = hi = rn16(x), plane 1 = lo =
 *          rn16(x - hi)); with split operands the kernels issue hi*hi + hi*lo + lo*hi,
 *          i.e. ~fp32-accurate products on the 16-bit tensor cores at 3x the MMA work.
 *          All 16-bit operands of one call share one dtype: bf16 (kDLBfloat) or IEEE fp16
 *          (kDLFloat, 16 bits; 11-bit significand per plane).  Backward calls (dgrad, wgrad,
 *          linear_bwd) use min(P) over their operands, so a P == 1 gradient operand against
 *          P == 2 activations / weights runs ONE pass on the hi planes.
 *          16-bit GRADIENT operands (the dz written by ugn_conv2d_bwd_act / ugn_act_mask_bwd)
 *          are stored multiplied by the ctx's gradient scale s and every consumer multiplies
 *          its f32 result by 1/s (ugn_grad_scale_*; s == 1 unless set) -- fp16 range
 *          management; conversions to fp16 saturate at +-65504.
 *
 * Layouts
 *  - activations: NHWC, channel count padded to a multiple of 32 (64-byte rows) in bf16
 *    mode ("Cp"); f32 mode uses the same padded shapes so one host plan serves both.
 *  - conv weights: OHWI  [Cout][kh][kw][Cp_in]  (master copy f32 unpadded
 *    [Cout][kh][kw][Cin]); dense weights [out][in].
 */
#ifndef UGAITNET_B200_H_
#define UGAITNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UGN_ABI_VERSION 1

typedef enum {
  UGN_OK = 0,
  UGN_ERR_INVALID = -1,      /* bad shape / dtype / argument */
  UGN_ERR_CUDA = -2,         /* CUDA runtime or driver error */
  UGN_ERR_UNSUPPORTED = -3,  /* valid request this build does not implement */
  UGN_ERR_DEVICE = -4        /* tensor is not on the ctx device (no CPU fallback) */
} ugn_status;

/* DLPack-compatible view (== DLTensor). */
typedef struct {
  void* data;
  int32_t device_type; /* kDLCUDA = 2 */
  int32_t device_id;
  int32_t ndim;
  uint8_t dtype_code; /* kDLInt=0, kDLUInt=1, kDLFloat=2 (32 | 16 | 64 bits), kDLBfloat=4 */
  uint8_t dtype_bits;
  uint16_t dtype_lanes;
  int64_t* shape;
  int64_t* strides; /* NULL = compact row-major */
  uint64_t byte_offset;
} ugn_tensor;

typedef struct ugn_ctx ugn_ctx;

enum { UGN_ACT_LINEAR = 0, UGN_ACT_RELU = 1, UGN_ACT_LEAKY = 2 };
enum { UGN_MERGE_MAX = 0, UGN_MERGE_AVG = 1, UGN_MERGE_SIGNMAX = 2 };
/* OR-ed into the `merge` argument of ugn_fuse3_fwd / ugn_fuse3_bwd: gate + fusion only, no l2_normalize -- the
 * postriplet == 2 graph normalises AFTER the Dense layer "signature" (nets/mj_uwyhNets_ba.py:814-832). */
enum { UGN_FUSE3_NO_NORM = 0x100 };

int ugn_abi_version(void);
const char* ugn_last_error(void);
int ugn_ctx_create(int device, ugn_ctx** out);
int ugn_ctx_destroy(ugn_ctx* ctx);
/* cudaDeviceSynchronize + report of asynchronous kernel-side failures (tests / debugging). */
int ugn_ctx_check(ugn_ctx* ctx);
/* 1 if the device is compute capability 10.x (tcgen05/TMEM/TMA path usable). */
int ugn_ctx_has_tcgen05(ugn_ctx* ctx);

/* ---- a0: step input contract (data/mj_dataGeneratorMMUWYHsingle.py:664-823) ----------
 * x_nchw f32 [B,C,H,W] (what the Keras Input layers receive, nets/mj_uwyhNets_ba.py:1069-1074)
 * -> NHWC with C padded to Cp: f32 [B,H,W,Cp] or bf16 [P,B,H,W,Cp]; pad channels = 0. */
int ugn_pack_input(ugn_ctx*, const ugn_tensor* x_nchw, ugn_tensor* x_nhwc, void* stream);
/* The same pack with the generator's missing-modality expansion and mirror augmentation done ON THE
 * DEVICE (data/mj_dataGeneratorMMUWYHsingle.py:780-812, data/mj_augmentation.py:12-32), so only the base
 * rows cross PCIe: output row b reads x_base row src_row[b] (i32 [B], nullable = identity);
 * enable f32 [B] (nullable; the modality's use-flag): 0 -> the row's whole volume is the constant
 * `noise` (the reference's 1e-9); mirror u8 [B] (nullable): != 0 -> every channel flipped left-right
 * and even channels negated (literally what mj_mirrorsequence does, for every modality). */
int ugn_pack_input_expand(ugn_ctx*, const ugn_tensor* x_base, const ugn_tensor* src_row,
                          const ugn_tensor* enable, const ugn_tensor* mirror, float noise,
                          ugn_tensor* x_nhwc, void* stream);
/* The same pack with the integer-exact rest of the generator's augmentation (data/mj_dataGeneratorMMUWYHsingle.py:718-746,
 * data/mj_augmentation.py:35-50, __load_dd :318-321), each nullable:
 *   shift i8 [B,2] = (tx, ty) of the random transform: out[y][x] = in[clamp(y+tx)][clamp(x+ty)] (ImageDataGenerator
 *     .apply_transform with integer displacements, order-1 interpolation, fill_mode='nearest'), applied before the mirror;
 *   clip u8 [B] != 0: the optical-flow magnitude clip, on decoded values: |v| > clip_hi or |v| < clip_lo -> clip_val
 *     (raw thresholds 2300 / 50 and 1e-8 before the 1/compressFactor * 0.1 scaling: 2.3 / 0.05 / 1e-11). */
int ugn_pack_input_augment(ugn_ctx*, const ugn_tensor* x_base, const ugn_tensor* src_row, const ugn_tensor* enable,
                           const ugn_tensor* mirror, const ugn_tensor* shift, const ugn_tensor* clip, float clip_lo,
                           float clip_hi, float clip_val, float noise, ugn_tensor* x_nhwc, void* stream);


/* ---- a17 use3D: Conv3D branches (UWYHSemiNet.build_3Dbranch{,LReLU}, nets/mj_uwyhNets_ba.py:336-417) --------------
 * Strided 'valid' channels-last 3-D convolution, f32 only (fp32 validation engine: the option is outside the
 * benchmarked configurations).  x [B,T,H,W,C], w [Cout,kt,kh,kw,C], bias [Cout] (nullable), y / dz
 * [B,To,Ho,Wo,Cout] with To = (T-kt)/st+1 ...; act / alpha as ugn_conv2d_fwd (applied after the bias).
 * wgrad: dw like w (overwritten), db [Cout] nullable = column sums of dz.  dgrad: dx like x (overwritten). */
int ugn_conv3d_fwd(ugn_ctx*, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* bias, ugn_tensor* y,
                   int stride_t, int stride_h, int stride_w, int act, float alpha, void* stream);
int ugn_conv3d_wgrad(ugn_ctx*, const ugn_tensor* x, const ugn_tensor* dz, ugn_tensor* dw, ugn_tensor* db,
                     int stride_t, int stride_h, int stride_w, void* stream);
int ugn_conv3d_dgrad(ugn_ctx*, const ugn_tensor* dz, const ugn_tensor* w, ugn_tensor* dx,
                     int stride_t, int stride_h, int stride_w, void* stream);
/* On-disk sample values -> the f32 volume of the generator (`__load_dd`, data/mj_dataGeneratorMMUWYHsingle.py:313-329),
 * on the device: raw int16 (optical flow, compressFactor 100) or uint8 (gray / depth / silhouette), any shape; out f32 with
 * the same number of elements.  x = float(raw); |x| > clip_max or |x| < clip_min (each when > 0) -> 1e-8; then
 * x / divisor, * mul, - sub, each one IEEE f32 operation (bit-identical to the numpy statements): optical flow
 * (divisor 100, mul 0.1, sub 0), gray / depth (255, 1, 0.5), silhouette (255, 1, 0).  Only the stored integers cross PCIe. */
int ugn_decode_samples(ugn_ctx*, const ugn_tensor* raw, float divisor, float mul, float sub, float clip_min,
                       float clip_max, ugn_tensor* out, void* stream);
/* master f32 conv kernel [Cout][kh][kw][Cin] -> compute copy f32 [Cout][kh][kw][Cp] or
 * bf16 [P][Cout][kh][kw][Cp]; also used for dense weights with w viewed as [out][1][1][in]. */
int ugn_pack_weight(ugn_ctx*, const ugn_tensor* w_master, ugn_tensor* w_packed, void* stream);

/* ---- a1: Conv2D(valid, stride 1)+bias+act [+MaxPooling2D 2x2] -------------------------
 * replaces Conv2D/LeakyReLU/MaxPooling2D of UWYHNet.buildBranch{,LReLU}
 * (nets/mj_uwyhNets_ba.py:82-92, :125-137).
 * x [.,B,H,W,Cp_in]; w packed; bias f32 [Cout]; y [.,B,Hp,Wp,Cout] (pooled if pool!=0,
 * floor); pool_idx u8 [B,Hp,Wp,Cout] = argmax position 0..3 inside the 2x2 window
 * (ties -> first in (dy,dx) scan order), required iff pool != 0. */
int ugn_conv2d_fwd(ugn_ctx*, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* bias,
                   ugn_tensor* y, ugn_tensor* pool_idx, int act, float alpha, int pool,
                   void* stream);

/* backward through pool+act: dy f32 [B,Hp,Wp,C] (grad wrt layer output), y = layer output,
 * -> dz (grad wrt conv pre-activation) [.,B,Ho,Wo,C] in the storage mode of `dz`.
 * db f32 [C] (nullable, OVERWRITTEN): the Conv2D bias gradient sum_pixels dz, reduced from the f32
 * values in the same pass (then pass db = NULL to ugn_conv2d_wgrad). */
int ugn_conv2d_bwd_act(ugn_ctx*, const ugn_tensor* dy, const ugn_tensor* y,
                       const ugn_tensor* pool_idx, ugn_tensor* dz, ugn_tensor* db, int act,
                       float alpha, int pool, void* stream);

/* dx f32 [B,H,W,Cp_in] = full correlation of dz with w (Keras Conv2D input gradient). */
int ugn_conv2d_dgrad(ugn_ctx*, const ugn_tensor* dz, const ugn_tensor* w, ugn_tensor* dx,
                     void* stream);

/* dw f32 [Cout][kh][kw][Cin] (master layout, OVERWRITTEN), db f32 [Cout]. */
int ugn_conv2d_wgrad(ugn_ctx*, const ugn_tensor* x, const ugn_tensor* dz, ugn_tensor* dw,
                     ugn_tensor* db, void* stream);

/* Flatten() in the reference is over (C,H,W) (channels_first, :94): convert the last conv
 * output NHWC [.,B,H,W,C] <-> flat [.,B,C*H*W] (forward) and the f32 gradient back. */
int ugn_flatten_chw(ugn_ctx*, const ugn_tensor* y_nhwc, ugn_tensor* flat, void* stream);
int ugn_unflatten_chw(ugn_ctx*, const ugn_tensor* dflat_f32, ugn_tensor* dy_nhwc_f32, void* stream);

/* ---- a1/a5/a6: Dense (nets/mj_uwyhNets_ba.py:97-105, :1194-1214) ----------------------
 * y f32 [B,N] = act((x [.,B,K] . w[.,N,K]^T + bias)) * drop_mask (f32 [B,N], nullable,
 * already scaled by 1/(1-p): Keras inverted dropout).  y16 (nullable) receives the bf16
 * copy [P,B,N] for the next tensor-core consumer. */
int ugn_linear_fwd(ugn_ctx*, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* bias,
                   const ugn_tensor* drop_mask, ugn_tensor* y, ugn_tensor* y16, int act,
                   float alpha, void* stream);
/* Backward through dropout mask and activation of a Dense layer:
 * dz = dy * drop_mask * act'(y)  (y = the layer output; nullable for linear layers).
 * dz f32 (nullable) and/or dz16 bf16 [P,...] (nullable) receive the result. */
int ugn_act_mask_bwd(ugn_ctx*, const ugn_tensor* dy, const ugn_tensor* y, const ugn_tensor* drop_mask,
                     ugn_tensor* dz, ugn_tensor* dz16, int act, float alpha, void* stream);
/* dz [.,B,N] is the gradient wrt the pre-activation output (storage mode of x / w).
 * Outputs (each nullable): dx f32 [B,K], dw f32 [N,K] (overwritten), db f32 [N]. */
int ugn_linear_bwd(ugn_ctx*, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* dz,
                   ugn_tensor* dx, ugn_tensor* dw, ugn_tensor* db, void* stream);
/* ugn_linear_bwd with a FUSED input-gradient epilogue (tensor-core storage mode, B <= 128): the dx GEMM multiplies its
 * result by dx_mask f32 [B,K] (the inverted-dropout mask of the layer below, nullable), writes dx16 [P,B,K] = the 16-bit
 * gradient operand of that layer (still scaled by the ctx's gradient scale, so it feeds the next ugn_linear_bwd
 * directly) and dbx f32 [K] = its bias gradient (column sums of the masked, unscaled dx).  dx f32 [B,K] receives the UNMASKED
 * gradient (it doubles as the landing buffer of the split-K partial sums, so it is required).
 * Replaces dx GEMM + ugn_act_mask_bwd + the bias column-sum pass of nets/mj_uwyhNets_ba.py:97-105's backward. */
int ugn_linear_bwd_ex(ugn_ctx*, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* dz, ugn_tensor* dx,
                      const ugn_tensor* dx_mask, ugn_tensor* dx16, ugn_tensor* dbx, ugn_tensor* dw, ugn_tensor* db,
                      void* stream);

/* ---- Dropout without mask tensors (north_star: "the dropout mask fused into the epilogue"; Keras Dropout,
 * nets/mj_uwyhNets_ba.py:100) -- a counter-based generator (Philox4x32-10) keyed by rng i64 [2] = {seed, step} (device
 * memory), the layer id and the element index: keep with probability `keep`, scaled 1/keep (inverted dropout).  The
 * forward and backward passes regenerate the same bits; ugn_dropout_advance (rng[1] += 1, a kernel: CUDA-graph replays
 * draw fresh masks) is called once per step; ugn_dropout_mask materialises the mask of a layer (tests, layers that want a
 * tensor).  *_philox: ugn_linear_fwd / ugn_linear_bwd_ex with the mask generated in the dense post pass (tensor-core
 * storage mode; in ugn_linear_bwd_philox the mask is that of the layer BELOW, applied to dx16 / dbx). */
int ugn_dropout_advance(ugn_ctx*, ugn_tensor* rng, void* stream);
int ugn_dropout_mask(ugn_ctx*, const ugn_tensor* rng, int layer, float keep, ugn_tensor* out, void* stream);
int ugn_linear_fwd_philox(ugn_ctx*, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* bias,
                          const ugn_tensor* rng, int layer, float keep, ugn_tensor* y, ugn_tensor* y16, int act,
                          float alpha, void* stream);
int ugn_linear_bwd_philox(ugn_ctx*, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* dz, ugn_tensor* dx,
                          const ugn_tensor* rng, int layer, float keep, ugn_tensor* dx16, ugn_tensor* dbx,
                          ugn_tensor* dw, ugn_tensor* db, void* stream);



/* ---- a2+a3+a4: gate x use-flag, fusion, l2_normalize ---------------------------------
 * replaces mj_tensor_times_scalar (:51-54), fMerge(name="fusion") (:1189; sign_max at
 * mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:169-178) and tf.math.l2_normalize (:1191).
 * br[m] f32 [B,d], flags[m] f32 [B,1] (m < nmods <= 4).  Outputs: sig f32 [B,d],
 * sig16 bf16 [P,B,d] (nullable), winner u8 [B,d] (selected modality; for AVG unused),
 * inv_norm f32 [B,2] = {1/norm, sum x^2}.  normalize == 0 skips l2_normalize (1-modality graph, :904). */
int ugn_fuse_fwd(ugn_ctx*, int nmods, const ugn_tensor* const* br, const ugn_tensor* const* flags,
                 ugn_tensor* sig, ugn_tensor* sig16, ugn_tensor* winner, ugn_tensor* inv_norm,
                 int merge, int normalize, void* stream);
/* a2+a3+a4+a5 in ONE kernel (north_star: "gated sign_max fusion plus FC1 is a single kernel"): ugn_fuse_fwd followed by
 * the FC1 layer "code" = act(sig . code_w^T + code_b) (nets/mj_uwyhNets_ba.py:1194-1203) computed from the normalised
 * row while it is still in shared memory, and dropcode = code * drop_mask (Dropout "dropcode", mask already scaled by
 * 1/(1-p); drop_mask nullable: dropcode = code; dropcode nullable).  code_w f32 [nc,d], code_b f32 [nc] (nullable),
 * code / dropcode / drop_mask f32 [B,nc]; d % 4 == 0.  A modality whose use-flag is 0 is not read. */
int ugn_fuse_fc1_fwd(ugn_ctx*, int nmods, const ugn_tensor* const* br, const ugn_tensor* const* flags,
                     ugn_tensor* sig, ugn_tensor* sig16, ugn_tensor* winner, ugn_tensor* inv_norm, int merge,
                     int normalize, const ugn_tensor* code_w, const ugn_tensor* code_b, ugn_tensor* code,
                     const ugn_tensor* drop_mask, ugn_tensor* dropcode, int act, float alpha, void* stream);
/* dsig f32 [B,d] -> dbr[m] f32 [B,d] (gradient wrt each branch output, gate applied). */
int ugn_fuse_bwd(ugn_ctx*, int nmods, const ugn_tensor* dsig, const ugn_tensor* sig,
                 const ugn_tensor* winner, const ugn_tensor* inv_norm,
                 const ugn_tensor* const* flags, ugn_tensor* const* dbr, int merge,
                 int normalize, void* stream);

/* ---- a6: softmax + categorical cross-entropy (:1214, :1243) ---------------------------
 * logits f32 [B,C], labels i32 [B].  loss_acc f32 [2] = {mean CE, accuracy};
 * dlogits f32 [B,C] = scale * (softmax - onehot)/B (nullable). */
int ugn_softmax_ce(ugn_ctx*, const ugn_tensor* logits, const ugn_tensor* labels,
                   ugn_tensor* loss_acc, ugn_tensor* dlogits, float scale, void* stream);
/* the same with tf.losses.CategoricalCrossentropy(label_smoothing=e) (`smoothlabels`, :1252-1262):
 * targets y*(1-e) + e/C; accuracy is still measured against the hard label. */
int ugn_softmax_ce_ls(ugn_ctx*, const ugn_tensor* logits, const ugn_tensor* labels,
                      ugn_tensor* loss_acc, ugn_tensor* dlogits, float scale, float label_smoothing,
                      void* stream);

/* ---- a7: batch-all triplet loss (nets/triplet_loss_all.py:8-77) -----------------------
 * emb f32 [n,B,d] (or [B,d] == n = 1), labels i32 [B].  out f32 [2] = {loss, total count of
 * active triplets}; demb f32 like emb (nullable) = scale * dLoss/dEmb.
 * workspace: ugn_triplet_workspace_bytes(n,B) bytes, f32-aligned device buffer. */
int64_t ugn_triplet_workspace_bytes(int n, int B);
int ugn_triplet_all(ugn_ctx*, const ugn_tensor* emb, const ugn_tensor* labels, float margin,
                    float scale, ugn_tensor* out, ugn_tensor* demb, ugn_tensor* workspace,
                    void* stream);
/* The same with the B x B Gram matrix as a tensor-core GEMM (north_star; nets/triplet_loss_all.py:70-77): emb16 = the
 * 16-bit hi/lo planes [2,B,d] of emb (ugn_fuse_fwd writes them next to the f32 signature; d % 64 == 0, one part).
 * tcgen05, three passes (~fp32 products), fp32 accumulation, no split-K: one accumulation order per element, identical
 * rows keep an exactly zero distance.  Hinge, count and the analytic backward are shared with ugn_triplet_all. */
int ugn_triplet_all_tc(ugn_ctx*, const ugn_tensor* emb, const ugn_tensor* emb16, const ugn_tensor* labels, float margin,
                       float scale, ugn_tensor* out, ugn_tensor* demb, ugn_tensor* workspace, void* stream);
/* Batch-HARD triplet loss: tfa.losses.TripletHardLoss(margin) as compiled by UWYHSemiNet3Mods.compile_hard
 * (nets/mj_uwyhNets_ba.py:1302-1306; soft = False, L2 distances): mean over the anchors of
 * max(farthest positive - nearest negative + margin, 0), tfa's masked_maximum / masked_minimum semantics (anchor without
 * positives: 0; without negatives: the row maximum), gradient split evenly among tied entries.  emb f32 [B,d] (tfa takes
 * rank-2 embeddings), emb16 nullable (given: tensor-core Gram as ugn_triplet_all_tc).  out f32 [2] = {loss, anchors
 * with a positive term}; demb, workspace as ugn_triplet_all (n = 1). */
int ugn_triplet_hard(ugn_ctx*, const ugn_tensor* emb, const ugn_tensor* emb16, const ugn_tensor* labels, float margin,
                     float scale, ugn_tensor* out, ugn_tensor* demb, ugn_tensor* workspace, void* stream);
/* Pair verification loss of the Siamese builder UWYHNet.build (nets/mj_uwyhNets_ba.py:154-245 -> VerifLossLayer,
 * nets/mj_loss.py:65-95): emb f32 [2B,d] (rows [0,B) = first element of every pair, [B,2B) = second), labels i32 [>= B]
 * (1 = same identity, 0 = different; other values: pair ignored).  loss = 0.5 * sum_{pos} (a-b)^2 + 0.5 * max(0, margin -
 * sqrt(sum over ALL negative rows of (a-b)^2))^2.  out f32 [2] = {loss, 1 if the negative hinge is active}; demb f32
 * [2B,d] (nullable) = scale * dLoss/dEmb; workspace >= 16 bytes (8-byte aligned). */
int ugn_pair_verif_loss(ugn_ctx*, const ugn_tensor* emb, const ugn_tensor* labels, float margin, float scale,
                        ugn_tensor* out, ugn_tensor* demb, ugn_tensor* workspace, void* stream);


/* ---- a8/a9: regulariser + optimiser --------------------------------------------------
 * One fused multi-tensor step over a FLAT parameter arena: for segment s covering
 * [off[s], off[s+1]) elements, g' = g*gscale + 2*l2[s]*w (Keras L2 regulariser gradient),
 * Keras Adam (mains/mj_trainUWYHGaitNet_DataGen_3mods.py:242): m,v update,
 * w -= lr_t * m / (sqrt(v)+eps).  seg_off i64 [S+1], seg_l2 f32 [S] device tensors.
 * reg_out f32 [1] (nullable) accumulates sum_s l2[s]*|w_s|^2 of the PRE-update weights.
 * lr_dev f32 [1] (nullable): when given, the learning rate is read from device memory
 * instead of lr_t, so a captured CUDA graph can be replayed with a new rate.
 * pack_table i64 [S,2] (nullable, device): {address, numel} of the 16-bit compute copy
 * [pack_planes][numel] of segment s (address 0 = none): the updated weights are re-split into it
 * in the same pass (pack_f16: 0 bf16, 1 fp16), replacing a separate ugn_pack_weight for
 * unpadded (dense) weights. */
int ugn_adam_step(ugn_ctx*, ugn_tensor* w, const ugn_tensor* g, ugn_tensor* m, ugn_tensor* v,
                  const ugn_tensor* seg_off, const ugn_tensor* seg_l2, float lr_t, float beta1,
                  float beta2, float eps, float gscale, ugn_tensor* reg_out,
                  const ugn_tensor* lr_dev, const ugn_tensor* pack_table, int pack_planes,
                  int pack_f16, void* stream);
/* The two Adam variants the reference mains select (mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:232-236):
 * vhat f32 arena (nullable): optimizers.Adam(amsgrad=True) -- vhat = max(vhat, v), w -= lr_t*m/(sqrt(vhat)+eps);
 * weight_decay > 0: tfa.optimizers.AdamW -- decoupled decay w -= weight_decay*w (not scaled by the learning
 * rate) applied together with the Adam update.  All other arguments as ugn_adam_step. */
int ugn_adam_step_ex(ugn_ctx*, ugn_tensor* w, const ugn_tensor* g, ugn_tensor* m, ugn_tensor* v,
                     ugn_tensor* vhat, float weight_decay, const ugn_tensor* seg_off,
                     const ugn_tensor* seg_l2, float lr_t, float beta1, float beta2, float eps, float gscale,
                     ugn_tensor* reg_out, const ugn_tensor* lr_dev, const ugn_tensor* pack_table,
                     int pack_planes, int pack_f16, void* stream);
/* ---- a10: data-parallel step (MirroredStrategy, mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:342-349) with the
 * gradient exchange FUSED into the optimiser over NVLink peer memory -- one kernel instead of all-reduce + optimiser:
 * rank r owns the r-th of `world` equal slices of the flat arena; for its slice it sums the gradient arenas of all
 * ranks through peer loads (reduce-scatter), applies the regulariser + Adam / SGD update with ITS slice of the
 * optimiser state (m, v, vhat are only ever touched inside the owner's slice, so the state is effectively sharded
 * and the optimiser's HBM traffic drops by 1/world), and stores the updated weights into every rank's weight arena
 * (all-gather).  g_peers / w_peers: HOST arrays [world] of device addresses of the ranks' gradient / weight arenas
 * mapped into this process (e.g. torch.distributed._symmetric_memory buffer_ptrs); w_peers[rank] must be `w`.
 * g_multicast / w_multicast (0 = none): NVSwitch multicast addresses of the two arenas; when both are given the
 * sum is formed inside the switch (multimem.ld_reduce) and the weights are broadcast by it (multimem.st), which
 * halves the NVLink traffic of the unicast path.
 * The caller brackets the call with two group barriers (all gradients written before; all weights written after)
 * and refreshes its 16-bit compute copies afterwards (ugn_pack_weight).  opt: 0 Adam family (vhat / weight_decay as
 * ugn_adam_step_ex), 1 SGD with momentum (beta1 = momentum).  The mean (1/world) is applied inside; reg_out receives
 * this rank's slice of the regulariser value -- or, when reg_peers (HOST array [world] of the device addresses of every
 * rank's reg_out scalar, peer-mapped; nullable) is given, every rank's scalar receives the sum over all slices through
 * system-scope atomic adds (no all-reduce on the step's critical path); the caller then zeroes its scalar BEFORE the
 * first barrier.  lr_dev (required): the step's learning rate in device memory.
 * stage (f32 [world * ceil(n/4/world)*4], nullable) + staged_ranges (HOST array of n_ranges <= 4 element ranges [lo, hi),
 * multiples of 4): gradients of those arena ranges were PUSHED beforehand -- every rank p copied its part of this rank's
 * slice to stage[p * slice_len + (e - slice_start)] (copy engines over NVLink, overlapped with the rest of the backward
 * pass) -- so the kernel sums them from local memory instead of pulling them from the peers.
 * pack_table (as ugn_adam_step, nullable) + cw_peers (HOST array [world] of the base addresses of every rank's
 * symmetric arena of 16-bit compute copies; the table's addresses point into this rank's arena) [+ cw_multicast]:
 * for the segments in the table the owner writes the hi / lo planes of its updated weights into EVERY rank's compute
 * copy and skips the f32 broadcast -- the caller needs no re-split pass afterwards, and the other ranks' f32 masters
 * of those segments are stale until refreshed (UGaitEngine.sync_master_weights).
 * cw_multicast == UGN_CW_DEFERRED: the kernel refreshes only THIS rank's copies of its slice; the caller sends those
 * plane ranges to the peers itself (copy engines, on a side stream underneath the next step's convolution forward) and
 * orders the next reader of the copies behind that transfer. */
#define UGN_CW_DEFERRED ((int64_t)-1)
int ugn_dp_optim_step(ugn_ctx*, int opt, int world, int rank, const int64_t* g_peers, const int64_t* w_peers,
                      int64_t g_multicast, int64_t w_multicast, ugn_tensor* w, const ugn_tensor* g, ugn_tensor* m, ugn_tensor* v, ugn_tensor* vhat,
                      float weight_decay, const ugn_tensor* seg_off, const ugn_tensor* seg_l2, float beta1,
                      float beta2, float eps, ugn_tensor* reg_out, const int64_t* reg_peers, const ugn_tensor* lr_dev,
                      const ugn_tensor* stage, const int64_t* staged_ranges, int n_ranges,
                      const ugn_tensor* pack_table, int pack_planes, int pack_f16, const int64_t* cw_peers,
                      int64_t cw_multicast, void* stream);
/* SGD with momentum (optimizers.SGD(lr, momentum, decay), :245): v = mom*v - lr*g'; w += v.  Keras' `decay`
 * is a learning-rate schedule, lr / (1 + decay*iterations): the caller passes the scheduled rate. */
int ugn_sgd_step(ugn_ctx*, ugn_tensor* w, const ugn_tensor* g, ugn_tensor* v,
                 const ugn_tensor* seg_off, const ugn_tensor* seg_l2, float lr, float momentum,
                 float gscale, ugn_tensor* reg_out, const ugn_tensor* lr_dev,
                 const ugn_tensor* pack_table, int pack_planes, int pack_f16, void* stream);

/* ---- a12: brute-force k-NN (mains/mj_testUWYHGaitNet_open_tum.py:331-341) -------------
 * queries f32 [Q,D], gallery f32 [N,D] (this rank's shard), k <= 32.
 * Stage 1 (ugn_knn_topk): candidate search, kc >= k candidates per query by approximate
 *   (fp32 / tensor-core) squared distance, fused top-k (the QxN matrix is never stored),
 *   then exact fp64 re-rank of the candidates by sum((q-g)^2) with ordering
 *   (distance, global index).  out_d2 f64 [Q,k], out_idx i64 [Q,k] (+ idx_base),
 *   out_lab i32 [Q,k] (gallery_labels i32 [N]).
 * Stage 2 (ugn_knn_merge_vote): merges G shards' [G,Q,k] candidate lists with the same
 *   ordering and votes (uniform weights, ties -> smallest label). pred i32 [Q]. */
int64_t ugn_knn_workspace_bytes(int64_t Q, int64_t N, int64_t D, int k);
/* clf.fit(): per-row squared norms of the gallery shard.  g2 f32 [Npad >= N]: rows >= N are set to
 * +inf (padding for the tensor-core scan, which wants Npad = roundup(N,256)); gmax2 f32 [1]
 * (nullable) receives max_row |g|^2. */
int ugn_knn_gallery_norms(ugn_ctx*, const ugn_tensor* gallery, ugn_tensor* g2, ugn_tensor* gmax2,
                          void* stream);
/* SIMT (fp32 FFMA) candidate scan + exact re-rank. */
int ugn_knn_topk(ugn_ctx*, const ugn_tensor* queries, const ugn_tensor* gallery, const ugn_tensor* g2,
                 const ugn_tensor* gallery_labels, int k, int64_t idx_base, ugn_tensor* out_d2,
                 ugn_tensor* out_idx, ugn_tensor* out_lab, ugn_tensor* workspace, void* stream);
/* f32 [rows,D] -> operand of the tensor-core scan: fp16 hi/lo planes, K-chunk major
 * x16 f16 [2, ceil(D/64), rows, 64] (zero padded), so that every TMA box is one contiguous 16 KB run. */
int ugn_knn_pack(ugn_ctx*, const ugn_tensor* x, ugn_tensor* x16, void* stream);
/* Tensor-core candidate scan: the distance GEMM Q.G^T on tcgen05 (q16 / g16 from ugn_knn_pack) with
 * the top-KC filter fused into the GEMM epilogue, then the same exact fp64 re-rank.  The re-rank PROVES per query that the true top-k lie
 * inside the candidate set (error bound of the split-fp16 GEMM vs the KC-th candidate score); queries
 * that fail the proof are flagged (flags i32 [Q], output) and recomputed by brute force in fp64, so
 * indices / labels are bit-exact in every case. */
int ugn_knn_topk_tc(ugn_ctx*, const ugn_tensor* queries, const ugn_tensor* q16,
                    const ugn_tensor* gallery, const ugn_tensor* g16, const ugn_tensor* g2,
                    const ugn_tensor* gmax2, const ugn_tensor* gallery_labels, int k, int64_t idx_base,
                    ugn_tensor* out_d2, ugn_tensor* out_idx, ugn_tensor* out_lab, ugn_tensor* flags,
                    ugn_tensor* workspace, void* stream);
int ugn_knn_merge_vote(ugn_ctx*, const ugn_tensor* d2, const ugn_tensor* idx,
                       const ugn_tensor* lab, int k, ugn_tensor* out_d2, ugn_tensor* out_idx,
                       ugn_tensor* out_lab, ugn_tensor* pred, void* stream);
/* The same merge on the buffer ONE all-gather of the sharded search produces: packed u8 [G, row] (row >= 20*Q*k, a
 * multiple of 8), row g = rank g's lists back to back: d2 f64 [Q,k] | idx i64 [Q,k] | lab i32 [Q,k]. */
int ugn_knn_merge_vote_packed(ugn_ctx*, const ugn_tensor* packed, long long Q, int k, ugn_tensor* out_d2,
                              ugn_tensor* out_idx, ugn_tensor* out_lab, ugn_tensor* pred, void* stream);


/* ---- a13: video-level summaries of the open-world test -----------------------------------
 * (mains/mj_testUWYHGaitNet_open_tum.py:355-420).  Rows are grouped by video through a CSR built
 * on the host the way the reference does (np.unique / np.where): order i32 [N] = row indices sorted
 * by video (stable), offsets i32 [V+1].
 * ugn_segment_pool: out f32 [V,D] = per-video mean (use_avg != 0) or max of codes f32 [N,D].
 * ugn_segment_mode: out i32 [V] = statistics.mode of labels i32 [N] per video: most frequent label,
 *   ties -> first encountered (Python >= 3.8); legacy_ties != 0 -> on a tie the video's first label
 *   (the reference's `except:` branch under Python < 3.8, where mode() raised on ties). */
int ugn_segment_pool(ugn_ctx*, const ugn_tensor* codes, const ugn_tensor* order,
                     const ugn_tensor* offsets, int use_avg, ugn_tensor* out, void* stream);
int ugn_segment_mode(ugn_ctx*, const ugn_tensor* labels, const ugn_tensor* order,
                     const ugn_tensor* offsets, int legacy_ties, ugn_tensor* out, void* stream);

/* ---- a16: GaitSet branch type (nets/mj_uwyhNets_ba.py:420-484, MatMul :23-48) --------------
 * The convolutions of build_gaitset_branch run on ugn_conv2d_* above: padding='same' (:428-466) is
 * realised with ZERO-BORDERED activation buffers (the caller allocates [.,N,H+2,W+2,C] zero-filled
 * once, producers fill the interior with ugn_pad_hw, gradients come back through ugn_crop_hw), the
 * bias argument is NULL (use_bias=False) and act = UGN_ACT_LEAKY, alpha 0.3 (layers.LeakyReLU()).
 *
 * Split layout: the tensor-core conv kernels keep an input row in at most 64 pixel slots; a zero-bordered
 * 66-wide row does not fit, so the 64-wide layers are kept as S = 2 overlapping column halves:
 * padded input [.,N*S,H+2,W/S+2,C] (half h = padded columns [h*W/S, h*W/S + W/S + 2)), valid-conv output
 * [.,N*S,H,W/S,C].  ugn_pad_hw / ugn_crop_hw convert between split and plain images (S = ratio of the
 * leading dimensions).
 *
 * ugn_gs_conv1_fwd: TimeDistributed(ZeroPadding2D(2)) + Conv2D(32, 5x5, 'same', no bias) + LeakyReLU
 *   (:427-429) fused, straight from the Keras input x f32 [B,T,H,W,c] (c = 1 | 2;
 *   data/mj_dataGeneratorMMUWYHsingle_repetitions.py:426-434) into the interior of the zero-bordered
 *   (split) input of the next convolution, yp [.,B*T*S,H+6,(W+4)/S+2,32].  w f32 [32,1,1,25c], tap-major
 *   (ky,kx,ci).  K = 25c is too small for the tensor cores: fp32 FFMA, exact products.
 * ugn_gs_conv1_wgrad: its kernel gradient with the LeakyReLU derivative and the crop folded in:
 *   dxp f32 (shape of yp) = gradient wrt the layer output on the padded (split) frame, as written by
 *   ugn_conv2d_dgrad of the next layer; dw f32 [32,1,1,25c], OVERWRITTEN. */
int ugn_gs_conv1_fwd(ugn_ctx*, const ugn_tensor* x, const ugn_tensor* w, ugn_tensor* yp, float alpha,
                     void* stream);
int ugn_gs_conv1_wgrad(ugn_ctx*, const ugn_tensor* x, const ugn_tensor* dxp, const ugn_tensor* yp,
                       ugn_tensor* dw, float alpha, void* stream);
/* interior copy src [.,N*S,H,W,C] -> dst [.,N,H+2p,W*S+2p,C] (any storage mode, p and S from the
 * shapes; the border of dst is left untouched) and its adjoint on f32 gradients:
 * dst [N*S,H,W,C] (+)= interior of src [N,H+2p,W*S+2p,C]. */
int ugn_pad_hw(ugn_ctx*, const ugn_tensor* src, ugn_tensor* dst, void* stream);
int ugn_crop_hw(ugn_ctx*, const ugn_tensor* src, ugn_tensor* dst, int accumulate, void* stream);
/* Set pooling Lambda(reduce_max(x, axis=1)) over the T frames of a sequence (:435,:454,:465) fused
 * with the layers.Add() that follows (:455,:466): a [.,B*T,H,W,C] -> m f32 [B,H,W,C] = max_t a
 * (nullable), y [.,B,H,W,C] = m + addend (nullable; addend [.,B,H,W,C] nullable, any storage mode).
 * Backward: da f32 [B*T,H,W,C] (+)= dm * [a == m] / #{t: a == m}  (tf reduce_max splits the gradient
 * evenly among ties). */
int ugn_setmax_fwd(ugn_ctx*, const ugn_tensor* a, int T, const ugn_tensor* addend, ugn_tensor* m,
                   ugn_tensor* y, void* stream);
int ugn_setmax_bwd(ugn_ctx*, const ugn_tensor* dm, const ugn_tensor* a, const ugn_tensor* m, int T,
                   ugn_tensor* da, int accumulate, void* stream);
/* Horizontal pyramid pooling (:468-479): x f32 [B,H,W,C] -> rows of feat f32 [62,B,C]; for
 * nb in {1,2,4,8,16} the H*W positions (row-major) are cut into nb strips, feature = mean + max;
 * part = 2*(nb-1) + which*nb + strip, which = 0 for the set-level map, 1 for the global map (the
 * reference's concatenation order, already transposed to [62,B,C], :481-482).
 * Backward: dx (+)= dfeat/len + dfeat * [x == strip max] / #ties, summed over the 5 levels. */
int ugn_hpp_fwd(ugn_ctx*, const ugn_tensor* x, int which, ugn_tensor* feat, void* stream);
int ugn_hpp_bwd(ugn_ctx*, const ugn_tensor* dfeat, const ugn_tensor* x, int which, ugn_tensor* dx,
                int accumulate, void* stream);
/* MatMul layer (:41-46) and its gradients: C f32 [n,M,N] = op(A)[n] . op(B)[n]  (fp32 FFMA);
 * a_t == 0: A [n,M,K], 1: A [n,K,M];  b_t == 0: B [n,K,N], 1: B [n,N,K]. */
int ugn_bmm_f32(ugn_ctx*, const ugn_tensor* A, int a_t, const ugn_tensor* B, int b_t, ugn_tensor* C,
                void* stream);
/* Gate x use-flag (:51-54), fusion (:1189) and tf.math.l2_normalize(axis=1) (:1191) on the GaitSet
 * layout: br[m] f32 [n,B,d], flags[m] f32 [B,1].  axis 1 of [n,B,d] is the BATCH axis: every
 * (part, feature) column is normalised over the rows of the batch -- the reference's literal
 * behaviour.  sig f32 [n,B,d], winner u8 [n,B,d], col_norm f32 [n,d,2] = {1/norm, sum x^2}.
 * merge = UGN_MERGE_* [| UGN_FUSE3_NO_NORM: sig is the un-normalised fusion, col_norm = {1, sum x^2}]. */
int ugn_fuse3_fwd(ugn_ctx*, int nmods, const ugn_tensor* const* br, const ugn_tensor* const* flags,
                  ugn_tensor* sig, ugn_tensor* winner, ugn_tensor* col_norm, int merge, void* stream);
int ugn_fuse3_bwd(ugn_ctx*, int nmods, const ugn_tensor* dsig, const ugn_tensor* sig,
                  const ugn_tensor* winner, const ugn_tensor* col_norm, const ugn_tensor* const* flags,
                  ugn_tensor* const* dbr, int merge, void* stream);
/* Lambda(tf.transpose(x, [1,0,2])) in front of Flatten + "classprob" (:1211-1213):
 * src f32 [A,B,d] -> dst f32 [B,A,d] (or its flattened view [B,A*d]). */
int ugn_permute102(ugn_ctx*, const ugn_tensor* src, ugn_tensor* dst, void* stream);

/* ---- generic tensor-core GEMM (building block exposed for tests / k-NN / triplet) ----
 * C f32 [M,N] (+)= A . B^T with bf16 (or f16) operands [P,rows,cols]:
 *   a_mn == 0: A is [P,M,K] (K contiguous);  a_mn == 1: A is [P,K,M] (M contiguous).
 *   b_mn == 0: B is [P,N,K];                 b_mn == 1: B is [P,K,N].
 * accumulate != 0 adds into C (split-K partials use red.global.add). */
int ugn_gemm_bf16(ugn_ctx*, const ugn_tensor* A, int a_mn, const ugn_tensor* B, int b_mn,
                  ugn_tensor* C, int accumulate, void* stream);
/* f32 -> 16-bit [P,...] split helper (P and bf16 | f16 taken from dst). */
int ugn_split_bf16(ugn_ctx*, const ugn_tensor* src_f32, ugn_tensor* dst_bf16, void* stream);

/* ---- gradient scale of the ctx (16-bit gradient operands, see "Storage modes") ---------
 * update: s = 2^floor(log2(target / max|ref|)) computed ON DEVICE from a reference gradient of
 * the step (f32, any shape; the engine passes dL/dsignature), so it is CUDA-graph capturable.
 * set: fixed value (1 = off).  Both are stream-ordered. */
int ugn_grad_scale_update(ugn_ctx*, const ugn_tensor* ref, float target, void* stream);
int ugn_grad_scale_set(ugn_ctx*, float scale, void* stream);

/* out f32 [cols] = column sums of x f32 [rows, cols]: the bias gradient of a Dense layer taken from the f32 output
 * gradient BEFORE it is rounded to a 16-bit GEMM operand.  Where the rows of dy cancel (sum_b dL/dsignature_b = 0 for the
 * translation-invariant triplet loss on an un-normalised signature, nets/mj_uwyhNets_ba.py:900-915), the sum of the
 * rounded operand carries the rounding noise of every row against a near-zero total. */
int ugn_colsum(ugn_ctx*, const ugn_tensor* x, ugn_tensor* out, void* stream);

/* MMA passes per product of the FORWARD convolution / dense layers when their operands are hi/lo split
 * 16-bit planes (host-side ctx state, not stream-ordered; applies to ugn_conv2d_fwd / ugn_linear_fwd):
 *   0 or 3: hi*hi + hi*lo + lo*hi (~fp32 products)      1: hi*hi only (11-bit operands, "TF32 class")
 *   2: hi*hi + hi*lo(weights): activations at 11 bits, weights at 22 -- the activation lo plane is not even loaded
 *   4: hi*hi + lo(activations)*hi: weights at 11 bits -- the weight lo plane is not loaded (weight-streaming dense layers)
 * Replaces nothing in the reference (Keras computes in fp32, nets/mj_uwyhNets_ba.py:82-105); it is the precision knob
 * of the tensor-core restatement, gated by tests/test_decisions_gpu.py. */
int ugn_set_fwd_passes(ugn_ctx*, int conv_passes, int dense_passes);

/* number of kernels this library has launched on this ctx since creation (bench.py's
 * gpu_launches claim is read from here). */
int64_t ugn_launch_count(ugn_ctx*);

#ifdef __cplusplus
}
#endif
#endif /* UGAITNET_B200_H_ */
