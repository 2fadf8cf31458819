// C-ABI entry points: validation + dispatch on storage mode (f32 -> SIMT validation kernels,
// bf16 -> tcgen05/TMA kernels).  No CPU fallback anywhere.
#include "common.cuh"
#include "tc.cuh"

static thread_local std::string g_last_error;

void ugn_set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

int ugn_validate(const ugn_ctx* ctx, const ugn_tensor* t, const char* name, UgnDType dt, int lo,
                 int hi) {
  if (!t) UGN_FAIL(UGN_ERR_INVALID, "%s: null tensor", name);
  if (t->device_type != UGN_DL_CUDA || t->device_id != ctx->device)
    UGN_FAIL(UGN_ERR_DEVICE, "%s: tensor is on device (%d,%d), expected CUDA device %d (no CPU fallback)",
             name, t->device_type, t->device_id, ctx->device);
  if (!t->data && ugn_numel(t) > 0) UGN_FAIL(UGN_ERR_INVALID, "%s: null data pointer", name);
  if (dt != DT_BAD && ugn_dtype(t) != dt)
    UGN_FAIL(UGN_ERR_INVALID, "%s: unexpected dtype (code %d, bits %d)", name, t->dtype_code, t->dtype_bits);
  if (t->ndim < lo || t->ndim > hi)
    UGN_FAIL(UGN_ERR_INVALID, "%s: rank %d outside [%d,%d]", name, t->ndim, lo, hi);
  if (t->strides) {
    int64_t expect = 1;
    for (int i = t->ndim - 1; i >= 0; --i) {
      if (t->shape[i] != 1 && t->strides[i] != expect)
        UGN_FAIL(UGN_ERR_INVALID, "%s: tensor must be dense row-major", name);
      expect *= t->shape[i];
    }
  }
  return UGN_OK;
}

int ugn_scratch(ugn_ctx* ctx, size_t bytes, void** out) {
  const int slot = ctx->scratch_next;
  ctx->scratch_next = (slot + 1) % ugn_ctx::kScratchSlots;
  if (bytes > ctx->scratch_bytes[slot]) {
    // grow EVERY slot to the new size at once: this happens in the warm-up step, so the captured step (which
    // requests the same sizes from other slots of the ring) never allocates inside a CUDA-graph capture.
    // Previous smaller buffers may still be referenced by in-flight kernels / captured graphs: they are kept
    // alive (leak-once) rather than freed under them.
    size_t want = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
    for (int i = 0; i < ugn_ctx::kScratchSlots; ++i) {
      if (ctx->scratch_bytes[i] >= want) continue;
      void* p = nullptr;
      UGN_CUDA(cudaMalloc(&p, want));
      ctx->scratch[i] = p;
      ctx->scratch_bytes[i] = want;
    }
  }
  *out = ctx->scratch[slot];
  return UGN_OK;
}

extern "C" int ugn_abi_version(void) { return UGN_ABI_VERSION; }
extern "C" const char* ugn_last_error(void) { return g_last_error.c_str(); }

extern "C" int ugn_ctx_create(int device, ugn_ctx** out) {
  UGN_CHECK(out, "ugn_ctx_create: null out pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    UGN_FAIL(UGN_ERR_DEVICE, "no CUDA device available (%s): ugaitnet_b200 has no CPU fallback",
             e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  UGN_CHECK(device >= 0 && device < count, "device %d out of range [0,%d)", device, count);
  UGN_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  UGN_CUDA(cudaGetDeviceProperties(&prop, device));
  ugn_ctx* c = new ugn_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  const float gs0[4] = {1.f, 1.f, 0.f, 0.f};
  if (cudaMalloc(&c->gscale, sizeof(gs0)) != cudaSuccess ||
      cudaMemcpy(c->gscale, gs0, sizeof(gs0), cudaMemcpyHostToDevice) != cudaSuccess) {
    delete c;
    UGN_FAIL(UGN_ERR_CUDA, "ugn_ctx_create: cannot allocate the gradient-scale buffer");
  }
  *out = c;
  return UGN_OK;
}
extern "C" int ugn_ctx_destroy(ugn_ctx* ctx) {
  if (ctx && ctx->err_flag) cudaFree(ctx->err_flag);
  if (ctx && ctx->gscale) cudaFree(ctx->gscale);
  if (ctx)
    for (int i = 0; i < ugn_ctx::kScratchSlots; ++i)
      if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);   // (outgrown buffers were leaked on purpose)
  delete ctx;
  return UGN_OK;
}
// Synchronises the device and reports asynchronous kernel-side failures.
extern "C" int ugn_ctx_check(ugn_ctx* ctx) {
  UGN_CHECK(ctx, "ugn_ctx_check: null ctx");
  UGN_CUDA(cudaDeviceSynchronize());
  if (ctx->err_flag) {
    int flag = 0;
    UGN_CUDA(cudaMemcpy(&flag, ctx->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
      UGN_CUDA(cudaMemset(ctx->err_flag, 0, sizeof(int)));
      UGN_FAIL(UGN_ERR_CUDA, "tensor-core kernel pipeline timeout (role %d: 1=TMA producer, 2=MMA issuer, 3=epilogue)", flag);
    }
  }
  return UGN_OK;
}
extern "C" int ugn_ctx_has_tcgen05(ugn_ctx* ctx) { return ctx && ctx->cc_major == 10; }
extern "C" int64_t ugn_launch_count(ugn_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- implementations living in other translation units --------------------------------
int ew_pack_input(ugn_ctx*, const float*, void*, int, int, int, int, int, int, int, const int*, const float*,
                  const uint8_t*, float, cudaStream_t, const int8_t* = nullptr, const uint8_t* = nullptr, float = 0.f, float = 0.f, float = 0.f);
int ew_pack_weight(ugn_ctx*, const float*, void*, int, int, long long, int, int, cudaStream_t);
int ew_split(ugn_ctx*, const float*, __nv_bfloat16*, int, int, long long, cudaStream_t);
int ew_bwd_act(ugn_ctx*, const float*, const void*, int, const uint8_t*, void*, float*, int*, int, int, int, int,
               int, int, int, int, int, float, int, cudaStream_t);
int simt_colsum(ugn_ctx*, const float*, long long, int, int, float*, cudaStream_t);
int simt_colsum_bf16(ugn_ctx*, const __nv_bfloat16*, int, int, long long, int, float*, cudaStream_t);
int ew_flatten(ugn_ctx*, const void*, void*, int, int, int, int, int, int, cudaStream_t);
int ew_decode_samples(ugn_ctx*, const void*, int, float*, long long, float, float, float, float, float, cudaStream_t);
int ew_act_mask_bwd(ugn_ctx*, const float*, const float*, const float*, float*, __nv_bfloat16*, int, int,
                    long long, int, float, cudaStream_t);
int ew_fuse_fwd(ugn_ctx*, const FusePtrs&, int, int, int, float*, __nv_bfloat16*, int, int, uint8_t*, float*,
                int, int, cudaStream_t);
int ew_gscale_update(ugn_ctx*, const float*, long long, float, float, cudaStream_t);
int ew_fuse_bwd(ugn_ctx*, const FusePtrs&, int, int, int, const float*, const float*, const uint8_t*,
                const float*, int, int, cudaStream_t);
int ew_softmax_ce(ugn_ctx*, const float*, const int*, float*, float*, int, int, float, float, cudaStream_t);
int ew_optim(ugn_ctx*, int, float*, const float*, float*, float*, const long long*, const float*, int,
             long long, float, float, float, float, float, float*, const float*, const long long*, int, int,
             float*, float, cudaStream_t, int, int, const long long*, const long long*, long long, long long, const long long* = nullptr, const float* = nullptr, const long long* = nullptr, int = 0, const long long* = nullptr, long long = 0);
int simt_conv_fwd(ugn_ctx*, const ConvGeom&, const float*, const float*, const float*, float*, uint8_t*,
                  int, float, int, cudaStream_t);
int simt_conv_dgrad(ugn_ctx*, const ConvGeom&, const float*, const float*, float*, cudaStream_t);
int simt_conv_wgrad(ugn_ctx*, const ConvGeom&, const float*, const float*, float*, float*, cudaStream_t);
int simt_linear_fwd(ugn_ctx*, int, int, int, const float*, const float*, const float*, const float*,
                    float*, int, float, cudaStream_t);
int simt_linear_bwd(ugn_ctx*, int, int, int, const float*, const float*, const float*, float*, float*,
                    float*, cudaStream_t);

// mode of an activation/weight operand: 0 = f32, 1 = 16-bit P=1, 2 = 16-bit P=2, -1 invalid.
// `rank` is the logical rank (without the plane dimension).  16-bit storage is bf16 or fp16 (is_f16).
static int storage_mode(const ugn_tensor* t, int rank) {
  UgnDType dt = ugn_dtype(t);
  if (dt == DT_F32 && t->ndim == rank) return 0;
  if ((dt == DT_BF16 || dt == DT_F16) && t->ndim == rank + 1 && (t->shape[0] == 1 || t->shape[0] == 2))
    return (int)t->shape[0];
  return -1;
}
static inline int is_f16(const ugn_tensor* t) { return ugn_dtype(t) == DT_F16; }
static inline bool is_16(const ugn_tensor* t) { UgnDType d = ugn_dtype(t); return d == DT_BF16 || d == DT_F16; }
// operands of one tensor-core call: all f32, or all 16-bit of the same format
#define UGN_SAME_FMT(a, b, what) \
  UGN_CHECK(ugn_dtype(a) == ugn_dtype(b), what ": operands must share one storage dtype (f32 | bf16 | f16)")
static inline const int64_t* lshape(const ugn_tensor* t, int rank) { return t->shape + (t->ndim - rank); }

static int pack_input_common(ugn_ctx* ctx, const ugn_tensor* x_nchw, const ugn_tensor* src_row,
                             const ugn_tensor* enable, const ugn_tensor* mirror, float noise, ugn_tensor* x_nhwc,
                             void* stream, const ugn_tensor* shift = nullptr, const ugn_tensor* clip = nullptr,
                             float clip_lo = 0.f, float clip_hi = 0.f, float clip_val = 0.f) {
  UGN_CHECK(ctx && x_nchw && x_nhwc, "ugn_pack_input: null argument");
  UGN_TENSOR(x_nchw, DT_F32, 4, 4);
  UGN_TENSOR(x_nhwc, DT_BAD, 4, 5);
  int mode = storage_mode(x_nhwc, 4);
  UGN_CHECK(mode >= 0, "x_nhwc must be f32 [B,H,W,Cp] or 16-bit [P,B,H,W,Cp]");
  const int64_t* s = lshape(x_nhwc, 4);
  int B0 = (int)x_nchw->shape[0], C = (int)x_nchw->shape[1], H = (int)x_nchw->shape[2], W = (int)x_nchw->shape[3];
  int B = (int)s[0];
  UGN_CHECK(s[1] == H && s[2] == W && s[3] >= C, "pack_input: shape mismatch");
  UGN_CHECK(src_row || B == B0, "pack_input: output rows %d != input rows %d (no src_row given)", B, B0);
  UGN_CHECK((size_t)C * (W + 1) * 4 <= 48 * 1024, "pack_input: C*W too large");
  if (src_row) { UGN_TENSOR(src_row, DT_I32, 1, 1); UGN_CHECK(src_row->shape[0] == B, "pack_input: src_row must be i32 [B]"); }
  if (enable) { UGN_TENSOR(enable, DT_F32, 1, 2); UGN_CHECK(ugn_numel(enable) == B, "pack_input: enable must be f32 [B]"); }
  if (mirror) { UGN_TENSOR(mirror, DT_U8, 1, 1); UGN_CHECK(mirror->shape[0] == B, "pack_input: mirror must be u8 [B]"); }
  if (shift) { UGN_TENSOR(shift, DT_I8, 2, 2); UGN_CHECK(shift->shape[0] == B && shift->shape[1] == 2, "pack_input: shift must be i8 [B,2]"); }
  if (clip) { UGN_TENSOR(clip, DT_U8, 1, 1); UGN_CHECK(clip->shape[0] == B && clip_hi > clip_lo, "pack_input: clip must be u8 [B], clip_hi > clip_lo"); }
  if (B == 0) return UGN_OK;
  return ew_pack_input(ctx, ugn_ptr<float>(x_nchw), ugn_ptr<void>(x_nhwc), mode, is_f16(x_nhwc), B, C, H, W, (int)s[3],
                       src_row ? ugn_ptr<int>(src_row) : nullptr, enable ? ugn_ptr<float>(enable) : nullptr,
                       mirror ? ugn_ptr<uint8_t>(mirror) : nullptr, noise, (cudaStream_t)stream,
                       shift ? ugn_ptr<int8_t>(shift) : nullptr, clip ? ugn_ptr<uint8_t>(clip) : nullptr, clip_lo, clip_hi,
                       clip_val);
}
extern "C" int ugn_pack_input_augment(ugn_ctx* ctx, const ugn_tensor* x_base, const ugn_tensor* src_row,
                                      const ugn_tensor* enable, const ugn_tensor* mirror, const ugn_tensor* shift,
                                      const ugn_tensor* clip, float clip_lo, float clip_hi, float clip_val, float noise,
                                      ugn_tensor* x_nhwc, void* stream) {
  return pack_input_common(ctx, x_base, src_row, enable, mirror, noise, x_nhwc, stream, shift, clip, clip_lo, clip_hi,
                           clip_val);
}
extern "C" int ugn_pack_input(ugn_ctx* ctx, const ugn_tensor* x_nchw, ugn_tensor* x_nhwc, void* stream) {
  return pack_input_common(ctx, x_nchw, nullptr, nullptr, nullptr, 0.f, x_nhwc, stream);
}
extern "C" int ugn_pack_input_expand(ugn_ctx* ctx, const ugn_tensor* x_base, const ugn_tensor* src_row,
                                     const ugn_tensor* enable, const ugn_tensor* mirror, float noise,
                                     ugn_tensor* x_nhwc, void* stream) {
  return pack_input_common(ctx, x_base, src_row, enable, mirror, noise, x_nhwc, stream);
}

extern "C" int ugn_decode_samples(ugn_ctx* ctx, const ugn_tensor* raw, float divisor, float mul, float sub, float clip_min,
                                  float clip_max, ugn_tensor* out, void* stream) {
  UGN_CHECK(ctx && raw && out, "ugn_decode_samples: null argument");
  UGN_TENSOR(raw, DT_BAD, 1, 6);
  UGN_TENSOR(out, DT_F32, 1, 6);
  const UgnDType dt = ugn_dtype(raw);
  UGN_CHECK(dt == DT_I16 || dt == DT_U8, "ugn_decode_samples: raw must be int16 or uint8");
  UGN_CHECK(ugn_numel(raw) == ugn_numel(out), "ugn_decode_samples: raw and out must have the same number of elements");
  UGN_CHECK(divisor != 0.f, "ugn_decode_samples: divisor must be non-zero");
  return ew_decode_samples(ctx, ugn_ptr<void>(raw), dt == DT_I16, ugn_ptr<float>(out), ugn_numel(raw), divisor, mul, sub,
                           clip_min, clip_max, (cudaStream_t)stream);
}

extern "C" int ugn_pack_weight(ugn_ctx* ctx, const ugn_tensor* w_master, ugn_tensor* w_packed, void* stream) {
  UGN_CHECK(ctx && w_master && w_packed, "ugn_pack_weight: null argument");
  UGN_TENSOR(w_master, DT_F32, 2, 4);
  UGN_TENSOR(w_packed, DT_BAD, 2, 5);
  int rank = w_master->ndim;
  int mode = storage_mode(w_packed, rank);
  UGN_CHECK(mode >= 0, "w_packed must be f32 (same rank) or bf16 with a leading plane dimension");
  const int64_t* s = lshape(w_packed, rank);
  long long R = 1;
  for (int i = 0; i < rank - 1; ++i) {
    UGN_CHECK(s[i] == w_master->shape[i], "pack_weight: leading dims mismatch");
    R *= s[i];
  }
  int Cin = (int)w_master->shape[rank - 1], Cp = (int)s[rank - 1];
  UGN_CHECK(Cp >= Cin, "pack_weight: padded width smaller than source");
  return ew_pack_weight(ctx, ugn_ptr<float>(w_master), ugn_ptr<void>(w_packed), mode, is_f16(w_packed), R, Cin, Cp,
                        (cudaStream_t)stream);
}

extern "C" int ugn_split_bf16(ugn_ctx* ctx, const ugn_tensor* src, ugn_tensor* dst, void* stream) {
  UGN_CHECK(ctx && src && dst, "ugn_split_bf16: null argument");
  UGN_TENSOR(src, DT_F32, 1, 8);
  UGN_TENSOR(dst, DT_BAD, 2, 8);
  UGN_CHECK(is_16(dst), "split_bf16: dst must be bf16 or f16");
  int P = (int)dst->shape[0];
  UGN_CHECK((P == 1 || P == 2) && ugn_numel(dst) == P * ugn_numel(src), "split_bf16: dst must be [P,...src]");
  return ew_split(ctx, ugn_ptr<float>(src), ugn_ptr<__nv_bfloat16>(dst), P, is_f16(dst), ugn_numel(src), (cudaStream_t)stream);
}

static int conv_geom(const ugn_tensor* x, const ugn_tensor* w, int pool, ConvGeom& g) {
  const int64_t* xs = lshape(x, 4);
  const int64_t* ws = lshape(w, 4);
  g.B = (int)xs[0]; g.H = (int)xs[1]; g.W = (int)xs[2]; g.Cp = (int)xs[3];
  g.Co = (int)ws[0]; g.KH = (int)ws[1]; g.KW = (int)ws[2];
  UGN_CHECK(ws[3] == g.Cp, "conv: weight inner dim %lld != activation channels %d", (long long)ws[3], g.Cp);
  g.Ho = g.H - g.KH + 1; g.Wo = g.W - g.KW + 1;
  UGN_CHECK(g.Ho > 0 && g.Wo > 0, "conv: kernel larger than input");
  g.Hp = pool ? g.Ho / 2 : g.Ho;
  g.Wp = pool ? g.Wo / 2 : g.Wo;
  g.Cin = g.Cp;
  return UGN_OK;
}

extern "C" int ugn_conv2d_fwd(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* bias,
                              ugn_tensor* y, ugn_tensor* pool_idx, int act, float alpha, int pool,
                              void* stream) {
  UGN_CHECK(ctx && x && w && y, "ugn_conv2d_fwd: null argument");
  UGN_TENSOR(x, DT_BAD, 4, 5);
  UGN_TENSOR(w, DT_BAD, 4, 5);
  UGN_TENSOR(y, DT_BAD, 4, 5);
  if (bias) UGN_TENSOR(bias, DT_F32, 1, 1);
  int mx = storage_mode(x, 4), mw = storage_mode(w, 4), my = storage_mode(y, 4);
  UGN_CHECK(mx >= 0 && mx == mw && (my == mx), "conv2d_fwd: x/w/y storage modes must match (%d,%d,%d)", mx, mw, my);
  UGN_SAME_FMT(x, w, "conv2d_fwd");
  UGN_SAME_FMT(x, y, "conv2d_fwd");
  ConvGeom g;
  int rc = conv_geom(x, w, pool, g);
  if (rc != UGN_OK) return rc;
  const int64_t* ys = lshape(y, 4);
  UGN_CHECK(ys[0] == g.B && ys[1] == g.Hp && ys[2] == g.Wp && ys[3] == g.Co,
            "conv2d_fwd: y must be [B=%d,%d,%d,%d]", g.B, g.Hp, g.Wp, g.Co);
  if (bias) UGN_CHECK(bias->shape[0] == g.Co, "conv2d_fwd: bias must be [Cout]");
  uint8_t* idx = nullptr;
  if (pool) {
    UGN_CHECK(pool_idx, "conv2d_fwd: pool_idx required when pool != 0");
    UGN_TENSOR(pool_idx, DT_U8, 4, 4);
    UGN_CHECK(ugn_numel(pool_idx) == (int64_t)g.B * g.Hp * g.Wp * g.Co, "conv2d_fwd: pool_idx shape mismatch");
    idx = ugn_ptr<uint8_t>(pool_idx);
  }
  if (g.B == 0) return UGN_OK;
  if (mx == 0)
    return simt_conv_fwd(ctx, g, ugn_ptr<float>(x), ugn_ptr<float>(w), bias ? ugn_ptr<float>(bias) : nullptr,
                         ugn_ptr<float>(y), idx, act, alpha, pool, (cudaStream_t)stream);
  return tc_conv_fwd(ctx, g, mx, is_f16(x), ugn_ptr<__nv_bfloat16>(x), ugn_ptr<__nv_bfloat16>(w),
                     bias ? ugn_ptr<float>(bias) : nullptr, ugn_ptr<__nv_bfloat16>(y), idx, act, alpha, pool,
                     (cudaStream_t)stream);
}

extern "C" int ugn_conv2d_bwd_act(ugn_ctx* ctx, const ugn_tensor* dy, const ugn_tensor* y,
                                  const ugn_tensor* pool_idx, ugn_tensor* dz, ugn_tensor* db, int act, float alpha,
                                  int pool, void* stream) {
  UGN_CHECK(ctx && dy && y && dz, "ugn_conv2d_bwd_act: null argument");
  UGN_TENSOR(dy, DT_F32, 4, 4);
  UGN_TENSOR(y, DT_BAD, 4, 5);
  UGN_TENSOR(dz, DT_BAD, 4, 5);
  int my = storage_mode(y, 4), mz = storage_mode(dz, 4);
  UGN_CHECK(my >= 0 && mz >= 0, "bwd_act: bad storage mode");
  const int64_t* ys = lshape(y, 4);
  const int64_t* zs = lshape(dz, 4);
  int B = (int)ys[0], Hp = (int)ys[1], Wp = (int)ys[2], C = (int)ys[3];
  int Ho = (int)zs[1], Wo = (int)zs[2];
  UGN_CHECK(zs[0] == B && zs[3] == C && ugn_numel(dy) == (int64_t)B * Hp * Wp * C, "bwd_act: shape mismatch");
  if (pool) {
    UGN_CHECK(Hp == Ho / 2 && Wp == Wo / 2 && pool_idx, "bwd_act: pooled shape mismatch");
    UGN_TENSOR(pool_idx, DT_U8, 4, 4);
  } else {
    UGN_CHECK(Hp == Ho && Wp == Wo, "bwd_act: shape mismatch (no pool)");
  }
  if (B == 0) return UGN_OK;
  if (my > 0 && mz > 0) UGN_SAME_FMT(y, dz, "bwd_act");
  if (db) { UGN_TENSOR(db, DT_F32, 1, 1); UGN_CHECK(db->shape[0] == C, "bwd_act: db must be [C]"); }
  int db_done = 0;
  int rc = ew_bwd_act(ctx, ugn_ptr<float>(dy), ugn_ptr<void>(y), my > 0, pool ? ugn_ptr<uint8_t>(pool_idx) : nullptr,
                      ugn_ptr<void>(dz), db ? ugn_ptr<float>(db) : nullptr, &db_done, mz,
                      mz > 0 ? is_f16(dz) : is_f16(y), B, Ho, Wo, Hp, Wp, C, act, alpha, pool, (cudaStream_t)stream);
  if (rc != UGN_OK || !db || db_done) return rc;
  // layouts the fused reduction does not cover: column sums of the dz just written
  if (mz == 0) return simt_colsum(ctx, ugn_ptr<float>(dz), (long long)B * Ho * Wo, C, C, ugn_ptr<float>(db), (cudaStream_t)stream);
  return simt_colsum_bf16(ctx, ugn_ptr<__nv_bfloat16>(dz), mz, is_f16(dz), (long long)B * Ho * Wo, C, ugn_ptr<float>(db),
                          (cudaStream_t)stream);
}

extern "C" int ugn_conv2d_dgrad(ugn_ctx* ctx, const ugn_tensor* dz, const ugn_tensor* w, ugn_tensor* dx,
                                void* stream) {
  UGN_CHECK(ctx && dz && w && dx, "ugn_conv2d_dgrad: null argument");
  UGN_TENSOR(dz, DT_BAD, 4, 5);
  UGN_TENSOR(w, DT_BAD, 4, 5);
  UGN_TENSOR(dx, DT_F32, 4, 4);
  int mz = storage_mode(dz, 4), mw = storage_mode(w, 4);
  // 16-bit: the MMA uses min(planes(dz), planes(w)) planes of each operand (1 = single pass on the hi planes)
  UGN_CHECK(mz >= 0 && mw >= 0 && (mz == 0) == (mw == 0), "conv2d_dgrad: dz/w storage modes must match");
  UGN_SAME_FMT(dz, w, "conv2d_dgrad");
  ConvGeom g;
  int rc = conv_geom(dx, w, 0, g);
  if (rc != UGN_OK) return rc;
  const int64_t* zs = lshape(dz, 4);
  UGN_CHECK(zs[0] == g.B && zs[1] == g.Ho && zs[2] == g.Wo && zs[3] == g.Co, "conv2d_dgrad: dz shape mismatch");
  if (g.B == 0) return UGN_OK;
  if (mz == 0)
    return simt_conv_dgrad(ctx, g, ugn_ptr<float>(dz), ugn_ptr<float>(w), ugn_ptr<float>(dx), (cudaStream_t)stream);
  return tc_conv_dgrad(ctx, g, mz < mw ? mz : mw, is_f16(dz), ugn_ptr<__nv_bfloat16>(dz), ugn_ptr<__nv_bfloat16>(w), ugn_ptr<float>(dx),
                       (cudaStream_t)stream);
}

// ---- Conv3D branches (use3D, nets/mj_uwyhNets_ba.py:336-417): fp32 validation engine ----
int simt_conv3d_fwd(ugn_ctx*, const Conv3Geom&, const float*, const float*, const float*, float*, int, float, cudaStream_t);
int simt_conv3d_wgrad(ugn_ctx*, const Conv3Geom&, const float*, const float*, float*, float*, cudaStream_t);
int simt_conv3d_dgrad(ugn_ctx*, const Conv3Geom&, const float*, const float*, float*, cudaStream_t);

static int conv3_geom(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* y, int st, int sh, int sw, Conv3Geom& g) {
  UGN_TENSOR(x, DT_F32, 5, 5);
  UGN_TENSOR(w, DT_F32, 5, 5);
  UGN_TENSOR(y, DT_F32, 5, 5);
  UGN_CHECK(st >= 1 && sh >= 1 && sw >= 1, "conv3d: strides must be >= 1");
  g.B = (int)x->shape[0]; g.T = (int)x->shape[1]; g.H = (int)x->shape[2]; g.W = (int)x->shape[3]; g.C = (int)x->shape[4];
  g.Co = (int)w->shape[0]; g.KT = (int)w->shape[1]; g.KH = (int)w->shape[2]; g.KW = (int)w->shape[3];
  g.ST = st; g.SH = sh; g.SW = sw;
  UGN_CHECK(w->shape[4] == g.C, "conv3d: weight inner dim %lld != input channels %d", (long long)w->shape[4], g.C);
  UGN_CHECK(g.T >= g.KT && g.H >= g.KH && g.W >= g.KW, "conv3d: kernel larger than the input volume");
  g.To = (g.T - g.KT) / st + 1; g.Ho = (g.H - g.KH) / sh + 1; g.Wo = (g.W - g.KW) / sw + 1;
  UGN_CHECK(y->shape[0] == g.B && y->shape[1] == g.To && y->shape[2] == g.Ho && y->shape[3] == g.Wo && y->shape[4] == g.Co,
            "conv3d: output must be [%d,%d,%d,%d,%d]", g.B, g.To, g.Ho, g.Wo, g.Co);
  UGN_CHECK((long long)g.B * g.T * g.H * g.W * g.C < 0x7fffffffLL && (long long)g.B * g.To * g.Ho * g.Wo * g.Co < 0x7fffffffLL,
            "conv3d: tensor too large for 32-bit offsets");
  return UGN_OK;
}

extern "C" int ugn_conv3d_fwd(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* bias, ugn_tensor* y,
                              int st, int sh, int sw, int act, float alpha, void* stream) {
  UGN_CHECK(ctx && x && w && y, "ugn_conv3d_fwd: null argument");
  Conv3Geom g;
  int rc = conv3_geom(ctx, x, w, y, st, sh, sw, g);
  if (rc != UGN_OK) return rc;
  if (bias) { UGN_TENSOR(bias, DT_F32, 1, 1); UGN_CHECK(bias->shape[0] == g.Co, "conv3d: bias must be f32 [Cout]"); }
  if (g.B == 0) return UGN_OK;
  return simt_conv3d_fwd(ctx, g, ugn_ptr<float>(x), ugn_ptr<float>(w), bias ? ugn_ptr<float>(bias) : nullptr, ugn_ptr<float>(y),
                         act, alpha, (cudaStream_t)stream);
}

extern "C" int ugn_conv3d_wgrad(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* dz, ugn_tensor* dw, ugn_tensor* db,
                                int st, int sh, int sw, void* stream) {
  UGN_CHECK(ctx && x && dz && dw, "ugn_conv3d_wgrad: null argument");
  Conv3Geom g;
  int rc = conv3_geom(ctx, x, dw, dz, st, sh, sw, g);
  if (rc != UGN_OK) return rc;
  if (db) { UGN_TENSOR(db, DT_F32, 1, 1); UGN_CHECK(db->shape[0] == g.Co, "conv3d: db must be f32 [Cout]"); }
  if (g.B == 0) return UGN_OK;
  return simt_conv3d_wgrad(ctx, g, ugn_ptr<float>(x), ugn_ptr<float>(dz), ugn_ptr<float>(dw), db ? ugn_ptr<float>(db) : nullptr,
                           (cudaStream_t)stream);
}

extern "C" int ugn_conv3d_dgrad(ugn_ctx* ctx, const ugn_tensor* dz, const ugn_tensor* w, ugn_tensor* dx, int st, int sh, int sw,
                                void* stream) {
  UGN_CHECK(ctx && dz && w && dx, "ugn_conv3d_dgrad: null argument");
  Conv3Geom g;
  int rc = conv3_geom(ctx, dx, w, dz, st, sh, sw, g);
  if (rc != UGN_OK) return rc;
  if (g.B == 0) return UGN_OK;
  return simt_conv3d_dgrad(ctx, g, ugn_ptr<float>(dz), ugn_ptr<float>(w), ugn_ptr<float>(dx), (cudaStream_t)stream);
}

extern "C" int ugn_conv2d_wgrad(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* dz, ugn_tensor* dw,
                                ugn_tensor* db, void* stream) {
  UGN_CHECK(ctx && x && dz && dw, "ugn_conv2d_wgrad: null argument");
  UGN_TENSOR(x, DT_BAD, 4, 5);
  UGN_TENSOR(dz, DT_BAD, 4, 5);
  UGN_TENSOR(dw, DT_F32, 4, 4);
  if (db) UGN_TENSOR(db, DT_F32, 1, 1);
  int mx = storage_mode(x, 4), mz = storage_mode(dz, 4);
  UGN_CHECK(mx >= 0 && mz >= 0 && (mx == 0) == (mz == 0), "conv2d_wgrad: x/dz storage modes must match");
  UGN_SAME_FMT(x, dz, "conv2d_wgrad");
  const int64_t* xs = lshape(x, 4);
  const int64_t* zs = lshape(dz, 4);
  ConvGeom g;
  g.B = (int)xs[0]; g.H = (int)xs[1]; g.W = (int)xs[2]; g.Cp = (int)xs[3];
  g.Co = (int)dw->shape[0]; g.KH = (int)dw->shape[1]; g.KW = (int)dw->shape[2]; g.Cin = (int)dw->shape[3];
  g.Ho = g.H - g.KH + 1; g.Wo = g.W - g.KW + 1; g.Hp = g.Ho; g.Wp = g.Wo;
  UGN_CHECK(g.Cin <= g.Cp, "conv2d_wgrad: dw inner dim larger than activation channels");
  UGN_CHECK(zs[0] == g.B && zs[1] == g.Ho && zs[2] == g.Wo && zs[3] == g.Co, "conv2d_wgrad: dz shape mismatch");
  if (db) UGN_CHECK(db->shape[0] == g.Co, "conv2d_wgrad: db must be [Cout]");
  if (mx == 0)
    return simt_conv_wgrad(ctx, g, ugn_ptr<float>(x), ugn_ptr<float>(dz), ugn_ptr<float>(dw),
                           db ? ugn_ptr<float>(db) : nullptr, (cudaStream_t)stream);
  return tc_conv_wgrad(ctx, g, mx < mz ? mx : mz, is_f16(x), ugn_ptr<__nv_bfloat16>(x), ugn_ptr<__nv_bfloat16>(dz), ugn_ptr<float>(dw),
                       db ? ugn_ptr<float>(db) : nullptr, (cudaStream_t)stream);
}

extern "C" int ugn_flatten_chw(ugn_ctx* ctx, const ugn_tensor* y, ugn_tensor* flat, void* stream) {
  UGN_CHECK(ctx && y && flat, "ugn_flatten_chw: null argument");
  UGN_TENSOR(y, DT_BAD, 4, 5);
  UGN_TENSOR(flat, DT_BAD, 2, 3);
  int my = storage_mode(y, 4), mf = storage_mode(flat, 2);
  UGN_CHECK(my >= 0 && my == mf, "flatten: storage modes must match");
  UGN_SAME_FMT(y, flat, "flatten");
  const int64_t* ys = lshape(y, 4);
  UGN_CHECK(lshape(flat, 2)[0] == ys[0] && lshape(flat, 2)[1] == ys[1] * ys[2] * ys[3], "flatten: shape mismatch");
  if (ys[0] == 0) return UGN_OK;
  return ew_flatten(ctx, ugn_ptr<void>(y), ugn_ptr<void>(flat), my > 0, my > 0 ? my : 1, (int)ys[0],
                    (int)(ys[1] * ys[2]), (int)ys[3], 1, (cudaStream_t)stream);
}
extern "C" int ugn_unflatten_chw(ugn_ctx* ctx, const ugn_tensor* dflat, ugn_tensor* dy, void* stream) {
  UGN_CHECK(ctx && dflat && dy, "ugn_unflatten_chw: null argument");
  UGN_TENSOR(dflat, DT_F32, 2, 2);
  UGN_TENSOR(dy, DT_F32, 4, 4);
  UGN_CHECK(dflat->shape[0] == dy->shape[0] && dflat->shape[1] == dy->shape[1] * dy->shape[2] * dy->shape[3],
            "unflatten: shape mismatch");
  if (dy->shape[0] == 0) return UGN_OK;
  return ew_flatten(ctx, ugn_ptr<void>(dflat), ugn_ptr<void>(dy), 0, 1, (int)dy->shape[0],
                    (int)(dy->shape[1] * dy->shape[2]), (int)dy->shape[3], 0, (cudaStream_t)stream);
}

extern "C" int ugn_linear_fwd(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* bias,
                              const ugn_tensor* drop_mask, ugn_tensor* y, ugn_tensor* y16, int act, float alpha,
                              void* stream) {
  UGN_CHECK(ctx && x && w && y, "ugn_linear_fwd: null argument");
  UGN_TENSOR(x, DT_BAD, 2, 3);
  UGN_TENSOR(w, DT_BAD, 2, 3);
  UGN_TENSOR(y, DT_F32, 2, 2);
  int mx = storage_mode(x, 2), mw = storage_mode(w, 2);
  UGN_CHECK(mx >= 0 && mx == mw, "linear_fwd: x/w storage modes must match");
  UGN_SAME_FMT(x, w, "linear_fwd");
  const int64_t* xs = lshape(x, 2);
  const int64_t* ws = lshape(w, 2);
  int B = (int)xs[0], K = (int)xs[1], N = (int)ws[0];
  UGN_CHECK(ws[1] == K && y->shape[0] == B && y->shape[1] == N, "linear_fwd: shape mismatch");
  if (bias) { UGN_TENSOR(bias, DT_F32, 1, 1); UGN_CHECK(bias->shape[0] == N, "linear_fwd: bias must be [N]"); }
  if (drop_mask) { UGN_TENSOR(drop_mask, DT_F32, 2, 2); UGN_CHECK(ugn_numel(drop_mask) == (int64_t)B * N, "linear_fwd: mask must be [B,N]"); }
  int P16 = 0;
  if (y16) {
    UGN_TENSOR(y16, DT_BAD, 3, 3);
    UGN_CHECK(is_16(y16), "linear_fwd: y16 must be bf16 or f16");
    P16 = (int)y16->shape[0];
    UGN_CHECK((P16 == 1 || P16 == 2) && y16->shape[1] == B && y16->shape[2] == N, "linear_fwd: y16 must be [P,B,N]");
  }
  if (B == 0) return UGN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (mx == 0) {
    rc = simt_linear_fwd(ctx, B, N, K, ugn_ptr<float>(x), ugn_ptr<float>(w), bias ? ugn_ptr<float>(bias) : nullptr,
                         drop_mask ? ugn_ptr<float>(drop_mask) : nullptr, ugn_ptr<float>(y), act, alpha, st);
  } else {
    if (y16) UGN_SAME_FMT(x, y16, "linear_fwd");
    return tc_linear_fwd(ctx, mx, is_f16(x), B, N, K, ugn_ptr<__nv_bfloat16>(x), ugn_ptr<__nv_bfloat16>(w),
                         bias ? ugn_ptr<float>(bias) : nullptr, drop_mask ? ugn_ptr<float>(drop_mask) : nullptr,
                         ugn_ptr<float>(y), act, alpha, st, y16 ? ugn_ptr<__nv_bfloat16>(y16) : nullptr, P16);
  }
  if (rc != UGN_OK) return rc;
  if (y16) return ew_split(ctx, ugn_ptr<float>(y), ugn_ptr<__nv_bfloat16>(y16), P16, is_f16(y16), (long long)B * N, st);
  return UGN_OK;
}

extern "C" int ugn_act_mask_bwd(ugn_ctx* ctx, const ugn_tensor* dy, const ugn_tensor* y, const ugn_tensor* drop_mask,
                                ugn_tensor* dz, ugn_tensor* dz16, int act, float alpha, void* stream) {
  UGN_CHECK(ctx && dy && (dz || dz16), "ugn_act_mask_bwd: null argument");
  UGN_TENSOR(dy, DT_F32, 1, 4);
  long long n = ugn_numel(dy);
  if (y) { UGN_TENSOR(y, DT_F32, 1, 4); UGN_CHECK(ugn_numel(y) == n, "act_mask_bwd: y shape mismatch"); }
  if (drop_mask) { UGN_TENSOR(drop_mask, DT_F32, 1, 4); UGN_CHECK(ugn_numel(drop_mask) == n, "act_mask_bwd: mask shape mismatch"); }
  if (dz) { UGN_TENSOR(dz, DT_F32, 1, 4); UGN_CHECK(ugn_numel(dz) == n, "act_mask_bwd: dz shape mismatch"); }
  int P = 0;
  if (dz16) {
    UGN_TENSOR(dz16, DT_BAD, 2, 5);
    UGN_CHECK(is_16(dz16), "act_mask_bwd: dz16 must be bf16 or f16");
    P = (int)dz16->shape[0];
    UGN_CHECK((P == 1 || P == 2) && ugn_numel(dz16) == P * n, "act_mask_bwd: dz16 must be [P,...]");
  }
  if (n == 0) return UGN_OK;
  return ew_act_mask_bwd(ctx, ugn_ptr<float>(dy), y ? ugn_ptr<float>(y) : nullptr,
                         drop_mask ? ugn_ptr<float>(drop_mask) : nullptr, dz ? ugn_ptr<float>(dz) : nullptr,
                         dz16 ? ugn_ptr<__nv_bfloat16>(dz16) : nullptr, P, dz16 ? is_f16(dz16) : 0, n, act, alpha,
                         (cudaStream_t)stream);
}

static int linear_bwd_common(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* dz,
                             ugn_tensor* dx, ugn_tensor* dw, ugn_tensor* db, void* stream, const ugn_tensor* dx_mask,
                             ugn_tensor* dx16, ugn_tensor* dbx) {
  UGN_CHECK(ctx && x && w && dz, "ugn_linear_bwd: null argument");
  UGN_TENSOR(x, DT_BAD, 2, 3);
  UGN_TENSOR(w, DT_BAD, 2, 3);
  UGN_TENSOR(dz, DT_BAD, 2, 3);
  int mx = storage_mode(x, 2), mw = storage_mode(w, 2), mz = storage_mode(dz, 2);
  UGN_CHECK(mx >= 0 && mw >= 0 && mz >= 0 && (mx == 0) == (mw == 0) && (mx == 0) == (mz == 0),
            "linear_bwd: x/w/dz storage modes must match");
  UGN_SAME_FMT(x, w, "linear_bwd");
  UGN_SAME_FMT(x, dz, "linear_bwd");
  const int mp = mx < mz ? (mx < mw ? mx : mw) : (mz < mw ? mz : mw);   // planes used by the MMAs
  const int64_t* xs = lshape(x, 2);
  const int64_t* ws = lshape(w, 2);
  const int64_t* zs = lshape(dz, 2);
  int B = (int)xs[0], K = (int)xs[1], N = (int)ws[0];
  UGN_CHECK(ws[1] == K && zs[0] == B && zs[1] == N, "linear_bwd: shape mismatch");
  if (dx) { UGN_TENSOR(dx, DT_F32, 2, 2); UGN_CHECK(dx->shape[0] == B && dx->shape[1] == K, "linear_bwd: dx must be [B,K]"); }
  if (dw) { UGN_TENSOR(dw, DT_F32, 2, 2); UGN_CHECK(dw->shape[0] == N && dw->shape[1] <= K, "linear_bwd: dw must be [N,<=K]"); }
  if (db) { UGN_TENSOR(db, DT_F32, 1, 1); UGN_CHECK(db->shape[0] == N, "linear_bwd: db must be [N]"); }
  if (dw) UGN_CHECK(dw->shape[1] == K, "linear_bwd: dw inner dim must equal K (dense weights are never padded)");
  cudaStream_t st = (cudaStream_t)stream;
  int P16 = 0;
  if (dx_mask) { UGN_TENSOR(dx_mask, DT_F32, 2, 2); UGN_CHECK(dx_mask->shape[0] == B && dx_mask->shape[1] == K, "linear_bwd: dx_mask must be [B,K]"); }
  if (dx16) {
    UGN_TENSOR(dx16, DT_BAD, 3, 3);
    UGN_CHECK(is_16(dx16) && mx > 0, "linear_bwd: dx16 needs 16-bit operands");
    UGN_SAME_FMT(x, dx16, "linear_bwd");
    P16 = (int)dx16->shape[0];
    UGN_CHECK((P16 == 1 || P16 == 2) && dx16->shape[1] == B && dx16->shape[2] == K, "linear_bwd: dx16 must be [P,B,K]");
  }
  if (dbx) { UGN_TENSOR(dbx, DT_F32, 1, 1); UGN_CHECK(dbx->shape[0] == K, "linear_bwd: dbx must be [K]"); }
  UGN_CHECK(mx > 0 || (!dx_mask && !dx16 && !dbx), "linear_bwd: the fused input-gradient outputs need the tensor-core storage mode");
  if (mx == 0)
    return simt_linear_bwd(ctx, B, N, K, ugn_ptr<float>(x), ugn_ptr<float>(w), ugn_ptr<float>(dz),
                           dx ? ugn_ptr<float>(dx) : nullptr, dw ? ugn_ptr<float>(dw) : nullptr,
                           db ? ugn_ptr<float>(db) : nullptr, st);
  return tc_linear_bwd(ctx, mp, is_f16(x), B, N, K, ugn_ptr<__nv_bfloat16>(x), ugn_ptr<__nv_bfloat16>(w),
                       ugn_ptr<__nv_bfloat16>(dz), dx ? ugn_ptr<float>(dx) : nullptr,
                       dw ? ugn_ptr<float>(dw) : nullptr, db ? ugn_ptr<float>(db) : nullptr, st,
                       dx_mask ? ugn_ptr<float>(dx_mask) : nullptr, dx16 ? ugn_ptr<__nv_bfloat16>(dx16) : nullptr, P16,
                       dbx ? ugn_ptr<float>(dbx) : nullptr);
}
extern "C" int ugn_linear_bwd(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* dz,
                              ugn_tensor* dx, ugn_tensor* dw, ugn_tensor* db, void* stream) {
  return linear_bwd_common(ctx, x, w, dz, dx, dw, db, stream, nullptr, nullptr, nullptr);
}
extern "C" int ugn_linear_bwd_ex(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* dz,
                                 ugn_tensor* dx, const ugn_tensor* dx_mask, ugn_tensor* dx16, ugn_tensor* dbx,
                                 ugn_tensor* dw, ugn_tensor* db, void* stream) {
  return linear_bwd_common(ctx, x, w, dz, dx, dw, db, stream, dx_mask, dx16, dbx);
}

// ---- dropout from a counter-based generator (no mask tensors) -----------------------------------
int ew_dropout_advance(ugn_ctx*, unsigned long long*, cudaStream_t);
int ew_dropout_mask(ugn_ctx*, const unsigned long long*, int, float, float*, long long, cudaStream_t);
static int rng_check(ugn_ctx* ctx, const ugn_tensor* rng, float keep) {
  UGN_CHECK(rng, "dropout: rng state required");
  UGN_TENSOR(rng, DT_I64, 1, 1);
  UGN_CHECK(rng->shape[0] >= 2 && keep > 0.f && keep <= 1.f, "dropout: rng must be i64 [2] = {seed, step}, keep in (0, 1]");
  return UGN_OK;
}
extern "C" int ugn_dropout_advance(ugn_ctx* ctx, ugn_tensor* rng, void* stream) {
  UGN_CHECK(ctx, "ugn_dropout_advance: null ctx");
  int rc = rng_check(ctx, rng, 1.f);
  if (rc != UGN_OK) return rc;
  return ew_dropout_advance(ctx, ugn_ptr<unsigned long long>(rng), (cudaStream_t)stream);
}
extern "C" int ugn_dropout_mask(ugn_ctx* ctx, const ugn_tensor* rng, int layer, float keep, ugn_tensor* out, void* stream) {
  UGN_CHECK(ctx && out, "ugn_dropout_mask: null argument");
  int rc = rng_check(ctx, rng, keep);
  if (rc != UGN_OK) return rc;
  UGN_TENSOR(out, DT_F32, 1, 4);
  return ew_dropout_mask(ctx, ugn_ptr<unsigned long long>(rng), layer, keep, ugn_ptr<float>(out), ugn_numel(out),
                         (cudaStream_t)stream);
}
extern "C" int ugn_linear_fwd_philox(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* bias,
                                     const ugn_tensor* rng, int layer, float keep, ugn_tensor* y, ugn_tensor* y16, int act,
                                     float alpha, void* stream) {
  UGN_CHECK(ctx && x, "ugn_linear_fwd_philox: null argument");
  int rc = rng_check(ctx, rng, keep);
  if (rc != UGN_OK) return rc;
  UGN_CHECK(is_16(x), "ugn_linear_fwd_philox: tensor-core storage mode only (fp32 validation mode takes mask tensors)");
  ctx->drop_rng = ugn_ptr<unsigned long long>(rng); ctx->drop_layer = layer; ctx->drop_keep = keep;
  rc = ugn_linear_fwd(ctx, x, w, bias, nullptr, y, y16, act, alpha, stream);
  ctx->drop_rng = nullptr;
  return rc;
}
extern "C" int ugn_linear_bwd_philox(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* w, const ugn_tensor* dz,
                                     ugn_tensor* dx, const ugn_tensor* rng, int layer, float keep, ugn_tensor* dx16,
                                     ugn_tensor* dbx, ugn_tensor* dw, ugn_tensor* db, void* stream) {
  UGN_CHECK(ctx && x && dx, "ugn_linear_bwd_philox: null argument (dx is required)");
  int rc = rng_check(ctx, rng, keep);
  if (rc != UGN_OK) return rc;
  UGN_CHECK(is_16(x), "ugn_linear_bwd_philox: tensor-core storage mode only");
  ctx->drop_rng = ugn_ptr<unsigned long long>(rng); ctx->drop_layer = layer; ctx->drop_keep = keep;
  rc = linear_bwd_common(ctx, x, w, dz, dx, dw, db, stream, nullptr, dx16, dbx);
  ctx->drop_rng = nullptr;
  return rc;
}

static int fuse_common(ugn_ctx* ctx, int nmods, const ugn_tensor* const* br, const ugn_tensor* const* flags,
                       int& B, int& d) {
  UGN_CHECK(nmods >= 1 && nmods <= 4, "fuse: 1..4 modalities supported");
  for (int m = 0; m < nmods; ++m) {
    UGN_TENSOR(br[m], DT_F32, 2, 2);
    UGN_TENSOR(flags[m], DT_F32, 1, 2);
    if (m == 0) { B = (int)br[0]->shape[0]; d = (int)br[0]->shape[1]; }
    UGN_CHECK(br[m]->shape[0] == B && br[m]->shape[1] == d && ugn_numel(flags[m]) == B, "fuse: shape mismatch at modality %d", m);
  }
  UGN_CHECK((size_t)d * 4 <= 48 * 1024, "fuse: signature dimension too large");
  return UGN_OK;
}

extern "C" int ugn_fuse_fwd(ugn_ctx* ctx, int nmods, const ugn_tensor* const* br, const ugn_tensor* const* flags,
                            ugn_tensor* sig, ugn_tensor* sig16, ugn_tensor* winner, ugn_tensor* inv_norm, int merge,
                            int normalize, void* stream) {
  UGN_CHECK(ctx && br && flags && sig, "ugn_fuse_fwd: null argument");
  int B = 0, d = 0;
  int rc = fuse_common(ctx, nmods, br, flags, B, d);
  if (rc != UGN_OK) return rc;
  UGN_TENSOR(sig, DT_F32, 2, 2);
  UGN_CHECK(sig->shape[0] == B && sig->shape[1] == d, "fuse_fwd: sig must be [B,d]");
  int P = 0;
  if (sig16) {
    UGN_TENSOR(sig16, DT_BAD, 3, 3);
    UGN_CHECK(is_16(sig16), "fuse_fwd: sig16 must be bf16 or f16");
    P = (int)sig16->shape[0];
    UGN_CHECK((P == 1 || P == 2) && sig16->shape[1] == B && sig16->shape[2] == d, "fuse_fwd: sig16 must be [P,B,d]");
  }
  if (winner) { UGN_TENSOR(winner, DT_U8, 2, 2); UGN_CHECK(ugn_numel(winner) == (int64_t)B * d, "fuse_fwd: winner must be [B,d]"); }
  if (inv_norm) { UGN_TENSOR(inv_norm, DT_F32, 2, 2); UGN_CHECK(inv_norm->shape[0] == B && inv_norm->shape[1] == 2, "fuse_fwd: inv_norm must be [B,2]"); }
  UGN_CHECK(merge >= 0 && merge <= 2, "fuse_fwd: unknown merge mode %d", merge);
  if (B == 0) return UGN_OK;
  FusePtrs p{};
  for (int m = 0; m < nmods; ++m) { p.br[m] = ugn_ptr<float>(br[m]); p.flag[m] = ugn_ptr<float>(flags[m]); }
  return ew_fuse_fwd(ctx, p, nmods, B, d, ugn_ptr<float>(sig), sig16 ? ugn_ptr<__nv_bfloat16>(sig16) : nullptr, P,
                     sig16 ? is_f16(sig16) : 0,
                     winner ? ugn_ptr<uint8_t>(winner) : nullptr, inv_norm ? ugn_ptr<float>(inv_norm) : nullptr,
                     merge, normalize, (cudaStream_t)stream);
}

int ew_fuse_fc1_fwd(ugn_ctx*, const FusePtrs&, int, int, int, float*, __nv_bfloat16*, int, int, uint8_t*, float*, int, int,
                    const float*, const float*, int, float*, const float*, float*, int, float, cudaStream_t);

extern "C" int ugn_fuse_fc1_fwd(ugn_ctx* ctx, int nmods, const ugn_tensor* const* br, const ugn_tensor* const* flags,
                                ugn_tensor* sig, ugn_tensor* sig16, ugn_tensor* winner, ugn_tensor* inv_norm, int merge,
                                int normalize, const ugn_tensor* code_w, const ugn_tensor* code_b, ugn_tensor* code,
                                const ugn_tensor* drop_mask, ugn_tensor* dropcode, int act, float alpha, void* stream) {
  UGN_CHECK(ctx && br && flags && sig && code_w && code, "ugn_fuse_fc1_fwd: null argument");
  int B = 0, d = 0;
  int rc = fuse_common(ctx, nmods, br, flags, B, d);
  if (rc != UGN_OK) return rc;
  UGN_CHECK(d % 4 == 0 && (size_t)d * 4 <= 48 * 1024, "fuse_fc1_fwd: signature width must be a multiple of 4 and <= 12288");
  UGN_TENSOR(sig, DT_F32, 2, 2);
  UGN_CHECK(sig->shape[0] == B && sig->shape[1] == d, "fuse_fc1_fwd: sig must be [B,d]");
  int P = 0;
  if (sig16) {
    UGN_TENSOR(sig16, DT_BAD, 3, 3);
    UGN_CHECK(is_16(sig16), "fuse_fc1_fwd: sig16 must be bf16 or f16");
    P = (int)sig16->shape[0];
    UGN_CHECK((P == 1 || P == 2) && sig16->shape[1] == B && sig16->shape[2] == d, "fuse_fc1_fwd: sig16 must be [P,B,d]");
  }
  if (winner) { UGN_TENSOR(winner, DT_U8, 2, 2); UGN_CHECK(ugn_numel(winner) == (int64_t)B * d, "fuse_fc1_fwd: winner must be [B,d]"); }
  if (inv_norm) { UGN_TENSOR(inv_norm, DT_F32, 2, 2); UGN_CHECK(inv_norm->shape[0] == B && inv_norm->shape[1] == 2, "fuse_fc1_fwd: inv_norm must be [B,2]"); }
  UGN_CHECK(merge >= 0 && merge <= 2, "fuse_fc1_fwd: unknown merge mode %d", merge);
  UGN_TENSOR(code_w, DT_F32, 2, 2);
  const int nc = (int)code_w->shape[0];
  UGN_CHECK(code_w->shape[1] == d, "fuse_fc1_fwd: code_w must be f32 [nc,d]");
  if (code_b) { UGN_TENSOR(code_b, DT_F32, 1, 1); UGN_CHECK(code_b->shape[0] == nc, "fuse_fc1_fwd: code_b must be f32 [nc]"); }
  UGN_TENSOR(code, DT_F32, 2, 2);
  UGN_CHECK(code->shape[0] == B && code->shape[1] == nc, "fuse_fc1_fwd: code must be f32 [B,nc]");
  if (drop_mask) { UGN_TENSOR(drop_mask, DT_F32, 2, 2); UGN_CHECK(ugn_numel(drop_mask) == (int64_t)B * nc && dropcode, "fuse_fc1_fwd: drop_mask [B,nc] needs dropcode"); }
  if (dropcode) { UGN_TENSOR(dropcode, DT_F32, 2, 2); UGN_CHECK(ugn_numel(dropcode) == (int64_t)B * nc, "fuse_fc1_fwd: dropcode must be f32 [B,nc]"); }
  if (B == 0) return UGN_OK;
  FusePtrs p{};
  for (int m = 0; m < nmods; ++m) { p.br[m] = ugn_ptr<float>(br[m]); p.flag[m] = ugn_ptr<float>(flags[m]); }
  return ew_fuse_fc1_fwd(ctx, p, nmods, B, d, ugn_ptr<float>(sig), sig16 ? ugn_ptr<__nv_bfloat16>(sig16) : nullptr, P,
                         sig16 ? is_f16(sig16) : 0, winner ? ugn_ptr<uint8_t>(winner) : nullptr,
                         inv_norm ? ugn_ptr<float>(inv_norm) : nullptr, merge, normalize, ugn_ptr<float>(code_w),
                         code_b ? ugn_ptr<float>(code_b) : nullptr, nc, ugn_ptr<float>(code),
                         drop_mask ? ugn_ptr<float>(drop_mask) : nullptr, dropcode ? ugn_ptr<float>(dropcode) : nullptr,
                         act, alpha, (cudaStream_t)stream);
}

extern "C" int ugn_fuse_bwd(ugn_ctx* ctx, int nmods, const ugn_tensor* dsig, const ugn_tensor* sig,
                            const ugn_tensor* winner, const ugn_tensor* inv_norm, const ugn_tensor* const* flags,
                            ugn_tensor* const* dbr, int merge, int normalize, void* stream) {
  UGN_CHECK(ctx && dsig && sig && flags && dbr, "ugn_fuse_bwd: null argument");
  int B = 0, d = 0;
  int rc = fuse_common(ctx, nmods, (const ugn_tensor* const*)dbr, flags, B, d);
  if (rc != UGN_OK) return rc;
  UGN_TENSOR(dsig, DT_F32, 2, 2);
  UGN_TENSOR(sig, DT_F32, 2, 2);
  UGN_CHECK(ugn_numel(dsig) == (int64_t)B * d && ugn_numel(sig) == (int64_t)B * d, "fuse_bwd: shape mismatch");
  if (merge != UGN_MERGE_AVG) { UGN_CHECK(winner, "fuse_bwd: winner required"); UGN_TENSOR(winner, DT_U8, 2, 2); }
  if (normalize) { UGN_CHECK(inv_norm, "fuse_bwd: inv_norm required"); UGN_TENSOR(inv_norm, DT_F32, 2, 2); }
  if (B == 0) return UGN_OK;
  FusePtrs p{};
  for (int m = 0; m < nmods; ++m) { p.dbr[m] = ugn_ptr<float>(dbr[m]); p.flag[m] = ugn_ptr<float>(flags[m]); }
  return ew_fuse_bwd(ctx, p, nmods, B, d, ugn_ptr<float>(dsig), ugn_ptr<float>(sig),
                     winner ? ugn_ptr<uint8_t>(winner) : nullptr, inv_norm ? ugn_ptr<float>(inv_norm) : nullptr, merge,
                     normalize, (cudaStream_t)stream);
}

extern "C" int ugn_softmax_ce_ls(ugn_ctx* ctx, const ugn_tensor* logits, const ugn_tensor* labels, ugn_tensor* loss_acc,
                                 ugn_tensor* dlogits, float scale, float label_smoothing, void* stream);
extern "C" int ugn_softmax_ce(ugn_ctx* ctx, const ugn_tensor* logits, const ugn_tensor* labels, ugn_tensor* loss_acc,
                              ugn_tensor* dlogits, float scale, void* stream) {
  return ugn_softmax_ce_ls(ctx, logits, labels, loss_acc, dlogits, scale, 0.f, stream);
}
extern "C" int ugn_softmax_ce_ls(ugn_ctx* ctx, const ugn_tensor* logits, const ugn_tensor* labels, ugn_tensor* loss_acc,
                                 ugn_tensor* dlogits, float scale, float label_smoothing, void* stream) {
  UGN_CHECK(ctx && logits && labels && loss_acc, "ugn_softmax_ce: null argument");
  UGN_CHECK(label_smoothing >= 0.f && label_smoothing < 1.f, "softmax_ce: label_smoothing must be in [0,1)");
  UGN_TENSOR(logits, DT_F32, 2, 2);
  UGN_TENSOR(labels, DT_I32, 1, 2);
  UGN_TENSOR(loss_acc, DT_F32, 1, 1);
  int B = (int)logits->shape[0], C = (int)logits->shape[1];
  UGN_CHECK(ugn_numel(labels) == B && loss_acc->shape[0] >= 2, "softmax_ce: shape mismatch");
  if (dlogits) { UGN_TENSOR(dlogits, DT_F32, 2, 2); UGN_CHECK(ugn_numel(dlogits) == (int64_t)B * C, "softmax_ce: dlogits shape mismatch"); }
  if (B == 0) return UGN_OK;
  return ew_softmax_ce(ctx, ugn_ptr<float>(logits), ugn_ptr<int>(labels), ugn_ptr<float>(loss_acc),
                       dlogits ? ugn_ptr<float>(dlogits) : nullptr, B, C, scale, label_smoothing, (cudaStream_t)stream);
}

static int optim_common(ugn_ctx* ctx, int opt, ugn_tensor* w, const ugn_tensor* g, ugn_tensor* m, ugn_tensor* v,
                        const ugn_tensor* seg_off, const ugn_tensor* seg_l2, float lr, float b1, float b2, float eps,
                        float gscale, ugn_tensor* reg_out, const ugn_tensor* lr_dev, const ugn_tensor* pack_table,
                        int pack_planes, int pack_f16, void* stream, ugn_tensor* vhat = nullptr, float wd = 0.f,
                        int world = 1, int rank = 0, const int64_t* g_peers = nullptr, const int64_t* w_peers = nullptr,
                        int64_t g_mc = 0, int64_t w_mc = 0, const int64_t* reg_peers = nullptr,
                        const ugn_tensor* stage = nullptr, const int64_t* staged_ranges = nullptr, int n_ranges = 0,
                        const int64_t* cw_peers = nullptr, int64_t cw_mc = 0) {
  UGN_CHECK(ctx && w && g && v && seg_off && seg_l2, "optimizer: null argument");
  if (vhat) { UGN_TENSOR(vhat, DT_F32, 1, 1); UGN_CHECK(vhat->shape[0] == w->shape[0], "optimizer: vhat arena length mismatch"); }
  UGN_CHECK(wd >= 0.f && wd < 1.f, "optimizer: decoupled weight decay must be in [0,1)");
  if (lr_dev) UGN_TENSOR(lr_dev, DT_F32, 1, 1);
  UGN_TENSOR(w, DT_F32, 1, 1);
  UGN_TENSOR(g, DT_F32, 1, 1);
  UGN_TENSOR(v, DT_F32, 1, 1);
  if (opt == 0) { UGN_CHECK(m, "adam: m required"); UGN_TENSOR(m, DT_F32, 1, 1); }
  UGN_TENSOR(seg_off, DT_I64, 1, 1);
  UGN_TENSOR(seg_l2, DT_F32, 1, 1);
  long long n = w->shape[0];
  int S = (int)seg_l2->shape[0];
  UGN_CHECK(g->shape[0] == n && v->shape[0] == n && (!m || m->shape[0] == n), "optimizer: arena length mismatch");
  UGN_CHECK(seg_off->shape[0] == S + 1 && S >= 1, "optimizer: seg_off must be i64[S+1]");
  if (reg_out) UGN_TENSOR(reg_out, DT_F32, 1, 1);
  if (pack_table) {
    UGN_TENSOR(pack_table, DT_I64, 2, 2);
    UGN_CHECK(pack_table->shape[0] == S && pack_table->shape[1] == 2, "optimizer: pack_table must be i64 [S,2]");
    UGN_CHECK(pack_planes == 1 || pack_planes == 2, "optimizer: pack_planes must be 1 or 2");
  }
  if (n == 0) return UGN_OK;
  return ew_optim(ctx, opt, ugn_ptr<float>(w), ugn_ptr<float>(g), m ? ugn_ptr<float>(m) : nullptr, ugn_ptr<float>(v),
                  ugn_ptr<long long>(seg_off), ugn_ptr<float>(seg_l2), S, n, lr, b1, b2, eps, gscale,
                  reg_out ? ugn_ptr<float>(reg_out) : nullptr, lr_dev ? ugn_ptr<float>(lr_dev) : nullptr,
                  pack_table ? ugn_ptr<long long>(pack_table) : nullptr, pack_planes, pack_f16,
                  vhat ? ugn_ptr<float>(vhat) : nullptr, wd, (cudaStream_t)stream, world, rank,
                  reinterpret_cast<const long long*>(g_peers), reinterpret_cast<const long long*>(w_peers), (long long)g_mc,
                  (long long)w_mc, reinterpret_cast<const long long*>(reg_peers), stage ? ugn_ptr<float>(stage) : nullptr,
                  reinterpret_cast<const long long*>(staged_ranges), n_ranges,
                  reinterpret_cast<const long long*>(cw_peers), (long long)cw_mc);
}

extern "C" int ugn_adam_step(ugn_ctx* ctx, ugn_tensor* w, const ugn_tensor* g, ugn_tensor* m, ugn_tensor* v,
                             const ugn_tensor* seg_off, const ugn_tensor* seg_l2, float lr_t, float beta1, float beta2,
                             float eps, float gscale, ugn_tensor* reg_out, const ugn_tensor* lr_dev,
                             const ugn_tensor* pack_table, int pack_planes, int pack_f16, void* stream) {
  return optim_common(ctx, 0, w, g, m, v, seg_off, seg_l2, lr_t, beta1, beta2, eps, gscale, reg_out, lr_dev, pack_table,
                      pack_planes, pack_f16, stream);
}
extern "C" int ugn_adam_step_ex(ugn_ctx* ctx, ugn_tensor* w, const ugn_tensor* g, ugn_tensor* m, ugn_tensor* v,
                                ugn_tensor* vhat, float weight_decay, const ugn_tensor* seg_off, const ugn_tensor* seg_l2,
                                float lr_t, float beta1, float beta2, float eps, float gscale, ugn_tensor* reg_out,
                                const ugn_tensor* lr_dev, const ugn_tensor* pack_table, int pack_planes, int pack_f16,
                                void* stream) {
  return optim_common(ctx, 0, w, g, m, v, seg_off, seg_l2, lr_t, beta1, beta2, eps, gscale, reg_out, lr_dev, pack_table,
                      pack_planes, pack_f16, stream, vhat, weight_decay);
}
// a10: data-parallel step with the gradient exchange fused into the optimiser (see the header)
extern "C" int ugn_dp_optim_step(ugn_ctx* ctx, int opt, int world, int rank, const int64_t* g_peers, const int64_t* w_peers,
                                 int64_t g_multicast, int64_t w_multicast, ugn_tensor* w, const ugn_tensor* g, ugn_tensor* m, ugn_tensor* v, ugn_tensor* vhat,
                                 float weight_decay, const ugn_tensor* seg_off, const ugn_tensor* seg_l2, float beta1,
                                 float beta2, float eps, ugn_tensor* reg_out, const int64_t* reg_peers, const ugn_tensor* lr_dev,
                                 const ugn_tensor* stage, const int64_t* staged_ranges, int n_ranges,
                                 const ugn_tensor* pack_table, int pack_planes, int pack_f16, const int64_t* cw_peers,
                                 int64_t cw_multicast, void* stream) {
  UGN_CHECK(opt == 0 || opt == 1, "ugn_dp_optim_step: opt must be 0 (Adam family) or 1 (SGD momentum)");
  UGN_CHECK(world >= 2 && world <= 8 && g_peers && w_peers && lr_dev, "ugn_dp_optim_step: world in [2,8], peer tables and lr_dev required");
  return optim_common(ctx, opt, w, g, opt == 0 ? m : nullptr, v, seg_off, seg_l2, 0.f, beta1, beta2, eps, 1.f / (float)world,
                      reg_out, lr_dev, pack_table, pack_table ? pack_planes : 1, pack_f16, stream, opt == 0 ? vhat : nullptr,
                      opt == 0 ? weight_decay : 0.f, world, rank, g_peers, w_peers, g_multicast, w_multicast, reg_peers,
                      stage, staged_ranges, n_ranges, cw_peers, cw_multicast);
}
extern "C" int ugn_sgd_step(ugn_ctx* ctx, ugn_tensor* w, const ugn_tensor* g, ugn_tensor* v, const ugn_tensor* seg_off,
                            const ugn_tensor* seg_l2, float lr, float momentum, float gscale, ugn_tensor* reg_out,
                            const ugn_tensor* lr_dev, const ugn_tensor* pack_table, int pack_planes, int pack_f16,
                            void* stream) {
  return optim_common(ctx, 1, w, g, nullptr, v, seg_off, seg_l2, lr, momentum, 0.f, 0.f, gscale, reg_out, lr_dev,
                      pack_table, pack_planes, pack_f16, stream);
}

extern "C" int ugn_gemm_bf16(ugn_ctx* ctx, const ugn_tensor* A, int a_mn, const ugn_tensor* B, int b_mn, ugn_tensor* C,
                             int accumulate, void* stream) {
  UGN_CHECK(ctx && A && B && C, "ugn_gemm_bf16: null argument");
  UGN_TENSOR(A, DT_BAD, 3, 3);
  UGN_TENSOR(B, DT_BAD, 3, 3);
  UGN_CHECK(is_16(A), "gemm_bf16: operands must be bf16 or f16");
  UGN_SAME_FMT(A, B, "gemm_bf16");
  UGN_TENSOR(C, DT_F32, 2, 2);
  int P = (int)A->shape[0];
  UGN_CHECK((P == 1 || P == 2) && B->shape[0] == P, "gemm_bf16: plane counts must match (1 or 2)");
  int M = (int)(a_mn ? A->shape[2] : A->shape[1]), K = (int)(a_mn ? A->shape[1] : A->shape[2]);
  int N = (int)(b_mn ? B->shape[2] : B->shape[1]), Kb = (int)(b_mn ? B->shape[1] : B->shape[2]);
  UGN_CHECK(K == Kb && C->shape[0] == M && C->shape[1] == N, "gemm_bf16: shape mismatch (M=%d N=%d K=%d/%d)", M, N, K, Kb);
  return tc_gemm(ctx, P, is_f16(A), M, N, K, ugn_ptr<__nv_bfloat16>(A), a_mn, ugn_ptr<__nv_bfloat16>(B), b_mn, ugn_ptr<float>(C),
                 accumulate, (cudaStream_t)stream);
}

// ---- gradient scale for 16-bit gradient operands ---------------------------------------------
extern "C" int ugn_grad_scale_update(ugn_ctx* ctx, const ugn_tensor* ref, float target, void* stream) {
  UGN_CHECK(ctx && ref, "ugn_grad_scale_update: null argument");
  UGN_TENSOR(ref, DT_F32, 1, 4);
  UGN_CHECK(target > 0.f, "ugn_grad_scale_update: target must be positive");
  return ew_gscale_update(ctx, ugn_ptr<float>(ref), ugn_numel(ref), target, 0.f, (cudaStream_t)stream);
}
extern "C" int ugn_grad_scale_set(ugn_ctx* ctx, float scale, void* stream) {
  UGN_CHECK(ctx && scale > 0.f, "ugn_grad_scale_set: scale must be positive");
  return ew_gscale_update(ctx, nullptr, 0, 1.f, scale, (cudaStream_t)stream);
}

// ---- f32 column sums (bias gradient of a Dense layer from the UNROUNDED f32 output gradient) ----
extern "C" int ugn_colsum(ugn_ctx* ctx, const ugn_tensor* x, ugn_tensor* out, void* stream) {
  UGN_CHECK(ctx && x && out, "ugn_colsum: null argument");
  UGN_TENSOR(x, DT_F32, 2, 2);
  UGN_TENSOR(out, DT_F32, 1, 1);
  UGN_CHECK(out->shape[0] == x->shape[1], "ugn_colsum: out must be [cols]");
  return simt_colsum(ctx, ugn_ptr<float>(x), x->shape[0], (int)x->shape[1], (int)x->shape[1], ugn_ptr<float>(out),
                     (cudaStream_t)stream);
}

// ---- MMA passes of the forward tensor-core layers over hi/lo split operands --------------------
extern "C" int ugn_set_fwd_passes(ugn_ctx* ctx, int conv_passes, int dense_passes) {
  UGN_CHECK(ctx, "ugn_set_fwd_passes: null ctx");
  UGN_CHECK((conv_passes >= 0 && conv_passes <= 4) && (dense_passes >= 0 && dense_passes <= 4),
            "ugn_set_fwd_passes: pass codes are 0 (all three) | 1 | 2 | 3 | 4");
  ctx->fwd_conv_pass = conv_passes;
  ctx->fwd_dense_pass = dense_passes;
  return UGN_OK;
}

// ---- a13: video-level pooling and vote --------------------------------------------------------
int ew_segment_pool(ugn_ctx*, const float*, const int*, const int*, int, int, int, float*, cudaStream_t);
int ew_segment_mode(ugn_ctx*, const int*, const int*, const int*, int, int, int*, cudaStream_t);

static int segment_common(ugn_ctx* ctx, const ugn_tensor* order, const ugn_tensor* offsets, long long N, int& V) {
  UGN_TENSOR(order, DT_I32, 1, 1);
  UGN_TENSOR(offsets, DT_I32, 1, 1);
  UGN_CHECK(order->shape[0] == N && offsets->shape[0] >= 1, "segment: order must be i32 [N], offsets i32 [V+1]");
  V = (int)offsets->shape[0] - 1;
  return UGN_OK;
}
extern "C" int ugn_segment_pool(ugn_ctx* ctx, const ugn_tensor* codes, const ugn_tensor* order,
                                const ugn_tensor* offsets, int use_avg, ugn_tensor* out, void* stream) {
  UGN_CHECK(ctx && codes && order && offsets && out, "ugn_segment_pool: null argument");
  UGN_TENSOR(codes, DT_F32, 2, 2);
  UGN_TENSOR(out, DT_F32, 2, 2);
  int V = 0;
  int rc = segment_common(ctx, order, offsets, codes->shape[0], V);
  if (rc != UGN_OK) return rc;
  UGN_CHECK(out->shape[0] == V && out->shape[1] == codes->shape[1], "segment_pool: out must be [V,D]");
  return ew_segment_pool(ctx, ugn_ptr<float>(codes), ugn_ptr<int>(order), ugn_ptr<int>(offsets), V,
                         (int)codes->shape[1], use_avg, ugn_ptr<float>(out), (cudaStream_t)stream);
}
extern "C" int ugn_segment_mode(ugn_ctx* ctx, const ugn_tensor* labels, const ugn_tensor* order,
                                const ugn_tensor* offsets, int legacy_ties, ugn_tensor* out, void* stream) {
  UGN_CHECK(ctx && labels && order && offsets && out, "ugn_segment_mode: null argument");
  UGN_TENSOR(labels, DT_I32, 1, 1);
  UGN_TENSOR(out, DT_I32, 1, 1);
  int V = 0;
  int rc = segment_common(ctx, order, offsets, labels->shape[0], V);
  if (rc != UGN_OK) return rc;
  UGN_CHECK(out->shape[0] == V, "segment_mode: out must be i32 [V]");
  return ew_segment_mode(ctx, ugn_ptr<int>(labels), ugn_ptr<int>(order), ugn_ptr<int>(offsets), V, legacy_ties,
                         ugn_ptr<int>(out), (cudaStream_t)stream);
}
