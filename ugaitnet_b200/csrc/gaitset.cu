// GaitSet branch type: the kernels of UWYHSemiNet.build_gaitset_branch
// (/root/reference/nets/mj_uwyhNets_ba.py:420-484) that are not convolutions, and the rank-3
// gate / fusion / l2_normalize of the 3-modality graph around it (:1163-1191), with their C-ABI entry
// points.  All of them are HBM-bound streaming kernels (coalesced along the channel axis); the
// convolutions themselves run on ugn_conv2d_* (tcgen05 or fp32 validation mode) over zero-bordered
// activation buffers, which is how padding='same' is realised here.
//
// Storage modes as everywhere in the library: f32, or 16-bit planes [P][...] (P == 2: hi + lo).
#include "simt.cuh"

static inline int gs_mode(const ugn_tensor* t, int rank) {
  UgnDType dt = ugn_dtype(t);
  if (dt == DT_F32 && t->ndim == rank) return 0;
  if ((dt == DT_BF16 || dt == DT_F16) && t->ndim == rank + 1 && (t->shape[0] == 1 || t->shape[0] == 2))
    return (int)t->shape[0];
  return -1;
}
static inline const int64_t* gs_shape(const ugn_tensor* t, int rank) { return t->shape + (t->ndim - rank); }
static inline int gs_f16(const ugn_tensor* t) { return ugn_dtype(t) == DT_F16; }
static inline int gs_grid(ugn_ctx* ctx, long long items, int block) {
  long long g = (items + block - 1) / block;
  long long cap = (long long)ctx->sm_count * 16;
  return (int)std::max<long long>(1, std::min(g, cap));
}

__device__ __forceinline__ float gs_load(const void* p, int mode, int f16, long long plane, long long i) {
  if (mode == 0) return reinterpret_cast<const float*>(p)[i];
  const u16* q = reinterpret_cast<const u16*>(p);
  float v = ugn_f16to32(q[i], f16);
  if (mode == 2) v += ugn_f16to32(q[plane + i], f16);
  return v;
}
__device__ __forceinline__ void gs_store(void* p, int mode, int f16, long long plane, long long i, float v) {
  if (mode == 0) { reinterpret_cast<float*>(p)[i] = v; return; }
  u16* q = reinterpret_cast<u16*>(p);
  u16 hi = ugn_cvt16(v, f16);
  q[i] = hi;
  if (mode == 2) q[plane + i] = ugn_cvt16(v - ugn_f16to32(hi, f16), f16);
}

// ---------------------------------------------------------------------------------------------------
// input pack: x f32 [B,T,H,W,c] (the Keras gaitset input) -> im2col of ZeroPadding2D(2) + the 5x5 'same'
// convolution: out [.,B*T,H+4,W+4,Kp], channel j = (ky*5 + kx)*c + ci holds x[y+ky-4, x+kx-4, ci]
// (0 outside the frame and for j >= 25c), so that conv "a1" is a 1x1 convolution with K = Kp.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gs_pack_kernel(const float* __restrict__ x, void* __restrict__ out, long long F,
                                                      int H, int W, int c, int Kp, int mode, int f16) {
  const int Hs = H + 4, Ws = W + 4;
  const long long total = F * Hs * Ws * Kp, plane = total;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int j = (int)(i % Kp);
    long long pix = i / Kp;
    int xo = (int)(pix % Ws);
    int yo = (int)((pix / Ws) % Hs);
    long long f = pix / ((long long)Ws * Hs);
    float v = 0.f;
    if (j < 25 * c) {
      int ci = j % c, tap = j / c;
      int yy = yo + tap / 5 - 4, xx = xo + tap % 5 - 4;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[((f * H + yy) * W + xx) * c + ci];
    }
    gs_store(out, mode, f16, plane, i, v);
  }
}

extern "C" int ugn_gs_pack_input(ugn_ctx* ctx, const ugn_tensor* x, ugn_tensor* out, void* stream) {
  UGN_CHECK(ctx && x && out, "ugn_gs_pack_input: null argument");
  UGN_TENSOR(x, DT_F32, 5, 5);
  UGN_TENSOR(out, DT_BAD, 4, 5);
  int mode = gs_mode(out, 4);
  UGN_CHECK(mode >= 0, "gs_pack_input: out must be f32 [F,Hs,Ws,Kp] or 16-bit [P,F,Hs,Ws,Kp]");
  const int64_t* s = gs_shape(out, 4);
  long long F = x->shape[0] * x->shape[1];
  int H = (int)x->shape[2], W = (int)x->shape[3], c = (int)x->shape[4];
  UGN_CHECK(s[0] == F && s[1] == H + 4 && s[2] == W + 4 && s[3] >= 25 * c, "gs_pack_input: out must be [B*T,H+4,W+4,>=25c]");
  if (F == 0) return UGN_OK;
  long long total = F * s[1] * s[2] * s[3];
  gs_pack_kernel<<<gs_grid(ctx, total, 256), 256, 0, (cudaStream_t)stream>>>(ugn_ptr<float>(x), ugn_ptr<void>(out), F, H, W, c,
                                                                           (int)s[3], mode, gs_f16(out));
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------------------
// zero-border copies: padding='same' of the 3x3 convolutions.  The destination border is written once
// (zero) by the owner of the buffer; pad copies only the interior, crop reads only the interior.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gs_pad_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int P,
                                                     long long N, int H, int W, int rowv, int p) {
  const int Hd = H + 2 * p, Wd = W + 2 * p;
  const long long per_plane = N * H * W * rowv, total = per_plane * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int v = (int)(i % rowv);
    long long pix = i / rowv;
    int xx = (int)(pix % W);
    int yy = (int)((pix / W) % H);
    long long n = pix / ((long long)W * H);          // plane-major image index (pl*N + n)
    dst[((n * Hd + yy + p) * Wd + xx + p) * rowv + v] = src[i];
  }
}

extern "C" int ugn_pad_hw(ugn_ctx* ctx, const ugn_tensor* src, ugn_tensor* dst, void* stream) {
  UGN_CHECK(ctx && src && dst, "ugn_pad_hw: null argument");
  UGN_TENSOR(src, DT_BAD, 4, 5);
  UGN_TENSOR(dst, DT_BAD, 4, 5);
  int ms = gs_mode(src, 4), md = gs_mode(dst, 4);
  UGN_CHECK(ms >= 0 && ms == md && ugn_dtype(src) == ugn_dtype(dst), "pad_hw: storage modes must match");
  const int64_t* a = gs_shape(src, 4);
  const int64_t* b = gs_shape(dst, 4);
  int p = (int)(b[1] - a[1]) / 2;
  UGN_CHECK(a[0] == b[0] && a[3] == b[3] && p >= 0 && b[1] == a[1] + 2 * p && b[2] == a[2] + 2 * p, "pad_hw: dst must be [N,H+2p,W+2p,C]");
  int es = ms == 0 ? 4 : 2;
  UGN_CHECK((a[3] * es) % 16 == 0, "pad_hw: channel row must be a multiple of 16 bytes");
  if (ugn_numel(src) == 0) return UGN_OK;
  int rowv = (int)(a[3] * es / 16), P = ms == 0 ? 1 : ms;
  long long total = (long long)P * a[0] * a[1] * a[2] * rowv;
  gs_pad_kernel<<<gs_grid(ctx, total, 256), 256, 0, (cudaStream_t)stream>>>(ugn_ptr<uint4>(src), ugn_ptr<uint4>(dst), P, a[0],
                                                                          (int)a[1], (int)a[2], rowv, p);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

__global__ void __launch_bounds__(256) gs_crop_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long long N,
                                                      int H, int W, int rowv, int p, int accumulate) {
  const int Hs = H + 2 * p, Ws = W + 2 * p;
  const long long total = N * H * W * rowv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int v = (int)(i % rowv);
    long long pix = i / rowv;
    int xx = (int)(pix % W);
    int yy = (int)((pix / W) % H);
    long long n = pix / ((long long)W * H);
    float4 s = src[((n * Hs + yy + p) * Ws + xx + p) * rowv + v];
    if (accumulate) {
      float4 d = dst[i];
      s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
    }
    dst[i] = s;
  }
}

extern "C" int ugn_crop_hw(ugn_ctx* ctx, const ugn_tensor* src, ugn_tensor* dst, int accumulate, void* stream) {
  UGN_CHECK(ctx && src && dst, "ugn_crop_hw: null argument");
  UGN_TENSOR(src, DT_F32, 4, 4);
  UGN_TENSOR(dst, DT_F32, 4, 4);
  const int64_t* a = src->shape;
  const int64_t* b = dst->shape;
  int p = (int)(a[1] - b[1]) / 2;
  UGN_CHECK(a[0] == b[0] && a[3] == b[3] && p >= 0 && a[1] == b[1] + 2 * p && a[2] == b[2] + 2 * p && b[3] % 4 == 0,
            "crop_hw: src must be [N,H+2p,W+2p,C]");
  if (ugn_numel(dst) == 0) return UGN_OK;
  int rowv = (int)b[3] / 4;
  long long total = b[0] * b[1] * b[2] * rowv;
  gs_crop_kernel<<<gs_grid(ctx, total, 256), 256, 0, (cudaStream_t)stream>>>(ugn_ptr<float4>(src), ugn_ptr<float4>(dst), b[0],
                                                                           (int)b[1], (int)b[2], rowv, p, accumulate);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------------------
// set pooling: Lambda(reduce_max(x, axis=1)) over the T frames of a sequence (:435,:454,:465) fused with
// the layers.Add() that follows it (:455,:466).  m = max_t a[b,t]; y = m + addend.
// Backward: tf reduce_max splits the gradient evenly among the frames that attain the maximum.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gs_setmax_fwd_kernel(const void* __restrict__ a, int amode, long long aplane,
                                                            const void* __restrict__ addend, int dmode, long long dplane,
                                                            float* __restrict__ m, void* __restrict__ y, int ymode,
                                                            long long yplane, int B, int T, long long Q, int f16) {
  const long long total = (long long)B * Q;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long b = i / Q, q = i - b * Q;
    float best = gs_load(a, amode, f16, aplane, (b * T) * Q + q);
    for (int t = 1; t < T; ++t) best = fmaxf(best, gs_load(a, amode, f16, aplane, (b * T + t) * Q + q));
    if (m) m[i] = best;
    if (y) {
      float v = best;
      if (addend) v += gs_load(addend, dmode, f16, dplane, i);
      gs_store(y, ymode, f16, yplane, i, v);
    }
  }
}

extern "C" int ugn_setmax_fwd(ugn_ctx* ctx, const ugn_tensor* a, int T, const ugn_tensor* addend, ugn_tensor* m,
                              ugn_tensor* y, void* stream) {
  UGN_CHECK(ctx && a && T > 0 && (m || y), "ugn_setmax_fwd: null argument");
  UGN_TENSOR(a, DT_BAD, 4, 5);
  int am = gs_mode(a, 4);
  UGN_CHECK(am >= 0, "setmax_fwd: bad storage mode of a");
  const int64_t* s = gs_shape(a, 4);
  UGN_CHECK(s[0] % T == 0, "setmax_fwd: leading dim %lld is not a multiple of T=%d", (long long)s[0], T);
  int B = (int)(s[0] / T);
  long long Q = s[1] * s[2] * s[3];
  int f16 = gs_f16(a), dm = 0, ym = 0;
  if (addend) {
    UGN_TENSOR(addend, DT_BAD, 4, 5);
    dm = gs_mode(addend, 4);
    UGN_CHECK(dm >= 0 && ugn_numel(addend) == (dm ? dm : 1) * B * Q, "setmax_fwd: addend shape mismatch");
    if (dm) { UGN_CHECK(!am || gs_f16(addend) == f16, "setmax_fwd: 16-bit formats differ"); f16 = gs_f16(addend); }
  }
  if (m) { UGN_TENSOR(m, DT_F32, 4, 4); UGN_CHECK(ugn_numel(m) == B * Q, "setmax_fwd: m must be f32 [B,H,W,C]"); }
  if (y) {
    UGN_TENSOR(y, DT_BAD, 4, 5);
    ym = gs_mode(y, 4);
    UGN_CHECK(ym >= 0 && ugn_numel(y) == (ym ? ym : 1) * B * Q, "setmax_fwd: y shape mismatch");
    if (ym) { UGN_CHECK((!am && !dm) || gs_f16(y) == f16, "setmax_fwd: 16-bit formats differ"); f16 = gs_f16(y); }
  }
  if (B == 0 || Q == 0) return UGN_OK;
  gs_setmax_fwd_kernel<<<gs_grid(ctx, B * Q, 256), 256, 0, (cudaStream_t)stream>>>(
      ugn_ptr<void>(a), am, s[0] * Q, addend ? ugn_ptr<void>(addend) : nullptr, dm, B * Q, m ? ugn_ptr<float>(m) : nullptr,
      y ? ugn_ptr<void>(y) : nullptr, ym, B * Q, B, T, Q, f16);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

__global__ void __launch_bounds__(256) gs_setmax_bwd_kernel(const float* __restrict__ dm, const void* __restrict__ a,
                                                            int amode, long long aplane, const float* __restrict__ m,
                                                            float* __restrict__ da, int B, int T, long long Q, int f16,
                                                            int accumulate) {
  const long long total = (long long)B * Q;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long b = i / Q, q = i - b * Q;
    const float mx = m[i];
    int cnt = 0;
    for (int t = 0; t < T; ++t) cnt += gs_load(a, amode, f16, aplane, (b * T + t) * Q + q) == mx;
    const float g = dm[i] / (float)max(cnt, 1);
    for (int t = 0; t < T; ++t) {
      long long o = (b * T + t) * Q + q;
      float v = gs_load(a, amode, f16, aplane, o) == mx ? g : 0.f;
      da[o] = accumulate ? da[o] + v : v;
    }
  }
}

extern "C" int ugn_setmax_bwd(ugn_ctx* ctx, const ugn_tensor* dm, const ugn_tensor* a, const ugn_tensor* m, int T,
                              ugn_tensor* da, int accumulate, void* stream) {
  UGN_CHECK(ctx && dm && a && m && da && T > 0, "ugn_setmax_bwd: null argument");
  UGN_TENSOR(dm, DT_F32, 4, 4);
  UGN_TENSOR(m, DT_F32, 4, 4);
  UGN_TENSOR(da, DT_F32, 4, 4);
  UGN_TENSOR(a, DT_BAD, 4, 5);
  int am = gs_mode(a, 4);
  UGN_CHECK(am >= 0, "setmax_bwd: bad storage mode of a");
  const int64_t* s = gs_shape(a, 4);
  UGN_CHECK(s[0] % T == 0, "setmax_bwd: leading dim is not a multiple of T");
  int B = (int)(s[0] / T);
  long long Q = s[1] * s[2] * s[3];
  UGN_CHECK(ugn_numel(dm) == B * Q && ugn_numel(m) == B * Q && ugn_numel(da) == s[0] * Q, "setmax_bwd: shape mismatch");
  if (B == 0 || Q == 0) return UGN_OK;
  gs_setmax_bwd_kernel<<<gs_grid(ctx, B * Q, 256), 256, 0, (cudaStream_t)stream>>>(
      ugn_ptr<float>(dm), ugn_ptr<void>(a), am, s[0] * Q, ugn_ptr<float>(m), ugn_ptr<float>(da), B, T, Q, gs_f16(a), accumulate);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------------------
// Horizontal pyramid pooling (:468-479): the S = H*W positions of an NHWC map, taken in row-major order,
// are cut into nb in {1,2,4,8,16} strips; feature = mean + max per strip.  Parts are laid out as the
// reference concatenates them: for each nb, nb strips of the set-level map ("which" 0) then nb strips of
// the global map ("which" 1): part = 2*(nb-1) + which*nb + strip.  out f32 [62,B,C].
// One thread per (b, c): positions are C floats apart, so a warp reads 128-byte rows.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gs_hpp_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int S,
                                                         int C, int which) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  const float* xp = x + (long long)b * S * C + c;
  const int L = S / 16;
  float s16[16], m16[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float s = 0.f, mx = -INFINITY;
    for (int q = 0; q < L; ++q) {
      float v = xp[(long long)(j * L + q) * C];
      s += v;
      mx = fmaxf(mx, v);
    }
    s16[j] = s; m16[j] = mx;
  }
#pragma unroll
  for (int lv = 0; lv < 5; ++lv) {
    const int nb = 1 << lv, w = 16 / nb;
#pragma unroll
    for (int k = 0; k < nb; ++k) {
      float s = 0.f, mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < w; ++j) { s += s16[k * w + j]; mx = fmaxf(mx, m16[k * w + j]); }
      const int part = 2 * (nb - 1) + which * nb + k;
      out[((long long)part * B + b) * C + c] = s / (float)(w * L) + mx;
    }
  }
}

// dx[b,pos,c] (+)= sum over the 5 levels of dfeat[part]/len + dfeat[part] * [x == strip max] / #ties
__global__ void __launch_bounds__(128) gs_hpp_bwd_kernel(const float* __restrict__ dfeat, const float* __restrict__ x,
                                                         float* __restrict__ dx, int B, int S, int C, int which,
                                                         int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  const float* xp = x + (long long)b * S * C + c;
  float* dp = dx + (long long)b * S * C + c;
  const int L = S / 16;
  float m16[16];
  int c16[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float mx = -INFINITY;
    int cn = 0;
    for (int q = 0; q < L; ++q) {
      float v = xp[(long long)(j * L + q) * C];
      if (v > mx) { mx = v; cn = 1; } else if (v == mx) ++cn;
    }
    m16[j] = mx; c16[j] = cn;
  }
  // per finest strip j: mean coefficient and, per level, (strip max, dfeat / #ties)
  float cmean[16], lmax[5][16], lg[5][16];
#pragma unroll
  for (int j = 0; j < 16; ++j) cmean[j] = 0.f;
#pragma unroll
  for (int lv = 0; lv < 5; ++lv) {
    const int nb = 1 << lv, w = 16 / nb;
#pragma unroll
    for (int k = 0; k < nb; ++k) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < w; ++j) mx = fmaxf(mx, m16[k * w + j]);
      int cn = 0;
#pragma unroll
      for (int j = 0; j < w; ++j) cn += (m16[k * w + j] == mx) ? c16[k * w + j] : 0;
      const int part = 2 * (nb - 1) + which * nb + k;
      const float g = dfeat[((long long)part * B + b) * C + c];
#pragma unroll
      for (int j = 0; j < w; ++j) {
        cmean[k * w + j] += g / (float)(w * L);
        lmax[lv][k * w + j] = mx;
        lg[lv][k * w + j] = g / (float)max(cn, 1);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    for (int q = 0; q < L; ++q) {
      const long long o = (long long)(j * L + q) * C;
      const float v = xp[o];
      float g = cmean[j];
#pragma unroll
      for (int lv = 0; lv < 5; ++lv) g += (v == lmax[lv][j]) ? lg[lv][j] : 0.f;
      dp[o] = accumulate ? dp[o] + g : g;
    }
  }
}

static int hpp_args(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* feat, int which, int& B, int& S, int& C) {
  UGN_TENSOR(x, DT_F32, 4, 4);
  UGN_TENSOR(feat, DT_F32, 3, 3);
  B = (int)x->shape[0]; S = (int)(x->shape[1] * x->shape[2]); C = (int)x->shape[3];
  UGN_CHECK(S % 16 == 0, "hpp: H*W = %d must be a multiple of 16", S);
  UGN_CHECK(feat->shape[0] == 62 && feat->shape[1] == B && feat->shape[2] == C, "hpp: feat must be f32 [62,B,C]");
  UGN_CHECK(which == 0 || which == 1, "hpp: which must be 0 (set-level map) or 1 (global map)");
  return UGN_OK;
}

extern "C" int ugn_hpp_fwd(ugn_ctx* ctx, const ugn_tensor* x, int which, ugn_tensor* feat, void* stream) {
  UGN_CHECK(ctx && x && feat, "ugn_hpp_fwd: null argument");
  int B, S, C, rc;
  if ((rc = hpp_args(ctx, x, feat, which, B, S, C)) != UGN_OK) return rc;
  if (B == 0) return UGN_OK;
  gs_hpp_fwd_kernel<<<ugn_cdiv((long long)B * C, 128), 128, 0, (cudaStream_t)stream>>>(ugn_ptr<float>(x), ugn_ptr<float>(feat), B, S, C,
                                                                                     which);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

extern "C" int ugn_hpp_bwd(ugn_ctx* ctx, const ugn_tensor* dfeat, const ugn_tensor* x, int which, ugn_tensor* dx,
                           int accumulate, void* stream) {
  UGN_CHECK(ctx && dfeat && x && dx, "ugn_hpp_bwd: null argument");
  int B, S, C, rc;
  if ((rc = hpp_args(ctx, x, dfeat, which, B, S, C)) != UGN_OK) return rc;
  UGN_TENSOR(dx, DT_F32, 4, 4);
  UGN_CHECK(ugn_numel(dx) == ugn_numel(x), "hpp_bwd: dx must have the shape of x");
  if (B == 0) return UGN_OK;
  gs_hpp_bwd_kernel<<<ugn_cdiv((long long)B * C, 128), 128, 0, (cudaStream_t)stream>>>(ugn_ptr<float>(dfeat), ugn_ptr<float>(x),
                                                                                     ugn_ptr<float>(dx), B, S, C, which, accumulate);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------------------
// MatMul layer (:23-48): out[n] = x[n] . kernel[n], n = 62 parts.  Small (0.4 GFLOP at B = 96): fp32 FFMA.
//   C f32 [n,M,N] = op(A)[n] . op(B)[n];  a_t == 0: A is [n,M,K], a_t == 1: A is [n,K,M];
//   b_t == 0: B is [n,K,N], b_t == 1: B is [n,N,K].
// ---------------------------------------------------------------------------------------------------
extern "C" int ugn_bmm_f32(ugn_ctx* ctx, const ugn_tensor* A, int a_t, const ugn_tensor* Bm, int b_t, ugn_tensor* C,
                           void* stream) {
  UGN_CHECK(ctx && A && Bm && C, "ugn_bmm_f32: null argument");
  UGN_TENSOR(A, DT_F32, 3, 3);
  UGN_TENSOR(Bm, DT_F32, 3, 3);
  UGN_TENSOR(C, DT_F32, 3, 3);
  int n = (int)C->shape[0], M = (int)C->shape[1], N = (int)C->shape[2];
  int K = (int)(a_t ? A->shape[1] : A->shape[2]);
  UGN_CHECK(A->shape[0] == n && Bm->shape[0] == n && (a_t ? A->shape[2] : A->shape[1]) == M &&
                (b_t ? Bm->shape[1] : Bm->shape[2]) == N && (b_t ? Bm->shape[2] : Bm->shape[1]) == K,
            "bmm_f32: shape mismatch");
  if (n == 0 || M == 0 || N == 0) return UGN_OK;
  SGemm p;
  p.A = ugn_ptr<float>(A); p.B = ugn_ptr<float>(Bm); p.C = ugn_ptr<float>(C);
  p.M = M; p.N = N; p.K = K;
  p.ar = radix1(a_t ? 1 : K); p.ak = radix1(a_t ? M : 1);
  p.br = radix1(b_t ? K : 1); p.bk = radix1(b_t ? 1 : N);
  p.ldc = N;
  p.batches = n; p.batch_a = (long long)M * K; p.batch_b = (long long)K * N; p.batch_c = (long long)M * N;
  if (n == 1) { p.batches = 1; }
  return simt_gemm_launch(ctx, p, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------
// rank-3 gate x use-flag (:51-54), fusion (:1189) and tf.math.l2_normalize(axis=1) (:1191) on
// [n, B, d] tensors: axis 1 is the BATCH axis in this layout, so every (part, feature) column is
// normalised over the rows of the batch -- kept literally.  One thread per column (n, j), coalesced
// over j; two sweeps over b.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gs_fuse3_fwd_kernel(FusePtrs ptrs, int nmods, int n, int B, int d,
                                                           float* __restrict__ sig, uint8_t* __restrict__ winner,
                                                           float* __restrict__ col, int merge) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * d) return;
  const int part = i / d, j = i - part * d;
  float ss = 0.f;
  for (int b = 0; b < B; ++b) {
    const long long o = ((long long)part * B + b) * d + j;
    float best = ptrs.br[0][o] * ptrs.flag[0][b];
    int win = 0;
    if (merge == UGN_MERGE_AVG) {
      for (int m = 1; m < nmods; ++m) best += ptrs.br[m][o] * ptrs.flag[m][b];
      best = best / (float)nmods;
    } else {
      float key = merge == UGN_MERGE_SIGNMAX ? fabsf(best) : best;
      for (int m = 1; m < nmods; ++m) {
        float v = ptrs.br[m][o] * ptrs.flag[m][b];
        float kv = merge == UGN_MERGE_SIGNMAX ? fabsf(v) : v;
        if (kv > key) { key = kv; best = v; win = m; }
      }
    }
    sig[o] = best;
    winner[o] = (uint8_t)win;
    ss += best * best;
  }
  const float inv = rsqrtf(fmaxf(ss, 1e-12f));
  col[2 * i] = inv;
  col[2 * i + 1] = ss;
  for (int b = 0; b < B; ++b) {
    const long long o = ((long long)part * B + b) * d + j;
    sig[o] *= inv;
  }
}

__global__ void __launch_bounds__(128) gs_fuse3_bwd_kernel(FusePtrs ptrs, int nmods, int n, int B, int d,
                                                           const float* __restrict__ dsig, const float* __restrict__ sig,
                                                           const uint8_t* __restrict__ winner,
                                                           const float* __restrict__ col, int merge) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * d) return;
  const int part = i / d, j = i - part * d;
  const float inv = col[2 * i];
  const bool clamped = !(col[2 * i + 1] > 1e-12f);
  float dot = 0.f;
  if (!clamped)
    for (int b = 0; b < B; ++b) {
      const long long o = ((long long)part * B + b) * d + j;
      dot += dsig[o] * sig[o];
    }
  for (int b = 0; b < B; ++b) {
    const long long o = ((long long)part * B + b) * d + j;
    float g = dsig[o];
    g = clamped ? g * inv : inv * (g - sig[o] * dot);
    if (merge == UGN_MERGE_AVG) {
      for (int m = 0; m < nmods; ++m) ptrs.dbr[m][o] = g * ptrs.flag[m][b] / (float)nmods;
    } else {
      const int w = winner[o];
      for (int m = 0; m < nmods; ++m) ptrs.dbr[m][o] = (m == w) ? g * ptrs.flag[m][b] : 0.f;
    }
  }
}

static int fuse3_ptrs(ugn_ctx* ctx, int nmods, const ugn_tensor* const* br, const ugn_tensor* const* flags,
                      ugn_tensor* const* dbr, int n, int B, int d, FusePtrs& P) {
  UGN_CHECK(nmods >= 1 && nmods <= 4, "fuse3: 1..4 modalities");
  for (int m = 0; m < nmods; ++m) {
    if (br) {
      UGN_TENSOR(br[m], DT_F32, 3, 3);
      UGN_CHECK(br[m]->shape[0] == n && br[m]->shape[1] == B && br[m]->shape[2] == d, "fuse3: branch shape mismatch");
      P.br[m] = ugn_ptr<float>(br[m]);
    }
    UGN_TENSOR(flags[m], DT_F32, 1, 2);
    UGN_CHECK(ugn_numel(flags[m]) == B, "fuse3: flags must be f32 [B,1]");
    P.flag[m] = ugn_ptr<float>(flags[m]);
    if (dbr) {
      UGN_TENSOR(dbr[m], DT_F32, 3, 3);
      UGN_CHECK(ugn_numel(dbr[m]) == (int64_t)n * B * d, "fuse3: dbr shape mismatch");
      P.dbr[m] = ugn_ptr<float>(dbr[m]);
    }
  }
  return UGN_OK;
}

extern "C" int ugn_fuse3_fwd(ugn_ctx* ctx, int nmods, const ugn_tensor* const* br, const ugn_tensor* const* flags,
                             ugn_tensor* sig, ugn_tensor* winner, ugn_tensor* col_norm, int merge, void* stream) {
  UGN_CHECK(ctx && br && flags && sig && winner && col_norm, "ugn_fuse3_fwd: null argument");
  UGN_TENSOR(sig, DT_F32, 3, 3);
  UGN_TENSOR(winner, DT_U8, 3, 3);
  UGN_TENSOR(col_norm, DT_F32, 3, 3);
  int n = (int)sig->shape[0], B = (int)sig->shape[1], d = (int)sig->shape[2];
  UGN_CHECK(ugn_numel(winner) == ugn_numel(sig) && col_norm->shape[0] == n && col_norm->shape[1] == d && col_norm->shape[2] == 2,
            "fuse3_fwd: winner must be u8 [n,B,d], col_norm f32 [n,d,2]");
  FusePtrs P{};
  int rc = fuse3_ptrs(ctx, nmods, br, flags, nullptr, n, B, d, P);
  if (rc != UGN_OK) return rc;
  if (ugn_numel(sig) == 0) return UGN_OK;
  gs_fuse3_fwd_kernel<<<ugn_cdiv((long long)n * d, 128), 128, 0, (cudaStream_t)stream>>>(P, nmods, n, B, d, ugn_ptr<float>(sig),
                                                                                       ugn_ptr<uint8_t>(winner),
                                                                                       ugn_ptr<float>(col_norm), merge);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

extern "C" int ugn_fuse3_bwd(ugn_ctx* ctx, int nmods, const ugn_tensor* dsig, const ugn_tensor* sig, const ugn_tensor* winner,
                             const ugn_tensor* col_norm, const ugn_tensor* const* flags, ugn_tensor* const* dbr, int merge,
                             void* stream) {
  UGN_CHECK(ctx && dsig && sig && winner && col_norm && flags && dbr, "ugn_fuse3_bwd: null argument");
  UGN_TENSOR(dsig, DT_F32, 3, 3);
  UGN_TENSOR(sig, DT_F32, 3, 3);
  UGN_TENSOR(winner, DT_U8, 3, 3);
  UGN_TENSOR(col_norm, DT_F32, 3, 3);
  int n = (int)sig->shape[0], B = (int)sig->shape[1], d = (int)sig->shape[2];
  UGN_CHECK(ugn_numel(dsig) == ugn_numel(sig) && ugn_numel(winner) == ugn_numel(sig) && ugn_numel(col_norm) == 2LL * n * d,
            "fuse3_bwd: shape mismatch");
  FusePtrs P{};
  int rc = fuse3_ptrs(ctx, nmods, nullptr, flags, dbr, n, B, d, P);
  if (rc != UGN_OK) return rc;
  if (ugn_numel(sig) == 0) return UGN_OK;
  gs_fuse3_bwd_kernel<<<ugn_cdiv((long long)n * d, 128), 128, 0, (cudaStream_t)stream>>>(
      P, nmods, n, B, d, ugn_ptr<float>(dsig), ugn_ptr<float>(sig), ugn_ptr<uint8_t>(winner), ugn_ptr<float>(col_norm), merge);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------------------
// Lambda(tf.transpose(x, [1,0,2])) before Flatten + "classprob" (:1211-1213): [n,B,d] <-> [B,n,d]
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gs_permute102_kernel(const float* __restrict__ src, float* __restrict__ dst, int A,
                                                            int Bd, int d) {
  const long long total = (long long)A * Bd * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int j = (int)(i % d);
    long long r = i / d;
    int b = (int)(r % Bd);
    int a = (int)(r / Bd);
    dst[((long long)b * A + a) * d + j] = src[i];
  }
}

extern "C" int ugn_permute102(ugn_ctx* ctx, const ugn_tensor* src, ugn_tensor* dst, void* stream) {
  UGN_CHECK(ctx && src && dst, "ugn_permute102: null argument");
  UGN_TENSOR(src, DT_F32, 3, 3);
  UGN_TENSOR(dst, DT_F32, 2, 3);
  UGN_CHECK(ugn_numel(dst) == ugn_numel(src) && dst->shape[0] == src->shape[1], "permute102: dst must be [B,n,d] (or [B,n*d])");
  if (ugn_numel(src) == 0) return UGN_OK;
  gs_permute102_kernel<<<gs_grid(ctx, ugn_numel(src), 256), 256, 0, (cudaStream_t)stream>>>(
      ugn_ptr<float>(src), ugn_ptr<float>(dst), (int)src->shape[0], (int)src->shape[1], (int)src->shape[2]);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}
