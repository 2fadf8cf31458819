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
// First convolution of the branch, fused: TimeDistributed(ZeroPadding2D(2)) + Conv2D(32, 5x5, 'same',
// no bias) + LeakyReLU (:427-429) straight from the Keras input x f32 [B,T,H,W,c] (c = 1 | 2) into the
// ZERO-BORDERED input buffer of the next convolution, y [.,F*S][Hs+2][Hs/S+2][32] (Hs = H+4, S column
// halves, see ugn_pad_hw below).  K = 25c is far too small for the tensor cores (a 32-channel im2col would cost
// 2.5 GB of traffic per modality): this is an fp32 FFMA kernel, exact products, HBM-bound on the output.
// One thread = one output pixel x 32 channels; the 5x5 halo tile of the frame and the kernel live in smem.
// ---------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256, 2) gs_conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           void* __restrict__ y, long long F, int H, int W, int S,
                                                           int mode, int f16, float alpha) {
  constexpr int K = 25 * C;
  extern __shared__ float sm[];
  const int Ws = W + 4, Hs = H + 4;
  const int TW = Ws + 4;                       // tile row: output columns 0..Ws-1 need input columns -4..Ws-1
  float* ws = sm;                              // [K][32]
  float* xs = sm + K * 32;                     // [8][TW][C]
  const int rows_per_block = 256 / Ws > 0 ? 256 / Ws : 1;     // output rows per block (Ws <= 256)
  for (int i = threadIdx.x; i < K * 32; i += 256) ws[i] = w[(i & 31) * K + (i >> 5)];   // w [32][K] -> [K][32]
  const int tiles_y = (Hs + rows_per_block - 1) / rows_per_block;
  const long long ntiles = F * tiles_y;
  const int Wh = Ws / S + 2, Hp = Hs + 2;
  const long long plane = F * S * (long long)Hp * Wh * 32;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long f = tile / tiles_y;
    const int y0 = (int)(tile % tiles_y) * rows_per_block;
    __syncthreads();
    const int nin = (rows_per_block + 4) * TW * C;
    for (int i = threadIdx.x; i < nin; i += 256) {
      int ci = i % C, xx = (i / C) % TW, yy = i / (C * TW);
      int gy = y0 + yy - 4, gx = xx - 4;
      xs[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? x[((f * H + gy) * W + gx) * C + ci] : 0.f;
    }
    __syncthreads();
    const int ly = threadIdx.x / Ws, lx = threadIdx.x - ly * Ws;
    if (ly >= rows_per_block || y0 + ly >= Hs) continue;
    float acc[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = 0.f;
    for (int ky = 0; ky < 5; ++ky)
#pragma unroll
      for (int kx = 0; kx < 5; ++kx)
#pragma unroll
        for (int ci = 0; ci < C; ++ci) {
          const float xv = xs[((ly + ky) * TW + lx + kx) * C + ci];
          const float4* wr = reinterpret_cast<const float4*>(ws + ((ky * 5 + kx) * C + ci) * 32);
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const float4 w4 = wr[o];
            acc[4 * o] = fmaf(xv, w4.x, acc[4 * o]);
            acc[4 * o + 1] = fmaf(xv, w4.y, acc[4 * o + 1]);
            acc[4 * o + 2] = fmaf(xv, w4.z, acc[4 * o + 2]);
            acc[4 * o + 3] = fmaf(xv, w4.w, acc[4 * o + 3]);
          }
        }
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = acc[o] > 0.f ? acc[o] : alpha * acc[o];
    // padded coordinates (y+1, x+1); the pixel belongs to every half whose column window contains it
    const int yc = y0 + ly + 1, xc = lx + 1, wh = Ws / S;
    for (int h = 0; h < S; ++h) {
      const int xl = xc - h * wh;
      if (xl < 0 || xl >= Wh) continue;
      const long long o0 = (((f * S + h) * Hp + yc) * Wh + xl) * 32;
      if (mode == 0) {
        float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + o0);
#pragma unroll
        for (int o = 0; o < 8; ++o) d[o] = make_float4(acc[4 * o], acc[4 * o + 1], acc[4 * o + 2], acc[4 * o + 3]);
      } else {
        // 8 channels (one 16-byte store per plane) at a time: keeps the live registers under the 128 that two
        // resident blocks per SM allow
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int o = 0; o < 8; o += 2) {
            u16 h0, l0, h1, l1;
            ugn_split16(acc[8 * q + o], f16, h0, l0);
            ugn_split16(acc[8 * q + o + 1], f16, h1, l1);
            hw[o >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lw[o >> 1] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
          reinterpret_cast<uint4*>(reinterpret_cast<u16*>(y) + o0)[q] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          if (mode == 2)
            reinterpret_cast<uint4*>(reinterpret_cast<u16*>(y) + plane + o0)[q] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
    }
  }
}

static int conv1_geom(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* yp, long long& F, int& H, int& W, int& c, int& S) {
  UGN_TENSOR(x, DT_F32, 5, 5);
  F = x->shape[0] * x->shape[1];
  H = (int)x->shape[2]; W = (int)x->shape[3]; c = (int)x->shape[4];
  UGN_CHECK(c == 1 || c == 2, "gs_conv1: per-frame channels must be 1 or 2");
  UGN_CHECK(W + 4 <= 256, "gs_conv1: frame too wide");
  const int64_t* s = gs_shape(yp, 4);
  UGN_CHECK(F > 0 && s[0] % F == 0, "gs_conv1: padded buffer leading dim must be B*T*S");
  S = (int)(s[0] / F);
  UGN_CHECK(S >= 1 && (W + 4) % S == 0 && s[1] == H + 6 && s[2] == (W + 4) / S + 2 && s[3] == 32,
            "gs_conv1: padded buffer must be [B*T*S, H+6, (W+4)/S+2, 32]");
  return UGN_OK;
}

extern "C" int ugn_gs_conv1_fwd(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* w, ugn_tensor* yp, float alpha,
                                void* stream) {
  UGN_CHECK(ctx && x && w && yp, "ugn_gs_conv1_fwd: null argument");
  UGN_TENSOR(yp, DT_BAD, 4, 5);
  UGN_TENSOR(w, DT_F32, 2, 4);
  int mode = gs_mode(yp, 4);
  UGN_CHECK(mode >= 0, "gs_conv1_fwd: bad storage mode of the output");
  long long F; int H, W, c, S, rc;
  if ((rc = conv1_geom(ctx, x, yp, F, H, W, c, S)) != UGN_OK) return rc;
  UGN_CHECK(ugn_numel(w) == 32 * 25 * c, "gs_conv1_fwd: w must be f32 [32,1,1,25c] (tap-major ky,kx,ci)");
  const int Ws = W + 4, rows = std::max(1, 256 / Ws);
  const long long ntiles = F * ugn_cdiv(H + 4, rows);
  size_t smem = sizeof(float) * (25 * c * 32 + (size_t)(rows + 4) * (Ws + 4) * c);
  UGN_CHECK(smem <= 48 * 1024, "gs_conv1_fwd: frame too wide");
  int grid = (int)std::min<long long>(ntiles, (long long)ctx->sm_count * 8);
  if (c == 1)
    gs_conv1_fwd_kernel<1><<<grid, 256, smem, (cudaStream_t)stream>>>(ugn_ptr<float>(x), ugn_ptr<float>(w), ugn_ptr<void>(yp), F, H, W, S,
                                                                     mode, gs_f16(yp), alpha);
  else
    gs_conv1_fwd_kernel<2><<<grid, 256, smem, (cudaStream_t)stream>>>(ugn_ptr<float>(x), ugn_ptr<float>(w), ugn_ptr<void>(yp), F, H, W, S,
                                                                     mode, gs_f16(yp), alpha);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// Kernel gradient of that convolution, with the LeakyReLU derivative and the crop of the padded input
// gradient folded in: dz = dxp * (y > 0 ? 1 : alpha) over the interior of the (split) padded frame, where
// dxp f32 [F*S][Hs+2][Hs/S+2][32] is what ugn_conv2d_dgrad of the next layer wrote.  A pixel that lives in
// two halves is simply visited twice (the sum over halves commutes with the reduction over pixels).
// dw f32 [32][25c] (tap-major), OVERWRITTEN.
// Per tile of 4 output rows: (1) every (position, 4 channels) item is loaded ONCE (16-byte gradient load +
// sign of the hi plane: sign(y) == sign(hi)), multiplied by the LeakyReLU derivative and parked in smem;
// (2) thread = (4 output channels, 25c/16 taps) sweeps the parked positions: one LDS.128 + TT broadcast reads
// of the frame tile for 4*TT FMAs.  Persistent blocks keep their partial sums in registers and issue one
// round of atomics at the end.
template <int C, int MODE>
__global__ void __launch_bounds__(128) gs_conv1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dxp,
                                                             const void* __restrict__ yp, float* __restrict__ dw,
                                                             long long F, int H, int W, int S, int f16, float alpha) {
  constexpr int K = 25 * C, TT = (K + 15) / 16, ROWS = 4;
  extern __shared__ float sm[];
  const int Ws = W + 4, Hs = H + 4, TW = Ws + 4;
  const int Wh = Ws / S + 2, Hp = Hs + 2, wh = Ws / S;
  float* xs = sm;                              // [ROWS+4][TW][C]
  float* dzs = sm + (ROWS + 4) * TW * C;       // [ROWS][npos][32]
  int* lxs = reinterpret_cast<int*>(dzs + ROWS * (Ws + 2 * S) * 32);   // output column of each position
  const int cg = threadIdx.x & 7, tg = threadIdx.x >> 3;
  int toff[TT];
#pragma unroll
  for (int j = 0; j < TT; ++j) {
    const int k = tg + 16 * j;
    const int ci = k % C, tap = (k / C) % 25;
    toff[j] = ((tap / 5) * TW + tap % 5) * C + ci;
  }
  float acc[TT][4];
#pragma unroll
  for (int j = 0; j < TT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  // positions of one output row: for every half h the local columns xl whose padded column h*wh + xl lies in
  // [1, Ws]; a pixel that lives in two halves appears twice (the sum over halves commutes with the reduction)
  int npos = 0;
  for (int h = 0; h < S; ++h) npos += min(Wh - 1, Ws - h * wh) - max(0, 1 - h * wh) + 1;
  if (threadIdx.x < 32)
    for (int i = threadIdx.x; i < npos; i += 32) {
      int rem = i, h = 0;
      for (; h < S; ++h) {
        const int n = min(Wh - 1, Ws - h * wh) - max(0, 1 - h * wh) + 1;
        if (rem < n) break;
        rem -= n;
      }
      const int xl = max(0, 1 - h * wh) + rem;
      lxs[i] = (h << 16) | xl;
    }
  const int tiles_y = (Hs + ROWS - 1) / ROWS;
  const long long ntiles = F * tiles_y;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long f = tile / tiles_y;
    const int y0 = (int)(tile % tiles_y) * ROWS;
    const int nrow = min(ROWS, Hs - y0);
    __syncthreads();
    const int nin = (ROWS + 4) * TW * C;
    for (int i = threadIdx.x; i < nin; i += 128) {
      int ci = i % C, xx = (i / C) % TW, yy = i / (C * TW);
      int gy = y0 + yy - 4, gx = xx - 4;
      xs[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? x[((f * H + gy) * W + gx) * C + ci] : 0.f;
    }
    // (1) park dz = dxp * act'(y): item = (row, position, channel quad)
    const int nitem = nrow * npos * 8;
    for (int i = threadIdx.x; i < nitem; i += 128) {
      const int q = i & 7, pos = (i >> 3) % npos, ly = (i >> 3) / npos;
      const int hx = lxs[pos], h = hx >> 16, xl = hx & 0xffff;
      const long long o = ((((f * S + h) * Hp + y0 + ly + 1) * Wh) + xl) * 32 + 4 * q;
      float4 g = *reinterpret_cast<const float4*>(dxp + o);
      bool p0, p1, p2, p3;
      if (MODE == 0) {
        const float4 yv = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(yp) + o);
        p0 = yv.x > 0.f; p1 = yv.y > 0.f; p2 = yv.z > 0.f; p3 = yv.w > 0.f;
      } else {
        const uint2 hv = *reinterpret_cast<const uint2*>(reinterpret_cast<const u16*>(yp) + o);
        // 16-bit sign test: positive <=> sign bit clear and magnitude non-zero (bf16 and fp16 alike)
        p0 = (hv.x & 0x8000u) == 0 && (hv.x & 0x7fffu) != 0;
        p1 = (hv.x & 0x80000000u) == 0 && (hv.x & 0x7fff0000u) != 0;
        p2 = (hv.y & 0x8000u) == 0 && (hv.y & 0x7fffu) != 0;
        p3 = (hv.y & 0x80000000u) == 0 && (hv.y & 0x7fff0000u) != 0;
      }
      g.x *= p0 ? 1.f : alpha; g.y *= p1 ? 1.f : alpha; g.z *= p2 ? 1.f : alpha; g.w *= p3 ? 1.f : alpha;
      *reinterpret_cast<float4*>(dzs + ((ly * npos + pos) * 32 + 4 * q)) = g;
    }
    __syncthreads();
    // (2) rank-1 updates
    for (int ly = 0; ly < nrow; ++ly) {
      const float* xrow = xs + ly * TW * C;
      const float* drow = dzs + (ly * npos) * 32 + 4 * cg;
#pragma unroll 4
      for (int pos = 0; pos < npos; ++pos) {
        const int hx = lxs[pos];
        const int lx = (hx >> 16) * wh + (hx & 0xffff) - 1;          // output column of this position
        const float4 g = *reinterpret_cast<const float4*>(drow + pos * 32);
        const float* xb = xrow + lx * C;
#pragma unroll
        for (int j = 0; j < TT; ++j) {
          const float xv = xb[toff[j]];
          acc[j][0] = fmaf(g.x, xv, acc[j][0]);
          acc[j][1] = fmaf(g.y, xv, acc[j][1]);
          acc[j][2] = fmaf(g.z, xv, acc[j][2]);
          acc[j][3] = fmaf(g.w, xv, acc[j][3]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < TT; ++j) {
    const int k = tg + 16 * j;
    if (k < K) {
#pragma unroll
      for (int q = 0; q < 4; ++q) atomicAdd(dw + (4 * cg + q) * K + k, acc[j][q]);
    }
  }
}

extern "C" int ugn_gs_conv1_wgrad(ugn_ctx* ctx, const ugn_tensor* x, const ugn_tensor* dxp, const ugn_tensor* yp,
                                  ugn_tensor* dw, float alpha, void* stream) {
  UGN_CHECK(ctx && x && dxp && yp && dw, "ugn_gs_conv1_wgrad: null argument");
  UGN_TENSOR(yp, DT_BAD, 4, 5);
  UGN_TENSOR(dxp, DT_F32, 4, 4);
  UGN_TENSOR(dw, DT_F32, 2, 4);
  int mode = gs_mode(yp, 4);
  UGN_CHECK(mode >= 0, "gs_conv1_wgrad: bad storage mode of y");
  long long F; int H, W, c, S, rc;
  if ((rc = conv1_geom(ctx, x, yp, F, H, W, c, S)) != UGN_OK) return rc;
  const int64_t* s = gs_shape(yp, 4);
  UGN_CHECK(dxp->shape[0] == s[0] && dxp->shape[1] == s[1] && dxp->shape[2] == s[2] && dxp->shape[3] == 32,
            "gs_conv1_wgrad: dxp must have the (split) padded shape of y");
  UGN_CHECK(ugn_numel(dw) == 32 * 25 * c, "gs_conv1_wgrad: dw must be f32 [32,1,1,25c]");
  cudaStream_t st = (cudaStream_t)stream;
  UGN_CUDA(cudaMemsetAsync(ugn_ptr<float>(dw), 0, sizeof(float) * 32 * 25 * c, st));
  const int Ws = W + 4;
  const long long ntiles = F * ugn_cdiv(H + 4, 4);
  size_t smem = sizeof(float) * ((size_t)8 * (Ws + 4) * c + (size_t)4 * (Ws + 2 * S) * 32) + sizeof(int) * (Ws + 2 * S);
  UGN_CHECK(smem <= 48 * 1024, "gs_conv1_wgrad: frame too wide");
  int grid = (int)std::min<long long>(ntiles, (long long)ctx->sm_count * 5);
#define GS_WG(C_, M_)                                                                                                  \
  gs_conv1_wgrad_kernel<C_, M_><<<grid, 128, smem, st>>>(ugn_ptr<float>(x), ugn_ptr<float>(dxp), ugn_ptr<void>(yp),     \
                                                        ugn_ptr<float>(dw), F, H, W, S, gs_f16(yp), alpha)
  if (c == 1) { if (mode == 0) GS_WG(1, 0); else GS_WG(1, 1); }
  else { if (mode == 0) GS_WG(2, 0); else GS_WG(2, 1); }
#undef GS_WG
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------------------
// zero-border copies: padding='same' of the 3x3 convolutions.  The destination border is written once
// (zero) by the owner of the buffer; pad copies only the interior, crop reads only the interior.
// ---------------------------------------------------------------------------------------------------
// Split layouts: a tensor-core convolution kernel wants rows of at most 64 pixel slots, so a 66-wide
// zero-bordered input is kept as S = 2 overlapping column halves, [N*S][H+2][W/S+2][C] (half h = padded
// columns [h*W/S, h*W/S + W/S + 2)), and its valid-convolution output as [N*S][H][W/S][C].  pad / crop
// convert between split and plain images: the split factor is the ratio of the leading dimensions.
__global__ void __launch_bounds__(128) gs_pad_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int H, int Ws,
                                                     int rowv, int p, int S) {
  // src [P][Ns = N*S][H][Ws][rowv] -> dst [P][N][H+2p][Ws*S+2p][rowv]; one block = one source pixel row
  const int Hd = H + 2 * p, Wd = Ws * S + 2 * p;
  const long long row = blockIdx.x;                 // (plane-major split image) * H + y
  const long long ns = row / H;
  const int yy = (int)(row - ns * H);
  const long long n = ns / S;
  const int h = (int)(ns - n * S);
  const uint4* sp = src + row * Ws * rowv;
  uint4* dp = dst + ((n * Hd + yy + p) * Wd + h * Ws + p) * rowv;
  for (int i = threadIdx.x; i < Ws * rowv; i += blockDim.x) dp[i] = sp[i];
}

extern "C" int ugn_pad_hw(ugn_ctx* ctx, const ugn_tensor* src, ugn_tensor* dst, void* stream) {
  UGN_CHECK(ctx && src && dst, "ugn_pad_hw: null argument");
  UGN_TENSOR(src, DT_BAD, 4, 5);
  UGN_TENSOR(dst, DT_BAD, 4, 5);
  int ms = gs_mode(src, 4), md = gs_mode(dst, 4);
  UGN_CHECK(ms >= 0 && ms == md && ugn_dtype(src) == ugn_dtype(dst), "pad_hw: storage modes must match");
  const int64_t* a = gs_shape(src, 4);
  const int64_t* b = gs_shape(dst, 4);
  UGN_CHECK(b[0] > 0 && a[0] % b[0] == 0, "pad_hw: leading dims must be N*S and N");
  int S = (int)(a[0] / b[0]);
  int p = (int)(b[1] - a[1]) / 2;
  UGN_CHECK(a[3] == b[3] && p >= 0 && b[1] == a[1] + 2 * p && b[2] == a[2] * S + 2 * p,
            "pad_hw: dst must be [N,H+2p,W*S+2p,C] for src [N*S,H,W,C]");
  int es = ms == 0 ? 4 : 2;
  UGN_CHECK((a[3] * es) % 16 == 0, "pad_hw: channel row must be a multiple of 16 bytes");
  if (ugn_numel(src) == 0) return UGN_OK;
  int rowv = (int)(a[3] * es / 16), P = ms == 0 ? 1 : ms;
  long long rows = (long long)P * a[0] * a[1];
  UGN_CHECK(rows < (1LL << 31), "pad_hw: too many rows");
  gs_pad_kernel<<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(ugn_ptr<uint4>(src), ugn_ptr<uint4>(dst), (int)a[1], (int)a[2],
                                                                rowv, p, S);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

__global__ void __launch_bounds__(128) gs_crop_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int H, int Ws,
                                                      int rowv, int p, int S, int accumulate) {
  // src [N][H+2p][Ws*S+2p][rowv] -> dst [Ns = N*S][H][Ws][rowv]; one block = one destination pixel row
  const int Hs = H + 2 * p, Wsrc = Ws * S + 2 * p;
  const long long row = blockIdx.x;
  const long long ns = row / H;
  const int yy = (int)(row - ns * H);
  const long long n = ns / S;
  const int h = (int)(ns - n * S);
  const float4* sp = src + ((n * Hs + yy + p) * Wsrc + h * Ws + p) * rowv;
  float4* dp = dst + row * Ws * rowv;
  for (int i = threadIdx.x; i < Ws * rowv; i += blockDim.x) {
    float4 s4 = sp[i];
    if (accumulate) {
      const float4 d = dp[i];
      s4.x += d.x; s4.y += d.y; s4.z += d.z; s4.w += d.w;
    }
    dp[i] = s4;
  }
}

extern "C" int ugn_crop_hw(ugn_ctx* ctx, const ugn_tensor* src, ugn_tensor* dst, int accumulate, void* stream) {
  UGN_CHECK(ctx && src && dst, "ugn_crop_hw: null argument");
  UGN_TENSOR(src, DT_F32, 4, 4);
  UGN_TENSOR(dst, DT_F32, 4, 4);
  const int64_t* a = src->shape;
  const int64_t* b = dst->shape;
  UGN_CHECK(a[0] > 0 && b[0] % a[0] == 0, "crop_hw: leading dims must be N and N*S");
  int S = (int)(b[0] / a[0]);
  int p = (int)(a[1] - b[1]) / 2;
  UGN_CHECK(a[3] == b[3] && p >= 0 && a[1] == b[1] + 2 * p && a[2] == b[2] * S + 2 * p && b[3] % 4 == 0,
            "crop_hw: src must be [N,H+2p,W*S+2p,C] for dst [N*S,H,W,C]");
  if (ugn_numel(dst) == 0) return UGN_OK;
  int rowv = (int)b[3] / 4;
  long long rows = b[0] * b[1];
  UGN_CHECK(rows < (1LL << 31), "crop_hw: too many rows");
  gs_crop_kernel<<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(ugn_ptr<float4>(src), ugn_ptr<float4>(dst), (int)b[1], (int)b[2],
                                                                 rowv, p, S, accumulate);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------------------
// set pooling: Lambda(reduce_max(x, axis=1)) over the T frames of a sequence (:435,:454,:465) fused with
// the layers.Add() that follows it (:455,:466).  m = max_t a[b,t]; y = m + addend.
// Backward: tf reduce_max splits the gradient evenly among the frames that attain the maximum.
// ---------------------------------------------------------------------------------------------------
// 8 consecutive elements (channels) per thread: 16-byte loads of each 16-bit plane / two float4 of an f32 tensor
__device__ __forceinline__ void gs_load8(const void* p, int mode, int f16, long long plane, long long i, float v[8]) {
  if (mode == 0) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    return;
  }
  const u16* q = reinterpret_cast<const u16*>(p);
  union { uint4 u; u16 h[8]; } hi, lo;
  hi.u = *reinterpret_cast<const uint4*>(q + i);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = ugn_f16to32(hi.h[k], f16);
  if (mode == 2) {
    lo.u = *reinterpret_cast<const uint4*>(q + plane + i);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += ugn_f16to32(lo.h[k], f16);
  }
}
__device__ __forceinline__ void gs_store8(void* p, int mode, int f16, long long plane, long long i, const float v[8]) {
  if (mode == 0) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
    return;
  }
  u16* q = reinterpret_cast<u16*>(p);
  union { uint4 u; u16 h[8]; } hi, lo;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    hi.h[k] = ugn_cvt16(v[k], f16);
    lo.h[k] = ugn_cvt16(v[k] - ugn_f16to32(hi.h[k], f16), f16);
  }
  *reinterpret_cast<uint4*>(q + i) = hi.u;
  if (mode == 2) *reinterpret_cast<uint4*>(q + plane + i) = lo.u;
}

__global__ void __launch_bounds__(256) gs_setmax_fwd_kernel(const void* __restrict__ a, int amode, long long aplane,
                                                            const void* __restrict__ addend, int dmode, long long dplane,
                                                            float* __restrict__ m, void* __restrict__ y, int ymode,
                                                            long long yplane, int B, int T, long long Q, int f16) {
  const long long Q8 = Q >> 3, total = (long long)B * Q8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / Q8, q = (i - b * Q8) << 3;
    float best[8], v[8];
    gs_load8(a, amode, f16, aplane, (b * T) * Q + q, best);
    for (int t = 1; t < T; ++t) {
      gs_load8(a, amode, f16, aplane, (b * T + t) * Q + q, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) best[k] = fmaxf(best[k], v[k]);
    }
    if (m) gs_store8(m, 0, 0, 0, b * Q + q, best);
    if (y) {
      if (addend) {
        gs_load8(addend, dmode, f16, dplane, b * Q + q, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) best[k] += v[k];
      }
      gs_store8(y, ymode, f16, yplane, b * Q + q, best);
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
  UGN_CHECK(s[3] % 8 == 0, "setmax_fwd: channel count must be a multiple of 8");
  gs_setmax_fwd_kernel<<<gs_grid(ctx, B * Q / 8, 256), 256, 0, (cudaStream_t)stream>>>(
      ugn_ptr<void>(a), am, s[0] * Q, addend ? ugn_ptr<void>(addend) : nullptr, dm, B * Q, m ? ugn_ptr<float>(m) : nullptr,
      y ? ugn_ptr<void>(y) : nullptr, ym, B * Q, B, T, Q, f16);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

__global__ void __launch_bounds__(256) gs_setmax_bwd_kernel(const float* __restrict__ dm, const void* __restrict__ a,
                                                            int amode, long long aplane, const float* __restrict__ m,
                                                            float* __restrict__ da, int B, int T, long long Q, int f16,
                                                            int accumulate) {
  const long long Q8 = Q >> 3, total = (long long)B * Q8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / Q8, q = (i - b * Q8) << 3;
    float mx[8], g[8], v[8];
    gs_load8(m, 0, 0, 0, b * Q + q, mx);
    gs_load8(dm, 0, 0, 0, b * Q + q, g);
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < T; ++t) {
      gs_load8(a, amode, f16, aplane, (b * T + t) * Q + q, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) cnt[k] += v[k] == mx[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] /= (float)max(cnt[k], 1);
    for (int t = 0; t < T; ++t) {
      const long long o = (b * T + t) * Q + q;
      gs_load8(a, amode, f16, aplane, o, v);
      float out[8];
      if (accumulate) gs_load8(da, 0, 0, 0, o, out);
#pragma unroll
      for (int k = 0; k < 8; ++k) out[k] = (accumulate ? out[k] : 0.f) + (v[k] == mx[k] ? g[k] : 0.f);
      gs_store8(da, 0, 0, 0, o, out);
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
  UGN_CHECK(s[3] % 8 == 0, "setmax_bwd: channel count must be a multiple of 8");
  gs_setmax_bwd_kernel<<<gs_grid(ctx, B * Q / 8, 256), 256, 0, (cudaStream_t)stream>>>(
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
  const bool no_norm = (merge & UGN_FUSE3_NO_NORM) != 0;      // postriplet == 2 (:814-816): the fusion stays un-normalised
  merge &= 0xff;
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
  const float inv = no_norm ? 1.f : rsqrtf(fmaxf(ss, 1e-12f));
  col[2 * i] = inv;
  col[2 * i + 1] = ss;
  if (no_norm) return;
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
  const bool no_norm = (merge & UGN_FUSE3_NO_NORM) != 0;      // g passes through unchanged (inv = 1, no projection)
  merge &= 0xff;
  const int part = i / d, j = i - part * d;
  const float inv = no_norm ? 1.f : col[2 * i];
  const bool clamped = no_norm || !(col[2 * i + 1] > 1e-12f);
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
