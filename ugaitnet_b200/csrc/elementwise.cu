// HBM-bound kernels of the step: layout packing, pool/activation backward, flatten,
// gated fusion + l2_normalize, softmax-CE, optimiser.  All are coalesced, vectorised where
// the layout allows, and sized from the SM count.
#include "common.cuh"

static inline int grid_for(ugn_ctx* ctx, long long work_items, int block) {
  long long g = (work_items + block - 1) / block;
  long long cap = (long long)ctx->sm_count * 16;
  return (int)std::max<long long>(1, std::min(g, cap));
}

// ---------------------------------------------------------------------------------------
// a0: NCHW f32 -> NHWC (f32 | bf16 P planes), channel padding zero-filled.
// One block per (b, y): coalesced row reads, smem transpose, coalesced channel-run writes.
// ---------------------------------------------------------------------------------------
template <int MODE>  // 0 f32, 1 16-bit P=1, 2 16-bit P=2
// Device-side missing-modality expansion / mirror augmentation fused into the pack (the reference builds the
// E-fold batch on the host, data/mj_dataGeneratorMMUWYHsingle.py:780-812, and mirrors with
// data/mj_augmentation.py:12-32): output row b reads base row src_row[b]; enable[b] == 0 -> the whole volume
// is the constant `noise` (1e-9, :102); mirror[b] != 0 -> every channel flipped left-right and, literally as
// mj_mirrorsequence does for EVERY modality, even channels negated.
// The rest of the generator's augmentation that is exact on integers (:718-746): shift[b] = (tx, ty) of the random
// transform -- ImageDataGenerator.apply_transform with integer displacements, order-1 interpolation and
// fill_mode='nearest' is out[y][x] = in[clamp(y + tx)][clamp(x + ty)], applied BEFORE the mirror -- and, on rows with
// clip[b] != 0, the optical-flow magnitude clip of __load_dd (:318-321: |raw| > 2300 or < 50 -> 1e-8 before the
// 1/compressFactor * 0.1 scaling), expressed on the decoded values: |v| > clip_hi or < clip_lo -> clip_val.
__global__ void pack_input_kernel(const float* __restrict__ x, void* __restrict__ out, int B, int C,
                                  int H, int W, int Cp, long long plane, int f16,
                                  const int* __restrict__ src_row, const float* __restrict__ enable,
                                  const uint8_t* __restrict__ mirror, float noise,
                                  const int8_t* __restrict__ shift, const uint8_t* __restrict__ clip,
                                  float clip_lo, float clip_hi, float clip_val) {
  extern __shared__ float sm[];  // [C][W+1]
  int by = blockIdx.x;
  int b = by / H, y = by % H;
  const int sb = src_row ? src_row[b] : b;
  const bool on = !enable || enable[b] != 0.f;
  const bool mir = mirror && mirror[b];
  const int sy = shift ? shift[2 * b] : 0, sx = shift ? shift[2 * b + 1] : 0;
  const bool clp = clip && clip[b];
  const int ys = min(max(y + sy, 0), H - 1);
  if ((W & 3) == 0 && !mir && sx == 0) {
    // no horizontal displacement: a row of one channel is read as float4 (four times the bytes in flight per thread)
    const int W4 = W >> 2;
    for (int e = threadIdx.x; e < C * W4; e += blockDim.x) {
      const int c = e / W4, x4 = e - c * W4;
      float4 v = make_float4(noise, noise, noise, noise);
      if (on) {
        v = __ldg(reinterpret_cast<const float4*>(x + (((long long)sb * C + c) * H + ys) * W) + x4);
        if (clp) {
          float a = fabsf(v.x); if (a > clip_hi || a < clip_lo) v.x = clip_val;
          a = fabsf(v.y); if (a > clip_hi || a < clip_lo) v.y = clip_val;
          a = fabsf(v.z); if (a > clip_hi || a < clip_lo) v.z = clip_val;
          a = fabsf(v.w); if (a > clip_hi || a < clip_lo) v.w = clip_val;
        }
      }
      float* dsm = sm + c * (W + 1) + 4 * x4;
      dsm[0] = v.x; dsm[1] = v.y; dsm[2] = v.z; dsm[3] = v.w;
    }
  } else
  for (int e = threadIdx.x; e < C * W; e += blockDim.x) {
    int c = e / W, xx = e % W;
    float v = noise;
    if (on) {
      const int xs = min(max((mir ? W - 1 - xx : xx) + sx, 0), W - 1);
      v = x[(((long long)sb * C + c) * H + ys) * W + xs];
      if (clp) { const float a = fabsf(v); if (a > clip_hi || a < clip_lo) v = clip_val; }
      if (mir && !(c & 1)) v = -v;
    }
    sm[c * (W + 1) + xx] = v;
  }
  __syncthreads();
  long long obase = ((long long)b * H + y) * W * Cp;
  if (MODE != 0 && (Cp & 7) == 0) {
    // 8 channels per thread: one 16-byte store per plane
    const int C8 = Cp >> 3;
    u16* o = reinterpret_cast<u16*>(out);
    for (int e = threadIdx.x; e < W * C8; e += blockDim.x) {
      const int xx = e / C8, c0 = (e % C8) * 8;
      __align__(16) u16 hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v = (c0 + i) < C ? sm[(c0 + i) * (W + 1) + xx] : 0.f;
        ugn_split16(v, f16, hi[i], lo[i]);
      }
      *reinterpret_cast<uint4*>(o + obase + (long long)e * 8) = *reinterpret_cast<const uint4*>(hi);
      if (MODE == 2) *reinterpret_cast<uint4*>(o + plane + obase + (long long)e * 8) = *reinterpret_cast<const uint4*>(lo);
    }
    return;
  }
  for (int e = threadIdx.x; e < W * Cp; e += blockDim.x) {
    int xx = e / Cp, c = e % Cp;
    float v = c < C ? sm[c * (W + 1) + xx] : 0.f;
    if (MODE == 0) {
      reinterpret_cast<float*>(out)[obase + e] = v;
    } else {
      u16 hi, lo;
      ugn_split16(v, f16, hi, lo);
      u16* o = reinterpret_cast<u16*>(out);
      o[obase + e] = hi;
      if (MODE == 2) o[plane + obase + e] = lo;
    }
  }
}

int ew_pack_input(ugn_ctx* ctx, const float* x, void* out, int mode, int f16, int B, int C, int H, int W,
                  int Cp, const int* src_row, const float* enable, const uint8_t* mirror, float noise,
                  cudaStream_t st, const int8_t* shift, const uint8_t* clip, float clip_lo, float clip_hi,
                  float clip_val) {
  size_t smem = sizeof(float) * C * (W + 1);
  long long plane = (long long)B * H * W * Cp;
#define UGN_PACK_ARGS x, out, B, C, H, W, Cp, plane, f16, src_row, enable, mirror, noise, shift, clip, clip_lo, clip_hi, clip_val
  if (mode == 0) pack_input_kernel<0><<<B * H, 256, smem, st>>>(UGN_PACK_ARGS);
  else if (mode == 1) pack_input_kernel<1><<<B * H, 256, smem, st>>>(UGN_PACK_ARGS);
  else pack_input_kernel<2><<<B * H, 256, smem, st>>>(UGN_PACK_ARGS);
#undef UGN_PACK_ARGS
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// On-disk sample values -> the f32 volume the generator would hand to Keras (__load_dd,
// data/mj_dataGeneratorMMUWYHsingle.py:313-329), computed on the DEVICE so that only the stored integers cross PCIe
// (int16 optical flow: 2 of 4 bytes per value; uint8 gray / depth / silhouette: 1 of 4):
//   x = float(raw); |x| > clip_max or |x| < clip_min (when > 0) -> 1e-8; x = x / divisor; x = x * mul; x = x - sub
// every step one IEEE f32 operation with its own rounding, as numpy evaluates the statements (no FMA contraction).
template <typename T>
__global__ void decode_samples_kernel(const T* __restrict__ raw, float* __restrict__ out, long long n, float divisor,
                                      float mul, float sub, float clip_min, float clip_max) {
  for (long long e = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; e < n;
       e += (long long)gridDim.x * blockDim.x * 4) {
    float v[4];
    const int m = (int)min(4ll, n - e);
    for (int i = 0; i < m; ++i) {
      float x = (float)raw[e + i];
      if (clip_max > 0.f && fabsf(x) > clip_max) x = 1e-8f;
      if (clip_min > 0.f && fabsf(x) < clip_min) x = 1e-8f;
      x = __fdiv_rn(x, divisor);
      x = __fmul_rn(x, mul);
      v[i] = __fsub_rn(x, sub);
    }
    if (m == 4 && (reinterpret_cast<uintptr_t>(out + e) & 15) == 0) {
      *reinterpret_cast<float4*>(out + e) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      for (int i = 0; i < m; ++i) out[e + i] = v[i];
    }
  }
}

int ew_decode_samples(ugn_ctx* ctx, const void* raw, int is_i16, float* out, long long n, float divisor, float mul,
                      float sub, float clip_min, float clip_max, cudaStream_t st) {
  if (n == 0) return UGN_OK;
  int grid = (int)std::min<long long>((n / 4 + 255) / 256 + 1, (long long)ctx->sm_count * 16);
  if (is_i16)
    decode_samples_kernel<short><<<grid, 256, 0, st>>>(reinterpret_cast<const short*>(raw), out, n, divisor, mul, sub, clip_min, clip_max);
  else
    decode_samples_kernel<unsigned char><<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(raw), out, n, divisor, mul, sub, clip_min, clip_max);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// master [R][Cin] -> packed [P][R][Cp]   (R = Cout*kh*kw, or out-features for dense)
template <int MODE>
__global__ void pack_weight_kernel(const float* __restrict__ w, void* __restrict__ out, long long R,
                                   int Cin, int Cp, int f16) {
  long long total = R * Cp;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    long long r = e / Cp;
    int c = (int)(e % Cp);
    float v = c < Cin ? w[r * Cin + c] : 0.f;
    if (MODE == 0) reinterpret_cast<float*>(out)[e] = v;
    else {
      u16 hi, lo;
      ugn_split16(v, f16, hi, lo);
      u16* o = reinterpret_cast<u16*>(out);
      o[e] = hi;
      if (MODE == 2) o[total + e] = lo;
    }
  }
}

// unpadded fast path: flat f32 -> bf16 planes, 4 elements per thread (dense weights are 92 % of the bytes)
template <int P>
__global__ void __launch_bounds__(256) split4_kernel(const float4* __restrict__ s, uint2* __restrict__ d,
                                                     long long n4, int f16) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    const float4 v = s[e];
    __align__(8) u16 hi[4], lo[4];
    ugn_split16(v.x, f16, hi[0], lo[0]);
    ugn_split16(v.y, f16, hi[1], lo[1]);
    ugn_split16(v.z, f16, hi[2], lo[2]);
    ugn_split16(v.w, f16, hi[3], lo[3]);
    d[e] = *reinterpret_cast<const uint2*>(hi);
    if (P == 2) d[n4 + e] = *reinterpret_cast<const uint2*>(lo);
  }
}

int ew_pack_weight(ugn_ctx* ctx, const float* w, void* out, int mode, int f16, long long R, int Cin, int Cp,
                   cudaStream_t st) {
  const long long n = R * Cp;
  if (mode > 0 && Cin == Cp && (n & 3) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)out & 7) == 0) {
    int g4 = grid_for(ctx, n / 4, 256);
    if (mode == 1) split4_kernel<1><<<g4, 256, 0, st>>>((const float4*)w, (uint2*)out, n / 4, f16);
    else split4_kernel<2><<<g4, 256, 0, st>>>((const float4*)w, (uint2*)out, n / 4, f16);
    UGN_LAUNCHED(ctx);
    return UGN_OK;
  }
  int g = grid_for(ctx, R * Cp, 256);
  if (mode == 0) pack_weight_kernel<0><<<g, 256, 0, st>>>(w, out, R, Cin, Cp, f16);
  else if (mode == 1) pack_weight_kernel<1><<<g, 256, 0, st>>>(w, out, R, Cin, Cp, f16);
  else pack_weight_kernel<2><<<g, 256, 0, st>>>(w, out, R, Cin, Cp, f16);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// generic f32 -> bf16 planes
template <int P>
__global__ void split_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, long long n, int f16) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    u16 hi, lo;
    ugn_split16(s[e], f16, hi, lo);
    d[e] = hi;
    if (P == 2) d[n + e] = lo;
  }
}
int ew_split(ugn_ctx* ctx, const float* s, __nv_bfloat16* d, int P, int f16, long long n, cudaStream_t st) {
  if ((n & 3) == 0 && ((uintptr_t)s & 15) == 0 && ((uintptr_t)d & 7) == 0) {
    int g4 = grid_for(ctx, n / 4, 256);
    if (P == 1) split4_kernel<1><<<g4, 256, 0, st>>>((const float4*)s, (uint2*)d, n / 4, f16);
    else split4_kernel<2><<<g4, 256, 0, st>>>((const float4*)s, (uint2*)d, n / 4, f16);
    UGN_LAUNCHED(ctx);
    return UGN_OK;
  }
  int g = grid_for(ctx, n, 256);
  if (P == 1) split_kernel<1><<<g, 256, 0, st>>>(s, d, n, f16);
  else split_kernel<2><<<g, 256, 0, st>>>(s, d, n, f16);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// backward through MaxPooling2D(2x2, floor) + activation:
// dz[b,yo,xo,c] = (pos(yo,xo) == idx[b,yo/2,xo/2,c]) ? dy * act'(y) : 0 ; rows/cols beyond the
// floor region get 0 (nets/mj_uwyhNets_ba.py:85,92 -- 23->11 and 9->4 drop the last row/col).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float yval(float v, int) { return v; }
__device__ __forceinline__ float yval(u16 v, int f16) { return ugn_f16to32(v, f16); }

template <int MODE, typename YT>
__global__ void bwd_act_kernel(const float* __restrict__ dy, const YT* __restrict__ y,
                               const uint8_t* __restrict__ idx, void* __restrict__ dz, int B, int Ho,
                               int Wo, int Hp, int Wp, int C, int act, float alpha, int pool, int f16,
                               const float* __restrict__ gs) {
  long long total = (long long)B * Ho * Wo * C;
  const float scale = (MODE > 0 && gs) ? gs[0] : 1.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    long long pix = e / C;
    int xo = (int)(pix % Wo);
    int yo = (int)((pix / Wo) % Ho);
    int b = (int)(pix / ((long long)Wo * Ho));
    float v = 0.f;
    if (pool) {
      int yp = yo >> 1, xp = xo >> 1;
      if (yp < Hp && xp < Wp) {
        long long o = (((long long)b * Hp + yp) * Wp + xp) * C + c;
        int pos = ((yo & 1) << 1) | (xo & 1);
        if (idx[o] == pos) v = dy[o] * ugn_act_bwd(yval(y[o], f16), act, alpha);
      }
    } else {
      v = dy[e] * ugn_act_bwd(yval(y[e], f16), act, alpha);
    }
    if (MODE == 0) reinterpret_cast<float*>(dz)[e] = v;
    else {
      u16 hi, lo;
      ugn_split16(v * scale, f16, hi, lo);
      u16* o = reinterpret_cast<u16*>(dz);
      o[e] = hi;
      if (MODE == 2) o[total + e] = lo;
    }
  }
}

// tensor-core storage variant: 8 channels per thread, 16-byte loads / stores.  A thread keeps ONE channel
// group (c8 = tid % C8) for its whole pixel loop, so the bias gradient db[c] = sum_pix dz[pix][c] (Keras
// Conv2D bias, nets/mj_uwyhNets_ba.py:82) is accumulated in registers from the f32 values and reduced
// through smem + one atomicAdd per channel per block -- no second pass over dz.
template <int P>
__global__ void __launch_bounds__(256) bwd_act_vec8_kernel(const float* __restrict__ dy,
                                                           const __nv_bfloat16* __restrict__ y,
                                                           const uint8_t* __restrict__ idx,
                                                           __nv_bfloat16* __restrict__ dz, float* __restrict__ db,
                                                           int B, int Ho, int Wo,
                                                           int Hp, int Wp, int C8, int act, float alpha, int pool,
                                                           int f16, const float* __restrict__ gs) {
  __shared__ float red[256 * 8];
  const float scale = gs ? gs[0] : 1.f;
  const unsigned npix = (unsigned)B * Ho * Wo;
  const long long plane = (long long)npix * C8 * 8;
  const unsigned lanes = 256u / C8, c8 = threadIdx.x % C8, lane = threadIdx.x / C8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (lane < lanes) {
    for (unsigned pix = blockIdx.x * lanes + lane; pix < npix; pix += gridDim.x * lanes) {
      const unsigned e = pix * C8 + c8;
      const unsigned xo = pix % Wo, t2 = pix / Wo, yo = t2 % Ho, b = t2 / Ho;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
      bool live = true;
      long long o = (long long)e * 8;
      int pos = 0;
      if (pool) {
        const unsigned yp = yo >> 1, xp = xo >> 1;
        live = yp < (unsigned)Hp && xp < (unsigned)Wp;
        o = ((((long long)b * Hp + yp) * Wp + xp) * C8 + c8) * 8;
        pos = ((yo & 1) << 1) | (xo & 1);
      }
      if (live) {
        const float4 d0 = *reinterpret_cast<const float4*>(dy + o), d1 = *reinterpret_cast<const float4*>(dy + o + 4);
        const uint4 yr = *reinterpret_cast<const uint4*>(y + o);
        const __nv_bfloat16* yh = reinterpret_cast<const __nv_bfloat16*>(&yr);
        const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        unsigned long long ib = 0;
        if (pool) ib = *reinterpret_cast<const unsigned long long*>(idx + o);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool sel = !pool || (int)((ib >> (8 * i)) & 0xff) == pos;
          if (sel) v[i] = dd[i] * ugn_act_bwd(ugn_f16to32(yh[i], f16), act, alpha);
          acc[i] += v[i];
        }
      }
      __align__(16) u16 hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ugn_split16(v[i] * scale, f16, hi[i], lo[i]);
      *reinterpret_cast<uint4*>(dz + (long long)e * 8) = *reinterpret_cast<const uint4*>(hi);
      if (P == 2) *reinterpret_cast<uint4*>(dz + plane + (long long)e * 8) = *reinterpret_cast<const uint4*>(lo);
    }
  }
  if (db) {
    const int C = C8 * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = acc[i];   // == red[lane * C + c8 * 8 + i] for live lanes
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float sacc = 0.f;
      for (unsigned l = 0; l < lanes; ++l) sacc += red[l * C + c];
      atomicAdd(db + c, sacc);
    }
  }
}

// POOLED layers, input-driven: a thread owns one pooled element group (8 channels) -- dy / y / idx are loaded ONCE
// (the output-driven kernel above reads them from each of the four window positions) -- and writes the four
// pre-pool pixels of its 2x2 window (the selected position gets dy * act'(y), the others zero).  Odd Ho / Wo leave a
// last row / column outside every window: the work items of the fringe (yp == Hp or xp == Wp) zero-fill it.
template <int P>
__global__ void __launch_bounds__(256) bwd_act_pool_vec8_kernel(const float* __restrict__ dy,
                                                                const __nv_bfloat16* __restrict__ y,
                                                                const uint8_t* __restrict__ idx,
                                                                __nv_bfloat16* __restrict__ dz, float* __restrict__ db,
                                                                int B, int Ho, int Wo, int Hp, int Wp, int C8, int act,
                                                                float alpha, int f16, const float* __restrict__ gs) {
  __shared__ float red[256 * 8];
  const float scale = gs ? gs[0] : 1.f;
  const int He = (Ho + 1) >> 1, We = (Wo + 1) >> 1;            // window grid incl. the fringe
  const unsigned nwin = (unsigned)B * He * We;
  const long long plane = (long long)B * Ho * Wo * C8 * 8;
  const unsigned lanes = 256u / C8, c8 = threadIdx.x % C8, lane = threadIdx.x / C8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (lane < lanes) {
    for (unsigned wdx = blockIdx.x * lanes + lane; wdx < nwin; wdx += gridDim.x * lanes) {
      const unsigned xp = wdx % We, t2 = wdx / We, yp = t2 % He, b = t2 / He;
      const bool live = yp < (unsigned)Hp && xp < (unsigned)Wp;
      float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      unsigned long long ib = 0;
      if (live) {
        const long long o = ((((long long)b * Hp + yp) * Wp + xp) * C8 + c8) * 8;
        const float4 d0 = *reinterpret_cast<const float4*>(dy + o), d1 = *reinterpret_cast<const float4*>(dy + o + 4);
        const uint4 yr = *reinterpret_cast<const uint4*>(y + o);
        const __nv_bfloat16* yh = reinterpret_cast<const __nv_bfloat16*>(&yr);
        ib = *reinterpret_cast<const unsigned long long*>(idx + o);
        const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          g[i] = dd[i] * ugn_act_bwd(ugn_f16to32(yh[i], f16), act, alpha);
          acc[i] += g[i];
        }
      }
#pragma unroll
      for (int pos = 0; pos < 4; ++pos) {
        const unsigned yo = 2 * yp + (pos >> 1), xo = 2 * xp + (pos & 1);
        if (yo >= (unsigned)Ho || xo >= (unsigned)Wo) continue;
        __align__(16) u16 hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v = (live && (int)((ib >> (8 * i)) & 0xff) == pos) ? g[i] * scale : 0.f;
          ugn_split16(v, f16, hi[i], lo[i]);
        }
        const long long e = ((((long long)b * Ho + yo) * Wo + xo) * C8 + c8) * 8;
        *reinterpret_cast<uint4*>(dz + e) = *reinterpret_cast<const uint4*>(hi);
        if (P == 2) *reinterpret_cast<uint4*>(dz + plane + e) = *reinterpret_cast<const uint4*>(lo);
      }
    }
  }
  if (db) {
    const int C = C8 * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = acc[i];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float sacc = 0.f;
      for (unsigned l = 0; l < lanes; ++l) sacc += red[l * C + c];
      atomicAdd(db + c, sacc);
    }
  }
}

int ew_bwd_act(ugn_ctx* ctx, const float* dy, const void* y, int y_bf16, const uint8_t* idx,
               void* dz, float* db, int* db_done, int mode, int f16, int B, int Ho, int Wo, int Hp, int Wp, int C,
               int act, float alpha, int pool, cudaStream_t st) {
  const float* gs = ctx->gscale;
  *db_done = 0;
  if (y_bf16 && mode > 0 && C % 8 == 0 && C / 8 <= 256 && (long long)B * Ho * Wo * (C / 8) < 0x7fffffffLL) {
    int g8 = grid_for(ctx, (long long)B * Ho * Wo * (C / 8), 256);
    if (db) {
      UGN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * C, st));
      *db_done = 1;
    }
    static const bool pool_in = !(getenv("UGN_BWDACT_OUT") && atoi(getenv("UGN_BWDACT_OUT")));
    if (pool && pool_in) {
      const long long nwin = (long long)B * ((Ho + 1) / 2) * ((Wo + 1) / 2) * (C / 8);
      const int gp = grid_for(ctx, nwin, 256);
      if (mode == 1)
        bwd_act_pool_vec8_kernel<1><<<gp, 256, 0, st>>>(dy, (const __nv_bfloat16*)y, idx, (__nv_bfloat16*)dz, db, B, Ho, Wo, Hp, Wp, C / 8, act, alpha, f16, gs);
      else
        bwd_act_pool_vec8_kernel<2><<<gp, 256, 0, st>>>(dy, (const __nv_bfloat16*)y, idx, (__nv_bfloat16*)dz, db, B, Ho, Wo, Hp, Wp, C / 8, act, alpha, f16, gs);
      UGN_LAUNCHED(ctx);
      return UGN_OK;
    }
    if (mode == 1)
      bwd_act_vec8_kernel<1><<<g8, 256, 0, st>>>(dy, (const __nv_bfloat16*)y, idx, (__nv_bfloat16*)dz, db, B, Ho, Wo, Hp, Wp, C / 8, act, alpha, pool, f16, gs);
    else
      bwd_act_vec8_kernel<2><<<g8, 256, 0, st>>>(dy, (const __nv_bfloat16*)y, idx, (__nv_bfloat16*)dz, db, B, Ho, Wo, Hp, Wp, C / 8, act, alpha, pool, f16, gs);
    UGN_LAUNCHED(ctx);
    return UGN_OK;
  }
  int g = grid_for(ctx, (long long)B * Ho * Wo * C, 256);
#define LAUNCH(M, YT) \
  bwd_act_kernel<M, YT><<<g, 256, 0, st>>>(dy, (const YT*)y, idx, dz, B, Ho, Wo, Hp, Wp, C, act, alpha, pool, f16, gs)
  if (y_bf16) {
    if (mode == 0) LAUNCH(0, __nv_bfloat16);
    else if (mode == 1) LAUNCH(1, __nv_bfloat16);
    else LAUNCH(2, __nv_bfloat16);
  } else {
    if (mode == 0) LAUNCH(0, float);
    else if (mode == 1) LAUNCH(1, float);
    else LAUNCH(2, float);
  }
#undef LAUNCH
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// Flatten over (C,H,W) (channels_first): NHWC [B,H,W,C] -> [B, C*H*W] and back.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void flatten_kernel(const T* __restrict__ src, T* __restrict__ dst, long long planes_stride,
                               int P, int B, int HW, int C, int to_chw) {
  long long total = (long long)B * HW * C;
  for (int pl = 0; pl < P; ++pl) {
    const T* s = src + pl * planes_stride;
    T* d = dst + pl * planes_stride;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
      int b = (int)(e / ((long long)HW * C));
      int r = (int)(e % ((long long)HW * C));
      if (to_chw) {  // e indexes dst [b][c][hw]
        int c = r / HW, hw = r % HW;
        d[e] = s[((long long)b * HW + hw) * C + c];
      } else {       // e indexes dst [b][hw][c]
        int hw = r / C, c = r % C;
        d[e] = s[((long long)b * C + c) * HW + hw];
      }
    }
  }
}

int ew_flatten(ugn_ctx* ctx, const void* src, void* dst, int bf16, int P, int B, int HW, int C,
               int to_chw, cudaStream_t st) {
  long long n = (long long)B * HW * C;
  int g = grid_for(ctx, n, 256);
  if (bf16)
    flatten_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n, P, B, HW, C, to_chw);
  else
    flatten_kernel<float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, n, P, B, HW, C, to_chw);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// dense backward helper: dz = dy * mask * act'(y)   (+ optional bf16 copy)
// ---------------------------------------------------------------------------------------
template <int P>
__global__ void act_mask_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                    const float* __restrict__ mask, float* __restrict__ dz,
                                    __nv_bfloat16* __restrict__ dz16, long long n, int act, float alpha,
                                    int f16, const float* __restrict__ gs) {
  const float scale = (P > 0 && gs) ? gs[0] : 1.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    float v = dy[e];
    if (mask) v *= mask[e];
    if (y) v *= ugn_act_bwd(y[e], act, alpha);
    if (dz) dz[e] = v;
    if (P > 0) {
      u16 hi, lo;
      ugn_split16(v * scale, f16, hi, lo);
      dz16[e] = hi;
      if (P == 2) dz16[n + e] = lo;
    }
  }
}
int ew_act_mask_bwd(ugn_ctx* ctx, const float* dy, const float* y, const float* mask, float* dz,
                    __nv_bfloat16* dz16, int P, int f16, long long n, int act, float alpha, cudaStream_t st) {
  int g = grid_for(ctx, n, 256);
  const float* gs = ctx->gscale;
  if (P == 0) act_mask_bwd_kernel<0><<<g, 256, 0, st>>>(dy, y, mask, dz, dz16, n, act, alpha, f16, gs);
  else if (P == 1) act_mask_bwd_kernel<1><<<g, 256, 0, st>>>(dy, y, mask, dz, dz16, n, act, alpha, f16, gs);
  else act_mask_bwd_kernel<2><<<g, 256, 0, st>>>(dy, y, mask, dz, dz16, n, act, alpha, f16, gs);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// a2+a3+a4: gate x flag -> merge (max | avg | sign_max) -> l2_normalize.  One CTA per row;
// float4 loads of every modality's signature, warp-shuffle + smem reduction of sum(x^2).
// ---------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(256) fuse_fwd_kernel(FusePtrs ptrs, int nmods, int d,
                                                       float* __restrict__ sig,
                                                       __nv_bfloat16* __restrict__ sig16,
                                                       uint8_t* __restrict__ winner,
                                                       float* __restrict__ inv_norm, int merge,
                                                       int normalize, long long plane, int f16) {
  extern __shared__ float row[];  // [d] fused values
  __shared__ float red[8];
  const int b = blockIdx.x;
  float fl[4];
  for (int m = 0; m < nmods; ++m) fl[m] = ptrs.flag[m][b];
  float ss = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float best = ptrs.br[0][(long long)b * d + j] * fl[0];
    int win = 0;
    if (merge == UGN_MERGE_AVG) {
      for (int m = 1; m < nmods; ++m) best += ptrs.br[m][(long long)b * d + j] * fl[m];
      best = best / (float)nmods;
    } else {
      float key = merge == UGN_MERGE_SIGNMAX ? fabsf(best) : best;
      for (int m = 1; m < nmods; ++m) {
        float v = ptrs.br[m][(long long)b * d + j] * fl[m];
        float kv = merge == UGN_MERGE_SIGNMAX ? fabsf(v) : v;
        if (kv > key) { key = kv; best = v; win = m; }   // strict: ties -> earlier modality
      }
    }
    row[j] = best;
    if (winner) winner[(long long)b * d + j] = (uint8_t)win;
    ss += best * best;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  ss = red[0];
  float inv = normalize ? rsqrtf(fmaxf(ss, 1e-12f)) : 1.f;
  if (threadIdx.x == 0 && inv_norm) {
    inv_norm[2 * b] = inv;
    inv_norm[2 * b + 1] = ss;
  }
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float v = row[j] * inv;
    sig[(long long)b * d + j] = v;
    if (P > 0) {
      u16 hi, lo;
      ugn_split16(v, f16, hi, lo);
      sig16[(long long)b * d + j] = hi;
      if (P == 2) sig16[plane + (long long)b * d + j] = lo;
    }
  }
}

int ew_fuse_fc1_fwd(ugn_ctx* ctx, const FusePtrs& ptrs, int nmods, int B, int d, float* sig, __nv_bfloat16* sig16, int P,
                    int f16, uint8_t* winner, float* inv_norm, int merge, int normalize, const float* Wc, const float* bc,
                    int nc, float* code, const float* cmask, float* dropcode, int act, float alpha, cudaStream_t st);

int ew_fuse_fwd(ugn_ctx* ctx, const FusePtrs& ptrs, int nmods, int B, int d, float* sig,
                __nv_bfloat16* sig16, int P, int f16, uint8_t* winner, float* inv_norm, int merge,
                int normalize, cudaStream_t st) {
  // vectorised kernel (float4 loads, missing modalities skipped) whenever the row width allows: the FC1-fused kernel
  // with no FC1 behind it
  if (d % 4 == 0 && (size_t)d * 4 <= 48 * 1024)
    return ew_fuse_fc1_fwd(ctx, ptrs, nmods, B, d, sig, sig16, P, f16, winner, inv_norm, merge, normalize, nullptr, nullptr,
                           0, nullptr, nullptr, nullptr, 0, 0.f, st);
  size_t smem = sizeof(float) * d;
  UGN_CHECK(smem <= 48 * 1024, "fuse_fwd: signature width %d exceeds the 12288 floats one shared-memory row holds", d);
  long long plane = (long long)B * d;
  if (P == 0) fuse_fwd_kernel<0><<<B, 256, smem, st>>>(ptrs, nmods, d, sig, sig16, winner, inv_norm, merge, normalize, plane, f16);
  else if (P == 1) fuse_fwd_kernel<1><<<B, 256, smem, st>>>(ptrs, nmods, d, sig, sig16, winner, inv_norm, merge, normalize, plane, f16);
  else fuse_fwd_kernel<2><<<B, 256, smem, st>>>(ptrs, nmods, d, sig, sig16, winner, inv_norm, merge, normalize, plane, f16);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// a2+a3+a4+a5 in ONE kernel (north_star): gate x flag -> merge -> l2_normalize -> FC1 "code" = act(sig . Wc^T + b)
// [-> dropcode = code * dropout mask] (nets/mj_uwyhNets_ba.py:1163-1203).  One CTA per row:
//   phase 1  float4 loads of every AVAILABLE modality's branch output (the use-flag is uniform over the row: a missing
//            modality -- flag 0, gated value 0 -- is not read at all), merge, winner as uchar4, sum(x^2) by warp
//            shuffles + smem
//   phase 2  normalise; the row stays in shared memory, sig (f32, and its 16-bit planes for the tensor-core Gram) is
//            written once
//   phase 3  FC1 from shared memory: a warp per output feature pair, float4 loads of the weight rows (L2-resident:
//            every CTA reads the same nc x d matrix), warp-shuffle reduction, bias + activation + dropout mask
// ---------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(256) fuse_fc1_fwd_kernel(FusePtrs ptrs, int nmods, int d, float* __restrict__ sig,
                                                           __nv_bfloat16* __restrict__ sig16,
                                                           uint8_t* __restrict__ winner, float* __restrict__ inv_norm,
                                                           int merge, int normalize, long long plane, int f16,
                                                           const float* __restrict__ Wc, const float* __restrict__ bc,
                                                           int nc, float* __restrict__ code,
                                                           const float* __restrict__ cmask, float* __restrict__ dropcode,
                                                           int act, float alpha) {
  extern __shared__ __align__(16) float row[];  // [d] fused values
  __shared__ float red[8];
  const int b = blockIdx.x, d4 = d >> 2;
  // (fully unrolled over the <= 4 modalities with constant indices: the pointer table stays in the parameter bank and
  // the flags in registers -- no local-memory copies)
  float fl[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int m = 0; m < 4; ++m)
    if (m < nmods) fl[m] = ptrs.flag[m][b];
  float ss = 0.f;
  for (int q = threadIdx.x; q < d4; q += blockDim.x) {
    float best[4] = {0.f, 0.f, 0.f, 0.f}, key[4];
    int win[4] = {0, 0, 0, 0};
    if (fl[0] != 0.f) {
      const float4 v = *reinterpret_cast<const float4*>(ptrs.br[0] + (long long)b * d + 4 * q);
      best[0] = v.x * fl[0]; best[1] = v.y * fl[0]; best[2] = v.z * fl[0]; best[3] = v.w * fl[0];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) key[i] = merge == UGN_MERGE_SIGNMAX ? fabsf(best[i]) : best[i];
#pragma unroll
    for (int m = 1; m < 4; ++m) {
      if (m >= nmods) break;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (fl[m] != 0.f) v = *reinterpret_cast<const float4*>(ptrs.br[m] + (long long)b * d + 4 * q);
      const float vv[4] = {v.x * fl[m], v.y * fl[m], v.z * fl[m], v.w * fl[m]};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (merge == UGN_MERGE_AVG) {
          best[i] += vv[i];
        } else {
          const float kv = merge == UGN_MERGE_SIGNMAX ? fabsf(vv[i]) : vv[i];
          if (kv > key[i]) { key[i] = kv; best[i] = vv[i]; win[i] = m; }   // strict: ties -> earlier modality
        }
      }
    }
    if (merge == UGN_MERGE_AVG) {
#pragma unroll
      for (int i = 0; i < 4; ++i) best[i] = best[i] / (float)nmods;
    }
    *reinterpret_cast<float4*>(row + 4 * q) = make_float4(best[0], best[1], best[2], best[3]);
    if (winner)
      *reinterpret_cast<uchar4*>(winner + (long long)b * d + 4 * q) = make_uchar4((unsigned char)win[0], (unsigned char)win[1],
                                                                                 (unsigned char)win[2], (unsigned char)win[3]);
    ss += best[0] * best[0] + best[1] * best[1] + best[2] * best[2] + best[3] * best[3];
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  ss = red[0];
  const float inv = normalize ? rsqrtf(fmaxf(ss, 1e-12f)) : 1.f;
  if (threadIdx.x == 0 && inv_norm) {
    inv_norm[2 * b] = inv;
    inv_norm[2 * b + 1] = ss;
  }
  for (int q = threadIdx.x; q < d4; q += blockDim.x) {
    float4 v = *reinterpret_cast<float4*>(row + 4 * q);
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    *reinterpret_cast<float4*>(row + 4 * q) = v;
    *reinterpret_cast<float4*>(sig + (long long)b * d + 4 * q) = v;
    if (P > 0) {
      const float vv[4] = {v.x, v.y, v.z, v.w};
      __align__(8) u16 hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ugn_split16(vv[i], f16, hi[i], lo[i]);
      *reinterpret_cast<uint2*>(sig16 + (long long)b * d + 4 * q) = *reinterpret_cast<const uint2*>(hi);
      if (P == 2) *reinterpret_cast<uint2*>(sig16 + plane + (long long)b * d + 4 * q) = *reinterpret_cast<const uint2*>(lo);
    }
  }
  __syncthreads();
  // FC1: two output features per warp iteration (two independent weight streams in flight)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = 2 * warp; j < nc; j += 2 * nw) {
    const bool two = j + 1 < nc;
    const float4* w0 = reinterpret_cast<const float4*>(Wc + (long long)j * d);
    const float4* w1 = reinterpret_cast<const float4*>(Wc + (long long)(two ? j + 1 : j) * d);
    float a0 = 0.f, a1 = 0.f;
    for (int q = lane; q < d4; q += 32) {
      const float4 s4 = *reinterpret_cast<const float4*>(row + 4 * q);
      const float4 x0 = __ldg(w0 + q), x1 = __ldg(w1 + q);
      a0 = fmaf(x0.x, s4.x, fmaf(x0.y, s4.y, fmaf(x0.z, s4.z, fmaf(x0.w, s4.w, a0))));
      a1 = fmaf(x1.x, s4.x, fmaf(x1.y, s4.y, fmaf(x1.z, s4.z, fmaf(x1.w, s4.w, a1))));
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane < (two ? 2 : 1)) {
      const int jj = j + lane;
      const float y = ugn_act_fwd((lane ? a1 : a0) + (bc ? bc[jj] : 0.f), act, alpha);
      const long long o = (long long)b * nc + jj;
      code[o] = y;
      if (dropcode) dropcode[o] = cmask ? y * cmask[o] : y;
    }
  }
}

int ew_fuse_fc1_fwd(ugn_ctx* ctx, const FusePtrs& ptrs, int nmods, int B, int d, float* sig, __nv_bfloat16* sig16, int P,
                    int f16, uint8_t* winner, float* inv_norm, int merge, int normalize, const float* Wc, const float* bc,
                    int nc, float* code, const float* cmask, float* dropcode, int act, float alpha, cudaStream_t st) {
  size_t smem = sizeof(float) * d;
  long long plane = (long long)B * d;
#define UGN_FF_ARGS ptrs, nmods, d, sig, sig16, winner, inv_norm, merge, normalize, plane, f16, Wc, bc, nc, code, cmask, dropcode, act, alpha
  if (P == 0) fuse_fc1_fwd_kernel<0><<<B, 256, smem, st>>>(UGN_FF_ARGS);
  else if (P == 1) fuse_fc1_fwd_kernel<1><<<B, 256, smem, st>>>(UGN_FF_ARGS);
  else fuse_fc1_fwd_kernel<2><<<B, 256, smem, st>>>(UGN_FF_ARGS);
#undef UGN_FF_ARGS
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// backward: y = x*inv, inv = rsqrt(max(ss,eps)).  ss > eps: dx = inv*(dy - y*sum(dy*y));
// clamped (ss <= eps): inv is a constant, dx = dy*inv.  Then route dx to the winning modality
// (max / sign_max) or spread it (avg), times the gate flag.
__global__ void __launch_bounds__(256) fuse_bwd_kernel(FusePtrs ptrs, int nmods, int d,
                                                       const float* __restrict__ dsig,
                                                       const float* __restrict__ sig,
                                                       const uint8_t* __restrict__ winner,
                                                       const float* __restrict__ inv_norm, int merge,
                                                       int normalize) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  float inv = 1.f, dot = 0.f;
  bool clamped = true;
  if (normalize) {
    inv = inv_norm[2 * b];
    clamped = !(inv_norm[2 * b + 1] > 1e-12f);
    if (!clamped) {
      for (int j = threadIdx.x; j < d; j += blockDim.x)
        dot += dsig[(long long)b * d + j] * sig[(long long)b * d + j];
      dot = warp_sum(dot);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
      __syncthreads();
      if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) red[0] = v;
      }
      __syncthreads();
      dot = red[0];
    }
  }
  float fl[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int m = 0; m < 4; ++m)          // constant indices: pointer table and flags never go through local memory
    if (m < nmods) fl[m] = ptrs.flag[m][b];
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    long long o = (long long)b * d + j;
    float g = dsig[o];
    if (normalize) g = clamped ? g * inv : inv * (g - sig[o] * dot);
    const int w = merge == UGN_MERGE_AVG ? -1 : (int)winner[o];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (m >= nmods) break;
      ptrs.dbr[m][o] = merge == UGN_MERGE_AVG ? g * fl[m] / (float)nmods : ((m == w) ? g * fl[m] : 0.f);
    }
  }
}

int ew_fuse_bwd(ugn_ctx* ctx, const FusePtrs& ptrs, int nmods, int B, int d, const float* dsig,
                const float* sig, const uint8_t* winner, const float* inv_norm, int merge,
                int normalize, cudaStream_t st) {
  fuse_bwd_kernel<<<B, 256, 0, st>>>(ptrs, nmods, d, dsig, sig, winner, inv_norm, merge, normalize);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// softmax + categorical CE (fused log-softmax), one warp per row, C <= 32*32.
// ---------------------------------------------------------------------------------------
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const int* __restrict__ labels,
                                  float* __restrict__ loss_acc, float* __restrict__ dlogits, int B, int C,
                                  float scale, float smooth) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* row = logits + (long long)warp * C;
  float mx = -INFINITY;
  int amax = 0;
  for (int j = lane; j < C; j += 32) {
    float v = row[j];
    if (v > mx) { mx = v; amax = j; }
  }
  // warp argmax with lowest-index tie rule
  for (int o = 16; o > 0; o >>= 1) {
    float om = __shfl_xor_sync(0xffffffffu, mx, o);
    int oa = __shfl_xor_sync(0xffffffffu, amax, o);
    if (om > mx || (om == mx && oa < amax)) { mx = om; amax = oa; }
  }
  float se = 0.f, sz = 0.f;
  for (int j = lane; j < C; j += 32) { se += expf(row[j] - mx); sz += row[j]; }
  se = warp_sum(se);
  sz = warp_sum(sz);
  float lse = mx + logf(se);
  int lab = labels[warp];
  // label smoothing (tf.losses.CategoricalCrossentropy(label_smoothing=e), nets/mj_uwyhNets_ba.py:1252-1262):
  // y_s = y*(1-e) + e/C;  loss = -sum_j y_s[j] log p[j] = lse - (1-e) z[lab] - (e/C) sum_j z[j]
  const float uni = smooth / (float)C;
  // an id outside [0, C) has an all-zero one-hot row (tf.one_hot semantics): its target mass is `smooth` only, and
  // nothing is read outside the row
  const bool valid = lab >= 0 && lab < C;
  const float hot = valid ? 1.f - smooth : 0.f, mass = hot + smooth;
  if (lane == 0) {
    atomicAdd(loss_acc, (mass * lse - (valid ? hot * row[lab] : 0.f) - uni * sz) / (float)B);
    atomicAdd(loss_acc + 1, (valid && amax == lab ? 1.f : 0.f) / (float)B);
  }
  if (dlogits) {
    float s = scale / (float)B;
    for (int j = lane; j < C; j += 32) {
      float p = expf(row[j] - lse);
      dlogits[(long long)warp * C + j] = s * (mass * p - (j == lab ? hot : 0.f) - uni);
    }
  }
}

int ew_softmax_ce(ugn_ctx* ctx, const float* logits, const int* labels, float* loss_acc,
                  float* dlogits, int B, int C, float scale, float smooth, cudaStream_t st) {
  UGN_CUDA(cudaMemsetAsync(loss_acc, 0, 2 * sizeof(float), st));
  int threads = 128, blocks = ugn_cdiv((long long)B * 32, threads);
  softmax_ce_kernel<<<blocks, threads, 0, st>>>(logits, labels, loss_acc, dlogits, B, C, scale, smooth);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// a8/a9: fused regulariser + optimiser over a flat arena (segments padded to 4 elements).
// 7 arena-sized streams for Adam (read w,g,m,v; write w,m,v) = 28 B/param -> HBM-bound.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int seg_find(const long long* __restrict__ off, int S, long long e) {
  int lo = 0, hi = S;  // off[lo] <= e < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// Arenas of the other ranks of a data-parallel group, mapped into this process (symmetric / peer memory).
struct PeerArenas {
  int n = 0;                 // 0: single-GPU step
  const float* g[8] = {};    // gradient arena of every rank (own rank included)
  float* w[8] = {};          // weight arena of every rank
  float* reg[8] = {};        // regulariser-value scalar of every rank (nullptr: only the local slice value is produced)
  // PUSHED gradients: for arena ranges whose gradients are final early in the backward pass (the dense layers, 92 % of
  // the bytes) every rank copies its part of the owner's slice into slot `rank` of the owner's staging buffer with the
  // copy engines WHILE the convolution backward still runs; the owner then sums local memory instead of pulling over
  // NVLink inside this kernel.  stage: [n][slice4] float4 slots of THIS rank, srange: staged [lo, hi) in float4 units
  const float4* stage = nullptr;
  long long slice4 = 0;
  int rank = 0, nsr = 0;
  long long srange[8] = {};
  // 16-bit compute copies of the dense weights live in ONE symmetric arena per rank: the owner of a slice writes the
  // hi / lo planes of its updated weights straight into every rank's copy (same 4 B / parameter on NVLink as the f32
  // all-gather it replaces; the non-owners' f32 masters of those segments are refreshed on demand only) -- no local
  // re-split pass after the exchange.  cw[p]: base of rank p's arena, cw_mc: its multicast address
  unsigned char* cw[8] = {};
  unsigned char* cw_mc = nullptr;
  // cw_defer: the kernel refreshes THIS rank's copies only; the caller then sends its slice of the planes to the peers
  // with the copy engines underneath the next step's convolution forward (the all-gather leaves the critical path)
  bool cw_defer = false;
  const float* mc_g = nullptr;   // NVSwitch multicast addresses of the two arenas (nullptr: unicast peer accesses):
  float* mc_w = nullptr;         // multimem.ld_reduce sums in the switch, multimem.st broadcasts -- half the link traffic
};

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* addr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(addr) : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st8(void* addr, uint2 v) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};"
               :: "l"(addr), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)) : "memory");
}
__device__ __forceinline__ void multimem_st(float* addr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// OPT: 0 adam, 1 sgd-momentum.  DP: the data-parallel instantiation (peer arenas); compile-time so that the
// single-GPU kernel carries none of it (the dynamically indexed peer table costs it 25 % otherwise)
template <int OPT, bool DP>
__global__ void __launch_bounds__(256) optim_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v,
                                                    const long long* __restrict__ off,
                                                    const float* __restrict__ l2, int S, long long n4,
                                                    float lr, float b1, float b2, float eps, float gscale,
                                                    float* __restrict__ reg_out,
                                                    const float* __restrict__ lr_dev,
                                                    const long long* __restrict__ pack, int packP, int f16,
                                                    float* __restrict__ vhat, float wd, PeerArenas peers,
                                                    long long q0) {
  float reg = 0.f;
  if (lr_dev) lr = *lr_dev;  // CUDA-graph friendly: the step-dependent rate lives in device memory
  // n4 = end of this launch's range in float4 units, q0 = its start (0 and the arena length unless this rank owns
  // one slice of a data-parallel arena, see ugn_dp_optim_step)
  for (long long q = q0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n4;
       q += (long long)gridDim.x * blockDim.x) {
    int s = seg_find(off, S, q * 4);
    float c = l2[s];
    // FROZEN segment (layer.trainable = False, nets/mj_uwyhNets_ba.py:1366-1391): the coefficient table stores
    // -(c + 1); the regulariser value still counts (Keras keeps the losses of non-trainable layers), weights and
    // optimiser state stay untouched, nothing is exchanged or re-split for it
    if (c < 0.f) {
      c = -c - 1.f;
      if (c > 0.f) {
        const float4 fw = reinterpret_cast<const float4*>(w)[q];
        reg += c * (fw.x * fw.x + fw.y * fw.y + fw.z * fw.z + fw.w * fw.w);
      }
      continue;
    }
    float4 wv = reinterpret_cast<float4*>(w)[q];
    float4 gv;
    bool staged = false;
    if (DP && peers.nsr) {
#pragma unroll 4
      for (int r = 0; r < peers.nsr; ++r) staged = staged || (q >= peers.srange[2 * r] && q < peers.srange[2 * r + 1]);
    }
    if (DP && staged) {
      gv = reinterpret_cast<const float4*>(g)[q];
      for (int p = 0; p < peers.n; ++p) {
        if (p == peers.rank) continue;
        const float4 o = peers.stage[p * peers.slice4 + (q - q0)];
        gv.x += o.x; gv.y += o.y; gv.z += o.z; gv.w += o.w;
      }
    } else if (DP && peers.mc_g) {
      gv = multimem_ld_reduce_add(peers.mc_g + 4 * q);       // reduced inside the NVSwitch
    } else if (DP && peers.n > 0) {
      // fused reduce-scatter: this rank owns the slice, the gradient is the sum over the ranks' arenas read
      // through NVLink peer memory (gscale carries the 1/N of the mean)
      gv = reinterpret_cast<const float4*>(peers.g[0])[q];
      for (int p = 1; p < peers.n; ++p) {
        const float4 o = reinterpret_cast<const float4*>(peers.g[p])[q];
        gv.x += o.x; gv.y += o.y; gv.z += o.z; gv.w += o.w;
      }
    } else {
      gv = reinterpret_cast<const float4*>(g)[q];
    }
    float ww[4] = {wv.x, wv.y, wv.z, wv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
    if (OPT == 0) {
      float4 mv = reinterpret_cast<float4*>(m)[q];
      float4 vv = reinterpret_cast<float4*>(v)[q];
      float mm[4] = {mv.x, mv.y, mv.z, mv.w}, v2[4] = {vv.x, vv.y, vv.z, vv.w};
      float vh[4] = {0.f, 0.f, 0.f, 0.f};
      if (vhat) {      // AMSGrad (optimizers.Adam(amsgrad=True)): the denominator uses the running maximum of v
        float4 hv = reinterpret_cast<float4*>(vhat)[q];
        vh[0] = hv.x; vh[1] = hv.y; vh[2] = hv.z; vh[3] = hv.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        reg += c * ww[i] * ww[i];
        float gr = gg[i] * gscale + 2.f * c * ww[i];
        mm[i] = b1 * mm[i] + (1.f - b1) * gr;
        v2[i] = b2 * v2[i] + (1.f - b2) * gr * gr;
        float den = v2[i];
        if (vhat) { vh[i] = fmaxf(vh[i], v2[i]); den = vh[i]; }
        // tfa.optimizers.AdamW: decoupled decay var -= wd * var (not scaled by the learning rate), then Adam
        ww[i] = ww[i] - wd * ww[i] - lr * mm[i] / (sqrtf(den) + eps);
      }
      reinterpret_cast<float4*>(m)[q] = make_float4(mm[0], mm[1], mm[2], mm[3]);
      reinterpret_cast<float4*>(v)[q] = make_float4(v2[0], v2[1], v2[2], v2[3]);
      if (vhat) reinterpret_cast<float4*>(vhat)[q] = make_float4(vh[0], vh[1], vh[2], vh[3]);
    } else {
      float4 vv = reinterpret_cast<float4*>(v)[q];
      float v2[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        reg += c * ww[i] * ww[i];
        float gr = gg[i] * gscale + 2.f * c * ww[i];
        v2[i] = b1 * v2[i] - lr * gr;
        ww[i] += v2[i];
      }
      reinterpret_cast<float4*>(v)[q] = make_float4(v2[0], v2[1], v2[2], v2[3]);
    }
    reinterpret_cast<float4*>(w)[q] = make_float4(ww[0], ww[1], ww[2], ww[3]);
    // fused all-gather: the owner stores the updated weights into every other rank's arena (segments whose 16-bit
    // compute copies are exchanged instead skip the f32 broadcast)
    const bool packed_x = DP && pack && pack[2 * s] && peers.cw[0];
    if (DP && !packed_x) {
      if (peers.mc_w) {
        multimem_st(peers.mc_w + 4 * q, make_float4(ww[0], ww[1], ww[2], ww[3]));
      } else {
        for (int p = 0; p < peers.n; ++p)
          if (peers.w[p] != w) reinterpret_cast<float4*>(peers.w[p])[q] = make_float4(ww[0], ww[1], ww[2], ww[3]);
      }
    }
    // fused refresh of the tensor-core compute copy of this segment ([P][numel] 16-bit planes): saves the
    // separate f32 re-read of ugn_pack_weight for the dense weights (92 % of the parameter bytes)
    if (pack && pack[2 * s]) {
      u16* dst = reinterpret_cast<u16*>(pack[2 * s]);
      const long long numel = pack[2 * s + 1], li = q * 4 - off[s];
      __align__(8) u16 hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ugn_split16(ww[i], f16, hi[i], lo[i]);
      if (li + 4 <= numel) {
        *reinterpret_cast<uint2*>(dst + li) = *reinterpret_cast<const uint2*>(hi);
        if (packP == 2) *reinterpret_cast<uint2*>(dst + numel + li) = *reinterpret_cast<const uint2*>(lo);
        if (packed_x && !peers.cw_defer) {
          const long long ob = reinterpret_cast<unsigned char*>(dst + li) - peers.cw[peers.rank];   // byte offset in the arena
          if (peers.cw_mc) {
            multimem_st8(peers.cw_mc + ob, *reinterpret_cast<const uint2*>(hi));
            if (packP == 2) multimem_st8(peers.cw_mc + ob + 2 * numel, *reinterpret_cast<const uint2*>(lo));
          } else {
            for (int p = 0; p < peers.n; ++p) {
              if (p == peers.rank) continue;
              *reinterpret_cast<uint2*>(peers.cw[p] + ob) = *reinterpret_cast<const uint2*>(hi);
              if (packP == 2) *reinterpret_cast<uint2*>(peers.cw[p] + ob + 2 * numel) = *reinterpret_cast<const uint2*>(lo);
            }
          }
        }
      } else {
        for (int i = 0; i < 4 && li + i < numel; ++i) {
          dst[li + i] = hi[i];
          if (packP == 2) dst[numel + li + i] = lo[i];
        }
      }
    }
  }
  if (reg_out) {
    __shared__ float red[8];
    reg = warp_sum(reg);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = reg;
    __syncthreads();
    if (threadIdx.x < 32) {
      float r = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
      r = warp_sum(r);
      if (threadIdx.x == 0) {
        if (DP && peers.reg[0]) {
          // every rank ends up with the sum over the slices: system-scope adds into all ranks' scalars (a few thousand
          // 4-byte NVLink atomics) instead of a 4-byte all-reduce on the step's critical path
          for (int p = 0; p < peers.n; ++p) atomicAdd_system(peers.reg[p], r);
        } else {
          atomicAdd(reg_out, r);
        }
      }
    }
  }
}

int ew_optim(ugn_ctx* ctx, int opt, float* w, const float* g, float* m, float* v,
             const long long* off, const float* l2, int S, long long n, float lr, float b1,
             float b2, float eps, float gscale, float* reg_out, const float* lr_dev, const long long* pack,
             int packP, int f16, float* vhat, float wd, cudaStream_t st, int world, int rank,
             const long long* g_peers, const long long* w_peers, long long g_mc, long long w_mc,
             const long long* reg_peers, const float* stage, const long long* staged_ranges, int n_ranges,
             const long long* cw_peers, long long cw_mc) {
  UGN_CHECK(n % 4 == 0, "optimizer arena length must be a multiple of 4 (got %lld)", n);
  // reg_peers: the caller zeroed every rank's scalar BEFORE the group barrier (a memset here would race with the adds
  // of a faster peer)
  if (reg_out && !(world > 1 && reg_peers)) UGN_CUDA(cudaMemsetAsync(reg_out, 0, sizeof(float), st));
  long long n4 = n / 4, q0 = 0;
  PeerArenas peers;
  if (world > 1) {
    UGN_CHECK(world <= 8 && rank >= 0 && rank < world && g_peers && w_peers, "dp optimizer: bad rank / world / peer tables");
    peers.n = world;
    for (int p = 0; p < world; ++p) {
      peers.g[p] = reinterpret_cast<const float*>(g_peers[p]);
      peers.w[p] = reinterpret_cast<float*>(w_peers[p]);
      if (reg_peers) peers.reg[p] = reinterpret_cast<float*>(reg_peers[p]);
    }
    UGN_CHECK(peers.w[rank] == w, "dp optimizer: w_peers[rank] must be this rank's own arena");
    if (g_mc && w_mc) {
      peers.mc_g = reinterpret_cast<const float*>(g_mc);
      peers.mc_w = reinterpret_cast<float*>(w_mc);
    }
    const long long per = (n4 + world - 1) / world;           // this rank's slice, in float4 units
    q0 = std::min(n4, per * rank);
    n4 = std::min(n4, per * (rank + 1));
    peers.rank = rank;
    if (cw_peers && pack) {
      for (int p = 0; p < world; ++p) peers.cw[p] = reinterpret_cast<unsigned char*>(cw_peers[p]);
      peers.cw_defer = cw_mc == -1;          // UGN_CW_DEFERRED
      peers.cw_mc = peers.cw_defer ? nullptr : reinterpret_cast<unsigned char*>(cw_mc);
    }
    if (stage && n_ranges > 0) {
      UGN_CHECK(n_ranges <= 4 && staged_ranges, "dp optimizer: at most 4 staged ranges");
      peers.stage = reinterpret_cast<const float4*>(stage);
      peers.slice4 = per; peers.rank = rank; peers.nsr = n_ranges;
      for (int r = 0; r < n_ranges; ++r) {
        UGN_CHECK(staged_ranges[2 * r] % 4 == 0 && staged_ranges[2 * r + 1] % 4 == 0, "staged ranges must be multiples of 4 elements");
        peers.srange[2 * r] = staged_ranges[2 * r] / 4;
        peers.srange[2 * r + 1] = staged_ranges[2 * r + 1] / 4;
      }
    }
  }
  int grid = (int)std::min<long long>((n4 - q0 + 255) / 256, (long long)ctx->sm_count * 8);
  grid = std::max(grid, 1);
#define UGN_OPTIM(O, D, VH, WD) \
  optim_kernel<O, D><<<grid, 256, 0, st>>>(w, g, m, v, off, l2, S, n4, lr, b1, b2, eps, gscale, reg_out, lr_dev, pack, packP, f16, VH, WD, peers, q0)
  if (world > 1) { if (opt == 0) UGN_OPTIM(0, true, vhat, wd); else UGN_OPTIM(1, true, nullptr, 0.f); }
  else { if (opt == 0) UGN_OPTIM(0, false, vhat, wd); else UGN_OPTIM(1, false, nullptr, 0.f); }
#undef UGN_OPTIM
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// y = act(y + bias[col]) * mask   (post pass of the split-K dense forward)
__global__ void bias_act_mask_kernel(float* __restrict__ y, const float* __restrict__ bias,
                                     const float* __restrict__ mask, long long n, int cols, int act, float alpha) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float v = y[e] + (bias ? bias[e % cols] : 0.f);
    v = ugn_act_fwd(v, act, alpha);
    if (mask) v *= mask[e];
    y[e] = v;
  }
}
int ew_bias_act_mask(ugn_ctx* ctx, float* y, const float* bias, const float* mask, long long rows, int cols,
                     int act, float alpha, cudaStream_t st) {
  long long n = rows * cols;
  bias_act_mask_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(y, bias, mask, n, cols, act, alpha);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// Dropout masks from a counter-based generator (Philox4x32-10, Salmon et al. 2011) instead of mask tensors: element e of
// layer `layer` at step rng[1] with seed rng[0] -> keep (scaled 1/keep, Keras inverted dropout, nets/mj_uwyhNets_ba.py:100)
// or 0.  Forward and backward post passes regenerate the same bits, so no mask tensor exists; the step counter lives in
// device memory and is advanced by a kernel, so a captured CUDA graph draws fresh masks on every replay.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3, unsigned k0, unsigned k1) {
  const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
  const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
  c1 = (unsigned)p1; c3 = (unsigned)p0; c0 = n0; c2 = n2;
}
__device__ __forceinline__ float philox_keep(const unsigned long long* __restrict__ rng, int layer, long long e, float keep) {
  const unsigned long long seed = rng[0], step = rng[1];
  unsigned c0 = (unsigned)(e >> 2), c1 = (unsigned)((e >> 34) | ((unsigned long long)layer << 16)), c2 = (unsigned)step,
           c3 = (unsigned)(step >> 32);
  unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const unsigned lane = (unsigned)(e & 3);
  const unsigned x = lane == 0 ? c0 : (lane == 1 ? c1 : (lane == 2 ? c2 : c3));
  const float u = (float)(x >> 8) * 5.9604644775390625e-08f;        // [0, 1)
  return u < keep ? 1.f / keep : 0.f;
}
__global__ void dropout_advance_kernel(unsigned long long* rng) { rng[1] += 1ull; }
__global__ void dropout_mask_kernel(const unsigned long long* __restrict__ rng, int layer, float keep, float* __restrict__ out,
                                    long long n) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    out[e] = philox_keep(rng, layer, e, keep);
}
int ew_dropout_advance(ugn_ctx* ctx, unsigned long long* rng, cudaStream_t st) {
  dropout_advance_kernel<<<1, 1, 0, st>>>(rng);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}
int ew_dropout_mask(ugn_ctx* ctx, const unsigned long long* rng, int layer, float keep, float* out, long long n,
                    cudaStream_t st) {
  dropout_mask_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(rng, layer, keep, out, n);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ONE post pass of a split-K dense GEMM (forward: bias + activation + dropout mask -> f32 result + its 16-bit planes;
// backward: dropout mask of the layer below -> that layer's scaled 16-bit gradient operand + its bias gradient).
// Block = 32 columns x 8 row lanes over ALL rows, so the column sums need no atomics.
template <int P16>
__global__ void __launch_bounds__(256) dense_post_kernel(float* __restrict__ y, const float* __restrict__ bias,
                                                         const float* __restrict__ mask, u16* __restrict__ out16,
                                                         long long plane, int rows, int cols, int act, float alpha,
                                                         int f16, const float* __restrict__ scale16,
                                                         float* __restrict__ colsum, int write_f32,
                                                         const unsigned long long* __restrict__ rng, int layer,
                                                         float keep) {
  __shared__ float sm[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const float s16 = scale16 ? *scale16 : 1.f;
  float acc = 0.f;
  if (j < cols) {
    const float bj = bias ? bias[j] : 0.f;
    for (int r = threadIdx.y; r < rows; r += 8) {
      const long long e = (long long)r * cols + j;
      float v = ugn_act_fwd(y[e] + bj, act, alpha);
      if (rng) v *= philox_keep(rng, layer, e, keep);
      else if (mask) v *= mask[e];
      if (write_f32) y[e] = v;
      if (P16 > 0) {
        u16 hi, lo;
        ugn_split16(v * s16, f16, hi, lo);
        out16[e] = hi;
        if (P16 == 2) out16[plane + e] = lo;
      }
      acc += v;
    }
  }
  if (colsum) {
    sm[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && j < cols) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += sm[q][threadIdx.x];
      colsum[j] = t;
    }
  }
}
int ew_dense_post(ugn_ctx* ctx, float* y, const float* bias, const float* mask, __nv_bfloat16* out16, int P16, int f16,
                  const float* scale16, float* colsum, int write_f32, long long rows, int cols, int act, float alpha,
                  cudaStream_t st, const unsigned long long* rng, int layer, float keep) {
  dim3 grid(ugn_cdiv(cols, 32)), block(32, 8);
  const long long plane = rows * cols;
  u16* o = reinterpret_cast<u16*>(out16);
#define UGN_DP_ARGS y, bias, mask, o, plane, (int)rows, cols, act, alpha, f16, scale16, colsum, write_f32, rng, layer, keep
  if (!out16 || P16 == 0) dense_post_kernel<0><<<grid, block, 0, st>>>(UGN_DP_ARGS);
  else if (P16 == 1) dense_post_kernel<1><<<grid, block, 0, st>>>(UGN_DP_ARGS);
  else dense_post_kernel<2><<<grid, block, 0, st>>>(UGN_DP_ARGS);
#undef UGN_DP_ARGS
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// Gradient scale for 16-bit (fp16) gradient operands.  ctx->gscale = {s, 1/s, amax bits, -}.
// ugn_grad_scale_update: s = 2^floor(log2(target / max|ref|)) from a reference gradient of the step
// (the signature gradient), so every dz = s * true gradient sits in fp16's normal range; consumers
// (dgrad, wgrad, dense backward, bias sums) multiply their f32 results by 1/s.  Power-of-two scales
// are exact in floating point, so the only effect is on range.
// ---------------------------------------------------------------------------------------
__global__ void amax_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ amax_bits) {
  float m = 0.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float a = fabsf(x[e]);
    if (a < INFINITY) m = fmaxf(m, a);   // NaN / inf never drive the scale
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
}
__global__ void gscale_set_kernel(float* __restrict__ gs, float target, float fixed) {
  float s = fixed;
  if (!(fixed > 0.f)) {
    float a = __uint_as_float(reinterpret_cast<unsigned*>(gs)[2]);
    s = a > 0.f ? exp2f(floorf(log2f(target / a))) : 1.f;
    s = fminf(fmaxf(s, 1.f / 1024.f), 16777216.f);
  }
  gs[0] = s;
  gs[1] = 1.f / s;
  reinterpret_cast<unsigned*>(gs)[2] = 0u;
}
int ew_gscale_update(ugn_ctx* ctx, const float* ref, long long n, float target, float fixed, cudaStream_t st) {
  UGN_CHECK(ctx->gscale, "gradient scale buffer missing");
  if (!(fixed > 0.f)) {
    amax_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(ref, n, reinterpret_cast<unsigned*>(ctx->gscale) + 2);
    UGN_LAUNCHED(ctx);
  }
  gscale_set_kernel<<<1, 1, 0, st>>>(ctx->gscale, target, fixed);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// split-K epilogue of a small convolution: f32 partial sums [rows][cols] -> act(acc + bias) -> 16-bit planes
__global__ void bias_act_split16_kernel(const float* __restrict__ acc, const float* __restrict__ bias,
                                        u16* __restrict__ out, long long n, int cols, int P, int f16, int act,
                                        float alpha) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float v = ugn_act_fwd(acc[e] + (bias ? bias[e % cols] : 0.f), act, alpha);
    u16 hi, lo;
    ugn_split16(v, f16, hi, lo);
    out[e] = hi;
    if (P == 2) out[n + e] = lo;
  }
}
int ew_bias_act_split16(ugn_ctx* ctx, const float* acc, const float* bias, __nv_bfloat16* out, long long rows,
                        int cols, int P, int f16, int act, float alpha, cudaStream_t st) {
  long long n = rows * cols;
  bias_act_split16_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(acc, bias, out, n, cols, P, f16, act, alpha);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// a13: video-level summaries of the open-world test (mains/mj_testUWYHGaitNet_open_tum.py:355-420).
// Rows are grouped by video through a CSR (order i32 [N] = row indices sorted by video, offsets
// i32 [V+1]); the host builds it with np.unique exactly as the reference does.
//   segment_pool : out[v] = mean (use_avg) or max over the video's sub-sequence descriptors, summed in
//                  row order in fp32 like numpy's axis-0 reduction
//   segment_mode : statistics.mode of the video's labels: most frequent, ties -> first encountered
//                  (Python >= 3.8); legacy != 0 -> on a tie the first element (the reference's
//                  `except: ...[idx][0]` branch under Python < 3.8)
// ---------------------------------------------------------------------------------------
__global__ void segment_pool_kernel(const float* __restrict__ codes, const int* __restrict__ order,
                                    const int* __restrict__ offsets, int D, int use_avg, float* __restrict__ out) {
  const int v = blockIdx.x, j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j >= D) return;
  const int a = offsets[v], b = offsets[v + 1];
  float acc = use_avg ? 0.f : -INFINITY;
  for (int e = a; e < b; ++e) {
    const float x = codes[(long long)order[e] * D + j];
    acc = use_avg ? acc + x : fmaxf(acc, x);
  }
  out[(long long)v * D + j] = use_avg ? acc / (float)(b - a) : acc;
}
__global__ void segment_mode_kernel(const int* __restrict__ labels, const int* __restrict__ order,
                                    const int* __restrict__ offsets, int V, int legacy, int* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int a = offsets[v], b = offsets[v + 1];
  int best = labels[order[a]], bestc = 0, ties = 0;
  for (int e = a; e < b; ++e) {
    const int le = labels[order[e]];
    bool first = true;
    for (int f = a; f < e; ++f) first = first && labels[order[f]] != le;
    if (!first) continue;                       // count every distinct label once, in order of appearance
    int c = 0;
    for (int f = e; f < b; ++f) c += labels[order[f]] == le;
    if (c > bestc) { bestc = c; best = le; ties = 0; }
    else if (c == bestc) ++ties;
  }
  out[v] = (legacy && ties) ? labels[order[a]] : best;
}
int ew_segment_pool(ugn_ctx* ctx, const float* codes, const int* order, const int* offsets, int V, int D,
                    int use_avg, float* out, cudaStream_t st) {
  if (V == 0 || D == 0) return UGN_OK;
  segment_pool_kernel<<<dim3(V, ugn_cdiv(D, 256)), 256, 0, st>>>(codes, order, offsets, D, use_avg, out);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}
int ew_segment_mode(ugn_ctx* ctx, const int* labels, const int* order, const int* offsets, int V, int legacy,
                    int* out, cudaStream_t st) {
  if (V == 0) return UGN_OK;
  segment_mode_kernel<<<ugn_cdiv(V, 128), 128, 0, st>>>(labels, order, offsets, V, legacy, out);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}
