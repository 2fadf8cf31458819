// fp32 validation-mode kernels: generic mixed-radix implicit GEMM + the conv / dense
// wrappers that use it when the operands are f32.
#include "simt.cuh"

#define TBM 64
#define TBN 64
#define TBK 16

__device__ __forceinline__ int radix_decode(const Radix& R, int idx, int& y, int& x) {
  int off = 0, rem = idx;
  y = 0; x = 0;
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    if (d < R.n) {
      int q;
      if (d == R.n - 1) { q = rem; }
      else { q = rem % R.r[d]; rem /= R.r[d]; }
      off += q * R.s[d];
      if (d == R.cy) y = q;
      if (d == R.cx) x = q;
    }
  }
  return off;
}

__global__ void __launch_bounds__(256) simt_gemm_kernel(const SGemm p) {
  __shared__ __align__(16) float As[TBK][TBM + 4];
  __shared__ __align__(16) float Bs[TBK][TBN + 4];
  __shared__ int a_off[TBM], b_off[TBN], a_yx[TBM];

  const int t = threadIdx.x;
  const int i0 = blockIdx.x * TBM;
  const int j0 = blockIdx.y * TBN;
  const float* A = p.A;
  const float* B = p.B;
  float* C = p.C;
  int kbeg = 0, kend = p.K;
  if (p.batches > 1) {
    A += (long long)blockIdx.z * p.batch_a;
    B += (long long)blockIdx.z * p.batch_b;
    C += (long long)blockIdx.z * p.batch_c;
  } else if (p.ksplit > 1) {
    int len = (p.K + p.ksplit - 1) / p.ksplit;
    len = (len + TBK - 1) / TBK * TBK;
    kbeg = blockIdx.z * len;
    kend = min(p.K, kbeg + len);
  }

  if (t < TBM) {
    int i = i0 + t, y = 0, x = 0;
    a_off[t] = (i < p.M) ? radix_decode(p.ar, i, y, x) : 0;
    a_yx[t] = (y << 16) | x;
  } else if (t < TBM + TBN) {
    int j = j0 + (t - TBM), y, x;
    b_off[t - TBM] = (j < p.N) ? radix_decode(p.br, j, y, x) : 0;
  }
  __syncthreads();

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  const int kk = t & 15, r = t >> 4;
  const int ty = t >> 4, tx = t & 15;

  for (int k0 = kbeg; k0 < kend; k0 += TBK) {
    int k = k0 + kk;
    bool kv = k < kend;
    int ky = 0, kx = 0, dy, dx;
    int aoffk = kv ? radix_decode(p.ak, k, ky, kx) : 0;
    int boffk = kv ? radix_decode(p.bk, k, dy, dx) : 0;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      int row = r + 16 * jj;
      bool va = kv && (i0 + row < p.M);
      if (va && p.check) {
        int yy = (a_yx[row] >> 16) - ky, xx = (a_yx[row] & 0xffff) - kx;
        va = (yy >= 0) && (yy < p.limY) && (xx >= 0) && (xx < p.limX);
      }
      As[kk][row] = va ? __ldg(A + (long long)a_off[row] + aoffk) : 0.f;
      bool vb = kv && (j0 + row < p.N);
      Bs[kk][row] = vb ? __ldg(B + (long long)b_off[row] + boffk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < TBK; ++q) {
      float4 av = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[q][tx * 4]);
      float a4[4] = {av.x, av.y, av.z, av.w};
      float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
    }
    __syncthreads();
  }

  const int ib = i0 + ty * 4, jb = j0 + tx * 4;
  if (p.epi == EPI_POOL4) {
    // rows ib..ib+3 are the (dy,dx) positions of one 2x2 window (pooled row order)
    if (ib >= p.M) return;
    long long orow = (long long)(ib >> 2) * p.ldc;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int j = jb + b;
      if (j >= p.N) continue;
      float bs = p.bias ? p.bias[j] : 0.f;
      float best = ugn_act_fwd(acc[0][b] + bs, p.act, p.alpha);
      int pos = 0;
#pragma unroll
      for (int a = 1; a < 4; ++a) {
        float v = ugn_act_fwd(acc[a][b] + bs, p.act, p.alpha);
        if (v > best) { best = v; pos = a; }
      }
      C[orow + j] = best;
      p.pool_idx[orow + j] = (uint8_t)pos;
    }
    return;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int i = ib + a;
    if (i >= p.M) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int j = jb + b;
      if (j >= p.N) continue;
      long long o = (long long)i * p.ldc + j;
      if (p.epi == EPI_ATOMIC) {
        atomicAdd(C + o, acc[a][b] * p.out_scale);
      } else {
        float v = acc[a][b] + (p.bias ? p.bias[j] : 0.f);
        v = ugn_act_fwd(v, p.act, p.alpha);
        if (p.mask) v *= p.mask[o];
        C[o] = v * p.out_scale;
      }
    }
  }
}

int simt_gemm_launch(ugn_ctx* ctx, const SGemm& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return UGN_OK;
  dim3 grid(ugn_cdiv(p.M, TBM), ugn_cdiv(p.N, TBN), p.batches > 1 ? p.batches : p.ksplit);
  UGN_CHECK(grid.y <= 65535 && grid.z <= 65535, "simt gemm grid too large");
  simt_gemm_kernel<<<grid, 256, 0, st>>>(p);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// column sums: out[j] = sum_i X[i*ld + j]   (bias gradients)
// ---------------------------------------------------------------------------------------
__global__ void colsum_kernel(const float* __restrict__ X, long long rows, int cols, int ld,
                              float* __restrict__ out) {
  // block: 32 columns x 8 row-lanes; grid.y splits the rows; atomics combine
  __shared__ float sm[8][33];
  int j = blockIdx.x * 32 + threadIdx.x;
  long long per = (rows + gridDim.y - 1) / gridDim.y;
  long long r0 = (long long)blockIdx.y * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  if (j < cols)
    for (long long i = r0 + threadIdx.y; i < r1; i += 8) s += X[i * ld + j];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j < cols) {
    float tsum = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) tsum += sm[q][threadIdx.x];
    atomicAdd(out + j, tsum);
  }
}

int simt_colsum(ugn_ctx* ctx, const float* X, long long rows, int cols, int ld, float* out,
                cudaStream_t st) {
  UGN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
  int gy = (int)std::min<long long>(std::max<long long>(rows / 256, 1), 256);
  dim3 grid(ugn_cdiv(cols, 32), gy), block(32, 8);
  colsum_kernel<<<grid, block, 0, st>>>(X, rows, cols, ld, out);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// conv / dense in f32 mode
// ---------------------------------------------------------------------------------------
int simt_conv_fwd(ugn_ctx* ctx, const ConvGeom& g, const float* x, const float* w,
                  const float* bias, float* y, uint8_t* idx, int act, float alpha, int pool,
                  cudaStream_t st) {
  SGemm p;
  p.A = x; p.B = w; p.C = y;
  p.N = g.Co; p.K = g.KH * g.KW * g.Cp;
  p.ak = radix3(g.Cp, 1, g.KW, g.Cp, g.W * g.Cp);
  p.br = radix1(p.K);
  p.bk = radix1(1);
  p.bias = bias; p.act = act; p.alpha = alpha; p.ldc = g.Co;
  if (pool) {
    p.M = g.B * g.Hp * g.Wp * 4;
    Radix R;
    R.n = 5;
    R.r[0] = 2; R.r[1] = 2; R.r[2] = g.Wp; R.r[3] = g.Hp;
    R.s[0] = g.Cp; R.s[1] = g.W * g.Cp; R.s[2] = 2 * g.Cp; R.s[3] = 2 * g.W * g.Cp;
    R.s[4] = g.H * g.W * g.Cp;
    p.ar = R;
    p.epi = EPI_POOL4;
    p.pool_idx = idx;
  } else {
    p.M = g.B * g.Ho * g.Wo;
    p.ar = radix3(g.Wo, g.Cp, g.Ho, g.W * g.Cp, g.H * g.W * g.Cp);
    p.epi = EPI_STORE;
  }
  return simt_gemm_launch(ctx, p, st);
}

int simt_conv_dgrad(ugn_ctx* ctx, const ConvGeom& g, const float* dz, const float* w, float* dx,
                    cudaStream_t st) {
  SGemm p;
  p.A = dz; p.B = w; p.C = dx;
  p.M = g.B * g.H * g.W; p.N = g.Cp; p.K = g.KH * g.KW * g.Co;
  p.ar = radix3(g.W, g.Co, g.H, g.Wo * g.Co, g.Ho * g.Wo * g.Co, /*cy=*/1, /*cx=*/0);
  p.ak = radix3(g.Co, 1, g.KW, -g.Co, -g.Wo * g.Co, /*cy=*/2, /*cx=*/1);
  p.check = 1; p.limY = g.Ho; p.limX = g.Wo;
  p.br = radix1(1);
  p.bk = radix3(g.Co, g.KH * g.KW * g.Cp, g.KW, g.Cp, g.KW * g.Cp);
  p.ldc = g.Cp; p.epi = EPI_STORE;
  return simt_gemm_launch(ctx, p, st);
}

int simt_colsum(ugn_ctx*, const float*, long long, int, int, float*, cudaStream_t);

int simt_conv_wgrad(ugn_ctx* ctx, const ConvGeom& g, const float* x, const float* dz, float* dw,
                    float* db, cudaStream_t st) {
  SGemm p;
  p.A = dz; p.B = x; p.C = dw;
  p.M = g.Co; p.N = g.KH * g.KW * g.Cin; p.K = g.B * g.Ho * g.Wo;
  p.ar = radix1(1);
  p.ak = radix1(g.Co);
  p.br = radix3(g.Cin, 1, g.KW, g.Cp, g.W * g.Cp);
  p.bk = radix3(g.Wo, g.Cp, g.Ho, g.W * g.Cp, g.H * g.W * g.Cp);
  p.ldc = p.N;
  int blocks = ugn_cdiv(p.M, TBM) * ugn_cdiv(p.N, TBN);
  int want = ugn_cdiv(4 * ctx->sm_count, blocks);
  int maxs = std::max(1, p.K / 128);
  p.ksplit = std::max(1, std::min(want, maxs));
  p.epi = EPI_ATOMIC;
  UGN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)p.M * p.N, st));
  int rc = simt_gemm_launch(ctx, p, st);
  if (rc != UGN_OK) return rc;
  if (db) return simt_colsum(ctx, dz, (long long)p.K, g.Co, g.Co, db, st);
  return UGN_OK;
}

int ew_bias_act_mask(ugn_ctx* ctx, float* y, const float* bias, const float* mask, long long rows, int cols,
                     int act, float alpha, cudaStream_t st);

int simt_linear_fwd(ugn_ctx* ctx, int B, int N, int K, const float* x, const float* w,
                    const float* bias, const float* mask, float* y, int act, float alpha,
                    cudaStream_t st) {
  SGemm p;
  p.A = x; p.B = w; p.C = y;
  p.M = B; p.N = N; p.K = K;
  p.ar = radix1(K); p.ak = radix1(1); p.br = radix1(K); p.bk = radix1(1);
  p.ldc = N;
  // few output tiles but a long reduction (the "classprob" head: 96 x 150 x 2048): spread K over the SMs
  // (split-K, red.add into zeroed y) and apply bias / activation / dropout mask in a post pass
  const int blocks = ugn_cdiv(B, TBM) * ugn_cdiv(N, TBN);
  if (blocks * 4 <= ctx->sm_count && K >= 512) {
    p.ksplit = std::max(1, std::min(2 * ctx->sm_count / blocks, K / 64));
    p.epi = EPI_ATOMIC;
    UGN_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)B * N, st));
    int rc = simt_gemm_launch(ctx, p, st);
    if (rc != UGN_OK) return rc;
    return ew_bias_act_mask(ctx, y, bias, mask, B, N, act, alpha, st);
  }
  p.bias = bias; p.mask = mask; p.act = act; p.alpha = alpha;
  return simt_gemm_launch(ctx, p, st);
}

int simt_linear_bwd(ugn_ctx* ctx, int B, int N, int K, const float* x, const float* w,
                    const float* dz, float* dx, float* dw, float* db, cudaStream_t st) {
  int rc;
  if (dx) {
    SGemm p;
    p.A = dz; p.B = w; p.C = dx;
    p.M = B; p.N = K; p.K = N;
    p.ar = radix1(N); p.ak = radix1(1); p.br = radix1(1); p.bk = radix1(K);
    p.ldc = K;
    if ((rc = simt_gemm_launch(ctx, p, st)) != UGN_OK) return rc;
  }
  if (dw) {
    SGemm p;
    p.A = dz; p.B = x; p.C = dw;
    p.M = N; p.N = K; p.K = B;
    p.ar = radix1(1); p.ak = radix1(N); p.br = radix1(1); p.bk = radix1(K);
    p.ldc = K;
    if ((rc = simt_gemm_launch(ctx, p, st)) != UGN_OK) return rc;
  }
  if (db) return simt_colsum(ctx, dz, B, N, N, db, st);
  return UGN_OK;
}

// column sums of a bf16 [P][rows][cols] tensor (hi + lo planes): bias gradients in tensor-core mode
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, int P, long long rows, int cols,
                                   float* __restrict__ out, int f16, const float* __restrict__ gs) {
  __shared__ float sm[8][33];
  int j = blockIdx.x * 32 + threadIdx.x;
  long long per = (rows + gridDim.y - 1) / gridDim.y;
  long long r0 = (long long)blockIdx.y * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  if (j < cols)
    for (int pl = 0; pl < P; ++pl)
      for (long long i = r0 + threadIdx.y; i < r1; i += 8) s += ugn_f16to32(X[((long long)pl * rows + i) * cols + j], f16);
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j < cols) {
    float tsum = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) tsum += sm[q][threadIdx.x];
    atomicAdd(out + j, tsum * (gs ? gs[1] : 1.f));
  }
}

// X is a (scaled) 16-bit gradient operand: the sum is multiplied by 1/s (ctx->gscale)
int simt_colsum_bf16(ugn_ctx* ctx, const __nv_bfloat16* X, int P, int f16, long long rows, int cols, float* out,
                     cudaStream_t st) {
  UGN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
  int gy = (int)std::min<long long>(std::max<long long>(rows / 256, 1), 256);
  dim3 grid(ugn_cdiv(cols, 32), gy), block(32, 8);
  colsum_bf16_kernel<<<grid, block, 0, st>>>(X, P, rows, cols, out, f16, ctx->gscale);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// Conv3D branches (use3D: UWYHSemiNet.build_3Dbranch{,LReLU}, nets/mj_uwyhNets_ba.py:336-417): strided 'valid'
// channels-last 3-D convolutions.  fp32 validation engine only (the builder option is outside the benchmarked
// configurations): forward and kernel gradient are the generic mixed-radix GEMM with one more digit, the input gradient
// is a gather with stride-divisibility tests in a kernel of its own.
//   x [B,T,H,W,C], w [Co,KT,KH,KW,C], y / dz [B,To,Ho,Wo,Co]
// ---------------------------------------------------------------------------------------
static Radix radix4(int r0, int s0, int r1, int s1, int r2, int s2, int s3) {
  Radix R;
  R.n = 4;
  R.r[0] = r0; R.r[1] = r1; R.r[2] = r2;
  R.s[0] = s0; R.s[1] = s1; R.s[2] = s2; R.s[3] = s3;
  return R;
}

int simt_conv3d_fwd(ugn_ctx* ctx, const Conv3Geom& g, const float* x, const float* w, const float* bias, float* y,
                    int act, float alpha, cudaStream_t st) {
  SGemm p;
  p.A = x; p.B = w; p.C = y;
  p.M = g.B * g.To * g.Ho * g.Wo; p.N = g.Co; p.K = g.KT * g.KH * g.KW * g.C;
  p.ar = radix4(g.Wo, g.SW * g.C, g.Ho, g.SH * g.W * g.C, g.To, g.ST * g.H * g.W * g.C, g.T * g.H * g.W * g.C);
  p.ak = radix4(g.C, 1, g.KW, g.C, g.KH, g.W * g.C, g.H * g.W * g.C);
  p.br = radix1(p.K);
  p.bk = radix1(1);
  p.bias = bias; p.act = act; p.alpha = alpha; p.ldc = g.Co;
  return simt_gemm_launch(ctx, p, st);
}

int simt_conv3d_wgrad(ugn_ctx* ctx, const Conv3Geom& g, const float* x, const float* dz, float* dw, float* db,
                      cudaStream_t st) {
  SGemm p;
  p.A = dz; p.B = x; p.C = dw;
  p.M = g.Co; p.N = g.KT * g.KH * g.KW * g.C; p.K = g.B * g.To * g.Ho * g.Wo;
  p.ar = radix1(1);
  p.ak = radix1(g.Co);
  p.br = radix4(g.C, 1, g.KW, g.C, g.KH, g.W * g.C, g.H * g.W * g.C);
  p.bk = radix4(g.Wo, g.SW * g.C, g.Ho, g.SH * g.W * g.C, g.To, g.ST * g.H * g.W * g.C, g.T * g.H * g.W * g.C);
  p.ldc = p.N;
  int blocks = ugn_cdiv(p.M, TBM) * ugn_cdiv(p.N, TBN);
  int want = ugn_cdiv(4 * ctx->sm_count, blocks);
  int maxs = std::max(1, p.K / 128);
  p.ksplit = std::max(1, std::min(want, maxs));
  p.epi = EPI_ATOMIC;
  UGN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)p.M * p.N, st));
  int rc = simt_gemm_launch(ctx, p, st);
  if (rc != UGN_OK) return rc;
  if (db) return simt_colsum(ctx, dz, (long long)p.K, g.Co, g.Co, db, st);
  return UGN_OK;
}

// dx[b,t,y,x,c] = sum_{kt,ky,kx,co} dz[b,(t-kt)/ST,(y-ky)/SH,(x-kx)/SW,co] * w[co,kt,ky,kx,c] over the taps whose
// differences are non-negative multiples of the strides and land inside the output volume.
// GEMM view: M = B*T*H*W rows, N = C, K = (co fastest, kx, ky, kt).  64 x 64 x 16 tiles as simt_gemm_kernel.
__global__ void __launch_bounds__(256) conv3d_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ w,
                                                           float* __restrict__ dx, const Conv3Geom g) {
  __shared__ __align__(16) float As[TBK][TBM + 4];
  __shared__ __align__(16) float Bs[TBK][TBN + 4];
  __shared__ int r_t[TBM], r_y[TBM], r_x[TBM], r_b[TBM];
  const int t = threadIdx.x;
  const int i0 = blockIdx.x * TBM, j0 = blockIdx.y * TBN;
  const int M = g.B * g.T * g.H * g.W, K3 = g.KT * g.KH * g.KW * g.C, K = g.KT * g.KH * g.KW * g.Co;
  if (t < TBM) {
    int i = i0 + t;
    if (i < M) {
      r_x[t] = i % g.W; i /= g.W;
      r_y[t] = i % g.H; i /= g.H;
      r_t[t] = i % g.T; r_b[t] = i / g.T;
    } else {
      r_x[t] = r_y[t] = r_t[t] = 0; r_b[t] = -1;
    }
  }
  __syncthreads();
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const int kk = t & 15, r = t >> 4;
  const int ty = t >> 4, tx = t & 15;
  for (int k0 = 0; k0 < K; k0 += TBK) {
    const int k = k0 + kk;
    const bool kv = k < K;
    int co = 0, kx = 0, ky = 0, kt = 0;
    if (kv) {
      int q = k;
      co = q % g.Co; q /= g.Co;
      kx = q % g.KW; q /= g.KW;
      ky = q % g.KH; kt = q / g.KH;
    }
    const long long wk = (long long)co * K3 + ((long long)(kt * g.KH + ky) * g.KW + kx) * g.C;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int row = r + 16 * jj;
      float av = 0.f;
      if (kv && r_b[row] >= 0) {
        const int dt = r_t[row] - kt, dy = r_y[row] - ky, dxx = r_x[row] - kx;
        if (dt >= 0 && dy >= 0 && dxx >= 0 && dt % g.ST == 0 && dy % g.SH == 0 && dxx % g.SW == 0) {
          const int to = dt / g.ST, yo = dy / g.SH, xo = dxx / g.SW;
          if (to < g.To && yo < g.Ho && xo < g.Wo)
            av = __ldg(dz + ((((long long)r_b[row] * g.To + to) * g.Ho + yo) * g.Wo + xo) * g.Co + co);
        }
      }
      As[kk][row] = av;
      const int j = j0 + row;
      Bs[kk][row] = (kv && j < g.C) ? __ldg(w + wk + j) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < TBK; ++q) {
      float4 a4v = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
      float4 b4v = *reinterpret_cast<const float4*>(&Bs[q][tx * 4]);
      float a4[4] = {a4v.x, a4v.y, a4v.z, a4v.w};
      float b4[4] = {b4v.x, b4v.y, b4v.z, b4v.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
    }
    __syncthreads();
  }
  const int ib = i0 + ty * 4, jb = j0 + tx * 4;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = ib + a;
    if (i >= M) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = jb + b;
      if (j < g.C) dx[(long long)i * g.C + j] = acc[a][b];
    }
  }
}

int simt_conv3d_dgrad(ugn_ctx* ctx, const Conv3Geom& g, const float* dz, const float* w, float* dx, cudaStream_t st) {
  const int M = g.B * g.T * g.H * g.W;
  if (M <= 0) return UGN_OK;
  dim3 grid(ugn_cdiv(M, TBM), ugn_cdiv(g.C, TBN));
  UGN_CHECK(grid.y <= 65535, "conv3d dgrad: too many input channels");
  conv3d_dgrad_kernel<<<grid, 256, 0, st>>>(dz, w, dx, g);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}
