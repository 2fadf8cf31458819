// fp32 "validation mode" implicit-GEMM engine (SIMT FFMA).
//
// One kernel computes  C[i,j] (op)= sum_k A[aoff(i) + akoff(k)] * B[boff(j) + bkoff(k)]
// where the four offset functions are mixed-radix decodes of the GEMM indices.  That single
// form covers Conv2D forward (nets/mj_uwyhNets_ba.py:82-89), its input / kernel gradients,
// the Dense layers (:97-105) and the Gram matrix of the triplet loss -- each caller only
// fills in the radices / strides.
#pragma once
#include "common.cuh"

struct Radix {
  int n = 1;                 // digits in use (least significant first); last digit unbounded
  int r[4] = {1, 1, 1, 1};   // radices of digits 0..3
  int s[5] = {0, 0, 0, 0, 0};// element stride of each digit
  int cy = -1, cx = -1;      // digits exported as (y, x) coordinates for the validity test
};

enum { EPI_STORE = 0, EPI_POOL4 = 1, EPI_ATOMIC = 2 };

struct SGemm {
  const float* A = nullptr;
  const float* B = nullptr;
  float* C = nullptr;
  int M = 0, N = 0, K = 0;
  Radix ar, ak, br, bk;
  int check = 0, limY = 0, limX = 0;  // A valid iff 0 <= ry-ky < limY and 0 <= rx-kx < limX
  int epi = EPI_STORE;
  int ldc = 0;
  const float* bias = nullptr;  // [N]
  const float* mask = nullptr;  // [M,ldc] multiplicative (dropout)
  int act = UGN_ACT_LINEAR;
  float alpha = 0.f;
  float out_scale = 1.f;
  uint8_t* pool_idx = nullptr;
  int ksplit = 1;               // grid.z; EPI_ATOMIC when > 1
  long long batch_a = 0, batch_b = 0, batch_c = 0;  // blockIdx.z strides when batched (ksplit==1)
  int batches = 1;
};

int simt_gemm_launch(ugn_ctx* ctx, const SGemm& p, cudaStream_t st);

static inline Radix radix1(int s0) {
  Radix R;
  R.n = 1;
  R.s[0] = s0;
  return R;
}
static inline Radix radix3(int r0, int s0, int r1, int s1, int s2, int cy = -1, int cx = -1) {
  Radix R;
  R.n = 3;
  R.r[0] = r0; R.r[1] = r1;
  R.s[0] = s0; R.s[1] = s1; R.s[2] = s2;
  R.cy = cy; R.cx = cx;
  return R;
}

// Conv3D (use3D branches), fp32 validation engine
int simt_conv3d_fwd(ugn_ctx* ctx, const Conv3Geom& g, const float* x, const float* w, const float* bias, float* y,
                    int act, float alpha, cudaStream_t st);
int simt_conv3d_wgrad(ugn_ctx* ctx, const Conv3Geom& g, const float* x, const float* dz, float* dw, float* db,
                      cudaStream_t st);
int simt_conv3d_dgrad(ugn_ctx* ctx, const Conv3Geom& g, const float* dz, const float* w, float* dx, cudaStream_t st);
