// Tensor-core (tcgen05 / TMEM / TMA) path entry points, used when operands are bf16.
#pragma once
#include "common.cuh"

int tc_gemm(ugn_ctx* ctx, int P, int f16, int M, int N, int K, const __nv_bfloat16* A, int a_mn,
            const __nv_bfloat16* B, int b_mn, float* C, int accumulate, cudaStream_t st);
int tc_conv_fwd(ugn_ctx* ctx, const ConvGeom& g, int P, int f16, const __nv_bfloat16* x, const __nv_bfloat16* w,
                const float* bias, __nv_bfloat16* y, uint8_t* idx, int act, float alpha, int pool,
                cudaStream_t st);
int tc_conv_dgrad(ugn_ctx* ctx, const ConvGeom& g, int P, int f16, const __nv_bfloat16* dz, const __nv_bfloat16* w,
                  float* dx, cudaStream_t st);
int tc_conv_wgrad(ugn_ctx* ctx, const ConvGeom& g, int P, int f16, const __nv_bfloat16* x, const __nv_bfloat16* dz,
                  float* dw, float* db, cudaStream_t st);
// y16 (nullable): [P16][B][N] 16-bit planes of y written by the same epilogue (next tensor-core consumer)
int tc_linear_fwd(ugn_ctx* ctx, int P, int f16, int B, int N, int K, const __nv_bfloat16* x, const __nv_bfloat16* w,
                  const float* bias, const float* mask, float* y, int act, float alpha, cudaStream_t st,
                  __nv_bfloat16* y16 = nullptr, int P16 = 0);
// dx_mask / dx16 / dbx (each nullable): the input-gradient GEMM's epilogue multiplies by the dropout mask of the layer
// below, writes the (still scaled) 16-bit gradient operand of that layer and its bias gradient (column sums);
// dx itself may then be null
int tc_linear_bwd(ugn_ctx* ctx, int P, int f16, int B, int N, int K, const __nv_bfloat16* x, const __nv_bfloat16* w,
                  const __nv_bfloat16* dz, float* dx, float* dw, float* db, cudaStream_t st,
                  const float* dx_mask = nullptr, __nv_bfloat16* dx16 = nullptr, int P16 = 0, float* dbx = nullptr);
