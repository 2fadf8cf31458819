// Shared host/device helpers for libugaitnet_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include "../../include/ugaitnet_b200.h"

#define UGN_DL_CUDA 2
#define UGN_DL_INT 0
#define UGN_DL_UINT 1
#define UGN_DL_FLOAT 2
#define UGN_DL_BFLOAT 4

struct ugn_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  long long launches = 0;
  // driver entry point for TMA descriptor encoding (resolved lazily; no libcuda link)
  void* encode_tiled = nullptr;
  // device flag set by a tensor-core kernel whose mbarrier wait timed out (protocol bug guard)
  int* err_flag = nullptr;
  // device-resident gradient scale {s, 1/s}: 16-bit gradient operands (dz) are stored multiplied by s
  // and every kernel that consumes them multiplies its f32 output by 1/s (fp16 range management;
  // lives in device memory so that a captured CUDA graph picks up each step's value)
  float* gscale = nullptr;
  // MMA passes per product of the FORWARD tensor-core layers over split (hi/lo) operands (ugn_set_fwd_passes;
  // 0 = all three): 1 hi*hi, 2 hi*hi + hi*lo(weights), 4 hi*hi + lo(activations)*hi.  gemm_npass is the transient
  // hand-over from tc_linear_fwd to tc_gemm_ex
  int fwd_conv_pass = 0, fwd_dense_pass = 0, gemm_npass = 0;
  // transient (set by the *_philox entry points around one call): Philox dropout source of the dense post pass
  const unsigned long long* drop_rng = nullptr;
  int drop_layer = 0;
  float drop_keep = 1.f;
  int gemm_nosplit = 0;        // transient: tc_gemm_ex without split-K (deterministic accumulation order: triplet Gram)
  // grow-only device scratch (split-K partial sums of small convolutions); sized on first use, i.e. in
  // the warm-up step before any CUDA-graph capture
  // round-robin pool (the engine runs the modality branches on concurrent streams: consecutive calls
  // get different buffers; a buffer comes round again only after kScratchSlots further calls, i.e. more
  // than two steps later, and steps are serialised by the stream joins)
  static constexpr int kScratchSlots = 8;
  void* scratch[kScratchSlots] = {};
  size_t scratch_bytes[kScratchSlots] = {};
  int scratch_next = 0;
};

int ugn_scratch(ugn_ctx* ctx, size_t bytes, void** out);

void ugn_set_error(const char* fmt, ...);

#define UGN_FAIL(code, ...)        \
  do {                             \
    ugn_set_error(__VA_ARGS__);    \
    return (code);                 \
  } while (0)

#define UGN_CHECK(cond, ...) \
  do {                       \
    if (!(cond)) UGN_FAIL(UGN_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define UGN_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess)                                                         \
      UGN_FAIL(UGN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
               __FILE__, __LINE__);                                                 \
  } while (0)

// Checks the launch that has just been issued and counts it.
#define UGN_LAUNCHED(ctx)                                                          \
  do {                                                                             \
    cudaError_t e__ = cudaPeekAtLastError();                                       \
    if (e__ != cudaSuccess) {                                                      \
      cudaGetLastError();                                                          \
      UGN_FAIL(UGN_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                   \
               cudaGetErrorString(e__), __FILE__, __LINE__);                       \
    }                                                                              \
    (ctx)->launches++;                                                             \
  } while (0)

enum UgnDType { DT_F32, DT_BF16, DT_I32, DT_I64, DT_U8, DT_F64, DT_F16, DT_I8, DT_I16, DT_BAD };

static inline UgnDType ugn_dtype(const ugn_tensor* t) {
  if (t->dtype_lanes != 1) return DT_BAD;
  if (t->dtype_code == UGN_DL_FLOAT && t->dtype_bits == 32) return DT_F32;
  if (t->dtype_code == UGN_DL_FLOAT && t->dtype_bits == 64) return DT_F64;
  if (t->dtype_code == UGN_DL_BFLOAT && t->dtype_bits == 16) return DT_BF16;
  if (t->dtype_code == UGN_DL_FLOAT && t->dtype_bits == 16) return DT_F16;
  if (t->dtype_code == UGN_DL_INT && t->dtype_bits == 32) return DT_I32;
  if (t->dtype_code == UGN_DL_INT && t->dtype_bits == 64) return DT_I64;
  if (t->dtype_code == UGN_DL_UINT && t->dtype_bits == 8) return DT_U8;
  if (t->dtype_code == UGN_DL_INT && t->dtype_bits == 8) return DT_I8;
  if (t->dtype_code == UGN_DL_INT && t->dtype_bits == 16) return DT_I16;
  return DT_BAD;
}

static inline int64_t ugn_numel(const ugn_tensor* t) {
  int64_t n = 1;
  for (int i = 0; i < t->ndim; ++i) n *= t->shape[i];
  return n;
}

template <class T>
static inline T* ugn_ptr(const ugn_tensor* t) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(t->data) + t->byte_offset);
}

// Validates device / dtype / rank / contiguity.  name is used in the error text.
int ugn_validate(const ugn_ctx* ctx, const ugn_tensor* t, const char* name, UgnDType dt,
                 int ndim_lo, int ndim_hi);

#define UGN_TENSOR(t, dt, lo, hi)                                  \
  do {                                                             \
    int rc__ = ugn_validate(ctx, (t), #t, (dt), (lo), (hi));       \
    if (rc__ != UGN_OK) return rc__;                               \
  } while (0)

static inline int ugn_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- device helpers ------------------------------------------------------------------
__device__ __forceinline__ float ugn_act_fwd(float z, int act, float alpha) {
  if (act == UGN_ACT_RELU) return fmaxf(z, 0.f);
  if (act == UGN_ACT_LEAKY) return z > 0.f ? z : alpha * z;
  return z;
}
// derivative expressed through the activation OUTPUT y (sign(y) == sign(z) for relu/leaky)
__device__ __forceinline__ float ugn_act_bwd(float y, int act, float alpha) {
  if (act == UGN_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == UGN_ACT_LEAKY) return y > 0.f ? 1.f : alpha;
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// bf16 split: hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void ugn_split(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// 16-bit storage is either bf16 (f16 == 0) or IEEE fp16 (f16 == 1); `u16` is the raw container.
// fp16 conversions saturate at +-65504 instead of producing inf (range guard for scaled gradients).
typedef __nv_bfloat16 u16;
__device__ __forceinline__ u16 ugn_cvt16(float x, int f16) {
  if (f16) {
    x = fminf(fmaxf(x, -65504.f), 65504.f);
    return __ushort_as_bfloat16(__half_as_ushort(__float2half_rn(x)));
  }
  return __float2bfloat16_rn(x);
}
__device__ __forceinline__ float ugn_f16to32(u16 v, int f16) {
  if (f16) return __half2float(__ushort_as_half(__bfloat16_as_ushort(v)));
  return __bfloat162float(v);
}
__device__ __forceinline__ void ugn_split16(float x, int f16, u16& hi, u16& lo) {
  hi = ugn_cvt16(x, f16);
  lo = ugn_cvt16(x - ugn_f16to32(hi, f16), f16);
}

// forward declarations of the per-file implementations (dispatch lives in abi.cu)
struct ConvGeom {
  int B, H, W, Cp;      // input (padded channels)
  int Co, KH, KW;       // filter
  int Ho, Wo;           // conv output
  int Hp, Wp;           // pooled output (== Ho,Wo when pool == 0)
  int Cin;              // unpadded input channels (master weight layout)
};

struct Conv3Geom {        // strided 'valid' channels-last Conv3D (use3D branches)
  int B, T, H, W, C;      // input
  int Co, KT, KH, KW;     // filter
  int ST, SH, SW;         // strides
  int To, Ho, Wo;         // output
};

struct FusePtrs {
  const float* br[4];
  const float* flag[4];
  float* dbr[4];
};
