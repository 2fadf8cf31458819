#include "tc.cuh"
#define TC_TODO(name) UGN_FAIL(UGN_ERR_UNSUPPORTED, name ": tensor-core path not built yet")
int tc_gemm(ugn_ctx*, int, int, int, int, const __nv_bfloat16*, int, const __nv_bfloat16*, int, float*, int, cudaStream_t) { TC_TODO("tc_gemm"); }
int tc_conv_fwd(ugn_ctx*, const ConvGeom&, int, const __nv_bfloat16*, const __nv_bfloat16*, const float*, __nv_bfloat16*, uint8_t*, int, float, int, cudaStream_t) { TC_TODO("tc_conv_fwd"); }
int tc_conv_dgrad(ugn_ctx*, const ConvGeom&, int, const __nv_bfloat16*, const __nv_bfloat16*, float*, cudaStream_t) { TC_TODO("tc_conv_dgrad"); }
int tc_conv_wgrad(ugn_ctx*, const ConvGeom&, int, const __nv_bfloat16*, const __nv_bfloat16*, float*, float*, cudaStream_t) { TC_TODO("tc_conv_wgrad"); }
int tc_linear_fwd(ugn_ctx*, int, int, int, int, const __nv_bfloat16*, const __nv_bfloat16*, const float*, const float*, float*, int, float, cudaStream_t) { TC_TODO("tc_linear_fwd"); }
int tc_linear_bwd(ugn_ctx*, int, int, int, int, const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, float*, float*, float*, cudaStream_t) { TC_TODO("tc_linear_bwd"); }
