// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) as a function of N,
// operand swizzle / majorness and the row alignment of the A descriptor.  One CTA per SM, operands are
// whatever is in shared memory (no loads): this isolates the tensor pipe + its shared-memory operand
// fetch.  Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ugaitnet_b200/csrc scripts/umma_rate.cu -o /tmp/umma_rate && /tmp/umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace tc;

struct Cfg { int N, rowbytes, a_mn, b_mn, shift_rows, nacc, iters; };

__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc16(128, c.N, c.a_mn, c.b_mn, 1);
    const uint32_t lay = c.rowbytes == 128 ? 2u : 4u;
    const uint32_t a_addr = smem_u32(smem) + c.shift_rows * c.rowbytes;
    const uint32_t b_addr = smem_u32(smem) + 48 * 1024;
    // K-major: SBO = 8 rows; MN-major: LBO = one 64-element block (64 K-rows x rowbytes), SBO = 8 K-rows
    const uint64_t a0 = c.a_mn ? make_smem_desc(a_addr, 64 * c.rowbytes, 8 * c.rowbytes, lay)
                               : make_smem_desc(a_addr, 0, 8 * c.rowbytes, lay);
    const uint64_t b0 = c.b_mn ? make_smem_desc(b_addr, 64 * c.rowbytes, 8 * c.rowbytes, lay)
                               : make_smem_desc(b_addr, 0, 8 * c.rowbytes, lay);
    long long t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
      const uint32_t d = tb + (it % c.nacc) * c.N;
      // 4 K-slices of one 64-wide stage
      for (int k = 0; k < 4; ++k)
        umma_f16(d, a0 + (c.a_mn ? (uint64_t)((16 * c.rowbytes * k) >> 4) : (uint64_t)(2 * k)),
                 b0 + (c.b_mn ? (uint64_t)((16 * c.rowbytes * k) >> 4) : (uint64_t)(2 * k)), idesc, 1);
    }
    umma_commit(&bar);
    while (!mbar_try_wait(&bar, 0)) {}
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("%4s %5s %4s %4s %5s %4s  %10s %9s\n", "N", "rowB", "a_mn", "b_mn", "shift", "nacc", "clk/MMA", "B/clk");
  const int Ns[] = {64, 96, 128, 160, 192, 256};
  for (int rb : {128, 64})
    for (int mn = 0; mn < 2; ++mn)
      for (int N : Ns)
        for (int sh : {0, 1, 3}) {
          if (mn && sh) continue;
          int nacc = 512 / N;
          if (nacc > 2) nacc = 2;
          Cfg c{N, rb, mn, mn, sh, nacc, 2000};
          rate_kernel<<<sms, 128, 100 * 1024>>>(c, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long cyc;
          cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
          double per = (double)cyc / (c.iters * 4.0);
          printf("%4d %5d %4d %4d %5d %4d  %10.1f %9.1f\n", N, rb, mn, mn, sh, nacc, per, (128 + N) * 32.0 / per);
        }
  return 0;
}
