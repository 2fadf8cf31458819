// Tensor-core path: ONE warp-specialised tcgen05 kernel (TMA producer warp -> smem ring ->
// single-thread tcgen05.mma issue into TMEM -> 4 epilogue warps via tcgen05.ld) specialised by
// MODE for
//   MODE_GEMM  : C[M,N] (+)= A.B^T, operands K-major or MN-major (dense layers, Gram/distance GEMMs)
//   MODE_CONV  : implicit-GEMM convolution over NHWC activations; the A tile of every filter tap is a
//                shifted TMA box of the activation tensor (no im2col buffer).  sgn=+1 forward
//                (bias+act[+2x2 max-pool+argmax] epilogue), sgn=-1 input gradient.
//   MODE_WGRAD : kernel gradient; K = output pixels, both operands MN-major straight from NHWC,
//                split-K over pixel boxes with fp32 red.add.
// Operands are bf16 with P planes (P=2: hi/lo split -> hi*hi + hi*lo + lo*hi, ~fp32 products).
// Replaces the Conv2D / Dense layers of UWYHNet.buildBranch (nets/mj_uwyhNets_ba.py:82-105).
#include "tc.cuh"
#include "tc_ptx.cuh"
#include <algorithm>

using namespace tc;

enum { MODE_GEMM = 0, MODE_CONV = 1, MODE_WGRAD = 2 };
enum { EPI_F32 = 0, EPI_F32_ATOMIC = 1, EPI_BF16_ACT = 2, EPI_BF16_POOL = 3 };

struct alignas(64) TcOp {
  CUtensorMap map;
  int major;        // 0 K-major, 1 MN-major
  int rowbytes;     // swizzle span == inner box bytes: 128 | 64
  int nbox;         // TMA boxes per plane per stage
  int box_bytes;
  int plane_bytes;  // 1024-aligned
  int lbo, sbo, kadv;
};

struct TcParams {
  TcOp a, b;
  int mode, epi;
  int M, N, block_n;
  int ksteps_total, ksplit, kslices, planes, stages;
  // MMA passes per product over split (hi/lo) operands and the planes each operand actually loads:
  //   1: hi*hi                     (pa = 1, pb = 1)     3: hi*hi + hi*lo + lo*hi (pa = 2, pb = 2)
  //   2: hi*hi + hi*lo(B)          (pa = 1, pb = 2)     4: hi*hi + lo(A)*hi      (pa = 2, pb = 1)
  // `planes` stays the plane count of the TENSORS (TMA map extent, epilogue output planes)
  int npass, pa, pb;
  // conv
  int bw, bh, bn, ntx, nty, KW, ncc, sgn, Wout, Hout, Bn, Cout;
  int Hp, Wp;
  // wgrad
  int nbx, nby, nch, cw, ntaps, Cin, Co;
  // epilogue
  float* out_f32;
  __nv_bfloat16* out_bf16;
  long long out_plane;  // elements between bf16 planes
  uint8_t* pool_idx;
  const float* bias;
  const float* mask;
  int ldc, act;
  float alpha;
  int* err;
  int f16;                 // 16-bit operand format: 0 bf16, 1 fp16
  const float* oscale;     // device scalar multiplied into f32 outputs (1/grad-scale for backward ops), nullable
  // fused GEMM epilogue extras (narrow-tile dense layers): 16-bit planes of the result for the next tensor-core
  // consumer (out16: [planes][M][ldc]; the value converted is acc * o16scale, o16scale nullable = oscale-independent
  // scaling of gradient operands), and per-column sums of the f32 result (bias gradient) by warp reduction + atomics
  u16* out16;
  long long out16_plane;
  int out16_planes;
  int out16_raw;           // 1: convert the RAW accumulator (x mask), not the oscale'd value (gradient operands stay scaled)
  float* colsum;
  // patch-resident conv (tc_convp_kernel)
  int T, SW, RH, PR, KH, xorg, yorg, tiles_y, tmem_cols;
  int patch_chunk_bytes, patch_plane_bytes;
  int cps, nslots, ngroups;   // channel chunks per patch slot, patch slots (1 | 2), groups per tile = ncc / cps
  // accumulator layout of tc_convp_kernel: nbuf TMEM buffers (2 = epilogue overlaps the next tile);
  // concat = 1: hi*[hi|lo] is issued as ONE N = 2*block_n MMA over the two adjacent weight planes (an
  // N <= 128 MMA costs 73 clk whatever N is, N = 192 costs 96: profiles/r01_umma_rate.txt), the hi*lo
  // partial sums live in columns [block_n, 2*block_n) of the tile and are added in the epilogue
  int nbuf, concat, acc_tile_cols;
  // thread-block cluster of `cluster` (1 | 2) CTAs that work on tiles with the SAME weights: every CTA loads 1/cluster of
  // each weight tile and TMA-multicasts it to all of them (b_half_rows / b_half_bytes: its share of a K-major tile)
  int cluster, b_half_rows, b_half_bytes;
  long long* dbg;   // optional [8] cycle counters (UGN_CONVP_PROF): where each role of tc_convp_kernel waits
};

// 16-byte vector reduction (sm_90+): one red.global.add.v4.f32 instead of four scalar atomics
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 16 accumulator columns -> global red.add; vector form when the destination run is 16-byte aligned
__device__ __forceinline__ void red_add_16(float* dst, const float* v, float os, int nvalid) {
  if (nvalid >= 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) red_add_v4(dst + i, v[i] * os, v[i + 1] * os, v[i + 2] * os, v[i + 3] * os);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) atomicAdd(dst + i, v[i] * os);
  }
}

static constexpr int kThreads = 192;
static constexpr int kTmemCols = 256;

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Epilogue of one 128-row conv accumulator tile whose rows are the (bn x bh x bw) pixel box at
// (nn0, y0, x0): bias + activation, then bf16 hi/lo store, f32 store (dgrad) or fused 2x2 max-pool.
__device__ __forceinline__ void conv_tile_epilogue(const TcParams& p, uint8_t* smem, uint32_t trow, int r, bool ok,
                                                   int x0, int y0, int nn0, int n0, int half = 0, int nhalf = 1) {
  float v[16];
  const float os = p.oscale ? *p.oscale : 1.f;
  const int xl = r % p.bw, yl = (r / p.bw) % p.bh, nl = r / (p.bw * p.bh);
    const int x = x0 + xl, y = y0 + yl, n = nn0 + nl;
    const bool rv = ok && nl < p.bn && x < p.Wout && y < p.Hout && n < p.Bn;
    if (p.epi != EPI_BF16_POOL) {
      const long long obase = (((long long)n * p.Hout + y) * p.Wout + x) * p.Cout;
      for (int c0 = 16 * half; c0 < p.block_n; c0 += 16 * nhalf) {
        tmem_ld16(trow + c0, v);
        if (p.concat) {
          float v2[16];
          tmem_ld16(trow + p.block_n + c0, v2);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += v2[i];
        }
        if (!rv) continue;
        if (p.epi == EPI_F32_ATOMIC) {
          red_add_16(p.out_f32 + obase + n0 + c0, v, os, p.Cout - (n0 + c0));
          continue;
        }
        if (p.epi == EPI_F32 && (p.Cout & 3) == 0 && n0 + c0 + 16 <= p.Cout) {
          float* dst = p.out_f32 + obase + n0 + c0;
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(dst + i) = make_float4(v[i] * os, v[i + 1] * os, v[i + 2] * os, v[i + 3] * os);
          continue;
        }
        if (p.epi == EPI_BF16_ACT && (p.Cout & 7) == 0 && n0 + c0 + 16 <= p.Cout) {
          // 16 channels of one pixel = 32 contiguous bytes per plane: two 16-byte stores instead of 16
          // scattered 2-byte stores (the per-frame GaitSet layers are dominated by this epilogue)
          const int cb = n0 + c0;
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float z0 = ugn_act_fwd(v[i] + (p.bias ? p.bias[cb + i] : 0.f), p.act, p.alpha);
            const float z1 = ugn_act_fwd(v[i + 1] + (p.bias ? p.bias[cb + i + 1] : 0.f), p.act, p.alpha);
            u16 h0, l0, h1, l1;
            ugn_split16(z0, p.f16, h0, l0);
            ugn_split16(z1, p.f16, h1, l1);
            hw[i >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lw[i >> 1] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
          uint4* dh = reinterpret_cast<uint4*>(p.out_bf16 + obase + cb);
          dh[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          dh[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
          if (p.planes == 2) {
            uint4* dl = reinterpret_cast<uint4*>(p.out_bf16 + p.out_plane + obase + cb);
            dl[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            dl[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
          }
          continue;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = n0 + c0 + i;
          if (c >= p.Cout) continue;
          if (p.epi == EPI_F32) {
            p.out_f32[obase + c] = v[i] * os;
          } else {
            float z = ugn_act_fwd(v[i] + (p.bias ? p.bias[c] : 0.f), p.act, p.alpha);
            u16 hi, lo;
            ugn_split16(z, p.f16, hi, lo);
            p.out_bf16[obase + c] = hi;
            if (p.planes == 2) p.out_bf16[p.out_plane + obase + c] = lo;
          }
        }
      }
    } else {
      // fused 2x2 max-pool: stage act(z+b) through the (now idle) stage-0 smem, 32 columns at a time
      float* stg = reinterpret_cast<float*>(smem);  // [128][33]
      const int t = threadIdx.x - 64;
      const int pw = p.bw >> 1, ph2 = p.bh >> 1, npool = pw * ph2 * p.bn;
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tmem_ld16(trow + c0 + 16 * h, v);
          if (p.concat) {
            float v2[16];
            tmem_ld16(trow + p.block_n + c0 + 16 * h, v2);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += v2[i];
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = n0 + c0 + 16 * h + i;
            float z = v[i] + ((p.bias && c < p.Cout) ? p.bias[c] : 0.f);
            stg[r * 33 + 16 * h + i] = ugn_act_fwd(z, p.act, p.alpha);
          }
        }
        epi_barrier();
        const int col = t & 31, c = n0 + c0 + col;
        for (int pp = t >> 5; pp < npool; pp += 4) {
          const int pxl = pp % pw, pyl = (pp / pw) % ph2, pnl = pp / (pw * ph2);
          const int r00 = (pnl * p.bh + 2 * pyl) * p.bw + 2 * pxl;
          float best = stg[r00 * 33 + col];
          int pos = 0;
          float o1 = stg[(r00 + 1) * 33 + col], o2 = stg[(r00 + p.bw) * 33 + col], o3 = stg[(r00 + p.bw + 1) * 33 + col];
          if (o1 > best) { best = o1; pos = 1; }
          if (o2 > best) { best = o2; pos = 2; }
          if (o3 > best) { best = o3; pos = 3; }
          const int xp = (x0 >> 1) + pxl, yp = (y0 >> 1) + pyl, nn = nn0 + pnl;
          if (ok && c < p.Cout && xp < p.Wp && yp < p.Hp && nn < p.Bn) {
            const long long o = (((long long)nn * p.Hp + yp) * p.Wp + xp) * p.Cout + c;
            u16 hi, lo;
            ugn_split16(best, p.f16, hi, lo);
            p.out_bf16[o] = hi;
            if (p.planes == 2) p.out_bf16[p.out_plane + o] = lo;
            p.pool_idx[o] = (uint8_t)pos;
          }
        }
        epi_barrier();
      }
    }
}

// Fused 2x2 max-pool epilogue of the persistent conv kernel, 8 warps (two per TMEM lane quadrant = two per
// SM sub-partition, so one hides the other's latencies).  Per pass of 32 channels: thread (row r, half h)
// pulls its 16 accumulator columns with ONE tcgen05.ld, adds the bias (4 broadcast float4 loads), applies
// the activation and parks the values in smem [128][33]; after a 256-thread barrier thread (channel,
// pooled pixel) takes the max of its 2x2 window (ties -> first in (dy,dx) order) and stores the 16-bit
// planes + arg-max byte, 32 consecutive channels per warp store.  The pooled-pixel geometry of a thread is
// the same in every pass and is computed once per tile.
static constexpr int kConvpThreads = 320;
__device__ __forceinline__ void epi_barrier256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void convp_pool_epilogue(const TcParams& p, float* stg, uint32_t trow, int r, int half,
                                                    bool ok, int y0, int img, int n0) {
  const int t = threadIdx.x - 64;                 // 0..255
  const int pw = p.bw >> 1, ph2 = p.bh >> 1, npool = pw * ph2 * p.bn;
  const int col = t & 31;
  // up to 4 pooled pixels per thread and pass (npool <= 32): smem row of the window's top-left pixel and
  // the output offset (without the channel)
  int r00[4];
  long long ob[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pp = (t >> 5) + 8 * j;
    r00[j] = -1;
    ob[j] = 0;
    if (pp < npool) {
      const int pxl = pp % pw, pyl = (pp / pw) % ph2, pnl = pp / (pw * ph2);
      const int xp = pxl, yp = (y0 >> 1) + pyl, nn = img + pnl;
      if (ok && xp < p.Wp && yp < p.Hp && nn < p.Bn) {
        r00[j] = (pnl * p.bh + 2 * pyl) * p.bw + 2 * pxl;
        ob[j] = (((long long)nn * p.Hp + yp) * p.Wp + xp) * p.Cout;
      }
    }
  }
  float v[16];
  for (int c0 = 0; c0 < p.block_n; c0 += 32) {
    tmem_ld16(trow + c0 + 16 * half, v);
    if (p.concat) {
      float v2[16];
      tmem_ld16(trow + p.block_n + c0 + 16 * half, v2);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += v2[i];
    }
    const int cb = n0 + c0 + 16 * half;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias && cb + i + 4 <= p.Cout) b4 = *reinterpret_cast<const float4*>(p.bias + cb + i);
      float* d = stg + r * 33 + 16 * half + i;
      d[0] = ugn_act_fwd(v[i] + b4.x, p.act, p.alpha);
      d[1] = ugn_act_fwd(v[i + 1] + b4.y, p.act, p.alpha);
      d[2] = ugn_act_fwd(v[i + 2] + b4.z, p.act, p.alpha);
      d[3] = ugn_act_fwd(v[i + 3] + b4.w, p.act, p.alpha);
    }
    epi_barrier256();
    const int c = n0 + c0 + col;
    if (c < p.Cout) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (r00[j] < 0) continue;
        const float* w0 = stg + r00[j] * 33 + col;
        float best = w0[0];
        int pos = 0;
        const float o1 = w0[33], o2 = w0[p.bw * 33], o3 = w0[(p.bw + 1) * 33];
        if (o1 > best) { best = o1; pos = 1; }
        if (o2 > best) { best = o2; pos = 2; }
        if (o3 > best) { best = o3; pos = 3; }
        u16 hi, lo;
        ugn_split16(best, p.f16, hi, lo);
        const long long o = ob[j] + c;
        p.out_bf16[o] = hi;
        if (p.planes == 2) p.out_bf16[p.out_plane + o] = lo;
        p.pool_idx[o] = (uint8_t)pos;
      }
    }
    epi_barrier256();
  }
}

// PL2 (split operands) and KSL (K slices per ring stage; 0 = run-time) are compile-time so that the MMA issue
// sequence of a stage is straight-line code (see tc_convp_kernel).
template <int MODE, int NPASS, int KSL>
__global__ void __launch_bounds__(kThreads, 1) tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A planes | B planes)] | barriers | tmem ptr
  const uint32_t a_stage = p.pa * p.a.plane_bytes, b_stage = p.pb * p.b.plane_bytes;
  const uint32_t stage_bytes = a_stage + b_stage;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;

  // K range of this CTA
  int per = (p.ksteps_total + p.ksplit - 1) / p.ksplit;
  int ks_beg = blockIdx.z * per, ks_end = min(p.ksteps_total, ks_beg + per);
  int nsteps = ks_end - ks_beg;
  if (nsteps <= 0) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
    prefetch_tmap(&p.a.map);
    prefetch_tmap(&p.b.map);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  // tile origin
  int m0 = tile_m * 128, n0 = tile_n * p.block_n;
  int x0 = 0, y0 = 0, nn0 = 0;
  if (MODE == MODE_CONV) {
    int tx = tile_m % p.ntx, ty = (tile_m / p.ntx) % p.nty, tn = tile_m / (p.ntx * p.nty);
    x0 = tx * p.bw; y0 = ty * p.bh; nn0 = tn * p.bn;
  }
  const int BK = p.kslices * 16;
  const int cwA = p.a.rowbytes >> 1, cwB = p.b.rowbytes >> 1;

  if (warp == 0) {
   if (elect_one()) {
    // ===================== TMA producer (one elected lane; `if (elect.sync)` lets ptxas treat the
    // region as single-threaded: descriptors / coordinates live in uniform registers) =====================
    int nb_b = p.b.nbox;
    if (MODE == MODE_WGRAD) {
      int remain = p.nch * p.ntaps - tile_n * p.b.nbox;
      nb_b = min(nb_b, remain);
    }
    const uint32_t tx_bytes = p.pa * p.a.nbox * p.a.box_bytes + p.pb * nb_b * p.b.box_bytes;
    // incremental K-step decode (no div/mod in the loop)
    int d0 = 0, d1 = 0, d2 = 0;   // CONV: cc, kw, kh ; WGRAD: bx, by, bb
    if (MODE == MODE_CONV) { d0 = ks_beg % p.ncc; int tap = ks_beg / p.ncc; d1 = tap % p.KW; d2 = tap / p.KW; }
    if (MODE == MODE_WGRAD) { d0 = ks_beg % p.nbx; d1 = (ks_beg / p.nbx) % p.nby; d2 = ks_beg / (p.nbx * p.nby); }
    int s = 0, ph = 0;
    for (int it = 0; it < nsteps; ++it) {
      if (!mbar_wait(&empty[s], ph ^ 1, p.err, 1)) break;
      mbar_expect_tx(&full[s], tx_bytes);
      const int ks = ks_beg + it;
      uint8_t* sa = smem + (size_t)s * stage_bytes;
      uint8_t* sb = sa + a_stage;
      const int plmax = p.pa > p.pb ? p.pa : p.pb;
      for (int pl = 0; pl < plmax; ++pl) {
        const bool la_ = pl < p.pa, lb_ = pl < p.pb;       // which operands load this plane
        if (MODE == MODE_GEMM) {
          for (int j = 0; la_ && j < p.a.nbox; ++j) {
            if (p.a.major == 0) tma_load_5d(&p.a.map, &full[s], sa + pl * p.a.plane_bytes, ks * BK, m0, pl, 0, 0);
            else tma_load_5d(&p.a.map, &full[s], sa + pl * p.a.plane_bytes + j * p.a.box_bytes, m0 + j * cwA, ks * BK, pl, 0, 0);
          }
          for (int j = 0; lb_ && j < p.b.nbox; ++j) {
            if (p.b.major == 0) tma_load_5d(&p.b.map, &full[s], sb + pl * p.b.plane_bytes, ks * BK, n0, pl, 0, 0);
            else tma_load_5d(&p.b.map, &full[s], sb + pl * p.b.plane_bytes + j * p.b.box_bytes, n0 + j * cwB, ks * BK, pl, 0, 0);
          }
        } else if (MODE == MODE_CONV) {
          const int cc = d0, kw = d1, kh = d2, tap = kh * p.KW + kw;
          if (la_) tma_load_5d(&p.a.map, &full[s], sa + pl * p.a.plane_bytes, cc * BK, x0 + p.sgn * kw, y0 + p.sgn * kh, nn0, pl);
          for (int j = 0; lb_ && j < p.b.nbox; ++j) {
            if (p.b.major == 0) tma_load_5d(&p.b.map, &full[s], sb + pl * p.b.plane_bytes, cc * BK, tap, n0, pl, 0);
            else tma_load_5d(&p.b.map, &full[s], sb + pl * p.b.plane_bytes + j * p.b.box_bytes, n0 + j * cwB, tap, cc * BK, pl, 0);
          }
        } else {
          const int px = d0 * p.bw, py = d1 * p.bh, pn = d2 * p.bn;
          for (int j = 0; la_ && j < p.a.nbox; ++j)
            tma_load_5d(&p.a.map, &full[s], sa + pl * p.a.plane_bytes + j * p.a.box_bytes, m0 + j * cwA, px, py, pn, pl);
          int chunk = (tile_n * p.b.nbox) % p.nch, tap = (tile_n * p.b.nbox) / p.nch;
          int tkw = tap % p.KW, tkh = tap / p.KW;
          for (int j = 0; lb_ && j < nb_b; ++j) {
            tma_load_5d(&p.b.map, &full[s], sb + pl * p.b.plane_bytes + j * p.b.box_bytes, chunk * p.cw,
                              px + tkw, py + tkh, pn, pl);
            if (++chunk == p.nch) { chunk = 0; if (++tkw == p.KW) { tkw = 0; ++tkh; } }
          }
        }
      }
      if (MODE == MODE_CONV) { if (++d0 == p.ncc) { d0 = 0; if (++d1 == p.KW) { d1 = 0; ++d2; } } }
      if (MODE == MODE_WGRAD) { if (++d0 == p.nbx) { d0 = 0; if (++d1 == p.nby) { d1 = 0; ++d2; } } }
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
   }
  } else if (warp == 1) {
   if (elect_one()) {
    // ===================== MMA issuer (one elected lane) =====================
    const uint32_t idesc = make_idesc16(128, p.block_n, p.a.major, p.b.major, p.f16);
    const uint32_t la = p.a.rowbytes == 128 ? 2u : 4u, lb = p.b.rowbytes == 128 ? 2u : 4u;
    const uint32_t smem0 = smem_u32(smem);
    // descriptors of stage 0 / plane 0 / slice 0; everything else is a 16-byte-unit add on the low word
    const uint64_t a0 = make_smem_desc(smem0, p.a.lbo, p.a.sbo, la);
    const uint64_t b0 = make_smem_desc(smem0 + a_stage, p.b.lbo, p.b.sbo, lb);
    const uint32_t st16 = stage_bytes >> 4, ka16 = p.a.kadv >> 4, kb16 = p.b.kadv >> 4;
    const uint32_t pa16 = p.a.plane_bytes >> 4, pb16 = p.b.plane_bytes >> 4;
    uint32_t accum = 0;
    int s = 0, ph = 0;
    for (int it = 0; it < nsteps; ++it) {
      if (!mbar_wait(&full[s], ph, p.err, 2)) break;
      fence_after_sync();
      uint64_t a_hi = a0 + (uint64_t)(s * st16), b_hi = b0 + (uint64_t)(s * st16);
      const int nk = KSL > 0 ? KSL : p.kslices;
#pragma unroll
      for (int k = 0; k < nk; ++k) {
        const uint64_t ak = a_hi + (uint64_t)(k * ka16), bk = b_hi + (uint64_t)(k * kb16);
        umma_f16(tmem_base, ak, bk, idesc, accum);
        accum = 1;
        if (NPASS == 2 || NPASS == 3) umma_f16(tmem_base, ak, bk + pb16, idesc, 1);
        if (NPASS == 3 || NPASS == 4) umma_f16(tmem_base, ak + pa16, bk, idesc, 1);
      }
      umma_commit(&empty[s]);  // frees the smem slot once these MMAs have read it
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    umma_commit(tmem_full);    // accumulator complete
   }
  } else if (warp >= 2) {
    // ===================== epilogue (4 warps, one TMEM lane quadrant each) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // accumulator row owned by this thread
    bool ok = mbar_wait(tmem_full, 0, p.err, 3);
    fence_after_sync();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float v[16];

    const float os = p.oscale ? *p.oscale : 1.f;
    if (MODE == MODE_GEMM) {
      const int m = m0 + r;
      const bool vec = (p.ldc & 3) == 0 && p.epi == EPI_F32 && !p.mask;
      const bool fused = p.out16 != nullptr || p.colsum != nullptr;
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        tmem_ld16(trow + c0, v);
        if (fused) {
          // narrow-tile dense layer: bias / activation / mask, f32 store (nullable), 16-bit planes, column sums.
          // No early exit: the column sums are reduced across the warp.
          const bool rv = ok && m < p.M;
          float z[16];
          __align__(16) u16 hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = n0 + c0 + i;
            const long long o = (long long)m * p.ldc + n;
            const bool cv = rv && n < p.N;
            const float mk = (cv && p.mask) ? p.mask[o] : 1.f;
            float zz = ugn_act_fwd(v[i] * os + ((p.bias && cv) ? p.bias[n] : 0.f), p.act, p.alpha) * mk;
            z[i] = cv ? zz : 0.f;
            if (p.out16) ugn_split16(p.out16_raw ? (cv ? v[i] * mk : 0.f) : z[i], p.f16, hi[i], lo[i]);
          }
          if (rv && n0 + c0 + 16 <= p.N && (p.ldc & 7) == 0) {
            const long long o = (long long)m * p.ldc + n0 + c0;
            if (p.out_f32) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(p.out_f32 + o + i) = make_float4(z[i], z[i + 1], z[i + 2], z[i + 3]);
            }
            if (p.out16) {
              *reinterpret_cast<uint4*>(p.out16 + o) = *reinterpret_cast<const uint4*>(hi);
              *reinterpret_cast<uint4*>(p.out16 + o + 8) = *reinterpret_cast<const uint4*>(hi + 8);
              if (p.out16_planes == 2) {
                *reinterpret_cast<uint4*>(p.out16 + p.out16_plane + o) = *reinterpret_cast<const uint4*>(lo);
                *reinterpret_cast<uint4*>(p.out16 + p.out16_plane + o + 8) = *reinterpret_cast<const uint4*>(lo + 8);
              }
            }
          } else if (rv) {
            for (int i = 0; i < 16; ++i) {
              const int n = n0 + c0 + i;
              if (n >= p.N) break;
              const long long o = (long long)m * p.ldc + n;
              if (p.out_f32) p.out_f32[o] = z[i];
              if (p.out16) { p.out16[o] = hi[i]; if (p.out16_planes == 2) p.out16[p.out16_plane + o] = lo[i]; }
            }
          }
          if (p.colsum) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float sacc = z[i];
#pragma unroll
              for (int d = 16; d > 0; d >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, d);
              if (lane == 0 && n0 + c0 + i < p.N) atomicAdd(p.colsum + n0 + c0 + i, sacc);
            }
          }
          continue;
        }
        if (!ok || m >= p.M) continue;
        if (p.epi == EPI_F32_ATOMIC) {
          red_add_16(p.out_f32 + (long long)m * p.ldc + n0 + c0, v, os, p.N - (n0 + c0));
          continue;
        }
        if (vec && n0 + c0 + 16 <= p.N) {
          float* dst = p.out_f32 + (long long)m * p.ldc + n0 + c0;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 o4;
            o4.x = ugn_act_fwd(v[i] * os + (p.bias ? p.bias[n0 + c0 + i] : 0.f), p.act, p.alpha);
            o4.y = ugn_act_fwd(v[i + 1] * os + (p.bias ? p.bias[n0 + c0 + i + 1] : 0.f), p.act, p.alpha);
            o4.z = ugn_act_fwd(v[i + 2] * os + (p.bias ? p.bias[n0 + c0 + i + 2] : 0.f), p.act, p.alpha);
            o4.w = ugn_act_fwd(v[i + 3] * os + (p.bias ? p.bias[n0 + c0 + i + 3] : 0.f), p.act, p.alpha);
            *reinterpret_cast<float4*>(dst + i) = o4;
          }
          continue;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = n0 + c0 + i;
          if (n >= p.N) continue;
          const long long o = (long long)m * p.ldc + n;
          if (p.epi == EPI_F32_ATOMIC) {
            atomicAdd(p.out_f32 + o, v[i] * os);
          } else {
            float z = v[i] * os + (p.bias ? p.bias[n] : 0.f);
            z = ugn_act_fwd(z, p.act, p.alpha);
            if (p.mask) z *= p.mask[o];
            p.out_f32[o] = z;
          }
        }
      }
    } else if (MODE == MODE_CONV) {
      conv_tile_epilogue(p, smem, trow, r, ok, x0, y0, nn0, n0);
    } else {  // MODE_WGRAD: rows = co, columns = (box j -> tap, ci chunk)
      const int co = m0 + r;
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        tmem_ld16(trow + c0, v);
        if (!ok || co >= p.Co) continue;
        const int j = c0 / p.cw, gb = tile_n * p.b.nbox + j;
        if (gb >= p.nch * p.ntaps) continue;
        const int chunk = gb % p.nch, tap = gb / p.nch;
        const int cib = chunk * p.cw + (c0 % p.cw);
        float* dst = p.out_f32 + ((long long)co * p.ntaps + tap) * p.Cin;
        red_add_16(dst + cib, v, os, p.Cin - cib);
      }
    }
    fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------
// Patch-resident implicit-GEMM convolution.
// One CTA owns T M-tiles = T*RH consecutive output rows of one image.  The activation patch
// (PR = T*RH+KH-1 image rows x SW pixel slots, every channel chunk, every plane) is loaded ONCE by TMA
// into 128B/64B-swizzled smem; the A operand of filter tap (kh,kw) for tile t is the SAME patch viewed
// through a UMMA descriptor whose start address is advanced by ((t*RH+kh)*SW+kw) rows (the swizzle is a
// function of the absolute smem address, so row-shifted views are exact -- scripts/shift_check.py).
// Only the weight tile of each (tap, channel chunk) streams through the mbarrier ring, and it is shared by
// the T accumulators in TMEM.  sgn=-1 (input gradient) flips the taps and shifts the patch origin.
// ---------------------------------------------------------------------------------------
// CONCAT / PROF are compile-time: a run-time branch or the cycle-counter bookkeeping inside the single-thread
// issue loops costs more than the MMAs they surround when there is only one MMA per K slice (dgrad).
// TT (tiles per work item), KSL (K slices per ring stage) and PL2 (split operands) are compile-time too, so
// that the MMA issue sequence of one ring stage is a straight line of UTCHMMAs with immediate descriptor
// offsets.
template <bool CONCAT, bool PROF, int TT, int KSL, int NPASS, int CL>
__global__ void __launch_bounds__(kConvpThreads, 1) tc_convp_kernel(const __grid_constant__ TcParams p) {
  // PERSISTENT: each CTA walks tiles blockIdx.x, +gridDim.x, ...; two TMEM accumulator sets so that the
  // epilogue of tile i (CUDA cores) overlaps the mainloop of tile i+1 (tensor pipe).
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // patch slot = cps channel chunks x planes; nslots == 2 streams the chunks through a 2-deep ring
  const uint32_t slot_bytes = p.pa * p.patch_plane_bytes;
  const uint32_t b_stage = p.pb * p.b.plane_bytes;
  uint8_t* stg = smem + p.nslots * slot_bytes;                 // epilogue staging, 128 x 33 floats (17 KB region)
  uint8_t* ring = stg + 17 * 1024;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)p.stages * b_stage);
  uint64_t* empty = full + p.stages;
  uint64_t* patch_full = empty + p.stages;     // [2]
  uint64_t* patch_empty = patch_full + 2;      // [2]
  uint64_t* tmem_full = patch_empty + 2;       // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsteps = p.ksteps_total;
  const int ntiles = p.M;          // total tiles = n_tiles_n * images * tiles_y
  const int tiles_mn = p.N;        // images * tiles_y

  if (threadIdx.x == 0) {
    // a ring stage is refilled by EVERY CTA of the cluster (multicast halves), so it is free only when every CTA's MMA
    // issuer has released it: CL arrivals
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&patch_full[b], 1); mbar_init(&patch_empty[b], 1);
      mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 256);
    }
    fence_mbar_init();
    prefetch_tmap(&p.a.map);
    prefetch_tmap(&p.b.map);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync();          // every CTA's barriers exist before a peer multicasts into / arrives on them
  fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const int BK = p.kslices * 16;
  const int cwB = p.b.rowbytes >> 1;
  const uint32_t acc_cols = (uint32_t)(p.T * p.acc_tile_cols);
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

  if (warp == 0) {
   if (elect_one()) {
    // ---- producer ----
    const uint32_t tx_bytes = p.pb * p.b.nbox * p.b.box_bytes;
    int s = 0, ph = 0, li = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++li) {
      const int tile_n = tile / tiles_mn, rem = tile % tiles_mn;
      const int img = rem / p.tiles_y, ty = rem % p.tiles_y;
      const int y0 = ty * p.T * p.RH, n0 = tile_n * p.block_n;
      for (int g = 0; g < p.ngroups; ++g) {
        const int u = li * p.ngroups + g, slot = u % p.nslots, sph = (u / p.nslots) & 1;
        uint8_t* pslot = smem + slot * slot_bytes;
        long long tw0 = PROF ? clock64() : 0;
        mbar_wait(&patch_empty[slot], sph ^ 1, p.err, 5);
        if (PROF) atomicAdd((unsigned long long*)p.dbg + 7, (unsigned long long)(clock64() - tw0));
        mbar_expect_tx(&patch_full[slot], slot_bytes);
        for (int pl = 0; pl < p.pa; ++pl)
          for (int c = 0; c < p.cps; ++c)
            tma_load_5d(&p.a.map, &patch_full[slot], pslot + pl * p.patch_plane_bytes + c * p.patch_chunk_bytes,
                        (g * p.cps + c) * (p.a.rowbytes >> 1), p.xorg, y0 + p.yorg, img, pl);
        int c = 0, tap = 0;
        for (int it = 0; it < nsteps / p.ngroups; ++it) {
          const int cc = g * p.cps + c;
          long long tw1 = PROF ? clock64() : 0;
          mbar_wait(&empty[s], ph ^ 1, p.err, 1);
          if (PROF) atomicAdd((unsigned long long*)p.dbg + 6, (unsigned long long)(clock64() - tw1));
          mbar_expect_tx(&full[s], tx_bytes);
          uint8_t* sb = ring + (size_t)s * b_stage;
          if (CL == 1) {
            for (int pl = 0; pl < p.pb; ++pl)
              for (int j = 0; j < p.b.nbox; ++j) {
                if (p.b.major == 0) tma_load_5d(&p.b.map, &full[s], sb + pl * p.b.plane_bytes, cc * BK, tap, n0, pl, 0);
                else tma_load_5d(&p.b.map, &full[s], sb + pl * p.b.plane_bytes + j * p.b.box_bytes, n0 + j * cwB, tap, cc * BK, pl, 0);
              }
          } else {
            // this CTA's share of the stage, delivered to every CTA of the cluster (the peers deliver the rest)
            for (int pl = 0; pl < p.pb; ++pl) {
              if (p.b.major == 0) {
                tma_load_5d_mc(&p.b.map, &full[s], sb + pl * p.b.plane_bytes + crank * p.b_half_bytes, cc * BK, tap,
                               n0 + (int)crank * p.b_half_rows, pl, 0, kMask);
              } else {
                for (int j = (int)crank; j < p.b.nbox; j += CL)
                  tma_load_5d_mc(&p.b.map, &full[s], sb + pl * p.b.plane_bytes + j * p.b.box_bytes, n0 + j * cwB, tap,
                                 cc * BK, pl, 0, kMask);
              }
            }
          }
          if (++c == p.cps) { c = 0; ++tap; }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
   }
  } else if (warp == 1) {
   if (elect_one()) {
    // ---- MMA issuer ----
    const uint32_t idesc = make_idesc16(128, p.block_n, 0, p.b.major, p.f16);
    const uint32_t idesc2 = make_idesc16(128, 2 * p.block_n, 0, p.b.major, p.f16);   // concat: both weight planes
    const uint32_t la = p.a.rowbytes == 128 ? 2u : 4u, lb = p.b.rowbytes == 128 ? 2u : 4u;
    const uint32_t smem0 = smem_u32(smem);
    const uint64_t a0 = make_smem_desc(smem0, 0, 8 * p.a.rowbytes, la);
    const uint64_t b0 = make_smem_desc(smem_u32(ring), p.b.lbo, p.b.sbo, lb);
    const uint32_t st16 = b_stage >> 4, kb16 = p.b.kadv >> 4, pb16 = p.b.plane_bytes >> 4;
    const uint32_t pa16 = p.patch_plane_bytes >> 4, ch16 = p.patch_chunk_bytes >> 4, row16 = p.a.rowbytes >> 4;
    const uint32_t tile16 = (uint32_t)(p.RH * p.SW) * row16;
    int s = 0, ph = 0, li = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++li) {
      const int buf = li % p.nbuf;
      long long tm0 = PROF ? clock64() : 0;
      mbar_wait(&tmem_empty[buf], ((li / p.nbuf) & 1) ^ 1, p.err, 6);
      if (PROF) atomicAdd((unsigned long long*)p.dbg + 2, (unsigned long long)(clock64() - tm0));
      const uint32_t tacc = tmem_base + buf * acc_cols;
      for (int g = 0; g < p.ngroups; ++g) {
      const int u = li * p.ngroups + g, slot = u % p.nslots, sph = (u / p.nslots) & 1;
      long long tm1 = PROF ? clock64() : 0;
      mbar_wait(&patch_full[slot], sph, p.err, 4);
      if (PROF) atomicAdd((unsigned long long*)p.dbg + 0, (unsigned long long)(clock64() - tm1));
      fence_after_sync();
      const uint64_t a_slot = a0 + (uint64_t)(slot * (slot_bytes >> 4));
      int cc = 0, kw = 0, kh = 0;
      for (int it = 0; it < nsteps / p.ngroups; ++it) {
        long long tm2 = PROF ? clock64() : 0;
        mbar_wait(&full[s], ph, p.err, 2);
        if (PROF) atomicAdd((unsigned long long*)p.dbg + 1, (unsigned long long)(clock64() - tm2));
        fence_after_sync();
        const int ay = p.sgn > 0 ? kh : (p.KH - 1 - kh), ax = p.sgn > 0 ? kw : (p.KW - 1 - kw);
        const uint64_t a_tap = a_slot + (uint64_t)(cc * ch16 + (uint32_t)(ay * p.SW + ax) * row16);
        const uint64_t b_st = b0 + (uint64_t)(s * st16);
        const uint32_t acc0 = (g > 0 || it > 0) ? 1u : 0u;
#pragma unroll
        for (int t = 0; t < TT; ++t) {
          const uint64_t a_t = a_tap + (uint64_t)(t * tile16);
          const uint32_t td = tacc + t * p.acc_tile_cols;
#pragma unroll
          for (int k = 0; k < KSL; ++k) {
            const uint64_t a_hi = a_t + (uint64_t)(2 * k);      // 32 bytes = one UMMA_K slice inside the swizzled row
            const uint64_t b_hi = b_st + (uint64_t)(k * kb16);
            if (CONCAT) {
              umma_f16(td, a_hi, b_hi, idesc2, (k > 0) ? 1u : acc0);   // [hi*hi | hi*lo], N = 2*block_n
              if (NPASS == 3) umma_f16(td, a_hi + pa16, b_hi, idesc, 1);   // lo*hi into the hi*hi columns
            } else {
              umma_f16(td, a_hi, b_hi, idesc, (k > 0) ? 1u : acc0);
              if (NPASS == 2 || NPASS == 3) umma_f16(td, a_hi, b_hi + pb16, idesc, 1);
              if (NPASS == 3 || NPASS == 4) umma_f16(td, a_hi + pa16, b_hi, idesc, 1);
            }
          }
        }
        if (CL == 1) umma_commit(&empty[s]);
        else umma_commit_mc(&empty[s], kMask);      // releases the stage in every CTA of the cluster
        if (++cc == p.cps) { cc = 0; if (++kw == p.KW) { kw = 0; ++kh; } }
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      umma_commit(&patch_empty[slot]);   // slot may be overwritten once these MMAs have read it
      }
      umma_commit(&tmem_full[buf]);    // accumulators of this tile complete
      if (PROF) atomicAdd((unsigned long long*)p.dbg + 3, (unsigned long long)(clock64() - tm0));
    }
   }
  } else {
    // ---- epilogue: T accumulator tiles per work item, each an (RH x SW) pixel box; 8 warps, warp (q, half)
    // owns TMEM lanes 32q.. and every second group of 16 columns ----
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    int li = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++li) {
      const int tile_n = tile / tiles_mn, rem = tile % tiles_mn;
      const int img = rem / p.tiles_y, ty = rem % p.tiles_y;
      const int y0 = ty * p.T * p.RH, n0 = tile_n * p.block_n;
      const int buf = li % p.nbuf;
      long long te0 = (PROF && threadIdx.x == 64) ? clock64() : 0;
      bool ok = mbar_wait(&tmem_full[buf], (li / p.nbuf) & 1, p.err, 3);
      long long te1 = (PROF && threadIdx.x == 64) ? clock64() : 0;
      fence_after_sync();
      for (int t = 0; t < p.T; ++t) {
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * acc_cols + t * p.acc_tile_cols;
        if (p.epi == EPI_BF16_POOL)
          convp_pool_epilogue(p, reinterpret_cast<float*>(stg), trow, r, half, ok, y0 + t * p.RH, img, n0);
        else
          conv_tile_epilogue(p, stg, trow, r, ok, 0, y0 + t * p.RH, img, n0, half, 2);
      }
      fence_before_sync();
      mbar_arrive(&tmem_empty[buf]);
      if (PROF && threadIdx.x == 64) {
        atomicAdd((unsigned long long*)p.dbg + 4, (unsigned long long)(te1 - te0));
        atomicAdd((unsigned long long*)p.dbg + 5, (unsigned long long)(clock64() - te1));
      }
    }
  }
  __syncthreads();
  if (CL > 1) cluster_sync();          // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    __syncwarp();
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------
// Kernel-gradient v2 ("row-resident" wgrad).
//   dw[co][kh][kw][ci] = sum_pix dz[pix][co] * x[pix + (kh,kw)][ci]
// K-step = one box of 64 output-pixel slots (PW pixel slots x bh rows x bn images, PW >= Wo+KW-1 so that
// the A rows which would pair with wrapped B rows are TMA zero-filled).  Per (kh, ci-chunk) ONE activation
// box is loaded; the KW taps of that filter row are NOT separate loads: they are the same box viewed
// through an MN-major UMMA descriptor whose N-blocks overlap -- leading-byte-offset = one pixel row -- so a
// single tcgen05.mma of N = nkw*cw columns produces the gradients of nkw taps at once.
// A (dz, MN-major) is shared by every segment of the CTA; accumulators of up to 512 TMEM columns.
// ---------------------------------------------------------------------------------------
struct WgSeg { int box, kh, chunk, kw0, nkw, col0; };
struct alignas(64) WgParams {
  CUtensorMap amap, bmap;
  int planes, stages, ksteps_total, ksplit;
  int PW, bh, bn, nby;                 // K-step box; steps decode: by = ks % nby, bb = ks / nby
  int cw, rowbytes_b, KW, ntaps, Cin, Co;
  int a_box_bytes, a_plane_bytes, b_box_bytes, b_plane_bytes;   // b_box_bytes includes the zero tail
  int ntiles_n;
  int tile_seg0[33];                   // segments of N-tile j: [tile_seg0[j], tile_seg0[j+1])
  int tile_box0[33];                   // activation boxes of N-tile j
  WgSeg seg[40];
  int box_kh[40], box_chunk[40];
  float* dw;
  int* err;
  int f16;
  const float* oscale;
};

template <bool PL2>
__global__ void __launch_bounds__(kThreads, 1) tc_wgradv_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;
  const int seg0 = p.tile_seg0[tile_n], seg1 = p.tile_seg0[tile_n + 1];
  const int box0 = p.tile_box0[tile_n], nboxes = p.tile_box0[tile_n + 1] - box0;
  const uint32_t a_stage = p.planes * p.a_plane_bytes, b_stage = p.planes * p.b_plane_bytes;
  const uint32_t stage_bytes = a_stage + b_stage;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int per = (p.ksteps_total + p.ksplit - 1) / p.ksplit;
  const int ks_beg = blockIdx.z * per, ks_end = min(p.ksteps_total, ks_beg + per);
  const int nsteps = ks_end - ks_beg;
  if (nsteps <= 0) return;

  // zero the tail rows of every activation box slot (TMA never writes them; shifted views read them)
  {
    const int tail = p.b_box_bytes - 64 * p.rowbytes_b;
    const int slots = p.stages * p.planes * 4;
    for (int e = threadIdx.x; e < slots * (tail / 16); e += blockDim.x) {
      int slot = e / (tail / 16), off = e % (tail / 16);
      int st = slot / (p.planes * 4), rem = slot % (p.planes * 4), pl = rem / 4, bx = rem % 4;
      uint8_t* base = smem + (size_t)st * stage_bytes + a_stage + pl * p.b_plane_bytes + bx * p.b_box_bytes +
                      64 * p.rowbytes_b;
      if (bx * p.b_box_bytes + p.b_box_bytes <= p.b_plane_bytes)
        reinterpret_cast<uint4*>(base)[off] = make_uint4(0, 0, 0, 0);
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
    prefetch_tmap(&p.amap);
    prefetch_tmap(&p.bmap);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  // make the generic-proxy zero fill visible to the async proxy (UMMA reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const int m0 = tile_m * 128;

  if (warp == 0) {
   if (elect_one()) {
    const uint32_t tx_bytes = p.planes * (2 * p.a_box_bytes + nboxes * 64 * p.rowbytes_b);
    int by = ks_beg % p.nby, bb = ks_beg / p.nby, s = 0, ph = 0;
    for (int it = 0; it < nsteps; ++it) {
      mbar_wait(&empty[s], ph ^ 1, p.err, 1);
      mbar_expect_tx(&full[s], tx_bytes);
      uint8_t* sa = smem + (size_t)s * stage_bytes;
      uint8_t* sb = sa + a_stage;
      const int py = by * p.bh, pn = bb * p.bn;
      for (int pl = 0; pl < p.planes; ++pl) {
        tma_load_5d(&p.amap, &full[s], sa + pl * p.a_plane_bytes, m0, 0, py, pn, pl);
        tma_load_5d(&p.amap, &full[s], sa + pl * p.a_plane_bytes + p.a_box_bytes, m0 + 64, 0, py, pn, pl);
        for (int j = 0; j < nboxes; ++j)
          tma_load_5d(&p.bmap, &full[s], sb + pl * p.b_plane_bytes + j * p.b_box_bytes,
                      p.box_chunk[box0 + j] * p.cw, 0, py + p.box_kh[box0 + j], pn, pl);
      }
      if (++by == p.nby) { by = 0; ++bb; }
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
   }
  } else if (warp == 1) {
   if (elect_one()) {
    const uint32_t lb = p.rowbytes_b == 128 ? 2u : 4u;
    const uint32_t smem0 = smem_u32(smem);
    // A: dz box, MN-major SW128: N-blocks (64 co) are a_box_bytes apart, 8 K-rows = 1024 B
    const uint64_t a0 = make_smem_desc(smem0, p.a_box_bytes, 1024, 2u);
    // B: activation box, MN-major; overlapping N-blocks: LBO = one pixel row
    const uint64_t b0 = make_smem_desc(smem0 + a_stage, p.rowbytes_b, 8 * p.rowbytes_b, lb);
    const uint32_t st16 = stage_bytes >> 4, pa16 = p.a_plane_bytes >> 4, pb16 = p.b_plane_bytes >> 4;
    const uint32_t ka16 = (16 * 128) >> 4, kb16 = (16 * p.rowbytes_b) >> 4, row16 = p.rowbytes_b >> 4;
    // per-segment constants of this N-tile, hoisted out of the K loop (<= 8 segments; the unrolled loops below
    // keep them in registers): instruction descriptor, accumulator columns, B-descriptor offset
    constexpr int kMaxSeg = 8;
    const int nseg = seg1 - seg0;
    uint32_t s_idesc[kMaxSeg], s_td[kMaxSeg], s_boff[kMaxSeg];
#pragma unroll
    for (int gi = 0; gi < kMaxSeg; ++gi) {
      const WgSeg sg = p.seg[seg0 + (gi < nseg ? gi : 0)];
      s_idesc[gi] = make_idesc16(128, sg.nkw * p.cw, 1, 1, p.f16);
      s_td[gi] = tmem_base + sg.col0;
      s_boff[gi] = (uint32_t)((sg.box - box0) * (p.b_box_bytes >> 4) + sg.kw0 * row16);
    }
    int s = 0, ph = 0;
    for (int it = 0; it < nsteps; ++it) {
      mbar_wait(&full[s], ph, p.err, 2);
      fence_after_sync();
      const uint64_t a_st = a0 + (uint64_t)(s * st16), b_st = b0 + (uint64_t)(s * st16);
      const uint32_t acc = it > 0 ? 1u : 0u;
#pragma unroll
      for (int gi = 0; gi < kMaxSeg; ++gi) {
        if (gi < nseg) {
          const uint64_t b_sg = b_st + (uint64_t)s_boff[gi];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a_hi = a_st + (uint64_t)(k * ka16), b_hi = b_sg + (uint64_t)(k * kb16);
            umma_f16(s_td[gi], a_hi, b_hi, s_idesc[gi], k > 0 ? 1u : acc);
            if (PL2) {
              umma_f16(s_td[gi], a_hi, b_hi + pb16, s_idesc[gi], 1);
              umma_f16(s_td[gi], a_hi + pa16, b_hi, s_idesc[gi], 1);
            }
          }
        }
      }
      umma_commit(&empty[s]);
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    umma_commit(tmem_full);
   }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int co = m0 + r;
    bool ok = mbar_wait(tmem_full, 0, p.err, 3);
    fence_after_sync();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const float os = p.oscale ? *p.oscale : 1.f;
    float v[16];
    for (int g = seg0; g < seg1; ++g) {
      const WgSeg sg = p.seg[g];
      const int ncols = sg.nkw * p.cw;
      for (int c0 = 0; c0 < ncols; c0 += 16) {
        tmem_ld16(trow + sg.col0 + c0, v);
        if (!ok || co >= p.Co) continue;
        const int tap = sg.kh * p.KW + sg.kw0 + c0 / p.cw;
        const int cib = sg.chunk * p.cw + (c0 % p.cw);
        float* dst = p.dw + ((long long)co * p.ntaps + tap) * p.Cin;
        red_add_16(dst + cib, v, os, p.Cin - cib);
      }
    }
    fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encoder(ugn_ctx* ctx, EncodeTiledFn* fn) {
  if (!ctx->encode_tiled) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    UGN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres));
    if (!f || qres != cudaDriverEntryPointSuccess) UGN_FAIL(UGN_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    ctx->encode_tiled = f;
  }
  *fn = (EncodeTiledFn)ctx->encode_tiled;
  return UGN_OK;
}

// 5-D bf16 tensor map; dims innermost first; strides[i] = byte stride of dim i+1.
static int make_map(ugn_ctx* ctx, CUtensorMap* map, const void* base, const uint64_t dims[5],
                    const uint64_t strides_bytes[4], const uint32_t box[5], int rowbytes) {
  EncodeTiledFn enc;
  int rc = get_encoder(ctx, &enc);
  if (rc != UGN_OK) return rc;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 5; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < 4; ++i) gs[i] = strides_bytes[i];
  UGN_CHECK(((uintptr_t)base & 15) == 0, "tensor-core operand must be 16-byte aligned");
  for (int i = 0; i < 4; ++i) UGN_CHECK(gs[i] % 16 == 0, "tensor-core operand stride %d (=%llu B) not a multiple of 16", i, (unsigned long long)gs[i]);
  for (int i = 0; i < 5; ++i) UGN_CHECK(bx[i] >= 1 && bx[i] <= 256, "TMA box dim %d = %u out of range", i, bx[i]);
  UGN_CHECK((int)(bx[0] * 2) == rowbytes, "inner box must span the swizzle width");
  CUtensorMapSwizzle sw = rowbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) UGN_FAIL(UGN_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return UGN_OK;
}

int tc_make_map(ugn_ctx* ctx, CUtensorMap* map, const void* base, const uint64_t dims[5],
                const uint64_t strides_bytes[4], const uint32_t box[5], int rowbytes) {
  return make_map(ctx, map, base, dims, strides_bytes, box, rowbytes);
}

static void finish_op(TcOp& op, int major, int rowbytes, int nbox, int box_rows, int plane_rows_reserved) {
  op.major = major;
  op.rowbytes = rowbytes;
  op.nbox = nbox;
  op.box_bytes = box_rows * rowbytes;
  int reserved = std::max(nbox * op.box_bytes, plane_rows_reserved * rowbytes);
  op.plane_bytes = (reserved + 1023) / 1024 * 1024;
  op.sbo = 8 * rowbytes;
  if (major == 0) { op.lbo = 0; op.kadv = 32; }
  else { op.lbo = op.box_bytes; op.kadv = 16 * rowbytes; }
}

template <int MODE>
static int launch(ugn_ctx* ctx, TcParams& p, dim3 grid, cudaStream_t st) {
  UGN_CHECK(ctx->cc_major == 10, "tensor-core path needs an sm_100 device (found sm_%d%d)", ctx->cc_major, ctx->cc_minor);
  UGN_CHECK(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= 256, "block_n=%d invalid", p.block_n);
  if (p.npass == 0) { p.npass = p.planes == 2 ? 3 : 1; p.pa = p.pb = p.planes; }
  size_t stage = (size_t)p.pa * p.a.plane_bytes + (size_t)p.pb * p.b.plane_bytes;
  int stages = (int)std::min<size_t>(8, (200 * 1024) / stage);
  UGN_CHECK(stages >= 2, "tensor-core tile does not fit shared memory (stage=%zu B)", stage);
  stages = std::min(stages, std::max(2, (p.ksteps_total + p.ksplit - 1) / p.ksplit));
  p.stages = stages;
  size_t smem = stages * stage + 1024 /*align*/ + (2 * stages + 1) * 8 + 16;
  smem = std::max(smem, (size_t)(128 * 33 * 4 + 2048));
  if (!ctx->err_flag) {
    UGN_CUDA(cudaMalloc(&ctx->err_flag, sizeof(int)));
    UGN_CUDA(cudaMemset(ctx->err_flag, 0, sizeof(int)));
  }
  p.err = ctx->err_flag;
#define TC_LAUNCH(NP, KSL)                                                                                      \
  do {                                                                                                         \
    UGN_CUDA(cudaFuncSetAttribute(tc_kernel<MODE, NP, KSL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    tc_kernel<MODE, NP, KSL><<<grid, kThreads, smem, st>>>(p);                                                 \
  } while (0)
#define TC_LAUNCH_K(NP)                                                                                         \
  do {                                                                                                         \
    if (p.kslices == 4) TC_LAUNCH(NP, 4); else if (p.kslices == 2) TC_LAUNCH(NP, 2); else TC_LAUNCH(NP, 0);    \
  } while (0)
  if (p.npass == 3) TC_LAUNCH_K(3);
  else if (p.npass == 1) TC_LAUNCH_K(1);
  else if (MODE != MODE_WGRAD && p.npass == 2) TC_LAUNCH_K(2);
  else if (MODE != MODE_WGRAD && p.npass == 4) TC_LAUNCH_K(4);
  else UGN_FAIL(UGN_ERR_UNSUPPORTED, "tc_kernel: npass=%d unsupported in mode %d", p.npass, MODE);
#undef TC_LAUNCH_K
#undef TC_LAUNCH
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---- plain GEMM -------------------------------------------------------------------------
static int gemm_operand(ugn_ctx* ctx, TcOp& op, const __nv_bfloat16* base, int P, int rows, int K, int mn_major,
                        int tile_rows, int BK) {
  uint64_t dims[5];
  uint64_t str[4];
  uint32_t box[5] = {1, 1, 1, 1, 1};
  if (!mn_major) {  // [P][rows][K]
    UGN_CHECK(K % 8 == 0, "K-major operand needs K %% 8 == 0 (K=%d)", K);
    dims[0] = K; dims[1] = rows; dims[2] = P; dims[3] = 1; dims[4] = 1;
    str[0] = (uint64_t)K * 2; str[1] = (uint64_t)rows * K * 2; str[2] = str[1] * P; str[3] = str[2];
    box[0] = 64; box[1] = tile_rows;
    finish_op(op, 0, 128, 1, tile_rows, tile_rows);
  } else {          // [P][K][rows]
    UGN_CHECK(rows % 8 == 0, "MN-major operand needs its contiguous extent %% 8 == 0 (got %d)", rows);
    UGN_CHECK(tile_rows >= 32, "MN-major operand tile must span >= 32 contiguous elements (64-byte swizzle rows)");
    const int cw = tile_rows >= 64 ? 64 : 32;     // narrow tiles (dense layers at small batch): 64-byte swizzle rows
    dims[0] = rows; dims[1] = K; dims[2] = P; dims[3] = 1; dims[4] = 1;
    str[0] = (uint64_t)rows * 2; str[1] = (uint64_t)rows * K * 2; str[2] = str[1] * P; str[3] = str[2];
    box[0] = cw; box[1] = BK;
    finish_op(op, 1, cw * 2, tile_rows / cw, BK, 0);
    return make_map(ctx, &op.map, base, dims, str, box, cw * 2);
  }
  return make_map(ctx, &op.map, base, dims, str, box, 128);
}

// Fused-epilogue request of a dense layer at small batch (M <= 128): NARROW output tiles (16 / 32 / 64 columns) give
// ~sm_count tiles without split-K, so the result goes straight from TMEM through bias / activation / mask into its f32
// and 16-bit destinations (no zero-fill, no red.add partial sums, no post pass), optionally with column sums.
struct GemmFuse {
  __nv_bfloat16* out16 = nullptr;   // [planes][M][ldc]
  int out16_planes = 0;
  int out16_raw = 0;                // convert acc * mask (scaled gradient operand) instead of the f32 result
  float* colsum = nullptr;          // [N], zeroed here
};

int tc_gemm_ex(ugn_ctx* ctx, int P, int f16, int M, int N, int K, const __nv_bfloat16* A, int a_mn,
               const __nv_bfloat16* B, int b_mn, float* C, int ldc, int accumulate, const float* bias,
               const float* mask, int act, float alpha, const float* oscale, cudaStream_t st,
               const GemmFuse* fz = nullptr) {
  TcParams p{};
  p.mode = MODE_GEMM; p.f16 = f16; p.oscale = oscale;
  p.M = M; p.N = N; p.planes = P;
  const bool narrow = fz != nullptr && M <= 128 && !accumulate && !getenv("UGN_NO_NARROW");
  if (P == 2 && ctx->gemm_npass) {       // forward dense layers: reduced pass count (ugn_set_fwd_passes)
    p.npass = ctx->gemm_npass;
    p.pa = (p.npass == 3 || p.npass == 4) ? 2 : 1;
    p.pb = (p.npass == 2 || p.npass == 3) ? 2 : 1;
  }
  p.block_n = N > 128 ? 256 : (N > 64 ? 128 : 64);
  if (P == 2 && p.block_n > 128 && !(p.npass == 1 || p.npass == 4)) p.block_n = 128;
  if (narrow) {
    // widest tile that still gives ~ one tile per SM; an MN-major B tile needs >= 32 contiguous columns
    p.block_n = b_mn ? 32 : 16;
    for (int bn : {64, 32}) if (ugn_cdiv(N, bn) * 10 >= ctx->sm_count * 8) { p.block_n = bn; break; }
  }
  p.kslices = 4;
  const int BK = 64;
  int rc;
  if ((rc = gemm_operand(ctx, p.a, A, P, M, K, a_mn, 128, BK)) != UGN_OK) return rc;
  if ((rc = gemm_operand(ctx, p.b, B, P, N, K, b_mn, p.block_n, BK)) != UGN_OK) return rc;
  p.ksteps_total = (K + BK - 1) / BK;
  int tiles = ugn_cdiv(M, 128) * ugn_cdiv(N, p.block_n);
  int split = 1;
  if (!bias && !mask && act == UGN_ACT_LINEAR && !narrow && !ctx->gemm_nosplit) {
    split = std::max(1, std::min(ctx->sm_count / std::max(tiles, 1), p.ksteps_total / 4));
    split = std::min(split, 32);
  }
  if (fz) {
    UGN_CHECK(narrow, "fused dense epilogue needs M <= 128 (got %d) and no accumulation", M);
    p.out16 = reinterpret_cast<u16*>(fz->out16); p.out16_planes = fz->out16_planes; p.out16_raw = fz->out16_raw;
    p.out16_plane = (long long)M * ldc;
    p.colsum = fz->colsum;
    if (p.colsum) UGN_CUDA(cudaMemsetAsync(p.colsum, 0, sizeof(float) * (size_t)N, st));
    UGN_CHECK(p.out16 || p.colsum || C, "fused dense epilogue without any output");
  }
  p.ksplit = split;
  p.epi = (split > 1 || accumulate) ? EPI_F32_ATOMIC : EPI_F32;
  p.out_f32 = C; p.ldc = ldc; p.bias = bias; p.mask = mask; p.act = act; p.alpha = alpha;
  if (split > 1 && !accumulate) UGN_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * ldc, st));
  dim3 grid(ugn_cdiv(M, 128), ugn_cdiv(N, p.block_n), split);
  return launch<MODE_GEMM>(ctx, p, grid, st);
}

// G[B,B] = X X^T from the hi/lo planes X16 [2][B][d]: three-pass split GEMM, fp32 accumulation in TMEM, NO split-K (one
// deterministic accumulation order per output element: the diagonal is consistent with the matrix)
int tc_gram(ugn_ctx* ctx, int f16, int B, int d, const __nv_bfloat16* X16, float* G, cudaStream_t st) {
  ctx->gemm_nosplit = 1;
  int rc = tc_gemm_ex(ctx, 2, f16, B, B, d, X16, 0, X16, 0, G, B, 0, nullptr, nullptr, UGN_ACT_LINEAR, 0.f, nullptr, st);
  ctx->gemm_nosplit = 0;
  return rc;
}

int tc_gemm(ugn_ctx* ctx, int P, int f16, int M, int N, int K, const __nv_bfloat16* A, int a_mn,
            const __nv_bfloat16* B, int b_mn, float* C, int accumulate, cudaStream_t st) {
  return tc_gemm_ex(ctx, P, f16, M, N, K, A, a_mn, B, b_mn, C, N, accumulate, nullptr, nullptr, UGN_ACT_LINEAR, 0.f,
                    nullptr, st);
}

// ---- convolution ------------------------------------------------------------------------
// activation tensor [P][B][H][W][C] as a 5-D map (C, W, H, B, P)
static int act_map(ugn_ctx* ctx, TcOp& op, const __nv_bfloat16* base, int P, int B, int H, int W, int C, int cbox,
                   int bw, int bh, int bn, int major, int nbox, int rows_reserved) {
  uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B, (uint64_t)P};
  uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)B * H * W * C * 2};
  uint32_t box[5] = {(uint32_t)cbox, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn, 1};
  finish_op(op, major, cbox * 2, nbox, bw * bh * bn, rows_reserved);
  return make_map(ctx, &op.map, base, dims, str, box, cbox * 2);
}

static void conv_box(int Wn, int Hn, int Bn, int pool, int& bw, int& bh, int& bn) {
  // spatial box of <= 128 output pixels covering the needed region Wn x Hn
  bw = Wn;
  if (pool && (bw & 1)) bw -= 1;
  bw = std::min(bw, 128);
  int rows = std::max(1, 128 / bw);
  bh = std::min(Hn, rows);
  if (pool) bh = std::max(2, bh & ~1);
  if (bh >= Hn) {
    bh = pool ? (Hn & ~1) : Hn;
    bn = std::max(1, std::min(Bn, 128 / (bw * bh)));
  } else {
    bn = 1;
  }
}

// patch-resident launch (forward with sgn=+1, input gradient with sgn=-1); returns UGN_ERR_UNSUPPORTED when
// the geometry does not fit so that the caller falls back to the per-tap-box kernel.
// one launch of the patch-resident conv kernel; the 2-CTA cluster (multicast weights) variant exists for the pass counts
// the training step uses (3-pass forward, 1-pass input gradient) and is only instantiated for them
template <bool C, bool PR, int TT, int KSL, int NP>
static int convp_do_launch(TcParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  if constexpr (!PR && (NP == 3 || NP == 1)) {
    if (p.cluster == 2) {
      auto kfn = tc_convp_kernel<C, false, TT, KSL, NP, 2>;
      UGN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaLaunchConfig_t lc = {};
      lc.gridDim = grid; lc.blockDim = dim3(kConvpThreads); lc.dynamicSmemBytes = smem; lc.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1;
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, kfn, &lc) == cudaSuccess && ncl > 0 && 2 * ncl < (int)lc.gridDim.x)
        lc.gridDim.x = 2 * ncl;      // persistent kernel: only as many clusters as are co-resident (GPC pairing)
      UGN_CUDA(cudaLaunchKernelEx(&lc, kfn, p));
      return UGN_OK;
    }
  }
  p.cluster = 1;
  UGN_CUDA(cudaFuncSetAttribute(tc_convp_kernel<C, PR, TT, KSL, NP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_convp_kernel<C, PR, TT, KSL, NP, 1><<<grid, kConvpThreads, smem, st>>>(p);
  return UGN_OK;
}

// description of the weight tensor map, so that convp_launch can re-encode it with a per-CTA share of the box when it
// decides to run thread-block clusters with multicast weight loads
struct WMapSpec {
  const void* base;
  uint64_t dims[5];
  uint64_t str[4];
  uint32_t box[5];
  int rowbytes;
};

static int convp_launch(ugn_ctx* ctx, TcParams& p, const __nv_bfloat16* act, int P, int B, int H, int W, int C,
                        int KH, int KW, int Wneed, int Hneed, int pool, int sgn, cudaStream_t st,
                        const WMapSpec* wspec = nullptr) {
  if (p.npass == 0) { p.npass = P == 2 ? 3 : 1; p.pa = p.pb = P; }
  const int PA = p.pa, PB = p.pb;
  const int cbox = (C % 64 == 0) ? 64 : 32;
  const int rowbytes = cbox * 2;
  int SW = 16;
  while (SW < (sgn > 0 ? W : Wneed + KW - 1)) SW <<= 1;
  if (SW > 64) return UGN_ERR_UNSUPPORTED;
  const int RH = 128 / SW;
  if (pool && (RH & 1)) return UGN_ERR_UNSUPPORTED;
  p.ncc = C / cbox; p.KW = KW; p.KH = KH; p.sgn = sgn;
  p.kslices = cbox / 16;
  p.ksteps_total = KH * KW * p.ncc; p.ksplit = 1;
  const size_t b_stage = (size_t)PB * p.b.plane_bytes;
  const size_t budget = 222 * 1024, fixed = 17 * 1024 + 2048;
  // accumulator columns per tile: block_n, or 2*block_n when hi*[hi|lo] is issued as one MMA (K-major
  // weight planes that are exactly adjacent in the ring stage)
  p.concat = (PB == 2 && sgn > 0 && p.b.major == 0 && p.block_n <= 128 && p.b.plane_bytes == p.block_n * rowbytes &&
              !getenv("UGN_NO_CONCAT")) ? 1 : 0;
  p.acc_tile_cols = p.concat ? 2 * p.block_n : p.block_n;
  int T = 0;
  p.cps = p.ncc; p.nslots = 1;
  for (int ring = 0; ring < 2 && !T; ++ring) {           // ring: stream one channel chunk at a time (2 slots)
    if (ring && p.ncc == 1) break;
    const int tforce = getenv("UGN_CONVP_T") ? atoi(getenv("UGN_CONVP_T")) : 0;     // (experiments)
    for (int cand : {2, 1}) {
      if (tforce && cand != tforce) continue;
      if (cand * p.acc_tile_cols > 512) continue;         // at least one accumulator set in the 512 TMEM columns
      if (cand > 1 && (cand - 1) * RH >= Hneed) continue;
      size_t chunk = (size_t)PA * (size_t)(cand * RH + KH - 1) * SW * rowbytes;
      size_t patch = ring ? 2 * chunk : p.ncc * chunk;
      if (patch + fixed + 2 * b_stage <= budget) { T = cand; p.cps = ring ? 1 : p.ncc; p.nslots = ring ? 2 : 1; break; }
    }
  }
  if (!T) return UGN_ERR_UNSUPPORTED;
  p.ngroups = p.ncc / p.cps;
  // hi*[hi|lo] concatenation doubles the accumulator columns: keep it only while TWO accumulator sets
  // still fit (measured: conv1-OF / conv3 with T = 1 gain 12-15 %; conv1-gray with T = 2 would fall back
  // to a single set and lose more to the exposed epilogue than the wider MMA wins)
  if (p.concat && 2 * T * p.acc_tile_cols > 512 && !getenv("UGN_FORCE_CONCAT")) { p.concat = 0; p.acc_tile_cols = p.block_n; }
  p.nbuf = (2 * T * p.acc_tile_cols <= 512) ? 2 : 1;      // single set (Cout = 192 tiles): the epilogue shares
                                                          // the tile-boundary bubble with the next patch load
  p.T = T; p.SW = SW; p.RH = RH; p.PR = T * RH + KH - 1;
  p.bw = SW; p.bh = RH; p.bn = 1;
  p.patch_chunk_bytes = p.PR * SW * rowbytes;
  p.patch_plane_bytes = p.cps * p.patch_chunk_bytes;
  p.xorg = sgn > 0 ? 0 : -(KW - 1);
  p.yorg = sgn > 0 ? 0 : -(KH - 1);
  p.tiles_y = ugn_cdiv(Hneed, T * RH);
  {
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B, (uint64_t)P};
    uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)B * H * W * C * 2};
    uint32_t box[5] = {(uint32_t)cbox, (uint32_t)SW, (uint32_t)p.PR, 1, 1};
    p.a.major = 0; p.a.rowbytes = rowbytes; p.a.nbox = 1; p.a.box_bytes = p.patch_chunk_bytes;
    int rc = make_map(ctx, &p.a.map, act, dims, str, box, rowbytes);
    if (rc != UGN_OK) return rc;
  }
  size_t patch = (size_t)p.nslots * PA * p.patch_plane_bytes;
  int stages = (int)std::min<size_t>(8, (budget - fixed - patch) / b_stage);
  stages = std::max(2, stages);
  p.stages = stages;
  size_t smem = patch + 17 * 1024 + stages * b_stage + 1024 + (2 * stages + 8) * 8 + 16;
  if (!ctx->err_flag) {
    UGN_CUDA(cudaMalloc(&ctx->err_flag, sizeof(int)));
    UGN_CUDA(cudaMemset(ctx->err_flag, 0, sizeof(int)));
  }
  p.err = ctx->err_flag;
  const int tiles_mn = p.tiles_y * B, tiles_n = ugn_cdiv(p.Cout, p.block_n);
  p.N = tiles_mn;
  p.M = tiles_mn * tiles_n;
  dim3 grid(std::min(p.M, ctx->sm_count), 1, 1);
  // ---- thread-block clusters of 2 with TMA-multicast weight tiles (OPT-IN, UGN_CONV_CLUSTER=1): two CTAs that work on
  // neighbouring tiles (same weights) each load half of every weight tile and multicast it to both -- half the L2 reads.
  // MEASURED (round 2, B = 96): no gain (conv1-OF 394 vs 394 us, conv1-gray 238 vs 237, conv3 60.8 vs 60.8).  The "wait B"
  // time of the MMA issuer in the role profile (UGN_CONVP_PROF: conv1-OF 216 of 644 kclk) is back-pressure, not
  // starvation: the issuer runs a 2-3 stage ring ahead of the tensor pipe and then waits for it to drain; per stage the
  // pipe is busy 676 of 751 clk.  The layer sits at 90 % of its MMA-issue bound (169 clk per K slice: the N = 96 floor).
  p.cluster = 1;
  const bool prof_ = getenv("UGN_CONVP_PROF") != nullptr;
  if (wspec && !prof_ && (p.npass == 3 || p.npass == 1) && getenv("UGN_CONV_CLUSTER") && tiles_mn % 2 == 0 &&
      p.M >= 2 * ctx->sm_count) {
    bool ok = false;
    if (p.b.major == 0) {
      const int half = p.block_n / 2;
      ok = p.block_n % 2 == 0 && (half * rowbytes) % 1024 == 0 && p.b.nbox == 1;
      if (ok) { p.b_half_rows = half; p.b_half_bytes = half * rowbytes; }
    } else {
      ok = p.b.nbox % 2 == 0;
    }
    if (ok) p.cluster = 2;
  }
  if (p.cluster == 2) {
    if (p.b.major == 0) {          // re-encode the weight map with this CTA's share of the box (half the rows)
      uint32_t box[5];
      for (int i = 0; i < 5; ++i) box[i] = wspec->box[i];
      box[2] = (uint32_t)p.b_half_rows;
      int rc = make_map(ctx, &p.b.map, wspec->base, wspec->dims, wspec->str, box, wspec->rowbytes);
      if (rc != UGN_OK) return rc;
    }
    grid.x = (unsigned)std::min<long long>(p.M, (ctx->sm_count / 2) * 2);
  }
  static long long* dbg_buf = nullptr;
  const bool prof = getenv("UGN_CONVP_PROF") != nullptr;
  if (prof) {
    if (!dbg_buf) UGN_CUDA(cudaMalloc(&dbg_buf, 8 * sizeof(long long)));
    UGN_CUDA(cudaMemsetAsync(dbg_buf, 0, 8 * sizeof(long long), st));
    p.dbg = dbg_buf;
  }
#define CONVP_LAUNCH(C, PR, TT, KSL, NP)                                                                       \
  do {                                                                                                         \
    int rcl = convp_do_launch<C, PR, TT, KSL, NP>(p, grid, smem, st);                                          \
    if (rcl != UGN_OK) return rcl;                                                                             \
  } while (0)
#define CONVP_T_K(C, PR, NP)                                                           \
  do {                                                                                 \
    if (p.T == 2 && p.kslices == 4) CONVP_LAUNCH(C, PR, 2, 4, NP);                     \
    else if (p.T == 2 && p.kslices == 2) CONVP_LAUNCH(C, PR, 2, 2, NP);                \
    else if (p.T == 1 && p.kslices == 4) CONVP_LAUNCH(C, PR, 1, 4, NP);                \
    else if (p.T == 1 && p.kslices == 2) CONVP_LAUNCH(C, PR, 1, 2, NP);                \
    else UGN_FAIL(UGN_ERR_UNSUPPORTED, "convp: unsupported T=%d kslices=%d", p.T, p.kslices); \
  } while (0)
  if (prof) {          // profiling build of the shapes of interest only (code size)
    if (p.concat && p.npass == 3) CONVP_T_K(true, true, 3);
    else if (p.concat) CONVP_T_K(true, true, 2);
    else if (p.npass == 3) CONVP_T_K(false, true, 3);
    else if (p.npass == 2) CONVP_T_K(false, true, 2);
    else CONVP_T_K(false, true, 1);
  } else if (p.concat && p.npass == 3) CONVP_T_K(true, false, 3);
  else if (p.concat) CONVP_T_K(true, false, 2);
  else if (p.npass == 3) CONVP_T_K(false, false, 3);
  else if (p.npass == 2) CONVP_T_K(false, false, 2);
  else if (p.npass == 4) CONVP_T_K(false, false, 4);
  else CONVP_T_K(false, false, 1);
#undef CONVP_T_K
#undef CONVP_LAUNCH
  UGN_LAUNCHED(ctx);
  if (prof) {   // development aid: synchronous read-back of the per-role wait cycles (averaged per CTA)
    long long h[8];
    UGN_CUDA(cudaStreamSynchronize(st));
    UGN_CUDA(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
    const double n = grid.x;
    fprintf(stderr, "[convp sgn=%d Co=%d bn=%d T=%d nbuf=%d concat=%d stages=%d nslots=%d iters/CTA=%.1f] per CTA kclk: "
            "mma{wait patch %.0f, wait B %.0f, wait tmem %.0f, total %.0f} epi{wait %.0f, work %.0f} "
            "prod{wait ring %.0f, wait patch slot %.0f}\n", sgn, p.Cout, p.block_n, p.T, p.nbuf, p.concat, p.stages,
            p.nslots, p.M / n, h[0] / n / 1e3, h[1] / n / 1e3, h[2] / n / 1e3, h[3] / n / 1e3, h[4] / n / 1e3,
            h[5] / n / 1e3, h[6] / n / 1e3, h[7] / n / 1e3);
  }
  return UGN_OK;
}

int ew_bias_act_split16(ugn_ctx* ctx, const float* acc, const float* bias, __nv_bfloat16* out, long long rows,
                        int cols, int P, int f16, int act, float alpha, cudaStream_t st);

int tc_conv_fwd(ugn_ctx* ctx, const ConvGeom& g, int P, int f16, const __nv_bfloat16* x, const __nv_bfloat16* w,
                const float* bias, __nv_bfloat16* y, uint8_t* idx, int act, float alpha, int pool,
                cudaStream_t st) {
  UGN_CHECK(g.Cp % 32 == 0 && g.Co % 16 == 0, "tensor-core conv needs Cin %% 32 == 0 and Cout %% 16 == 0");
  TcParams p{};
  p.mode = MODE_CONV; p.planes = P; p.sgn = 1; p.f16 = f16;
  if (P == 2 && ctx->fwd_conv_pass) {    // reduced pass count of the forward convolutions (ugn_set_fwd_passes)
    p.npass = ctx->fwd_conv_pass;
    p.pa = (p.npass == 3 || p.npass == 4) ? 2 : 1;
    p.pb = (p.npass == 2 || p.npass == 3) ? 2 : 1;
  }
  const int cbox = (g.Cp % 64 == 0) ? 64 : 32;
  p.kslices = cbox / 16;
  p.ncc = g.Cp / cbox; p.KW = g.KW;
  p.ksteps_total = g.KH * g.KW * p.ncc; p.ksplit = 1;
  const int Wn = pool ? g.Wp * 2 : g.Wo, Hn = pool ? g.Hp * 2 : g.Ho;
  p.Wout = pool ? Wn : g.Wo; p.Hout = pool ? Hn : g.Ho; p.Bn = g.B; p.Cout = g.Co;
  p.Hp = g.Hp; p.Wp = g.Wp;
  p.block_n = g.Co > 128 ? ((g.Co % 256 == 0 || g.Co > 192) ? 256 : (g.Co + 15) / 16 * 16) : (g.Co + 15) / 16 * 16;
  if (p.block_n > 256) p.block_n = 256;
  // split operands: N = 192 tiles run at the full tensor rate (96 clk per MMA), N <= 128 tiles cost 73 clk
  // each and are issued as hi*[hi|lo] pairs instead (convp_launch: concat)
  if (P == 2 && p.block_n > 128 && p.npass != 1)
    p.block_n = (g.Co % 192 == 0 && !getenv("UGN_NO_N192")) ? 192 : (g.Co % 128 == 0) ? 128 : ((g.Co % 96 == 0) ? 96 : 64);
  int rc;
  WMapSpec wspec{};
  {  // weights [P][Co][taps][Cp] K-major: dims (Cp, taps, Co, P, 1)
    const int taps = g.KH * g.KW;
    uint64_t dims[5] = {(uint64_t)g.Cp, (uint64_t)taps, (uint64_t)g.Co, (uint64_t)P, 1};
    uint64_t str[4] = {(uint64_t)g.Cp * 2, (uint64_t)taps * g.Cp * 2, (uint64_t)g.Co * taps * g.Cp * 2,
                       (uint64_t)P * g.Co * taps * g.Cp * 2};
    uint32_t box[5] = {(uint32_t)cbox, 1, (uint32_t)p.block_n, 1, 1};
    finish_op(p.b, 0, cbox * 2, 1, p.block_n, p.block_n);
    if ((rc = make_map(ctx, &p.b.map, w, dims, str, box, cbox * 2)) != UGN_OK) return rc;
    wspec.base = w; wspec.rowbytes = cbox * 2;
    for (int i = 0; i < 5; ++i) { wspec.dims[i] = dims[i]; wspec.box[i] = box[i]; }
    for (int i = 0; i < 4; ++i) wspec.str[i] = str[i];
  }
  p.epi = pool ? EPI_BF16_POOL : EPI_BF16_ACT;
  p.out_bf16 = y; p.out_plane = (long long)g.B * g.Hp * g.Wp * g.Co;
  p.pool_idx = idx; p.bias = bias; p.act = act; p.alpha = alpha;
  if (!getenv("UGN_NO_CONVP") && g.H * g.W > 16) {
    rc = convp_launch(ctx, p, x, P, g.B, g.H, g.W, g.Cp, g.KH, g.KW, Wn, Hn, pool, +1, st, &wspec);
    if (rc != UGN_ERR_UNSUPPORTED) return rc;
  }
  conv_box(Wn, Hn, g.B, pool, p.bw, p.bh, p.bn);
  p.ntx = ugn_cdiv(Wn, p.bw); p.nty = ugn_cdiv(Hn, p.bh);
  if ((rc = act_map(ctx, p.a, x, P, g.B, g.H, g.W, g.Cp, cbox, p.bw, p.bh, p.bn, 0, 1, 128)) != UGN_OK) return rc;
  dim3 grid(p.ntx * p.nty * ugn_cdiv(g.B, p.bn), ugn_cdiv(g.Co, p.block_n), 1);
  const int tiles = grid.x * grid.y;
  if (!pool && tiles * 2 <= ctx->sm_count && p.ksteps_total >= 8 && !getenv("UGN_NO_CONV_SPLITK")) {
    // few output tiles, long reduction (conv4: 3x3 maps, K = 2048): split the (tap, channel-chunk) steps
    // over the SMs into f32 partial sums (red.add), then bias + activation + 16-bit split in a post pass
    const long long rows = (long long)g.B * g.Ho * g.Wo;
    void* scratch = nullptr;
    if ((rc = ugn_scratch(ctx, sizeof(float) * rows * g.Co, &scratch)) != UGN_OK) return rc;
    UGN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * rows * g.Co, st));
    p.ksplit = std::max(1, std::min(ctx->sm_count / tiles, p.ksteps_total / 4));
    p.epi = EPI_F32_ATOMIC; p.out_f32 = reinterpret_cast<float*>(scratch);
    grid.z = p.ksplit;
    if ((rc = launch<MODE_CONV>(ctx, p, grid, st)) != UGN_OK) return rc;
    return ew_bias_act_split16(ctx, reinterpret_cast<float*>(scratch), bias, y, rows, g.Co, P, f16, act, alpha, st);
  }
  return launch<MODE_CONV>(ctx, p, grid, st);
}

int tc_conv_dgrad(ugn_ctx* ctx, const ConvGeom& g, int P, int f16, const __nv_bfloat16* dz, const __nv_bfloat16* w,
                  float* dx, cudaStream_t st) {
  UGN_CHECK(g.Co % 32 == 0 && g.Cp % 32 == 0, "tensor-core dgrad needs Cout %% 32 == 0 and Cin %% 32 == 0");
  const int kc = (g.Co % 64 == 0) ? 64 : 32;   // K stage = kc output channels
  TcParams p{};
  p.mode = MODE_CONV; p.planes = P; p.sgn = -1; p.f16 = f16; p.oscale = ctx->gscale ? ctx->gscale + 1 : nullptr;
  p.kslices = kc / 16;
  p.ncc = g.Co / kc; p.KW = g.KW;
  p.ksteps_total = g.KH * g.KW * p.ncc; p.ksplit = 1;
  p.Wout = g.W; p.Hout = g.H; p.Bn = g.B; p.Cout = g.Cp;
  const int cw = (g.Cp % 64 == 0) ? 64 : 32;
  int bn_cols = std::min(g.Cp, P == 2 ? 128 : 256);
  bn_cols = bn_cols / cw * cw;
  p.block_n = bn_cols;
  int rc;
  {  // weights as MN-major B: N = ci (contiguous), K = co: dims (Cp, taps, Co, P, 1), box (cw, 1, 64, 1, 1)
    const int taps = g.KH * g.KW;
    uint64_t dims[5] = {(uint64_t)g.Cp, (uint64_t)taps, (uint64_t)g.Co, (uint64_t)P, 1};
    uint64_t str[4] = {(uint64_t)g.Cp * 2, (uint64_t)taps * g.Cp * 2, (uint64_t)g.Co * taps * g.Cp * 2,
                       (uint64_t)P * g.Co * taps * g.Cp * 2};
    uint32_t box[5] = {(uint32_t)cw, 1, (uint32_t)kc, 1, 1};
    finish_op(p.b, 1, cw * 2, p.block_n / cw, kc, 0);
    if ((rc = make_map(ctx, &p.b.map, w, dims, str, box, cw * 2)) != UGN_OK) return rc;
  }
  p.epi = EPI_F32; p.out_f32 = dx;
  if (!getenv("UGN_NO_CONVP") && g.H * g.W >= 400) {   // small maps: the per-tap-box kernel is faster
    rc = convp_launch(ctx, p, dz, P, g.B, g.Ho, g.Wo, g.Co, g.KH, g.KW, g.W, g.H, 0, -1, st);
    if (rc != UGN_ERR_UNSUPPORTED) return rc;
    p.ncc = g.Co / kc; p.kslices = kc / 16; p.ksteps_total = g.KH * g.KW * p.ncc;
  }
  conv_box(g.W, g.H, g.B, 0, p.bw, p.bh, p.bn);
  p.ntx = ugn_cdiv(g.W, p.bw); p.nty = ugn_cdiv(g.H, p.bh);
  if ((rc = act_map(ctx, p.a, dz, P, g.B, g.Ho, g.Wo, g.Co, kc, p.bw, p.bh, p.bn, 0, 1, 128)) != UGN_OK) return rc;
  dim3 grid(p.ntx * p.nty * ugn_cdiv(g.B, p.bn), ugn_cdiv(g.Cp, p.block_n), 1);
  const int tiles = grid.x * grid.y;
  if (tiles * 2 <= ctx->sm_count && p.ksteps_total >= 8 && !getenv("UGN_NO_CONV_SPLITK")) {
    p.ksplit = std::max(1, std::min(ctx->sm_count / tiles, p.ksteps_total / 4));
    p.epi = EPI_F32_ATOMIC;
    grid.z = p.ksplit;
    UGN_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)g.B * g.H * g.W * g.Cp, st));
  }
  return launch<MODE_CONV>(ctx, p, grid, st);
}

int simt_colsum_bf16(ugn_ctx* ctx, const __nv_bfloat16* X, int P, int f16, long long rows, int cols, float* out,
                     cudaStream_t st);

static int wgradv_launch(ugn_ctx* ctx, const ConvGeom& g, int P, int f16, const __nv_bfloat16* x, const __nv_bfloat16* dz,
                         float* dw, cudaStream_t st) {
  WgParams p{};
  p.planes = P; p.f16 = f16; p.oscale = ctx->gscale ? ctx->gscale + 1 : nullptr;
  p.cw = (g.Cp % 64 == 0) ? 64 : 32;
  p.rowbytes_b = p.cw * 2;
  const int nch = g.Cp / p.cw;
  int PW = 8;
  while (PW < g.Wo + g.KW - 1) PW <<= 1;
  if (PW > 64 || g.Ho * g.Wo <= 16) return UGN_ERR_UNSUPPORTED;   // tiny maps: the per-tap-box kernel is faster
  int bh = 64 / PW, bn = 1;
  int hcap = 1;
  while (hcap * 2 <= g.Ho) hcap <<= 1;
  if (bh > hcap) { bn = bh / hcap; bh = hcap; }
  p.PW = PW; p.bh = bh; p.bn = bn;
  p.nby = ugn_cdiv(g.Ho, bh);
  p.ksteps_total = p.nby * ugn_cdiv(g.B, bn);
  p.KW = g.KW; p.ntaps = g.KH * g.KW; p.Cin = g.Cin; p.Co = g.Co;
  p.a_box_bytes = 64 * 128;
  p.a_plane_bytes = 2 * p.a_box_bytes;
  p.b_box_bytes = (64 + 8) * p.rowbytes_b;
  // segments
  const int max_nkw = std::max(1, 256 / p.cw);
  int nseg = 0, nbox = 0, ntile = 0;
  int cols = 0, boxes_in_tile = 0;
  p.tile_seg0[0] = 0; p.tile_box0[0] = 0;
  for (int kh = 0; kh < g.KH; ++kh)
    for (int ch = 0; ch < nch; ++ch) {
      int need = g.KW * p.cw;
      if (cols + need > 512 || boxes_in_tile == 4) {   // close the current N-tile
        if (ntile >= 32) return UGN_ERR_UNSUPPORTED;
        ++ntile; p.tile_seg0[ntile] = nseg; p.tile_box0[ntile] = nbox; cols = 0; boxes_in_tile = 0;
      }
      if (nbox >= 40) return UGN_ERR_UNSUPPORTED;
      p.box_kh[nbox] = kh; p.box_chunk[nbox] = ch;
      for (int kw0 = 0; kw0 < g.KW; kw0 += max_nkw) {
        if (nseg >= 40) return UGN_ERR_UNSUPPORTED;
        int nkw = std::min(max_nkw, g.KW - kw0);
        if ((nkw * p.cw) % 16 != 0) return UGN_ERR_UNSUPPORTED;
        p.seg[nseg] = WgSeg{nbox, kh, ch, kw0, nkw, cols};
        cols += nkw * p.cw;
        ++nseg;
      }
      ++nbox; ++boxes_in_tile;
    }
  ++ntile; p.tile_seg0[ntile] = nseg; p.tile_box0[ntile] = nbox;
  p.ntiles_n = ntile;
  int max_boxes = 0;
  for (int j = 0; j < ntile; ++j) max_boxes = std::max(max_boxes, p.tile_box0[j + 1] - p.tile_box0[j]);
  if (max_boxes > 4) return UGN_ERR_UNSUPPORTED;
  for (int j = 0; j < ntile; ++j)
    if (p.tile_seg0[j + 1] - p.tile_seg0[j] > 8) return UGN_ERR_UNSUPPORTED;   // kMaxSeg of the issue loop
  // ring stage sized by the boxes an N-tile really uses (1-3), not the 4-slot maximum: more stages in flight
  p.b_plane_bytes = (max_boxes * p.b_box_bytes + 1023) / 1024 * 1024;
  size_t stage = (size_t)P * (p.a_plane_bytes + p.b_plane_bytes);
  int stages = (int)std::min<size_t>(8, (220 * 1024) / stage);
  if (stages < 2) return UGN_ERR_UNSUPPORTED;
  p.stages = stages;
  int rc;
  EncodeTiledFn enc;
  if ((rc = get_encoder(ctx, &enc)) != UGN_OK) return rc;
  {
    uint64_t dims[5] = {(uint64_t)g.Co, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.B, (uint64_t)P};
    uint64_t str[4] = {(uint64_t)g.Co * 2, (uint64_t)g.Wo * g.Co * 2, (uint64_t)g.Ho * g.Wo * g.Co * 2,
                       (uint64_t)g.B * g.Ho * g.Wo * g.Co * 2};
    uint32_t box[5] = {64, (uint32_t)PW, (uint32_t)bh, (uint32_t)bn, 1};
    if ((rc = make_map(ctx, &p.amap, dz, dims, str, box, 128)) != UGN_OK) return rc;
  }
  {
    uint64_t dims[5] = {(uint64_t)g.Cp, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B, (uint64_t)P};
    uint64_t str[4] = {(uint64_t)g.Cp * 2, (uint64_t)g.W * g.Cp * 2, (uint64_t)g.H * g.W * g.Cp * 2,
                       (uint64_t)g.B * g.H * g.W * g.Cp * 2};
    uint32_t box[5] = {(uint32_t)p.cw, (uint32_t)PW, (uint32_t)bh, (uint32_t)bn, 1};
    if ((rc = make_map(ctx, &p.bmap, x, dims, str, box, p.rowbytes_b)) != UGN_OK) return rc;
  }
  const int ntile_m = ugn_cdiv(g.Co, 128);
  int split = std::max(1, std::min(ctx->sm_count / std::max(1, ntile * ntile_m), p.ksteps_total / 2));
  p.ksplit = std::min(split, 65535);
  p.dw = dw;
  if (!ctx->err_flag) {
    UGN_CUDA(cudaMalloc(&ctx->err_flag, sizeof(int)));
    UGN_CUDA(cudaMemset(ctx->err_flag, 0, sizeof(int)));
  }
  p.err = ctx->err_flag;
  size_t smem = stages * stage + 1024 + (2 * stages + 1) * 8 + 16;
  UGN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)g.Co * p.ntaps * g.Cin, st));
  dim3 grid(ntile_m, ntile, p.ksplit);
  if (P == 2) {
    UGN_CUDA(cudaFuncSetAttribute(tc_wgradv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_wgradv_kernel<true><<<grid, kThreads, smem, st>>>(p);
  } else {
    UGN_CUDA(cudaFuncSetAttribute(tc_wgradv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_wgradv_kernel<false><<<grid, kThreads, smem, st>>>(p);
  }
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

int tc_conv_wgrad(ugn_ctx* ctx, const ConvGeom& g, int P, int f16, const __nv_bfloat16* x, const __nv_bfloat16* dz,
                  float* dw, float* db, cudaStream_t st) {
  UGN_CHECK(g.Cp % 32 == 0 && g.Co % 8 == 0, "tensor-core wgrad needs Cin %% 32 == 0");
  if (!getenv("UGN_NO_WGRADV")) {
    int rcv = wgradv_launch(ctx, g, P, f16, x, dz, dw, st);
    if (rcv == UGN_OK) return db ? simt_colsum_bf16(ctx, dz, P, f16, (long long)g.B * g.Ho * g.Wo, g.Co, db, st) : UGN_OK;
    if (rcv != UGN_ERR_UNSUPPORTED) return rcv;
  }
  TcParams p{};
  p.mode = MODE_WGRAD; p.planes = P; p.f16 = f16; p.oscale = ctx->gscale ? ctx->gscale + 1 : nullptr;
  // K stage = a box of output pixels (rows zero-filled by TMA beyond the dz extent)
  int bw = 1;
  while (bw < g.Wo && bw < 64) bw <<= 1;
  int bh = std::max(1, 64 / bw);
  bh = std::min(bh, 1 << (31 - __builtin_clz(std::max(1, g.Ho))));
  if (bh < 1) bh = 1;
  int bn = std::max(1, 64 / (bw * bh));
  while (bw * bh * bn > 64) bn >>= 1;
  if (bn < 1) bn = 1;
  while ((bw * bh * bn) % 16 != 0) bn *= 2;
  p.bw = bw; p.bh = bh; p.bn = bn;
  p.kslices = bw * bh * bn / 16;
  p.nbx = ugn_cdiv(g.Wo, bw); p.nby = ugn_cdiv(g.Ho, bh);
  p.ksteps_total = p.nbx * p.nby * ugn_cdiv(g.B, bn);
  p.KW = g.KW; p.ntaps = g.KH * g.KW; p.Cin = g.Cin; p.Co = g.Co;
  p.cw = (g.Cp % 64 == 0) ? 64 : 32;
  p.nch = g.Cp / p.cw;
  const int maxcols = P == 2 ? 128 : 256;
  int nboxB = std::max(1, std::min(maxcols / p.cw, p.nch * p.ntaps));
  p.block_n = nboxB * p.cw;
  int rc = act_map(ctx, p.a, dz, P, g.B, g.Ho, g.Wo, g.Co, 64, bw, bh, bn, 1, 2, 0);
  if (rc != UGN_OK) return rc;
  if ((rc = act_map(ctx, p.b, x, P, g.B, g.H, g.W, g.Cp, p.cw, bw, bh, bn, 1, nboxB, 0)) != UGN_OK) return rc;
  int ntile_n = ugn_cdiv(p.nch * p.ntaps, nboxB), ntile_m = ugn_cdiv(g.Co, 128);
  int split = std::max(1, std::min((2 * ctx->sm_count) / std::max(1, ntile_n * ntile_m), p.ksteps_total / 2));
  split = std::min(split, 65535);
  p.ksplit = split;
  p.epi = EPI_F32_ATOMIC; p.out_f32 = dw;
  UGN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)g.Co * p.ntaps * g.Cin, st));
  dim3 grid(ntile_m, ntile_n, split);
  rc = launch<MODE_WGRAD>(ctx, p, grid, st);
  if (rc != UGN_OK) return rc;
  if (db) return simt_colsum_bf16(ctx, dz, P, f16, (long long)g.B * g.Ho * g.Wo, g.Co, db, st);
  return UGN_OK;
}

// ---- dense ------------------------------------------------------------------------------
int ew_bias_act_mask(ugn_ctx* ctx, float* y, const float* bias, const float* mask, long long rows, int cols,
                     int act, float alpha, cudaStream_t st);
int ew_split(ugn_ctx*, const float*, __nv_bfloat16*, int, int, long long, cudaStream_t);
int ew_dense_post(ugn_ctx* ctx, float* y, const float* bias, const float* mask, __nv_bfloat16* out16, int P16, int f16,
                  const float* scale16, float* colsum, int write_f32, long long rows, int cols, int act, float alpha,
                  cudaStream_t st, const unsigned long long* rng = nullptr, int layer = 0, float keep = 1.f);
int ew_dropout_mask(ugn_ctx* ctx, const unsigned long long* rng, int layer, float keep, float* out, long long n,
                    cudaStream_t st);

int tc_linear_fwd(ugn_ctx* ctx, int P, int f16, int B, int N, int K, const __nv_bfloat16* x, const __nv_bfloat16* w,
                  const float* bias, const float* mask, float* y, int act, float alpha, cudaStream_t st,
                  __nv_bfloat16* y16, int P16) {
  int rc;
  ctx->gemm_npass = ctx->fwd_dense_pass;
  if (B <= 128 && N % 16 == 0 && getenv("UGN_NARROW")) {
    // EXPERIMENT (opt-in, measured slower: 0.36 vs 0.23 ms/step of dense forward at B = 96): narrow N tiles spread the
    // weight stream over the SMs WITHOUT split-K and the whole epilogue leaves the accumulator in one pass -- but every
    // CTA then re-reads the full 128-row activation tile per K step (32 KB next to 8 KB of weights), and the per-SM TMA
    // ingest, not HBM, bounds the kernel.  Default below: 128-wide tiles + split-K + ONE fused post pass.
    GemmFuse fz;
    fz.out16 = y16; fz.out16_planes = P16;
    rc = tc_gemm_ex(ctx, P, f16, B, N, K, x, 0, w, 0, y, N, 0, bias, mask, act, alpha, nullptr, st, &fz);
    ctx->gemm_npass = 0;
    return rc;
  }
  int tiles = ugn_cdiv(B, 128) * ugn_cdiv(N, P == 2 ? 128 : 256);
  const bool post_route = tiles * 2 <= ctx->sm_count && K >= 512;
  if (ctx->drop_rng && !post_route) {
    // large batch: the GEMM epilogue applies the mask itself -> materialise the Philox mask once (same bits as the
    // backward pass regenerates)
    void* scr = nullptr;
    if ((rc = ugn_scratch(ctx, sizeof(float) * (size_t)B * N, &scr)) != UGN_OK) return rc;
    if ((rc = ew_dropout_mask(ctx, ctx->drop_rng, ctx->drop_layer, ctx->drop_keep, reinterpret_cast<float*>(scr),
                              (long long)B * N, st)) != UGN_OK) return rc;
    mask = reinterpret_cast<const float*>(scr);
    ctx->drop_rng = nullptr;
  }
  if (post_route) {
    rc = tc_gemm_ex(ctx, P, f16, B, N, K, x, 0, w, 0, y, N, 0, nullptr, nullptr, UGN_ACT_LINEAR, 0.f, nullptr, st);
    ctx->gemm_npass = 0;
    if (rc != UGN_OK) return rc;
    // split-K partial sums -> ONE post pass: bias + activation + dropout (mask tensor | Philox) + f32 result + 16-bit planes
    if (bias || mask || act != UGN_ACT_LINEAR || y16 || ctx->drop_rng)
      return ew_dense_post(ctx, y, bias, mask, y16, P16, f16, nullptr, nullptr, 1, B, N, act, alpha, st, ctx->drop_rng,
                           ctx->drop_layer, ctx->drop_keep);
    return UGN_OK;
  } else {
    rc = tc_gemm_ex(ctx, P, f16, B, N, K, x, 0, w, 0, y, N, 0, bias, mask, act, alpha, nullptr, st);
    ctx->gemm_npass = 0;
  }
  if (rc != UGN_OK || !y16) return rc;
  return ew_split(ctx, y, y16, P16, f16, (long long)B * N, st);
}

int tc_linear_bwd(ugn_ctx* ctx, int P, int f16, int B, int N, int K, const __nv_bfloat16* x, const __nv_bfloat16* w,
                  const __nv_bfloat16* dz, float* dx, float* dw, float* db, cudaStream_t st, const float* dx_mask,
                  __nv_bfloat16* dx16, int P16, float* dbx) {
  int rc;
  const float* os = ctx->gscale ? ctx->gscale + 1 : nullptr;   // dz is a scaled gradient operand
  const bool want_dx = dx || dx16 || dbx;
  // dx[B,K] = dz[B,N] . w[N,K]      : A = dz K-major (K'=N), B = w as MN-major [K'=N rows][K contiguous]
  if (want_dx && B <= 128 && K % 32 == 0 && getenv("UGN_NARROW")) {      // (experiment, see tc_linear_fwd)
    GemmFuse fz;
    fz.out16 = dx16; fz.out16_planes = P16; fz.out16_raw = 1; fz.colsum = dbx;
    if ((rc = tc_gemm_ex(ctx, P, f16, B, K, N, dz, 0, w, 1, dx, K, 0, nullptr, dx_mask, 0, 0.f, os, st, &fz)) != UGN_OK) return rc;
  } else if (want_dx) {
    UGN_CHECK(dx, "linear_bwd: dx (f32 [B,K], may be scratch) is required: the split-K partial sums land there");
    if ((rc = tc_gemm_ex(ctx, P, f16, B, K, N, dz, 0, w, 1, dx, K, 0, nullptr, nullptr, 0, 0.f, os, st)) != UGN_OK) return rc;
    // ONE post pass over dx: x dropout mask of the layer below -> its 16-bit gradient operand (re-scaled by the
    // gradient scale) + its bias gradient; dx itself stays the unmasked f32 gradient
    if (dx_mask || dx16 || dbx || ctx->drop_rng)
      if ((rc = ew_dense_post(ctx, dx, nullptr, dx_mask, dx16, P16, f16, ctx->gscale, dbx, 0, B, K, UGN_ACT_LINEAR, 0.f, st,
                              ctx->drop_rng, ctx->drop_layer, ctx->drop_keep)) != UGN_OK) return rc;
  }
  // dw[N,K] = dz^T . x             : A = dz MN-major [K'=B rows][N contiguous], B = x MN-major [B rows][K contiguous]
  if (dw && (rc = tc_gemm_ex(ctx, P, f16, N, K, B, dz, 1, x, 1, dw, K, 0, nullptr, nullptr, 0, 0.f, os, st)) != UGN_OK) return rc;
  if (db) return simt_colsum_bf16(ctx, dz, P, f16, B, N, db, st);
  return UGN_OK;
}
