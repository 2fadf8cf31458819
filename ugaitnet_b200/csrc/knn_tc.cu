// Tensor-core candidate scan of the open-world k-NN (mains/mj_testUWYHGaitNet_open_tum.py:331-341,
// KNeighborsClassifier.predict): the distance GEMM  S = Q . G^T  runs on tcgen05 (fp16 hi/lo split
// operands, 3 MMA passes, fp32 accumulate in TMEM) and the top-KC filter is FUSED into its epilogue:
// the Q x N score matrix only ever exists as 128 x 256 accumulator tiles in TMEM.
//
//   warp 0      TMA producer : RESIDENT mode (Dp <= 256): the 128-query tile (all K chunks, both planes,
//                              <= 128 KB) is loaded ONCE and stays in smem; only gallery boxes [128 x 64]
//                              stream through a 3-stage ring -- half the L2->SM bytes per MMA.
//                              STREAMING mode (any Dp): per K-step one query box [128 x 64] and one gallery
//                              box [256 x 64] per plane into a 2-stage ring.  Per gallery tile the row norms
//                              (cp.async.bulk) go into a double-buffered smem strip.
//   warp 1      MMA issuer   : hi*hi + hi*lo + lo*hi into one of two TMEM accumulators (128 x TN fp32)
//   warps 2..9  epilogue     : two warps per TMEM lane quadrant (two per SM sub-partition, so that one
//                              hides the other's latencies), each owning one half of the tile's columns:
//                              a thread tcgen05.ld's 16 scores of ITS query row, forms
//                              s = |g|^2 - 2 q.g, min-reduces the 16 and compares ONCE with the row's
//                              threshold (a register); only a hit walks the 16 values and updates the KC
//                              best (score, index) pairs kept in REGISTERS (unsorted, replace-worst) -- no
//                              shared-memory lists, no atomics.  Overlaps the next tile's MMAs.
// The candidate lists go to the same exact fp64 re-rank as the SIMT scan (knn.cu); a containment check
// there proves that the true top-k lie inside the candidate set, else the query is recomputed exactly.
#include "tc.cuh"
#include "tc_ptx.cuh"
#include <float.h>
#include <algorithm>

using namespace tc;

struct Cand {
  float s;
  int i;
};

struct alignas(64) KnnTcParams {
  CUtensorMap qmap, gmap;      // fp16 [2 planes][K chunks][rows][64]: every (plane, chunk, row tile) box is contiguous
  const float* g2;             // [>= roundup(N,256)], +inf beyond N
  Cand* out;                   // [Q][chunks][2][kc]
  int* gthr;                   // [Q] shared per-query threshold (order-preserving int key of a score)
  int Q, kc, ksteps, chunks;
  long long N, rows_per_chunk;
  int skew;                    // rotate each query tile's sweep through its chunk (de-synchronises the CTAs)
  int* err;
};

static constexpr int KT_THREADS = 320;                  // producer warp + MMA warp + 8 epilogue warps
static constexpr int KT_A_PLANE = 128 * 128;            // 128 query rows x 64 fp16
static constexpr int KT_A_CHUNK = 2 * KT_A_PLANE;       // both planes of one K chunk: 32 KB
// RES (query tile resident): gallery tile 128 rows, ring stage = gallery planes only (32 KB), 3 stages
// streaming              : gallery tile 256 rows, ring stage = query + gallery planes (96 KB), 2 stages
template <bool RES> struct KtCfg {
  static constexpr int TN = RES ? 128 : 256;
  static constexpr int STAGES = RES ? 3 : 2;
  static constexpr int B_PLANE = TN * 128;
  static constexpr int STAGE = 2 * B_PLANE + (RES ? 0 : KT_A_CHUNK);
};

// order-preserving float <-> int key (atomicMin on scores of either sign)
__device__ __forceinline__ int score_key(float f) {
  int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key_score(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int KC, bool RES>
__global__ void __launch_bounds__(KT_THREADS, 1) knn_tc_scan_kernel(const __grid_constant__ KnnTcParams p) {
  constexpr int KT_TN = KtCfg<RES>::TN, KT_STAGES = KtCfg<RES>::STAGES, KT_B_PLANE = KtCfg<RES>::B_PLANE,
                KT_STAGE = KtCfg<RES>::STAGE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ares = smem + KT_STAGES * KT_STAGE;                                  // RES: [ksteps][2 planes][16 KB]
  float* g2s = reinterpret_cast<float*>(ares + (RES ? p.ksteps * KT_A_CHUNK : 0));   // [2][TN]
  uint64_t* full = reinterpret_cast<uint64_t*>(g2s + 2 * KT_TN);
  uint64_t* empty = full + KT_STAGES;
  uint64_t* tmem_full = empty + KT_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint64_t* g2_full = tmem_empty + 2;        // [2]
  uint64_t* a_full = g2_full + 2;            // [1] resident query tile landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const long long c0 = (long long)blockIdx.y * p.rows_per_chunk;
  const long long c1 = min(p.N, c0 + p.rows_per_chunk);
  const int ntiles = c1 > c0 ? (int)((c1 - c0 + KT_TN - 1) / KT_TN) : 0;
  // CTAs of different query tiles sweep the same chunk: start each at a different tile so that they do not
  // all request the same L2 lines at the same moment
  const int t_rot = (p.skew && ntiles > 0) ? (int)(((long long)blockIdx.x * ntiles) / gridDim.x) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < KT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 256); mbar_init(&g2_full[b], 1); }
    mbar_init(a_full, 1);
    fence_mbar_init();
    prefetch_tmap(&p.qmap);
    prefetch_tmap(&p.gmap);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 2 * KT_TN);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0, ph = 0;
      if (RES && ntiles > 0) {
        mbar_expect_tx(a_full, p.ksteps * KT_A_CHUNK);
        for (int ks = 0; ks < p.ksteps; ++ks)
          for (int pl = 0; pl < 2; ++pl)
            tma_load_5d(&p.qmap, a_full, ares + ks * KT_A_CHUNK + pl * KT_A_PLANE, 0, q0, ks, pl, 0);
      }
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const long long g0 = c0 + (long long)((t + t_rot) % ntiles) * KT_TN;
        // the norm strip of this accumulator buffer is free once its previous tile has been drained
        mbar_wait(&tmem_empty[buf], ((t >> 1) & 1) ^ 1, p.err, 6);
        mbar_expect_tx(&g2_full[buf], KT_TN * 4);
        bulk_load(g2s + buf * KT_TN, p.g2 + g0, KT_TN * 4, &g2_full[buf]);
        for (int ks = 0; ks < p.ksteps; ++ks) {
          mbar_wait(&empty[s], ph ^ 1, p.err, 1);
          mbar_expect_tx(&full[s], KT_STAGE);
          uint8_t* sa = smem + (size_t)s * KT_STAGE;
          uint8_t* sb = sa + (RES ? 0 : KT_A_CHUNK);
          for (int pl = 0; pl < 2; ++pl) {
            if (!RES) tma_load_5d(&p.qmap, &full[s], sa + pl * KT_A_PLANE, 0, q0, ks, pl, 0);
            tma_load_5d(&p.gmap, &full[s], sb + pl * KT_B_PLANE, 0, (int)g0, ks, pl, 0);
          }
          if (++s == KT_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc16(128, KT_TN, 0, 0, 1);
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t a0 = make_smem_desc(RES ? smem_u32(ares) : smem0, 0, 1024, 2u);
      const uint64_t b0 = make_smem_desc(smem0 + (RES ? 0 : KT_A_CHUNK), 0, 1024, 2u);
      const uint32_t st16 = KT_STAGE >> 4, pa16 = KT_A_PLANE >> 4, pb16 = KT_B_PLANE >> 4, ac16 = KT_A_CHUNK >> 4;
      int s = 0, ph = 0;
      if (RES && ntiles > 0) {
        mbar_wait(a_full, 0, p.err, 8);
        fence_after_sync();
      }
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        mbar_wait(&tmem_empty[buf], ((t >> 1) & 1) ^ 1, p.err, 6);
        fence_after_sync();
        const uint32_t tacc = tmem_base + buf * KT_TN;
        for (int ks = 0; ks < p.ksteps; ++ks) {
          mbar_wait(&full[s], ph, p.err, 2);
          fence_after_sync();
          uint64_t a_hi = a0 + (uint64_t)(RES ? ks * ac16 : s * st16), b_hi = b0 + (uint64_t)(s * st16);
          for (int k = 0; k < 4; ++k) {
            umma_f16(tacc, a_hi, b_hi, idesc, (ks > 0 || k > 0) ? 1u : 0u);
            umma_f16(tacc, a_hi, b_hi + pb16, idesc, 1);
            umma_f16(tacc, a_hi + pa16, b_hi, idesc, 1);
            a_hi += 2;
            b_hi += 2;
          }
          umma_commit(&empty[s]);
          if (++s == KT_STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ---- epilogue: thread <-> (query row, column half); KC best candidates in registers ----
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    constexpr int HC = KT_TN / 2;       // columns per epilogue thread and tile
    float cs[KC];
    int ci[KC];
#pragma unroll
    for (int j = 0; j < KC; ++j) { cs[j] = FLT_MAX; ci[j] = 0x7fffffff; }
    // own_s/own_i/own_j: worst entry of THIS thread's list = lexicographically largest (score, index).
    // thr: the acceptance threshold = min(own worst, shared per-query threshold).  The shared threshold is
    // the smallest "KC-th best so far" any CTA / column half working on this query has published: a score
    // above it can never be among the query's KC best, so the list warm-up (KC ln n updates) is paid once
    // per query instead of once per chunk.  Rows beyond Q (zero-filled by TMA) never accept anything.
    const bool live = (q0 + r) < p.Q;
    float own_s = FLT_MAX, thr = live ? FLT_MAX : -FLT_MAX;
    int own_i = 0x7fffffff, own_j = 0, thr_i = 0x7fffffff;
    int* gq = p.gthr + (live ? q0 + r : 0);
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const long long g0 = c0 + (long long)((t + t_rot) % ntiles) * KT_TN + half * HC;
      if (live) {
        const float gs = key_score(*reinterpret_cast<volatile int*>(gq));
        if (gs < thr) { thr = gs; thr_i = 0x7fffffff; }     // ties with a foreign threshold are kept
      }
      mbar_wait(&g2_full[buf], (t >> 1) & 1, p.err, 7);
      bool ok = mbar_wait(&tmem_full[buf], (t >> 1) & 1, p.err, 3);
      fence_after_sync();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * KT_TN + half * HC;
      const float* gn = g2s + buf * KT_TN + half * HC;
      float v[16];
      for (int cb = 0; cb < HC; cb += 16) {
        tmem_ld16(trow + cb, v);
        if (!ok) continue;
        float m = FLT_MAX;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 n4 = *reinterpret_cast<const float4*>(gn + cb + i);
          v[i] = fmaf(-2.f, v[i], n4.x);            // +inf for rows beyond N (norm strip padding)
          v[i + 1] = fmaf(-2.f, v[i + 1], n4.y);
          v[i + 2] = fmaf(-2.f, v[i + 2], n4.z);
          v[i + 3] = fmaf(-2.f, v[i + 3], n4.w);
          m = fminf(m, fminf(fminf(v[i], v[i + 1]), fminf(v[i + 2], v[i + 3])));
        }
        if (m <= thr) {      // rare once the threshold has tightened
          bool changed = false;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float sc = v[i];
            const int gi = (int)(g0 + cb + i);
            if (sc < thr || (sc == thr && gi < thr_i)) {
              // replace this list's worst entry, then find the new worst (ties -> larger index is worse)
#pragma unroll
              for (int j = 0; j < KC; ++j)
                if (j == own_j) { cs[j] = sc; ci[j] = gi; }
              float ws = cs[0];
              int wi = ci[0], wj = 0;
#pragma unroll
              for (int j = 1; j < KC; ++j)
                if (cs[j] > ws || (cs[j] == ws && ci[j] > wi)) { ws = cs[j]; wi = ci[j]; wj = j; }
              own_s = ws; own_i = wi; own_j = wj;
              if (own_s < thr || (own_s == thr && own_i < thr_i)) { thr = own_s; thr_i = own_i; }
              changed = true;
            }
          }
          // publish a full list's worst score (an upper bound of the query's KC-th best)
          if (changed && own_s < FLT_MAX) atomicMin(gq, score_key(own_s));
        }
      }
      fence_before_sync();
      mbar_arrive(&tmem_empty[buf]);
    }
    const int qi = q0 + r;
    if (qi < p.Q) {
      Cand* o = p.out + (((long long)qi * p.chunks + blockIdx.y) * 2 + half) * KC;
#pragma unroll
      for (int j = 0; j < KC; ++j) { o[j].s = cs[j]; o[j].i = ci[j]; }
    }
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    fence_after_sync();
    tmem_dealloc(tmem_base, 2 * KT_TN);
  }
}

int tc_make_map(ugn_ctx* ctx, CUtensorMap* map, const void* base, const uint64_t dims[5],
                const uint64_t strides_bytes[4], const uint32_t box[5], int rowbytes);

// f32 [rows][D] -> fp16 hi/lo planes in the scan's operand layout [2][KCH][rows][64] (KCH = ceil(D/64),
// zero padded): one thread per (row, chunk, 8 columns), 16-byte stores.
__global__ void knn_pack_kernel(const float* __restrict__ X, long long rows, int D, int KCH,
                                __nv_bfloat16* __restrict__ out) {
  const long long total = rows * KCH * 8;
  const long long plane = rows * KCH * 64;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(e & 7);
    const long long rc = e >> 3;
    const int kc = (int)(rc % KCH);
    const long long row = rc / KCH;
    const int col0 = kc * 64 + c8 * 8;
    __align__(16) u16 hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float v = (col0 + i) < D ? X[row * D + col0 + i] : 0.f;
      ugn_split16(v, 1, hi[i], lo[i]);
    }
    const long long o = ((long long)kc * rows + row) * 64 + c8 * 8;
    *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(out + plane + o) = *reinterpret_cast<const uint4*>(lo);
  }
}

int knn_tc_pack(ugn_ctx* ctx, const float* X, long long rows, int D, __nv_bfloat16* out, cudaStream_t st) {
  const int KCH = (D + 63) / 64;
  const long long total = rows * KCH * 8;
  if (total == 0) return UGN_OK;
  int grid = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16);
  knn_pack_kernel<<<grid, 256, 0, st>>>(X, rows, D, KCH, out);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// q16 / g16: fp16 planes [2][KCH][rows][64] (knn_tc_pack); cands [Q][chunks][2][kc].
int knn_tc_scan(ugn_ctx* ctx, const __nv_bfloat16* q16, const __nv_bfloat16* g16, const float* g2, int Q,
                long long N, int KCH, int kc, int chunks, long long rows_per_chunk, void* cands, int* gthr,
                cudaStream_t st) {
  UGN_CHECK(ctx->cc_major == 10, "tensor-core k-NN needs an sm_100 device");
  KnnTcParams p{};
  const bool res = KCH <= 4 && !getenv("UGN_KNN_STREAM");   // query tile (KCH x 32 KB) fits next to a 3-stage ring
  int rc;
  {
    uint64_t dims[5] = {64, (uint64_t)Q, (uint64_t)KCH, 2, 1};
    uint64_t str[4] = {128, (uint64_t)Q * 128, (uint64_t)Q * KCH * 128, (uint64_t)Q * KCH * 256};
    uint32_t box[5] = {64, 128, 1, 1, 1};
    if ((rc = tc_make_map(ctx, &p.qmap, q16, dims, str, box, 128)) != UGN_OK) return rc;
  }
  {
    uint64_t dims[5] = {64, (uint64_t)N, (uint64_t)KCH, 2, 1};
    uint64_t str[4] = {128, (uint64_t)N * 128, (uint64_t)N * KCH * 128, (uint64_t)N * KCH * 256};
    uint32_t box[5] = {64, (uint32_t)(res ? 128 : 256), 1, 1, 1};
    if ((rc = tc_make_map(ctx, &p.gmap, g16, dims, str, box, 128)) != UGN_OK) return rc;
  }
  p.g2 = g2; p.out = reinterpret_cast<Cand*>(cands); p.gthr = gthr;
  UGN_CUDA(cudaMemsetAsync(gthr, 0x7f, sizeof(int) * (size_t)Q, st));   // key of 3.4e38: "no threshold yet"
  p.Q = Q; p.kc = kc; p.ksteps = KCH; p.chunks = chunks;
  p.N = N; p.rows_per_chunk = rows_per_chunk;
  // measured: lock-step sweeps (all query tiles of a chunk read the same gallery lines at the same time)
  // are FASTER (4.92 vs 6.35 ms on 1 M x 256, Q = 4096): L2 serves the shared lines once; skew is opt-in
  p.skew = getenv("UGN_KNN_SKEW") ? 1 : 0;
  if (!ctx->err_flag) {
    UGN_CUDA(cudaMalloc(&ctx->err_flag, sizeof(int)));
    UGN_CUDA(cudaMemset(ctx->err_flag, 0, sizeof(int)));
  }
  p.err = ctx->err_flag;
  size_t smem = res ? (size_t)KtCfg<true>::STAGES * KtCfg<true>::STAGE + (size_t)p.ksteps * KT_A_CHUNK + 2 * 128 * 4
                    : (size_t)KtCfg<false>::STAGES * KtCfg<false>::STAGE + 2 * 256 * 4;
  smem += 16 * 8 + 16 + 1024;
  dim3 grid((Q + 127) / 128, chunks);
#define KNN_TC_LAUNCH(KC)                                                                                     \
  do {                                                                                                        \
    if (res) {                                                                                                \
      UGN_CUDA(cudaFuncSetAttribute(knn_tc_scan_kernel<KC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      knn_tc_scan_kernel<KC, true><<<grid, KT_THREADS, smem, st>>>(p);                                        \
    } else {                                                                                                  \
      UGN_CUDA(cudaFuncSetAttribute(knn_tc_scan_kernel<KC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      knn_tc_scan_kernel<KC, false><<<grid, KT_THREADS, smem, st>>>(p);                                       \
    }                                                                                                         \
  } while (0)
  if (kc == 8) KNN_TC_LAUNCH(8);
  else if (kc == 16) KNN_TC_LAUNCH(16);
  else if (kc == 32) KNN_TC_LAUNCH(32);
  else UGN_FAIL(UGN_ERR_INVALID, "k-NN candidate count %d unsupported", kc);
#undef KNN_TC_LAUNCH
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}
