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
//   warps 2..5  epilogue     : thread r owns query row r: tcgen05.ld its 256 scores,
//                              s = |g|^2 - 2 q.g, compare with the row's threshold (a register) and keep
//                              the KC best (score, index) pairs in REGISTERS (unsorted, replace-worst) --
//                              no shared-memory lists, no atomics.  Overlaps the next tile's MMAs.
// The candidate lists go to the same exact fp64 re-rank as the SIMT scan (knn.cu); a containment check
// there proves that the true top-k lie inside the candidate set, else the query is recomputed exactly.
#include "tc.cuh"
#include "tc_ptx.cuh"
#include <float.h>

using namespace tc;

struct Cand {
  float s;
  int i;
};

struct alignas(64) KnnTcParams {
  CUtensorMap qmap, gmap;      // [P][rows][Dp] fp16, box (64, 128|256, 1)
  const float* g2;             // [>= roundup(N,256)], +inf beyond N
  Cand* out;                   // [Q][chunks][kc]
  int Q, kc, ksteps, chunks;
  long long N, rows_per_chunk;
  int* err;
};

static constexpr int KT_THREADS = 192;
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

  if (threadIdx.x == 0) {
    for (int s = 0; s < KT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 128); mbar_init(&g2_full[b], 1); }
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
            tma_load_5d(&p.qmap, a_full, ares + ks * KT_A_CHUNK + pl * KT_A_PLANE, ks * 64, q0, pl, 0, 0);
      }
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const long long g0 = c0 + (long long)t * KT_TN;
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
            if (!RES) tma_load_5d(&p.qmap, &full[s], sa + pl * KT_A_PLANE, ks * 64, q0, pl, 0, 0);
            tma_load_5d(&p.gmap, &full[s], sb + pl * KT_B_PLANE, ks * 64, (int)g0, pl, 0, 0);
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
    // ---- epilogue: thread <-> query row; KC best candidates in registers ----
    const int q = warp & 3;
    const int r = q * 32 + lane;
    float cs[KC];
    int ci[KC];
#pragma unroll
    for (int j = 0; j < KC; ++j) { cs[j] = FLT_MAX; ci[j] = 0x7fffffff; }
    float thr = FLT_MAX;     // worst (largest) score in the list; candidates must be strictly better
    int thr_j = 0;           // its slot
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const long long g0 = c0 + (long long)t * KT_TN;
      mbar_wait(&g2_full[buf], (t >> 1) & 1, p.err, 7);
      bool ok = mbar_wait(&tmem_full[buf], (t >> 1) & 1, p.err, 3);
      fence_after_sync();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * KT_TN;
      const float* gn = g2s + buf * KT_TN;
      float v[16];
      for (int cb = 0; cb < KT_TN; cb += 16) {
        tmem_ld16(trow + cb, v);
        if (!ok) continue;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float sc = fmaf(-2.f, v[i], gn[cb + i]);   // +inf for rows beyond N (norm strip padding)
          if (sc < thr) {
            // replace the worst entry, then find the new worst (ties -> the larger index is worse;
            // columns arrive in increasing index order, so `sc < thr` keeps the lower index on ties)
            const int gi = (int)(g0 + cb + i);
#pragma unroll
            for (int j = 0; j < KC; ++j)
              if (j == thr_j) { cs[j] = sc; ci[j] = gi; }
            float ws = cs[0];
            int wi = ci[0], wj = 0;
#pragma unroll
            for (int j = 1; j < KC; ++j)
              if (cs[j] > ws || (cs[j] == ws && ci[j] > wi)) { ws = cs[j]; wi = ci[j]; wj = j; }
            thr = ws;
            thr_j = wj;
          }
        }
      }
      fence_before_sync();
      mbar_arrive(&tmem_empty[buf]);
    }
    const int qi = q0 + r;
    if (qi < p.Q) {
      Cand* o = p.out + ((long long)qi * p.chunks + blockIdx.y) * KC;
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

// q16 [2][Q][Dp], g16 [2][N][Dp] fp16 planes; cands [Q][chunks][kc].  Returns the chunk count used.
int knn_tc_scan(ugn_ctx* ctx, const __nv_bfloat16* q16, const __nv_bfloat16* g16, const float* g2, int Q,
                long long N, int Dp, int kc, int chunks, long long rows_per_chunk, void* cands, cudaStream_t st) {
  UGN_CHECK(ctx->cc_major == 10, "tensor-core k-NN needs an sm_100 device");
  UGN_CHECK(Dp % 8 == 0, "k-NN operand planes need Dp %% 8 == 0 (got %d)", Dp);
  KnnTcParams p{};
  int rc;
  {
    uint64_t dims[5] = {(uint64_t)Dp, (uint64_t)Q, 2, 1, 1};
    uint64_t str[4] = {(uint64_t)Dp * 2, (uint64_t)Q * Dp * 2, (uint64_t)Q * Dp * 4, (uint64_t)Q * Dp * 4};
    uint32_t box[5] = {64, 128, 1, 1, 1};
    if ((rc = tc_make_map(ctx, &p.qmap, q16, dims, str, box, 128)) != UGN_OK) return rc;
  }
  {
    uint64_t dims[5] = {(uint64_t)Dp, (uint64_t)N, 2, 1, 1};
    uint64_t str[4] = {(uint64_t)Dp * 2, (uint64_t)N * Dp * 2, (uint64_t)N * Dp * 4, (uint64_t)N * Dp * 4};
    uint32_t box[5] = {64, (uint32_t)(((Dp + 63) / 64 <= 4 && !getenv("UGN_KNN_STREAM")) ? 128 : 256), 1, 1, 1};
    if ((rc = tc_make_map(ctx, &p.gmap, g16, dims, str, box, 128)) != UGN_OK) return rc;
  }
  p.g2 = g2; p.out = reinterpret_cast<Cand*>(cands);
  p.Q = Q; p.kc = kc; p.ksteps = (Dp + 63) / 64; p.chunks = chunks;
  p.N = N; p.rows_per_chunk = rows_per_chunk;
  if (!ctx->err_flag) {
    UGN_CUDA(cudaMalloc(&ctx->err_flag, sizeof(int)));
    UGN_CUDA(cudaMemset(ctx->err_flag, 0, sizeof(int)));
  }
  p.err = ctx->err_flag;
  const bool res = p.ksteps <= 4 && !getenv("UGN_KNN_STREAM");   // query tile (ksteps x 32 KB) fits next to a 3-stage ring
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
