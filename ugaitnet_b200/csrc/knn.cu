// Open-world k-NN (mains/mj_testUWYHGaitNet_open_tum.py:331-341):
//   KNeighborsClassifier(n_neighbors=k).fit(gallery, labels).predict(probe)
// brute-force Euclidean, uniform vote, vote ties -> smallest label.
//
// Stage 1  scan   : approximate score s(q,g) = |g|^2 - 2 q.g (|q|^2 is rank-irrelevant) tile by
//                   tile with a FUSED top-KC filter -- the Q x N matrix is never materialised.
//                   Every CTA keeps, per query row, a sorted (score, idx) list in shared memory
//                   and a threshold; tile scores that beat the threshold are appended with a
//                   shared-memory atomic and merged between tiles (the append rate decays as KC/n).
// Stage 2  rerank : the CTAs' candidate lists are merged per query, the surviving KC candidates
//                   are re-ranked with EXACT fp64 sum((q-g)^2) and ordered (distance, index) --
//                   the same rule as oracle/knn_oracle.c, so indices/labels are bit-exact.
// Stage 3  vote   : merge of G shards' results + uniform vote (multi-GPU gallery sharding).
#include "common.cuh"
#include <float.h>

#define KNN_TQ 64     // queries per CTA
#define KNN_TG 64     // gallery rows per tile
#define KNN_TK 16
#define KNN_MAXKC 32
#define KNN_CB 64     // append buffer entries per row (>= KNN_TG so one tile can never overflow)

struct Cand {
  float s;
  int i;
};
__device__ __forceinline__ bool cand_less(float s0, int i0, float s1, int i1) {
  return s0 < s1 || (s0 == s1 && i0 < i1);
}

// sorted insert of (s,i) into list[0..kc) (ascending by (s,i)); list is full-length with
// +inf sentinels.  Called by ONE thread per row.
__device__ __forceinline__ void list_insert(Cand* list, int kc, float s, int i) {
  if (!cand_less(s, i, list[kc - 1].s, list[kc - 1].i)) return;
  int p = kc - 1;
  while (p > 0 && cand_less(s, i, list[p - 1].s, list[p - 1].i)) {
    list[p] = list[p - 1];
    --p;
  }
  list[p].s = s;
  list[p].i = i;
}

// g2[row] = |g_row|^2 for row < N, +inf for the padding rows N <= row < Npad (the tensor-core scan reads
// whole 256-row strips; +inf keeps padding rows out of every candidate list); gmax2 = max_row |g|^2.
__global__ void knn_norms_kernel(const float* __restrict__ G, long long N, long long Npad, int D,
                                 float* __restrict__ g2, unsigned* __restrict__ gmax2) {
  long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= Npad) return;
  if (row >= N) {
    if (lane == 0) g2[row] = INFINITY;
    return;
  }
  const float* g = G + row * D;
  float s = 0.f;
  for (int j = lane; j < D; j += 32) s = fmaf(g[j], g[j], s);
  s = warp_sum(s);
  if (lane == 0) {
    g2[row] = s;
    if (gmax2 && s > 0.f && s < INFINITY) atomicMax(gmax2, __float_as_uint(s));
  }
}

// dynamic smem layout (bytes): top[TQ][kc] Cand | buf[TQ][CB] Cand | As | Bs | cnt[TQ] | thr[TQ]
__global__ void __launch_bounds__(256) knn_scan_kernel(const float* __restrict__ Qm,
                                                       const float* __restrict__ Gm,
                                                       const float* __restrict__ g2, int Q, long long N,
                                                       int D, int kc, long long rows_per_chunk,
                                                       Cand* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  Cand* top = reinterpret_cast<Cand*>(smraw);
  Cand* buf = top + KNN_TQ * kc;
  float* As = reinterpret_cast<float*>(buf + KNN_TQ * KNN_CB);       // [TK][TQ+4]
  float* Bs = As + KNN_TK * (KNN_TQ + 4);                            // [TK][TG+4]
  int* cnt = reinterpret_cast<int*>(Bs + KNN_TK * (KNN_TG + 4));
  float* thr = reinterpret_cast<float*>(cnt + KNN_TQ);

  const int t = threadIdx.x;
  const int q0 = blockIdx.x * KNN_TQ;
  const long long c0 = (long long)blockIdx.y * rows_per_chunk;
  const long long c1 = min(N, c0 + rows_per_chunk);
  for (int e = t; e < KNN_TQ * kc; e += 256) { top[e].s = FLT_MAX; top[e].i = 0x7fffffff; }
  if (t < KNN_TQ) { cnt[t] = 0; thr[t] = FLT_MAX; }
  __syncthreads();

  const int kk = t & 15, r = t >> 4, ty = t >> 4, tx = t & 15;
  for (long long gt = c0; gt < c1; gt += KNN_TG) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int k0 = 0; k0 < D; k0 += KNN_TK) {
      int k = k0 + kk;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        int row = r + 16 * jj;
        int qi = q0 + row;
        long long gi = gt + row;
        As[kk * (KNN_TQ + 4) + row] = (k < D && qi < Q) ? __ldg(Qm + (long long)qi * D + k) : 0.f;
        Bs[kk * (KNN_TG + 4) + row] = (k < D && gi < c1) ? __ldg(Gm + gi * D + k) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < KNN_TK; ++q) {
        float4 av = *reinterpret_cast<const float4*>(&As[q * (KNN_TQ + 4) + ty * 4]);
        float4 bv = *reinterpret_cast<const float4*>(&Bs[q * (KNN_TG + 4) + tx * 4]);
        float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
      }
      __syncthreads();
    }
    // fused top-k filter
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      long long gi = gt + tx * 4 + b;
      if (gi >= c1) continue;
      float n2 = g2[gi];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        int row = ty * 4 + a;
        float s = n2 - 2.f * acc[a][b];
        if (s <= thr[row]) {
          int pos = atomicAdd(&cnt[row], 1);
          buf[row * KNN_CB + pos].s = s;
          buf[row * KNN_CB + pos].i = (int)gi;
        }
      }
    }
    __syncthreads();
    if (t < KNN_TQ) {
      int c = cnt[t];
      if (c) {
        Cand* lst = top + t * kc;
        for (int e = 0; e < c; ++e) list_insert(lst, kc, buf[t * KNN_CB + e].s, buf[t * KNN_CB + e].i);
        thr[t] = lst[kc - 1].s;
        cnt[t] = 0;
      }
    }
    __syncthreads();
  }
  // write this CTA's lists: out[q][chunk][kc]
  for (int e = t; e < KNN_TQ * kc; e += 256) {
    int row = e / kc, j = e % kc;
    int qi = q0 + row;
    if (qi < Q) out[((long long)qi * gridDim.y + blockIdx.y) * kc + j] = top[e];
  }
}

// one CTA per query: select the KC best approximate candidates over all chunks, re-rank them
// exactly in fp64 and emit the k nearest as (d2, global idx, label).
__global__ void __launch_bounds__(256) knn_rerank_kernel(const float* __restrict__ Qm,
                                                         const float* __restrict__ Gm,
                                                         const int* __restrict__ labels,
                                                         const Cand* __restrict__ cands, int ncand, int kc,
                                                         int k, int D, long long idx_base,
                                                         double* __restrict__ out_d2,
                                                         long long* __restrict__ out_idx,
                                                         int* __restrict__ out_lab,
                                                         const float* __restrict__ gmax2, int Dp,
                                                         int* __restrict__ flags) {
  __shared__ Cand sel[KNN_MAXKC];
  __shared__ double q2s;
  __shared__ double ex[KNN_MAXKC];
  __shared__ float rs[8];
  __shared__ int ri[8], rp[8];
  const int q = blockIdx.x, t = threadIdx.x;
  const Cand* cq = cands + (long long)q * ncand;
  // KC rounds of block-wide lexicographic argmin over entries greater than the previous pick
  float ps = -FLT_MAX;
  int pi = -1;
  for (int round = 0; round < kc; ++round) {
    float bs = FLT_MAX;
    int bi = 0x7fffffff;
    for (int e = t; e < ncand; e += 256) {
      float s = cq[e].s;
      int i = cq[e].i;
      if (cand_less(ps, pi, s, i) && cand_less(s, i, bs, bi)) { bs = s; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      float os = __shfl_xor_sync(0xffffffffu, bs, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (cand_less(os, oi, bs, bi)) { bs = os; bi = oi; }
    }
    if ((t & 31) == 0) { rs[t >> 5] = bs; ri[t >> 5] = bi; }
    __syncthreads();
    if (t == 0) {
      for (int w = 1; w < 8; ++w)
        if (cand_less(rs[w], ri[w], bs, bi)) { bs = rs[w]; bi = ri[w]; }
      sel[round].s = bs;
      sel[round].i = bi;
    }
    __syncthreads();
    ps = sel[round].s;
    pi = sel[round].i;
  }
  const float ps_last = (pi == 0x7fffffff) ? FLT_MAX : ps;   // KC-th best approximate score (tau)
  // exact fp64 distances, one warp per candidate
  const float* qv = Qm + (long long)q * D;
  if (flags && t < 32) {
    double s = 0.0;
    for (int j = t; j < D; j += 32) s = fma((double)qv[j], (double)qv[j], s);
    s = warp_sum_d(s);
    if (t == 0) q2s = s;
  }
  for (int c = t >> 5; c < kc; c += 8) {
    int gi = sel[c].i;
    double s = 0.0;
    if (gi != 0x7fffffff) {
      const float* gv = Gm + (long long)gi * D;
      for (int j = t & 31; j < D; j += 32) {
        double df = (double)qv[j] - (double)gv[j];
        s = fma(df, df, s);
      }
      s = warp_sum_d(s);
    } else {
      s = DBL_MAX;
    }
    if ((t & 31) == 0) ex[c] = s;
  }
  __syncthreads();
  if (t == 0) {
    // selection sort of <= 32 entries by (d2, idx)
    (void)rp;
    double kth_d2 = 0.0;
    for (int j = 0; j < k; ++j) {
      int best = -1;
      for (int c = 0; c < kc; ++c) {
        if (sel[c].i == -2) continue;
        if (best < 0 || ex[c] < ex[best] || (ex[c] == ex[best] && sel[c].i < sel[best].i)) best = c;
      }
      int gi = sel[best].i;
      out_d2[(long long)q * k + j] = ex[best];
      kth_d2 = ex[best];
      out_idx[(long long)q * k + j] = (gi == 0x7fffffff) ? -1 : idx_base + gi;
      out_lab[(long long)q * k + j] = (gi == 0x7fffffff) ? -1 : labels[gi];
      sel[best].i = -2;
    }
    if (flags) {
      // Containment proof for the tensor-core candidate scan.  Every gallery row that is NOT a candidate
      // has approximate score >= tau (the KC-th best approximate score over all chunks), hence exact
      // score >= tau - eps.  If the exact k-th best candidate score is below that, no outsider can belong
      // to the true top-k and the result is exact; otherwise the query is flagged and recomputed by
      // brute force in fp64 (knn_exact_kernel).
      //   eps: fp16 hi/lo split drops lo*lo and rounds lo (<= 2^-20 |q||g| in total), fp32 accumulation
      //   of Dp terms (16-sigma random-walk bound), fp16 subnormal lo planes (2^-24 absolute per element),
      //   the fp32 row norm (Dp/32+8 roundings) and the final fmaf.
      const float tau = ps_last;
      int bad = 0;
      if (tau < FLT_MAX) {
        const double qn = sqrt(q2s), gm2 = (double)*gmax2, gn = sqrt(gm2), sq = sqrt((double)Dp);
        const double u = 5.9604644775390625e-08;  // 2^-24
        double eps_dot = qn * gn * (16.0 * u + 16.0 * sq * u) + u * sq * (qn + gn);
        double eps = 2.0 * eps_dot + gm2 * ((double)Dp / 32.0 + 8.0) * 2.0 * u + 4.0 * u * (gm2 + 2.0 * qn * gn);
        const double sk = kth_d2 - q2s;          // exact score |g|^2 - 2 q.g of the k-th neighbour
        bad = !(sk + eps < (double)tau);
      }
      flags[q] = bad;
    }
  }
}

// merge G shards' [G,Q,k] lists by (d2, idx), emit the k best and the uniform vote.
// sd / si / sl: element stride between two shards' lists (Q*k when the three arrays are separate and contiguous; the
// packed all-gather buffer of the sharded search keeps {d2 | idx | lab} of one rank back to back)
__global__ void knn_merge_vote_kernel(const double* __restrict__ d2, const long long* __restrict__ idx,
                                      const int* __restrict__ lab, long long sd, long long si, long long sl,
                                      int G, int Q, int k,
                                      double* __restrict__ od2, long long* __restrict__ oidx,
                                      int* __restrict__ olab, int* __restrict__ pred) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  int head[16];
  for (int g = 0; g < G; ++g) head[g] = 0;
  int labs[KNN_MAXKC];
  for (int j = 0; j < k; ++j) {
    int bg = -1;
    double bd = 0.0;
    long long bi = 0;
    for (int g = 0; g < G; ++g) {
      if (head[g] >= k) continue;
      const long long o = (long long)q * k + head[g];
      long long ii = idx[g * si + o];
      if (ii < 0) continue;
      double dd = d2[g * sd + o];
      if (bg < 0 || dd < bd || (dd == bd && ii < bi)) { bg = g; bd = dd; bi = ii; }
    }
    int lb = -1;
    if (bg >= 0) {
      lb = lab[bg * sl + (long long)q * k + head[bg]];
      head[bg]++;
    } else {
      bd = DBL_MAX; bi = -1;
    }
    labs[j] = lb;
    if (od2) od2[(long long)q * k + j] = bd;
    if (oidx) oidx[(long long)q * k + j] = bi;
    if (olab) olab[(long long)q * k + j] = lb;
  }
  // uniform vote, ties -> smallest label
  int best = 0x7fffffff, bestc = -1;
  for (int a = 0; a < k; ++a) {
    if (labs[a] < 0) continue;
    int c = 0;
    for (int b = 0; b < k; ++b) c += labs[b] == labs[a];
    if (c > bestc || (c == bestc && labs[a] < best)) { bestc = c; best = labs[a]; }
  }
  pred[q] = bestc < 0 ? -1 : best;
}

static int knn_kc(int k) { return k <= 4 ? 8 : (k <= 8 ? 16 : 32); }
static int knn_chunks(ugn_ctx* ctx, long long Q, long long N) {
  long long qt = (Q + KNN_TQ - 1) / KNN_TQ;
  long long want = std::max<long long>(1, (2LL * ctx->sm_count + qt - 1) / qt);
  long long maxc = std::max<long long>(1, N / (4 * KNN_TG));
  return (int)std::min<long long>(std::min(want, maxc), 1024);
}

extern "C" int64_t ugn_knn_workspace_bytes(int64_t Q, int64_t N, int64_t D, int k) {
  (void)D;
  // worst case chunk count (sm_count unknown here): 1024
  return Q * 1024 * (int64_t)knn_kc(k) * (int64_t)sizeof(Cand) + Q * 4;
}

extern "C" int ugn_knn_gallery_norms(ugn_ctx* ctx, const ugn_tensor* gallery, ugn_tensor* g2,
                                     ugn_tensor* gmax2, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UGN_CHECK(ctx && gallery && g2, "ugn_knn_gallery_norms: null argument");
  UGN_TENSOR(gallery, DT_F32, 2, 2);
  UGN_TENSOR(g2, DT_F32, 1, 1);
  long long N = gallery->shape[0], Npad = g2->shape[0];
  int D = (int)gallery->shape[1];
  UGN_CHECK(Npad >= N, "g2 must be f32[>= N]");
  unsigned* gm = nullptr;
  if (gmax2) {
    UGN_TENSOR(gmax2, DT_F32, 1, 1);
    gm = reinterpret_cast<unsigned*>(ugn_ptr<float>(gmax2));
    UGN_CUDA(cudaMemsetAsync(gm, 0, sizeof(float), st));
  }
  if (Npad == 0) return UGN_OK;
  knn_norms_kernel<<<ugn_cdiv(Npad * 32, 256), 256, 0, st>>>(ugn_ptr<float>(gallery), N, Npad, D,
                                                              ugn_ptr<float>(g2), gm);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

extern "C" int ugn_knn_topk(ugn_ctx* ctx, const ugn_tensor* queries, const ugn_tensor* gallery,
                            const ugn_tensor* g2, const ugn_tensor* gallery_labels, int k,
                            int64_t idx_base, ugn_tensor* out_d2, ugn_tensor* out_idx,
                            ugn_tensor* out_lab, ugn_tensor* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UGN_CHECK(ctx && queries && gallery && g2 && gallery_labels && out_d2 && out_idx && out_lab && workspace,
            "ugn_knn_topk: null argument");
  UGN_TENSOR(queries, DT_F32, 2, 2);
  UGN_TENSOR(gallery, DT_F32, 2, 2);
  UGN_TENSOR(g2, DT_F32, 1, 1);
  UGN_TENSOR(gallery_labels, DT_I32, 1, 1);
  UGN_TENSOR(out_d2, DT_F64, 2, 2);
  UGN_TENSOR(out_idx, DT_I64, 2, 2);
  UGN_TENSOR(out_lab, DT_I32, 2, 2);
  UGN_TENSOR(workspace, DT_BAD, 1, 8);
  long long Q = queries->shape[0], N = gallery->shape[0];
  int D = (int)queries->shape[1];
  UGN_CHECK(gallery->shape[1] == D, "gallery/query dimension mismatch (%lld vs %d)", (long long)gallery->shape[1], D);
  UGN_CHECK(k >= 1 && k <= KNN_MAXKC, "k must be in [1,%d]", KNN_MAXKC);
  UGN_CHECK(N >= k, "gallery shard has fewer rows (%lld) than k=%d", N, k);
  UGN_CHECK(N < 0x7fffffffLL, "gallery shard too large for 32-bit local indices");
  UGN_CHECK(g2->shape[0] >= N && gallery_labels->shape[0] == N, "g2/labels must have N entries");
  UGN_CHECK(out_d2->shape[0] == Q && out_d2->shape[1] == k && out_idx->shape[0] == Q && out_idx->shape[1] == k &&
                out_lab->shape[0] == Q && out_lab->shape[1] == k, "outputs must be [Q,k]");
  if (Q == 0) return UGN_OK;
  int kc = knn_kc(k);
  int chunks = knn_chunks(ctx, Q, N);
  long long rows = (N + chunks - 1) / chunks;
  rows = (rows + KNN_TG - 1) / KNN_TG * KNN_TG;
  chunks = (int)((N + rows - 1) / rows);
  long long need = Q * chunks * (long long)kc * (long long)sizeof(Cand);
  long long have = ugn_numel(workspace) * (workspace->dtype_bits / 8);
  UGN_CHECK(have >= need, "knn workspace too small: %lld < %lld", have, need);
  Cand* cands = ugn_ptr<Cand>(workspace);
  size_t smem = sizeof(Cand) * KNN_TQ * (kc + KNN_CB) + sizeof(float) * KNN_TK * (KNN_TQ + 4 + KNN_TG + 4) +
                sizeof(int) * KNN_TQ + sizeof(float) * KNN_TQ;
  UGN_CUDA(cudaFuncSetAttribute(knn_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ugn_cdiv(Q, KNN_TQ), chunks);
  knn_scan_kernel<<<grid, 256, smem, st>>>(ugn_ptr<float>(queries), ugn_ptr<float>(gallery), ugn_ptr<float>(g2),
                                           (int)Q, N, D, kc, rows, cands);
  UGN_LAUNCHED(ctx);
  knn_rerank_kernel<<<(int)Q, 256, 0, st>>>(ugn_ptr<float>(queries), ugn_ptr<float>(gallery),
                                            ugn_ptr<int>(gallery_labels), cands, chunks * kc, kc, k, D,
                                            idx_base, ugn_ptr<double>(out_d2), ugn_ptr<long long>(out_idx),
                                            ugn_ptr<int>(out_lab), nullptr, D, nullptr);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// ---------------------------------------------------------------------------------------
// Exact recomputation of flagged queries, PARALLEL over the shard: a handful of flagged queries is the normal case
// (measured: 1 of 4096 on one of 8 shards at D = 2048) and one CTA scanning 1 GB of gallery for it took 225 ms --
// 55x the whole search.  knn_flag_compact lists the first KX_SLOTS flagged queries; knn_exact_part gives every
// (flagged query, gallery part) pair its own CTA (KX_PARTS parts) and leaves a part-local top-k; knn_exact_merge
// merges the parts in (distance, index) order and marks the query done (flags = 2).  Anything beyond KX_SLOTS
// flagged queries falls through to the one-CTA-per-query kernel below, which is efficient when MANY are flagged.
// ---------------------------------------------------------------------------------------
static constexpr int KX_SLOTS = 128, KX_PARTS = 128;
__global__ void knn_flag_compact_kernel(const int* __restrict__ flags, int Q, int* __restrict__ list) {
  // list[0] = count (<= KX_SLOTS), list[1..] = query ids; single block, order-preserving
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  for (int q0 = 0; q0 < Q; q0 += blockDim.x) {
    const int q = q0 + threadIdx.x;
    const int f = (q < Q && flags[q] == 1) ? 1 : 0;
    if (f) {
      const int slot = atomicAdd(&cnt, 1);
      if (slot < KX_SLOTS) list[1 + slot] = q;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) list[0] = min(cnt, KX_SLOTS);
}
__global__ void __launch_bounds__(256) knn_exact_part_kernel(const float* __restrict__ Qm, const float* __restrict__ Gm,
                                                             const int* __restrict__ list, long long N, int D, int k,
                                                             double* __restrict__ pd, int* __restrict__ pi) {
  __shared__ double ld[8][KNN_MAXKC];
  __shared__ int li[8][KNN_MAXKC];
  const int t = threadIdx.x, w = t >> 5, lane = t & 31, part = blockIdx.x;
  const int count = list[0];
  const long long per = (N + KX_PARTS - 1) / KX_PARTS, r0 = part * per, r1 = min(N, r0 + per);
  for (int f = blockIdx.y; f < count; f += gridDim.y) {
    const int q = list[1 + f];
    if (lane == 0)
      for (int j = 0; j < k; ++j) { ld[w][j] = DBL_MAX; li[w][j] = 0x7fffffff; }
    __syncwarp();
    const float* qv = Qm + (long long)q * D;
    for (long long gi = r0 + w; gi < r1; gi += 8) {
      const float* gv = Gm + gi * D;
      double s = 0.0;
      for (int j = lane; j < D; j += 32) {
        double df = (double)qv[j] - (double)gv[j];
        s = fma(df, df, s);
      }
      s = warp_sum_d(s);
      if (lane == 0) {
        int ii = (int)gi;
        if (s < ld[w][k - 1] || (s == ld[w][k - 1] && ii < li[w][k - 1])) {
          int p = k - 1;
          while (p > 0 && (s < ld[w][p - 1] || (s == ld[w][p - 1] && ii < li[w][p - 1]))) {
            ld[w][p] = ld[w][p - 1]; li[w][p] = li[w][p - 1]; --p;
          }
          ld[w][p] = s; li[w][p] = ii;
        }
      }
      __syncwarp();
    }
    __syncthreads();
    if (t == 0) {                                   // 8 warp lists -> the part's list
      int head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int j = 0; j < k; ++j) {
        int bw = -1;
        for (int x = 0; x < 8; ++x) {
          if (head[x] >= k) continue;
          if (bw < 0 || ld[x][head[x]] < ld[bw][head[bw]] ||
              (ld[x][head[x]] == ld[bw][head[bw]] && li[x][head[x]] < li[bw][head[bw]])) bw = x;
        }
        const long long o = ((long long)f * KX_PARTS + part) * k + j;
        pd[o] = ld[bw][head[bw]];
        pi[o] = li[bw][head[bw]];
        head[bw]++;
      }
    }
    __syncthreads();
  }
}
__global__ void knn_exact_merge_kernel(const int* __restrict__ list, const double* __restrict__ pd,
                                       const int* __restrict__ pi, const int* __restrict__ labels, int k,
                                       long long idx_base, double* __restrict__ out_d2, long long* __restrict__ out_idx,
                                       int* __restrict__ out_lab, int* __restrict__ flags) {
  const int f = blockIdx.x;
  if (f >= list[0] || threadIdx.x != 0) return;
  const int q = list[1 + f];
  int head[KX_PARTS];
  for (int x = 0; x < KX_PARTS; ++x) head[x] = 0;
  for (int j = 0; j < k; ++j) {
    int bw = -1;
    double bd = 0.0;
    int bi = 0;
    for (int x = 0; x < KX_PARTS; ++x) {
      if (head[x] >= k) continue;
      const long long o = ((long long)f * KX_PARTS + x) * k + head[x];
      const double dd = pd[o];
      const int ii = pi[o];
      if (bw < 0 || dd < bd || (dd == bd && ii < bi)) { bw = x; bd = dd; bi = ii; }
    }
    out_d2[(long long)q * k + j] = bd;
    out_idx[(long long)q * k + j] = (bi == 0x7fffffff) ? -1 : idx_base + bi;
    out_lab[(long long)q * k + j] = (bi == 0x7fffffff) ? -1 : labels[bi];
    head[bw]++;
  }
  flags[q] = 2;                                       // flagged and recomputed
}
// one CTA per remaining flagged query (flags == 1: more than KX_SLOTS were flagged)
__global__ void __launch_bounds__(256) knn_exact_kernel(const float* __restrict__ Qm, const float* __restrict__ Gm,
                                                        const int* __restrict__ labels,
                                                        const int* __restrict__ flags, long long N, int D, int k,
                                                        long long idx_base, double* __restrict__ out_d2,
                                                        long long* __restrict__ out_idx,
                                                        int* __restrict__ out_lab) {
  __shared__ double ld[8][KNN_MAXKC];
  __shared__ int li[8][KNN_MAXKC];
  const int q = blockIdx.x, t = threadIdx.x, w = t >> 5, lane = t & 31;
  if (flags[q] != 1) return;
  if (lane == 0)
    for (int j = 0; j < k; ++j) { ld[w][j] = DBL_MAX; li[w][j] = 0x7fffffff; }
  __syncwarp();
  const float* qv = Qm + (long long)q * D;
  for (long long gi = w; gi < N; gi += 8) {
    const float* gv = Gm + gi * D;
    double s = 0.0;
    for (int j = lane; j < D; j += 32) {
      double df = (double)qv[j] - (double)gv[j];
      s = fma(df, df, s);
    }
    s = warp_sum_d(s);
    if (lane == 0) {
      int ii = (int)gi;
      if (s < ld[w][k - 1] || (s == ld[w][k - 1] && ii < li[w][k - 1])) {
        int p = k - 1;
        while (p > 0 && (s < ld[w][p - 1] || (s == ld[w][p - 1] && ii < li[w][p - 1]))) {
          ld[w][p] = ld[w][p - 1]; li[w][p] = li[w][p - 1]; --p;
        }
        ld[w][p] = s; li[w][p] = ii;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (t == 0) {
    int head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < k; ++j) {
      int bw = -1;
      for (int x = 0; x < 8; ++x) {
        if (head[x] >= k) continue;
        if (bw < 0 || ld[x][head[x]] < ld[bw][head[bw]] ||
            (ld[x][head[x]] == ld[bw][head[bw]] && li[x][head[x]] < li[bw][head[bw]])) bw = x;
      }
      int gi = li[bw][head[bw]];
      out_d2[(long long)q * k + j] = ld[bw][head[bw]];
      out_idx[(long long)q * k + j] = (gi == 0x7fffffff) ? -1 : idx_base + gi;
      out_lab[(long long)q * k + j] = (gi == 0x7fffffff) ? -1 : labels[gi];
      head[bw]++;
    }
  }
}

int knn_tc_scan(ugn_ctx* ctx, const __nv_bfloat16* q16, const __nv_bfloat16* g16, const float* g2, int Q,
                long long N, int KCH, int kc, int chunks, long long rows_per_chunk, void* cands, int* gthr,
                cudaStream_t st);
int knn_tc_pack(ugn_ctx* ctx, const float* X, long long rows, int D, __nv_bfloat16* out, cudaStream_t st);

// f32 [rows,D] -> the tensor-core scan's operand: fp16 hi/lo planes, K-chunk major [2, ceil(D/64), rows, 64]
extern "C" int ugn_knn_pack(ugn_ctx* ctx, const ugn_tensor* x, ugn_tensor* x16, void* stream) {
  UGN_CHECK(ctx && x && x16, "ugn_knn_pack: null argument");
  UGN_TENSOR(x, DT_F32, 2, 2);
  UGN_TENSOR(x16, DT_F16, 4, 4);
  long long rows = x->shape[0];
  int D = (int)x->shape[1], KCH = (D + 63) / 64;
  UGN_CHECK(x16->shape[0] == 2 && x16->shape[1] == KCH && x16->shape[2] == rows && x16->shape[3] == 64,
            "ugn_knn_pack: x16 must be f16 [2, ceil(D/64)=%d, rows=%lld, 64]", KCH, rows);
  return knn_tc_pack(ctx, ugn_ptr<float>(x), rows, D, ugn_ptr<__nv_bfloat16>(x16), (cudaStream_t)stream);
}

// Tensor-core variant of ugn_knn_topk: q16 / g16 are the fp16 hi/lo planes of queries / gallery made by
// ugn_knn_pack, g2 is padded to a multiple of 256 rows (+inf), gmax2 = max |g|^2.
extern "C" int ugn_knn_topk_tc(ugn_ctx* ctx, const ugn_tensor* queries, const ugn_tensor* q16,
                               const ugn_tensor* gallery, const ugn_tensor* g16, const ugn_tensor* g2,
                               const ugn_tensor* gmax2, const ugn_tensor* gallery_labels, int k,
                               int64_t idx_base, ugn_tensor* out_d2, ugn_tensor* out_idx, ugn_tensor* out_lab,
                               ugn_tensor* flags, ugn_tensor* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UGN_CHECK(ctx && queries && q16 && gallery && g16 && g2 && gmax2 && gallery_labels && out_d2 && out_idx &&
                out_lab && flags && workspace, "ugn_knn_topk_tc: null argument");
  UGN_TENSOR(queries, DT_F32, 2, 2);
  UGN_TENSOR(gallery, DT_F32, 2, 2);
  UGN_TENSOR(q16, DT_F16, 4, 4);
  UGN_TENSOR(g16, DT_F16, 4, 4);
  UGN_TENSOR(g2, DT_F32, 1, 1);
  UGN_TENSOR(gmax2, DT_F32, 1, 1);
  UGN_TENSOR(gallery_labels, DT_I32, 1, 1);
  UGN_TENSOR(out_d2, DT_F64, 2, 2);
  UGN_TENSOR(out_idx, DT_I64, 2, 2);
  UGN_TENSOR(out_lab, DT_I32, 2, 2);
  UGN_TENSOR(flags, DT_I32, 1, 1);
  UGN_TENSOR(workspace, DT_BAD, 1, 8);
  long long Q = queries->shape[0], N = gallery->shape[0];
  int D = (int)queries->shape[1], KCH = (D + 63) / 64, Dp = KCH * 64;
  UGN_CHECK(gallery->shape[1] == D, "gallery/query dimension mismatch");
  UGN_CHECK(q16->shape[0] == 2 && q16->shape[1] == KCH && q16->shape[2] == Q && q16->shape[3] == 64 &&
                g16->shape[0] == 2 && g16->shape[1] == KCH && g16->shape[2] == N && g16->shape[3] == 64,
            "q16/g16 must be f16 [2, ceil(D/64), rows, 64] (ugn_knn_pack)");
  UGN_CHECK(k >= 1 && k <= KNN_MAXKC && N >= k && N < 0x7fffffffLL, "k-NN: bad k / shard size");
  UGN_CHECK(g2->shape[0] >= (N + 255) / 256 * 256, "g2 must be padded to a multiple of 256 rows");
  UGN_CHECK(gallery_labels->shape[0] == N && flags->shape[0] >= Q, "labels [N], flags [Q] expected");
  UGN_CHECK(out_d2->shape[0] == Q && out_d2->shape[1] == k && out_idx->shape[0] == Q && out_idx->shape[1] == k &&
                out_lab->shape[0] == Q && out_lab->shape[1] == k, "outputs must be [Q,k]");
  if (Q == 0) return UGN_OK;
  int kc = knn_kc(k);
  // gallery chunks per query tile: one CTA per SM, so pick the chunk count that minimises
  // (waves of CTAs) x (rows per chunk) -- e.g. 32 query tiles x 37 chunks = 8 full waves of 148
  long long qt = (Q + 127) / 128;
  long long maxc = std::max<long long>(1, std::min<long long>(N / 2048, 512));
  long long best_cost = -1, rows = 0;
  int chunks = 1;
  long long cmin = 1;
  if (const char* e = getenv("UGN_KNN_CHUNKS")) { cmin = std::max(1LL, std::min<long long>(atoll(e), maxc)); maxc = cmin; }   // (experiments)
  for (long long c = cmin; c <= maxc; ++c) {
    long long r = ((N + c - 1) / c + 255) / 256 * 256;
    long long cc = (N + r - 1) / r;
    long long waves = (qt * cc + ctx->sm_count - 1) / ctx->sm_count;
    long long cost = waves * (r + 2048);          // + fixed per-CTA cost (prologue, list warm-up) in row units
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; chunks = (int)cc; rows = r; }
  }
  long long need = Q * chunks * 2 * (long long)kc * (long long)sizeof(Cand) + Q * 4;   // two column halves per chunk + thresholds
  long long have = ugn_numel(workspace) * (workspace->dtype_bits / 8);
  UGN_CHECK(have >= need, "knn workspace too small: %lld < %lld", have, need);
  Cand* cands = ugn_ptr<Cand>(workspace);
  int rc = knn_tc_scan(ctx, ugn_ptr<__nv_bfloat16>(q16), ugn_ptr<__nv_bfloat16>(g16), ugn_ptr<float>(g2), (int)Q, N,
                       KCH, kc, chunks, rows, cands,
                       reinterpret_cast<int*>(cands + Q * chunks * 2 * (long long)kc), st);
  if (rc != UGN_OK) return rc;
  knn_rerank_kernel<<<(int)Q, 256, 0, st>>>(ugn_ptr<float>(queries), ugn_ptr<float>(gallery),
                                            ugn_ptr<int>(gallery_labels), cands, chunks * 2 * kc, kc, k, D,
                                            idx_base, ugn_ptr<double>(out_d2), ugn_ptr<long long>(out_idx),
                                            ugn_ptr<int>(out_lab), ugn_ptr<float>(gmax2), Dp, ugn_ptr<int>(flags));
  UGN_LAUNCHED(ctx);
  {
    // flagged queries: parallel exact recomputation (see knn_exact_part_kernel); scratch = list + part lists
    const size_t nlist = (size_t)KX_SLOTS * KX_PARTS * KNN_MAXKC;
    void* scr = nullptr;
    if ((rc = ugn_scratch(ctx, 1024 + nlist * (sizeof(double) + sizeof(int)), &scr)) != UGN_OK) return rc;
    int* list = reinterpret_cast<int*>(scr);
    double* pd = reinterpret_cast<double*>(reinterpret_cast<char*>(scr) + 1024);
    int* pi = reinterpret_cast<int*>(pd + nlist);
    knn_flag_compact_kernel<<<1, 1024, 0, st>>>(ugn_ptr<int>(flags), (int)Q, list);
    UGN_LAUNCHED(ctx);
    knn_exact_part_kernel<<<dim3(KX_PARTS, 8), 256, 0, st>>>(ugn_ptr<float>(queries), ugn_ptr<float>(gallery), list, N,
                                                             D, k, pd, pi);
    UGN_LAUNCHED(ctx);
    knn_exact_merge_kernel<<<KX_SLOTS, 32, 0, st>>>(list, pd, pi, ugn_ptr<int>(gallery_labels), k, idx_base,
                                                    ugn_ptr<double>(out_d2), ugn_ptr<long long>(out_idx),
                                                    ugn_ptr<int>(out_lab), ugn_ptr<int>(flags));
    UGN_LAUNCHED(ctx);
  }
  knn_exact_kernel<<<(int)Q, 256, 0, st>>>(ugn_ptr<float>(queries), ugn_ptr<float>(gallery),
                                           ugn_ptr<int>(gallery_labels), ugn_ptr<int>(flags), N, D, k, idx_base,
                                           ugn_ptr<double>(out_d2), ugn_ptr<long long>(out_idx),
                                           ugn_ptr<int>(out_lab));
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

extern "C" int ugn_knn_merge_vote(ugn_ctx* ctx, const ugn_tensor* d2, const ugn_tensor* idx,
                                  const ugn_tensor* lab, int k, ugn_tensor* out_d2, ugn_tensor* out_idx,
                                  ugn_tensor* out_lab, ugn_tensor* pred, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UGN_CHECK(ctx && d2 && idx && lab && pred, "ugn_knn_merge_vote: null argument");
  UGN_TENSOR(d2, DT_F64, 3, 3);
  UGN_TENSOR(idx, DT_I64, 3, 3);
  UGN_TENSOR(lab, DT_I32, 3, 3);
  UGN_TENSOR(pred, DT_I32, 1, 1);
  int G = (int)d2->shape[0];
  long long Q = d2->shape[1];
  UGN_CHECK(G >= 1 && G <= 16, "1..16 shards supported");
  UGN_CHECK(d2->shape[2] == k && k <= KNN_MAXKC, "lists must be [G,Q,k]");
  UGN_CHECK(ugn_numel(idx) == ugn_numel(d2) && ugn_numel(lab) == ugn_numel(d2) && pred->shape[0] == Q,
            "merge_vote shape mismatch");
  if (out_d2) UGN_TENSOR(out_d2, DT_F64, 2, 2);
  if (out_idx) UGN_TENSOR(out_idx, DT_I64, 2, 2);
  if (out_lab) UGN_TENSOR(out_lab, DT_I32, 2, 2);
  if (Q == 0) return UGN_OK;
  knn_merge_vote_kernel<<<ugn_cdiv(Q, 128), 128, 0, st>>>(
      ugn_ptr<double>(d2), ugn_ptr<long long>(idx), ugn_ptr<int>(lab), Q * k, Q * k, Q * k, G, (int)Q, k,
      out_d2 ? ugn_ptr<double>(out_d2) : nullptr, out_idx ? ugn_ptr<long long>(out_idx) : nullptr,
      out_lab ? ugn_ptr<int>(out_lab) : nullptr, ugn_ptr<int>(pred));
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

// Sharded search: `packed` u8 [G, 20*Q*k (+pad)] is the all-gathered buffer whose row g holds rank g's lists back to
// back: d2 f64 [Q,k] | idx i64 [Q,k] | lab i32 [Q,k] -- ONE collective instead of three.
extern "C" int ugn_knn_merge_vote_packed(ugn_ctx* ctx, const ugn_tensor* packed, long long Q, int k, ugn_tensor* out_d2,
                                         ugn_tensor* out_idx, ugn_tensor* out_lab, ugn_tensor* pred, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UGN_CHECK(ctx && packed && pred, "ugn_knn_merge_vote_packed: null argument");
  UGN_TENSOR(packed, DT_U8, 2, 2);
  UGN_TENSOR(pred, DT_I32, 1, 1);
  const int G = (int)packed->shape[0];
  const long long row = packed->shape[1], n = Q * k;
  UGN_CHECK(G >= 1 && G <= 16 && k >= 1 && k <= KNN_MAXKC, "1..16 shards, k <= %d", KNN_MAXKC);
  UGN_CHECK(row >= 20 * n && row % 8 == 0 && pred->shape[0] == Q, "packed rows must hold 20*Q*k bytes and be 8-byte multiples");
  if (out_d2) UGN_TENSOR(out_d2, DT_F64, 2, 2);
  if (out_idx) UGN_TENSOR(out_idx, DT_I64, 2, 2);
  if (out_lab) UGN_TENSOR(out_lab, DT_I32, 2, 2);
  if (Q == 0) return UGN_OK;
  const uint8_t* base = ugn_ptr<uint8_t>(packed);
  knn_merge_vote_kernel<<<ugn_cdiv(Q, 128), 128, 0, st>>>(
      reinterpret_cast<const double*>(base), reinterpret_cast<const long long*>(base + 8 * n),
      reinterpret_cast<const int*>(base + 16 * n), row / 8, row / 8, row / 4, G, (int)Q, k,
      out_d2 ? ugn_ptr<double>(out_d2) : nullptr, out_idx ? ugn_ptr<long long>(out_idx) : nullptr,
      out_lab ? ugn_ptr<int>(out_lab) : nullptr, ugn_ptr<int>(pred));
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}
