// Batch-all triplet loss, forward + analytic backward.
// Replaces triplet_loss(margin)/batch_dist of the reference (nets/triplet_loss_all.py:8-77).
//
//   dist[n,a,b] = sqrt(max(|x_a|^2 + |x_b|^2 - 2 x_a.x_b, 0)), entries <= 0 forced to 0 with
//                 zero gradient (:72-76)
//   L_n = sum_{a, p: lab_p==lab_a (p==a included), q: lab_q!=lab_a} max(margin + (d_ap - d_aq), 0)
//         / #{terms > 0}          (0 when no term is active, :55-59);   loss = mean_n L_n (:61)
//
// Mask semantics are implemented directly (any label multiset); on the balanced P x K batches the
// reference's boolean_mask+reshape requires, both agree.
// Backward (the active count is a constant under autodiff):
//   dL/dd_ab = (+#{q active} if same(a,b) else -#{p active}) / (n_parts * c_n)
//   dd_ab/dx_a = (x_a - x_b)/d_ab,  dd_ab/dx_b = -(x_a - x_b)/d_ab, both 0 where d_ab == 0.
#include <cfloat>
#include "simt.cuh"

struct TripAcc {
  double sum;
  unsigned long long cnt;
  unsigned long long act;   // batch-hard loss: anchors with a positive term (cnt is the batch size there)
};

__global__ void trip_diag_kernel(const float* __restrict__ G, float* __restrict__ x2, int B) {
  int n = blockIdx.y;
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < B) x2[(long long)n * B + a] = G[((long long)n * B + a) * B + a];
}

__global__ void trip_dist_kernel(float* __restrict__ G, const float* __restrict__ x2, int B) {
  int n = blockIdx.z, a = blockIdx.y;
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long long o = ((long long)n * B + a) * B + b;
  float d2 = x2[(long long)n * B + a] + x2[(long long)n * B + b] - 2.0f * G[o];
  d2 = fmaxf(d2, 0.f);
  G[o] = d2 > 0.f ? sqrtf(d2) : 0.f;
}

__global__ void __launch_bounds__(128) trip_hinge_kernel(const float* __restrict__ D,
                                                         const int* __restrict__ labels,
                                                         float* __restrict__ Wc, TripAcc* __restrict__ acc,
                                                         int B, float margin) {
  extern __shared__ float sm[];  // drow[B] | lab[B]
  float* drow = sm;
  int* lab = reinterpret_cast<int*>(sm + B);
  const int a = blockIdx.x, n = blockIdx.y;
  const float* Da = D + ((long long)n * B + a) * B;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    drow[b] = Da[b];
    lab[b] = labels[b];
  }
  __syncthreads();
  const int la = lab[a];
  double sum = 0.0;
  unsigned long long cnt = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float db = drow[b];
    int c_act = 0;
    if (lab[b] == la) {
      float s = 0.f;
      for (int q = 0; q < B; ++q) {
        if (lab[q] != la) {
          float t = margin + (db - drow[q]);
          if (t > 0.f) { s += t; ++c_act; }
        }
      }
      sum += (double)s;
      cnt += (unsigned long long)c_act;
      Wc[((long long)n * B + a) * B + b] = (float)c_act;
    } else {
      for (int p = 0; p < B; ++p) {
        if (lab[p] == la) {
          float t = margin + (drow[p] - db);
          if (t > 0.f) ++c_act;
        }
      }
      Wc[((long long)n * B + a) * B + b] = -(float)c_act;
    }
  }
  // block reduction (4 warps)
  __shared__ double rs[4];
  __shared__ unsigned long long rc[4];
  sum = warp_sum_d(sum);
  unsigned int lo = (unsigned int)cnt;  // per-thread counts fit 32 bits (<= B*B/128)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lo += __shfl_xor_sync(0xffffffffu, lo, o);
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = sum; rc[threadIdx.x >> 5] = lo; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = rs[0] + rs[1] + rs[2] + rs[3];
    unsigned long long c = rc[0] + rc[1] + rc[2] + rc[3];
    if (c) {
      atomicAdd(&acc[n].sum, s);
      atomicAdd(&acc[n].cnt, c);
    }
  }
}

// Batch-HARD triplet loss: tfa.losses.TripletHardLoss(margin) as compiled by UWYHSemiNet3Mods.compile_hard
// (nets/mj_uwyhNets_ba.py:1302-1306; soft = False, distance_metric = "L2"; tensorflow_addons is not vendored, the
// published algorithm of tfa/losses/triplet.py is followed):
//   hard_p(a) = masked_maximum(pdist, same label minus the diagonal) = the farthest positive (0 without positives)
//   hard_n(a) = masked_minimum(pdist, other label) = min_j((d_aj - M_a) * mask) + M_a, M_a = max_j d_aj: the nearest
//               negative; M_a itself when the anchor has no negative
//   loss = mean_a max(hard_p - hard_n + margin, 0)
// One block per anchor.  Wc receives dLoss/dd_ab * B (trip_bwd_kernel divides by acc.cnt = B); reduce_max / reduce_min
// split the gradient evenly among tied entries, and the masked_minimum form also routes gradient through M_a when the
// selected entry ties with the masked zeros -- both are kept literally.
__global__ void __launch_bounds__(128) trip_hard_kernel(const float* __restrict__ D, const int* __restrict__ labels,
                                                        float* __restrict__ Wc, TripAcc* __restrict__ acc, int B,
                                                        float margin) {
  extern __shared__ float sm[];  // drow[B] | lab[B]
  float* drow = sm;
  int* lab = reinterpret_cast<int*>(sm + B);
  __shared__ float r_hp[4], r_hn[4], r_M[4];
  __shared__ int r_c[4][4];
  const int a = blockIdx.x;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    drow[b] = b == a ? 0.f : D[(long long)a * B + b];      // pairwise_distance zeroes the diagonal
    lab[b] = labels[b];
  }
  __syncthreads();
  const int la = lab[a];
  float hp = -1.f, hn = FLT_MAX, M = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float db = drow[b];
    M = fmaxf(M, db);
    if (b == a) continue;
    if (lab[b] == la) hp = fmaxf(hp, db); else hn = fminf(hn, db);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    hp = fmaxf(hp, __shfl_xor_sync(0xffffffffu, hp, o));
    hn = fminf(hn, __shfl_xor_sync(0xffffffffu, hn, o));
    M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  }
  if ((threadIdx.x & 31) == 0) { r_hp[threadIdx.x >> 5] = hp; r_hn[threadIdx.x >> 5] = hn; r_M[threadIdx.x >> 5] = M; }
  __syncthreads();
  hp = fmaxf(fmaxf(r_hp[0], r_hp[1]), fmaxf(r_hp[2], r_hp[3]));
  hn = fminf(fminf(r_hn[0], r_hn[1]), fminf(r_hn[2], r_hn[3]));
  M = fmaxf(fmaxf(r_M[0], r_M[1]), fmaxf(r_M[2], r_M[3]));
  const bool has_pos = hp >= 0.f, has_neg = hn < FLT_MAX;
  // tie counts: positives at hp, negatives at hn, entries at M, masked (non-negative) entries
  int np = 0, nn = 0, nM = 0, nmask = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float db = drow[b];
    const bool same = lab[b] == la;
    nM += db == M;
    nmask += same;                       // the diagonal is a masked entry of adjacency_not too
    if (b == a) continue;
    np += same && db == hp;
    nn += !same && db == hn;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    np += __shfl_xor_sync(0xffffffffu, np, o);
    nn += __shfl_xor_sync(0xffffffffu, nn, o);
    nM += __shfl_xor_sync(0xffffffffu, nM, o);
    nmask += __shfl_xor_sync(0xffffffffu, nmask, o);
  }
  if ((threadIdx.x & 31) == 0) { int w = threadIdx.x >> 5; r_c[w][0] = np; r_c[w][1] = nn; r_c[w][2] = nM; r_c[w][3] = nmask; }
  __syncthreads();
  np = r_c[0][0] + r_c[1][0] + r_c[2][0] + r_c[3][0];
  nn = r_c[0][1] + r_c[1][1] + r_c[2][1] + r_c[3][1];
  nM = r_c[0][2] + r_c[1][2] + r_c[2][2] + r_c[3][2];
  nmask = r_c[0][3] + r_c[1][3] + r_c[2][3] + r_c[3][3];
  const float hp_v = (has_pos && hp > 0.f) ? hp : 0.f;            // max over (d * mask): never below the masked zeros
  const float hn_v = has_neg ? hn : M;
  const float t = hp_v - hn_v + margin;
  const bool active = t > 0.f;
  // masked_minimum: the selected set is {negatives at hn} when hn < M, else those AND every masked entry (all tie at 0)
  const bool tie0 = !has_neg || hn == M;
  const float sel = tie0 ? (float)((has_neg ? nn : 0) + nmask) : (float)nn;
  const float w_neg = (has_neg && active) ? 1.f / sel : 0.f;                                  // direct part
  const float w_M = (active && tie0) ? (1.f - (has_neg ? (float)nn : 0.f) / sel) / (float)nM : 0.f;   // through M_a
  const float w_pos = (active && hp_v > 0.f) ? 1.f / (float)np : 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float db = drow[b];
    float w = 0.f;
    if (b != a) {
      const bool same = lab[b] == la;
      if (same && db == hp) w += w_pos;
      if (!same && db == hn) w -= w_neg;
      if (db == M) w -= w_M;
    }
    Wc[(long long)a * B + b] = w;
  }
  if (threadIdx.x == 0) {
    if (active) { atomicAdd(&acc[0].sum, (double)t); atomicAdd(&acc[0].act, 1ull); }
    atomicAdd(&acc[0].cnt, 1ull);
  }
}

__global__ void trip_finalize_kernel(const TripAcc* __restrict__ acc, int nparts, float* __restrict__ out, int hard) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double loss = 0.0, total = 0.0;
    for (int n = 0; n < nparts; ++n) {
      if (acc[n].cnt) loss += (double)((float)acc[n].sum / (float)acc[n].cnt);
      total += (double)(hard ? acc[n].act : acc[n].cnt);
    }
    out[0] = (float)(loss / nparts);
    out[1] = (float)total;
  }
}

__global__ void __launch_bounds__(256) trip_bwd_kernel(const float* __restrict__ emb,
                                                       const float* __restrict__ D,
                                                       const float* __restrict__ Wc,
                                                       const TripAcc* __restrict__ acc,
                                                       float* __restrict__ demb, int nparts, int B, int d,
                                                       float scale) {
  extern __shared__ float coef[];  // [B]
  const int a = blockIdx.x, n = blockIdx.y;
  const long long base = (long long)n * B * B;
  unsigned long long c = acc[n].cnt;
  float s = c ? scale / ((float)nparts * (float)c) : 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float dab = D[base + (long long)a * B + b], dba = D[base + (long long)b * B + a];
    float v = 0.f;
    if (dab > 0.f) v += Wc[base + (long long)a * B + b] / dab;
    if (dba > 0.f) v += Wc[base + (long long)b * B + a] / dba;
    coef[b] = v;
  }
  __syncthreads();
  const float* X = emb + (long long)n * B * d;
  // blockIdx.z slices the feature axis (256 features per block) so that B x d/256 blocks fill the SMs;
  // four independent accumulators hide the L2 latency of the row sweep
  const int j = blockIdx.z * blockDim.x + threadIdx.x;
  if (j >= d) return;
  const float xa = X[(long long)a * d + j];
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
  int b = 0;
  for (; b + 4 <= B; b += 4) {
    g0 = fmaf(coef[b], xa - X[(long long)b * d + j], g0);
    g1 = fmaf(coef[b + 1], xa - X[(long long)(b + 1) * d + j], g1);
    g2 = fmaf(coef[b + 2], xa - X[(long long)(b + 2) * d + j], g2);
    g3 = fmaf(coef[b + 3], xa - X[(long long)(b + 3) * d + j], g3);
  }
  for (; b < B; ++b) g0 = fmaf(coef[b], xa - X[(long long)b * d + j], g0);
  demb[((long long)n * B + a) * d + j] = s * ((g0 + g1) + (g2 + g3));
}

// ---------------------------------------------------------------------------------------------------------------
// Pair verification loss of the Siamese builder UWYHNet.build (nets/mj_uwyhNets_ba.py:154-245): VerifLossLayer
// (nets/mj_loss.py:65-95) on the two normalised signatures a = emb[:B], b = emb[B:] and the pair labels:
//   loss = 0.5 * sum_{label == 1} (a - b)^2  +  0.5 * max(0, m - sqrt(sum_{label == 0} (a - b)^2))^2
// (the negative term takes the square root of the sum over ALL negative rows, as the reference does).
// acc[0] = positive sum, acc[1] = negative sum S.  Backward: d/da = (a - b) on positive rows, -(m - sqrt S)/sqrt S (a - b)
// on negative rows while m > sqrt S > 0; d/db = -d/da.
__global__ void __launch_bounds__(256) pair_sums_kernel(const float* __restrict__ emb, const int* __restrict__ labels,
                                                        double* __restrict__ acc, int B, int d) {
  const int r = blockIdx.x;
  const int lab = labels[r];
  if (lab != 0 && lab != 1) return;
  const float* a = emb + (long long)r * d;
  const float* b = emb + (long long)(r + B) * d;
  double s = 0.0;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float df = a[j] - b[j];
    s += (double)(df * df);
  }
  __shared__ double red[8];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(&acc[lab == 1 ? 0 : 1], t);
  }
}

__global__ void __launch_bounds__(256) pair_finish_kernel(const float* __restrict__ emb, const int* __restrict__ labels,
                                                          const double* __restrict__ acc, float* __restrict__ out,
                                                          float* __restrict__ demb, int B, int d, float margin,
                                                          float scale) {
  const float S = (float)acc[1];
  const float rs = sqrtf(S);
  const float gap = fmaxf(margin - rs, 0.f);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out[0] = 0.5f * (float)acc[0] + 0.5f * gap * gap;
    out[1] = gap > 0.f ? 1.f : 0.f;
  }
  if (!demb) return;
  const int r = blockIdx.x;
  const int lab = labels[r];
  float c = 0.f;
  if (lab == 1) c = 1.f;
  else if (lab == 0 && gap > 0.f && rs > 0.f) c = -gap / rs;
  const float* a = emb + (long long)r * d;
  const float* b = emb + (long long)(r + B) * d;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float g = scale * c * (a[j] - b[j]);
    demb[(long long)r * d + j] = g;
    demb[(long long)(r + B) * d + j] = -g;
  }
}

// emb f32 [2B,d] (rows [0,B) = first element of every pair, [B,2B) = second), labels i32 [>= B] (1 same / 0 different;
// anything else: the pair is ignored), out f32 [2] = {loss, 1 if the negative hinge is active}, demb [2B,d] nullable,
// workspace >= 16 bytes
extern "C" int ugn_pair_verif_loss(ugn_ctx* ctx, const ugn_tensor* emb, const ugn_tensor* labels, float margin, float scale,
                                   ugn_tensor* out, ugn_tensor* demb, ugn_tensor* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UGN_CHECK(ctx && emb && labels && out && workspace, "ugn_pair_verif_loss: null argument");
  UGN_TENSOR(emb, DT_F32, 2, 2);
  UGN_TENSOR(labels, DT_I32, 1, 2);
  UGN_TENSOR(out, DT_F32, 1, 1);
  UGN_TENSOR(workspace, DT_BAD, 1, 8);
  UGN_CHECK(emb->shape[0] % 2 == 0, "ugn_pair_verif_loss: emb must hold 2B rows");
  const int B = (int)(emb->shape[0] / 2), d = (int)emb->shape[1];
  UGN_CHECK(ugn_numel(labels) >= B && out->shape[0] >= 2, "ugn_pair_verif_loss: labels [>=B], out [2] expected");
  UGN_CHECK(ugn_numel(workspace) * (workspace->dtype_bits / 8) >= 16, "ugn_pair_verif_loss: workspace too small");
  if (demb) { UGN_TENSOR(demb, DT_F32, 2, 2); UGN_CHECK(ugn_numel(demb) == ugn_numel(emb), "demb shape mismatch"); }
  double* acc = ugn_ptr<double>(workspace);
  UGN_CUDA(cudaMemsetAsync(acc, 0, 16, st));
  if (B > 0) {
    pair_sums_kernel<<<B, 256, 0, st>>>(ugn_ptr<float>(emb), ugn_ptr<int>(labels), acc, B, d);
    UGN_LAUNCHED(ctx);
  }
  pair_finish_kernel<<<std::max(B, 1), 256, 0, st>>>(ugn_ptr<float>(emb), ugn_ptr<int>(labels), acc, ugn_ptr<float>(out),
                                                     (demb && B > 0) ? ugn_ptr<float>(demb) : nullptr, B, d, margin, scale);
  UGN_LAUNCHED(ctx);
  return UGN_OK;
}

extern "C" int64_t ugn_triplet_workspace_bytes(int n, int B) {
  int64_t accb = ((int64_t)n * sizeof(TripAcc) + 255) / 256 * 256;
  int64_t x2b = ((int64_t)n * B * 4 + 255) / 256 * 256;
  return accb + x2b + 2 * (int64_t)n * B * B * 4;
}

int tc_gram(ugn_ctx* ctx, int f16, int B, int d, const __nv_bfloat16* X16, float* G, cudaStream_t st);

static int triplet_common(ugn_ctx* ctx, const ugn_tensor* emb, const ugn_tensor* emb16, const ugn_tensor* labels,
                          float margin, float scale, ugn_tensor* out, ugn_tensor* demb, ugn_tensor* workspace,
                          void* stream, int hard = 0);
// tfa.losses.TripletHardLoss (compile_hard): emb [B,d] (tfa takes rank-2 embeddings), emb16 nullable (tensor-core Gram)
extern "C" int ugn_triplet_hard(ugn_ctx* ctx, const ugn_tensor* emb, const ugn_tensor* emb16, const ugn_tensor* labels,
                                float margin, float scale, ugn_tensor* out, ugn_tensor* demb, ugn_tensor* workspace,
                                void* stream) {
  UGN_CHECK(emb && emb->ndim == 2, "ugn_triplet_hard: embeddings must be [B,d]");
  return triplet_common(ctx, emb, emb16, labels, margin, scale, out, demb, workspace, stream, 1);
}
extern "C" int ugn_triplet_all(ugn_ctx* ctx, const ugn_tensor* emb, const ugn_tensor* labels,
                               float margin, float scale, ugn_tensor* out, ugn_tensor* demb,
                               ugn_tensor* workspace, void* stream) {
  return triplet_common(ctx, emb, nullptr, labels, margin, scale, out, demb, workspace, stream);
}
// The same loss with the B x B Gram matrix on the tensor cores: emb16 = fp16 | bf16 hi/lo planes [2,B,d] of emb (written by
// ugn_fuse_fwd next to the f32 signature), one tcgen05 GEMM with three passes (hi*hi + hi*lo + lo*hi, ~fp32 products,
// fp32 accumulation in TMEM) and NO split-K, so that every element has ONE accumulation order: the diagonal
// the squared norms are read from is consistent with the matrix and identical rows give an exactly zero distance, as in
// the FFMA path.
extern "C" int ugn_triplet_all_tc(ugn_ctx* ctx, const ugn_tensor* emb, const ugn_tensor* emb16, const ugn_tensor* labels,
                                  float margin, float scale, ugn_tensor* out, ugn_tensor* demb, ugn_tensor* workspace,
                                  void* stream) {
  UGN_CHECK(emb16, "ugn_triplet_all_tc: emb16 required");
  return triplet_common(ctx, emb, emb16, labels, margin, scale, out, demb, workspace, stream);
}

static int triplet_common(ugn_ctx* ctx, const ugn_tensor* emb, const ugn_tensor* emb16, const ugn_tensor* labels,
                          float margin, float scale, ugn_tensor* out, ugn_tensor* demb, ugn_tensor* workspace,
                          void* stream, int hard) {
  cudaStream_t st = (cudaStream_t)stream;
  UGN_CHECK(ctx && emb && labels && out && workspace, "ugn_triplet_all: null argument");
  UGN_TENSOR(emb, DT_F32, 2, 3);
  UGN_TENSOR(labels, DT_I32, 1, 2);
  UGN_TENSOR(out, DT_F32, 1, 1);
  UGN_TENSOR(workspace, DT_BAD, 1, 8);
  int n = emb->ndim == 3 ? (int)emb->shape[0] : 1;
  int B = (int)emb->shape[emb->ndim - 2], d = (int)emb->shape[emb->ndim - 1];
  UGN_CHECK(ugn_numel(labels) == B, "labels must have B=%d entries", B);
  UGN_CHECK(out->shape[0] >= 2, "out must be f32[2]");
  UGN_CHECK(B <= 8192, "triplet batch too large");
  int64_t need = ugn_triplet_workspace_bytes(n, B);
  int64_t have = ugn_numel(workspace) * (workspace->dtype_bits / 8);
  UGN_CHECK(have >= need, "triplet workspace too small: %lld < %lld", (long long)have, (long long)need);
  if (demb) {
    UGN_TENSOR(demb, DT_F32, 2, 3);
    UGN_CHECK(ugn_numel(demb) == ugn_numel(emb), "demb shape mismatch");
  }
  char* ws = ugn_ptr<char>(workspace);
  int64_t accb = ((int64_t)n * sizeof(TripAcc) + 255) / 256 * 256;
  int64_t x2b = ((int64_t)n * B * 4 + 255) / 256 * 256;
  TripAcc* acc = reinterpret_cast<TripAcc*>(ws);
  float* x2 = reinterpret_cast<float*>(ws + accb);
  float* D = reinterpret_cast<float*>(ws + accb + x2b);
  float* Wc = D + (int64_t)n * B * B;
  const float* X = ugn_ptr<float>(emb);
  UGN_CUDA(cudaMemsetAsync(acc, 0, n * sizeof(TripAcc), st));

  SGemm p;  // Gram: G[n] = X[n] X[n]^T  (fp32 FFMA; fp32-accurate Gram is required)
  p.A = X; p.B = X; p.C = D;
  p.M = B; p.N = B; p.K = d;
  p.ar = radix1(d); p.ak = radix1(1); p.br = radix1(d); p.bk = radix1(1);
  p.ldc = B;
  p.batches = n; p.batch_a = (long long)B * d; p.batch_b = (long long)B * d; p.batch_c = (long long)B * B;
  if (n == 1 && B <= 256 && d >= 512) {
    // small batch: only (B/64)^2 output tiles -> spread the feature dimension over the SMs (split-K + red.add)
    p.batches = 1;
    p.ksplit = std::min(32, d / 64);
    p.epi = EPI_ATOMIC;
    UGN_CUDA(cudaMemsetAsync(D, 0, sizeof(float) * (size_t)B * B, st));
  }
  int rc;
  if (emb16) {
    UGN_TENSOR(emb16, DT_BAD, 3, 3);
    UGN_CHECK(n == 1 && emb16->shape[0] == 2 && emb16->shape[1] == B && emb16->shape[2] == d && d % 64 == 0,
              "triplet_all_tc: emb16 must be 16-bit [2,B,d] of a single-part embedding with d %% 64 == 0");
    UGN_CHECK(ugn_dtype(emb16) == DT_F16 || ugn_dtype(emb16) == DT_BF16, "triplet_all_tc: emb16 must be f16 or bf16");
    rc = tc_gram(ctx, ugn_dtype(emb16) == DT_F16, B, d, ugn_ptr<__nv_bfloat16>(emb16), D, st);
  } else {
    rc = simt_gemm_launch(ctx, p, st);
  }
  if (rc != UGN_OK) return rc;
  trip_diag_kernel<<<dim3(ugn_cdiv(B, 128), n), 128, 0, st>>>(D, x2, B);
  UGN_LAUNCHED(ctx);
  trip_dist_kernel<<<dim3(ugn_cdiv(B, 128), B, n), 128, 0, st>>>(D, x2, B);
  UGN_LAUNCHED(ctx);
  if (2 * (size_t)B * sizeof(float) > 48 * 1024) {      // B > 6144: opt in to the large dynamic shared memory carve-out
    static bool attr_set = false;
    if (!attr_set) {
      UGN_CUDA(cudaFuncSetAttribute(trip_hinge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * (int)sizeof(float)));
      UGN_CUDA(cudaFuncSetAttribute(trip_hard_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * (int)sizeof(float)));
      attr_set = true;
    }
  }
  if (hard)
    trip_hard_kernel<<<B, 128, 2 * B * sizeof(float), st>>>(D, ugn_ptr<int>(labels), Wc, acc, B, margin);
  else
    trip_hinge_kernel<<<dim3(B, n), 128, 2 * B * sizeof(float), st>>>(D, ugn_ptr<int>(labels), Wc, acc, B, margin);
  UGN_LAUNCHED(ctx);
  trip_finalize_kernel<<<1, 32, 0, st>>>(acc, n, ugn_ptr<float>(out), hard);
  UGN_LAUNCHED(ctx);
  if (demb) {
    trip_bwd_kernel<<<dim3(B, n, ugn_cdiv(d, 256)), 256, B * sizeof(float), st>>>(X, D, Wc, acc, ugn_ptr<float>(demb), n, B, d, scale);
    UGN_LAUNCHED(ctx);
  }
  return UGN_OK;
}
