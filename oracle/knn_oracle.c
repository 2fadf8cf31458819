/* CPU oracle (TEST INFRASTRUCTURE ONLY) for the open-world k-NN step.
 *
 * Plain-C restatement of what the reference executes at
 *   /root/reference/mains/mj_testUWYHGaitNet_open_tum.py:331-341
 *     clf = KNeighborsClassifier(n_neighbors=knn); clf.fit(G, y); clf.predict(Q)
 * i.e. brute-force Euclidean k nearest neighbours with a uniform vote
 * (vote ties -> smallest label).  Distances are exact fp64 sum((q-g)^2) of the fp32
 * inputs; neighbour order is (distance, then lower gallery index).  scikit-learn's own
 * order among EXACTLY tied distances is unspecified (heap artefact), so index parity is
 * asserted on tie-free queries and label parity on all of them (see tests/test_oracle.py).
 *
 * Pinned against scikit-learn 1.9.0 run in the build container (tests/golden/knn_*.npz).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* gallery [N,D] f32 row-major, queries [Q,D] f32, out_idx [Q,k] i64, out_d2 [Q,k] f64 */
int knn_oracle_search(const float *gallery, int64_t N, int64_t D, const float *queries,
                      int64_t Q, int k, int64_t *out_idx, double *out_d2)
{
    if (k <= 0 || k > N || k > 64) return -1;
    for (int64_t q = 0; q < Q; ++q) {
        double bd[64];
        int64_t bi[64];
        int have = 0;
        const float *qv = queries + q * D;
        for (int64_t g = 0; g < N; ++g) {
            const float *gv = gallery + g * D;
            double s = 0.0;
            for (int64_t j = 0; j < D; ++j) {
                double t = (double)qv[j] - (double)gv[j];
                s += t * t;
            }
            /* insert (s, g) keeping ascending (dist, idx); g increases so ties keep order */
            if (have < k || s < bd[have - 1]) {
                int p = have < k ? have : k - 1;
                while (p > 0 && bd[p - 1] > s) {
                    bd[p] = bd[p - 1];
                    bi[p] = bi[p - 1];
                    --p;
                }
                bd[p] = s;
                bi[p] = g;
                if (have < k) ++have;
            }
        }
        for (int j = 0; j < k; ++j) {
            out_idx[q * k + j] = bi[j];
            out_d2[q * k + j] = bd[j];
        }
    }
    return 0;
}

/* uniform vote over neighbour labels, ties -> smallest label */
int knn_oracle_vote(const int64_t *idx, int64_t Q, int k, const int32_t *labels, int32_t *out)
{
    for (int64_t q = 0; q < Q; ++q) {
        int32_t best = 0;
        int bestc = -1;
        for (int a = 0; a < k; ++a) {
            int32_t la = labels[idx[q * k + a]];
            int c = 0;
            for (int b = 0; b < k; ++b) c += labels[idx[q * k + b]] == la;
            if (c > bestc || (c == bestc && la < best)) {
                bestc = c;
                best = la;
            }
        }
        out[q] = best;
    }
    return 0;
}
