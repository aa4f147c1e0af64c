/*
 * oracle.c -- CPU restatement of the reference's descriptor-space kNN
 * correspondence search (aleksandrina-streltsova/lidar-global-registration).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it; the product path (libb200match.so) never links or calls it.
 *
 * Pin status (see DESIGN.md "Oracle"):
 *   - KNNResult / top-k container: PINNED by the reference's own known-answer
 *     vectors (tests/knn_result.cpp:28-49), replayed in tests/test_oracle.py.
 *   - kNN indices: PINNED against OpenCV cv2.BFMatcher(NORM_L2).knnMatch (the
 *     same third-party arithmetic matchBF calls, include/matching.h:600,612)
 *     via committed fixtures tests/golden/ (generator: tests/golden/make_golden.py).
 *   - distances: the reference pins none (isclose() in tests/flann_bf_matcher.h:12
 *     is never called); we pin to cv2 within 1e-6 relative.
 *   - ratio filter: PARITY UNPINNED -- RatioMatcher::match_impl is a stub
 *     (include/matching.h:470-473); semantics defined in SURVEY.md 8a.
 *
 * The reference itself cannot be compiled here: include/matching.h:8-16 pulls
 * OpenCV features2d + PCL (kdtree, search, transforms), none of which is in
 * this image.  The arithmetic restated below is the one the reference code
 * spells out itself (pcl::L2_Norm loop as used at include/matching.h:663) and
 * the published behaviour of the un-vendored third parties at its call sites:
 *   OpenCV 4.5.1 cv::BFMatcher / batchDistance (call sites :600,:612)
 *   PCL 1.12.1 pcl::KdTreeFLANN over FLANN 1.9.1 L2_Simple (call sites :567-585)
 *
 * Distance arithmetic (all paths): FP32, sequential over dimensions,
 *      s = 0; for d: diff = a[d]-b[d]; s = s + diff*diff;   (no FMA contraction;
 *      the reference's CMakeLists.txt sets no -march, so x86-64 baseline)
 *      dist = sqrtf(s)
 * which is literally pcl::L2_Norm (matchLocal, :663) and FLANN's L2_Simple
 * (matchFLANN).  OpenCV's SIMD normL2Sqr_ sums in a different lane order, so
 * matchBF distances may differ from this in the last bits; the reference's own
 * test (tests/flann_bf_matcher.h:73-88) asserts index equality only.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* 16-byte record == reference `Correspondence` (include/common.h:120-131):
 * pcl::Correspondence{int index_query; int index_match; float distance} + float threshold */
typedef struct {
    int32_t index_query;
    int32_t index_match;
    float distance;
    float threshold;
} orc_corr;

static inline const float *row_at(const float *base, size_t stride_bytes, size_t i) {
    return (const float *) ((const char *) base + i * stride_bytes);
}

/* pcl::PointRepresentation::isValid -- every one of the D values must be finite
 * (used at include/matching.h:576, :655, :661). */
ORC_API int orc_is_valid(const float *row, int dim) {
    for (int d = 0; d < dim; ++d)
        if (!isfinite(row[d])) return 0;
    return 1;
}

/* pcl::L2_Norm(a, b, dim) as called at include/matching.h:663-664. */
ORC_API float orc_l2_norm(const float *a, const float *b, int dim) {
    float s = 0.f;
    for (int d = 0; d < dim; ++d) {
        float diff = a[d] - b[d];
        s = s + diff * diff;
    }
    return sqrtf(s);
}

/* ORC_UNROLL independent pcl::L2_Norm chains side by side: each pair keeps the
 * exact sequential per-pair arithmetic above (bit-identical results); the
 * interleaving only hides FP-add latency so the timed CPU baseline is not
 * artificially slow. */
#define ORC_UNROLL 8
static inline void l2_norm_block(const float *q, const float *train, size_t t_stride, size_t j0,
                                 int dim, float *out) {
    const float *t[ORC_UNROLL];
    float s[ORC_UNROLL];
    for (int u = 0; u < ORC_UNROLL; ++u) { t[u] = row_at(train, t_stride, j0 + u); s[u] = 0.f; }
    for (int d = 0; d < dim; ++d) {
        float qd = q[d];
        for (int u = 0; u < ORC_UNROLL; ++u) {
            float diff = qd - t[u][d];
            s[u] = s[u] + diff * diff;
        }
    }
    for (int u = 0; u < ORC_UNROLL; ++u) out[u] = sqrtf(s[u]);
}

/* torchrun exports OMP_NUM_THREADS=1; the timed CPU baseline asks for all host cores explicitly. */
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void) n;
#endif
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- KNNResult<float> (include/matching.h:44-94) -------------------------
 * Bounded ascending list; a new point goes AFTER existing equal distances
 * ("ties keep earlier insertion first", tests/knn_result.cpp:46-47). */
typedef struct {
    int capacity, count;
    int32_t *indices;
    float *dists;
} orc_knn_result;

ORC_API void orc_knn_result_init(orc_knn_result *r, int capacity, int32_t *idx_buf, float *dist_buf) {
    r->capacity = capacity;
    r->count = 0;
    r->indices = idx_buf;
    r->dists = dist_buf;
}

ORC_API void orc_knn_result_add(orc_knn_result *r, float dist, int32_t index) {
    int i;
    for (i = r->count; i > 0; --i) {
        if (r->dists[i - 1] > dist) {
            if (i < r->capacity) {
                r->dists[i] = r->dists[i - 1];
                r->indices[i] = r->indices[i - 1];
            }
        } else {
            break;
        }
    }
    if (i < r->capacity) {
        r->dists[i] = dist;
        r->indices[i] = index;
    }
    if (r->count < r->capacity) r->count++;
}

/* ---- updateMultivaluedCorrespondence (src/common.cpp:517-529) -------------
 * Insert before the first entry whose distance is NOT < the new one, i.e. a
 * new entry goes BEFORE existing equal distances; then truncate to k.
 * idx/dist hold up to k+1 entries; *count is updated. */
ORC_API void orc_update_multivalued(int32_t *idx, float *dist, int *count, int k,
                                    int32_t match_idx, float distance) {
    int n = *count, pos = 0;
    while (pos != n && dist[pos] < distance) pos++;
    for (int i = n; i > pos; --i) {
        idx[i] = idx[i - 1];
        dist[i] = dist[i - 1];
    }
    idx[pos] = match_idx;
    dist[pos] = distance;
    n++;
    if (n > k) n = k;
    *count = n;
}

/* ---- canonical exact kNN ---------------------------------------------------
 * == matchLocal with match_search_radius = FLT_MAX and train rows visited in
 * ascending index order (include/matching.h:637-678, the way the reference's
 * test calls it, tests/flann_bf_matcher.h:66-72), == matchFLANN's result set
 * (include/matching.h:562-592; exact kd-tree search, sqrt at :586-588).
 * Rules:  invalid (non-finite) query -> empty list (:576,:655);
 *         invalid train rows are never candidates (:661; PCL KdTreeFLANN drops
 *         them when building the tree);
 *         list ascending by (distance, train index), at most k entries.
 * Output: idx[nq*k] (-1 padded), dist[nq*k] (+inf padded... written as 0 with
 *         idx -1), count[nq]. */
ORC_API void orc_knn_scalar(const float *query, size_t nq, size_t q_stride,
                            const float *train, size_t nt, size_t t_stride,
                            int dim, int k, int32_t *idx, float *dist, int32_t *count) {
    uint8_t *tvalid = (uint8_t *) malloc(nt ? nt : 1);
#pragma omp parallel for schedule(static)
    for (long j = 0; j < (long) nt; ++j) tvalid[j] = (uint8_t) orc_is_valid(row_at(train, t_stride, j), dim);

#pragma omp parallel for schedule(dynamic, 16)
    for (long i = 0; i < (long) nq; ++i) {
        int32_t *oi = idx + (size_t) i * k;
        float *od = dist + (size_t) i * k;
        for (int m = 0; m < k; ++m) { oi[m] = -1; od[m] = 0.f; }
        count[i] = 0;
        const float *q = row_at(query, q_stride, i);
        if (!orc_is_valid(q, dim)) continue;
        orc_knn_result r;
        orc_knn_result_init(&r, k, oi, od);
        size_t j = 0;
        for (; j + ORC_UNROLL <= nt; j += ORC_UNROLL) {
            float d[ORC_UNROLL];
            l2_norm_block(q, train, t_stride, j, dim, d);
            for (int u = 0; u < ORC_UNROLL; ++u) {
                if (!tvalid[j + u]) continue;
                /* cheap reject keeps the O(k) insertion off the common path; equal
                 * distances are rejected too, which is KNNResult's rule when full */
                if (r.count == k && !(d[u] < od[k - 1])) continue;
                orc_knn_result_add(&r, d[u], (int32_t) (j + u));
            }
        }
        for (; j < nt; ++j) {
            if (!tvalid[j]) continue;
            float d = orc_l2_norm(q, row_at(train, t_stride, j), dim);
            if (r.count == k && !(d < od[k - 1])) continue;
            orc_knn_result_add(&r, d, (int32_t) j);
        }
        count[i] = r.count;
    }
    free(tvalid);
}

/* ---- matchLocal with a finite match_search_radius (include/matching.h:637-678) -------
 * Per valid query: radiusSearch on the train keypoints around the (already guess-transformed) query keypoint --
 * FLANN L2_Simple squared distance (sequential FP32 sum) strictly below radius*radius, results sorted by that
 * distance (PCL's default) -- then pcl::L2_Norm on the descriptors of the gated, valid train rows, inserted into
 * KNNResult in that order: equal descriptor distances keep the spatially nearer row first.  (Equal spatial distances
 * too: lower index first here; FLANN's order among them is unspecified -- unpinned.)
 * query_xyz / train_xyz are [n][3]. */
typedef struct { float d2; int32_t j; } orc_gate;
static int gate_cmp(const void *a, const void *b) {
    const orc_gate *x = (const orc_gate *) a, *y = (const orc_gate *) b;
    if (x->d2 < y->d2) return -1;
    if (x->d2 > y->d2) return 1;
    return (x->j > y->j) - (x->j < y->j);
}
ORC_API void orc_match_local(const float *query, size_t nq, size_t q_stride, const float *train, size_t nt,
                             size_t t_stride, int dim, int k, const float *query_xyz, const float *train_xyz,
                             float radius, int32_t *idx, float *dist, int32_t *count) {
    const float r2 = radius * radius;
#pragma omp parallel
    {
        orc_gate *g = (orc_gate *) malloc(sizeof(orc_gate) * (nt ? nt : 1));
#pragma omp for schedule(dynamic, 16)
        for (long i = 0; i < (long) nq; ++i) {
            int32_t *oi = idx + (size_t) i * k;
            float *od = dist + (size_t) i * k;
            for (int m = 0; m < k; ++m) { oi[m] = -1; od[m] = 0.f; }
            count[i] = 0;
            const float *q = row_at(query, q_stride, i);
            if (!orc_is_valid(q, dim)) continue;
            const float *a = query_xyz + 3 * (size_t) i;
            size_t ng = 0;
            for (size_t j = 0; j < nt; ++j) {
                const float *b = train_xyz + 3 * j;
                float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
                float d2 = 0.f;
                d2 += dx * dx;
                d2 += dy * dy;
                d2 += dz * dz;
                if (d2 < r2) { g[ng].d2 = d2; g[ng].j = (int32_t) j; ng++; }
            }
            qsort(g, ng, sizeof(orc_gate), gate_cmp);
            orc_knn_result r;
            orc_knn_result_init(&r, k, oi, od);
            for (size_t e = 0; e < ng; ++e) {
                const float *t = row_at(train, t_stride, g[e].j);
                if (!orc_is_valid(t, dim)) continue;
                orc_knn_result_add(&r, orc_l2_norm(q, t, dim), g[e].j);
            }
            count[i] = r.count;
        }
        free(g);
    }
}

/* Same function, same per-pair arithmetic (bit-identical output, asserted in
 * tests/test_oracle.py), laid out for the CPU baseline timing: 8 train rows per
 * SIMD vector (one lane == one (query, train) pair, still a sequential
 * pcl::L2_Norm chain over d), train rows transposed per block so the loads are
 * contiguous, 4 vectors in flight to hide FP latency.  This is what bench.py
 * times as "the reference's CPU matcher" -- roughly the speed class of OpenCV's
 * SIMD batchDistance that matchBF calls. */
typedef float v8f __attribute__((vector_size(32)));
#define ORC_TB 512 /* train rows per transposed block (multiple of 32) */
#define ORC_QB 32  /* query rows per work item */

ORC_API void orc_knn(const float *query, size_t nq, size_t q_stride,
                     const float *train, size_t nt, size_t t_stride,
                     int dim, int k, int32_t *idx, float *dist, int32_t *count) {
    uint8_t *tvalid = (uint8_t *) malloc(nt ? nt : 1);
#pragma omp parallel for schedule(static)
    for (long j = 0; j < (long) nt; ++j) tvalid[j] = (uint8_t) orc_is_valid(row_at(train, t_stride, j), dim);
    long n_qb = (long) ((nq + ORC_QB - 1) / ORC_QB);

#pragma omp parallel
    {
        /* tr[g][d] = vector of train rows j0+8g .. j0+8g+7 at dimension d */
        v8f *tr = (v8f *) aligned_alloc(32, sizeof(v8f) * (size_t) (ORC_TB / 8) * (size_t) dim);
        uint8_t qvalid[ORC_QB];
#pragma omp for schedule(dynamic, 1)
        for (long qb = 0; qb < n_qb; ++qb) {
            size_t i0 = (size_t) qb * ORC_QB;
            size_t i1 = i0 + ORC_QB < nq ? i0 + ORC_QB : nq;
            for (size_t i = i0; i < i1; ++i) {
                for (int m = 0; m < k; ++m) { idx[i * k + m] = -1; dist[i * k + m] = 0.f; }
                count[i] = 0;
                qvalid[i - i0] = (uint8_t) orc_is_valid(row_at(query, q_stride, i), dim);
            }
            for (size_t j0 = 0; j0 < nt; j0 += ORC_TB) {
                size_t jn = j0 + ORC_TB < nt ? ORC_TB : nt - j0;
                size_t ng = (jn + 7) / 8;
                for (size_t g = 0; g < ng; ++g)
                    for (int l = 0; l < 8; ++l) {
                        size_t j = j0 + 8 * g + l;
                        const float *t = row_at(train, t_stride, j < nt ? j : nt - 1);
                        for (int d = 0; d < dim; ++d) ((float *) &tr[g * dim + d])[l] = t[d];
                    }
                for (size_t i = i0; i < i1; ++i) {
                    if (!qvalid[i - i0]) continue;
                    const float *q = row_at(query, q_stride, i);
                    int32_t *oi = idx + i * k;
                    float *od = dist + i * k;
                    orc_knn_result r;
                    r.capacity = k; r.count = count[i]; r.indices = oi; r.dists = od;
                    for (size_t g = 0; g < ng; g += 4) {
                        size_t gn = g + 4 <= ng ? 4 : ng - g;
                        v8f s0 = {0}, s1 = {0}, s2 = {0}, s3 = {0};
                        const v8f *t0 = tr + (g + 0) * dim, *t1 = tr + (g + (gn > 1 ? 1 : 0)) * dim;
                        const v8f *t2 = tr + (g + (gn > 2 ? 2 : 0)) * dim, *t3 = tr + (g + (gn > 3 ? 3 : 0)) * dim;
                        for (int d = 0; d < dim; ++d) {
                            float qd = q[d];
                            v8f qv = {qd, qd, qd, qd, qd, qd, qd, qd};
                            v8f d0 = qv - t0[d], d1 = qv - t1[d], d2 = qv - t2[d], d3 = qv - t3[d];
                            s0 = s0 + d0 * d0; s1 = s1 + d1 * d1; s2 = s2 + d2 * d2; s3 = s3 + d3 * d3;
                        }
                        float sq[32];
                        memcpy(sq, &s0, 32); memcpy(sq + 8, &s1, 32); memcpy(sq + 16, &s2, 32); memcpy(sq + 24, &s3, 32);
                        for (size_t u = 0; u < gn * 8; ++u) {
                            size_t j = j0 + 8 * g + u;
                            if (j >= nt || !tvalid[j]) continue;
                            float dd = sqrtf(sq[u]);
                            if (r.count == k && !(dd < od[k - 1])) continue;
                            orc_knn_result_add(&r, dd, (int32_t) j);
                        }
                    }
                    count[i] = r.count;
                }
            }
        }
        free(tr);
    }
    free(tvalid);
}

/* ---- matchBF (include/matching.h:594-634) ---------------------------------
 * Query and train sets are cut into blocks of `block_size` rows; each block
 * pair goes through OpenCV knnMatch (per-block top-k, ties -> lower train index
 * first, NaN rows never inserted -- behaviour verified against cv2 4.13 in
 * tests/golden/make_golden.py); per-block results are merged with
 * updateMultivaluedCorrespondence in (rank-within-block) order, which puts a
 * LATER block's entry before an earlier block's entry of exactly equal distance
 * (src/common.cpp:520). */
ORC_API void orc_match_bf(const float *query, size_t nq, size_t q_stride,
                          const float *train, size_t nt, size_t t_stride,
                          int dim, int k, int block_size,
                          int32_t *idx, float *dist, int32_t *count) {
    uint8_t *tvalid = (uint8_t *) malloc(nt ? nt : 1);
    for (size_t j = 0; j < nt; ++j) tvalid[j] = (uint8_t) orc_is_valid(row_at(train, t_stride, j), dim);
    size_t n_tblocks = (nt + block_size - 1) / block_size;

#pragma omp parallel
    {
        int32_t *bi = (int32_t *) malloc(sizeof(int32_t) * (k + 1));
        float *bd = (float *) malloc(sizeof(float) * (k + 1));
        int32_t *mi = (int32_t *) malloc(sizeof(int32_t) * (k + 1));
        float *md = (float *) malloc(sizeof(float) * (k + 1));
#pragma omp for schedule(dynamic, 16)
        for (long i = 0; i < (long) nq; ++i) {
            int32_t *oi = idx + (size_t) i * k;
            float *od = dist + (size_t) i * k;
            for (int m = 0; m < k; ++m) { oi[m] = -1; od[m] = 0.f; }
            count[i] = 0;
            const float *q = row_at(query, q_stride, i);
            if (!orc_is_valid(q, dim)) continue;
            int mcount = 0;
            for (size_t tb = 0; tb < n_tblocks; ++tb) {
                size_t j0 = tb * (size_t) block_size;
                size_t j1 = j0 + block_size < nt ? j0 + block_size : nt;
                orc_knn_result r;
                orc_knn_result_init(&r, k, bi, bd);
                for (size_t j = j0; j < j1; ++j) {
                    if (!tvalid[j]) continue;
                    float d = orc_l2_norm(q, row_at(train, t_stride, j), dim);
                    if (r.count == k && !(d < bd[k - 1])) continue;
                    orc_knn_result_add(&r, d, (int32_t) j);
                }
                for (int m = 0; m < r.count; ++m)
                    orc_update_multivalued(mi, md, &mcount, k, bi[m], bd[m]);
            }
            for (int m = 0; m < mcount; ++m) { oi[m] = mi[m]; od[m] = md[m]; }
            count[i] = mcount;
        }
        free(bi); free(bd); free(mi); free(md);
    }
    free(tvalid);
}

/* ---- spatial vote (match_multiscale, include/matching.h:327-352) ----------
 * Collapses each query's candidate list to at most one match: candidate m1
 * scores sum over m2>=m1 with |p_m1-p_m2| < 32*iss_radius of
 * iss_radius / max(|p_m1-p_m2|, iss_radius); best score wins, ties -> smaller
 * descriptor distance.  train_xyz is [nt][3]. In-place on idx/dist/count with
 * row pitch k. */
ORC_API void orc_spatial_vote(size_t nq, int k, int32_t *idx, float *dist, int32_t *count,
                              const float *train_xyz, float iss_radius) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long) nq; ++i) {
        int32_t *oi = idx + (size_t) i * k;
        float *od = dist + (size_t) i * k;
        int n = count[i];
        float best_count = 0.f, best_dist = 0.f;
        int best = -1;
        for (int m1 = 0; m1 < n; ++m1) {
            float c = 0.f;
            for (int m2 = m1; m2 < n; ++m2) {
                if (oi[m1] < 0 || oi[m2] < 0) continue;
                const float *p1 = train_xyz + 3 * (size_t) oi[m1];
                const float *p2 = train_xyz + 3 * (size_t) oi[m2];
                float dx = p1[0] - p2[0], dy = p1[1] - p2[1], dz = p1[2] - p2[2];
                /* Eigen's Vector3f::norm(): the unrolled fixed-size reduction adds x^2 + (y^2 + z^2); separate roundings
                 * (whether the reference's build contracts these into FMAs depends on PCL's exported -march flags:
                 * unpinned) */
                float yz = dy * dy + dz * dz;
                float d = sqrtf(dx * dx + yz);
                if (d < 32 * iss_radius) c += iss_radius / fmaxf(d, iss_radius);
            }
            if (c > best_count || (c == best_count && od[m1] < best_dist)) {
                best_count = c;
                best_dist = od[m1];
                best = m1;
            }
        }
        if (best >= 0) {
            int32_t bi = oi[best];
            float bd = od[best];
            oi[0] = bi; od[0] = bd;
            for (int m = 1; m < k; ++m) { oi[m] = -1; od[m] = 0.f; }
            count[i] = 1;
        } else {
            for (int m = 0; m < k; ++m) { oi[m] = -1; od[m] = 0.f; }
            count[i] = 0;
        }
    }
}

static inline float corr_threshold(const float *thr_q, const float *thr_t, long i, long j, float distance_thr) {
    /* std::min(std::max(thresholds_src[i], thresholds_tgt[j]), distance_thr)
     * (include/matching.h:404-405, :441-442); without density thresholds the
     * result is distance_thr. */
    if (!thr_q || !thr_t) return distance_thr;
    float t = thr_q[i] > thr_t[j] ? thr_q[i] : thr_t[j];
    return t < distance_thr ? t : distance_thr;
}

/* ---- OneSidedMatcher::match_impl (include/matching.h:395-411) ------------- */
ORC_API size_t orc_filter_one_sided(size_t nq, int k, const int32_t *fidx, const float *fdist,
                                    const int32_t *fcount, const float *thr_q, const float *thr_t,
                                    float distance_thr, orc_corr *out) {
    size_t n = 0;
    for (size_t i = 0; i < nq; ++i) {
        if (fcount[i] == 0) continue;
        int32_t j = fidx[i * k];
        out[n].index_query = (int32_t) i;
        out[n].index_match = j;
        out[n].distance = fdist[i * k];
        out[n].threshold = corr_threshold(thr_q, thr_t, (long) i, j, distance_thr);
        n++;
    }
    return n;
}

/* ---- LeftToRightMatcher::match_impl (include/matching.h:428-453) ----------
 * General k-list form: for i ascending, for j in fwd[i] in list order, emit
 * (i, j, rev[j].dist[m], thr) for the first m with rev[j].idx[m]==i.  The
 * emitted distance is the REVERSE list's (:443). */
ORC_API size_t orc_filter_mutual(size_t nq, int k, const int32_t *fidx, const int32_t *fcount,
                                 size_t nt, const int32_t *ridx, const float *rdist, const int32_t *rcount,
                                 const float *thr_q, const float *thr_t, float distance_thr, orc_corr *out) {
    (void) nt;
    size_t n = 0;
    for (size_t i = 0; i < nq; ++i) {
        for (int a = 0; a < fcount[i]; ++a) {
            int32_t j = fidx[i * k + a];
            for (int m = 0; m < rcount[j]; ++m) {
                if (ridx[(size_t) j * k + m] == (int32_t) i) {
                    out[n].index_query = (int32_t) i;
                    out[n].index_match = j;
                    out[n].distance = rdist[(size_t) j * k + m];
                    out[n].threshold = corr_threshold(thr_q, thr_t, (long) i, j, distance_thr);
                    n++;
                    break;
                }
            }
        }
    }
    return n;
}

/* ---- ClusterMatcher (include/matching.h:480-551) --------------------------
 * 3-D neighbourhoods: pcl::search::KdTree<PointN>::nearestKSearch(index, k) on the
 * keypoint clouds (:524-528) -- PCL's KdTreeFLANN, exact, FLANN L2_Simple squared
 * distance (sequential FP32 sum over x, y, z), the point itself included.  Only the
 * SET of neighbours is used (std::unordered_set, :521-528); members are chosen here
 * by (squared distance, lower index), so a tie AT the k-th distance resolves to the
 * lower index (the kd-tree's choice among exactly equidistant points is unspecified:
 * parity unpinned for that case).  nbr is [n][k], -1 padded when n < k. */
ORC_API void orc_knn3d(size_t n, const float *xyz, int k, int32_t *nbr) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long) n; ++i) {
        int32_t *oi = nbr + (size_t) i * k;
        float bd[64];
        int cnt = 0;
        const float *a = xyz + 3 * (size_t) i;
        for (size_t j = 0; j < n; ++j) {
            const float *b = xyz + 3 * j;
            float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
            float d = 0.f;
            d += dx * dx;
            d += dy * dy;
            d += dz * dz;
            if (cnt == k && !(d < bd[k - 1])) continue;   /* ascending j: an equal distance never displaces */
            int pos = cnt < k ? cnt : k - 1;
            while (pos > 0 && d < bd[pos - 1]) {
                bd[pos] = bd[pos - 1];
                oi[pos] = oi[pos - 1];
                pos--;
            }
            bd[pos] = d;
            oi[pos] = (int32_t) j;
            if (cnt < k) cnt++;
        }
        for (int m = cnt; m < k; ++m) oi[m] = -1;
    }
}

/* calculateCorrespondenceDistance (:519-550): 1 - (consistent pairs / pairs) over the matches of
 * i's 3-D neighbours, a pair being consistent when its match is a 3-D neighbour of j. */
static float cluster_distance(long i, long j, int ck, int k, const int32_t *fidx, const int32_t *fcount,
                              const int32_t *nbr_a, const int32_t *nbr_b) {
    const int32_t *ni = nbr_a + (size_t) i * ck, *nj = nbr_b + (size_t) j * ck;
    int consistent = 0, pairs = 0;
    for (int a = 0; a < ck; ++a) {
        if (ni[a] < 0) continue;
        const size_t ia = (size_t) ni[a];
        for (int m = 0; m < fcount[ia]; ++m) {
            int32_t match = fidx[ia * k + m];
            for (int b = 0; b < ck; ++b)
                if (nj[b] >= 0 && nj[b] == match) { consistent++; break; }
            pairs++;
        }
    }
    if (pairs == 0) return 0.f;
    return 1.f - (float) consistent / (float) pairs;
}

/* ClusterMatcher::match_impl (:492-517), k-list form: for i, for j in fwd[i]: keep (i, j,
 * max(d_i, d_j), threshold) iff both cluster distances < cluster_thr (MATCHING_CLUSTER_THRESHOLD). */
ORC_API size_t orc_filter_cluster(size_t nq, int k, const int32_t *fidx, const int32_t *fcount, size_t nt,
                                  const int32_t *ridx, const int32_t *rcount, int ck, const int32_t *nbr_src,
                                  const int32_t *nbr_tgt, float cluster_thr, const float *thr_q, const float *thr_t,
                                  float distance_thr, orc_corr *out) {
    (void) nt;
    size_t n = 0;
    for (size_t i = 0; i < nq; ++i) {
        for (int a = 0; a < fcount[i]; ++a) {
            int32_t j = fidx[i * k + a];
            float di = cluster_distance((long) i, j, ck, k, fidx, fcount, nbr_src, nbr_tgt);
            float dj = cluster_distance(j, (long) i, ck, k, ridx, rcount, nbr_tgt, nbr_src);
            if (di < cluster_thr && dj < cluster_thr) {
                out[n].index_query = (int32_t) i;
                out[n].index_match = j;
                out[n].distance = di > dj ? di : dj;
                out[n].threshold = corr_threshold(thr_q, thr_t, (long) i, j, distance_thr);
                n++;
            }
        }
    }
    return n;
}

/* ---- ratio filter: PARITY UNPINNED ---------------------------------------
 * RatioMatcher::match_impl is a stub in the reference (include/matching.h:
 * 470-473).  Defined here from the declared constants MATCHING_RATIO_K 2 and
 * MATCHING_RATIO_THRESHOLD 1.1f (include/common.h:50-51): with the two nearest
 * neighbours d1<=d2 keep (i, j1, d1) iff d2 >= ratio_thr * d1 (FP32 product);
 * queries with fewer than 2 neighbours are dropped. */
ORC_API size_t orc_filter_ratio(size_t nq, int k, const int32_t *fidx, const float *fdist,
                                const int32_t *fcount, float ratio_thr, const float *thr_q,
                                const float *thr_t, float distance_thr, orc_corr *out) {
    size_t n = 0;
    for (size_t i = 0; i < nq; ++i) {
        if (fcount[i] < 2) continue;
        float d1 = fdist[i * k], d2 = fdist[i * k + 1];
        if (!(d2 >= ratio_thr * d1)) continue;
        int32_t j = fidx[i * k];
        out[n].index_query = (int32_t) i;
        out[n].index_match = j;
        out[n].distance = d1;
        out[n].threshold = corr_threshold(thr_q, thr_t, (long) i, j, distance_thr);
        n++;
    }
    return n;
}

/* ---- FeatureBasedMatcher::printDebugInfo (src/matching.cpp:3-19) ----------
 * mean first-NN distance over non-empty lists, FP32 running sum in index
 * order; FLT_MAX when no list is non-empty (include/matching.h:41). */
ORC_API float orc_average_distance(size_t nq, int k, const float *fdist, const int32_t *fcount) {
    float sum = 0.f;
    int n = 0;
    for (size_t i = 0; i < nq; ++i)
        if (fcount[i] > 0) { sum += fdist[i * k]; n++; }
    if (n == 0) return 3.402823466e+38F;
    return sum / (float) n;
}

/* ---- FeatureBasedMatcherImpl::finalize (include/matching.h:356-362) ------- */
ORC_API void orc_finalize(orc_corr *corrs, size_t n, const int32_t *kps_src, const int32_t *kps_tgt) {
    for (size_t c = 0; c < n; ++c) {
        corrs[c].index_query = kps_src[corrs[c].index_query];
        corrs[c].index_match = kps_tgt[corrs[c].index_match];
    }
}
