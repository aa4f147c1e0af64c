// b200match_shim.hpp -- header-only C++ host mirror of the reference's matcher interface on top of
// the C-ABI in b200match.h.  It puts the reference's own names and signatures back:
//
//   matchBF<FeatureT> / matchFLANN<FeatureT> / matchLocal<FeatureT>(radius = inf)
//                                  reference include/matching.h:368-383, :562-678
//   MultivaluedCorrespondence      reference include/common.h:192-195
//   Correspondence                 reference include/common.h:120-127
//   FeatureBasedMatcher, OneSidedMatcher / LeftToRightMatcher / RatioMatcher  (match(), getAverageDistance(),
//   getClassName())                reference include/matching.h:25-42, :385-478
//
// With -DB200MATCH_WITH_PCL the feature types are PCL's (pcl::FPFHSignature33, pcl::SHOT352,
// pcl::Histogram<135>) and clouds are pcl::PointCloud<FeatureT>::ConstPtr exactly as in the reference; without
// it (this repository's tests: PCL is not installed) layout-identical mirror structs are used and a "cloud"
// is a std::vector<FeatureT>.  Errors surface as std::runtime_error -- the convention of the reference's
// rassert (include/utils.h:9).  There is no CPU fallback anywhere below.
#pragma once
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200match.h"

#ifdef B200MATCH_WITH_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/point_representation.h>
#endif

namespace b200match {

#ifndef B200MATCH_WITH_PCL
// Layout mirrors of the PCL point structs the reference matches on (sizeof must equal PCL's).
struct FPFHSignature33 { float histogram[33]; static constexpr int descriptorSize() { return 33; } };
struct SHOT352 { float descriptor[352]; float rf[9]; static constexpr int descriptorSize() { return 352; } };
struct Histogram135 { float histogram[135]; static constexpr int descriptorSize() { return 135; } };
static_assert(sizeof(FPFHSignature33) == 132, "pcl::FPFHSignature33 is 132 bytes");
static_assert(sizeof(SHOT352) == 1444, "pcl::SHOT352 is 1444 bytes");
static_assert(sizeof(Histogram135) == 540, "pcl::Histogram<135> is 540 bytes");
template <typename FeatureT> using FeatureCloud = std::vector<FeatureT>;
template <typename FeatureT> inline const FeatureT *cloud_data(const FeatureCloud<FeatureT> &c) { return c.data(); }
template <typename FeatureT> inline size_t cloud_size(const FeatureCloud<FeatureT> &c) { return c.size(); }
template <typename FeatureT> inline int feature_dim() { return FeatureT::descriptorSize(); }
#else
template <typename FeatureT> using FeatureCloud = typename pcl::PointCloud<FeatureT>::ConstPtr;
template <typename FeatureT> inline const FeatureT *cloud_data(const FeatureCloud<FeatureT> &c) { return c->points.data(); }
template <typename FeatureT> inline size_t cloud_size(const FeatureCloud<FeatureT> &c) { return c->size(); }
template <typename FeatureT> inline int feature_dim() {
    return pcl::DefaultPointRepresentation<FeatureT>().getNumberOfDimensions();   // as include/matching.h:598-599
}
#endif

// reference include/common.h:192-195 (pcl::Indices == std::vector<int>)
struct MultivaluedCorrespondence {
    std::vector<int> match_indices;
    std::vector<float> distances;
};

// reference include/common.h:120-127; binary-identical to b200m_corr
struct Correspondence {
    int index_query = -1, index_match = -1;
    float distance = std::numeric_limits<float>::max();
    float threshold = 0.f;
};
static_assert(sizeof(Correspondence) == sizeof(b200m_corr), "Correspondence must stay 16 bytes");
typedef std::vector<Correspondence> Correspondences;
typedef std::shared_ptr<Correspondences> CorrespondencesPtr;

// The fields of AlignmentParameters (reference include/common.h:135-163) read on this path.
struct AlignmentParameters {
    bool use_bfmatcher = true;      // :144  either backend maps to the same exact GPU search
    int bf_block_size = 10000;      // :145  accepted, unused
    int ratio_k = 2;                // :146  MATCHING_RATIO_K
    int cluster_k = 40;             // :146  MATCHING_CLUSTER_K (include/common.h:53)
    int randomness = 1;             // :147  k
    float distance_thr = std::numeric_limits<float>::max();   // :139
    std::string matching_id = "lr"; // :149  one_sided | lr | ratio
    float ratio_thr = 1.1f;         // MATCHING_RATIO_THRESHOLD, include/common.h:50
    int device = 0;
    int precision = B200M_PREC_TC_F16;
};

class Context {   // RAII b200m_ctx; one per host thread
public:
    explicit Context(int device = 0) {
        if (b200m_create(&ctx_, device) != 0) throw std::runtime_error(b200m_last_error(nullptr));
    }
    ~Context() { b200m_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    b200m_ctx *get() const { return ctx_; }
    void check(int rc) const {
        if (rc != 0) throw std::runtime_error(b200m_last_error(ctx_));
    }
    template <typename FeatureT>
    void upload(int side, const FeatureCloud<FeatureT> &cloud) {
        // &cloud.points[0] with stride sizeof(FeatureT): what pcl2cv reads (reference include/matching.h:556-558)
        check(b200m_upload(ctx_, side, reinterpret_cast<const float *>(cloud_data<FeatureT>(cloud)), cloud_size<FeatureT>(cloud),
                           sizeof(FeatureT), feature_dim<FeatureT>(), 0));
    }

private:
    b200m_ctx *ctx_ = nullptr;
};

inline b200m_params make_params(const AlignmentParameters &p, int mode, int k) {
    b200m_params q;
    q.k = k;
    q.mode = mode;
    q.ratio_thr = p.ratio_thr;
    q.distance_thr = p.distance_thr;
    q.precision = p.precision;
    q.cand_cap = 0;
    return q;
}

// matchBF<FeatureT> (reference include/matching.h:594-634): one entry per query, <= k matches, ascending L2.
template <typename FeatureT>
std::vector<MultivaluedCorrespondence> matchBF(const FeatureCloud<FeatureT> &query_features,
                                               const FeatureCloud<FeatureT> &train_features,
                                               const AlignmentParameters &parameters) {
    Context ctx(parameters.device);
    ctx.upload<FeatureT>(0, query_features);
    ctx.upload<FeatureT>(1, train_features);
    const size_t nq = cloud_size<FeatureT>(query_features);
    const int k = parameters.randomness;
    std::vector<int32_t> idx(nq * k);
    std::vector<float> dist(nq * k);
    std::vector<int32_t> cnt(nq);
    b200m_params p = make_params(parameters, B200M_MODE_KNN_ONLY, k);
    ctx.check(b200m_knn(ctx.get(), &p, 0, 0, 0, idx.data(), dist.data(), cnt.data()));
    std::vector<MultivaluedCorrespondence> out(nq);
    for (size_t i = 0; i < nq; ++i) {
        out[i].match_indices.assign(idx.begin() + i * k, idx.begin() + i * k + cnt[i]);
        out[i].distances.assign(dist.begin() + i * k, dist.begin() + i * k + cnt[i]);
    }
    return out;
}

// matchFLANN<FeatureT> (reference include/matching.h:562-592): same exact result set.
template <typename FeatureT>
std::vector<MultivaluedCorrespondence> matchFLANN(const FeatureCloud<FeatureT> &query_features,
                                                  const FeatureCloud<FeatureT> &train_features,
                                                  const AlignmentParameters &parameters) {
    return matchBF<FeatureT>(query_features, train_features, parameters);
}

// matchLocal<FeatureT> with match_search_radius = inf (reference include/matching.h:637-678 as called by
// tests/flann_bf_matcher.h:66-72).  The spatially gated variant is a next-row item.
template <typename FeatureT>
std::vector<MultivaluedCorrespondence> matchLocal(const FeatureCloud<FeatureT> &query_features,
                                                  const FeatureCloud<FeatureT> &train_features,
                                                  const AlignmentParameters &parameters) {
    return matchBF<FeatureT>(query_features, train_features, parameters);
}

// match_multiscale (reference include/matching.h:264-354) over precomputed per-scale descriptors: one exact kNN per
// scale (k = parameters.randomness) over that scale's keypoint subset, per-scale row numbers renamed to keypoint ids
// through the scale's kps_indices_multiscale (:313-321), candidates concatenated per query keypoint in scale order, and
// the spatial vote over the train keypoints' xyz (:327-352) keeps at most ONE match per query keypoint.
//   query_features[s] / train_features[s]   kps_features_multiscale[s] of the two Storage objects (common scales only)
//   query_indices[s] / train_indices[s]     kps_indices_multiscale[s] (row of scale s -> keypoint id)
//   train_kps_xyz                           st_train.kps: n_train_kps rows, `xyz_stride_bytes` apart (pcl::PointXYZ: 16)
// Returns st_query.kps->size() entries, each empty or {one match index, its descriptor distance}.
template <typename FeatureT>
std::vector<MultivaluedCorrespondence> match_multiscale(const std::vector<FeatureCloud<FeatureT>> &query_features,
                                                        const std::vector<FeatureCloud<FeatureT>> &train_features,
                                                        const std::vector<std::vector<int>> &query_indices,
                                                        const std::vector<std::vector<int>> &train_indices,
                                                        size_t n_query_kps, const float *train_kps_xyz, size_t n_train_kps,
                                                        size_t xyz_stride_bytes, float iss_radius,
                                                        const AlignmentParameters &parameters) {
    const size_t n_scales = query_features.size();
    if (n_scales == 0 || train_features.size() != n_scales || query_indices.size() != n_scales || train_indices.size() != n_scales)
        throw std::runtime_error("match_multiscale: one descriptor cloud and one index map per scale and side are needed");
    Context ctx(parameters.device);
    const int k = parameters.randomness;
    ctx.check(b200m_multiscale_begin(ctx.get(), n_query_kps, (int) n_scales, k));
    b200m_params p = make_params(parameters, B200M_MODE_KNN_ONLY, k);
    static_assert(sizeof(int) == sizeof(int32_t), "index maps are passed as int32");
    for (size_t s = 0; s < n_scales; ++s) {
        if (query_indices[s].size() != cloud_size<FeatureT>(query_features[s]) ||
            train_indices[s].size() != cloud_size<FeatureT>(train_features[s]))
            throw std::runtime_error("match_multiscale: index map length != number of descriptors of the scale");
        ctx.upload<FeatureT>(0, query_features[s]);
        ctx.upload<FeatureT>(1, train_features[s]);
        ctx.check(b200m_multiscale_add(ctx.get(), &p, 0, (int) s, reinterpret_cast<const int32_t *>(query_indices[s].data()),
                                       reinterpret_cast<const int32_t *>(train_indices[s].data()), n_train_kps));
    }
    std::vector<int32_t> idx(n_query_kps), cnt(n_query_kps);
    std::vector<float> dist(n_query_kps);
    ctx.check(b200m_multiscale_vote(ctx.get(), train_kps_xyz, n_train_kps, xyz_stride_bytes, iss_radius, idx.data(), dist.data(),
                                    cnt.data()));
    std::vector<MultivaluedCorrespondence> out(n_query_kps);
    for (size_t i = 0; i < n_query_kps; ++i)
        if (cnt[i] > 0) {
            out[i].match_indices.push_back(idx[i]);
            out[i].distances.push_back(dist[i]);
        }
    return out;
}

// FeatureBasedMatcher (reference include/matching.h:25-42) at the descriptor seam.
class FeatureBasedMatcher {
public:
    using Ptr = std::shared_ptr<FeatureBasedMatcher>;
    virtual ~FeatureBasedMatcher() = default;
    virtual CorrespondencesPtr match() = 0;
    inline float getAverageDistance() const { return average_distance_; }
    virtual std::string getClassName() = 0;

protected:
    float average_distance_ = std::numeric_limits<float>::max();
};

template <typename FeatureT>
class FeatureBasedMatcherImpl : public FeatureBasedMatcher {
public:
    // thresholds_*: calculateSmoothedDensities outputs (may be empty); kps_indices_*: keypoint-local -> cloud-global
    // index maps applied by finalize (reference include/matching.h:356-362); may be empty.
    FeatureBasedMatcherImpl(FeatureCloud<FeatureT> src, FeatureCloud<FeatureT> tgt, AlignmentParameters parameters,
                            std::vector<float> thresholds_src = {}, std::vector<float> thresholds_tgt = {},
                            std::vector<int> kps_indices_src = {}, std::vector<int> kps_indices_tgt = {})
        : src_(std::move(src)), tgt_(std::move(tgt)), parameters_(std::move(parameters)),
          thr_src_(std::move(thresholds_src)), thr_tgt_(std::move(thresholds_tgt)),
          kps_src_(std::move(kps_indices_src)), kps_tgt_(std::move(kps_indices_tgt)) {}

    CorrespondencesPtr match() override {
        Context ctx(parameters_.device);
        ctx.upload<FeatureT>(0, src_);
        ctx.upload<FeatureT>(1, tgt_);
        const int k = mode() == B200M_MODE_RATIO ? (parameters_.ratio_k < 2 ? 2 : parameters_.ratio_k) : parameters_.randomness;
        b200m_params p = make_params(parameters_, mode(), k);
        const size_t nq = cloud_size<FeatureT>(src_);
        auto out = std::make_shared<Correspondences>(nq * (mode() == B200M_MODE_MUTUAL ? k : 1));
        size_t n = 0;
        const bool thr = !thr_src_.empty() && !thr_tgt_.empty();
        ctx.check(b200m_match(ctx.get(), &p, thr ? thr_src_.data() : nullptr, thr ? thr_tgt_.data() : nullptr,
                              reinterpret_cast<b200m_corr *>(out->data()), out->size(), &n, &average_distance_));
        out->resize(n);
        for (auto &c : *out) {   // finalize
            if (!kps_src_.empty()) c.index_query = kps_src_[c.index_query];
            if (!kps_tgt_.empty()) c.index_match = kps_tgt_[c.index_match];
        }
        return out;
    }

protected:
    virtual int mode() const = 0;
    FeatureCloud<FeatureT> src_, tgt_;
    AlignmentParameters parameters_;
    std::vector<float> thr_src_, thr_tgt_;
    std::vector<int> kps_src_, kps_tgt_;
};

template <typename FeatureT>
class OneSidedMatcher : public FeatureBasedMatcherImpl<FeatureT> {   // reference include/matching.h:385-416
public:
    using FeatureBasedMatcherImpl<FeatureT>::FeatureBasedMatcherImpl;
    std::string getClassName() override { return "OneSidedMatcher"; }
protected:
    int mode() const override { return B200M_MODE_ONE_SIDED; }
};

template <typename FeatureT>
class LeftToRightMatcher : public FeatureBasedMatcherImpl<FeatureT> {   // reference include/matching.h:418-458
public:
    using FeatureBasedMatcherImpl<FeatureT>::FeatureBasedMatcherImpl;
    std::string getClassName() override { return "LeftToRightMatcher"; }
protected:
    int mode() const override { return B200M_MODE_MUTUAL; }
};

template <typename FeatureT>
class RatioMatcher : public FeatureBasedMatcherImpl<FeatureT> {   // reference include/matching.h:460-478 (stub there)
public:
    using FeatureBasedMatcherImpl<FeatureT>::FeatureBasedMatcherImpl;
    std::string getClassName() override { return "RatioMatcher"; }
protected:
    int mode() const override { return B200M_MODE_RATIO; }
};

// ClusterMatcher (reference include/matching.h:480-551, the default matching_id): needs the keypoint coordinates of
// both sides (st_src_.kps / st_tgt_.kps), one row per descriptor, `xyz_stride_bytes` apart (pcl::PointXYZ / PointN rows).
template <typename FeatureT>
class ClusterMatcher : public FeatureBasedMatcherImpl<FeatureT> {
public:
    ClusterMatcher(FeatureCloud<FeatureT> src, FeatureCloud<FeatureT> tgt, AlignmentParameters parameters,
                   const float *src_kps_xyz, const float *tgt_kps_xyz, size_t xyz_stride_bytes,
                   std::vector<float> thresholds_src = {}, std::vector<float> thresholds_tgt = {},
                   std::vector<int> kps_indices_src = {}, std::vector<int> kps_indices_tgt = {})
        : FeatureBasedMatcherImpl<FeatureT>(std::move(src), std::move(tgt), std::move(parameters), std::move(thresholds_src),
                                            std::move(thresholds_tgt), std::move(kps_indices_src), std::move(kps_indices_tgt)),
          sx_(src_kps_xyz), tx_(tgt_kps_xyz), stride_(xyz_stride_bytes) {}
    std::string getClassName() override { return "ClusterMatcher"; }

    CorrespondencesPtr match() override {
        Context ctx(this->parameters_.device);
        ctx.template upload<FeatureT>(0, this->src_);
        ctx.template upload<FeatureT>(1, this->tgt_);
        const int k = this->parameters_.randomness;
        b200m_params p = make_params(this->parameters_, B200M_MODE_CLUSTER, k);
        auto out = std::make_shared<Correspondences>(cloud_size<FeatureT>(this->src_) * k);
        size_t n = 0;
        const bool thr = !this->thr_src_.empty() && !this->thr_tgt_.empty();
        ctx.check(b200m_match_cluster(ctx.get(), &p, this->parameters_.cluster_k, sx_, tx_, stride_,
                                      thr ? this->thr_src_.data() : nullptr, thr ? this->thr_tgt_.data() : nullptr,
                                      reinterpret_cast<b200m_corr *>(out->data()), out->size(), &n, &this->average_distance_));
        out->resize(n);
        for (auto &c : *out) {   // finalize
            if (!this->kps_src_.empty()) c.index_query = this->kps_src_[c.index_query];
            if (!this->kps_tgt_.empty()) c.index_match = this->kps_tgt_[c.index_match];
        }
        return out;
    }

protected:
    int mode() const override { return B200M_MODE_CLUSTER; }
    const float *sx_, *tx_;
    size_t stride_;
};

// getFeatureBasedMatcherFromParameters (reference src/matching.cpp:21-76) for one feature type.  ("cluster" needs the
// keypoint coordinates: construct b200match::ClusterMatcher<FeatureT> directly.)
template <typename FeatureT, typename... Args>
FeatureBasedMatcher::Ptr getFeatureBasedMatcherFromParameters(const FeatureCloud<FeatureT> &src, const FeatureCloud<FeatureT> &tgt,
                                                              const AlignmentParameters &parameters, Args &&...rest) {
    if (parameters.matching_id == "one_sided")
        return std::make_shared<OneSidedMatcher<FeatureT>>(src, tgt, parameters, std::forward<Args>(rest)...);
    if (parameters.matching_id == "lr")
        return std::make_shared<LeftToRightMatcher<FeatureT>>(src, tgt, parameters, std::forward<Args>(rest)...);
    if (parameters.matching_id == "ratio")
        return std::make_shared<RatioMatcher<FeatureT>>(src, tgt, parameters, std::forward<Args>(rest)...);
    throw std::runtime_error("Matching method " + parameters.matching_id + " isn't supported by the B200 matcher");
}

}  // namespace b200match
