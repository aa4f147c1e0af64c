// Compile check of the bindings INTEGRATION.md shows (tests/test_integration_doc.py): the snippets between the
// "snippet-begin" / "snippet-end" markers are the text of INTEGRATION.md sections 2 and 3, compiled here against
// stand-ins that carry the reference's own member names (include/matching.h:114-127 Storage, :296-311 the backend
// choice, include/common.h:192-200 MultivaluedCorrespondence).  PCL itself is not installed in this image, so the
// stand-ins only mirror the members the snippets touch.
#include <array>
#include <memory>
#include <optional>
#include <vector>

#include "b200match_shim.hpp"

namespace pcl {   // the few PCL names the snippets use
using Indices = std::vector<int>;
using IndicesConstPtr = std::shared_ptr<const Indices>;
struct PointNormal { float x, y, z, pad0, nx, ny, nz, pad1, curvature, pad2[3]; };
template <typename T> struct PointCloud {
    std::vector<T> points;
    size_t size() const { return points.size(); }
    using Ptr = std::shared_ptr<PointCloud<T>>;
    using ConstPtr = std::shared_ptr<const PointCloud<T>>;
};
}  // namespace pcl
static_assert(sizeof(pcl::PointNormal) == 48, "pcl::PointNormal is 48 bytes");
using PointN = pcl::PointNormal;
using PointNCloud = pcl::PointCloud<PointN>;

// reference include/common.h:192-200
struct MultivaluedCorrespondence {
    pcl::Indices match_indices;
    std::vector<float> distances;
};
// the fields of the reference's AlignmentParameters the snippets read (include/common.h:135-163) + the two new ones
struct AlignmentParameters {
    int randomness = 1;
    bool use_bfmatcher = true;
    float distance_thr = 0.f;
    int cluster_k = 40;
    std::string matching_id = "cluster";
    std::optional<std::array<float, 16>> guess;
    float match_search_radius = 0.f;
    bool use_b200 = true;            // new: YAML key `b200: true`
    std::vector<int> b200_devices;   // new: YAML key `b200_devices: [0, 1, ...]`
};
using Correspondences = std::vector<b200match::Correspondence>;
using CorrespondencesPtr = std::shared_ptr<Correspondences>;

template <typename FeatureT>
struct FeatureBasedMatcherImplStandIn {
    using FeatureCloud = std::vector<FeatureT>;   // pcl::PointCloud<FeatureT> in the reference; the shim's non-PCL mirror here
    struct Storage {                              // reference include/matching.h:114-127
        int min_log2_radius{0}, max_log2_radius{0};
        PointNCloud::Ptr kps{new PointNCloud};
        pcl::IndicesConstPtr kps_indices;
        std::vector<pcl::Indices> kps_indices_multiscale;
        std::vector<PointNCloud::Ptr> kps_multiscale;
        std::vector<std::shared_ptr<FeatureCloud>> kps_features_multiscale;
        float iss_radius{1.f};
    };
    Storage st_src_, st_tgt_;
    AlignmentParameters parameters_;
    float average_distance_ = 0.f;

    // ---- section 2: the narrow seam, one more branch in match_multiscale's backend choice (include/matching.h:296-311) ----
    std::vector<MultivaluedCorrespondence> backend_choice(const Storage &st_query, const Storage &st_train, int idx_query,
                                                          int idx_train) {
        std::vector<MultivaluedCorrespondence> mv_corrs_fixed_level;
        // snippet-begin narrow
        if (parameters_.use_b200) {                                  // new flag (YAML key `b200: true`)
            b200match::AlignmentParameters p;
            p.randomness = parameters_.randomness;                   // k
            p.devices = parameters_.b200_devices;                    // more than one id: rows sharded over these GPUs
            auto mv = b200match::matchBF<FeatureT>(*st_query.kps_features_multiscale[idx_query],
                                                   *st_train.kps_features_multiscale[idx_train], p);
            mv_corrs_fixed_level.resize(mv.size());
            for (size_t q = 0; q < mv.size(); ++q) {                 // same field names, same meaning
                mv_corrs_fixed_level[q].match_indices.assign(mv[q].match_indices.begin(), mv[q].match_indices.end());
                mv_corrs_fixed_level[q].distances = std::move(mv[q].distances);
            }
        }
        // snippet-end narrow
        return mv_corrs_fixed_level;
    }

    // ---- section 3: the wide seam, match_impl of LeftToRight / OneSided / ClusterMatcher in one call ----
    CorrespondencesPtr match_impl(const std::vector<float> &thresholds_src, const std::vector<float> &thresholds_tgt) {
        // snippet-begin wide
        using B200 = b200match::FeatureBasedMatcherImpl<FeatureT>;
        auto to_b200 = [](const Storage &st, const std::vector<float> &thresholds) {
            typename B200::Storage s;
            s.kps_xyz = &st.kps->points[0].x;                        // st.kps: keypoint coordinates
            s.n_kps = st.kps->size();
            s.kps_stride_bytes = sizeof(PointN);
            s.kps_indices = *st.kps_indices;                         // finalize (include/matching.h:356-362)
            s.kps_indices_multiscale = st.kps_indices_multiscale;    // per scale: row -> keypoint id
            for (const auto &f : st.kps_features_multiscale) s.kps_features_multiscale.push_back(*f);
            s.min_log2_radius = st.min_log2_radius;
            s.max_log2_radius = st.max_log2_radius;
            s.iss_radius = st.iss_radius;
            s.thresholds = thresholds;                               // calculateSmoothedDensities(st.kps)
            return s;
        };
        b200match::AlignmentParameters p;
        p.randomness = parameters_.randomness;
        p.distance_thr = parameters_.distance_thr;
        p.cluster_k = parameters_.cluster_k;
        p.matching_id = parameters_.matching_id;                     // one_sided | lr | cluster
        auto m = b200match::getFeatureBasedMatcherFromParameters<FeatureT>(to_b200(st_src_, thresholds_src),
                                                                           to_b200(st_tgt_, thresholds_tgt), p);
        auto corrs = m->match();                                     // match_impl + finalize: ascending index_query, cloud indices
        average_distance_ = m->getAverageDistance();                 // printDebugInfo's average (src/matching.cpp:3-19)
        CorrespondencesPtr correspondences(new Correspondences(corrs->begin(), corrs->end()));
        // snippet-end wide
        return correspondences;
    }

    // ---- section 2b: matchLocal when parameters_.guess is set (include/matching.h:297-304) ----
    std::vector<b200match::MultivaluedCorrespondence> local(const Storage &st_query, const Storage &st_train, int idx_query,
                                                            int idx_train, bool inverse_tn) {
        // snippet-begin local
        b200match::AlignmentParameters p;
        p.randomness = parameters_.randomness;
        p.match_search_radius = parameters_.match_search_radius;
        std::array<float, 16> guess = parameters_.guess.value();     // row-major Eigen::Matrix4f; invert it first when inverse_tn
        (void) inverse_tn;
        auto mv = b200match::matchLocal<FeatureT>(&st_query.kps_multiscale[idx_query]->points[0].x,
                                                  &st_train.kps_multiscale[idx_train]->points[0].x, sizeof(PointN),
                                                  *st_query.kps_features_multiscale[idx_query],
                                                  *st_train.kps_features_multiscale[idx_train], p, guess);
        // snippet-end local
        return mv;
    }
};

template struct FeatureBasedMatcherImplStandIn<b200match::FPFHSignature33>;
template struct FeatureBasedMatcherImplStandIn<b200match::SHOT352>;

int main() { return 0; }
