// C++ user of the shim, written like the reference's own tests/flann_bf_matcher.h:31-90: BF == FLANN == Local
// indices in both directions, then the matcher classes.  Descriptors come from a raw float32 dump written by the
// pytest driver (tests/test_cpp_shim.py), results go back as text so the driver can compare with the oracle.
//   shim_test <fpfh|shot|rops> <src.bin> <n_src> <tgt.bin> <n_tgt> <k> <out.txt>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "b200match_shim.hpp"

using namespace b200match;

static int g_devices = 1;

template <typename FeatureT>
static std::vector<FeatureT> load(const char *path, size_t n) {
    std::vector<FeatureT> v(n);
    std::ifstream f(path, std::ios::binary);
    f.read(reinterpret_cast<char *>(v.data()), (std::streamsize) (n * sizeof(FeatureT)));
    if (!f) throw std::runtime_error(std::string("cannot read ") + path);
    return v;
}

static void assertCorrespondencesEqual(int i, const MultivaluedCorrespondence &a, const MultivaluedCorrespondence &b) {
    if (a.match_indices != b.match_indices) {   // tests/flann_bf_matcher.h:16-29
        std::cerr << "[" << i << "] match indices differ\n";
        abort();
    }
}

template <typename FeatureT>
static int run(char **argv) {
    size_t ns = std::stoul(argv[3]), nt = std::stoul(argv[5]);
    auto src = load<FeatureT>(argv[2], ns);
    auto tgt = load<FeatureT>(argv[4], nt);
    AlignmentParameters parameters;
    parameters.randomness = std::stoi(argv[6]);
    FILE *out = fopen(argv[7], "w");
    for (int dir = 0; dir < 2; ++dir) {
        const auto &q = dir ? tgt : src;
        const auto &t = dir ? src : tgt;
        auto bf = matchBF<FeatureT>(q, t, parameters);
        auto flann = matchFLANN<FeatureT>(q, t, parameters);
        auto local = matchLocal<FeatureT>(q, t, parameters);
        if (bf.size() != q.size()) abort();   // rassert at include/matching.h:312
        for (size_t i = 0; i < q.size(); ++i) {
            assertCorrespondencesEqual((int) i, bf[i], flann[i]);
            assertCorrespondencesEqual((int) i, bf[i], local[i]);
            fprintf(out, "knn %d %zu %zu", dir, i, bf[i].match_indices.size());
            for (size_t m = 0; m < bf[i].match_indices.size(); ++m)
                fprintf(out, " %d %.9g", bf[i].match_indices[m], bf[i].distances[m]);
            fprintf(out, "\n");
        }
    }
    // keypoints laid out on a deterministic lattice (the pytest driver builds the same coordinates): pcl::PointXYZ-like
    // rows of 4 floats
    auto lattice = [](size_t n, float step) {
        std::vector<float> xyz(4 * n, 0.f);
        for (size_t i = 0; i < n; ++i) {
            xyz[4 * i + 0] = step * (float) (i % 17);
            xyz[4 * i + 1] = step * (float) ((i / 17) % 13);
            xyz[4 * i + 2] = step * (float) (i / 221);
        }
        return xyz;
    };
    const auto sx = lattice(ns, 0.5f), tx = lattice(nt, 0.25f);
    using Impl = FeatureBasedMatcherImpl<FeatureT>;
    // the matcher classes, one scale: match_multiscale both ways + vote + filter (reference include/matching.h:395-517)
    parameters.cluster_k = 12;
    for (const char *id : {"one_sided", "lr", "cluster", "ratio"}) {
        if (!strcmp(id, "ratio") && parameters.randomness < 2) continue;
        parameters.matching_id = id;
        auto matcher = getFeatureBasedMatcherFromParameters<FeatureT>(Impl::makeStorage(src, sx.data(), 16, 0.3f),
                                                                      Impl::makeStorage(tgt, tx.data(), 16, 0.2f), parameters);
        auto corrs = matcher->match();
        fprintf(out, "matcher %s %s %zu %.9g\n", id, matcher->getClassName().c_str(), corrs->size(), matcher->getAverageDistance());
        for (const auto &c : *corrs) fprintf(out, "corr %s %d %d %.9g\n", id, c.index_query, c.index_match, c.distance);
    }
    try {   // more than one candidate per keypoint and no coordinates for the vote
        parameters.matching_id = "lr";
        AlignmentParameters p2 = parameters;
        p2.randomness = 2;
        getFeatureBasedMatcherFromParameters<FeatureT>(Impl::makeStorage(src), Impl::makeStorage(tgt), p2)->match();
        abort();
    } catch (const std::runtime_error &) {
    }
    // two "scales" (all keypoints, then every second one on both sides): the narrow-seam match_multiscale and the
    // LeftToRightMatcher over the same Storage contents, finalize included
    {
        std::vector<FeatureCloud<FeatureT>> qf{src, {}}, tf{tgt, {}};
        std::vector<std::vector<int>> qi(2), ti(2);
        for (size_t i = 0; i < ns; ++i) qi[0].push_back((int) i);
        for (size_t i = 0; i < nt; ++i) ti[0].push_back((int) i);
        for (size_t i = 0; i < ns; i += 2) { qf[1].push_back(src[i]); qi[1].push_back((int) i); }
        for (size_t i = 0; i < nt; i += 2) { tf[1].push_back(tgt[i]); ti[1].push_back((int) i); }
        auto mv = match_multiscale<FeatureT>(qf, tf, qi, ti, ns, tx.data(), nt, 16, 0.3f, parameters);
        if (mv.size() != ns) abort();
        for (size_t i = 0; i < ns; ++i) {
            fprintf(out, "ms %zu %zu", i, mv[i].match_indices.size());
            for (size_t m = 0; m < mv[i].match_indices.size(); ++m) fprintf(out, " %d %.9g", mv[i].match_indices[m], mv[i].distances[m]);
            fprintf(out, "\n");
        }
        typename Impl::Storage ss, st;
        ss.kps_xyz = sx.data(); ss.n_kps = ns; ss.iss_radius = 0.3f; ss.kps_features_multiscale = qf; ss.kps_indices_multiscale = qi;
        st.kps_xyz = tx.data(); st.n_kps = nt; st.iss_radius = 0.2f; st.kps_features_multiscale = tf; st.kps_indices_multiscale = ti;
        ss.min_log2_radius = st.min_log2_radius = -3;
        ss.max_log2_radius = st.max_log2_radius = -2;
        for (size_t i = 0; i < ns; ++i) ss.kps_indices.push_back((int) (3 * i + 1));   // keypoint id -> cloud index
        for (size_t i = 0; i < nt; ++i) st.kps_indices.push_back((int) (2 * i + 5));
        parameters.matching_id = "lr";
        auto matcher = getFeatureBasedMatcherFromParameters<FeatureT>(ss, st, parameters);
        auto corrs = matcher->match();
        fprintf(out, "matcher lr2 %s %zu %.9g\n", matcher->getClassName().c_str(), corrs->size(), matcher->getAverageDistance());
        for (const auto &c : *corrs) fprintf(out, "corr lr2 %d %d %.9g\n", c.index_query, c.index_match, c.distance);
    }
    // matchLocal as match_multiscale calls it with a guess (reference include/matching.h:297-304, :637-678)
    {
        parameters.match_search_radius = 1.75f;
        const std::array<float, 16> guess{1.f, 0.f, 0.f, 0.25f, 0.f, 1.f, 0.f, -0.5f, 0.f, 0.f, 1.f, 0.125f, 0.f, 0.f, 0.f, 1.f};
        auto loc = matchLocal<FeatureT>(sx.data(), tx.data(), 16, src, tgt, parameters, guess);
        if (loc.size() != ns) abort();
        for (size_t i = 0; i < ns; ++i) {
            fprintf(out, "local %zu %zu", i, loc[i].match_indices.size());
            for (size_t m = 0; m < loc[i].match_indices.size(); ++m) fprintf(out, " %d %.9g", loc[i].match_indices[m], loc[i].distances[m]);
            fprintf(out, "\n");
        }
    }
    // several GPUs (optional 9th argument): the same calls partitioned inside the library must give the same answers
    if (g_devices > 1) {
        AlignmentParameters pm;
        pm.randomness = std::stoi(argv[6]);
        for (int d = 0; d < g_devices; ++d) pm.devices.push_back(d);
        auto bf = matchBF<FeatureT>(src, tgt, pm);
        if (bf.size() != ns) abort();
        for (size_t i = 0; i < ns; ++i) {
            fprintf(out, "mknn %zu %zu", i, bf[i].match_indices.size());
            for (size_t m = 0; m < bf[i].match_indices.size(); ++m) fprintf(out, " %d %.9g", bf[i].match_indices[m], bf[i].distances[m]);
            fprintf(out, "\n");
        }
        pm.randomness = 1;
        for (const char *id : {"one_sided", "lr"}) {
            pm.matching_id = id;
            auto matcher = getFeatureBasedMatcherFromParameters<FeatureT>(Impl::makeStorage(src), Impl::makeStorage(tgt), pm);
            auto corrs = matcher->match();
            fprintf(out, "matcher m_%s %s %zu %.9g\n", id, matcher->getClassName().c_str(), corrs->size(), matcher->getAverageDistance());
            for (const auto &c : *corrs) fprintf(out, "corr m_%s %d %d %.9g\n", id, c.index_query, c.index_match, c.distance);
        }
    }
    fclose(out);
    return 0;
}

int main(int argc, char **argv) {
    if (argc != 8 && argc != 9) {
        std::cerr << "usage: shim_test <fpfh|shot|rops> src.bin n_src tgt.bin n_tgt k out.txt [n_gpus]\n";
        return 2;
    }
    if (argc == 9) g_devices = std::stoi(argv[8]);
    try {
        if (!strcmp(argv[1], "fpfh")) return run<FPFHSignature33>(argv);
        if (!strcmp(argv[1], "shot")) return run<SHOT352>(argv);
        if (!strcmp(argv[1], "rops")) return run<Histogram135>(argv);
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    return 2;
}
