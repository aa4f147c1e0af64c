// wide.cu -- the reference's matcher classes at the WIDE seam: FeatureBasedMatcherImpl<FeatureT>::match_impl of
// OneSidedMatcher / LeftToRightMatcher / ClusterMatcher (reference include/matching.h:395-411, :428-453, :492-517),
// composed exactly as the reference composes them:
//
//     mv_corrs_ij = match_multiscale(st_src_, st_tgt_)            per-scale kNN, remap to keypoint ids, concatenate,
//     mv_corrs_ji = match_multiscale(st_tgt_, st_src_, true)      spatial vote -> AT MOST ONE match per keypoint (:264-354)
//     printDebugInfo(mv_corrs_ij)                                 average first-NN distance of the VOTED forward lists
//     filter loop over the voted lists                            one-sided / mutual / cluster
//
// Everything between the descriptor upload and the correspondence records stays in HBM: per scale the two clouds'
// descriptors are packed, the forward (and, for the two-way matchers, the reverse = inverse_tn) kNN runs through the
// tensor-core candidate pass + exact re-rank, its k-lists are filed under the keypoint ids (ms_scatter_kernel), the two
// votes reduce both tables to one-entry lists (ms_vote_kernel), and the filter kernels of filter.cu / cluster.cu run with
// k = 1 on those.  With one scale and k = 1 the vote is the identity and the result equals b200m_match's.
//
// The single-scale mutual matcher skips reverse rows no voted forward list names (the masked reverse pass of api.cu):
// LeftToRightMatcher reads mv_corrs_ji[j] only for j in mv_corrs_ij[i] (:437-438), and a target keypoint's vote only
// depends on its own lists.
#include <string>

#include "internal.cuh"

namespace {

struct WideState {
    MultiscaleState fwd, rev;
    DevBuf xyz_s, xyz_t, smap, tmap, kidx, kdist, kcnt;
    DevBuf vfi, vfd, vfc, vri, vrd, vrc, kp_flags, row_flags;
};

WideState *wide_state(b200m_ctx *ctx) {
    if (!ctx->wide) ctx->wide = new WideState();
    return static_cast<WideState *>(ctx->wide);
}

// flags over the target keypoints that a voted forward list names
__global__ void mark_voted_kernel(const int32_t *__restrict__ vidx, const int32_t *__restrict__ vcnt, size_t n, uint8_t *flags,
                                  size_t n_flags) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || vcnt[i] == 0) return;
    const int32_t j = vidx[i];
    if (j >= 0 && (size_t) j < n_flags) flags[j] = 1;
}

// keypoint flags -> flags of this scale's descriptor rows (row r is keypoint map[r])
__global__ void expand_flags_kernel(const uint8_t *__restrict__ kp_flags, size_t n_kps, const int32_t *__restrict__ map,
                                    size_t n_rows, uint8_t *__restrict__ row_flags) {
    const size_t r = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const long long kp = map ? (long long) map[r] : (long long) r;
    row_flags[r] = (kp >= 0 && (size_t) kp < n_kps) ? kp_flags[kp] : 0;
}

int copy_rows(b200m_ctx *ctx, DevBuf &dst, const float *host, size_t n, size_t stride_bytes) {
    CK(dst.reserve(n * stride_bytes + 16));
    if (n) CK(cudaMemcpyAsync(dst.p, host, (n - 1) * stride_bytes + 12, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

}  // namespace

void wide_release(b200m_ctx *ctx) {
    WideState *ws = static_cast<WideState *>(ctx->wide);
    if (!ws) return;
    for (MultiscaleState *ms : {&ws->fwd, &ws->rev}) {
        DevBuf *b[] = {&ms->idx, &ms->dist, &ms->cnt, &ms->bad, &ms->qmap, &ms->tmap, &ms->xyz, &ms->oidx, &ms->odist, &ms->ocnt,
                       &ms->kidx, &ms->kdist, &ms->kcnt};
        for (DevBuf *x : b) x->release();
    }
    DevBuf *b[] = {&ws->xyz_s, &ws->xyz_t, &ws->smap, &ws->tmap, &ws->kidx, &ws->kdist, &ws->kcnt, &ws->vfi, &ws->vfd, &ws->vfc,
                   &ws->vri, &ws->vrd, &ws->vrc, &ws->kp_flags, &ws->row_flags};
    for (DevBuf *x : b) x->release();
    delete ws;
    ctx->wide = nullptr;
}

extern "C" int b200m_match_multiscale(b200m_ctx *ctx, const b200m_params *p, const b200m_scale *scales, int n_scales,
                                      size_t stride_bytes, int dim, const float *src_kps_xyz, size_t n_src_kps,
                                      const float *tgt_kps_xyz, size_t n_tgt_kps, size_t xyz_stride_bytes, float iss_radius_src,
                                      float iss_radius_tgt, int cluster_k, const float *thr_src, const float *thr_tgt,
                                      b200m_corr *out, size_t cap, size_t *n_out, float *avg_first_dist) {
    if (!ctx) return b200m_fail_msg(nullptr, "null context");
    CK(cudaSetDevice(ctx->device));
    if (!p || !n_out) return b200m_fail_msg(ctx, "b200m_match_multiscale: null params / n_out");
    *n_out = 0;
    const int mode = p->mode;
    if (mode != B200M_MODE_ONE_SIDED && mode != B200M_MODE_MUTUAL && mode != B200M_MODE_CLUSTER)
        return b200m_fail_msg(ctx, "b200m_match_multiscale: mode must be ONE_SIDED, MUTUAL or CLUSTER (the reference's implemented matchers)");
    if (p->k < 1 || p->k > B200M_MAX_K) return b200m_fail_msg(ctx, "b200m_match_multiscale: params.k must be in [1, 32]");
    if (!scales || n_scales < 1) return b200m_fail_msg(ctx, "b200m_match_multiscale: need at least one scale");
    if ((thr_src == nullptr) != (thr_tgt == nullptr))
        return b200m_fail_msg(ctx, "b200m_match_multiscale: give both threshold arrays or neither");
    if (xyz_stride_bytes % 4 != 0 || xyz_stride_bytes < 12)
        return b200m_fail_msg(ctx, "b200m_match_multiscale: xyz stride must be a multiple of 4 and >= 12 bytes");
    if (n_src_kps >= ((size_t) 1 << 31) || n_tgt_kps >= ((size_t) 1 << 31))
        return b200m_fail_msg(ctx, "b200m_match_multiscale: more than 2^31-1 keypoints");
    if (n_src_kps == 0) {
        if (avg_first_dist) *avg_first_dist = 3.402823466e+38F;   // FLT_MAX, reference include/matching.h:41
        return 0;
    }
    if ((n_tgt_kps && !tgt_kps_xyz) || (mode != B200M_MODE_ONE_SIDED && !src_kps_xyz))
        return b200m_fail_msg(ctx, "b200m_match_multiscale: null keypoint coordinates");
    const bool two_way = mode != B200M_MODE_ONE_SIDED;
    const int k = p->k;
    cudaStream_t st = ctx->stream;
    WideState *ws = wide_state(ctx);
    if (ms_begin(ctx, &ws->fwd, n_src_kps, n_scales, k)) return 1;
    if (two_way && ms_begin(ctx, &ws->rev, n_tgt_kps, n_scales, k)) return 1;
    if (copy_rows(ctx, ws->xyz_t, tgt_kps_xyz, n_tgt_kps, xyz_stride_bytes)) return 1;
    if (two_way && copy_rows(ctx, ws->xyz_s, src_kps_xyz, n_src_kps, xyz_stride_bytes)) return 1;
    CK(ws->vfi.reserve(sizeof(int32_t) * n_src_kps));
    CK(ws->vfd.reserve(sizeof(float) * n_src_kps));
    CK(ws->vfc.reserve(sizeof(int32_t) * n_src_kps));
    if (two_way) {
        CK(ws->vri.reserve(sizeof(int32_t) * (n_tgt_kps + 1)));
        CK(ws->vrd.reserve(sizeof(float) * (n_tgt_kps + 1)));
        CK(ws->vrc.reserve(sizeof(int32_t) * (n_tgt_kps + 1)));
    }
    b200m_params pk = *p;
    pk.mode = B200M_MODE_KNN_ONLY;
    // the mutual filter of a single-scale run reads the reverse list of target keypoint j only if a voted forward list
    // names j: do the forward vote first and answer only those rows in the reverse pass
    const bool masked = mode == B200M_MODE_MUTUAL && n_scales == 1 &&
                        (double) scales[0].n_src * (double) scales[0].n_tgt >= ctx->masked_min_pairs;
    for (int s = 0; s < n_scales; ++s) {
        const b200m_scale &sc = scales[s];
        if ((sc.n_src && !sc.src_desc) || (sc.n_tgt && !sc.tgt_desc))
            return b200m_fail_msg(ctx, "b200m_match_multiscale: null descriptor pointer in scale " + std::to_string(s));
        if (b200m_upload(ctx, 0, sc.src_desc, sc.n_src, stride_bytes, dim, 0)) return 1;
        if (b200m_upload(ctx, 1, sc.tgt_desc, sc.n_tgt, stride_bytes, dim, 0)) return 1;
        const int32_t *d_smap = nullptr, *d_tmap = nullptr;
        if (sc.src_map && sc.n_src) {
            CK(ws->smap.reserve(sizeof(int32_t) * sc.n_src));
            CK(cudaMemcpyAsync(ws->smap.p, sc.src_map, sizeof(int32_t) * sc.n_src, cudaMemcpyHostToDevice, st));
            d_smap = ws->smap.as<int32_t>();
        }
        if (sc.tgt_map && sc.n_tgt) {
            CK(ws->tmap.reserve(sizeof(int32_t) * sc.n_tgt));
            CK(cudaMemcpyAsync(ws->tmap.p, sc.tgt_map, sizeof(int32_t) * sc.n_tgt, cudaMemcpyHostToDevice, st));
            d_tmap = ws->tmap.as<int32_t>();
        }
        const size_t n_max = sc.n_src > sc.n_tgt ? sc.n_src : sc.n_tgt;
        CK(ws->kidx.reserve(sizeof(int32_t) * (n_max * k + 1)));
        CK(ws->kdist.reserve(sizeof(float) * (n_max * k + 1)));
        CK(ws->kcnt.reserve(sizeof(int32_t) * (n_max + 1)));
        if (sc.n_src) {
            if (b200m_knn_device(ctx, &pk, 0, 0, sc.n_src, ws->kidx.as<int32_t>(), ws->kdist.as<float>(), ws->kcnt.as<int32_t>()))
                return 1;
            if (ms_add_device(ctx, &ws->fwd, s, sc.n_src, ws->kidx.as<int32_t>(), ws->kdist.as<float>(), ws->kcnt.as<int32_t>(),
                              d_smap, d_tmap, sc.n_tgt, 0, n_tgt_kps))
                return 1;
        }
        if (masked) {   // single scale: forward vote -> referenced target keypoints -> rows of this scale
            if (ms_vote_device(ctx, &ws->fwd, ws->xyz_t.as<float>(), xyz_stride_bytes, iss_radius_tgt, ws->vfi.as<int32_t>(),
                               ws->vfd.as<float>(), ws->vfc.as<int32_t>()))
                return 1;
            CK(ws->kp_flags.reserve(n_tgt_kps + 1));
            CK(ws->row_flags.reserve(sc.n_tgt + 1));
            CK(cudaMemsetAsync(ws->kp_flags.p, 0, n_tgt_kps + 1, st));
            mark_voted_kernel<<<(unsigned) ((n_src_kps + 255) / 256), 256, 0, st>>>(ws->vfi.as<int32_t>(), ws->vfc.as<int32_t>(),
                                                                                 n_src_kps, ws->kp_flags.as<uint8_t>(), n_tgt_kps);
            if (sc.n_tgt)
                expand_flags_kernel<<<(unsigned) ((sc.n_tgt + 255) / 256), 256, 0, st>>>(ws->kp_flags.as<uint8_t>(), n_tgt_kps, d_tmap,
                                                                                      sc.n_tgt, ws->row_flags.as<uint8_t>());
            CK(cudaGetLastError());
            ctx->stats.launches += 2;
        }
        if (two_way && sc.n_tgt) {
            if (masked) {
                if (b200m_knn_masked_device(ctx, &pk, 1, 0, sc.n_tgt, ws->row_flags.as<uint8_t>(), ws->kidx.as<int32_t>(),
                                            ws->kdist.as<float>(), ws->kcnt.as<int32_t>()))
                    return 1;
            } else if (b200m_knn_device(ctx, &pk, 1, 0, sc.n_tgt, ws->kidx.as<int32_t>(), ws->kdist.as<float>(),
                                        ws->kcnt.as<int32_t>())) {
                return 1;
            }
            if (ms_add_device(ctx, &ws->rev, s, sc.n_tgt, ws->kidx.as<int32_t>(), ws->kdist.as<float>(), ws->kcnt.as<int32_t>(),
                              d_tmap, d_smap, sc.n_src, 0, n_src_kps))
                return 1;
        }
    }
    // the two votes (match_multiscale's tail): train keypoints = the OTHER cloud's, iss_radius = the train side's (:336-343)
    if (!masked && ms_vote_device(ctx, &ws->fwd, ws->xyz_t.as<float>(), xyz_stride_bytes, iss_radius_tgt, ws->vfi.as<int32_t>(),
                                  ws->vfd.as<float>(), ws->vfc.as<int32_t>()))
        return 1;
    if (two_way && n_tgt_kps &&
        ms_vote_device(ctx, &ws->rev, ws->xyz_s.as<float>(), xyz_stride_bytes, iss_radius_src, ws->vri.as<int32_t>(),
                       ws->vrd.as<float>(), ws->vrc.as<int32_t>()))
        return 1;
    const float *d_thr_s = nullptr, *d_thr_t = nullptr;
    if (thr_src && n_tgt_kps) {
        CK(ctx->ws_thr[0].reserve(sizeof(float) * n_src_kps));
        CK(ctx->ws_thr[1].reserve(sizeof(float) * n_tgt_kps));
        CK(cudaMemcpyAsync(ctx->ws_thr[0].p, thr_src, sizeof(float) * n_src_kps, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->ws_thr[1].p, thr_tgt, sizeof(float) * n_tgt_kps, cudaMemcpyHostToDevice, st));
        d_thr_s = ctx->ws_thr[0].as<float>();
        d_thr_t = ctx->ws_thr[1].as<float>();
    }
    CK(ctx->ws_misc.reserve(64));
    float *d_avg = ctx->ws_misc.as<float>();
    unsigned long long *d_n = reinterpret_cast<unsigned long long *>(ctx->ws_misc.as<char>() + 16);
    CK(ctx->ws_corr.reserve(sizeof(b200m_corr) * n_src_kps));
    b200m_params p1 = *p;
    p1.k = 1;   // the voted lists hold at most one match
    if (mode == B200M_MODE_CLUSTER) {
        if (cluster_k < 1) return b200m_fail_msg(ctx, "b200m_match_multiscale: cluster_k must be in [1, 64]");
        if (b200m_cluster_filter_device(ctx, &p1, cluster_k, 0.95f /* MATCHING_CLUSTER_THRESHOLD, include/common.h:52 */, n_src_kps,
                                        n_tgt_kps, ws->vfi.as<int32_t>(), ws->vfd.as<float>(), ws->vfc.as<int32_t>(),
                                        ws->vri.as<int32_t>(), ws->vrc.as<int32_t>(), ws->xyz_s.as<float>(), ws->xyz_t.as<float>(),
                                        xyz_stride_bytes, d_thr_s, d_thr_t, ctx->ws_corr.as<b200m_corr>(), n_src_kps, d_n,
                                        avg_first_dist ? d_avg : nullptr))
            return 1;
        if (n_tgt_kps == 0 && avg_first_dist)
            CK(launch_average(ws->vfd.as<float>(), ws->vfc.as<int32_t>(), n_src_kps, 1, d_avg, st));
    } else {
        if (b200m_filter_device(ctx, &p1, 0, n_src_kps, ws->vfi.as<int32_t>(), ws->vfd.as<float>(), ws->vfc.as<int32_t>(),
                                ws->vri.as<int32_t>(), ws->vrd.as<float>(), ws->vrc.as<int32_t>(),
                                mode == B200M_MODE_MUTUAL ? n_tgt_kps : 0, d_thr_s, d_thr_t, ctx->ws_corr.as<b200m_corr>(), n_src_kps,
                                d_n, avg_first_dist ? d_avg : nullptr))
            return 1;
    }
    struct { float avg; float pad[3]; unsigned long long n; int bad_f, bad_r; } h{};
    CK(cudaMemcpyAsync(&h, ctx->ws_misc.p, 24, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h.bad_f, ws->fwd.bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (two_way) CK(cudaMemcpyAsync(&h.bad_r, ws->rev.bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h.bad_f || h.bad_r)
        return b200m_fail_msg(ctx, "b200m_match_multiscale: an index map entry is outside its keypoint range");
    if (avg_first_dist) *avg_first_dist = h.avg;
    *n_out = (size_t) h.n;
    if (h.n > cap) return b200m_fail_msg(ctx, "b200m_match_multiscale: output capacity too small (" + std::to_string(h.n) + " correspondences)");
    if (h.n) {
        if (!out) return b200m_fail_msg(ctx, "b200m_match_multiscale: null output buffer");
        CK(cudaMemcpyAsync(out, ctx->ws_corr.p, sizeof(b200m_corr) * h.n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return 0;
}
