"""ctypes front-end of oracle/liboracle.so (the CPU restatement of the reference matcher).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(lidar_global_registration_b200) never imports this module.

Every wrapper names the reference lines the C function restates; see
oracle/oracle.c for the statement of the arithmetic and the pin status.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

CORR_DTYPE = np.dtype([("index_query", "<i4"), ("index_match", "<i4"),
                       ("distance", "<f4"), ("threshold", "<f4")])


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc, OpenMP); a library whose stamp
    matches the current source is kept."""
    import hashlib
    src = os.path.join(_HERE, "oracle.c")
    stamp = _LIB_PATH + ".stamp"
    digest = hashlib.sha256(open(src, "rb").read() + open(os.path.join(_HERE, "Makefile"), "rb").read()).hexdigest()
    if (not force and os.path.exists(_LIB_PATH) and os.path.exists(stamp)
            and open(stamp).read().strip() == digest):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                   stdout=subprocess.DEVNULL)
    with open(stamp, "w") as f:
        f.write(digest + "\n")
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int32)
        sz = C.c_size_t
        _lib.orc_knn.argtypes = [fp, sz, sz, fp, sz, sz, C.c_int, C.c_int, ip, fp, ip]
        _lib.orc_knn.restype = None
        _lib.orc_knn_scalar.argtypes = _lib.orc_knn.argtypes
        _lib.orc_knn_scalar.restype = None
        _lib.orc_match_bf.argtypes = [fp, sz, sz, fp, sz, sz, C.c_int, C.c_int, C.c_int, ip, fp, ip]
        _lib.orc_match_bf.restype = None
        _lib.orc_l2_norm.argtypes = [fp, fp, C.c_int]
        _lib.orc_l2_norm.restype = C.c_float
        _lib.orc_is_valid.argtypes = [fp, C.c_int]
        _lib.orc_is_valid.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_update_multivalued.argtypes = [ip, fp, C.POINTER(C.c_int), C.c_int, C.c_int32, C.c_float]
        _lib.orc_update_multivalued.restype = None
        _lib.orc_spatial_vote.argtypes = [sz, C.c_int, ip, fp, ip, fp, C.c_float]
        _lib.orc_spatial_vote.restype = None
        vp = C.c_void_p
        _lib.orc_filter_one_sided.argtypes = [sz, C.c_int, ip, fp, ip, fp, fp, C.c_float, vp]
        _lib.orc_filter_one_sided.restype = sz
        _lib.orc_filter_mutual.argtypes = [sz, C.c_int, ip, ip, sz, ip, fp, ip, fp, fp, C.c_float, vp]
        _lib.orc_filter_mutual.restype = sz
        _lib.orc_filter_ratio.argtypes = [sz, C.c_int, ip, fp, ip, C.c_float, fp, fp, C.c_float, vp]
        _lib.orc_filter_ratio.restype = sz
        _lib.orc_match_local.argtypes = [fp, sz, sz, fp, sz, sz, C.c_int, C.c_int, fp, fp, C.c_float, ip, fp, ip]
        _lib.orc_match_local.restype = None
        _lib.orc_knn3d.argtypes = [sz, fp, C.c_int, ip]
        _lib.orc_knn3d.restype = None
        _lib.orc_filter_cluster.argtypes = [sz, C.c_int, ip, ip, sz, ip, ip, C.c_int, ip, ip, C.c_float, fp, fp, C.c_float, vp]
        _lib.orc_filter_cluster.restype = sz
        _lib.orc_average_distance.argtypes = [sz, C.c_int, fp, ip]
        _lib.orc_average_distance.restype = C.c_float
        _lib.orc_finalize.argtypes = [vp, sz, ip, ip]
        _lib.orc_finalize.restype = None
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _rows(a):
    """(base pointer array, n rows, row stride in bytes) of a 2-D float32 array whose
    rows are contiguous -- rows may be strided (AoS point structs, e.g. SHOT352's 1444 B)."""
    assert a.dtype == np.float32 and a.ndim == 2 and (a.shape[1] == 0 or a.strides[1] == 4)
    return a, a.shape[0], a.strides[0] if a.shape[0] > 1 else max(a.strides[0], 4 * a.shape[1])


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads.argtypes = [C.c_int]
    lib().orc_set_num_threads.restype = None
    lib().orc_set_num_threads(int(n))


def knn(query, train, k, scalar=False):
    """Canonical exact kNN == matchLocal(radius=inf) == matchFLANN result set
    (include/matching.h:637-678, :562-592).  Returns (idx[nq,k] int32 -1 padded,
    dist[nq,k] float32, count[nq] int32).  scalar=True runs the plain loop nest
    (orc_knn_scalar); the default is the SIMD-blocked layout of the same arithmetic."""
    q, nq, qs = _rows(query)
    t, nt, ts = _rows(train)
    dim = query.shape[1]
    idx = np.empty((nq, k), np.int32)
    dist = np.empty((nq, k), np.float32)
    cnt = np.empty((nq,), np.int32)
    (lib().orc_knn_scalar if scalar else lib().orc_knn)(_fp(q), nq, qs, _fp(t), nt, ts, dim, k, _ip(idx), _fp(dist), _ip(cnt))
    return idx, dist, cnt


def match_bf(query, train, k, block_size=10000):
    """matchBF with its blocking and cross-block merge quirk (include/matching.h:594-634,
    src/common.cpp:517-529)."""
    q, nq, qs = _rows(query)
    t, nt, ts = _rows(train)
    dim = query.shape[1]
    idx = np.empty((nq, k), np.int32)
    dist = np.empty((nq, k), np.float32)
    cnt = np.empty((nq,), np.int32)
    lib().orc_match_bf(_fp(q), nq, qs, _fp(t), nt, ts, dim, k, block_size, _ip(idx), _fp(dist), _ip(cnt))
    return idx, dist, cnt


def l2_norm(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return float(lib().orc_l2_norm(_fp(a), _fp(b), a.shape[0]))


class KNNResult:
    """KNNResult<float> (include/matching.h:44-94)."""

    class _S(C.Structure):
        _fields_ = [("capacity", C.c_int), ("count", C.c_int),
                    ("indices", C.POINTER(C.c_int32)), ("dists", C.POINTER(C.c_float))]

    def __init__(self, capacity):
        self._idx = np.zeros(capacity, np.int32)
        self._dist = np.zeros(capacity, np.float32)
        self._s = self._S()
        L = lib()
        L.orc_knn_result_init.argtypes = [C.POINTER(self._S), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        L.orc_knn_result_add.argtypes = [C.POINTER(self._S), C.c_float, C.c_int32]
        L.orc_knn_result_init(C.byref(self._s), capacity, _ip(self._idx), _fp(self._dist))

    def add_point(self, dist, index):
        lib().orc_knn_result_add(C.byref(self._s), dist, index)

    def size(self):
        return self._s.count

    def indices(self):
        return self._idx[:self._s.count].tolist()

    def distances(self):
        return self._dist[:self._s.count].tolist()


def update_multivalued(idx_list, dist_list, k, match_idx, distance):
    """updateMultivaluedCorrespondence (src/common.cpp:517-529) on python lists."""
    idx = np.zeros(k + 1, np.int32)
    dist = np.zeros(k + 1, np.float32)
    n = len(idx_list)
    idx[:n] = idx_list
    dist[:n] = dist_list
    cnt = C.c_int(n)
    lib().orc_update_multivalued(_ip(idx), _fp(dist), C.byref(cnt), k, match_idx, distance)
    return idx[:cnt.value].tolist(), dist[:cnt.value].tolist()


def spatial_vote(idx, dist, count, train_xyz, iss_radius):
    """match_multiscale's vote (include/matching.h:327-352); returns new arrays."""
    idx = np.ascontiguousarray(idx.copy()); dist = np.ascontiguousarray(dist.copy()); count = count.copy()
    xyz = np.ascontiguousarray(train_xyz, np.float32)
    lib().orc_spatial_vote(idx.shape[0], idx.shape[1], _ip(idx), _fp(dist), _ip(count), _fp(xyz), iss_radius)
    return idx, dist, count


def _thr(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def filter_one_sided(fidx, fdist, fcount, distance_thr, thr_q=None, thr_t=None):
    """OneSidedMatcher::match_impl (include/matching.h:395-411)."""
    nq, k = fidx.shape
    out = np.empty(nq, CORR_DTYPE)
    tq, tt = _thr(thr_q), _thr(thr_t)
    n = lib().orc_filter_one_sided(nq, k, _ip(fidx), _fp(fdist), _ip(fcount), _fp(tq), _fp(tt),
                                   distance_thr, out.ctypes.data)
    return out[:n].copy()


def filter_mutual(fidx, fcount, ridx, rdist, rcount, distance_thr, thr_q=None, thr_t=None):
    """LeftToRightMatcher::match_impl, k-list form (include/matching.h:428-453)."""
    nq, k = fidx.shape
    nt = ridx.shape[0]
    out = np.empty(nq * k, CORR_DTYPE)
    tq, tt = _thr(thr_q), _thr(thr_t)
    n = lib().orc_filter_mutual(nq, k, _ip(fidx), _ip(fcount), nt, _ip(ridx), _fp(rdist), _ip(rcount),
                                _fp(tq), _fp(tt), distance_thr, out.ctypes.data)
    return out[:n].copy()


def filter_ratio(fidx, fdist, fcount, ratio_thr, distance_thr, thr_q=None, thr_t=None):
    """Ratio filter as DEFINED in SURVEY 8a (reference stub include/matching.h:470-473) -- parity unpinned."""
    nq, k = fidx.shape
    out = np.empty(nq, CORR_DTYPE)
    tq, tt = _thr(thr_q), _thr(thr_t)
    n = lib().orc_filter_ratio(nq, k, _ip(fidx), _fp(fdist), _ip(fcount), ratio_thr, _fp(tq), _fp(tt),
                               distance_thr, out.ctypes.data)
    return out[:n].copy()


def match_local(query, train, k, query_xyz, train_xyz, radius):
    """matchLocal with a finite match_search_radius (include/matching.h:637-678); query_xyz are the query keypoints
    AFTER the guess transform.  Returns (idx, dist, count) like knn()."""
    q, _, qs = _rows(query)
    t, _, ts = _rows(train)
    qx = np.ascontiguousarray(np.asarray(query_xyz, np.float32)[:, :3])
    tx = np.ascontiguousarray(np.asarray(train_xyz, np.float32)[:, :3])
    assert qx.shape[0] == q.shape[0] and tx.shape[0] == t.shape[0]
    dim = q.shape[1]
    idx = np.empty((q.shape[0], k), np.int32)
    dist = np.empty((q.shape[0], k), np.float32)
    cnt = np.empty((q.shape[0],), np.int32)
    lib().orc_match_local(_fp(q), q.shape[0], qs, _fp(t), t.shape[0], ts, dim, k, _fp(qx), _fp(tx), radius, _ip(idx),
                          _fp(dist), _ip(cnt))
    return idx, dist, cnt


def knn3d(xyz, k):
    """3-D neighbourhoods of ClusterMatcher (pcl KdTree nearestKSearch(index, k), include/matching.h:524-528):
    [n, k] int32, the point itself included, -1 padded when n < k."""
    p = np.ascontiguousarray(np.asarray(xyz, np.float32)[:, :3])
    if not 1 <= k <= 64:
        raise ValueError("k must be in [1, 64]")
    out = np.empty((p.shape[0], k), np.int32)
    lib().orc_knn3d(p.shape[0], _fp(p), k, _ip(out))
    return out


def filter_cluster(fidx, fcount, ridx, rcount, nbr_src, nbr_tgt, distance_thr, cluster_thr=np.float32(0.95), thr_q=None,
                   thr_t=None):
    """ClusterMatcher::match_impl, k-list form (include/matching.h:492-517); MATCHING_CLUSTER_THRESHOLD = 0.95f."""
    nq, k = fidx.shape
    nt = ridx.shape[0]
    out = np.empty(max(nq * k, 1), CORR_DTYPE)
    tq, tt = _thr(thr_q), _thr(thr_t)
    ns, ng = np.ascontiguousarray(nbr_src, np.int32), np.ascontiguousarray(nbr_tgt, np.int32)
    n = lib().orc_filter_cluster(nq, k, _ip(fidx), _ip(fcount), nt, _ip(ridx), _ip(rcount), ns.shape[1], _ip(ns), _ip(ng),
                                 cluster_thr, _fp(tq), _fp(tt), distance_thr, out.ctypes.data)
    return out[:n].copy()


def average_distance(fdist, fcount):
    """FeatureBasedMatcher::printDebugInfo (src/matching.cpp:3-19)."""
    nq, k = fdist.shape
    return float(lib().orc_average_distance(nq, k, _fp(fdist), _ip(fcount)))


def finalize(corrs, kps_src, kps_tgt):
    """FeatureBasedMatcherImpl::finalize (include/matching.h:356-362); returns a new array."""
    out = corrs.copy()
    ks = np.ascontiguousarray(kps_src, np.int32)
    kt = np.ascontiguousarray(kps_tgt, np.int32)
    lib().orc_finalize(out.ctypes.data, out.shape[0], _ip(ks), _ip(kt))
    return out


def match_multiscale(query_scales, train_scales, n_query_kps, train_xyz, iss_radius, k):
    """FeatureBasedMatcherImpl::match_multiscale (include/matching.h:264-354): per scale an exact kNN (k = randomness) over
    that scale's descriptor rows, row numbers renamed to keypoint ids through kps_indices_multiscale (:313-321), candidates
    concatenated per query keypoint in scale order, then the spatial vote (:327-352).  `*_scales`: lists of
    (descriptors [n_s, dim], row -> keypoint id map [n_s] or None).  Returns one-entry lists (idx [n,1], dist [n,1], count [n])."""
    comb_i = [[] for _ in range(n_query_kps)]
    comb_d = [[] for _ in range(n_query_kps)]
    for (qf, qmap), (tf, tmap) in zip(query_scales, train_scales):
        if qf.shape[0] == 0:
            continue
        ei, ed, ec = knn(qf, tf, k)
        for r in range(qf.shape[0]):
            kp = r if qmap is None else int(qmap[r])
            for m in range(ec[r]):
                j = int(ei[r, m])
                comb_i[kp].append(j if tmap is None else int(tmap[j]))
                comb_d[kp].append(ed[r, m])
    width = max(max((len(c) for c in comb_i), default=0), 1)
    ci = np.full((n_query_kps, width), -1, np.int32)
    cd = np.zeros((n_query_kps, width), np.float32)
    cc = np.zeros(n_query_kps, np.int32)
    for i in range(n_query_kps):
        cc[i] = len(comb_i[i])
        ci[i, :cc[i]] = comb_i[i]
        cd[i, :cc[i]] = comb_d[i]
    xyz = np.ascontiguousarray(np.asarray(train_xyz, np.float32)[:, :3])
    if xyz.shape[0] == 0:
        return ci[:, :1].copy(), cd[:, :1].copy(), np.zeros(n_query_kps, np.int32)
    vi, vd, vc = spatial_vote(ci, cd, cc, xyz, iss_radius)
    return np.ascontiguousarray(vi[:, :1]), np.ascontiguousarray(vd[:, :1]), vc


def match_wide(mode, src_scales, tgt_scales, src_xyz, tgt_xyz, iss_radius_src, iss_radius_tgt, k, cluster_k=40,
               distance_thr=np.float32(3.4e38), thr_q=None, thr_t=None):
    """match_impl of OneSidedMatcher / LeftToRightMatcher / ClusterMatcher (include/matching.h:395-411, :428-453, :492-517)
    as the reference composes it: match_multiscale both ways (the reverse call is the reference's inverse_tn one), the
    average over the voted forward lists (printDebugInfo), then the matcher's loop over the voted lists.
    mode in {'one_sided', 'mutual', 'cluster'}.  Returns (corrs, avg_first_distance)."""
    n_s, n_t = np.asarray(src_xyz).shape[0], np.asarray(tgt_xyz).shape[0]
    fi, fd, fc = match_multiscale(src_scales, tgt_scales, n_s, tgt_xyz, iss_radius_tgt, k)
    avg = average_distance(fd, fc)
    if mode == "one_sided":
        return filter_one_sided(fi, fd, fc, distance_thr, thr_q, thr_t), avg
    ri, rd, rc = match_multiscale(tgt_scales, src_scales, n_t, src_xyz, iss_radius_src, k)
    if mode == "mutual":
        return filter_mutual(fi, fc, ri, rd, rc, distance_thr, thr_q, thr_t), avg
    if mode == "cluster":
        if n_s == 0 or n_t == 0:
            return np.empty(0, CORR_DTYPE), avg
        return filter_cluster(fi, fc, ri, rc, knn3d(src_xyz, cluster_k), knn3d(tgt_xyz, cluster_k), distance_thr,
                              thr_q=thr_q, thr_t=thr_t), avg
    raise ValueError(mode)


def match(query, train, k, mode, ratio_thr=1.1, distance_thr=np.float32(3.4e38), thr_q=None, thr_t=None):
    """Whole matcher call at the k-list seam: mode in {'one_sided','mutual','ratio'}.
    Returns (corrs, avg_first_distance)."""
    fidx, fdist, fcnt = knn(query, train, k)
    avg = average_distance(fdist, fcnt)
    if mode == "one_sided":
        return filter_one_sided(fidx, fdist, fcnt, distance_thr, thr_q, thr_t), avg
    if mode == "ratio":
        return filter_ratio(fidx, fdist, fcnt, ratio_thr, distance_thr, thr_q, thr_t), avg
    if mode == "mutual":
        ridx, rdist, rcnt = knn(train, query, k)
        return filter_mutual(fidx, fcnt, ridx, rdist, rcnt, distance_thr, thr_q, thr_t), avg
    raise ValueError(mode)
