"""The oracle pinned against the reference's own known-answer vectors and against
frozen cv2.BFMatcher outputs (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

CASES = ["fpfh_k1", "fpfh_k2_blocked", "fpfh_k5", "rops_k3", "shot_k2", "shot_k1_blocked"]


def test_knn_result_known_answers():
    """Replays tests/knn_result.cpp:28-49 of the reference verbatim."""
    r = orc.KNNResult(3)
    assert (r.indices(), r.distances()) == ([], [])
    r.add_point(3.0, 3)
    assert (r.indices(), r.distances()) == ([3], [3.0])
    r.add_point(2.0, 2)
    assert (r.indices(), r.distances()) == ([2, 3], [2.0, 3.0])
    r.add_point(4.0, 4)
    assert (r.indices(), r.distances()) == ([2, 3, 4], [2.0, 3.0, 4.0])
    r.add_point(1.0, 1)
    assert (r.indices(), r.distances()) == ([1, 2, 3], [1.0, 2.0, 3.0])
    r.add_point(1.0, 5)
    assert (r.indices(), r.distances()) == ([1, 5, 2], [1.0, 1.0, 2.0])


def test_update_multivalued_tie_goes_first():
    """src/common.cpp:520 -- `while (dist[pos] < distance)`: an equal distance is inserted BEFORE."""
    i, d = orc.update_multivalued([], [], 2, 7, 1.0)
    i, d = orc.update_multivalued(i, d, 2, 9, 1.0)
    assert i == [9, 7]
    i, d = orc.update_multivalued(i, d, 2, 3, 0.5)
    assert i == [3, 9] and d == [0.5, 1.0]


@pytest.mark.parametrize("name", CASES)
def test_match_bf_equals_cv2(golden_dir, name):
    """orc.match_bf == the reference's matchBF loop run over real cv2.BFMatcher: indices exact,
    distances within 1e-6 relative (OpenCV sums in SIMD lane order), both directions --
    the assertion form of tests/flann_bf_matcher.h:73-88."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    dim, k, block = int(g["dim"]), int(g["k"]), int(g["block"])
    src, tgt = g["src"], g["tgt"]
    for (q, t, gi, gd, gc) in ((src, tgt, g["bf_idx"], g["bf_dist"], g["bf_cnt"]),
                               (tgt, src, g["bf_ridx"], g["bf_rdist"], g["bf_rcnt"])):
        idx, dist, cnt = orc.match_bf(q[:, :dim], t[:, :dim], k, block)
        assert np.array_equal(cnt, gc)
        assert np.array_equal(idx, gi)
        np.testing.assert_allclose(dist, gd, rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", CASES)
def test_bf_flann_local_agree(golden_dir, name):
    """BF == FLANN == Local(inf) indices (tests/flann_bf_matcher.h:73-76): the canonical exact
    kNN (matchLocal/matchFLANN restatement) equals the blocked matchBF restatement and cv2."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    dim, k = int(g["dim"]), int(g["k"])
    src, tgt = g["src"], g["tgt"]
    idx, dist, cnt = orc.knn(src[:, :dim], tgt[:, :dim], k)
    assert np.array_equal(idx, g["bf_idx"]) and np.array_equal(cnt, g["bf_cnt"])
    ridx, rdist, rcnt = orc.knn(tgt[:, :dim], src[:, :dim], k)
    assert np.array_equal(ridx, g["bf_ridx"]) and np.array_equal(rcnt, g["bf_rcnt"])
    # ascending, sqrt'ed L2
    for i in range(idx.shape[0]):
        assert np.all(np.diff(dist[i, :cnt[i]]) >= 0)


def test_invalid_rows(golden_dir):
    g = np.load(os.path.join(golden_dir, "fpfh_k2_blocked.npz"))
    dim, k = int(g["dim"]), int(g["k"])
    src, tgt = g["src"][:, :dim], g["tgt"][:, :dim]
    idx, dist, cnt = orc.knn(src, tgt, k)
    bad_q = ~np.isfinite(src).all(1)
    bad_t = np.nonzero(~np.isfinite(tgt).all(1))[0]
    assert bad_q.any() and len(bad_t)
    assert np.all(cnt[bad_q] == 0) and np.all(idx[bad_q] == -1)      # include/matching.h:576
    assert np.all(cnt[~bad_q] == k)
    assert not np.isin(idx, bad_t).any()                                # never a candidate (:661)


def test_ties_and_k_larger_than_train(golden_dir):
    g = np.load(os.path.join(golden_dir, "ties_small.npz"))
    q, t = g["q"], g["t"]
    # matchBF: merge reverses exact ties (documented quirk) -- equals cv2-driven loop
    idx, dist, cnt = orc.match_bf(q, t, 3, 10000)
    assert np.array_equal(idx, g["idx"]) and idx[0].tolist() == [30, 17, 5]
    # canonical (KNNResult / matchLocal / FLANN order): lower train index first
    idx, dist, cnt = orc.knn(q, t, 3)
    assert idx[0].tolist() == [5, 17, 30] and np.all(dist[0] == 0)
    assert np.array_equal(np.sort(idx, 1), np.sort(g["idx"], 1))
    # k > nt -> nt results
    idx, dist, cnt = orc.knn(q, t[:2], 3)
    assert cnt.tolist() == g["cnt_kgt"].tolist() == [2, 2]
    assert np.array_equal(idx, g["idx_kgt"])


def test_strided_aos_rows(golden_dir):
    """SHOT352 AoS rows (1444 B stride, rf[9] junk tail) give the same result as dense rows."""
    g = np.load(os.path.join(golden_dir, "shot_k2.npz"))
    dim, k = int(g["dim"]), int(g["k"])
    src, tgt = g["src"], g["tgt"]
    assert src.shape[1] == 361
    a = orc.knn(src[:, :dim], tgt[:, :dim], k)
    b = orc.knn(np.ascontiguousarray(src[:, :dim]), np.ascontiguousarray(tgt[:, :dim]), k)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_l2_norm_is_sequential_fp32():
    rng = np.random.default_rng(0)
    a = rng.random(352).astype(np.float32)
    b = rng.random(352).astype(np.float32)
    s = np.float32(0)
    for d in range(352):
        diff = np.float32(a[d] - b[d])
        s = np.float32(s + np.float32(diff * diff))
    assert orc.l2_norm(a, b) == float(np.sqrt(s, dtype=np.float32))


def _brute_lists(q, t, k):
    d = np.sqrt(((q[:, None, :].astype(np.float64) - t[None, :, :].astype(np.float64)) ** 2).sum(-1))
    return np.argsort(d, axis=1, kind="stable")[:, :k]


def test_filters_against_python_restatement(golden_dir):
    g = np.load(os.path.join(golden_dir, "fpfh_k5.npz"))
    dim, k = int(g["dim"]), int(g["k"])
    src, tgt = g["src"][:, :dim], g["tgt"][:, :dim]
    fi, fd, fc = orc.knn(src, tgt, k)
    ri, rd, rc = orc.knn(tgt, src, k)
    rng = np.random.default_rng(1)
    thr_q = rng.random(src.shape[0]).astype(np.float32)
    thr_t = rng.random(tgt.shape[0]).astype(np.float32)
    dthr = np.float32(0.6)
    # mutual (include/matching.h:437-449) in plain python
    exp = []
    for i in range(src.shape[0]):
        for j in fi[i, :fc[i]]:
            for m in range(rc[j]):
                if ri[j, m] == i:
                    exp.append((i, j, rd[j, m], min(max(thr_q[i], thr_t[j]), dthr)))
                    break
    got = orc.filter_mutual(fi, fc, ri, rd, rc, dthr, thr_q, thr_t)
    assert len(got) == len(exp) and len(exp) > 10
    assert [tuple(x) for x in got.tolist()] == [(int(a), int(b), float(c), float(d)) for a, b, c, d in exp]
    assert np.all(np.diff(got["index_query"]) >= 0)
    # one-sided (:399-407)
    got = orc.filter_one_sided(fi, fd, fc, dthr, thr_q, thr_t)
    exp = [(i, int(fi[i, 0]), float(fd[i, 0]), float(min(max(thr_q[i], thr_t[fi[i, 0]]), dthr)))
           for i in range(src.shape[0]) if fc[i] > 0]
    assert [tuple(x) for x in got.tolist()] == exp
    # ratio (defined, unpinned): d2 >= 1.1f * d1
    got = orc.filter_ratio(fi, fd, fc, 1.1, dthr)
    keep = [i for i in range(src.shape[0]) if fc[i] >= 2 and fd[i, 1] >= np.float32(1.1) * fd[i, 0]]
    assert got["index_query"].tolist() == keep and 0 < len(keep) < src.shape[0]
    assert np.all(got["threshold"] == dthr)
    # average first-NN distance (src/matching.cpp:3-19)
    s = np.float32(0)
    n = 0
    for i in range(src.shape[0]):
        if fc[i] > 0:
            s = np.float32(s + fd[i, 0]); n += 1
    assert orc.average_distance(fd, fc) == float(np.float32(s / np.float32(n)))
    # finalize (:356-362)
    ks = rng.permutation(10 * src.shape[0])[:src.shape[0]].astype(np.int32)
    kt = rng.permutation(10 * tgt.shape[0])[:tgt.shape[0]].astype(np.int32)
    fin = orc.finalize(got, ks, kt)
    assert np.array_equal(fin["index_query"], ks[got["index_query"]])
    assert np.array_equal(fin["index_match"], kt[got["index_match"]])


def test_spatial_vote_single_candidate_is_identity():
    """With one scale and k=1 the vote is the identity (include/matching.h:327-352)."""
    rng = np.random.default_rng(2)
    idx = rng.integers(0, 50, size=(20, 1)).astype(np.int32)
    dist = rng.random((20, 1)).astype(np.float32)
    cnt = np.ones(20, np.int32); cnt[3] = 0; idx[3] = -1
    xyz = rng.random((50, 3)).astype(np.float32)
    i2, d2, c2 = orc.spatial_vote(idx, dist, cnt, xyz, 0.05)
    assert np.array_equal(c2, cnt) and np.array_equal(i2[cnt > 0], idx[cnt > 0]) and np.array_equal(d2[cnt > 0], dist[cnt > 0])


def test_spatial_vote_prefers_clustered_candidates():
    xyz = np.array([[0, 0, 0], [0.01, 0, 0], [0.02, 0, 0], [5, 5, 5]], np.float32)
    # candidate list: far point first (best descriptor), then three mutually-close points
    idx = np.array([[3, 0, 1, 2]], np.int32)
    dist = np.array([[0.1, 0.2, 0.3, 0.4]], np.float32)
    cnt = np.array([4], np.int32)
    i2, d2, c2 = orc.spatial_vote(idx, dist, cnt, xyz, 0.05)
    assert c2[0] == 1 and i2[0, 0] == 0 and d2[0, 0] == np.float32(0.2)


@pytest.mark.parametrize("desc,nq,nt,k", [("fpfh", 1000, 2051, 5), ("shot", 300, 1037, 2), ("rops", 200, 515, 3)])
def test_simd_layout_is_bit_identical_to_scalar(desc, nq, nt, k):
    from lidar_global_registration_b200 import synth
    s, t, d = synth.make_pair(desc, nq, nt, nan_frac=0.01)
    a = orc.knn(s[:, :d], t[:, :d], k)
    b = orc.knn(s[:, :d], t[:, :d], k, scalar=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_cluster_filter_matches_a_literal_restatement():
    """orc_knn3d / orc_filter_cluster against ClusterMatcher::match_impl + calculateCorrespondenceDistance written out
    with Python sets, line by line (reference include/matching.h:492-550)."""
    rng = np.random.default_rng(9)
    nq, nt, k, ck = 120, 140, 2, 7
    sx, tx = rng.random((nq, 3)).astype(np.float32), rng.random((nt, 3)).astype(np.float32)
    src, tgt = rng.random((nq, 8)).astype(np.float32), rng.random((nt, 8)).astype(np.float32)
    fidx, fdist, fcnt = orc.knn(src, tgt, k)
    ridx, rdist, rcnt = orc.knn(tgt, src, k)
    ns, ng = orc.knn3d(sx, ck), orc.knn3d(tx, ck)

    def knn_set(p, i):          # nearestKSearch(i, ck): the ck nearest by FLANN's L2_Simple, the point itself included
        d = np.zeros(p.shape[0], np.float32)
        for c in range(3):
            diff = p[i, c] - p[:, c]
            d = d + diff * diff
        return set(np.lexsort((np.arange(p.shape[0]), d))[:ck].tolist())
    for i in (0, 17, nq - 1):
        assert set(ns[i].tolist()) == knn_set(sx, i)

    def corr_distance(i, j, lists, counts, p_a, p_b):
        i_nb, j_nb = knn_set(p_a, i), knn_set(p_b, j)
        consistent = pairs = 0
        for a in i_nb:
            for m in range(counts[a]):
                if int(lists[a, m]) in j_nb:
                    consistent += 1
                pairs += 1
        return np.float32(0) if pairs == 0 else np.float32(1) - np.float32(consistent) / np.float32(pairs)
    exp = []
    for i in range(nq):
        for a in range(fcnt[i]):
            j = int(fidx[i, a])
            di, dj = corr_distance(i, j, fidx, fcnt, sx, tx), corr_distance(j, i, ridx, rcnt, tx, sx)
            if di < np.float32(0.95) and dj < np.float32(0.95):
                exp.append((i, j, max(di, dj)))
    got = orc.filter_cluster(fidx, fcnt, ridx, rcnt, ns, ng, np.float32(1.0))
    assert [(int(c["index_query"]), int(c["index_match"]), np.float32(c["distance"])) for c in got] == exp
    assert 0 < len(exp) < fcnt.sum()
