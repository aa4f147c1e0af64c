"""Parity of the CUDA path against the CPU oracle, through the C-ABI (run with -m gpu on a B200).

Bar (BASELINE.json north_star): neighbour indices bit-exact, distances bit-exact too (the re-rank
runs the reference's own sequential FP32 arithmetic), for both the exact CUDA-core path and the
tensor-core candidate path; filters bit-exact record for record."""
import os

import numpy as np
import pytest

from lidar_global_registration_b200 import build as b200_build
from lidar_global_registration_b200 import matcher as M
from lidar_global_registration_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

CASES = ["fpfh_k1", "fpfh_k2_blocked", "fpfh_k5", "rops_k3", "shot_k2", "shot_k1_blocked"]
PRECS = [M.PREC_F32_EXACT, M.PREC_TC_F16]


@pytest.fixture(scope="module", autouse=True)
def _built():
    b200_build.build()
    M.load_library()


def _same(a, b):
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and x.shape == y.shape
        assert np.array_equal(x, y)


def _dense(a, dim):
    return a[:, :dim]


@pytest.mark.parametrize("prec", PRECS, ids=["exact", "tc"])
@pytest.mark.parametrize("name", CASES)
def test_knn_equals_oracle_on_golden_inputs(golden_dir, name, prec):
    """GPU k-lists == oracle k-lists (idx, dist, count) bit for bit, both directions; and == the frozen
    cv2.BFMatcher indices (the assertion of the reference's tests/flann_bf_matcher.h:73-88)."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    dim, k = int(g["dim"]), int(g["k"])
    src, tgt = g["src"], g["tgt"]
    with M.Context(0) as ctx:
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        fwd = ctx.knn(k, 0, precision=prec)
        rev = ctx.knn(k, 1, precision=prec)
    _same(fwd, orc.knn(_dense(src, dim), _dense(tgt, dim), k))
    _same(rev, orc.knn(_dense(tgt, dim), _dense(src, dim), k))
    if int(g["block"]) >= max(src.shape[0], tgt.shape[0]):   # single block: no cross-block tie reversal possible
        assert np.array_equal(fwd[0], g["bf_idx"]) and np.array_equal(rev[0], g["bf_ridx"])
    np.testing.assert_allclose(fwd[1], g["bf_dist"], rtol=1e-5, atol=0)   # the north star's FP32 tolerance vs cv2


@pytest.mark.parametrize("prec", PRECS, ids=["exact", "tc"])
def test_ties_kgt_nt_and_empty(golden_dir, prec):
    g = np.load(os.path.join(golden_dir, "ties_small.npz"))
    q, t = g["q"], g["t"]
    with M.Context(0) as ctx:
        ctx.upload(0, q, 33)
        ctx.upload(1, t, 33)
        idx, dist, cnt = ctx.knn(3, 0, precision=prec)
        assert idx[0].tolist() == [5, 17, 30] and np.all(dist[0] == 0)      # canonical: lower index first
        _same((idx, dist, cnt), orc.knn(q, t, 3))
        ctx.upload(1, t[:2], 33)                                            # k > nt -> nt results (cv2: idx_kgt)
        idx, dist, cnt = ctx.knn(3, 0, precision=prec)
        assert cnt.tolist() == [2, 2] and np.array_equal(idx, g["idx_kgt"])
        _same((idx, dist, cnt), orc.knn(q, t[:2], 3))
        ctx.upload(1, t[:0], 33)                                            # empty train set -> empty lists
        idx, dist, cnt = ctx.knn(2, 0, precision=prec)
        assert np.all(cnt == 0) and np.all(idx == -1)
        ctx.upload(0, q[:0], 33)                                            # empty query set
        idx, dist, cnt = ctx.knn(2, 0, precision=prec)
        assert idx.shape == (0, 2)
        corrs, avg = ctx.match(1, M.MODE_ONE_SIDED, precision=prec)
        assert len(corrs) == 0 and avg == M.FLT_MAX


@pytest.mark.parametrize("prec", PRECS, ids=["exact", "tc"])
def test_all_rows_invalid_or_duplicated(prec):
    rng = np.random.default_rng(3)
    t = rng.random((300, 33)).astype(np.float32)
    q = t[:40].copy()
    q[7, 3] = np.nan
    t[11, 0] = np.inf
    t[100:200] = t[0]          # 101 exact duplicates of row 0: a long run of exact ties
    with M.Context(0) as ctx:
        ctx.upload(0, q, 33)
        ctx.upload(1, t, 33)
        for k in (1, 4, 8):
            _same(ctx.knn(k, 0, precision=prec), orc.knn(q, t, k))
        bad = np.full((5, 33), np.nan, np.float32)
        ctx.upload(1, bad, 33)     # every train row invalid -> every list empty
        idx, dist, cnt = ctx.knn(2, 0, precision=prec)
        assert np.all(cnt == 0)
        _same((idx, dist, cnt), orc.knn(q, bad, 2))


@pytest.mark.parametrize("desc,nq,nt,k", [("fpfh", 777, 1500, 2), ("shot", 300, 900, 2), ("rops", 260, 700, 3),
                                          ("fpfh", 130, 5000, 16)])
def test_tc_operands_and_accumulators(desc, nq, nt, k):
    """White-box check of the tensor-core pass: operand tiles as documented in pack.cu and raw
    tcgen05 accumulators == |b16|^2 - 2 a16.b16 within the slop the candidate threshold budgets."""
    src, tgt, dim = synth.make_pair(desc, nq, nt, nan_frac=0.01)
    with M.Context(0) as ctx:
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        oq, nq16, scale = ctx.debug_operands(0, True)
        ot, nt16, _ = ctx.debug_operands(1, False)
        assert oq.shape[1] % 64 == 0 and oq.shape[0] % 256 == 0 and ot.shape[0] % 256 == 0
        ot_q, _, _ = ctx.debug_operands(0, False)
        valid = np.isfinite(src[:, :dim]).all(1)
        # query form = -2 * train form, then ones picking up the norm triple
        assert np.array_equal(oq[:nq, :dim].astype(np.float32), -2.0 * ot_q[:nq, :dim].astype(np.float32))
        assert np.all(oq[:, dim:dim + 3] == 1) and np.all(oq[:, dim + 3:] == 0) and np.all(ot[:, dim + 3:] == 0)
        x16 = ot_q[:nq, :dim].astype(np.float64)
        np.testing.assert_allclose(nq16[:nq][valid], (x16[valid] ** 2).sum(1), rtol=3e-7)
        assert np.abs(x16).max() <= 1.0
        tvalid = np.isfinite(tgt[:, :dim]).all(1)
        triple = ot[:nt, dim:dim + 3].astype(np.float64).sum(1)
        np.testing.assert_allclose(triple[tvalid], (ot[:nt, :dim].astype(np.float64)[tvalid] ** 2).sum(1), rtol=3e-7)
        assert np.all(triple[~tvalid] > 5e4) and np.all(ot[nt:, dim].astype(np.float64) > 5e4)   # sentinel rows
        # the centred/scaled FP16 rows approximate scale*(x - c) to FP16 rounding
        c = np.concatenate([src[valid, :dim], tgt[tvalid, :dim]]).astype(np.float64).mean(0)
        ref = (src[valid, :dim].astype(np.float64) - c) * scale
        assert np.abs(x16[valid] - ref).max() <= 2.0 ** -11 * 1.01 + 1e-6
        # raw accumulators of a few tiles vs float64 on the same FP16 operands
        worst = 0.0
        for (q0, tt) in [(0, 0), (128, 1), (nq - nq % 128 if nq % 128 else nq - 128, ot.shape[0] // 256 - 1)]:
            acc = ctx.debug_tc_tile(0, q0, tt)
            a = oq[q0:q0 + 128].astype(np.float64)
            b = ot[tt * 256:(tt + 1) * 256].astype(np.float64)
            exp = a @ b.T
            rows = min(128, oq.shape[0] - q0)
            na = np.sqrt(np.maximum((a[:rows, :dim] ** 2).sum(1) / 4.0, 0))[:, None]
            nb = np.sqrt((b[:, :dim] ** 2).sum(1))[None, :]
            ok = (np.abs(exp[:rows]) < 1e4)          # leave the sentinel rows out of the relative measure
            err = np.abs(acc[:rows] - exp[:rows])
            budget = (na + nb) ** 2 * 2.0 ** -16 + 1e-6
            assert np.all(err[ok] <= np.broadcast_to(budget, err.shape)[ok])
            worst = max(worst, float((err / np.broadcast_to((na + nb) ** 2 + 1e-30, err.shape))[ok].max()))
            np.testing.assert_allclose(acc[:rows][~ok], exp[:rows][~ok], rtol=1e-3)
        print("max tensor-core accumulation error / (|a|+|b|)^2 = %.3e (budget 2^-16 = %.3e)" % (worst, 2.0 ** -16))
        # and the whole pass agrees with the oracle
        _same(ctx.knn(k, 0), orc.knn(_dense(src, dim), _dense(tgt, dim), k))
        _same(ctx.knn(k, 1), orc.knn(_dense(tgt, dim), _dense(src, dim), k))


@pytest.mark.parametrize("prec", PRECS, ids=["exact", "tc"])
def test_match_mutual_with_masked_reverse_pass(monkeypatch, golden_dir, prec):
    """b200m_match's large-problem path (reverse kNN only for the target rows the forward lists name), forced on at
    test size: the same records, distances and thresholds as the oracle."""
    monkeypatch.setenv("B200M_MASKED_MIN_PAIRS", "0")
    for case in ("fpfh_k5", "shot_k2", "fpfh_k1"):
        g = np.load(os.path.join(golden_dir, case + ".npz"))
        dim, k = int(g["dim"]), int(g["k"])
        src, tgt = g["src"], g["tgt"]
        with M.Context(0) as ctx:
            ctx.upload(0, src, dim)
            ctx.upload(1, tgt, dim)
            got, avg = ctx.match(k, M.MODE_MUTUAL, 1.1, np.float32(0.7), precision=prec)
            got_rm, _ = ctx.match(max(k, 2), M.MODE_RATIO_MUTUAL, 1.1, np.float32(0.7), precision=prec)
        exp, eavg = orc.match(_dense(src, dim), _dense(tgt, dim), k, "mutual", 1.1, np.float32(0.7))
        assert got.tobytes() == exp.tobytes() and avg == eavg and len(exp) > 0
        monkeypatch.setenv("B200M_MASKED_MIN_PAIRS", "1e30")
        with M.Context(0) as ctx:
            ctx.upload(0, src, dim)
            ctx.upload(1, tgt, dim)
            ref_rm, _ = ctx.match(max(k, 2), M.MODE_RATIO_MUTUAL, 1.1, np.float32(0.7), precision=prec)
        monkeypatch.setenv("B200M_MASKED_MIN_PAIRS", "0")
        assert got_rm.tobytes() == ref_rm.tobytes()


@pytest.mark.parametrize("prec", PRECS, ids=["exact", "tc"])
@pytest.mark.parametrize("mode,name", [("one_sided", M.MODE_ONE_SIDED), ("mutual", M.MODE_MUTUAL), ("ratio", M.MODE_RATIO)])
def test_match_equals_oracle(golden_dir, mode, name, prec):
    """b200m_match == OneSided/LeftToRight(/Ratio) match_impl restated in the oracle: same records, same order,
    same distances and thresholds, same average first-NN distance."""
    for case in ("fpfh_k5", "shot_k2", "fpfh_k1"):
        g = np.load(os.path.join(golden_dir, case + ".npz"))
        dim, k = int(g["dim"]), int(g["k"])
        if mode == "ratio" and k < 2:
            continue
        src, tgt = g["src"], g["tgt"]
        rng = np.random.default_rng(5)
        thr_s = rng.random(src.shape[0]).astype(np.float32)
        thr_t = rng.random(tgt.shape[0]).astype(np.float32)
        dthr = np.float32(0.7)
        with M.Context(0) as ctx:
            ctx.upload(0, src, dim)
            ctx.upload(1, tgt, dim)
            got, avg = ctx.match(k, name, 1.1, dthr, thr_s, thr_t, precision=prec)
            got2, avg2 = ctx.match(k, name, 1.1, dthr, precision=prec)
        exp, eavg = orc.match(_dense(src, dim), _dense(tgt, dim), k, mode, 1.1, dthr, thr_s, thr_t)
        assert len(got) == len(exp) and len(exp) > 0
        assert got.tobytes() == exp.tobytes()
        assert avg == eavg and avg2 == eavg
        assert np.all(got2["threshold"] == dthr) and np.array_equal(got2["index_match"], exp["index_match"])
        assert np.all(np.diff(got["index_query"]) >= 0)


def test_ratio_mutual_mode(golden_dir):
    g = np.load(os.path.join(golden_dir, "fpfh_k2_blocked.npz"))
    dim, k = int(g["dim"]), int(g["k"])
    src, tgt = g["src"], g["tgt"]
    with M.Context(0) as ctx:
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        got, _ = ctx.match(k, M.MODE_RATIO_MUTUAL, 1.1)
    fi, fd, fc = orc.knn(_dense(src, dim), _dense(tgt, dim), k)
    ri, rd, rc = orc.knn(_dense(tgt, dim), _dense(src, dim), k)
    exp = []
    for i in range(src.shape[0]):
        if fc[i] < 2 or not (fd[i, 1] >= np.float32(1.1) * fd[i, 0]):
            continue
        j = fi[i, 0]
        for m in range(rc[j]):
            if ri[j, m] == i:
                exp.append((i, j, rd[j, m]))
                break
    assert [(int(a), int(b), float(c)) for a, b, c in zip(got["index_query"], got["index_match"], got["distance"])] == \
           [(int(a), int(b), float(c)) for a, b, c in exp]
    assert 0 < len(exp) < src.shape[0]


def test_reference_style_interface(golden_dir):
    """The host mirror of the reference interface: matchBF/matchFLANN/matchLocal agree (tests/flann_bf_matcher.h),
    matcher classes finalize to cloud-global indices (include/matching.h:356-362)."""
    g = np.load(os.path.join(golden_dir, "rops_k3.npz"))
    dim, k = int(g["dim"]), int(g["k"])
    src, tgt = g["src"], g["tgt"]
    params = M.AlignmentParameters(randomness=k, matching_id=M.MATCHING_LEFT_TO_RIGHT)
    bf = M.match_bf(src, tgt, params, dim)
    fl = M.match_flann(src, tgt, params, dim)
    lo = M.match_local(src, tgt, params, dim)
    _same(bf, fl)
    _same(bf, lo)
    assert np.array_equal(bf[0], g["bf_idx"])
    rng = np.random.default_rng(1)
    ks = rng.permutation(10 * src.shape[0])[:src.shape[0]].astype(np.int32)
    kt = rng.permutation(10 * tgt.shape[0])[:tgt.shape[0]].astype(np.int32)
    # randomness = 1, one scale: match_multiscale's vote is the identity, no keypoint coordinates needed
    p1 = M.AlignmentParameters(randomness=1, matching_id=M.MATCHING_LEFT_TO_RIGHT)
    m = M.get_feature_based_matcher_from_parameters(src, tgt, p1, dim=dim, kps_indices_src=ks, kps_indices_tgt=kt)
    assert m.get_class_name() == "LeftToRightMatcher" and m.get_average_distance() == M.FLT_MAX
    corrs = m.match()
    exp, eavg = orc.match(_dense(src, dim), _dense(tgt, dim), 1, "mutual", distance_thr=np.float32(M.FLT_MAX))
    exp = orc.finalize(exp, ks, kt)
    assert corrs.tobytes() == exp.tobytes() and m.get_average_distance() == eavg
    # randomness > 1 feeds the spatial vote (include/matching.h:327-352): the classes refuse to run without coordinates
    with pytest.raises(M.B200MatchError):
        M.get_feature_based_matcher_from_parameters(src, tgt, params, dim=dim).match()
    with pytest.raises(M.B200MatchError):
        M.get_feature_based_matcher_from_parameters(src, tgt, M.AlignmentParameters(matching_id="cluster"))


def test_error_behaviour():
    with M.Context(0) as ctx:
        a = np.zeros((8, 33), np.float32)
        ctx.upload(0, a, 33)
        ctx.upload(1, np.zeros((8, 34), np.float32), 34)
        with pytest.raises(M.B200MatchError):
            ctx.knn(1, 0)                       # descriptor lengths differ
        ctx.upload(1, a, 33)
        with pytest.raises(M.B200MatchError):
            ctx.knn(0, 0)                       # k out of range
        with pytest.raises(M.B200MatchError):
            ctx.match(1, M.MODE_RATIO)          # ratio needs k >= 2
        with pytest.raises(M.B200MatchError):
            ctx.knn(1, 0, row_begin=4, row_end=100)
        idx, dist, cnt = ctx.knn(2, 0, row_begin=2, row_end=6)
        assert idx.shape == (4, 2) and idx[0].tolist() == [0, 1]


@pytest.mark.parametrize("desc,n,k", [("fpfh", 20000, 1), ("shot", 6000, 2)])
def test_config1_scale_mutual(desc, n, k):
    """BASELINE config 1 (FPFH-33 ~20k x 20k, k=1, mutual) in full against the oracle, and a SHOT-352 case:
    bit-exact lists both directions, bit-exact correspondences; candidate-list overflow stays rare."""
    src, tgt, dim = synth.make_pair(desc, n, n + 137)
    with M.Context(0) as ctx:
        ctx.set_profiling(True)
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        fwd = ctx.knn(k, 0)
        rev = ctx.knn(k, 1)
        st = ctx.stats()
        got, avg = ctx.match(k, M.MODE_MUTUAL)
    efwd = orc.knn(_dense(src, dim), _dense(tgt, dim), k)
    erev = orc.knn(_dense(tgt, dim), _dense(src, dim), k)
    _same(fwd, efwd)
    _same(rev, erev)
    exp = orc.filter_mutual(efwd[0], efwd[2], erev[0], erev[1], erev[2], np.float32(M.FLT_MAX))
    assert got.tobytes() == exp.tobytes() and avg == orc.average_distance(efwd[1], efwd[2])
    print(desc, "candidates/row %.1f" % (st["candidates"] / st["rows_total"]), "flagged", st["rows_flagged"], st)
    assert 0.2 * n < len(got) <= n * k
    assert st["rows_flagged"] <= 0.01 * st["rows_total"], st


def test_properties_at_bench_scale():
    """Size-independent properties on a larger FPFH run (oracle on a row subsample only)."""
    n = 100000
    src, tgt, dim = synth.make_pair("fpfh", n, n)
    k = 2
    with M.Context(0) as ctx:
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        idx, dist, cnt = ctx.knn(k, 0)
        idx2, dist2, cnt2 = ctx.knn(k, 0)
        sub = ctx.knn(k, 0, row_begin=5000, row_end=5000 + 512)
        ex = ctx.knn(k, 0, row_begin=5000, row_end=5000 + 512, precision=M.PREC_F32_EXACT)
    _same((idx, dist, cnt), (idx2, dist2, cnt2))                       # idempotent
    _same(sub, (idx[5000:5512], dist[5000:5512], cnt[5000:5512]))      # row ranges are consistent
    _same(sub, ex)                                                     # tensor-core path == exact CUDA-core path
    good = np.isfinite(src[:, :dim]).all(1)
    assert np.all(cnt[good] == k) and np.all(cnt[~good] == 0)
    assert np.all(np.diff(dist[good], axis=1) >= 0)                    # ascending
    assert np.all(idx[good] >= 0) and np.all(idx[good] < n)
    rows = np.random.default_rng(0).choice(n, 1024, replace=False)
    eo = orc.knn(np.ascontiguousarray(src[rows, :dim]), _dense(tgt, dim), k)
    _same((idx[rows], dist[rows], cnt[rows]), eo)


FPFH_LAYOUTS = {"eight-warps": 0, "alternating-tiles": 1, "quarter-columns": 2, "chunk-entries": 4, "chunk-entries-x16": 5}


@pytest.mark.parametrize("layout", list(FPFH_LAYOUTS))
@pytest.mark.parametrize("nq,nt,k", [(2500, 40000, 2), (700, 21000, 5), (300, 600, 1), (900, 12345, 11)])
def test_fpfh_epilogue_layouts(monkeypatch, layout, nq, nt, k):
    """The production (split-N) FPFH candidate kernel with each epilogue layout -- eight warps of 128 accumulators, sixteen
    warps bound to one accumulator buffer (alternating tiles), sixteen warps of 64 columns of every tile (all three with
    columns as list entries), and the default: alternating tiles with 32-column CHUNKS as list entries (the re-rank evaluates
    every train row of a surviving chunk; 12345 train rows: a ragged last chunk), drained by 32- or by 16-column TMEM loads
    -- oracle-exact lists in both directions."""
    monkeypatch.setenv("B200M_TC_ALT", str(FPFH_LAYOUTS[layout]))
    src, tgt, dim = synth.make_pair("fpfh", nq, nt, nan_frac=0.01)
    with M.Context(0) as ctx:
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        _same(ctx.knn(k, 0), orc.knn(_dense(src, dim), _dense(tgt, dim), k))
        _same(ctx.knn(k, 1), orc.knn(_dense(tgt, dim), _dense(src, dim), k))
        assert ctx.stats()["rows_flagged"] == 0


@pytest.mark.parametrize("lag", [1, 5, 20])
@pytest.mark.parametrize("layout", ["chunk-entries", "chunk-entries-x16"])
@pytest.mark.parametrize("nq,nt,k,splits", [(30000, 9000, 2, 0), (40000, 1300, 5, 0), (700, 70000, 3, 3)])
def test_fpfh_rotated_sweep(monkeypatch, layout, lag, nq, nt, k, splits):
    """The chunk-entry kernels with the rotated sweep forced on (by default only train sets beyond L2 use it): several waves
    of CTA pairs (30000 rows = 118 pairs on 74 slots) so that late starters read a non-zero front position, train sets
    shorter than the followers' distances (5 tiles: every start wraps around, negative positions), and several train
    splits -- oracle-exact lists in both directions."""
    monkeypatch.setenv("B200M_TC_ALT", str(FPFH_LAYOUTS[layout]))
    monkeypatch.setenv("B200M_TC_SWEEP_LAG", str(lag))
    if splits:
        monkeypatch.setenv("B200M_TC_SPLITS", str(splits))
    src, tgt, dim = synth.make_pair("fpfh", nq, nt, nan_frac=0.01)
    with M.Context(0) as ctx:
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        rows = np.random.default_rng(3).choice(nq, min(nq, 3000), replace=False)
        idx, dist, cnt = ctx.knn(k, 0)
        _same((idx[rows], dist[rows], cnt[rows]), orc.knn(np.ascontiguousarray(src[rows, :dim]), _dense(tgt, dim), k))
        rows = np.random.default_rng(4).choice(nt, min(nt, 3000), replace=False)
        idx, dist, cnt = ctx.knn(k, 1)
        _same((idx[rows], dist[rows], cnt[rows]), orc.knn(np.ascontiguousarray(tgt[rows, :dim]), _dense(src, dim), k))
        assert ctx.stats()["rows_flagged"] == 0


@pytest.mark.parametrize("epi", ["eh1", "eh2"])
@pytest.mark.parametrize("mode", ["pair", "mcast1", "mcast2", "mcast4"])
@pytest.mark.parametrize("desc,nq,nt,k", [("fpfh", 3000, 9000, 2), ("shot", 1100, 2600, 5)])
def test_candidate_kernel_modes(monkeypatch, mode, epi, desc, nq, nt, k):
    """The candidate kernel's operand-sharing schemes -- CTA pair (tcgen05 cta_group::2, default) and cta_group::1
    alone or with TMA multicast across a cluster of 2 or 4 -- and its two epilogue shapes (4 warps, a thread owns a
    row; 8 warps, two threads share a row) all give the same (oracle-exact) lists."""
    if mode == "pair":
        monkeypatch.setenv("B200M_TC_MODE", "pair")
    else:
        monkeypatch.setenv("B200M_TC_CLUSTER", mode[-1])
    monkeypatch.setenv("B200M_TC_DEBUG", "64" if epi == "eh1" else "128")   # 64 / 128 only force the epilogue shape
    src, tgt, dim = synth.make_pair(desc, nq, nt, nan_frac=0.01)
    with M.Context(0) as ctx:
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        _same(ctx.knn(k, 0), orc.knn(_dense(src, dim), _dense(tgt, dim), k))
        _same(ctx.knn(k, 1), orc.knn(_dense(tgt, dim), _dense(src, dim), k))


@pytest.mark.parametrize("desc,nq,nt,k,regime", [("fpfh", 200, 5000, 2, "few"), ("shot", 150, 3000, 5, "few"),
                                                 ("fpfh", 2000, 5000, 2, "many"), ("rops", 700, 2500, 3, "many")])
def test_overflowed_candidate_lists_fall_back_exactly(desc, nq, nt, k, regime):
    """A candidate list that overflows its slots sends the row to the exact CUDA-core path: split over the whole grid
    when few rows overflow (<= 256), one CTA per row when many do.  cand_cap = k makes every row overflow."""
    src, tgt, dim = synth.make_pair(desc, nq, nt, nan_frac=0.01)
    with M.Context(0) as ctx:
        ctx.set_profiling(True)
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        ctx.reset_stats()
        got = ctx.knn(k, 0, cand_cap=k)
        flagged = ctx.stats()["rows_flagged"]
        assert (0 < flagged <= 256) if regime == "few" else flagged > 256
        _same(got, orc.knn(_dense(src, dim), _dense(tgt, dim), k))
        _same(ctx.knn(k, 0, cand_cap=k), got)   # the split kernel's completion counters are reusable


@pytest.mark.parametrize("desc,k,n_scales", [("fpfh", 1, 1), ("fpfh", 3, 3), ("shot", 2, 2), ("rops", 5, 4)])
def test_match_multiscale_vote_equals_oracle(desc, k, n_scales):
    """match_multiscale's tail (reference include/matching.h:264-354): per-scale kNN over keypoint subsets, remap to
    keypoint ids, concatenate in scale order, spatial vote -> one match per query keypoint; == the oracle's restatement
    (orc.knn per scale + orc.spatial_vote on the concatenated lists), bit for bit."""
    rng = np.random.default_rng(17 + k)
    n_qk, n_tk = 900, 1100                      # keypoints of the two clouds
    xyz = (rng.random((n_tk, 4)) * 2.0).astype(np.float32)      # pcl::PointXYZ rows (16 B)
    xyz[: n_tk // 3, :3] = xyz[0, :3] + 0.01 * rng.standard_normal((n_tk // 3, 3)).astype(np.float32)   # a tight cluster
    iss_radius = 0.05
    q_scales, t_scales, dim = [], [], None
    comb_i = [[] for _ in range(n_qk)]
    comb_d = [[] for _ in range(n_qk)]
    for s in range(n_scales):
        qmap = np.sort(rng.choice(n_qk, size=n_qk - 60 * s, replace=False)).astype(np.int32)   # the scale's keypoint subset
        tmap = np.sort(rng.choice(n_tk, size=n_tk - 45 * s, replace=False)).astype(np.int32)
        src, tgt, dim = synth.make_pair(desc, qmap.shape[0], tmap.shape[0], seed=100 + s, nan_frac=0.01)
        q_scales.append((src, qmap))
        t_scales.append((tgt, tmap))
        ei, ed, ec = orc.knn(_dense(src, dim), _dense(tgt, dim), k)
        for r in range(qmap.shape[0]):
            for m in range(ec[r]):
                comb_i[qmap[r]].append(int(tmap[ei[r, m]]))
                comb_d[qmap[r]].append(ed[r, m])
    width = max(max(len(c) for c in comb_i), 1)
    ci = np.full((n_qk, width), -1, np.int32)
    cd = np.zeros((n_qk, width), np.float32)
    cc = np.zeros(n_qk, np.int32)
    for i in range(n_qk):
        cc[i] = len(comb_i[i])
        ci[i, :cc[i]] = comb_i[i]
        cd[i, :cc[i]] = comb_d[i]
    vi, vd, vc = orc.spatial_vote(ci, cd, cc, xyz[:, :3], iss_radius)
    params = M.AlignmentParameters(randomness=k)
    gi, gd, gc = M.match_multiscale(q_scales, t_scales, n_qk, xyz, iss_radius, params, dim=dim)
    assert np.array_equal(gc, vc) and np.array_equal(gi, vi[:, 0]) and np.array_equal(gd, vd[:, 0])
    assert gc.sum() > 0.9 * n_qk - 60 * n_scales
    if k * n_scales > 1:   # the vote is not the identity: some keypoints do not keep their nearest descriptor
        first = np.array([c[0] if c else -1 for c in comb_i])
        assert np.any(gi != first)


@pytest.mark.parametrize("masked", [False, True], ids=["full-reverse", "masked-reverse"])
@pytest.mark.parametrize("mode", ["one_sided", "mutual", "cluster"])
@pytest.mark.parametrize("desc,k,n_scales", [("fpfh", 1, 1), ("fpfh", 2, 1), ("shot", 2, 2), ("rops", 3, 3), ("fpfh", 5, 2)])
def test_matcher_classes_at_the_wide_seam_equal_oracle(monkeypatch, desc, k, n_scales, mode, masked):
    """match_impl of OneSided / LeftToRight / ClusterMatcher as the reference composes it (include/matching.h:395-411,
    :428-453, :492-517): match_multiscale in BOTH directions (the reverse one is the reference's inverse_tn call) -- per-scale
    kNN, remap, concatenation, spatial vote -- the average over the voted forward lists, then the filter loop over the voted
    lists; records and average == the oracle composed the same way, byte for byte."""
    monkeypatch.setenv("B200M_MASKED_MIN_PAIRS", "1" if masked else "1e30")
    rng = np.random.default_rng(31 + 7 * k + n_scales)
    n_sk, n_tk = 800, 950
    sx = np.zeros((n_sk, 4), np.float32)
    tx = np.zeros((n_tk, 4), np.float32)
    sx[:, :3] = rng.random((n_sk, 3)) * 4
    tx[:, :3] = rng.random((n_tk, 3)) * 4
    tx[: n_tk // 4, :3] = tx[0, :3] + 0.02 * rng.standard_normal((n_tk // 4, 3)).astype(np.float32)   # tight clusters:
    sx[: n_sk // 4, :3] = sx[0, :3] + 0.02 * rng.standard_normal((n_sk // 4, 3)).astype(np.float32)   # the vote matters
    iss_s, iss_t = np.float32(0.06), np.float32(0.05)
    s_sc, t_sc, dim = [], [], None
    for s in range(n_scales):
        smap = None if (s == 0 and n_scales == 1 and k == 1) else np.sort(rng.choice(n_sk, n_sk - 50 * s, replace=False)).astype(np.int32)
        tmap = None if (s == 0 and n_scales == 1 and k == 1) else np.sort(rng.choice(n_tk, n_tk - 40 * s, replace=False)).astype(np.int32)
        src, tgt, dim = synth.make_pair(desc, n_sk if smap is None else smap.shape[0], n_tk if tmap is None else tmap.shape[0],
                                        seed=200 + s, nan_frac=0.01)
        s_sc.append((np.ascontiguousarray(_dense(src, dim)), smap))
        t_sc.append((np.ascontiguousarray(_dense(tgt, dim)), tmap))
    thr_s, thr_t = rng.random(n_sk).astype(np.float32), rng.random(n_tk).astype(np.float32)
    dthr = np.float32(0.7)
    exp, eavg = orc.match_wide(mode, s_sc, t_sc, sx[:, :3], tx[:, :3], iss_s, iss_t, k, 40, dthr, thr_s, thr_t)
    gmode = {"one_sided": M.MODE_ONE_SIDED, "mutual": M.MODE_MUTUAL, "cluster": M.MODE_CLUSTER}[mode]
    with M.Context(0) as ctx:
        got, avg = ctx.match_wide(k, gmode, s_sc, t_sc, sx, tx, iss_s, iss_t, dim, 40, dthr, thr_s, thr_t)
    assert len(exp) > 0 and got.tobytes() == exp.tobytes() and avg == eavg
    assert np.all(np.diff(got["index_query"]) > 0)       # at most one correspondence per source keypoint, ascending
    # the reference-shaped classes, finalize included
    ks = rng.permutation(5 * n_sk)[:n_sk].astype(np.int32)
    kt = rng.permutation(5 * n_tk)[:n_tk].astype(np.int32)
    params = M.AlignmentParameters(randomness=k, distance_thr=float(dthr), cluster_k=40,
                                   matching_id={"one_sided": M.MATCHING_ONE_SIDED, "mutual": M.MATCHING_LEFT_TO_RIGHT,
                                                "cluster": M.MATCHING_CLUSTER}[mode])
    m = M.get_feature_based_matcher_from_parameters(
        [a for a, _ in s_sc], [a for a, _ in t_sc], params, dim=dim, thresholds_src=thr_s, thresholds_tgt=thr_t,
        kps_indices_src=ks, kps_indices_tgt=kt, kps_xyz_src=sx, kps_xyz_tgt=tx, iss_radius_src=iss_s, iss_radius_tgt=iss_t,
        kps_indices_multiscale_src=[m_ for _, m_ in s_sc], kps_indices_multiscale_tgt=[m_ for _, m_ in t_sc])
    assert m.match().tobytes() == orc.finalize(exp, ks, kt).tobytes() and m.get_average_distance() == eavg


@pytest.mark.parametrize("desc,nq,nt,k,ck", [("fpfh", 1500, 1700, 1, 40), ("shot", 600, 500, 2, 40), ("rops", 400, 450, 3, 10),
                                             ("fpfh", 30, 25, 1, 40)])
def test_cluster_matcher_equals_oracle(desc, nq, nt, k, ck):
    """ClusterMatcher::match_impl (reference include/matching.h:492-517, the default matching_id): forward and reverse
    kNN, 3-D neighbourhoods of the keypoints (cluster_k nearest, the point itself included), consistency distances both
    ways, threshold 0.95 -- the same records as the oracle's restatement; the last case has fewer keypoints than cluster_k."""
    src, tgt, dim = synth.make_pair(desc, nq, nt, nan_frac=0.01)
    rng = np.random.default_rng(3)
    # keypoints: the target cloud is the source cloud moved rigidly + noise where descriptors correspond, so that the
    # geometric consistency check keeps a real fraction; PointXYZ rows (16 B)
    sx = np.zeros((nq, 4), np.float32)
    tx = np.zeros((nt, 4), np.float32)
    sx[:, :3] = rng.random((nq, 3)) * 10
    fi, fd, fc = orc.knn(_dense(src, dim), _dense(tgt, dim), 1)
    tx[:, :3] = rng.random((nt, 3)) * 10
    good = fc > 0
    tx[fi[good, 0], :3] = sx[good, :3] + np.float32(0.01) * rng.standard_normal((int(good.sum()), 3)).astype(np.float32)
    thr_s = rng.random(nq).astype(np.float32)
    thr_t = rng.random(nt).astype(np.float32)
    dthr = np.float32(0.6)
    with M.Context(0) as ctx:
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        got, avg = ctx.match_cluster(k, ck, sx, tx, dthr, thr_s, thr_t)
    fidx, fdist, fcnt = orc.knn(_dense(src, dim), _dense(tgt, dim), k)
    ridx, rdist, rcnt = orc.knn(_dense(tgt, dim), _dense(src, dim), k)
    ns, ng = orc.knn3d(sx[:, :3], ck), orc.knn3d(tx[:, :3], ck)
    exp = orc.filter_cluster(fidx, fcnt, ridx, rcnt, ns, ng, dthr, thr_q=thr_s, thr_t=thr_t)
    assert len(exp) > 0
    if min(nq, nt) > ck:
        assert len(exp) < fcnt.sum()          # the filter keeps some pairs and rejects some
    assert got.tobytes() == exp.tobytes()
    assert avg == orc.average_distance(fdist, fcnt)
    # the reference-shaped class goes through match_multiscale's vote first (at most one match per keypoint)
    params = M.AlignmentParameters(randomness=k, matching_id=M.MATCHING_CLUSTER, cluster_k=ck, distance_thr=float(dthr))
    m = M.get_feature_based_matcher_from_parameters(src, tgt, params, dim=dim, kps_xyz_src=sx, kps_xyz_tgt=tx,
                                                    thresholds_src=thr_s, thresholds_tgt=thr_t, iss_radius_src=0.3,
                                                    iss_radius_tgt=0.25)
    exp_w, avg_w = orc.match_wide("cluster", [(_dense(src, dim), None)], [(_dense(tgt, dim), None)], sx[:, :3], tx[:, :3],
                                  np.float32(0.3), np.float32(0.25), k, ck, dthr, thr_s, thr_t)
    assert m.get_class_name() == "ClusterMatcher" and m.match().tobytes() == exp_w.tobytes()
    assert m.get_average_distance() == avg_w
    if k == 1:
        assert exp_w.tobytes() == exp.tobytes()      # one candidate per keypoint: the vote is the identity


@pytest.mark.parametrize("desc,nq,nt,k,radius", [("fpfh", 900, 1200, 2, 1.5), ("shot", 300, 500, 5, 3.0), ("rops", 250, 300, 1, 0.4),
                                                 ("fpfh", 400, 300, 3, 100.0)])
@pytest.mark.parametrize("path", ["cell-list", "all-rows"])
def test_match_local_with_search_radius_equals_oracle(monkeypatch, desc, nq, nt, k, radius, path):
    """matchLocal with a finite match_search_radius (reference include/matching.h:637-678): only train rows whose keypoint
    is within the radius of the (transformed) query keypoint compete; lists shorter than k where the gate leaves fewer
    rows; the last case's radius covers everything (== plain kNN up to the order of exactly tied distances).  Both device
    paths: the cell list of the train keypoints (csrc/local.cu; radii 3.0 and 100 are too large for a grid of this cloud
    and fall through to the other kernel by themselves) and the gate tested against every train row (csrc/exact.cu)."""
    monkeypatch.setenv("B200M_LOCAL_MIN_ROWS", "1" if path == "cell-list" else "1000000000")
    src, tgt, dim = synth.make_pair(desc, nq, nt, nan_frac=0.01)
    rng = np.random.default_rng(5)
    qx = np.zeros((nq, 4), np.float32)
    tx = np.zeros((nt, 4), np.float32)
    qx[:, :3] = rng.random((nq, 3)) * 6
    tx[:, :3] = rng.random((nt, 3)) * 6
    with M.Context(0) as ctx:
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        got = ctx.knn_local(k, qx, tx, radius)
        got_rev = ctx.knn_local(k, tx, qx, radius, direction=1)
    exp = orc.match_local(_dense(src, dim), _dense(tgt, dim), k, qx[:, :3], tx[:, :3], radius)
    _same(got, exp)
    _same(got_rev, orc.match_local(_dense(tgt, dim), _dense(src, dim), k, tx[:, :3], qx[:, :3], radius))
    if radius < 50:
        assert (exp[2] < k).any() and (exp[2] > 0).any()      # the gate bites
        full = orc.knn(_dense(src, dim), _dense(tgt, dim), k)
        assert not np.array_equal(full[0], exp[0])
    # the reference-shaped free function, with a guess: a pure translation moves the query keypoints into place
    guess = np.eye(4, dtype=np.float32)
    guess[:3, 3] = [0.5, -0.25, 1.0]
    moved = qx.copy()
    moved[:, :3] = qx[:, :3] - guess[:3, 3]
    params = M.AlignmentParameters(randomness=k)
    got2 = M.match_local(src, tgt, params, dim=dim, query_kps_xyz=moved, train_kps_xyz=tx, guess=guess, match_search_radius=radius)
    back = moved[:, :3] @ guess[:3, :3].T + guess[:3, 3]
    _same(got2, orc.match_local(_dense(src, dim), _dense(tgt, dim), k, back, tx[:, :3], radius))


@pytest.mark.parametrize("desc,nq,nt,k,radius", [("fpfh", 6000, 20000, 2, 0.9), ("shot", 1500, 9000, 3, 2.5)])
def test_match_local_cell_list_on_a_clustered_cloud(monkeypatch, desc, nq, nt, k, radius):
    """The cell-list path on a cloud big enough to use it by default: keypoints with dense clumps and empty space, NaN
    coordinates, queries outside the train cloud's bounding box -- records equal the oracle's in both directions."""
    src, tgt, dim = synth.make_pair(desc, nq, nt, nan_frac=0.01)
    rng = np.random.default_rng(9)
    qx = np.zeros((nq, 4), np.float32)
    tx = np.zeros((nt, 4), np.float32)
    tx[:, :3] = rng.random((nt, 3)) * np.float32([40, 25, 8])
    tx[: nt // 5, :3] = tx[0, :3] + 0.3 * rng.standard_normal((nt // 5, 3)).astype(np.float32)     # a dense clump
    qx[:, :3] = rng.random((nq, 3)) * np.float32([44, 27, 9]) - np.float32([2, 1, 0.5])                # some outside the box
    qx[: nq // 5, :3] = tx[0, :3] + 0.4 * rng.standard_normal((nq // 5, 3)).astype(np.float32)
    qx[7, 0] = np.nan
    tx[11, 1] = np.nan
    with M.Context(0) as ctx:
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        got = ctx.knn_local(k, qx, tx, radius)
        got_rev = ctx.knn_local(k, tx, qx, radius, direction=1)
    exp = orc.match_local(_dense(src, dim), _dense(tgt, dim), k, qx[:, :3], tx[:, :3], radius)
    _same(got, exp)
    _same(got_rev, orc.match_local(_dense(tgt, dim), _dense(src, dim), k, tx[:, :3], qx[:, :3], radius))
    assert (exp[2] == 0).any() and (exp[2] == k).any() and got[2][7] == 0


def test_upload_from_pageable_memory_goes_through_the_bounce_buffers():
    """b200m_upload of a large set from ordinary (pageable) host memory -- what a pcl::PointCloud is -- is staged through
    pinned bounce buffers by several host threads (csrc/api.cu staged_h2d; 40 MB here = three chunks, the last one
    partial): the descriptors must arrive intact, so the kNN against them equals the oracle's."""
    src, tgt, dim = synth.make_pair("fpfh", 300000, 1500, nan_frac=0.001)
    src = np.ascontiguousarray(src)          # numpy memory is pageable
    assert src.nbytes >= 32 << 20
    with M.Context(0) as ctx:
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        got = ctx.knn(2, 1)                  # the 1500 target rows against all 300k uploaded source rows
    _same(got, orc.knn(_dense(tgt, dim), _dense(src, dim), 2))


@pytest.mark.parametrize("workload", ["c2", "c3", "c4"])
def test_full_size_workloads_match_oracle_on_sampled_rows(workload):
    """BASELINE.json's full sizes (C2 FPFH-33 200k x 200k k=2, C3 SHOT-352 500k x 500k k=2, C4 FPFH-33 2M x 2M k=5): the
    device-resident kNN of the whole problem in both directions -- idempotent, ascending, complete, row-range consistent
    on ALL rows -- and bit-exact against the CPU oracle on 4096 random query rows per direction vs the full train set
    (tools/fullsize_parity.py; the 8M-row target-sharded C5 needs 8 GPUs: tools/multigpu_check.py c5,
    profiles/r02_multigpu_check_c5_8gpu.log)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fullsize_parity.py"), workload, "4096"],
                       capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("bit-exact") == 2


@pytest.mark.parametrize("desc,nq,nt,k", [("fpfh", 2500, 3100, 2), ("shot", 900, 1300, 2), ("rops", 700, 600, 5)])
def test_masked_reverse_pass_gives_the_same_mutual_records(desc, nq, nt, k):
    """The mutual filter reads rev[j] only for targets j named by a forward list, so the reverse kNN may skip the other
    target rows (b200m_knn_masked_device + b200m_mark_referenced_device; b200m_match does it for large problems):
    skipped rows come back empty, answered rows are the oracle's, and the filter output is unchanged."""
    import torch
    from lidar_global_registration_b200 import device as D
    src, tgt, dim = synth.make_pair(desc, nq, nt, nan_frac=0.01)
    be = D.GpuBackend(0)
    try:
        be.upload_device(0, torch.from_numpy(src).cuda(), dim)
        be.upload_device(1, torch.from_numpy(tgt).cuda(), dim)
        fwd = be.knn(k, 0, 0, nq)
        flags = be.referenced_rows(k, fwd, nt)
        fi, fd, fc = orc.knn(_dense(src, dim), _dense(tgt, dim), k)
        exp_flags = np.zeros(nt, np.uint8)
        for i in range(nq):
            exp_flags[fi[i, :fc[i]]] = 1
        assert np.array_equal(flags.cpu().numpy(), exp_flags) and 0 < exp_flags.sum() < nt
        lo, hi = nt // 5, nt - 3                       # a row range, as a sharded rank would ask for
        ridx, rdist, rcnt = [x.cpu().numpy() for x in be.knn_masked(k, 1, lo, hi, flags)]
        ei, ed, ec = orc.knn(_dense(tgt, dim), _dense(src, dim), k)
        keep = (exp_flags[lo:hi] == 1) & np.isfinite(tgt[lo:hi, :dim]).all(1)
        assert np.array_equal(ridx[keep], ei[lo:hi][keep]) and np.array_equal(rdist[keep], ed[lo:hi][keep])
        assert np.array_equal(rcnt[keep], ec[lo:hi][keep]) and np.all(rcnt[~keep] == 0) and np.all(ridx[~keep] == -1)
        old = D.MASKED_REVERSE_MIN_PAIRS
        D.MASKED_REVERSE_MIN_PAIRS = 0
        try:
            rec, n_out, _ = be.match_device(k, M.MODE_MUTUAL)
        finally:
            D.MASKED_REVERSE_MIN_PAIRS = old
        got = D.records_to_numpy(rec, n_out)
        exp, _ = orc.match(_dense(src, dim), _dense(tgt, dim), k, "mutual", 1.1, np.float32(M.FLT_MAX))
        assert got.tobytes() == exp.tobytes()
    finally:
        be.close()
