"""The C++ host mirror (include/b200match_shim.hpp): compiles against the C-ABI on CPU; on the GPU box the
reference-style C++ program (tests/cpp/shim_test.cpp) must reproduce the oracle."""
import os
import subprocess

import numpy as np
import pytest

from lidar_global_registration_b200 import build as b200_build
from lidar_global_registration_b200 import synth
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "shim_test")


def _build_exe():
    b200_build.build()
    src = os.path.join(ROOT, "tests", "cpp", "shim_test.cpp")
    deps = [src, os.path.join(ROOT, "include", "b200match_shim.hpp"), os.path.join(ROOT, "include", "b200match.h")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return EXE
    libdir = os.path.join(ROOT, "lidar_global_registration_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                    "-L", libdir, "-l:libb200match.so", "-Wl,-rpath," + libdir], check=True)
    return EXE


def test_shim_compiles_and_links():
    exe = _build_exe()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.parametrize("desc,ns,nt,k", [("fpfh", 700, 900, 2), ("shot", 300, 420, 1), ("rops", 280, 300, 3)])
def test_cpp_shim_matches_oracle(tmp_path, desc, ns, nt, k):
    exe = _build_exe()
    src, tgt, dim = synth.make_pair(desc, ns, nt, nan_frac=0.01)
    sp, tp, op = tmp_path / "s.bin", tmp_path / "t.bin", tmp_path / "o.txt"
    src.tofile(sp)
    tgt.tofile(tp)
    n_gpus = min(_n_gpus(), 8)     # with several GPUs visible the program also runs its multi-GPU section (b200m_create_multi)
    r = subprocess.run([exe, desc, str(sp), str(ns), str(tp), str(nt), str(k), str(op)] + ([str(n_gpus)] if n_gpus > 1 else []),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    s_d, t_d = src[:, :dim], tgt[:, :dim]
    exp = {0: orc.knn(s_d, t_d, k), 1: orc.knn(t_d, s_d, k)}
    got_corr = {"one_sided": [], "lr": [], "ratio": [], "cluster": [], "lr2": [], "m_one_sided": [], "m_lr": []}
    got_ms, got_local, got_mknn = {}, {}, {}
    heads = {}
    n_knn = 0
    for line in open(op):
        w = line.split()
        if w[0] == "knn":
            d, i, c = int(w[1]), int(w[2]), int(w[3])
            idx, dist, cnt = exp[d]
            assert c == cnt[i]
            assert [int(x) for x in w[4::2]] == idx[i, :c].tolist()
            assert [np.float32(x) for x in w[5::2]] == dist[i, :c].tolist()
            n_knn += 1
        elif w[0] == "matcher":
            heads[w[1]] = (w[2], int(w[3]), np.float32(w[4]))
        elif w[0] == "corr":
            got_corr[w[1]].append((int(w[2]), int(w[3]), np.float32(w[4])))
        elif w[0] == "ms":
            got_ms[int(w[1])] = [(int(a), np.float32(b)) for a, b in zip(w[3::2], w[4::2])]
        elif w[0] == "mknn":
            got_mknn[int(w[1])] = [(int(a), np.float32(b)) for a, b in zip(w[3::2], w[4::2])]
        elif w[0] == "local":
            got_local[int(w[1])] = [(int(a), np.float32(b)) for a, b in zip(w[3::2], w[4::2])]
    assert n_knn == ns + nt
    fmax = np.float32(np.finfo(np.float32).max)

    # the same keypoint lattice shim_test.cpp builds
    def lattice(n, step):
        i = np.arange(n)
        return np.stack([step * (i % 17), step * ((i // 17) % 13), step * (i // 221)], 1).astype(np.float32)
    sx, tx = lattice(ns, np.float32(0.5)), lattice(nt, np.float32(0.25))
    iss_s, iss_t = np.float32(0.3), np.float32(0.2)

    def triples(e):
        return [(int(a), int(b), np.float32(c)) for a, b, c in zip(e["index_query"], e["index_match"], e["distance"])]
    # matcher classes: match_impl as the reference composes it (match_multiscale both ways + vote + filter)
    one = ([(np.ascontiguousarray(s_d), None)], [(np.ascontiguousarray(t_d), None)])
    for mid, mode, cls in (("one_sided", "one_sided", "OneSidedMatcher"), ("lr", "mutual", "LeftToRightMatcher"),
                           ("cluster", "cluster", "ClusterMatcher")):
        e, eavg = orc.match_wide(mode, one[0], one[1], sx, tx, iss_s, iss_t, k, 12, fmax)
        assert heads[mid][0] == cls and heads[mid][1] == len(e) and heads[mid][2] == np.float32(eavg)
        assert got_corr[mid] == triples(e)
    if k >= 2:   # the ratio filter stays on the raw k-lists (a stub in the reference)
        e, eavg = orc.match(s_d, t_d, k, "ratio", 1.1, fmax)
        assert heads["ratio"][0] == "RatioMatcher" and heads["ratio"][1] == len(e) and heads["ratio"][2] == np.float32(eavg)
        assert got_corr["ratio"] == triples(e)
    # two scales: the narrow-seam match_multiscale and the LeftToRightMatcher over the same Storage, finalize included
    q2, t2 = np.arange(0, ns, 2).astype(np.int32), np.arange(0, nt, 2).astype(np.int32)
    s_sc = [(np.ascontiguousarray(s_d), np.arange(ns, dtype=np.int32)), (np.ascontiguousarray(s_d[q2]), q2)]
    t_sc = [(np.ascontiguousarray(t_d), np.arange(nt, dtype=np.int32)), (np.ascontiguousarray(t_d[t2]), t2)]
    vi, vd, vc = orc.match_multiscale(s_sc, t_sc, ns, tx, np.float32(0.3), k)
    assert len(got_ms) == ns
    for i in range(ns):
        assert got_ms[i] == ([(int(vi[i, 0]), np.float32(vd[i, 0]))] if vc[i] else [])
    e, eavg = orc.match_wide("mutual", s_sc, t_sc, sx, tx, iss_s, iss_t, k, 12, fmax)
    e = orc.finalize(e, (3 * np.arange(ns) + 1).astype(np.int32), (2 * np.arange(nt) + 5).astype(np.int32))
    assert heads["lr2"][0] == "LeftToRightMatcher" and heads["lr2"][1] == len(e) and heads["lr2"][2] == np.float32(eavg)
    assert got_corr["lr2"] == triples(e)
    # the gated matchLocal: query keypoints moved by the guess (a translation), radius 1.75
    g = np.array([0.25, -0.5, 0.125], np.float32)
    moved = ((np.float32(1) * sx + np.float32(0)) + g).astype(np.float32)   # 1*x + 0*y + 0*z + t in float32
    li, ld, lc = orc.match_local(s_d, t_d, k, moved, tx, np.float32(1.75))
    assert len(got_local) == ns and lc.min() < k   # the gate bites somewhere
    for i in range(ns):
        assert got_local[i] == [(int(li[i, m]), np.float32(ld[i, m])) for m in range(lc[i])]
    if n_gpus > 1:   # the multi-GPU section: same answers as one GPU / the oracle
        idx, dist, cnt = exp[0]
        assert len(got_mknn) == ns
        for i in range(ns):
            assert got_mknn[i] == [(int(idx[i, m]), np.float32(dist[i, m])) for m in range(cnt[i])]
        for mid, mode in (("m_one_sided", "one_sided"), ("m_lr", "mutual")):
            e, eavg = orc.match(s_d, t_d, 1, mode, 1.1, fmax)
            assert heads[mid][1] == len(e) and heads[mid][2] == np.float32(eavg)
            assert got_corr[mid] == triples(e)
