"""Seeded fuzz of the C-ABI kNN against the oracle: descriptor lengths on the padding / K-atom edges (1 .. 700, the last
ones beyond the tensor-core pass), k up to 32 (k > 16 takes the exact CUDA-core path), odd row counts around the tile
sizes, data with duplicates (exact ties), far-from-origin values (stress on the centring), tiny and huge magnitudes
(unusable for FP16: exact path), NaN / Inf rows, strided AoS rows.  Everything must be bit-identical."""
import numpy as np
import pytest

from lidar_global_registration_b200 import build as b200_build
from lidar_global_registration_b200 import matcher as M
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    b200_build.build()
    M.load_library()


DIMS = [1, 2, 3, 13, 16, 33, 45, 61, 62, 64, 77, 125, 126, 135, 200, 352, 353, 500, 637, 638, 700]
KINDS = ["uniform", "clustered", "duplicates", "offset", "tiny", "huge", "sparse", "integers", "near_ties"]


def _make(rng, kind, n, dim):
    if kind == "uniform":
        a = rng.random((n, dim))
    elif kind == "clustered":
        protos = rng.random((max(n // 40, 2), dim))
        a = protos[rng.integers(0, protos.shape[0], n)] + 0.01 * rng.standard_normal((n, dim))
    elif kind == "duplicates":          # many exactly equal rows: ties at every rank
        base = rng.random((max(n // 5, 1), dim))
        a = base[rng.integers(0, base.shape[0], n)]
    elif kind == "offset":              # spread 1 around 1e4: the common centre has to absorb the offset
        a = 1e4 + rng.random((n, dim))
    elif kind == "tiny":
        a = 1e-32 * rng.random((n, dim))
    elif kind == "huge":
        a = 1e32 * rng.random((n, dim))
    elif kind == "near_ties":           # hundreds of rows within 1e-6 of each other: far below FP16 resolution, so the
        base = rng.random((max(n // 200, 1), dim))   # candidate lists overflow and the exact fallback has to answer
        a = base[rng.integers(0, base.shape[0], n)] + 1e-6 * rng.standard_normal((n, dim))
    elif kind == "sparse":
        a = rng.random((n, dim)) * (rng.random((n, dim)) < 0.2)
    else:                               # small integers: lots of equal distances between different rows
        a = rng.integers(0, 3, (n, dim)).astype(np.float64)
    return a.astype(np.float32)


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    dim = DIMS[seed % len(DIMS)]
    kind = KINDS[(seed // 3) % len(KINDS)]
    nq = int(rng.choice([1, 7, 127, 128, 129, 255, 257, 300, 777, 1500]))
    nt = int(rng.choice([1, 5, 255, 256, 257, 511, 513, 1000, 2049, 3000]))
    k = int(rng.choice([1, 2, 3, 4, 5, 8, 9, 16, 17, 32]))
    stride = dim + int(rng.choice([0, 0, 1, 9]))          # AoS padding (e.g. SHOT352's rf[9])
    q = np.zeros((nq, stride), np.float32)
    t = np.zeros((nt, stride), np.float32)
    both = _make(rng, kind, nq + nt, dim)
    q[:, :dim], t[:, :dim] = both[:nq], both[nq:]
    q[:, dim:] = np.nan                                   # whatever sits behind the descriptor must not matter
    t[:, dim:] = np.inf
    for a in (q, t):                                      # a few invalid rows
        bad = rng.random(a.shape[0]) < 0.03
        a[bad, rng.integers(0, dim)] = rng.choice([np.nan, np.inf, -np.inf])
    return q, t, dim, k, kind


@pytest.mark.parametrize("seed", range(81))
def test_knn_fuzz(seed):
    q, t, dim, k, kind = _case(seed)
    with M.Context(0) as ctx:
        ctx.upload(0, q, dim)
        ctx.upload(1, t, dim)
        got = ctx.knn(k, 0)
        got_rev = ctx.knn(k, 1)
        exact = ctx.knn(k, 0, precision=M.PREC_F32_EXACT)
    exp = orc.knn(np.ascontiguousarray(q[:, :dim]), np.ascontiguousarray(t[:, :dim]), k)
    exp_rev = orc.knn(np.ascontiguousarray(t[:, :dim]), np.ascontiguousarray(q[:, :dim]), k)
    for a, b, c in zip(got, exp, exact):
        assert np.array_equal(a, b), "tensor-core path differs from the oracle (%s, dim %d, k %d)" % (kind, dim, k)
        assert np.array_equal(c, b), "exact path differs from the oracle (%s, dim %d, k %d)" % (kind, dim, k)
    for a, b in zip(got_rev, exp_rev):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("seed", range(12))
def test_match_fuzz(seed):
    q, t, dim, k, kind = _case(100 + seed)
    k = max(2, min(k, 8))
    mode = ["one_sided", "mutual", "ratio"][seed % 3]
    name = {"one_sided": M.MODE_ONE_SIDED, "mutual": M.MODE_MUTUAL, "ratio": M.MODE_RATIO}[mode]
    with M.Context(0) as ctx:
        ctx.upload(0, q, dim)
        ctx.upload(1, t, dim)
        got, avg = ctx.match(k, name, 1.1, np.float32(0.9))
    exp, eavg = orc.match(np.ascontiguousarray(q[:, :dim]), np.ascontiguousarray(t[:, :dim]), k, mode, 1.1, np.float32(0.9))
    assert got.tobytes() == exp.tobytes()
    assert avg == eavg or (np.isnan(avg) and np.isnan(eavg))
