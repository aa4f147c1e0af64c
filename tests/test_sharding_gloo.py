"""Host-side multi-GPU logic (lidar_global_registration_b200/device.py: ShardedMatcher) on CPU: two gloo ranks and a
stand-in backend that answers kNN/filter/merge with the CPU oracle.  Checks that sharding + the one all-gather +
per-rank filtering reproduce the single-process result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lidar_global_registration_b200 import device as D
from lidar_global_registration_b200 import matcher as M
from lidar_global_registration_b200 import synth
from oracle import oracle as orc


class OracleBackend:
    """Same interface as device.GpuBackend, computed by the oracle on CPU tensors (tests only)."""

    def __init__(self, src, tgt, dim, tgt_offset=0):
        self.s = np.ascontiguousarray(src[:, :dim])
        self.t = np.ascontiguousarray(tgt[:, :dim])
        self.n = [self.s.shape[0], self.t.shape[0]]
        self.tgt_offset = tgt_offset

    def knn(self, k, direction, row_begin, row_end):
        q, t = (self.s, self.t) if direction == 0 else (self.t, self.s)
        idx, dist_, cnt = orc.knn(q[row_begin:row_end], t, k)
        if direction == 0 and self.tgt_offset:
            idx = np.where(idx >= 0, idx + self.tgt_offset, idx).astype(np.int32)
        return torch.from_numpy(idx), torch.from_numpy(dist_), torch.from_numpy(cnt)

    def referenced_rows(self, k, fwd, n_flags, index_offset=0):
        idx, cnt = fwd[0].numpy(), fwd[2].numpy()
        flags = np.zeros(n_flags, np.uint8)
        for i in range(idx.shape[0]):
            flags[idx[i, :cnt[i]] - index_offset] = 1
        return torch.from_numpy(flags)

    def knn_masked(self, k, direction, row_begin, row_end, flags):
        idx, dist_, cnt = [x.numpy().copy() for x in self.knn(k, direction, row_begin, row_end)]
        skip = flags.numpy()[row_begin:row_end] == 0
        idx[skip], dist_[skip], cnt[skip] = -1, 0, 0
        return torch.from_numpy(idx), torch.from_numpy(dist_), torch.from_numpy(cnt)

    def filter(self, k, mode, row_begin, row_end, fwd, rev, n_rev_rows, ratio_thr=1.1, distance_thr=M.FLT_MAX,
               thr_src=None, thr_tgt=None, want_avg=False):
        nq = self.n[0]
        fi = np.full((nq, k), -1, np.int32); fd = np.zeros((nq, k), np.float32); fc = np.zeros(nq, np.int32)
        fi[row_begin:row_end], fd[row_begin:row_end], fc[row_begin:row_end] = [x.numpy() for x in fwd]
        dthr = np.float32(distance_thr)
        if mode == M.MODE_MUTUAL:
            ri, rd, rc = [np.ascontiguousarray(x.numpy()) for x in rev]
            out = orc.filter_mutual(fi, fc, ri, rd, rc, dthr)
        elif mode == M.MODE_RATIO:
            out = orc.filter_ratio(fi, fd, fc, ratio_thr, dthr)
        else:
            out = orc.filter_one_sided(fi, fd, fc, dthr)
        rec = torch.from_numpy(out.view(np.int32).reshape(-1, 4).copy())
        cap = max((row_end - row_begin) * (k if mode == M.MODE_MUTUAL else 1), 1)
        pad = torch.zeros((cap, 4), dtype=torch.int32)
        pad[:rec.shape[0]] = rec
        return pad, torch.tensor([rec.shape[0]], dtype=torch.int64), None

    def merge(self, k, idx_in, dist_in, cnt_in):
        n_lists, nq = cnt_in.shape
        idx = np.full((nq, k), -1, np.int32); dist_ = np.zeros((nq, k), np.float32); cnt = np.zeros(nq, np.int32)
        ii, dd, cc = idx_in.numpy(), dist_in.numpy(), cnt_in.numpy()
        for q in range(nq):
            items = sorted((float(dd[l, q, m]), int(ii[l, q, m])) for l in range(n_lists) for m in range(cc[l, q]))[:k]
            cnt[q] = len(items)
            for m, (d, i) in enumerate(items):
                idx[q, m], dist_[q, m] = i, np.float32(d)
        return torch.from_numpy(idx), torch.from_numpy(dist_), torch.from_numpy(cnt)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, desc, ns, nt, k, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        src, tgt, dim = synth.make_pair(desc, ns, nt, nan_frac=0.01)
        res = {}
        # query-sharded / target replicated
        sm = D.ShardedMatcher(OracleBackend(src, tgt, dim), rank, world)
        for name, mode in (("mutual", M.MODE_MUTUAL), ("ratio", M.MODE_RATIO), ("one_sided", M.MODE_ONE_SIDED)):
            if name == "ratio" and k < 2:
                continue
            rec, n_out, _ = sm.match_query_sharded(k, mode)
            allrec, n = sm.gather_records(rec, n_out)
            res[name] = allrec[:n].numpy().view(M.CORR_DTYPE).reshape(-1).copy()
        # the same mutual run with the reverse pass restricted to the target rows the forward lists name
        # (flags max-reduced over the ranks): identical records
        D.MASKED_REVERSE_MIN_PAIRS = 0
        rec, n_out, _ = sm.match_query_sharded(k, M.MODE_MUTUAL)
        allrec, n = sm.gather_records(rec, n_out)
        res["mutual_masked"] = allrec[:n].numpy().view(M.CORR_DTYPE).reshape(-1).copy()
        D.MASKED_REVERSE_MIN_PAIRS = 10 ** 9
        # target-sharded: this rank holds target rows [t0, t1)
        t0, t1 = D.shard_bounds(nt, rank, world)
        sm2 = D.ShardedMatcher(OracleBackend(src, tgt[t0:t1], dim, tgt_offset=t0), rank, world)
        res["tknn"] = tuple(x.numpy().copy() for x in sm2.knn_target_sharded(k))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("desc,ns,nt,k", [("fpfh", 301, 411, 2), ("shot", 150, 97, 1)])
def test_sharded_matcher_two_ranks_gloo(desc, ns, nt, k):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, desc, ns, nt, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    src, tgt, dim = synth.make_pair(desc, ns, nt, nan_frac=0.01)
    s_d, t_d = src[:, :dim], tgt[:, :dim]
    fmax = np.float32(M.FLT_MAX)
    for name in ("mutual", "ratio", "one_sided"):
        if name == "ratio" and k < 2:
            continue
        exp, _ = orc.match(s_d, t_d, k, name, 1.1, fmax)
        for r in range(2):
            assert got[r][name].tobytes() == exp.tobytes(), (name, r)
            if name == "mutual":
                assert got[r]["mutual_masked"].tobytes() == exp.tobytes()
        assert len(exp) > 0
    e = orc.knn(s_d, t_d, k)
    for r in range(2):
        for a, b in zip(got[r]["tknn"], e):
            assert np.array_equal(a, b)


def test_shard_bounds_cover_and_match_the_library():
    """Contiguous cover of [0, n); every rank but the last non-empty one holds ceil(n / world) rows (so that all-gathered
    slots are the row-major table of all rows); identical to the library's own partition (b200m_shard_rows: host code)."""
    from lidar_global_registration_b200 import build as b200_build
    from lidar_global_registration_b200 import matcher as M
    b200_build.build()
    for n in (0, 1, 7, 8, 10, 500000, 500001):
        for w in (1, 2, 3, 8):
            b = [D.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            per = -(-n // w)
            assert all(hi - lo == per for lo, hi in b if hi < n)
            assert b == [M.shard_rows(n, w, r) for r in range(w)]
