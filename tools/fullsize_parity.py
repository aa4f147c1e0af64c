"""Parity at BASELINE.json's full sizes (run on a B200): the device-resident kNN of the whole problem, checked
  * bit for bit (indices, distances, counts) against the CPU oracle on a random subsample of query rows vs the FULL
    train set, both directions, and
  * through size-independent properties on ALL rows: ascending distances, counts, index range, idempotence (a second
    run gives the same checksum), row-range consistency.
    python tools/fullsize_parity.py [c3|c4|c2] [rows_sampled]
The oracle is the checker only; nothing here reads /root/reference."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from lidar_global_registration_b200 import device as D  # noqa: E402
from lidar_global_registration_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
n_sample = int(sys.argv[2]) if len(sys.argv) > 2 else 384
desc, n_src, n_tgt, k, mode_name, cfg = WORKLOADS[wl]
be = D.GpuBackend(0)
src, tgt, dim = synth.make_pair_torch(desc, n_src, n_tgt, be.device)
be.upload_device(0, src, dim)
be.upload_device(1, tgt, dim)
ok = True
src_h, tgt_h = src.cpu().numpy(), tgt.cpu().numpy()
for direction, (q_h, t_h, nq, nt) in enumerate([(src_h, tgt_h, n_src, n_tgt), (tgt_h, src_h, n_tgt, n_src)]):
    t0 = time.time()
    idx, dist, cnt = be.knn(k, direction, 0, nq)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    idx2, dist2, cnt2 = be.knn(k, direction, 0, nq)
    same = bool(torch.equal(idx, idx2) and torch.equal(dist, dist2) and torch.equal(cnt, cnt2))
    lo = nq // 3
    sub = be.knn(k, direction, lo, lo + 1000)
    ranges = bool(torch.equal(sub[0], idx[lo:lo + 1000]) and torch.equal(sub[1], dist[lo:lo + 1000]))
    idx_h, dist_h, cnt_h = idx.cpu().numpy(), dist.cpu().numpy(), cnt.cpu().numpy()
    good = np.isfinite(q_h[:, :dim]).all(1)
    props = bool(np.all(cnt_h[good] == k) and np.all(cnt_h[~good] == 0) and np.all(np.diff(dist_h[good], axis=1) >= 0)
                 and np.all(idx_h[good] >= 0) and np.all(idx_h[good] < nt))
    rows = np.sort(np.random.default_rng(7 + direction).choice(nq, n_sample, replace=False))
    t0 = time.time()
    e = orc.knn(np.ascontiguousarray(q_h[rows, :dim]), np.ascontiguousarray(t_h[:, :dim]), k)
    t_cpu = time.time() - t0
    exact = all(np.array_equal(a[rows], b) for a, b in zip((idx_h, dist_h, cnt_h), e))
    print("%s direction %d: %d x %d, D=%d, k=%d | GPU %.3f s | idempotent %s | row ranges %s | properties %s | "
          "%d sampled rows vs oracle (%.1f s CPU): %s" % (wl, direction, nq, nt, dim, k, t_gpu, same, ranges, props, n_sample,
                                                           t_cpu, "bit-exact" if exact else "MISMATCH"))
    ok = ok and same and ranges and props and exact
be.close()
sys.exit(0 if ok else 1)
