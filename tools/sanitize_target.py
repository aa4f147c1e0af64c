"""Small, fixed program for compute-sanitizer (one tool per process): every kernel family of the hot path on inputs a
sanitised run finishes in a minute or two -- the candidate kernel in pair / split-N / alternating-tile mode (FPFH), pair
mode with a four-warp epilogue (SHOT, RoPS), the multicast modes (B200M_TC_MODE=mcast, B200M_TC_CLUSTER=2|4 in the
environment), the re-rank, the overflow fallback (a tiny candidate capacity forces exact_rows_split_kernel), the masked
reverse pass and the filters.  Results are compared with the oracle so a sanitised run is also a parity run.
    compute-sanitizer --tool memcheck python tools/sanitize_target.py [fpfh|shot|rops|all]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_global_registration_b200 import matcher as M  # noqa: E402
from lidar_global_registration_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
cases = [("fpfh", 1500, 2100, 2), ("fpfh", 700, 900, 5), ("shot", 600, 800, 2), ("rops", 500, 700, 3)]
ok = True
for desc, nq, nt, k in cases:
    if which not in ("all", desc):
        continue
    src, tgt, dim = synth.make_pair(desc, nq, nt, nan_frac=0.01)
    exp = orc.knn(src[:, :dim], tgt[:, :dim], k)
    exp_m, _ = orc.match(src[:, :dim], tgt[:, :dim], k, "mutual", distance_thr=np.float32(M.FLT_MAX))
    os.environ["B200M_MASKED_MIN_PAIRS"] = "1"   # the masked reverse pass at this size too
    with M.Context(0) as ctx:
        ctx.upload(0, src, dim)
        ctx.upload(1, tgt, dim)
        got = ctx.knn(k, 0)
        same = all(np.array_equal(a, b) for a, b in zip(got, exp))
        got_small = ctx.knn(k, 0, cand_cap=k)     # overflowing candidate lists: exact fallback kernels
        same_small = all(np.array_equal(a, b) for a, b in zip(got_small, exp))
        rec, _ = ctx.match(k, M.MODE_MUTUAL)
        same_m = rec.tobytes() == exp_m.tobytes()
        st = ctx.stats()
    print("%s %dx%d k=%d: knn %s | overflow path %s | mutual (masked reverse) %s | launches %d" % (
        desc, nq, nt, k, same, same_small, same_m, st["launches"]))
    ok = ok and same and same_small and same_m
print("SANITIZE_TARGET_OK" if ok else "SANITIZE_TARGET_MISMATCH")
sys.exit(0 if ok else 1)
