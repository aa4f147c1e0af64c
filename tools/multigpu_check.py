"""Multi-GPU parity on real GPUs (run under torchrun, one rank per GPU):
  * query-sharded / target-replicated matcher (mutual, ratio): concatenated rank slices == oracle records
  * target-sharded kNN (per-rank exact top-k, NCCL all-gather, merge kernel) == oracle k-lists
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/multigpu_check.py
Nothing here reads /root/reference; the oracle is the checker only."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_global_registration_b200 import device as D  # noqa: E402
from lidar_global_registration_b200 import matcher as M  # noqa: E402
from lidar_global_registration_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
dist.init_process_group("nccl", device_id=dev)
be = D.GpuBackend(local_rank)
sm = D.ShardedMatcher(be, rank, world, None)
ok = True
for desc, nq, nt, k in [("fpfh", 5003, 7001, 2), ("shot", 2111, 3005, 2), ("rops", 1500, 2200, 5)]:
    src, tgt, dim = synth.make_pair(desc, nq, nt, nan_frac=0.01)
    s_d, t_d = torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev)
    sd, td = np.ascontiguousarray(src[:, :dim]), np.ascontiguousarray(tgt[:, :dim])
    # --- query-sharded matcher ---
    # (descriptors replicated from pinned HOST memory: 1/world slice per rank over PCIe + NCCL all-gather)
    sm.upload_host_sharded(0, torch.from_numpy(src).pin_memory(), dim)
    sm.upload_host_sharded(1, torch.from_numpy(tgt).pin_memory(), dim)
    for mode, oname in [(M.MODE_MUTUAL, "mutual"), (M.MODE_RATIO, "ratio"), (M.MODE_MUTUAL, "mutual-masked")]:
        # "mutual-masked": the reverse pass answers only the target rows that forward lists name (flags max-reduced over
        # the ranks with NCCL) -- what large runs do by default
        D.MASKED_REVERSE_MIN_PAIRS = 0 if oname == "mutual-masked" else 10 ** 18
        rec, n_out = sm.match_query_sharded(k, mode)[:2]
        allrec, n_all = sm.gather_records(rec, n_out)
        got = allrec[:n_all].contiguous().cpu().numpy().view(M.CORR_DTYPE).reshape(-1)
        if rank == 0:
            exp = orc.match(sd, td, k, oname.split("-")[0], M.MATCHING_RATIO_THRESHOLD, np.float32(M.FLT_MAX))[0]
            same = got.shape == exp.shape and all(np.array_equal(got[f], exp[f]) for f in got.dtype.names)
            print("query-sharded %-13s %s %dx%d k=%d world=%d: %d records %s" % (oname, desc, nq, nt, k, world, n_all,
                                                                           "PASS" if same else "FAIL"))
            if not same:
                print("   expected %d records; differing fields: %s" % (exp.shape[0], [f for f in got.dtype.names
                      if got.shape != exp.shape or not np.array_equal(got[f], exp[f])]))
            ok = ok and same
    # --- target-sharded kNN ---
    t0, t1 = D.shard_bounds(nt, rank, world)
    be.upload_device(0, s_d, dim)
    be.upload_device(1, t_d[t0:t1].contiguous(), dim, index_offset=t0)
    idx, dst, cnt = sm.knn_target_sharded(k)
    if rank == 0:
        e = orc.knn(sd, td, k)
        g = (idx.cpu().numpy(), dst.cpu().numpy(), cnt.cpu().numpy())
        same = all(np.array_equal(a, b) for a, b in zip(g, e))
        print("target-sharded kNN   %s %dx%d k=%d world=%d: %s" % (desc, nq, nt, k, world, "PASS" if same else "FAIL"))
        ok = ok and same
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
be.close()
sys.exit(0 if int(flag.item()) else 1)
