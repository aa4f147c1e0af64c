"""Multi-GPU parity on real GPUs (run under torchrun, one rank per GPU); the partitioning and the NCCL exchange run INSIDE
libb200match.so (csrc/multi.cu), torch.distributed only carries the communicator id:
  * query-sharded / target-replicated matcher (mutual, ratio; full and masked reverse pass): rank slices in rank order ==
    oracle records; the average over all source rows == the oracle's
  * target-sharded kNN (per-rank exact top-k, all-gather, merge kernel) == oracle k-lists
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/multigpu_check.py
    ... tools/multigpu_check.py c5 [rows]     BASELINE configs[4] at full size (1M target rows per rank): sampled rows vs the
                                              oracle over the WHOLE target set, plus properties on all rows
Nothing here reads /root/reference; the oracle is the checker only."""
import os
import sys
import time

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
ok = True


def gather_slices(sm, rec, n_out):
    allrec, n_all = sm.gather_records(rec, n_out)
    return allrec[:n_all].contiguous().cpu().numpy().view(M.CORR_DTYPE).reshape(-1)


if len(sys.argv) > 1 and sys.argv[1] == "c5":
    from bench import WORKLOADS
    n_sample = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    desc, n_src, n_tgt, k, _, cfg = WORKLOADS["c5"]
    be = D.GpuBackend(local_rank)
    sm = D.ShardedMatcher(be, rank, world, None)
    src, tgt, dim = synth.make_pair_torch(desc, n_src, n_tgt, dev)
    if rank > 0:   # bench.py's recipe: a different shard per rank
        tgt = (torch.roll(tgt, shifts=1237 * rank, dims=0) * (1.0 + 2e-3 * rank)).contiguous()
    be.upload_device(0, src, dim)
    be.upload_device(1, tgt, dim, index_offset=rank * n_tgt)
    t0 = time.time()
    idx, dst, cnt = sm.knn_target_sharded(k)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    idx2, dst2, cnt2 = sm.knn_target_sharded(k)
    same = bool(torch.equal(idx, idx2) and torch.equal(dst, dst2) and torch.equal(cnt, cnt2))
    # the whole target set on rank 0's host (world x 1.4 GB) for the oracle
    shard_h = np.ascontiguousarray(tgt[:, :dim].cpu().numpy())
    parts = [None] * world
    dist.gather_object(shard_h, parts if rank == 0 else None, dst=0)
    if rank == 0:
        full = np.concatenate(parts)
        del parts
        idx_h, dst_h, cnt_h = idx.cpu().numpy(), dst.cpu().numpy(), cnt.cpu().numpy()
        q_h = src[:, :dim].cpu().numpy()
        good = np.isfinite(q_h).all(1)
        props = bool(np.all(cnt_h[good] == k) and np.all(cnt_h[~good] == 0) and np.all(np.diff(dst_h[good], axis=1) >= 0)
                     and np.all(idx_h[good] >= 0) and np.all(idx_h[good] < world * n_tgt))
        rows = np.sort(np.random.default_rng(5).choice(n_src, n_sample, replace=False))
        t0 = time.time()
        e = orc.knn(np.ascontiguousarray(q_h[rows]), full, k)
        t_cpu = time.time() - t0
        exact = all(np.array_equal(a[rows], b) for a, b in zip((idx_h, dst_h, cnt_h), e))
        print("c5 target-sharded: %d queries x %d target rows (%d ranks x %d), D=%d, k=%d | GPU %.3f s | idempotent %s | "
              "properties %s | %d sampled rows vs oracle over the whole target (%.1f s CPU): %s" % (
                  n_src, world * n_tgt, world, n_tgt, dim, k, t_gpu, same, props, n_sample, t_cpu,
                  "bit-exact" if exact else "MISMATCH"))
        ok = same and props and exact
    be.close()
else:
    for desc, nq, nt, k in [("fpfh", 5003, 7001, 2), ("shot", 2111, 3005, 2), ("rops", 1500, 2200, 5)]:
        src, tgt, dim = synth.make_pair(desc, nq, nt, nan_frac=0.01)
        s_d, t_d = torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev)
        sd, td = np.ascontiguousarray(src[:, :dim]), np.ascontiguousarray(tgt[:, :dim])
        for masked in (False, True):
            # masked: the reverse pass answers only the target rows that forward lists name (flags max-reduced over the
            # ranks with NCCL) -- what large runs do by default
            os.environ["B200M_MASKED_MIN_PAIRS"] = "1" if masked else "1e30"
            be = D.GpuBackend(local_rank)
            sm = D.ShardedMatcher(be, rank, world, None)
            # descriptors replicated from pinned HOST memory: 1/world slice per rank over PCIe + NVLink all-gather
            sm.upload_host_sharded(0, torch.from_numpy(src).pin_memory(), dim)
            sm.upload_host_sharded(1, torch.from_numpy(tgt).pin_memory(), dim)
            for mode, oname in [(M.MODE_MUTUAL, "mutual"), (M.MODE_RATIO, "ratio"), (M.MODE_ONE_SIDED, "one_sided")]:
                rec, n_out = sm.match_query_sharded(k, mode)[:2]
                got = gather_slices(sm, rec, n_out)
                # host-buffer form: this rank's slice + the average over ALL source rows
                mine, avg = be.ctx.match_sharded(k, mode)
                if rank == 0:
                    exp, eavg = orc.match(sd, td, k, oname, M.MATCHING_RATIO_THRESHOLD, np.float32(M.FLT_MAX))
                    q0, q1 = D.shard_bounds(nq, 0, world)
                    exp0 = exp[(exp["index_query"] >= q0) & (exp["index_query"] < q1)]
                    same = got.tobytes() == exp.tobytes() and mine.tobytes() == exp0.tobytes() and avg == eavg
                    print("query-sharded %-9s %s %s %dx%d k=%d world=%d: %d records %s" % (
                        oname, "masked" if masked else "full  ", desc, nq, nt, k, world, len(got), "PASS" if same else "FAIL"))
                    ok = ok and same
            if not masked:   # target-sharded kNN
                t0, t1 = D.shard_bounds(nt, rank, world)
                be.upload_device(0, s_d, dim)
                be.upload_device(1, t_d[t0:t1].contiguous(), dim, index_offset=t0)
                idx, dst, cnt = sm.knn_target_sharded(k)
                if rank == 0:
                    e = orc.knn(sd, td, k)
                    g = (idx.cpu().numpy(), dst.cpu().numpy(), cnt.cpu().numpy())
                    same = all(np.array_equal(a, b) for a, b in zip(g, e))
                    print("target-sharded kNN      %s %dx%d k=%d world=%d: %s" % (desc, nq, nt, k, world, "PASS" if same else "FAIL"))
                    ok = ok and same
            be.close()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
