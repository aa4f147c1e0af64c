#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 descriptor matcher (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl b200|reference]

A "step" is one whole matcher call over one synthetic descriptor pair of the named BASELINE.json
configuration: descriptor upload/packing + forward kNN (+ reverse kNN for mutual) + filter.
  value : queries/s, descriptors (AoS, as PCL lays them out) already resident in HBM; CUDA events.
  e2e   : the same through the host-facing C-ABI (b200m_upload from pinned host memory, result
          records copied back to the host) -- host<->device copies inside the timed region.
Multi-GPU (torchrun, one rank per GPU): source rows (and, for the mutual pass, target rows) are sharded
across ranks, both descriptor sets replicated; one NCCL all-gather of the reverse table; strong scaling.
`--impl reference` times the reference's CPU matcher semantics (the oracle port; the reference itself
cannot be built here: PCL/OpenCV/FLANN absent) on the host cores, bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (descriptor, n_src, n_tgt, k, mode, BASELINE.json config string)
    "c1": ("fpfh", 20000, 20000, 1, "mutual", "FPFH-33 20k x 20k k=1 mutual (BASELINE configs[0])"),
    "c2": ("fpfh", 200000, 200000, 2, "ratio", "FPFH-33 200k x 200k k=2 + ratio 1.1 (BASELINE configs[1])"),
    "c3": ("shot", 500000, 500000, 2, "mutual", "SHOT-352 500k x 500k k=2 + mutual (BASELINE configs[2])"),
    "c3s": ("shot", 40000, 40000, 2, "mutual", "SHOT-352 40k x 40k k=2 + mutual (reduced copy of configs[2], for debugging)"),
    "c4": ("fpfh", 2000000, 2000000, 5, "mutual", "FPFH-33 2M x 2M k=5 mutual k-lists (BASELINE configs[3])"),
    # configs[4]: the target set is sharded 1M rows per GPU (8M at 8 GPUs) and every rank holds all queries; per-rank exact
    # top-k with global indices, one NCCL all-gather of the k-lists, merge kernel.  n_tgt here is PER RANK (weak scaling).
    "c5": ("shot", 500000, 1000000, 2, "knn_target_sharded",
           "SHOT-352 kNN k=2, 500k queries x (1M target rows per GPU), target-sharded + NCCL top-k merge (BASELINE configs[4])"),
}
METRIC = "descriptor queries/sec (k=2 + mutual)"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE candidate-kernel launch, from the ncu --set full capture of the
# same workload on 1 GPU (profiles/r01b_ncu_c3_cand.txt, profiles/r01c_ncu_c2_cand.txt).  The kernel is tensor bound; the traffic is the FP16 train
# operand array streaming through L2 once per wave of query tiles (algorithmic operand bytes: 0.77 GB for c3).
NCU_TRAFFIC_BYTES_PER_LAUNCH = {"c3": 19.784716e9 + 219.15264e6, "c2": 61.222400e6 + 21.200384e6}
NCU_TRAFFIC_SOURCE = {"c3": "profiles/r01b_ncu_c3_cand.txt", "c2": "profiles/r01c_ncu_c2_cand.txt"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d.get("bf16_tflops_sustained", 1388.3), "bf16_burst": d.get("bf16_tflops", 1658.5),
                "hbm": d.get("hbm_gbs", 6450.3), "source": "MEASURED_PEAKS.json (of measured)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region.  The sampler process is started before the
    warm-up (nvidia-smi needs a few hundred ms before its first line) and streams a time-stamped line every 20 ms;
    stop() keeps the lines whose time stamps fall inside the timed region's host-clock window."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t_begin = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def begin(self):
        """Call right before the timed region starts (after the barrier)."""
        self.t_begin = time.time()

    @staticmethod
    def _stamp(text):
        import datetime
        return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()

    def stop(self):
        t_end = time.time()
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        parsed = []
        for r in rows:
            try:
                try:
                    ts = self._stamp(r[0])
                except Exception:
                    ts = t_end   # unparsable time stamp: count the line as inside the window
                parsed.append((ts, float(r[1]), float(r[2]), float(r[3]),
                               [n for n, v in zip(names, r[4:8]) if v.strip().lower() == "active"]))
            except Exception:
                pass
        if not parsed:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        t0 = self.t_begin if self.t_begin is not None else parsed[0][0]
        window = "timed region"
        inside = [q for q in parsed if t0 <= q[0] <= t_end]
        if not inside:
            # a timed region shorter than the sampling period: the samples closest to it (the warm-up runs the same
            # steps back to back, so these are still samples under load)
            mid = 0.5 * (t0 + t_end)
            inside = sorted(parsed, key=lambda q: abs(q[0] - mid))[:3]
            window = "nearest samples (timed region of %.0f ms is shorter than the sampling period)" % (1e3 * (t_end - t0))
        reasons = set()
        for q in inside:
            reasons.update(q[4])
        return {"sm_mhz": float(np.median([q[1] for q in inside])), "sm_max_mhz": float(max(q[2] for q in inside)),
                "power_w_max": float(max(q[3] for q in inside)), "samples": len(inside), "window": window,
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
class CpuReference:
    """Oracle (CPU port of the reference matcher) on a bounded query sample against the FULL opposite
    set, both directions for mutual.  Data are generated once per process."""

    def __init__(self, desc, n_src, n_tgt, k, mode):
        from lidar_global_registration_b200 import synth
        from oracle import oracle as orc
        self.orc, self.k, self.both = orc, k, mode == "mutual"
        try:
            orc.set_num_threads(len(os.sched_getaffinity(0)))   # all host cores, whatever OMP_NUM_THREADS torchrun exported
        except AttributeError:
            orc.set_num_threads(os.cpu_count() or 1)
        self.cores = orc.num_threads()
        self.n_src, self.n_tgt = n_src, n_tgt
        # numpy generation is slow at the named sizes: generate <= 200k rows and tile them with small
        # perturbations up to the named row counts (the CPU cost per query only depends on the row counts)
        n_gen_s, n_gen_t = min(n_src, 200000), min(n_tgt, 200000)
        src, tgt, dim = synth.make_pair(desc, n_gen_s, n_gen_t, nan_frac=0.001)
        rng = np.random.default_rng(11)

        def tile(a, n):
            a = np.ascontiguousarray(a[:, :dim])
            if a.shape[0] >= n:
                return a[:n]
            reps = (n + a.shape[0] - 1) // a.shape[0]
            return np.concatenate([a * np.float32(1 + 1e-3 * rng.standard_normal()) for _ in range(reps)])[:n]
        self.s, self.t, self.dim = tile(src, n_src), tile(tgt, n_tgt), dim
        self.probe = max(self.cores * 4, 64)
        self.t_probe = self._run(self.probe)

    def _run(self, nq):
        t0 = time.perf_counter()
        self.orc.knn(self.s[:nq], self.t, self.k)
        if self.both:
            self.orc.knn(self.t[:nq], self.s, self.k)
        return time.perf_counter() - t0

    def rate(self, budget_s):
        nq = int(max(self.probe, min(self.n_src, self.probe * budget_s / max(self.t_probe, 1e-6))))
        secs = self._run(nq)
        for _ in range(2):   # the probe includes thread start-up: re-aim once or twice if the sample came out short
            if secs >= 0.4 * budget_s or nq >= self.n_src:
                break
            nq = int(min(self.n_src, nq * 0.8 * budget_s / max(secs, 1e-6)))
            secs = self._run(nq)
        sample = "%d source queries vs all %d target rows%s (D=%d, k=%d), oracle port of the reference matcher, OpenMP" % (
            nq, self.n_tgt, (" + %d target queries vs all %d source rows" % (nq, self.n_src)) if self.both else "",
            self.dim, self.k)
        return nq / secs, self.cores, sample, secs


def run_reference(args, wl):
    desc, n_src, n_tgt, k, mode, cfg = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(desc, n_src, n_tgt, k, mode)
    rates, secs_all = [], []
    sample, cores = "", ref.cores
    for i in range(args.warmup + args.steps):
        r, cores, sample, secs = ref.rate(6.0)
        if i >= args.warmup:
            rates.append(r)
            secs_all.append(secs)
    value = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs_all)),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg, "timing": "wall clock around the CPU matcher call, bounded query sample per step"},
            "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    from lidar_global_registration_b200 import build as b200_build
    from lidar_global_registration_b200 import device as D
    from lidar_global_registration_b200 import matcher as M
    from lidar_global_registration_b200 import synth

    desc, n_src, n_tgt, k, mode_name, cfg = WORKLOADS[wl]
    tsharded = mode_name == "knn_target_sharded"
    mode = {"mutual": M.MODE_MUTUAL, "ratio": M.MODE_RATIO, "one_sided": M.MODE_ONE_SIDED,
            "knn_target_sharded": M.MODE_KNN_ONLY}[mode_name]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 matcher has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        if "B200M_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["B200M_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)   # keep NCCL's version banner off stdout (rank 0 prints ONE JSON line)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        b200_build.build()
    if world > 1:
        dist.barrier()
    be = D.GpuBackend(local_rank)
    sm = D.ShardedMatcher(be, rank, world, group)

    # target-sharded: every rank generates its own shard of n_tgt rows (seeded by rank), global row offset rank * n_tgt
    src, tgt, dim = synth.make_pair_torch(desc, n_src, n_tgt, dev)
    if tsharded and rank > 0:   # a different shard per rank: the common recipe, rows rotated and slightly rescaled
        tgt = (torch.roll(tgt, shifts=1237 * rank, dims=0) * (1.0 + 2e-3 * rank)).contiguous()
    stride_b = src.stride(0) * 4
    t_off = rank * n_tgt if tsharded else 0
    torch.cuda.synchronize()

    def step_device():
        be.upload_device(0, src, dim)
        be.upload_device(1, tgt, dim, index_offset=t_off)
        if tsharded:
            idx, dst, cnt = sm.knn_target_sharded(k)
            return idx, cnt.sum(), dst
        return sm.match_query_sharded(k, mode)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        barrier()
        if sampler:
            sampler.begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    # ---- device-resident value (per-kernel CUDA events are recorded on the stream without host syncs) ----
    be.ctx.set_profiling(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step_device()
    be.ctx.reset_stats()
    ms_total, out = timed(step_device, args.steps, sampler)
    clocks = sampler.stop() if sampler else None
    st = be.ctx.stats()
    be.ctx.set_profiling(False)
    launches = st["launches"]
    n_corr = int(out[1].item())
    value = n_src * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host-facing C-ABI ----
    src_h = src.cpu().pin_memory()
    tgt_h = tgt.cpu().pin_memory()
    kk = k if mode == M.MODE_MUTUAL else 1
    q0, q1 = D.shard_bounds(n_src, rank, world)
    out_h = torch.empty((max((q1 - q0) * kk, 1), 4), dtype=torch.int32).pin_memory()
    lists_h = None
    if tsharded:
        lists_h = [torch.empty((n_src, k), dtype=torch.int32).pin_memory(), torch.empty((n_src, k), dtype=torch.float32).pin_memory(),
                   torch.empty((n_src,), dtype=torch.int32).pin_memory()]
    d2h = [0]

    h2d_rank = [0]

    def step_e2e():
        if tsharded:   # queries replicated (slices over PCIe, all-gather over NVLink), this rank's own target shard
            h2d_rank[0] = sm.upload_host_sharded(0, src_h, dim) + be.upload_host(1, tgt_h, dim, index_offset=t_off)
        else:          # both sets replicated
            h2d_rank[0] = sm.upload_host_sharded(0, src_h, dim) + sm.upload_host_sharded(1, tgt_h, dim)
        if tsharded:
            idx, dst, cnt = sm.knn_target_sharded(k)
            lists_h[0].copy_(idx, non_blocking=True)
            lists_h[1].copy_(dst, non_blocking=True)
            lists_h[2].copy_(cnt, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            d2h[0] = (lists_h[0].numel() + lists_h[1].numel() + lists_h[2].numel()) * 4 if rank == 0 else 0
            return idx, cnt, None
        rec, n_out, _ = sm.match_query_sharded(k, mode)
        n = int(n_out.item())                       # 8-byte D2H + sync: the count the caller needs
        out_h[:n].copy_(rec[:n], non_blocking=True)
        d2h[0] = n * 16 + 8
        return rec, n_out, None

    for _ in range(2):
        step_e2e()
    e2e_steps = max(2, min(args.steps, 5))
    ms_e2e, _ = timed(step_e2e, e2e_steps)
    e2e_value = n_src * e2e_steps / (ms_e2e * 1e-3)
    d2h_t = torch.tensor([d2h[0], h2d_rank[0]], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(d2h_t)
    h2d = int(d2h_t[1].item())   # bytes all ranks together copied host -> device per step

    # ---- roofline of the dominant kernel (tcgen05 candidate pass): CUDA events around every launch of the
    #      timed region above, on the stream the kernel is launched on ----
    both = mode in (M.MODE_MUTUAL, M.MODE_RATIO_MUTUAL)
    t0, t1 = D.shard_bounds(n_tgt, rank, world)
    flops_per_step = 2.0 * dim * ((q1 - q0) * n_tgt + ((t1 - t0) * n_src if both else 0))
    if tsharded:   # every rank: all queries against its own target shard
        flops_per_step = 2.0 * dim * n_src * n_tgt
    # FLOPs the candidate kernel actually had to do: the masked reverse pass answers only the target rows that forward
    # lists name (st["pairs_scored"] = sum over the launches of rows searched x train rows)
    if st.get("pairs_scored", 0) > 0:
        flops_per_step = 2.0 * dim * st["pairs_scored"] / args.steps
    cand_ms_per_step = st["ms_candidates"] / args.steps
    pk = peaks()
    achieved = flops_per_step / (cand_ms_per_step * 1e-3) / 1e12 if cand_ms_per_step > 0 else 0.0
    n_launch = st["candidate_launches"] / args.steps
    roofline = {"bound": "tensor", "kernel": "tc_candidates_kernel", "achieved": achieved, "peak": pk["bf16_sustained"],
                "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH.get(wl) if world == 1 else None,
                "traffic_source": "bytes per launch, %s (ncu --set full, 1 GPU)" % NCU_TRAFFIC_SOURCE.get(wl, "-")
                                  if world == 1 and wl in NCU_TRAFFIC_BYTES_PER_LAUNCH else None,
                "peak_source": pk["source"] + ", bf16_tflops_sustained (kernel timed inside a long step)",
                "launch_ms": cand_ms_per_step / max(n_launch, 1), "launches_per_step": n_launch,
                "algorithmic_flops_per_step": flops_per_step,
                "breakdown_ms_per_step": {x: st[x] / args.steps for x in ("ms_pack", "ms_prepare", "ms_candidates", "ms_rerank",
                                                                     "ms_fallback", "ms_filter")},
                "candidates_per_row": st["candidates"] / max(st.get("rows_answered") or st["rows_total"], 1),
                "query_rows_searched_frac": (st.get("rows_answered") or st["rows_total"]) / max(st["rows_total"], 1),
                "rows_overflowed_frac": st["rows_flagged"] / max(st["rows_total"], 1)}

    # Second yardstick (for short descriptors the limiter is the accumulator drain + select epilogue, not the
    # tensor pipe): every (query, train) pair's accumulator has to pass the min pipe once.  Measured on this part
    # (tools/ubench/min_ubench.cu): one three-input FMNMX3 (two new values per lane) per 2.25 cycles per scheduler.
    pairs_per_step = flops_per_step / (2.0 * dim)
    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    select_peak = 148 * 4 * (64.0 / 2.25) * sm_clock * 1e6
    select_ach = pairs_per_step / (cand_ms_per_step * 1e-3) if cand_ms_per_step > 0 else 0.0
    roofline["select_epilogue"] = {"achieved": select_ach, "peak": select_peak, "unit": "pairs/s", "frac": select_ach / select_peak,
                                   "peak_source": "148 SMs x 4 schedulers x 64 values per 2.25 cycles (measured FMNMX3 rate) x SM clock under load"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r, cores, sample, secs = CpuReference(desc, n_src, n_tgt, k, "one_sided" if tsharded else mode_name).rate(12.0)
        cpu = {"value": r, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample, "seconds": secs}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak" if tsharded else "strong", "vs_baseline": None,
                "dtype": "f32 (FP16 tensor-core candidates, exact FP32 re-rank)",
                "data": "synthetic",
                "config": {"workload": cfg, "descriptor": desc, "dim": dim, "n_src": n_src, "n_tgt": n_tgt, "k": k,
                           "filter": mode_name, "row_stride_bytes": stride_b, "correspondences": n_corr,
                           "sharding": ("target-sharded (n_tgt rows per rank, %d in total), queries replicated, NCCL all-gather + merge"
                                        % (n_tgt * world)) if tsharded else
                                       ("query-sharded, target replicated" if world > 1 else "single GPU"),
                           "l2": "inputs larger than L2 (no flush needed)" if n_tgt * dim * 2 > 126e6 else
                                 "inputs smaller than L2; the step rewrites >126 MB of operands/candidates between kNN passes"},
                "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h_t[0].item()), "ms_per_step": ms_e2e / e2e_steps,
                        "replication": "each rank copies 1/N of a replicated set over PCIe, NCCL all-gather over NVLink"
                                       if world > 1 else "single GPU"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    be.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, args.workload)
    else:
        run_b200(args, args.workload)


if __name__ == "__main__":
    main()
