"""Device-resident and multi-GPU drivers of the matcher on torch tensors (HBM buffers, the current torch stream).

    GpuBackend      one b200m context working on torch device buffers (no host copies)
    ShardedMatcher  SURVEY 8e, one process per GPU: query-sharded / target-replicated matcher and the target-sharded kNN.
                    With a GpuBackend this is a THIN caller of the library -- partitioning, the NCCL exchange step and
                    the merge live in libb200match.so (csrc/multi.cu: b200m_match_sharded_device,
                    b200m_knn_target_sharded_device, b200m_upload_replicated); torch.distributed only carries the
                    128-byte communicator id to the ranks.  The same partitioning written out in Python over
                    torch.distributed collectives is kept for backends without native sharding: the CPU tests drive it
                    with gloo and a stand-in backend as an executable statement of the exchange protocol.
"""
import torch

from . import matcher as M


# below this many (query, train) pairs the row selection's host round trip costs more than the skipped rows save
MASKED_REVERSE_MIN_PAIRS = 10 ** 9


def shard_bounds(n, rank, world):
    """The library's row partition (b200m_shard_rows): rank r owns [r*R, min(n, (r+1)*R)), R = ceil(n / world) -- the
    all-gathered slots are then the row-major table of all rows."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


class GpuBackend:
    """kNN / filter / merge on device buffers through the C-ABI *_device entry points."""
    native_sharding = True    # partitioning + NCCL inside libb200match.so

    def __init__(self, device_index=0, precision=M.PREC_TC_F16, cand_cap=0):
        self.device = torch.device("cuda", device_index)
        torch.cuda.set_device(self.device)
        self.ctx = M.Context(device_index)
        self.precision = precision
        self.cand_cap = cand_cap
        self.use_torch_stream()

    def use_torch_stream(self):
        """Queue the library's work on torch's current stream.  torch reports the legacy default stream as
        handle 0, which the C-ABI reads as "use the context's own stream": pass cudaStreamLegacy (0x1) instead,
        otherwise torch's ops (NCCL waits, .item(), events) would not be ordered with the library's kernels."""
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream or 1)

    def close(self):
        self.ctx.close()

    @property
    def n(self):
        return self.ctx.n

    def attach_comm(self, rank, world, group=None):
        """Give the context its rank in an NCCL communicator of its own: rank 0 makes the id, torch.distributed carries
        the 128 bytes to the others (plumbing only)."""
        import torch.distributed as dist
        ident = torch.zeros(M.UNIQUE_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            ident = torch.frombuffer(bytearray(M.comm_unique_id()), dtype=torch.uint8).clone()
        on_dev = dist.get_backend(group) == "nccl"
        t = ident.to(self.device) if on_dev else ident
        dist.broadcast(t, src=0, group=group)
        self.ctx.comm_attach(world, rank, bytes(t.cpu().numpy().tobytes()))

    # -- descriptors ----------------------------------------------------------
    def upload_device(self, side, aos, dim, index_offset=0):
        """aos: float32 CUDA tensor [n, stride_floats] (AoS point structs resident in HBM)."""
        assert aos.is_cuda and aos.dtype == torch.float32 and aos.dim() == 2 and aos.stride(1) == 1
        self.ctx.upload_device(side, aos.data_ptr(), aos.shape[0], aos.stride(0) * 4, dim, index_offset)

    def upload_host(self, side, aos_host, dim, index_offset=0):
        """aos_host: float32 CPU tensor (pinned for a true async copy) [n, stride_floats]."""
        assert not aos_host.is_cuda and aos_host.dtype == torch.float32 and aos_host.stride(1) == 1
        self.ctx._ck(self.ctx._L.b200m_upload(self.ctx._h, side, aos_host.data_ptr(), aos_host.shape[0],
                                              aos_host.stride(0) * 4, dim, index_offset))
        self.ctx.n[side] = aos_host.shape[0]
        self.ctx.dim = dim
        return aos_host.shape[0] * aos_host.stride(0) * 4

    # -- kernels ----------------------------------------------------------------
    def knn(self, k, direction, row_begin, row_end):
        rows = row_end - row_begin
        idx = torch.empty((rows, k), dtype=torch.int32, device=self.device)
        dist = torch.empty((rows, k), dtype=torch.float32, device=self.device)
        cnt = torch.empty((rows,), dtype=torch.int32, device=self.device)
        if rows:
            self.ctx.knn_device(k, direction, row_begin, row_end, idx.data_ptr(), dist.data_ptr(), cnt.data_ptr(),
                                self.precision, self.cand_cap)
        return idx, dist, cnt

    def knn_masked(self, k, direction, row_begin, row_end, flags):
        """kNN of the rows of [row_begin, row_end) whose flag is set (uint8 CUDA tensor over ALL rows of the query side);
        the other rows get empty lists."""
        rows = row_end - row_begin
        idx = torch.empty((rows, k), dtype=torch.int32, device=self.device)
        dist = torch.empty((rows, k), dtype=torch.float32, device=self.device)
        cnt = torch.empty((rows,), dtype=torch.int32, device=self.device)
        if rows:
            self.ctx.knn_masked_device(k, direction, row_begin, row_end, flags.data_ptr(), idx.data_ptr(), dist.data_ptr(),
                                       cnt.data_ptr(), self.precision, self.cand_cap)
        return idx, dist, cnt

    def referenced_rows(self, k, fwd, n_flags, index_offset=0):
        """uint8 flags [n_flags]: 1 for every row of the other side that the k-lists `fwd` name."""
        flags = torch.zeros((n_flags,), dtype=torch.uint8, device=self.device)
        if fwd[0].shape[0]:
            self.ctx.mark_referenced_device(k, fwd[0].data_ptr(), fwd[2].data_ptr(), fwd[0].shape[0], index_offset,
                                            flags.data_ptr(), n_flags)
        return flags

    def filter(self, k, mode, row_begin, row_end, fwd, rev, n_rev_rows, ratio_thr=M.MATCHING_RATIO_THRESHOLD,
               distance_thr=M.FLT_MAX, thr_src=None, thr_tgt=None, want_avg=False):
        """-> (records int32 [cap, 4] (reinterpret as CORR_DTYPE), n_out uint64-as-int64 [1], avg float32 [1] or None)."""
        rows = row_end - row_begin
        kk = k if mode == M.MODE_MUTUAL else 1
        cap = max(rows * kk, 1)
        out = torch.empty((cap, 4), dtype=torch.int32, device=self.device)
        n_out = torch.zeros((1,), dtype=torch.int64, device=self.device)
        avg = torch.zeros((1,), dtype=torch.float32, device=self.device) if want_avg else None
        r = rev if rev is not None else (None, None, None)
        ptr = lambda t: 0 if t is None else t.data_ptr()
        self.ctx.filter_device(k, mode, row_begin, row_end, ptr(fwd[0]), ptr(fwd[1]), ptr(fwd[2]), ptr(r[0]), ptr(r[1]),
                               ptr(r[2]), n_rev_rows, out.data_ptr(), cap, n_out.data_ptr(), ptr(avg), ptr(thr_src),
                               ptr(thr_tgt), ratio_thr, distance_thr)
        return out, n_out, avg

    def merge(self, k, idx_in, dist_in, cnt_in):
        """idx_in/dist_in [n_lists, nq, k], cnt_in [n_lists, nq] -> the k best per query by (dist, idx)."""
        n_lists, nq = cnt_in.shape
        idx = torch.empty((nq, k), dtype=torch.int32, device=self.device)
        dist = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        cnt = torch.empty((nq,), dtype=torch.int32, device=self.device)
        self.ctx.merge_device(k, n_lists, nq, idx_in.data_ptr(), dist_in.data_ptr(), cnt_in.data_ptr(), idx.data_ptr(),
                              dist.data_ptr(), cnt.data_ptr())
        return idx, dist, cnt

    def match_device(self, k, mode, ratio_thr=M.MATCHING_RATIO_THRESHOLD, distance_thr=M.FLT_MAX, want_avg=False):
        """Whole matcher call on one GPU, everything resident in HBM."""
        nq, nt = self.n
        fwd = self.knn(k, 0, 0, nq)
        rev = None
        if mode in (M.MODE_MUTUAL, M.MODE_RATIO_MUTUAL):
            if nq * nt >= MASKED_REVERSE_MIN_PAIRS:
                rev = self.knn_masked(k, 1, 0, nt, self.referenced_rows(k, fwd, nt))
            else:
                rev = self.knn(k, 1, 0, nt)
        return self.filter(k, mode, 0, nq, fwd, rev, nt if rev is not None else 0, ratio_thr, distance_thr,
                           want_avg=want_avg)


def records_to_numpy(records, n_out):
    """int32 [cap,4] device tensor + count -> CORR_DTYPE host array."""
    n = int(n_out.item())
    return records[:n].contiguous().cpu().numpy().view(M.CORR_DTYPE).reshape(-1)


class ShardedMatcher:
    """One process per GPU.  `dist` is torch.distributed (or None for a single process)."""

    def __init__(self, backend, rank=0, world=1, group=None):
        self.b = backend
        self.rank, self.world, self.group = rank, world, group
        self.native = world > 1 and getattr(backend, "native_sharding", False)
        if self.native:
            backend.attach_comm(rank, world, group)

    # -- collectives ------------------------------------------------------------
    def _all_gather_rows(self, t, n_total):
        """All-gather row-sharded tensors (shard_bounds layout) into the full [n_total, ...] tensor."""
        if self.world == 1:
            return t
        import torch.distributed as dist
        max_rows = (n_total + self.world - 1) // self.world
        pad = torch.zeros((max_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        out = torch.empty((self.world * max_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        if n_total % self.world == 0:
            return out
        parts = []
        for r in range(self.world):
            lo, hi = shard_bounds(n_total, r, self.world)
            parts.append(out[r * max_rows:r * max_rows + (hi - lo)])
        return torch.cat(parts, 0)

    # -- descriptor replication over NVLink instead of PCIe ---------------------------
    def upload_host_sharded(self, side, aos_host, dim, index_offset=0):
        """Replicate a HOST descriptor set on every rank: rank r copies only its 1/world slice over PCIe and the
        slices are all-gathered over NVLink (8 ranks pulling the whole set through the host's memory system at
        once is what limits the end-to-end rate of a replicated run).  aos_host: pinned float32 [n, stride_floats]."""
        if self.world == 1:
            return self.b.upload_host(side, aos_host, dim, index_offset)
        n = aos_host.shape[0]
        if self.native:
            assert index_offset == 0
            c = self.b.ctx
            c._ck(c._L.b200m_upload_replicated(c._h, side, aos_host.data_ptr(), n, aos_host.stride(0) * 4, dim))
            c.n[side], c.dim = n, dim
            lo, hi = shard_bounds(n, self.rank, self.world)
            return (hi - lo) * aos_host.stride(0) * 4
        lo, hi = shard_bounds(n, self.rank, self.world)
        shard = aos_host[lo:hi].to(self.b.device, non_blocking=True)
        full = self._all_gather_rows(shard, n)
        self.b.upload_device(side, full, dim, index_offset)     # the pack kernel reads `full` on this same stream
        return (hi - lo) * aos_host.stride(0) * 4               # bytes this rank moved host -> device

    # -- query-sharded, target replicated -----------------
    def match_query_sharded(self, k, mode, ratio_thr=M.MATCHING_RATIO_THRESHOLD, distance_thr=M.FLT_MAX):
        """Both descriptor sets are resident on every rank.  Returns this rank's slice of the
        correspondence list (records, n_out): concatenating the slices in rank order gives the
        single-GPU output (ascending index_query)."""
        nq, nt = self.b.n
        q0, q1 = shard_bounds(nq, self.rank, self.world)
        if self.native:
            cap = max((q1 - q0) * (k if mode == M.MODE_MUTUAL else 1), 1)
            out = torch.empty((cap, 4), dtype=torch.int32, device=self.b.device)
            n_out = torch.zeros((1,), dtype=torch.int64, device=self.b.device)
            self.b.ctx.match_sharded_device(k, mode, out.data_ptr(), cap, n_out.data_ptr(), ratio_thr=ratio_thr,
                                            distance_thr=distance_thr, precision=self.b.precision, cand_cap=self.b.cand_cap)
            return out, n_out, None
        fwd = self.b.knn(k, 0, q0, q1)
        rev = None
        if mode in (M.MODE_MUTUAL, M.MODE_RATIO_MUTUAL):
            t0, t1 = shard_bounds(nt, self.rank, self.world)
            if nq * nt >= MASKED_REVERSE_MIN_PAIRS:
                # the mutual test reads rev[j] only for targets j that a forward list names: every rank marks the ones its
                # own source rows name, the flags are max-reduced (nt bytes), and the reverse pass skips the rest
                flags = self.b.referenced_rows(k, fwd, nt)
                if self.world > 1:
                    import torch.distributed as dist
                    dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=self.group)
                r = self.b.knn_masked(k, 1, t0, t1, flags)
            else:
                r = self.b.knn(k, 1, t0, t1)
            rev = tuple(self._all_gather_rows(x, nt) for x in r)     # the path's one exchange step
        return self.b.filter(k, mode, q0, q1, fwd, rev, nt if rev is not None else 0, ratio_thr, distance_thr)

    def gather_records(self, records, n_out):
        """Concatenate every rank's correspondence slice in rank order on all ranks -> (records, n)."""
        if self.world == 1:
            return records[:int(n_out.item())], int(n_out.item())
        import torch.distributed as dist
        counts = torch.empty((self.world,), dtype=torch.int64, device=records.device)
        dist.all_gather_into_tensor(counts, n_out.to(torch.int64), group=self.group)
        counts_h = counts.cpu().tolist()
        mx = max(max(counts_h), 1)
        pad = torch.zeros((mx, 4), dtype=records.dtype, device=records.device)
        pad[:counts_h[self.rank]] = records[:counts_h[self.rank]]
        out = torch.empty((self.world * mx, 4), dtype=records.dtype, device=records.device)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        parts = [out[r * mx:r * mx + counts_h[r]] for r in range(self.world)]
        return torch.cat(parts, 0), sum(counts_h)

    # -- target-sharded (config C5): each rank holds a slice of the target set ---------
    def knn_target_sharded(self, k):
        """The backend's side 1 holds THIS rank's target shard (uploaded with index_offset = shard start), side 0
        all queries.  Exact local top-k with global indices -> all-gather -> merge kernel (canonical tie rule)."""
        nq = self.b.n[0]
        if self.native:
            idx = torch.empty((nq, k), dtype=torch.int32, device=self.b.device)
            dist_ = torch.empty((nq, k), dtype=torch.float32, device=self.b.device)
            cnt = torch.empty((nq,), dtype=torch.int32, device=self.b.device)
            self.b.ctx.knn_target_sharded_device(k, idx.data_ptr(), dist_.data_ptr(), cnt.data_ptr(), self.b.precision,
                                                 self.b.cand_cap)
            return idx, dist_, cnt
        local = self.b.knn(k, 0, 0, nq)
        if self.world == 1:
            return local
        import torch.distributed as dist
        gathered = []
        for x in local:
            out = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            dist.all_gather_into_tensor(out, x.contiguous(), group=self.group)
            gathered.append(out.view((self.world,) + tuple(x.shape)))
        return self.b.merge(k, gathered[0], gathered[1], gathered[2])
