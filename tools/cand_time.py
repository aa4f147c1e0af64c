"""Candidate-kernel timing of one BASELINE workload, forward pass only, device-resident inputs (run on a B200):
    python tools/cand_time.py c2 [reps]
Prints the mean CUDA-event time of a tc_candidates_kernel launch and of the re-rank (b200m_stats, no host syncs inside).
A/B switches are environment variables read by b200m_create (B200M_TC_ALT, B200M_TC_SPLITN, B200M_TC_SPLITS, B200M_TC_DEBUG)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from lidar_global_registration_b200 import device as D  # noqa: E402
from lidar_global_registration_b200 import synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
desc, n_src, n_tgt, k, _, _ = WORKLOADS[wl]
be = D.GpuBackend(0)
src, tgt, dim = synth.make_pair_torch(desc, n_src, n_tgt, be.device)
be.upload_device(0, src, dim)
be.upload_device(1, tgt, dim)
for _ in range(2):
    be.knn(k, 0, 0, n_src)
torch.cuda.synchronize()
be.ctx.set_profiling(True)
be.ctx.reset_stats()
for _ in range(reps):
    be.knn(k, 0, 0, n_src)
torch.cuda.synchronize()
st = be.ctx.stats()
n = max(st["candidate_launches"], 1)
ms = st["ms_candidates"] / n
flops = 2.0 * dim * n_src * n_tgt
print("%s env{%s}: cand %.3f ms/launch = %.0f TFLOP/s algorithmic | rerank %.3f ms | fallback %.3f ms | cand/row %.1f | overflowed rows %d" % (
    wl, " ".join("%s=%s" % (k_, v) for k_, v in sorted(os.environ.items()) if k_.startswith("B200M_")), ms,
    flops / (ms * 1e-3) / 1e12, st["ms_rerank"] / n, st["ms_fallback"] / n, st["candidates"] / max(st["rows_total"], 1),
    st["rows_flagged"]))
be.close()
