"""Short, fixed program for ncu: N matcher calls of one BASELINE workload on device-resident inputs.
    python tools/profile_target.py c3 [calls]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from lidar_global_registration_b200 import device as D  # noqa: E402
from lidar_global_registration_b200 import matcher as M  # noqa: E402
from lidar_global_registration_b200 import synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
desc, n_src, n_tgt, k, mode_name, _ = WORKLOADS[wl]
mode = {"mutual": M.MODE_MUTUAL, "ratio": M.MODE_RATIO, "one_sided": M.MODE_ONE_SIDED}[mode_name]
be = D.GpuBackend(0)
src, tgt, dim = synth.make_pair_torch(desc, n_src, n_tgt, be.device)
for _ in range(calls):
    be.upload_device(0, src, dim)
    be.upload_device(1, tgt, dim)
    rec, n_out, _ = be.match_device(k, mode)
torch.cuda.synchronize()
print(wl, "correspondences", int(n_out.item()))
be.close()
