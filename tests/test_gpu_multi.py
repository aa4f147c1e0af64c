"""Single-process multi-GPU (b200m_create_multi: one context and one host thread per device INSIDE the library, NCCL over
NVLink) == the oracle.  Needs at least two GPUs (gpurun --gpus 2); skipped on a one-GPU box."""
import numpy as np
import pytest

from lidar_global_registration_b200 import build as b200_build
from lidar_global_registration_b200 import matcher as M
from lidar_global_registration_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    return list(range(min(torch.cuda.device_count(), 8))) if torch.cuda.is_available() else []


@pytest.fixture(scope="module", autouse=True)
def _built():
    b200_build.build()
    M.load_library()
    if len(_devices()) < 2:
        pytest.skip("needs at least two GPUs")


@pytest.mark.parametrize("masked", [False, True], ids=["full-reverse", "masked-reverse"])
@pytest.mark.parametrize("desc,nq,nt,k", [("fpfh", 5003, 7001, 2), ("shot", 1111, 1405, 2), ("rops", 901, 1200, 5), ("fpfh", 3, 5, 1)])
def test_group_match_and_knn_equal_oracle(monkeypatch, desc, nq, nt, k, masked):
    monkeypatch.setenv("B200M_MASKED_MIN_PAIRS", "1" if masked else "1e30")
    src, tgt, dim = synth.make_pair(desc, nq, nt, nan_frac=0.01 if nq > 10 else 0.0)
    sd, td = np.ascontiguousarray(src[:, :dim]), np.ascontiguousarray(tgt[:, :dim])
    rng = np.random.default_rng(2)
    thr_s, thr_t = rng.random(nq).astype(np.float32), rng.random(nt).astype(np.float32)
    dthr = np.float32(0.5)
    with M.Group(_devices()) as g:
        g.upload(0, src, dim)
        g.upload(1, tgt, dim)
        for mode, name in ((M.MODE_MUTUAL, "mutual"), (M.MODE_ONE_SIDED, "one_sided"), (M.MODE_RATIO, "ratio")):
            if name == "ratio" and k < 2:
                continue
            got, avg = g.match(k, mode, distance_thr=float(dthr), thr_src=thr_s, thr_tgt=thr_t)
            exp, eavg = orc.match(sd, td, k, name, 1.1, dthr, thr_s, thr_t)
            assert got.tobytes() == exp.tobytes() and avg == eavg, name
        e = orc.knn(sd, td, k)
        for a, b in zip(g.knn(k, M.SHARD_QUERY), e):
            assert np.array_equal(a, b)
        g.upload(1, tgt, dim, sharded=True)      # target rows split over the devices, global index offsets
        for a, b in zip(g.knn(k, M.SHARD_TARGET), e):
            assert np.array_equal(a, b)
