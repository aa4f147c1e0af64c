"""Generate the committed golden fixtures under tests/golden/ with OpenCV's
cv2.BFMatcher(NORM_L2).knnMatch -- the third-party routine the reference's matchBF
calls (include/matching.h:600,612).  cv2 exists in the build container and cannot be
assumed on the GPU box, so its outputs are frozen here.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

The blocking loop below restates matchBF's own (include/matching.h:604-632) around
the real cv2 call, including the merge through updateMultivaluedCorrespondence
(src/common.cpp:517-529), so `bf_*` arrays are what the reference's matchBF returns
for these inputs (with cv2 4.13 standing in for the CI's OpenCV 4.5.1).
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from lidar_global_registration_b200 import synth  # noqa: E402


def update_multivalued(idx, dist, k, match_idx, distance):
    # src/common.cpp:517-529
    pos = 0
    while pos != len(idx) and dist[pos] < distance:
        pos += 1
    idx.insert(pos, match_idx)
    dist.insert(pos, distance)
    if len(idx) > k:
        del idx[k:]
        del dist[k:]


def match_bf_cv2(query, train, k, block_size):
    matcher = cv2.BFMatcher(cv2.NORM_L2)
    nq, nt = query.shape[0], train.shape[0]
    res_i = [[] for _ in range(nq)]
    res_d = [[] for _ in range(nq)]
    for qb in range(0, nq, block_size):
        for tb in range(0, nt, block_size):
            q = np.ascontiguousarray(query[qb:qb + block_size])
            t = np.ascontiguousarray(train[tb:tb + block_size])
            matches = matcher.knnMatch(q, t, k)
            for ms in matches:
                if len(ms) == 0 or ms[0].queryIdx == -1:
                    continue
                qi = qb + ms[0].queryIdx
                for m in ms:
                    update_multivalued(res_i[qi], res_d[qi], k, tb + m.trainIdx, np.float32(m.distance))
    idx = np.full((nq, k), -1, np.int32)
    dist = np.zeros((nq, k), np.float32)
    cnt = np.zeros(nq, np.int32)
    for i in range(nq):
        n = len(res_i[i])
        cnt[i] = n
        idx[i, :n] = res_i[i]
        dist[i, :n] = res_d[i]
    return idx, dist, cnt


CASES = [
    # name, descriptor, n_src, n_tgt, k, block
    ("fpfh_k1", "fpfh", 600, 700, 1, 10000),
    ("fpfh_k2_blocked", "fpfh", 500, 650, 2, 200),
    ("fpfh_k5", "fpfh", 400, 450, 5, 10000),
    ("rops_k3", "rops", 300, 320, 3, 10000),
    ("shot_k2", "shot", 300, 400, 2, 10000),
    ("shot_k1_blocked", "shot", 256, 300, 1, 128),
]


def main():
    for name, desc, ns, nt, k, block in CASES:
        src, tgt, dim = synth.make_pair(desc, ns, nt, seed=synth.SEED + len(name), nan_frac=0.01)
        q = np.ascontiguousarray(src[:, :dim])
        t = np.ascontiguousarray(tgt[:, :dim])
        fi, fd, fc = match_bf_cv2(q, t, k, block)
        ri, rd, rc = match_bf_cv2(t, q, k, block)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), src=src, tgt=tgt, dim=dim, k=k, block=block,
                            bf_idx=fi, bf_dist=fd, bf_cnt=fc, bf_ridx=ri, bf_rdist=rd, bf_rcnt=rc,
                            cv2_version=cv2.__version__)
        print(name, src.shape, tgt.shape, "k", k, "nonempty", int((fc > 0).sum()))

    # exact ties inside one block: duplicated train rows -> lower train index first (cv2)
    rng = np.random.default_rng(7)
    t = rng.random((40, 33)).astype(np.float32)
    t[17] = t[5]
    t[30] = t[5]
    q = t[[5, 9]] .copy()
    fi, fd, fc = match_bf_cv2(q, t, 3, 10000)
    # k larger than the train set -> nt results
    fi2, fd2, fc2 = match_bf_cv2(q, t[:2], 3, 10000)
    np.savez_compressed(os.path.join(HERE, "ties_small.npz"), q=q, t=t, idx=fi, dist=fd, cnt=fc,
                        idx_kgt=fi2, dist_kgt=fd2, cnt_kgt=fc2)
    print("ties", fi.tolist(), fc.tolist(), "k>nt", fi2.tolist(), fc2.tolist())


if __name__ == "__main__":
    main()
