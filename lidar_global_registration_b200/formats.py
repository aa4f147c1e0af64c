"""The reference's descriptor / correspondence text formats, so that descriptor dumps produced on
a machine that has PCL can be replayed through the GPU matcher and its output fed back into the reference's `metric` /
`debug` commands.

  saveFeatures<FeatureT>          include/feature_analysis.h:12-27   "histograms[scale]_{src,tgt}.csv": NO header, one
                                  line per descriptor: `index,f0,...,f{m-1}`; floats through operator<<(float), i.e. "%g"
                                  with 6 significant digits (the dump is lossy by design of the reference).
  saveCorrespondencesToCSV        src/common.cpp:1247-1265           header `query_idx,match_idx,distance,threshold,x_s,y_s,
                                  z_s,x_t,y_t,z_t`, then one line per correspondence, same float formatting.
  readCorrespondencesFromCSV      src/common.cpp:1223-1245           skips the header, reads the first four tokens of every
                                  line: stoi, stoi, stof, stof.
"""
import numpy as np

from .matcher import CORR_DTYPE

CORR_HEADER = "query_idx,match_idx,distance,threshold,x_s,y_s,z_s,x_t,y_t,z_t"


def _g(x):
    """std::ostream << float with default flags: printf("%g") at precision 6."""
    return "%g" % float(np.float32(x))


def write_features_csv(path, features, indices=None, dim=None):
    """saveFeatures: `features` [n, >= dim] float32 (AoS rows allowed: only the first `dim` columns are written)."""
    f = np.asarray(features, np.float32)
    dim = dim or f.shape[1]
    with open(path, "w") as out:
        for i in range(f.shape[0]):
            out.write(str(int(indices[i]) if indices is not None else i))
            out.write("".join("," + _g(v) for v in f[i, :dim]))
            out.write("\n")


def read_features_csv(path):
    """-> (indices int32 [n], features float32 [n, m]) of a saveFeatures dump (values as std::stof would read them)."""
    idx, rows = [], []
    with open(path) as fin:
        for line in fin:
            line = line.strip()
            if not line:
                continue
            tok = line.split(",")
            idx.append(int(tok[0]))
            rows.append([np.float32(t) for t in tok[1:]])
    if not rows:
        return np.zeros(0, np.int32), np.zeros((0, 0), np.float32)
    m = len(rows[0])
    if any(len(r) != m for r in rows):
        raise ValueError("%s: rows of different length" % path)
    return np.asarray(idx, np.int32), np.asarray(rows, np.float32)


def write_correspondences_csv(path, correspondences, src_xyz, tgt_xyz):
    """saveCorrespondencesToCSV: `correspondences` is a CORR_DTYPE array (cloud-global indices, i.e. after finalize);
    src_xyz / tgt_xyz are the clouds' point coordinates [n, >= 3]."""
    s, t = np.asarray(src_xyz, np.float32), np.asarray(tgt_xyz, np.float32)
    with open(path, "w") as out:
        out.write(CORR_HEADER + "\n")
        for c in correspondences:
            i, j = int(c["index_query"]), int(c["index_match"])
            vals = [_g(c["distance"]), _g(c["threshold"])] + [_g(v) for v in s[i, :3]] + [_g(v) for v in t[j, :3]]
            out.write("%d,%d,%s\n" % (i, j, ",".join(vals)))


def read_correspondences_csv(path):
    """readCorrespondencesFromCSV: header skipped, tokens 0..3 of every line -> CORR_DTYPE array."""
    recs = []
    with open(path) as fin:
        fin.readline()
        for line in fin:
            tok = line.rstrip("\n").split(",")
            if len(tok) < 4:
                continue
            recs.append((int(tok[0]), int(tok[1]), np.float32(tok[2]), np.float32(tok[3])))
    return np.asarray(recs, CORR_DTYPE) if recs else np.zeros(0, CORR_DTYPE)
