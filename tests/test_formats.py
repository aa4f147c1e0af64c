"""The reference's text formats (saveFeatures CSV, correspondences CSV): known-answer strings for the float
formatting (operator<<(float) == "%g") and round trips at the precision the reference's own dump keeps."""
import numpy as np

from lidar_global_registration_b200 import formats as F
from lidar_global_registration_b200 import matcher as M
from lidar_global_registration_b200 import synth


def test_float_formatting_is_ostream_default():
    # std::cout << 0.1f, 1e-7f, 123456789.f, 100.f, 1.5f, 0.f, 3.4028235e38f
    assert [F._g(v) for v in (0.1, 1e-7, 123456789.0, 100.0, 1.5, 0.0, 3.4028235e38)] == \
        ["0.1", "1e-07", "1.23457e+08", "100", "1.5", "0", "3.40282e+38"]


def test_features_csv_known_answer_and_round_trip(tmp_path):
    p = tmp_path / "histograms_src.csv"
    feats = np.array([[0.5, 12.25, 0.0], [100.0, 1e-7, 33.333332]], np.float32)
    F.write_features_csv(p, feats, indices=[7, 42])
    assert open(p).read() == "7,0.5,12.25,0\n42,100,1e-07,33.3333\n"     # no header (include/feature_analysis.h:18-25)
    idx, back = F.read_features_csv(p)
    assert idx.tolist() == [7, 42] and back.dtype == np.float32
    np.testing.assert_allclose(back, feats, rtol=1e-5)
    # an AoS descriptor set: only the descriptor columns are dumped; 6 significant digits survive
    src, _, dim = synth.make_pair("fpfh", 50, 10, nan_frac=0.0)
    F.write_features_csv(p, src, dim=dim)
    idx, back = F.read_features_csv(p)
    assert back.shape == (50, dim) and idx.tolist() == list(range(50))
    np.testing.assert_allclose(back, src[:, :dim], rtol=5e-6)


def test_correspondences_csv_known_answer_and_round_trip(tmp_path):
    p = tmp_path / "correspondences.csv"
    corrs = np.array([(0, 2, 0.125, 0.5), (3, 1, 2.5, 3.4028235e38)], M.CORR_DTYPE)
    sxyz = np.arange(12, dtype=np.float32).reshape(4, 3) / 4
    txyz = -np.arange(9, dtype=np.float32).reshape(3, 3)
    F.write_correspondences_csv(p, corrs, sxyz, txyz)
    lines = open(p).read().splitlines()
    assert lines[0] == "query_idx,match_idx,distance,threshold,x_s,y_s,z_s,x_t,y_t,z_t"      # src/common.cpp:1252
    assert lines[1] == "0,2,0.125,0.5,0,0.25,0.5,-6,-7,-8"
    assert lines[2] == "3,1,2.5,3.40282e+38,2.25,2.5,2.75,-3,-4,-5"
    back = F.read_correspondences_csv(p)
    assert back.dtype == M.CORR_DTYPE and back["index_query"].tolist() == [0, 3] and back["index_match"].tolist() == [2, 1]
    assert back["distance"].tolist() == [0.125, 2.5]
    np.testing.assert_allclose(back["threshold"], corrs["threshold"], rtol=1e-5)
    assert F.read_correspondences_csv(tmp_path / "c2.csv" if False else p).shape == (2,)
