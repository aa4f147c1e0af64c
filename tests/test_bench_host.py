"""Host-side pieces of bench.py that need no GPU: the clock sampler's time-window logic."""
import os
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _sampler_with(lines, t_begin):
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    import tempfile

    class _Done:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass
    s.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    s.f.write("\n".join(lines) + "\n")
    s.f.flush()
    s.p = _Done()
    s.t_begin = t_begin
    return s


def _line(t, sm, reasons=("Not Active",) * 4):
    import datetime
    ts = datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
    return "%s, %d, 1965, 500.0, %s" % (ts, sm, ", ".join(reasons))


def test_stamp_round_trip():
    t = 1792335361.123
    import datetime
    text = datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
    assert abs(bench.ClockSampler._stamp(text) - t) < 2e-3


def test_only_samples_inside_the_timed_region_count():
    now = time.time()
    lines = [_line(now - 2.0, 1000),                                                    # warm-up: outside
             _line(now - 0.5, 1400, ("Not Active", "Not Active", "Not Active", "Active")),
             _line(now - 0.3, 1500)]
    out = _sampler_with(lines, now - 1.0).stop()
    assert out["samples"] == 2 and out["window"] == "timed region"
    assert out["sm_mhz"] == 1450.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]


def test_short_region_falls_back_to_nearest_samples_and_says_so():
    now = time.time()
    lines = [_line(now - 3.0, 1000), _line(now - 2.0, 1100), _line(now - 1.5, 1200), _line(now - 1.0, 1300)]
    out = _sampler_with(lines, now - 0.01).stop()
    assert out["samples"] == 3 and out["window"].startswith("nearest samples")
    assert out["sm_mhz"] == 1200.0


def test_no_lines_is_reported():
    out = _sampler_with([], time.time()).stop()
    assert out["reasons"] == ["no samples"]


def test_both_arms_describe_the_workload_with_the_same_config():
    for wl in bench.WORKLOADS:
        assert wl in bench.METRICS
        for world in (1, 8):
            c = bench.workload_config(wl, world)
            assert c["workload"] == bench.WORKLOADS[wl][5] and c["dim"] in (33, 135, 352)
            assert set(c) == {"workload", "descriptor", "dim", "n_src", "n_tgt", "k", "filter", "row_stride_bytes", "sharding", "l2"}


def test_roofline_traffic_is_read_from_a_committed_profile():
    b, src = bench.ncu_traffic("c3")
    assert b and b > 1e9 and os.path.exists(os.path.join(ROOT, src.split(" ")[0]))
    assert bench.ncu_traffic("no-such-workload") == (None, None)


def test_opencv_baseline_path_agrees_with_the_oracle():
    """bench.py's second CPU engine -- cv2.BFMatcher in the reference's blocking, the routine matchBF calls -- returns the
    oracle's neighbours (indices on tie-free data, distances within the 1e-5 the north star states) on small random
    descriptors: the two CPU baselines time the same computation."""
    cv2 = pytest.importorskip("cv2")
    import bench
    from oracle import oracle as orc
    rng = np.random.default_rng(5)
    for dim, k in ((33, 2), (352, 5)):
        q = rng.random((300, dim), dtype=np.float32)
        t = rng.random((2500, dim), dtype=np.float32)
        idx, dist = bench.opencv_knn(q, t, k, block=1000)
        oi, od, oc = orc.knn(q, t, k)
        assert np.array_equal(idx, oi.astype(np.int64))
        assert np.allclose(dist, od, rtol=1e-5, atol=0)
        assert np.all(oc == k)
    assert cv2.__version__


def test_reference_arm_prints_exactly_one_json_line():
    """The driver parses stdout as ONE JSON line: run the CPU (reference) arm on the small C1 workload and check that
    stdout is exactly that -- whatever the libraries print goes to stderr -- and that the line carries the contract's keys."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[:2000]
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "queries/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("FPFH-33 20k")
