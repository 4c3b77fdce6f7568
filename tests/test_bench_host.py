"""CPU: host-side pieces of bench.py and of the ctypes layer that need no GPU -- the ncu traffic figures read from the
committed capture, the clock sampler's line parsing and bracket window, the two-value argument conversion."""
import ctypes
import os
import time

import numpy
import pytest

from conftest import ROOT


def test_ncu_traffic_is_read_from_the_committed_capture():
    import bench
    per_row, src = bench.ncu_traffic_bytes_per_row()
    assert os.path.exists(bench.NCU_CAPTURE) and "read from" in src
    # DRAM traffic per row can exceed the algorithmic bytes only by the wasted sectors of the gather, and the fused
    # resample / update stay below theirs because part of their streams lives in L2 between the kernels
    for stage, lo, hi in (("predict", 20.0, 60.0), ("update", 6.0, 14.0), ("resample", 4.0, 12.0)):
        assert lo < per_row[stage] < hi, (stage, per_row[stage])
    # stages the capture does not hold keep the constants
    assert per_row["scan"] == bench.NCU_TRAFFIC_BYTES_PER_ROW["scan"]


def test_ncu_traffic_falls_back_to_the_constants(monkeypatch, tmp_path):
    import bench
    monkeypatch.setattr(bench, "NCU_CAPTURE", str(tmp_path / "missing.csv"))
    per_row, src = bench.ncu_traffic_bytes_per_row()
    assert per_row == bench.NCU_TRAFFIC_BYTES_PER_ROW and "constants" in src


def test_clock_sampler_keeps_the_samples_that_bracket_the_region():
    import bench

    class Done:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    s = bench.ClockSampler(0)
    s.proc = Done()
    t0 = time.perf_counter()
    line = "0, %d, 1965, 700.0, 0x0000000000000000, Not Active, Not Active, Not Active, %s"
    s.lines = [(t0 - 1.0, line % (1200, "Active")),           # long before the region: ignored
               (t0 - 0.002, line % (1965, "Not Active")),     # right in front of it
               (t0 + 0.030, line % (1950, "Active")),         # inside: sw_power_cap is kept and reported
               (t0 + 0.052, line % (1965, "Not Active")),     # right behind it
               (t0 + 0.500, line % (1000, "Not Active"))]     # after the bracket: ignored
    out = s.stop(t0, t0 + 0.05)
    assert out["samples"] == 3 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
    none = bench.ClockSampler(0)
    assert none.stop(0.0, 1.0)["sm_mhz"] is None               # nvidia-smi not available


def test_as_double2_accepts_what_the_reference_callers_pass():
    from gpu_se_b200 import _lib
    for v in ([0.06, 0.2], (0.06, 0.2), numpy.array([0.06, 0.2]), numpy.array([0.06, 0.2], dtype=numpy.float32),
              numpy.array([[0.06], [0.2]])):
        a = _lib.as_double2(v)
        assert isinstance(a, ctypes.c_double * 2)
        assert a[0] == pytest.approx(0.06, rel=1e-6) and a[1] == pytest.approx(0.2, rel=1e-6)
    for bad in ([1.0], [1.0, 2.0, 3.0], numpy.zeros((2, 2)), 1.0):
        with pytest.raises((ValueError, TypeError)):
            _lib.as_double2(bad)


def test_every_cited_profile_exists():
    """DESIGN.md / BASELINE.md / README.md / bench.py cite files under profiles/ as evidence: none of them may dangle."""
    import glob
    import re
    cited = set()
    for doc in ("DESIGN.md", "BASELINE.md", "README.md", "INTEGRATION.md", "bench.py"):
        text = open(os.path.join(ROOT, doc)).read()
        cited |= set(re.findall(r"profiles/([A-Za-z0-9_.*-]+\.(?:json|csv|txt))", text))
        # the tables name files without the directory once it is clear from the context: `r2_launches.csv`
        cited |= set(re.findall(r"`(r[12]_[A-Za-z0-9_.*-]+\.(?:json|csv|txt))`", text))
    assert len(cited) >= 10
    missing = sorted(c for c in cited if not glob.glob(os.path.join(ROOT, "profiles", c)))
    assert not missing, "cited but not under profiles/: %s" % ", ".join(missing)
