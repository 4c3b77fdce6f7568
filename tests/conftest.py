import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return numpy.load(os.path.join(GOLDEN, name))


def ulp32(err, ref):
    """|err| in units of float32 ulp at max(|ref|, 1) (compare STATES, not increments)."""
    scale = numpy.spacing(numpy.maximum(numpy.abs(ref), 1).astype(numpy.float32)).astype(numpy.float64)
    return numpy.abs(err) / scale


@pytest.fixture(scope="session")
def noise_pdfs():
    from oracle import mixture
    return mixture.benchmark_noise()
