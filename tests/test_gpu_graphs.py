"""-m gpu: CUDA-graph replay of predict -> update -> resample (filter/_base.py: enable_graphs) is
bit-identical to eager execution, whatever is interleaved with the three calls."""
import numpy
import pytest

from gpu_common import consistent_measurement, make_gsf, make_pf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gpu_se_b200
    return gpu_se_b200


def _inputs(c):
    rng = numpy.random.default_rng(1000 + c)
    u = numpy.array([rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.3)])
    return u, consistent_measurement(u, 0.1, rng), float(rng.random())


@pytest.mark.parametrize("kind,N", [("pf", 1000), ("pf", 65536), ("pf", 300001), ("gsf", 700)])
def test_graph_replay_is_bit_identical(g, kind, N):
    make = make_pf if kind == "pf" else make_gsf
    a, b = make(g, N, seed=5), make(g, N, seed=5)
    b.enable_graphs()
    state = (lambda f: f.particles.get()) if kind == "pf" else (lambda f: numpy.concatenate(
        [f.means.get(), f.covariances.get().reshape(N, 25)], axis=1))
    for c in range(9):
        u, z, r = _inputs(c)
        dt = 0.1 if c != 4 else 0.25                      # dt is a per-step scalar too
        for f in (a, b):
            f.predict(u, dt)
            if c == 3:
                f.point_estimate()                        # something between the calls: the recorded predict is flushed
            f.update(u, z)
            f.resample(r=r)
        if c in (2, 6):
            assert numpy.array_equal(a.point_estimate(), b.point_estimate())      # through the pending index
        if c == 5:
            assert numpy.array_equal(state(a), state(b))  # materialises both: the next cycle starts in place
    assert numpy.array_equal(state(a), state(b))
    assert numpy.array_equal(a.weights.get(), b.weights.get())
    # one graph per state-buffer parity, and per flavour of the resample kernel (with / without the estimate)
    assert b.graph_replays >= 4 and 2 <= len(b._graphs) <= 4
    assert a.graph_replays == 0


def test_graph_mode_random_offsets_follow_numpy(g):
    """r defaults to numpy.random.rand() in graph mode as well (particle.py:93)."""
    a, b = make_pf(g, 5000, seed=9), make_pf(g, 5000, seed=9)
    b.enable_graphs()
    for f in (a, b):
        numpy.random.seed(77)
        for c in range(5):
            u, z, _ = _inputs(c)
            f.predict(u, 0.1)
            f.update(u, z)
            f.resample()
    assert numpy.array_equal(a.particles.get(), b.particles.get())
    b.enable_graphs(False)
    u, z, r = _inputs(11)
    for f in (a, b):
        f.predict(u, 0.1)
        f.update(u, z)
        f.resample(r=r)
    assert numpy.array_equal(a.particles.get(), b.particles.get())
