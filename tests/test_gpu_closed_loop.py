"""-m gpu: the closed-loop driver (gpu_se_b200/sim_base.py, mirroring sim_base.Simulation:207-309)
keeps the filter locked onto the plant and inside its real-time budget."""
import numpy
import pytest

@pytest.mark.gpu
@pytest.mark.parametrize("pf,N", [(True, 1 << 14), (False, 1 << 8)])
def test_closed_loop_tracks_plant(pf, N):
    from gpu_se_b200.sim_base import Simulation
    sim = Simulation(N, dt_control=0.1, dt_predict=0.1, end_time=20, pf=pf, seed=3, controller="pi").simulate()
    assert sim.predict_count == len(sim.ts) - 1 and sim.update_count >= len(sim.ts) - 2
    assert numpy.isfinite(sim.xs_f).all() and numpy.isfinite(sim.covariance_point_size).all()
    err = numpy.abs(sim.ys_f - sim.ys[:, list(sim.OUTPUTS)])[20:]
    # ys_f is the estimate BEFORE the step's measurement is assimilated, so it trails the plant by one step of process
    # noise: sigma_Cg = sqrt(.75e-4 + .25e-3) = 0.0180 -> 3.24 mg/L glucose, sigma_Cfa = sqrt(.75e-3 + .25e-2) = 0.0570
    # -> 6.61 mg/L fatty acid (sim_base.py:141-151).  The median of |N(0, sigma)| is 0.6745 sigma = 2.19 / 4.46 mg/L;
    # with 180 samples the sample median scatters by ~0.1 sigma, the filter's own error adds ~0.25 mg/L.
    assert numpy.median(err[:, 0]) < 3.0 and numpy.median(err[:, 1]) < 6.0
    assert sim.utilisation() < 0.05                                             # 6 s control period
    assert sim.performance >= 0.0


@pytest.mark.gpu
def test_closed_loop_with_the_mpc_reaches_the_set_point():
    """BASELINE.json configs[4] with the reference's controller (host ADMM in place of OSQP): the filter feeds the MPC,
    every QP is solved, and the fumaric-acid output heads for its 850 mg/L set point (sim_base.py:80)."""
    from gpu_se_b200.sim_base import Simulation
    sim = Simulation(1 << 14, dt_control=1.0, dt_predict=0.1, end_time=150, pf=True, seed=5).simulate()
    assert sim.K.mpc_frac == 1.0 and sim.update_count >= 149
    y = sim.ys[:, list(sim.OUTPUTS)]
    assert abs(y[-100:, 0].mean() - 280.0) < 10.0
    assert y[-1, 1] > y[0, 1] + 50.0                              # started at 611 mg/L, driven upwards
    assert (sim.us >= -1e-9).all()


def test_performance_is_time_weighted_simpson():
    """sim_base.py:181-185: the integrand is (ys - r)**2 * ts, integrated with scipy.integrate.simps (even='avg')."""
    from gpu_se_b200.sim_base import _simpson_avg, performance
    ts = numpy.linspace(0, 2, 21)
    ys = numpy.zeros((21, 1))
    r = (ts ** 2)[:, None]                    # integral of t * t^4 over [0, 2] = 32 / 3
    assert performance(ys, r, ts) == pytest.approx(32 / 3, rel=1e-4)
    ts = numpy.linspace(0, 2, 20)             # even sample count: the 'avg' rule
    assert performance(numpy.zeros((20, 1)), (ts ** 2)[:, None], ts) == pytest.approx(32 / 3, rel=1e-3)
    # two axes add up; Simpson is exact for cubics on an odd number of samples
    ts = numpy.linspace(0, 3, 7)
    two = performance(numpy.zeros((7, 2)), numpy.stack([ts, numpy.ones(7)], axis=1), ts)
    assert two == pytest.approx(3 ** 4 / 4 + 3 ** 2 / 2, rel=1e-12)
    try:
        from scipy.integrate import simpson
    except ImportError:
        return
    rng = numpy.random.default_rng(0)
    x = numpy.sort(rng.random(31)) * 5
    y = rng.random(31)
    assert _simpson_avg(y, x) == pytest.approx(float(simpson(y, x=x)), rel=1e-12)     # odd N: the same composite rule
