"""-m gpu: the closed-loop driver (gpu_se_b200/sim_base.py, mirroring sim_base.Simulation:207-309)
keeps the filter locked onto the plant and inside its real-time budget."""
import numpy
import pytest

@pytest.mark.gpu
@pytest.mark.parametrize("pf,N", [(True, 1 << 14), (False, 1 << 8)])
def test_closed_loop_tracks_plant(pf, N):
    from gpu_se_b200.sim_base import Simulation
    sim = Simulation(N, dt_control=0.1, dt_predict=0.1, end_time=20, pf=pf, seed=3).simulate()
    assert sim.predict_count == len(sim.ts) - 1 and sim.update_count >= len(sim.ts) - 2
    assert numpy.isfinite(sim.xs_f).all() and numpy.isfinite(sim.covariance_point_size).all()
    err = numpy.abs(sim.ys_f - sim.ys[:, list(sim.OUTPUTS)])[20:]
    # mg/L: the plant moves by its process noise (sigma 1.8 / 3.7 mg/L per step) after the last measurement
    assert numpy.median(err[:, 0]) < 4.0 and numpy.median(err[:, 1]) < 7.0
    assert sim.utilisation() < 0.05                                             # 6 s control period
    assert sim.performance >= 0.0


def test_performance_is_simpson():
    from gpu_se_b200.sim_base import performance
    ts = numpy.linspace(0, 2, 21)
    ys = numpy.zeros((21, 1))
    r = (ts ** 2)[:, None]                    # integral of t^4 over [0, 2] = 32 / 5
    assert performance(ys, r, ts) == pytest.approx(32 / 5, rel=1e-4)
    ts = numpy.linspace(0, 2, 20)
    assert performance(numpy.zeros((20, 1)), (ts ** 2)[:, None], ts) == pytest.approx(32 / 5, rel=1e-3)
