"""CPU tests of the host MPC (gpu_se_b200/controller.py): the quadratic programme is the
reference's (matrices compared entry for entry with controller.MPC built from /root/reference where that exists),
the ADMM solution satisfies the KKT conditions, and the closed loop on the linear model reaches the set point."""
import os
import subprocess
import sys

import numpy
import pytest
import scipy.sparse

from conftest import ROOT


class _Lin:
    def __init__(self, A, B, C, D):
        self.A, self.B, self.C, self.D = A, B, C, D


def _toy(seed=0, nx=2, ni=2, no=2, feedthrough=False):
    rng = numpy.random.default_rng(seed)
    A = numpy.diag(rng.uniform(0.5, 0.95, nx)) + 0.05 * rng.normal(size=(nx, nx))
    B = rng.normal(size=(nx, ni))
    C = rng.normal(size=(no, nx))
    D = 0.1 * rng.normal(size=(no, ni)) if feedthrough else numpy.zeros((no, ni))
    return _Lin(A, B, C, D)


def test_qp_solution_satisfies_kkt_and_the_model():
    from gpu_se_b200.controller import MPC
    lin = _toy(1)
    P, M = 12, 5
    K = MPC(P, M, numpy.diag([1.0, 2.0]), numpy.diag([0.1, 0.1]), lin, ysp=numpy.array([1.0, -0.5]),
            u_bounds=[(-0.3, 0.3), (-0.2, 0.4)], u_step_bounds=[(-0.25, 0.25), (-0.25, 0.25)], eps_abs=1e-7, eps_rel=1e-7)
    x0, um1, y0 = numpy.array([0.3, -0.2]), numpy.array([0.05, -0.05]), lin.C @ numpy.array([0.3, -0.2])
    u = K.step(x0, um1, y0)
    w, sol = K.last_solution, K.prob
    # KKT: primal feasibility, stationarity, complementarity
    Aw = K.A_matrix @ w
    assert (Aw >= K.l_matrix - 1e-5).all() and (Aw <= K.u_matrix + 1e-5).all()
    assert numpy.abs(K.H @ w + K.q + K.A_matrix.T @ sol.y).max() < 1e-5
    slack_lo, slack_hi = Aw - K.l_matrix, K.u_matrix - Aw
    assert (sol.y[slack_hi > 1e-4] <= 1e-5).all() and (sol.y[slack_lo > 1e-4] >= -1e-5).all()    # y > 0 only on an active upper bound
    # the programme means what its docstring says: outputs in w are those of the linear model under the planned inputs
    Nx, Ni, No = 2, 2, 2
    du = w[K._c_first_move:].reshape(M + 1, Ni)
    ys = w[K._c_first_output:K._c_first_output + P * No].reshape(P, No)
    x, uk = x0.copy(), um1.copy()
    for k in range(P):
        if k <= M - 1:
            uk = uk + du[k]
        x = lin.A @ x + lin.B @ uk
        assert numpy.allclose(ys[k], lin.C @ x, atol=1e-5), k
    assert numpy.allclose(u, um1 + du[0])
    assert -0.3 - 1e-6 <= u[0] <= 0.3 + 1e-6 and (numpy.abs(du) <= 0.25 + 1e-6).all()
    # bias correction on the next call (controller.py:258-266)
    y1_meas = ys[0] + numpy.array([0.02, -0.01])
    K.step(lin.A @ x0 + lin.B @ u, u, y1_meas)
    assert K.y_predicted is not None and K.prob.iterations > 0


def test_unconstrained_solution_equals_the_dense_kkt_solve():
    from gpu_se_b200.controller import MPC
    lin = _toy(2, feedthrough=True)
    K = MPC(8, 3, numpy.eye(2), 0.5 * numpy.eye(2), lin, ysp=numpy.array([0.4, 0.1]), eps_abs=1e-9, eps_rel=1e-9)
    x0, um1 = numpy.array([0.1, 0.2]), numpy.array([0.0, 0.1])
    K.step(x0, um1, numpy.zeros(2))
    eq = K.l_matrix == K.u_matrix
    Ae = K.A_matrix[eq].toarray()
    n, me = K.H.shape[0], Ae.shape[0]
    kkt = numpy.block([[K.H.toarray(), Ae.T], [Ae, numpy.zeros((me, me))]])
    rhs = numpy.concatenate([-K.q, K.l_matrix[eq]])
    dense = numpy.linalg.lstsq(kkt, rhs, rcond=None)[0][:n]
    assert numpy.allclose(K.last_solution, dense, atol=1e-6)


def test_mpc_drives_the_linearised_bioreactor_to_its_set_point():
    from gpu_se_b200.controller import get_controller
    lin, K = get_controller(dt_control=1.0)
    assert lin.A.shape == (2, 2) and K.P == 300 and K.M == 200
    assert abs(lin.y_bar[0] - 280.0) < 1.0                      # glucose is regulated to 280 mg/L by the organism
    x = numpy.zeros(2)
    u = numpy.zeros(2)
    for k in range(120):
        y = lin.C @ x
        u = K.step(x, u, y)
        assert (u + lin.u_bar >= -1e-6).all()                   # inputs stay non-negative
        x = lin.A @ x + lin.B @ u
    y = lin.yd2n(lin.C @ x)
    assert abs(y[1] - 850.0) < 0.05 * 850.0                     # fumaric acid set point (sim_base.py:80)


def test_programme_matches_the_reference_controller_entry_for_entry():
    """Where /root/reference exists: build controller.MPC from the reference (with a stand-in ``osqp`` module that
    only records the matrices) on the same linear model and compare H, q, A, l, u."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree not present")
    script = r'''
import sys, types, warnings
warnings.simplefilter("ignore")
import numpy, scipy.sparse
sys.path.insert(0, %(root)r)
captured = {}
class _OSQP:
    def setup(self, H, q, A, l, u, **kw): captured.update(H=H, q=q, A=A, l=l.copy(), u=u.copy())
    def update(self, **kw): pass
osqp = types.ModuleType("osqp"); osqp.OSQP = _OSQP; sys.modules["osqp"] = osqp
cupy = types.ModuleType("cupy"); cupy.float32 = "stub"; sys.modules["cupy"] = cupy
sys.path.insert(0, "/root/reference")
import controller as ref_controller
from gpu_se_b200.controller import MPC, get_controller
lin, mine = get_controller(dt_control=10.0)
bounds = [numpy.array([0, numpy.inf]) - lin.u_bar[0], numpy.array([0, numpy.inf]) - lin.u_bar[1]]
for kw in (dict(u_bounds=bounds), dict(u_bounds=bounds, y_bounds=[(-5.0, 7.0), (-100.0, 50.0)], u_step_bounds=[(-0.01, 0.02), (-0.03, 0.04)])):
    theirs = ref_controller.MPC(P=mine.P, M=mine.M, Q=mine.Q, R=mine.R, lin_model=lin, ysp=mine.ysp, **kw)
    ours = MPC(mine.P, mine.M, mine.Q, mine.R, lin, mine.ysp, **kw)
    assert abs(captured["H"] - ours.H).max() == 0
    assert numpy.array_equal(captured["q"], ours.q)
    assert abs(scipy.sparse.csc_matrix(captured["A"]) - ours.A_matrix).max() == 0
    assert numpy.array_equal(captured["l"], ours.l_matrix) and numpy.array_equal(captured["u"], ours.u_matrix)
# feed-through pattern (D != 0) as well
lin.D = numpy.array([[0.1, -0.2], [0.3, 0.05]])
theirs = ref_controller.MPC(P=7, M=4, Q=mine.Q, R=mine.R, lin_model=lin, ysp=mine.ysp)
ours = MPC(7, 4, mine.Q, mine.R, lin, mine.ysp)
assert abs(scipy.sparse.csc_matrix(captured["A"]) - ours.A_matrix).max() == 0
print("MATCH")
''' % {"root": ROOT}
    res = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "MATCH" in res.stdout, res.stdout + res.stderr
