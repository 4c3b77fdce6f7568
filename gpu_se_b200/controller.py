"""Host-side linear MPC for the closed loop: the reference's controller without OSQP.

The reference's ``controller.MPC`` (controller.py:9-279) states a sparse quadratic programme over a prediction
horizon and hands it to OSQP, which is not installable in this image.  This module keeps the interface
(``MPC(P, M, Q, R, lin_model, ysp, y_bounds, u_bounds, u_step_bounds)``, ``step(x0, um1, y0) -> u``) and the
programme -- same decision vector, same constraint rows, same cost, so ``step`` reads the same entries of the
solution -- and solves it with the algorithm OSQP implements (Stellato et al., "OSQP: an operator splitting solver
for quadratic programs", 2020): ADMM on the splitting ``A x = z, l <= z <= u`` with one sparse LU factorisation of
the quasi-definite KKT matrix (``scipy.sparse.linalg.splu``), over-relaxation, step sizes 10^3 times larger on
equality rows, and residual-balancing updates of the step size.  Everything runs on the host: one problem of
~1.6e3 (control period 1 min) to ~1.6e4 (0.1 min) variables per control period is not GPU work.

``LinearBioreactor`` is the internal model ``sim_base.get_parts`` builds with ``model.LinearModel``
(sim_base.py:56-73; model/LinearModel.py:69-159): the low-nitrogen bioreactor linearised about a steady state by
central differences, discretised with a zero-order hold, reduced to the states / inputs / outputs the controller
uses, with the deviation-variable conversions of LinearModel.py:161-272.
"""
import numpy
import scipy.linalg
import scipy.optimize
import scipy.signal
import scipy.sparse
import scipy.sparse.linalg

from gpu_se_b200.model.BioreactorModel import Bioreactor

MOLAR_MASS = numpy.array([180.0, 24.6, 116.0, 46.0, 1.0])          # BioreactorModel.py:119-121


def bioreactor_rates(x, u):
    """dx/dt of the low-nitrogen bioreactor (homeostatic_DEs with dt = 1, BioreactorModel.py:170-231)."""
    return numpy.array(Bioreactor.homeostatic_DEs(numpy.asarray(x, dtype=numpy.float64), u, 1.0), dtype=numpy.float64)


def steady_state(u_op, x_guess):
    """Bioreactor.find_SS (BioreactorModel.py:137-168): a root of dx/dt near ``x_guess`` with the biomass held at its
    guess (its rate is identically zero in this regime)."""
    x_guess = numpy.asarray(x_guess, dtype=numpy.float64)

    def residual(x):
        x = numpy.array(x, dtype=numpy.float64)
        x[1] = x_guess[1]
        return bioreactor_rates(x, u_op)

    root = scipy.optimize.fsolve(residual, x_guess)
    root[1] = x_guess[1]
    return root


def _central_gradient(fun, tol=1e-8, h=0.1):
    """(fun(h) - fun(-h)) / 2h, h halved until two successive estimates agree to ``tol`` in the maximum norm
    (LinearModel.py:93-106)."""
    grad = (fun(h) - fun(-h)) / (2 * h)
    for _ in range(60):
        h *= 0.5
        new = (fun(h) - fun(-h)) / (2 * h)
        done = numpy.max(numpy.abs(new - grad)) <= tol
        grad = new
        if done:
            break
    return grad


class LinearBioreactor:
    """Discrete linear model ``x+ = A x + B u, y = C x + D u`` in deviation variables about ``(x_bar, u_bar)``."""

    def __init__(self, x_bar, u_bar, T, states=(0, 2), inputs=(0, 1), outputs=(0, 2)):
        x_bar = numpy.asarray(x_bar, dtype=numpy.float64)
        u_bar = numpy.asarray(u_bar, dtype=numpy.float64)
        nx, nu = len(x_bar), len(u_bar)
        Ac = numpy.zeros((nx, nx))
        Bc = numpy.zeros((nx, nu))
        for k in range(nx):
            e = numpy.zeros(nx)
            e[k] = 1.0
            Ac[:, k] = _central_gradient(lambda s: bioreactor_rates(x_bar + s * e, u_bar))
        for k in range(nu):
            e = numpy.zeros(nu)
            e[k] = 1.0
            Bc[:, k] = _central_gradient(lambda s: bioreactor_rates(x_bar, u_bar + s * e))
        Cc = numpy.diag(MOLAR_MASS)                                 # outputs = concentrations in mg/L (:111-122)
        Dc = numpy.zeros((nx, nu))
        Ad, Bd, Cd, Dd, _ = scipy.signal.cont2discrete((Ac, Bc, Cc, Dc), T)
        self.dt = T
        self.states, self.inputs, self.outputs = list(states), list(inputs), list(outputs)
        self.A = Ad[numpy.ix_(self.states, self.states)]
        self.B = Bd[numpy.ix_(self.states, self.inputs)]
        self.C = Cd[numpy.ix_(self.outputs, self.states)]
        self.D = Dd[numpy.ix_(self.outputs, self.inputs)]
        self.x_bar_full, self.u_bar_full = x_bar, u_bar
        self.y_bar_full = x_bar * MOLAR_MASS
        self.x_bar, self.u_bar, self.y_bar = x_bar[self.states], u_bar[self.inputs], self.y_bar_full[self.outputs]
        self.Nx, self.Ni, self.No = len(self.states), len(self.inputs), len(self.outputs)

    # deviation-variable conversions (LinearModel.py:161-272)
    def xn2d(self, x, subselect=True):
        x = numpy.asarray(x, dtype=numpy.float64)
        return x[self.states] - self.x_bar if subselect else x - self.x_bar

    def yn2d(self, y, subselect=True):
        y = numpy.asarray(y, dtype=numpy.float64)
        return y[self.outputs] - self.y_bar if subselect else y - self.y_bar

    def un2d(self, u, subselect=True):
        u = numpy.asarray(u, dtype=numpy.float64)
        return u[self.inputs] - self.u_bar if subselect else u - self.u_bar

    def xd2n(self, x_hat):
        return numpy.asarray(x_hat) + self.x_bar

    def yd2n(self, y_hat):
        return numpy.asarray(y_hat) + self.y_bar

    def ud2n(self, u_hat):
        return numpy.asarray(u_hat) + self.u_bar


class QPNotSolved(ValueError):
    pass


class ADMMSolver:
    """min 1/2 x' P x + q' x  s.t.  l <= A x <= u  by the OSQP iteration; the matrices are fixed, ``l`` / ``u``
    change between solves and every solve is warm-started from the previous one.  Defaults are OSQP's (the reference
    calls it with defaults): tolerances 1e-3, 4000 iterations, 10 passes of Ruiz equilibration, sigma 1e-6, alpha 1.6."""

    def __init__(self, P, q, A, l, u, rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3, max_iter=4000,
                 check_every=10, scaling=10):
        P = scipy.sparse.csc_matrix(P, dtype=numpy.float64)
        A = scipy.sparse.csc_matrix(A, dtype=numpy.float64)
        q = numpy.asarray(q, dtype=numpy.float64)
        self.n, self.m = P.shape[0], A.shape[0]
        # Ruiz equilibration of the KKT matrix [[P, A'], [A, 0]]: x = D xs, y = E ys / c (OSQP section 5.1)
        D, E = numpy.ones(self.n), numpy.ones(self.m)
        Ps, As = P.copy(), A.copy()
        for _ in range(int(scaling)):
            col_p = abs(Ps).max(axis=0).toarray().ravel()
            col_a = abs(As).max(axis=0).toarray().ravel()
            row_a = abs(As).max(axis=1).toarray().ravel()
            d = 1.0 / numpy.sqrt(numpy.maximum(numpy.maximum(col_p, col_a), 1e-4))
            e = 1.0 / numpy.sqrt(numpy.maximum(row_a, 1e-4))
            Dm, Em = scipy.sparse.diags(d), scipy.sparse.diags(e)
            Ps, As = (Dm @ Ps @ Dm).tocsc(), (Em @ As @ Dm).tocsc()
            D, E = D * d, E * e
        self.D, self.E = D, E
        self.P, self.A, self.q = Ps, As, D * q
        self.P0, self.A0, self.q0 = P, A, q
        self.sigma, self.alpha = sigma, alpha
        self.eps_abs, self.eps_rel, self.max_iter, self.check_every = eps_abs, eps_rel, max_iter, check_every
        self.x = numpy.zeros(self.n)                # scaled iterates
        self.z = numpy.zeros(self.m)
        self.ys = numpy.zeros(self.m)
        self.y = numpy.zeros(self.m)                # multipliers of the original problem
        self.iterations = 0
        self.factorisations = 0
        self._rho = rho
        self._eq = None
        self._set_bounds(l, u)

    def _set_bounds(self, l, u):
        self.l0 = numpy.asarray(l, dtype=numpy.float64).copy()
        self.u0 = numpy.asarray(u, dtype=numpy.float64).copy()
        self.l, self.u = self.E * self.l0, self.E * self.u0
        eq = self.l0 == self.u0
        if self._eq is None or not numpy.array_equal(eq, self._eq):
            self._eq = eq
            self._factorise()

    def _factorise(self):
        self.rho_vec = numpy.where(self._eq, 1e3 * self._rho, self._rho)
        kkt = scipy.sparse.bmat([[self.P + self.sigma * scipy.sparse.identity(self.n), self.A.T],
                                 [self.A, scipy.sparse.diags(-1.0 / self.rho_vec)]], format="csc")
        self._lu = scipy.sparse.linalg.splu(kkt)
        self.factorisations += 1

    def solve(self, l=None, u=None):
        if l is not None:
            self._set_bounds(l, u)
        x, z, y = self.x, numpy.clip(self.z, self.l, self.u), self.ys
        P, A, q, n = self.P, self.A, self.q, self.n
        r_prim = r_dual = numpy.inf
        for it in range(1, self.max_iter + 1):
            rhs = numpy.concatenate([self.sigma * x - q, z - y / self.rho_vec])
            sol = self._lu.solve(rhs)
            xt = sol[:n]
            zt = z + (sol[n:] - y) / self.rho_vec
            x = self.alpha * xt + (1 - self.alpha) * x
            zr = self.alpha * zt + (1 - self.alpha) * z
            z_new = numpy.clip(zr + y / self.rho_vec, self.l, self.u)
            y = y + self.rho_vec * (zr - z_new)
            z = z_new
            if it % self.check_every and it != self.max_iter:
                continue
            # residuals of the ORIGINAL problem
            Ax, Px, Aty = (A @ x) / self.E, (P @ x) / self.D, (A.T @ y) / self.D
            zu = z / self.E
            r_prim = numpy.max(numpy.abs(Ax - zu)) if self.m else 0.0
            r_dual = numpy.max(numpy.abs(Px + self.q0 + Aty))
            n_prim = max(numpy.max(numpy.abs(Ax)), numpy.max(numpy.abs(zu)), 1e-30)
            n_dual = max(numpy.max(numpy.abs(Px)), numpy.max(numpy.abs(Aty)), numpy.max(numpy.abs(self.q0)), 1e-30)
            if r_prim <= self.eps_abs + self.eps_rel * n_prim and r_dual <= self.eps_abs + self.eps_rel * n_dual:
                break
            # residual balancing (OSQP section 5.2): refactorise when the estimate moved by more than 5x
            ratio = numpy.sqrt((r_prim / n_prim + 1e-30) / (r_dual / n_dual + 1e-30))
            if ratio > 5.0 or ratio < 0.2:
                self._rho = float(min(max(self._rho * ratio, 1e-6), 1e6))
                self._factorise()
        else:
            self.x, self.z, self.ys, self.y, self.iterations = x, z, y, self.E * y, self.max_iter
            raise QPNotSolved("ADMM did not reach the tolerances in %d iterations (primal %.2e, dual %.2e)"
                              % (self.max_iter, r_prim, r_dual))
        self.x, self.z, self.ys, self.y, self.iterations = x, z, y, self.E * y, it
        return self.D * x


class MPC:
    """Linear MPC with the reference's programme (controller.py:63-238).

    Decision vector  w = [ mu_0, dmu_1 .. dmu_P | y_1 .. y_P | u_-1 | du_0 .. du_M ]  (deviation variables; state
    INCREMENTS from the second block on, as the reference formulates it), cost
    1/2 sum_k (y_k - ysp)' Q (y_k - ysp) + 1/2 sum_k du_k' R du_k, rows in the reference's order:

      (a) u_-1 = um1                                  (b) -mu_0 = -x0;  (A - I) mu_0 - dmu_1 + B (u_-1 + du_0) = 0;
      (c) y_k - y_{k-1} = C dmu_k + D(..) + bias          A dmu_{k-1} - dmu_k + B du_{k-1} = 0  (no input beyond M)
      (d) y_min <= y_k <= y_max     (e) du_min <= du_k <= du_max     (f) u_min <= u_-1 + du_0 <= u_max
    """

    def __init__(self, P, M, Q, R, lin_model, ysp, y_bounds=None, u_bounds=None, u_step_bounds=None, **solver_options):
        self.P, self.M, self.Q, self.R, self.model = int(P), int(M), numpy.asarray(Q, float), numpy.asarray(R, float), lin_model
        self.ysp = numpy.asarray(ysp, dtype=numpy.float64)
        A, B, C, D = (numpy.asarray(m, dtype=numpy.float64) for m in (lin_model.A, lin_model.B, lin_model.C, lin_model.D))
        Nx, Ni = B.shape
        No = C.shape[0]
        P_, M_ = self.P, self.M
        if not (1 <= M_ <= P_):
            raise ValueError("need 1 <= M <= P")

        def bounds(b, n):
            if b is None:
                return numpy.full(n, -numpy.inf), numpy.full(n, numpy.inf)
            lo, hi = zip(*b)
            return numpy.asarray(lo, dtype=numpy.float64), numpy.asarray(hi, dtype=numpy.float64)

        y_min, y_max = bounds(y_bounds, No)
        u_min, u_max = bounds(u_bounds, Ni)
        du_min, du_max = bounds(u_step_bounds, Ni)

        # column offsets of the decision vector
        c_x = lambda k: k * Nx                                     # noqa: E731  mu_0 / dmu_k
        c_y = lambda k: (P_ + 1) * Nx + k * No                     # noqa: E731  y_{k+1}, k = 0 .. P-1
        c_um1 = (P_ + 1) * Nx + P_ * No
        c_du = lambda k: c_um1 + Ni + k * Ni                       # noqa: E731  du_k, k = 0 .. M
        n = c_du(M_ + 1)
        self._c_first_move, self._c_first_output = c_du(0), c_y(0)

        rows, cols, vals = [], [], []

        def put(r0, c0, block):
            block = numpy.atleast_2d(block)
            for i in range(block.shape[0]):
                for j in range(block.shape[1]):
                    if block[i, j] != 0.0:
                        rows.append(r0 + i)
                        cols.append(c0 + j)
                        vals.append(block[i, j])

        eye_x, eye_y, eye_u = numpy.eye(Nx), numpy.eye(No), numpy.eye(Ni)
        r = 0
        # (a) the previous input
        put(r, c_um1, eye_u)
        self._r_um1 = r
        r += Ni
        # (b) state recursion in increments
        self._r_x0 = r
        put(r, c_x(0), -eye_x)
        r += Nx
        for k in range(1, P_ + 1):
            if k == 1:
                put(r, c_x(0), A - eye_x)
                put(r, c_um1, B)
            else:
                put(r, c_x(k - 1), A)
            put(r, c_x(k), -eye_x)
            if k <= M_:
                put(r, c_du(k - 1), B)
            r += Nx
        # (c) outputs accumulate the increments; the feed-through pattern follows controller.py:168-177
        self._r_out = r
        for k in range(P_):
            if k == 0:
                put(r, c_x(0), C)
            put(r, c_x(k + 1), C)
            put(r, c_y(k), -eye_y)
            if k >= 1:
                put(r, c_y(k - 1), eye_y)
            if k < M_:
                if k == 0:
                    put(r, c_um1, D)
                    put(r, c_du(0), D)
                put(r, c_du(k + 1), D)
            r += No
        n_eq = r
        # (d) output limits, (e) move limits, (f) limits on the first input
        put(r, c_y(0), numpy.eye(P_ * No))
        r += P_ * No
        put(r, c_du(0), numpy.eye((M_ + 1) * Ni))
        r += (M_ + 1) * Ni
        put(r, c_um1, eye_u)
        put(r, c_du(0), eye_u)
        r += Ni
        m = r
        self.A_matrix = scipy.sparse.csc_matrix((vals, (rows, cols)), shape=(m, n))
        self.l_matrix = numpy.concatenate([numpy.zeros(n_eq), numpy.tile(y_min, P_), numpy.tile(du_min, M_ + 1), u_min])
        self.u_matrix = numpy.concatenate([numpy.zeros(n_eq), numpy.tile(y_max, P_), numpy.tile(du_max, M_ + 1), u_max])
        hdiag = [scipy.sparse.csc_matrix(((P_ + 1) * Nx, (P_ + 1) * Nx)), scipy.sparse.kron(scipy.sparse.identity(P_), self.Q),
                 scipy.sparse.csc_matrix((Ni, Ni)), scipy.sparse.kron(scipy.sparse.identity(M_ + 1), self.R)]
        self.H = scipy.sparse.block_diag(hdiag, format="csc")
        self.q = numpy.concatenate([numpy.zeros((P_ + 1) * Nx), numpy.tile(-self.Q @ self.ysp, P_), numpy.zeros((M_ + 2) * Ni)])
        self._dims = (Nx, Ni, No)
        self.prob = ADMMSolver(self.H, self.q, self.A_matrix, self.l_matrix, self.u_matrix, **solver_options)
        self.y_predicted = None
        self.last_solution = None

    def step(self, x0, um1, y0):
        """The control input for the next period (controller.py:240-279): deviation-variable state estimate ``x0``,
        previous input ``um1`` and measured output ``y0`` in, input (deviation variables) out.  Raises ``ValueError``
        when the programme is not solved to tolerance, as the reference does."""
        Nx, Ni, No = self._dims
        x0 = numpy.clip(numpy.asarray(x0, dtype=numpy.float64), -1e10, 1e10)
        um1 = numpy.clip(numpy.asarray(um1, dtype=numpy.float64), -1e10, 1e10)
        y0 = numpy.clip(numpy.asarray(y0, dtype=numpy.float64), -1e10, 1e10)
        bias = y0 - self.y_predicted if self.y_predicted is not None else numpy.zeros_like(y0)     # :258-261
        for vec in (self.l_matrix, self.u_matrix):
            vec[self._r_um1:self._r_um1 + Ni] = um1
            vec[self._r_x0:self._r_x0 + Nx] = -x0
            vec[self._r_out:self._r_out + self.P * No] = numpy.tile(-bias, self.P)
        w = self.prob.solve(self.l_matrix, self.u_matrix)
        self.last_solution = w
        ctrl = w[self._c_first_move:self._c_first_move + Ni] + um1
        self.y_predicted = w[self._c_first_output:self._c_first_output + No] - bias
        return ctrl


def get_controller(dt_control=1.0, **solver_options):
    """(lin_model, K) as ``sim_base.get_parts`` builds them (sim_base.py:46-87): linearised about the steady state of
    u = (0.04, 0.1), states (Cg, Cfa), outputs (glucose, fumaric acid), horizons 300 / 200 min, Q = diag(0.1, 1),
    R = I, set point (280, 850) mg/L, inputs non-negative."""
    x_bar = steady_state(numpy.array([0.04, 0.1]), numpy.array([260 / 180, 640 / 24.6, 1000 / 116, 0, 0]))
    lin_model = LinearBioreactor(x_bar, numpy.array([0.04, 0.1]), dt_control)
    K = MPC(P=int(300 // dt_control), M=max(int(200 // dt_control), 1), Q=numpy.diag([0.1, 1.0]), R=numpy.diag([1.0, 1.0]),
            lin_model=lin_model, ysp=lin_model.yn2d(numpy.array([280.0, 850.0]), subselect=False),
            u_bounds=[numpy.array([0, numpy.inf]) - lin_model.u_bar[0], numpy.array([0, numpy.inf]) - lin_model.u_bar[1]],
            **solver_options)
    return lin_model, K


class MPCController:
    """The plug-in ``controller(x_estimate, u_previous, y_measured) -> u`` of ``gpu_se_b200.sim_base.Simulation`` around
    ``MPC.step`` with the deviation-variable conversions and the fallback input of sim_base.py:264-275."""

    def __init__(self, dt_control=1.0, **solver_options):
        self.lin_model, self.K = get_controller(dt_control, **solver_options)
        self.converged, self.failed = 0, 0
        self.iterations = []

    def __call__(self, x_estimate, u_previous, y_measured):
        lm = self.lin_model
        u = numpy.array(u_previous, dtype=numpy.float64)
        try:
            du = self.K.step(lm.xn2d(x_estimate), lm.un2d(u_previous), lm.yn2d(y_measured))
            self.converged += 1
            self.iterations.append(self.K.prob.iterations)
            # ADMM at OSQP's default tolerance (1e-3 of the largest row: the outputs, hundreds of mg/L) leaves the first
            # move feasible only to a few 1e-2 L/h; feed rates cannot be negative, so the move is projected onto the bound
            u[lm.inputs] = numpy.maximum(lm.ud2n(du), 0.0)
        except ValueError:
            self.failed += 1
            u[:] = [0.06, 0.2]                                     # sim_base.py:271-273
        return u

    @property
    def mpc_frac(self):
        return self.converged / max(self.converged + self.failed, 1)
