"""Closed-loop bioreactor simulation around the B200 filters: the caller side of the hot path.

Mirrors the loop of the reference's ``sim_base.Simulation`` (sim_base.py:207-309) -- predict every
simulation step, update + resample + point_estimate every control period, point_estimate and
point_covariance logged every step -- and its noise / factory definitions (``get_noise``,
sim_base.py:117-161; ``get_parts``, :10-114) for the parts that are on the state-estimation path.

The controller is the reference's linear MPC (controller.py:9-279 with the internal model of sim_base.py:56-87),
restated in ``gpu_se_b200/controller.py`` with an ADMM solver in place of OSQP (not installable in this image); any
callable ``controller(x_estimate, u_previous, y_measured) -> u`` can be plugged in instead (``controller="pi"``: a small
proportional-integral law for timing studies).  The plant is the low-nitrogen bioreactor
(``Bioreactor.homeostatic_DEs``, the regime the filters model) stepped on the host.
"""
import time

import numpy

from gpu_se_b200.gaussian_sum_dist import MultivariateGaussianSum
from gpu_se_b200.model.BioreactorModel import X_STEADY, Bioreactor

MOLAR_MASS = numpy.array([180.0, 24.6, 116.0, 46.0, 1.0])          # BioreactorModel.py:119-121


def get_noise(device=None, seed=None):
    """(state_pdf, measurement_pdf) of the benchmark (sim_base.py:141-160)."""
    state_pdf = MultivariateGaussianSum(
        means=numpy.zeros((2, 5)),
        covariances=numpy.array([numpy.diag([1e-4, 1e-7, 1e-3, 1e-3, 1e-7]), numpy.diag([1e-3, 1e-6, 1e-2, 1e-2, 1e-6])]),
        weights=numpy.array([0.75, 0.25]), device=device, seed=seed)
    measurement_pdf = MultivariateGaussianSum(
        means=numpy.array([[1e-1, 0], [0, -1e-1]]),
        covariances=numpy.array([[[6e-2, 0], [0, 8e-2]], [[500, 100], [100, 700]]]),
        weights=numpy.array([0.85, 0.15]), device=device, seed=None if seed is None else seed + 1)
    return state_pdf, measurement_pdf


def get_filter(N_particles=2 * 15, pf=True, device=None, seed=None):
    """The filter ``get_parts`` builds (sim_base.py:89-112): x0 = state noise shifted by the steady
    state, f / g the static bioreactor functions.  ``N_particles`` defaults as the reference (:10)."""
    import gpu_se_b200 as g
    state_pdf, measurement_pdf = get_noise(device, seed)
    x0 = MultivariateGaussianSum(means=state_pdf.means + numpy.array(X_STEADY)[None, :],
                                 covariances=state_pdf._covariances64, weights=state_pdf.weights, device=device)
    cls = g.ParallelParticleFilter if pf else g.ParallelGaussianSumUnscentedKalmanFilter
    return cls(f=Bioreactor.homeostatic_DEs, g=Bioreactor.static_outputs, N_particles=N_particles, x0=x0,
               state_pdf=state_pdf, measurement_pdf=measurement_pdf, device=device, seed=seed)


class HostBioreactor:
    """Plant: the low-nitrogen bioreactor state stepped with explicit Euler (BioreactorModel.py:95-109)."""

    def __init__(self, X0=X_STEADY):
        self.X = numpy.array(X0, dtype=numpy.float64)
        self.t = 0.0

    def step(self, dt, inputs):
        self.t += dt
        self.X += numpy.array(Bioreactor.homeostatic_DEs(self.X, inputs, dt))
        self.X[:4] = numpy.maximum(self.X[:4], 0)                              # :109

    def outputs(self, inputs):
        return self.X * MOLAR_MASS                                              # :111-122


class GlucosePI:
    """PI law on the measured glucose concentration (mg/L) that trims the glucose feed around its nominal value: a
    cheap stand-in for the MPC when only the filter is being timed (``controller="pi"``)."""

    def __init__(self, setpoint=280.0, kp=2e-4, ki=2e-5, u_nominal=(0.06, 0.2), limits=(0.0, 0.2)):
        self.setpoint, self.kp, self.ki = setpoint, kp, ki
        self.u_nominal = numpy.array(u_nominal, dtype=numpy.float64)
        self.limits = limits
        self.integral = 0.0

    def __call__(self, x_estimate, u_previous, y_measured):
        y = Bioreactor.static_outputs(x_estimate, u_previous)[0]
        err = self.setpoint - y
        self.integral += err
        u = self.u_nominal.copy()
        u[0] = min(max(u[0] + self.kp * err + self.ki * self.integral, self.limits[0]), self.limits[1])
        return u


class Simulation:
    """Holds details of a closed-loop simulation (constructor as sim_base.Simulation, :209)."""

    OUTPUTS = (0, 2)           # measured outputs: glucose and fumaric acid (lin_model.outputs)

    def __init__(self, N_particles, dt_control, dt_predict, end_time=50, pf=True, controller=None, device=None,
                 seed=0):
        self.ts = numpy.linspace(0, end_time, int(end_time * 10))               # :211
        self.dt = self.ts[1]
        self.dt_control, self.dt_predict = dt_control, dt_predict
        self.bioreactor = HostBioreactor()
        if controller is None or controller == "mpc":           # the reference's closed loop (sim_base.py:75-87)
            from gpu_se_b200.controller import MPCController
            controller = MPCController(dt_control)
        elif controller == "pi":
            controller = GlucosePI()
        self.K = controller
        self.f = get_filter(N_particles, pf, device, seed)
        self.state_pdf, self.measurement_pdf = get_noise(device, seed + 100)
        self._rng = numpy.random.default_rng(seed)
        self.us = [numpy.array([0.06, 0.2])]
        self.xs = [self.bioreactor.X.copy()]
        self.ys = [self.bioreactor.outputs(self.us[-1])]
        self.ys_meas = [self.bioreactor.outputs(self.us[-1])]
        self.xs_f = [self.f.point_estimate()]
        self.ys_f = [numpy.array(Bioreactor.static_outputs(self.f.point_estimate(), self.us[-1]))]
        self.covariance_point_size = [self.f.point_covariance()]
        self.predict_count, self.update_count = 0, 0
        self.filter_seconds = []       # wall time of the filter calls in each control period (synchronised)
        self.performance = None

    def _host_draw(self, pdf):
        """One sample of a mixture for the host-side plant: drawn by the device sampler and copied over, exactly as
        the reference does it (``state_pdf.draw().get()``, sim_base.py:281-284)."""
        return numpy.asarray(pdf.draw().get()[0], dtype=numpy.float64)

    def simulate(self):
        """The loop of sim_base.Simulation.simulate (:243-309)."""
        t_next_control, t_next_predict = 0, 0
        for t in self.ts[1:]:
            spent = 0.0
            if t > t_next_predict:
                t0 = time.perf_counter()
                self.f.predict(self.us[-1], self.dt)                            # :249 (self.dt, not dt_predict)
                spent += time.perf_counter() - t0
                self.predict_count += 1
                t_next_predict += self.dt_predict
            if t > t_next_control:
                t0 = time.perf_counter()
                self.f.update(self.us[-1], self.ys_meas[-1][list(self.OUTPUTS)])  # :258
                self.f.resample()
                self.update_count += 1
                self.xs_f.append(self.f.point_estimate())                       # :262 (synchronises)
                spent += time.perf_counter() - t0
                self.us.append(numpy.asarray(self.K(self.xs_f[-1], self.us[-1], self.ys_meas[-1]), dtype=numpy.float64))
                t_next_control += self.dt_control
            else:
                self.us.append(self.us[-1])
            self.bioreactor.step(self.dt, self.us[-1])                          # :280
            self.bioreactor.X += self._host_draw(self.state_pdf)
            outputs = self.bioreactor.outputs(self.us[-1])
            self.ys.append(outputs.copy())
            outputs[list(self.OUTPUTS)] += self._host_draw(self.measurement_pdf)
            self.ys_meas.append(outputs)
            self.xs.append(self.bioreactor.X.copy())
            t0 = time.perf_counter()
            est = self.f.point_estimate()                                       # :287-295, every step
            self.covariance_point_size.append(self.f.point_covariance())
            spent += time.perf_counter() - t0
            self.ys_f.append(numpy.array(Bioreactor.static_outputs(est, self.us[-1])))
            self.filter_seconds.append(spent)
        for name in ("us", "xs", "ys", "ys_meas", "xs_f", "ys_f", "covariance_point_size"):
            setattr(self, name, numpy.array(getattr(self, name)))
        self.performance = performance(self.ys[:, list(self.OUTPUTS)], self.ys_f, self.ts)
        return self

    def utilisation(self):
        """Filter run time per control period over the period itself, periods in minutes
        (results/pf_closedloop/bioreactor_performance_pf.py:157)."""
        steps_per_period = max(1, int(round(self.dt_control / self.dt)))
        per_period = numpy.add.reduceat(numpy.asarray(self.filter_seconds),
                                        numpy.arange(0, len(self.filter_seconds), steps_per_period))
        return float(numpy.median(per_period) / (self.dt_control * 60.0))


def _simpson_avg(y, x):
    """scipy.integrate.simps(y, x) as the reference's SciPy evaluated it (``even='avg'``): composite Simpson for an
    odd number of samples; for an even number the average of (Simpson on the first N-1 samples + trapezoid on the
    last interval) and (trapezoid on the first interval + Simpson on the last N-1 samples)."""
    y, x = numpy.asarray(y, dtype=numpy.float64), numpy.asarray(x, dtype=numpy.float64)
    n = len(x)
    if n < 2:
        return 0.0
    if n == 2:
        return 0.5 * (x[1] - x[0]) * (y[0] + y[1])

    def odd(yy, xx):                                   # composite Simpson, non-uniform spacing allowed
        h0, h1 = numpy.diff(xx)[0::2], numpy.diff(xx)[1::2]
        hs, hp, hd = h0 + h1, h0 * h1, h0 / h1
        return float(numpy.sum(hs / 6.0 * (yy[0:-2:2] * (2.0 - 1.0 / hd) + yy[1:-1:2] * (hs * hs / hp)
                                           + yy[2::2] * (2.0 - hd))))

    if n % 2 == 1:
        return odd(y, x)
    first = odd(y[:-1], x[:-1]) + 0.5 * (x[-1] - x[-2]) * (y[-1] + y[-2])
    last = odd(y[1:], x[1:]) + 0.5 * (x[1] - x[0]) * (y[0] + y[1])
    return 0.5 * (first + last)


def performance(ys, r, ts):
    """The reference's closed-loop performance measure (sim_base.py:164-185): per output axis the TIME-WEIGHTED
    squared error ``(ys - r)**2 * ts`` integrated over ``ts`` with Simpson's rule, summed over the axes."""
    se = (numpy.asarray(ys, dtype=numpy.float64) - numpy.asarray(r, dtype=numpy.float64)) ** 2
    ts = numpy.asarray(ts, dtype=numpy.float64)
    return float(sum(_simpson_avg(se_ax * ts, ts) for se_ax in numpy.rollaxis(se, 1)))
