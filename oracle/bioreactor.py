"""Oracle restatement of the bioreactor state-transition increment and output map.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference model/BioreactorModel.py:
  * ``homeostatic_DEs`` :170-231  -> :func:`increment`   (one explicit-Euler increment, already
    multiplied by ``dt``; quirk Q1 of SURVEY.md)
  * ``static_outputs``  :233-253  -> :func:`outputs`

All arithmetic is float64 on whatever (float32) states are passed in; the reference's own
arithmetic is a float32/float64 mixture that depends on the numpy promotion rules in force
(SURVEY.md §7 "Mixed precision"), so parity with it is stated as a tolerance, not bit-exact.
"""
import numpy

# steady state used for x0 by sim_base.get_parts (sim_base.py:46-53):
#   Bioreactor.find_SS([0.06, 0.2], [260/180, 640/24.6, 1000/116, 0, 0])
# value produced by the reference in the build container (tests/golden/make_golden.py stores it too)
X_STEADY = numpy.array([1.5555555555555556, 26.016260162601625, 5.2711537990850506,
                        0.0, 15.188571428736536])


def increment(x, u, dt):
    """Vectorised ``homeostatic_DEs`` (BioreactorModel.py:170-231).

    x : (..., 5) array (Cg, Cx, Cfa, Ce, Ch); u : (2,) (Fg_in, Fm_in); dt : float.
    Returns the (..., 5) float64 increment (dCg, dCx, dCfa, dCe, dCh).
    """
    x = numpy.asarray(x, dtype=numpy.float64)
    u = numpy.asarray(u, dtype=numpy.float64)
    dt = float(dt)
    # :192-193 -- only the rate expressions see the clamp; Ch is not clamped (quirk Q2)
    Cg = numpy.maximum(x[..., 0], 0.0)
    Cx = numpy.maximum(x[..., 1], 0.0)
    Cfa = numpy.maximum(x[..., 2], 0.0)
    Ce = numpy.maximum(x[..., 3], 0.0)
    Ch = x[..., 4]

    Fg_in, Fm_in = u[0], u[1]          # :195
    Cg_in = 5000 / 180                 # :196
    F_out = Fg_in + Fm_in              # :197
    V = 1.0                            # :199

    rX = 0.0 * Cx                      # :201
    rH = 280 / 180 - Cg                # :202

    rFA_max = 0.25 / 116 * Cx * 24.6 * V          # :205
    sat = Cg / (1e-2 + Cg)
    rFA = rFA_max * sat                           # :206

    r_theta1_max = (0.4 - 0.25) / 180 * Cx * 24.6 * V                              # :209
    r_theta1_req = r_theta1_max - (r_theta1_max / 2000 / (0.28 / 180) * rH + 0.01 * Ch)  # :210
    r_theta1 = numpy.minimum(r_theta1_max, numpy.maximum(0.0, r_theta1_req)) * sat  # :211

    r_E_max = 0.025 / 46 * Cx * 24.6 * V          # :214
    rE_req = r_theta1_req - r_theta1_max          # :215
    rE = numpy.minimum(r_E_max, numpy.maximum(0.0, rE_req))     # :216

    r_theta2_max = (0.1 - 0.025) / 180 * Cx * 24.6 * V          # :219
    r_theta2_req = r_theta1_req - r_theta1_max - rE             # :220
    r_theta2 = numpy.minimum(r_theta2_max, numpy.maximum(0.0, r_theta2_req))  # :221

    rG = -rFA * (116 / 180) - r_theta1 - rE * (46 / 180) - r_theta2           # :223

    out = numpy.empty(x.shape, dtype=numpy.float64)
    out[..., 0] = (Fg_in * Cg_in - F_out * Cg + rG) / V * dt    # :225
    out[..., 1] = rX / V * dt                                   # :226
    out[..., 2] = (-F_out * Cfa + rFA) / V * dt                 # :227
    out[..., 3] = (-F_out * Ce + rE) / V * dt                   # :228
    out[..., 4] = rH / V * dt                                   # :229
    return out


def outputs(x, u=None, round32=True):
    """Vectorised ``static_outputs`` (BioreactorModel.py:233-253): (Cg*180, Cfa*116), unclamped.

    The reference multiplies a float32 state by a Python int, which yields a float32 under both
    the legacy and the NEP-50 promotion rules (and its GPU gufunc stores the outputs as f4,
    particle.py:193-203), so the outputs are float32-rounded products (``round32``); they are
    returned widened to float64 because everything downstream (``e = z - y``, the pdf) is float64.
    """
    x = numpy.asarray(x)
    out = numpy.empty(x.shape[:-1] + (2,), dtype=numpy.float64)
    out[..., 0] = x[..., 0].astype(numpy.float64) * 180
    out[..., 1] = x[..., 2].astype(numpy.float64) * 116
    if round32:
        out = out.astype(numpy.float32).astype(numpy.float64)
    return out


def increment_scalar(x, u, dt=1):
    """Per-particle form with the reference's call signature (used by the loop-faithful
    timing port in oracle/particle.py; same expressions as :func:`increment`)."""
    Cg, Cx, Cfa, Ce, Ch = x
    Cg, Cx, Cfa, Ce = max(Cg, 0), max(Cx, 0), max(Cfa, 0), max(Ce, 0)
    Fg_in, Fm_in = u
    Cg_in = 5000 / 180
    F_out = Fg_in + Fm_in
    rH = 280 / 180 - Cg
    sat = Cg / (1e-2 + Cg)
    rFA = 0.25 / 116 * Cx * 24.6 * sat
    t1max = (0.4 - 0.25) / 180 * Cx * 24.6
    t1req = t1max - (t1max / 2000 / (0.28 / 180) * rH + 0.01 * Ch)
    t1 = min(t1max, max(0, t1req)) * sat
    rE = min(0.025 / 46 * Cx * 24.6, max(0, t1req - t1max))
    t2 = min((0.1 - 0.025) / 180 * Cx * 24.6, max(0, t1req - t1max - rE))
    rG = -rFA * (116 / 180) - t1 - rE * (46 / 180) - t2
    return ((Fg_in * Cg_in - F_out * Cg + rG) * dt, 0.0 * Cx * dt, (-F_out * Cfa + rFA) * dt,
            (-F_out * Ce + rE) * dt, rH * dt)


def outputs_scalar(x, u=None):
    return x[0] * 180, x[2] * 116
