"""CPU oracle for the gpu_se state-estimation hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy (float64 arithmetic on the float32 arrays the
reference stores), the algorithms of the reference's CPU classes:

* ``oracle.bioreactor``  <- model/BioreactorModel.py:170-253
* ``oracle.mixture``     <- gaussian_sum_dist/MultivariateGaussianSum.py:27-97
* ``oracle.particle``    <- filter/particle.py:43-114
* ``oracle.gs_ukf``      <- filter/gs_ukf.py:45-183
* ``oracle.philox``      <- specification of THIS repo's counter-based sampler
                            (Philox4x32-10, Salmon et al. SC'11; Random123 known-answer vectors)

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the reference's own CPU
classes from ``/root/reference`` (with an empty ``cupy`` stub) in the build container, runs them
on seeded inputs and stores inputs + outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every oracle function against those vectors, and
``tests/test_oracle_vs_reference.py`` re-runs the comparison live whenever ``/root/reference``
is present.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the timed CPU
baseline.  Nothing under ``gpu_se_b200/`` imports it: the product path has no CPU fallback.
"""
