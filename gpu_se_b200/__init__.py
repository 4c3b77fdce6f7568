"""gpu_se_b200 -- the B200 (sm_100a) state-estimation hot path of AlgorithmicAmoeba/gpu_se.

``filter``, ``gaussian_sum_dist`` and ``model`` mirror the reference's packages of the same names
for the classes on the hot path.  Importing this package loads libgse_b200.so and fails loudly if
it is missing; nothing here falls back to a CPU implementation.
"""
from gpu_se_b200 import _lib  # noqa: F401  (load the shared library first: fail early and loudly)
from gpu_se_b200 import filter, gaussian_sum_dist, model  # noqa: A004
from gpu_se_b200.filter import (GaussianSumUnscentedKalmanFilter, ParallelGaussianSumUnscentedKalmanFilter,
                                ParallelParticleFilter, ParticleFilter)
from gpu_se_b200.gaussian_sum_dist import DeterministicGaussianSum, MultivariateGaussianSum
from gpu_se_b200.model import Bioreactor

__all__ = ["filter", "gaussian_sum_dist", "model", "ParticleFilter", "ParallelParticleFilter",
           "GaussianSumUnscentedKalmanFilter", "ParallelGaussianSumUnscentedKalmanFilter",
           "MultivariateGaussianSum", "DeterministicGaussianSum", "Bioreactor"]
__version__ = "0.1.0"
