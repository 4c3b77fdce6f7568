"""Drop-in for the reference's ``filter`` package (filter/__init__.py:1-7): the same four names.
The CPU and the "Parallel" class of each pair are the same B200 implementation -- there is no
CPU path in this build."""
from gpu_se_b200.filter.particle import ParticleFilter
from gpu_se_b200.filter.particle import ParallelParticleFilter
from gpu_se_b200.filter.gs_ukf import GaussianSumUnscentedKalmanFilter
from gpu_se_b200.filter.gs_ukf import ParallelGaussianSumUnscentedKalmanFilter
from gpu_se_b200.filter.resample import resample_from_cumsum

__all__ = ['ParticleFilter', 'ParallelParticleFilter',
           'GaussianSumUnscentedKalmanFilter', 'ParallelGaussianSumUnscentedKalmanFilter',
           'resample_from_cumsum']
