from gpu_se_b200.model.BioreactorModel import Bioreactor

__all__ = ["Bioreactor"]
