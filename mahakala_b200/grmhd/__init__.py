from .grmhd import GRMHDFluidModel
from .athenak import AnalyticTorusFluidModel, AthenakFluidModel

__all__ = ["GRMHDFluidModel", "AthenakFluidModel", "AnalyticTorusFluidModel"]
