from .grmhd import GRMHDFluidModel
from .athenak import AthenakFluidModel

__all__ = ["GRMHDFluidModel", "AthenakFluidModel"]
