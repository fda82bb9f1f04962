from .grmhd import GRMHDFluidModel
from .athenak import AnalyticTorusFluidModel, AthenakFluidModel, build_block_grid, fill_ghost_zones

__all__ = ["GRMHDFluidModel", "AthenakFluidModel", "AnalyticTorusFluidModel", "build_block_grid", "fill_ghost_zones"]
