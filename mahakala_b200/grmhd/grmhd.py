"""Fluid-model base class (reference: /root/reference/mahakala/grmhd/grmhd.py:35-55)."""
import numpy as np

from ..constants import CL, GNEWT, ME, MP


class GRMHDFluidModel:
    """Duck type consumed by ``images.make_image``: attributes ``bhspin`` and ``fluid_gamma``, methods
    ``get_fluid_scalars_from_geodesics(S)``, ``get_prims_from_geodesics(S)`` and ``get_units``."""

    def __init__(self):
        pass

    def get_units(self, M_BH, mass_scale):
        """grmhd.py:40-55: code -> cgs unit factors (host scalars)."""
        L_unit = GNEWT * M_BH / CL**2
        T_unit = L_unit / CL
        dens_unit = mass_scale / L_unit**3
        Ne_unit = dens_unit / (MP + ME)
        B_unit = CL * np.sqrt(4. * np.pi * dens_unit)
        return dict(L_unit=L_unit, T_unit=T_unit, dens_unit=dens_unit, Ne_unit=Ne_unit, B_unit=B_unit)
