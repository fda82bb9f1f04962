"""mahakala_b200 — B200-native drop-in for the per-ray hot path of Mahakala.

Mirrors the export list of /root/reference/mahakala/__init__.py:25-45 (the reference's API contract);
float64 throughout (the reference forces ``jax_enable_x64`` at import, __init__.py:22-23).  Compute runs
in hand-written sm_100a CUDA kernels behind the C ABI of ``include/mahakala_b200.h``; there is no JAX,
no Triton and no CPU fallback on that path.

    import mahakala_b200 as ma            # instead of: import mahakala as ma
    from mahakala_b200.images import make_image
    from mahakala_b200.grmhd import AthenakFluidModel
"""
from . import constants
from . import geodesics
from . import electrons
from . import transfer
from . import images
from . import grmhd

from .geodesics import find_shadow_bisection
from .geodesics import find_shadow_bisection_angles
from .geodesics import geodesic_integrator
from .geodesics import initialize_geodesics_at_camera

from .transfer import synchrotron_coefficients
from .transfer import solve_specific_intensity
from .transfer import solve_attenuated_emissivity

__all__ = [
    "find_shadow_bisection",
    "find_shadow_bisection_angles",
    "geodesic_integrator",
    "initialize_geodesics_at_camera",
    "synchrotron_coefficients",
    "solve_specific_intensity",
    "solve_attenuated_emissivity",
]


def install_as_mahakala():
    """Register this package under the name ``mahakala`` so that existing scripts
    (``import mahakala as ma``, ``from mahakala.images import make_image``) run unchanged."""
    import sys
    me = sys.modules[__name__]
    sys.modules.setdefault("mahakala", me)
    for sub in ("constants", "geodesics", "electrons", "transfer", "images", "grmhd"):
        sys.modules.setdefault(f"mahakala.{sub}", getattr(me, sub))
    sys.modules.setdefault("mahakala.grmhd.athenak", grmhd.athenak)
    sys.modules.setdefault("mahakala.grmhd.grmhd", grmhd.grmhd)
    return me
