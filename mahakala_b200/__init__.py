"""mahakala_b200 — B200-native drop-in for the per-ray hot path of Mahakala.

Mirrors the export list of /root/reference/mahakala/__init__.py:25-45 (the reference's API contract);
float64 throughout (the reference forces ``jax_enable_x64`` at import, __init__.py:22-23).  Compute runs
in hand-written sm_100a CUDA kernels behind the C ABI of ``include/mahakala_b200.h``; there is no JAX,
no Triton and no CPU fallback on that path.
"""
from . import constants
from . import geodesics
from .geodesics import find_shadow_bisection
from .geodesics import find_shadow_bisection_angles
from .geodesics import geodesic_integrator
from .geodesics import initialize_geodesics_at_camera

__all__ = [
    "find_shadow_bisection",
    "find_shadow_bisection_angles",
    "geodesic_integrator",
    "initialize_geodesics_at_camera",
]
