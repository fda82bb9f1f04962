"""Drop-in for ``mahakala.transfer`` (reference: /root/reference/mahakala/transfer.py:30-144)."""
import ctypes

import torch

from . import _cabi
from .constants import CL, EE, HPL, ME, MP
from ._device import DeviceArray, as_device, empty, stream_ptr


class EmissionParams(ctypes.Structure):
    """Mirror of ``mk_emission_params`` (include/mahakala_b200.h)."""
    _fields_ = [(k, ctypes.c_double) for k in (
        "fluid_gamma", "r_low", "r_high", "electron_gamma", "ion_gamma", "Ne_unit", "B_unit", "L_unit",
        "sigma_cut", "EE", "CL", "ME", "MP", "HPL", "two_11_12")]


def emission_params(fluid_gamma=4. / 3, r_low=1, r_high=40, electron_gamma=4. / 3, ion_gamma=5. / 3,
                    Ne_unit=1., B_unit=1., L_unit=1., sigma_cut=100.):
    return EmissionParams(float(fluid_gamma), float(r_low), float(r_high), float(electron_gamma),
                          float(ion_gamma), float(Ne_unit), float(B_unit), float(L_unit), float(sigma_cut),
                          EE, CL, ME, MP, HPL, 2.0**(11. / 12))


def _broadcast_device(*xs):
    ts = [as_device(x) if not isinstance(x, (int, float)) else x for x in xs]
    shape = torch.broadcast_shapes(*[t.shape for t in ts if isinstance(t, torch.Tensor)])
    dev = next(t.device for t in ts if isinstance(t, torch.Tensor))
    out = []
    for t in ts:
        if not isinstance(t, torch.Tensor):
            t = torch.full(shape, float(t), dtype=torch.float64, device=dev)
        out.append(t.expand(shape).contiguous())
    return out, shape


def synchrotron_coefficients(Ne, Theta_e, B, pitch_angle, nu, invariant=True, rescale_nu=1.):
    """transfer.py:30-86: thermal synchrotron emissivity and absorptivity (cgs; invariant by default)."""
    (ne, th, b, pa, nu_), shape = _broadcast_device(Ne, Theta_e, B, pitch_angle, nu)
    em = empty(shape)
    ab = empty(shape)
    P = emission_params()
    _cabi.call("mk_synchrotron", P, ne, th, b, pa, nu_, em.numel(), 1 if invariant else 0, float(rescale_nu),
               em, ab, stream_ptr())
    return DeviceArray.wrap(em), DeviceArray.wrap(ab)


def solve_specific_intensity(emissivity, absorptivity, dt, L_unit, dIs=False):
    """transfer.py:89-119: back-to-front explicit-Euler transfer; returns I_nu (npx,) [and dI per step]."""
    em = as_device(emissivity)
    ab = as_device(absorptivity)
    d = as_device(dt)
    nsteps, npx = em.shape
    I = empty((npx,))
    dI = empty((max(nsteps - 1, 0), npx)) if dIs else None
    _cabi.call("mk_solve_specific_intensity", em, ab, d, nsteps, npx, float(L_unit), I, dI, stream_ptr())
    if dIs:
        return DeviceArray.wrap(I), DeviceArray.wrap(dI)
    return DeviceArray.wrap(I)


def solve_attenuated_emissivity(emissivity, absorptivity, dt, L_unit):
    """transfer.py:122-144: attenuated (observed) emissivity contribution of every step."""
    em = as_device(emissivity)
    ab = as_device(absorptivity)
    d = as_device(dt)
    nsteps, npx = em.shape
    out = empty((max(nsteps - 1, 0), npx))
    _cabi.call("mk_solve_attenuated_emissivity", em, ab, d, nsteps, npx, float(L_unit), out, stream_ptr())
    return DeviceArray.wrap(out)


def emission_probe(S, prims, bhspin, params, observing_frequencies, fast=True):
    """Invariant (j, alpha) of arbitrary (state, primitives) pairs through the emission code of the fused render
    kernel (``fast=True``: ``emission_fast<NF>``) or through the literal IEEE chain of images.py:87-118 +
    athenak.py:760-794 + transfer.py:56-86 (``fast=False``).  ``S`` (n, 8), ``prims`` (n, 8) in canonical order
    dens, eint, U1..3, B1..3; ``params`` from ``emission_params``.  Returns em, ab of shape (nfreq, n).  A test and
    diagnosis hook: the special cases of the reference (sigma cut, Theta_e floor, X limit, NaN -> 0) can be put in
    front of the kernel code directly instead of hoping a snapshot contains them."""
    import numpy as np
    s = as_device(S).contiguous()
    p = as_device(prims).contiguous()
    nus = np.atleast_1d(np.asarray(observing_frequencies, dtype=np.float64))
    n = s.shape[0]
    em = empty((nus.size, n))
    ab = empty((nus.size, n))
    c_nu = (ctypes.c_double * nus.size)(*nus)
    _cabi.call("mk_emission_probe", params, float(bhspin), s, p, n, int(nus.size), c_nu, 1 if fast else 0, em, ab,
               stream_ptr())
    return DeviceArray.wrap(em), DeviceArray.wrap(ab)
