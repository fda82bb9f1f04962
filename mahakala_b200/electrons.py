"""Drop-in for ``mahakala.electrons`` (reference: /root/reference/mahakala/electrons.py:32-50)."""
from . import _cabi
from .constants import CL, ME, MP
from ._device import DeviceArray, as_device, empty, stream_ptr


def rlow_rhigh_model(dens, u, beta, r_low=1, r_high=40, electron_gamma=4. / 3, ion_gamma=5. / 3):
    """electrons.py:32-50: dimensionless electron temperature Theta_e from the R_low/R_high model."""
    d = as_device(dens)
    shape = d.shape
    out = empty(shape)
    _cabi.call("mk_rlow_rhigh", d, as_device(u), as_device(beta), d.numel(), float(r_low), float(r_high),
               float(electron_gamma), float(ion_gamma), CL, MP, ME, out, stream_ptr())
    return DeviceArray.wrap(out)
