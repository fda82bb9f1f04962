"""Shared builders for the fluid-model tests (oracle model and device model from the same arrays)."""
import numpy as np

from mahakala_b200.synthetic import make_synthetic_snapshot

M_BH = 6.2e9 * 1.989e33
MASS_SCALE = 1.e26


def snapshot_arrays(ncells=32, block=16, extent=16.0, seed=0, **kw):
    return make_synthetic_snapshot(ncells=ncells, block=block, extent=extent, seed=seed, **kw)


def oracle_model(arr, bhspin):
    from oracle import mahakala_oracle as onp
    return onp.AthenakFluidModel(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                 arr["x3f"], arr["LogicalLocations"], arr["Levels"], bhspin,
                                 fluid_gamma=arr["fluid_gamma"], variable_names=arr["VariableNames"])


def device_model(arr, bhspin, **kw):
    from mahakala_b200.grmhd import AthenakFluidModel
    return AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"],
                                         arr["x2f"], arr["x3f"], arr["LogicalLocations"], arr["Levels"], bhspin,
                                         fluid_gamma=arr["fluid_gamma"], VariableNames=arr["VariableNames"], **kw)


def close(a, b, rtol, atol):
    a = np.asarray(a); b = np.asarray(b)
    return np.all(np.abs(a - b) <= atol + rtol * np.abs(b))
