"""Freeze a few oracle outputs as golden fixtures (tests/golden/oracle_fixtures.npz).

The reference (JAX) cannot be imported in this image, so these vectors come from the audited CPU restatement
(oracle/), NOT from the reference itself; they guard the oracle against accidental drift and give the GPU
tests a fixed target that does not depend on the oracle being rebuilt.  Regenerate with
    python tests/golden/make_oracle_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def compute():
    from helpers import M_BH, MASS_SCALE, oracle_model, snapshot_arrays
    from oracle import c_oracle, mahakala_oracle as onp
    a = 0.94
    out = {}
    s0 = onp.initialize_geodesics_at_camera(a, 60, 1000, -10, 10, 12)
    o = c_oracle.integrate(2000, s0, 40, 1e-2, a)
    out["geo_s0"] = s0
    out["geo_final"] = o["final"]
    out["geo_nsteps"] = o["nsteps"]
    out["geo_r_last"] = o["r_last"]
    arr = snapshot_arrays(ncells=32, block=16, extent=16.0)
    om = oracle_model(arr, a)
    units = om.get_units(M_BH, MASS_SCALE)
    img, nsteps, nin = c_oracle.render(om, s0, units, [230e9, 345e9])
    out["img_230_345"] = img
    out["img_in_domain_samples"] = np.int64(nin)
    tor = onp.AnalyticTorusFluidModel(a)
    timg, _, _ = c_oracle.render(tor, s0, tor.get_units(M_BH, MASS_SCALE), [230e9])
    out["torus_img_230"] = timg
    rng = np.random.default_rng(11)
    Ne = np.exp(rng.normal(10, 2, 64)); Th = np.exp(rng.normal(1, 1.5, 64)); B = np.exp(rng.normal(1, 1, 64))
    pitch = rng.uniform(0, np.pi, 64); nu = 230e9 * np.exp(rng.normal(0, 0.5, 64))
    em, ab = onp.synchrotron_coefficients(Ne, Th, B, pitch, nu, invariant=True, rescale_nu=1 / 230e9)
    out.update(syn_Ne=Ne, syn_Th=Th, syn_B=B, syn_pitch=pitch, syn_nu=nu, syn_em=em, syn_ab=ab)
    return out


if __name__ == "__main__":
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_fixtures.npz")
    np.savez(dst, **compute())
    print("wrote", dst)
