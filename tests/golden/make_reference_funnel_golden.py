"""A 10x10 image by the reference's own make_image on a snapshot WITH A MAGNETISED FUNNEL, so that the sigma > 100
cut of images.py:116-118 fires (the smooth torus of the other fixtures never exceeds sigma = 0.72).

Same stand-ins as make_reference_fluid_golden.py (whole reference package, NumPy-backed jax, in-memory h5py).  Snapshot:
make_synthetic_snapshot(ncells=32, block=16, extent=16, seed=0, funnel={}) -- the 32^3 GPU-test fixture plus the polar
funnel of mahakala_b200/synthetic.py.  Frozen in tests/golden/reference_funnel_golden.npz:

    image_res10        the reference's make_image(model, resolution=10)
    sigma_gt_100       number of trajectory samples with sigma > 100 according to the reference's own
                       geodesic_integrator + get_fluid_scalars_from_geodesics (images.py:84-92), and
    in_domain          the number of samples with dens > 0 on moving rows

    python tests/golden/make_reference_funnel_golden.py        (about 15 minutes of Python loops)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_fluid_golden as F  # noqa: E402

if __name__ == "__main__":
    F.install()
    import mahakala as ma
    from mahakala.grmhd.athenak import AthenakFluidModel
    from mahakala.images import make_image
    from mahakala_b200.synthetic import make_synthetic_snapshot
    snap = make_synthetic_snapshot(ncells=32, block=16, extent=16.0, seed=0, funnel={})
    F.register("funnel32.athdf", snap)
    model = AthenakFluidModel("funnel32.athdf", 0.94, fluid_gamma=snap["fluid_gamma"])
    res = 10
    img = np.asarray(make_image(model, resolution=res))
    # the sigma statistics of the same trajectories, by the reference's own functions (images.py:56-92)
    s0 = ma.initialize_geodesics_at_camera(0.94, 60, 1000, -10., 10., res)
    S, final_dt = ma.geodesic_integrator(10000, s0, 40, 1e-4, 0.94)
    fs = model.get_fluid_scalars_from_geodesics(S)
    dens, b = np.asarray(fs['dens']), np.asarray(fs['b'])
    moving = np.zeros(dens.shape, dtype=bool)
    moving[1:] = np.asarray(final_dt)[:-1] != 0           # row i is weighted by dt[i-1] (transfer.py:107)
    with np.errstate(all='ignore'):
        sigma = b * b / dens
    m = moving & (dens > 0)
    np.savez_compressed(os.path.join(HERE, "reference_funnel_golden.npz"), image_res10=img,
                        sigma_gt_100=int((sigma[m] > 100.).sum()), in_domain=int(m.sum()),
                        sigma_max=float(np.nanmax(sigma[m])))
    print("wrote funnel image", img.shape, img.max(), img.sum(), int((sigma[m] > 100.).sum()), int(m.sum()))
