"""A 12x12 image by the reference's own make_image (whole package under the stand-ins of make_reference_fluid_golden.py)
on the 32^3 synthetic snapshot the GPU tests use as their fixture (ncells=32, block=16, extent=16): 144 rays, tol 1e-4,
N = 10000, 230 GHz.  -> tests/golden/reference_image12_golden.npz   (about 8 minutes)

    python tests/golden/make_reference_image_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_fluid_golden as F  # noqa: E402

if __name__ == "__main__":
    F.install()
    from mahakala.grmhd.athenak import AthenakFluidModel
    from mahakala.images import make_image
    from mahakala_b200.synthetic import make_synthetic_snapshot
    snap = make_synthetic_snapshot(ncells=32, block=16, extent=16.0, seed=0)
    F.register("snap32.athdf", snap)
    model = AthenakFluidModel("snap32.athdf", 0.94, fluid_gamma=snap["fluid_gamma"])
    img = np.asarray(make_image(model, resolution=12))
    np.savez_compressed(os.path.join(HERE, "reference_image12_golden.npz"), image_res12=img)
    print("wrote image", img.shape, img.max(), img.sum())
