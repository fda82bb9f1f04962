"""Short workload for ncu captures (see profiles/README.md): final-mode integrate at 1024^2, paged dump at
512^2, fused render at 1024^2 on a 256^3 snapshot."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mahakala_b200 as ma
from mahakala_b200 import geodesics as geo, images
from mahakala_b200.grmhd import AthenakFluidModel
from mahakala_b200.synthetic import make_synthetic_snapshot

what = sys.argv[1] if len(sys.argv) > 1 else "all"
a = 0.94
if what in ("all", "integrate"):
    s0 = ma.initialize_geodesics_at_camera(a, 60, 1000, -10, 10, 1024)
    for _ in range(2):
        geo.integrate_final(10000, s0, 40, 1e-4, a)
    torch.cuda.synchronize()
if what in ("all", "paged"):
    s0 = ma.initialize_geodesics_at_camera(a, 60, 1000, -10, 10, 512)
    store = geo.TrajectoryStore.allocate(s0.shape[0], 10000, mem_fraction=0.2)
    for _ in range(2):
        store.reset()
        geo.integrate_paged(10000, s0, 40, 1e-4, a, store=store)
    torch.cuda.synchronize()
    del store
if what == "long":
    # the warp-specialised long-patch kernel on the 592 longest patches of the 1024^2 cfg4 frame (learned order)
    arr = make_synthetic_snapshot(ncells=256, block=32, extent=32.0, seed=0)
    m = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                      arr["x3f"], arr["LogicalLocations"], arr["Levels"], a, fluid_gamma=arr["fluid_gamma"],
                                      storage="f64")
    images.learn_patch_order(a, resolution=1024)
    for _ in range(2):
        images.render(m, resolution=1024, observing_frequencies=(230e9,), long_patches=592)
    torch.cuda.synchronize()
if what in ("all", "render"):
    nc = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    arr = make_synthetic_snapshot(ncells=nc, block=32, extent=32.0, seed=0)
    m = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                      arr["x3f"], arr["LogicalLocations"], arr["Levels"], a, fluid_gamma=arr["fluid_gamma"],
                                      storage=os.environ.get("MK_RENDER_STORAGE", "f64"))     # as bench.py's render leg
    for _ in range(2):
        images.render(m, resolution=1024, observing_frequencies=(230e9,))
    torch.cuda.synchronize()
