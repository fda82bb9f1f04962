"""Experiment: lane-level refill in the fused kernel vs whole patches.  Build the library with
    scripts/build_variant.sh "-DMK_EXPERIMENTS"
then run with MK_RENDER_REFILL_THR=k (k idle lanes trigger a refill; unset = product path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mahakala_b200 import images
from mahakala_b200.grmhd import AthenakFluidModel
from mahakala_b200.synthetic import make_synthetic_snapshot
arr = make_synthetic_snapshot(ncells=256, block=32, extent=32.0, seed=0)
m = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                  arr["x3f"], arr["LogicalLocations"], arr["Levels"], 0.94, fluid_gamma=arr["fluid_gamma"])
def timeit(fn, n=4):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
img = images.render(m, resolution=1024)
print(os.environ.get("MK_RENDER_REFILL_THR", "patches"), "ms", timeit(lambda: images.render(m, resolution=1024)), "flux", float(img.sum()))
