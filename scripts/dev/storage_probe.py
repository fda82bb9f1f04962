"""f32 vs f64 cell storage at 512^3 (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mahakala_b200 import images
from mahakala_b200.grmhd import AthenakFluidModel
from mahakala_b200.synthetic import make_synthetic_snapshot
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 512
res = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
arr = make_synthetic_snapshot(ncells=nc, block=32, extent=32.0, seed=0)
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
base = None
for st in ("f32", "f64"):
    m = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                      arr["x3f"], arr["LogicalLocations"], arr["Levels"], 0.94, fluid_gamma=arr["fluid_gamma"], storage=st)
    m.snapshot()
    for nus in ((230e9,), (43e9, 86e9, 130e9, 230e9, 345e9, 460e9, 690e9, 870e9)):
        t = timeit(lambda: images.render(m, resolution=res, observing_frequencies=nus, mass_scale=2e24))
        print(f"{nc}^3 {st} cells, {res}^2, {len(nus)} freq: {t:.2f} ms")
    m.release()
