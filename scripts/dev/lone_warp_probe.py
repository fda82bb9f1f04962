"""Per-step latency of a LONE warp (nothing else on the GPU) against the per-step time at full occupancy: how much
of the longest photon-ring ray's 3765 dependent RK4 steps could a low-occupancy launch save?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import mahakala_b200 as ma
from mahakala_b200 import geodesics as geo, images
from mahakala_b200.grmhd import AthenakFluidModel
from mahakala_b200.synthetic import make_synthetic_snapshot
a = 0.94
s0 = ma.initialize_geodesics_at_camera(a, 60, 1000, -10, 10, 1024)
f, n, r = geo.integrate_final(10000, s0, 40, 1e-4, a)
n = n.cpu().numpy()
order = np.argsort(-n)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
arr = make_synthetic_snapshot(ncells=256, block=32, extent=32.0, seed=0, dtype=np.float32)
m = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                  arr["x3f"], arr["LogicalLocations"], arr["Levels"], a, fluid_gamma=arr["fluid_gamma"], storage="f64")
m.snapshot()
for label, idx in (("32 longest rays (one warp)", order[:32]), ("1 longest ray", order[:1]), ("128 longest (4 warps, one CTA)", order[:128]),
                   ("592 x 32 longest (one warp per SMSP)", order[:592 * 32]), ("32 median rays", order[len(order) // 2:len(order) // 2 + 32])):
    sub = s0[torch.from_numpy(idx.copy()).cuda()].contiguous()
    mx = int(n[idx].max())
    ti = timeit(lambda: geo.integrate_final(10000, sub, 40, 1e-4, a))
    tr = timeit(lambda: images.render(m, s0=sub))
    print(f"{label:40s} max steps {mx:5d}: integrate {ti:7.3f} ms = {1e3 * ti / mx:5.3f} us/step; render {tr:7.3f} ms = {1e3 * tr / mx:5.3f} us/step")
tf = timeit(lambda: geo.integrate_final(10000, s0, 40, 1e-4, a))
tr = timeit(lambda: images.render(m, resolution=1024))
wsteps = n.reshape(256, 4, 128, 8).max(axis=(1, 3)).sum()
print(f"full frame: integrate {tf:.2f} ms, render {tr:.2f} ms; render warp-steps {wsteps} over 2368 warps -> {1e3 * tr / (wsteps / 2368):.3f} us per warp-step at full occupancy")
