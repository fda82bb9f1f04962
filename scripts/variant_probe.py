"""Time integrate (final / paged) and render with whatever library is currently built (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mahakala_b200 as ma
from mahakala_b200 import geodesics as geo, images
from mahakala_b200.grmhd import AthenakFluidModel
from mahakala_b200.synthetic import make_synthetic_snapshot
a = 0.94
tag = sys.argv[1] if len(sys.argv) > 1 else ""
s0 = ma.initialize_geodesics_at_camera(a, 60, 1000, -10, 10, 1024)
def timeit(fn, n=4):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
if "norender" not in tag:
    nc = 256
    arr = make_synthetic_snapshot(ncells=nc, block=32, extent=32.0, seed=0)
    m = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                      arr["x3f"], arr["LogicalLocations"], arr["Levels"], a, fluid_gamma=arr["fluid_gamma"])
    import time
    for gf in ("device", "host"):
        t0 = time.time()
        mt = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                           arr["x3f"], arr["LogicalLocations"], arr["Levels"], a, fluid_gamma=arr["fluid_gamma"],
                                           ghost_fill=gf)
        mt.snapshot(); torch.cuda.synchronize()
        print(tag, f"snapshot setup ({nc}^3, ghost_fill={gf}) s", round(time.time() - t0, 3), mt.storage)
        mt.release()
    m.snapshot()
    torch.save({k: v for k, v in arr.items() if hasattr(v, "shape")}, "/tmp/snap.pt") if False else None
    print(tag, "render 1f ms", timeit(lambda: images.render(m, resolution=1024)))
    import time as _t
    for it in range(5):
        t0 = _t.perf_counter(); h = images.make_image(m, resolution=1024); t1 = _t.perf_counter()
        print(tag, "make_image host-to-host ms", round(1e3 * (t1 - t0), 2))
    m64 = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                        arr["x3f"], arr["LogicalLocations"], arr["Levels"], a, fluid_gamma=arr["fluid_gamma"], storage="f64")
    print(tag, "render 1f f64-cells ms", timeit(lambda: images.render(m64, resolution=1024)))
    m64.release()
    print(tag, "render 8f ms", timeit(lambda: images.render(m, resolution=1024, observing_frequencies=[43e9, 86e9, 130e9, 230e9, 345e9, 460e9, 690e9, 870e9])))
print(tag, "final ms", timeit(lambda: geo.integrate_final(10000, s0, 40, 1e-4, a)))
store = geo.TrajectoryStore.allocate(s0.shape[0], 10000, mem_fraction=0.4)
def paged():
    store.reset(); geo.integrate_paged(10000, s0, 40, 1e-4, a, store=store)
print(tag, "paged ms", timeit(paged))
