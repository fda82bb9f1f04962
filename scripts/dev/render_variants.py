"""Time the fused render with every library build under mahakala_b200/variants/ (scripts/build_variant.sh)."""
import glob, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
child = r'''
import sys, torch
sys.path.insert(0, %r)
from mahakala_b200 import images
from mahakala_b200.grmhd import AthenakFluidModel
from mahakala_b200.synthetic import make_synthetic_snapshot
arr = make_synthetic_snapshot(ncells=256, block=32, extent=32.0, seed=0)
def model(**kw):
    return AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                         arr["x3f"], arr["LogicalLocations"], arr["Levels"], 0.94, fluid_gamma=arr["fluid_gamma"], **kw)
def timeit(fn, n=4):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
m = model(); m64 = model(storage="f64")
print("%%-8s image sums f32 cells %%.15e  f64 cells %%.15e" %% (sys.argv[1], float(images.render(m, resolution=1024).sum()), float(images.render(m64, resolution=1024).sum())), flush=True)
F8 = [43e9, 86e9, 130e9, 230e9, 345e9, 460e9, 690e9, 870e9]
print("%%-8s 1f %%.2f  1f-f64cells %%.2f  2f %%.2f  4f %%.2f  8f %%.2f ms" %% (sys.argv[1], timeit(lambda: images.render(m, resolution=1024)),
      timeit(lambda: images.render(m64, resolution=1024)), timeit(lambda: images.render(m, resolution=1024, observing_frequencies=F8[2:4])),
      timeit(lambda: images.render(m, resolution=1024, observing_frequencies=F8[2:6])),
      timeit(lambda: images.render(m, resolution=1024, observing_frequencies=F8))), flush=True)
m.release(); m64.release()
import mahakala_b200 as ma
from mahakala_b200 import geodesics as geo
s0 = ma.initialize_geodesics_at_camera(0.94, 60, 1000, -10, 10, 1024)
store = geo.TrajectoryStore.allocate(s0.shape[0], 10000, mem_fraction=0.5)
print("%%-8s integrate final %%.2f  paged dump %%.2f ms" %% (sys.argv[1], timeit(lambda: geo.integrate_final(10000, s0, 40, 1e-4, 0.94)),
      timeit(lambda: geo.integrate_paged(10000, s0, 40, 1e-4, 0.94, store=store))), flush=True)
''' % root
libs = [("default", None)] + [(os.path.basename(p)[3:-3], p) for p in sorted(glob.glob(os.path.join(root, "mahakala_b200/variants/lib*.so")))]
for name, path in libs:
    env = dict(os.environ)
    if path:
        env["MAHAKALA_B200_LIB"] = path
    subprocess.run([sys.executable, "-c", child, name], env=env)
