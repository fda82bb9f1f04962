"""Latency of the longest patches through the fused kernel against the warp-specialised long-patch kernel
(mk_render_long), alone on the GPU, and a whole 1024^2 frame with a learned order with / without the long-patch launch."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import mahakala_b200 as ma
from mahakala_b200 import _cabi, geodesics as geo, images
from mahakala_b200._device import stream_ptr
from mahakala_b200.constants import Msun
from mahakala_b200.grmhd import AthenakFluidModel
from mahakala_b200.synthetic import make_synthetic_snapshot
a = 0.94
res = 1024
s0 = ma.initialize_geodesics_at_camera(a, 60, 1000, -10, 10, res)
f, n, r = geo.integrate_final(10000, s0, 40, 1e-4, a)
n = n.cpu().numpy()
order = np.argsort(-n)
arr = make_synthetic_snapshot(ncells=256, block=32, extent=32.0, seed=0, dtype=np.float32)
m = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                  arr["x3f"], arr["LogicalLocations"], arr["Levels"], a, fluid_gamma=arr["fluid_gamma"], storage="f64")
snap = m.snapshot()
P, _ = images._params_for(m, 6.2e9 * Msun, 1e26, 40)
nu = (ctypes.c_double * 8)(*([230e9] * 8))
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
def long_only(sub, img):
    _cabi.call("mk_render_long", a, float(np.cos(np.pi / 3)), float(np.sin(np.pi / 3)), 1000.0, -10.0, 10.0, 0, sub, sub.shape[0],
               10000, 40.0, 1e-4, snap, P, 1, nu, img, None, None, None, None, 0, -1, 1, None, int(os.environ.get('MK_LONG_EXCLUSIVE', '1')), 0, stream_ptr())
for label, idx in (("32 longest rays (one patch)", order[:32]), ("1 longest ray", order[:1]), ("148 x 32 longest", order[:148 * 32]),
                   ("592 x 32 longest", order[:592 * 32]), ("32 median rays", order[len(order) // 2:len(order) // 2 + 32])):
    sub = s0[torch.from_numpy(idx.copy()).cuda()].contiguous()
    mx = int(n[idx].max())
    ref = images.render(m, s0=sub)
    img = torch.empty_like(ref)
    long_only(sub, img); torch.cuda.synchronize()
    tf = timeit(lambda: images.render(m, s0=sub))
    tl = timeit(lambda: long_only(sub, img))
    print(f"{label:32s} max steps {mx:5d}: fused {tf:7.3f} ms = {1e3 * tf / mx:5.3f} us/step; pipeline {tl:7.3f} ms = {1e3 * tl / mx:5.3f} us/step; identical {torch.equal(img, ref)}")
plain = images.render(m, resolution=res)
t_plain = timeit(lambda: images.render(m, resolution=res))
images.learn_patch_order(a, resolution=res)
key = next(iter(images._learned_lengths))
L = images._learned_lengths[key]
print("patch lengths: max", int(L[0]), "counts >= 0.5/0.4/0.3/0.2 of max:", [int((L >= t * L[0]).sum()) for t in (0.5, 0.4, 0.3, 0.2)], "of", len(L))
t_learn = timeit(lambda: images.render(m, resolution=res, long_patches=0))
print(f"whole frame: centre-out {t_plain:.2f} ms; learned order, fused only {t_learn:.2f} ms")
for nl in (64, 148, 296, 592, 1184):
    img = images.render(m, resolution=res, long_patches=nl)
    t = timeit(lambda: images.render(m, resolution=res, long_patches=nl))
    print(f"  learned order + {nl:5d} long patches: {t:.2f} ms; identical {torch.equal(img, plain)}")
