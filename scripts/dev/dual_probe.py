import os, sys
sys.path.insert(0, "/root/repo")
import torch
import mahakala_b200 as ma
from mahakala_b200 import geodesics as geo
a = 0.94
s0 = ma.initialize_geodesics_at_camera(a, 60, 1000, -10, 10, 512)
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for name in ("kerr_schild", "kerr_schild_dual"):
    geo.set_metric(name)
    t = timeit(lambda: geo.integrate_final(10000, s0, 40, 1e-4, a))
    f, n, rl, tot = geo.integrate_final(10000, s0, 40, 1e-4, a, want_total=True)
    print(name, "512^2 final ms", t, "G ray-steps/s", int(tot) / t / 1e6)
sys.path.insert(0, "/root/repo/tests")
from test_geodesics_gpu import KERR_SCHILD_USER
geo.register_metric("ks_user", KERR_SCHILD_USER); geo.set_metric("ks_user")
t = timeit(lambda: geo.integrate_final(10000, s0, 40, 1e-4, a))
print("ks_user (NVRTC)", "512^2 final ms", t)
