"""N = 1: does handing out the long rays first (multigpu.longest_first_ray_order) shorten the tail of the cfg2 launch?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mahakala_b200 as ma
from mahakala_b200 import geodesics as geo, multigpu
a = 0.94
s0 = ma.initialize_geodesics_at_camera(a, 60, 1000, -10, 10, 1024)
store = geo.TrajectoryStore.allocate(s0.shape[0], 10000, mem_fraction=0.5)
order = torch.from_numpy(multigpu.longest_first_ray_order(1024, 1)).cuda()
queue = torch.zeros(64, dtype=torch.int32, device="cuda")
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        queue.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)
print("pixel order   paged ms (min, mean)", timeit(lambda: geo.integrate_paged(10000, s0, 40, 1e-4, a, store=store)))
ref = (store.final.clone(), store.nsteps.clone())
print("longest first paged ms (min, mean)", timeit(lambda: geo.integrate_paged(10000, s0, 40, 1e-4, a, store=store, queue=queue, ray_order=order)))
print("identical results:", torch.equal(ref[0], store.final) and torch.equal(ref[1], store.nsteps))
