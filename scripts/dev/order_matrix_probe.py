"""N = 1, cfg2 paged dump: local queue / shared-queue variant x pixel order / centre-out order, with an L2 flush between
launches, several repetitions, interleaved so that box drift shows up in every row alike."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import mahakala_b200 as ma
from mahakala_b200 import geodesics as geo, multigpu
a = 0.94
res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
s0 = ma.initialize_geodesics_at_camera(a, 60, 1000, -10, 10, res)
store = geo.TrajectoryStore.allocate(s0.shape[0], 10000, mem_fraction=0.5)
order = torch.from_numpy(multigpu.longest_first_ray_order(res, 1)).cuda()
ident = torch.arange(s0.shape[0], dtype=torch.int32, device="cuda")
queue = torch.zeros(64, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
variants = {
    "local  pixel     ": lambda: geo.integrate_paged(10000, s0, 40, 1e-4, a, store=store),
    "shared pixel     ": lambda: geo.integrate_paged(10000, s0, 40, 1e-4, a, store=store, queue=queue, ray_order=ident),
    "shared centre-out": lambda: geo.integrate_paged(10000, s0, 40, 1e-4, a, store=store, queue=queue, ray_order=order),
}
times = {k: [] for k in variants}
for k, fn in variants.items():
    queue.zero_(); fn()
torch.cuda.synchronize()
for r in range(reps):
    for k, fn in variants.items():
        flush.fill_(r); queue.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        times[k].append(e0.elapsed_time(e1))
for k, t in times.items():
    print(k, "ms min %.3f median %.3f max %.3f" % (min(t), float(np.median(t)), max(t)))
