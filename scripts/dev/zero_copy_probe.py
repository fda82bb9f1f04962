"""Does the integrate kernel run with its inputs / per-ray outputs in pinned HOST memory (zero-copy over PCIe)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mahakala_b200 as ma
from mahakala_b200 import _cabi, geodesics as geo
a = 0.94
res = 1024
s0 = ma.initialize_geodesics_at_camera(a, 60, 1000, -10, 10, res)
npx = s0.shape[0]
s0_host = torch.empty((npx, 8), dtype=torch.float64, pin_memory=True); s0_host.copy_(s0)
out = {"final": torch.empty((npx, 8), dtype=torch.float64, pin_memory=True),
       "nsteps": torch.empty((npx,), dtype=torch.int32, pin_memory=True),
       "r_last": torch.empty((npx,), dtype=torch.float64, pin_memory=True)}
store = geo.TrajectoryStore.allocate(npx, 10000)
def zero_copy():
    store.reset()
    _cabi.call("mk_integrate_paged", 0, a, 10000, npx, s0_host, 40.0, 1e-4, out["final"], out["nsteps"], out["r_last"],
               store.pages, store.page_next, store.page_first, store.ctrl[0:1], store.max_pages, store.ctrl[1:2],
               store.total_steps, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return int(out["nsteps"].sum())
def streamed(ch):
    store.reset()
    geo.integrate_paged_streamed(10000, s0_host, 40, 1e-4, a, store, out, chunks=ch)
    return int(out["nsteps"].sum())
ref = streamed(4)
fin_ref = out["final"].clone()
for name, fn in (("zero-copy", zero_copy), ("streamed x4", lambda: streamed(4)), ("streamed x1", lambda: streamed(1))):
    fn()
    ts = []
    for _ in range(5):
        out["final"].zero_()
        t0 = time.perf_counter(); n = fn(); ts.append(1e3 * (time.perf_counter() - t0))
    print(name, "ms", [round(t, 2) for t in ts], "steps ok", n == ref, "final identical", bool(torch.equal(out["final"], fin_ref)))
