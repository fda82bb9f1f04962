"""torchrun --nproc-per-node N scripts/dev/e2e_mix_probe.py: where does the host-to-host (zero-copy) step lose time when N
processes share the host?  cfg2 paged dump per rank with s0 read from {device, pinned host} and the per-ray results
written to {device, pinned host}."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist
import mahakala_b200 as ma
from mahakala_b200 import _cabi, geodesics as geo
from mahakala_b200._device import stream_ptr
local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if world > 1 else 0
a = 0.94
incl = [60.0, 17.0, 30.0, 80.0, 45.0, 70.0, 25.0, 52.0][rank % 8]
s0 = ma.initialize_geodesics_at_camera(a, incl, 1000, -10, 10, 1024)
npx = s0.shape[0]
s0h = torch.empty((npx, 8), dtype=torch.float64, pin_memory=True); s0h.copy_(s0)
store = geo.TrajectoryStore.allocate(npx, 10000, mem_fraction=0.6)
host = {"final": torch.empty((npx, 8), dtype=torch.float64, pin_memory=True), "nsteps": torch.empty((npx,), dtype=torch.int32, pin_memory=True),
        "r_last": torch.empty((npx,), dtype=torch.float64, pin_memory=True)}
devo = {"final": store.final, "nsteps": store.nsteps, "r_last": store.r_last}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def launch(src, out):
    store.reset()
    _cabi.call("mk_integrate_paged", 0, a, 10000, npx, src, 40.0, 1e-4, out["final"], out["nsteps"], out["r_last"], store.pages,
               store.page_next, store.page_first, store.ctrl[0:1], store.max_pages, store.ctrl[1:2], store.total_steps, stream_ptr())
for name, src, out in (("device -> device", s0, devo), ("host   -> device", s0h, devo), ("device -> host  ", s0, host), ("host   -> host  ", s0h, host)):
    ts = []
    for it in range(5):
        flush.fill_(it)
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(src, out); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it: ts.append(float(t))
    if rank == 0:
        print(f"[{world} GPUs] s0 / results {name}: min {min(ts):.2f} ms mean {np.mean(ts):.2f} ms (max over ranks)", flush=True)
# the public host-to-host call by wall clock, as bench.py's e2e leg times it, split into its parts (per rank)
import time
parts = {"call": [], "sum": []}
for it in range(8):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    geo.integrate_paged_host(10000, s0h, 40, 1e-4, a, store, host)
    t1 = time.perf_counter()
    n = int(host["nsteps"].sum())
    t2 = time.perf_counter()
    if it >= 2:
        parts["call"].append(1e3 * (t1 - t0)); parts["sum"].append(1e3 * (t2 - t1))
mine = torch.tensor([np.mean(parts["call"]), np.mean(parts["sum"])], dtype=torch.float64, device="cuda")
allr = [torch.zeros_like(mine) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, mine)
else:
    allr = [mine]
if rank == 0:
    print(f"[{world} GPUs] integrate_paged_host wall ms per rank: {[round(float(t[0]), 2) for t in allr]}; nsteps.sum() ms per rank: {[round(float(t[1]), 2) for t in allr]}", flush=True)
# back to back without barriers (the bench loop)
if world > 1: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(10):
    geo.integrate_paged_host(10000, s0h, 40, 1e-4, a, store, host)
    n = int(host["nsteps"].sum())
torch.cuda.synchronize()
tt = torch.tensor([1e2 * (time.perf_counter() - t0)], dtype=torch.float64, device="cuda")
allr = [torch.zeros_like(tt) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, tt)
else:
    allr = [tt]
if rank == 0:
    print(f"[{world} GPUs] 10 back-to-back steps, ms per step per rank: {[round(float(t[0]), 2) for t in allr]}", flush=True)
if world > 1:
    dist.destroy_process_group()
