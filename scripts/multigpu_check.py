"""torchrun --nproc-per-node N scripts/multigpu_check.py [res] [ncells]

Checks that the multi-GPU paths reproduce the single-GPU results bit for bit and prints timings:
  * the distributed render: shared NVLink tile queue with in-kernel gather, and static sharding + reduce;
  * the distributed integration: one ray queue for all ranks, per-ray results gathered in rank 0's memory by the
    kernels, trajectories paged where they are computed and found again through the rank-tagged page locator.
MK_SAME_GPU=1 runs every rank on GPU 0 (gloo instead of NCCL, which refuses two ranks on one device): the queue
counter and the result buffers are then CUDA-IPC mappings of the same device, i.e. exactly the code path of the
multi-GPU run, and a single-GPU box can execute the test."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist
import mahakala_b200 as ma
from mahakala_b200 import geodesics as geo, images, multigpu
from mahakala_b200.grmhd import AthenakFluidModel
from mahakala_b200.synthetic import make_synthetic_snapshot

res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 64
same_gpu = os.environ.get("MK_SAME_GPU", "0") == "1"
local = 0 if same_gpu else int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if same_gpu:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
m = None
if rank == 0:
    arr = make_synthetic_snapshot(ncells=nc, block=32 if nc % 32 == 0 else 16, extent=32.0, seed=0)
    m = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                      arr["x3f"], arr["LogicalLocations"], arr["Levels"], 0.94,
                                      fluid_gamma=arr["fluid_gamma"], storage="f32")
t0 = time.time()
m = multigpu.replicate_snapshot(m)      # geometry-only replicas on the other ranks + one broadcast of the cells
t_bcast = time.time() - t0
kw = dict(resolution=res, observing_frequencies=(230e9, 345e9))
ref = images.render(m, **kw) if rank == 0 else None
shared = multigpu.SharedImage(2, res * res)
out = {}
for mode in ("queue", "static"):
    for rep in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        img = multigpu.render_distributed(m, mode=mode, shared=shared if mode == "queue" else None, **kw)
        torch.cuda.synchronize(); dist.barrier()
        out[mode] = 1e3 * (time.perf_counter() - t0)
    if rank == 0:
        same = bool(torch.equal(img, ref))
        print(f"[{world} ranks] mode={mode}: identical to single-GPU image: {same}; {out[mode]:.2f} ms (wall, incl. barriers)")
        assert same
if rank == 0:
    torch.cuda.synchronize(); t0 = time.perf_counter(); images.render(m, **kw); torch.cuda.synchronize()
    print(f"single GPU: {1e3 * (time.perf_counter() - t0):.2f} ms; snapshot broadcast {t_bcast * 1e3:.1f} ms; flux {float(ref.sum()):.6e}")
dist.barrier()
shared.close()

# ---- one frequency with a learned patch order: the photon-ring patches go through the warp-specialised long-patch
# kernel (its own shared queue counter), the rest through the fused kernel; all ranks together vs rank 0 alone, and
# both against the plain fused kernel with no learned order
kw1 = dict(resolution=res, observing_frequencies=(230e9,))
plain = images.render(m, **kw1) if rank == 0 else None
images.learn_patch_order(0.94, resolution=res)
alone = images.render(m, **kw1) if rank == 0 else None
shared1 = multigpu.SharedImage(1, res * res)
img = multigpu.render_distributed(m, mode="queue", shared=shared1, **kw1)
if rank == 0:
    key = next(iter(images._learned_lengths))
    n_long = images.long_patch_count(images._learned_lengths[key], world)
    same = bool(torch.equal(img, alone)) and bool(torch.equal(img, plain))
    print(f"[{world} ranks] long-patch pipeline ({n_long} patches): identical to the fused single-GPU image: {same}")
    assert same and n_long > 0
images.forget_patch_orders()
# nothing learned: render() runs the coarse, capped pre-pass (images.quick_patch_order) on every rank by itself
img = multigpu.render_distributed(m, mode="queue", shared=shared1, **kw1)
if rank == 0:
    nq = list(images._quick_long.values())
    same = bool(torch.equal(img, plain))
    print(f"[{world} ranks] quick patch order ({nq} long patches): identical to the fused single-GPU image: {same}")
    assert same and len(nq) == 1 and nq[0] > 0
images.forget_patch_orders()
dist.barrier()
shared1.close()

# ---- distributed integration: `world` frames in one job ----
a, N, tol = 0.94, 10000, 1e-4
incl = [60.0, 17.0, 30.0, 80.0, 45.0, 70.0, 25.0, 52.0]
npx = res * res
s0_all = torch.cat([ma.initialize_geodesics_at_camera(a, incl[f % 8], 1000, -10, 10, res) for f in range(world)])
order = torch.from_numpy(multigpu.longest_first_ray_order(res, world)).cuda()
rays = multigpu.SharedRays(world * npx)
store = geo.TrajectoryStore.allocate(world * npx, N, mem_fraction=0.2)
for rep in range(2):
    views = multigpu.integrate_distributed(N, s0_all, 40, tol, a, store, rays, ray_order=order)
cpu = "cpu" if same_gpu else "cuda"
mine = torch.tensor([float(store.total_steps.item())], dtype=torch.float64, device=cpu)
per_rank = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(per_rank, mine)
# every rank gets the gathered step counts and page locators, integrates the job alone and checks (a) on rank 0 the
# gathered per-ray results, (b) on every rank the trajectories of rays IT holds, found through the rank-tagged locator
n = world * npx
nsteps_all = views["nsteps"].clone() if rank == 0 else torch.empty(n, dtype=torch.int32, device="cuda")
pf_all = views["page_first"].clone() if rank == 0 else torch.empty((n, 2), dtype=torch.int32, device="cuda")
dist.broadcast(nsteps_all, src=0)
dist.broadcast(pf_all, src=0)
alone = geo.TrajectoryStore.allocate(n, N, mem_fraction=0.2)
geo.integrate_paged(N, s0_all, 40, tol, a, store=alone)
owner = pf_all[:, 0] >> multigpu.PAGE_RANK_SHIFT
mine_idx = torch.nonzero(owner == rank).flatten()
mine_idx = mine_idx[torch.linspace(0, max(mine_idx.numel() - 1, 0), min(96, mine_idx.numel())).long()] if mine_idx.numel() else mine_idx
traj_ok = 1.0
if mine_idx.numel():
    store.page_first[mine_idx, 0] = pf_all[mine_idx, 0] & ((1 << multigpu.PAGE_RANK_SHIFT) - 1)
    store.page_first[mine_idx, 1] = pf_all[mine_idx, 1]
    store.nsteps.copy_(nsteps_all)
    S, dt = store.padded(mine_idx.cpu().numpy())
    Sa, dta = alone.padded(mine_idx.cpu().numpy())
    traj_ok = 1.0 if (torch.equal(S, Sa) and torch.equal(dt, dta)) else 0.0
chk = torch.tensor([traj_ok, float(mine_idx.numel())], dtype=torch.float64, device=cpu)
chks = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(chks, chk)
if rank == 0:
    same = (torch.equal(views["final"], alone.final) and torch.equal(views["nsteps"], alone.nsteps)
            and torch.equal(views["r_last"], alone.r_last))
    steps = [int(t.item()) for t in per_rank]
    print(f"[{world} ranks] shared ray queue: identical to single-GPU integration: {same}; rays per rank "
          f"{[int((owner == r).sum()) for r in range(world)]}, ray-steps per rank {steps}, total {sum(steps)} "
          f"(single GPU: {int(alone.total_steps.item())})")
    assert same and sum(steps) == int(alone.total_steps.item()) and int(owner.min()) >= 0 and int(owner.max()) < world
    traj_same = all(float(c[0]) == 1.0 for c in chks)
    print(f"[{world} ranks] trajectories of {[int(c[1]) for c in chks]} rays checked by the ranks holding them: "
          f"identical to the single-GPU dump: {traj_same}")
    assert traj_same and sum(int(c[1]) for c in chks) > 0
dist.barrier()
rays.close()
dist.destroy_process_group()
