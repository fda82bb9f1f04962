"""torchrun --nproc-per-node N scripts/dev/strong_probe.py [res]: ONE res^2 cfg4 frame rendered by all ranks together
(shared tile queue + in-kernel gather), learned patch order, for several sizes of the long-patch launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist
from mahakala_b200 import images, multigpu
from mahakala_b200.grmhd import AthenakFluidModel
from mahakala_b200.synthetic import make_synthetic_snapshot
res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
NUS = [43e9, 86e9, 130e9, 230e9, 345e9, 460e9, 690e9, 870e9][:int(sys.argv[2])] if len(sys.argv) > 2 else [230e9]
MS = 2e24 if len(NUS) > 1 else 1e26
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
m = None
if rank == 0:
    arr = make_synthetic_snapshot(ncells=256, block=32, extent=32.0, seed=0, dtype=np.float32)
    m = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                      arr["x3f"], arr["LogicalLocations"], arr["Levels"], 0.94, fluid_gamma=arr["fluid_gamma"], storage="f64")
m = multigpu.replicate_snapshot(m)
ref = images.render(m, resolution=res, observing_frequencies=NUS, mass_scale=MS) if rank == 0 else None
images.learn_patch_order(0.94, resolution=res)
key = next(iter(images._learned_lengths))
L = images._learned_lengths[key]
shared = multigpu.SharedImage(len(NUS), res * res)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def leg(n_long, reps=4, patch_range=(0, -1, 1), skip_bulk=False):
    images._DEV_SKIP_BULK = skip_bulk
    ts = []
    for it in range(reps + 1):
        flush.fill_(it)
        shared.reset(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        images.render(m, resolution=res, observing_frequencies=NUS, mass_scale=MS, image_out=shared.image_ptr, queue=shared.queue_ptr, long_queue=shared.ring_queue_ptr,
                      participants=world, long_patches=n_long, patch_range=patch_range)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        if it: ts.append(float(t))
    same = bool(torch.equal(shared.local_view()[1], ref)) if rank == 0 else None
    return min(ts), float(np.mean(ts)), same
for thr in (None, 0.3, 0.25):
    n_long = 0 if thr is None else min(int((L >= thr * L[0]).sum()), world * 148)
    r = leg(n_long)
    if rank == 0:
        print(f"[{world} GPUs, {res}^2, {len(NUS)} frequencies, exclusive={images._LONG_EXCLUSIVE}] threshold {thr}: {n_long:5d} long patches: min {r[0]:.2f} ms mean {r[1]:.2f} ms identical {r[2]}", flush=True)
n_long = min(int((L >= 0.3 * L[0]).sum()), world * 148)
r = leg(n_long, skip_bulk=True)
if rank == 0: print(f"[{world} GPUs] long-patch launch alone ({n_long} patches): min {r[0]:.2f} ms mean {r[1]:.2f}", flush=True)
r = leg(0, patch_range=(n_long, -1, 1))
if rank == 0: print(f"[{world} GPUs] bulk launch alone (patches {n_long}..): min {r[0]:.2f} ms mean {r[1]:.2f}", flush=True)
r = leg(0, patch_range=(len(L) - 64, -1, 1))
if rank == 0: print(f"[{world} GPUs] 64 shortest patches only (fixed cost of a frame): min {r[0]:.2f} ms mean {r[1]:.2f}", flush=True)
dist.barrier(); shared.close(); dist.destroy_process_group()
