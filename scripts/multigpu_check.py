"""torchrun --nproc-per-node N scripts/multigpu_check.py [res] [ncells]
Checks that the distributed render (shared NVLink tile queue, and static + NCCL reduce) reproduces the
single-GPU image bit for bit, and prints timings."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist
from mahakala_b200 import images, multigpu
from mahakala_b200.grmhd import AthenakFluidModel
from mahakala_b200.synthetic import make_synthetic_snapshot

res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 64
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
m = None
if rank == 0:
    arr = make_synthetic_snapshot(ncells=nc, block=32 if nc % 32 == 0 else 16, extent=32.0, seed=0)
    m = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                      arr["x3f"], arr["LogicalLocations"], arr["Levels"], 0.94,
                                      fluid_gamma=arr["fluid_gamma"], storage="f32")
t0 = time.time()
m = multigpu.replicate_snapshot(m)      # geometry-only replicas on the other ranks + one NCCL broadcast of the cells
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
        print(f"[{world} GPUs] mode={mode}: identical to single-GPU image: {same}; {out[mode]:.2f} ms (wall, incl. barriers)")
        assert same
if rank == 0:
    torch.cuda.synchronize(); t0 = time.perf_counter(); images.render(m, **kw); torch.cuda.synchronize()
    print(f"single GPU: {1e3 * (time.perf_counter() - t0):.2f} ms; snapshot broadcast {t_bcast * 1e3:.1f} ms; flux {float(ref.sum()):.6e}")
dist.barrier()
shared.close()
dist.destroy_process_group()
