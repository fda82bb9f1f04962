"""torchrun --nproc-per-node N scripts/dev/bcast_probe.py: 0.64 GB float32 from rank 0 to everybody: NCCL broadcast
against scatter + all-gather (multigpu.broadcast_large)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from mahakala_b200 import multigpu
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
multigpu.warm_communicator()
n = 160989184
src = torch.arange(n, dtype=torch.float32, device="cuda") if rank == 0 else None
def run(mode):
    os.environ["MK_BCAST"] = mode
    ts = []
    for it in range(5):
        w = src.clone() if rank == 0 else torch.zeros(n, dtype=torch.float32, device="cuda")
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); multigpu.broadcast_large(w, 0); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = torch.tensor([float(w[12345] == 12345.0 and w[-1] == float(n - 1) and float(w.sum(dtype=torch.float64)) == n * (n - 1) / 2)], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if it: ts.append(float(t))
    if rank == 0:
        print(f"[{world} GPUs] {mode:5s}: min {min(ts):.2f} ms = {4 * n / min(ts) / 1e6:.0f} GB/s, mean {sum(ts) / len(ts):.2f} ms, correct on all ranks: {bool(ok.item())}", flush=True)
run("nccl"); run("sag"); run("nccl"); run("sag")
dist.destroy_process_group()
