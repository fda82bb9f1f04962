"""Write-only HBM bandwidth (what bounds the trajectory dump's 72 B per ray-step) next to the copy bandwidth."""
import torch, time
n = 8 << 30   # 8 Gi elements of uint8 = 8 GiB per buffer
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
tf = t(lambda: a.fill_(7)); tz = t(lambda: a.zero_()); tc = t(lambda: b.copy_(a))
af = a.view(torch.float64)
print(f"fill_  {n / tf / 1e6:8.1f} GB/s written   zero_ {n / tz / 1e6:8.1f} GB/s   copy_ {2 * n / tc / 1e6:8.1f} GB/s read+write ({n / tc / 1e6:.1f} each way)")
