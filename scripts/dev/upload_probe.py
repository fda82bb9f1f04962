"""Host -> device upload strategies for a 537 MB pageable float32 array (the interior cells of the 256^3 snapshot)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from mahakala_b200 import _cabi
n = 5 * 512 * 32**3
h = np.random.default_rng(0).random(n, dtype=np.float32)
d = torch.empty(n, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
def t(fn, reps=3):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); out.append(1e3 * (time.perf_counter() - t0))
    return [round(x, 1) for x in out]
print("bytes", h.nbytes, "cpus", len(os.sched_getaffinity(0)))
print("torch pageable copy_        ms", t(lambda: d.copy_(torch.from_numpy(h))))
for th in (1, 4, 8, 16):
    print(f"staged, {th:2d} host threads     ms", t(lambda: _cabi.call("mk_upload_pageable", d, h.ctypes.data, int(h.nbytes), th, None)))
print("registered in place         ms", t(lambda: _cabi.call("mk_upload_pageable", d, h.ctypes.data, int(h.nbytes), -1, None)))
print("check", bool(torch.equal(d.cpu(), torch.from_numpy(h))))
hp = torch.from_numpy(h).pin_memory()
print("already pinned copy_        ms", t(lambda: d.copy_(hp, non_blocking=True)))
