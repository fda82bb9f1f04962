"""Where does the snapshot set-up time go?  (host arrays -> device snapshot of the 256^3 cfg4 snapshot, three times)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from mahakala_b200.grmhd import AthenakFluidModel
from mahakala_b200.synthetic import make_synthetic_snapshot
torch.zeros(1, device="cuda")
for dtype in (np.float32, np.float64):
    arr = make_synthetic_snapshot(ncells=256, block=32, extent=32.0, seed=0, dtype=dtype)
    for storage in ("f64", "auto"):
        for rep in range(3):
            t0 = time.perf_counter()
            m = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                              arr["x3f"], arr["LogicalLocations"], arr["Levels"], 0.94,
                                              fluid_gamma=arr["fluid_gamma"], storage=storage)
            t1 = time.perf_counter()
            m.snapshot(); torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(np.dtype(dtype).name, storage, "ctor ms %.1f snapshot ms %.1f" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)),
                  {k: round(v, 1) for k, v in m.setup_timing.items()}, m.storage, flush=True)
            m.release()
