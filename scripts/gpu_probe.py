"""Quick GPU probe: FP64 DFMA peak and integrate-kernel throughput (not the bench; a development aid)."""
import ctypes
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mahakala_b200 as ma
from mahakala_b200 import _cabi, geodesics as geo

tf = ctypes.c_double(0); ms = ctypes.c_double(0)
for it in (2000, 20000):
    _cabi.call("mk_measure_fp64_peak", it, tf, ms)
    print(f"fp64 peak: iters={it} {tf.value:.2f} TFLOP/s ({ms.value:.3f} ms)")

a = 0.94
for res, tol, N in ((256, 1e-2, 2000), (1024, 1e-4, 10000)):
    s0 = ma.initialize_geodesics_at_camera(a, 60, 1000, -10, 10, res)
    torch.cuda.synchronize()
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        final, nsteps, r_last, total = geo.integrate_final(N, s0, 40, tol, a, want_total=True)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        steps = int(total.item())
        print(f"res={res} tol={tol}: {t:.2f} ms, {steps} ray-steps, {steps/t/1e6:.2f} G ray-steps/s, "
              f"{steps/t/1e6*859/1e3:.2f} TFLOP/s algorithmic; max steps {int(nsteps.max())} captured {(r_last<100).sum().item()}")
