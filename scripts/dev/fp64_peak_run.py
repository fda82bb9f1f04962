"""The roofline denominator on its own: mk_measure_fp64_peak (dfma_peak_kernel, 8 independent DFMA chains per thread,
148 x 4 x 256 threads) with the SM clock read next to it.  Run plain for the number, under ncu for the pipe utilisation:
  ncu --metrics sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:dfma_peak python scripts/dev/fp64_peak_run.py"""
import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mahakala_b200 import _cabi
torch.zeros(1, device="cuda")
for k in range(3):
    tf, ms = ctypes.c_double(0), ctypes.c_double(0)
    _cabi.call("mk_measure_fp64_peak", 20000, tf, ms)
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw",
                          "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    print(f"run {k}: {tf.value:.3f} TFLOP/s FP64 (DFMA = 2 flop), {ms.value:.3f} ms; nvidia-smi after: {clk}")
print("nominal: 148 SMs x 64 FP64 lanes x 2 flop x 1.965 GHz = 37.22 TFLOP/s")
