"""Parity report (run on the GPU box): error distributions of the CUDA path against the C oracle, with the
oracle-vs-oracle spread (NumPy/LAPACK restatement vs C restatement) as the noise floor.  Writes
gpurun_out/parity_report.txt (copied to profiles/ for the record)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mahakala_b200 as ma
from mahakala_b200 import geodesics as geo, images
from oracle import c_oracle, mahakala_oracle as onp
from helpers import M_BH, MASS_SCALE, device_model, oracle_model, snapshot_arrays

A = 0.94
lines = []


def pct(x):
    x = np.asarray(x).ravel()
    return f"median {np.median(x):.2e}  p90 {np.percentile(x, 90):.2e}  p99 {np.percentile(x, 99):.2e}  max {x.max():.2e}"


def state_err(a, b):
    return np.abs(a - b).max(axis=1) / np.abs(b).max(axis=1)


for (name, res, tol, N, stride) in (("cfg1 64x64 tol=1e-2 N=2000", 64, 1e-2, 2000, 1),
                                    ("cfg2 1024x1024 tol=1e-4 N=10000, every 8th pixel", 1024, 1e-4, 10000, 8)):
    s0 = ma.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, res)
    f, n, rl = geo.integrate_final(N, s0, 40, tol, A)
    idx = (np.arange(0, res, stride)[:, None] * res + np.arange(0, res, stride)[None, :]).reshape(-1)
    s0h = np.asarray(s0)[idx]
    f, n, rl = (np.asarray(q.cpu())[idx] for q in (f, n, rl))
    ref = c_oracle.integrate(N, s0h, 40, tol, A)
    cap, cap_ref = rl < 100, ref["r_last"] < 100
    esc = ~cap_ref
    lines.append(f"== {name}: {idx.size} rays, {int(cap_ref.sum())} captured")
    lines.append(f"   shadow classification mismatches (CUDA vs C oracle): {int((cap != cap_ref).sum())}")
    lines.append(f"   step-count mismatches, escaped rays: {int((n[esc] != ref['nsteps'][esc]).sum())} of {int(esc.sum())};"
                 f" captured rays: {int((n[~esc] != ref['nsteps'][~esc]).sum())} of {int((~esc).sum())}")
    lines.append(f"   final state rel. err, escaped rays : {pct(state_err(f[esc], ref['final'][esc]))}")
    same = (~esc) & (n == ref["nsteps"])
    if same.any():
        lines.append(f"   final state rel. err, captured rays with equal step count ({int(same.sum())}): {pct(state_err(f[same], ref['final'][same]))}")
    if idx.size <= 4096:
        S1, d1 = onp.geodesic_integrator(N, s0h[::8], 40, tol, A)
        o2 = c_oracle.integrate(N, s0h[::8], 40, tol, A)
        e2 = onp.last_point_radius(S1, d1, A) >= 100
        lines.append(f"   NOISE FLOOR NumPy-oracle vs C-oracle, escaped rays ({int(e2.sum())}): {pct(state_err(S1[-1][e2], o2['final'][e2]))}")

arr = snapshot_arrays(ncells=64, block=16, extent=32.0)
om, dm = oracle_model(arr, A), device_model(arr, A)
units = om.get_units(M_BH, MASS_SCALE)
res = 64
s0 = onp.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, res)
ref, _, nin = c_oracle.render(om, s0, units, [230e9])
img = images.make_image(dm, resolution=res).reshape(-1)
scale = ref[0].max()
m = ref[0] > 1e-6 * scale
lines.append(f"== image 64x64, synthetic 64^3 snapshot, 230 GHz: {int(m.sum())} lit pixels, {nin} in-domain samples")
lines.append(f"   per-pixel rel. err (pixels > 1e-6 of max): {pct(np.abs(img[m] - ref[0][m]) / ref[0][m])}")
lines.append(f"   total flux rel. err: {abs(img.sum() - ref[0].sum()) / ref[0].sum():.2e}   (north-star: 1e-6 per pixel, 1e-8 flux)")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "parity_report.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
