"""BASELINE config 1 by the reference's own code: the a = 0.94, i = 60 deg, 64x64 grid camera (fov +-10 M, d = 1000 M),
geodesic_integrator(2000, s0, 40, 1e-2, a) and the last-point rule of select_photons_integrator (geodesics.py:370-378),
executed from /root/reference under the NumPy stand-in of make_reference_geodesics_golden.py.  4096 rays in Python
loops: the bundle is cut into chunks that run as separate processes (rays are independent; the reference's truncation
depends on the whole bundle, so only per-ray quantities are frozen: step count, classifier radius, end state).

    python tests/golden/make_reference_cfg1_golden.py            (spawns 8 workers, ~15 minutes)
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "reference_cfg1_golden.npz")
A, INC, RES, NCHUNK = 0.94, 60, 64, 8


def worker(k):
    sys.path.insert(0, HERE)
    import make_reference_geodesics_golden as G
    from make_reference_golden import arr, load
    G.install_stand_in()
    geo = load("mahakala.geodesics", "geodesics.py")
    s0 = np.asarray(geo.initialize_geodesics_at_camera(A, INC, 1000, -10, 10, RES))
    lo, hi = k * len(s0) // NCHUNK, (k + 1) * len(s0) // NCHUNK
    S, dt = geo.geodesic_integrator(2000, arr(s0[lo:hi]), 40, 1e-2, A)
    S, dt = np.asarray(S), np.asarray(dt)
    r = np.asarray(geo.radius_cal(arr(S), A))
    maxi = np.argmax(dt, axis=0) - 1                    # geodesics.py:373-378 (its two .at[].set() results are discarded)
    n = (dt != 0).sum(axis=0)
    np.savez(os.path.join("/tmp", "cfg1_chunk%d.npz" % k), s0=s0[lo:hi], nsteps=n, r_last=r[maxi, np.arange(hi - lo)],
             final=S[np.minimum(n, S.shape[0] - 1), np.arange(hi - lo)])


if __name__ == "__main__":
    if len(sys.argv) > 1:
        worker(int(sys.argv[1]))
    else:
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), str(k)]) for k in range(NCHUNK)]
        assert all(p.wait() == 0 for p in procs)
        parts = [np.load(os.path.join("/tmp", "cfg1_chunk%d.npz" % k)) for k in range(NCHUNK)]
        res = {key: np.concatenate([p[key] for p in parts]) for key in ("s0", "nsteps", "r_last", "final")}
        np.savez_compressed(OUT, **res)
        print("wrote", OUT, {k: v.shape for k, v in res.items()}, "captured:", int((res["r_last"] < 100).sum()),
              "ray-steps:", int(res["nsteps"].sum()))
