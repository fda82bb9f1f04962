"""Is the NumPy stand-in for JAX (make_reference_geodesics_golden.py) faithful?  Two numbers in the reference repo were
produced by REAL JAX: the trajectory shape (819, 60, 8) printed in demos/shadows.ipynb (equatorial camera, a = 0,
N = 10000, tol = 1e-4) and the golden shadow radii of tests/data/shadow_data.npy.  This script runs the reference's own
geodesics.py under the stand-in for exactly those calls -- i.e. it runs the reference's own tests/test_shadows.py
criterion, np.allclose(find_shadow_bisection_angles(...), radii, rtol=1e-2), for the golden cases -- and stores what
comes out in tests/golden/standin_validation.npz (checked by tests/test_oracle_cpu.py).  Result: shape (819, 60, 8);
all four cases pass, max relative errors 2.16e-3 / 2.33e-4 / 3.20e-4 / 1.85e-4 (test1..test4).  Python loops: about
six minutes for the notebook call + test4, 20-35 minutes per case for test1..3 (run them as separate processes).

    python tests/golden/validate_stand_in.py [test4 [test1 ...]]
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import make_reference_geodesics_golden as G
from make_reference_golden import arr, load
G.install_stand_in()
geo = load("mahakala.geodesics", "geodesics.py")
t0=time.time()
s0 = geo.initialize_geodesics_at_camera(0.0, 60, 1000, -15, 15, 60, camera_type='Equator')
S, dt = geo.geodesic_integrator(10000, s0, 40, 1e-4, 0.0)
print("notebook shape", np.asarray(S).shape, time.time()-t0, flush=True)
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shadow_golden.npz'))
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'standin_validation.npz')
res = dict(np.load(dst)) if os.path.exists(dst) else {}
res["notebook_shape"] = np.array(np.asarray(S).shape)
for case in (sys.argv[1:] or ["test4"]):
    t0 = time.time()
    r = np.asarray(geo.find_shadow_bisection_angles(float(z[case + '__bhspin']), float(z[case + '__inclination']), z[case + '__angles']))
    print(case, "max rel err vs the reference's golden radii", np.max(np.abs(r - z[case + '__radii']) / z[case + '__radii']),
          "reference test criterion:", np.allclose(r, z[case + '__radii'], rtol=1e-2), time.time() - t0, flush=True)
    res[case + "_radii"] = r
np.savez(dst, **res)
