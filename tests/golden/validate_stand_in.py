"""Is the NumPy stand-in for JAX (make_reference_geodesics_golden.py) faithful?  Two numbers in the reference repo were
produced by REAL JAX: the trajectory shape (819, 60, 8) printed in demos/shadows.ipynb (equatorial camera, a = 0,
N = 10000, tol = 1e-4) and the golden shadow radii of tests/data/shadow_data.npy.  This script runs the reference's own
geodesics.py under the stand-in for exactly those calls (case test4: a = 0.71, i = 13 deg, 25 angles) and stores what
comes out in tests/golden/standin_validation.npz (checked by tests/test_oracle_cpu.py).  Result: shape (819, 60, 8);
radii within 1.8e-4 relative of the golden ones (the reference's own test allows 1e-2).  About six minutes.

    python tests/golden/validate_stand_in.py
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
t0=time.time()
r = geo.find_shadow_bisection_angles(float(z['test4__bhspin']), float(z['test4__inclination']), z['test4__angles'])
print("test4 max rel err vs the reference's golden radii", np.max(np.abs(np.asarray(r)-z['test4__radii'])/z['test4__radii']), time.time()-t0, flush=True)
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'standin_validation.npz'), notebook_shape=np.array(np.asarray(S).shape), test4_radii=np.asarray(r))
