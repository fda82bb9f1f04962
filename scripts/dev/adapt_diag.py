import sys, ctypes, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from host_harness import build
from oracle import mahakala_oracle as onp
from mahakala_b200 import geodesics as geo
from test_host_harness_cpu import _adaptive
hk = build.lib()
A=0.94
s0 = np.ascontiguousarray(onp.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, 64))[:1]
for N in list(range(1,12))+[20,40,60,80]:
    f,n,nr,rl = geo.integrate_adaptive(N, s0, 1e-2, A)
    hf,hn,hr,hl = _adaptive(hk, s0, 1e-9, N=N)
    print(N, 'dev r', float(rl[0]), 'n', int(n[0]), 'rej', int(nr[0]), '| host r', hl[0], 'n', hn[0], 'rej', hr[0], '| dt', float(f[0,0]), hf[0,0])
