"""Regenerate tests/golden/shadow_golden.npz from the reference's own golden file.

Run in the build container only (needs /root/reference):
    python tests/golden/make_shadow_golden.py

Source: /root/reference/tests/data/shadow_data.npy (pickled dict of 4 cases, consumed by
/root/reference/tests/test_shadows.py:28-45).  We re-save it as a pickle-free .npz so that the
fixtures can be loaded with allow_pickle=False on the GPU box, where /root/reference is absent.
"""
import os
import numpy as np

SRC = "/root/reference/tests/data/shadow_data.npy"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shadow_golden.npz")

if __name__ == "__main__":
    cases = np.load(SRC, allow_pickle=True).item()
    out = {}
    for key, c in cases.items():
        out[f"{key}__bhspin"] = np.float64(c["bhspin"])
        out[f"{key}__inclination"] = np.float64(c["inclination"])
        out[f"{key}__angles"] = np.asarray(c["angles"], dtype=np.float64)
        out[f"{key}__radii"] = np.asarray(c["radii"], dtype=np.float64)
    np.savez(DST, **out)
    print("wrote", DST, sorted(out))
