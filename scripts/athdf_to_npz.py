"""Convert an AthenaK .athdf dump into the .npz container AthenakFluidModel reads without h5py.

    python scripts/athdf_to_npz.py snapshot.athdf snapshot.npz        (needs h5py on the converting machine)

Copies exactly the datasets the reference's loader reads (/root/reference/mahakala/grmhd/athenak.py:79-103):
x{1,2,3}v, x{1,2,3}f, uov, B, LogicalLocations, Levels and the VariableNames attribute."""
import sys

import numpy as np


def convert(src, dst):
    import h5py
    with h5py.File(src, 'r') as f:
        out = {k: np.array(f[k]) for k in ('x1v', 'x2v', 'x3v', 'x1f', 'x2f', 'x3f', 'uov', 'B', 'LogicalLocations', 'Levels')}
        out['VariableNames'] = np.array([n.decode('utf-8') for n in f.attrs['VariableNames']])
    np.savez(dst, **out)


if __name__ == "__main__":
    convert(sys.argv[1], sys.argv[2])
