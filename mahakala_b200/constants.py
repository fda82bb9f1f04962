"""Physical constants in cgs units.

The values are the reference's, digit for digit (/root/reference/mahakala/constants.py:23-31; ``MP`` and
``EC`` are the extra entries the reference repeats in electrons.py:23-29 and grmhd/grmhd.py:26-32): parity
of Theta_e, j_nu and alpha_nu depends on every digit, so they are kept in one table that is also handed to
the CUDA kernels (``mk_emission_params``) instead of being retyped there.
"""

CGS = {
    "EE": 4.8032e-10,       # electron charge [esu]
    "KB": 1.3807e-16,       # Boltzmann constant [erg/K]
    "CL": 2.99792458e10,    # speed of light [cm/s]
    "ME": 9.1094e-28,       # electron mass [g]
    "MP": 1.6726e-24,       # proton mass [g]
    "EC": 4.8032e-10,       # electron charge again, as electrons.py / grmhd.py name it
    "HPL": 6.6261e-27,      # Planck constant [erg s]
    "GNEWT": 6.6743e-8,     # gravitational constant [cm^3 g^-1 s^-2]
    "Msun": 1.989e33,       # solar mass [g]
}

globals().update(CGS)
__all__ = sorted(CGS)
