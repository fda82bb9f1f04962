"""cgs constants, digit for digit as /root/reference/mahakala/constants.py:23-31 (plus MP, EC which the
reference duplicates in electrons.py:23-29 and grmhd/grmhd.py:26-32)."""

# universal constants
EE = 4.8032e-10
KB = 1.3807e-16
CL = 2.99792458e10
ME = 9.1094e-28
MP = 1.6726e-24
EC = 4.8032e-10
HPL = 6.6261e-27
GNEWT = 6.6743e-8

# other quantities
Msun = 1.989e33
