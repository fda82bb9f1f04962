"""ORACLE PINNING AID — TEST INFRASTRUCTURE ONLY.

Line-for-line torch.func transliteration of the reference's rhs (geodesics.py:294-347): the metric is
typed exactly as geodesics.py:88-104, its Jacobian comes from ``torch.func.jacfwd`` (as ``jax.jacfwd``),
the inverse from ``torch.linalg.inv`` (as ``jnp.linalg.inv``).  Slow; used only to pin the NumPy and C
oracles at a handful of states (tests/test_oracle_pinning.py)."""
import torch
from torch.func import jacfwd, vmap


def metric(x, bhspin):
    eta = torch.diag(torch.tensor([-1., 1., 1., 1.], dtype=x.dtype))
    a = bhspin
    aa = a * a
    zz = x[3]**2.
    kk = 0.5 * (x[1] * x[1] + x[2] * x[2] + zz - aa)
    rr = torch.sqrt(kk * kk + aa * zz) + kk
    r = torch.sqrt(rr)
    f = (2.0 * rr * r) / (rr * rr + aa * zz)
    l = torch.stack([torch.ones_like(r), (r * x[1] + a * x[2]) / (rr + aa), (r * x[2] - a * x[1]) / (rr + aa), x[3] / r])
    return eta + f * (l[:, None] * l[None, :])


def imetric(x, bhspin):
    return torch.linalg.inv(metric(x, bhspin))


def rhs(state1, bhspin):
    x = state1[:4]
    v = state1[4:]
    ig = imetric(x, bhspin)
    jg = jacfwd(metric)(x, bhspin)
    a = ig @ (- (jg @ v) @ v + 0.5 * v @ (v @ jg))
    return torch.cat([v, a])


def vectorized_rhs(s0, bhspin):
    return vmap(rhs, in_dims=(0, None))(torch.as_tensor(s0, dtype=torch.float64), bhspin).numpy()
