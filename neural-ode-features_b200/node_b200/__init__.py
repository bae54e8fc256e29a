"""node_b200 - host side of the B200-native dopri5 / ODE-Net hot path.

`odeint`, `odeint_adjoint`: drop-in for the reference's torchdiffeq (see solver.py);
`models`: mirror of the reference's ODENet / ODEBlock / ODEfunc module surface;
`native`: ctypes binding of the C ABI (include/node_b200.h); `distributed`: batch sharding; `wide`: 128 / 192 / 256 filters;
`unrolled`: gradients through the non-adjoint odeint; `each`: per-sample solves; `retrieval`, `caller_ops`, `caller_grad`: the
callers either side of the path.
"""
from .solver import invalidate_caches, odeint, odeint_adjoint  # noqa: F401

__all__ = ['odeint', 'odeint_adjoint', 'invalidate_caches']
