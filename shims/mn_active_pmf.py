"""Shim so the reference's own scripts (`from mn_active_pmf import ...`) resolve to the GPU-backed mirror.
Put this directory first on sys.path."""
from active_matrix_factorization_b200.mn_active_pmf import *  # noqa: F401,F403
import active_matrix_factorization_b200.mn_active_pmf as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})
