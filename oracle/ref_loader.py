"""Loader for the compiled reference in ``oracle/_ref`` -- TEST INFRASTRUCTURE ONLY.

Imports the reference's own modules (``pmf_cy``, ``normal_exps_cy``, ``active_pmf``,
``bayes_pmf``; built by ``oracle/build_ref.py``) as top-level modules, exactly as the
reference's driver scripts do (``active_pmf.py:19-31``, ``bayes_pmf.py:32``).  The product
package lives under ``active_matrix_factorization_b200`` so the two never collide.
"""
import importlib
import os
import sys

import numpy as np

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    from . import build_ref
    return build_ref.built()


def load():
    """Returns a namespace with the reference modules; raises ImportError if not built."""
    if not available():
        raise ImportError("oracle/_ref is not built; run python oracle/build_ref.py")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    if not hasattr(np, "float"):       # SURVEY.md 8c patch 1 for the pure-python twins
        np.float = float
    class NS:
        pass
    ns = NS()
    for name in ("pmf_cy", "normal_exps_cy", "matrix_normal_exps_cy", "active_pmf",
                 "mn_active_pmf", "bayes_pmf"):
        mod = importlib.import_module(name)
        if not os.path.abspath(mod.__file__).startswith(REF_DIR):
            raise ImportError("%s resolved to %s, not oracle/_ref" % (name, mod.__file__))
        setattr(ns, name, mod)
    return ns
