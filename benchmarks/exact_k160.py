"""Exact-mode (full covariance) variational fit on SURVEY.md 8d's shrunken config-2 instance
(12x20, d=5, k=160): GPU vs the reference's CPU path, bounded number of accepted steps."""
import os, sys, time, random
from itertools import islice
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_matrix_factorization_b200 import active_pmf as A
from oracle import build_ref, ref_loader
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "random_12x20_d5.npz"))
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 30

def run(mod, cap):
    np.random.seed(7)
    a = mod.ActivePMF(g["ratings"], 5, rating_values={-1, 1} if False else None)
    a.users, a.items = g["users"].copy(), g["items"].copy()
    a.initialize_approx()
    t0 = time.perf_counter()
    if hasattr(a, "max_normal_steps"):
        a.max_normal_steps = cap; kls = list(a.fit_normal_kls())
    else:
        kls = list(islice(a.fit_normal_kls(), cap))
    torch.cuda.synchronize()
    return kls, time.perf_counter() - t0

kg, tg = run(A, STEPS)
kg, tg = run(A, STEPS)
print("gpu: %d steps in %.3f s (%.1f ms/step), kl[-1]=%.10g" % (len(kg), tg, 1e3 * tg / max(1, len(kg)), kg[-1]))
if build_ref.built():
    ref = ref_loader.load()
    kr, tr = run(ref.active_pmf, STEPS)
    k = min(len(kg), len(kr))
    print("ref: %d steps in %.3f s (%.1f ms/step), kl[-1]=%.10g; max rel diff over %d common steps %.2e" % (
        len(kr), tr, 1e3 * tr / max(1, len(kr)), kr[-1], k, np.abs(np.array(kg[:k]) - np.array(kr[:k])).max() / abs(kr[k - 1])))
