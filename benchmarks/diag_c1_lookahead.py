"""Diagnostic: per-candidate comparison of the C1 uv-entropy lookahead against the reference
(which candidates differ, and how the two line searches diverged)."""
import contextlib, io, os, random, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from copy import deepcopy
from oracle import ref_loader
ref = ref_loader.load()
from active_matrix_factorization_b200 import active_pmf as A
from active_matrix_factorization_b200 import normal as NM

def build(mod):
    np.random.seed(0); random.seed(0)
    real, ratings, vals = ref.active_pmf.make_fake_data(noise=.25, num_users=10, num_items=10, rank=2,
                                             data_type='binary', mask_type='diag')
    a = mod.ActivePMF(ratings, latent_d=2, rating_values=vals, discrete_expectations=True)
    a.do_fit(); a.initialize_approx(); a.fit_normal()
    return a
g, r = build(A), build(ref.active_pmf)
print("state diff mean/cov after each side's own initial fit:", np.abs(g.mean - r.mean).max(), np.abs(g.cov - r.cov).max())
if "--same-state" in sys.argv:
    g.mean, g.cov = r.mean.copy(), r.cov.copy()      # identical starting point for the lookahead
    print("lookahead started from the reference's fitted state on both sides")
pool = sorted(g.unrated)[:12]
for (i, j) in pool:
    for v in (0.0, 1.0):
        rr = deepcopy(r); rr.add_rating(i, j, v)
        with contextlib.redirect_stdout(io.StringIO()):
            rk = list(rr.fit_normal_kls())
        re = rr._approx_entropy()
        b = NM.NormalBatch(g.ratings, g._fit_params(), g.mean[None], g.cov[None],
                           extra=(np.array([i], np.int32), np.array([j], np.int32), np.array([v])))
        res = b.fit(trace_len=4096, want_entropy=True)
        gk = res['trace'][0][:int(res['steps'][0])]
        k = min(len(gk), len(rk))
        dk = np.abs(np.array(gk[:k]) - np.array(rk[:k])) / np.abs(np.array(rk[:k])) if k else np.zeros(1)
        print((i, j, v), "steps gpu/ref", len(gk), len(rk), "entropy gpu/ref %.6f %.6f" % (res['entropy'][0], re),
              "max kl rel diff over common steps %.2e" % dk.max(), "first>1e-8 at", int(np.argmax(dk > 1e-8)) if (dk > 1e-8).any() else None)
