#!/usr/bin/env python3
"""Evaluates the criteria that tests/golden pins only loosely so far (one-step lookahead, the
prediction-entropy bound, Bayesian expected variance) on the fixture problems and writes them to
gpurun_out/more_gpu.npz, to be compared offline with the reference's values
(tests/golden/make_golden.py more).  Results are saved after every stage."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from active_matrix_factorization_b200 import active_pmf as A, bayes_pmf as Bm  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "more_gpu.npz")
os.makedirs(os.path.dirname(OUT), exist_ok=True)
out = {}


def save():
    np.savez_compressed(OUT, **out)


g = np.load(os.path.join(ROOT, "tests", "golden", "lookahead_6x7_d2.npz"))
a = A.ActivePMF(g["ratings"], 2, rating_values={0, 1}, discrete_expectations=True)
a.compute_dtype = "f64"
a.users, a.items = g["users"].copy(), g["items"].copy()
a.mean, a.cov = g["mean"].copy(), g["cov"].copy()
cand = list(zip(g["cand_i"].tolist(), g["cand_j"].tolist()))[:4]
with contextlib.redirect_stdout(io.StringIO()):
    out["onestep_ge_half"] = np.array([a.onestep_ge_half(c) for c in cand]); save()
    out["onestep_ge_half_approx"] = np.array([a.onestep_ge_half_approx(c) for c in cand]); save()
    out["pred_covs"] = a.approx_pred_covs()
    out["pred_entropy_bound"] = a._pred_entropy_bound(); save()
    out["exp_pred_entropy_bound"] = np.array([a.exp_pred_entropy_bound(c) for c in cand[:2]]); save()

g = np.load(os.path.join(ROOT, "tests", "golden", "gibbs_15x12_d3.npz"))
for tag, kw in (("disc", dict(rating_values=(1, 2, 3, 4, 5), discrete_expectations=True)),
                ("cont", dict(rating_values=None, discrete_expectations=False, num_integration_pts=5))):
    b = Bm.BayesianPMF(g["ratings"], 3, **kw)
    b.compute_dtype = "f64"
    b.users, b.items = g["users"].copy(), g["items"].copy()
    samples = list(zip(g["samples_u"], g["samples_v"]))
    which = tuple(np.array(sorted(b.unrated)[:2]).T)
    np.random.seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        out["ev_" + tag] = b.exp_variance(samples, which=which, num_samps=3, fit_first=False)
    save()
print("wrote", OUT, sorted(out))
