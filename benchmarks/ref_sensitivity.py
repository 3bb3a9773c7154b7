"""CPU-only: how sensitive the REFERENCE's own exact-mode variational fit is to noise at the level of a
backward-stable eigensolver (1e-16 * norm): re-runs the reference (oracle/_ref) with perturbed
starting covariances.  Used to interpret GPU-vs-reference differences (parity caveat)."""
import random, sys
import numpy as np
sys.path.insert(0, '/root/repo')
from copy import deepcopy
from oracle import ref_loader
ref = ref_loader.load()
np.random.seed(0); random.seed(0)
real, ratings, vals = ref.active_pmf.make_fake_data(noise=.25, num_users=10, num_items=10, rank=2, data_type='binary', mask_type='diag')
r = ref.active_pmf.ActivePMF(ratings, latent_d=2, rating_values=vals, discrete_expectations=True)
r.do_fit(); r.initialize_approx()
cov0 = r.cov.copy(); mean0 = r.mean.copy()
r.fit_normal()
w = np.linalg.eigvalsh(r.cov); print("fitted cov: norm %.3g, eigenvalues min %.3g, #<=1.01e-5: %d, cond %.3g" % (np.abs(r.cov).max(), w.min(), (w <= 1.01e-5).sum(), w.max() / w.min()))
rng = np.random.RandomState(1)
# (a) the reference re-run with its INITIAL random covariance perturbed by symmetric noise of 1e-16 * norm
for scale in (1e-16, 1e-14):
    p = deepcopy(r); p.mean = mean0.copy()
    E = rng.normal(size=cov0.shape); E = (E + E.T) / 2
    p.cov = cov0 + scale * np.abs(cov0).max() * E
    p.fit_normal()
    print("initial fit, noise %.0e*norm: |cov - cov_ref|max = %.2e, entropy %.8f vs %.8f" % (scale, np.abs(p.cov - r.cov).max(), p._approx_entropy(), r._approx_entropy()))
# (b) lookahead re-fits from the fitted state perturbed the same way
for (i, j, val) in [(0, 2, 0.), (0, 8, 0.), (1, 2, 1.), (1, 3, 0.)]:
    base = deepcopy(r); base.add_rating(i, j, val); base.fit_normal(); e0 = base._approx_entropy()
    outs = []
    for t in range(3):
        p = deepcopy(r)
        E = rng.normal(size=cov0.shape); E = (E + E.T) / 2
        p.cov = p.cov + 1e-16 * np.abs(p.cov).max() * E
        p.add_rating(i, j, val); p.fit_normal(); outs.append(p._approx_entropy())
    print((i, j, val), "reference entropy %.6f; with 1e-16*norm noise on the starting cov:" % e0, ["%.6f" % o for o in outs])
