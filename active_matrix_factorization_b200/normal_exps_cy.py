"""Host mirror of the reference's ``normal_exps_cy`` module (python-pmf/normal_exps_cy.pyx).

The single-index Isserlis formulas are scalar expressions kept for API compatibility (the
reference's test_normal_exps.py and notebooks call them one index tuple at a time); everything
that is evaluated over ratings or candidates -- ``exp_dotprod_sq`` and ``normal_gradient`` --
runs in libamf_b200 on the GPU.
"""
import numpy as np

from . import _native as N
from . import normal as _normal
from . import scoring as _scoring


def tripexpect(mean, cov, a, b, c):
    '''E[X_a X_b X_c] for N(mean, cov)                       (normal_exps_cy.pyx:9-14)'''
    return (mean[a] * mean[b] * mean[c]
            + mean[a] * cov[b, c] + mean[b] * cov[a, c] + mean[c] * cov[a, b])


def quadexpect(mean, cov, a, b, c, d):
    '''E[X_a X_b X_c X_d], distinct indices (Isserlis)        (normal_exps_cy.pyx:42-73)'''
    ma, mb, mc, md = mean[a], mean[b], mean[c], mean[d]
    pairs = cov[a, b] * cov[c, d] + cov[a, c] * cov[b, d] + cov[a, d] * cov[b, c]
    mixed = (ma * mb * cov[c, d] + ma * mc * cov[b, d] + ma * md * cov[b, c]
             + mb * mc * cov[a, d] + mb * md * cov[a, c] + mc * md * cov[a, b])
    return ma * mb * mc * md + mixed + pairs


def exp_squared(mean, cov, a, b):
    '''E[X_a^2 X_b^2]                                          (normal_exps_cy.pyx:77-87)'''
    return (4 * mean[a] * mean[b] * cov[a, b] + 2 * cov[a, b] ** 2
            + (mean[a] ** 2 + cov[a, a]) * (mean[b] ** 2 + cov[b, b]))


def exp_a2bc(mean, cov, a, b, c):
    '''E[X_a^2 X_b X_c]                                        (normal_exps_cy.pyx:91-107)'''
    ma, mb, mc = mean[a], mean[b], mean[c]
    return ((ma ** 2 + cov[a, a]) * (mb * mc + cov[b, c])
            + 2 * ma * mc * cov[a, b] + 2 * ma * mb * cov[a, c] + 2 * cov[a, b] * cov[a, c])


def exp_dotprod_sq(u, v, mean, cov, i, j):
    '''E[(U_i^T V_j)^2] = Var + E^2, from the scoring kernel   (normal_exps_cy.pyx:111-135)'''
    d, n = u.shape
    m = v.shape[1]
    e, _ = _scoring.score_normal(N.CRIT_APPROX_MEAN, mean, cov, n, m, d, [i], [j], "f64")
    var, _ = _scoring.score_normal(N.CRIT_PRED_VARIANCE, mean, cov, n, m, d, [i], [j], "f64")
    return float(var[0] + e[0] ** 2)


def normal_gradient(apmf):
    '''Gradient of the KL w.r.t. (mean, cov) of the model's approximation
    (normal_exps_cy.pyx:140-303), including its triangular-half convention.'''
    if apmf is None:
        raise TypeError("Argument 'apmf' must not be None")
    if apmf.mean is None or apmf.cov is None:
        raise TypeError("mean, cov are None; run initialize_approx first")
    p = _normal.fit_params(apmf.num_users, apmf.num_items, apmf.latent_d, apmf.sigma_sq,
                           apmf.sigma_u_sq, apmf.sigma_v_sq)
    batch = _normal.NormalBatch(apmf.ratings, p, apmf.mean[None], apmf.cov[None])
    gm, gc = batch.gradient()
    return gm[0], gc[0]
