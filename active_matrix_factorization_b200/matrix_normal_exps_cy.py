"""Host mirror of the reference's ``matrix_normal_exps_cy`` module
(python-pmf/matrix_normal_exps_cy.pyx): moments of X ~ MN(mean, cov_rows, cov_cols) with
Cov(X_ai, X_bj) = cov_rows[a, b] * cov_cols[i, j].  The single-entry formulas are scalar
expressions kept for API compatibility; what runs over ratings or candidates
(``exp_dotprod_sq``, ``mn_kl_divergence``, ``matrixnormal_gradient``) runs on the GPU."""
import numpy as np

from . import _native as N
from . import normal as _normal


def _c(cr, cc, a, b):
    return cr[a[0], b[0]] * cc[a[1], b[1]]


def tripexpect(mean, cov_rows, cov_cols, a_i, a_j, b_i, b_j, c_i, c_j):
    '''E[a b c]                                        (matrix_normal_exps_cy.pyx:9-24)'''
    a, b, c = (a_i, a_j), (b_i, b_j), (c_i, c_j)
    ma, mb, mc = mean[a], mean[b], mean[c]
    return (ma * mb * mc + ma * _c(cov_rows, cov_cols, b, c) + mb * _c(cov_rows, cov_cols, a, c)
            + mc * _c(cov_rows, cov_cols, a, b))


def quadexpect(mean, cov_rows, cov_cols, a_i, a_j, b_i, b_j, c_i, c_j, d_i, d_j):
    '''E[a b c d]                                      (matrix_normal_exps_cy.pyx:27-71)'''
    a, b, c, d = (a_i, a_j), (b_i, b_j), (c_i, c_j), (d_i, d_j)
    ma, mb, mc, md = mean[a], mean[b], mean[c], mean[d]
    cv = lambda x, y: _c(cov_rows, cov_cols, x, y)
    return (ma * mb * mc * md
            + ma * mb * cv(c, d) + ma * mc * cv(b, d) + ma * md * cv(b, c)
            + mb * mc * cv(a, d) + mb * md * cv(a, c) + mc * md * cv(a, b)
            + cv(a, b) * cv(c, d) + cv(a, c) * cv(b, d) + cv(a, d) * cv(b, c))


def exp_squared(mean, cov_rows, cov_cols, a_i, a_j, b_i, b_j):
    '''E[a^2 b^2]                                      (matrix_normal_exps_cy.pyx:75-94)'''
    a, b = (a_i, a_j), (b_i, b_j)
    ma, mb = mean[a], mean[b]
    cab = _c(cov_rows, cov_cols, a, b)
    return (4 * ma * mb * cab + 2 * cab ** 2
            + (ma ** 2 + _c(cov_rows, cov_cols, a, a)) * (mb ** 2 + _c(cov_rows, cov_cols, b, b)))


def exp_a2bc(mean, cov_rows, cov_cols, a_i, a_j, b_i, b_j, c_i, c_j):
    '''E[a^2 b c]                                      (matrix_normal_exps_cy.pyx:98-121)'''
    a, b, c = (a_i, a_j), (b_i, b_j), (c_i, c_j)
    ma, mb, mc = mean[a], mean[b], mean[c]
    cv = lambda x, y: _c(cov_rows, cov_cols, x, y)
    return ((ma ** 2 + cv(a, a)) * (mb * mc + cv(b, c))
            + 2 * ma * mc * cv(a, b) + 2 * ma * mb * cv(a, c) + 2 * cv(a, b) * cv(a, c))


def exp_dotprod_sq(num_users, mean, cov_useritems, cov_latents, i, j):
    '''E[(U_i^T V_j)^2] = Var + E^2                    (matrix_normal_exps_cy.pyx:126-154)'''
    n = int(num_users)
    m = mean.shape[0] - n
    d = mean.shape[1]
    e, _ = _normal.mn_score(N.CRIT_APPROX_MEAN, mean, cov_useritems, cov_latents, n, m, d, [i], [j])
    v, _ = _normal.mn_score(N.CRIT_PRED_VARIANCE, mean, cov_useritems, cov_latents, n, m, d, [i], [j])
    return float(v[0] + e[0] ** 2)


def mn_kl_divergence(num_users, ratings, mean, cov_useritems, cov_latents, sigma_sq, sigma_u_sq,
                     sigma_v_sq):
    '''KL(MN(mean, cov_useritems, cov_latents) || PMF) up to a constant
    (matrix_normal_exps_cy.pyx:159-216, including its two quirks)'''
    n = int(num_users)
    p = _normal.fit_params(n, mean.shape[0] - n, mean.shape[1], sigma_sq, sigma_u_sq, sigma_v_sq)
    batch = _normal.MnBatch(ratings, p, mean[None], cov_useritems[None], cov_latents[None])
    return float(batch.kl_divergence()[0])


def matrixnormal_gradient(mn_apmf):
    '''(matrix_normal_exps_cy.pyx:219-247) -> (g_mean, g_cov_useritems, g_cov_latents)'''
    if mn_apmf is None:
        raise TypeError("Argument 'mn_apmf' must not be None")
    if mn_apmf.mean is None or mn_apmf.cov_useritems is None or mn_apmf.cov_latents is None:
        raise TypeError("mean, cov are None; run initialize_approx first")
    p = _normal.fit_params(mn_apmf.num_users, mn_apmf.num_items, mn_apmf.latent_d,
                           mn_apmf.sigma_sq, mn_apmf.sigma_u_sq, mn_apmf.sigma_v_sq)
    batch = _normal.MnBatch(mn_apmf.ratings, p, mn_apmf.mean[None], mn_apmf.cov_useritems[None],
                            mn_apmf.cov_latents[None])
    gm, gs, go = batch.gradient()
    return gm[0], gs[0], go[0]
