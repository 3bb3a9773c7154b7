"""CPU restatement of the SCALABLE-MODE variational posterior and lookahead -- TEST
INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg; never by the product package).

The reference's Gaussian approximation is a full k x k covariance, k = (N+M)d
(python-pmf/active_pmf.py:136,190-288); it minimises

    KL(mean, cov) = [sum_r E(U_i.V_j)^2 - 2 r E(U_i.V_j) + r^2] / (2 sigma^2)
                    + (|m_u|^2 + tr cov_uu) / (2 sigma_u^2) + (|m_v|^2 + tr cov_vv) / (2 sigma_v^2)
                    - 1/2 log det cov                                   (active_pmf.py:202-240)

which cannot be stored beyond toy sizes ("Full-covariance posterior does
not scale").  Scalable mode is THE SAME objective restricted to block-diagonal covariances --
one d x d block A_i per user row and B_j per item column.  On that family the objective has
closed-form coordinate minimisers (set the derivative w.r.t. one row's (mean, block) to zero):

    Lambda_i = I / sigma_u^2 + sum_{j in rated(i)} (n_j n_j^T + B_j) / sigma^2
    h_i      = sum_{j in rated(i)} r_ij n_j / sigma^2
    A_i = Lambda_i^-1,   m_i = A_i h_i                      (and symmetrically for items)

`kl_blocks` evaluates the reference's KL at such a posterior; tests check it against the
reference's own `kl_divergence` on the embedded k x k matrix and check stationarity with the
reference's `normal_gradient` (tests/test_oracle_blocks.py).

Lookahead (`_exp_with_rij`, active_pmf.py:635-704) adds one rating (i, j, v) and re-fits; here
the re-fit is `rounds` coordinate updates of the two blocks the new rating touches (row i, then
column j), everything else held fixed -- a rank-d update of each d x d precision and a Cholesky
per update.  The criteria are the reference's: `_approx_entropy` = log det cov = sum of the
block log-determinants (:526-530) and `_total_variance` = sum over all cells of
Var[U_i.V_j] (:605-606, normal_exps_cy.pyx:111-135 with a zero cross block), which for a block
posterior is <SA, SB + SNN> + <SMM, SB> with SA = sum A_i, SMM = sum m_i m_i^T, SB, SNN alike.
"""
import numpy as np
from scipy import stats


class Blocks(object):
    """means (n,d)/(m,d); covariance blocks A (n,d,d), B (m,d,d); their precisions and the
    natural-parameter vectors h = Lambda @ mean."""

    def __init__(self, mu, mv, A, B, Lu, Lv, hu, hv, sigma_sq, sigma_u_sq, sigma_v_sq):
        self.mu, self.mv, self.A, self.B = mu, mv, A, B
        self.Lu, self.Lv, self.hu, self.hv = Lu, Lv, hu, hv
        self.sigma_sq, self.sigma_u_sq, self.sigma_v_sq = sigma_sq, sigma_u_sq, sigma_v_sq

    def embed(self):
        """(mean (k,), cov (k,k)) in the reference's layout (active_pmf.py:136-142)"""
        from scipy.linalg import block_diag
        mean = np.hstack((self.mu.ravel(), self.mv.ravel()))
        return mean, block_diag(*(list(self.A) + list(self.B)))


def _half(R_rows, R_cols, R_vals, rows, other_mean, other_cov, prior_var, sigma_sq, cov_term):
    d = other_mean.shape[1]
    L = np.tile(np.eye(d) / prior_var, (rows, 1, 1))
    h = np.zeros((rows, d))
    for i, j, r in zip(R_rows, R_cols, R_vals):
        nj = other_mean[j]
        L[i] += np.outer(nj, nj) / sigma_sq
        if cov_term:
            L[i] += other_cov[j] / sigma_sq
        h[i] += r * nj / sigma_sq
    cov = np.linalg.inv(L)
    mean = np.einsum("nkl,nl->nk", cov, h)
    return L, h, cov, mean


def fit_blocks(R, n, m, d, users, items, sigma_sq=1., sigma_u_sq=10., sigma_v_sq=10., sweeps=50,
               tol=1e-10, cov_term=True, update_mean=True, trace=None):
    """Coordinate descent on the block-restricted KL: all user rows given the items' posterior,
    then all item columns given the users', `sweeps` times or until the means move by < tol.
    cov_term=False, update_mean=False, sweeps=1 is the cheap variant: curvature of the MAP
    objective at the MAP factors (means stay the MAP factors)."""
    ri, rj, rr = R[:, 0].astype(int), R[:, 1].astype(int), R[:, 2].astype(float)
    mu, mv = np.array(users, float), np.array(items, float)
    A = np.zeros((n, d, d))
    B = np.zeros((m, d, d))
    for _ in range(sweeps):
        Lu, hu, A, mu_new = _half(ri, rj, rr, n, mv, B, sigma_u_sq, sigma_sq, cov_term)
        if not update_mean:
            mu_new = mu
        Lv, hv, B, mv_new = _half(rj, ri, rr, m, mu_new, A, sigma_v_sq, sigma_sq, cov_term)
        if not update_mean:
            mv_new = mv
        move = max(np.abs(mu_new - mu).max(), np.abs(mv_new - mv).max())
        mu, mv = mu_new, mv_new
        st = Blocks(mu, mv, A, B, Lu, Lv, hu, hv, sigma_sq, sigma_u_sq, sigma_v_sq)
        if trace is not None:
            trace.append(kl_blocks(st, R))
        if update_mean and move < tol:
            break
    return st


def pred_mean_var(st, ii, jj):
    """E and Var of U_i.V_j under the block posterior: the closed form of SURVEY.md 8d row S2
    with C = 0 (normal_exps_cy.pyx:111-135, active_pmf.py:392-400,502-524)."""
    mu, mv, A, B = st.mu[ii], st.mv[jj], st.A[ii], st.B[jj]
    mean = np.einsum("nk,nk->n", mu, mv)
    var = (np.einsum("nkl,nkl->n", A, B) + np.einsum("nk,nkl,nl->n", mv, A, mv) +
           np.einsum("nk,nkl,nl->n", mu, B, mu))
    return mean, var


def kl_blocks(st, R):
    """the reference's KL (active_pmf.py:202-240) at the block posterior"""
    ri, rj, rr = R[:, 0].astype(int), R[:, 1].astype(int), R[:, 2].astype(float)
    mean, var = pred_mean_var(st, ri, rj)
    div = ((var + mean ** 2) - 2 * rr * mean + rr ** 2).sum() / (2 * st.sigma_sq)
    div += ((st.mu ** 2).sum() + np.trace(st.A, axis1=1, axis2=2).sum()) / (2 * st.sigma_u_sq)
    div += ((st.mv ** 2).sum() + np.trace(st.B, axis1=1, axis2=2).sum()) / (2 * st.sigma_v_sq)
    div -= (np.linalg.slogdet(st.A)[1].sum() + np.linalg.slogdet(st.B)[1].sum()) / 2
    return div


def entropy(st):
    """_approx_entropy (active_pmf.py:526-530): log det of the block-diagonal covariance"""
    return np.linalg.slogdet(st.A)[1].sum() + np.linalg.slogdet(st.B)[1].sum()


def total_variance(st):
    """_total_variance (active_pmf.py:605-606) over ALL cells, through the four d x d sums"""
    SA, SB = st.A.sum(0), st.B.sum(0)
    SMM, SNN = st.mu.T @ st.mu, st.mv.T @ st.mv
    return (SA * (SB + SNN)).sum() + (SMM * SB).sum()


def refit_pair(st, i, j, v, rounds=1):
    """posterior of row i and column j after adding the rating (i, j, v): `rounds` x (row i
    given column j, column j given row i).  Returns (m_i', A_i', n_j', B_j')."""
    s2 = st.sigma_sq
    nj, Bj = st.mv[j], st.B[j]
    for _ in range(rounds):
        Li = st.Lu[i] + (np.outer(nj, nj) + Bj) / s2
        Ai = np.linalg.inv(Li)
        mi = Ai @ (st.hu[i] + v * nj / s2)
        Lj = st.Lv[j] + (np.outer(mi, mi) + Ai) / s2
        Bj = np.linalg.inv(Lj)
        nj = Bj @ (st.hv[j] + v * mi / s2)
    return mi, Ai, nj, Bj


def lookahead_evals(st, ii, jj, values, what, rounds=1):
    """fn(model + (i, j, v)) for every candidate and value: (ncand, nvalues); what is 'entropy'
    or 'total_variance' (the `fn` of active_pmf.py:669-676)."""
    values = np.asarray(values, float)
    if values.ndim == 1:
        values = np.broadcast_to(values, (len(ii), len(values)))
    out = np.empty(values.shape)
    H0 = entropy(st)
    SA, SB = st.A.sum(0), st.B.sum(0)
    SMM, SNN = st.mu.T @ st.mu, st.mv.T @ st.mv
    ldA, ldB = np.linalg.slogdet(st.A)[1], np.linalg.slogdet(st.B)[1]
    for c, (i, j) in enumerate(zip(ii, jj)):
        for q, v in enumerate(values[c]):
            mi, Ai, nj, Bj = refit_pair(st, i, j, v, rounds)
            if what == 'entropy':
                out[c, q] = H0 + np.linalg.slogdet(Ai)[1] - ldA[i] + np.linalg.slogdet(Bj)[1] - ldB[j]
            else:
                SA2 = SA + Ai - st.A[i]
                SB2 = SB + Bj - st.B[j]
                SMM2 = SMM + np.outer(mi, mi) - np.outer(st.mu[i], st.mu[i])
                SNN2 = SNN + np.outer(nj, nj) - np.outer(st.mv[j], st.mv[j])
                out[c, q] = (SA2 * (SB2 + SNN2)).sum() + (SMM2 * SB2).sum()
    return out


def rij_distribution(st, ii, jj, use_map, users=None, items=None):
    """mean and variance assumed for R_ij (active_pmf.py:656-666)"""
    if use_map:
        return np.einsum("nk,nk->n", users[ii], items[jj]), np.full(len(ii), float(st.sigma_sq))
    return pred_mean_var(st, ii, jj)


def discrete_weights(rating_values, mu, var):
    """Delta-cdf at the rating bounds (active_pmf.py:171-185, :687-689)"""
    vals = np.array(sorted(rating_values), float)
    edges = np.empty(len(vals) + 2)
    edges[0], edges[-1] = -np.inf, np.inf
    edges[1:-1] = vals
    bounds = (edges[1:] + edges[:-1]) / 2
    cdfs = stats.norm.cdf(bounds[None, :], loc=np.asarray(mu)[:, None],
                          scale=np.sqrt(np.asarray(var))[:, None])
    return vals, np.diff(cdfs, axis=1)


# 2-sigma window of active_pmf.py:691-699 integrated with fixed Gauss-Legendre nodes:
# est = int_{-2}^{2} f(mu + sigma t) phi(t) dt
def gauss_nodes(nq=16):
    t, w = np.polynomial.legendre.leggauss(nq)
    t, w = 2 * t, 2 * w
    return t, w * stats.norm.pdf(t)


def lookahead(st, ii, jj, what, use_map=True, rating_values=None, users=None, items=None,
              rounds=1, nq=16):
    """E_v[fn(model + (i, j, v))] per candidate: discrete values weighted by Delta-cdf, or (no
    rating_values) the 2-sigma window with `nq` Gauss-Legendre nodes."""
    mu, var = rij_distribution(st, ii, jj, use_map, users, items)
    if rating_values:
        vals, w = discrete_weights(rating_values, mu, var)
        ev = lookahead_evals(st, ii, jj, vals, what, rounds)
    else:
        t, w1 = gauss_nodes(nq)
        vals = mu[:, None] + np.sqrt(var)[:, None] * t[None, :]
        w = np.broadcast_to(w1, vals.shape)
        ev = lookahead_evals(st, ii, jj, vals, what, rounds)
    return (ev * w).sum(1)


# ------------------------------------------------------------------------------------------
# Converged full-covariance optimum of the same KL (validation target for the block family).
# The reference's own optimiser (projected gradient steps of 1e-4 from a random covariance,
# stopping at a KL gain below .005, active_pmf.py:251-288) stops far from the optimum: on the
# 6x7 fixture it ends at KL = 80.5 while the block-restricted optimum reaches 14.1.  To compare
# the block posterior with what exact mode is aiming at, the stationary point is computed here
# by the variational-Gaussian fixed point (Bonnet / Price):
#     cov^-1 = E_q[Hessian of the energy],   E_q[gradient of the energy] = 0
# with damped updates; tests check the result with the REFERENCE's kl_divergence /
# normal_gradient.
# ------------------------------------------------------------------------------------------
def _exp_grad_hess(R, n, m, d, mean, cov, sigma_sq, sigma_u_sq, sigma_v_sq):
    k = (n + m) * d
    g = np.zeros(k)
    H = np.zeros((k, k))
    nu = n * d
    g[:nu] = mean[:nu] / sigma_u_sq
    g[nu:] = mean[nu:] / sigma_v_sq
    H[np.arange(nu), np.arange(nu)] = 1 / sigma_u_sq
    H[np.arange(nu, k), np.arange(nu, k)] = 1 / sigma_v_sq
    for i, j, r in R:
        i, j = int(i), int(j)
        us = slice(i * d, (i + 1) * d)
        vs = slice(nu + j * d, nu + (j + 1) * d)
        mu, mv = mean[us], mean[vs]
        A, Bv, Cm = cov[us, us], cov[vs, vs], cov[us, vs]          # Cm[k,l] = Cov(u_k, v_l)
        Ee = mu @ mv + np.trace(Cm) - r
        # E[e v] = E[(u.v) v] - r n ;  E[(u.v) v_k] = sum_l E[u_l v_l v_k]
        Euvv = (mu * mv).sum() * mv + Bv @ mu + Cm.T @ mv + np.trace(Cm) * mv
        Euvu = (mu * mv).sum() * mu + A @ mv + Cm @ mu + np.trace(Cm) * mu
        g[us] += (Euvv - r * mv) / sigma_sq
        g[vs] += (Euvu - r * mu) / sigma_sq
        H[us, us] += (np.outer(mv, mv) + Bv) / sigma_sq
        H[vs, vs] += (np.outer(mu, mu) + A) / sigma_sq
        X = (np.outer(mv, mu) + Cm.T + Ee * np.eye(d)) / sigma_sq  # d2/du_k dv_l = v_k u_l + e delta
        H[us, vs] += X
        H[vs, us] += X.T
    return g, H


def kl_full(R, n, m, d, mean, cov, sigma_sq=1., sigma_u_sq=10., sigma_v_sq=10.):
    """the reference's KL (active_pmf.py:202-240) at a full (mean, cov), by the closed-form
    second moment E(U_i.V_j)^2 = Var + E^2 (pmf_oracle.pred_mean_var_closed)"""
    nu = n * d
    acc = 0.0
    for i, j, r in R:
        i, j = int(i), int(j)
        us, vs = slice(i * d, (i + 1) * d), slice(nu + j * d, nu + (j + 1) * d)
        mu, mv, A, Bv, Cm = mean[us], mean[vs], cov[us, us], cov[vs, vs], cov[us, vs]
        e = mu @ mv + np.trace(Cm)
        var = (A * Bv).sum() + (Cm * Cm.T).sum() + mv @ A @ mv + mu @ Bv @ mu + 2 * (mu @ Cm.T @ mv)
        acc += var + e * e - 2 * r * e + r * r
    dg = np.diag(cov)
    div = acc / (2 * sigma_sq)
    div += ((mean[:nu] ** 2).sum() + dg[:nu].sum()) / (2 * sigma_u_sq)
    div += ((mean[nu:] ** 2).sum() + dg[nu:].sum()) / (2 * sigma_v_sq)
    return div - np.linalg.slogdet(cov)[1] / 2


def fit_full_converged(R, n, m, d, mean, cov, sigma_sq=1., sigma_u_sq=10., sigma_v_sq=10.,
                       iters=400, tol=1e-9, kl=None):
    """stationary point of the full-covariance KL reached from (mean, cov); `kl` is a callable
    (mean, cov) -> objective used for the damping line search"""
    R = np.asarray(R, float)
    if kl is None:
        kl = lambda a, b: kl_full(R, n, m, d, a, b, sigma_sq, sigma_u_sq, sigma_v_sq)  # noqa: E731
    P = np.linalg.inv(cov)
    cur = kl(mean, cov)
    for _ in range(iters):
        g, H = _exp_grad_hess(R, n, m, d, mean, cov, sigma_sq, sigma_u_sq, sigma_v_sq)
        step = 1.0
        while step > 1e-6:
            P2 = (1 - step) * P + step * H
            try:
                np.linalg.cholesky(P2)
            except np.linalg.LinAlgError:
                step /= 2
                continue
            cov2 = np.linalg.inv(P2)
            cov2 = (cov2 + cov2.T) / 2
            mean2 = mean - step * cov2 @ g
            new = kl(mean2, cov2)
            if new <= cur:
                break
            step /= 2
        else:
            break
        gain = cur - new
        mean, cov, P, cur = mean2, cov2, P2, new
        if gain < tol:
            break
    return mean, cov, cur
