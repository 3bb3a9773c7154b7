"""Host mirror of the reference's ``mn_active_pmf`` module (python-pmf/mn_active_pmf.py):
active learning on PMF with a MATRIX-NORMAL approximate posterior
MN(mean, cov_useritems, cov_latents) -- the variant the reference's drugbank and movielens
experiments run (results/drugbank-94x425/Makefile:66-76).  SURVEY.md 8f-1.

Same class surface as the reference (``MNActivePMF``, the 13 ``KEY_FUNCS``; its drivers are run from
the reference's own source by ``drivers.load``); the numerics run in csrc/mn.cu: the whole ``fit_normal`` line search of
every (candidate, value) lookahead problem is one CTA, and the cheap criteria read three
scalars of Sigma plus Omega per candidate.  No CPU path.
"""
from copy import deepcopy
import itertools
import operator

import numpy as np

from . import _native as N
from . import normal as _normal
from . import active_pmf as _apmf
from .active_pmf import (ActivePMF,  # noqa: F401
                         do_normal_fit, spawn_processes, nice_name, minimize, maximize, strictmap)
from .pmf_cy import parse_fit_type
from .matrix_normal_exps_cy import (quadexpect, exp_a2bc, exp_dotprod_sq,  # noqa: F401
                                    mn_kl_divergence, matrixnormal_gradient)


def project_psd(mat, min_eig=0, destroy=False):
    '''(mn_active_pmf.py:42-67) symmetrise and clamp the spectrum, on the device'''
    return _normal.project_psd_device(np.array(mat, dtype=float), float(min_eig))


class MNActivePMFEvaluator(_apmf.ActivePMFEvaluator):
    pass


class MNActivePMF(ActivePMF):
    # problems with more than this many rows in Sigma use the host-driven "wide" fit (cuSOLVER
    # for the (N+M)^2 algebra) instead of the one-CTA-per-problem kernel
    wide_threshold = 96
    max_normal_steps = 0      # > 0 caps the accepted steps of one fit_normal (0: to convergence)

    def __init__(self, rating_tuples, latent_d=1, rating_values=None,
                 discrete_expectations=False, refit_lookahead=False, knowable=None,
                 fit_type=('batch',)):
        super(MNActivePMF, self).__init__(rating_tuples, latent_d=latent_d,
                                          rating_values=rating_values,
                                          discrete_expectations=discrete_expectations,
                                          refit_lookahead=refit_lookahead, knowable=knowable,
                                          fit_type=fit_type)
        # the matrix-normal parameters replace the full covariance of the parent class
        for name in ('cov', 'u', 'v', 'num_params'):
            self.__dict__.pop(name, None)
        self.mean = None
        self.cov_useritems = None
        self.cov_latents = None

    def __copy__(self):
        res = MNActivePMF(self.ratings, self.latent_d, self.rating_values, self.discrete_expectations)
        res.__setstate__(self.__getstate__())
        return res

    def __deepcopy__(self, memodict):
        res = MNActivePMF(self.ratings, self.latent_d, self.rating_values, self.discrete_expectations)
        res.__setstate__(deepcopy(self.__getstate__(), memodict))
        return res

    ############################################################################
    ### The approximation

    def initialize_approx(self, random_cov=False):
        '''(mn_active_pmf.py:202-219): mean <- MAP factors, identity (or random) covariances'''
        self.mean = np.vstack((self.users, self.items))
        nui = self.num_users + self.num_items
        if random_cov:
            a = np.random.normal(size=(nui, nui))
            b = np.random.normal(size=(self.latent_d, self.latent_d))
            self.cov_useritems = np.dot(a, a.T)
            self.cov_latents = np.dot(b, b.T)
        else:
            self.cov_useritems = np.eye(nui)
            self.cov_latents = np.eye(self.latent_d)

    def _require_approx(self):
        if self.mean is None or self.cov_useritems is None or self.cov_latents is None:
            raise ValueError("run initialize_approx first")

    def kl_divergence(self, mean=None, cov_useritems=None, cov_latents=None):
        '''KL(approximation || PMF model), up to an additive constant (mn_active_pmf.py:221-231)'''
        mean = self.mean if mean is None else mean
        cov_useritems = self.cov_useritems if cov_useritems is None else cov_useritems
        cov_latents = self.cov_latents if cov_latents is None else cov_latents
        if mean is None or cov_useritems is None or cov_latents is None:
            raise ValueError("run initialize_approx first")
        return mn_kl_divergence(self.num_users, self.ratings, mean, cov_useritems, cov_latents,
                                self.sigma_sq, self.sigma_u_sq, self.sigma_v_sq)

    def fit_normal_kls(self):
        '''(mn_active_pmf.py:242-288) one launch; the KL of each accepted step is yielded after'''
        self._require_approx()
        if self.num_users + self.num_items > self.wide_threshold:
            wide = _normal.MnWide(self.ratings, self._fit_params(), self.mean, self.cov_useritems,
                                  self.cov_latents)
            mean, sig, om, kls = wide.fit(max_steps=self.max_normal_steps)
            if kls:
                self.mean, self.cov_useritems, self.cov_latents = mean, sig, om
            for kl in kls:
                yield float(kl)
            return
        batch = _normal.MnBatch(self.ratings, self._fit_params(max_steps=self.max_normal_steps),
                                self.mean[None], self.cov_useritems[None], self.cov_latents[None])
        trace_len = 1 << 16
        res = batch.fit(trace_len=trace_len)
        steps = int(res['steps'][0])
        if steps > 0:
            mean, sig, om = batch.state()
            self.mean, self.cov_useritems, self.cov_latents = mean[0], sig[0], om[0]
        for kl in res['trace'][0][:min(steps, trace_len)]:
            yield float(kl)

    def mean_meandiff(self):
        return np.abs(self.mean - np.vstack((self.users, self.items))).mean()

    def _normal_scores(self, criterion, ii, jj, cutoff=0., maximize_=True):
        self._require_approx()
        return _normal.mn_score(criterion, self.mean, self.cov_useritems, self.cov_latents,
                                self.num_users, self.num_items, self.latent_d, ii, jj, "f64",
                                cutoff=cutoff, maximize=maximize_)

    def approx_pred_covs(self):
        raise NotImplementedError("approx_pred_covs is not implemented for the matrix-normal "
                                  "approximation (commented out in mn_active_pmf.py:332-404)")

    def _approx_entropy(self):
        '''(mn_active_pmf.py:513-521)'''
        ui_sign, ui_logdet = _apmf._slogdet(self.cov_useritems)
        l_sign, l_logdet = _apmf._slogdet(self.cov_latents)
        assert ui_sign == 1
        assert l_sign == 1
        return 0.5 * (self.latent_d * ui_logdet + (self.num_users + self.num_items) * l_logdet)

    def _pred_entropy_bound(self):
        raise NotImplementedError("pred-entropy-bound is not available for MNActivePMF "
                                  "(mn_active_pmf.py:550-595 is commented out)")

    ############################################################################
    ### Batched lookahead (mn_active_pmf.py:627-697)

    def _refits(self, pairs_vals, what):
        self._require_approx()
        B = len(pairs_vals)
        if B == 0:
            return np.zeros(0)
        if self.refit_lookahead:
            return np.array([self._refit_one_host_driven(i, j, v, what) for i, j, v in pairs_vals])
        if what == 'pred_entropy_bound':
            self._pred_entropy_bound()
        if self.num_users + self.num_items > self.wide_threshold:
            # large Sigma: each re-fit is one wide (host-driven) fit
            return np.array([self._refit_one_wide(i, j, v, what) for i, j, v in pairs_vals])
        ei = np.array([p[0] for p in pairs_vals], dtype=np.int32)
        ej = np.array([p[1] for p in pairs_vals], dtype=np.int32)
        er = np.array([p[2] for p in pairs_vals], dtype=np.float64)
        nui, d = self.num_users + self.num_items, self.latent_d
        per = 8 * (6 * nui * nui + 3 * nui * d + 6 * d * d)
        chunk = max(1, min(B, int(6e9 // per)))
        out = np.empty(B)
        for s in range(0, B, chunk):
            e = min(B, s + chunk)
            nb = e - s
            batch = _normal.MnBatch(self.ratings, self._fit_params(),
                                    np.broadcast_to(self.mean, (nb, nui, d)),
                                    np.broadcast_to(self.cov_useritems, (nb, nui, nui)),
                                    np.broadcast_to(self.cov_latents, (nb, d, d)),
                                    extra=(ei[s:e], ej[s:e], er[s:e]))
            res = batch.fit(want_entropy=(what == 'entropy'), want_totvar=(what == 'total_variance'))
            if what == 'entropy':
                out[s:e] = res['entropy']
            elif what == 'total_variance':
                out[s:e] = res['total_variance']
            else:
                mean, sig, om = batch.state()
                for b in range(nb):
                    out[s + b] = self._criterion_on((mean[b], sig[b], om[b]), None, what,
                                                    (int(ei[s + b]), int(ej[s + b])), er[s + b])
        return out

    def _refit_one_wide(self, i, j, v, what):
        r2 = np.append(self.ratings, [[i, j, v]], 0)
        wide = _normal.MnWide(r2, self._fit_params(), self.mean, self.cov_useritems, self.cov_latents)
        mean, sig, om, _ = wide.fit(max_steps=self.max_normal_steps)
        if what == 'entropy':
            su, lu = _apmf._slogdet(sig)
            sl, ll = _apmf._slogdet(om)
            return 0.5 * (self.latent_d * lu + (self.num_users + self.num_items) * ll)
        if what == 'total_variance':
            ii, jj = self._all_cells()
            var, _ = _normal.mn_score(N.CRIT_PRED_VARIANCE, mean, sig, om, self.num_users,
                                      self.num_items, self.latent_d, ii, jj)
            return float(var.sum())
        return self._criterion_on((mean, sig, om), None, what, (i, j), v)

    def _criterion_on(self, state, _unused, what, ij, v):
        mean, sig, om = state
        n, m, d = self.num_users, self.num_items, self.latent_d
        if isinstance(what, tuple) and what[0] == 'fn':
            apmf = deepcopy(self)
            apmf.add_rating(ij[0], ij[1], v)
            apmf.mean, apmf.cov_useritems, apmf.cov_latents = mean, sig, om
            return what[1](apmf, v=v) if what[2] else what[1](apmf)
        if isinstance(what, tuple) and what[0] == 'onestep':
            cutoff = what[1]
            pool = [c for c in self.unrated if c != ij]
            if not pool:
                raise ValueError("max() arg is an empty sequence")
            ii, jj = zip(*pool)
            _, (best, _i) = _normal.mn_score(N.CRIT_PROB_GE, mean, sig, om, n, m, d, ii, jj,
                                             cutoff=cutoff)
            return int(v >= cutoff) + best
        raise ValueError("unknown lookahead quantity %r" % (what,))

    _FN_NAMES = {'_approx_entropy': 'entropy', '_total_variance': 'total_variance'}


################################################################################
### Registry (mn_active_pmf.py:1005-1025).  The drivers of mn_active_pmf.py:785-1132 are run from
### the reference's own source by drivers.load("mn_active_pmf", ref_dir).

KEY_FUNCS = {
    "random": MNActivePMF.random_weighting,
    "pred-variance": MNActivePMF.pred_variance,

    "total-variance": MNActivePMF.exp_total_variance,
    "total-variance-approx": MNActivePMF.exp_total_variance_byapprox,

    "uv-entropy": MNActivePMF.exp_approx_entropy,
    "uv-entropy-approx": MNActivePMF.exp_approx_entropy_byapprox,

    "pred": MNActivePMF.pred,
    "prob-ge-3.5": MNActivePMF.prob_ge_3_5,
    "prob-ge-.5": MNActivePMF.prob_ge_half,

    "1step-ge-3.5": MNActivePMF.onestep_ge_3_5,
    "1step-ge-3.5-approx": MNActivePMF.onestep_ge_3_5_approx,

    "1step-ge-.5": MNActivePMF.onestep_ge_half,
    "1step-ge-.5-approx": MNActivePMF.onestep_ge_half_approx,
}
