"""Host mirror of the reference's ``active_pmf`` module (python-pmf/active_pmf.py).

Same names and conventions -- ``ActivePMF``, the criterion methods with their
``do_normal_fit / spawn_processes / nice_name / chooser`` attributes, ``KEY_FUNCS``,
``pick_query_point / _get_key_vals / get_key_evals`` (the experiment drivers of the same file
are run from the reference's own source by ``drivers.load``) -- but the pool of candidates is evaluated in batched GPU launches instead of a Python map over
``multiprocessing.Pool`` workers:

* cheap criteria (pred, prob-ge-*, pred-variance): one scoring launch with a fused arg-best;
* lookahead criteria (``_exp_with_rij``, active_pmf.py:635-704): every (candidate, value)
  pair is an independent variational re-fit; the whole batch runs as ONE launch with one CTA
  per problem (csrc/normal.cu), replacing deepcopy + host ``fit_normal`` per pair.

There is no CPU path: every numeric method raises if libamf_b200 or the GPU is missing.
"""
from copy import deepcopy
import functools
import os
import itertools
import math
import numbers
import operator
import random
import warnings

import numpy as np
from scipy import stats
import scipy.integrate
import torch

from . import _native as N
from . import blocks as _blocks
from . import device as _D
from . import normal as _normal
from . import scoring as _scoring
from .pmf_cy import ProbabilisticMatrixFactorization, parse_fit_type
from .normal_exps_cy import (quadexpect, exp_a2bc, exp_dotprod_sq,  # noqa: F401 (API parity)
                             normal_gradient)


################################################################################
### Helpers

def project_psd(mat, min_eig=0):
    '''
    Symmetrise `mat` and clamp its spectrum from below at `min_eig`
    (active_pmf.py:36-50); the eigendecomposition is the parallel Jacobi of csrc/normal.cu.
    '''
    return _normal.project_psd_device(np.asarray(mat, dtype=float), float(min_eig))


class ActivePMFEvaluator(object):
    '''Kept for API parity (active_pmf.py:54-67): evaluates one criterion for one pair.'''
    def __init__(self, apmf, key):
        self.apmf = apmf
        self.key_name = key.__name__

    def __call__(self, ij):
        return getattr(self.apmf, self.key_name)(ij)


def strictmap(*args, **kwargs):
    return list(map(*args, **kwargs))


# decorators carrying the criterion metadata the drivers read (active_pmf.py:73-96)
def do_normal_fit(val):
    def decorator(f):
        f.do_normal_fit = val
        return f
    return decorator


def spawn_processes(val):
    def decorator(f):
        f.spawn_processes = val
        return f
    return decorator


def nice_name(name):
    def decorator(f):
        f.nice_name = name
        return f
    return decorator


def minimize(f):
    f.chooser = min
    return f


def maximize(f):
    f.chooser = max
    return f


def _criterion(name, normal_fit, spawn, chooser):
    def decorator(f):
        do_normal_fit(normal_fit)(f)
        spawn_processes(spawn)(f)
        nice_name(name)(f)
        (maximize if chooser is max else minimize)(f)
        return f
    return decorator


################################################################################
### Main code

def _pool_arrays(pool):
    """(i, j) int32 arrays of a list of index pairs; one pass over a flattened iterator is the
    cheapest way through the per-tuple Python cost that dominates scoring of large pools."""
    import itertools
    n = len(pool)
    try:
        flat = np.fromiter(itertools.chain.from_iterable(pool), dtype=np.int64, count=2 * n)
    except (TypeError, ValueError):          # pairs that are not plain integer 2-tuples
        ii, jj = zip(*pool)
        return np.asarray(ii, dtype=np.int32), np.asarray(jj, dtype=np.int32)
    flat = flat.reshape(n, 2)
    return flat[:, 0].astype(np.int32), flat[:, 1].astype(np.int32)


class ActivePMF(ProbabilisticMatrixFactorization):
    verbose_lookahead = False   # the reference prints one line per lookahead (active_pmf.py:702-703)
    max_normal_steps = 0        # > 0 caps the accepted steps of one fit_normal (0: to convergence)

    # Which family the Gaussian approximation lives in.  'exact': the reference's k x k
    # covariance, k = (N+M)d (active_pmf.py:136,190-288) -- bit-for-bit algorithm, feasible up to
    # a few hundred dimensions.  'blocks': the same KL objective restricted to one d x d block per
    # row / column (blocks.py, csrc/blocks.cu) -- O((N+M) d^2) state, lookahead by local re-fits.
    # 'auto' = exact while k <= exact_max_dim, blocks above (drugbank: k = 2595).
    approx_mode = os.environ.get("AMF_B200_APPROX", "auto")
    exact_max_dim = 320
    lookahead_rounds = 1        # coordinate rounds (row i, column j) of a scalable-mode re-fit
    blocks_max_sweeps = 500
    blocks_tol = 1e-10          # largest movement of a posterior mean that still counts as moving
    quadrature_nodes = 16       # Gauss-Legendre nodes of the 2-sigma window in scalable mode

    def __init__(self, rating_tuples, latent_d=1, rating_values=None,
                 discrete_expectations=False, refit_lookahead=False, knowable=None,
                 fit_type=('batch',)):
        super(ActivePMF, self).__init__(rating_tuples, latent_d=latent_d, subtract_mean=False,
                                        knowable=knowable, fit_type=fit_type)
        self.ratings = np.asarray(self.ratings, dtype=float)

        if rating_values is not None:
            rating_values = set(map(float, rating_values))
            if not rating_values.issuperset(self.ratings[:, 2]):
                raise ValueError("got ratings not in rating_values")

        self.rating_values = rating_values
        self.discrete_expectations = discrete_expectations
        self.refit_lookahead = refit_lookahead

        self.mean = None
        self.cov = None

        n, m, d = self.num_users, self.num_items, self.latent_d
        self.approx_dim = k = (n + m) * d
        self.num_params = k + k * (k + 1) / 2
        # positions of U_ki / V_kj in the k-vector (active_pmf.py:141-142)
        self.u = np.arange(0, n * d).reshape(n, d).T
        self.v = np.arange(n * d, (n + m) * d).reshape(m, d).T

        self.normal_learning_rate = 1e-4
        self.min_eig = 1e-5

    def __copy__(self):
        res = ActivePMF(self.ratings, self.latent_d, self.rating_values, self.discrete_expectations)
        res.__setstate__(self.__getstate__())
        return res

    def __deepcopy__(self, memodict):
        res = ActivePMF(self.ratings, self.latent_d, self.rating_values, self.discrete_expectations)
        res.__setstate__(deepcopy(self.__getstate__(), memodict))
        return res

    def __getstate__(self):
        state = super().__getstate__()
        state['__dict__'] = {k: v for k, v in self.__dict__.items()
                             if k not in ('_dev', '_users', '_items', '_ratings')}
        return state

    rating_values = property(lambda self: self._rating_values)
    rating_bounds = property(lambda self: self._rating_bounds)

    @rating_values.setter
    def rating_values(self, vals):
        if vals:
            vals = tuple(sorted(vals))
            self._rating_values = vals
            edges = np.empty(len(vals) + 2)
            edges[0], edges[-1] = -np.inf, np.inf
            edges[1:-1] = vals
            self._rating_bounds = (edges[1:] + edges[:-1]) / 2
        else:
            self._rating_values = None
            self._rating_bounds = None

    ############################################################################
    ### Normal approximation

    def _fit_params(self, max_steps=0):
        return _normal.fit_params(self.num_users, self.num_items, self.latent_d, self.sigma_sq,
                                  self.sigma_u_sq, self.sigma_v_sq,
                                  learning_rate=self.normal_learning_rate, min_eig=self.min_eig,
                                  max_steps=max_steps)

    def _use_blocks(self):
        mode = self.approx_mode
        if mode not in ('auto', 'exact', 'blocks'):
            raise ValueError("approx_mode must be 'auto', 'exact' or 'blocks'")
        return mode == 'blocks' or (mode == 'auto' and self.approx_dim > self.exact_max_dim)

    def _in_blocks(self, cov=None):
        return isinstance(getattr(self, 'cov', None) if cov is None else cov, _blocks.BlockDiagonal)

    def _new_block_posterior(self):
        return _blocks.BlockPosterior(self.num_users, self.num_items, self.latent_d, self.sigma_sq,
                                      self.sigma_u_sq, self.sigma_v_sq)

    def _block_posterior(self, cov=None):
        '''device tables of the BlockDiagonal in self.cov (cached until cov is replaced)'''
        cov = self.cov if cov is None else cov
        hit = self._dev.get('blocks')
        hyper = (self.sigma_sq, self.sigma_u_sq, self.sigma_v_sq)
        if hit is None or hit[0] is not cov or hit[2] != hyper:
            post = _blocks.BlockPosterior.from_host(cov, *hyper)
            hit = self._dev['blocks'] = (cov, post, hyper)
        return hit[1]

    def _adopt(self, post):
        bd = post.to_host()
        self.cov = bd
        self.mean = bd.stacked_mean()
        self._dev['blocks'] = (bd, post, (self.sigma_sq, self.sigma_u_sq, self.sigma_v_sq))

    def initialize_approx(self):
        '''(active_pmf.py:190-200): mean <- MAP factors, cov <- random PSD matrix.  Scalable
        mode: cov <- the block curvature of the MAP objective at the MAP factors (no k x k draw).'''
        if self._use_blocks():
            post = self._new_block_posterior()
            post.fit(self._rating_handle(), self.users, self.items, sweeps=1, cov_term=False,
                     update_mean=False)
            self._adopt(post)
            return
        self.mean = np.hstack((self.users.reshape(-1), self.items.reshape(-1)))
        s = np.random.normal(0, 2, (self.approx_dim, self.approx_dim))
        self.cov = project_psd(s, min_eig=self.min_eig)

    def _rating_arrays_device(self):
        r = np.asarray(self.ratings, dtype=float)
        return (_D.to_device(r[:, 0], np.int32), _D.to_device(r[:, 1], np.int32),
                _D.to_device(r[:, 2], np.float64))

    def kl_divergence(self, mean=None, cov=None):
        '''KL(PMF model || approximation), up to an additive constant (active_pmf.py:202-240)'''
        if mean is None:
            mean = self.mean
        if cov is None:
            cov = self.cov
        if mean is None or cov is None:
            raise ValueError("run initialize_approx first")
        if self._in_blocks(cov):
            return self._block_posterior(cov).kl(*self._rating_arrays_device())
        batch = _normal.NormalBatch(self.ratings, self._fit_params(), mean[None], cov[None])
        return float(batch.kl_divergence()[0])

    def fit_normal(self):
        if self._in_blocks():
            # no KL is asked for: the whole coordinate descent is one library call (amf_blocks_fit),
            # same sweeps and stopping rule as fit_normal_kls
            if self.mean is None or self.cov is None:
                raise ValueError("run initialize_approx first")
            post = self._new_block_posterior()
            n, d = self.num_users, self.latent_d
            try:
                post.fit(self._rating_handle(), self.mean[:n * d].reshape(n, d),
                         self.mean[n * d:].reshape(-1, d), sweeps=self.blocks_max_sweeps,
                         tol=self.blocks_tol)
            finally:
                self._adopt(post)
            return
        for _kl in self.fit_normal_kls():
            pass

    def fit_normal_kls(self):
        '''
        Line-search descent on the KL (active_pmf.py:251-288).  The whole search runs in one
        kernel launch; the KL after each accepted step is yielded afterwards.
        '''
        if self.mean is None or self.cov is None:
            raise ValueError("run initialize_approx first")
        if self._in_blocks():
            # coordinate descent on the block-restricted KL from the current means; every sweep
            # lowers the reference's objective, which is yielded like the accepted steps there
            post = self._new_block_posterior()
            n, d = self.num_users, self.latent_d
            ri, rj, rr = self._rating_arrays_device()
            try:
                for kl in post.fit_sweeps(self._rating_handle(), self.mean[:n * d].reshape(n, d),
                                          self.mean[n * d:].reshape(-1, d),
                                          sweeps=self.blocks_max_sweeps, tol=self.blocks_tol,
                                          kl_of=lambda p: p.kl(ri, rj, rr)):
                    yield kl
            finally:
                self._adopt(post)
            return
        batch = _normal.NormalBatch(self.ratings, self._fit_params(max_steps=self.max_normal_steps),
                                    self.mean[None], self.cov[None])
        trace_len = 1 << 14
        res = batch.fit(trace_len=trace_len)
        steps = int(res['steps'][0])
        if steps > 0:
            self.mean = batch.means()[0]
            self.cov = batch.covs()[0]
        for kl in res['trace'][0][:min(steps, trace_len)]:
            yield float(kl)

    ############################################################################
    ### Quantities under the current approximation

    def _require_approx(self):
        if self.mean is None or self.cov is None:
            raise ValueError("run initialize_approx first")

    def mean_meandiff(self):
        p = np.hstack((self.users.reshape(-1), self.items.reshape(-1)))
        return np.abs(self.mean - p).mean()

    def _all_cells(self):
        ii, jj = np.meshgrid(np.arange(self.num_users), np.arange(self.num_items), indexing='ij')
        return ii.reshape(-1), jj.reshape(-1)

    def _normal_scores(self, criterion, ii, jj, cutoff=0., maximize_=True):
        self._require_approx()
        if self._in_blocks():
            ci, cj = _scoring._cands(ii, jj)
            scores, best = self._block_posterior().score(criterion, ci, cj, "f64", cutoff=cutoff,
                                                         maximize=maximize_)
            return scores.cpu().numpy(), _scoring.unpack_best(best)
        return _scoring.score_normal(criterion, self.mean, self.cov, self.num_users,
                                     self.num_items, self.latent_d, ii, jj, "f64",
                                     cutoff=cutoff, maximize=maximize_)

    def approx_pred_means_vars(self):
        '''(active_pmf.py:301-322) mean and variance of every predicted cell'''
        ii, jj = self._all_cells()
        shape = (self.num_users, self.num_items)
        mn, _ = self._normal_scores(N.CRIT_APPROX_MEAN, ii, jj)
        var, _ = self._normal_scores(N.CRIT_PRED_VARIANCE, ii, jj)
        return mn.reshape(shape), var.reshape(shape)

    def approx_pred_mean_var(self, i, j):
        '''(active_pmf.py:392-400)'''
        mn, _ = self._normal_scores(N.CRIT_APPROX_MEAN, [i], [j])
        var, _ = self._normal_scores(N.CRIT_PRED_VARIANCE, [i], [j])
        return float(mn[0]), float(var[0])

    def approx_pred_covs(self):
        '''(active_pmf.py:324-390) covariance between all pairs of predicted cells'''
        self._require_approx()
        return _pred_covs(self.mean, self.cov, self.num_users, self.num_items, self.latent_d)

    ############################################################################
    ### Criteria -- each is usable on one pair; pools go through _get_key_vals

    @_criterion("Random", False, False, max)
    def random_weighting(self, ij):
        return random.random()

    @_criterion("Pred Mag", False, False, max)
    def pred(self, ij):
        '''The MAP estimate of R_ij (active_pmf.py:416-421).'''
        return self._get_key_vals([ij], ActivePMF.pred, 1, None)[0]

    def _prob_ge_cutoff(self, ij, cutoff):
        '''norm.sf(cutoff, loc=mean, scale=var) -- the reference passes the variance as the
        scale (active_pmf.py:432-439); reproduced.'''
        vals, _ = self._normal_scores(N.CRIT_PROB_GE, [ij[0]], [ij[1]], cutoff=cutoff)
        return float(vals[0])

    @_criterion("Prob >= 3.5", True, False, max)
    def prob_ge_3_5(self, ij):
        return self._prob_ge_cutoff(ij, 3.5)

    @_criterion("Prob >= .5", True, False, max)
    def prob_ge_half(self, ij):
        return self._prob_ge_cutoff(ij, .5)

    def _onestep_ge_cutoff(self, ij, cutoff, use_map):
        '''One-step lookahead utility (active_pmf.py:460-474); always discretised.'''
        return self._lookahead([ij], ('onestep', cutoff), use_map, discretize=True)[0]

    @_criterion("1 step >= 3.5 (MAP)", True, True, max)
    def onestep_ge_3_5(self, ij):
        return self._onestep_ge_cutoff(ij, 3.5, True)

    @_criterion("1 step >= 3.5 (Approx)", True, True, max)
    def onestep_ge_3_5_approx(self, ij):
        return self._onestep_ge_cutoff(ij, 3.5, False)

    @_criterion("1 step >= .5 (MAP)", True, True, max)
    def onestep_ge_half(self, ij):
        return self._onestep_ge_cutoff(ij, .5, True)

    @_criterion("1 step >= .5 (Approx)", True, True, max)
    def onestep_ge_half_approx(self, ij):
        return self._onestep_ge_cutoff(ij, .5, False)

    def _last_step_lookahead_helper(self, cutoff, v):
        '''(active_pmf.py:492-500)'''
        if not self.unrated:
            raise ValueError("max() arg is an empty sequence")
        pool = list(self.unrated)
        ii, jj = zip(*pool)
        _, (best, _idx) = self._normal_scores(N.CRIT_PROB_GE, ii, jj, cutoff=cutoff)
        return int(v >= cutoff) + best

    @_criterion("Pred Variance", True, False, max)
    def pred_variance(self, ij):
        '''Variance of the prediction for R_ij under the approximation (active_pmf.py:502-524).'''
        vals, _ = self._normal_scores(N.CRIT_PRED_VARIANCE, [ij[0]], [ij[1]])
        return float(vals[0])

    def _approx_entropy(self):
        '''(active_pmf.py:526-530) log det cov'''
        if self._in_blocks():
            return self._block_posterior().entropy()
        sign, logdet = _slogdet(self.cov)
        assert sign == 1
        return logdet

    @_criterion("E[U/V Entropy] (MAP)", True, True, min)
    def exp_approx_entropy(self, ij):
        return self._lookahead([ij], 'entropy', True)[0]

    @_criterion("E[U/V Entropy] (Approx)", True, True, min)
    def exp_approx_entropy_byapprox(self, ij):
        return self._lookahead([ij], 'entropy', False)[0]

    def _pred_entropy_bound(self):
        '''(active_pmf.py:559-574)'''
        s, logdet = _slogdet(self.approx_pred_covs())
        return _entropy_bound_from(s, logdet)

    @_criterion("E[Pred Entropy Bound] (MAP)", True, True, min)
    def exp_pred_entropy_bound(self, ij):
        return self._lookahead([ij], 'pred_entropy_bound', True)[0]

    @_criterion("E[Pred Entropy Bound] (Approx)", True, True, min)
    def exp_pred_entropy_bound_byapprox(self, ij):
        return self._lookahead([ij], 'pred_entropy_bound', False)[0]

    def _total_variance(self):
        if self._in_blocks():
            return self._block_posterior().total_variance()
        return self.approx_pred_means_vars()[1].sum()

    @_criterion("E[Pred Total Variance] (MAP)", True, True, min)
    def exp_total_variance(self, ij):
        return self._lookahead([ij], 'total_variance', True)[0]

    @_criterion("E[Pred Total Variance] (Approx)", True, True, min)
    def exp_total_variance_byapprox(self, ij):
        return self._lookahead([ij], 'total_variance', False)[0]

    # name of the quantity -> the reference's helper, for _exp_with_rij(fn=...) callers
    _FN_NAMES = {'_approx_entropy': 'entropy', '_total_variance': 'total_variance',
                 '_pred_entropy_bound': 'pred_entropy_bound'}

    def _exp_with_rij(self, ij, fn, use_map=True, discretize=None, pass_v=False):
        '''E[fn(apmf with R_ij)] (active_pmf.py:635-704) for one pair.'''
        name = getattr(fn, '__name__', '')
        if name in self._FN_NAMES:
            what = self._FN_NAMES[name]
        elif name.startswith('_1step_') or pass_v:
            cutoff = getattr(fn, 'keywords', {}).get('cutoff')
            what = ('onestep', cutoff)
        else:
            what = ('fn', fn, pass_v)   # arbitrary callable: evaluated on a host copy of each re-fit
        return self._lookahead([ij], what, use_map, discretize=discretize, pass_v=pass_v)[0]

    ############################################################################
    ### Batched lookahead

    def _rij_distribution(self, pool, use_map):
        '''mean and variance of the distribution assumed for each R_ij (active_pmf.py:656-666)'''
        ii, jj = zip(*pool)
        if use_map:
            mu, _ = _scoring.score_pred(self.users, self.items, ii, jj, "f64")
            var = np.full(len(pool), float(self.sigma_sq))
        else:
            mu, _ = self._normal_scores(N.CRIT_APPROX_MEAN, ii, jj)
            var, _ = self._normal_scores(N.CRIT_PRED_VARIANCE, ii, jj)
        return mu, var

    def _refits(self, pairs_vals, what):
        '''fn(model + (i, j, v)) for a list of (i, j, v): one variational re-fit each.'''
        self._require_approx()
        B = len(pairs_vals)
        if B == 0:
            return np.zeros(0)
        if self.refit_lookahead:
            return np.array([self._refit_one_host_driven(i, j, v, what) for i, j, v in pairs_vals])
        ei = np.array([p[0] for p in pairs_vals], dtype=np.int32)
        ej = np.array([p[1] for p in pairs_vals], dtype=np.int32)
        er = np.array([p[2] for p in pairs_vals], dtype=np.float64)
        k = self.approx_dim
        # bound the device workspace: (5 k^2 + 2k) doubles of scratch + k^2 + k of state each
        per = 8 * (6 * k * k + 3 * k)
        chunk = max(1, min(B, int(6e9 // per)))
        out = np.empty(B)
        # accepted line-search steps of every re-fit (a diagnostic: two implementations can only
        # agree on a re-fit when their accept / reject sequences do)
        steps_out = self._last_refit_steps = np.zeros(B, dtype=np.int64)
        for s in range(0, B, chunk):
            e = min(B, s + chunk)
            nb = e - s
            batch = _normal.NormalBatch(self.ratings, self._fit_params(),
                                        np.broadcast_to(self.mean, (nb, k)),
                                        np.broadcast_to(self.cov, (nb, k, k)),
                                        extra=(ei[s:e], ej[s:e], er[s:e]))
            res = batch.fit(want_entropy=(what == 'entropy'),
                            want_totvar=(what == 'total_variance'))
            steps_out[s:e] = res['steps']
            if what == 'entropy':
                out[s:e] = res['entropy']
            elif what == 'total_variance':
                out[s:e] = res['total_variance']
            elif what == 'pred_entropy_bound':
                # all re-fits of the chunk at once, where they are: Isserlis covariances of the
                # predictions and their log-determinants, both on the device
                pc = _pred_covs_device(batch.mean, batch.cov, self.num_users, self.num_items,
                                       self.latent_d)
                sg, ld = _slogdet_device(pc)
                out[s:e] = [_entropy_bound_from(int(a), float(b)) for a, b in zip(sg, ld)]
            else:
                means, covs = batch.means(), batch.covs()
                for b in range(nb):
                    out[s + b] = self._criterion_on(means[b], covs[b], what,
                                                    (int(ei[s + b]), int(ej[s + b])), er[s + b])
        return out

    def _criterion_on(self, mean, cov, what, ij, v):
        '''criteria that need more than the fit kernel's own outputs'''
        n, m, d = self.num_users, self.num_items, self.latent_d
        if what == 'pred_entropy_bound':
            s, logdet = _slogdet(_pred_covs(mean, cov, n, m, d))
            return _entropy_bound_from(s, logdet)
        if isinstance(what, tuple) and what[0] == 'fn':
            apmf = deepcopy(self)
            apmf.add_rating(ij[0], ij[1], v)
            apmf.mean, apmf.cov = mean, cov
            return what[1](apmf, v=v) if what[2] else what[1](apmf)
        if isinstance(what, tuple) and what[0] == 'onestep':
            cutoff = what[1]
            pool = [c for c in self.unrated if c != ij]
            if not pool:
                raise ValueError("max() arg is an empty sequence")
            ii, jj = zip(*pool)
            _, (best, _i) = _scoring.score_normal(N.CRIT_PROB_GE, mean, cov, n, m, d, ii, jj,
                                                  "f64", cutoff=cutoff)
            return int(v >= cutoff) + best
        raise ValueError("unknown lookahead quantity %r" % (what,))

    def _refit_one_host_driven(self, i, j, v, what):
        '''refit_lookahead=True (active_pmf.py:669-676): MAP refit + fresh random covariance per
        problem, consuming the global RNG in the reference's order.'''
        apmf = deepcopy(self)
        apmf.add_rating(i, j, v)
        apmf.do_fit()
        apmf.initialize_approx()
        apmf.fit_normal()
        if what == 'entropy':
            return apmf._approx_entropy()
        if what == 'total_variance':
            return apmf._total_variance()
        if what == 'pred_entropy_bound':
            return apmf._pred_entropy_bound()
        if what[0] == 'fn':
            return what[1](apmf, v=v) if what[2] else what[1](apmf)
        return apmf._last_step_lookahead_helper(what[1], v)

    def _lookahead_blocks(self, ii, jj, what, use_map, discretize=None, want_scores=True):
        '''_exp_with_rij (active_pmf.py:635-704) for every candidate in scalable mode: one launch,
        one lane group per candidate, the expectation over R_ij and the arg-min fused.  Returns
        (scores ndarray, (best value, best index)).'''
        self._require_approx()
        if what not in ('entropy', 'total_variance'):
            raise ValueError("criterion %r needs the exact (full-covariance) mode: "
                             "set approx_mode = 'exact'" % (what,))
        if self.refit_lookahead:
            raise ValueError("refit_lookahead needs the exact mode (approx_mode = 'exact')")
        if discretize is None:
            discretize = self.discrete_expectations
        post = self._block_posterior()
        ci, cj = _scoring._cands(ii, jj)
        if use_map:          # R_ij ~ N(U_i . V_j, sigma^2) at the MAP factors (:656-659)
            U, V = _D.to_padded(self.users, "f64"), _D.to_padded(self.items, "f64")
            mu, _ = _scoring.score_device(N.CRIT_PRED, "f64", ci, cj, self.latent_d, U, V)
            sd = torch.full_like(mu, math.sqrt(self.sigma_sq))
        else:                # ... or the approximation's own mean and variance (:660-666)
            mu, _ = post.score(N.CRIT_APPROX_MEAN, ci, cj, "f64")
            var, _ = post.score(N.CRIT_PRED_VARIANCE, ci, cj, "f64")
            sd = var.sqrt()
        code = N.LOOK_ENTROPY if what == 'entropy' else N.LOOK_TOTAL_VARIANCE
        points = self.rating_values
        rounds = self.lookahead_rounds
        if discretize and points:
            vals = np.array(points, dtype=float)
            if discretize == 'simps':
                evals, _, _ = post.lookahead(code, ci, cj, vals, rounds=rounds, want_evals=True)
                pdfs = stats.norm.pdf(vals[None, :], loc=mu.cpu().numpy()[:, None],
                                      scale=sd.cpu().numpy()[:, None])
                est = scipy.integrate.simpson(evals.cpu().numpy() * pdfs, x=vals, axis=1)
                return est, _argbest(est, False)
            _, scores, best = post.lookahead(code, ci, cj, vals, N.WEIGHTS_DISCRETE,
                                             self.rating_bounds, mu, sd, rounds=rounds,
                                             want_scores=want_scores)
        else:
            if discretize and points is None:
                warnings.warn("ActivePMF has no rating_values; doing integral")
            t, w = _blocks.gauss_nodes(self.quadrature_nodes)
            _, scores, best = post.lookahead(code, ci, cj, t, N.WEIGHTS_NODES, w, mu, sd,
                                             rounds=rounds, want_scores=want_scores)
        return (scores.cpu().numpy() if scores is not None else None), _scoring.unpack_best(best)

    def _lookahead(self, pool, what, use_map, discretize=None, pass_v=False):
        '''_exp_with_rij for every pair of `pool`.'''
        if self._in_blocks():
            ii, jj = _pool_arrays(pool) if not isinstance(pool, np.ndarray) else (pool[:, 0], pool[:, 1])
            return self._lookahead_blocks(ii, jj, what, use_map, discretize)[0].tolist()
        pool = [(int(i), int(j)) for i, j in pool]
        if discretize is None:
            discretize = self.discrete_expectations
        mu, var = self._rij_distribution(pool, use_map)
        std = np.sqrt(var)
        points = self.rating_values
        if discretize and points:
            vals = np.array(points, dtype=float)
            trip = [(i, j, v) for (i, j) in pool for v in vals]
            evals = self._refits(trip, what).reshape(len(pool), len(vals))
            if discretize == 'simps':
                pdfs = stats.norm.pdf(vals[None, :], loc=mu[:, None], scale=std[:, None])
                est = scipy.integrate.simpson(evals * pdfs, x=vals, axis=1)
                how = "simps'ed"
            else:
                cdfs = stats.norm.cdf(self.rating_bounds[None, :], loc=mu[:, None],
                                      scale=std[:, None])
                est = (evals * np.diff(cdfs, axis=1)).sum(1)
                how = "summed"
        else:
            if discretize and points is None:
                warnings.warn("ActivePMF has no rating_values; doing integral")
            est = np.empty(len(pool))
            for t, (i, j) in enumerate(pool):
                left, right = mu[t] - 2 * std[t], mu[t] + 2 * std[t]
                est[t] = stats.norm.expect(
                    lambda v: float(self._refits([(i, j, float(v))], what)[0]),
                    loc=mu[t], scale=std[t], lb=left, ub=right, epsrel=.02)
            how = "integrated"
        if self.verbose_lookahead:
            name = what if isinstance(what, str) else getattr(what, '__name__', str(what))
            for (i, j), e in zip(pool, est):
                print("\t{:>20}({},{}) {}: {: 10.2f}".format(name, i, j, how, e))
        return [float(e) for e in est]

    ############################################################################
    ### Picking a query point

    def pick_query_point(self, pool=None, key=None, procs=None, worker_pool=None):
        '''(active_pmf.py:709-737); procs / worker_pool are accepted and ignored.  The winner is
        the one the scoring launch itself reduced (best value, first in pool order on ties --
        what `chooser(zip(pool, vals), key=itemgetter(1))` returns); no score list is built.'''
        if pool is None:
            pool = self.unrated
        if key is None:
            key = ActivePMF.pred_variance
        if len(pool) == 0:
            raise ValueError("can't pick a query point from an empty pool")
        elif len(pool) == 1:
            first = next(iter(pool))
            return (int(first[0]), int(first[1])) if isinstance(pool, np.ndarray) else first
        pool, vals, best = self._key_scores(pool, key, want_scores=False)
        if best is None:
            chooser = getattr(key, 'chooser', max)
            best = chooser(range(len(vals)), key=vals.__getitem__)
        ij = pool[best]
        return (int(ij[0]), int(ij[1])) if isinstance(pool, np.ndarray) else ij

    _LOOKAHEADS = {
        'exp_approx_entropy': ('entropy', True),
        'exp_approx_entropy_byapprox': ('entropy', False),
        'exp_total_variance': ('total_variance', True),
        'exp_total_variance_byapprox': ('total_variance', False),
        'exp_pred_entropy_bound': ('pred_entropy_bound', True),
        'exp_pred_entropy_bound_byapprox': ('pred_entropy_bound', False),
    }
    _ONESTEPS = {'onestep_ge_3_5': (3.5, True), 'onestep_ge_3_5_approx': (3.5, False),
                 'onestep_ge_half': (.5, True), 'onestep_ge_half_approx': (.5, False)}
    _CELL_CRITERIA = {'pred': (N.CRIT_PRED, 0.), 'pred_variance': (N.CRIT_PRED_VARIANCE, 0.),
                      'prob_ge_3_5': (N.CRIT_PROB_GE, 3.5), 'prob_ge_half': (N.CRIT_PROB_GE, .5)}

    def _pool_device(self, pool):
        '''(indexable pool, device i, device j).  A CandidatePool is already resident; an
        (n, 2) integer array is uploaded once and remembered while the same object comes back;
        anything else is listed and converted like the reference's iteration over it.'''
        if isinstance(pool, _scoring.CandidatePool):
            pools = self._dev.setdefault('candidate_pools', [])
            if not any(p is pool for p in pools):
                pools.append(pool)           # add_rating(s) removes queried cells from it
            ci, cj = pool.device_arrays()
            return pool, ci, cj
        if isinstance(pool, np.ndarray) and pool.ndim == 2 and pool.shape[1] == 2:
            hit = self._dev.get('pool_array')
            if hit is not None and hit[0] is pool:
                return pool, hit[1], hit[2]
            if pool.dtype == np.int32 and pool.flags.c_contiguous:
                # one H2D copy of the (n, 2) array, split into the two index vectors on the device
                both = torch.from_numpy(pool).to(_D.device())
                ci, cj = both[:, 0].contiguous(), both[:, 1].contiguous()
            else:
                ci = _D.to_device(pool[:, 0], np.int32)
                cj = _D.to_device(pool[:, 1], np.int32)
            self._dev['pool_array'] = (pool, ci, cj)
            return pool, ci, cj
        pool = list(pool)
        ii, jj = _pool_arrays(pool)
        return pool, _D.to_device(ii, np.int32), _D.to_device(jj, np.int32)

    def _map_factors_device(self, name):
        '''padded MAP factors on the device: the tensors of a device-resident fit when they are
        the newest copy, else an upload of the host arrays'''
        dev = self._dev
        if dev.get('host_stale') and dev.get('U') is not None and \
                dev['U'].dtype == _D.torch_dtype(name):
            return dev['U'], dev['V']
        return _D.to_padded(self.users, name), _D.to_padded(self.items, name)

    def _key_scores(self, pool, key, want_scores=True):
        '''(pool as an indexable sequence, float64 ndarray of criterion values aligned with it
        or None when not wanted and a fused winner exists, index of the winner or None).'''
        name = getattr(key, '__name__', None)
        maximize_ = getattr(key, 'chooser', max) is max
        if name == 'random_weighting':
            pool = pool if isinstance(pool, (np.ndarray, _scoring.CandidatePool)) else list(pool)
            return pool, np.array([random.random() for _ in range(len(pool))]), None
        if name in self._CELL_CRITERIA:
            crit, cutoff = self._CELL_CRITERIA[name]
            pool, ci, cj = self._pool_device(pool)
            if crit == N.CRIT_PRED:
                U, V = self._map_factors_device(self.dtype_name)
                scores, best = _scoring.score_device(crit, self.dtype_name, ci, cj, self.latent_d,
                                                     U, V, want_scores=want_scores,
                                                     maximize=maximize_)
            else:
                self._require_approx()
                if self._in_blocks():
                    scores, best = self._block_posterior().score(
                        crit, ci, cj, "f64", cutoff=cutoff, want_scores=want_scores,
                        maximize=maximize_)
                else:
                    vals, (_bv, bi) = self._normal_scores(crit, ci, cj, cutoff=cutoff,
                                                          maximize_=maximize_)
                    return pool, vals, bi
            vals = scores.to(torch.float64).cpu().numpy() if scores is not None else None
            return pool, vals, _scoring.unpack_best(best)[1]
        if name in self._LOOKAHEADS or name in self._ONESTEPS:
            if name in self._LOOKAHEADS:
                what, use_map = self._LOOKAHEADS[name]
                discretize = None
            else:
                cutoff, use_map = self._ONESTEPS[name]
                what, discretize = ('onestep', cutoff), True
            if self._in_blocks():
                pool, ci, cj = self._pool_device(pool)
                vals, (_bv, bi) = self._lookahead_blocks(ci, cj, what, use_map, discretize,
                                                         want_scores=want_scores)
                return pool, vals, bi
            pool = [(int(i), int(j)) for i, j in pool]
            return pool, np.array(self._lookahead(pool, what, use_map, discretize=discretize)), None
        # unknown criterion: evaluate it pair by pair like the reference's serial path
        pool = list(pool)
        return pool, np.array([key(self, ij) for ij in pool], dtype=float), None

    def _get_key_vals(self, pool, key, procs=None, worker_pool=None):
        '''Criterion value for every pair of `pool`, aligned with its iteration order
        (active_pmf.py:739-770) -- evaluated in batched GPU launches.  Returns a float64 array
        (the reference returns a list; every caller indexes, zips or assigns it).'''
        if len(pool) == 0:
            return np.zeros(0)
        return self._key_scores(pool, key)[1]

    def get_key_evals(self, pool=None, key=None, procs=None, worker_pool=None):
        '''(active_pmf.py:772-787) NaN-filled (N, M) matrix of criterion values'''
        if pool is None:
            pool = self.unrated
        if key is None:
            key = ActivePMF.pred_variance
        evals = np.empty((self.num_users, self.num_items))
        evals.fill(np.nan)
        if len(pool):
            pool, vals, _ = self._key_scores(pool, key)
            if isinstance(pool, np.ndarray):
                ii, jj = pool[:, 0], pool[:, 1]
            elif isinstance(pool, _scoring.CandidatePool):
                ii, jj = pool.i[:pool.n], pool.j[:pool.n]
            else:
                ii, jj = _pool_arrays(pool)
            evals[ii, jj] = vals
        return evals


def _argbest(vals, maximize_):
    '''(value, index) of the best entry, lowest index on ties, NaN never wins'''
    vals = np.asarray(vals, dtype=float)
    ok = ~np.isnan(vals)
    if not ok.any():
        return 0.0, -1
    idx = int(np.nanargmax(vals) if maximize_ else np.nanargmin(vals))
    return float(vals[idx]), idx


def _slogdet_device(mats_t):
    '''np.linalg.slogdet of a (B, k, k) device tensor (overwritten): slogdet_kernel, LU with
    partial pivoting, one CTA per matrix.  Returns host arrays (sign, logdet).'''
    import torch
    from . import device as D
    lib = N.require_device()
    B, k = int(mats_t.shape[0]), int(mats_t.shape[1])
    sign = torch.empty(B, dtype=torch.int32, device=mats_t.device)
    logdet = torch.empty(B, dtype=torch.float64, device=mats_t.device)
    N.check(lib.amf_slogdet_batched(k, B, D.ptr(mats_t), D.ptr(sign), D.ptr(logdet), D.stream_ptr()))
    return sign.cpu().numpy(), logdet.cpu().numpy()


def _slogdet(mat):
    '''np.linalg.slogdet semantics, computed on the device'''
    from . import device as D
    s, ld = _slogdet_device(D.to_device(np.asarray(mat, dtype=np.float64)[None], np.float64))
    return float(s[0]), float(ld[0])


def _pred_covs_device(means_t, covs_t, n, m, d):
    '''(active_pmf.py:324-390) for a batch: device tensors (B, k), (B, k, k) -> (B, NM, NM).

    With X = [vec U; vec V] ~ N(mean, cov) every entry is a sum of 4th moments minus a product of
    2nd moments; by Isserlis (x1 = U_ki, x2 = V_kj, x3 = U_la, x4 = V_lb)
      Cov(x1 x2, x3 x4) = m1 m3 C24 + m1 m4 C23 + m2 m3 C14 + m2 m4 C13 + C13 C24 + C14 C23,
    summed over k, l by pred_covs_kernel (csrc/predcov.cu), one thread per entry.'''
    import torch
    from . import device as D
    lib = N.require_device()
    B = int(means_t.shape[0])
    out = torch.empty((B, n * m, n * m), dtype=torch.float64, device=means_t.device)
    N.check(lib.amf_pred_covs(n, m, d, B, D.ptr(means_t), D.ptr(covs_t), D.ptr(out), D.stream_ptr()))
    return out


def _pred_covs(mean, cov, n, m, d):
    from . import device as D
    out = _pred_covs_device(D.to_device(np.asarray(mean, dtype=np.float64)[None], np.float64),
                            D.to_device(np.asarray(cov, dtype=np.float64)[None], np.float64), n, m, d)
    return out[0].cpu().numpy()


def _entropy_bound_from(sign, logdet):
    '''the sign logic of _pred_entropy_bound (active_pmf.py:559-574)'''
    if sign != 1:
        if sign == -1 and logdet < -50:
            return -1000
        raise ValueError("prediction cov has det with sign {}, log {}".format(sign, logdet))
    return logdet


################################################################################
### Registry (active_pmf.py:901-923).  The experiment drivers of active_pmf.py:796-1257
### (full_test, compare, make_fake_data, get_ratings, main) are not restated here:
### drivers.load("active_pmf", ref_dir) runs the reference's own against these classes.

KEY_FUNCS = {
    "random": ActivePMF.random_weighting,
    "pred-variance": ActivePMF.pred_variance,

    "total-variance": ActivePMF.exp_total_variance,
    "total-variance-approx": ActivePMF.exp_total_variance_byapprox,

    "uv-entropy": ActivePMF.exp_approx_entropy,
    "uv-entropy-approx": ActivePMF.exp_approx_entropy_byapprox,

    "pred-entropy-bound": ActivePMF.exp_pred_entropy_bound,
    "pred-entropy-bound-approx": ActivePMF.exp_pred_entropy_bound_byapprox,

    "pred": ActivePMF.pred,
    "prob-ge-3.5": ActivePMF.prob_ge_3_5,
    "prob-ge-.5": ActivePMF.prob_ge_half,

    "1step-ge-3.5": ActivePMF.onestep_ge_3_5,
    "1step-ge-3.5-approx": ActivePMF.onestep_ge_3_5_approx,

    "1step-ge-.5": ActivePMF.onestep_ge_half,
    "1step-ge-.5-approx": ActivePMF.onestep_ge_half_approx,
}
