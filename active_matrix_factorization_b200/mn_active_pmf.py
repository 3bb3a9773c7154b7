"""Host mirror of the reference's ``mn_active_pmf`` module (python-pmf/mn_active_pmf.py):
active learning on PMF with a MATRIX-NORMAL approximate posterior
MN(mean, cov_useritems, cov_latents) -- the variant the reference's drugbank and movielens
experiments run (results/drugbank-94x425/Makefile:66-76).  SURVEY.md 8f-1.

Same class surface as the reference (``MNActivePMF``, the 13 ``KEY_FUNCS``, ``full_test``,
``compare``, ``main``); the numerics run in csrc/mn.cu: the whole ``fit_normal`` line search of
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
from .active_pmf import (ActivePMF, _InlinePool, add_bool_opt,  # noqa: F401
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
### Drivers (mn_active_pmf.py:785-1132)

def _mean_info(apmf):
    return "Mean diff of means: %g; mean useritems cov %g, latents cov %g" % (
        apmf.mean_meandiff(), np.abs(apmf.cov_useritems.mean()), np.abs(apmf.cov_latents.mean()))


def full_test(apmf, real, picker_key=MNActivePMF.pred_variance, fit_normal=True,
              fit_sigmas=False, processes=None, test_on=None):
    '''(mn_active_pmf.py:795-846)'''
    print("Training PMF")
    if fit_sigmas:
        apmf.fit_with_sigmas()
    else:
        apmf.do_fit()
    apmf.initialize_approx()
    if fit_normal:
        print("Fitting normal")
        apmf.fit_normal()
        print(_mean_info(apmf))

    total = apmf.num_users * apmf.num_items
    rmse = apmf.rmse(real, test_on)
    print("RMSE: {:.5}".format(rmse))
    yield len(apmf.rated), rmse, None, None, None

    while apmf.unrated:
        print()
        print("Picking a query point...")
        if len(apmf.unrated) == 1:
            i, j = next(iter(apmf.unrated))
            vals = None
        else:
            pool = list(apmf.unrated)
            vals = apmf._get_key_vals(pool, picker_key, processes, None)
            i, j = picker_key.chooser(zip(pool, vals), key=operator.itemgetter(1))[0]
        apmf.add_rating(i, j, real[i, j])
        print("Queried (%d, %d); %d/%d known" % (i, j, len(apmf.rated), total))
        print("Training PMF")
        for _ll in apmf.fit_lls():
            pass
        if fit_normal:
            print("Fitting normal")
            for kl in apmf.fit_normal_kls():
                assert kl > -1e5
            print(_mean_info(apmf))
        rmse = apmf.rmse(real, test_on)
        print("RMSE: {:.5}".format(rmse))
        yield len(apmf.rated), rmse, (i, j), vals, apmf.predicted_matrix()


_in_between_work = _apmf._in_between_work


def _full_test_threaded(apmf, real, picker_key, fit_normal, fit_sigmas, worker_pool, test_on=None):
    '''(mn_active_pmf.py:867-894)'''
    total = real.size
    name = picker_key.nice_name
    rmse = apmf.rmse(real, test_on)
    print("{:<40} Initial RMSE: {:.5}".format(name, rmse))
    yield len(apmf.rated), rmse, None, None, None
    while apmf.unrated:
        n = len(apmf.rated) + 1
        print("{:<40} Picking query point {}...".format(name, n))
        if len(apmf.unrated) == 1:
            vals = np.empty((apmf.num_users, apmf.num_items))
            vals.fill(np.nan)
            i, j = next(iter(apmf.unrated))
        else:
            vals = apmf.get_key_evals(key=picker_key, worker_pool=worker_pool)
            i, j = picker_key.chooser(apmf.unrated, key=vals.__getitem__)
        apmf = worker_pool.apply(_in_between_work,
                                 (apmf, i, j, real[i, j], total, fit_normal, fit_sigmas, name))
        rmse = apmf.rmse(real, test_on)
        print("{:<40} RMSE {}: {:.5}".format(picker_key.nice_name, n, rmse))
        yield len(apmf.rated), rmse, (i, j), vals, apmf.predicted_matrix()


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


def compare(key_names, real, ratings, rating_vals=None, latent_d=5, knowable=None, test_on=None,
            processes=None, do_threading=True, steps=None, discrete_exp=False,
            refit_lookahead=False, fit_sigmas=False, apmf=None,
            sig_u_mean=0, sig_u_var=-1, sig_v_mean=0, sig_v_var=-1, fit_type=('batch',)):
    '''(mn_active_pmf.py:922-1003); `processes` is accepted and ignored (GPU fan-out).'''
    from threading import Thread, Lock
    if apmf:
        assert (apmf.num_users, apmf.num_items) == real.shape
        assert np.all(apmf.ratings == ratings)
        assert set(apmf.rating_values) == set(rating_vals)
        apmf.discrete_expectations = discrete_exp
    else:
        apmf = MNActivePMF(ratings, latent_d=latent_d, rating_values=rating_vals,
                           discrete_expectations=discrete_exp, refit_lookahead=refit_lookahead,
                           knowable=knowable, fit_type=fit_type)
        apmf.sig_u_mean, apmf.sig_u_var = sig_u_mean, sig_u_var
        apmf.sig_v_mean, apmf.sig_v_var = sig_v_mean, sig_v_var
        print("Doing initial fit")
        if fit_sigmas:
            apmf.fit_with_sigmas()
        else:
            apmf.do_fit()
        if any(KEY_FUNCS[name].do_normal_fit for name in key_names):
            apmf.initialize_approx()
            print("Initial approximation fit")
            apmf.fit_normal()
            print(_mean_info(apmf))

    results = {'_real': real, '_ratings': ratings, '_rating_vals': rating_vals,
               '_initial_apmf': deepcopy(apmf)}
    if do_threading:
        worker_pool = _InlinePool()
        worker_pool.access_lock = Lock()

        def eval_key(key_name):
            key = KEY_FUNCS[key_name]
            res = _full_test_threaded(deepcopy(apmf), real, key, key.do_normal_fit, fit_sigmas,
                                      worker_pool, test_on)
            results[key_name] = list(itertools.islice(res, steps))

        threads = [Thread(name=k, target=eval_key, args=(k,)) for k in key_names]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    else:
        for key_name in key_names:
            key = KEY_FUNCS[key_name]
            res = full_test(deepcopy(apmf), real, key, key.do_normal_fit, fit_sigmas, processes, test_on)
            results[key_name] = list(itertools.islice(res, steps))
    return results


def main(argv=None):
    '''Same command line as the reference (mn_active_pmf.py:1011-1132).'''
    import os
    import pickle
    import sys

    key_names = set(KEY_FUNCS.keys())
    spec = [e for e in _apmf._CLI if e[0] == "Model Options" or e[0] == "Running"]
    spec += [("Problem", ('--load-data',), dict(default=None, metavar='FILE')),
             ("Results", ('--save-results',), dict(default=True, metavar='FILE')),
             ("Results", ('--no-save-results',), dict(action='store_false', dest='save_results')),
             ("Results", ('--note',), dict(action='append'))]
    parser = _apmf.build_parser(spec, key_names)
    args = parser.parse_args(argv)

    for k in args.keys:
        if k not in key_names:
            sys.stderr.write("Invalid key name %s; options are %s.\n" % (k, ', '.join(sorted(key_names))))
            sys.exit(1)
    if not args.keys:
        args.keys = sorted(key_names)
    if args.save_results is True:
        args.save_results = 'results.pkl'
    elif args.save_results:
        dirname = os.path.dirname(args.save_results)
        if dirname and not os.path.exists(dirname):
            os.makedirs(dirname)

    with open(args.load_data, 'rb') as f:
        data = np.load(f, allow_pickle=True)
        if isinstance(data, np.ndarray):
            data = {'_real': data}
        real = data['_real']
        ratings = data['_ratings']
        rating_vals = data['_rating_vals'] if '_rating_vals' in data else None
        test_on = data['_test_on'] if '_test_on' in data else None

    knowable = np.isfinite(real)
    knowable[real == 0] = False
    if test_on is not None:
        knowable[test_on] = False
    knowable = zip(*knowable.nonzero())

    results = compare(args.keys, real=real, ratings=ratings, rating_vals=rating_vals,
                      latent_d=args.latent_d, knowable=knowable, test_on=test_on,
                      discrete_exp=args.discrete_integration, refit_lookahead=args.refit_lookahead,
                      fit_sigmas=args.fit_sigmas, sig_u_mean=args.sig_u_mean,
                      sig_u_var=args.sig_u_var, sig_v_mean=args.sig_v_mean, sig_v_var=args.sig_v_var,
                      steps=args.steps, fit_type=parse_fit_type(args.fit),
                      processes=args.processes, do_threading=args.threading)
    if args.save_results:
        print("saving results in '{}'".format(args.save_results))
        results['_args'] = args
        with open(args.save_results, 'wb') as f:
            pickle.dump(results, f)
    return results


if __name__ == '__main__':
    main()
