"""Host mirror of the reference's ``bayes_pmf`` module (python-pmf/bayes_pmf.py + .pxd):
Bayesian PMF by Gibbs sampling (Salakhutdinov & Mnih), with the active-learning criteria
computed from posterior samples.

The two O(data) pieces run in libamf_b200:
  * every row/column conditional of a half-sweep (``sample_feature``, bayes_pmf.py:189-216) in
    one launch, one CTA per row (csrc/gibbs.cu);
  * predict / pred_variance / prob_ge_cutoff over S samples at the requested cells as one
    streaming pass (no S dense N x M matrices).
Random numbers are drawn on the host from numpy's global stream in the reference's order, so
a seeded chain reproduces the reference's samples.  No CPU path exists for the numerics.
"""
from collections import namedtuple
from copy import deepcopy
from itertools import islice, repeat
import ctypes as C
import random
from threading import Thread
import warnings

import numpy as np
from scipy import stats, integrate
import torch

from . import _native as N
from . import device as D
from .pmf_cy import ProbabilisticMatrixFactorization, rmse, parse_fit_type  # noqa: F401


################################################################################
### Utilities

def sample_wishart(sigma, dof):
    '''
    Draw from Wishart(sigma, dof) (bayes_pmf.py:41-59): d x d host algebra on draws from the
    global numpy stream (direct scheme if dof <= 81+n and integral, else Bartlett).  `dof` is a
    C int in the compiled reference (bayes_pmf.pxd:7), so a fractional value is truncated.
    '''
    dof = int(dof)
    n = sigma.shape[0]
    chol = np.linalg.cholesky(sigma)
    if dof <= 81 + n and dof == round(dof):
        X = np.dot(chol, np.random.normal(size=(n, int(dof))))
    else:
        A = np.diag(np.sqrt(np.random.chisquare(dof - np.arange(0, n), size=n)))
        A[np.tri(n, k=-1, dtype=bool)] = np.random.normal(size=(n * (n - 1) // 2))
        X = np.dot(chol, A)
    return np.dot(X, X.T)


def iter_mean(iterable):
    i = iter(iterable)
    total = next(i)
    count = -1
    for count, x in enumerate(i):
        total += x
    return total / (count + 2)


################################################################################

class BayesianPMF(ProbabilisticMatrixFactorization):
    def __init__(self, rating_tuples, latent_d=5, subtract_mean=True, rating_values=None,
                 discrete_expectations=True, num_integration_pts=50, knowable=None,
                 fit_type=('batch',)):
        super(BayesianPMF, self).__init__(rating_tuples, latent_d=latent_d,
                                          subtract_mean=subtract_mean, knowable=knowable,
                                          fit_type=fit_type)
        if rating_values is not None:
            rating_values = set(map(float, rating_values))
            if not rating_values.issuperset(self.ratings[:, 2]):
                raise ValueError("got ratings not in rating_values")
        self.rating_values = rating_values
        self.discrete_expectations = discrete_expectations
        self.num_integration_pts = num_integration_pts

        self.beta = 2  # observation noise precision

        # (wishart scale, b0, degrees of freedom, mu0)   (bayes_pmf.py:97-109)
        self.u_hyperparams = (np.eye(latent_d), 2, latent_d, np.zeros(latent_d))
        self.v_hyperparams = (np.eye(latent_d), 2, latent_d, np.zeros(latent_d))

    def __copy__(self):
        res = BayesianPMF(self.ratings, self.latent_d)
        res.__setstate__(self.__getstate__())
        return res

    def __deepcopy__(self, memodict):
        res = BayesianPMF(self.ratings, self.latent_d)
        res.__setstate__(deepcopy(self.__getstate__(), memodict))
        return res

    def __getstate__(self):
        # the five extra keys of the compiled reference (bayes_pmf.py:123-130)
        state = super(BayesianPMF, self).__getstate__()
        state['discrete_expectations'] = self.discrete_expectations
        state['rating_values'] = self._rating_values
        state['beta'] = self.beta
        state['u_hyperparams'] = self.u_hyperparams
        state['v_hyperparams'] = self.v_hyperparams
        state['num_integration_pts'] = self.num_integration_pts
        return state

    def _set_rating_values(self, vals):
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

    rating_values = property(lambda self: self._rating_values, _set_rating_values)
    rating_bounds = property(lambda self: self._rating_bounds)

    ############################################################################
    ### Gibbs sampler

    def sample_hyperparam(self, feats, do_users):
        '''
        Normal-Wishart posterior draw of (mu, alpha) given a factor matrix
        (bayes_pmf.py:157-186).  d x d host algebra; the reference's scalar inner product at
        :176 (np.dot of two 1-D vectors) is kept.
        '''
        # the O(N d^2) moments of the factor matrix on the device; d x d algebra + RNG on the host
        if isinstance(feats, torch.Tensor):        # fast mode: the sample never left the device
            ft = feats.to(torch.float64)
        else:
            ft = torch.from_numpy(np.ascontiguousarray(feats, dtype=np.float64)).to(D.device())
        x_bar = ft.mean(dim=0).cpu().numpy()
        s_bar = np.atleast_2d(torch.cov(ft.T).cpu().numpy()) if feats.shape[1] > 1 else \
            np.array(float(ft.var(unbiased=True).item()))
        return self._hyperparam_from_moments(feats.shape[0], x_bar, s_bar, do_users)

    def _hyperparam_from_moments(self, n, x_bar, s_bar, do_users):
        '''the d x d part of sample_hyperparam (bayes_pmf.py:166-186) given mean and covariance'''
        wi, b0, df, mu0 = self.u_hyperparams if do_users else self.v_hyperparams
        diff = mu0 - x_bar
        wi_post = np.linalg.inv(np.linalg.inv(wi) + n * s_bar
                                + (b0 * n) / (b0 + n) * np.dot(diff, diff.T))
        wi_post /= 2
        wi_post = wi_post + wi_post.T
        alpha = sample_wishart(wi_post, df + n)
        mu_temp = (b0 * mu0 + n * x_bar) / (b0 + n)
        lam = np.linalg.cholesky(np.linalg.inv((b0 + n) * alpha))
        mu = np.dot(lam, np.random.normal(0, 1, self.latent_d)) + mu_temp
        return mu, alpha

    def _mean_offset(self):
        return self.mean_rating if self.subtract_mean else 0.

    # Multi-GPU Gibbs is OPT-IN: set `shard_group` to True (default process
    # group) or to a torch.distributed group and the rows of every half-sweep of samples() are
    # split over its ranks.  mu, alpha and z are then rank 0's draws, broadcast, so the ranks
    # need not be seeded alike and cannot drift apart; single-row calls (sample_feature) and
    # models without the attribute set never touch the process group.
    shard_group = None

    def _shard(self):
        """(world, rank, group) of the opt-in row sharding, (1, 0, None) when it is off"""
        if self.shard_group is None or self.shard_group is False:
            return 1, 0, None
        from . import parallel as P
        if not (P.dist is not None and P.dist.is_available() and P.dist.is_initialized()):
            raise RuntimeError("shard_group is set but torch.distributed is not initialised")
        group = None if self.shard_group is True else self.shard_group
        return P.dist.get_world_size(group), P.dist.get_rank(group), group

    def _shared_draw(self, t):
        """rank 0's value of a device tensor on every rank of the shard group"""
        world, _rank, group = self._shard()
        if world > 1:
            from . import parallel as P
            P.dist.broadcast(t, P.dist.get_global_rank(group, 0) if group is not None else 0,
                             group=group)
        return t

    def _half_sweep(self, rat, side, other_t, mu, alpha, z, shard=False):
        '''all conditionals of one side in one launch; returns the new (rows, d) device tensor'''
        lib = N.require_device()
        name = rat.name
        dt = D.np_dtype(name)
        rows = rat.n_users if side == 0 else rat.n_items
        out = torch.empty((rows, self.latent_d), dtype=D.torch_dtype(name), device=other_t.device)
        alpha_t = D.to_device(np.atleast_2d(alpha), dt)
        mu_t = D.to_device(np.atleast_1d(mu), dt)
        z_t = D.to_device(z, dt)
        world, rank, group = self._shard() if shard else (1, 0, None)
        lo, hi = 0, rows
        if world > 1:
            from . import parallel as P
            for t in (alpha_t, mu_t, z_t):
                self._shared_draw(t)
            lo, hi = P.shard_bounds(rows, world, rank)
        N.check(lib.amf_gibbs_half_sweep_rows(rat.handle, side, D.code(name), self.latent_d,
                                              D.ptr(other_t), D.ptr(alpha_t), D.ptr(mu_t),
                                              float(self.beta), float(self._mean_offset()),
                                              D.ptr(z_t), D.ptr(out), lo, hi, D.stream_ptr()))
        if world > 1:       # rows of this side are split over the ranks
            out = P.all_gather_rows(out, rows, world, group)
        return out

    def _check_gibbs(self, rat, shard=False):
        '''raises like np.linalg.cholesky (bayes_pmf.py:215) if any half-sweep since the last
        check met a non-positive-definite matrix; the device flag is sticky, and in a sharded
        chain it is max-reduced first so that every rank raises together'''
        failed = C.c_int(0)
        N.check(N.load().amf_gibbs_status(rat.handle, C.byref(failed), D.stream_ptr()))
        bad = failed.value
        world, _rank, group = self._shard() if shard else (1, 0, None)
        if world > 1:
            from . import parallel as P
            dev = torch.device('cuda', torch.cuda.current_device())
            flag = torch.tensor([bad], dtype=torch.int32, device=dev)
            P.dist.all_reduce(flag, op=P.dist.ReduceOp.MAX, group=group)
            bad = int(flag.item())
        if bad:
            raise np.linalg.LinAlgError("Matrix is not positive definite")

    def sample_feature(self, n, is_user, mu, alpha, oth_feats, rated_indices, ratings):
        '''
        One row's conditional draw (bayes_pmf.py:189-216): z ~ N(0, I) from the global stream,
        then chol(inv(alpha + beta F'F)) z + mean on the device.
        '''
        name = self.dtype_name
        rated_indices = np.asarray(rated_indices, dtype=np.int32)
        rat = D.Ratings(1, oth_feats.shape[0], np.zeros(len(rated_indices), np.int32),
                        rated_indices, np.asarray(ratings, dtype=float), name)
        other_t = D.to_device(oth_feats, D.np_dtype(name))
        z = np.random.normal(0, 1, self.latent_d)
        out = self._half_sweep(rat, 0, other_t, mu, alpha, z[None])
        self._check_gibbs(rat)
        res = out[0].to(torch.float64).cpu().numpy()
        rat.close()
        return res

    def samples(self, num_gibbs=2, fit_first=False):
        '''
        The Markov chain of bayes_pmf.py:227-302, started at the current MAP factors; yields
        (user_sample, item_sample) forever.  Ratings added after the generator started are
        ignored, like in the reference.
        '''
        if self.rng_mode == 'device':
            for us, vs in self.samples_device(num_gibbs=num_gibbs, fit_first=fit_first):
                yield us.to(torch.float64).cpu().numpy(), vs.to(torch.float64).cpu().numpy()
            return
        if self.rng_mode != 'host':
            raise ValueError("rng_mode must be 'host' or 'device'")
        name = self.dtype_name
        dt = D.np_dtype(name)
        n, m, d = self.num_users, self.num_items, self.latent_d
        rat = D.Ratings.from_tuples(self.ratings, n, m, name)   # adjacency, rating-list order

        if fit_first:
            self.do_fit()

        user_sample = self.users.copy()
        item_sample = self.items.copy()
        items_t = D.to_device(item_sample, dt)

        while True:
            mu_u, alpha_u = self.sample_hyperparam(user_sample, True)
            mu_v, alpha_v = self.sample_hyperparam(item_sample, False)
            for _gibbs in range(num_gibbs):
                z = np.random.normal(0, 1, (n, d))          # row-major = the reference's per-row draws
                users_t = self._half_sweep(rat, 0, items_t, mu_u, alpha_u, z, shard=True)
                z = np.random.normal(0, 1, (m, d))
                items_t = self._half_sweep(rat, 1, users_t, mu_v, alpha_v, z, shard=True)
            self._check_gibbs(rat, shard=True)
            user_sample = users_t.to(torch.float64).cpu().numpy()
            item_sample = items_t.to(torch.float64).cpu().numpy()
            yield user_sample, item_sample

    # 'host': every normal comes from numpy's legacy global stream in the reference's draw order
    # (seeded chains reproduce the reference's samples).  'device': the (N+M) d normals of every
    # half-sweep are generated inside the kernel (Philox4x32-10) and each row needs one Cholesky
    # (amf_gibbs_half_sweep_device_rng): the same Markov chain in law, several times faster.
    rng_mode = 'host'
    device_seed = 0

    # where the d x d Normal-Wishart draws of the fast-mode chain are made: 'device'
    # (amf_gibbs_hyper_device: nothing of a sample touches the host) or 'host' (numpy, global stream)
    hyper_mode = 'device'
    device_chunk = 64         # samples per amf_gibbs_chain_device call of the fast-mode chain (at most 256 MB)

    def samples_device(self, num_gibbs=2, fit_first=False, seed=None, hyper=None):
        '''Fast-mode chain (bayes_pmf.py:227-302 in law): yields (user_sample, item_sample) as
        DEVICE tensors of the compute dtype; the criteria (`predict`, `pred_variance`, ...) take
        them as they are.  With hyper='device' (default, d <= 32) the whole sample -- moments of
        the factor matrices, Normal-Wishart draws, half-sweeps -- is stream-ordered device work
        and the host only enqueues; hyper='host' makes the d x d draws in numpy (one device->host
        read and one host->device copy per sample).'''
        lib = N.require_device()
        name = self.dtype_name
        dt = D.np_dtype(name)
        n, m, d = self.num_users, self.num_items, self.latent_d
        rat = D.Ratings.from_tuples(self.ratings, n, m, name)
        if fit_first:
            self.do_fit()
        seed = int(self.device_seed if seed is None else seed)
        hyper = self.hyper_mode if hyper is None else hyper
        if hyper not in ('device', 'host'):
            raise ValueError("hyper must be 'device' or 'host'")
        if d > 32 or min(n, m) < 2:
            hyper = 'host'
        users_t = D.to_device(self.users, dt)
        items_t = D.to_device(self.items, dt)
        tdt = D.torch_dtype(name)
        stream_id = 0
        sample_no = 0
        dd = d * d
        hyper_h = torch.empty(2 * (d + dd), dtype=tdt).pin_memory()
        priors = []
        if hyper == 'device':
            for wi, b0, df, mu0 in (self.u_hyperparams, self.v_hyperparams):
                pr = np.concatenate((np.linalg.inv(np.atleast_2d(wi)).reshape(-1),
                                     np.atleast_1d(mu0).astype(float).reshape(-1), [float(b0), float(df)]))
                priors.append(D.to_device(pr, np.float64))

        scratch = [torch.empty((n, d), dtype=tdt, device=users_t.device),
                   torch.empty((m, d), dtype=tdt, device=users_t.device)]

        def half(side, other_t, hyper_t, rows, keep):
            # only the last round's rows are handed to the caller; earlier rounds reuse scratch
            nonlocal stream_id
            out = torch.empty((rows, d), dtype=tdt, device=other_t.device) if keep else scratch[side]
            o = side * (d + dd)
            N.check(lib.amf_gibbs_half_sweep_device_rng(
                rat.handle, side, D.code(name), d, D.ptr(other_t), D.ptr(hyper_t[o + d:o + d + dd]),
                D.ptr(hyper_t[o:o + d]), float(self.beta), float(self._mean_offset()), seed,
                stream_id, D.ptr(out), 0, -1, D.stream_ptr()))
            stream_id += 1
            return out

        def moments(t):
            t64 = t.to(torch.float64)
            mean = t64.mean(dim=0)
            c = t64 - mean
            return mean, (c.T @ c) / (t.shape[0] - 1)

        # samples are produced `chunk` at a time by ONE library call (amf_gibbs_chain_device): the
        # host does not take part in a sample at all; the failure flag is read once per chunk
        chunk = max(1, int(self.device_chunk))      # 1: drive every kernel group from here instead
        if chunk > 1:
            chunk = max(2, min(chunk, (256 << 20) // max(1, (n + m) * d * (4 if name == 'f32' else 8))))
        while hyper == 'device' and chunk > 1:
            us = torch.empty((chunk, n, d), dtype=tdt, device=users_t.device)
            vs = torch.empty((chunk, m, d), dtype=tdt, device=users_t.device)
            N.check(lib.amf_gibbs_chain_device(
                rat.handle, D.code(name), d, chunk, int(num_gibbs), D.ptr(users_t), D.ptr(items_t),
                D.ptr(priors[0]), D.ptr(priors[1]), float(self.beta), float(self._mean_offset()), seed,
                stream_id, D.ptr(us), D.ptr(vs), D.stream_ptr()))
            stream_id += chunk * (2 + 2 * int(num_gibbs))
            self._check_gibbs(rat)
            users_t, items_t = us[chunk - 1], vs[chunk - 1]
            for k in range(chunk):
                yield us[k], vs[k]
        while True:
            if hyper == 'device':      # device_chunk = 1: the same chain, one library call per kernel group
                hyper_t = torch.empty(2 * (d + dd), dtype=tdt, device=users_t.device)
                for side, feats in ((0, users_t), (1, items_t)):
                    o = side * (d + dd)
                    N.check(lib.amf_gibbs_hyper_device(
                        rat.handle, D.code(name), d, int(feats.shape[0]), D.ptr(feats),
                        D.ptr(priors[side]), seed, (1 << 40) + stream_id, D.ptr(hyper_t[o:o + d]),
                        D.ptr(hyper_t[o + d:o + d + dd]), D.stream_ptr()))
                    stream_id += 1
            else:
                # ONE device->host read per sample: means and covariances of both factor matrices
                mu_m, mu_c = moments(users_t)
                mv_m, mv_c = moments(items_t)
                mom = torch.cat((mu_m, mu_c.reshape(-1), mv_m, mv_c.reshape(-1))).cpu().numpy()
                sc_u = mom[d:d + dd].reshape(d, d) if d > 1 else np.array(float(mom[d]))
                sc_v = mom[2 * d + dd:].reshape(d, d) if d > 1 else np.array(float(mom[2 * d + dd]))
                mu_u, alpha_u = self._hyperparam_from_moments(n, mom[:d], sc_u, True)
                mu_v, alpha_v = self._hyperparam_from_moments(m, mom[d + dd:2 * d + dd], sc_v, False)
                # ... and ONE host->device copy of both sides' (mu, alpha)
                hyper_h.copy_(torch.from_numpy(np.concatenate(
                    (np.atleast_1d(mu_u), np.atleast_2d(alpha_u).reshape(-1),
                     np.atleast_1d(mu_v), np.atleast_2d(alpha_v).reshape(-1))).astype(dt)))
                hyper_t = hyper_h.to(users_t.device, non_blocking=True)
            for r in range(num_gibbs):
                users_t = half(0, items_t, hyper_t, n, r == num_gibbs - 1)
                items_t = half(1, users_t, hyper_t, m, r == num_gibbs - 1)
            sample_no += 1
            if sample_no % 16 == 0:          # the failure flag is sticky: one read covers 16 samples
                self._check_gibbs(rat)
            yield users_t, items_t

    # ---- batched lookahead chains (fast mode) ---------------------------------------------------
    def _hyperparams_batched(self, n_rows, x_bar, s_bar, do_users):
        """_hyperparam_from_moments for P chains at once (numpy batched d x d algebra, global
        numpy stream): x_bar (P, d), s_bar (P, d, d) -> mu (P, d), alpha (P, d, d)"""
        wi, b0, df, mu0 = self.u_hyperparams if do_users else self.v_hyperparams
        P, d = x_bar.shape
        n = n_rows
        diff = mu0[None, :] - x_bar
        quirk = (diff * diff).sum(1)                     # np.dot of two 1-D vectors (bayes_pmf.py:176)
        wi_post = np.linalg.inv(np.linalg.inv(wi)[None] + n * s_bar
                                + ((b0 * n) / (b0 + n)) * quirk[:, None, None])
        wi_post = wi_post / 2
        wi_post = wi_post + wi_post.transpose(0, 2, 1)
        dof = int(df + n)
        chol = np.linalg.cholesky(wi_post)
        if dof <= 81 + d:
            X = chol @ np.random.normal(size=(P, d, dof))
        else:                                            # Bartlett (bayes_pmf.py:52-58)
            A = np.zeros((P, d, d))
            A[:, np.arange(d), np.arange(d)] = np.sqrt(np.random.chisquare(dof - np.arange(d), size=(P, d)))
            lower = np.tri(d, k=-1, dtype=bool)
            A[:, lower] = np.random.normal(size=(P, d * (d - 1) // 2))
            X = chol @ A
        alpha = X @ X.transpose(0, 2, 1)
        mu_temp = (b0 * mu0[None, :] + n * x_bar) / (b0 + n)
        lam = np.linalg.cholesky(np.linalg.inv((b0 + n) * alpha))
        mu = np.einsum('pkl,pl->pk', lam, np.random.normal(size=(P, d))) + mu_temp
        return mu, alpha

    # device bytes one chunk of lookahead chains may hold (samples of every chain are kept until
    # its total variance is reduced)
    lookahead_chunk_bytes = 2 << 30

    def _lookahead_total_variance(self, cells_i, cells_j, values, num_samps, num_gibbs=2, seed=None):
        """total_variance (bayes_pmf.py:450-451) of `num_samps` samples of the chain of
        model + (i, j, v) for every (cell, value): values (ncell, nval).  All chains of a chunk run
        in the same launches (amf_gibbs_half_sweep_batched), each from this model's current
        factors; the sum over all N x M cells of the sample variance is reduced from d x d Gram
        matrices of the stacked samples, never from N x M predictions."""
        lib = N.require_device()
        name = self.dtype_name
        dt, tdt = D.np_dtype(name), D.torch_dtype(name)
        n, m, d = self.num_users, self.num_items, self.latent_d
        if d > 32:
            raise ValueError("the batched lookahead supports latent_d <= 32")
        values = np.asarray(values, dtype=float)
        ncell, nval = values.shape
        ei = np.repeat(np.asarray(cells_i, dtype=np.int32), nval)
        ej = np.repeat(np.asarray(cells_j, dtype=np.int32), nval)
        ev = values.reshape(-1)
        P_all = ei.shape[0]
        rat = D.Ratings.from_tuples(self.ratings, n, m, name)
        nnz = self.ratings.shape[0]
        seed = int(self.device_seed if seed is None else seed)
        S = int(num_samps)
        per_chain = (S + 2) * (n + m) * d * np.dtype(dt).itemsize + 2 * (S * d) ** 2 * 8
        chunk = int(max(1, min(P_all, self.lookahead_chunk_bytes // per_chain)))
        base_u, base_v = D.to_device(self.users, dt), D.to_device(self.items, dt)
        out = np.empty(P_all)
        stream_id = 1 << 32            # apart from the counters of samples_device()
        dd = d * d
        for lo in range(0, P_all, chunk):
            hi = min(P_all, lo + chunk)
            P = hi - lo
            us = base_u.unsqueeze(0).expand(P, n, d).contiguous()
            vs = base_v.unsqueeze(0).expand(P, m, d).contiguous()
            ex_i, ex_j = D.to_device(ei[lo:hi], np.int32), D.to_device(ej[lo:hi], np.int32)
            ex_v = D.to_device(ev[lo:hi], np.float64)
            offs = None
            if self.subtract_mean:       # add_rating moves the mean rating (pmf_cy.pyx:155)
                offs = D.to_device((self.mean_rating * nnz + ev[lo:hi]) / (nnz + 1), np.float64)
            keep_u = torch.empty((P, n, S * d), dtype=tdt, device=us.device)
            keep_v = torch.empty((P, m, S * d), dtype=tdt, device=us.device)

            def half(side, other_t, alpha_t, mu_t, rows):
                nonlocal stream_id
                res = torch.empty((P, rows, d), dtype=tdt, device=other_t.device)
                N.check(lib.amf_gibbs_half_sweep_batched(
                    rat.handle, side, D.code(name), d, P, D.ptr(other_t), D.ptr(alpha_t), D.ptr(mu_t),
                    float(self.beta), float(self._mean_offset()), D.ptr(offs),
                    D.ptr(ex_i if side == 0 else ex_j), D.ptr(ex_j if side == 0 else ex_i), D.ptr(ex_v),
                    seed, stream_id, D.ptr(res), D.stream_ptr()))
                stream_id += 1
                return res

            def moments(t):
                t64 = t.to(torch.float64)
                mean = t64.mean(dim=1)
                c = t64 - mean.unsqueeze(1)
                return mean, torch.bmm(c.transpose(1, 2), c) / (t.shape[1] - 1)

            for sidx in range(S):
                mu_m, mu_c = moments(us)
                mv_m, mv_c = moments(vs)
                mom = torch.cat((mu_m, mu_c.reshape(P, dd), mv_m, mv_c.reshape(P, dd)), dim=1).cpu().numpy()
                mu_u, al_u = self._hyperparams_batched(n, mom[:, :d], mom[:, d:d + dd].reshape(P, d, d), True)
                mu_v, al_v = self._hyperparams_batched(m, mom[:, d + dd:2 * d + dd],
                                                       mom[:, 2 * d + dd:].reshape(P, d, d), False)
                au, mu_ut = D.to_device(al_u, dt), D.to_device(mu_u, dt)
                av, mu_vt = D.to_device(al_v, dt), D.to_device(mu_v, dt)
                for _g in range(num_gibbs):
                    us = half(0, vs, au, mu_ut, n)
                    vs = half(1, us, av, mu_vt, m)
                keep_u[:, :, sidx * d:(sidx + 1) * d] = us
                keep_v[:, :, sidx * d:(sidx + 1) * d] = vs
            self._check_gibbs(rat)
            out[lo:hi] = _total_variance_of_stacks(keep_u, keep_v, S, d).cpu().numpy()
            del keep_u, keep_v
        rat.close()
        return out.reshape(ncell, nval)

    def samples_parallel(self, num_gibbs=2, pool=None, multiproc_mode=None, fit_first=False):
        '''(bayes_pmf.py:306-424) the row fan-out is the GPU launch; `pool` is not needed.'''
        if multiproc_mode == 'force' and pool is None:
            raise ValueError("need a process pool if multiproc is forced")
        yield from self.samples(num_gibbs=num_gibbs, fit_first=fit_first)

    ############################################################################
    ### Criteria over samples

    def _which_indices(self, which):
        n, m = self.num_users, self.num_items
        if which is None:
            which = Ellipsis
        # the two forms the drivers use (bayes_pmf.py:702-712: a pair of index arrays; a boolean
        # mask) without building the n x m index grids
        if isinstance(which, tuple) and len(which) == 2:
            ia, ja = (np.asarray(w) for w in which)
            if (ia.dtype.kind in 'iu' and ja.dtype.kind in 'iu' and ia.shape == ja.shape and ia.ndim >= 1
                    and (ia.size == 0 or (ia.min() >= 0 and ja.min() >= 0 and ia.max() < n and ja.max() < m))):
                return ia.reshape(-1), ja.reshape(-1), ia.shape
        elif isinstance(which, np.ndarray) and which.dtype == bool and which.shape == (n, m):
            ia, ja = np.nonzero(which)
            return ia, ja, ia.shape
        ii, jj = np.meshgrid(np.arange(n), np.arange(m), indexing='ij')
        i_idx, j_idx = ii[which], jj[which]
        return i_idx.reshape(-1), j_idx.reshape(-1), i_idx.shape

    # `which` covering at least this fraction of the matrix goes through the dense (blocked) form
    # of amf_bayes_sample_stats and is gathered from its n x m outputs
    _DENSE_WHICH_FRACTION = 0.125

    def _stacked_samples(self, samples, name):
        """(S, N, d) and (S, M, d) device tensors of a list of samples.  The drivers call several
        criteria on the SAME list (picker, then bayes_rmse: bayes_pmf.py:702-724), so the last
        stacking is kept and reused when the same sample objects come back."""
        tdt, dt = D.torch_dtype(name), D.np_dtype(name)
        hit = self._dev.get('stacked')
        if hit is not None and hit[0] == name and len(hit[1]) == len(samples) and \
                all(a[0] is b[0] and a[1] is b[1] for a, b in zip(hit[1], samples)):
            return hit[2], hit[3]
        if isinstance(samples[0][0], torch.Tensor):   # samples_device(): stack where they are
            us = torch.stack([u for u, _ in samples]).to(tdt).contiguous()
            vs = torch.stack([v for _, v in samples]).to(tdt).contiguous()
        else:
            us = D.to_device(np.stack([np.asarray(u) for u, _ in samples]), dt)
            vs = D.to_device(np.stack([np.asarray(v) for _, v in samples]), dt)
        self._dev['stacked'] = (name, [(u, v) for u, v in samples], us, vs)
        return us, vs

    def _sample_stats(self, samples_iter, which, cutoff=0., want=('mean', 'var', 'prob')):
        lib = N.require_device()
        name = self.dtype_name
        dt = D.np_dtype(name)
        samples = list(samples_iter)
        if not samples:
            raise StopIteration
        n, m, d = self.num_users, self.num_items, self.latent_d
        tdt = D.torch_dtype(name)
        us, vs = self._stacked_samples(samples, name)
        whole = which is None or which is Ellipsis
        if whole:
            i_idx = j_idx = None
            shape, nsel = (n, m), n * m
        else:
            i_idx, j_idx, shape = self._which_indices(which)
            nsel = i_idx.shape[0]
        dense = (whole or nsel >= self._DENSE_WHICH_FRACTION * n * m) and 98 * d * np.dtype(dt).itemsize <= 200 * 1024
        nc = n * m if dense else nsel
        if dense:
            ci = cj = None
        else:
            ci, cj = D.to_device(i_idx, np.int32), D.to_device(j_idx, np.int32)
        outs = {k: torch.empty(nc, dtype=tdt, device=us.device) if k in want else None
                for k in ('mean', 'var', 'prob')}
        N.check(lib.amf_bayes_sample_stats(
            D.code(name), nc, D.ptr(ci), D.ptr(cj), len(samples), n, m, d, D.ptr(us), D.ptr(vs),
            float(self._mean_offset()), float(cutoff),
            D.ptr(outs['mean']), D.ptr(outs['var']), D.ptr(outs['prob']), 1, 1, 0, None,
            D.stream_ptr()))
        if dense and not whole:                 # pick the wanted cells out of the dense outputs
            flat = torch.from_numpy(np.asarray(i_idx, dtype=np.int64) * m + np.asarray(j_idx, dtype=np.int64)).to(us.device)
            outs = {k: (v[flat] if v is not None else None) for k, v in outs.items()}
        return {k: v.to(torch.float64).cpu().numpy().reshape(shape)
                for k, v in outs.items() if v is not None}

    def matrix_results(self, vals, which):
        res = np.empty((self.num_users, self.num_items))
        res.fill(np.nan)
        res[which] = vals
        return res

    def predict(self, samples_iter, which=Ellipsis):
        '''Mean reconstruction over the samples (bayes_pmf.py:433-438).'''
        return self._sample_stats(samples_iter, which, want=('mean',))['mean']

    def pred_variance(self, samples_iter, which=Ellipsis):
        '''Population variance of each prediction over the samples (bayes_pmf.py:440-448).'''
        return self._sample_stats(samples_iter, which, want=('var',))['var']

    def total_variance(self, samples_iter, which=Ellipsis):
        return self.pred_variance(samples_iter, which=which).sum()

    def prob_ge_cutoff(self, samples_iter, cutoff, which=Ellipsis):
        '''Fraction of samples predicting >= cutoff (bayes_pmf.py:528-538).'''
        return self._sample_stats(samples_iter, which, cutoff=cutoff, want=('prob',))['prob']

    def random(self, samples_iter, which=Ellipsis):
        shape = np.empty((self.num_users, self.num_items))[which].shape
        return np.random.rand(*shape)

    def bayes_rmse(self, samples_iter, true_r, which=Ellipsis):
        return rmse(self.predict(samples_iter, which), true_r[which])

    def exp_variance(self, samples_iter, which=Ellipsis, pool=None, fit_first=True, num_samps=30):
        '''Expected total variance after learning each R_ij (bayes_pmf.py:457-468).'''
        return self._distribute(_exp_variance_helper, samples_iter, which, pool, fit_first, num_samps)

    def _distribute_batched(self, i_idx, j_idx, shape, discrete, params, num_samps):
        res = np.empty(shape)
        res.fill(np.nan)
        cells = [(int(i), int(j)) for i, j in zip(i_idx.flat, j_idx.flat)]
        live = [t for t, c in enumerate(cells) if c not in self.rated]
        if len(live) < len(cells):
            warnings.warn("Asked to check a known entry; returning NaN")
        if live:
            est = _batched_integration(self, [cells[t] for t in live], discrete,
                                       [params[t] for t in live], num_samps)
            res.flat[live] = np.float32(est)           # `exp = cython.float` in bayes_pmf.pxd
        return res

    def _distribute(self, fn, samples_iter, which, pool, fit_first, num_samps):
        '''(bayes_pmf.py:470-525); alpha and denom are C floats in the compiled reference.'''
        samples = list(samples_iter)
        i_idx, j_idx, shape = self._which_indices(which)
        name = self.dtype_name
        dt = D.np_dtype(name)
        # R_ij samples at the requested cells: (S, ncand)
        vals = np.stack([
            self._sample_stats([s], which, want=('mean',))['mean'].reshape(-1) for s in samples])

        if self.discrete_expectations and self.rating_values is not None:
            discrete = True
            alpha = float(np.float32(.1))
            prev_samps = vals.shape[0]
            denom = float(np.float32(prev_samps + alpha * len(self.rating_values)))
            params = [(np.histogram(v, bins=self.rating_bounds)[0] + alpha) / denom for v in vals.T]
        else:
            if self.discrete_expectations and self.rating_values is None:
                warnings.warn("have no rating_values; doing continuous")
            discrete = False
            params = list(zip(np.mean(vals, 0).flat, np.var(vals, 0).flat))

        if self.rng_mode == 'device' and fn is _exp_variance_helper:
            return self._distribute_batched(i_idx, j_idx, shape, discrete, params, num_samps)
        exps = map(fn, zip(repeat(self), i_idx.flat, j_idx.flat, repeat(discrete), params,
                           repeat(fit_first), repeat(num_samps)))
        res = np.empty(shape)
        res.fill(np.nan)
        for idx, exp in enumerate(exps):
            res.flat[idx] = np.float32(exp)          # `exp = cython.float` in bayes_pmf.pxd
        return res


def _total_variance_of_stacks(keep_u, keep_v, S, d):
    """sum over ALL cells (i, j) of the population variance over S samples of U_s[i] . V_s[j], for
    P chains at once, from the samples stacked along columns: keep_u (P, n, S d), keep_v
    (P, m, S d).  With Gu = X_u' X_u (S d x S d, blocks Gu_st = U_s' U_t):
        sum_ij Var = (1/S) sum_s <Gu_ss, Gv_ss> - (1/S^2) sum_st <Gu_st, Gv_st>
    -- O((n + m) (S d)^2) instead of S products of n x m, in fp64."""
    ku, kv = keep_u.to(torch.float64), keep_v.to(torch.float64)
    prod = torch.bmm(ku.transpose(1, 2), ku) * torch.bmm(kv.transpose(1, 2), kv)
    mask = torch.block_diag(*[torch.ones((d, d), dtype=torch.float64, device=ku.device)] * S)
    return (prod * mask.unsqueeze(0)).sum(dim=(1, 2)) / S - prod.sum(dim=(1, 2)) / (S * S)


def _batched_integration(bpmf, cells, discrete, params, num_samps):
    """_integrate_lookahead (bayes_pmf.py:560-598) for many cells at once in fast mode: every
    (cell, value) model is a chain of ONE batched launch sequence; chains start from the parent's
    factors (the reference's `fit_first` MAP re-fit per model is not done)."""
    ci = np.array([c[0] for c in cells], dtype=np.int32)
    cj = np.array([c[1] for c in cells], dtype=np.int32)
    if discrete:
        vals = np.tile(np.asarray(bpmf.rating_values, dtype=float), (len(cells), 1))
        evals = bpmf._lookahead_total_variance(ci, cj, vals, num_samps)
        return (evals * np.asarray(params)).sum(1)
    means = np.array([p[0] for p in params])
    sds = np.sqrt(np.array([p[1] for p in params]))
    q = stats.norm.ppf(np.linspace(.001, .999, bpmf.num_integration_pts))
    pts = means[:, None] + sds[:, None] * q[None, :]
    evals = bpmf._lookahead_total_variance(ci, cj, pts, num_samps)
    pdfs = stats.norm.pdf(pts, loc=means[:, None], scale=sds[:, None])
    return integrate.trapezoid(evals * pdfs, pts, axis=1)


def _integrate_lookahead(fn, bpmf, i, j, discrete, params, fit_first, num_samps):
    '''(bayes_pmf.py:560-598)'''
    i, j = int(i), int(j)
    if (i, j) in bpmf.rated:
        warnings.warn("Asked to check a known entry; returning NaN")
        return np.nan

    def calculate_fn(v):
        b = deepcopy(bpmf)
        b.add_rating(i, j, v)
        samps = b.samples(fit_first=fit_first)
        return fn(b, islice(samps, num_samps))

    if discrete:
        evals = np.array([calculate_fn(v) for v in bpmf.rating_values])
        est = (evals * params).sum()
    else:
        mean, var = params
        dist = stats.norm(loc=mean, scale=np.sqrt(var))
        pts = dist.ppf(np.linspace(.001, .999, bpmf.num_integration_pts))
        evals = np.fromiter(map(calculate_fn, pts), float, pts.size)
        est = integrate.trapezoid(evals * dist.pdf(pts), pts)
    return est


def _exp_variance_helper(args):
    return _integrate_lookahead(BayesianPMF.total_variance, *args)


################################################################################

Key = namedtuple('Key', ['nice_name', 'key_fn', 'choose_max', 'wants_pool', 'args'])

KEYS = {
    'random': Key("Random", 'random', True, False, ()),
    'pred-variance': Key("Var[R_ij]", 'pred_variance', True, False, ()),

    'exp-variance': Key("E[Var[R]]", 'exp_variance', False, True, ()),

    'pred': Key("Pred", 'predict', True, False, ()),
    'prob-ge-3.5': Key("Prob >= 3.5", 'prob_ge_cutoff', True, False, (3.5,)),
    'prob-ge-.5': Key("Prob >= .5", 'prob_ge_cutoff', True, False, (.5,)),
    'prob-ge-0': Key("Prob >= 0", 'prob_ge_cutoff', True, False, (0,)),
}


# The experiment drivers of bayes_pmf.py:675-938 (fetch_samples, full_test, compare_active, main)
# are not restated here: drivers.load("bayes_pmf", ref_dir) runs the reference's own against
# BayesianPMF / KEYS above.
