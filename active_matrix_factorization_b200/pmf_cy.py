"""Host mirror of the reference's ``pmf_cy`` module (python-pmf/pmf_cy.pyx + pmf_cy.pxd).

Same class, attributes, method names, return conventions and exceptions as the Cython
original, so the reference's driver scripts can import this module in its place; the numeric
loops run in libamf_b200 (hand-written sm_100a kernels) and there is no CPU path -- every
compute method raises if the library or the GPU is missing.

Residency: ``users`` / ``items`` / ``ratings`` stay observable and assignable numpy arrays
(callers do ``a.users, a.items = ...``).  The device copies are synchronised lazily in both
directions; inside ``fit_lls`` the whole line search runs on the device and only three
scalars per trial cross PCIe.
"""
import itertools
import random
import warnings
from copy import deepcopy

import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import device as D


def rmse(exp, obs):
    """cpdef float rmse(...) -- note the C float return type (pmf_cy.pyx:25-26)."""
    return float(np.float32(np.sqrt(np.mean((np.asarray(obs) - np.asarray(exp)) ** 2))))


def rmse_on(exp, obs, on):
    """(pmf_cy.pyx:28-29)"""
    return float(np.float32(np.sqrt(np.mean((obs[on] - exp[on]) ** 2))))


class ProbabilisticMatrixFactorization(object):
    # attribute surface of pmf_cy.pxd:7-15 ; defaults of __cinit__ (pmf_cy.pyx:35-47)
    compute_dtype = None      # None -> device.default_dtype() ('f64' parity / 'f32' fast)

    def __init__(self, rating_tuples, latent_d=1, subtract_mean=False, knowable=None,
                 fit_type=('batch',)):
        if rating_tuples is None:
            raise TypeError("Argument 'rating_tuples' must not be None")
        self.learning_rate = 1e-4
        self.min_learning_rate = 1e-10
        self.stop_thresh = 1e-2
        self.sigma_sq = 1.
        self.sigma_u_sq = 10.
        self.sigma_v_sq = 10.
        self.sig_u_mean = self.sig_v_mean = 0.
        self.sig_u_var = self.sig_v_var = -1.

        self.latent_d = int(latent_d)
        self.subtract_mean = bool(subtract_mean)
        if fit_type is None:
            warnings.warn('passed None fit_type; using batch')
            fit_type = ('batch',)
        self.fit_type = tuple(fit_type)

        self._dev = {}                 # device-side state, never pickled
        ratings = np.asarray(rating_tuples, dtype=float)
        if ratings.ndim != 2 or ratings.shape[1] != 3:
            raise TypeError("invalid rating tuple length")
        self.ratings = ratings
        self.mean_rating = float(np.mean(ratings[:, 2]))

        self.num_users = n = int(np.max(ratings[:, 0]) + 1)
        self.num_items = m = int(np.max(ratings[:, 1]) + 1)

        self.rated = set(zip(ratings[:, 0].astype(int).tolist(), ratings[:, 1].astype(int).tolist()))
        if knowable is None:
            knowable = itertools.product(range(n), range(m))
        self.unrated = set(knowable).difference(self.rated)

        # same two draws, same order, as pmf_cy.pyx:74-75
        self.users = np.random.random((n, self.latent_d))
        self.items = np.random.random((m, self.latent_d))

    @classmethod
    def from_coo(cls, i, j, r, num_users, num_items, latent_d=1, subtract_mean=False,
                 fit_type=('batch',), init=None):
        """Large-data constructor: the rating list is given as three arrays
        (numpy or CUDA torch tensors: user ids, item ids, values) and goes straight into the
        device layout -- no (nnz, 3) float64 array, no Python sets of all N*M cells (the
        reference's constructor is O(N*M), pmf_cy.pyx:69-72).  ``ratings`` is materialised
        lazily only if somebody reads it; ``rated`` / ``unrated`` are empty (pass explicit
        candidate pools).  ``init`` = (users, items) or None for the reference's U(0,1) draws."""
        import torch
        self = cls.__new__(cls)
        self._dev = {}
        self.learning_rate, self.min_learning_rate, self.stop_thresh = 1e-4, 1e-10, 1e-2
        self.sigma_sq, self.sigma_u_sq, self.sigma_v_sq = 1., 10., 10.
        self.sig_u_mean = self.sig_v_mean = 0.
        self.sig_u_var = self.sig_v_var = -1.
        self.latent_d, self.subtract_mean = int(latent_d), bool(subtract_mean)
        self.fit_type = tuple(fit_type)
        self.num_users, self.num_items = int(num_users), int(num_items)
        name = self.dtype_name
        if isinstance(i, torch.Tensor):
            ti, tj = i.to(torch.int32), j.to(torch.int32)
            tr = r.to(D.torch_dtype(name))
        else:
            ti, tj = D.to_device(np.asarray(i), np.int32), D.to_device(np.asarray(j), np.int32)
            tr = D.to_device(np.asarray(r), D.np_dtype(name))
        rat = D.Ratings(self.num_users, self.num_items, ti, tj, tr, name)
        self._coo = (ti, tj, tr)
        self._ratings = None
        self._dev['rat'] = rat
        self._dev['rat_key'] = (name, 'coo', rat.nnz)
        self.mean_rating = rat.mean()
        self.rated, self.unrated = set(), set()
        if init is None:
            self.users = np.random.random((self.num_users, self.latent_d))
            self.items = np.random.random((self.num_items, self.latent_d))
        else:
            self.users, self.items = init
        return self

    @classmethod
    def from_coo_file(cls, path, latent_d=1, subtract_mean=False, **kw):
        """``from_coo`` on the rating list of a data file (see ``load_coo`` for the formats)."""
        i, j, r, n, m = load_coo(path)
        return cls.from_coo(i, j, r, n, m, latent_d, subtract_mean, **kw)

    # ---- host <-> device bookkeeping -----------------------------------------------------
    @property
    def dtype_name(self):
        return self.compute_dtype or D.default_dtype()

    @property
    def ratings(self):
        if self._ratings is None and self.__dict__.get('_coo') is not None:
            ti, tj, tr = self._coo            # from_coo(): build the (nnz, 3) array on demand
            self._ratings = np.column_stack((ti.cpu().numpy(), tj.cpu().numpy(),
                                             tr.double().cpu().numpy())).astype(float)
        return self._ratings

    def _num_ratings(self):
        if self._ratings is None and self.__dict__.get('_coo') is not None:
            return int(self._coo[0].numel())
        return self._ratings.shape[0]

    @ratings.setter
    def ratings(self, value):
        self._ratings = value
        self._coo = None
        self._drop_device('rat')

    @property
    def users(self):
        self._pull()
        return self._users

    @users.setter
    def users(self, value):
        self._pull()
        self._users = value
        self._dev['host_set'] = True

    @property
    def items(self):
        self._pull()
        return self._items

    @items.setter
    def items(self, value):
        self._pull()
        self._items = value
        self._dev['host_set'] = True

    def _drop_device(self, *keys):
        dev = self.__dict__.get('_dev')
        if dev is None:
            self._dev = dev = {}
        for k in keys:
            old = dev.pop(k, None)
            if k == 'rat' and old is not None:
                old.close()

    def _pull(self):
        """If a device-resident fit holds newer factors than the host arrays, download them."""
        dev = self.__dict__.get('_dev')
        if dev and dev.get('host_stale'):
            dev['host_stale'] = False
            self._users = D.from_padded(dev['U'], self.latent_d)
            self._items = D.from_padded(dev['V'], self.latent_d)

    def _rating_handle(self):
        if self._ratings is None and self.__dict__.get('_coo') is not None:
            if self._dev.get('rat') is None or self._dev.get('rat_key', (None,))[0] != self.dtype_name:
                ti, tj, tr = self._coo
                self._drop_device('rat')
                self._dev['rat'] = D.Ratings(self.num_users, self.num_items, ti, tj,
                                             tr.to(D.torch_dtype(self.dtype_name)), self.dtype_name)
                self._dev['rat_key'] = (self.dtype_name, 'coo', self._dev['rat'].nnz)
            return self._dev['rat']
        key = (self.dtype_name, id(self._ratings), self._ratings.shape[0])
        rat = self._dev.get('rat')
        if rat is None or self._dev.get('rat_key') != key:
            self._drop_device('rat')
            rat = D.Ratings.from_tuples(self._ratings, self.num_users, self.num_items, self.dtype_name)
            self._dev['rat'] = rat
            self._dev['rat_key'] = key
        return rat

    def _params(self):
        return D.pmf_params(self.sigma_sq, self.sigma_u_sq, self.sigma_v_sq,
                            self.mean_rating if self.subtract_mean else 0.)

    def _check_factors(self, users, items):
        d = self.latent_d
        if users.shape != (self.num_users, d) or items.shape != (self.num_items, d):
            raise ValueError("factor matrices have shapes %r, %r; expected %r, %r" % (
                users.shape, items.shape, (self.num_users, d), (self.num_items, d)))

    # ---- copying / pickling (pmf_cy.pyx:78-126) ------------------------------------------
    def __copy__(self):
        res = type(self)(self.ratings, self.latent_d)
        res.__setstate__(res.__getstate__())
        return res

    def __deepcopy__(self, memodict):
        res = type(self)(self.ratings.copy())
        res.__setstate__(deepcopy(self.__getstate__(), memodict))
        return res

    def __setstate__(self, state):
        if state is None:
            raise TypeError("Argument 'state' must not be None")
        if '_dev' not in self.__dict__:
            self._dev = {}
        for k, v in state.items():
            if k == '__dict__':
                for real_k, real_v in v.items():
                    if real_k != '_dev':
                        setattr(self, real_k, real_v)
            else:
                setattr(self, k, v)

    def __getstate__(self):
        return dict(
            latent_d=self.latent_d, num_users=self.num_users, num_items=self.num_items,
            learning_rate=self.learning_rate, min_learning_rate=self.min_learning_rate,
            stop_thresh=self.stop_thresh,
            sigma_sq=self.sigma_sq, sigma_u_sq=self.sigma_u_sq, sigma_v_sq=self.sigma_v_sq,
            fit_type=self.fit_type,
            sig_u_mean=self.sig_u_mean, sig_u_var=self.sig_u_var,
            sig_v_mean=self.sig_v_mean, sig_v_var=self.sig_v_var,
            ratings=self.ratings, users=self.users, items=self.items,
            subtract_mean=self.subtract_mean, mean_rating=self.mean_rating,
            rated=self.rated, unrated=self.unrated,
        )

    def __reduce__(self):
        return (_rebuild, (type(self), self.__getstate__()))

    # ---- ratings bookkeeping (pmf_cy.pyx:128-155) ----------------------------------------
    def add_rating(self, i, j, rating):
        self.add_ratings([int(i), int(j), float(rating)])

    def add_ratings(self, extra):
        cols = 3 if self._ratings is None else self._ratings.shape[1]
        extra = np.array(extra, ndmin=2)
        if len(extra.shape) != 2 or extra.shape[1] != cols:
            raise TypeError("bad shape for extra")
        assert np.max(extra[:, 0] + 1) <= self.num_users
        assert np.max(extra[:, 1] + 1) <= self.num_items

        rating_vals = getattr(self, 'rating_values', None)
        if rating_vals is not None:
            if not set(rating_vals).issuperset(extra[:, 2]):
                raise ValueError("got ratings with bad values")

        new_items = set((int(i), int(j)) for i, j in extra[:, :2])
        if not new_items.isdisjoint(self.rated):
            raise ValueError("can't rate already rated items")
        self.rated.update(new_items)
        self.unrated.difference_update(new_items)
        # device-resident candidate pools this model has scored (scoring.CandidatePool): the
        # queried cells leave them too, as they leave `unrated` (pmf_cy.pyx:152)
        for pool in list(self.__dict__.get('_dev', {}).get('candidate_pools', ())):
            for i, j in new_items:
                pool.remove(i, j)

        # active-loop residency: a rating list that already lives on the device
        # takes the new ratings as an appended tail instead of being re-uploaded and re-sorted
        rat = self.__dict__.get('_dev', {}).get('rat')
        key = self._dev.get('rat_key') if rat is not None else None
        coo = self.__dict__.get('_coo')
        if coo is not None and self._ratings is None:
            # from_coo() model: no host (nnz, 3) array is kept; the device list is the list
            import torch
            ti = torch.as_tensor(extra[:, 0].astype(np.int32)).to(coo[0].device)
            tj = torch.as_tensor(extra[:, 1].astype(np.int32)).to(coo[0].device)
            tr = torch.as_tensor(extra[:, 2]).to(coo[2].dtype).to(coo[0].device)
            total = float(self.mean_rating) * int(coo[0].numel()) + float(extra[:, 2].sum())
            self._coo = (torch.cat((coo[0], ti)), torch.cat((coo[1], tj)), torch.cat((coo[2], tr)))
            if rat is not None and key is not None and key[0] == self.dtype_name:
                rat.append(ti, tj, tr)
                self._dev['rat_key'] = (key[0], 'coo', rat.nnz)
            else:
                self._drop_device('rat')
            self.mean_rating = total / int(self._coo[0].numel())
            return
        old = self._ratings
        new_ratings = np.append(old, extra, 0)
        if rat is not None and key == (self.dtype_name, id(old), old.shape[0]):
            rat.append(extra[:, 0].astype(np.int32), extra[:, 1].astype(np.int32), extra[:, 2])
            self._ratings = new_ratings            # not through the setter: the handle stays
            self._dev['rat_key'] = (self.dtype_name, id(new_ratings), new_ratings.shape[0])
        else:
            self.ratings = new_ratings
        self.mean_rating = float(np.mean(self._ratings[:, 2]))

    # ---- numerics ------------------------------------------------------------------------
    def prediction_for(self, i, j, users=None, items=None):
        """(pmf_cy.pyx:158-168) -- single cell; goes through the candidate-scoring kernel."""
        from .scoring import score_pred
        users = self.users if users is None else users
        items = self.items if items is None else items
        val = score_pred(users, items, [int(i)], [int(j)], self.dtype_name)[0][0]
        return float(val + self.mean_rating) if self.subtract_mean else float(val)

    def _loss_grad_host(self, users, items, want_grad):
        lib = N.require_device()
        name = self.dtype_name
        users = np.ascontiguousarray(users, dtype=D.np_dtype(name))
        items = np.ascontiguousarray(items, dtype=D.np_dtype(name))
        self._check_factors(users, items)
        rat = self._rating_handle()
        gu = np.empty_like(users) if want_grad else None
        gv = np.empty_like(items) if want_grad else None
        sums = np.zeros(3)
        params = self._params()
        import ctypes as C
        N.check(lib.amf_pmf_loss_grad_host(
            rat.handle, D.code(name), self.latent_d, N.host_ptr(users), N.host_ptr(items),
            C.byref(params), N.host_ptr(gu) if want_grad else None,
            N.host_ptr(gv) if want_grad else None, N.host_ptr(sums)))
        return sums, gu, gv

    def log_likelihood(self, users=None, items=None):
        """(pmf_cy.pyx:170-193)"""
        users = self.users if users is None else users
        items = self.items if items is None else items
        sums, _, _ = self._loss_grad_host(users, items, False)
        return float(D.log_likelihood_from_sums(sums, self.sigma_sq, self.sigma_u_sq, self.sigma_v_sq))

    def ll_prior_adjustment(self):
        """(pmf_cy.pyx:195-199)"""
        return float(-.5 * (
            np.log(self.sigma_sq) * self._num_ratings()
            + self.num_users * self.latent_d * np.log(self.sigma_u_sq)
            + self.num_items * self.latent_d * np.log(self.sigma_v_sq)))

    def full_ll(self, users=None, items=None):
        return self.log_likelihood(users, items) + self.ll_prior_adjustment()

    def gradient(self, ratings=None):
        """(pmf_cy.pyx:204-223); ``ratings`` may be an explicit mini-batch."""
        if ratings is None or ratings is self.ratings:
            _, gu, gv = self._loss_grad_host(self.users, self.items, True)
            return gu.astype(np.float64, copy=False), gv.astype(np.float64, copy=False)
        name = self.dtype_name
        U = D.to_padded(self.users, name)
        V = D.to_padded(self.items, name)
        gu, gv = self._batch_gradient_device(np.asarray(ratings, dtype=float), U, V)
        return D.from_padded(gu, self.latent_d), D.from_padded(gv, self.latent_d)

    def _batch_gradient_device(self, batch, U, V, coo=None):
        """prior + COO data term for a mini-batch, on device tensors."""
        import ctypes as C
        lib = N.require_device()
        name = self.dtype_name
        if coo is None:
            coo = (D.to_device(batch[:, 0], np.int32), D.to_device(batch[:, 1], np.int32),
                   D.to_device(batch[:, 2], D.np_dtype(name)))
        bi, bj, br = coo
        gu, gv = torch.empty_like(U), torch.empty_like(V)
        st = D.stream_ptr()
        N.check(lib.amf_pmf_prior(D.code(name), U.numel(), D.ptr(U), self.sigma_u_sq, D.ptr(gu), None, st))
        N.check(lib.amf_pmf_prior(D.code(name), V.numel(), D.ptr(V), self.sigma_v_sq, D.ptr(gv), None, st))
        params = self._params()
        N.check(lib.amf_pmf_grad_coo(D.code(name), bi.numel(), D.ptr(bi), D.ptr(bj), D.ptr(br),
                                     self.latent_d, U.shape[1], D.ptr(U), D.ptr(V), C.byref(params),
                                     D.ptr(gu), D.ptr(gv), None, st))
        return gu, gv

    def _sq_error(self):
        sums, _, _ = self._loss_grad_host(self.users, self.items, False)
        return float(sums[0])

    def update_sigma(self):
        """(pmf_cy.pyx:225-234)"""
        self.sigma_sq = self._sq_error() / self._num_ratings()

    def update_sigma_uv(self):
        """(pmf_cy.pyx:236-255)"""
        d, n, m = self.latent_d, self.num_users, self.num_items
        sums, _, _ = self._loss_grad_host(self.users, self.items, False)   # |U|^2, |V|^2 on device
        user_norm2, item_norm2 = float(sums[1]), float(sums[2])
        if self.sig_u_var > 0:
            self.sigma_u_sq = user_norm2 / (n * d + 2 + 2 * (np.log(self.sigma_u_sq) - self.sig_u_mean) / self.sig_u_var)
        else:
            self.sigma_u_sq = user_norm2 / n / d
        if self.sig_v_var > 0:
            self.sigma_v_sq = item_norm2 / (m * d + 2 + 2 * (np.log(self.sigma_v_sq) - self.sig_v_mean) / self.sig_v_var)
        else:
            self.sigma_v_sq = item_norm2 / m / d

    def fit_lls(self):
        """Line-search gradient ascent (pmf_cy.pyx:257-291), device resident.

        Per trial: X' = X + lr*G, then ONE fused pass gives both LL(X') and the gradient at X';
        an accepted trial therefore already holds the next iteration's gradient.  Only the
        three objective sums cross PCIe.  Yields the log-likelihood of each accepted step.
        """
        name = self.dtype_name
        d = self.latent_d
        rat = self._rating_handle()
        dev = self._dev
        self._check_factors(self._users, self._items)
        U, V = D.to_padded(self._users, name), D.to_padded(self._items, name)
        gU, gV = torch.empty_like(U), torch.empty_like(V)
        U2, V2, gU2, gV2 = (torch.empty_like(U), torch.empty_like(V),
                            torch.empty_like(U), torch.empty_like(V))
        sums = torch.empty(3, dtype=torch.float64, device=U.device)

        def ll_of(s):
            s = s.cpu().numpy()
            return float(D.log_likelihood_from_sums(s, self.sigma_sq, self.sigma_u_sq, self.sigma_v_sq))

        lr = self.learning_rate
        dev.pop('host_set', None)
        D.loss_grad(rat, d, U, V, self._params(), gU, gV, sums)
        old_ll = ll_of(sums)
        hyper = (self.sigma_sq, self.sigma_u_sq, self.sigma_v_sq)
        rat_nnz = rat.nnz

        converged = False
        while not converged:
            while True:
                D.axpy(U, gU, lr, U2, name)
                D.axpy(V, gV, lr, V2, name)
                D.loss_grad(rat, d, U2, V2, self._params(), gU2, gV2, sums)
                new_ll = ll_of(sums)
                if new_ll > old_ll:
                    U, U2, V, V2 = U2, U, V2, V
                    gU, gU2, gV, gV2 = gU2, gU, gV2, gV
                    dev['U'], dev['V'], dev['host_stale'] = U, V, True
                    lr *= 1.25
                    if new_ll - old_ll < self.stop_thresh:
                        converged = True
                    yield new_ll
                    old_ll = new_ll
                    # the caller may have changed hyper-parameters, factors or ratings between
                    # steps (fit_with_sigmas_lls does).  The reference then takes its next
                    # gradient from the new state but keeps comparing against the objective it
                    # yielded (pmf_cy.pyx:265,284-285: old_ll is never re-evaluated), so only
                    # the gradient is refreshed here
                    # (add_ratings appends to the SAME device list object: compare its length too)
                    changed = dev.get('rat') is not rat or rat.nnz != rat_nnz
                    if dev.pop('host_set', False):
                        self._pull()
                        U, V = D.to_padded(self._users, name), D.to_padded(self._items, name)
                        changed = True
                    if changed or hyper != (self.sigma_sq, self.sigma_u_sq, self.sigma_v_sq):
                        rat = self._rating_handle()
                        rat_nnz = rat.nnz
                        D.loss_grad(rat, d, U, V, self._params(), gU, gV, sums)
                    hyper = (self.sigma_sq, self.sigma_u_sq, self.sigma_v_sq)
                    break
                else:
                    lr *= .5
                    if lr < self.min_learning_rate:
                        converged = True
                        break
        self._pull()

    # problems up to this size run the whole line search in one launch (csrc/fit.cu)
    _DEVICE_FIT_MAX_NNZ = 200_000
    _DEVICE_FIT_MAX_TABLE = 2_000_000

    def fit(self):
        """(pmf_cy.pyx:293-295) run fit_lls to convergence.  Small problems (the reference's own
        sizes) do the entire line search on the device in one launch -- same control flow, no
        host round trip per trial (amf_pmf_fit_lls); larger ones drive the fused pass per trial."""
        name = self.dtype_name
        ld = D.padded_ld(self.latent_d, name)
        if (self._num_ratings() > self._DEVICE_FIT_MAX_NNZ or
                (self.num_users + self.num_items) * ld > self._DEVICE_FIT_MAX_TABLE):
            for _ll in self.fit_lls():
                pass
            return
        lib = N.require_device()
        rat = self._rating_handle()
        self._check_factors(self._users, self._items)
        U, V = D.to_padded(self._users, name), D.to_padded(self._items, name)
        nbytes = int(lib.amf_pmf_fit_workspace_bytes(rat.handle, D.code(name), ld))
        work = torch.empty(nbytes, dtype=torch.uint8, device=U.device)
        result = torch.zeros(4, dtype=torch.float64, device=U.device)       # amf_fit_result_t
        params = self._params()
        N.check(lib.amf_pmf_fit_lls(rat.handle, D.code(name), self.latent_d, ld, D.ptr(U), D.ptr(V),
                                    C.byref(params), float(self.learning_rate),
                                    float(self.min_learning_rate), float(self.stop_thresh), 0,
                                    None, 0, D.ptr(result), D.ptr(work), nbytes, D.stream_ptr()))
        self._dev.pop('host_set', None)
        self._dev['U'], self._dev['V'], self._dev['host_stale'] = U, V, True
        self._pull()

    def do_fit(self):
        kind, *args = self.fit_type
        if kind == 'batch':
            self.fit(*args)
        elif kind == 'mini-valid':
            self.fit_minibatches_until_validation(*args)
        else:
            raise ValueError("unknown fit type '{}'".format(kind))

    def fit_minibatches(self, batch_size, lr=1, momentum=.8, ratings=None):
        """Momentum SGD (pmf_cy.pyx:308-351).  lr and momentum are C floats in the reference;
        the shuffle draws from the global numpy RNG and permutes ``ratings`` in place."""
        import ctypes as C
        lib = N.require_device()
        batch_size = int(batch_size)
        lr = float(np.float32(lr))
        momentum = float(np.float32(momentum))
        if ratings is None:
            ratings = self.ratings
        num_ratings = ratings.shape[0]
        name = self.dtype_name
        code = D.code(name)
        dev = self._dev

        U, V = D.to_padded(self.users, name), D.to_padded(self.items, name)
        dev.pop('host_set', None)
        u_inc, v_inc = torch.zeros_like(U), torch.zeros_like(V)
        while True:
            np.random.shuffle(ratings)
            if ratings is self._ratings:
                self._drop_device('rat')
            bi = D.to_device(ratings[:, 0], np.int32)
            bj = D.to_device(ratings[:, 1], np.int32)
            br = D.to_device(ratings[:, 2], D.np_dtype(name))
            st = D.stream_ptr()
            for start in range(0, num_ratings, batch_size):
                end = min(start + batch_size, num_ratings)
                n = end - start
                gu, gv = self._batch_gradient_device(None, U, V, (bi[start:end], bj[start:end], br[start:end]))
                step = float(np.float32(lr) / np.float32(n))   # C float / C int in the reference
                N.check(lib.amf_momentum_step(code, U.numel(), D.ptr(u_inc), D.ptr(gu), momentum, step, D.ptr(U), st))
                N.check(lib.amf_momentum_step(code, V.numel(), D.ptr(v_inc), D.ptr(gv), momentum, step, D.ptr(V), st))
            dev['U'], dev['V'], dev['host_stale'] = U, V, True
            # training error over ALL of self.ratings (pmf_cy.pyx:347-349)
            sums = D.loss_grad(self._rating_handle(), self.latent_d, U, V, self._params())
            err = float(np.float32(np.sqrt(float(sums[0].item()) / self._num_ratings())))
            yield err
            if dev.pop('host_set', False):     # caller replaced the factors
                self._pull()
                U, V = D.to_padded(self._users, name), D.to_padded(self._items, name)

    def fit_minibatches_validation(self, batch_size, valid_size, **kwargs):
        """(pmf_cy.pyx:353-372)"""
        from .scoring import score_pred
        total = self.ratings.shape[0]
        valid_idx_set = set(random.sample(range(total), int(valid_size)))
        train_idx = tuple(i for i in range(total) if i not in valid_idx_set)
        train = self.ratings[train_idx, :]
        valid_idx = list(valid_idx_set)
        vi = self.ratings[valid_idx, 0].astype(int)
        vj = self.ratings[valid_idx, 1].astype(int)
        valid_real = self.ratings[valid_idx, 2]
        for train_err in self.fit_minibatches(batch_size, ratings=train, **kwargs):
            valid_pred = score_pred(self.users, self.items, vi, vj, self.dtype_name)[0]
            if self.subtract_mean:
                valid_pred = valid_pred + self.mean_rating
            valid_err = float(np.float32(np.sqrt(np.mean((valid_pred - valid_real) ** 2))))
            yield train_err, valid_err

    def fit_minibatches_until_validation(self, *args, stop_thresh=1e-3, **kw):
        """(pmf_cy.pyx:374-381); the comparisons are made in C floats there."""
        last_valid = np.float32(np.inf)
        for _train, valid in self.fit_minibatches_validation(*args, **kw):
            valid = np.float32(valid)
            if valid > last_valid - stop_thresh:
                break
            last_valid = valid

    def fit_with_sigmas_lls(self, noise_every=5, users_every=2):
        """(pmf_cy.pyx:384-403)"""
        cont = True
        while cont:
            cont = False
            for i, ll in enumerate(self.fit_lls()):
                if i % noise_every == 0:
                    self.update_sigma()
                if i % users_every == 0:
                    self.update_sigma_uv()
                yield ll
                cont = True
            self.update_sigma()
            self.update_sigma_uv()

    def fit_with_sigmas(self, noise_every=10, users_every=5):
        for _ll in self.fit_with_sigmas_lls(noise_every, users_every):
            pass

    def predicted_matrix(self, u=None, v=None):
        """Dense U V' (+ mean) (pmf_cy.pyx:410-420), fp64: dense_pred_kernel (csrc/dense.cu)."""
        u = self.users if u is None else u
        v = self.items if v is None else v
        lib = N.require_device()
        ut, vt = D.to_device(u, np.float64), D.to_device(v, np.float64)
        n, d = ut.shape
        m = vt.shape[0]
        out = torch.empty((n, m), dtype=torch.float64, device=ut.device)
        N.check(lib.amf_predicted_matrix(N.F64, n, m, d, d, D.ptr(ut), D.ptr(vt),
                                         float(self.mean_rating) if self.subtract_mean else 0.0,
                                         D.ptr(out), D.stream_ptr()))
        return out.cpu().numpy()

    def rmse(self, real, on=None):
        """(pmf_cy.pyx:422-426): squared error summed over the selected cells by one fused kernel
        (no N x M prediction matrix); C float result"""
        lib = N.require_device()
        real = np.ascontiguousarray(real, dtype=np.float64)
        n, m = self.num_users, self.num_items
        mask = None
        if on is not None:
            on_arr = np.asarray(on) if not isinstance(on, tuple) else None
            if on_arr is not None and on_arr.dtype == bool and on_arr.shape == real.shape:
                mask = on_arr
            else:
                # index arrays may name a cell more than once: count cells the way real[on] does
                picked = np.zeros(real.shape, dtype=np.int64)
                np.add.at(picked, on, 1)
                if picked.max(initial=0) > 1:
                    pred = self.predicted_matrix()
                    return float(np.float32(np.sqrt(np.mean((real[on] - pred[on]) ** 2))))
                mask = picked.astype(bool)
        ut, vt = D.to_device(self.users, np.float64), D.to_device(self.items, np.float64)
        real_t = D.to_device(real, np.float64)
        mask_t = D.to_device(mask.astype(np.uint8), np.uint8) if mask is not None else None
        sums = torch.empty(2, dtype=torch.float64, device=ut.device)
        N.check(lib.amf_sq_error_dense(N.F64, n, m, self.latent_d, self.latent_d, D.ptr(ut), D.ptr(vt),
                                       float(self.mean_rating) if self.subtract_mean else 0.0,
                                       D.ptr(real_t), D.ptr(mask_t), D.ptr(sums), D.stream_ptr()))
        sq, cnt = sums.cpu().numpy()
        return float(np.float32(np.sqrt(sq / cnt)))

    def print_latent_vectors(self):
        print("Users:")
        for i in range(self.num_users):
            print("%d: %s" % (i, self.users[i, :]))
        print("\nItems:")
        for j in range(self.num_items):
            print("%d: %s" % (j, self.items[j, :]))

    def save_latent_vectors(self, prefix):
        self.users.dump(prefix + "%sd_users.pickle" % self.latent_d)
        self.items.dump(prefix + "%sd_items.pickle" % self.latent_d)


def load_coo(path):
    """Rating list of a data file as (i, j, r, num_users, num_items) for ``from_coo`` -- the
    large-data companion of the reference's ``.npz`` / ``.pkl`` dictionaries
    (choose_training.py:215-259, active_pmf.py:1200-1219), which always pass through an (nnz, 3)
    float64 array.  Accepted: ``.npz`` with arrays ``i, j, r`` (any integer / float dtypes; loaded without
    unpickling) and optionally ``shape``; ``.npz`` / ``.pkl``
    with the reference's ``_ratings`` (nnz, 3) table and optionally ``_real`` (its shape gives the
    matrix shape); ``.npy`` with an (nnz, 3) table; anything else is read as whitespace-
    separated ``i j r`` text lines (``.npy`` tables are memory-mapped; numpy cannot map ``.npz``
    members).  Only the reference's ``_ratings`` dictionaries are unpickled -- load those from
    trusted files only.  Without an explicit shape it is ``max id + 1`` per side, as in
    the reference's constructor (pmf_cy.pyx:65-66).  Ids must be non-negative."""
    import pickle
    shape = None
    if path.endswith('.npz') or path.endswith('.pkl'):
        if path.endswith('.pkl'):
            with open(path, 'rb') as f:
                data = pickle.load(f)
        else:
            data = np.load(path, allow_pickle=False)
            if '_ratings' in data.files and not {'i', 'j', 'r'} <= set(data.files):
                data = np.load(path, allow_pickle=True)     # the reference's dictionaries hold objects
        keys = set(data.keys())
        if {'i', 'j', 'r'} <= keys:
            i, j, r = data['i'], data['j'], data['r']
            if 'shape' in keys:
                shape = tuple(int(x) for x in np.asarray(data['shape']).reshape(-1)[:2])
        elif '_ratings' in keys:
            table = np.asarray(data['_ratings'])
            i, j, r = table[:, 0], table[:, 1], table[:, 2]
            if '_real' in keys and data['_real'] is not None and np.ndim(data['_real']) == 2:
                shape = tuple(np.shape(data['_real']))
        else:
            raise ValueError("%s holds neither i/j/r arrays nor a _ratings table" % path)
    elif path.endswith('.npy'):
        table = np.load(path, mmap_mode='r')
        if table.ndim != 2 or table.shape[1] != 3:
            raise TypeError("invalid rating tuple length")
        i, j, r = table[:, 0], table[:, 1], table[:, 2]
    else:
        table = np.loadtxt(path, ndmin=2)
        if table.shape[1] != 3:
            raise TypeError("invalid rating tuple length")
        i, j, r = table[:, 0], table[:, 1], table[:, 2]
    i, j, r = np.asarray(i), np.asarray(j), np.ascontiguousarray(r)
    if not (i.shape == j.shape == r.shape and i.ndim == 1):
        raise TypeError("i, j, r must be three vectors of one length")
    # range checks on the ORIGINAL dtype: a cast to int32 first would wrap ids >= 2^31
    if i.size and (i.min() < 0 or j.min() < 0):
        raise ValueError("negative user / item id")
    if shape is None:
        shape = (int(i.max()) + 1 if i.size else 0, int(j.max()) + 1 if j.size else 0)
    if i.size and (i.max() >= shape[0] or j.max() >= shape[1]):
        raise ValueError("ids outside the %d x %d matrix" % shape)
    if max(shape) > np.iinfo(np.int32).max:
        raise ValueError("matrix sides above 2^31 - 1 are not supported")
    i = np.ascontiguousarray(i).astype(np.int32, copy=False)
    j = np.ascontiguousarray(j).astype(np.int32, copy=False)
    return i, j, r, int(shape[0]), int(shape[1])


def _rebuild(cls, state):
    obj = cls.__new__(cls)
    obj._dev = {}
    obj.__setstate__(state)
    return obj


def parse_fit_type(string):
    """'mini-valid,100,30' -> ('mini-valid', 100, 30)   (pmf_cy.pyx:444-456)"""
    out = []
    for part in string.split(','):
        for conv in (int, float):
            try:
                out.append(conv(part))
                break
            except ValueError:
                continue
        else:
            out.append(part)
    return tuple(out)


def fake_ratings(noise=.25, num_users=100, num_items=100, num_ratings=30, latent_dimension=10):
    """Synthetic low-rank ratings, same draw order as pmf_cy.pyx:461-474."""
    u = np.random.normal(0, 2, (num_users, latent_dimension))
    v = np.random.normal(0, 2, (num_items, latent_dimension))
    ratings = []
    for i in range(num_users):
        for j in random.sample(range(num_items), num_ratings):
            ratings.append((i, j, np.dot(u[i], v[j]) + np.random.normal(scale=noise)))
    return (np.array(ratings), u, v)
