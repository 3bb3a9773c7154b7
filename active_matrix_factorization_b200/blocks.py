"""Scalable-mode variational posterior: the reference's KL objective (active_pmf.py:202-240)
restricted to block-diagonal covariances -- one d x d block per user row / item column -- and
its lookahead (active_pmf.py:635-704) by local re-fits on the device (csrc/blocks.cu).

The reference's full covariance is k x k with k = (N+M)d (active_pmf.py:136,190-200): 2595^2 at
the drugbank configuration, 26,250^2 at movielens-100k; every `fit_normal` step factorises it.
Here the state is O((N+M) d^2):

    BlockDiagonal   host container (numpy, picklable): means, covariance blocks, precisions and
                    natural parameters; stands in for the `cov` attribute of ActivePMF
    BlockPosterior  the same tables resident on the device + the kernels over them

Nothing here runs on the CPU: fitting, scoring and lookahead are launches of libamf_b200.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import device as D


class BlockDiagonal(object):
    """Block-diagonal Gaussian over (U, V): what `ActivePMF.cov` holds in scalable mode.

    `A[i]` = Cov(U_i), `B[j]` = Cov(V_j) (d x d); `Lu/Lv` their inverses (precisions) and
    `hu/hv` = precision @ mean, kept because the lookahead updates them by one rating."""

    def __init__(self, mean_u, mean_v, A, B, Lu, Lv, hu, hv, logdet_u, logdet_v):
        self.mean_u, self.mean_v, self.A, self.B = mean_u, mean_v, A, B
        self.Lu, self.Lv, self.hu, self.hv = Lu, Lv, hu, hv
        self.logdet_u, self.logdet_v = logdet_u, logdet_v

    @property
    def shape(self):
        k = self.A.shape[0] * self.A.shape[1] + self.B.shape[0] * self.B.shape[1]
        return (k, k)

    def mean(self):
        """mean of all k*k entries of the covariance (the drivers print `abs(cov.mean())`)"""
        k = self.shape[0]
        return (self.A.sum() + self.B.sum()) / (float(k) * k)

    def toarray(self, max_dim=8192):
        """the k x k matrix in the reference's layout (active_pmf.py:136-142); small k only"""
        k = self.shape[0]
        if k > max_dim:
            raise MemoryError("a %d x %d covariance is what scalable mode avoids; use the blocks" % (k, k))
        out = np.zeros((k, k))
        d = self.A.shape[1]
        for b, blk in enumerate(list(self.A) + list(self.B)):
            out[b * d:(b + 1) * d, b * d:(b + 1) * d] = blk
        return out

    def __array__(self, dtype=None, copy=None):
        arr = self.toarray()
        return arr.astype(dtype) if dtype is not None else arr

    def stacked_mean(self):
        return np.hstack((self.mean_u.reshape(-1), self.mean_v.reshape(-1)))


def _f64(x):
    return D.to_device(np.ascontiguousarray(x, dtype=np.float64), np.float64)


class BlockPosterior(object):
    """Device-resident block posterior (all tables fp64)."""

    _SIDES = ("u", "v")

    def __init__(self, n, m, d, sigma_sq, sigma_u_sq, sigma_v_sq, mean_offset=0.0):
        if d > 32:
            raise ValueError("scalable mode supports latent_d <= 32")
        self.n, self.m, self.d = int(n), int(m), int(d)
        self.sigma_sq, self.sigma_u_sq, self.sigma_v_sq = float(sigma_sq), float(sigma_u_sq), float(sigma_v_sq)
        self.mean_offset = float(mean_offset)
        dev = D.device()
        f = dict(dtype=torch.float64, device=dev)
        for s, rows in (("u", self.n), ("v", self.m)):
            setattr(self, "mean_" + s, torch.zeros((rows, d), **f))
            setattr(self, "cov_" + s, torch.zeros((rows, d, d), **f))
            setattr(self, "prec_" + s, torch.zeros((rows, d, d), **f))
            setattr(self, "h_" + s, torch.zeros((rows, d), **f))
            setattr(self, "logdet_" + s, torch.zeros((rows,), **f))
        self._fail = torch.zeros(1, dtype=torch.int32, device=dev)
        self._sums = None
        self._packed = {}

    # ---------------------------------------------------------------- construction ------------
    @classmethod
    def from_host(cls, bd, sigma_sq, sigma_u_sq, sigma_v_sq, mean_offset=0.0):
        n, d = bd.mean_u.shape
        self = cls(n, bd.mean_v.shape[0], d, sigma_sq, sigma_u_sq, sigma_v_sq, mean_offset)
        self.mean_u, self.mean_v = _f64(bd.mean_u), _f64(bd.mean_v)
        self.cov_u, self.cov_v = _f64(bd.A), _f64(bd.B)
        self.prec_u, self.prec_v = _f64(bd.Lu), _f64(bd.Lv)
        self.h_u, self.h_v = _f64(bd.hu), _f64(bd.hv)
        self.logdet_u, self.logdet_v = _f64(bd.logdet_u), _f64(bd.logdet_v)
        return self

    def to_host(self):
        g = lambda t: t.cpu().numpy()  # noqa: E731
        return BlockDiagonal(g(self.mean_u), g(self.mean_v), g(self.cov_u), g(self.cov_v),
                             g(self.prec_u), g(self.prec_v), g(self.h_u), g(self.h_v),
                             g(self.logdet_u), g(self.logdet_v))

    def _invalidate(self):
        self._sums = None
        self._packed = {}

    def _check(self):
        if int(self._fail.item()):
            self._fail.zero_()
            raise np.linalg.LinAlgError("a block precision is not positive definite")

    def half_sweep(self, rat, side, cov_term=True, update_mean=True):
        """one coordinate half-sweep (amf_blocks_half_sweep): side 0 users, 1 items"""
        lib = N.require_device()
        me, other = ("u", "v") if side == 0 else ("v", "u")
        prior = self.sigma_u_sq if side == 0 else self.sigma_v_sq
        N.check(lib.amf_blocks_half_sweep(
            rat.handle, side, self.d, D.ptr(getattr(self, "mean_" + other)),
            D.ptr(getattr(self, "cov_" + other)) if cov_term else None, prior, self.sigma_sq,
            self.mean_offset, D.ptr(getattr(self, "prec_" + me)), D.ptr(getattr(self, "h_" + me)),
            D.ptr(getattr(self, "cov_" + me)),
            D.ptr(getattr(self, "mean_" + me)) if update_mean else None,
            D.ptr(getattr(self, "logdet_" + me)), D.ptr(self._fail), D.stream_ptr()))
        self._invalidate()

    def fit(self, rat, users, items, sweeps=500, tol=1e-10, cov_term=True, update_mean=True,
            kl_of=None):
        """fit_sweeps run to the end.  Without a per-sweep callback the whole loop is one library
        call (amf_blocks_fit): same sweeps, same stopping rule, no Python per sweep."""
        if kl_of is not None or not update_mean:
            for _ in self.fit_sweeps(rat, users, items, sweeps=sweeps, tol=tol, cov_term=cov_term,
                                     update_mean=update_mean, kl_of=kl_of):
                pass
            return self
        import ctypes as C
        lib = N.require_device()
        self.mean_u.copy_(_f64(users))
        self.mean_v.copy_(_f64(items))
        self.cov_u.zero_()
        self.cov_v.zero_()
        done = C.c_int(0)
        N.check(lib.amf_blocks_fit(
            rat.handle, self.d, self.sigma_u_sq, self.sigma_v_sq, self.sigma_sq, self.mean_offset,
            1 if cov_term else 0, int(sweeps), float(tol),
            D.ptr(self.mean_u), D.ptr(self.cov_u), D.ptr(self.prec_u), D.ptr(self.h_u), D.ptr(self.logdet_u),
            D.ptr(self.mean_v), D.ptr(self.cov_v), D.ptr(self.prec_v), D.ptr(self.h_v), D.ptr(self.logdet_v),
            D.ptr(self._fail), C.byref(done), D.stream_ptr()))
        self.sweeps_done = int(done.value)
        self._invalidate()
        self._check()
        return self

    def fit_sweeps(self, rat, users, items, sweeps=500, tol=1e-10, cov_term=True,
                   update_mean=True, kl_of=None):
        """Coordinate descent on the block-restricted KL from the factors (users, items): all
        user rows given the items' posterior, then all item columns; stops when no mean moves
        by more than `tol`.  A generator: yields kl_of(self) (or None) after each sweep."""
        self.mean_u.copy_(_f64(users))
        self.mean_v.copy_(_f64(items))
        self.cov_u.zero_()
        self.cov_v.zero_()
        for _ in range(int(sweeps)):
            old_u, old_v = self.mean_u.clone(), self.mean_v.clone()
            self.half_sweep(rat, 0, cov_term, update_mean)
            self.half_sweep(rat, 1, cov_term, update_mean)
            # ONE device->host read per sweep: the failure flag and the largest move of a mean
            state = torch.stack(((self._fail != 0).any().to(torch.float64),
                                 (self.mean_u - old_u).abs().max(),
                                 (self.mean_v - old_v).abs().max())).cpu().numpy()
            if state[0]:
                self._fail.zero_()
                raise np.linalg.LinAlgError("a block precision is not positive definite")
            yield kl_of(self) if kl_of is not None else None
            if update_mean and max(state[1], state[2]) < tol:
                break

    # ---------------------------------------------------------------- scalar summaries --------
    def entropy(self):
        """_approx_entropy (active_pmf.py:526-530): log det of the block-diagonal covariance"""
        return float(self.logdet_u.sum() + self.logdet_v.sum())

    def sums(self):
        """device tensor (4, d, d): sum A_i, sum m_i m_i^T, sum B_j, sum n_j n_j^T"""
        if self._sums is None:
            lib = N.require_device()
            out = torch.empty((4, self.d, self.d), dtype=torch.float64, device=self.mean_u.device)
            N.check(lib.amf_blocks_sums(self.n, self.d, D.ptr(self.mean_u), D.ptr(self.cov_u),
                                        D.ptr(out[0:2]), D.stream_ptr()))
            N.check(lib.amf_blocks_sums(self.m, self.d, D.ptr(self.mean_v), D.ptr(self.cov_v),
                                        D.ptr(out[2:4]), D.stream_ptr()))
            self._sums = out
        return self._sums

    def total_variance(self):
        """_total_variance (active_pmf.py:605-606) = <SA, SB + SNN> + <SMM, SB>"""
        s = self.sums()
        return float((s[0] * (s[2] + s[3])).sum() + (s[1] * s[2]).sum())

    def view(self, need_sums=False):
        return N.BlocksView(self.n, self.m, self.d,
                            self.mean_u.data_ptr(), self.cov_u.data_ptr(), self.prec_u.data_ptr(),
                            self.h_u.data_ptr(), self.logdet_u.data_ptr(),
                            self.mean_v.data_ptr(), self.cov_v.data_ptr(), self.prec_v.data_ptr(),
                            self.h_v.data_ptr(), self.logdet_v.data_ptr(),
                            self.sums().data_ptr() if need_sums else None,
                            self.sigma_sq, self.entropy())

    # ---------------------------------------------------------------- cell criteria -----------
    def packed(self, name):
        """the two SDDMM tables of pred_variance (amf_blocks_pack), cached per dtype"""
        if name not in self._packed:
            lib = N.require_device()
            d2 = self.d * (self.d + 1)
            ld = D.padded_ld(d2, name)
            out = []
            for side, (rows, mean, cov) in enumerate(((self.n, self.mean_u, self.cov_u),
                                                      (self.m, self.mean_v, self.cov_v))):
                t = torch.empty((rows, ld), dtype=D.torch_dtype(name), device=mean.device)
                N.check(lib.amf_blocks_pack(D.code(name), rows, self.d, D.ptr(mean), D.ptr(cov),
                                            side, ld, D.ptr(t), D.stream_ptr()))
                out.append(t)
            self._packed[name] = tuple(out)
        return self._packed[name]

    def _padded_means(self, name):
        key = "mean_" + name
        if key not in self._packed:
            ld = D.padded_ld(self.d, name)
            out = []
            for mean in (self.mean_u, self.mean_v):
                t = torch.zeros((mean.shape[0], ld), dtype=D.torch_dtype(name), device=mean.device)
                t[:, :self.d].copy_(mean)
                out.append(t)
            self._packed[key] = tuple(out)
        return self._packed[key]

    def normal_view(self, name):
        """amf_normal_view_t over the block tables (cov_uv = NULL); `name` tables"""
        key = "view_" + name
        if key not in self._packed:
            dt = D.torch_dtype(name)
            tabs = tuple(t.to(dt).contiguous() for t in (self.mean_u, self.mean_v, self.cov_u, self.cov_v))
            d = self.d
            self._packed[key] = (tabs, N.NormalView(tabs[0].data_ptr(), d, tabs[1].data_ptr(), d,
                                                    tabs[2].data_ptr(), d * d, d,
                                                    tabs[3].data_ptr(), d * d, d, None, 0, 0, 0))
        return self._packed[key][1]

    def score(self, criterion, ci, cj, name="f64", cutoff=0.0, want_scores=True, maximize=True,
              index_base=0):
        """approx mean / pred_variance / prob_ge for device candidate arrays; returns (scores
        tensor or None, best record tensor)."""
        from . import scoring as S
        d2 = self.d * (self.d + 1)
        if criterion == N.CRIT_APPROX_MEAN:
            mu, mv = self._padded_means(name)
            return S.score_device(N.CRIT_PRED, name, ci, cj, self.d, mu, mv,
                                  want_scores=want_scores, maximize=maximize, index_base=index_base)
        if criterion == N.CRIT_PRED_VARIANCE:
            pu, pv = self.packed(name)
            return S.score_device(N.CRIT_PRED, name, ci, cj, d2, pu, pv, want_scores=want_scores,
                                  maximize=maximize, index_base=index_base)
        if criterion == N.CRIT_PROB_GE:
            e, _ = self.score(N.CRIT_APPROX_MEAN, ci, cj, name)
            var, _ = self.score(N.CRIT_PRED_VARIANCE, ci, cj, name)
            out = torch.empty_like(e) if want_scores else None
            best = torch.empty(2, dtype=torch.int64, device=e.device)
            N.check(N.require_device().amf_prob_ge(D.code(name), int(e.numel()), D.ptr(e), D.ptr(var),
                                                   float(cutoff), D.ptr(out), 1 if maximize else 0,
                                                   int(index_base), D.ptr(best), D.stream_ptr()))
            return out, best
        raise ValueError("unknown criterion %r" % (criterion,))

    def kl(self, ri, rj, rr):
        """the reference's KL (active_pmf.py:202-240) at this posterior; ri, rj int32 and rr
        float64 device tensors of the rating list"""
        mean, _ = self.score(N.CRIT_APPROX_MEAN, ri, rj, "f64")
        var, _ = self.score(N.CRIT_PRED_VARIANCE, ri, rj, "f64")
        r = rr - self.mean_offset
        div = ((var + mean * mean) - 2 * r * mean + r * r).sum() / (2 * self.sigma_sq)
        tr = lambda c: torch.diagonal(c, dim1=1, dim2=2).sum()  # noqa: E731
        div = div + ((self.mean_u ** 2).sum() + tr(self.cov_u)) / (2 * self.sigma_u_sq)
        div = div + ((self.mean_v ** 2).sum() + tr(self.cov_v)) / (2 * self.sigma_v_sq)
        div = div - (self.logdet_u.sum() + self.logdet_v.sum()) / 2
        return float(div)

    # ---------------------------------------------------------------- lookahead ---------------
    def lookahead(self, what, ci, cj, values, weight_mode=N.WEIGHTS_NONE, bounds_or_weights=None,
                  rij_mean=None, rij_sd=None, rounds=1, want_evals=False, want_scores=True,
                  maximize=False, index_base=0):
        """amf_blocks_lookahead; device candidate arrays in, (evals (ncand, nv) or None, scores
        (ncand,) or None, best record) out."""
        lib = N.require_device()
        ncand = int(ci.numel())
        vals = _f64(values).reshape(-1)
        nv = int(vals.numel())
        dev = ci.device
        evals = torch.empty((ncand, nv), dtype=torch.float64, device=dev) if want_evals else None
        scores = torch.empty(ncand, dtype=torch.float64, device=dev) \
            if (want_scores and weight_mode != N.WEIGHTS_NONE) else None
        best = torch.empty(2, dtype=torch.int64, device=dev)
        wb = _f64(bounds_or_weights) if bounds_or_weights is not None else None
        if wb is not None:
            wb = torch.nan_to_num(wb, posinf=1e300, neginf=-1e300)     # the kernel ignores the two ends
        view = self.view(need_sums=(what == N.LOOK_TOTAL_VARIANCE))
        N.check(lib.amf_blocks_lookahead(C.byref(view), int(what), int(rounds), ncand, D.ptr(ci),
                                         D.ptr(cj), nv, D.ptr(vals), int(weight_mode), D.ptr(wb),
                                         D.ptr(rij_mean), D.ptr(rij_sd), D.ptr(evals), D.ptr(scores),
                                         1 if maximize else 0, int(index_base), D.ptr(best),
                                         D.ptr(self._fail), D.stream_ptr()))
        self._check()
        return evals, scores, best


# 2-sigma window of active_pmf.py:691-699 with fixed Gauss-Legendre nodes:
#   est = int_{-2}^{2} f(mu + sigma t) phi(t) dt
def gauss_nodes(nq=16):
    t, w = np.polynomial.legendre.leggauss(int(nq))
    t, w = 2 * t, 2 * w
    return t, w * np.exp(-t * t / 2) / np.sqrt(2 * np.pi)
