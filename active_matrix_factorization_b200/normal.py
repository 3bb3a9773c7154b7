"""Device driver for the batched variational approximation (csrc/normal.cu)."""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import device as D


def fit_params(n, m, d, sigma_sq, sigma_u_sq, sigma_v_sq, learning_rate=1e-4, min_eig=1e-5,
               kl_stop=.005, min_lr=1e-10, max_steps=0):
    return N.NormalFitParams(int(n), int(m), int(d), float(sigma_sq), float(sigma_u_sq),
                             float(sigma_v_sq), float(learning_rate), float(min_eig),
                             float(kl_stop), float(min_lr), int(max_steps))


class NormalBatch:
    """B problems sharing one COO rating list; each may append one extra rating."""

    def __init__(self, ratings, params, means, covs, extra=None):
        """ratings (nnz,3) host array; means (B,k), covs (B,k,k) host float64; extra = optional
        (ei, ej, er) host arrays of length B."""
        self.lib = N.require_device()
        self.p = params
        self.k = (params.n + params.m) * params.d
        ratings = np.asarray(ratings, dtype=np.float64).reshape(-1, 3)
        self.nnz = ratings.shape[0]
        self.ri = D.to_device(ratings[:, 0], np.int32)
        self.rj = D.to_device(ratings[:, 1], np.int32)
        self.rr = D.to_device(ratings[:, 2], np.float64)
        means = np.ascontiguousarray(means, dtype=np.float64).reshape(-1, self.k)
        covs = np.ascontiguousarray(covs, dtype=np.float64).reshape(-1, self.k, self.k)
        self.B = means.shape[0]
        assert covs.shape[0] == self.B
        self.mean = D.to_device(means, np.float64)
        self.cov = D.to_device(covs, np.float64)
        if extra is not None:
            self.ei = D.to_device(extra[0], np.int32)
            self.ej = D.to_device(extra[1], np.int32)
            self.er = D.to_device(extra[2], np.float64)
        else:
            self.ei = self.ej = self.er = None
        self.ws = int(self.lib.amf_normal_workspace_doubles(params.n, params.m, params.d))
        dev = self.mean.device
        self.work = torch.empty(self.B * self.ws, dtype=torch.float64, device=dev)
        self.kl = torch.zeros(self.B, dtype=torch.float64, device=dev)
        self.steps = torch.zeros(self.B, dtype=torch.int32, device=dev)

    def _run(self, mode, trace=None, entropy=None, totvar=None):
        N.check(self.lib.amf_normal_batched(
            mode, self.B, self.nnz, D.ptr(self.ri), D.ptr(self.rj), D.ptr(self.rr),
            D.ptr(self.ei), D.ptr(self.ej), D.ptr(self.er), C.byref(self.p), D.ptr(self.mean),
            D.ptr(self.cov), D.ptr(self.work), D.ptr(self.kl), D.ptr(self.steps),
            D.ptr(trace), 0 if trace is None else trace.shape[1], D.ptr(entropy), D.ptr(totvar),
            D.stream_ptr()))

    def kl_divergence(self):
        self._run(N.NORMAL_KL)
        return self.kl.cpu().numpy()

    def gradient(self):
        self._run(N.NORMAL_GRADIENT)
        w = self.work.view(self.B, self.ws)
        k = self.k
        return (w[:, :k].cpu().numpy().copy(),
                w[:, 2 * k:2 * k + k * k].reshape(self.B, k, k).cpu().numpy().copy())

    def project(self):
        self._run(N.NORMAL_PROJECT)
        return self.cov.cpu().numpy()

    def fit(self, trace_len=0, want_entropy=False, want_totvar=False):
        """Runs fit_normal_kls on every problem.  Returns dict of host arrays."""
        dev = self.mean.device
        trace = torch.full((self.B, trace_len), float('nan'), dtype=torch.float64, device=dev) \
            if trace_len else None
        ent = torch.zeros(self.B, dtype=torch.float64, device=dev) if want_entropy else None
        tv = torch.zeros(self.B, dtype=torch.float64, device=dev) if want_totvar else None
        self._run(N.NORMAL_FIT, trace, ent, tv)
        out = dict(kl=self.kl.cpu().numpy(), steps=self.steps.cpu().numpy())
        if trace is not None:
            out['trace'] = trace.cpu().numpy()
        if ent is not None:
            out['entropy'] = ent.cpu().numpy()
        if tv is not None:
            out['total_variance'] = tv.cpu().numpy()
        return out

    def means(self):
        return self.mean.cpu().numpy()

    def covs(self):
        return self.cov.cpu().numpy()


def project_psd_device(mat, min_eig):
    mat = np.ascontiguousarray(mat, dtype=np.float64)
    k = mat.shape[0]
    p = fit_params(k, 0, 1, 1., 1., 1., min_eig=min_eig)
    batch = NormalBatch(np.zeros((0, 3)), p, np.zeros((1, k)), mat[None])
    return batch.project()[0]


class MnBatch:
    """B matrix-normal problems MN(mean, Sigma, Omega) sharing one COO rating list (csrc/mn.cu)."""

    def __init__(self, ratings, params, means, sigs, oms, extra=None):
        self.lib = N.require_device()
        self.p = params
        self.nui = params.n + params.m
        self.d = params.d
        ratings = np.asarray(ratings, dtype=np.float64).reshape(-1, 3)
        self.nnz = ratings.shape[0]
        self.ri = D.to_device(ratings[:, 0], np.int32)
        self.rj = D.to_device(ratings[:, 1], np.int32)
        self.rr = D.to_device(ratings[:, 2], np.float64)
        means = np.ascontiguousarray(means, dtype=np.float64).reshape(-1, self.nui, self.d)
        self.B = means.shape[0]
        self.mean = D.to_device(means, np.float64)
        self.sig = D.to_device(np.ascontiguousarray(sigs, dtype=np.float64).reshape(self.B, self.nui, self.nui), np.float64)
        self.om = D.to_device(np.ascontiguousarray(oms, dtype=np.float64).reshape(self.B, self.d, self.d), np.float64)
        if extra is not None:
            self.ei = D.to_device(extra[0], np.int32)
            self.ej = D.to_device(extra[1], np.int32)
            self.er = D.to_device(extra[2], np.float64)
        else:
            self.ei = self.ej = self.er = None
        self.ws = int(self.lib.amf_mn_workspace_doubles(params.n, params.m, params.d))
        dev = self.mean.device
        self.work = torch.empty(self.B * self.ws, dtype=torch.float64, device=dev)
        self.kl = torch.zeros(self.B, dtype=torch.float64, device=dev)
        self.steps = torch.zeros(self.B, dtype=torch.int32, device=dev)

    def _run(self, mode, trace=None, entropy=None, totvar=None):
        N.check(self.lib.amf_mn_batched(
            mode, self.B, self.nnz, D.ptr(self.ri), D.ptr(self.rj), D.ptr(self.rr),
            D.ptr(self.ei), D.ptr(self.ej), D.ptr(self.er), C.byref(self.p), D.ptr(self.mean),
            D.ptr(self.sig), D.ptr(self.om), D.ptr(self.work), D.ptr(self.kl), D.ptr(self.steps),
            D.ptr(trace), 0 if trace is None else trace.shape[1], D.ptr(entropy), D.ptr(totvar),
            D.stream_ptr()))

    def kl_divergence(self):
        self._run(N.NORMAL_KL)
        return self.kl.cpu().numpy()

    def gradient(self):
        self._run(N.NORMAL_GRADIENT)
        w = self.work.view(self.B, self.ws)
        nd, n2, d2 = self.nui * self.d, self.nui * self.nui, self.d * self.d
        gm = w[:, :nd].reshape(self.B, self.nui, self.d).cpu().numpy().copy()
        gs = w[:, 2 * nd:2 * nd + n2].reshape(self.B, self.nui, self.nui).cpu().numpy().copy()
        o0 = 2 * nd + 5 * n2
        go = w[:, o0:o0 + d2].reshape(self.B, self.d, self.d).cpu().numpy().copy()
        return gm, gs, go

    def fit(self, trace_len=0, want_entropy=False, want_totvar=False):
        dev = self.mean.device
        trace = torch.full((self.B, trace_len), float('nan'), dtype=torch.float64, device=dev) \
            if trace_len else None
        ent = torch.zeros(self.B, dtype=torch.float64, device=dev) if want_entropy else None
        tv = torch.zeros(self.B, dtype=torch.float64, device=dev) if want_totvar else None
        self._run(N.NORMAL_FIT, trace, ent, tv)
        out = dict(kl=self.kl.cpu().numpy(), steps=self.steps.cpu().numpy())
        if trace is not None:
            out['trace'] = trace.cpu().numpy()
        if ent is not None:
            out['entropy'] = ent.cpu().numpy()
        if tv is not None:
            out['total_variance'] = tv.cpu().numpy()
        return out

    def state(self):
        return self.mean.cpu().numpy(), self.sig.cpu().numpy(), self.om.cpu().numpy()


def mn_score(criterion, mean, sig, om, n, m, d, ii, jj, name="f64", cutoff=0.0, maximize=True):
    """criteria over candidates under MN(mean, Sigma, Omega); returns (scores, (best, index))"""
    from . import scoring as S
    lib = N.require_device()
    dt = D.np_dtype(name)
    mean_t, sig_t, om_t = D.to_device(mean, dt), D.to_device(sig, dt), D.to_device(om, dt)
    ci, cj = S._cands(ii, jj)
    nc = int(ci.numel())
    scores = torch.empty(nc, dtype=D.torch_dtype(name), device=ci.device)
    best = torch.empty(2, dtype=torch.int64, device=ci.device)
    N.check(lib.amf_mn_score_candidates(criterion, D.code(name), nc, D.ptr(ci), D.ptr(cj), n, m, d,
                                        D.ptr(mean_t), D.ptr(sig_t), D.ptr(om_t), float(cutoff),
                                        D.ptr(scores), 1 if maximize else 0, 0, D.ptr(best),
                                        D.stream_ptr()))
    return scores.to(torch.float64).cpu().numpy(), S.unpack_best(best)


class MnWide:
    """One LARGE matrix-normal problem (N+M in the hundreds or thousands): the line search of
    mn_active_pmf.py:242-288 driven from the host, with the rating/prior terms from csrc/mn.cu
    (sparse modes) and the (N+M)^2 dense algebra -- eigendecomposition for project_psd, Cholesky
    for log-det and inverse -- done grid-wide by cuSOLVER through torch.  The single-CTA kernel
    is the right tool for thousands of small lookahead problems, not for one big one."""

    MN_KL_SPARSE, MN_GRADIENT_SPARSE = 4, 5

    def __init__(self, ratings, params, mean, sig, om):
        self.b = MnBatch(ratings, params, mean[None], sig[None], om[None])
        self.p = params
        self.nui, self.d = self.b.nui, self.b.d
        self._clamped = {}

    @staticmethod
    def _logdet(mat):
        chol, info = torch.linalg.cholesky_ex(mat)
        if int(info.item()) != 0:
            return None, None
        return 2.0 * torch.log(torch.diagonal(chol)).sum(), chol

    def kl(self, mean, sig, om):
        b = self.b
        b.mean, b.sig, b.om = mean.reshape(1, self.nui, self.d), sig.reshape(1, self.nui, self.nui), om.reshape(1, self.d, self.d)
        b._run(self.MN_KL_SPARSE)
        ls, _ = self._logdet(sig)
        lo, _ = self._logdet(om)
        if ls is None or lo is None:
            return float('nan')
        return float((b.kl[0] - (ls * self.d + lo * self.nui) / 2.0).item())

    def gradient(self, mean, sig, om):
        b = self.b
        b.mean, b.sig, b.om = mean.reshape(1, self.nui, self.d), sig.reshape(1, self.nui, self.nui), om.reshape(1, self.d, self.d)
        b._run(self.MN_GRADIENT_SPARSE)
        w = b.work.view(1, b.ws)
        nd, n2, d2 = self.nui * self.d, self.nui * self.nui, self.d * self.d
        gm = w[0, :nd].reshape(self.nui, self.d).clone()
        gs = w[0, 2 * nd:2 * nd + n2].reshape(self.nui, self.nui).clone()
        o0 = 2 * nd + 5 * n2
        go = w[0, o0:o0 + d2].reshape(self.d, self.d).clone()
        for g, mat, scale in ((gs, sig, self.d / 2.0), (go, om, self.nui / 2.0)):
            inv = torch.linalg.inv(mat)
            eye = torch.eye(mat.shape[0], dtype=mat.dtype, device=mat.device)
            g -= scale * (inv + inv.T * (1 - eye))
        return gm, gs, go

    def project(self, mat, min_eig, key):
        """project_psd of mn_active_pmf.py:42-67.  When the previous projection of this matrix
        did not need clamping, first ask a Cholesky of (mat - min_eig*I) whether lambda_min >=
        min_eig -- then the reference returns the symmetrised matrix unchanged and the
        eigendecomposition is skipped."""
        mat = (mat + mat.T) / 2
        if not self._clamped.get(key, True):
            eye = torch.eye(mat.shape[0], dtype=mat.dtype, device=mat.device)
            _, info = torch.linalg.cholesky_ex(mat - min_eig * eye)
            if int(info.item()) == 0:
                return mat
        w, q = torch.linalg.eigh(mat)
        clamp = float(w.min().item()) < min_eig
        self._clamped[key] = clamp
        if clamp:
            mat = (q * torch.clamp(w, min=min_eig)) @ q.T
            mat = (mat + mat.T) / 2
        return mat

    def fit(self, max_steps=0, callback=None):
        """Returns (mean, sig, om, [kl per accepted step]) as host arrays."""
        p = self.p
        mean, sig, om = self.b.mean[0].clone(), self.b.sig[0].clone(), self.b.om[0].clone()
        lr = p.learning_rate
        old = self.kl(mean, sig, om)
        kls, done = [], False
        while not done:
            gm, gs, go = self.gradient(mean, sig, om)
            while True:
                nm = mean - lr * gm
                ns = self.project(sig - lr * gs, p.min_eig, 'sig')
                no = self.project(om - lr * go, p.min_eig, 'om')
                new = self.kl(nm, ns, no)
                if new < old:
                    mean, sig, om = nm, ns, no
                    lr *= 1.25
                    if old - new < p.kl_stop:
                        done = True
                    kls.append(new)
                    if callback is not None:
                        callback(new)
                    old = new
                    break
                lr *= .5
                if lr < p.min_lr:
                    done = True
                    break
            if max_steps and len(kls) >= max_steps:
                break
        return mean.cpu().numpy(), sig.cpu().numpy(), om.cpu().numpy(), kls
