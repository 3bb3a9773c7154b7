"""Batched candidate scoring on the device (active_pmf.py:739-770 `_get_key_vals`)."""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import device as D


def _cands(ii, jj):
    if isinstance(ii, torch.Tensor):
        return ii, jj
    return D.to_device(np.asarray(ii), np.int32), D.to_device(np.asarray(jj), np.int32)


def score_device(criterion, name, ci, cj, d, U=None, V=None, view=None, cutoff=0.0,
                 want_scores=True, maximize=True, index_base=0):
    """Low-level: device tensors in, (scores tensor or None, best tensor bytes) out.
    Returns (scores, (best_value, best_index)) after synchronising on the winner."""
    lib = N.require_device()
    ncand = int(ci.numel())
    scores = torch.empty(ncand, dtype=D.torch_dtype(name), device=ci.device) if want_scores else None
    best = torch.empty(2, dtype=torch.int64, device=ci.device)   # {double, int64} record
    ld = U.shape[1] if U is not None else 0
    N.check(lib.amf_score_candidates(criterion, D.code(name), ncand, D.ptr(ci), D.ptr(cj), d, ld,
                                     D.ptr(U), D.ptr(V), C.byref(view) if view is not None else None,
                                     float(cutoff), D.ptr(scores), 1 if maximize else 0,
                                     int(index_base), D.ptr(best), D.stream_ptr()))
    return scores, best


def unpack_best(best):
    raw = best.cpu().numpy()
    return float(raw[:1].view(np.float64)[0]), int(raw[1])


def score_pred(users, items, ii, jj, name, maximize=True):
    """U_i . V_j for host factor matrices; returns (float64 scores, (best value, best index))."""
    U, V = D.to_padded(users, name), D.to_padded(items, name)
    ci, cj = _cands(ii, jj)
    scores, best = score_device(N.CRIT_PRED, name, ci, cj, users.shape[1], U, V, maximize=maximize)
    return scores.to(torch.float64).cpu().numpy(), unpack_best(best)


def full_cov_view(mean_t, cov_t, n, m, d):
    """amf_normal_view_t over the reference's k-vector / k x k matrix (active_pmf.py:136-142)."""
    k = (n + m) * d
    es = mean_t.element_size()
    mp, cp = mean_t.data_ptr(), cov_t.data_ptr()
    return N.NormalView(
        mp, d, mp + n * d * es, d,
        cp, d * k + d, k,
        cp + (n * d * k + n * d) * es, d * k + d, k,
        cp + n * d * es, d * k, d, k)


def score_normal(criterion, mean, cov, n, m, d, ii, jj, name, cutoff=0.0, maximize=True):
    """Approximation-based criteria for a host (mean, cov) pair."""
    mean_t = D.to_device(mean, D.np_dtype(name))
    cov_t = D.to_device(cov, D.np_dtype(name))
    view = full_cov_view(mean_t, cov_t, n, m, d)
    ci, cj = _cands(ii, jj)
    scores, best = score_device(criterion, name, ci, cj, d, view=view, cutoff=cutoff, maximize=maximize)
    return scores.to(torch.float64).cpu().numpy(), unpack_best(best)
