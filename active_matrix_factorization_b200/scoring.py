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


def score_pred_host_csr(users, items, cand_ptr, cand_j, name, want_scores=False, maximize=True):
    """End-to-end host call for a pool sorted by user and given as row offsets (the candidates of
    user i are cand_j[cand_ptr[i]:cand_ptr[i+1]]): host arrays in, (scores or None, (best value,
    best position in cand_j)) out.  amf_score_pred_host_csr; half the PCIe bytes of (i, j) pairs.
    A ``cand_j`` of dtype uint16 (at most 65536 items) goes through amf_score_pred_host_csr16 and
    halves them again."""
    import ctypes as C
    lib = N.require_device()
    dt = D.np_dtype(name)
    users = np.ascontiguousarray(users, dtype=dt)
    items = np.ascontiguousarray(items, dtype=dt)
    cand_ptr = np.ascontiguousarray(cand_ptr, dtype=np.int64)
    narrow = getattr(cand_j, 'dtype', None) == np.uint16
    cand_j = np.ascontiguousarray(cand_j, dtype=np.uint16 if narrow else np.int32)
    n, d = users.shape
    if cand_ptr.shape[0] != n + 1 or cand_ptr[0] != 0 or cand_ptr[-1] != cand_j.shape[0]:
        raise ValueError("cand_ptr must hold n_users + 1 offsets running from 0 to len(cand_j)")
    scores = np.empty(cand_j.shape[0], dtype=dt) if want_scores else None
    best = N.Best()
    call = lib.amf_score_pred_host_csr16 if narrow else lib.amf_score_pred_host_csr
    N.check(call(D.code(name), N.host_ptr(cand_ptr), N.host_ptr(cand_j), n, items.shape[0], d,
                 N.host_ptr(users), N.host_ptr(items), N.host_ptr(scores) if want_scores else None,
                 1 if maximize else 0, C.byref(best)))
    return scores, (best.value, best.index)


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


class CandidatePool(object):
    """A candidate set that stays on the device across active-learning steps: what `unrated`
    (a Python set of tuples, pmf_cy.pyx:69-72) is to the reference's loop
    (active_pmf.py:880-898), without the per-step list(set) -> array -> H2D round trip.

    Pass it as `pool=` to `pick_query_point / _get_key_vals / get_key_evals`; `remove(i, j)`
    (or `add_rating` on a model the pool is attached to) drops a queried cell in O(1) by moving
    the last candidate into its slot, on the host mirror and on the device alike.  Iteration
    and indexing give (i, j) tuples in the current order, like a list."""

    def __init__(self, ii, jj=None):
        if jj is None:
            arr = np.asarray(ii)
            if arr.ndim != 2 or arr.shape[1] != 2:
                arr = np.array(sorted(ii), dtype=np.int64).reshape(-1, 2)
            ii, jj = arr[:, 0], arr[:, 1]
        self.i = np.ascontiguousarray(ii, dtype=np.int32).copy()
        self.j = np.ascontiguousarray(jj, dtype=np.int32).copy()
        self.n = int(self.i.shape[0])
        self.ci = D.to_device(self.i, np.int32)
        self.cj = D.to_device(self.j, np.int32)
        self._where = None       # (i, j) -> slot, built on the first remove()

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        if idx < 0:
            idx += self.n
        if not 0 <= idx < self.n:
            raise IndexError(idx)
        return (int(self.i[idx]), int(self.j[idx]))

    def __iter__(self):
        return iter(zip(self.i[:self.n].tolist(), self.j[:self.n].tolist()))

    def device_arrays(self):
        return self.ci[:self.n], self.cj[:self.n]

    def remove(self, i, j):
        """drop the cell (i, j) if it is a candidate"""
        if self._where is None:
            self._where = {ij: t for t, ij in enumerate(self)}
        slot = self._where.pop((int(i), int(j)), None)
        if slot is None:
            return False
        last = self.n - 1
        if slot != last:
            li, lj = int(self.i[last]), int(self.j[last])
            self.i[slot], self.j[slot] = li, lj
            self._where[(li, lj)] = slot
            self.ci[slot:slot + 1].copy_(self.ci[last:last + 1])
            self.cj[slot:slot + 1].copy_(self.cj[last:last + 1])
        self.n = last
        return True


class Pool:
    """Device-resident candidate pool bucketed by item tile (csrc/pool.cu); build it once and
    score it every active-learning step.  Scores and the winner are reported in the caller's
    order, exactly like the unbucketed path."""

    def __init__(self, ii, jj, n_users, n_items, name, d, tile_bytes=224 * 1024):
        import ctypes as C
        lib = N.require_device()
        self.name, self.d = name, int(d)
        self.n_users, self.n_items = int(n_users), int(n_items)
        ci, cj = _cands(ii, jj)
        assert ci.dtype == torch.int32 and cj.dtype == torch.int32
        self.ncand = int(ci.numel())
        # the pool kernel wants a row of 1, 2, 4, 8 or 16 16-byte vectors
        vec = D.vec_elems(name)
        nvec = (d + vec - 1) // vec
        nvec = 1 << (nvec - 1).bit_length()
        if nvec > 16:
            raise ValueError("latent_d=%d is too large for the bucketed pool" % d)
        self.ld = ld = nvec * vec
        row_bytes = ld * (4 if name == "f32" else 8)
        # as many item rows as the kernel can keep in shared memory next to its staging buffers
        # (any smaller tile height works: tile_bytes is a knob for tests and benchmarks)
        most = int(lib.amf_pool_max_tile_rows(row_bytes))
        self.tile_rows = int(max(1, min(most, tile_bytes // row_bytes)))
        self._h = C.c_void_p()
        torch.cuda.current_stream().synchronize()
        N.check(lib.amf_pool_create(C.byref(self._h), self.ncand, D.ptr(ci), D.ptr(cj),
                                    self.n_users, self.n_items, self.tile_rows, D.stream_ptr()))

    def score_pred(self, U, V, want_scores=False, maximize=True, index_base=0, best=None, peer=None):
        """U, V: device tensors padded to (rows, self.ld) -- see pad().  Returns (scores tensor
        or None, best record tensor).  peer = a connected parallel.PeerWinnerExchange: the winner
        over all ranks' shards, exchanged inside the scoring kernel (a collective call)."""
        lib = N.require_device()
        assert U.shape[1] == self.ld and V.shape[1] == self.ld, "use Pool.pad() for the factors"
        scores = torch.empty(self.ncand, dtype=D.torch_dtype(self.name), device=U.device) \
            if want_scores else None
        if best is None:
            best = torch.empty(2, dtype=torch.int64, device=U.device)
        if peer is not None:
            N.check(lib.amf_pool_score_pred_peer(self._h, D.code(self.name), self.d, U.shape[1], D.ptr(U),
                                                 D.ptr(V), D.ptr(scores), 1 if maximize else 0,
                                                 int(index_base), peer._h, D.ptr(best), D.stream_ptr()))
            return scores, best
        N.check(lib.amf_pool_score_pred(self._h, D.code(self.name), self.d, U.shape[1], D.ptr(U),
                                        D.ptr(V), D.ptr(scores), 1 if maximize else 0,
                                        int(index_base), D.ptr(best), D.stream_ptr()))
        return scores, best

    def remove(self, indices):
        """Drop candidates (positions in the caller's order) from the pool, O(1) each: what
        `unrated.difference_update` does for the reference's set (pmf_cy.pyx:152)."""
        idx = torch.as_tensor(np.atleast_1d(np.asarray(indices, dtype=np.int64))).to(D.device())
        N.check(N.require_device().amf_pool_remove(self._h, int(idx.numel()), D.ptr(idx),
                                                   D.stream_ptr()))

    def pad(self, arr):
        """host (rows, d) factors -> zero-padded (rows, self.ld) device tensor for this pool"""
        arr = np.ascontiguousarray(arr, dtype=D.np_dtype(self.name))
        out = torch.zeros((arr.shape[0], self.ld), dtype=D.torch_dtype(self.name), device=D.device())
        out[:, :arr.shape[1]].copy_(torch.from_numpy(arr))
        return out

    def close(self):
        if self._h:
            N.load().amf_pool_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
