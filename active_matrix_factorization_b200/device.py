"""Device-memory plumbing for the host mirror: torch tensors own HBM, libamf_b200 does the math."""
import ctypes as C
import os

import numpy as np
import torch

from . import _native as N

_DEFAULT_DTYPE = os.environ.get("AMF_B200_DTYPE", "f64")


def default_dtype():
    """'f64' (parity mode: the reference's precision) or 'f32' (fast mode)."""
    return _DEFAULT_DTYPE


def set_default_dtype(name):
    global _DEFAULT_DTYPE
    if name not in ("f32", "f64"):
        raise ValueError("dtype must be 'f32' or 'f64'")
    _DEFAULT_DTYPE = name


def np_dtype(name):
    return np.float32 if name == "f32" else np.float64


def torch_dtype(name):
    return torch.float32 if name == "f32" else torch.float64


def code(name):
    return N.F32 if name == "f32" else N.F64


def vec_elems(name):
    return 4 if name == "f32" else 2


def padded_ld(d, name):
    v = vec_elems(name)
    return (d + v - 1) // v * v


def device():
    N.require_device()
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def to_padded(arr, name):
    """host (rows, d) array -> zero-padded (rows, ld) device tensor of the compute dtype"""
    arr = np.ascontiguousarray(arr, dtype=np_dtype(name))
    rows, d = arr.shape
    ld = padded_ld(d, name)
    out = torch.zeros((rows, ld), dtype=torch_dtype(name), device=device())
    out[:, :d].copy_(torch.from_numpy(arr))
    return out


def from_padded(t, d):
    """(rows, ld) device tensor -> host (rows, d) float64 array"""
    return t[:, :d].to(torch.float64).cpu().numpy().copy()


def to_device(arr, dtype):
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(device())


class Ratings:
    """Owner of an amf_ratings_t handle (device CSR + CSC of the rating list)."""

    def __init__(self, n_users, n_items, i, j, r, name):
        lib = N.require_device()
        self.name = name
        self.n_users, self.n_items = int(n_users), int(n_items)
        self._h = C.c_void_p()
        if isinstance(i, torch.Tensor):
            assert i.dtype == torch.int32 and j.dtype == torch.int32 and r.dtype == torch_dtype(name)
            i, j, r = i.contiguous(), j.contiguous(), r.contiguous()
            self.nnz = int(i.numel())
            torch.cuda.current_stream().synchronize()
            N.check(lib.amf_ratings_create(C.byref(self._h), self.n_users, self.n_items, self.nnz,
                                           ptr(i), ptr(j), ptr(r), code(name), stream_ptr()))
        else:
            i = np.ascontiguousarray(i, dtype=np.int32)
            j = np.ascontiguousarray(j, dtype=np.int32)
            r = np.ascontiguousarray(r, dtype=np_dtype(name))
            self.nnz = int(i.shape[0])
            N.check(lib.amf_ratings_create_host(C.byref(self._h), self.n_users, self.n_items,
                                                self.nnz, N.host_ptr(i), N.host_ptr(j),
                                                N.host_ptr(r), code(name)))
        layout = os.environ.get("AMF_B200_LAYOUT", "auto")
        if layout != "auto":
            self.set_layout(layout)

    def append(self, i, j, r):
        """Add ratings to the device-resident list without re-sorting it (amf_ratings_append);
        i, j, r: numpy arrays or CUDA tensors."""
        if isinstance(i, torch.Tensor):
            ti, tj, tr = i.to(torch.int32).contiguous(), j.to(torch.int32).contiguous(), \
                r.to(torch_dtype(self.name)).contiguous()
        else:
            ti, tj = to_device(np.atleast_1d(i), np.int32), to_device(np.atleast_1d(j), np.int32)
            tr = to_device(np.atleast_1d(r), np_dtype(self.name))
        torch.cuda.current_stream().synchronize()
        N.check(N.load().amf_ratings_append(self.handle, int(ti.numel()), ptr(ti), ptr(tj), ptr(tr),
                                            stream_ptr()))
        torch.cuda.current_stream().synchronize()      # the inputs may be freed by the caller
        self.nnz += int(ti.numel())

    def compact(self):
        """Fold appended ratings into the sorted lists now (amf_ratings_compact)."""
        N.check(N.load().amf_ratings_compact(self.handle, stream_ptr()))

    def set_layout(self, mode):
        """'auto' | 'rows' | 'tiled': which copy of the list the fused loss+gradient runs on
        (amf_ratings_set_layout)."""
        N.check(N.load().amf_ratings_set_layout(self.handle, {"auto": 0, "rows": 1, "tiled": 2}[mode]))

    @classmethod
    def from_tuples(cls, ratings, n_users, n_items, name):
        ratings = np.asarray(ratings, dtype=np.float64)
        return cls(n_users, n_items, ratings[:, 0].astype(np.int32), ratings[:, 1].astype(np.int32),
                   ratings[:, 2], name)

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("rating handle already destroyed")
        return self._h

    def mean(self):
        out = C.c_double()
        N.check(N.load().amf_ratings_mean(self.handle, C.byref(out), stream_ptr()))
        return out.value

    def close(self):
        if self._h:
            N.load().amf_ratings_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pmf_params(sigma_sq, sigma_u_sq, sigma_v_sq, mean_offset):
    return N.PmfParams(float(sigma_sq), float(sigma_u_sq), float(sigma_v_sq), float(mean_offset))


def loss_grad(rat, d, U, V, params, dU=None, dV=None, sums=None):
    """Enqueues the fused loss+gradient on device tensors; returns the (3,) float64 sums tensor."""
    lib = N.require_device()
    if sums is None:
        sums = torch.empty(3, dtype=torch.float64, device=U.device)
    N.check(lib.amf_pmf_loss_grad(rat.handle, code(rat.name), d, U.shape[1], ptr(U), ptr(V),
                                  C.byref(params), ptr(dU), ptr(dV), ptr(sums), stream_ptr()))
    return sums


def loss_grad_part(rat, d, U, V, params, dU, dV, sums, part, max_ctas=0):
    """One half of the fused loss+gradient (amf_pmf_loss_grad_part): part 0 completes dU and the
    sums, part 1 completes dV."""
    N.check(N.require_device().amf_pmf_loss_grad_part(
        rat.handle, code(rat.name), d, U.shape[1], ptr(U), ptr(V), C.byref(params), ptr(dU), ptr(dV),
        ptr(sums), int(part), int(max_ctas), stream_ptr()))


def axpy(X, G, lr, out, name):
    N.check(N.require_device().amf_axpy(code(name), X.numel(), ptr(X), ptr(G), float(lr), ptr(out),
                                        stream_ptr()))


def log_likelihood_from_sums(s, sigma_sq, sigma_u_sq, sigma_v_sq):
    return (-s[0] / (2. * sigma_sq) - s[1] / (2. * sigma_u_sq) - s[2] / (2. * sigma_v_sq))
