"""Fast-mode Gibbs half-sweeps (in-kernel Philox, one Cholesky per row) at C5 scale, per side."""
import ctypes as C, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from active_matrix_factorization_b200 import _native as N, device as D
a = types.SimpleNamespace(users=200_000, items=50_000, latent_d=32, nnz=50_000_000, ncand=1_000_000, dtype="f32")
torch.cuda.set_device(0)
p = bench.make_problem(a, 0, torch)
n, m, d = a.users, a.items, a.latent_d
rat = D.Ratings(n, m, p["ri"], p["rj"], p["r"], "f32")
lib = N.require_device()
alpha = torch.eye(d, device="cuda") * 2.0
mu = torch.zeros(d, device="cuda")
for side, rows, other in ((0, n, p["V"]), (1, m, p["U"])):
    out = torch.empty((rows, d), device="cuda")
    oc = other.contiguous()
    def sweep(k=0):
        N.check(lib.amf_gibbs_half_sweep_device_rng(rat.handle, side, N.F32, d, D.ptr(oc), D.ptr(alpha), D.ptr(mu),
                                                    2.0, 0.0, 7, k, D.ptr(out), 0, -1, D.stream_ptr()))
    sweep(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(3): sweep(k + 1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("fast side %d: %d rows, %.2f ms per half-sweep, %.2e rows/s, finite=%s" % (
        side, rows, ms, rows / ms * 1e3, bool(torch.isfinite(out).all())))
