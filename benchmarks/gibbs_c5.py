"""Gibbs half-sweeps (bayes_pmf.py:189-216 for every row) at C5 scale: 200k x 50k, rank 32,
~50M ratings, fp32 Gram accumulation + fp64 d x d algebra."""
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
    z = torch.randn((rows, d), device="cuda")
    out = torch.empty((rows, d), device="cuda")
    def sweep():
        N.check(lib.amf_gibbs_half_sweep(rat.handle, side, N.F32, d, D.ptr(other.contiguous()), D.ptr(alpha), D.ptr(mu),
                                         2.0, 0.0, D.ptr(z), D.ptr(out), D.stream_ptr()))
    sweep(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): sweep()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    flops = rat.nnz * (2 * d * d + 2 * d)
    print("side %d: %d rows, %.2f ms per half-sweep, %.2e rows/s, %.1f TFLOP/s counting only the Gram flops, finite=%s" % (
        side, rows, ms, rows / ms * 1e3, flops / ms / 1e9, bool(torch.isfinite(out).all())))
