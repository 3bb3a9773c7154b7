"""Device-resident active-learning step at C5 scale: one gradient-ascent step on
the fused loss+gradient, one scoring pass over the resident pool with fused arg-max, removal of the
queried candidate from the pool (amf_pool_remove) and append of its rating to the list
(amf_ratings_append) -- against the reference's per-step pattern of rebuilding everything
(here: re-creating the rating handle and the pool, which is what a host-side add_rating costs)."""
import ctypes as C, os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from active_matrix_factorization_b200 import _native as N, device as D, scoring as S

a = types.SimpleNamespace(users=200_000, items=50_000, latent_d=32, nnz=50_000_000, ncand=100_000_000, dtype="f32")
torch.cuda.set_device(0)
p = bench.make_problem(a, 0, torch)
n, m, d = a.users, a.items, a.latent_d
lib = N.require_device()
U, V, ci, cj = p["U"].clone(), p["V"].clone(), p["ci"], p["cj"]
dU, dV = torch.empty_like(U), torch.empty_like(V)
sums = torch.zeros(3, dtype=torch.float64, device="cuda")
best = torch.zeros(2, dtype=torch.int64, device="cuda")
params = D.pmf_params(1.0, 10.0, 10.0, 0.0)

t0 = time.perf_counter()
rat = D.Ratings(n, m, p["ri"], p["rj"], p["r"], "f32")
torch.cuda.synchronize(); t_rat = time.perf_counter() - t0
t0 = time.perf_counter()
pool = S.Pool(ci, cj, n, m, "f32", d)
torch.cuda.synchronize(); t_pool = time.perf_counter() - t0


def step(lr=1e-6):
    N.check(lib.amf_pmf_loss_grad(rat.handle, N.F32, d, d, D.ptr(U), D.ptr(V), C.byref(params),
                                  D.ptr(dU), D.ptr(dV), D.ptr(sums), D.stream_ptr()))
    N.check(lib.amf_axpy(N.F32, U.numel(), D.ptr(U), D.ptr(dU), lr, D.ptr(U), D.stream_ptr()))
    N.check(lib.amf_axpy(N.F32, V.numel(), D.ptr(V), D.ptr(dV), lr, D.ptr(V), D.stream_ptr()))
    pool.score_pred(U, V, best=best)
    v, idx = S.unpack_best(best)                      # the one host sync of the step: the query
    i, j = int(ci[idx]), int(cj[idx])
    pool.remove([idx])
    rat.append(np.array([i], np.int32), np.array([j], np.int32), np.array([v], np.float32))
    return idx

step(); torch.cuda.synchronize()
K = 20
t0 = time.perf_counter()
picked = [step() for _ in range(K)]
torch.cuda.synchronize()
ms = 1e3 * (time.perf_counter() - t0) / K
assert len(set(picked)) == K, "a removed candidate was picked again"
print("resident step: %.2f ms (loss+grad, 2 axpy, scoring of %d candidates, remove, append); "
      "rebuilding instead would add %.0f ms (rating list: upload-free re-sort) + %.0f ms (pool) per step"
      % (ms, pool.ncand, 1e3 * t_rat, 1e3 * t_pool))
