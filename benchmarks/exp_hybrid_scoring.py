"""Experiment: score one pool with the L2-bound flat kernel and the shared-memory-bound bucketed
kernel running CONCURRENTLY on disjoint parts (two streams)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from active_matrix_factorization_b200 import _native as N, scoring as S

a = types.SimpleNamespace(users=200_000, items=50_000, latent_d=32, nnz=1_000_000, ncand=100_000_000, dtype="f32")
torch.cuda.set_device(0)
p = bench.make_problem(a, 0, torch)
U, V, ci, cj = p["U"], p["V"], p["ci"], p["cj"]
nc = ci.numel()
n, m, d = a.users, a.items, a.latent_d

def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

best = torch.zeros(2, dtype=torch.int64, device="cuda")
flat_all = timeit(lambda: S.score_device(N.CRIT_PRED, "f32", ci, cj, d, U, V, want_scores=False))
pool_all = S.Pool(ci, cj, n, m, "f32", d)
tiled_all = timeit(lambda: pool_all.score_pred(U, V, False, True, 0, best))
print("flat all %.3f ms, tiled all %.3f ms" % (flat_all, tiled_all))
pool_all.close()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for frac in (0.5, 0.6, 0.7, 0.8):
    cut = int(nc * frac) // 32 * 32
    pool = S.Pool(ci[:cut], cj[:cut], n, m, "f32", d)
    fi, fj = ci[cut:].contiguous(), cj[cut:].contiguous()
    b1 = torch.zeros(2, dtype=torch.int64, device="cuda"); 
    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            pool.score_pred(U, V, False, True, 0, b1)
        with torch.cuda.stream(s2):
            S.score_device(N.CRIT_PRED, "f32", fi, fj, d, U, V, want_scores=False, index_base=cut)
        cur.wait_stream(s1); cur.wait_stream(s2)
    t = timeit(both)
    t_pool = timeit(lambda: pool.score_pred(U, V, False, True, 0, b1))
    t_flat = timeit(lambda: S.score_device(N.CRIT_PRED, "f32", fi, fj, d, U, V, want_scores=False))
    print("tiled %.0f%% + flat %.0f%%: concurrent %.3f ms (alone: tiled %.3f, flat %.3f, sum %.3f)" % (
        100 * frac, 100 * (1 - frac), t, t_pool, t_flat, t_pool + t_flat))
    pool.close()
