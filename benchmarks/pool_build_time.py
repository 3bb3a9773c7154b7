import os, sys, time, types
sys.path.insert(0, "/root/repo")
import torch, bench
from active_matrix_factorization_b200 import scoring as S
a = types.SimpleNamespace(users=200_000, items=50_000, latent_d=32, nnz=1_000_000, ncand=100_000_000, dtype="f32")
torch.cuda.set_device(0)
p = bench.make_problem(a, 0, torch)
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pool = S.Pool(p["ci"], p["cj"], a.users, a.items, "f32", 32)
    torch.cuda.synchronize(); print("pool build %d: %.1f ms" % (k, 1e3 * (time.perf_counter() - t0)))
    pool.close()
