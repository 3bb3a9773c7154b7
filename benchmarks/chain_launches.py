import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from itertools import islice
from benchmarks import config_lines as CL
from active_matrix_factorization_b200 import bayes_pmf as Bm
g = CL.golden("c4_movielens_bayes") if hasattr(CL, "golden") else None
R = g["ratings"] if g is not None and "ratings" in g else None
if R is None:
    rng = np.random.RandomState(0)
    n, m = 943, 1682
    cells = rng.permutation(n * m)[:5000]
    R = np.column_stack((cells // m, cells % m, rng.randint(1, 6, 5000))).astype(float)
b = Bm.BayesianPMF(R, 15, subtract_mean=True)
list(islice(b.samples_device(num_gibbs=2), 16))
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    out = list(islice(b.samples_device(num_gibbs=2), 192))
    e1.record(); torch.cuda.synchronize()
    print("192 samples: wall %.1f ms, device %.1f ms" % (1e3 * (time.perf_counter() - t0), e0.elapsed_time(e1)))
