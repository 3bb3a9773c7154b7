"""Times the pieces of one exact-mode variational step (k = 40) on the device."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_matrix_factorization_b200 import normal as NM, _native as N
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "known_answer_10x10_d2.npz"))
p = NM.fit_params(10, 10, 2, 1., 10., 10.)
for B in (1, 148, 180):
    b = NM.NormalBatch(g["ratings"], p, np.broadcast_to(g["mean"], (B, 40)), np.broadcast_to(g["cov"], (B, 40, 40)))
    for mode, name in ((N.NORMAL_KL, "kl"), (N.NORMAL_GRADIENT, "gradient"), (N.NORMAL_PROJECT, "project")):
        b._run(mode); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            b._run(mode)
        torch.cuda.synchronize()
        print("B=%d %-9s %.1f us per launch" % (B, name, (time.perf_counter() - t0) / 50 * 1e6))
    # projection of a matrix that actually needs clamping (negative eigenvalues)
    rng = np.random.RandomState(0)
    s = rng.normal(0, 2, (40, 40))
    b2 = NM.NormalBatch(g["ratings"], p, np.broadcast_to(g["mean"], (B, 40)), np.broadcast_to(s, (B, 40, 40)))
    covs = b2.cov.clone()
    b2._run(N.NORMAL_PROJECT); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        b2.cov.copy_(covs); b2._run(N.NORMAL_PROJECT)
    torch.cuda.synchronize()
    print("B=%d project(random) %.1f us per launch" % (B, (time.perf_counter() - t0) / 20 * 1e6))
