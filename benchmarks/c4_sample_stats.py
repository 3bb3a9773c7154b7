#!/usr/bin/env python3
"""C4-shape timing of the sample-based criteria (bayes_pmf.py:433-455): kernel alone (device
resident inputs, CUDA events) against the whole BayesianPMF.pred_variance call (host sample list
in, numpy array out).  943 x 1682, rank 15, 200 samples, all cells."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from active_matrix_factorization_b200 import build  # noqa: E402

build.build()
from active_matrix_factorization_b200 import _native as N, device as D, bayes_pmf  # noqa: E402

n, m, d, S = 943, 1682, 15, 200
rng = np.random.RandomState(0)
samples = [(rng.normal(size=(n, d)), rng.normal(size=(m, d))) for _ in range(S)]
R = np.column_stack((rng.randint(0, n, 5000), rng.randint(0, m, 5000), rng.randint(1, 6, 5000))).astype(float)
R[0, :2] = (n - 1, m - 1)
lib = N.require_device()
for name in ("f64", "f32"):
    b = bayes_pmf.BayesianPMF(R, d)
    b.compute_dtype = name
    dt = D.np_dtype(name)
    us = D.to_device(np.stack([u for u, _ in samples]), dt)
    vs = D.to_device(np.stack([v for _, v in samples]), dt)
    ii, jj = np.meshgrid(np.arange(n), np.arange(m), indexing="ij")
    ci, cj = D.to_device(ii.ravel(), np.int32), D.to_device(jj.ravel(), np.int32)
    nc = ci.numel()
    var = torch.empty(nc, dtype=D.torch_dtype(name), device=ci.device)

    def launch(dense):
        N.check(lib.amf_bayes_sample_stats(D.code(name), nc, None if dense else D.ptr(ci),
                                           None if dense else D.ptr(cj), S, n, m, d,
                                           D.ptr(us), D.ptr(vs), 0.0, 0.0, None, D.ptr(var), None,
                                           1, 1, 0, None, D.stream_ptr()))

    def timed(dense):
        launch(dense)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            launch(dense)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5
    g_ms, k_ms = timed(False), timed(True)
    b.pred_variance(samples)
    t0 = time.perf_counter()
    out = b.pred_variance(samples)
    call_ms = (time.perf_counter() - t0) * 1e3
    ref = np.var([u[:7] @ v[:9].T for u, v in samples], 0)
    err = np.abs(out[:7, :9] - ref).max() / np.abs(ref).max()
    flop = nc * S * (2 * d + 6)
    print("%s: per-candidate kernel %.3f ms, dense kernel %.3f ms (%.2f TFLOP/s), whole call %.1f ms, "
          "rel err %.1e" % (name, g_ms, k_ms, flop / k_ms / 1e9, call_ms, err))
