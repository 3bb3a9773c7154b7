#!/usr/bin/env python3
"""C4-shape dense sample variance (943 x 1682, rank 15, 200 samples, fp32) on the tensor-core
kernel and on the CUDA-core kernel it replaces: the command the ncu captures under profiles/ run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from active_matrix_factorization_b200 import build
    build.build()
    from active_matrix_factorization_b200 import _native as N, device as D
    lib = N.require_device()
    rng = np.random.RandomState(0)
    S_, n, m, d = 200, 943, 1682, int(os.environ.get("D", "15"))
    us = D.to_device(rng.normal(0, .5, (S_, n, d)), np.float32)
    vs = D.to_device(rng.normal(0, .5, (S_, m, d)), np.float32)
    var = torch.empty(n * m, dtype=torch.float32, device="cuda")
    prob = torch.empty(n * m, dtype=torch.float32, device="cuda")
    best = torch.zeros(2, dtype=torch.int64, device="cuda")

    def run(with_prob):
        N.check(lib.amf_bayes_sample_stats(N.F32, n * m, None, None, S_, n, m, d, D.ptr(us), D.ptr(vs), 0.0, 0.0,
                                           None, D.ptr(var), D.ptr(prob) if with_prob else None, 1, 1, 0,
                                           D.ptr(best), D.stream_ptr()))
    for with_prob in (False, True):
        for _ in range(3):
            run(with_prob)
        import time
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        t0 = time.perf_counter()
        for _ in range(10):
            run(with_prob)
        host = (time.perf_counter() - t0) / 10
        e1.record(); torch.cuda.synchronize()
        print("%s: %.3f ms per call on the device, %.3f ms of host time to enqueue it" % (
            "CUDA cores (prob output requested)" if with_prob else "tensor cores", e0.elapsed_time(e1) / 10, host * 1e3))


if __name__ == "__main__":
    main()
