"""The dense sample statistics on the tensor cores (csrc/dense_tc.cu: tcgen05.mma kind::tf32 with
hi/lo split operands, TMEM accumulators, 2-D TMA) against numpy in fp64, against the reference's
own Bayesian criteria fixture, and against the CUDA-core dense kernel it replaces."""
import ctypes as C

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def K():
    from active_matrix_factorization_b200 import build
    build.build()
    import torch
    from active_matrix_factorization_b200 import _native as N, device as D, scoring as S

    class NS:
        pass
    ns = NS()
    ns.N, ns.D, ns.S, ns.torch, ns.lib = N, D, S, torch, N.require_device()
    return ns


def _tc(K, Us, Vs, offset=0.0, select=1, maximize=True):
    S_, n, d = Us.shape
    m = Vs.shape[1]
    t = K.torch
    us, vs = K.D.to_device(Us, np.float32), K.D.to_device(Vs, np.float32)
    mean = t.empty(n * m, dtype=t.float32, device="cuda")
    var = t.empty(n * m, dtype=t.float32, device="cuda")
    best = t.zeros(2, dtype=t.int64, device="cuda")
    K.N.check(K.lib.amf_bayes_sample_stats_dense_tc(S_, n, m, d, K.D.ptr(us), K.D.ptr(vs), float(offset),
                                                    K.D.ptr(mean), K.D.ptr(var), select, 1 if maximize else 0,
                                                    0, K.D.ptr(best), K.D.stream_ptr()))
    t.cuda.synchronize()
    return mean.double().cpu().numpy().reshape(n, m), var.double().cpu().numpy().reshape(n, m), \
        K.S.unpack_best(best)


@pytest.mark.parametrize("n,m,d,S_", [(15, 12, 3, 5), (128, 128, 8, 2), (130, 257, 15, 7),
                                       (300, 140, 32, 9), (257, 129, 24, 4), (943, 1682, 15, 20)])
def test_dense_tc_matches_numpy(K, n, m, d, S_):
    rng = np.random.RandomState(n + d)
    Us = rng.normal(0, .6, (S_, n, d)).astype(np.float32)
    Vs = rng.normal(0, .6, (S_, m, d)).astype(np.float32)
    Us += rng.normal(0, 1, (1, n, d)).astype(np.float32)           # a common part: |mean| >> spread
    preds = np.einsum("snd,smd->snm", Us.astype(np.float64), Vs.astype(np.float64)) + 3.25
    mean, var, (bv, bi) = _tc(K, Us, Vs, offset=3.25)
    want_m, want_v = preds.mean(0), preds.var(0)
    scale = np.abs(preds).max()
    # 3xTF32: every product carries ~2^-21 relative error; the variance is accumulated about
    # the first sample, so its error is relative to the spread, not to the mean
    assert np.abs(mean - want_m).max() <= 4e-6 * scale
    assert np.abs(var - want_v).max() <= 2e-5 * want_v.max()
    assert bi == int(np.argmax(var)) and bv == pytest.approx(var.max(), rel=1e-12)
    _, _, (bv, bi) = _tc(K, Us, Vs, offset=3.25, select=0, maximize=False)
    assert bi == int(np.argmin(mean))


def test_dense_tc_against_reference_fixture(K, golden):
    """the reference's own predict / pred_variance over its seeded 3 samples (gibbs_15x12_d3)"""
    g = golden("gibbs_15x12_d3")
    us = np.stack([g["samples_u"][s] for s in range(g["samples_u"].shape[0])]).astype(np.float32)
    vs = np.stack([g["samples_v"][s] for s in range(g["samples_v"].shape[0])]).astype(np.float32)
    off = float(g["mean_rating"]) if "mean_rating" in g else 0.0
    mean, var, _ = _tc(K, us, vs, offset=off)
    ii, jj = g["cand_i"], g["cand_j"]
    np.testing.assert_allclose(var[ii, jj], g["bayes_pred_variance"], rtol=2e-4, atol=2e-5 * g["bayes_pred_variance"].max())
    np.testing.assert_allclose(mean[ii, jj], g["bayes_predict"], rtol=1e-5, atol=1e-5)


def test_dense_route_and_speed(K):
    """amf_bayes_sample_stats routes dense fp32 calls to the tensor-core kernel; both dense
    kernels agree; timing at the C4 shape (943 x 1682, rank 15, 200 samples)"""
    import os
    t = K.torch
    rng = np.random.RandomState(0)
    S_, n, m, d = 200, 943, 1682, 15
    us = K.D.to_device(rng.normal(0, .5, (S_, n, d)), np.float32)
    vs = K.D.to_device(rng.normal(0, .5, (S_, m, d)), np.float32)
    var = t.empty(n * m, dtype=t.float32, device="cuda")
    prob = t.empty(n * m, dtype=t.float32, device="cuda")
    best = t.zeros(2, dtype=t.int64, device="cuda")

    def run(with_prob):
        K.N.check(K.lib.amf_bayes_sample_stats(K.N.F32, n * m, None, None, S_, n, m, d, K.D.ptr(us), K.D.ptr(vs),
                                               0.0, 0.0, None, K.D.ptr(var), K.D.ptr(prob) if with_prob else None,
                                               1, 1, 0, K.D.ptr(best), K.D.stream_ptr()))

    def ms(fn, reps=10):
        fn(); fn()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        t.cuda.synchronize(); e0.record()
        for _ in range(reps):
            fn()
        e1.record(); t.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    run(False)
    v_tc, b_tc = var.clone(), K.S.unpack_best(best)
    run(True)                                   # a prob output keeps the call on the CUDA cores
    v_cc, b_cc = var.clone(), K.S.unpack_best(best)
    assert (v_tc - v_cc).abs().max().item() <= 2e-5 * v_cc.max().item()
    assert b_tc[1] == b_cc[1]
    t_tc, t_cc = ms(lambda: run(False)), ms(lambda: run(True))
    print("dense sample variance 943x1682 d=15 S=200: tensor cores %.3f ms (incl. split pre-pass), "
          "CUDA cores %.3f ms" % (t_tc, t_cc))
    assert t_tc < t_cc
