"""BASELINE.json configurations 1-4 at their real shapes plus the SURVEY.md 8d kernels that the
C5 step does not exercise (S2 block-posterior pred_variance, S3 sample variance, B1 Gibbs
half-sweep, L1 lookahead), each timed on the device with CUDA events.  bench.py puts the result
into its JSON line as `configs`; `python benchmarks/config_lines.py` prints it alone.

Inputs: the reference's own splits of its own data files (drugbank 94x425, movielens-100k),
committed as fixtures by tests/golden/make_golden_configs.py -- tests/golden/*.npz are data,
nothing under oracle/ is imported here.
"""
import json
import os
import sys
import time
from itertools import islice

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def cuda_ms(fn, reps=10, warm=2):
    """average device time of fn() on the current stream (CUDA events, synchronised)"""
    import torch
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def wall_ms(fn, reps=3, warm=1):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


def unknown_cells(R, n, m):
    known = np.zeros((n, m), bool)
    known[R[:, 0].astype(int), R[:, 1].astype(int)] = True
    return np.nonzero(~known)


def roof(bytes_, flops, ms, hbm_peak):
    out = {"kernel_ms": ms, "algorithmic_bytes": int(bytes_), "algorithmic_flops": int(flops),
           "achieved_gbs": bytes_ / (ms * 1e-3) / 1e9, "achieved_tflops": flops / (ms * 1e-3) / 1e12}
    out["frac_hbm"] = out["achieved_gbs"] / hbm_peak
    return out


def c1_exact_lookahead():
    """C1: 10x10, two rating levels, rank 2, diagonal known -- exact (full-covariance) mode: MAP
    fit, variational fit, pred-variance pick, and the uv-entropy lookahead over the whole pool as
    one launch of one CTA per (candidate, value) variational re-fit"""
    from active_matrix_factorization_b200 import active_pmf as A
    rng = np.random.RandomState(0)
    u, v = rng.normal(0, 2, (10, 2)), rng.normal(0, 2, (10, 2))
    real = ((u @ v.T + rng.normal(0, .25, (10, 10))) > 0).astype(float)
    ratings = np.array([(i, i, real[i, i]) for i in range(10)])
    np.random.seed(0)
    a = A.ActivePMF(ratings, 2, rating_values={0, 1}, discrete_expectations=True)
    a.approx_mode = 'exact'
    t0 = time.perf_counter()
    a.fit()
    a.initialize_approx()
    steps = len(list(a.fit_normal_kls()))
    fit_s = time.perf_counter() - t0
    pool = sorted(a.unrated)
    ms = wall_ms(lambda: a._get_key_vals(pool, A.ActivePMF.exp_approx_entropy), reps=2)
    pv_ms = wall_ms(lambda: a.pick_query_point(pool, A.ActivePMF.pred_variance))
    return {"config": "C1 10x10, 2 levels, rank 2, exact mode (k = 40)", "candidates": len(pool),
            "map_plus_variational_fit_s": fit_s, "fit_normal_steps": steps,
            "uv_entropy_ms": ms, "uv_entropy_cand_per_s": len(pool) / (ms * 1e-3),
            "values": 2, "refits_per_s": 2 * len(pool) / (ms * 1e-3), "pick_pred_variance_ms": pv_ms}


def c2_drugbank(hbm_peak):
    """C2: drugbank 94x425, 500 known, rank 5, uv-entropy over all 39,450 unknown cells
    (scalable mode; exact mode would need a 2595 x 2595 re-fit per candidate and value)"""
    import torch
    from active_matrix_factorization_b200 import _native as N, active_pmf as A, device as D, scoring as S
    g = golden("c2_drugbank")
    R = g["ratings"].astype(float)
    n, m, d = 94, 425, 5
    a = A.ActivePMF(R, d, rating_values={-1, 1}, discrete_expectations=True)
    a.users, a.items = g["users"].copy(), g["items"].copy()
    a.blocks_tol, a.blocks_max_sweeps = 1e-12, 2000
    t0 = time.perf_counter()
    a.initialize_approx()
    sweeps = len(list(a.fit_normal_kls()))
    fit_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    a.initialize_approx()
    a.fit_normal()                      # one library call (amf_blocks_fit), no KL per sweep
    fit_fast_s = time.perf_counter() - t0
    ii, jj = unknown_cells(R, n, m)
    pool = np.column_stack((ii, jj)).astype(np.int32)
    post = a._block_posterior()
    ci, cj = D.to_device(ii, np.int32), D.to_device(jj, np.int32)
    U, V = D.to_padded(a.users, "f64"), D.to_padded(a.items, "f64")
    mu, _ = S.score_device(N.CRIT_PRED, "f64", ci, cj, d, U, V)
    sd = torch.ones_like(mu)
    vals, bounds = np.array([-1., 1.]), a.rating_bounds
    out = {"config": "C2 drugbank 94x425, 500 known, rank 5, scalable mode (k = 2595 never formed)",
           "candidates": len(ii), "values": 2,
           "block_fit": {"sweeps": sweeps, "seconds": fit_s, "fit_normal_seconds": fit_fast_s,
                         "note": "seconds: fit_normal_kls (KL evaluated and yielded after every sweep); fit_normal_seconds: fit_normal(), one library call"}}
    for what, code in (("uv_entropy", N.LOOK_ENTROPY), ("total_variance", N.LOOK_TOTAL_VARIANCE)):
        ms = cuda_ms(lambda: post.lookahead(code, ci, cj, vals, N.WEIGHTS_DISCRETE, bounds, mu, sd))
        out[what] = {"kernel_ms": ms, "cand_per_s": len(ii) / (ms * 1e-3),
                     "refits_per_s": 2 * len(ii) / (ms * 1e-3)}
    _, sc, best = post.lookahead(N.LOOK_ENTROPY, ci, cj, vals, N.WEIGHTS_DISCRETE, bounds, mu, sd)
    sc = sc.cpu().numpy()
    out["parity"] = {"uv_entropy_max_rel_vs_fixture": float(np.abs(sc / g["b_uv_entropy"] - 1).max()),
                     "same_pick": int(S.unpack_best(best)[1]) == int(np.argmin(g["b_uv_entropy"])),
                     "kl_rel_vs_reference_kl_divergence": abs(a.kl_divergence() / float(g["ref_kl_at_blocks"]) - 1)}
    out["class_api_pick_ms"] = wall_ms(lambda: a.pick_query_point(pool, A.ActivePMF.exp_approx_entropy))
    out["numpy_oracle_seconds_same_pool"] = float(g["oracle_lookahead_seconds"])
    # per (candidate, value): one d x d precision update + Cholesky; per candidate one more with
    # its inverse (fp64)
    flops = len(ii) * (2 * (d ** 3 / 3 + 4 * d * d) + (d ** 3 / 3 + 2 * d ** 3))
    out["uv_entropy"]["approx_fp64_gflops"] = flops / (out["uv_entropy"]["kernel_ms"] * 1e-3) / 1e9
    return out


def c3_movielens(hbm_peak):
    """C3: 943x1682, 5,000 known, rank 10: MAP fit, `pred` (S1) and block-posterior pred_variance
    (S2) over all 1,581,126 unknown cells, and the class API end to end"""
    import ctypes as C
    from active_matrix_factorization_b200 import _native as N, active_pmf as A, device as D, scoring as S
    g = golden("c3_movielens")
    R = g["ratings"].astype(float)
    n, m, d = 943, 1682, 10
    np.random.seed(0)
    a = A.ActivePMF(R, d, rating_values={1, 2, 3, 4, 5}, discrete_expectations=True, knowable=())
    a.log_likelihood()
    t0 = time.perf_counter()
    a.fit()
    fit_s = time.perf_counter() - t0
    out = {"config": "C3 movielens-100k 943x1682, 5,000 known, rank 10; all 1,581,126 unknown cells scored",
           "map_fit_one_launch_s": fit_s, "final_ll": a.log_likelihood()}
    a.blocks_tol, a.blocks_max_sweeps = 1e-9, 300
    t0 = time.perf_counter()
    a.initialize_approx()
    sweeps = len(list(a.fit_normal_kls()))
    out["block_fit"] = {"sweeps": sweeps, "seconds": time.perf_counter() - t0,
                        "note": "fit_normal_kls: the KL is evaluated and yielded after every sweep"}
    t0 = time.perf_counter()
    a.initialize_approx()
    a.fit_normal()
    out["block_fit"]["fit_normal_seconds"] = time.perf_counter() - t0   # one library call, no KL per sweep
    ii, jj = unknown_cells(R, n, m)
    nc = len(ii)
    out["candidates"] = nc
    allpool = np.column_stack((ii, jj)).astype(np.int32)
    ci, cj = D.to_device(ii, np.int32), D.to_device(jj, np.int32)
    post = a._block_posterior()
    d2 = d * (d + 1)
    for name, es in (("f32", 4), ("f64", 8)):
        U, V = D.to_padded(a.users, name), D.to_padded(a.items, name)
        ms = cuda_ms(lambda: S.score_device(N.CRIT_PRED, name, ci, cj, d, U, V, want_scores=False))
        out["S1_pred_" + name] = roof(nc * 8 + (n + m) * d * es, nc * 2 * d, ms, hbm_peak)
        out["S1_pred_" + name]["cand_per_s"] = nc / (ms * 1e-3)
        post.packed(name)
        ms = cuda_ms(lambda: post.score(N.CRIT_PRED_VARIANCE, ci, cj, name, want_scores=False))
        out["S2_pred_variance_" + name] = roof(nc * 8 + (n + m) * d2 * es, nc * 2 * d2, ms, hbm_peak)
        out["S2_pred_variance_" + name]["cand_per_s"] = nc / (ms * 1e-3)
    # scalable-mode lookahead over ALL unknown cells (5 rating values each): rows L1 of SURVEY 8d
    import torch
    Um, Vm = D.to_padded(a.users, "f64"), D.to_padded(a.items, "f64")
    mu, _ = S.score_device(N.CRIT_PRED, "f64", ci, cj, d, Um, Vm)
    sd = torch.ones_like(mu)
    vals5 = np.array([1., 2., 3., 4., 5.])
    for what, code in (("uv_entropy", N.LOOK_ENTROPY), ("total_variance", N.LOOK_TOTAL_VARIANCE)):
        ms = cuda_ms(lambda: post.lookahead(code, ci, cj, vals5, N.WEIGHTS_DISCRETE, a.rating_bounds, mu, sd,
                                            want_scores=False), reps=3, warm=1)
        out["L1_scalable_" + what] = {"kernel_ms": ms, "cand_per_s": nc / (ms * 1e-3),
                                      "refits_per_s": 5 * nc / (ms * 1e-3)}
    # class API end to end against the host-buffer C-ABI call on the same pool
    lib = N.require_device()
    U_h = np.ascontiguousarray(a.users, dtype=np.float64)
    V_h = np.ascontiguousarray(a.items, dtype=np.float64)
    ii32, jj32 = np.ascontiguousarray(ii, np.int32), np.ascontiguousarray(jj, np.int32)
    best_h = N.Best()

    def abi():
        N.check(lib.amf_score_pred_host(N.F64, nc, N.host_ptr(ii32), N.host_ptr(jj32), n, m, d,
                                        N.host_ptr(U_h), N.host_ptr(V_h), None, 1, C.byref(best_h)))
    abi_ms = wall_ms(abi, reps=5)

    def fresh():
        a._dev.pop('pool_array', None)
        return a.pick_query_point(allpool, A.ActivePMF.pred)
    fresh_ms = wall_ms(fresh, reps=5)
    cached_ms = wall_ms(lambda: a.pick_query_point(allpool, A.ActivePMF.pred), reps=5)
    pick = a.pick_query_point(allpool, A.ActivePMF.pred)
    out["e2e_class_api"] = {
        "criterion": "pred", "c_abi_host_call_ms": abi_ms, "pick_query_point_ms_pool_uploaded_each_call": fresh_ms,
        "pick_query_point_ms_same_pool_object": cached_ms, "cand_per_s": nc / (fresh_ms * 1e-3),
        "ratio_to_c_abi": fresh_ms / abi_ms,
        "same_pick": pick == (int(ii[best_h.index]), int(jj[best_h.index]))}
    pv_ms = wall_ms(lambda: a.pick_query_point(allpool, A.ActivePMF.pred_variance), reps=5)
    out["e2e_class_api"]["pick_pred_variance_ms"] = pv_ms
    cp = S.CandidatePool(allpool)
    out["e2e_class_api"]["pick_pred_variance_ms_resident_pool"] = wall_ms(
        lambda: a.pick_query_point(cp, A.ActivePMF.pred_variance), reps=5)
    return out


def c4_bayes(hbm_peak):
    """C4: BayesianPMF rank 15 on the same split, 200 Gibbs samples (B1), variance selection over
    all unrated cells (S3)"""
    import torch
    from active_matrix_factorization_b200 import _native as N, bayes_pmf as Bm, device as D
    g = golden("c4_movielens_bayes")
    R = g["ratings"].astype(float)
    n, m, d, S_ = 943, 1682, 15, 200
    np.random.seed(0)
    b = Bm.BayesianPMF(R, d, subtract_mean=True, rating_values={1, 2, 3, 4, 5}, knowable=())
    b.fit()
    np.random.seed(7)
    list(islice(b.samples(num_gibbs=2), 3))
    np.random.seed(7)
    t0 = time.perf_counter()
    samples = list(islice(b.samples(num_gibbs=2), S_))
    gibbs_s = time.perf_counter() - t0
    out = {"config": "C4 BayesianPMF rank 15, movielens-100k split, 200 samples, variance selection over all unrated cells",
           "gibbs_200_samples_s": gibbs_s, "row_conditionals_per_s": S_ * 2 * (n + m) / gibbs_s,
           "rng": "host (numpy legacy stream, the reference's draw order)"}
    if hasattr(b, "samples_device"):
        torch.manual_seed(0)
        list(islice(b.samples_device(num_gibbs=2), 3))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fast = list(islice(b.samples_device(num_gibbs=2), S_))
        torch.cuda.synchronize()
        fs = time.perf_counter() - t0
        out["gibbs_200_samples_device_rng_s"] = fs
        out["row_conditionals_per_s_device_rng"] = S_ * 2 * (n + m) / fs
        out["rng_device"] = "Philox in the kernels; Normal-Wishart hyper-parameters drawn on the device (amf_gibbs_hyper_device): no host round trip per sample"
        del fast
        list(islice(b.samples_device(num_gibbs=2, hyper='host'), 3))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fast = list(islice(b.samples_device(num_gibbs=2, hyper='host'), S_))
        torch.cuda.synchronize()
        out["gibbs_200_samples_device_rng_host_hyper_s"] = time.perf_counter() - t0
        del fast
    lib = N.require_device()
    for name, es in (("f64", 8), ("f32", 4)):
        rat = D.Ratings.from_tuples(R, n, m, name)
        dt = D.np_dtype(name)
        it = D.to_device(samples[-1][1], dt)
        us = D.to_device(samples[-1][0], dt)
        alpha, mu = D.to_device(np.eye(d) * 2, dt), D.to_device(np.zeros(d), dt)
        zu, zv = torch.randn((n, d), dtype=it.dtype, device=it.device), torch.randn((m, d), dtype=it.dtype, device=it.device)
        ou, ov = torch.empty_like(zu), torch.empty_like(zv)

        def sweep():
            N.check(lib.amf_gibbs_half_sweep(rat.handle, 0, D.code(name), d, D.ptr(it), D.ptr(alpha),
                                             D.ptr(mu), 2.0, float(b.mean_rating), D.ptr(zu), D.ptr(ou),
                                             D.stream_ptr()))
            N.check(lib.amf_gibbs_half_sweep(rat.handle, 1, D.code(name), d, D.ptr(us), D.ptr(alpha),
                                             D.ptr(mu), 2.0, float(b.mean_rating), D.ptr(zv), D.ptr(ov),
                                             D.stream_ptr()))
        ms = cuda_ms(sweep)
        nnz = len(R)
        bytes_ = 2 * nnz * (4 + es) + (n + m) * (3 * d * es)
        flops = 2 * nnz * (2 * d * d + 2 * d) + (n + m) * d ** 3
        out["B1_gibbs_sweep_" + name] = roof(bytes_, flops, ms, hbm_peak)
        out["B1_gibbs_sweep_" + name]["rows_per_s"] = (n + m) / (ms * 1e-3)

        def fast_sweep():
            N.check(lib.amf_gibbs_half_sweep_device_rng(rat.handle, 0, D.code(name), d, D.ptr(it), D.ptr(alpha),
                                                        D.ptr(mu), 2.0, float(b.mean_rating), 1, 2, D.ptr(ou),
                                                        0, -1, D.stream_ptr()))
            N.check(lib.amf_gibbs_half_sweep_device_rng(rat.handle, 1, D.code(name), d, D.ptr(us), D.ptr(alpha),
                                                        D.ptr(mu), 2.0, float(b.mean_rating), 1, 3, D.ptr(ov),
                                                        0, -1, D.stream_ptr()))
        ms = cuda_ms(fast_sweep)
        out["B1_gibbs_sweep_device_rng_" + name] = roof(bytes_ - (n + m) * d * es, flops, ms, hbm_peak)
        out["B1_gibbs_sweep_device_rng_" + name]["rows_per_s"] = (n + m) / (ms * 1e-3)
        rat.close()
    known = np.zeros((n, m), bool)
    known[R[:, 0].astype(int), R[:, 1].astype(int)] = True
    which = np.nonzero(~known)
    for name, es in (("f32", 4), ("f64", 8)):
        dt = D.np_dtype(name)
        Us = D.to_device(np.stack([s[0] for s in samples]), dt)
        Vs = D.to_device(np.stack([s[1] for s in samples]), dt)
        var = torch.empty(n * m, dtype=Us.dtype, device=Us.device)
        best = torch.empty(2, dtype=torch.int64, device=Us.device)

        def stats():
            N.check(lib.amf_bayes_sample_stats(D.code(name), n * m, None, None, S_, n, m, d, D.ptr(Us),
                                               D.ptr(Vs), float(b.mean_rating), 0.0, None, D.ptr(var),
                                               None, 1, 1, 0, D.ptr(best), D.stream_ptr()))
        ms = cuda_ms(stats)
        out["S3_sample_variance_dense_" + name] = roof(S_ * (n + m) * d * es + n * m * es,
                                                       n * m * S_ * (2 * d + 3), ms, hbm_peak)
        out["S3_sample_variance_dense_" + name]["cells_per_s"] = n * m / (ms * 1e-3)
    b.compute_dtype = "f64"
    out["pred_variance_call_ms"] = wall_ms(lambda: b.pred_variance(samples, which=which), reps=3)
    return out


def c5_extra(rat, n, m, d, ci, cj, hbm_peak, name="f32"):
    """C5 scale (200k x 50k, rank 32): block-posterior pred_variance over the bench's candidate
    shard (S2: 4.2 KB packed rows) and one Gibbs sweep over its rating list (B1)"""
    import torch
    from active_matrix_factorization_b200 import _native as N, blocks as BL, device as D
    es = 4 if name == "f32" else 8
    dev = ci.device
    out = {}
    lib = N.require_device()
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    dt = torch.float32 if name == "f32" else torch.float64
    it = torch.randn((m, d), generator=g, device=dev, dtype=dt) * .4
    us = torch.randn((n, d), generator=g, device=dev, dtype=dt) * .4
    alpha = torch.eye(d, device=dev, dtype=dt) * 2
    mu = torch.zeros(d, device=dev, dtype=dt)
    zu = torch.randn((n, d), generator=g, device=dev, dtype=dt)
    zv = torch.randn((m, d), generator=g, device=dev, dtype=dt)
    ou, ov = torch.empty_like(zu), torch.empty_like(zv)

    def sweep():
        N.check(lib.amf_gibbs_half_sweep(rat.handle, 0, D.code(name), d, D.ptr(it), D.ptr(alpha), D.ptr(mu),
                                         2.0, 0.0, D.ptr(zu), D.ptr(ou), D.stream_ptr()))
        N.check(lib.amf_gibbs_half_sweep(rat.handle, 1, D.code(name), d, D.ptr(us), D.ptr(alpha), D.ptr(mu),
                                         2.0, 0.0, D.ptr(zv), D.ptr(ov), D.stream_ptr()))
    ms = cuda_ms(sweep, reps=3, warm=1)
    nnz = rat.nnz
    out["B1_gibbs_sweep_c5_" + name] = roof(2 * nnz * (4 + es) + (n + m) * 3 * d * es,
                                            2 * nnz * (2 * d * d + 2 * d) + (n + m) * d ** 3, ms, hbm_peak)
    out["B1_gibbs_sweep_c5_" + name]["rows_per_s"] = (n + m) / (ms * 1e-3)

    def fast_sweep():
        N.check(lib.amf_gibbs_half_sweep_device_rng(rat.handle, 0, D.code(name), d, D.ptr(it), D.ptr(alpha),
                                                    D.ptr(mu), 2.0, 0.0, 1, 2, D.ptr(ou), 0, -1, D.stream_ptr()))
        N.check(lib.amf_gibbs_half_sweep_device_rng(rat.handle, 1, D.code(name), d, D.ptr(us), D.ptr(alpha),
                                                    D.ptr(mu), 2.0, 0.0, 1, 3, D.ptr(ov), 0, -1, D.stream_ptr()))
    ms = cuda_ms(fast_sweep, reps=3, warm=1)
    out["B1_gibbs_sweep_device_rng_c5_" + name] = roof(2 * nnz * (4 + es) + (n + m) * 2 * d * es,
                                                       2 * nnz * (2 * d * d + 2 * d) + (n + m) * d ** 3, ms, hbm_peak)
    out["B1_gibbs_sweep_device_rng_c5_" + name]["rows_per_s"] = (n + m) / (ms * 1e-3)
    del zu, zv, ou, ov
    # block posterior at the MAP curvature (one Gram pass per side), then the variance criterion
    post = BL.BlockPosterior(n, m, d, 1.0, 10.0, 10.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    post.fit(rat, us.double().cpu().numpy(), it.double().cpu().numpy(), sweeps=1, cov_term=False,
             update_mean=False)
    torch.cuda.synchronize()
    out["block_gram_pass_c5_s"] = time.perf_counter() - t0
    post.packed(name)
    torch.cuda.synchronize()
    nc = int(ci.numel())
    d2 = d * (d + 1)
    ms = cuda_ms(lambda: post.score(N.CRIT_PRED_VARIANCE, ci, cj, name, want_scores=False), reps=3, warm=1)
    out["S2_pred_variance_c5_" + name] = roof(nc * 8 + (n + m) * d2 * es, nc * 2 * d2, ms, hbm_peak)
    out["S2_pred_variance_c5_" + name]["cand_per_s"] = nc / (ms * 1e-3)
    out["S2_pred_variance_c5_" + name]["note"] = "packed rows of d(d+1) = %d numbers (%.1f KB): one random item row per candidate through L2" % (d2, d2 * es / 1024)
    return out


def all_configs(hbm_peak):
    out = {}
    for key, fn in (("c1", c1_exact_lookahead), ("c2", lambda: c2_drugbank(hbm_peak)),
                    ("c3", lambda: c3_movielens(hbm_peak)), ("c4", lambda: c4_bayes(hbm_peak))):
        try:
            out[key] = fn()
        except Exception as exc:                      # a failed config must not lose the C5 line
            out[key] = {"failed": repr(exc)[:300]}
    return out


if __name__ == "__main__":
    from active_matrix_factorization_b200 import build
    build.build()
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
    except Exception:
        peak = 6650.0
    print(json.dumps(all_configs(peak), default=float))
