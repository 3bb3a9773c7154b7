"""GPU parity of the Gibbs sampler and the sample-based criteria against the reference."""
from itertools import islice

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Bm():
    from active_matrix_factorization_b200 import build
    build.build()
    from active_matrix_factorization_b200 import bayes_pmf
    return bayes_pmf


def test_known_answer_table(Bm, golden):
    """SURVEY.md 8c Bayesian rows."""
    g = golden("known_answer_10x10_d2")
    R, U, V = g["ratings"], g["users"], g["items"]
    b = Bm.BayesianPMF(R, 2, subtract_mean=False)
    b.users, b.items = U.copy(), V.copy()
    sel = R[:, 0] == 0
    np.random.seed(0)
    x = b.sample_feature(0, True, np.zeros(2), np.eye(2), V, R[sel, 1].astype(int), R[sel, 2])
    np.testing.assert_allclose(x, [-28.04639559940905, 49.27729944075128], rtol=1e-10)
    np.random.seed(0)
    mu, alpha = b.sample_hyperparam(U, True)
    np.testing.assert_allclose(mu, [1.8946544328590074, 1.9165495158619739], rtol=1e-12)
    np.testing.assert_allclose(alpha, g["hyper_alpha"], rtol=1e-11)
    np.random.seed(0)
    s = list(islice(b.samples(num_gibbs=2), 3))
    np.testing.assert_allclose(s[2][0][0], [5.911244993127693, 17.59631201803792], rtol=1e-8)
    np.testing.assert_allclose(s[2][1][9], [-1.7869175224238054, 3.3705880307551848], rtol=1e-8)
    np.testing.assert_allclose(np.array([x[0] for x in s]), g["samples_u"], rtol=1e-7, atol=1e-9)
    pv = b.pred_variance(s)
    np.testing.assert_allclose(pv, g["bayes_pred_variance"], rtol=1e-8)
    assert pv.sum() == pytest.approx(16816.729160266303, rel=1e-8)
    assert b.predict(s)[4, 7] == pytest.approx(84.793209488960429, rel=1e-9)
    np.testing.assert_array_equal(b.prob_ge_cutoff(s, .5), g["bayes_prob_ge_half"])


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-7), ("f32", 2e-3)])
def test_chain_subtract_mean(Bm, golden, dtype, tol):
    g = golden("gibbs_15x12_d3")
    b = Bm.BayesianPMF(g["ratings"], 3)
    b.compute_dtype = dtype
    b.users, b.items = g["users"].copy(), g["items"].copy()
    np.random.seed(int(g["seed"]))
    s = list(islice(b.samples(num_gibbs=2), 6))
    scale = np.abs(g["samples_u"]).max()
    np.testing.assert_allclose(np.array([x[0] for x in s]), g["samples_u"], rtol=tol, atol=tol * scale)
    np.testing.assert_allclose(np.array([x[1] for x in s]), g["samples_v"], rtol=tol, atol=tol * scale)
    if dtype == "f64":
        ii, jj = g["cand_i"], g["cand_j"]
        which = (ii, jj)
        np.testing.assert_allclose(b.predict(s, which=which), g["bayes_predict"], rtol=1e-8)
        np.testing.assert_allclose(b.pred_variance(s, which=which), g["bayes_pred_variance"], rtol=1e-7)
        np.testing.assert_array_equal(b.prob_ge_cutoff(s, 3.5, which=which), g["bayes_prob_ge_3_5"])
        assert b.total_variance(s) == pytest.approx(float(g["bayes_total_variance"]), rel=1e-7)
        # selection as in bayes_pmf.full_test (:702-712)
        evals = b.pred_variance(s, which=which)
        assert int(np.argmax(evals)) == int(np.argmax(g["bayes_pred_variance"]))
        assert b.matrix_results(evals, which).shape == (15, 12)


def test_keys_registry(Bm):
    assert set(Bm.KEYS) == {'random', 'pred-variance', 'exp-variance', 'pred', 'prob-ge-3.5',
                            'prob-ge-.5', 'prob-ge-0'}
    k = Bm.KEYS['prob-ge-3.5']
    assert (k.key_fn, k.choose_max, k.wants_pool, k.args) == ('prob_ge_cutoff', True, False, (3.5,))


def test_active_loop_two_steps(Bm, ref_drivers):
    """The reference's own bayes_pmf.compare_active (loaded by drivers.load) end to end on a
    tiny problem (pred-variance + exp-variance), driving the GPU-backed BayesianPMF."""
    drv = ref_drivers("bayes_pmf")
    assert drv.BayesianPMF is Bm.BayesianPMF
    import random
    np.random.seed(1); random.seed(1)
    n, m, d = 6, 5, 2
    u, v = np.random.normal(0, 1, (n, d)), np.random.normal(0, 1, (m, d))
    real = np.clip(np.round(u @ v.T + 3), 1, 5)
    known = [(i, (i * 2 + t) % m) for i in range(n) for t in range(2)] + [(0, 3), (1, 1)]
    known = sorted(set(known))
    ratings = np.array([(i, j, real[i, j]) for i, j in known], dtype=float)
    assert set(ratings[:, 1].astype(int)) == set(range(m))
    res = drv.compare_active(['pred-variance'], d, real, ratings, rating_vals=(1, 2, 3, 4, 5),
                             num_steps=3, num_samps=8, threaded=False, procs=0)
    steps = res['pred-variance']
    assert len(steps) == 3 and steps[1][0] == len(known) + 1 and steps[2][0] == len(known) + 2
    assert steps[1][3].shape == (n, m) and np.isfinite(steps[1][1])
    b = res['_initial_bpmf']
    samples = list(islice(b.samples(), 4))
    cand = sorted(b.unrated)[:2]
    which = tuple(np.array(cand).T)
    ev = b.exp_variance(samples, which=which, num_samps=4, fit_first=False)
    assert ev.shape == (2,) and np.all(np.isfinite(ev)) and np.all(ev > 0)


@pytest.mark.parametrize("d", [32, 64])
def test_tensor_core_gram_half_sweep(d):
    """fp32, d = 32 / 64: the Gram matrix of the row conditionals (bayes_pmf.py:189-216) runs on
    the tensor cores (3xTF32).  Against the fp64 kernel on the same inputs the sampled rows agree
    to fp32 accuracy; ragged rows (0, 1, 7, 8, 9, 64, 65, 200 ratings) cover the zero-filled
    k-steps and the multi-tile loop."""
    import ctypes as C
    import torch
    from active_matrix_factorization_b200 import build
    build.build()
    from active_matrix_factorization_b200 import _native as N, device as D
    lib = N.require_device()
    rng = np.random.RandomState(d)
    counts = [0, 1, 7, 8, 9, 64, 65, 200, 31, 130]
    n, m = len(counts), 260
    ii = np.concatenate([np.full(c, r) for r, c in enumerate(counts)]).astype(np.int32)
    jj = np.concatenate([rng.permutation(m)[:c] for c in counts]).astype(np.int32)
    r = rng.normal(3, 1, len(ii))
    V = rng.normal(0, .4, (m, d))
    W = rng.normal(size=(d, d)); alpha = W @ W.T / d + np.eye(d) * 2.0
    mu = rng.normal(0, .1, d)
    z = rng.normal(size=(n, d))
    out = {}
    for dtype in ("f64", "f32"):
        dt = D.torch_dtype(dtype)
        rat = D.Ratings(n, m, ii, jj, r, dtype)
        Vt = torch.from_numpy(V).to(dt).cuda()
        al, mt, zt = (torch.from_numpy(x).to(dt).cuda() for x in (alpha, mu, z))
        o = torch.empty((n, d), dtype=dt, device="cuda")
        N.check(lib.amf_gibbs_half_sweep(rat.handle, 0, D.code(dtype), d, D.ptr(Vt), D.ptr(al), D.ptr(mt),
                                         2.0, 3.0, D.ptr(zt), D.ptr(o), D.stream_ptr()))
        failed = C.c_int(0)
        N.check(lib.amf_gibbs_status(rat.handle, C.byref(failed), D.stream_ptr()))
        assert failed.value == 0
        out[dtype] = o.double().cpu().numpy()
    # and against the oracle's sample_feature for the heaviest row
    row = 7
    sel = ii == row
    from oracle import pmf_oracle as O
    ref = O.sample_feature(mu, alpha, V, jj[sel], r[sel] - 3.0, beta=2.0, z=z[row])
    np.testing.assert_allclose(out["f64"][row], ref, rtol=1e-9, atol=1e-11)
    scale = np.abs(out["f64"]).max()
    assert np.abs(out["f32"] - out["f64"]).max() <= 2e-5 * scale


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-12), ("f32", 2e-5)])
@pytest.mark.parametrize("n,m,d,S", [(45, 70, 3, 7), (33, 129, 15, 20), (64, 64, 40, 5), (1, 5, 2, 1)])
def test_dense_sample_stats_match_numpy_and_the_gather_kernel(Bm, n, m, d, S, dtype, tol):
    """amf_bayes_sample_stats, dense form (blocked product over shared-memory tiles; `which` =
    the whole matrix or most of it) against numpy over the samples (bayes_pmf.py:433-455,528-538)
    and against the per-candidate kernel that sparse `which` sets use; ragged tile edges."""
    rng = np.random.RandomState(n + d)
    R = np.column_stack((rng.randint(0, n, 60), rng.randint(0, m, 60), rng.randint(1, 6, 60))).astype(float)
    R[0, :2] = (n - 1, m - 1)
    b = Bm.BayesianPMF(R, d)
    b.compute_dtype = dtype
    samples = [(rng.normal(size=(n, d)), rng.normal(size=(m, d))) for _ in range(S)]
    preds = np.array([u @ v.T + b.mean_rating for u, v in samples])
    scale = np.abs(preds).max()
    want = dict(mean=preds.mean(0), var=preds.var(0), prob=(preds >= 3.5).mean(0))
    got = dict(mean=b.predict(samples), var=b.pred_variance(samples), prob=b.prob_ge_cutoff(samples, 3.5))
    for k in want:
        assert got[k].shape == (n, m)
        np.testing.assert_allclose(got[k], want[k], rtol=tol, atol=tol * scale * scale, err_msg=k)
    assert b.total_variance(samples) == pytest.approx(want["var"].sum(), rel=tol * 10)
    # most of the matrix as index arrays -> dense kernel + gather; a few cells -> gather kernel
    mask = rng.uniform(size=(n, m)) < .7
    mask[0, 0] = True
    big = np.nonzero(mask)
    few = (np.array([0, n - 1, n // 2]), np.array([m - 1, 0, m // 3]))
    for which in (big, few, mask):
        np.testing.assert_allclose(b.pred_variance(samples, which=which), want["var"][which],
                                   rtol=tol, atol=tol * scale * scale)
        np.testing.assert_allclose(b.predict(samples, which=which), want["mean"][which],
                                   rtol=tol, atol=tol * scale)
    b._DENSE_WHICH_FRACTION = 2.0                       # force the per-candidate kernel everywhere
    np.testing.assert_allclose(b.pred_variance(samples, which=big), got["var"][big],
                               rtol=tol, atol=tol * scale * scale)
    np.testing.assert_array_equal(b.prob_ge_cutoff(samples, 3.5, which=big), got["prob"][big])


@pytest.mark.parametrize("tag,kw", [("disc", dict(rating_values=(1, 2, 3, 4, 5), discrete_expectations=True)),
                                    ("cont", dict(rating_values=None, discrete_expectations=False,
                                                  num_integration_pts=5))])
def test_exp_variance_golden(Bm, golden, tag, kw):
    """Bayesian lookahead (bayes_pmf.py:457-525,560-602): for every candidate and rating value a
    copy of the model gets the rating and runs its own short chain from the global numpy stream,
    in the reference's draw order -- seeded, it reproduces the reference's expected total variance
    (categorical fit summed over the values; normal fit integrated over ppf points).  Measured on
    B200: 4e-8 relative (the compiled reference rounds alpha / denom to C floats)."""
    g, c = golden("gibbs_15x12_d3"), golden("more_criteria")
    b = Bm.BayesianPMF(g["ratings"], 3, **kw)
    b.compute_dtype = "f64"
    b.users, b.items = g["users"].copy(), g["items"].copy()
    samples = list(zip(g["samples_u"], g["samples_v"]))
    which = (c["ev_cand_i"], c["ev_cand_j"])
    assert [tuple(x) for x in np.array(which).T] == sorted(b.unrated)[:2]
    np.random.seed(5)
    ev = b.exp_variance(samples, which=which, num_samps=3, fit_first=False)
    np.testing.assert_allclose(ev, c["ev_" + tag], rtol=1e-6)


# ---- fast mode: device random numbers, one Cholesky per row -----------------------------------
def _philox4x32_10(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = list(ctr)
    k = list(key)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xffffffff, p1 & 0xffffffff,
             ((p0 >> 32) ^ c[3] ^ k[1]) & 0xffffffff, p0 & 0xffffffff]
        k = [(k[0] + W0) & 0xffffffff, (k[1] + W1) & 0xffffffff]
    return c


def _device_normals(seed, stream_id, rows, d, dtype="f64"):
    import torch
    from active_matrix_factorization_b200 import _native as N, device as D
    out = torch.empty((rows, d), dtype=D.torch_dtype(dtype), device="cuda")
    N.check(N.require_device().amf_philox_normal(D.code(dtype), seed, stream_id, rows, d, D.ptr(out),
                                                 D.stream_ptr()))
    return out.double().cpu().numpy()


def test_philox_known_answer_and_normals():
    """Philox4x32-10: the published test vectors (Random123 kat_vectors) for the host restatement,
    and the kernel's normals against Box-Muller on that restatement's words"""
    assert _philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    seed, stream = 0x0123456789abcdef, (7 << 32) | 5
    z = _device_normals(seed, stream, 6, 4)
    for row in range(6):
        for k in range(4):
            w = _philox4x32_10([row, k, stream & 0xffffffff, stream >> 32], [seed & 0xffffffff, seed >> 32])
            u1, u2 = (w[0] + 1.0) * 2.0 ** -32, (w[1] + 1.0) * 2.0 ** -32
            assert z[row, k] == pytest.approx(np.sqrt(-2 * np.log(u1)) * np.cos(2 * np.pi * u2), rel=1e-12, abs=1e-13)
    big = _device_normals(3, 1, 20000, 16)
    assert abs(big.mean()) < 0.01 and abs(big.var() - 1) < 0.01
    assert abs(np.corrcoef(big[:, 0], big[:, 1])[0, 1]) < 0.03


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-9), ("f32", 2e-4)])
@pytest.mark.parametrize("d", [3, 15, 32])
def test_fast_half_sweep_is_the_same_conditional(d, dtype, tol):
    """amf_gibbs_half_sweep_device_rng: with z the kernel's own normals (amf_philox_normal, same
    counters) every row equals Lambda^-1 rhs + R^-T z for Lambda = R R' (bayes_pmf.py:189-216
    with the one-factor map); its mean and covariance are those of the reference's conditional"""
    import torch
    from active_matrix_factorization_b200 import _native as N, device as D
    lib = N.require_device()
    rng = np.random.RandomState(d)
    n, m = 40, 30
    cells = rng.permutation(n * m)[:500]
    R = np.column_stack((cells // m, cells % m, rng.normal(3, 1, 500)))
    R = R[R[:, 0] != 2]                                     # a row without ratings
    other = rng.normal(0, .5, (m, d))
    a0 = rng.normal(0, 1, (d, d))
    alpha, mu, beta, off = a0 @ a0.T / d + np.eye(d), rng.normal(0, .3, d), 2.0, 3.0
    rat = D.Ratings.from_tuples(R, n, m, dtype)
    dt = D.np_dtype(dtype)
    out = torch.empty((n, d), dtype=D.torch_dtype(dtype), device="cuda")
    seed, stream = 99, 12
    other_t, alpha_t, mu_t = D.to_device(other, dt), D.to_device(alpha, dt), D.to_device(mu, dt)
    N.check(lib.amf_gibbs_half_sweep_device_rng(
        rat.handle, 0, D.code(dtype), d, D.ptr(other_t), D.ptr(alpha_t), D.ptr(mu_t), beta, off,
        seed, stream, D.ptr(out), 0, -1, D.stream_ptr()))
    got = out.double().cpu().numpy()
    z = _device_normals(seed, stream, n, d)
    for i in range(n):
        rows = R[R[:, 0] == i]
        F = other[rows[:, 1].astype(int)]
        lam = alpha + beta * F.T @ F
        rhs = beta * F.T @ (rows[:, 2] - off) + alpha @ mu
        Rf = np.linalg.cholesky(lam)
        want = np.linalg.solve(Rf.T, np.linalg.solve(Rf, rhs) + z[i])
        assert np.abs(got[i] - want).max() <= tol * max(1.0, np.abs(want).max())
    rat.close()


def test_fast_chain_matches_host_chain_in_law(Bm, golden):
    """device-RNG chain against the host-RNG (reference-order) chain on the same model: posterior
    mean and variance of the predictions agree within Monte-Carlo error"""
    g = golden("gibbs_15x12_d3")
    b = Bm.BayesianPMF(g["ratings"], 3, subtract_mean=True)
    b.users, b.items = g["users"].copy(), g["items"].copy()
    S, burn = 1500, 100
    np.random.seed(11)
    host = list(islice(b.samples(num_gibbs=2), S + burn))[burn:]
    np.random.seed(12)
    b.device_seed = 5
    dev = list(islice(b.samples_device(num_gibbs=2), S + burn))[burn:]
    import torch
    assert isinstance(dev[0][0], torch.Tensor)
    mh, md = b.predict(host), b.predict(dev)
    vh, vd = b.pred_variance(host), b.pred_variance(dev)
    # Monte-Carlo error of a mean over ~S/10 effective samples of spread sqrt(v)
    assert np.abs(mh - md).max() < 6 * np.sqrt(vh.max() / (S / 10))
    assert np.abs(np.log(vd / vh)).max() < 0.6
    # drop-in switch: samples() in device mode yields host arrays
    b.rng_mode = 'device'
    us, vs = next(b.samples())
    assert isinstance(us, np.ndarray) and us.shape == (15, 3) and vs.shape == (12, 3)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("rows,d", [(400, 6), (37, 3), (5000, 15)])
def test_device_hyperparameter_draws_follow_the_normal_wishart_posterior(Bm, dtype, rows, d):
    """amf_gibbs_hyper_device against the posterior of bayes_pmf.py:158-186 in law: with W = inv(M),
    nu = dof0 + n,  E[alpha] = nu W,  Var[alpha_kl] = nu (W_kl^2 + W_kk W_ll),  E[mu] = mu*,
    Cov[mu] = inv(W) / ((b0 + n)(nu - d - 1)) -- each checked within Monte-Carlo error; one draw with
    hyper='host' stream-for-stream is not expected to be equal."""
    import ctypes as C
    import torch
    from active_matrix_factorization_b200 import _native as N, device as D
    lib = N.require_device()
    rng = np.random.RandomState(rows + d)
    A = rng.normal(size=(d, d)) / np.sqrt(d)
    feats = rng.normal(size=(rows, d)) @ (np.eye(d) + A) + rng.normal(size=d)
    wi = np.eye(d) + 0.3 * (A @ A.T)
    b0, df, mu0 = 2.0, float(d) + 0.7, rng.normal(size=d) * 0.3     # fractional dof: truncated
    rat = D.Ratings(rows, 4, np.array([0], np.int32), np.array([0], np.int32), np.array([1.0]), dtype)
    x = D.to_device(feats, D.np_dtype(dtype))
    prior = D.to_device(np.concatenate((np.linalg.inv(wi).reshape(-1), mu0, [b0, df])), np.float64)
    K = 4000
    out = torch.empty((K, d + d * d), dtype=D.torch_dtype(dtype), device=x.device)
    for k in range(K):
        N.check(lib.amf_gibbs_hyper_device(rat.handle, D.code(dtype), d, rows, D.ptr(x), D.ptr(prior), 77, 1000 + k,
                                           D.ptr(out[k, :d]), D.ptr(out[k, d:]), D.stream_ptr()))
    failed = C.c_int(0)
    N.check(lib.amf_gibbs_status(rat.handle, C.byref(failed), D.stream_ptr()))
    assert failed.value == 0
    o = out.double().cpu().numpy()
    mus, alphas = o[:, :d], o[:, d:].reshape(K, d, d)
    xs = feats.astype(D.np_dtype(dtype)).astype(float)
    x_bar, s_bar, n = xs.mean(0), np.cov(xs, rowvar=0), rows
    diff = mu0 - x_bar
    W = np.linalg.inv(np.linalg.inv(wi) + n * s_bar + (b0 * n) / (b0 + n) * np.dot(diff, diff.T))
    nu = int(df + n)
    np.testing.assert_allclose(alphas, alphas.transpose(0, 2, 1), rtol=1e-5, atol=1e-7 * np.abs(alphas).max())
    se = np.sqrt(nu * (W ** 2 + np.outer(np.diag(W), np.diag(W))) / K)
    assert (np.abs(alphas.mean(0) - nu * W) < 5 * se).all()
    assert np.abs(alphas.var(0) / (nu * (W ** 2 + np.outer(np.diag(W), np.diag(W)))) - 1).max() < 0.2
    mu_star = (b0 * mu0 + n * x_bar) / (b0 + n)
    cov_mu = np.linalg.inv(W) / ((b0 + n) * (nu - d - 1))
    assert (np.abs(mus.mean(0) - mu_star) < 5 * np.sqrt(np.diag(cov_mu) / K)).all()
    assert np.abs(np.diag(np.cov(mus, rowvar=0)) / np.diag(cov_mu) - 1).max() < 0.2
    # all eigenvalues of every draw are positive
    assert np.linalg.eigvalsh(alphas).min() > 0


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_chain_driver_equals_the_host_driven_loop(Bm, golden, dtype, monkeypatch):
    """amf_gibbs_chain_device (16 samples per call) and the same chain driven kernel group by
    kernel group from Python use the same Philox counters: equal draw for draw"""
    monkeypatch.setenv("AMF_B200_DTYPE", dtype)
    g = golden("gibbs_15x12_d3")
    chains = []
    for chunk in (16, 1, 5):
        b = Bm.BayesianPMF(g["ratings"], 3, subtract_mean=True)
        b.compute_dtype = dtype
        b.users, b.items = g["users"].copy(), g["items"].copy()
        b.device_seed, b.device_chunk = 9, chunk
        chains.append([(u.cpu().numpy().copy(), v.cpu().numpy().copy())
                       for u, v in islice(b.samples_device(num_gibbs=2), 21)])
    for other in chains[1:]:
        for (u0, v0), (u1, v1) in zip(chains[0], other):
            np.testing.assert_array_equal(u0, u1)
            np.testing.assert_array_equal(v0, v1)
    assert np.isfinite(chains[0][-1][0]).all()


def test_fast_chain_with_host_hyperparameters_still_available(Bm, golden):
    g = golden("gibbs_15x12_d3")
    b = Bm.BayesianPMF(g["ratings"], 3, subtract_mean=True)
    b.users, b.items = g["users"].copy(), g["items"].copy()
    np.random.seed(3)
    us, vs = next(b.samples_device(num_gibbs=1, hyper='host'))
    assert us.shape == (15, 3) and vs.shape == (12, 3) and bool(us.isfinite().all())
    with pytest.raises(ValueError):
        next(b.samples_device(hyper='nowhere'))


# ---- batched lookahead chains (fast mode) ------------------------------------------------------
@pytest.mark.parametrize("dtype,tol", [("f64", 1e-9), ("f32", 2e-4)])
def test_batched_half_sweep_is_each_chains_own_conditional(dtype, tol):
    """amf_gibbs_half_sweep_batched: chain p = the common rating list + its one extra rating, its
    own factors, hyper-parameters and mean offset; every row equals the conditional of THAT model
    with the normals of its own Philox counters (bayes_pmf.py:560-598 without the copies)"""
    import torch
    from active_matrix_factorization_b200 import _native as N, device as D
    lib = N.require_device()
    rng = np.random.RandomState(3)
    n, m, d, P = 12, 9, 4, 5
    cells = rng.permutation(n * m)[:50]
    R = np.column_stack((cells // m, cells % m, rng.normal(3, 1, 50)))
    free = [c for c in range(n * m) if c not in set(cells.tolist())]
    ex = rng.permutation(free)[:P]
    ex_i, ex_j, ex_v = (ex // m).astype(np.int32), (ex % m).astype(np.int32), rng.normal(3, 1, P)
    others = rng.normal(0, .5, (P, m, d))
    a0 = rng.normal(0, 1, (P, d, d))
    alphas = a0 @ a0.transpose(0, 2, 1) / d + np.eye(d)
    mus, offs, beta = rng.normal(0, .3, (P, d)), rng.normal(3, .1, P), 2.0
    rat = D.Ratings.from_tuples(R, n, m, dtype)
    dt = D.np_dtype(dtype)
    t = [D.to_device(x, dt) for x in (others, alphas, mus)]
    ti, tj, tv, to = (D.to_device(ex_i, np.int32), D.to_device(ex_j, np.int32),
                      D.to_device(ex_v, np.float64), D.to_device(offs, np.float64))
    out = torch.empty((P, n, d), dtype=D.torch_dtype(dtype), device="cuda")
    seed, stream = 1234, 77
    N.check(lib.amf_gibbs_half_sweep_batched(rat.handle, 0, D.code(dtype), d, P, D.ptr(t[0]), D.ptr(t[1]),
                                             D.ptr(t[2]), beta, 0.0, D.ptr(to), D.ptr(ti), D.ptr(tj),
                                             D.ptr(tv), seed, stream, D.ptr(out), D.stream_ptr()))
    got = out.double().cpu().numpy()
    for p in range(P):
        Rp = np.vstack((R, [ex_i[p], ex_j[p], ex_v[p]]))
        for i in range(n):
            rows = Rp[Rp[:, 0] == i]
            F = others[p][rows[:, 1].astype(int)]
            lam = alphas[p] + beta * F.T @ F
            rhs = beta * F.T @ (rows[:, 2] - offs[p]) + alphas[p] @ mus[p]
            z = np.empty(d)
            for k in range(d):
                w = _philox4x32_10([i, k + 32 * p, stream & 0xffffffff, stream >> 32],
                                   [seed & 0xffffffff, seed >> 32])
                z[k] = np.sqrt(-2 * np.log((w[0] + 1.0) * 2.0 ** -32)) * np.cos(2 * np.pi * (w[1] + 1.0) * 2.0 ** -32)
            Rf = np.linalg.cholesky(lam)
            want = np.linalg.solve(Rf.T, np.linalg.solve(Rf, rhs) + z)
            assert np.abs(got[p, i] - want).max() <= tol * max(1.0, np.abs(want).max()), (p, i)
    rat.close()


def test_total_variance_from_gram_matrices(Bm):
    """sum over all cells of the sample variance, reduced from d x d Gram blocks of the stacked
    samples, against bayes_pmf.py:440-451 computed directly"""
    import torch
    rng = np.random.RandomState(1)
    P, n, m, d, S = 3, 17, 11, 4, 6
    us, vs = rng.normal(size=(P, S, n, d)), rng.normal(1, .5, size=(P, S, m, d))
    ku = torch.tensor(us.transpose(0, 2, 1, 3).reshape(P, n, S * d), device="cuda")
    kv = torch.tensor(vs.transpose(0, 2, 1, 3).reshape(P, m, S * d), device="cuda")
    got = Bm._total_variance_of_stacks(ku, kv, S, d).cpu().numpy()
    want = [np.einsum("snd,smd->snm", us[p], vs[p]).var(0).sum() for p in range(P)]
    np.testing.assert_allclose(got, want, rtol=1e-11)


def test_exp_variance_batched_fast_mode_in_law(Bm, golden):
    """exp_variance in fast mode (all (cell, value) chains in the same launches, device random
    numbers) against the host path that replays the reference draw for draw: same expectation
    within Monte-Carlo error; known cells give NaN"""
    g = golden("gibbs_15x12_d3")
    kw = dict(rating_values=(1, 2, 3, 4, 5), discrete_expectations=True)
    b = Bm.BayesianPMF(g["ratings"], 3, **kw)
    b.users, b.items = g["users"].copy(), g["items"].copy()
    np.random.seed(3)
    samples = list(islice(b.samples(num_gibbs=2), 40))
    cand = sorted(b.unrated)[:3]
    which = tuple(np.array(cand).T)
    S = 250
    np.random.seed(4)
    host = b.exp_variance(samples, which=which, num_samps=S, fit_first=False)
    b.rng_mode = 'device'
    np.random.seed(5)
    fast = b.exp_variance(samples, which=which, num_samps=S, fit_first=False)
    assert fast.shape == host.shape == (3,)
    assert np.all(np.isfinite(fast)) and np.all(fast > 0)
    assert np.abs(fast / host - 1).max() < 0.12, (fast, host)
    known = tuple(np.array([tuple(map(int, g["ratings"][0, :2]))]).T)
    with pytest.warns(UserWarning):
        assert np.isnan(b.exp_variance(samples, which=known, num_samps=5)[0])
    # continuous R_ij: normal fit, trapezoid over the ppf points
    c = Bm.BayesianPMF(g["ratings"], 3, rating_values=None, discrete_expectations=False, num_integration_pts=7)
    c.users, c.items = g["users"].copy(), g["items"].copy()
    c.rng_mode = 'device'
    cont = c.exp_variance(samples, which=which, num_samps=20)
    assert cont.shape == (3,) and np.all(np.isfinite(cont)) and np.all(cont > 0)
