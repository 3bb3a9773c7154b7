import sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from active_matrix_factorization_b200 import bayes_pmf as Bm, _native as N, device as D
rng = np.random.RandomState(0)
n, m, d = 943, 1682, 15
cells = rng.permutation(n * m)[:5000]
R = np.column_stack((cells // m, cells % m, rng.randint(1, 6, 5000))).astype(float)
for name in ("f64", "f32"):
    b = Bm.BayesianPMF(R, d, subtract_mean=True); b.compute_dtype = name
    lib = N.require_device()
    rat = D.Ratings.from_tuples(b.ratings, n, m, name)
    dt = D.np_dtype(name); tdt = D.torch_dtype(name)
    users_t, items_t = D.to_device(b.users, dt), D.to_device(b.items, dt)
    priors = []
    for wi, b0, df, mu0 in (b.u_hyperparams, b.v_hyperparams):
        priors.append(D.to_device(np.concatenate((np.linalg.inv(wi).reshape(-1), mu0.astype(float), [float(b0), float(df)])), np.float64))
    chunk = 16
    us = torch.empty((chunk, n, d), dtype=tdt, device="cuda"); vs = torch.empty((chunk, m, d), dtype=tdt, device="cuda")
    def call(sid):
        N.check(lib.amf_gibbs_chain_device(rat.handle, D.code(name), d, chunk, 2, D.ptr(users_t), D.ptr(items_t),
                D.ptr(priors[0]), D.ptr(priors[1]), 2.0, 0.0, 5, sid, D.ptr(us), D.ptr(vs), D.stream_ptr()))
    call(0); torch.cuda.synchronize()
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); call(100 * (rep + 1)); e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(name, "16 samples: enqueue %.2f ms, until done %.2f ms, device %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t0), e0.elapsed_time(e1)))

# the same through the class generator, chunk by chunk
from itertools import islice
b = Bm.BayesianPMF(R, d, subtract_mean=True)
gen = b.samples_device(num_gibbs=2)
next(gen); torch.cuda.synchronize()
for rep in range(5):
    t0 = time.perf_counter()
    out = list(islice(gen, 16))
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("generator, 16 samples: python %.2f ms, until done %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t0)))
t0 = time.perf_counter(); out = list(islice(b.samples_device(num_gibbs=2), 192)); torch.cuda.synchronize()
print("fresh generator, 192 samples: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
t0 = time.perf_counter(); rat = D.Ratings.from_tuples(b.ratings, n, m, "f64"); torch.cuda.synchronize()
print("Ratings.from_tuples: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
