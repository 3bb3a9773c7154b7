import time, sys, json
import numpy as np
sys.path.insert(0, '.')
import torch
from active_matrix_factorization_b200 import mn_active_pmf as M
np.random.seed(0)
n, m, d = 94, 425, 5
tu, tv = np.random.normal(0, 1, (n, d)), np.random.normal(0, 1, (m, d))
real = np.where(tu @ tv.T > 1.5, 1., -1.)
cells = np.random.permutation(n * m)
known = list(cells[:500])
ii, jj = np.array(known) // m, np.array(known) % m
# every row/col at least once
for i in range(n):
    if i not in ii: ii = np.append(ii, i); jj = np.append(jj, np.random.randint(m))
for j in range(m):
    if j not in jj: jj = np.append(jj, j); ii = np.append(ii, np.random.randint(n))
R = np.unique(np.column_stack((ii, jj)), axis=0)
R = np.column_stack((R, real[R[:, 0], R[:, 1]])).astype(float)
a = M.MNActivePMF(R, d, rating_values={-1, 1}, discrete_expectations=True)
t0 = time.time(); a.fit(); t_fit = time.time() - t0
a.initialize_approx()
a.max_normal_steps = 200          # bounded: the reference needs ~0.38 s per accepted step here
t0 = time.time(); kls = list(a.fit_normal_kls()); torch.cuda.synchronize(); t_normal = time.time() - t0
pool = sorted(a.unrated)
t0 = time.time(); pv = a._get_key_vals(pool, M.MNActivePMF.pred_variance, None, None); t_pv = time.time() - t0
t0 = time.time(); pv = a._get_key_vals(pool, M.MNActivePMF.pred_variance, None, None); t_pv2 = time.time() - t0
print(json.dumps(dict(nnz=len(R), pool=len(pool), map_fit_s=t_fit, fit_normal_steps=len(kls),
                      fit_normal_s=t_normal, s_per_step=t_normal / max(1, len(kls)), kl_last=kls[-1],
                      pred_variance_all_s=[t_pv, t_pv2], cand_per_s=len(pool) / t_pv2)))
