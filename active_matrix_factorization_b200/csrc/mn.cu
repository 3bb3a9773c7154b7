// Matrix-normal approximation of the PMF posterior: the variant the reference
// runs on its drugbank / movielens experiments (mn_active_pmf.py, matrix_normal_exps_cy.pyx).
//
//   X = [U; V] ~ MN(mean, Sigma, Omega),  Cov(X_ak, X_bl) = Sigma[a,b] * Omega[k,l]
//   mean (N+M, d), Sigma = cov_useritems (N+M, N+M), Omega = cov_latents (d, d)
//
//   mn_kl_divergence            matrix_normal_exps_cy.pyx:159-216
//   matrixnormal_gradient       matrix_normal_exps_cy.pyx:219-485
//   fit_normal_kls              mn_active_pmf.py:242-288 (project_psd :42-67 on Sigma and Omega)
//   _approx_entropy             mn_active_pmf.py:513-521
//   approx_pred_mean_var        mn_active_pmf.py:300-315 (+ exp_dotprod_sq pyx:126-154)
//
// Same execution model as normal.cu: one CTA per (candidate, value) problem runs the whole
// line search on the device.  Per rating the gradient touches only Sigma[i,i], Sigma[j,j],
// Sigma[i,j] and the d x d Omega, so the rating loop is O(nnz d^2) and the dense work is the two
// eigendecompositions per trial.  fp64.
//
// Reference quirks reproduced in the KL (pyx:176,192,197): the item trace term is dropped
// (num_items evaluates to 0) and both prior terms use sigma_u_sq.
#include "common.cuh"
#include "dense_blk.cuh"

namespace amf {

constexpr int MN_THREADS = 256;

struct MnProblem {
  int n, m, d, nui;
  int64_t nnz;
  const int32_t* ri; const int32_t* rj; const double* rr;
  int ei, ej; double er;
  double sigma_sq, sigma_u_sq, sigma_v_sq;
};

__device__ __forceinline__ void mn_rating(const MnProblem& P, int64_t t, int& i, int& j, double& r) {
  if (t < P.nnz) { i = P.ri[t]; j = P.rj[t]; r = P.rr[t]; }
  else { i = P.ei; j = P.ej; r = P.er; }
}

// E[(U_i . V_j)^2] restricted to the terms owned by latent index k (k and all l > k)
__device__ __forceinline__ double mn_e2_k(const double* mu, const double* mv, const double* om,
                                          int d, int k, double sii, double sjj, double sij) {
  const double okk = om[k * d + k];
  const double cab = sij * okk;
  double t = 4 * mu[k] * mv[k] * cab + 2 * cab * cab +
             (mu[k] * mu[k] + sii * okk) * (mv[k] * mv[k] + sjj * okk);
  for (int l = k + 1; l < d; ++l) {
    const double okl = om[k * d + l], oll = om[l * d + l];
    const double c_ab = sij * okk, c_ac = sii * okl, c_ad = sij * okl;
    const double c_bc = sij * okl, c_bd = sjj * okl, c_cd = sij * oll;
    t += 2 * (mu[k] * mv[k] * mu[l] * mv[l] + mu[k] * mv[k] * c_cd + mu[k] * mu[l] * c_bd +
              mu[k] * mv[l] * c_bc + mv[k] * mu[l] * c_ad + mv[k] * mv[l] * c_ac +
              mu[l] * mv[l] * c_ab + c_ab * c_cd + c_ac * c_bd + c_ad * c_bc);
  }
  return t;
}

__device__ double mn_kl(const MnProblem& P, const double* mean, const double* sig,
                        const double* om, double* wsig, double* wom, double* red, int* flag,
                        bool dense = true) {
  const int d = P.d, nui = P.nui, nu = P.n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  // entropy term (skipped when the caller does the dense algebra itself)
  double kl = 0;
  if (dense) {
    for (int t = tid; t < nui * nui; t += nt) wsig[t] = sig[t];
    for (int t = tid; t < d * d; t += nt) wom[t] = om[t];
    __syncthreads();
    const double ld_sig = blk_cholesky(wsig, nui, red, flag);
    const double ld_om = blk_cholesky(wom, d, red, flag);
    kl = -(ld_sig * d + ld_om * nui) / 2.;
  }
  // prior terms (quirks: no item trace, sigma_u_sq twice)
  double tr_om = 0;
  for (int k = 0; k < d; ++k) tr_om += om[k * d + k];
  double acc = 0;
  for (int t = tid; t < nui * d; t += nt) acc += mean[t] * mean[t];
  for (int t = tid; t < nu; t += nt) acc += sig[(int64_t)t * nui + t] * tr_om;
  kl += blk_sum(acc, red) / (2 * P.sigma_u_sq);
  // rating terms
  const int64_t total = P.nnz + (P.ei >= 0 ? 1 : 0);
  acc = 0;
  for (int64_t t = warp; t < total; t += nwarps) {
    int i, j; double r;
    mn_rating(P, t, i, j, r);
    const int j_ = nu + j;
    const double* mu = mean + (int64_t)i * d;
    const double* mv = mean + (int64_t)j_ * d;
    const double sii = sig[(int64_t)i * nui + i], sjj = sig[(int64_t)j_ * nui + j_];
    const double sij = sig[(int64_t)i * nui + j_];
    for (int k = lane; k < d; k += 32)
      acc += mn_e2_k(mu, mv, om, d, k, sii, sjj, sij) - 2 * r * (mu[k] * mv[k] + sij * om[k * d + k]);
    if (lane == 0) acc += r * r;
  }
  kl += blk_sum(acc, red) / (2 * P.sigma_sq);
  return kl;
}

__device__ void mn_grad(const MnProblem& P, const double* mean, const double* sig,
                        const double* om, double* gm, double* gs, double* go, double* w1,
                        double* w2, double* wo1, double* wo2, double* red, int* flag,
                        bool dense = true) {
  const int d = P.d, nui = P.nui, nu = P.n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  for (int t = tid; t < nui * d; t += nt) gm[t] = 0;
  for (int t = tid; t < nui * nui; t += nt) gs[t] = 0;
  for (int t = tid; t < d * d; t += nt) go[t] = 0;
  __syncthreads();
  const int64_t total = P.nnz + (P.ei >= 0 ? 1 : 0);
  const double inv_s = 1.0 / P.sigma_sq;
  for (int64_t t = warp; t < total; t += nwarps) {
    int i, j; double rating;
    mn_rating(P, t, i, j, rating);
    const int j_ = nu + j;
    const double* mi = mean + (int64_t)i * d;
    const double* mj = mean + (int64_t)j_ * d;
    const double sii = sig[(int64_t)i * nui + i], sjj = sig[(int64_t)j_ * nui + j_];
    const double sij = sig[(int64_t)i * nui + j_];
    double g_ii = 0, g_jj = 0, g_ij = 0;            // this lane's share of the three Sigma entries
    for (int k = lane; k < d; k += 32) {
      const double Mik = mi[k], Mjk = mj[k], vk = om[k * d + k];
      double gmik = 0, gmjk = 0, gokk = 0;
      for (int l = k + 1; l < d; ++l) {             // _quadexp_grad, mult = 1/sigma^2
        const double Mil = mi[l], Mjl = mj[l], ckl = om[k * d + l], vl = om[l * d + l];
        gmik += inv_s * (Mjk * Mil * Mjl + Mjl * sij * ckl + Mil * sjj * ckl + Mjk * sij * vl);
        atomicAdd(gm + (int64_t)i * d + l,
                  inv_s * (Mik * Mjk * Mjl + Mjl * sij * vk + Mjk * sij * ckl + Mik * sjj * ckl));
        gmjk += inv_s * (Mik * Mil * Mjl + Mjl * sii * ckl + Mil * sij * ckl + Mik * sij * vl);
        atomicAdd(gm + (int64_t)j_ * d + l,
                  inv_s * (Mik * Mjk * Mil + Mil * sij * vk + Mjk * sii * ckl + Mik * sij * ckl));
        g_ii += inv_s * (Mjk * Mjl * ckl + sjj * ckl * ckl);
        g_jj += inv_s * (Mik * Mil * ckl + sii * ckl * ckl);
        g_ij += inv_s * (Mil * Mjl * vk + Mjk * Mil * ckl + Mik * Mjl * ckl + Mik * Mjk * vl +
                         2 * sij * vk * vl + 2 * sij * ckl * ckl);
        gokk += inv_s * (Mil * Mjl * sij + sij * sij * vl);
        atomicAdd(go + l * d + l, inv_s * (Mik * Mjk * sij + sij * sij * vk));
        const double inc = inv_s * (Mjk * Mjl * sii + Mjk * Mil * sij + Mik * Mjl * sij +
                                    Mik * Mil * sjj + 2 * sii * sjj * ckl + 2 * sij * sij * ckl);
        atomicAdd(go + k * d + l, inc);
        atomicAdd(go + l * d + k, inc);
      }
      // _squareexp_grad, mult = 1/(2 sigma^2)
      const double h = inv_s / 2;
      const double e_ik = Mik * Mik + sii * vk, e_jk = Mjk * Mjk + sjj * vk;
      gmik += h * (4 * Mjk * sij * vk + 2 * Mik * e_jk);
      gmjk += h * (4 * Mik * sij * vk + e_ik * 2 * Mjk);
      g_ii += h * (vk * e_jk);
      g_jj += h * (e_ik * vk);
      g_ij += h * (4 * (Mik * Mjk + sij * vk) * vk);
      gokk += h * (4 * Mik * Mjk * sij + 4 * sij * sij * vk + sii * e_jk + e_ik * sjj);
      // -R_ij E[U_ik V_jk] / sigma^2
      const double mr = -rating * inv_s;
      gmik += mr * Mjk;
      gmjk += mr * Mik;
      g_ij += mr * vk;
      gokk += mr * sij;
      atomicAdd(gm + (int64_t)i * d + k, gmik);
      atomicAdd(gm + (int64_t)j_ * d + k, gmjk);
      atomicAdd(go + k * d + k, gokk);
    }
    g_ii = warp_sum(g_ii); g_jj = warp_sum(g_jj); g_ij = warp_sum(g_ij);
    if (lane == 0) {
      atomicAdd(gs + (int64_t)i * nui + i, g_ii);
      atomicAdd(gs + (int64_t)j_ * nui + j_, g_jj);
      atomicAdd(gs + (int64_t)i * nui + j_, g_ij);
      atomicAdd(gs + (int64_t)j_ * nui + i, g_ij);
    }
  }
  __syncthreads();
  // priors (pyx:442-458)
  double tr_om = 0;
  for (int k = 0; k < d; ++k) tr_om += om[k * d + k];
  double su = 0, sv = 0;
  for (int t = tid; t < nui; t += nt) {
    const double v = sig[(int64_t)t * nui + t];
    if (t < nu) su += v; else sv += v;
  }
  su = blk_sum(su, red);
  sv = blk_sum(sv, red);
  for (int t = tid; t < nui * d; t += nt)
    gm[t] += mean[t] / ((t / d) < nu ? P.sigma_u_sq : P.sigma_v_sq);
  for (int t = tid; t < nui; t += nt)
    gs[(int64_t)t * nui + t] += tr_om / (2 * (t < nu ? P.sigma_u_sq : P.sigma_v_sq));
  for (int t = tid; t < d; t += nt)
    go[t * d + t] += su / (2 * P.sigma_u_sq) + sv / (2 * P.sigma_v_sq);
  if (!dense) { __syncthreads(); return; }
  // entropy terms: g -= scale/2 * (inv + inv' o (1 - I))   (pyx:475-485)
  for (int t = tid; t < nui * nui; t += nt) w1[t] = sig[t];
  for (int t = tid; t < d * d; t += nt) wo1[t] = om[t];
  __syncthreads();
  blk_cholesky(w1, nui, red, flag);
  blk_tri_inverse(w1, w2, nui);
  for (int t = tid; t < nui * nui; t += nt) {
    const int r = t / nui, c = t % nui;
    double s = 0;
    for (int q = max(r, c); q < nui; ++q) s += w2[(int64_t)q * nui + r] * w2[(int64_t)q * nui + c];
    gs[t] -= (d / 2.) * ((r == c) ? s : 2 * s);
  }
  blk_cholesky(wo1, d, red, flag);
  blk_tri_inverse(wo1, wo2, d);
  for (int t = tid; t < d * d; t += nt) {
    const int r = t / d, c = t % d;
    double s = 0;
    for (int q = max(r, c); q < d; ++q) s += wo2[q * d + r] * wo2[q * d + c];
    go[t] -= (nui / 2.) * ((r == c) ? s : 2 * s);
  }
  __syncthreads();
}

// sum_ij Var[U_i . V_j]  (mn_active_pmf.py:317-330, :597-598), E[x^2] - E[x]^2 like the reference
__device__ double mn_total_variance(const MnProblem& P, const double* mean, const double* sig,
                                    const double* om, double* red) {
  const int d = P.d, nui = P.nui, nu = P.n;
  double tr_om = 0;
  for (int k = 0; k < d; ++k) tr_om += om[k * d + k];
  double acc = 0;
  for (int cell = threadIdx.x; cell < P.n * P.m; cell += blockDim.x) {
    const int i = cell / P.m, j_ = nu + cell % P.m;
    const double* mu = mean + (int64_t)i * d;
    const double* mv = mean + (int64_t)j_ * d;
    const double sii = sig[(int64_t)i * nui + i], sjj = sig[(int64_t)j_ * nui + j_];
    const double sij = sig[(int64_t)i * nui + j_];
    double e = sij * tr_om, e2 = 0;
    for (int k = 0; k < d; ++k) { e += mu[k] * mv[k]; e2 += mn_e2_k(mu, mv, om, d, k, sii, sjj, sij); }
    acc += e2 - e * e;
  }
  return blk_sum(acc, red);
}

struct MnArgs {
  int B, n, m, d;
  int64_t nnz;
  const int32_t* ri; const int32_t* rj; const double* rr;
  const int32_t* ei; const int32_t* ej; const double* er;
  double sigma_sq, sigma_u_sq, sigma_v_sq, lr0, min_eig, kl_stop, min_lr;
  int max_steps;
  double *mean, *sig, *om, *work, *kl_out;
  int* steps_out;
  double* kl_trace; int trace_len;
  double *entropy_out, *totvar_out;
  int mode;
  bool smem_state;
};

__host__ __device__ inline int64_t mn_workspace(int64_t nui, int64_t d) {
  return 2 * nui * d + 5 * nui * nui + 5 * d * d;
}

__global__ void __launch_bounds__(MN_THREADS) mn_fit_kernel(MnArgs a) {
  extern __shared__ double sh[];
  double* red = sh;
  double* cs = sh + 32;
  __shared__ int flag;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    MnProblem P;
    P.n = a.n; P.m = a.m; P.d = a.d; P.nui = a.n + a.m; P.nnz = a.nnz;
    P.ri = a.ri; P.rj = a.rj; P.rr = a.rr;
    P.ei = a.ei ? a.ei[b] : -1; P.ej = a.ej ? a.ej[b] : -1; P.er = a.er ? a.er[b] : 0.0;
    P.sigma_sq = a.sigma_sq; P.sigma_u_sq = a.sigma_u_sq; P.sigma_v_sq = a.sigma_v_sq;
    const int nui = P.nui, d = P.d;
    const int64_t n2 = (int64_t)nui * nui, d2 = (int64_t)d * d, nd = (int64_t)nui * d;
    double* mean_g = a.mean + b * nd;
    double* sig_g = a.sig + b * n2;
    double* om_g = a.om + b * d2;
    // small problems keep the state and all work matrices in shared memory (see normal.cu)
    const bool in_smem = a.smem_state;
    const int64_t kmax_ = nui > d ? nui : d;
    double* S = sh + blk_scratch_doubles(kmax_);
    double* mean = in_smem ? S : mean_g;
    double* sig = in_smem ? S + nd : sig_g;
    double* om = in_smem ? S + nd + n2 : om_g;
    double* W = in_smem ? S + nd + n2 + d2 : a.work + b * mn_workspace(nui, d);
    if (in_smem) {
      for (int t = tid; t < nd; t += nt) mean[t] = mean_g[t];
      for (int64_t t = tid; t < n2; t += nt) sig[t] = sig_g[t];
      for (int t = tid; t < d2; t += nt) om[t] = om_g[t];
      __syncthreads();
    }
    double* gm = W;  double* nmean = W + nd;
    double* gs = W + 2 * nd;  double* nsig = gs + n2;
    double* w1 = nsig + n2;  double* w2 = w1 + n2;  double* w3 = w2 + n2;
    double* go = w3 + n2;  double* nom = go + d2;
    double* wo1 = nom + d2;  double* wo2 = wo1 + d2;  double* wo3 = wo2 + d2;

    if (a.mode == 1 || a.mode == 4) {
      const double kl = mn_kl(P, mean, sig, om, w1, wo1, red, &flag, a.mode == 1);
      if (tid == 0) a.kl_out[b] = kl;
      continue;
    }
    if (a.mode == 2 || a.mode == 5) {
      mn_grad(P, mean, sig, om, gm, gs, go, w1, w2, wo1, wo2, red, &flag, a.mode == 2);
      if (in_smem) {            // the caller reads the gradient from the global workspace
        double* Wg = a.work + b * mn_workspace(nui, d);
        for (int t = tid; t < nd; t += nt) Wg[t] = gm[t];
        for (int64_t t = tid; t < n2; t += nt) Wg[2 * nd + t] = gs[t];
        for (int t = tid; t < d2; t += nt) Wg[2 * nd + 5 * n2 + t] = go[t];
        __syncthreads();
      }
      continue;
    }
    double lr = a.lr0;
    double old_kl = mn_kl(P, mean, sig, om, w1, wo1, red, &flag);
    int steps = 0;
    bool converged = false;
    while (!converged) {
      mn_grad(P, mean, sig, om, gm, gs, go, w1, w2, wo1, wo2, red, &flag);
      while (true) {
        for (int t = tid; t < nd; t += nt) nmean[t] = mean[t] - lr * gm[t];
        for (int64_t t = tid; t < n2; t += nt) nsig[t] = sig[t] - lr * gs[t];
        for (int t = tid; t < d2; t += nt) nom[t] = om[t] - lr * go[t];
        __syncthreads();
        blk_project_psd(nsig, nui, a.min_eig, w1, w2, cs, red);
        blk_project_psd(nom, d, a.min_eig, wo1, wo2, cs, red);
        const double new_kl = mn_kl(P, nmean, nsig, nom, w3, wo3, red, &flag);
        if (new_kl < old_kl) {
          for (int t = tid; t < nd; t += nt) mean[t] = nmean[t];
          for (int64_t t = tid; t < n2; t += nt) sig[t] = nsig[t];
          for (int t = tid; t < d2; t += nt) om[t] = nom[t];
          __syncthreads();
          lr *= 1.25;
          if (old_kl - new_kl < a.kl_stop) converged = true;
          if (a.kl_trace && steps < a.trace_len && tid == 0)
            a.kl_trace[(int64_t)b * a.trace_len + steps] = new_kl;
          old_kl = new_kl;
          ++steps;
          break;
        } else {
          lr *= 0.5;
          if (lr < a.min_lr) { converged = true; break; }
        }
      }
      if (a.max_steps > 0 && steps >= a.max_steps) break;
    }
    if (tid == 0) { a.kl_out[b] = old_kl; a.steps_out[b] = steps; }
    if (in_smem) {
      for (int t = tid; t < nd; t += nt) mean_g[t] = mean[t];
      for (int64_t t = tid; t < n2; t += nt) sig_g[t] = sig[t];
      for (int t = tid; t < d2; t += nt) om_g[t] = om[t];
    }
    if (a.entropy_out) {
      for (int64_t t = tid; t < n2; t += nt) w1[t] = sig[t];
      for (int t = tid; t < d2; t += nt) wo1[t] = om[t];
      __syncthreads();
      const double ls = blk_cholesky(w1, nui, red, &flag);
      const double lo = blk_cholesky(wo1, d, red, &flag);
      if (tid == 0) a.entropy_out[b] = 0.5 * (d * ls + nui * lo);
    }
    if (a.totvar_out) {
      const double tv = mn_total_variance(P, mean, sig, om, red);
      if (tid == 0) a.totvar_out[b] = tv;
    }
    __syncthreads();
  }
}

// criteria over a candidate pool under MN(mean, Sigma, Omega): one thread per candidate
template <typename T, int CRIT, bool MAX>
__global__ void __launch_bounds__(128)
mn_score_kernel(const int32_t* __restrict__ ci, const int32_t* __restrict__ cj, int64_t ncand,
                int n, int nui, int d, const T* __restrict__ mean, const T* __restrict__ sig,
                const T* __restrict__ om, T cutoff, T* __restrict__ scores, int64_t index_base,
                Best* __restrict__ part) {
  Best best{0.0, -1};
  T tr_om = 0, tr_om2 = 0;
  for (int k = 0; k < d; ++k) {
    tr_om += om[k * d + k];
    for (int l = 0; l < d; ++l) tr_om2 = fma(om[k * d + l], om[l * d + k], tr_om2);
  }
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncand;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int i = ci[c], j_ = n + cj[c];
    const T* mu = mean + (int64_t)i * d;
    const T* mv = mean + (int64_t)j_ * d;
    const T sii = sig[(int64_t)i * nui + i], sjj = sig[(int64_t)j_ * nui + j_];
    const T sij = sig[(int64_t)i * nui + j_];
    T e = sij * tr_om, quu = 0, qvv = 0, quv = 0;
    for (int k = 0; k < d; ++k) {
      e = fma(mu[k], mv[k], e);
      if (CRIT != AMF_CRIT_APPROX_MEAN) {
        T ou = 0, ov = 0;
        for (int l = 0; l < d; ++l) { ou = fma(om[k * d + l], mu[l], ou); ov = fma(om[k * d + l], mv[l], ov); }
        quu = fma(mu[k], ou, quu); qvv = fma(mv[k], ov, qvv); quv = fma(mu[k], ov, quv);
      }
    }
    // Var = (Sii Sjj + Sij^2) tr(Om^2) + Sii mv'Om mv + Sjj mu'Om mu + 2 Sij mu'Om mv
    const T var = (sii * sjj + sij * sij) * tr_om2 + sii * qvv + sjj * quu + 2 * sij * quv;
    T out;
    if (CRIT == AMF_CRIT_APPROX_MEAN) out = e;
    else if (CRIT == AMF_CRIT_PRED_VARIANCE) out = var;
    else out = var > 0 ? T(0.5) * erfc((cutoff - e) / (var * T(1.4142135623730951))) : T(NAN);
    if (scores) scores[c] = out;
    if (better<MAX>((double)out, c + index_base, best.v, best.i)) { best.v = (double)out; best.i = c + index_base; }
  }
  best = block_best<MAX>(best);
  if (threadIdx.x == 0) part[blockIdx.x] = best;
}

int acquire_partials(Best** out, cudaStream_t s);
int launch_best_final(const Best* part_d, int nparts, bool maximize, amf_best_t* out_d,
                      cudaStream_t s);

template <typename T, bool MAX>
static int mn_score(int crit, int64_t ncand, const int32_t* ci, const int32_t* cj, int n, int m,
                    int d, const T* mean, const T* sig, const T* om, double cutoff, T* scores,
                    int64_t index_base, Best* part, int grid, cudaStream_t s) {
#define MNS(C) mn_score_kernel<T, C, MAX><<<grid, 128, 0, s>>>(ci, cj, ncand, n, n + m, d, mean, sig, \
                                                               om, (T)cutoff, scores, index_base, part)
  if (crit == AMF_CRIT_APPROX_MEAN) MNS(AMF_CRIT_APPROX_MEAN);
  else if (crit == AMF_CRIT_PRED_VARIANCE) MNS(AMF_CRIT_PRED_VARIANCE);
  else MNS(AMF_CRIT_PROB_GE);
#undef MNS
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int64_t amf_mn_workspace_doubles(int32_t n, int32_t m, int d) {
  return mn_workspace((int64_t)n + m, d);
}

int amf_mn_batched(int mode, int B, int64_t nnz, const int32_t* ri_d, const int32_t* rj_d,
                   const double* rr_d, const int32_t* extra_i_d, const int32_t* extra_j_d,
                   const double* extra_r_d, const amf_normal_fit_params_t* p, double* mean_d,
                   double* sig_d, double* om_d, double* work_d, double* kl_out_d,
                   int32_t* steps_out_d, double* kl_trace_d, int trace_len, double* entropy_out_d,
                   double* totvar_out_d, void* stream) {
  AMF_REQUIRE(p && mean_d && sig_d && om_d && work_d, "amf_mn_batched: NULL argument");
  AMF_REQUIRE((mode >= 0 && mode <= 2) || mode == 4 || mode == 5, "amf_mn_batched: bad mode %d", mode);
  AMF_REQUIRE(B >= 0 && nnz >= 0 && p->n > 0 && p->m > 0 && p->d > 0, "amf_mn_batched: bad sizes");
  AMF_REQUIRE(mode != 0 || (kl_out_d && steps_out_d), "amf_mn_batched: fit needs kl_out/steps_out");
  AMF_REQUIRE((mode != 1 && mode != 4) || kl_out_d, "amf_mn_batched: kl mode needs kl_out");
  if (B == 0) return AMF_OK;
  const int64_t nui = (int64_t)p->n + p->m;
  AMF_REQUIRE(nui * nui < (1ll << 31), "amf_mn_batched: N+M=%lld too large", (long long)nui);
  MnArgs a{};
  a.B = B; a.n = p->n; a.m = p->m; a.d = p->d; a.nnz = nnz;
  a.ri = ri_d; a.rj = rj_d; a.rr = rr_d; a.ei = extra_i_d; a.ej = extra_j_d; a.er = extra_r_d;
  a.sigma_sq = p->sigma_sq; a.sigma_u_sq = p->sigma_u_sq; a.sigma_v_sq = p->sigma_v_sq;
  a.lr0 = p->learning_rate; a.min_eig = p->min_eig; a.kl_stop = p->kl_stop; a.min_lr = p->min_lr;
  a.max_steps = p->max_steps;
  a.mean = mean_d; a.sig = sig_d; a.om = om_d; a.work = work_d; a.kl_out = kl_out_d;
  a.steps_out = steps_out_d; a.kl_trace = kl_trace_d; a.trace_len = trace_len;
  a.entropy_out = entropy_out_d; a.totvar_out = totvar_out_d; a.mode = mode;
  const int64_t kmax = nui > p->d ? nui : p->d;
  size_t smem = sizeof(double) * (size_t)blk_scratch_doubles(kmax);
  const size_t state = sizeof(double) * (size_t)(nui * p->d + nui * nui + (int64_t)p->d * p->d +
                                                mn_workspace(nui, p->d));
  a.smem_state = smem + state <= 200 * 1024;
  if (a.smem_state) smem += state;
  AMF_CUDA(cudaFuncSetAttribute(mn_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
  mn_fit_kernel<<<B, MN_THREADS, smem, (cudaStream_t)stream>>>(a);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

int amf_mn_score_candidates(int criterion, int dtype, int64_t ncand, const int32_t* ci_d,
                            const int32_t* cj_d, int32_t n, int32_t m, int d, const void* mean_d,
                            const void* sig_d, const void* om_d, double cutoff, void* scores_d,
                            int maximize, int64_t index_base, amf_best_t* best_d, void* stream) {
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_mn_score_candidates: bad dtype");
  AMF_REQUIRE(criterion >= AMF_CRIT_APPROX_MEAN && criterion <= AMF_CRIT_PROB_GE,
              "amf_mn_score_candidates: criterion must be APPROX_MEAN, PRED_VARIANCE or PROB_GE");
  AMF_REQUIRE(mean_d && sig_d && om_d && best_d, "amf_mn_score_candidates: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  Best* part = nullptr;
  int rc = acquire_partials(&part, s);
  if (rc != AMF_OK) return rc;
  const int64_t blocks = (ncand + 127) / 128;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? (blocks > 0 ? blocks : 1)
                                                          : (int64_t)num_sms() * 16);
  if (dtype == AMF_F32)
    rc = maximize ? mn_score<float, true>(criterion, ncand, ci_d, cj_d, n, m, d, (const float*)mean_d, (const float*)sig_d, (const float*)om_d, cutoff, (float*)scores_d, index_base, part, grid, s)
                  : mn_score<float, false>(criterion, ncand, ci_d, cj_d, n, m, d, (const float*)mean_d, (const float*)sig_d, (const float*)om_d, cutoff, (float*)scores_d, index_base, part, grid, s);
  else
    rc = maximize ? mn_score<double, true>(criterion, ncand, ci_d, cj_d, n, m, d, (const double*)mean_d, (const double*)sig_d, (const double*)om_d, cutoff, (double*)scores_d, index_base, part, grid, s)
                  : mn_score<double, false>(criterion, ncand, ci_d, cj_d, n, m, d, (const double*)mean_d, (const double*)sig_d, (const double*)om_d, cutoff, (double*)scores_d, index_base, part, grid, s);
  if (rc != AMF_OK) return rc;
  return launch_best_final(part, grid, maximize != 0, best_d, s);
}

#pragma GCC visibility pop
}  // extern "C"
