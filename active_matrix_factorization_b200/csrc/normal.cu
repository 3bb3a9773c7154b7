// Full-covariance Gaussian approximation of the PMF posterior (exact mode), batched.
//
//   KL(PMF || N(mean, cov))        active_pmf.py:202-240  kl_divergence
//   its gradient                   normal_exps_cy.pyx:140-303  normal_gradient / _normal_grad
//   PSD projection                 active_pmf.py:36-50  project_psd  (np.linalg.eigh)
//   the line-search fit            active_pmf.py:251-288  fit_normal_kls
//   lookahead criteria             active_pmf.py:526-530 _approx_entropy, :605-606 _total_variance
//
// The reference evaluates a lookahead criterion by deep-copying the model once per candidate
// and rating value and re-running fit_normal on the host (active_pmf.py:668-676).  Here every
// (candidate, value) pair is one independent problem handled by ONE CTA from start to finish:
// the problems share the rating list and differ by one appended rating; mean (k) and the k x k
// matrices live in a per-problem workspace (L1/L2 resident for the sizes this mode is usable
// at), the eigendecomposition is a parallel cyclic Jacobi, inverse and log-determinant come
// from a Cholesky factorisation, and the accept/reject decisions of the line search are taken
// on the device -- nothing returns to the host until the batch is done.  fp64 throughout.
//
// Reference quirks reproduced in the gradient (see oracle/pmf_oracle.py normal_gradient): for
// latent_d > 2 the l>k cross terms add the l-SUM to every l>k target.
#include "common.cuh"
#include "dense_blk.cuh"

namespace amf {

constexpr int NRM_THREADS = 256;

struct NormalProblem {
  int n, m, d, k;
  int64_t nnz;
  const int32_t* ri;
  const int32_t* rj;
  const double* rr;
  int ei, ej;      // extra rating (ei < 0: none)
  double er;
  double sigma_sq, sigma_u_sq, sigma_v_sq;
};

// ---- rating access: base list followed by the optional extra rating ----------------------
__device__ __forceinline__ void get_rating(const NormalProblem& P, int64_t t, int& i, int& j,
                                           double& r) {
  if (t < P.nnz) { i = P.ri[t]; j = P.rj[t]; r = P.rr[t]; }
  else { i = P.ei; j = P.ej; r = P.er; }
}

// data part of the KL: sum_r E[(Ui.Vj)^2] - 2 r E[Ui.Vj] + r^2     (active_pmf.py:218-229)
__device__ double kl_data(const NormalProblem& P, const double* mean,
                          const double* cov, double* red) {
  const int d = P.d, k = P.k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int64_t total = P.nnz + (P.ei >= 0 ? 1 : 0);
  double acc = 0;
  for (int64_t t = warp; t < total; t += nwarps) {
    int i, j; double r;
    get_rating(P, t, i, j, r);
    const int a0 = i * d, b0 = P.n * d + j * d;
    for (int kk = lane; kk < d; kk += 32) {
      const int a = a0 + kk, b = b0 + kk;
      const double ma = mean[a], mb = mean[b];
      const double cab = cov[(int64_t)a * k + b];
      // exp_squared (normal_exps_cy.pyx:77-87)
      double term = 4 * ma * mb * cab + 2 * cab * cab +
                    (ma * ma + cov[(int64_t)a * k + a]) * (mb * mb + cov[(int64_t)b * k + b]);
      term -= 2 * r * (ma * mb + cab);
      for (int l = kk + 1; l < d; ++l) {
        // 2 * quadexpect(u_k, v_k, u_l, v_l)  (normal_exps_cy.pyx:42-73)
        const int c = a0 + l, e = b0 + l;
        const double mc = mean[c], me = mean[e];
        const double ccd = cov[(int64_t)c * k + e], cbd = cov[(int64_t)b * k + e];
        const double cbc = cov[(int64_t)b * k + c], cad = cov[(int64_t)a * k + e];
        const double cac = cov[(int64_t)a * k + c];
        term += 2 * (ma * mb * mc * me + ma * mb * ccd + ma * mc * cbd + ma * me * cbc +
                     mb * mc * cad + mb * me * cac + mc * me * cab + cab * ccd + cac * cbd +
                     cad * cbc);
      }
      acc += term;
    }
    if (lane == 0) acc += r * r;
  }
  return blk_sum(acc, red);
}

// prior part: (|m_u|^2 + tr S_uu)/(2 su) + same for v          (active_pmf.py:231-234)
__device__ double kl_prior(const NormalProblem& P, const double* mean,
                           const double* cov, double* red) {
  const int nu = P.n * P.d, k = P.k;
  double acc = 0;
  for (int t = threadIdx.x; t < k; t += blockDim.x) {
    const double v = mean[t] * mean[t] + cov[(int64_t)t * k + t];
    acc += v / (2 * (t < nu ? P.sigma_u_sq : P.sigma_v_sq));
  }
  return blk_sum(acc, red);
}

// KL(mean, cov); `work` (k*k) is scratch for the Cholesky factor
__device__ double kl_full(const NormalProblem& P, const double* mean, const double* cov,
                          double* work, double* red, int* flag) {
  const int k = P.k;
  double div = kl_data(P, mean, cov, red) / (2 * P.sigma_sq);
  div += kl_prior(P, mean, cov, red);
  for (int t = threadIdx.x; t < k * k; t += blockDim.x) work[t] = cov[t];
  __syncthreads();
  const double logdet = blk_cholesky(work, k, red, flag);
  return div - logdet / 2;
}

__device__ __forceinline__ double trip(const double* mean,
                                       const double* cov, int k, int a, int b, int c) {
  return mean[a] * mean[b] * mean[c] + mean[a] * cov[(int64_t)b * k + c] +
         mean[b] * cov[(int64_t)a * k + c] + mean[c] * cov[(int64_t)a * k + b];
}
__device__ __forceinline__ double mom2(const double* mean,
                                       const double* cov, int k, int a, int b) {
  return mean[a] * mean[b] + cov[(int64_t)a * k + b];
}
__device__ __forceinline__ void sym_add(double* gc, int k, int a, int b, double inc) {
  atomicAdd(gc + (int64_t)a * k + b, inc);
  atomicAdd(gc + (int64_t)b * k + a, inc);
}

// gradient of the KL; `work`/`work2` (k*k each) are scratch.  gm (k), gc (k*k) are overwritten.
__device__ void kl_gradient(const NormalProblem& P, const double* mean,
                            const double* cov, double* gm, double* gc,
                            double* work, double* work2, double* red, int* flag) {
  const int d = P.d, k = P.k;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const double sig = P.sigma_sq;
  for (int t = tid; t < k; t += nt) gm[t] = 0;
  for (int t = tid; t < k * k; t += nt) gc[t] = 0;
  __syncthreads();
  const int64_t total = P.nnz + (P.ei >= 0 ? 1 : 0);
  // Short lists (the toy problems the exact mode is meant for) are walked by ONE warp in
  // rating-list order, like the reference's loop: the sums then do not depend on the order in
  // which warps reach the atomics, and the fit -- which amplifies 1e-16 differences to 1e-5 in
  // the criteria (parity caveat) -- is reproducible from run to run.
  const int nw_eff = total <= 64 ? 1 : nwarps;
  for (int64_t t = warp; t < total && warp < nw_eff; t += nw_eff) {
    int i, j; double rating;
    get_rating(P, t, i, j, rating);
    const int a0 = i * d, b0 = P.n * d + j * d;
    for (int kk = lane; kk < d; kk += 32) {
      const int a = a0 + kk, b = b0 + kk;
      if (kk < d - 1) {
        // sums over l > k                                  (normal_exps_cy.pyx:239-258)
        double s1 = 0, s2 = 0, s3 = 0, s4 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
        for (int l = kk + 1; l < d; ++l) {
          const int c = a0 + l, e = b0 + l;
          s1 += trip(mean, cov, k, b, c, e);
          s2 += trip(mean, cov, k, a, c, e);
          s3 += trip(mean, cov, k, a, b, e);
          s4 += trip(mean, cov, k, a, b, c);
          t1 += mom2(mean, cov, k, c, e);
          t2 += mom2(mean, cov, k, b, e);
          t3 += mom2(mean, cov, k, b, c);
          t4 += mom2(mean, cov, k, a, e);
          t5 += mom2(mean, cov, k, a, c);
        }
        atomicAdd(gm + a, s1 / sig);
        atomicAdd(gm + b, s2 / sig);
        sym_add(gc, k, a, b, t1 / sig);
        const double pair = mom2(mean, cov, k, a, b) / sig;
        for (int l = kk + 1; l < d; ++l) {
          const int c = a0 + l, e = b0 + l;
          atomicAdd(gm + c, s3 / sig);        // quirk: the l-sum goes to every l > k
          atomicAdd(gm + e, s4 / sig);        // quirk
          sym_add(gc, k, a, c, t2 / sig);     // quirk
          sym_add(gc, k, a, e, t3 / sig);     // quirk
          sym_add(gc, k, b, c, t4 / sig);     // quirk
          sym_add(gc, k, b, e, t5 / sig);     // quirk
          sym_add(gc, k, c, e, pair);
        }
      }
      // terms vectorised over k in the reference            (normal_exps_cy.pyx:260-283)
      const double ma = mean[a], mb = mean[b];
      const double cab = cov[(int64_t)a * k + b];
      const double caa = cov[(int64_t)a * k + a], cbb = cov[(int64_t)b * k + b];
      atomicAdd(gm + a, (2 * mb * cab + ma * (mb * mb + cbb)) / sig - mb * (rating / sig));
      atomicAdd(gm + b, (2 * ma * cab + mb * (ma * ma + caa)) / sig - ma * (rating / sig));
      atomicAdd(gc + (int64_t)a * k + a, (mb * mb + cbb) / (2 * sig));
      atomicAdd(gc + (int64_t)b * k + b, (ma * ma + caa) / (2 * sig));
      sym_add(gc, k, a, b, 2 * (ma * mb + cab) / sig - rating / sig);
    }
  }
  __syncthreads();
  // priors (:287-291)
  const int nu = P.n * d;
  for (int t = tid; t < k; t += nt) {
    const double s = t < nu ? P.sigma_u_sq : P.sigma_v_sq;
    gm[t] += mean[t] / s;
    gc[(int64_t)t * k + t] += 1 / (2 * s);
  }
  // entropy term: grad_cov -= (inv + inv' o (1 - I)) / 2   (:302-303)
  for (int t = tid; t < k * k; t += nt) work[t] = cov[t];
  __syncthreads();
  blk_cholesky(work, k, red, flag);
  blk_tri_inverse(work, work2, k);
  for (int t = tid; t < k * k; t += nt) {
    const int r = t / k, c = t % k;
    double s = 0;
    for (int q = max(r, c); q < k; ++q) s += work2[(int64_t)q * k + r] * work2[(int64_t)q * k + c];
    gc[t] -= (r == c) ? s / 2 : s;
  }
  __syncthreads();
}

// total predictive variance sum_ij Var[Ui.Vj]  (active_pmf.py:301-322, :605-606)
__device__ double blk_total_variance(const NormalProblem& P, const double* mean,
                                     const double* cov, double* red) {
  const int d = P.d, k = P.k;
  double acc = 0;
  for (int cell = threadIdx.x; cell < P.n * P.m; cell += blockDim.x) {
    const int i = cell / P.m, j = cell % P.m;
    const int a0 = i * d, b0 = P.n * d + j * d;
    double e = 0, ex2 = 0;
    for (int kk = 0; kk < d; ++kk) {
      const int a = a0 + kk, b = b0 + kk;
      const double ma = mean[a], mb = mean[b], cab = cov[(int64_t)a * k + b];
      e += ma * mb + cab;
      ex2 += 4 * ma * mb * cab + 2 * cab * cab +
             (ma * ma + cov[(int64_t)a * k + a]) * (mb * mb + cov[(int64_t)b * k + b]);
      for (int l = kk + 1; l < d; ++l) {
        const int c = a0 + l, f = b0 + l;
        const double mc = mean[c], mf = mean[f];
        ex2 += 2 * (ma * mb * mc * mf + ma * mb * cov[(int64_t)c * k + f] +
                    ma * mc * cov[(int64_t)b * k + f] + ma * mf * cov[(int64_t)b * k + c] +
                    mb * mc * cov[(int64_t)a * k + f] + mb * mf * cov[(int64_t)a * k + c] +
                    mc * mf * cab + cab * cov[(int64_t)c * k + f] +
                    cov[(int64_t)a * k + c] * cov[(int64_t)b * k + f] +
                    cov[(int64_t)a * k + f] * cov[(int64_t)b * k + c]);
      }
    }
    acc += ex2 - e * e;       // E[x^2] - E[x]^2 like the reference (active_pmf.py:320)
  }
  return blk_sum(acc, red);
}

struct FitArgs {
  int B;
  int n, m, d;
  int64_t nnz;
  const int32_t* ri; const int32_t* rj; const double* rr;
  const int32_t* ei; const int32_t* ej; const double* er;   // per problem, may be NULL
  double sigma_sq, sigma_u_sq, sigma_v_sq;
  double lr0, min_eig, kl_stop, min_lr;
  int max_steps;
  double* mean;        // B x k   in/out
  double* cov;         // B x k x k  in/out
  double* work;        // B x (2k + 5 k^2) scratch
  double* kl_out;      // B       final KL
  int* steps_out;      // B       accepted steps
  double* kl_trace;    // B x trace_len or NULL
  int trace_len;
  double* entropy_out; // B or NULL : slogdet(cov) after the fit
  double* totvar_out;  // B or NULL : total predictive variance after the fit
  int mode;            // 0 fit, 1 kl only, 2 gradient only (grad -> work), 3 project only
  bool smem_state;     // state + work matrices live in shared memory (small k)
};

__device__ NormalProblem make_problem(const FitArgs& a, int b) {
  NormalProblem P;
  P.n = a.n; P.m = a.m; P.d = a.d; P.k = (a.n + a.m) * a.d;
  P.nnz = a.nnz; P.ri = a.ri; P.rj = a.rj; P.rr = a.rr;
  P.ei = a.ei ? a.ei[b] : -1; P.ej = a.ej ? a.ej[b] : -1; P.er = a.er ? a.er[b] : 0.0;
  P.sigma_sq = a.sigma_sq; P.sigma_u_sq = a.sigma_u_sq; P.sigma_v_sq = a.sigma_v_sq;
  return P;
}

__global__ void __launch_bounds__(NRM_THREADS) normal_fit_kernel(FitArgs a) {
  extern __shared__ double sh[];
  double* red = sh;                 // 32
  double* cs = sh + 32;             // 2 * (k/2 + 1)
  __shared__ int flag;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const NormalProblem P = make_problem(a, b);
    const int k = P.k;
    const int64_t kk2 = (int64_t)k * k;
    double* mean_g = a.mean + (int64_t)b * k;
    double* cov_g = a.cov + (int64_t)b * kk2;
    // Small problems (the lookahead batches) keep the state and all work matrices in shared
    // memory: every Jacobi / Cholesky phase is then a shared-memory round trip instead of an L2
    // one.  Larger k falls back to the per-problem global workspace.
    const bool in_smem = a.smem_state;
    double* S = sh + blk_scratch_doubles(k);
    double* mean = in_smem ? S : mean_g;
    double* cov = in_smem ? S + k : cov_g;
    double* W = in_smem ? S + k + kk2 : a.work + (int64_t)b * (2 * k + 5 * kk2);
    if (in_smem) {
      for (int t = tid; t < k; t += nt) mean[t] = mean_g[t];
      for (int64_t t = tid; t < kk2; t += nt) cov[t] = cov_g[t];
      __syncthreads();
    }
    double* gm = W;               // k
    double* nmean = W + k;        // k
    double* gc = W + 2 * k;       // k*k
    double* ncov = gc + kk2;      // k*k
    double* w1 = ncov + kk2;      // scratch
    double* w2 = w1 + kk2;
    double* w3 = w2 + kk2;

    if (a.mode == 1) {
      const double kl = kl_full(P, mean, cov, w1, red, &flag);
      if (tid == 0) a.kl_out[b] = kl;
      continue;
    }
    if (a.mode == 2) {
      kl_gradient(P, mean, cov, gm, gc, w1, w2, red, &flag);
      if (in_smem) {            // the caller reads the gradient from the global workspace
        double* Wg = a.work + (int64_t)b * (2 * k + 5 * kk2);
        for (int t = tid; t < k; t += nt) Wg[t] = gm[t];
        for (int64_t t = tid; t < kk2; t += nt) Wg[2 * k + t] = gc[t];
        __syncthreads();
      }
      continue;
    }
    if (a.mode == 3) {
      blk_project_psd(cov, k, a.min_eig, w1, w2, cs, red);
      if (in_smem) {
        for (int64_t t = tid; t < kk2; t += nt) cov_g[t] = cov[t];
        __syncthreads();
      }
      continue;
    }

    double lr = a.lr0;
    double old_kl = kl_full(P, mean, cov, w1, red, &flag);
    int steps = 0;
    bool converged = false;
    while (!converged) {
      kl_gradient(P, mean, cov, gm, gc, w1, w2, red, &flag);
      while (true) {
        for (int t = tid; t < k; t += nt) nmean[t] = mean[t] - lr * gm[t];
        for (int64_t t = tid; t < kk2; t += nt) ncov[t] = cov[t] - lr * gc[t];
        __syncthreads();
        blk_project_psd(ncov, k, a.min_eig, w1, w2, cs, red);
        const double new_kl = kl_full(P, nmean, ncov, w3, red, &flag);
        if (new_kl < old_kl) {
          for (int t = tid; t < k; t += nt) mean[t] = nmean[t];
          for (int64_t t = tid; t < kk2; t += nt) cov[t] = ncov[t];
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
      for (int t = tid; t < k; t += nt) mean_g[t] = mean[t];
      for (int64_t t = tid; t < kk2; t += nt) cov_g[t] = cov[t];
    }
    if (a.entropy_out) {
      for (int64_t t = tid; t < kk2; t += nt) w1[t] = cov[t];
      __syncthreads();
      const double ld = blk_cholesky(w1, k, red, &flag);
      if (tid == 0) a.entropy_out[b] = ld;
    }
    if (a.totvar_out) {
      const double tv = blk_total_variance(P, mean, cov, red);
      if (tid == 0) a.totvar_out[b] = tv;
    }
    __syncthreads();
  }
}

}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int64_t amf_normal_workspace_doubles(int32_t n, int32_t m, int d) {
  const int64_t k = (int64_t)(n + m) * d;
  return 2 * k + 5 * k * k;
}

int amf_normal_batched(int mode, int B, int64_t nnz, const int32_t* ri_d, const int32_t* rj_d,
                       const double* rr_d, const int32_t* extra_i_d, const int32_t* extra_j_d,
                       const double* extra_r_d, const amf_normal_fit_params_t* p, double* mean_d,
                       double* cov_d, double* work_d, double* kl_out_d, int32_t* steps_out_d,
                       double* kl_trace_d, int trace_len, double* entropy_out_d,
                       double* totvar_out_d, void* stream) {
  AMF_REQUIRE(p && mean_d && cov_d && work_d, "amf_normal_batched: NULL argument");
  AMF_REQUIRE(mode >= 0 && mode <= 3, "amf_normal_batched: bad mode %d", mode);
  AMF_REQUIRE(B >= 0 && nnz >= 0 && p->n > 0 && p->d > 0 && (p->m > 0 || (mode == 3 && p->m == 0)),
              "amf_normal_batched: bad sizes");
  AMF_REQUIRE(mode != 0 || (kl_out_d && steps_out_d), "amf_normal_batched: fit needs kl_out/steps_out");
  AMF_REQUIRE(mode != 1 || kl_out_d, "amf_normal_batched: kl mode needs kl_out");
  if (B == 0) return AMF_OK;
  const int64_t k = (int64_t)(p->n + p->m) * p->d;
  AMF_REQUIRE(k * k < (1ll << 31), "amf_normal_batched: k=%lld too large for exact mode", (long long)k);
  FitArgs a{};
  a.B = B; a.n = p->n; a.m = p->m; a.d = p->d; a.nnz = nnz;
  a.ri = ri_d; a.rj = rj_d; a.rr = rr_d; a.ei = extra_i_d; a.ej = extra_j_d; a.er = extra_r_d;
  a.sigma_sq = p->sigma_sq; a.sigma_u_sq = p->sigma_u_sq; a.sigma_v_sq = p->sigma_v_sq;
  a.lr0 = p->learning_rate; a.min_eig = p->min_eig; a.kl_stop = p->kl_stop; a.min_lr = p->min_lr;
  a.max_steps = p->max_steps;
  a.mean = mean_d; a.cov = cov_d; a.work = work_d; a.kl_out = kl_out_d; a.steps_out = steps_out_d;
  a.kl_trace = kl_trace_d; a.trace_len = trace_len; a.entropy_out = entropy_out_d;
  a.totvar_out = totvar_out_d; a.mode = mode;
  size_t smem = sizeof(double) * (size_t)blk_scratch_doubles(k);
  const size_t state = sizeof(double) * (size_t)(3 * k + 6 * k * k);   // mean, cov + workspace
  a.smem_state = smem + state <= 200 * 1024;
  if (a.smem_state) smem += state;
  AMF_CUDA(cudaFuncSetAttribute(normal_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
  cudaStream_t s = (cudaStream_t)stream;
  normal_fit_kernel<<<B, NRM_THREADS, smem, s>>>(a);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

#pragma GCC visibility pop
}  // extern "C"
