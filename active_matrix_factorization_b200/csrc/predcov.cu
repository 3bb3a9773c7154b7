// Covariance between all pairs of predicted cells under the Gaussian approximation, and the
// log-determinant the prediction-entropy bound takes of it (active_pmf.py:324-390 approx_pred_covs,
// :559-574 _pred_entropy_bound; np.linalg.slogdet), batched over the (candidate, value) re-fits of
// the lookahead.  Exact mode only (k = (N+M)d <= a few hundred; the matrix is NM x NM).
//
// With X = [vec U; vec V] ~ N(mean, cov), x1 = U_ki, x2 = V_kj, x3 = U_la, x4 = V_lb, Isserlis gives
//   Cov(x1 x2, x3 x4) = m1 m3 C24 + m1 m4 C23 + m2 m3 C14 + m2 m4 C13 + C13 C24 + C14 C23
// and Cov(U_i.V_j, U_a.V_b) is its sum over k, l.
#include "common.cuh"

namespace amf {
namespace {

// one thread per entry ((i, j), (a, b)) of one problem; mean (k), cov (k, k) in the reference's
// layout (active_pmf.py:136-142): U_ki at i*d + k, V_kj at n*d + j*d + k
__global__ void __launch_bounds__(256)
pred_covs_kernel(int n, int m, int d, const double* __restrict__ mean, const double* __restrict__ cov,
                 double* __restrict__ out) {
  const int64_t nm = (int64_t)n * m, kdim = (int64_t)(n + m) * d;
  const double* mu = mean + blockIdx.y * kdim;
  const double* c = cov + blockIdx.y * kdim * kdim;
  double* o = out + blockIdx.y * nm * nm;
  const int64_t nu = (int64_t)n * d;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nm * nm;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / nm, col = e % nm;
    const int i = (int)(row / m), j = (int)(row % m), a = (int)(col / m), b = (int)(col % m);
    double s = 0;
    for (int k = 0; k < d; ++k) {
      const int64_t p1 = (int64_t)i * d + k, p2 = nu + (int64_t)j * d + k;
      const double m1 = mu[p1], m2 = mu[p2];
      for (int l = 0; l < d; ++l) {
        const int64_t p3 = (int64_t)a * d + l, p4 = nu + (int64_t)b * d + l;
        const double m3 = mu[p3], m4 = mu[p4];
        const double c13 = c[p1 * kdim + p3], c14 = c[p1 * kdim + p4];
        const double c23 = c[p2 * kdim + p3], c24 = c[p2 * kdim + p4];
        s += m1 * m3 * c24 + m1 * m4 * c23 + m2 * m3 * c14 + m2 * m4 * c13 + c13 * c24 + c14 * c23;
      }
    }
    o[e] = s;
  }
}

// sign and log|det| of one k x k matrix per CTA by LU with partial pivoting, in place in global
// memory (what LAPACK's getrf does for np.linalg.slogdet); a zero pivot gives sign 0, log -inf
__global__ void __launch_bounds__(256)
slogdet_kernel(int k, double* __restrict__ a_all, int* __restrict__ sign_out,
               double* __restrict__ logdet_out) {
  double* a = a_all + (int64_t)blockIdx.x * k * k;
  __shared__ double s_val[256];
  __shared__ int s_idx[256];
  __shared__ int s_piv;
  int sign = 1;
  double logdet = 0;
  for (int c = 0; c < k; ++c) {
    // pivot: largest |a[r][c]|, r >= c (lowest row on ties, as the sequential search does)
    double best = -1;
    int bi = c;
    for (int r = c + threadIdx.x; r < k; r += blockDim.x) {
      const double v = fabs(a[(int64_t)r * k + c]);
      if (v > best) { best = v; bi = r; }
    }
    s_val[threadIdx.x] = best; s_idx[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        const double ov = s_val[threadIdx.x + o];
        const int oi = s_idx[threadIdx.x + o];
        if (ov > s_val[threadIdx.x] || (ov == s_val[threadIdx.x] && oi < s_idx[threadIdx.x])) {
          s_val[threadIdx.x] = ov; s_idx[threadIdx.x] = oi;
        }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) s_piv = s_idx[0];
    __syncthreads();
    const int p = s_piv;
    if (p != c) {
      for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const double x = a[(int64_t)c * k + t];
        a[(int64_t)c * k + t] = a[(int64_t)p * k + t];
        a[(int64_t)p * k + t] = x;
      }
      sign = -sign;
    }
    __syncthreads();
    const double piv = a[(int64_t)c * k + c];
    if (piv == 0.0 || piv != piv) { sign = 0; logdet = -INFINITY; break; }
    if (piv < 0) sign = -sign;
    logdet += log(fabs(piv));
    // trailing update: a[r][t] -= (a[r][c] / piv) * a[c][t]
    const int rem = k - c - 1;
    for (int64_t e = threadIdx.x; e < (int64_t)rem * rem; e += blockDim.x) {
      const int r = c + 1 + (int)(e / rem), t = c + 1 + (int)(e % rem);
      a[(int64_t)r * k + t] -= a[(int64_t)r * k + c] / piv * a[(int64_t)c * k + t];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { sign_out[blockIdx.x] = sign; logdet_out[blockIdx.x] = logdet; }
}

}  // namespace
}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int amf_pred_covs(int32_t n, int32_t m, int d, int count, const double* mean_d,
                  const double* cov_d, double* out_d, void* stream) {
  AMF_REQUIRE(mean_d && cov_d && out_d, "amf_pred_covs: NULL argument");
  AMF_REQUIRE(n > 0 && m > 0 && d >= 1 && count >= 0, "amf_pred_covs: bad sizes");
  const int64_t nm2 = (int64_t)n * m * n * m, kdim = (int64_t)(n + m) * d;
  const int64_t blocks = (nm2 + 255) / 256;
  for (int done = 0; done < count; done += 65535) {      // gridDim.y holds at most 65535 problems
    const int now = count - done < 65535 ? count - done : 65535;
    const dim3 grid((unsigned)(blocks < 1024 ? blocks : 1024), (unsigned)now);
    pred_covs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, d, mean_d + done * kdim,
                                                           cov_d + done * kdim * kdim,
                                                           out_d + done * nm2);
    AMF_LAUNCH_CHECK();
  }
  return AMF_OK;
}

int amf_slogdet_batched(int k, int count, double* a_d, int* sign_d, double* logdet_d, void* stream) {
  AMF_REQUIRE(a_d && sign_d && logdet_d, "amf_slogdet_batched: NULL argument");
  AMF_REQUIRE(k >= 1 && count >= 0, "amf_slogdet_batched: bad sizes");
  if (count == 0) return AMF_OK;
  slogdet_kernel<<<count, 256, 0, (cudaStream_t)stream>>>(k, a_d, sign_d, logdet_d);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

#pragma GCC visibility pop
}  // extern "C"
