// Fast-mode BayesianPMF.sample_hyperparam on the device (bayes_pmf.py:158-186, sample_wishart
// :41-59): the mean and covariance of the factor rows, the Normal-Wishart posterior and ONE draw
// (mu, alpha) from it -- so that a fast-mode Gibbs chain never waits for the host: per sample the
// stream carries  moments -> draw -> half-sweep  for each side and nothing else.
//
//   x_bar, S   mean and covariance (ddof = 1) of the rows
//   M     = inv(W0) + n S + (b0 n / (b0 + n)) * <mu0 - x_bar, mu0 - x_bar>      (the reference adds
//           the SCALAR np.dot(diff, diff.T) to every entry: bayes_pmf.py:176; reproduced)
//   alpha ~ Wishart(inv(M), dof0 + n)      by Bartlett:  M = L L',  X = L^-T B,  alpha = X X'
//           (B lower triangular, B_ii = sqrt(chi2(dof - i)), B_ij ~ N(0, 1))
//   mu    ~ N((b0 mu0 + n x_bar) / (b0 + n), inv((b0 + n) alpha))  =  mu* + L B^-T z / sqrt(b0 + n)
// Same distributions as the reference's host code, not the same draws: Philox4x32-10 counters
// (philox.cuh) instead of numpy's global stream, and the Bartlett scheme for every dof.
#include "common.cuh"
#include "philox.cuh"

namespace amf {
namespace {

constexpr int HYPER_SLAB = 32;        // rows per shared-memory slab of the moment pass

// sums of (x - x0) and (x - x0)(x - x0)' over all rows, x0 = row 0 (shifted to avoid cancellation):
// work[0 .. d*d) = second moments, work[d*d .. d*d + d) = first moments
template <typename T>
__global__ void __launch_bounds__(256)
hyper_moments_kernel(const T* __restrict__ x, int64_t rows, int d, double* __restrict__ work) {
  extern __shared__ double slab[];                     // [HYPER_SLAB][d]
  const int np = d * d + d;
  double acc[5] = {0, 0, 0, 0, 0};                     // d <= 32: (d*d + d) / 256 <= 5 sums per thread
  for (int64_t r0 = (int64_t)blockIdx.x * HYPER_SLAB; r0 < rows; r0 += (int64_t)gridDim.x * HYPER_SLAB) {
    const int nr = (int)min((int64_t)HYPER_SLAB, rows - r0);
    __syncthreads();
    for (int t = threadIdx.x; t < nr * d; t += blockDim.x)
      slab[t] = (double)x[r0 * d + t] - (double)x[t % d];
    __syncthreads();
    int q = 0;
    for (int p = threadIdx.x; p < np; p += blockDim.x, ++q) {
      double s = 0;
      if (p < d * d) {
        const int k = p / d, l = p % d;
        for (int r = 0; r < nr; ++r) s = fma(slab[r * d + k], slab[r * d + l], s);
      } else {
        const int k = p - d * d;
        for (int r = 0; r < nr; ++r) s += slab[r * d + k];
      }
      acc[q] += s;
    }
  }
  int q = 0;
  for (int p = threadIdx.x; p < np; p += blockDim.x, ++q) atomicAdd(work + p, acc[q]);
}

// chi-square(k) = 2 Gamma(k / 2) by Marsaglia & Tsang (2000); every attempt has its own counter
__device__ double philox_chi2(double k, unsigned long long seed, unsigned long long stream,
                              uint32_t row) {
  const double a = 0.5 * k, dd = a - 1.0 / 3.0, c = rsqrt(9.0 * dd);
  for (uint32_t attempt = 0; attempt < 64; ++attempt) {
    double u1, u2, u3;
    philox_uniform3(seed, stream, row, 64u + attempt, u1, u2, u3);
    const double xn = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    const double v1 = 1.0 + c * xn;
    if (v1 <= 0) continue;
    const double v = v1 * v1 * v1;
    if (log(u3) < 0.5 * xn * xn + dd - dd * v + dd * log(v)) return 2.0 * dd * v;
  }
  return 2.0 * dd;                                     // unreachable in practice (acceptance > 95 %)
}

// 128 threads draw the random numbers of the Bartlett factor (a Philox normal costs a few hundred
// fp64 instructions: spread over the CTA they take one round instead of d), then warp 0 does the
// algebra with lane = row / column; d <= 32; everything in fp64
template <typename T>
__global__ void __launch_bounds__(128)
hyper_draw_kernel(const T* __restrict__ x, int64_t rows, int d, const double* __restrict__ work,
                  const double* __restrict__ prior, unsigned long long seed,
                  unsigned long long stream, T* __restrict__ mu_out, T* __restrict__ alpha_out,
                  int* __restrict__ fail) {
  __shared__ double L[32][33], B[32][33], X[32][33];
  __shared__ double tvec[32], xbar[32], zvec[32];
  const double n = (double)rows;
  const double* winv0 = prior;
  const double* mu0 = prior + d * d;
  const double b0 = prior[d * d + d], dof = floor(prior[d * d + d + 1] + n);
  // Bartlett factor and the normals of mu: entry (i, j) of B by thread i * d + j (strided)
  for (int e = threadIdx.x; e < d * d + d; e += blockDim.x) {
    if (e < d * d) {
      const int i = e / d, j = e % d;
      double v = 0.0;
      if (j < i) v = philox_normal(seed, stream, (uint32_t)i, (uint32_t)j);
      else if (j == i) v = sqrt(philox_chi2(dof - i, seed, stream, (uint32_t)i));
      B[i][j] = v;
    } else {
      zvec[e - d * d] = philox_normal(seed, stream, (uint32_t)(e - d * d), 32u);
    }
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  // mean, covariance, posterior scale matrix
  double sh = 0, xb = 0;
  if (lane < d) {
    sh = work[d * d + lane] / n;                       // mean of the shifted rows
    xb = (double)x[lane] + sh;
    xbar[lane] = xb;
  }
  double diff = lane < d ? mu0[lane] - xb : 0.0;
  double dd = diff * diff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
  tvec[lane] = sh;
  __syncwarp();
  const double cross = b0 * n / (b0 + n) * dd;
  if (lane < d) {
    for (int l = 0; l < d; ++l) {
      const double s_kl = (work[lane * d + l] - n * sh * tvec[l]) / (n - 1.0);
      L[lane][l] = winv0[lane * d + l] + n * s_kl + cross;
    }
  }
  __syncwarp();
  // Cholesky M = L L' in place (lower), lane = row
  bool ok = true;
  for (int j = 0; j < d; ++j) {
    double piv = L[j][j];
    for (int k = 0; k < j; ++k) piv -= L[j][k] * L[j][k];
    ok = ok && piv > 0 && piv == piv;
    const double lj = sqrt(piv > 0 ? piv : 1.0);
    __syncwarp();
    if (lane == j) L[j][j] = lj;
    if (lane > j && lane < d) {
      double v = L[lane][j];
      for (int k = 0; k < j; ++k) v -= L[lane][k] * L[j][k];
      L[lane][j] = v / lj;
    }
    __syncwarp();
  }
  // X = L^-T B: lane = column of X, back substitution with the upper triangular L'
  if (lane < d) {
    for (int i = d - 1; i >= 0; --i) {
      double v = B[i][lane];
      for (int k = i + 1; k < d; ++k) v -= L[k][i] * X[k][lane];
      X[i][lane] = v / L[i][i];
    }
  }
  __syncwarp();
  // alpha = X X'
  if (lane < d) {
    for (int l = 0; l < d; ++l) {
      double v = 0;
      for (int c = 0; c < d; ++c) v = fma(X[lane][c], X[l][c], v);
      alpha_out[lane * d + l] = (T)v;
    }
  }
  // mu = mu* + L (B^-T z) / sqrt(b0 + n)
  tvec[lane] = lane < d ? zvec[lane] : 0.0;
  __syncwarp();
  if (lane == 0) {
    for (int i = d - 1; i >= 0; --i) {                 // B' t = z, B' upper triangular
      double v = tvec[i];
      for (int k = i + 1; k < d; ++k) v -= B[k][i] * tvec[k];
      tvec[i] = v / B[i][i];
    }
  }
  __syncwarp();
  if (lane < d) {
    double y = 0;
    for (int k = 0; k <= lane; ++k) y = fma(L[lane][k], tvec[k], y);
    mu_out[lane] = (T)((b0 * mu0[lane] + n * xbar[lane]) / (b0 + n) + y * rsqrt(b0 + n));
  }
  if (!ok && lane == 0) atomicExch(fail, 1);
}

template <typename T>
int hyper_launch(const amf_ratings* h, int d, int64_t rows, const T* x, const double* prior,
                 unsigned long long seed, unsigned long long stream_id, T* mu, T* alpha,
                 cudaStream_t s) {
  // the handle's own workspace: no allocation per draw (a chain makes two draws per sample)
  double* work = h->sums_d + 8;
  const size_t wbytes = sizeof(double) * (size_t)(d * d + d);
  AMF_CUDA(cudaMemsetAsync(work, 0, wbytes, s));
  int64_t slabs = (rows + HYPER_SLAB - 1) / HYPER_SLAB;
  const int grid = (int)(slabs < num_sms() ? slabs : num_sms());
  hyper_moments_kernel<T><<<grid, 256, sizeof(double) * HYPER_SLAB * d, s>>>(x, rows, d, work);
  AMF_LAUNCH_CHECK();
  int* fail = reinterpret_cast<int*>(h->sums_d + 6);   // the handle's sticky Gibbs failure flag
  hyper_draw_kernel<T><<<1, 128, 0, s>>>(x, rows, d, work, prior, seed, stream_id, mu, alpha, fail);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

}  // namespace

// typed entry for the chain driver of gibbs.cu
int gibbs_hyper_draw(const amf_ratings* h, int dtype, int d, int64_t rows, const void* feats,
                     const double* prior, unsigned long long seed, unsigned long long stream_id,
                     void* mu, void* alpha, cudaStream_t s) {
  if (dtype == AMF_F32)
    return hyper_launch<float>(h, d, rows, (const float*)feats, prior, seed, stream_id, (float*)mu,
                               (float*)alpha, s);
  return hyper_launch<double>(h, d, rows, (const double*)feats, prior, seed, stream_id, (double*)mu,
                              (double*)alpha, s);
}

}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int amf_gibbs_hyper_device(const amf_ratings_t* h, int dtype, int d, int64_t rows,
                           const void* feats_d, const double* prior_d, uint64_t seed,
                           uint64_t stream_id, void* mu_out_d, void* alpha_out_d, void* stream) {
  AMF_REQUIRE(h && feats_d && prior_d && mu_out_d && alpha_out_d, "amf_gibbs_hyper_device: NULL argument");
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_gibbs_hyper_device: bad dtype");
  AMF_REQUIRE(d >= 1 && d <= 32, "amf_gibbs_hyper_device: d must be in [1, 32]");
  AMF_REQUIRE(rows >= 2, "amf_gibbs_hyper_device: the covariance needs at least two rows");
  return gibbs_hyper_draw(h, dtype, d, rows, feats_d, prior_d, seed, stream_id, mu_out_d,
                          alpha_out_d, (cudaStream_t)stream);
}

#pragma GCC visibility pop
}  // extern "C"
