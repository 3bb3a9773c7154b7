// Dense linear algebra on one k x k fp64 matrix per CTA (global or shared memory operands):
// block sum, Cholesky (log-det), triangular inverse, parallel cyclic Jacobi eigensolver and the
// PSD projection of active_pmf.py:36-50 / mn_active_pmf.py:42-67.  Shared by normal.cu and mn.cu.
#pragma once
#include "common.cuh"

namespace amf {

__device__ __forceinline__ double blk_sum(double v, double* red) {
  // all threads get the result
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t = 0;
  const int nw = blockDim.x >> 5;
  for (int q = 0; q < nw; ++q) t += red[q];
  return t;
}

// in-place lower Cholesky in (global or shared) memory; returns log det = 2 sum log L_ii, NaN if
// the matrix is not positive definite.  Upper triangle is left untouched.
__device__ inline double blk_cholesky(double* A, int k, double* red, int* flag) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) *flag = 1;
  __syncthreads();
  for (int j = 0; j < k; ++j) {
    if (tid == 0) {
      const double pv = A[(int64_t)j * k + j];
      if (!(pv > 0.0)) *flag = 0;
      A[(int64_t)j * k + j] = sqrt(pv);
    }
    __syncthreads();
    const double inv = 1.0 / A[(int64_t)j * k + j];
    for (int i = j + 1 + tid; i < k; i += nt) A[(int64_t)i * k + j] *= inv;
    __syncthreads();
    const int rem = k - j - 1;
    for (int t = tid; t < rem * rem; t += nt) {
      const int i = j + 1 + t / rem, c = j + 1 + t % rem;
      if (c <= i) A[(int64_t)i * k + c] -= A[(int64_t)i * k + j] * A[(int64_t)c * k + j];
    }
    __syncthreads();
  }
  double acc = 0;
  for (int t = tid; t < k; t += nt) acc += log(A[(int64_t)t * k + t]);
  const double ld = 2 * blk_sum(acc, red);
  return *flag ? ld : NAN;
}

// Linv (lower) = L^-1 ; one thread per column
__device__ inline void blk_tri_inverse(const double* L, double* Linv, int k) {
  for (int c = threadIdx.x; c < k; c += blockDim.x) {
    for (int i = 0; i < k; ++i) {
      if (i < c) { Linv[(int64_t)i * k + c] = 0.0; continue; }
      double s = (i == c) ? 1.0 : 0.0;
      for (int q = c; q < i; ++q) s -= L[(int64_t)i * k + q] * Linv[(int64_t)q * k + c];
      Linv[(int64_t)i * k + c] = s / L[(int64_t)i * k + i];
    }
  }
  __syncthreads();
}

// Parallel cyclic Jacobi eigendecomposition of the symmetric k x k matrix A (destroyed: its
// diagonal ends up holding the eigenvalues); Q receives the eigenvectors (columns).
// cs: shared scratch of 2*(k/2+1) doubles.  Returns the smallest eigenvalue.
__device__ inline double blk_jacobi(double* A, double* Q, int k, double* cs, double* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int K = (k + 1) & ~1;          // players in the round-robin (one dummy if k is odd)
  const int npairs = K / 2;
  for (int t = tid; t < k * k; t += nt) Q[t] = (t / k == t % k) ? 1.0 : 0.0;
  __syncthreads();
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0, dg = 0;
    for (int t = tid; t < k * k; t += nt) {
      const int r = t / k, c = t % k;
      const double v = A[t];
      if (r == c) dg += v * v; else off += v * v;
    }
    off = blk_sum(off, red);
    dg = blk_sum(dg, red);
    if (off <= 1e-28 * dg || off == 0.0) break;   // relative off-diagonal norm 1e-14
    for (int round = 0; round < K - 1; ++round) {
      // pair t of this round: (K-1, round) for t == 0, else ((round+t) % (K-1), (round-t) % (K-1))
      for (int t = tid; t < npairs; t += nt) {
        int p = t == 0 ? K - 1 : (round + t) % (K - 1);
        int q = t == 0 ? round : (round - t + (K - 1)) % (K - 1);
        if (p > q) { const int tmp = p; p = q; q = tmp; }
        double c = 1.0, s = 0.0;
        if (q < k) {
          const double apq = A[(int64_t)p * k + q];
          if (apq != 0.0) {
            const double theta = (A[(int64_t)q * k + q] - A[(int64_t)p * k + p]) / (2 * apq);
            const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = 1.0 / sqrt(tt * tt + 1.0);
            s = tt * c;
          }
        }
        cs[2 * t] = c; cs[2 * t + 1] = s;
      }
      __syncthreads();
      // rows p, q of every pair
      for (int w = tid; w < npairs * k; w += nt) {
        const int t = w / k, col = w % k;
        int p = t == 0 ? K - 1 : (round + t) % (K - 1);
        int q = t == 0 ? round : (round - t + (K - 1)) % (K - 1);
        if (p > q) { const int tmp = p; p = q; q = tmp; }
        if (q >= k) continue;
        const double c = cs[2 * t], s = cs[2 * t + 1];
        const double x = A[(int64_t)p * k + col], y = A[(int64_t)q * k + col];
        A[(int64_t)p * k + col] = c * x - s * y;
        A[(int64_t)q * k + col] = s * x + c * y;
      }
      __syncthreads();
      // columns p, q of every pair, and the eigenvector accumulation
      for (int w = tid; w < npairs * k; w += nt) {
        const int t = w / k, row = w % k;
        int p = t == 0 ? K - 1 : (round + t) % (K - 1);
        int q = t == 0 ? round : (round - t + (K - 1)) % (K - 1);
        if (p > q) { const int tmp = p; p = q; q = tmp; }
        if (q >= k) continue;
        const double c = cs[2 * t], s = cs[2 * t + 1];
        double x = A[(int64_t)row * k + p], y = A[(int64_t)row * k + q];
        A[(int64_t)row * k + p] = c * x - s * y;
        A[(int64_t)row * k + q] = s * x + c * y;
        x = Q[(int64_t)row * k + p]; y = Q[(int64_t)row * k + q];
        Q[(int64_t)row * k + p] = c * x - s * y;
        Q[(int64_t)row * k + q] = s * x + c * y;
      }
      __syncthreads();
    }
  }
  double mn = INFINITY;
  for (int t = tid; t < k; t += nt) mn = fmin(mn, A[(int64_t)t * k + t]);
  // block min through the sum helper's buffer
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = mn;
  __syncthreads();
  double r = INFINITY;
  for (int q = 0; q < (nt >> 5); ++q) r = fmin(r, red[q]);
  __syncthreads();
  return r;
}

// M <- project_psd(M, min_eig)  (active_pmf.py:36-50).  work, work2: k*k scratch each.
__device__ inline void blk_project_psd(double* M, int k, double min_eig, double* work, double* work2,
                                double* cs, double* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  // symmetrise into M, copy to work
  for (int t = tid; t < k * k; t += nt) {
    const int r = t / k, c = t % k;
    if (c >= r) {
      const double v = (M[(int64_t)r * k + c] + M[(int64_t)c * k + r]) / 2;
      work[(int64_t)r * k + c] = v;
      work[(int64_t)c * k + r] = v;
    }
  }
  __syncthreads();
  for (int t = tid; t < k * k; t += nt) M[t] = work[t];
  __syncthreads();
  const double mn = blk_jacobi(work, work2, k, cs, red);   // work diag = eigenvalues, work2 = Q
  if (mn < min_eig) {
    for (int t = tid; t < k * k; t += nt) {
      const int r = t / k, c = t % k;
      if (c >= r) {
        double s = 0;
        for (int l = 0; l < k; ++l)
          s += fmax(work[(int64_t)l * k + l], min_eig) * work2[(int64_t)r * k + l] *
               work2[(int64_t)c * k + l];
        M[(int64_t)r * k + c] = s;
        M[(int64_t)c * k + r] = s;
      }
    }
  }
  __syncthreads();
}


}  // namespace amf
