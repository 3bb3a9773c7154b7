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

// Linv (lower) = L^-1.  Four lanes per column (forward substitution L x = e_c), the inner
// product of row i split over the lanes and reduced with two shuffles.
__device__ inline void blk_tri_inverse(const double* L, double* Linv, int k) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int sub = tid & 3;
  const unsigned mask = 0xFu << ((tid & 31) & ~3);
  for (int c = tid >> 2; c < k; c += nt >> 2) {
    for (int i = sub; i < c; i += 4) Linv[(int64_t)i * k + c] = 0.0;
    for (int i = c; i < k; ++i) {
      double s = 0;
      for (int q = c + sub; q < i; q += 4) s += L[(int64_t)i * k + q] * Linv[(int64_t)q * k + c];
      s += __shfl_xor_sync(mask, s, 1);
      s += __shfl_xor_sync(mask, s, 2);
      if (sub == 0) Linv[(int64_t)i * k + c] = ((i == c ? 1.0 : 0.0) - s) / L[(int64_t)i * k + i];
      __syncwarp(mask);
    }
  }
  __syncthreads();
}

// doubles of shared scratch the helpers below need for a k x k problem: 32 for block reductions,
// then per rotation pair (c, s) as doubles and (p, q) as ints
__host__ __device__ inline int64_t blk_scratch_doubles(int64_t k) { return 32 + 4 * (k / 2 + 2); }

// Parallel cyclic Jacobi eigendecomposition of the symmetric k x k matrix A (destroyed: its
// diagonal ends up holding the eigenvalues); Q receives the eigenvectors (columns).
// cs: shared scratch of 4*(k/2+2) doubles.  Returns the smallest eigenvalue.
// Round-robin ordering: K-1 rounds of K/2 disjoint pairs per sweep.  Per round: (1) one thread
// per pair computes the rotation and publishes (p, q, c, s); (2) rows p, q -- one warp per pair,
// lanes over columns; (3) columns p, q of A and Q -- warps over rows, one lane per pair, so
// shared-memory accesses of a warp fall in one row (no bank conflicts for any k).
__device__ inline double blk_jacobi(double* A, double* Q, int k, double* cs, double* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const int K = (k + 1) & ~1;          // players in the round-robin (one dummy if k is odd)
  const int npairs = K / 2;
  int* pq = reinterpret_cast<int*>(cs + 2 * (k / 2 + 2));
  for (int r = warp; r < k; r += nwarps)
    for (int c = lane; c < k; c += 32) Q[(int64_t)r * k + c] = (r == c) ? 1.0 : 0.0;
  __syncthreads();
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0, dg = 0;
    for (int r = warp; r < k; r += nwarps)
      for (int c = lane; c < k; c += 32) {
        const double v = A[(int64_t)r * k + c];
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
        pq[2 * t] = p; pq[2 * t + 1] = (q < k) ? q : -1;
      }
      __syncthreads();
      // A <- J' A J in one phase: the 2x2 block (pair a) x (pair b) is owned by one thread, which
      // applies pair a's rotation to its rows and then pair b's rotation to its columns -- the
      // same operations, in the same order, as a row pass followed by a column pass.
      for (int w = tid; w < npairs * npairs; w += nt) {
        const int ta = w / npairs, tb = w - ta * npairs;
        const int pa = pq[2 * ta], qa = pq[2 * ta + 1], pb = pq[2 * tb], qb = pq[2 * tb + 1];
        const double ca = cs[2 * ta], sa = cs[2 * ta + 1], cb = cs[2 * tb], sb = cs[2 * tb + 1];
        const double x00 = A[(int64_t)pa * k + pb];
        const double x01 = qb >= 0 ? A[(int64_t)pa * k + qb] : 0.0;
        const double x10 = qa >= 0 ? A[(int64_t)qa * k + pb] : 0.0;
        const double x11 = (qa >= 0 && qb >= 0) ? A[(int64_t)qa * k + qb] : 0.0;
        const double r00 = ca * x00 - sa * x10, r01 = ca * x01 - sa * x11;   // rows
        const double r10 = sa * x00 + ca * x10, r11 = sa * x01 + ca * x11;
        A[(int64_t)pa * k + pb] = cb * r00 - sb * r01;                        // columns
        if (qb >= 0) A[(int64_t)pa * k + qb] = sb * r00 + cb * r01;
        if (qa >= 0) {
          A[(int64_t)qa * k + pb] = cb * r10 - sb * r11;
          if (qb >= 0) A[(int64_t)qa * k + qb] = sb * r10 + cb * r11;
        }
      }
      // eigenvectors: columns p, q of Q
      for (int row = warp; row < k; row += nwarps) {
        double* qr = Q + (int64_t)row * k;
        for (int t = lane; t < npairs; t += 32) {
          const int p2 = pq[2 * t], q2 = pq[2 * t + 1];
          if (q2 < 0) continue;
          const double c = cs[2 * t], s2 = cs[2 * t + 1];
          const double x = qr[p2], y = qr[q2];
          qr[p2] = c * x - s2 * y;
          qr[q2] = s2 * x + c * y;
        }
      }
      __syncthreads();
    }
  }
  double mn = INFINITY;
  for (int t = tid; t < k; t += nt) mn = fmin(mn, A[(int64_t)t * k + t]);
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
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // symmetrise into M and work
  for (int r = warp; r < k; r += nwarps)
    for (int c = lane; c < k; c += 32)
      work[(int64_t)r * k + c] = (M[(int64_t)r * k + c] + M[(int64_t)c * k + r]) / 2;
  __syncthreads();
  for (int r = warp; r < k; r += nwarps)
    for (int c = lane; c < k; c += 32) {
      const double v = work[(int64_t)r * k + c];
      M[(int64_t)r * k + c] = v;
      work2[(int64_t)r * k + c] = (r == c) ? v - min_eig : v;
    }
  __syncthreads();
  // Fast exit: the reference returns the symmetrised matrix unchanged when its smallest
  // eigenvalue is >= min_eig, i.e. when M - min_eig*I is positive definite -- which a Cholesky
  // factorisation decides at a fraction of the cost of the eigendecomposition.  This is the
  // common case in the lookahead re-fits (small steps from a fitted covariance).
  __shared__ int psd_flag;
  if (!isnan(blk_cholesky(work2, k, red, &psd_flag))) return;
  const double mn = blk_jacobi(work, work2, k, cs, red);   // work diag = eigenvalues, work2 = Q
  if (mn < min_eig) {
    // clamped eigenvalues into cs-independent scratch: reuse the first row of `work` beyond the
    // diagonal is unsafe, so recompute fmax per term (k is small)
    for (int r = warp; r < k; r += nwarps)
      for (int c = lane; c < k; c += 32) {
        if (c < r) continue;
        double s = 0;
        for (int l = 0; l < k; ++l)
          s += fmax(work[(int64_t)l * k + l], min_eig) * work2[(int64_t)r * k + l] *
               work2[(int64_t)c * k + l];
        M[(int64_t)r * k + c] = s;
        M[(int64_t)c * k + r] = s;
      }
  }
  __syncthreads();
}

}  // namespace amf
