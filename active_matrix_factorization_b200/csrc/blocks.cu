// Scalable-mode variational posterior: the reference's KL objective (active_pmf.py:202-240)
// restricted to block-diagonal covariances, one d x d block per user row / item column, and the
// lookahead criteria of active_pmf.py:635-704 (_exp_with_rij with _approx_entropy :526-530 and
// _total_variance :605-606) evaluated by local re-fits: per (candidate, value) a rank-d update
// of the two d x d precisions the new rating touches and one Cholesky per update, all in shared
// memory / registers of one lane group -- no k x k matrix is ever formed (k = (N+M)d is 2595 at
// the drugbank config and 8e6 at the 200k x 50k one).
//
// Lane groups: G = 8, 16 or 32 lanes (d <= G) work on one row / candidate; lane l owns row l of
// the d x d matrices.  fp64 throughout: the accept tests downstream compare near-equal
// criteria.  A group's scratch lives in shared memory with row stride d+1 (odd strides keep the
// row-owner accesses conflict-free for even d).
#include <algorithm>
#include <math.h>

#include "common.cuh"

namespace amf {

int acquire_partials(Best** out, cudaStream_t s);

namespace {

constexpr int kWarpsPerCta = 4;

__device__ __forceinline__ double group_sum(double v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// In-place lower Cholesky of the d x d matrix M (row stride ld) by one lane group; lane l owns
// row l.  Returns log det M (sum of 2 log L_kk) on every lane; *ok cleared on a non-positive pivot.
__device__ __forceinline__ double group_cholesky(double* M, int ld, int d, int l, bool* ok) {
  double logdet = 0.0;
  for (int k = 0; k < d; ++k) {
    const double piv = M[k * ld + k];
    if (!(piv > 0.0)) *ok = false;
    const double r = sqrt(piv);
    logdet += 2.0 * log(r);
    __syncwarp();
    if (l == k) M[k * ld + k] = r;
    if (l > k && l < d) M[l * ld + k] /= r;
    __syncwarp();
    if (l > k && l < d) {
      const double lk = M[l * ld + k];
      for (int c = k + 1; c <= l; ++c) M[l * ld + c] = fma(-lk, M[c * ld + k], M[l * ld + c]);
    }
    __syncwarp();
  }
  return logdet;
}

// Inv = (L L^T)^-1 from the lower factor L: lane c solves for column c (forward then backward
// substitution inside its own column of Inv).
__device__ __forceinline__ void group_inverse(const double* L, int ld, double* Inv, int ldi, int d,
                                              int c) {
  if (c < d) {
    for (int r = 0; r < d; ++r) {
      double s = (r == c) ? 1.0 : 0.0;
      for (int t = 0; t < r; ++t) s = fma(-L[r * ld + t], Inv[t * ldi + c], s);
      Inv[r * ldi + c] = s / L[r * ld + r];
    }
    for (int r = d - 1; r >= 0; --r) {
      double s = Inv[r * ldi + c];
      for (int t = r + 1; t < d; ++t) s = fma(-L[t * ld + r], Inv[t * ldi + c], s);
      Inv[r * ldi + c] = s / L[r * ld + r];
    }
  }
  __syncwarp();
}

__device__ __forceinline__ double normal_cdf(double x) { return 0.5 * erfc(-x * 0.7071067811865476); }

// ---------------------------------------------------------------------------------------------
// One coordinate half-sweep: for every row i of the side
//   Lambda_i = I / prior_var + sum_{j in rated(i)} (n_j n_j^T + B_j) / sigma^2
//   h_i = sum_j r_ij n_j / sigma^2,  A_i = Lambda_i^-1,  m_i = A_i h_i,  logdet A_i
// (other_cov == NULL leaves the B_j term out: curvature of the MAP objective at fixed means).
// ---------------------------------------------------------------------------------------------
template <int G, typename RT>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
blocks_half_sweep_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                         const RT* __restrict__ val, int rows, int d, double mean_offset,
                         const double* __restrict__ other_mean, const double* __restrict__ other_cov,
                         double prior_var, double sigma_sq, double* __restrict__ prec,
                         double* __restrict__ hvec, double* __restrict__ cov,
                         double* __restrict__ mean, double* __restrict__ logdet,
                         int* __restrict__ fail) {
  extern __shared__ double smem[];
  constexpr int GPW = 32 / G;
  const int lane = threadIdx.x & 31, l = lane % G, g = lane / G;
  const int ld = d + 1;
  const int group_doubles = 2 * d * ld + 2 * d;
  double* M = smem + ((threadIdx.x >> 5) * GPW + g) * group_doubles;
  double* Inv = M + d * ld;
  double* hs = Inv + d * ld;
  const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
  const double inv_s2 = 1.0 / sigma_sq;
  const int64_t rows_pad = ((rows + GPW - 1) / GPW) * (int64_t)GPW;
  bool ok = true;
  for (int64_t row0 = group; row0 < rows_pad; row0 += ngroups) {
    const bool live = row0 < rows;
    const int64_t row = live ? row0 : rows - 1;
    double hl = 0.0;
    if (l < d) {
      for (int c = 0; c < d; ++c) M[l * ld + c] = (c == l) ? 1.0 / prior_var : 0.0;
      for (int64_t p = ptr[row]; p < ptr[row + 1]; ++p) {
        const int32_t j = idx[p];
        AMF_DBG_ASSERT(j >= 0);
        const double r = (double)val[p] - mean_offset;
        const double* nj = other_mean + (int64_t)j * d;
        const double nl = nj[l];
        hl = fma(r * inv_s2, nl, hl);
        if (other_cov) {
          const double* Bj = other_cov + (int64_t)j * d * d + l * d;
          for (int c = 0; c < d; ++c) M[l * ld + c] += (nl * nj[c] + Bj[c]) * inv_s2;
        } else {
          for (int c = 0; c < d; ++c) M[l * ld + c] = fma(nl * inv_s2, nj[c], M[l * ld + c]);
        }
      }
      hs[l] = hl;
      if (live) {
        for (int c = 0; c < d; ++c) prec[(row * d + l) * d + c] = M[l * ld + c];
        hvec[row * d + l] = hl;
      }
    }
    __syncwarp();
    const double ldet = group_cholesky(M, ld, d, l, &ok);
    group_inverse(M, ld, Inv, ld, d, l);
    if (l < d && live) {
      double mu = 0.0;
      for (int c = 0; c < d; ++c) {
        const double a = Inv[l * ld + c];
        cov[(row * d + l) * d + c] = a;
        mu = fma(a, hs[c], mu);
      }
      if (mean) mean[row * d + l] = mu;
      if (l == 0) logdet[row] = -ldet;
    }
    __syncwarp();
  }
  if (!ok && fail) atomicExch(fail, 1);
}

// SA = sum_i A_i, SMM = sum_i m_i m_i^T (one side per launch): the four d x d sums the total
// variance is a bilinear form of.  out[0 : d*d] += SA, out[d*d : 2 d*d] += SMM.
__global__ void __launch_bounds__(256)
blocks_sums_kernel(const double* __restrict__ mean, const double* __restrict__ cov, int64_t rows,
                   int d, double* __restrict__ out) {
  const int dd = d * d;
  for (int e = threadIdx.x; e < 2 * dd; e += blockDim.x) {
    const int which = e / dd, kl = e % dd, k = kl / d, c = kl % d;
    double s = 0.0;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x)
      s += which == 0 ? cov[r * dd + kl] : mean[r * d + k] * mean[r * d + c];
    atomicAdd(out + e, s);
  }
}

struct BlocksView {
  int32_t n, m, d;
  const double *mean_u, *cov_u, *prec_u, *h_u, *logdet_u;
  const double *mean_v, *cov_v, *prec_v, *h_v, *logdet_v;
  const double* sums;   // SA, SMM, SB, SNN (d*d each)
  double sigma_sq, entropy0;
};

// ---------------------------------------------------------------------------------------------
// Lookahead: for candidate (i, j) and every value v the posterior of row i and column j after
// adding the rating (i, j, v):   `rounds` x [ row i given column j ; column j given row i ]
//   Lambda_i' = Lambda_i + (n_j n_j^T + B_j)/sigma^2,  m_i' = A_i' (h_i + v n_j / sigma^2)
//   Lambda_j' = Lambda_j + (m_i' m_i'^T + A_i')/sigma^2,  n_j' = B_j' (h_j + v m_i' / sigma^2)
// criterion WHAT 0: log det cov' = entropy0 + dlogdet_i + dlogdet_j
//           WHAT 1: sum over all cells of Var[U_i.V_j] = <SA', SB' + SNN'> + <SMM', SB'>
// then score = sum_q w_q f(v_q): weight mode 1 = Delta-cdf of N(mu_c, sd_c^2) at the rating
// bounds (active_pmf.py:687-689), 2 = values are mu_c + sd_c t_q with given weights (2-sigma
// window, :691-699), 0 = no reduction (raw evals only).
// ---------------------------------------------------------------------------------------------
template <int G, int WHAT, bool MAX>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
blocks_lookahead_kernel(BlocksView bv, int rounds, int64_t ncand, const int32_t* __restrict__ ci,
                        const int32_t* __restrict__ cj, int nv, const double* __restrict__ values,
                        int weight_mode, const double* __restrict__ wb,
                        const double* __restrict__ rij_mean, const double* __restrict__ rij_sd,
                        double* __restrict__ evals, double* __restrict__ scores, int64_t index_base,
                        Best* __restrict__ part, int* __restrict__ fail) {
  extern __shared__ double smem[];
  constexpr int GPW = 32 / G;
  const int d = bv.d, ld = d + 1, dd = d * d;
  const int lane = threadIdx.x & 31, l = lane % G, g = lane / G;
  // CTA-wide: the four d x d sums; per group: M (factor scratch), S1 = A_i', S2 = B_j', and the
  // vectors n_j (current), m_i', rhs
  double* sums_s = smem;
  const int group_doubles = 3 * d * ld + 3 * d;
  double* M = smem + 4 * dd + ((threadIdx.x >> 5) * GPW + g) * group_doubles;
  double* S1 = M + d * ld;
  double* S2 = S1 + d * ld;
  double* njs = S2 + d * ld;
  double* mis = njs + d;
  double* rhs = mis + d;
  if (WHAT == 1) {
    for (int e = threadIdx.x; e < 4 * dd; e += blockDim.x) sums_s[e] = bv.sums[e];
  }
  __syncthreads();
  const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
  const double inv_s2 = 1.0 / bv.sigma_sq;
  const int64_t cand_pad = ((ncand + GPW - 1) / GPW) * (int64_t)GPW;
  Best best{0.0, -1};
  bool ok = true;
  for (int64_t c0 = group; c0 < cand_pad; c0 += ngroups) {
    const bool live = c0 < ncand;
    const int64_t c = live ? c0 : ncand - 1;
    const int32_t i = ci[c], j = cj[c];
    AMF_DBG_ASSERT((uint32_t)i < (uint32_t)bv.n && (uint32_t)j < (uint32_t)bv.m);
    const double* Lu = bv.prec_u + (int64_t)i * dd;
    const double* Lv = bv.prec_v + (int64_t)j * dd;
    const double* hu = bv.h_u + (int64_t)i * d;
    const double* hv = bv.h_v + (int64_t)j * d;
    const double* Bj0 = bv.cov_v + (int64_t)j * dd;
    const double* nj0 = bv.mean_v + (int64_t)j * d;
    const double mu_c = (weight_mode != 0) ? rij_mean[c] : 0.0;
    const double sd_c = (weight_mode != 0) ? rij_sd[c] : 1.0;
    double ld_i = 0.0;
    // rounds == 1: the row-i block does not depend on v -- factor it once per candidate
    auto row_i_block = [&](const double* nj, const double* Bj, int ldb) {
      if (l < d) {
        const double nl = nj[l];
        for (int k = 0; k < d; ++k)
          M[l * ld + k] = Lu[l * d + k] + (nl * nj[k] + Bj[l * ldb + k]) * inv_s2;
      }
      __syncwarp();
      const double det = group_cholesky(M, ld, d, l, &ok);
      group_inverse(M, ld, S1, ld, d, l);
      return -det;
    };
    if (rounds == 1) ld_i = row_i_block(nj0, Bj0, d);
    double acc = 0.0;
    for (int q = 0; q < nv; ++q) {
      const double v = (weight_mode == 2) ? fma(sd_c, values[q], mu_c) : values[q];
      double ld_j = 0.0;
      for (int round = 0; round < rounds; ++round) {
        const bool first = round == 0;
        if (rounds > 1) ld_i = row_i_block(first ? nj0 : njs, first ? Bj0 : S2, first ? d : ld);
        // m_i' = A_i' (h_i + v n_j / sigma^2)
        if (l < d) rhs[l] = fma(v * inv_s2, first ? nj0[l] : njs[l], hu[l]);
        __syncwarp();
        if (l < d) {
          double s = 0.0;
          for (int k = 0; k < d; ++k) s = fma(S1[l * ld + k], rhs[k], s);
          mis[l] = s;
        }
        __syncwarp();
        // column j given the new row i
        if (l < d) {
          const double ml = mis[l];
          for (int k = 0; k < d; ++k)
            M[l * ld + k] = Lv[l * d + k] + (ml * mis[k] + S1[l * ld + k]) * inv_s2;
        }
        __syncwarp();
        ld_j = -group_cholesky(M, ld, d, l, &ok);
        const bool last = round == rounds - 1;
        if (WHAT == 1 || !last) {
          group_inverse(M, ld, S2, ld, d, l);
          if (l < d) rhs[l] = fma(v * inv_s2, mis[l], hv[l]);
          __syncwarp();
          if (l < d) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) s = fma(S2[l * ld + k], rhs[k], s);
            njs[l] = s;
          }
          __syncwarp();
        }
      }
      double f;
      if (WHAT == 0) {
        f = bv.entropy0 + (ld_i - bv.logdet_u[i]) + (ld_j - bv.logdet_v[j]);
      } else {
        double s = 0.0;
        if (l < d) {
          const double* A0 = bv.cov_u + (int64_t)i * dd + l * d;
          const double* B0 = Bj0 + l * d;
          const double m0l = bv.mean_u[(int64_t)i * d + l], n0l = nj0[l];
          const double ml = mis[l], nl = njs[l];
          for (int k = 0; k < d; ++k) {
            const int e = l * d + k;
            const double sa = sums_s[e] + S1[l * ld + k] - A0[k];
            const double smm = sums_s[dd + e] + ml * mis[k] - m0l * bv.mean_u[(int64_t)i * d + k];
            const double sb = sums_s[2 * dd + e] + S2[l * ld + k] - B0[k];
            const double snn = sums_s[3 * dd + e] + nl * njs[k] - n0l * nj0[k];
            s = fma(sa, sb + snn, s);
            s = fma(smm, sb, s);
          }
        }
        f = group_sum(s, G);
      }
      if (evals && live && l == 0) evals[c * nv + q] = f;
      if (weight_mode == 1) {
        const double hi = (q == nv - 1) ? 1.0 : normal_cdf((wb[q + 1] - mu_c) / sd_c);
        const double lo = (q == 0) ? 0.0 : normal_cdf((wb[q] - mu_c) / sd_c);
        acc = fma(hi - lo, f, acc);
      } else if (weight_mode == 2) {
        acc = fma(wb[q], f, acc);
      }
      __syncwarp();
    }
    if (weight_mode != 0 && live && l == 0) {
      if (scores) scores[c] = acc;
      if (better<MAX>(acc, c + index_base, best.v, best.i)) { best.v = acc; best.i = c + index_base; }
    }
  }
  if (!ok && fail) atomicExch(fail, 1);
  best = block_best<MAX>(best);
  if (threadIdx.x == 0) part[blockIdx.x] = best;
}

// ---------------------------------------------------------------------------------------------
// pred_variance under the block posterior as a dot product:
//   Var_ij = <A_i, B_j + n_j n_j^T> + <m_i m_i^T, B_j>           (every term >= 0)
// packed per row into d(d+1) numbers -- symmetric halves, off-diagonals doubled on the user side:
//   user row  [ vech2(A_i)            ; vech2(m_i m_i^T) ]
//   item row  [ vech(B_j + n_j n_j^T) ; vech(B_j)        ]
// so the criterion is the same SDDMM as `pred` and runs on the same scoring kernels.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
blocks_pack_kernel(const double* __restrict__ mean, const double* __restrict__ cov, int64_t rows,
                   int d, int side, int ld_out, T* __restrict__ out) {
  const int half = d * (d + 1) / 2;
  const int64_t total = rows * (int64_t)ld_out;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / ld_out;
    const int e = (int)(t % ld_out);
    double x = 0.0;
    if (e < 2 * half) {
      const int h = e % half, second = e / half;
      // (k, c) of packed index h: row-major upper triangle
      int k = 0, rem = h;
      while (rem >= d - k) { rem -= d - k; ++k; }
      const int c = k + rem;
      const double a = cov[(r * d + k) * d + c];
      const double mm = mean[r * d + k] * mean[r * d + c];
      const double dbl = (k == c) ? 1.0 : 2.0;
      if (side == 0) x = dbl * (second ? mm : a);
      else x = second ? a : a + mm;
    }
    out[t] = (T)x;
  }
}

// norm.sf(cutoff, loc = E, scale = Var) elementwise with the arg-best (the reference passes the
// variance as the scale, active_pmf.py:438-439)
template <typename T, bool MAX>
__global__ void __launch_bounds__(256)
prob_ge_kernel(const T* __restrict__ e, const T* __restrict__ var, int64_t n, double cutoff,
               T* __restrict__ out, int64_t index_base, Best* __restrict__ part) {
  Best best{0.0, -1};
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n;
       c += (int64_t)gridDim.x * blockDim.x) {
    const double vr = (double)var[c];
    const double o = vr > 0 ? 0.5 * erfc((cutoff - (double)e[c]) / (vr * 1.4142135623730951)) : NAN;
    if (out) out[c] = (T)o;
    const double oo = (double)(T)o;
    if (better<MAX>(oo, c + index_base, best.v, best.i)) { best.v = oo; best.i = c + index_base; }
  }
  best = block_best<MAX>(best);
  if (threadIdx.x == 0) part[blockIdx.x] = best;
}

int pick_group(int d) { return d <= 8 ? 8 : (d <= 16 ? 16 : 32); }

}  // namespace
}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int amf_blocks_half_sweep(const amf_ratings_t* hc, int side, int d, const double* other_mean_d,
                          const double* other_cov_d, double prior_var, double sigma_sq,
                          double mean_offset, double* prec_d, double* h_d, double* cov_d,
                          double* mean_d, double* logdet_d, int* fail_d, void* stream) {
  amf_ratings* h = const_cast<amf_ratings*>(hc);
  AMF_REQUIRE(h && (side == 0 || side == 1), "amf_blocks_half_sweep: bad handle / side");
  AMF_REQUIRE(d >= 1 && d <= 32, "amf_blocks_half_sweep: d=%d outside 1..32", d);
  AMF_REQUIRE(other_mean_d && prec_d && h_d && cov_d && logdet_d,
              "amf_blocks_half_sweep: NULL argument");
  AMF_REQUIRE(prior_var > 0 && sigma_sq > 0, "amf_blocks_half_sweep: variances must be positive");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = ratings_compact(h, s);
  if (rc != AMF_OK) return rc;
  const int rows = side == 0 ? h->n_users : h->n_items;
  if (rows == 0) return AMF_OK;
  const int G = pick_group(d);
  const int groups_per_cta = kWarpsPerCta * (32 / G);
  const size_t smem = sizeof(double) * groups_per_cta * (2 * d * (d + 1) + 2 * d);
  const int64_t want = (rows + groups_per_cta - 1) / groups_per_cta;
  const int grid = (int)std::min<int64_t>(want, (int64_t)num_sms() * 8);
#define SWEEP(G_, RT_)                                                                          \
  do {                                                                                          \
    AMF_CUDA(cudaFuncSetAttribute(blocks_half_sweep_kernel<G_, RT_>,                            \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    blocks_half_sweep_kernel<G_, RT_><<<grid, kWarpsPerCta * 32, smem, s>>>(                    \
        h->ptr[side], h->idx[side], (const RT_*)h->val[side], rows, d, mean_offset,             \
        other_mean_d, other_cov_d, prior_var, sigma_sq, prec_d, h_d, cov_d, mean_d, logdet_d,   \
        fail_d);                                                                                \
  } while (0)
  if (h->dtype == AMF_F32) {
    if (G == 8) SWEEP(8, float); else if (G == 16) SWEEP(16, float); else SWEEP(32, float);
  } else {
    if (G == 8) SWEEP(8, double); else if (G == 16) SWEEP(16, double); else SWEEP(32, double);
  }
#undef SWEEP
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

// largest |a - b| over n doubles, written to *out (single CTA: the tables of a posterior are small)
__global__ void __launch_bounds__(1024)
blocks_maxdiff_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                      const double* __restrict__ a2, const double* __restrict__ b2, int64_t n2,
                      const int* __restrict__ fail, double* __restrict__ out) {
  __shared__ double part[32];
  double m = 0;
  for (int64_t t = threadIdx.x; t < n; t += blockDim.x) m = fmax(m, fabs(a[t] - b[t]));
  for (int64_t t = threadIdx.x; t < n2; t += blockDim.x) m = fmax(m, fabs(a2[t] - b2[t]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) { out[0] = m; out[1] = fail && *fail ? 1.0 : 0.0; }
  }
}

int amf_blocks_fit(const amf_ratings_t* hc, int d, double sigma_u_sq, double sigma_v_sq,
                   double sigma_sq, double mean_offset, int use_cov_term, int max_sweeps, double tol,
                   double* mean_u_d, double* cov_u_d, double* prec_u_d, double* h_u_d,
                   double* logdet_u_d, double* mean_v_d, double* cov_v_d, double* prec_v_d,
                   double* h_v_d, double* logdet_v_d, int* fail_d, int* sweeps_done, void* stream) {
  AMF_REQUIRE(hc && mean_u_d && cov_u_d && prec_u_d && h_u_d && logdet_u_d && mean_v_d && cov_v_d &&
                  prec_v_d && h_v_d && logdet_v_d && fail_d && sweeps_done,
              "amf_blocks_fit: NULL argument");
  AMF_REQUIRE(max_sweeps >= 0 && d >= 1 && d <= 32, "amf_blocks_fit: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nu = (int64_t)hc->n_users * d, nv = (int64_t)hc->n_items * d;
  double* old = nullptr;                    // previous means of both sides + {move, failed}
  AMF_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&old), sizeof(double) * (size_t)(nu + nv + 2), s));
  double* state_h = nullptr;
  if (cudaMallocHost(reinterpret_cast<void**>(&state_h), 2 * sizeof(double)) != cudaSuccess) {
    cudaFreeAsync(old, s);
    set_error("amf_blocks_fit: no pinned host memory");
    return AMF_ERR_CUDA;
  }
  int rc = AMF_OK, done = 0;
  for (int sw = 0; sw < max_sweeps && rc == AMF_OK; ++sw) {
    cudaMemcpyAsync(old, mean_u_d, sizeof(double) * (size_t)nu, cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(old + nu, mean_v_d, sizeof(double) * (size_t)nv, cudaMemcpyDeviceToDevice, s);
    rc = amf_blocks_half_sweep(hc, 0, d, mean_v_d, use_cov_term ? cov_v_d : nullptr, sigma_u_sq,
                               sigma_sq, mean_offset, prec_u_d, h_u_d, cov_u_d, mean_u_d, logdet_u_d,
                               fail_d, stream);
    if (rc == AMF_OK)
      rc = amf_blocks_half_sweep(hc, 1, d, mean_u_d, use_cov_term ? cov_u_d : nullptr, sigma_v_sq,
                                 sigma_sq, mean_offset, prec_v_d, h_v_d, cov_v_d, mean_v_d,
                                 logdet_v_d, fail_d, stream);
    if (rc != AMF_OK) break;
    blocks_maxdiff_kernel<<<1, 1024, 0, s>>>(mean_u_d, old, nu, mean_v_d, old + nu, nv, fail_d,
                                             old + nu + nv);
    cudaMemcpyAsync(state_h, old + nu + nv, 2 * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) { rc = AMF_ERR_CUDA; set_error("amf_blocks_fit: sweep failed"); break; }
    done = sw + 1;
    if (state_h[1] != 0.0 || state_h[0] < tol) break;   // not positive definite (caller reads fail_d) / converged
  }
  *sweeps_done = done;
  cudaFreeHost(state_h);
  cudaFreeAsync(old, s);
  return rc;
}

int amf_blocks_sums(int64_t rows, int d, const double* mean_d, const double* cov_d,
                    double* out_d, void* stream) {
  AMF_REQUIRE(rows >= 0 && d >= 1 && mean_d && cov_d && out_d, "amf_blocks_sums: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  AMF_CUDA(cudaMemsetAsync(out_d, 0, sizeof(double) * 2 * d * d, s));
  if (rows == 0) return AMF_OK;
  const int grid = (int)std::min<int64_t>(rows, (int64_t)num_sms() * 4);
  blocks_sums_kernel<<<grid, 256, 0, s>>>(mean_d, cov_d, rows, d, out_d);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

int amf_blocks_lookahead(const amf_blocks_view_t* v, int what, int rounds, int64_t ncand,
                         const int32_t* ci_d, const int32_t* cj_d, int nv, const double* values_d,
                         int weight_mode, const double* bounds_or_weights_d,
                         const double* rij_mean_d, const double* rij_sd_d, double* evals_d,
                         double* scores_d, int maximize, int64_t index_base, amf_best_t* best_d,
                         int* fail_d, void* stream) {
  AMF_REQUIRE(v && v->d >= 1 && v->d <= 32, "amf_blocks_lookahead: d outside 1..32");
  AMF_REQUIRE(what == AMF_LOOK_ENTROPY || what == AMF_LOOK_TOTAL_VARIANCE,
              "amf_blocks_lookahead: unknown criterion %d", what);
  AMF_REQUIRE(rounds >= 1 && nv >= 1 && ncand >= 0 && values_d && best_d,
              "amf_blocks_lookahead: bad arguments");
  AMF_REQUIRE(weight_mode >= 0 && weight_mode <= 2, "amf_blocks_lookahead: bad weight mode");
  AMF_REQUIRE(weight_mode == 0 || ncand == 0 || (bounds_or_weights_d && rij_mean_d && rij_sd_d),
              "amf_blocks_lookahead: weights need the R_ij distribution");
  AMF_REQUIRE(v->mean_u && v->cov_u && v->prec_u && v->h_u && v->logdet_u && v->mean_v &&
                  v->cov_v && v->prec_v && v->h_v && v->logdet_v,
              "amf_blocks_lookahead: incomplete view");
  AMF_REQUIRE(what == AMF_LOOK_ENTROPY || v->sums, "amf_blocks_lookahead: total variance needs sums");
  cudaStream_t s = (cudaStream_t)stream;
  Best* part = nullptr;
  int rc = acquire_partials(&part, s);
  if (rc != AMF_OK) return rc;
  const int d = v->d, G = pick_group(d);
  const int groups_per_cta = kWarpsPerCta * (32 / G);
  const size_t smem = sizeof(double) * (4 * d * d + groups_per_cta * (3 * d * (d + 1) + 3 * d));
  const int64_t want = std::max<int64_t>(1, (ncand + groups_per_cta - 1) / groups_per_cta);
  const int grid = (int)std::min<int64_t>(want, (int64_t)num_sms() * 8);
  BlocksView bv{v->n, v->m, v->d, v->mean_u, v->cov_u, v->prec_u, v->h_u, v->logdet_u,
                v->mean_v, v->cov_v, v->prec_v, v->h_v, v->logdet_v, v->sums, v->sigma_sq,
                v->entropy0};
  if (ncand == 0) {
    cudaFreeAsync(part, s);
    const amf_best_t none{0.0, -1};
    AMF_CUDA(cudaMemcpyAsync(best_d, &none, sizeof(none), cudaMemcpyHostToDevice, s));
    AMF_CUDA(cudaStreamSynchronize(s));
    return AMF_OK;
  }
#define LOOK(G_, W_, M_)                                                                          \
  do {                                                                                            \
    cudaError_t e__ = cudaFuncSetAttribute(blocks_lookahead_kernel<G_, W_, M_>,                   \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e__ != cudaSuccess) { cudaFreeAsync(part, s); AMF_CUDA(e__); }                            \
    blocks_lookahead_kernel<G_, W_, M_><<<grid, kWarpsPerCta * 32, smem, s>>>(                    \
        bv, rounds, ncand, ci_d, cj_d, nv, values_d, weight_mode, bounds_or_weights_d,            \
        rij_mean_d, rij_sd_d, evals_d, scores_d, index_base, part, fail_d);                       \
  } while (0)
#define LOOK_G(G_)                                                                                \
  do {                                                                                            \
    if (what == 0) { if (maximize) LOOK(G_, 0, true); else LOOK(G_, 0, false); }                  \
    else { if (maximize) LOOK(G_, 1, true); else LOOK(G_, 1, false); }                            \
  } while (0)
  if (G == 8) LOOK_G(8); else if (G == 16) LOOK_G(16); else LOOK_G(32);
#undef LOOK_G
#undef LOOK
  if (cudaGetLastError() != cudaSuccess) { cudaFreeAsync(part, s); AMF_LAUNCH_CHECK(); }
  return launch_best_final(part, grid, maximize != 0, best_d, s);
}

int amf_blocks_pack(int dtype, int64_t rows, int d, const double* mean_d, const double* cov_d,
                    int side, int ld_out, void* out_d, void* stream) {
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_blocks_pack: bad dtype");
  AMF_REQUIRE(rows >= 0 && d >= 1 && mean_d && cov_d && out_d && (side == 0 || side == 1),
              "amf_blocks_pack: bad arguments");
  AMF_REQUIRE(ld_out >= d * (d + 1), "amf_blocks_pack: ld_out=%d < d(d+1)=%d", ld_out, d * (d + 1));
  if (rows == 0) return AMF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t total = rows * (int64_t)ld_out;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16);
  if (dtype == AMF_F32)
    blocks_pack_kernel<float><<<grid, 256, 0, s>>>(mean_d, cov_d, rows, d, side, ld_out, (float*)out_d);
  else
    blocks_pack_kernel<double><<<grid, 256, 0, s>>>(mean_d, cov_d, rows, d, side, ld_out, (double*)out_d);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

int amf_prob_ge(int dtype, int64_t n, const void* mean_d, const void* var_d, double cutoff,
                void* out_d, int maximize, int64_t index_base, amf_best_t* best_d, void* stream) {
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_prob_ge: bad dtype");
  AMF_REQUIRE(n >= 0 && mean_d && var_d && best_d, "amf_prob_ge: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  Best* part = nullptr;
  int rc = acquire_partials(&part, s);
  if (rc != AMF_OK) return rc;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 8));
#define PG(T_, M_) prob_ge_kernel<T_, M_><<<grid, 256, 0, s>>>((const T_*)mean_d, (const T_*)var_d, n, cutoff, (T_*)out_d, index_base, part)
  if (dtype == AMF_F32) { if (maximize) PG(float, true); else PG(float, false); }
  else { if (maximize) PG(double, true); else PG(double, false); }
#undef PG
  if (cudaGetLastError() != cudaSuccess) { cudaFreeAsync(part, s); AMF_LAUNCH_CHECK(); }
  return launch_best_final(part, grid, maximize != 0, best_d, s);
}

#pragma GCC visibility pop
}
