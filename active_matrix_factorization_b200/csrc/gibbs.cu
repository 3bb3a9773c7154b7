// Bayesian PMF: batched row/column conditionals of the Gibbs sampler and streaming sample
// statistics over posterior samples.
//
//   amf_gibbs_half_sweep : bayes_pmf.py:189-216 sample_feature for every row of one side
//                          (the loops at :286-292 / :294-300): one warp per row for d <= 32,
//                          one CTA per row above.
//   amf_bayes_sample_stats : bayes_pmf.py:433-455 predict / pred_variance and :528-538
//                          prob_ge_cutoff without materialising S dense N x M matrices.
//
// The reference draws z ~ N(0, I_d) per row from numpy's global stream, in row order; the host
// passes those draws in (z_d) so a seeded chain reproduces the reference's samples.  As in the
// reference the sample is  chol(inv(Lambda)) z + mean  (lower factor of the COVARIANCE), not the
// cheaper  Lambda = R R', x = mean + R^-T z, which has the same law but different values.
#include <algorithm>

#include "common.cuh"
#include "philox.cuh"

namespace amf {

constexpr int GIBBS_THREADS = 128;
constexpr int GIBBS_TILE = 32;   // rated rows staged per tile
constexpr int GIBBS_MAXD = 64;
// the Gram matrix F'F is accumulated in registers as 2 x 2 blocks of its upper triangle:
// (d/2)(d/2+1)/2 blocks, GIBBS_MAXBLK per thread at the largest d
constexpr int GIBBS_MAXBLK = ((GIBBS_MAXD / 2) * (GIBBS_MAXD / 2 + 1) / 2 + GIBBS_THREADS - 1) / GIBBS_THREADS;

// ---- CTA-wide dense helpers (d > 32: one CTA per row, matrices in shared memory) -------------
// in-place lower Cholesky of the d x d matrix A (leading dimension lda) in shared memory.
// Returns false (to all threads of the group) if a pivot is not positive.
__device__ bool chol_lower(double* A, int d, int lda, int* flag) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) *flag = 1;
  __syncthreads();
  for (int j = 0; j < d; ++j) {
    if (tid == 0) {
      double pv = A[j * lda + j];
      if (!(pv > 0.0)) *flag = 0;
      A[j * lda + j] = sqrt(pv);
    }
    __syncthreads();
    const double inv = 1.0 / A[j * lda + j];
    for (int i = j + 1 + tid; i < d; i += nt) A[i * lda + j] *= inv;
    __syncthreads();
    const int rem = d - j - 1;
    for (int t = tid; t < rem * rem; t += nt) {
      const int i = j + 1 + t / rem, k = j + 1 + t % rem;
      if (k <= i) A[i * lda + k] -= A[i * lda + j] * A[k * lda + j];
    }
    __syncthreads();
  }
  for (int t = tid; t < d * d; t += nt) {
    const int i = t / d, k = t % d;
    if (k > i) A[i * lda + k] = 0.0;
  }
  __syncthreads();
  return *flag != 0;
}

// Linv = inverse of the lower-triangular L (both d x d in shared memory, distinct buffers).
// Four lanes cooperate on one column (forward substitution L x = e_c): the inner product of
// row i is split over the lanes and reduced with two shuffles, so a column costs ~d short steps
// instead of ~d^2/2 dependent shared-memory round trips.
__device__ void tri_inverse_lower(const double* L, double* Linv, int d, int lda) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int sub = tid & 3;
  const unsigned mask = 0xFu << ((tid & 31) & ~3);
  for (int c = tid >> 2; c < d; c += nt >> 2) {
    for (int i = sub; i < c; i += 4) Linv[i * lda + c] = 0.0;
    for (int i = c; i < d; ++i) {
      double s = 0;
      for (int k = c + sub; k < i; k += 4) s += L[i * lda + k] * Linv[k * lda + c];
      s += __shfl_xor_sync(mask, s, 1);
      s += __shfl_xor_sync(mask, s, 2);
      if (sub == 0) Linv[i * lda + c] = ((i == c ? 1.0 : 0.0) - s) / L[i * lda + i];
      __syncwarp(mask);
    }
  }
  __syncthreads();
}

// Given Lambda (A, d x d in shared memory) and rhs: cov = Lambda^-1 via Cholesky
// (Lambda = R R', cov = R^-T R^-1), mean = cov rhs, L = chol(cov), out = L z + mean
// (bayes_pmf.py:205-216: the reference inverts and factors the covariance, not Lambda).
template <typename T>
__device__ bool solve_and_sample(double* A, double* Bm, double* rhs, double* mean, int d, int lda,
                                 int* flag, const T* __restrict__ z, T* __restrict__ out) {
  const int tid = threadIdx.x, nt = blockDim.x;
  bool ok = chol_lower(A, d, lda, flag);
  tri_inverse_lower(A, Bm, d, lda);      // Bm = R^-1
  for (int t = tid; t < d * d; t += nt) {
    const int k = t / d, l = t % d;
    double s = 0;
    for (int q = max(k, l); q < d; ++q) s += Bm[q * lda + k] * Bm[q * lda + l];
    A[k * lda + l] = s;                        // cov
  }
  __syncthreads();
  if (tid < d) {
    double s = 0;
    for (int l = 0; l < d; ++l) s += A[tid * lda + l] * rhs[l];
    mean[tid] = s;
  }
  __syncthreads();
  ok = chol_lower(A, d, lda, flag) && ok;   // A = lower chol of cov
  if (tid < d) {
    double s = mean[tid];
    for (int l = 0; l <= tid; ++l) s += A[tid * lda + l] * (double)z[l];
    out[tid] = (T)s;
  }
  return ok;
}

template <typename T>
__global__ void __launch_bounds__(GIBBS_THREADS)
gibbs_rows_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                  const T* __restrict__ val, int row_begin, int rows, int d,
                  const T* __restrict__ other,
                  const T* __restrict__ alpha, const T* __restrict__ mu, double beta,
                  double mean_offset, const T* __restrict__ z, T* __restrict__ out,
                  int* __restrict__ fail) {
  extern __shared__ double smem[];
  const int lda = d + 1;
  const int dp = (d + 1) & ~1;      // d rounded up to even: the tile is read two columns at a time
  const int ldt = dp + 2;           // even row stride keeps the 2-element loads aligned
  const int hb = dp / 2;            // 2 x 2 blocks per side
  const int nblk = hb * (hb + 1) / 2;
  double* A = smem;                 // d x lda : Lambda -> chol -> ...
  double* Bm = A + d * lda;         // d x lda : scratch
  double* rhs = Bm + d * lda;       // d
  double* mean = rhs + d;           // d
  double* tile_r = mean + d;        // GIBBS_TILE
  T* tile = reinterpret_cast<T*>(tile_r + GIBBS_TILE);   // GIBBS_TILE x ldt
  __shared__ int flag;
  const int tid = threadIdx.x;

  // this thread's blocks (bk <= bl) of the upper triangle
  int bks[GIBBS_MAXBLK], bls[GIBBS_MAXBLK];
#pragma unroll
  for (int a = 0; a < GIBBS_MAXBLK; ++a) {
    int blk = tid + a * GIBBS_THREADS, bk = 0;
    if (blk < nblk) {
      while (blk >= hb - bk) { blk -= hb - bk; ++bk; }
      bks[a] = bk; bls[a] = bk + blk;
    } else {
      bks[a] = -1; bls[a] = 0;
    }
  }

  for (int row = row_begin + blockIdx.x; row < rows; row += gridDim.x) {
    const int64_t p0 = ptr[row], p1 = ptr[row + 1];
    // ---- Gram matrix F'F (2 x 2 register blocks) and F'(r - offset) ------------------------
    T acc[GIBBS_MAXBLK][4];
#pragma unroll
    for (int a = 0; a < GIBBS_MAXBLK; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0;
    T racc = 0;                                  // thread k < d accumulates rhs[k]
    for (int64_t base = p0; base < p1; base += GIBBS_TILE) {
      const int cnt = (int)min((int64_t)GIBBS_TILE, p1 - base);
      for (int t = tid; t < cnt * dp; t += GIBBS_THREADS) {
        const int e = t / dp, k = t - e * dp;
        tile[e * ldt + k] = k < d ? other[(int64_t)idx[base + e] * d + k] : T(0);
      }
      for (int t = tid; t < cnt; t += GIBBS_THREADS) tile_r[t] = (double)val[base + t] - mean_offset;
      __syncthreads();
#pragma unroll
      for (int a = 0; a < GIBBS_MAXBLK; ++a) {
        if (bks[a] >= 0) {
          const T* pa = tile + 2 * bks[a];
          const T* pb = tile + 2 * bls[a];
          T s00 = acc[a][0], s01 = acc[a][1], s10 = acc[a][2], s11 = acc[a][3];
          for (int e = 0; e < cnt; ++e) {
            const T a0 = pa[e * ldt], a1 = pa[e * ldt + 1];
            const T b0 = pb[e * ldt], b1 = pb[e * ldt + 1];
            s00 = fma(a0, b0, s00); s01 = fma(a0, b1, s01);
            s10 = fma(a1, b0, s10); s11 = fma(a1, b1, s11);
          }
          acc[a][0] = s00; acc[a][1] = s01; acc[a][2] = s10; acc[a][3] = s11;
        }
      }
      if (tid < d) {
        T s = racc;
        for (int e = 0; e < cnt; ++e) s = fma(tile[e * ldt + tid], (T)tile_r[e], s);
        racc = s;
      }
      __syncthreads();
    }
    // ---- Lambda = alpha + beta F'F ; rhs = beta F'r + alpha mu ---------------------------
#pragma unroll
    for (int a = 0; a < GIBBS_MAXBLK; ++a) {
      if (bks[a] >= 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = 2 * bks[a] + (q >> 1), l = 2 * bls[a] + (q & 1);
          if (k < d && l < d) {
            A[k * lda + l] = (double)alpha[k * d + l] + beta * (double)acc[a][q];
            if (k != l && bks[a] != bls[a])
              A[l * lda + k] = (double)alpha[l * d + k] + beta * (double)acc[a][q];
          }
        }
      }
    }
    if (tid < d) {
      double s = beta * (double)racc;
      for (int l = 0; l < d; ++l) s += (double)alpha[tid * d + l] * (double)mu[l];
      rhs[tid] = s;
    }
    __syncthreads();
    const bool ok = solve_and_sample<T>(A, Bm, rhs, mean, d, lda, &flag, z + (int64_t)row * d,
                                        out + (int64_t)row * d);
    if (!ok && tid == 0) atomicExch(fail, 1);
    __syncthreads();
  }
}

// ---- fp32, d = 32 or 64: the Gram matrix F'F on the tensor cores ----------------------------
// F'F over the rated rows is a dense (d x n_i) x (n_i x d) contraction; with d >= 32 it fills
// whole m16n8k8 TF32 MMA tiles.  fp32 accuracy is kept with the 3xTF32 split
// (x = hi + lo; hi*hi + hi*lo + lo*hi, fp32 accumulate).  The four warps of the CTA take the
// 8-rating k-steps of a staged tile in turn and keep the upper-triangle tiles of the Gram
// matrix in registers for the whole row; partial sums meet in shared memory once per row.
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int GIBBS_TC_TILE = 64;   // rated rows staged per tile (8 k-steps)

template <int D>
__global__ void __launch_bounds__(GIBBS_THREADS)
gibbs_rows_tc_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                     const float* __restrict__ val, int row_begin, int rows,
                     const float* __restrict__ other, const float* __restrict__ alpha,
                     const float* __restrict__ mu, double beta, double mean_offset,
                     const float* __restrict__ z, float* __restrict__ out, int* __restrict__ fail) {
  constexpr int d = D, lda = D + 1;
  constexpr int LDT = D + 8;               // == 8 (mod 32): the fragment loads hit 32 banks
  constexpr int MT = D / 16, NT8 = D / 8;  // 16-row and 8-column tiles per side
  constexpr int NTILES = MT * (MT + 1);    // tiles (M, N >= 2M) covering the upper triangle
  constexpr int NW = GIBBS_THREADS / 32;
  extern __shared__ double smem[];
  double* A = smem;                        // d x lda
  double* Bm = A + d * lda;                // d x lda
  double* rhs = Bm + d * lda;              // d
  double* mean = rhs + d;                  // d
  float* tile_r = reinterpret_cast<float*>(mean + d);   // GIBBS_TC_TILE
  float* tile = tile_r + GIBBS_TC_TILE;                   // GIBBS_TC_TILE x LDT (16-byte aligned)
  float* part = tile;        // NW x NTILES x 128 partial Gram tiles, after the last staged tile
  __shared__ int flag;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, tig = lane & 3;

  for (int row = row_begin + blockIdx.x; row < rows; row += gridDim.x) {
    const int64_t p0 = ptr[row], p1 = ptr[row + 1];
    float c[NTILES][4];
#pragma unroll
    for (int t = 0; t < NTILES; ++t) c[t][0] = c[t][1] = c[t][2] = c[t][3] = 0.f;
    float racc = 0.f;
    for (int64_t base = p0; base < p1; base += GIBBS_TC_TILE) {
      const int cnt = (int)min((int64_t)GIBBS_TC_TILE, p1 - base);
      const int cnt8 = (cnt + 7) & ~7;     // rows staged: whole k-steps, zero-filled past cnt
      for (int t = tid; t < cnt8 * (D / 4); t += GIBBS_THREADS) {
        const int e = t / (D / 4), q = t - e * (D / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < cnt) v = __ldg(reinterpret_cast<const float4*>(other + (int64_t)idx[base + e] * D) + q);
        *reinterpret_cast<float4*>(tile + e * LDT + 4 * q) = v;
      }
      for (int t = tid; t < cnt; t += GIBBS_THREADS)
        tile_r[t] = (float)((double)val[base + t] - mean_offset);
      __syncthreads();
      for (int e0 = w * 8; e0 < cnt8; e0 += NW * 8) {
        uint32_t xh[NT8], xl[NT8], yh[NT8], yl[NT8];
#pragma unroll
        for (int j = 0; j < NT8; ++j) {
          const float x = tile[(e0 + tig) * LDT + 8 * j + g];
          const float y = tile[(e0 + tig + 4) * LDT + 8 * j + g];
          xh[j] = to_tf32(x); xl[j] = to_tf32(x - __uint_as_float(xh[j]));
          yh[j] = to_tf32(y); yl[j] = to_tf32(y - __uint_as_float(yh[j]));
        }
        int t = 0;
#pragma unroll
        for (int M = 0; M < MT; ++M) {
          const uint32_t ah[4] = {xh[2 * M], xh[2 * M + 1], yh[2 * M], yh[2 * M + 1]};
          const uint32_t al[4] = {xl[2 * M], xl[2 * M + 1], yl[2 * M], yl[2 * M + 1]};
#pragma unroll
          for (int N = 2 * M; N < NT8; ++N, ++t) {
            mma_tf32(c[t], al, xh[N], yh[N]);
            mma_tf32(c[t], ah, xl[N], yl[N]);
            mma_tf32(c[t], ah, xh[N], yh[N]);
          }
        }
      }
      if (tid < d) {
        float s = racc;
        for (int e = 0; e < cnt; ++e) s = fmaf(tile[e * LDT + tid], tile_r[e], s);
        racc = s;
      }
      __syncthreads();
    }
    // ---- partial Gram tiles of the four warps -> Lambda = alpha + beta F'F ------------------
#pragma unroll
    for (int t = 0; t < NTILES; ++t)
      *reinterpret_cast<float4*>(part + ((w * NTILES + t) * 32 + lane) * 4) =
          make_float4(c[t][0], c[t][1], c[t][2], c[t][3]);
    __syncthreads();
    for (int e = tid; e < d * d; e += GIBBS_THREADS) {
      const int k = e / d, l = e - k * d;
      if (k > l) continue;
      const int M = k >> 4, N = l >> 3;                     // N >= 2M because l >= k
      const int t = M * NT8 - M * (M - 1) + (N - 2 * M);    // tiles before row M: sum (NT8 - 2m)
      const int rr = k & 15, cc = l & 7;
      const int src = ((rr & 7) * 4 + (cc >> 1)) * 4 + 2 * (rr >> 3) + (cc & 1);
      float gsum = 0.f;
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) gsum += part[(ww * NTILES + t) * 128 + src];
      A[k * lda + l] = (double)alpha[k * d + l] + beta * (double)gsum;
      if (k != l) A[l * lda + k] = (double)alpha[l * d + k] + beta * (double)gsum;
    }
    if (tid < d) {
      double s = beta * (double)racc;
      for (int l = 0; l < d; ++l) s += (double)alpha[tid * d + l] * (double)mu[l];
      rhs[tid] = s;
    }
    __syncthreads();
    const bool ok = solve_and_sample<float>(A, Bm, rhs, mean, d, lda, &flag, z + (int64_t)row * d,
                                            out + (int64_t)row * d);
    if (!ok && tid == 0) atomicExch(fail, 1);
    __syncthreads();
  }
}

// ---- register-resident warp solve (d <= 32, padded to 32 with the identity) -----------------
// Run by one warp, shared-memory formulations of the solve spend ~36k instructions per row on
// index arithmetic and predicated read-modify-write loops.  Here a lane keeps ITS row (Cholesky) or column (inverse)
// of the 32 x 32 matrix in registers; the one vector every lane needs per step (the current
// Cholesky column, a row of L, a row of L^-1) goes through shared memory as a broadcast read.
// All loops are fully unrolled (static register indices): ~4k instructions per row.
constexpr int S32 = 34;                     // row stride of the 32 x 32 scratch (16-byte aligned rows)

// fp64 sqrt / divide expand to ~40 instructions each; the unrolled solve needs 128 of them, so
// they are kept out of line (scalar arguments: no register arrays are forced to memory)
__device__ __noinline__ double sqrt_then_inv(double pv, double* inv) {
  const double r = sqrt(pv);
  *inv = 1.0 / r;
  return r;
}
__device__ __noinline__ double div_f64(double a, double b) { return a / b; }

// in-place lower Cholesky; lane i holds row i in a[0..31] (entries right of the diagonal are
// don't-care on entry and junk on exit).  col: 2 x 32 doubles of shared memory.
__device__ __forceinline__ bool chol32_rows(double (&a)[32], double* col, int lane) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double pv = __shfl_sync(0xffffffffu, a[j], j);
    ok = ok & (pv > 0.0);                             // no short circuit: straight-line code
    double inv;
    const double r = sqrt_then_inv(pv, &inv);
    double lij = lane == j ? r : a[j] * inv;
    if (lane < j) lij = 0.0;
    a[j] = lij;
    double* cb = col + (j & 1) * 32;                  // double-buffered: one sync per column
    cb[lane] = lij;
    __syncwarp();
#pragma unroll
    for (int k = j + 1; k < 32; ++k) a[k] = fma(-lij, cb[k], a[k]);   // rows above k: junk, unused
  }
  return ok;
}

// out = chol(inv(Lambda)) z + inv(Lambda) rhs for the row whose Lambda (d x d, leading dimension
// lda) sits in shared memory at Lam.  scr: 32 x S32 doubles (may alias Lam), col: 2 x 32, vec: 32.
template <typename T>
__device__ bool warp_solve32(const double* Lam, int d, int lda, double* scr, double* col, double* vec,
                             double rhs_lane, const T* __restrict__ z, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  double a[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) a[k] = (lane < d && k < d) ? Lam[lane * lda + k] : (k == lane ? 1.0 : 0.0);
  __syncwarp();                                        // Lam may be overwritten from here on
  bool ok = chol32_rows(a, col, lane);                 // Lambda = R R'
  // R (row per lane) -> scratch, then R^-1 by forward substitution, lane c owning column c
#pragma unroll
  for (int k = 0; k < 32; ++k) scr[lane * S32 + k] = k <= lane ? a[k] : 0.0;
  __syncwarp();
  double x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    double s0 = 0, s1 = 0;
#pragma unroll
    for (int k = 0; k + 1 < i; k += 2) {
      const double2 l2 = *reinterpret_cast<const double2*>(scr + i * S32 + k);
      s0 = fma(l2.x, x[k], s0);
      s1 = fma(l2.y, x[k + 1], s1);
    }
    if (i & 1) s0 = fma(scr[i * S32 + i - 1], x[i - 1], s0);
    x[i] = div_f64((i == lane ? 1.0 : 0.0) - (s0 + s1), scr[i * S32 + i]);   // 0 above the diagonal
  }
  __syncwarp();
  // R^-1 (column per lane) -> scratch rows; cov = R^-T R^-1, lane k owning row k
#pragma unroll
  for (int q = 0; q < 32; ++q) scr[q * S32 + lane] = x[q];
  vec[lane] = rhs_lane;
  __syncwarp();
  double c[32];
#pragma unroll
  for (int l = 0; l < 32; ++l) c[l] = 0.0;
#pragma unroll
  for (int l = 0; l < 32; l += 2) {
#pragma unroll
    for (int q = l; q < 32; ++q) {                     // terms with q < max(k, l) are zeros
      const double2 b2 = *reinterpret_cast<const double2*>(scr + q * S32 + l);
      c[l] = fma(x[q], b2.x, c[l]);
      c[l + 1] = fma(x[q], b2.y, c[l + 1]);            // b2.y = 0 at q == l
    }
  }
  double mean = 0;
#pragma unroll
  for (int l = 0; l < 32; ++l) mean = fma(c[l], vec[l], mean);
  __syncwarp();
  vec[lane] = lane < d ? (double)z[lane] : 0.0;
  ok = chol32_rows(c, col, lane) & ok;                 // cov = L L'
  __syncwarp();
  double sres = mean;
#pragma unroll
  for (int l = 0; l < 32; ++l) sres = fma(l <= lane ? c[l] : 0.0, vec[l], sres);
  if (lane < d) out[lane] = (T)sres;
  return ok;
}

// ---- fast mode: device random numbers (philox.cuh) and one factorisation ------------------------
template <typename T>
__global__ void __launch_bounds__(256)
philox_normal_kernel(unsigned long long seed, unsigned long long stream, int64_t rows, int d,
                     T* __restrict__ out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < rows * d;
       t += (int64_t)gridDim.x * blockDim.x)
    out[t] = (T)philox_normal(seed, stream, (uint32_t)(t / d), (uint32_t)(t % d));
}

// Fast-mode row sample: Lambda = R R', sample = R^-T (R^-1 rhs + z) ~ N(Lambda^-1 rhs, Lambda^-1).
// One Cholesky and two triangular solves instead of the reference's inv + second Cholesky
// (bayes_pmf.py:208-216) -- the same distribution, a different map from z to the sample.
template <typename T>
__device__ bool warp_solve32_fast(const double* Lam, int d, int lda, double* scr, double* col,
                                  double rhs_lane, double z_lane, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  double a[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) a[k] = (lane < d && k < d) ? Lam[lane * lda + k] : (k == lane ? 1.0 : 0.0);
  __syncwarp();
  const bool ok = chol32_rows(a, col, lane);
  // lane i holds row i of R; forward substitution R y = rhs, then w = y + z
  double diag = 1.0;
#pragma unroll
  for (int k = 0; k < 32; ++k) if (k == lane) diag = a[k];
  const double inv_diag = 1.0 / diag;
  double sacc = rhs_lane;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const double yk = __shfl_sync(0xffffffffu, sacc * inv_diag, k);   // final once lanes < k are done
    if (lane > k) sacc = fma(-a[k], yk, sacc);
  }
  double w = sacc * inv_diag + (lane < d ? z_lane : 0.0);
  // R' x = w: row k of R is column k of R'; rows go through the scratch
#pragma unroll
  for (int k = 0; k < 32; ++k) scr[lane * S32 + k] = k <= lane ? a[k] : 0.0;
  __syncwarp();
#pragma unroll
  for (int k = 31; k >= 0; --k) {
    const double xk = __shfl_sync(0xffffffffu, w * inv_diag, k);
    if (lane < k) w = fma(-scr[k * S32 + lane], xk, w);
  }
  if (lane < d) out[lane] = (T)(w * inv_diag);
  return ok;
}

// ---- d <= 32: one WARP per row ---------------------------------------------------------------
// The per-row work after the Gram matrix is a chain of ~100 short dependent steps (two Cholesky
// factorisations, a triangular inverse) on a matrix of at most 32 x 32: with a CTA per row the
// time goes to block barriers.  Here every warp owns a row, solves it in registers
// (warp_solve32) and synchronises with __syncwarp only; the four warps of a CTA never meet.
//   TC (fp32, d == 32): Gram matrix on the tensor cores as in gibbs_rows_tc_kernel, all k-steps
//   of the row taken by the one warp; otherwise 2 x 2 register blocks spread over the 32 lanes.
constexpr int GIBBS_WARP_MAXBLK = (16 * 17 / 2 + 31) / 32;   // 2 x 2 blocks per lane at d = 32

// A batch of chains over the same rating list, chain p with ONE extra rating (the lookahead of
// bayes_pmf.py:560-598: model + (i, j, v)): its own factor matrices, hyper-parameters and mean
// offset.  count <= 1 with NULL arrays is the plain half-sweep.
struct GibbsBatch {
  int count;                  // chains (>= 1)
  int other_rows;             // rows of `other` per chain
  const int32_t* ex_row;      // row of the side being sampled that holds chain p's extra rating
  const int32_t* ex_col;      // row of `other` it pairs with
  const double* ex_val;       // its value
  const double* offsets;      // mean offset of chain p (NULL: the common one)
};

// MINB = resident CTAs per SM the register budget is set for: 3 (168 registers, no spills) is the
// faster code for one wave of rows (C4: 2625 rows, 0.108 against 0.132 ms per sweep), 4 (128
// registers, ~0.9 KB of spills, 16 warps per SM) wins when the rows fill the GPU many times over
// (C5 user side, fast / parity: 6.80 / 12.19 against 7.60 / 12.65 ms; 5 and 6 are far slower).
template <typename T, bool TC, bool FAST, int MINB>
__global__ void __launch_bounds__(GIBBS_THREADS, MINB)
gibbs_rows_warp_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                       const T* __restrict__ val, int row_begin, int rows, int d,
                       const T* __restrict__ other, const T* __restrict__ alpha,
                       const T* __restrict__ mu, double beta, double mean_offset,
                       const T* __restrict__ z, T* __restrict__ out, int* __restrict__ fail,
                       int a_doubles, int warp_doubles, unsigned long long seed,
                       unsigned long long stream_id, GibbsBatch gb) {
  extern __shared__ double smem[];
  constexpr int NW = GIBBS_THREADS / 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int lda = d + 1;
  // per warp: [ staged tile | Lambda (d x lda) | 32 x S32 scratch of the solve ] share one region
  // (each is dead before the next is written), then the column / vector buffers of the solve
  double* A = smem + (size_t)w * warp_doubles;
  double* col = A + a_doubles;                      // 2 x 32
  double* vec = col + 64;                           // 32
  T* tile = reinterpret_cast<T*>(A);
  const int dp = (d + 1) & ~1;
  const int ldt = TC ? 40 : dp + 2;                 // TC: == 8 (mod 32), conflict-free fragments
  const int hb = dp / 2, nblk = hb * (hb + 1) / 2;
  const int g = lane >> 2, tig = lane & 3;

  int bks[GIBBS_WARP_MAXBLK], bls[GIBBS_WARP_MAXBLK];
  if (!TC) {
#pragma unroll
    for (int a = 0; a < GIBBS_WARP_MAXBLK; ++a) {
      int blk = lane + a * 32, bk = 0;
      if (blk < nblk) {
        while (blk >= hb - bk) { blk -= hb - bk; ++bk; }
        bks[a] = bk; bls[a] = bk + blk;
      } else {
        bks[a] = -1; bls[a] = 0;
      }
    }
  }

  const int span = rows - row_begin;
  const int64_t tasks = (int64_t)span * (gb.count > 1 ? gb.count : 1);
  const T* const other0 = other;
  const T* const alpha0 = alpha;
  const T* const mu0 = mu;
  T* const out0 = out;
  const double mean_offset0 = mean_offset;
#pragma unroll 1
  for (int64_t task = (int64_t)blockIdx.x * NW + w; task < tasks; task += (int64_t)gridDim.x * NW) {
    const int chain = (int)(task / span);
    const int row = row_begin + (int)(task - (int64_t)chain * span);
    if (gb.count > 1) {
      other = other0 + (int64_t)chain * gb.other_rows * d;
      alpha = alpha0 + (int64_t)chain * d * d;
      mu = mu0 + (int64_t)chain * d;
      out = out0 + (int64_t)chain * rows * d;
      if (gb.offsets) mean_offset = gb.offsets[chain];
    }
    (void)mean_offset0;
    const int64_t p0 = ptr[row], p1 = ptr[row + 1];
    T acc[GIBBS_WARP_MAXBLK][4];
    float c[6][4];
#pragma unroll
    for (int a = 0; a < GIBBS_WARP_MAXBLK; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0;
#pragma unroll
    for (int t = 0; t < 6; ++t) c[t][0] = c[t][1] = c[t][2] = c[t][3] = 0.f;
    T racc = 0;                                      // lane k < d accumulates rhs[k]
    // lane e holds index and rating of the tile's e-th entry; the next tile's are loaded one
    // tile ahead so that only the gather of the rated rows waits on memory inside a tile
    int32_t jnext = p0 + lane < p1 ? idx[p0 + lane] : 0;
    T vnext = p0 + lane < p1 ? val[p0 + lane] : T(0);
#pragma unroll 1
    for (int64_t base = p0; base < p1; base += 32) {
      const int cnt = (int)min((int64_t)32, p1 - base);
      const int32_t jreg = jnext;
      const T rreg = lane < cnt ? (T)((double)vnext - mean_offset) : T(0);
      if (base + 32 + lane < p1) {
        jnext = idx[base + 32 + lane];
        vnext = val[base + 32 + lane];
      }
      if (TC) {
        const int cnt8 = (cnt + 7) & ~7;             // whole k-steps, zero-filled past cnt
        // a lane stages the q-th 16-byte slice of rows e = it*4 + lane/8: all eight gathers are
        // issued before the first store (one L2 round trip per tile, not eight)
        float4 v[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int e = it * 4 + (lane >> 3);
          const int32_t j = __shfl_sync(0xffffffffu, jreg, e);
          v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e < cnt) v[it] = __ldg(reinterpret_cast<const float4*>(other + (int64_t)j * 32) + (lane & 7));
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int e = it * 4 + (lane >> 3);
          if (e < cnt8)
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(tile) + e * 40 + 4 * (lane & 7)) = v[it];
        }
        __syncwarp();
        const float* tf = reinterpret_cast<const float*>(tile);
        for (int e0 = 0; e0 < cnt8; e0 += 8) {
          uint32_t xh[4], xl[4], yh[4], yl[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x = tf[(e0 + tig) * 40 + 8 * j + g];
            const float y = tf[(e0 + tig + 4) * 40 + 8 * j + g];
            xh[j] = to_tf32(x); xl[j] = to_tf32(x - __uint_as_float(xh[j]));
            yh[j] = to_tf32(y); yl[j] = to_tf32(y - __uint_as_float(yh[j]));
          }
          int t = 0;
#pragma unroll
          for (int M = 0; M < 2; ++M) {
            const uint32_t ah[4] = {xh[2 * M], xh[2 * M + 1], yh[2 * M], yh[2 * M + 1]};
            const uint32_t al[4] = {xl[2 * M], xl[2 * M + 1], yl[2 * M], yl[2 * M + 1]};
#pragma unroll
            for (int N = 2 * M; N < 4; ++N, ++t) {
              mma_tf32(c[t], al, xh[N], yh[N]);
              mma_tf32(c[t], ah, xl[N], yl[N]);
              mma_tf32(c[t], ah, xh[N], yh[N]);
            }
          }
        }
      } else {
        for (int t = lane; t < cnt * dp; t += 32) {
          const int e = t / dp, k = t - e * dp;
          tile[e * ldt + k] = k < d ? other[(int64_t)idx[base + e] * d + k] : T(0);
        }
        __syncwarp();
#pragma unroll
        for (int a = 0; a < GIBBS_WARP_MAXBLK; ++a) {
          if (bks[a] >= 0) {
            const T* pa = tile + 2 * bks[a];
            const T* pb = tile + 2 * bls[a];
            T s00 = acc[a][0], s01 = acc[a][1], s10 = acc[a][2], s11 = acc[a][3];
            for (int e = 0; e < cnt; ++e) {
              const T a0 = pa[e * ldt], a1 = pa[e * ldt + 1];
              const T b0 = pb[e * ldt], b1 = pb[e * ldt + 1];
              s00 = fma(a0, b0, s00); s01 = fma(a0, b1, s01);
              s10 = fma(a1, b0, s10); s11 = fma(a1, b1, s11);
            }
            acc[a][0] = s00; acc[a][1] = s01; acc[a][2] = s10; acc[a][3] = s11;
          }
        }
      }
      {
        T sres = racc;
        for (int e = 0; e < cnt; ++e) {
          const T re = __shfl_sync(0xffffffffu, rreg, e);
          if (lane < d) sres = fma(tile[e * ldt + lane], re, sres);
        }
        racc = sres;
      }
      __syncwarp();
    }
    // ---- Lambda = alpha + beta F'F ; rhs = beta F'r + alpha mu ---------------------------
    if (TC) {
#pragma unroll
      for (int M = 0, t = 0; M < 2; ++M)
#pragma unroll
        for (int N = 2 * M; N < 4; ++N, ++t)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = 16 * M + g + 8 * (q >> 1), l = 8 * N + 2 * tig + (q & 1);
            if (k <= l) {
              A[k * lda + l] = (double)alpha[k * d + l] + beta * (double)c[t][q];
              if (k != l) A[l * lda + k] = (double)alpha[l * d + k] + beta * (double)c[t][q];
            }
          }
    } else {
#pragma unroll
      for (int a = 0; a < GIBBS_WARP_MAXBLK; ++a) {
        if (bks[a] >= 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = 2 * bks[a] + (q >> 1), l = 2 * bls[a] + (q & 1);
            if (k < d && l < d) {
              A[k * lda + l] = (double)alpha[k * d + l] + beta * (double)acc[a][q];
              if (k != l && bks[a] != bls[a])
                A[l * lda + k] = (double)alpha[l * d + k] + beta * (double)acc[a][q];
            }
          }
        }
      }
    }
    double rhs_lane = 0;
    if (lane < d) {
      rhs_lane = beta * (double)racc;
      for (int l = 0; l < d; ++l) rhs_lane += (double)alpha[lane * d + l] * (double)mu[l];
    }
    __syncwarp();
    if (gb.ex_row && gb.ex_row[chain] == row) {
      // this chain's extra rating: Lambda += beta f f', rhs += beta f (v - offset)
      const T* f = other + (int64_t)gb.ex_col[chain] * d;
      if (lane < d) {
        const double fl = (double)f[lane];
        for (int l = 0; l < d; ++l) A[lane * lda + l] += beta * fl * (double)f[l];
        rhs_lane += beta * fl * (gb.ex_val[chain] - mean_offset);
      }
      __syncwarp();
    }
    bool ok;
    if (FAST) {
      const double zl = lane < d ? philox_normal(seed, stream_id, (uint32_t)row,
                                                 (uint32_t)lane + 32u * (uint32_t)chain) : 0.0;
      ok = warp_solve32_fast<T>(A, d, lda, A, col, rhs_lane, zl, out + (int64_t)row * d);
    } else {
      ok = warp_solve32<T>(A, d, lda, A, col, vec, rhs_lane, z + (int64_t)row * d,
                           out + (int64_t)row * d);
    }
    if (!ok && lane == 0) atomicExch(fail, 1);
    __syncwarp();
  }
}

// mean / population variance / exceedance frequency of U_s[i].V_s[j] + offset over S samples
template <typename T, bool MAX>
__global__ void __launch_bounds__(128)
sample_stats_kernel(const int32_t* __restrict__ ci, const int32_t* __restrict__ cj, int64_t ncand,
                    int S, int64_t n, int64_t m, int d, const T* __restrict__ Us,
                    const T* __restrict__ Vs, T offset, T cutoff, T* __restrict__ mean_out,
                    T* __restrict__ var_out, T* __restrict__ prob_out, int select,
                    int64_t index_base, Best* __restrict__ part) {
  Best best{0.0, -1};
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncand;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = ci[c], j = cj[c];
    double sum = 0, wmean = 0, m2 = 0;
    int cnt_ge = 0;
    for (int s = 0; s < S; ++s) {
      const T* u = Us + ((int64_t)s * n + i) * d;
      const T* v = Vs + ((int64_t)s * m + j) * d;
      T dot = 0;
      for (int k = 0; k < d; ++k) dot = fma(u[k], v[k], dot);
      const T pred = dot + offset;
      cnt_ge += (pred >= cutoff);
      sum += (double)pred;
      const double delta = (double)pred - wmean;
      wmean += delta / (double)(s + 1);
      m2 += delta * ((double)pred - wmean);
    }
    const double mean = sum / (double)S, var = m2 / (double)S, prob = (double)cnt_ge / (double)S;
    if (mean_out) mean_out[c] = (T)mean;
    if (var_out) var_out[c] = (T)var;
    if (prob_out) prob_out[c] = (T)prob;
    const double sel = select == 0 ? mean : (select == 1 ? var : prob);
    if (better<MAX>(sel, c + index_base, best.v, best.i)) { best.v = sel; best.i = c + index_base; }
  }
  best = block_best<MAX>(best);
  if (threadIdx.x == 0) part[blockIdx.x] = best;
}

// The same statistics for ALL n x m cells (cell c = i*m + j) -- what the reference's criteria
// compute when `which` is the whole matrix or most of it (predicted_matrix(u, v) per sample,
// bayes_pmf.py:433-455).  A blocked product instead of a gather: a CTA owns a DENSE_TU x DENSE_TV
// tile of cells, stages the two factor tiles of one sample in shared memory (k-major, so the
// item reads of a warp are consecutive words and the user reads are broadcasts) and every
// thread carries a 2 x 4 block of cells with its running statistics in registers; the factor
// rows are read once per tile and sample instead of once per cell and sample.  The dot product
// is the same fma chain as sample_stats_kernel, the moments are accumulated about the first
// sample's prediction (no division in the loop).
constexpr int DENSE_TU = 32, DENSE_TV = 64, DENSE_THREADS = 256;

template <typename T, bool MAX>
__global__ void __launch_bounds__(DENSE_THREADS)
sample_stats_dense_kernel(int S, int n, int m, int d, const T* __restrict__ Us,
                          const T* __restrict__ Vs, T offset, T cutoff, T* __restrict__ mean_out,
                          T* __restrict__ var_out, T* __restrict__ prob_out, int select,
                          int64_t index_base, Best* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char dense_smem[];
  // k-major with one word of padding per row: the transposing stores of the staging loops
  // (consecutive threads = consecutive k of one factor row) fall on different banks
  constexpr int LU = DENSE_TU + 1, LV = DENSE_TV + 1;
  T* sU = reinterpret_cast<T*>(dense_smem);            // [d][LU]
  T* sV = sU + (size_t)d * LU;                         // [d][LV]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int tiles_v = (m + DENSE_TV - 1) / DENSE_TV;
  const int64_t n_tiles = (int64_t)((n + DENSE_TU - 1) / DENSE_TU) * tiles_v;
  Best best{0.0, -1};
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int i0 = (int)(tile / tiles_v) * DENSE_TU, j0 = (int)(tile % tiles_v) * DENSE_TV;
    double shift[2][4], s1[2][4], s2[2][4];
    int cnt[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int q = 0; q < 4; ++q) { shift[a][q] = 0; s1[a][q] = 0; s2[a][q] = 0; cnt[a][q] = 0; }
    for (int s = 0; s < S; ++s) {
      __syncthreads();                                 // the previous sample's tiles are done with
      const T* us = Us + ((int64_t)s * n + i0) * d;
      const T* vs = Vs + ((int64_t)s * m + j0) * d;
      const int nu = min(DENSE_TU, n - i0), nv = min(DENSE_TV, m - j0);
      for (int t = threadIdx.x; t < DENSE_TU * d; t += DENSE_THREADS) {
        const int r = t / d, k = t - r * d;
        sU[k * LU + r] = r < nu ? us[t] : T(0);
      }
      for (int t = threadIdx.x; t < DENSE_TV * d; t += DENSE_THREADS) {
        const int r = t / d, k = t - r * d;
        sV[k * LV + r] = r < nv ? vs[t] : T(0);
      }
      __syncthreads();
      T dot[2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) dot[a][q] = 0;
      for (int k = 0; k < d; ++k) {
        const T u0 = sU[k * LU + ty * 2], u1 = sU[k * LU + ty * 2 + 1];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const T v = sV[k * LV + tx + 16 * q];
          dot[0][q] = fma(u0, v, dot[0][q]);
          dot[1][q] = fma(u1, v, dot[1][q]);
        }
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const T pred = dot[a][q] + offset;
          cnt[a][q] += (pred >= cutoff);
          if (s == 0) shift[a][q] = (double)pred;
          const double x = (double)pred - shift[a][q];
          s1[a][q] += x;
          s2[a][q] = fma(x, x, s2[a][q]);
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + ty * 2 + a, j = j0 + tx + 16 * q;
        if (i < n && j < m) {
          const double mean = shift[a][q] + s1[a][q] / (double)S;
          double var = (s2[a][q] - s1[a][q] * s1[a][q] / (double)S) / (double)S;
          var = var < 0 ? 0 : var;
          const double prob = (double)cnt[a][q] / (double)S;
          const int64_t c = (int64_t)i * m + j;
          if (mean_out) mean_out[c] = (T)mean;
          if (var_out) var_out[c] = (T)var;
          if (prob_out) prob_out[c] = (T)prob;
          const double sel = select == 0 ? mean : (select == 1 ? var : prob);
          if (better<MAX>(sel, c + index_base, best.v, best.i)) { best.v = sel; best.i = c + index_base; }
        }
      }
  }
  best = block_best<MAX>(best);
  if (threadIdx.x == 0) part[blockIdx.x] = best;
}

int acquire_partials(Best** out, cudaStream_t s);
int launch_best_final(const Best* part_d, int nparts, bool maximize, amf_best_t* out_d,
                      cudaStream_t s);
// dense_tc.cu: the dense form on the tensor cores (tcgen05, TMEM accumulators, 2-D TMA)
bool dense_tc_applicable(int dtype, int S, int d, const void* prob_d, int select);
int dense_tc_launch(int S, int32_t n, int32_t m, int d, const float* Us, const float* Vs,
                    float offset, float* mean_d, float* var_d, int select, int maximize,
                    int64_t index_base, amf_best_t* best_d, cudaStream_t s);

template <typename T>
static int gibbs_launch(const amf_ratings* h, int side, int d, const T* other, const T* alpha,
                        const T* mu, double beta, double mean_offset, const T* z, T* out,
                        int row_begin, int row_end, cudaStream_t s, bool fast = false,
                        unsigned long long seed = 0, unsigned long long stream_id = 0,
                        GibbsBatch gb = GibbsBatch{1, 0, nullptr, nullptr, nullptr, nullptr}) {
  const int all_rows = side == 0 ? h->n_users : h->n_items;
  if (row_end < 0 || row_end > all_rows) row_end = all_rows;
  if (row_begin < 0) row_begin = 0;
  if (row_begin >= row_end) return AMF_OK;
  const int rows = row_end;
  // sticky: set by any row of any half-sweep on this handle, cleared only when
  // amf_gibbs_status reads it (a chain step is several half-sweeps and one status call)
  int* fail = reinterpret_cast<int*>(h->sums_d + 6);
  const int span = row_end - row_begin;
  const int grid = span < num_sms() * 8 ? span : num_sms() * 8;
  if (d <= 32) {
    // one warp per row; per-warp shared memory: A, Bm (or the staged tile), rhs, mean, flag
    const bool tc = sizeof(T) == 4 && d == 32;
    const size_t tile_bytes = tc ? 32 * 40 * sizeof(float) : 32 * (size_t)(((d + 1) & ~1) + 2) * sizeof(T);
    const int a_doubles = (int)std::max({(size_t)d * (d + 1), (tile_bytes + 7) / 8, (size_t)32 * 34});
    const int warp_doubles = (a_doubles + 64 + 32 + 1) & ~1;                    // 16-byte multiple
    const size_t smem_w = sizeof(double) * (size_t)warp_doubles * (GIBBS_THREADS / 32);
    const int64_t nwarp_rows = ((int64_t)span * (gb.count > 1 ? gb.count : 1) + GIBBS_THREADS / 32 - 1) /
                               (GIBBS_THREADS / 32);
    const int grid_w = (int)(nwarp_rows < (int64_t)num_sms() * 8 ? nwarp_rows : (int64_t)num_sms() * 8);
#define GIBBS_WARP_B(TC_, FAST_, MINB_)                                                          \
  do {                                                                                           \
    AMF_CUDA(cudaFuncSetAttribute(gibbs_rows_warp_kernel<T, TC_, FAST_, MINB_>,                  \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));    \
    gibbs_rows_warp_kernel<T, TC_, FAST_, MINB_><<<grid_w, GIBBS_THREADS, smem_w, s>>>(          \
        h->ptr[side], h->idx[side], (const T*)h->val[side], row_begin, rows, d, other, alpha, mu, \
        beta, mean_offset, z, out, fail, a_doubles, warp_doubles, seed, stream_id, gb);          \
  } while (0)
    // many waves of rows: the 16-warp build; otherwise the spill-free one
    const bool many = nwarp_rows > (int64_t)num_sms() * 3 * 8;
#define GIBBS_WARP(TC_, FAST_)                                                                   \
  do { if (many) GIBBS_WARP_B(TC_, FAST_, 4); else GIBBS_WARP_B(TC_, FAST_, 3); } while (0)
    if constexpr (sizeof(T) == 4) {
      if (tc) { if (fast) GIBBS_WARP(true, true); else GIBBS_WARP(true, false); }
      else { if (fast) GIBBS_WARP(false, true); else GIBBS_WARP(false, false); }
    } else {
      if (fast) GIBBS_WARP(false, true); else GIBBS_WARP(false, false);
    }
#undef GIBBS_WARP_B
#undef GIBBS_WARP
    AMF_LAUNCH_CHECK();
    return AMF_OK;
  }
  T* z_tmp = nullptr;
  if (fast) {   // d > 32: the CTA-per-row kernels take z from memory; fill it with the same counters
    AMF_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&z_tmp), sizeof(T) * (size_t)rows * d, s));
    philox_normal_kernel<T><<<num_sms() * 8, 256, 0, s>>>(seed, stream_id, rows, d, z_tmp);
    z = z_tmp;
  }
  struct FreeZ { T* p; cudaStream_t s; ~FreeZ() { if (p) cudaFreeAsync(p, s); } } free_z{z_tmp, s};
  if constexpr (sizeof(T) == 4) {
    if (d == 64) {                       // dense enough for whole MMA tiles: tensor-core Gram
      const int mt = d / 16, ntiles = mt * (mt + 1);
      const size_t stage = (size_t)GIBBS_TC_TILE * (d + 8), parts = (size_t)(GIBBS_THREADS / 32) * ntiles * 128;
      const size_t smem_tc = sizeof(double) * (2 * d * (d + 1) + 2 * d) +
                             sizeof(float) * (GIBBS_TC_TILE + (stage > parts ? stage : parts));
      AMF_CUDA(cudaFuncSetAttribute(gibbs_rows_tc_kernel<64>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));
      gibbs_rows_tc_kernel<64><<<grid, GIBBS_THREADS, smem_tc, s>>>(
          h->ptr[side], h->idx[side], (const float*)h->val[side], row_begin, rows, other, alpha,
          mu, beta, mean_offset, z, out, fail);
      AMF_LAUNCH_CHECK();
      return AMF_OK;
    }
  }
  const size_t smem = sizeof(double) * (2 * d * (d + 1) + 2 * d + GIBBS_TILE) +
                      sizeof(T) * GIBBS_TILE * (((d + 1) & ~1) + 2);
  AMF_CUDA(cudaFuncSetAttribute(gibbs_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  gibbs_rows_kernel<T><<<grid, GIBBS_THREADS, smem, s>>>(h->ptr[side], h->idx[side],
                                                         (const T*)h->val[side], row_begin, rows,
                                                         d, other, alpha, mu, beta, mean_offset,
                                                         z, out, fail);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int amf_gibbs_half_sweep_rows(const amf_ratings_t* h, int side, int dtype, int d,
                              const void* other_d, const void* alpha_d, const void* mu_d,
                              double beta, double mean_offset, const void* z_d, void* out_d,
                              int32_t row_begin, int32_t row_end, void* stream);

int amf_gibbs_half_sweep(const amf_ratings_t* h, int side, int dtype, int d, const void* other_d,
                         const void* alpha_d, const void* mu_d, double beta, double mean_offset,
                         const void* z_d, void* out_d, void* stream) {
  return amf_gibbs_half_sweep_rows(h, side, dtype, d, other_d, alpha_d, mu_d, beta, mean_offset,
                                   z_d, out_d, 0, -1, stream);
}

int amf_gibbs_half_sweep_rows(const amf_ratings_t* h, int side, int dtype, int d,
                              const void* other_d, const void* alpha_d, const void* mu_d,
                              double beta, double mean_offset, const void* z_d, void* out_d,
                              int32_t row_begin, int32_t row_end, void* stream) {
  AMF_REQUIRE(h && other_d && alpha_d && mu_d && z_d && out_d, "amf_gibbs_half_sweep: NULL argument");
  AMF_REQUIRE(side == 0 || side == 1, "amf_gibbs_half_sweep: side must be 0 or 1");
  AMF_REQUIRE(dtype == h->dtype, "amf_gibbs_half_sweep: dtype does not match the rating list");
  AMF_REQUIRE(d >= 1 && d <= GIBBS_MAXD, "amf_gibbs_half_sweep: latent_d=%d unsupported (max %d)", d,
              GIBBS_MAXD);
  cudaStream_t s = (cudaStream_t)stream;
  {
    int rc = ratings_compact(const_cast<amf_ratings*>(h), s);
    if (rc != AMF_OK) return rc;
  }
  if (dtype == AMF_F32)
    return gibbs_launch<float>(h, side, d, (const float*)other_d, (const float*)alpha_d,
                               (const float*)mu_d, beta, mean_offset, (const float*)z_d,
                               (float*)out_d, row_begin, row_end, s);
  return gibbs_launch<double>(h, side, d, (const double*)other_d, (const double*)alpha_d,
                              (const double*)mu_d, beta, mean_offset, (const double*)z_d,
                              (double*)out_d, row_begin, row_end, s);
}

int amf_gibbs_half_sweep_device_rng(const amf_ratings_t* h, int side, int dtype, int d,
                                    const void* other_d, const void* alpha_d, const void* mu_d,
                                    double beta, double mean_offset, uint64_t seed,
                                    uint64_t stream_id, void* out_d, int32_t row_begin,
                                    int32_t row_end, void* stream) {
  AMF_REQUIRE(h && other_d && alpha_d && mu_d && out_d, "amf_gibbs_half_sweep_device_rng: NULL argument");
  AMF_REQUIRE(side == 0 || side == 1, "amf_gibbs_half_sweep_device_rng: side must be 0 or 1");
  AMF_REQUIRE(dtype == h->dtype, "amf_gibbs_half_sweep_device_rng: dtype does not match the rating list");
  AMF_REQUIRE(d >= 1 && d <= GIBBS_MAXD, "amf_gibbs_half_sweep_device_rng: latent_d=%d unsupported (max %d)",
              d, GIBBS_MAXD);
  cudaStream_t s = (cudaStream_t)stream;
  {
    int rc = ratings_compact(const_cast<amf_ratings*>(h), s);
    if (rc != AMF_OK) return rc;
  }
  if (dtype == AMF_F32)
    return gibbs_launch<float>(h, side, d, (const float*)other_d, (const float*)alpha_d,
                               (const float*)mu_d, beta, mean_offset, nullptr, (float*)out_d,
                               row_begin, row_end, s, true, seed, stream_id);
  return gibbs_launch<double>(h, side, d, (const double*)other_d, (const double*)alpha_d,
                              (const double*)mu_d, beta, mean_offset, nullptr, (double*)out_d,
                              row_begin, row_end, s, true, seed, stream_id);
}

int amf_gibbs_half_sweep_batched(const amf_ratings_t* h, int side, int dtype, int d, int count,
                                 const void* other_d, const void* alpha_d, const void* mu_d,
                                 double beta, double mean_offset, const double* offsets_d,
                                 const int32_t* ex_row_d, const int32_t* ex_col_d,
                                 const double* ex_val_d, uint64_t seed, uint64_t stream_id,
                                 void* out_d, void* stream) {
  AMF_REQUIRE(h && other_d && alpha_d && mu_d && out_d, "amf_gibbs_half_sweep_batched: NULL argument");
  AMF_REQUIRE(side == 0 || side == 1, "amf_gibbs_half_sweep_batched: side must be 0 or 1");
  AMF_REQUIRE(dtype == h->dtype, "amf_gibbs_half_sweep_batched: dtype does not match the rating list");
  AMF_REQUIRE(d >= 1 && d <= 32, "amf_gibbs_half_sweep_batched: latent_d=%d outside 1..32", d);
  AMF_REQUIRE(count >= 1, "amf_gibbs_half_sweep_batched: count must be positive");
  AMF_REQUIRE(!ex_row_d == !ex_col_d && !ex_row_d == !ex_val_d,
              "amf_gibbs_half_sweep_batched: the extra-rating arrays come together");
  cudaStream_t s = (cudaStream_t)stream;
  {
    int rc = ratings_compact(const_cast<amf_ratings*>(h), s);
    if (rc != AMF_OK) return rc;
  }
  const GibbsBatch gb{count, side == 0 ? h->n_items : h->n_users, ex_row_d, ex_col_d, ex_val_d, offsets_d};
  if (dtype == AMF_F32)
    return gibbs_launch<float>(h, side, d, (const float*)other_d, (const float*)alpha_d,
                               (const float*)mu_d, beta, mean_offset, nullptr, (float*)out_d, 0, -1, s,
                               true, seed, stream_id, gb);
  return gibbs_launch<double>(h, side, d, (const double*)other_d, (const double*)alpha_d,
                              (const double*)mu_d, beta, mean_offset, nullptr, (double*)out_d, 0, -1, s,
                              true, seed, stream_id, gb);
}

int amf_gibbs_chain_device(const amf_ratings_t* h, int dtype, int d, int n_samples, int num_gibbs,
                           const void* users_d, const void* items_d, const double* prior_u_d,
                           const double* prior_v_d, double beta, double mean_offset, uint64_t seed,
                           uint64_t stream_id0, void* out_users_d, void* out_items_d, void* stream) {
  AMF_REQUIRE(h && users_d && items_d && prior_u_d && prior_v_d && out_users_d && out_items_d,
              "amf_gibbs_chain_device: NULL argument");
  AMF_REQUIRE(dtype == h->dtype, "amf_gibbs_chain_device: dtype does not match the rating list");
  AMF_REQUIRE(d >= 1 && d <= 32, "amf_gibbs_chain_device: d must be in [1, 32]");
  AMF_REQUIRE(n_samples >= 0 && num_gibbs >= 1, "amf_gibbs_chain_device: bad counts");
  AMF_REQUIRE(h->n_users >= 2 && h->n_items >= 2, "amf_gibbs_chain_device: needs at least two rows per side");
  if (n_samples == 0) return AMF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  {
    int rc = ratings_compact(const_cast<amf_ratings*>(h), s);
    if (rc != AMF_OK) return rc;
  }
  const size_t es = dtype == AMF_F32 ? 4 : 8;
  const size_t bu = (size_t)h->n_users * d * es, bv = (size_t)h->n_items * d * es;
  const size_t bh = (size_t)(d + d * d) * es;           // one side's (mu, alpha)
  // scratch: both sides' hyper-parameters, and the rows of the rounds that are not kept
  unsigned char* scratch = nullptr;
  AMF_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), 2 * bh + bu + bv, s));
  unsigned char* hyper = scratch;
  unsigned char* tmp_u = scratch + 2 * bh;
  unsigned char* tmp_v = tmp_u + bu;
  const unsigned char* cur_u = static_cast<const unsigned char*>(users_d);
  const unsigned char* cur_v = static_cast<const unsigned char*>(items_d);
  uint64_t sid = stream_id0;
  int rc = AMF_OK;
  auto half = [&](int side, const unsigned char* other, unsigned char* out) -> int {
    const unsigned char* mu = hyper + side * bh;
    const unsigned char* alpha = mu + (size_t)d * es;
    const uint64_t id = sid++;
    if (dtype == AMF_F32)
      return gibbs_launch<float>(h, side, d, (const float*)other, (const float*)alpha, (const float*)mu,
                                 beta, mean_offset, nullptr, (float*)out, 0, -1, s, true, seed, id);
    return gibbs_launch<double>(h, side, d, (const double*)other, (const double*)alpha,
                                (const double*)mu, beta, mean_offset, nullptr, (double*)out, 0, -1, s,
                                true, seed, id);
  };
  for (int smp = 0; smp < n_samples && rc == AMF_OK; ++smp) {
    // the same counters the host-driven loop of samples_device() uses: draw for draw equal to it
    rc = gibbs_hyper_draw(h, dtype, d, h->n_users, cur_u, prior_u_d, seed, (1ull << 40) + sid, hyper,
                          hyper + (size_t)d * es, s);
    ++sid;
    if (rc == AMF_OK)
      rc = gibbs_hyper_draw(h, dtype, d, h->n_items, cur_v, prior_v_d, seed, (1ull << 40) + sid,
                            hyper + bh, hyper + bh + (size_t)d * es, s);
    ++sid;
    unsigned char* keep_u = static_cast<unsigned char*>(out_users_d) + (size_t)smp * bu;
    unsigned char* keep_v = static_cast<unsigned char*>(out_items_d) + (size_t)smp * bv;
    for (int r = 0; r < num_gibbs && rc == AMF_OK; ++r) {
      const bool last = r == num_gibbs - 1;
      unsigned char* nu = last ? keep_u : tmp_u;
      unsigned char* nv = last ? keep_v : tmp_v;
      rc = half(0, cur_v, nu);
      cur_u = nu;
      if (rc == AMF_OK) rc = half(1, cur_u, nv);
      cur_v = nv;
    }
  }
  cudaFreeAsync(scratch, s);
  return rc;
}

int amf_philox_normal(int dtype, uint64_t seed, uint64_t stream_id, int64_t rows, int d,
                      void* out_d, void* stream) {
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_philox_normal: bad dtype");
  AMF_REQUIRE(rows >= 0 && d >= 1 && out_d, "amf_philox_normal: bad arguments");
  if (rows == 0) return AMF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == AMF_F32)
    philox_normal_kernel<float><<<num_sms() * 8, 256, 0, s>>>(seed, stream_id, rows, d, (float*)out_d);
  else
    philox_normal_kernel<double><<<num_sms() * 8, 256, 0, s>>>(seed, stream_id, rows, d, (double*)out_d);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

int amf_gibbs_status(const amf_ratings_t* h, int* failed, void* stream) {
  AMF_REQUIRE(h && failed, "amf_gibbs_status: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  int* flag = reinterpret_cast<int*>(h->sums_d + 6);
  AMF_CUDA(cudaMemcpyAsync(failed, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  AMF_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
  AMF_CUDA(cudaStreamSynchronize(s));
  return AMF_OK;
}

int amf_bayes_sample_stats(int dtype, int64_t ncand, const int32_t* ci_d, const int32_t* cj_d,
                           int S, int32_t n, int32_t m, int d, const void* Us_d, const void* Vs_d,
                           double mean_offset, double cutoff, void* mean_d, void* var_d,
                           void* prob_d, int select, int maximize, int64_t index_base,
                           amf_best_t* best_d, void* stream) {
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_bayes_sample_stats: bad dtype");
  AMF_REQUIRE(S >= 1 && d >= 1 && ncand >= 0, "amf_bayes_sample_stats: bad sizes");
  AMF_REQUIRE(select >= 0 && select <= 2, "amf_bayes_sample_stats: bad select");
  cudaStream_t s = (cudaStream_t)stream;
  const bool dense = !ci_d && !cj_d && ncand > 0;     // every cell of the matrix, c = i*m + j
  const size_t dense_smem = (size_t)(DENSE_TU + DENSE_TV + 2) * d * (dtype == AMF_F32 ? 4 : 8);
  if (dense) {
    AMF_REQUIRE(ncand == (int64_t)n * m, "amf_bayes_sample_stats: the dense form (NULL candidate "
                "arrays) needs ncand = n*m");
    AMF_REQUIRE(dense_smem <= 200 * 1024, "amf_bayes_sample_stats: latent_d=%d is too large for "
                "the dense form", d);
  } else {
    AMF_REQUIRE(ncand == 0 || (ci_d && cj_d), "amf_bayes_sample_stats: NULL candidate array");
  }
  if (dense && dense_tc_applicable(dtype, S, d, prob_d, select))
    return dense_tc_launch(S, n, m, d, (const float*)Us_d, (const float*)Vs_d, (float)mean_offset,
                           (float*)mean_d, (float*)var_d, select, maximize, index_base, best_d, s);
  Best* part = nullptr;
  int rc = acquire_partials(&part, s);
  if (rc != AMF_OK) return rc;
  if (dense) {
    const size_t smem = dense_smem;
    const int64_t tiles = (int64_t)((n + DENSE_TU - 1) / DENSE_TU) * ((m + DENSE_TV - 1) / DENSE_TV);
    const int dgrid = (int)(tiles < (int64_t)num_sms() * 8 ? tiles : (int64_t)num_sms() * 8);
#define DENSE(T, MAXV)                                                                          \
  do {                                                                                          \
    AMF_CUDA(cudaFuncSetAttribute(sample_stats_dense_kernel<T, MAXV>,                           \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    sample_stats_dense_kernel<T, MAXV><<<dgrid, DENSE_THREADS, smem, s>>>(                      \
        S, n, m, d, (const T*)Us_d, (const T*)Vs_d, (T)mean_offset, (T)cutoff, (T*)mean_d,      \
        (T*)var_d, (T*)prob_d, select, index_base, part);                                       \
  } while (0)
    if (dtype == AMF_F32) { if (maximize) DENSE(float, true); else DENSE(float, false); }
    else { if (maximize) DENSE(double, true); else DENSE(double, false); }
#undef DENSE
    AMF_LAUNCH_CHECK();
    if (best_d) return launch_best_final(part, dgrid, maximize != 0, best_d, s);
    AMF_CUDA(cudaFreeAsync(part, s));
    return AMF_OK;
  }
  const int64_t blocks = (ncand + 127) / 128;
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? (blocks > 0 ? blocks : 1)
                                                          : (int64_t)num_sms() * 16);
#define STATS(T, MAXV)                                                                          \
  sample_stats_kernel<T, MAXV><<<grid, 128, 0, s>>>(ci_d, cj_d, ncand, S, n, m, d,              \
      (const T*)Us_d, (const T*)Vs_d, (T)mean_offset, (T)cutoff, (T*)mean_d, (T*)var_d,         \
      (T*)prob_d, select, index_base, part)
  if (dtype == AMF_F32) { if (maximize) STATS(float, true); else STATS(float, false); }
  else { if (maximize) STATS(double, true); else STATS(double, false); }
#undef STATS
  AMF_LAUNCH_CHECK();
  if (best_d) return launch_best_final(part, grid, maximize != 0, best_d, s);
  AMF_CUDA(cudaFreeAsync(part, s));
  return AMF_OK;
}

#pragma GCC visibility pop
}  // extern "C"
