// Candidate scoring with a fused arg-best (active_pmf.py:739-770 _get_key_vals and the
// chooser at :737; criteria pred :416-421, approx_pred_mean_var :392-400, pred_variance
// :502-524 = exp_dotprod_sq (normal_exps_cy.pyx:111-135) - E^2, _prob_ge_cutoff :432-439).
//
// The reference maps a Python method over the pool (optionally through multiprocessing.Pool,
// pickling the model per chunk).  Here the pool is a pair of int32 device arrays; one launch
// scores every candidate, optionally stores the scores, and reduces (value, index) to the
// winner with the lowest-index tie-break, so selection costs no second pass.
#include "common.cuh"

namespace amf {

template <int E>
__device__ __forceinline__ void load_idx_vec(const int32_t* __restrict__ p, int32_t (&out)[E]) {
  if constexpr (E % 4 == 0) {
#pragma unroll
    for (int q = 0; q < E / 4; ++q) {
      const int4 v = __ldcs(reinterpret_cast<const int4*>(p) + q);
      out[4 * q] = v.x; out[4 * q + 1] = v.y; out[4 * q + 2] = v.z; out[4 * q + 3] = v.w;
    }
  } else if constexpr (E == 2) {
    const int2 v = __ldcs(reinterpret_cast<const int2*>(p));
    out[0] = v.x; out[1] = v.y;
  } else {
    out[0] = __ldcs(p);
  }
}

template <bool MAX>
__global__ void best_final_kernel(const Best* __restrict__ part, int nparts,
                                  amf_best_t* __restrict__ out) {
  Best b{0.0, -1};
  for (int t = threadIdx.x; t < nparts; t += blockDim.x) {
    Best o = part[t];
    if (better<MAX>(o.v, o.i, b.v, b.i)) b = o;
  }
  b = block_best<MAX>(b);
  if (threadIdx.x == 0) { out->value = b.v; out->index = b.i; }
}

int launch_best_final(const Best* part_d, int nparts, bool maximize, amf_best_t* out_d,
                      cudaStream_t s) {
  if (maximize) best_final_kernel<true><<<1, 256, 0, s>>>(part_d, nparts, out_d);
  else best_final_kernel<false><<<1, 256, 0, s>>>(part_d, nparts, out_d);
  AMF_LAUNCH_CHECK();
  AMF_CUDA(cudaFreeAsync(const_cast<Best*>(part_d), s));   // pairs with acquire_partials
  return AMF_OK;
}

// MAP prediction U_i . V_j.  A warp takes 32 consecutive candidates per iteration; each group
// of LPR lanes owns LPR of them and every lane holds one 16-byte slice (x VPL) of the factor
// rows.  All LPR row gathers of a group are issued before any is used (memory-level
// parallelism), the user row is re-used while consecutive candidates share i (the pool is
// sorted by user), and a transpose-reduce leaves exactly one finished dot product per lane
// (candidate base+lane), so stores are coalesced and the arg-best costs one compare per lane.
template <typename T, int LPR, int VPL, bool MAX, bool VECIDX>
__global__ void __launch_bounds__(256)
score_pred_kernel(const int32_t* __restrict__ ci, const int32_t* __restrict__ cj, int64_t ncand,
                  const T* __restrict__ U, const T* __restrict__ Vm, int ld, int nvec,
                  T* __restrict__ scores, int64_t index_base, Best* __restrict__ part) {
  using V = typename Vec<T>::type;
  const int lane = threadIdx.x & 31;
  const int g = lane / LPR, l = lane % LPR;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  T best_v = 0;
  int64_t best_c = -1;
  for (int64_t base = warp * 32; base < ncand; base += nwarps * 32) {
    const int64_t c0 = base + g * LPR;        // first candidate of this lane group
    int32_t is[LPR], js[LPR];
    if (VECIDX && LPR >= 2 && base + 32 <= ncand) {
      load_idx_vec<LPR>(ci + c0, is);
      load_idx_vec<LPR>(cj + c0, js);
    } else {
#pragma unroll
      for (int s = 0; s < LPR; ++s) {
        const bool ok = c0 + s < ncand;
        is[s] = ok ? __ldcs(ci + c0 + s) : 0;
        js[s] = ok ? __ldcs(cj + c0 + s) : 0;
      }
    }
    T p[LPR];
#pragma unroll
    for (int s = 0; s < LPR; ++s) p[s] = 0;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int chunk = l + v * LPR;
      const bool have = chunk < nvec;
      V b[LPR];
#pragma unroll
      for (int s = 0; s < LPR; ++s)
        b[s] = have ? reinterpret_cast<const V*>(Vm + (int64_t)js[s] * ld)[chunk] : vzero(V());
      V a = vzero(V());
#pragma unroll
      for (int s = 0; s < LPR; ++s) {
        if (have && (s == 0 || is[s] != is[s - 1]))
          a = reinterpret_cast<const V*>(U + (int64_t)is[s] * ld)[chunk];
        p[s] += vdot(a, b[s]);
      }
    }
    // transpose-reduce across the LPR lanes of the group: lane l ends with candidate l's sum
#pragma unroll
    for (int half = LPR >> 1; half >= 1; half >>= 1) {
      const bool upper = (l & half) != 0;
#pragma unroll
      for (int t = 0; t < half; ++t) {
        const T send = upper ? p[t] : p[t + half];
        const T keep = upper ? p[t + half] : p[t];
        p[t] = keep + __shfl_xor_sync(0xffffffffu, send, half);
      }
    }
    const int64_t c = base + lane;
    if (c < ncand) {
      if (scores) __stcs(scores + c, p[0]);
      if (MAX ? (p[0] > best_v || best_c < 0) && (p[0] == p[0])
              : (p[0] < best_v || best_c < 0) && (p[0] == p[0])) {
        best_v = p[0]; best_c = c;
      }
    }
  }
  Best best{(double)best_v, best_c < 0 ? -1 : best_c + index_base};
  best = block_best<MAX>(best);
  if (threadIdx.x == 0) part[blockIdx.x] = best;
}

// Rows longer than 64 vectors (the packed pred_variance tables of blocks.cu at d >= 16: d(d+1)
// numbers per row, 4.2 KB at d = 32): a warp per group of four candidates, every lane strides
// over the row in 16-byte vectors with the four item-row loads of a step in flight together;
// the user row is re-read only when the user changes (sorted pools: it stays in L1).
template <typename T, bool MAX>
__global__ void __launch_bounds__(256)
score_pred_long_kernel(const int32_t* __restrict__ ci, const int32_t* __restrict__ cj,
                       int64_t ncand, const T* __restrict__ U, const T* __restrict__ Vm, int ld,
                       int nvec, T* __restrict__ scores, int64_t index_base,
                       Best* __restrict__ part) {
  using V = typename Vec<T>::type;
  constexpr int B = 4;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  Best best{0.0, -1};
  for (int64_t base = warp * B; base < ncand; base += nwarps * B) {
    int32_t is[B], js[B];
#pragma unroll
    for (int s = 0; s < B; ++s) {
      const int64_t c = base + s < ncand ? base + s : ncand - 1;
      is[s] = ci[c]; js[s] = cj[c];
    }
    T p[B];
#pragma unroll
    for (int s = 0; s < B; ++s) p[s] = 0;
    for (int chunk = lane; chunk < nvec; chunk += 32) {
      V b[B];
#pragma unroll
      for (int s = 0; s < B; ++s) b[s] = reinterpret_cast<const V*>(Vm + (int64_t)js[s] * ld)[chunk];
      V a = vzero(V());
#pragma unroll
      for (int s = 0; s < B; ++s) {
        if (s == 0 || is[s] != is[s - 1]) a = reinterpret_cast<const V*>(U + (int64_t)is[s] * ld)[chunk];
        p[s] += vdot(a, b[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < B; ++s) p[s] = warp_sum(p[s]);
    if (lane < B && base + lane < ncand) {
      const T mine = lane == 0 ? p[0] : (lane == 1 ? p[1] : (lane == 2 ? p[2] : p[3]));
      const int64_t c = base + lane;
      if (scores) __stcs(scores + c, mine);
      if (better<MAX>((double)mine, c + index_base, best.v, best.i)) {
        best.v = (double)mine; best.i = c + index_base;
      }
    }
  }
  best = block_best<MAX>(best);
  if (threadIdx.x == 0) part[blockIdx.x] = best;
}

struct NormalView {
  const void *mean_u, *mean_v, *cov_uu, *cov_vv, *cov_uv;
  long long mean_u_stride, mean_v_stride, uu_stride, uu_ld, vv_stride, vv_ld, uv_stride_i,
      uv_stride_j, uv_ld;
};

// Moments of U_i . V_j under the Gaussian approximation, one thread per candidate.
//   E   = mu.mv + tr C
//   Var = <A,B> + <C,C'> + mv'A mv + mu'B mu + 2 mu'C'mv        (== exp_dotprod_sq - E^2)
template <typename T, int CRIT, bool MAX>
__global__ void __launch_bounds__(128)
score_normal_kernel(const int32_t* __restrict__ ci, const int32_t* __restrict__ cj, int64_t ncand,
                    int d, NormalView nv, T cutoff, T* __restrict__ scores, int64_t index_base,
                    Best* __restrict__ part) {
  Best best{0.0, -1};
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncand;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int32_t i = ci[c], j = cj[c];
    const T* mu = (const T*)nv.mean_u + i * nv.mean_u_stride;
    const T* mv = (const T*)nv.mean_v + j * nv.mean_v_stride;
    const T* A = (const T*)nv.cov_uu + i * nv.uu_stride;
    const T* B = (const T*)nv.cov_vv + j * nv.vv_stride;
    const T* C = nv.cov_uv ? (const T*)nv.cov_uv + i * nv.uv_stride_i + j * nv.uv_stride_j : nullptr;
    T e = 0, var = 0;
    for (int k = 0; k < d; ++k) {
      const T muk = mu[k], mvk = mv[k];
      e = fma(muk, mvk, e);
      if (C) e += C[k * nv.uv_ld + k];
      if (CRIT != AMF_CRIT_APPROX_MEAN) {
        T ab = 0, amv = 0, bmu = 0, cc = 0, cmv = 0;
        for (int l = 0; l < d; ++l) {
          const T a = A[k * nv.uu_ld + l], b = B[k * nv.vv_ld + l];
          ab = fma(a, b, ab);
          amv = fma(a, mv[l], amv);
          bmu = fma(b, mu[l], bmu);
          if (C) {
            const T clk = C[l * nv.uv_ld + k];
            cc = fma(C[k * nv.uv_ld + l], clk, cc);
            cmv = fma(clk, mv[l], cmv);
          }
        }
        var += ab + cc + mvk * amv + muk * bmu + 2 * muk * cmv;
      }
    }
    T out;
    if (CRIT == AMF_CRIT_APPROX_MEAN) out = e;
    else if (CRIT == AMF_CRIT_PRED_VARIANCE) out = var;
    else {
      // scipy.stats.norm.sf(cutoff, loc=e, scale=var): the reference passes the variance as the
      // scale (active_pmf.py:438-439); scale <= 0 gives nan there too
      out = var > 0 ? T(0.5) * erfc((cutoff - e) / (var * T(1.4142135623730951))) : T(NAN);
    }
    if (scores) scores[c] = out;
    if (better<MAX>((double)out, c + index_base, best.v, best.i)) {
      best.v = (double)out; best.i = c + index_base;
    }
  }
  best = block_best<MAX>(best);
  if (threadIdx.x == 0) part[blockIdx.x] = best;
}

static inline int pow2c(int x) { int p = 1; while (p < x) p <<= 1; return p; }

template <typename T, bool MAX>
static int score_pred(int64_t ncand, const int32_t* ci, const int32_t* cj, int d, int ld,
                      const T* U, const T* V, T* scores, int64_t index_base, Best* part, int grid,
                      cudaStream_t s) {
  constexpr int N = Vec<T>::N;
  AMF_REQUIRE(ld >= d && ld % N == 0, "ld=%d must be >= d=%d and a multiple of %d", ld, d, N);
  const int nvec = ld / N;
  int lpr = pow2c(nvec), vpl = 1;
  if (lpr > 8) { vpl = lpr / 8; lpr = 8; }    // at most 8 candidates (and partial sums) per lane
  if (vpl > 8) {
    score_pred_long_kernel<T, MAX><<<grid, 256, 0, s>>>(ci, cj, ncand, U, V, ld, nvec, scores,
                                                        index_base, part);
    AMF_LAUNCH_CHECK();
    return AMF_OK;
  }
  const bool vec = (reinterpret_cast<uintptr_t>(ci) % 16 == 0) && (reinterpret_cast<uintptr_t>(cj) % 16 == 0);
#define PRED(LPR_, VPL_)                                                                       \
  do {                                                                                         \
    if (vec) score_pred_kernel<T, LPR_, VPL_, MAX, true><<<grid, 256, 0, s>>>(                 \
        ci, cj, ncand, U, V, ld, nvec, scores, index_base, part);                              \
    else score_pred_kernel<T, LPR_, VPL_, MAX, false><<<grid, 256, 0, s>>>(                    \
        ci, cj, ncand, U, V, ld, nvec, scores, index_base, part);                              \
  } while (0)
  if (vpl == 1) {
    switch (lpr) {
      case 1: PRED(1, 1); break;
      case 2: PRED(2, 1); break;
      case 4: PRED(4, 1); break;
      default: PRED(8, 1); break;
    }
  } else if (vpl == 2) {
    PRED(8, 2);
  } else if (vpl <= 4) {
    PRED(8, 4);
  } else {
    PRED(8, 8);
  }
#undef PRED
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

template <typename T, bool MAX>
static int score_normal(int crit, int64_t ncand, const int32_t* ci, const int32_t* cj, int d,
                        const NormalView& nv, double cutoff, T* scores, int64_t index_base,
                        Best* part, int grid, cudaStream_t s) {
  switch (crit) {
    case AMF_CRIT_APPROX_MEAN:
      score_normal_kernel<T, AMF_CRIT_APPROX_MEAN, MAX><<<grid, 128, 0, s>>>(
          ci, cj, ncand, d, nv, (T)cutoff, scores, index_base, part);
      break;
    case AMF_CRIT_PRED_VARIANCE:
      score_normal_kernel<T, AMF_CRIT_PRED_VARIANCE, MAX><<<grid, 128, 0, s>>>(
          ci, cj, ncand, d, nv, (T)cutoff, scores, index_base, part);
      break;
    default:
      score_normal_kernel<T, AMF_CRIT_PROB_GE, MAX><<<grid, 128, 0, s>>>(
          ci, cj, ncand, d, nv, (T)cutoff, scores, index_base, part);
      break;
  }
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

// scratch for the per-block partial winners: a stream-ordered allocation per call, released by
// launch_best_final on the same stream, so concurrent callers (one host thread per criterion,
// active_pmf.py:1064-1079) never share a buffer
int acquire_partials(Best** out, cudaStream_t s) {
  // keep freed blocks in the device's default pool across synchronisations: with the default
  // release threshold (0) every sync trims the pool and the next allocation goes to the driver
  static bool tuned[64] = {false};
  int dev = 0;
  AMF_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !tuned[dev]) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = 1ull << 30;   // workspaces of the dense kernels are tens of MB per call
      uint64_t cur = 0;
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur);
      if (cur < keep) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    tuned[dev] = true;
  }
  AMF_CUDA(cudaMallocAsync(reinterpret_cast<void**>(out), sizeof(Best) * 8192, s));
  return AMF_OK;
}

}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int amf_best_reduce(const amf_best_t* recs_d, int n, int maximize, amf_best_t* out_d,
                    void* stream) {
  AMF_REQUIRE(recs_d && out_d && n >= 0, "amf_best_reduce: bad arguments");
  static_assert(sizeof(Best) == sizeof(amf_best_t), "record layouts must agree");
  cudaStream_t s = (cudaStream_t)stream;
  const Best* part = reinterpret_cast<const Best*>(recs_d);
  if (maximize) best_final_kernel<true><<<1, 256, 0, s>>>(part, n, out_d);
  else best_final_kernel<false><<<1, 256, 0, s>>>(part, n, out_d);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

int amf_score_candidates(int criterion, int dtype, int64_t ncand, const int32_t* ci_d,
                         const int32_t* cj_d, int d, int ld, const void* U_d, const void* V_d,
                         const amf_normal_view_t* nvp, double cutoff, void* scores_d,
                         int maximize, int64_t index_base, amf_best_t* best_d, void* stream) {
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_score_candidates: bad dtype %d", dtype);
  AMF_REQUIRE(criterion >= AMF_CRIT_PRED && criterion <= AMF_CRIT_PROB_GE,
              "amf_score_candidates: unknown criterion %d", criterion);
  AMF_REQUIRE(best_d != nullptr, "amf_score_candidates: best_d is NULL");
  AMF_REQUIRE(ncand >= 0 && d > 0, "amf_score_candidates: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  Best* part = nullptr;
  int rc = acquire_partials(&part, s);
  if (rc != AMF_OK) return rc;
  int grid;
  if (criterion == AMF_CRIT_PRED) {
    AMF_REQUIRE(U_d && V_d, "amf_score_candidates: U/V are NULL");
    const int64_t blocks = (ncand + 255) / 256;  // one 32-candidate batch per warp before capping
    grid = (int)(blocks < (int64_t)num_sms() * 8 ? (blocks > 0 ? blocks : 1) : (int64_t)num_sms() * 8);
  } else {
    AMF_REQUIRE(nvp && nvp->mean_u && nvp->mean_v && nvp->cov_uu && nvp->cov_vv,
                "amf_score_candidates: normal view is incomplete");
    const int64_t blocks = (ncand + 127) / 128;
    grid = (int)(blocks < (int64_t)num_sms() * 16 ? (blocks > 0 ? blocks : 1) : (int64_t)num_sms() * 16);
  }
  NormalView nv{};
  if (nvp) {
    nv.mean_u = nvp->mean_u; nv.mean_v = nvp->mean_v; nv.cov_uu = nvp->cov_uu;
    nv.cov_vv = nvp->cov_vv; nv.cov_uv = nvp->cov_uv;
    nv.mean_u_stride = nvp->mean_u_stride; nv.mean_v_stride = nvp->mean_v_stride;
    nv.uu_stride = nvp->uu_stride; nv.uu_ld = nvp->uu_ld;
    nv.vv_stride = nvp->vv_stride; nv.vv_ld = nvp->vv_ld;
    nv.uv_stride_i = nvp->uv_stride_i; nv.uv_stride_j = nvp->uv_stride_j; nv.uv_ld = nvp->uv_ld;
  }
#define DISPATCH(T)                                                                              \
  if (criterion == AMF_CRIT_PRED) {                                                              \
    rc = maximize ? score_pred<T, true>(ncand, ci_d, cj_d, d, ld, (const T*)U_d, (const T*)V_d,  \
                                        (T*)scores_d, index_base, part, grid, s)                 \
                  : score_pred<T, false>(ncand, ci_d, cj_d, d, ld, (const T*)U_d, (const T*)V_d, \
                                         (T*)scores_d, index_base, part, grid, s);               \
  } else {                                                                                       \
    rc = maximize ? score_normal<T, true>(criterion, ncand, ci_d, cj_d, d, nv, cutoff,           \
                                          (T*)scores_d, index_base, part, grid, s)               \
                  : score_normal<T, false>(criterion, ncand, ci_d, cj_d, d, nv, cutoff,          \
                                           (T*)scores_d, index_base, part, grid, s);             \
  }
  if (dtype == AMF_F32) { DISPATCH(float) } else { DISPATCH(double) }
#undef DISPATCH
  if (rc != AMF_OK) return rc;
  return launch_best_final(part, grid, maximize != 0, best_d, s);
}

constexpr int HOST_CHUNKS = 8;   // pieces the candidate arrays cross PCIe in (large pools)

// 16-bit item ids (pools over at most 65536 items cross PCIe at 2 bytes per candidate) -> the
// int32 ids the scoring kernel reads.  Pure streaming: 2 bytes in, 4 bytes out per candidate.
__global__ void __launch_bounds__(256) widen_u16_kernel(const uint16_t* __restrict__ in,
                                                        int64_t count, int32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride)
    out[t] = (int32_t)in[t];
}

// shared body of the host-buffer scoring calls: the candidate users come either as an array
// (ci_h) or as row offsets into cj_h (ptr_h, n+1 entries), expanded on the device
static int score_pred_host(int dtype, int64_t ncand, const int32_t* ci_h, const int64_t* ptr_h,
                           const int32_t* cj_h, const uint16_t* cj16_h, int32_t n, int32_t m, int d,
                           const void* U_h, const void* V_h, void* scores_h, int maximize,
                           amf_best_t* best_h) {
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_score_pred_host: bad dtype");
  AMF_REQUIRE(best_h && U_h && V_h, "amf_score_pred_host: NULL argument");
  AMF_REQUIRE(ncand == 0 || ((cj_h || cj16_h) && (ci_h || ptr_h)),
              "amf_score_pred_host: NULL candidate arrays");
  const size_t es = dtype == AMF_F32 ? 4 : 8;
  const int vecn = dtype == AMF_F32 ? 4 : 2;
  const int ld = (d + vecn - 1) / vecn * vecn;
  // grow-only staging buffers, one set per host thread and device (freed at process exit)
  struct Stage { void* p[8]; size_t n[8]; int dev; };
  static thread_local Stage st = {{nullptr}, {0}, -1};
  int dev = 0;
  AMF_CUDA(cudaGetDevice(&dev));
  if (st.dev != dev) {
    for (int q = 0; q < 8; ++q) { if (st.p[q]) cudaFree(st.p[q]); st.p[q] = nullptr; st.n[q] = 0; }
    st.dev = dev;
  }
  auto need = [&](int q, size_t bytes) -> int {
    if (st.n[q] >= bytes) return AMF_OK;
    if (st.p[q]) cudaFree(st.p[q]);
    st.p[q] = nullptr; st.n[q] = 0;
    AMF_CUDA(cudaMalloc(&st.p[q], bytes));
    st.n[q] = bytes;
    return AMF_OK;
  };
  const size_t nc = ncand > 0 ? (size_t)ncand : 1;
  int rc;
  if ((rc = need(0, (size_t)n * ld * es)) || (rc = need(1, (size_t)m * ld * es)) ||
      (rc = need(2, 4 * nc)) || (rc = need(3, 4 * nc)) ||
      (rc = need(4, sizeof(amf_best_t) * (HOST_CHUNKS + 1))) ||
      (scores_h && (rc = need(5, es * nc))) || (ptr_h && (rc = need(6, 8 * ((size_t)n + 1)))) ||
      (cj16_h && (rc = need(7, 2 * nc))))
    return rc;
  uint16_t* cj16_d = (uint16_t*)st.p[7];
  void *U_d = st.p[0], *V_d = st.p[1], *sc_d = scores_h ? st.p[5] : nullptr;
  int32_t *ci_d = (int32_t*)st.p[2], *cj_d = (int32_t*)st.p[3];
  amf_best_t* best_d = (amf_best_t*)st.p[4];          // [HOST_CHUNKS] per-chunk winners + result
  // copy stream + compute stream: the candidate arrays cross PCIe in HOST_CHUNKS pieces and
  // every piece is scored while the next one is still in flight
  static thread_local cudaStream_t s_copy = nullptr, s_comp = nullptr;
  static thread_local cudaEvent_t ev[HOST_CHUNKS + 1] = {nullptr};
  if (!s_copy) {
    AMF_CUDA(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
    AMF_CUDA(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
    for (int k = 0; k <= HOST_CHUNKS; ++k)
      AMF_CUDA(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
  }
  if (ld != d) {
    AMF_CUDA(cudaMemsetAsync(U_d, 0, (size_t)n * ld * es, s_copy));
    AMF_CUDA(cudaMemsetAsync(V_d, 0, (size_t)m * ld * es, s_copy));
  }
  AMF_CUDA(cudaMemcpy2DAsync(U_d, ld * es, U_h, d * es, d * es, n, cudaMemcpyHostToDevice, s_copy));
  AMF_CUDA(cudaMemcpy2DAsync(V_d, ld * es, V_h, d * es, d * es, m, cudaMemcpyHostToDevice, s_copy));
  if (ncand > 0 && ptr_h) {
    AMF_REQUIRE(ptr_h[0] == 0 && ptr_h[n] == ncand, "amf_score_pred_host_csr: row offsets must "
                "run from 0 to ncand");
    AMF_CUDA(cudaMemcpyAsync(st.p[6], ptr_h, 8 * ((size_t)n + 1), cudaMemcpyHostToDevice, s_copy));
  }
  AMF_CUDA(cudaEventRecord(ev[HOST_CHUNKS], s_copy));
  AMF_CUDA(cudaStreamWaitEvent(s_comp, ev[HOST_CHUNKS], 0));
  if (ncand > 0 && ptr_h) {
    expand_rows_kernel<int32_t><<<num_sms() * 8, 256, 0, s_comp>>>((const int64_t*)st.p[6], n, ci_d);
    AMF_LAUNCH_CHECK();
  }
  const int nchunks = ncand >= (int64_t)HOST_CHUNKS * (1 << 20) ? HOST_CHUNKS : 1;
  for (int k = 0; k < nchunks; ++k) {
    const int64_t lo = ncand * k / nchunks, hi = ncand * (k + 1) / nchunks;
    if (hi > lo) {
      if (!ptr_h)
        AMF_CUDA(cudaMemcpyAsync(ci_d + lo, ci_h + lo, 4 * (hi - lo), cudaMemcpyHostToDevice, s_copy));
      if (cj16_h)
        AMF_CUDA(cudaMemcpyAsync(cj16_d + lo, cj16_h + lo, 2 * (hi - lo), cudaMemcpyHostToDevice, s_copy));
      else
        AMF_CUDA(cudaMemcpyAsync(cj_d + lo, cj_h + lo, 4 * (hi - lo), cudaMemcpyHostToDevice, s_copy));
    }
    AMF_CUDA(cudaEventRecord(ev[k], s_copy));
    AMF_CUDA(cudaStreamWaitEvent(s_comp, ev[k], 0));
    if (cj16_h && hi > lo) {
      widen_u16_kernel<<<num_sms() * 8, 256, 0, s_comp>>>(cj16_d + lo, hi - lo, cj_d + lo);
      AMF_LAUNCH_CHECK();
    }
    rc = amf_score_candidates(AMF_CRIT_PRED, dtype, hi - lo, ci_d + lo, cj_d + lo, d, ld, U_d, V_d,
                              nullptr, 0.0, sc_d ? (char*)sc_d + es * lo : nullptr, maximize, lo,
                              best_d + 1 + k, s_comp);
    if (rc != AMF_OK) return rc;
  }
  rc = amf_best_reduce(best_d + 1, nchunks, maximize, best_d, s_comp);
  if (rc == AMF_OK) {
    if (scores_h && ncand > 0)
      AMF_CUDA(cudaMemcpyAsync(scores_h, sc_d, es * ncand, cudaMemcpyDeviceToHost, s_comp));
    AMF_CUDA(cudaMemcpyAsync(best_h, best_d, sizeof(amf_best_t), cudaMemcpyDeviceToHost, s_comp));
    AMF_CUDA(cudaStreamSynchronize(s_comp));
  }
  return rc;
}

int amf_score_pred_host(int dtype, int64_t ncand, const int32_t* ci_h, const int32_t* cj_h,
                        int32_t n, int32_t m, int d, const void* U_h, const void* V_h,
                        void* scores_h, int maximize, amf_best_t* best_h) {
  return score_pred_host(dtype, ncand, ci_h, nullptr, cj_h, nullptr, n, m, d, U_h, V_h, scores_h,
                         maximize, best_h);
}

int amf_score_pred_host_csr(int dtype, const int64_t* cand_ptr_h, const int32_t* cj_h, int32_t n,
                            int32_t m, int d, const void* U_h, const void* V_h, void* scores_h,
                            int maximize, amf_best_t* best_h) {
  AMF_REQUIRE(cand_ptr_h && n > 0, "amf_score_pred_host_csr: NULL row offsets");
  return score_pred_host(dtype, cand_ptr_h[n], nullptr, cand_ptr_h, cj_h, nullptr, n, m, d, U_h,
                         V_h, scores_h, maximize, best_h);
}

int amf_score_pred_host_csr16(int dtype, const int64_t* cand_ptr_h, const uint16_t* cj16_h,
                              int32_t n, int32_t m, int d, const void* U_h, const void* V_h,
                              void* scores_h, int maximize, amf_best_t* best_h) {
  AMF_REQUIRE(cand_ptr_h && n > 0, "amf_score_pred_host_csr16: NULL row offsets");
  AMF_REQUIRE(m <= 65536, "amf_score_pred_host_csr16: %d items do not fit 16-bit ids", m);
  return score_pred_host(dtype, cand_ptr_h[n], nullptr, cand_ptr_h, nullptr, cj16_h, n, m, d, U_h,
                         V_h, scores_h, maximize, best_h);
}

#pragma GCC visibility pop
}  // extern "C"
