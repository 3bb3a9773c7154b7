// Fused PMF objective + gradient (pmf_cy.pyx:170-193 log_likelihood, :204-223 gradient).
//
// The reference walks the (nnz,3) rating list once per call in Python.  Here one "side pass"
// streams one sorted copy of the list (8 B/entry: other index + rating), keeps the row being
// reduced (U_i and its gradient accumulator) in registers, gathers the other side's factor row
// (one 128 B line at d=32 fp32) and reduces
//     e = r - U_i.V_j - mean,   sum e^2,   dU_i += (e / sigma^2) V_j.
// The same kernel run on the item-major copy with the roles of U and V swapped gives dV, so no
// transposed scatter is needed and results do not depend on atomics ordering except for rows
// that straddle a 32-entry sub-chunk boundary (their partial sums are combined with RED.ADD).
//
// Thread mapping: LPR lanes cooperate on one rating, each owning VPL 16-byte vectors of the
// factor row (d=32 fp32 -> LPR=8, VPL=1: one float4 per lane, a warp gathers 4 full lines per
// load instruction).  Work is split by entries, not rows: a lane group owns sub-chunks of
// AMF_SUB=32 consecutive entries and finds its starting row in the precomputed sub_row table.
#include "common.cuh"

#ifndef AMF_SIDE_MIN_BLOCKS
#define AMF_SIDE_MIN_BLOCKS 4   // caps the side pass at 64 registers: 4 CTAs (32 warps) per SM
#endif

namespace amf {

// tiled.cu: the shared-memory-tiled copy of the list and the fused pass that runs on it
int tiled_prepare(amf_ratings* h, size_t row_bytes, const void* U, const void* V, const void* dU,
                  const void* dV, bool* use, cudaStream_t s);
template <typename T>
int tiled_loss_grad(const amf_ratings* h, int ld, const T* U, const T* V, T inv_sigma,
                    T mean_offset, T* dU, T* dV, double* sq_err, cudaStream_t s,
                    cudaEvent_t dU_done, int sides, int max_ctas);

template <typename T>
__global__ void __launch_bounds__(256)
prior_kernel(const T* __restrict__ X, int64_t count, T neg_inv_sigma, T* __restrict__ dX,
             double* __restrict__ norm2) {
  using V = typename Vec<T>::type;
  constexpr int N = Vec<T>::N;
  const int64_t nvec = count / N;
  double acc = 0;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nvec;
       t += (int64_t)gridDim.x * blockDim.x) {
    V x = reinterpret_cast<const V*>(X)[t];
    T s = vdot(x, x);
    acc += (double)s;
    if (dX) {
      V g = vzero(x);
      vfma(g, neg_inv_sigma, x);
      reinterpret_cast<V*>(dX)[t] = g;
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0 && norm2) atomicAdd(norm2, acc);
}

template <typename T>
__global__ void __launch_bounds__(256)
axpy_kernel(const T* __restrict__ X, const T* __restrict__ G, T lr, int64_t count,
            T* __restrict__ Xn) {
  using V = typename Vec<T>::type;
  constexpr int N = Vec<T>::N;
  const int64_t nvec = count / N;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nvec;
       t += (int64_t)gridDim.x * blockDim.x) {
    V x = reinterpret_cast<const V*>(X)[t];
    V g = reinterpret_cast<const V*>(G)[t];
    vfma(x, lr, g);
    reinterpret_cast<V*>(Xn)[t] = x;
  }
}

// momentum SGD update of fit_minibatches (pmf_cy.pyx:338-344):
//   inc = momentum * inc + scale * G ;  X += inc
template <typename T>
__global__ void __launch_bounds__(256)
momentum_kernel(T* __restrict__ inc, const T* __restrict__ G, T momentum, T scale, int64_t count,
                T* __restrict__ X) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count;
       t += (int64_t)gridDim.x * blockDim.x) {
    T v = fma(scale, G[t], momentum * inc[t]);
    inc[t] = v;
    X[t] += v;
  }
}

// loads E consecutive entries starting at a multiple of E (16-byte aligned vector loads; the
// same address is read by all lanes of a group, so a warp touches at most 2 lines per load)
template <int E>
__device__ __forceinline__ void load_idx(const int32_t* __restrict__ p, int32_t (&out)[E]) {
  if constexpr (E % 4 == 0) {
#pragma unroll
    for (int q = 0; q < E / 4; ++q) {
      const int4 v = __ldcs(reinterpret_cast<const int4*>(p) + q);
      out[4 * q] = v.x; out[4 * q + 1] = v.y; out[4 * q + 2] = v.z; out[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int q = 0; q < E / 2; ++q) {
      const int2 v = __ldcs(reinterpret_cast<const int2*>(p) + q);
      out[2 * q] = v.x; out[2 * q + 1] = v.y;
    }
  }
}
template <int E>
__device__ __forceinline__ void load_val(const float* __restrict__ p, float (&out)[E]) {
  if constexpr (E % 4 == 0) {
#pragma unroll
    for (int q = 0; q < E / 4; ++q) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(p) + q);
      out[4 * q] = v.x; out[4 * q + 1] = v.y; out[4 * q + 2] = v.z; out[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int q = 0; q < E / 2; ++q) {
      const float2 v = __ldcs(reinterpret_cast<const float2*>(p) + q);
      out[2 * q] = v.x; out[2 * q + 1] = v.y;
    }
  }
}
template <int E>
__device__ __forceinline__ void load_val(const double* __restrict__ p, double (&out)[E]) {
#pragma unroll
  for (int q = 0; q < E / 2; ++q) {
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p) + q);
    out[2 * q] = v.x; out[2 * q + 1] = v.y;
  }
}

// One side of the fused loss+gradient.  A lane group owns sub-chunks of AMF_SUB=32 consecutive
// entries of the sorted list and walks them in batches of E: the E (index, rating) pairs come
// in with two vector loads, the E factor rows of the other side are gathered back to back
// (E independent 16-byte loads per lane in flight), then the batch is reduced in order so the
// row accumulator and the row-change flush see the entries sequentially.
template <typename T, int LPR, int VPL, bool GRAD>
__global__ void __launch_bounds__(256, AMF_SIDE_MIN_BLOCKS)
side_pass_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                 const T* __restrict__ val, const int32_t* __restrict__ sub_row,
                 const T* __restrict__ Self, const T* __restrict__ Other, int ld, int nvec,
                 T inv_sigma, T mean_offset, T* __restrict__ dSelf,
                 double* __restrict__ sq_err, int64_t nnz, int64_t n_sub) {
  using V = typename Vec<T>::type;
  constexpr int N = Vec<T>::N;
  constexpr int G = 32 / LPR;                  // lane groups per warp
  constexpr int E = (VPL == 1) ? 8 : 2;        // entries per batch
  const int lane = threadIdx.x & 31;
  const int g = lane / LPR, l = lane % LPR;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;

  bool have[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) have[v] = (l + v * LPR) < nvec;

  double local_sq = 0;
  for (int64_t base = warp * G; base < n_sub; base += nwarps * G) {
    T sub_sq = 0;
    const int64_t sub = base + g;
    const bool live = sub < n_sub;
    const int64_t p0 = live ? sub * AMF_SUB : 0;
    // everything below is relative to p0 so the inner loop runs on 32-bit offsets
    const int cnt = live ? (int)min((int64_t)AMF_SUB, nnz - p0) : 0;
    int32_t row = live ? sub_row[sub] : 0;
    int row_end = live ? (int)min(ptr[row + 1] - p0, (int64_t)(AMF_SUB + 1)) : AMF_SUB + 1;
    const int32_t* __restrict__ idx0 = idx + p0;
    const T* __restrict__ val0 = val + p0;
    V self[VPL], acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      acc[v] = vzero(V());
      self[v] = (live && have[v])
                    ? reinterpret_cast<const V*>(Self + (int64_t)row * ld)[l + v * LPR]
                    : vzero(V());
    }
    const bool full = __all_sync(0xffffffffu, cnt == AMF_SUB);   // warp-uniform fast path
    for (int b = 0; b < AMF_SUB; b += E) {
      int32_t js[E];
      T rs[E];
      if (full || b + E <= cnt) {
        load_idx<E>(idx0 + b, js);
        load_val<E>(val0 + b, rs);
      } else {
#pragma unroll
        for (int s = 0; s < E; ++s) {
          const bool ok = b + s < cnt;
          js[s] = ok ? ld_stream(idx0 + b + s) : 0;
          rs[s] = ok ? ld_stream(val0 + b + s) : T(0);
        }
      }
      V o[E][VPL];
#pragma unroll
      for (int s = 0; s < E; ++s)
#pragma unroll
        for (int v = 0; v < VPL; ++v)
          o[s][v] = ((full || b + s < cnt) && have[v])
                        ? reinterpret_cast<const V*>(Other + (int64_t)js[s] * ld)[l + v * LPR]
                        : vzero(V());
#pragma unroll
      for (int s = 0; s < E; ++s) {
        const int t = b + s;
        const bool valid = full || t < cnt;
        if (valid && t >= row_end) {
          // leave the finished row: publish its partial gradient, move to the row holding t
          if (GRAD) {
#pragma unroll
            for (int v = 0; v < VPL; ++v)
              if (have[v]) vred_add(dSelf + (int64_t)row * ld + (l + v * LPR) * N, acc[v]);
          }
          int64_t re;
          do { ++row; re = ptr[row + 1] - p0; } while ((int64_t)t >= re);
          row_end = (int)min(re, (int64_t)(AMF_SUB + 1));
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            acc[v] = vzero(V());
            if (have[v]) self[v] = reinterpret_cast<const V*>(Self + (int64_t)row * ld)[l + v * LPR];
          }
        }
        T dot = 0;
#pragma unroll
        for (int v = 0; v < VPL; ++v) dot += vdot(self[v], o[s][v]);
#pragma unroll
        for (int off = LPR >> 1; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
        const T e = valid ? (rs[s] - mean_offset) - dot : T(0);
        sub_sq = fma(e, e, sub_sq);
        if (GRAD) {
          const T w = e * inv_sigma;
#pragma unroll
          for (int v = 0; v < VPL; ++v) vfma(acc[v], w, o[s][v]);
        }
      }
    }
    if (GRAD && cnt > 0) {
#pragma unroll
      for (int v = 0; v < VPL; ++v)
        if (have[v]) vred_add(dSelf + (int64_t)row * ld + (l + v * LPR) * N, acc[v]);
    }
    if (l == 0) local_sq += (double)sub_sq;
  }
  if (sq_err) {
    double s = block_sum(local_sq);
    if (threadIdx.x == 0) atomicAdd(sq_err, s);
  }
}

// COO mini-batch variant with atomics into both sides (gradient(ratings=batch))
template <typename T, int LPR, int VPL>
__global__ void __launch_bounds__(256)
coo_grad_kernel(const int32_t* __restrict__ ci, const int32_t* __restrict__ cj,
                const T* __restrict__ val, int64_t nnz, const T* __restrict__ U,
                const T* __restrict__ Vm, int ld, int nvec, T inv_sigma, T mean_offset,
                T* __restrict__ dU, T* __restrict__ dV, double* __restrict__ sq_err) {
  using V = typename Vec<T>::type;
  constexpr int N = Vec<T>::N;
  constexpr int G = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int g = lane / LPR, l = lane % LPR;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  T local_sq = 0;
  for (int64_t base = warp * G; base < nnz; base += nwarps * G) {
    const int64_t p = base + g;
    const bool valid = p < nnz;
    const int32_t i = valid ? ci[p] : 0, j = valid ? cj[p] : 0;
    const T r = valid ? val[p] : T(0);
    V a[VPL], b[VPL];
    T dot = 0;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const bool h = valid && (l + v * LPR) < nvec;
      a[v] = h ? reinterpret_cast<const V*>(U + (int64_t)i * ld)[l + v * LPR] : vzero(V());
      b[v] = h ? reinterpret_cast<const V*>(Vm + (int64_t)j * ld)[l + v * LPR] : vzero(V());
      dot += vdot(a[v], b[v]);
    }
#pragma unroll
    for (int off = LPR >> 1; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
    const T e = valid ? (r - dot - mean_offset) : T(0);
    if (l == 0) local_sq = fma(e, e, local_sq);
    const T w = e * inv_sigma;
    if (valid && dU) {
#pragma unroll
      for (int v = 0; v < VPL; ++v)
        if ((l + v * LPR) < nvec) {
          V gu = vzero(V()), gv = vzero(V());
          vfma(gu, w, b[v]);
          vfma(gv, w, a[v]);
          vred_add(dU + (int64_t)i * ld + (l + v * LPR) * N, gu);
          vred_add(dV + (int64_t)j * ld + (l + v * LPR) * N, gv);
        }
    }
  }
  if (sq_err) {
    double s = block_sum((double)local_sq);
    if (threadIdx.x == 0) atomicAdd(sq_err, s);
  }
}

static inline int pow2_ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// grid: enough warps to cover the sub-chunks, capped at a multiple of the SM count
static inline int grid_for(int64_t units_per_block_total, int blocks_per_sm_cap) {
  int64_t want = units_per_block_total;
  int64_t cap = (int64_t)num_sms() * blocks_per_sm_cap;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

template <typename T, bool GRAD>
static int launch_side(const amf_ratings* h, int side, const T* Self, const T* Other, int ld,
                       T inv_sigma, T mean_offset, T* dSelf, double* sq_err, cudaStream_t s) {
  constexpr int N = Vec<T>::N;
  const int nvec = ld / N;
  int lpr = pow2_ceil(nvec);
  int vpl = 1;
  if (lpr > 32) { vpl = lpr / 32; lpr = 32; }
  if (vpl > 4) { set_error("latent dimension too large (ld=%d)", ld); return AMF_ERR_UNSUPPORTED; }
  const int G = 32 / lpr;
  const int64_t warps_needed = (h->n_sub + G - 1) / G;
  const int grid = grid_for((warps_needed + 7) / 8, 8);
  if (GRAD) AMF_DBG_RANGE(0, dSelf, sizeof(T) * (size_t)(side == 0 ? h->n_users : h->n_items) * ld, s);
#define SIDE(LPR_, VPL_)                                                                      \
  side_pass_kernel<T, LPR_, VPL_, GRAD><<<grid, 256, 0, s>>>(                                 \
      h->ptr[side], h->idx[side], (const T*)h->val[side], h->sub_row[side], Self, Other, ld,  \
      nvec, inv_sigma, mean_offset, dSelf, sq_err, h->nnz, h->n_sub)
  if (vpl == 1) {
    switch (lpr) {
      case 1: SIDE(1, 1); break;
      case 2: SIDE(2, 1); break;
      case 4: SIDE(4, 1); break;
      case 8: SIDE(8, 1); break;
      case 16: SIDE(16, 1); break;
      default: SIDE(32, 1); break;
    }
  } else if (vpl == 2) {
    SIDE(32, 2);
  } else {
    SIDE(32, 4);
  }
#undef SIDE
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

template <typename T>
static int grad_coo(int64_t nnz, const int32_t* i_d, const int32_t* j_d, const T* r_d, int d,
                    int ld, const T* U, const T* V, const amf_pmf_params_t* p, T* dU, T* dV,
                    double* sums, cudaStream_t s, bool reset_sums = true);

// the data terms of the ratings appended since the sorted lists were built (amf_ratings_append):
// COO kernel with atomics into both sides, on top of what the sorted passes wrote
template <typename T>
static int tail_terms(const amf_ratings* h, int d, int ld, const T* U, const T* V,
                      const amf_pmf_params_t* p, T* dU, T* dV, double* sums, cudaStream_t s) {
  if (h->tail_n == 0) return AMF_OK;
  return grad_coo<T>(h->tail_n, h->tail_i, h->tail_j, (const T*)h->tail_r, d, ld, U, V, p, dU, dV,
                     sums, s, false);
}

// dU_done (optional): recorded on s as soon as dU holds its final value (before the pass that
// produces dV), so a host-buffer caller can start copying dU back while dV is computed.
// parts: bit 0 = prior terms of both sides + the pass that completes dU and the squared error
// (+ the appended tail), bit 1 = the pass that completes dV; a caller that runs them as two
// calls can put a collective on dU in between.  max_ctas > 0 caps the grid of the tiled passes.
template <typename T>
static int loss_grad(const amf_ratings* h, int d, int ld, const T* U, const T* V,
                     const amf_pmf_params_t* p, T* dU, T* dV, double* sums, cudaStream_t s,
                     cudaEvent_t dU_done = nullptr, int parts = 3, int max_ctas = 0) {
  constexpr int N = Vec<T>::N;
  AMF_REQUIRE(ld >= d && ld % N == 0, "ld=%d must be >= d=%d and a multiple of %d", ld, d, N);
  AMF_REQUIRE((dU == nullptr) == (dV == nullptr), "dU and dV must both be given or both NULL");
  AMF_REQUIRE(parts >= 1 && parts <= 3, "parts must be 1, 2 or 3");
  if (parts & 1) {
    AMF_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(double), s));
    const int64_t cu = (int64_t)h->n_users * ld, cv = (int64_t)h->n_items * ld;
    const int gu = grid_for((cu / N + 255) / 256, 8), gv = grid_for((cv / N + 255) / 256, 8);
    prior_kernel<T><<<gu, 256, 0, s>>>(U, cu, (T)(-1.0 / p->sigma_u_sq), dU, sums + 1);
    AMF_LAUNCH_CHECK();
    prior_kernel<T><<<gv, 256, 0, s>>>(V, cv, (T)(-1.0 / p->sigma_v_sq), dV, sums + 2);
    AMF_LAUNCH_CHECK();
  }
  // the appended tail (if any) touches dU last: then dU is final only at the very end
  cudaEvent_t early = h->tail_n > 0 ? nullptr : dU_done;
  int rc = AMF_OK;
  if (h->nnz > 0) {
    const T inv_sigma = (T)(1.0 / p->sigma_sq), mo = (T)p->mean_offset;
    bool tiled = false;
    rc = tiled_prepare(const_cast<amf_ratings*>(h), (size_t)ld * sizeof(T), U, V, dU, dV, &tiled, s);
    if (rc != AMF_OK) return rc;
    if (tiled) {
      rc = tiled_loss_grad<T>(h, ld, U, V, inv_sigma, mo, dU, dV, sums, s, early, parts, max_ctas);
    } else if (dU) {
      if (parts & 1) {
        rc = launch_side<T, true>(h, 0, U, V, ld, inv_sigma, mo, dU, sums, s);
        if (rc != AMF_OK) return rc;
        if (early) AMF_CUDA(cudaEventRecord(early, s));
      }
      if (parts & 2) rc = launch_side<T, true>(h, 1, V, U, ld, inv_sigma, mo, dV, nullptr, s);
    } else if (parts & 1) {
      rc = launch_side<T, false>(h, 0, U, V, ld, inv_sigma, mo, nullptr, sums, s);
    }
    if (rc != AMF_OK) return rc;
  }
  if (parts & 1) rc = tail_terms<T>(h, d, ld, U, V, p, dU, dV, sums, s);
  if (rc == AMF_OK && dU_done && (parts & 1) && (!early || !dU || h->nnz == 0))
    AMF_CUDA(cudaEventRecord(dU_done, s));
  return rc;
}

template <typename T>
static int grad_coo(int64_t nnz, const int32_t* i_d, const int32_t* j_d, const T* r_d, int d,
                    int ld, const T* U, const T* V, const amf_pmf_params_t* p, T* dU, T* dV,
                    double* sums, cudaStream_t s, bool reset_sums) {
  constexpr int N = Vec<T>::N;
  AMF_REQUIRE(ld >= d && ld % N == 0, "ld=%d must be >= d=%d and a multiple of %d", ld, d, N);
  if (sums && reset_sums) AMF_CUDA(cudaMemsetAsync(sums, 0, sizeof(double), s));
  if (nnz == 0) return AMF_OK;
  const int nvec = ld / N;
  int lpr = pow2_ceil(nvec), vpl = 1;
  if (lpr > 32) { vpl = lpr / 32; lpr = 32; }
  if (vpl > 4) { set_error("latent dimension too large (ld=%d)", ld); return AMF_ERR_UNSUPPORTED; }
  const int G = 32 / lpr;
  const int grid = grid_for((((nnz + G - 1) / G) + 7) / 8, 8);
  const T inv_sigma = (T)(1.0 / p->sigma_sq), mo = (T)p->mean_offset;
#ifdef AMF_BOUNDS_CHECK
  // the mini-batch entry does not know the table heights: the caller's ids were range-checked
  // when the list was built, so the widest legal write is bounded by the largest id present
  {
    AMF_DBG_RANGE(0, dU, (size_t)1 << 46, s);
    AMF_DBG_RANGE(1, dV, (size_t)1 << 46, s);
  }
#endif
#define COO(LPR_, VPL_)                                                                        \
  coo_grad_kernel<T, LPR_, VPL_><<<grid, 256, 0, s>>>(i_d, j_d, r_d, nnz, U, V, ld, nvec,      \
                                                      inv_sigma, mo, dU, dV, sums)
  if (vpl == 1) {
    switch (lpr) {
      case 1: COO(1, 1); break;
      case 2: COO(2, 1); break;
      case 4: COO(4, 1); break;
      case 8: COO(8, 1); break;
      case 16: COO(16, 1); break;
      default: COO(32, 1); break;
    }
  } else if (vpl == 2) {
    COO(32, 2);
  } else {
    COO(32, 4);
  }
#undef COO
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

// (rows, d) tightly packed host matrix <-> zero-padded (rows, ld) device matrix
static int ensure_stage(amf_ratings* h, int slot, size_t bytes) {
  if (h->stage_bytes[slot] >= bytes) return AMF_OK;
  cudaFree(h->stage[slot]);
  h->stage[slot] = nullptr;
  h->stage_bytes[slot] = 0;
  AMF_CUDA(cudaMalloc(&h->stage[slot], bytes));
  h->stage_bytes[slot] = bytes;
  return AMF_OK;
}

}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int amf_pmf_loss_grad(const amf_ratings_t* h, int dtype, int d, int ld, const void* U_d,
                      const void* V_d, const amf_pmf_params_t* p, void* dU_d, void* dV_d,
                      double* sums_d, void* stream) {
  AMF_REQUIRE(h && U_d && V_d && p && sums_d, "amf_pmf_loss_grad: NULL argument");
  AMF_REQUIRE(dtype == h->dtype, "amf_pmf_loss_grad: dtype %d does not match the rating list's %d",
              dtype, h->dtype);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == AMF_F32)
    return loss_grad<float>(h, d, ld, (const float*)U_d, (const float*)V_d, p, (float*)dU_d,
                            (float*)dV_d, sums_d, s);
  return loss_grad<double>(h, d, ld, (const double*)U_d, (const double*)V_d, p, (double*)dU_d,
                           (double*)dV_d, sums_d, s);
}

int amf_pmf_loss_grad_part(const amf_ratings_t* h, int dtype, int d, int ld, const void* U_d,
                           const void* V_d, const amf_pmf_params_t* p, void* dU_d, void* dV_d,
                           double* sums_d, int part, int max_ctas, void* stream) {
  AMF_REQUIRE(h && U_d && V_d && p && sums_d && dU_d && dV_d, "amf_pmf_loss_grad_part: NULL argument");
  AMF_REQUIRE(dtype == h->dtype, "amf_pmf_loss_grad_part: dtype does not match the rating list");
  AMF_REQUIRE(part == 0 || part == 1, "amf_pmf_loss_grad_part: part must be 0 or 1");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == AMF_F32)
    return loss_grad<float>(h, d, ld, (const float*)U_d, (const float*)V_d, p, (float*)dU_d,
                            (float*)dV_d, sums_d, s, nullptr, 1 << part, max_ctas);
  return loss_grad<double>(h, d, ld, (const double*)U_d, (const double*)V_d, p, (double*)dU_d,
                           (double*)dV_d, sums_d, s, nullptr, 1 << part, max_ctas);
}

int amf_axpy(int dtype, int64_t count, const void* X_d, const void* G_d, double lr, void* Xnew_d,
             void* stream) {
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_axpy: bad dtype");
  cudaStream_t s = (cudaStream_t)stream;
  if (count == 0) return AMF_OK;
  if (dtype == AMF_F32) {
    AMF_REQUIRE(count % 4 == 0, "amf_axpy: count must be a multiple of 4");
    axpy_kernel<float><<<grid_for((count / 4 + 255) / 256, 8), 256, 0, s>>>(
        (const float*)X_d, (const float*)G_d, (float)lr, count, (float*)Xnew_d);
  } else {
    AMF_REQUIRE(count % 2 == 0, "amf_axpy: count must be a multiple of 2");
    axpy_kernel<double><<<grid_for((count / 2 + 255) / 256, 8), 256, 0, s>>>(
        (const double*)X_d, (const double*)G_d, lr, count, (double*)Xnew_d);
  }
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

int amf_momentum_step(int dtype, int64_t count, void* inc_d, const void* G_d, double momentum,
                      double scale, void* X_d, void* stream) {
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_momentum_step: bad dtype");
  cudaStream_t s = (cudaStream_t)stream;
  if (count == 0) return AMF_OK;
  const int grid = grid_for((count + 255) / 256, 8);
  if (dtype == AMF_F32)
    momentum_kernel<float><<<grid, 256, 0, s>>>((float*)inc_d, (const float*)G_d, (float)momentum,
                                                (float)scale, count, (float*)X_d);
  else
    momentum_kernel<double><<<grid, 256, 0, s>>>((double*)inc_d, (const double*)G_d, momentum,
                                                 scale, count, (double*)X_d);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

int amf_pmf_prior(int dtype, int64_t count, const void* X_d, double sigma_x_sq, void* dX_d,
                  double* norm2_d, void* stream) {
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_pmf_prior: bad dtype");
  cudaStream_t s = (cudaStream_t)stream;
  if (count == 0) return AMF_OK;
  if (dtype == AMF_F32) {
    AMF_REQUIRE(count % 4 == 0, "amf_pmf_prior: count must be a multiple of 4");
    prior_kernel<float><<<grid_for((count / 4 + 255) / 256, 8), 256, 0, s>>>(
        (const float*)X_d, count, (float)(-1.0 / sigma_x_sq), (float*)dX_d, norm2_d);
  } else {
    AMF_REQUIRE(count % 2 == 0, "amf_pmf_prior: count must be a multiple of 2");
    prior_kernel<double><<<grid_for((count / 2 + 255) / 256, 8), 256, 0, s>>>(
        (const double*)X_d, count, -1.0 / sigma_x_sq, (double*)dX_d, norm2_d);
  }
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

int amf_pmf_grad_coo(int dtype, int64_t nnz, const int32_t* i_d, const int32_t* j_d,
                     const void* r_d, int d, int ld, const void* U_d, const void* V_d,
                     const amf_pmf_params_t* p, void* dU_d, void* dV_d, double* sums_d,
                     void* stream) {
  AMF_REQUIRE(U_d && V_d && p, "amf_pmf_grad_coo: NULL argument");
  AMF_REQUIRE((dU_d == nullptr) == (dV_d == nullptr), "dU and dV must both be given or both NULL");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == AMF_F32)
    return grad_coo<float>(nnz, i_d, j_d, (const float*)r_d, d, ld, (const float*)U_d,
                           (const float*)V_d, p, (float*)dU_d, (float*)dV_d, sums_d, s);
  AMF_REQUIRE(dtype == AMF_F64, "amf_pmf_grad_coo: bad dtype");
  return grad_coo<double>(nnz, i_d, j_d, (const double*)r_d, d, ld, (const double*)U_d,
                          (const double*)V_d, p, (double*)dU_d, (double*)dV_d, sums_d, s);
}

int amf_pmf_loss_grad_host(const amf_ratings_t* hc, int dtype, int d, const void* U_h,
                           const void* V_h, const amf_pmf_params_t* p, void* dU_h, void* dV_h,
                           double* sums_h) {
  AMF_REQUIRE(hc && U_h && V_h && p && sums_h, "amf_pmf_loss_grad_host: NULL argument");
  AMF_REQUIRE(dtype == hc->dtype, "amf_pmf_loss_grad_host: dtype mismatch");
  AMF_REQUIRE((dU_h == nullptr) == (dV_h == nullptr), "dU and dV must both be given or both NULL");
  amf_ratings* h = const_cast<amf_ratings*>(hc);
  const size_t es = dtype == AMF_F32 ? 4 : 8;
  const int vecn = dtype == AMF_F32 ? 4 : 2;
  const int ld = (d + vecn - 1) / vecn * vecn;
  const size_t bu = (size_t)h->n_users * ld * es, bv = (size_t)h->n_items * ld * es;
  int rc;
  if ((rc = ensure_stage(h, 0, bu)) || (rc = ensure_stage(h, 1, bv))) return rc;
  if (dU_h && ((rc = ensure_stage(h, 2, bu)) || (rc = ensure_stage(h, 3, bv)))) return rc;
  // two streams: dU goes back to the host while the second pass is still producing dV.  They are
  // cached per host thread and belong to ONE device: a thread that moves to another GPU gets new
  // ones, and the handle must live on the current device.  Work the caller enqueued on its own
  // streams for this handle (amf_ratings_append, ...) must be synchronised before this call.
  static thread_local cudaStream_t s = nullptr, s_copy = nullptr;
  static thread_local cudaEvent_t ev_dU = nullptr, ev_copied = nullptr;
  static thread_local int cached_dev = -1;
  int dev = -1;
  AMF_CUDA(cudaGetDevice(&dev));
  AMF_REQUIRE(dev == h->device, "amf_pmf_loss_grad_host: the rating list lives on device %d, the "
              "current device is %d", h->device, dev);
  if (cached_dev != dev) {
    if (s) { cudaStreamDestroy(s); cudaStreamDestroy(s_copy); cudaEventDestroy(ev_dU); cudaEventDestroy(ev_copied); }
    s = s_copy = nullptr; ev_dU = ev_copied = nullptr;
    AMF_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    AMF_CUDA(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
    AMF_CUDA(cudaEventCreateWithFlags(&ev_dU, cudaEventDisableTiming));
    AMF_CUDA(cudaEventCreateWithFlags(&ev_copied, cudaEventDisableTiming));
    cached_dev = dev;
  }
  if (ld != d) {
    AMF_CUDA(cudaMemsetAsync(h->stage[0], 0, bu, s));
    AMF_CUDA(cudaMemsetAsync(h->stage[1], 0, bv, s));
  }
  AMF_CUDA(cudaMemcpy2DAsync(h->stage[0], ld * es, U_h, d * es, d * es, h->n_users,
                             cudaMemcpyHostToDevice, s));
  AMF_CUDA(cudaMemcpy2DAsync(h->stage[1], ld * es, V_h, d * es, d * es, h->n_items,
                             cudaMemcpyHostToDevice, s));
  void* dU_d = dU_h ? h->stage[2] : nullptr;
  void* dV_d = dU_h ? h->stage[3] : nullptr;
  if (dtype == AMF_F32)
    rc = loss_grad<float>(h, d, ld, (const float*)h->stage[0], (const float*)h->stage[1], p,
                          (float*)dU_d, (float*)dV_d, h->sums_d, s, dU_h ? ev_dU : nullptr);
  else
    rc = loss_grad<double>(h, d, ld, (const double*)h->stage[0], (const double*)h->stage[1], p,
                           (double*)dU_d, (double*)dV_d, h->sums_d, s, dU_h ? ev_dU : nullptr);
  if (rc != AMF_OK) return rc;
  if (dU_h) {
    AMF_CUDA(cudaStreamWaitEvent(s_copy, ev_dU, 0));
    AMF_CUDA(cudaMemcpy2DAsync(dU_h, d * es, h->stage[2], ld * es, d * es, h->n_users,
                               cudaMemcpyDeviceToHost, s_copy));
    AMF_CUDA(cudaEventRecord(ev_copied, s_copy));
    AMF_CUDA(cudaMemcpy2DAsync(dV_h, d * es, h->stage[3], ld * es, d * es, h->n_items,
                               cudaMemcpyDeviceToHost, s));
    AMF_CUDA(cudaStreamWaitEvent(s, ev_copied, 0));
  }
  AMF_CUDA(cudaMemcpyAsync(sums_h, h->sums_d, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  AMF_CUDA(cudaStreamSynchronize(s));
  return AMF_OK;
}

#pragma GCC visibility pop
}  // extern "C"
