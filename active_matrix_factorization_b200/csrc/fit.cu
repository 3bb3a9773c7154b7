// Whole line-search fit on the device for SMALL problems (pmf_cy.pyx:257-305 fit_lls / fit).
//
// The host-driven fit (pmf_cy.py fit_lls) costs one fused loss+gradient launch, two axpy
// launches and one 24-byte read-back per trial: fine when a trial is milliseconds of kernel
// time, but at the reference's own sizes (10x10 ... movielens-100k with a few thousand ratings)
// a trial is microseconds of work behind ~100 us of launches and synchronisation.  Here ONE
// cooperative launch runs the entire loop -- trial point, objective, gradient, accept / reject,
// step-size update, convergence test -- with the reference's control flow and fp64 step-size
// arithmetic, the CTAs meeting at a grid-wide barrier three times per trial, and returns the
// accepted-step objective trace.  No host round trip inside the fit.
#include <algorithm>

#include "common.cuh"

namespace amf {

constexpr int FIT_THREADS = 256;

// ---- grid-wide barrier for a cooperative launch (all CTAs resident) --------------------------
// Monotonic arrival counter: the e-th barrier is passed once (e+1) * gridDim.x CTAs arrived.
// The tables the CTAs exchange are read with ld.global.cg (L2) everywhere in this file: the
// same addresses are rewritten by other SMs every trial, so an L1 copy could be stale.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int target = (epoch + 1) * gridDim.x;
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
    __threadfence();
  }
  ++epoch;
  __syncthreads();
}

struct FitState {            // in the workspace, zeroed before the launch
  double sums[3][4];         // rotating accumulators {sum e^2, |U|^2, |V|^2, -}
  unsigned int barrier;
  unsigned int pad_[7];
};

// G = -X / sigma_x^2 (prior term) and this thread's share of |X|^2
template <typename T>
__device__ double prior_init(const T* X, int64_t count, T neg_inv_sigma, T* G) {
  double acc = 0;
  for (int64_t t = blockIdx.x * (int64_t)FIT_THREADS + threadIdx.x; t < count;
       t += (int64_t)gridDim.x * FIT_THREADS) {
    const T x = __ldcg(X + t);
    acc += (double)(x * x);
    G[t] = neg_inv_sigma * x;
  }
  return acc;
}

// objective and gradient at (U, V), by the whole grid; returns the log-likelihood to every
// thread of every CTA (the same bits everywhere: all read the same three sums)
template <typename T>
__device__ double loss_grad_grid(int64_t nnz, const int32_t* __restrict__ own,
                                 const int32_t* __restrict__ idx, const T* __restrict__ val,
                                 int32_t n, int32_t m, int d, int ld, const T* U, const T* V,
                                 T* gU, T* gV, double sigma_sq, double sigma_u_sq,
                                 double sigma_v_sq, T mean_offset, FitState* st, unsigned int& epoch,
                                 int& slot, double* s_ll) {
  const double nu = prior_init<T>(U, (int64_t)n * ld, (T)(-1.0 / sigma_u_sq), gU);
  const double nv = prior_init<T>(V, (int64_t)m * ld, (T)(-1.0 / sigma_v_sq), gV);
  grid_barrier(&st->barrier, epoch);             // prior terms are in place everywhere
  if (blockIdx.x == 0 && threadIdx.x == 0) {     // the slot after next is free: clear it
    double* z = st->sums[(slot + 1) % 3];
    z[0] = z[1] = z[2] = 0.0;
  }
  const T inv_sigma = (T)(1.0 / sigma_sq);
  double sq = 0;
  for (int64_t p = blockIdx.x * (int64_t)FIT_THREADS + threadIdx.x; p < nnz;
       p += (int64_t)gridDim.x * FIT_THREADS) {
    const int32_t i = own[p], j = idx[p];
    AMF_DBG_ASSERT((uint32_t)i < (uint32_t)n && (uint32_t)j < (uint32_t)m);
    const T* u = U + (int64_t)i * ld;
    const T* v = V + (int64_t)j * ld;
    T dot = 0;
    for (int k = 0; k < d; ++k) dot = fma(__ldcg(u + k), __ldcg(v + k), dot);
    const T e = (val[p] - mean_offset) - dot;
    sq += (double)(e * e);
    const T w = e * inv_sigma;
    for (int k = 0; k < d; ++k) {
      atomicAdd(gU + (int64_t)i * ld + k, w * __ldcg(v + k));
      atomicAdd(gV + (int64_t)j * ld + k, w * __ldcg(u + k));
    }
  }
  const double s0 = block_sum(sq), s1 = block_sum(nu), s2 = block_sum(nv);
  if (threadIdx.x == 0) {
    double* acc = st->sums[slot];
    atomicAdd(acc + 0, s0);
    atomicAdd(acc + 1, s1);
    atomicAdd(acc + 2, s2);
  }
  grid_barrier(&st->barrier, epoch);             // gradient and sums are complete
  if (threadIdx.x == 0) {
    const double* acc = st->sums[slot];
    const double t0 = __ldcg(acc + 0), t1 = __ldcg(acc + 1), t2 = __ldcg(acc + 2);
    *s_ll = -t0 / (2. * sigma_sq) - t1 / (2. * sigma_u_sq) - t2 / (2. * sigma_v_sq);
  }
  slot = (slot + 1) % 3;
  __syncthreads();
  const double ll = *s_ll;
  __syncthreads();
  return ll;
}

template <typename T>
__global__ void __launch_bounds__(FIT_THREADS)
fit_lls_kernel(int64_t nnz, const int32_t* __restrict__ own, const int32_t* __restrict__ idx,
               const T* __restrict__ val, int32_t n, int32_t m, int d, int ld, T* U, T* V, T* U2,
               T* V2, T* gU, T* gV, T* gU2, T* gV2, double sigma_sq, double sigma_u_sq,
               double sigma_v_sq, double mean_offset, double lr, double min_lr,
               double stop_thresh, int max_steps, double* __restrict__ trace, int trace_cap,
               amf_fit_result_t* __restrict__ result, FitState* st) {
  __shared__ double s_ll;
  T* const U_out = U;
  T* const V_out = V;
  const T mo = (T)mean_offset;
  const int64_t cu = (int64_t)n * ld, cv = (int64_t)m * ld;
  const int64_t t_first = blockIdx.x * (int64_t)FIT_THREADS + threadIdx.x;
  const int64_t t_step = (int64_t)gridDim.x * FIT_THREADS;
  unsigned int epoch = 0;
  int slot = 0;
  double old_ll = loss_grad_grid<T>(nnz, own, idx, val, n, m, d, ld, U, V, gU, gV, sigma_sq,
                                    sigma_u_sq, sigma_v_sq, mo, st, epoch, slot, &s_ll);
  int steps = 0, trials = 0;
  bool converged = false;
  while (!converged && (max_steps <= 0 || steps < max_steps)) {
    for (;;) {
      const T lrT = (T)lr;
      for (int64_t t = t_first; t < cu; t += t_step) U2[t] = fma(lrT, __ldcg(gU + t), __ldcg(U + t));
      for (int64_t t = t_first; t < cv; t += t_step) V2[t] = fma(lrT, __ldcg(gV + t), __ldcg(V + t));
      // (the barrier after the prior terms inside loss_grad_grid also publishes U2 / V2: no CTA
      // gathers rows of the trial point before it)
      const double new_ll = loss_grad_grid<T>(nnz, own, idx, val, n, m, d, ld, U2, V2, gU2, gV2,
                                              sigma_sq, sigma_u_sq, sigma_v_sq, mo, st, epoch,
                                              slot, &s_ll);
      ++trials;
      if (new_ll > old_ll) {                 // accept: the trial point becomes the iterate
        T* t0;
        t0 = U; U = U2; U2 = t0;   t0 = V; V = V2; V2 = t0;
        t0 = gU; gU = gU2; gU2 = t0;   t0 = gV; gV = gV2; gV2 = t0;
        lr *= 1.25;
        if (new_ll - old_ll < stop_thresh) converged = true;
        if (blockIdx.x == 0 && threadIdx.x == 0 && steps < trace_cap) trace[steps] = new_ll;
        ++steps;
        old_ll = new_ll;
        break;
      }
      lr *= .5;
      if (lr < min_lr) { converged = true; break; }
    }
  }
  if (U != U_out) {                          // an odd number of accepted steps: copy back
    for (int64_t t = t_first; t < cu; t += t_step) U_out[t] = __ldcg(U + t);
    for (int64_t t = t_first; t < cv; t += t_step) V_out[t] = __ldcg(V + t);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    result->lr = lr;
    result->ll = old_ll;
    result->steps = steps;
    result->trials = trials;
    result->converged = converged ? 1 : 0;
  }
}

}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int64_t amf_pmf_fit_workspace_bytes(const amf_ratings_t* h, int dtype, int ld) {
  if (!h || ld <= 0) return -1;
  const size_t es = dtype == AMF_F32 ? 4 : 8;
  const size_t tables = (size_t)(h->n_users + h->n_items) * ld * es;
  return (int64_t)(256 + 3 * tables + 256 + 4 * (size_t)(h->nnz + h->tail_n) + 256);
}

int amf_pmf_fit_lls(const amf_ratings_t* hc, int dtype, int d, int ld, void* U_d, void* V_d,
                    const amf_pmf_params_t* p, double lr, double min_lr, double stop_thresh,
                    int max_steps, double* trace_d, int trace_cap, amf_fit_result_t* result_d,
                    void* workspace_d, int64_t workspace_bytes, void* stream) {
  AMF_REQUIRE(hc && U_d && V_d && p && result_d && workspace_d, "amf_pmf_fit_lls: NULL argument");
  AMF_REQUIRE(dtype == hc->dtype, "amf_pmf_fit_lls: dtype does not match the rating list");
  AMF_REQUIRE(ld >= d && d >= 1, "amf_pmf_fit_lls: ld=%d must be >= d=%d >= 1", ld, d);
  AMF_REQUIRE(trace_cap == 0 || trace_d, "amf_pmf_fit_lls: trace buffer is NULL");
  AMF_REQUIRE(workspace_bytes >= amf_pmf_fit_workspace_bytes(hc, dtype, ld),
              "amf_pmf_fit_lls: workspace too small (see amf_pmf_fit_workspace_bytes)");
  AMF_REQUIRE(((uintptr_t)workspace_d & 255) == 0, "amf_pmf_fit_lls: workspace must be 256-byte aligned");
  amf_ratings* h = const_cast<amf_ratings*>(hc);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = ratings_compact(h, s);            // the loop walks the user-major list
  if (rc != AMF_OK) return rc;
  const size_t es = dtype == AMF_F32 ? 4 : 8;
  const size_t tu = (size_t)h->n_users * ld * es, tv = (size_t)h->n_items * ld * es;
  char* w = (char*)workspace_d;
  FitState* st = (FitState*)w;
  static_assert(sizeof(FitState) <= 256, "FitState must fit its slot");
  w += 256;
  void *U2 = w, *gU = w + tu, *gU2 = w + 2 * tu;
  char* w2 = w + 3 * tu;
  void *V2 = w2, *gV = w2 + tv, *gV2 = w2 + 2 * tv;
  int32_t* own = (int32_t*)(((uintptr_t)(w2 + 3 * tv) + 255) & ~(uintptr_t)255);
  AMF_CUDA(cudaMemsetAsync(st, 0, sizeof(FitState), s));
  if (h->nnz > 0) {
    expand_rows_kernel<int32_t><<<num_sms() * 4, 256, 0, s>>>(h->ptr[0], h->n_users, own);
    AMF_LAUNCH_CHECK();
  }
  // enough CTAs to give every thread about one rating / a few table entries per trial, never
  // more than fit the device at once (cooperative launch: the grid barrier needs them resident)
  const int64_t work = std::max<int64_t>(h->nnz, (int64_t)(h->n_users + h->n_items) * ld / 4);
  int64_t grid64 = (work + FIT_THREADS - 1) / FIT_THREADS;
  if (grid64 < 1) grid64 = 1;
  if (grid64 > num_sms()) grid64 = num_sms();
  int grid = (int)grid64;
  int64_t nnz = h->nnz;
  const int32_t* idx = h->idx[0];
  const void* val = h->val[0];
  int32_t n = h->n_users, m = h->n_items;
  double sigma_sq = p->sigma_sq, sigma_u_sq = p->sigma_u_sq, sigma_v_sq = p->sigma_v_sq;
  double mean_offset = p->mean_offset;
  void* args[] = {&nnz, &own, &idx, &val, &n, &m, &d, &ld, &U_d, &V_d, &U2, &V2, &gU, &gV, &gU2,
                  &gV2, &sigma_sq, &sigma_u_sq, &sigma_v_sq, &mean_offset, &lr, &min_lr,
                  &stop_thresh, &max_steps, &trace_d, &trace_cap, &result_d, &st};
  const void* fn = dtype == AMF_F32 ? (const void*)fit_lls_kernel<float>
                                    : (const void*)fit_lls_kernel<double>;
  AMF_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(FIT_THREADS), args, 0, s));
  return AMF_OK;
}

#pragma GCC visibility pop
}  // extern "C"
