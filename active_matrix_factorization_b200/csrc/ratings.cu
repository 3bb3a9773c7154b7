// Rating-list handle: device COO -> user-major CSR + item-major CSC (stable), plus the
// per-sub-chunk starting rows the nnz-balanced kernels use.  Replaces the (nnz,3) float64
// `ratings` array walked row by row in pmf_cy.pyx:184-186,217-221 and the adjacency dicts
// of bayes_pmf.py:241-255.
#include <cub/cub.cuh>
#include <stdarg.h>

#include "common.cuh"

namespace amf {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev < 64 && cached[dev]) return cached[dev];
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 64) cached[dev] = n;
  return n;
}

__global__ void iota_kernel(uint32_t* p, int64_t n) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x)
    p[t] = (uint32_t)t;
}

// keys sorted ascending; ptr[r] = first position whose key >= r, ptr[rows] = nnz
__global__ void row_ptr_kernel(const int32_t* __restrict__ keys, int64_t nnz, int32_t rows,
                               int64_t* __restrict__ ptr) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p <= nnz;
       p += (int64_t)gridDim.x * blockDim.x) {
    int32_t lo = (p == 0) ? -1 : keys[p - 1];
    int32_t hi = (p == nnz) ? rows : keys[p];
    for (int32_t r = lo + 1; r <= hi; ++r) ptr[r] = p;
  }
}

template <typename T>
__global__ void gather_kernel(const uint32_t* __restrict__ perm, const int32_t* __restrict__ other,
                              const T* __restrict__ r, int64_t nnz, int32_t* __restrict__ idx,
                              T* __restrict__ val) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz;
       p += (int64_t)gridDim.x * blockDim.x) {
    uint32_t q = perm[p];
    idx[p] = other[q];
    val[p] = r[q];
  }
}

// sub_row[s] = row containing entry s*AMF_SUB (largest r with ptr[r] <= pos, skipping empties)
__global__ void sub_row_kernel(const int64_t* __restrict__ ptr, int32_t rows, int64_t n_sub,
                               int32_t* __restrict__ sub_row) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n_sub;
       s += (int64_t)gridDim.x * blockDim.x) {
    int64_t pos = s * AMF_SUB;
    int32_t lo = 0, hi = rows;  // invariant: ptr[lo] <= pos < ptr[hi]
    while (hi - lo > 1) {
      int32_t mid = lo + ((hi - lo) >> 1);
      if (ptr[mid] <= pos) lo = mid; else hi = mid;
    }
    sub_row[s] = lo;
  }
}

template <typename T>
__global__ void sum_kernel(const T* __restrict__ v, int64_t n, double* __restrict__ out) {
  double acc = 0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x)
    acc += (double)v[p];
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

static int bits_for(int32_t rows) {
  int b = 1;
  while (b < 31 && (1ll << b) < rows) ++b;
  return b;
}

template <typename T>
static int build_side(amf_ratings* h, int side, const int32_t* key_d, const int32_t* other_d,
                      const T* r_d, cudaStream_t s) {
  const int64_t nnz = h->nnz;
  const int32_t rows = side == 0 ? h->n_users : h->n_items;
  AMF_CUDA(cudaMalloc(&h->ptr[side], sizeof(int64_t) * (rows + 1)));
  AMF_CUDA(cudaMalloc(&h->idx[side], sizeof(int32_t) * (nnz > 0 ? nnz : 1)));
  AMF_CUDA(cudaMalloc(&h->val[side], sizeof(T) * (nnz > 0 ? nnz : 1)));
  AMF_CUDA(cudaMalloc(&h->sub_row[side], sizeof(int32_t) * (h->n_sub > 0 ? h->n_sub : 1)));
  const int grid = num_sms() * 8;
  if (nnz == 0) {
    AMF_CUDA(cudaMemsetAsync(h->ptr[side], 0, sizeof(int64_t) * (rows + 1), s));
    return AMF_OK;
  }
  int32_t* keys_sorted = nullptr;
  uint32_t *perm_in = nullptr, *perm_out = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  struct Temps {                            // released on every exit path (AMF_CUDA returns early)
    int32_t*& a; uint32_t*& b; uint32_t*& c; void*& d;
    ~Temps() { cudaFree(a); cudaFree(b); cudaFree(c); cudaFree(d); }
  } temps{keys_sorted, perm_in, perm_out, tmp};
  AMF_CUDA(cudaMalloc(&keys_sorted, sizeof(int32_t) * nnz));
  AMF_CUDA(cudaMalloc(&perm_in, sizeof(uint32_t) * nnz));
  AMF_CUDA(cudaMalloc(&perm_out, sizeof(uint32_t) * nnz));
  iota_kernel<<<grid, 256, 0, s>>>(perm_in, nnz);
  AMF_LAUNCH_CHECK();
  const int end_bit = bits_for(rows);
  AMF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_d, keys_sorted, perm_in,
                                           perm_out, nnz, 0, end_bit, s));
  AMF_CUDA(cudaMalloc(&tmp, tmp_bytes > 0 ? tmp_bytes : 1));
  AMF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_d, keys_sorted, perm_in, perm_out,
                                           nnz, 0, end_bit, s));
  row_ptr_kernel<<<grid, 256, 0, s>>>(keys_sorted, nnz, rows, h->ptr[side]);
  AMF_LAUNCH_CHECK();
  gather_kernel<T><<<grid, 256, 0, s>>>(perm_out, other_d, r_d, nnz, h->idx[side],
                                        (T*)h->val[side]);
  AMF_LAUNCH_CHECK();
  sub_row_kernel<<<grid, 256, 0, s>>>(h->ptr[side], rows, h->n_sub, h->sub_row[side]);
  AMF_LAUNCH_CHECK();
  AMF_CUDA(cudaStreamSynchronize(s));
  return AMF_OK;
}

int launch_best_final(const Best* part_d, int nparts, bool maximize, amf_best_t* out_d,
                      cudaStream_t s);
void tiled_free(amf_ratings* h);

static void free_sides(amf_ratings* h) {
  for (int s = 0; s < 2; ++s) {
    cudaFree(h->ptr[s]); cudaFree(h->idx[s]); cudaFree(h->val[s]); cudaFree(h->sub_row[s]);
    h->ptr[s] = nullptr; h->idx[s] = nullptr; h->val[s] = nullptr; h->sub_row[s] = nullptr;
  }
}

// Sorted lists + tail -> sorted lists of everything.  The user-major list keeps the order of
// arrival inside a row and the tail is placed after it, so the stable sorts give the lists a
// fresh amf_ratings_create of the whole rating list (in order of arrival) would give.
int ratings_compact(amf_ratings* h, cudaStream_t s) {
  if (h->tail_n == 0) return AMF_OK;
  const int64_t n0 = h->nnz, n1 = h->tail_n, total = n0 + n1;
  AMF_REQUIRE(total < (1ll << 32), "rating list would exceed 2^32 entries");
  const size_t es = h->dtype == AMF_F32 ? 4 : 8;
  int32_t *i_d = nullptr, *j_d = nullptr;
  void* r_d = nullptr;
  AMF_CUDA(cudaMalloc(&i_d, 4 * (size_t)total));
  AMF_CUDA(cudaMalloc(&j_d, 4 * (size_t)total));
  AMF_CUDA(cudaMalloc(&r_d, es * (size_t)total));
  if (n0 > 0) {
    expand_rows_kernel<int32_t><<<num_sms() * 8, 256, 0, s>>>(h->ptr[0], h->n_users, i_d);
    AMF_LAUNCH_CHECK();
    AMF_CUDA(cudaMemcpyAsync(j_d, h->idx[0], 4 * (size_t)n0, cudaMemcpyDeviceToDevice, s));
    AMF_CUDA(cudaMemcpyAsync(r_d, h->val[0], es * (size_t)n0, cudaMemcpyDeviceToDevice, s));
  }
  AMF_CUDA(cudaMemcpyAsync(i_d + n0, h->tail_i, 4 * (size_t)n1, cudaMemcpyDeviceToDevice, s));
  AMF_CUDA(cudaMemcpyAsync(j_d + n0, h->tail_j, 4 * (size_t)n1, cudaMemcpyDeviceToDevice, s));
  AMF_CUDA(cudaMemcpyAsync((char*)r_d + es * n0, h->tail_r, es * (size_t)n1, cudaMemcpyDeviceToDevice, s));
  AMF_CUDA(cudaStreamSynchronize(s));
  // the merged lists are built beside the old ones and swapped in only on success: a failed
  // build (out of memory) leaves the handle as it was, tail included
  struct Sides { int64_t* ptr[2]; int32_t* idx[2]; void* val[2]; int32_t* sub_row[2]; } old_sides;
  for (int t = 0; t < 2; ++t) {
    old_sides.ptr[t] = h->ptr[t]; old_sides.idx[t] = h->idx[t];
    old_sides.val[t] = h->val[t]; old_sides.sub_row[t] = h->sub_row[t];
    h->ptr[t] = nullptr; h->idx[t] = nullptr; h->val[t] = nullptr; h->sub_row[t] = nullptr;
  }
  const int64_t old_nnz = h->nnz, old_sub = h->n_sub;
  h->nnz = total;
  h->n_sub = (total + AMF_SUB - 1) / AMF_SUB;
  int rc;
  if (h->dtype == AMF_F32) {
    rc = build_side<float>(h, 0, i_d, j_d, (const float*)r_d, s);
    if (rc == AMF_OK) rc = build_side<float>(h, 1, j_d, i_d, (const float*)r_d, s);
  } else {
    rc = build_side<double>(h, 0, i_d, j_d, (const double*)r_d, s);
    if (rc == AMF_OK) rc = build_side<double>(h, 1, j_d, i_d, (const double*)r_d, s);
  }
  cudaFree(i_d); cudaFree(j_d); cudaFree(r_d);
  if (rc != AMF_OK) {
    free_sides(h);                          // whatever the failed build allocated
    for (int t = 0; t < 2; ++t) {
      h->ptr[t] = old_sides.ptr[t]; h->idx[t] = old_sides.idx[t];
      h->val[t] = old_sides.val[t]; h->sub_row[t] = old_sides.sub_row[t];
    }
    h->nnz = old_nnz; h->n_sub = old_sub;
    return rc;
  }
  for (int t = 0; t < 2; ++t) {
    cudaFree(old_sides.ptr[t]); cudaFree(old_sides.idx[t]);
    cudaFree(old_sides.val[t]); cudaFree(old_sides.sub_row[t]);
  }
  tiled_free(h);                            // the tiled copies are rebuilt on their next use
  h->tail_n = 0;
  return AMF_OK;
}

}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

const char* amf_last_error(void) { return amf::g_err; }
int amf_version(void) { return 100; }

int amf_device_info(int* n_devices, int* sm_arch, int* n_sms) {
  int n = 0;
  AMF_CUDA(cudaGetDeviceCount(&n));
  AMF_REQUIRE(n > 0, "no CUDA device visible; this library has no CPU path");
  int dev = 0, major = 0, minor = 0, sms = 0;
  AMF_CUDA(cudaGetDevice(&dev));
  AMF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  AMF_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  AMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (n_devices) *n_devices = n;
  if (sm_arch) *sm_arch = major * 10 + minor;
  if (n_sms) *n_sms = sms;
  return AMF_OK;
}

int amf_ratings_create(amf_ratings_t** out, int32_t n_users, int32_t n_items, int64_t nnz,
                       const int32_t* i_d, const int32_t* j_d, const void* r_d, int dtype,
                       void* stream) {
  AMF_REQUIRE(out != nullptr, "amf_ratings_create: out is NULL");
  AMF_REQUIRE(n_users > 0 && n_items > 0, "amf_ratings_create: empty matrix %d x %d", n_users,
              n_items);
  AMF_REQUIRE(nnz >= 0 && nnz < (1ll << 32), "amf_ratings_create: nnz=%lld out of range",
              (long long)nnz);
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_ratings_create: bad dtype %d", dtype);
  cudaStream_t s = (cudaStream_t)stream;
  AMF_REQUIRE(nnz == 0 || (i_d && j_d && r_d), "amf_ratings_create: NULL rating arrays");
  AMF_CHECK_ID_RANGE("amf_ratings_create", i_d, n_users, j_d, n_items, nnz, s);
  amf_ratings* h = new amf_ratings();
  memset(h, 0, sizeof(*h));
  h->n_users = n_users; h->n_items = n_items; h->nnz = nnz; h->dtype = dtype;
  h->n_sub = (nnz + AMF_SUB - 1) / AMF_SUB;
  cudaGetDevice(&h->device);
  int rc;
  if (dtype == AMF_F32) {
    rc = build_side<float>(h, 0, i_d, j_d, (const float*)r_d, s);
    if (rc == AMF_OK) rc = build_side<float>(h, 1, j_d, i_d, (const float*)r_d, s);
  } else {
    rc = build_side<double>(h, 0, i_d, j_d, (const double*)r_d, s);
    if (rc == AMF_OK) rc = build_side<double>(h, 1, j_d, i_d, (const double*)r_d, s);
  }
  if (rc == AMF_OK && cudaMalloc(&h->sums_d, sizeof(double) * AMF_SUMS_DOUBLES) != cudaSuccess) rc = AMF_ERR_CUDA;
  if (rc == AMF_OK && cudaMemsetAsync(h->sums_d, 0, sizeof(double) * AMF_SUMS_DOUBLES, s) != cudaSuccess) rc = AMF_ERR_CUDA;
  if (rc != AMF_OK) { amf_ratings_destroy(h); return rc; }
  *out = h;
  return AMF_OK;
}

int amf_ratings_create_host(amf_ratings_t** out, int32_t n_users, int32_t n_items, int64_t nnz,
                            const int32_t* i_h, const int32_t* j_h, const void* r_h, int dtype) {
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_ratings_create_host: bad dtype %d", dtype);
  const size_t es = dtype == AMF_F32 ? 4 : 8;
  int32_t *i_d = nullptr, *j_d = nullptr;
  void* r_d = nullptr;
  const size_t cnt = nnz > 0 ? (size_t)nnz : 1;
  AMF_CUDA(cudaMalloc(&i_d, 4 * cnt));
  AMF_CUDA(cudaMalloc(&j_d, 4 * cnt));
  AMF_CUDA(cudaMalloc(&r_d, es * cnt));
  if (nnz > 0) {
    AMF_CUDA(cudaMemcpy(i_d, i_h, 4 * nnz, cudaMemcpyHostToDevice));
    AMF_CUDA(cudaMemcpy(j_d, j_h, 4 * nnz, cudaMemcpyHostToDevice));
    AMF_CUDA(cudaMemcpy(r_d, r_h, es * nnz, cudaMemcpyHostToDevice));
  }
  int rc = amf_ratings_create(out, n_users, n_items, nnz, i_d, j_d, r_d, dtype, nullptr);
  cudaFree(i_d); cudaFree(j_d); cudaFree(r_d);
  return rc;
}

int amf_ratings_append(amf_ratings_t* h, int64_t n_new, const int32_t* i_d, const int32_t* j_d,
                       const void* r_d, void* stream) {
  AMF_REQUIRE(h && n_new >= 0 && (n_new == 0 || (i_d && j_d && r_d)), "amf_ratings_append: bad arguments");
  if (n_new == 0) return AMF_OK;
  AMF_REQUIRE(h->nnz + h->tail_n + n_new < (1ll << 32), "amf_ratings_append: list would exceed 2^32 entries");
  cudaStream_t s = (cudaStream_t)stream;
  AMF_CHECK_ID_RANGE("amf_ratings_append", i_d, h->n_users, j_d, h->n_items, n_new, s);
  const size_t es = h->dtype == AMF_F32 ? 4 : 8;
  if (h->tail_n + n_new > h->tail_cap) {
    int64_t cap = h->tail_cap > 0 ? h->tail_cap : 4096;
    while (cap < h->tail_n + n_new) cap *= 2;
    int32_t *ni = nullptr, *nj = nullptr;
    void* nr = nullptr;
    AMF_CUDA(cudaMalloc(&ni, 4 * (size_t)cap));
    AMF_CUDA(cudaMalloc(&nj, 4 * (size_t)cap));
    AMF_CUDA(cudaMalloc(&nr, es * (size_t)cap));
    if (h->tail_n > 0) {
      AMF_CUDA(cudaMemcpyAsync(ni, h->tail_i, 4 * (size_t)h->tail_n, cudaMemcpyDeviceToDevice, s));
      AMF_CUDA(cudaMemcpyAsync(nj, h->tail_j, 4 * (size_t)h->tail_n, cudaMemcpyDeviceToDevice, s));
      AMF_CUDA(cudaMemcpyAsync(nr, h->tail_r, es * (size_t)h->tail_n, cudaMemcpyDeviceToDevice, s));
      AMF_CUDA(cudaStreamSynchronize(s));
    }
    cudaFree(h->tail_i); cudaFree(h->tail_j); cudaFree(h->tail_r);
    h->tail_i = ni; h->tail_j = nj; h->tail_r = nr; h->tail_cap = cap;
  }
  AMF_CUDA(cudaMemcpyAsync(h->tail_i + h->tail_n, i_d, 4 * (size_t)n_new, cudaMemcpyDeviceToDevice, s));
  AMF_CUDA(cudaMemcpyAsync(h->tail_j + h->tail_n, j_d, 4 * (size_t)n_new, cudaMemcpyDeviceToDevice, s));
  AMF_CUDA(cudaMemcpyAsync((char*)h->tail_r + es * h->tail_n, r_d, es * (size_t)n_new,
                           cudaMemcpyDeviceToDevice, s));
  h->tail_n += n_new;
  // keep the unsorted tail a small fraction of the list: it is walked with atomics into both sides
  const int64_t limit = h->nnz / 32 > 65536 ? h->nnz / 32 : 65536;
  if (h->tail_n > limit) return amf::ratings_compact(h, s);
  return AMF_OK;
}

int amf_ratings_compact(amf_ratings_t* h, void* stream) {
  AMF_REQUIRE(h, "amf_ratings_compact: NULL handle");
  return amf::ratings_compact(h, (cudaStream_t)stream);
}

int amf_ratings_destroy(amf_ratings_t* h) {
  if (!h) return AMF_OK;
  for (int s = 0; s < 2; ++s) {
    cudaFree(h->ptr[s]); cudaFree(h->idx[s]); cudaFree(h->val[s]); cudaFree(h->sub_row[s]);
  }
  cudaFree(h->tail_i); cudaFree(h->tail_j); cudaFree(h->tail_r);
  for (int k = 0; k < 8; ++k) cudaFree(h->stage[k]);
  cudaFree(h->sums_d);
  amf::tiled_free(h);
  delete h;
  return AMF_OK;
}

int64_t amf_ratings_nnz(const amf_ratings_t* h) { return h ? h->nnz + h->tail_n : -1; }

int amf_ratings_layout(const amf_ratings_t* h, int side, const int64_t** ptr_d,
                       const int32_t** idx_d, const void** val_d) {
  AMF_REQUIRE(h && (side == 0 || side == 1), "amf_ratings_layout: bad arguments");
  int rc = amf::ratings_compact(const_cast<amf_ratings*>(h), nullptr);
  if (rc != AMF_OK) return rc;
  if (ptr_d) *ptr_d = h->ptr[side];
  if (idx_d) *idx_d = h->idx[side];
  if (val_d) *val_d = h->val[side];
  return AMF_OK;
}

int amf_ratings_mean(const amf_ratings_t* h, double* mean_out, void* stream) {
  AMF_REQUIRE(h && mean_out, "amf_ratings_mean: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  {
    int rc = amf::ratings_compact(const_cast<amf_ratings*>(h), s);
    if (rc != AMF_OK) return rc;
  }
  AMF_CUDA(cudaMemsetAsync(h->sums_d + 4, 0, sizeof(double), s));
  if (h->nnz > 0) {
    if (h->dtype == AMF_F32)
      sum_kernel<float><<<num_sms() * 4, 256, 0, s>>>((const float*)h->val[0], h->nnz, h->sums_d + 4);
    else
      sum_kernel<double><<<num_sms() * 4, 256, 0, s>>>((const double*)h->val[0], h->nnz, h->sums_d + 4);
    AMF_LAUNCH_CHECK();
  }
  double total = 0;
  AMF_CUDA(cudaMemcpyAsync(&total, h->sums_d + 4, sizeof(double), cudaMemcpyDeviceToHost, s));
  AMF_CUDA(cudaStreamSynchronize(s));
  *mean_out = h->nnz > 0 ? total / (double)h->nnz : 0.0;
  return AMF_OK;
}

#pragma GCC visibility pop
}  // extern "C"
