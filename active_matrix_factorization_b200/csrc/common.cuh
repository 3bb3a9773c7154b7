// Shared internals of libamf_b200 (sm_100a).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "../../include/amf_b200.h"

namespace amf {

void set_error(const char* fmt, ...);

#define AMF_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t err__ = (call);                                                          \
    if (err__ != cudaSuccess) {                                                          \
      amf::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                       \
                     cudaGetErrorString(err__));                                         \
      return AMF_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define AMF_REQUIRE(cond, ...)                                                           \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      amf::set_error(__VA_ARGS__);                                                       \
      return AMF_ERR_INVALID;                                                            \
    }                                                                                    \
  } while (0)

#define AMF_LAUNCH_CHECK() AMF_CUDA(cudaGetLastError())

int num_sms();

// ---- debug build (-DAMF_BOUNDS_CHECK, `AMF_B200_DEBUG=1 python -m ...build --force`) --------------
// compute-sanitizer is closed on the GPU pool this library is developed on, so the kernels that
// write through computed addresses (vector RED / atomics into the gradient tables, score stores)
// carry their own checks: the host entry registers the address ranges a launch may write
// (AMF_DBG_RANGE), every such write asserts that it falls inside one of them (AMF_DBG_WRITE), and
// index decodes assert their bounds (AMF_DBG_ASSERT).  A violation prints and traps, which the
// caller sees as a CUDA error.  All of it compiles to nothing in the release build.
#ifdef AMF_BOUNDS_CHECK
struct DbgRanges { unsigned long long lo[4], hi[4]; };
static __device__ DbgRanges amf_dbg_ranges;            // one copy per translation unit
static inline void dbg_set_range(int slot, const void* p, size_t bytes, cudaStream_t s) {
  const unsigned long long lo = (unsigned long long)(uintptr_t)p, hi = lo + bytes;
  cudaMemcpyToSymbolAsync(amf_dbg_ranges, &lo, 8, offsetof(DbgRanges, lo) + 8 * slot,
                          cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(amf_dbg_ranges, &hi, 8, offsetof(DbgRanges, hi) + 8 * slot,
                          cudaMemcpyHostToDevice, s);
}
#define AMF_DBG_RANGE(slot, p, bytes, s) amf::dbg_set_range((slot), (p), (bytes), (s))
#define AMF_DBG_ASSERT(cond)                                                                  \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      printf("amf bounds check failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__,     \
             __LINE__, (int)blockIdx.x, (int)threadIdx.x);                                    \
      __trap();                                                                               \
    }                                                                                         \
  } while (0)
#define AMF_DBG_WRITE(p, bytes)                                                               \
  do {                                                                                        \
    const unsigned long long a__ = (unsigned long long)(uintptr_t)(p);                        \
    bool in__ = false;                                                                        \
    for (int q__ = 0; q__ < 4; ++q__)                                                         \
      in__ |= a__ >= amf::amf_dbg_ranges.lo[q__] && a__ + (bytes) <= amf::amf_dbg_ranges.hi[q__]; \
    AMF_DBG_ASSERT(in__ && "write outside the registered output ranges");                      \
  } while (0)
#else
#define AMF_DBG_RANGE(slot, p, bytes, s) ((void)0)
#define AMF_DBG_ASSERT(cond) ((void)0)
#define AMF_DBG_WRITE(p, bytes) ((void)0)
#endif

// ---- 16-byte vectors of the compute type ---------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float> {
  using type = float4;
  static constexpr int N = 4;
};
template <> struct Vec<double> {
  using type = double2;
  static constexpr int N = 2;
};

__device__ __forceinline__ float4 vzero(float4) { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ double2 vzero(double2) { return make_double2(0., 0.); }
__device__ __forceinline__ float vdot(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ double vdot(const double2& a, const double2& b) {
  return fma(a.x, b.x, a.y * b.y);
}
__device__ __forceinline__ void vfma(float4& acc, float s, const float4& b) {
  acc.x = fmaf(s, b.x, acc.x); acc.y = fmaf(s, b.y, acc.y);
  acc.z = fmaf(s, b.z, acc.z); acc.w = fmaf(s, b.w, acc.w);
}
__device__ __forceinline__ void vfma(double2& acc, double s, const double2& b) {
  acc.x = fma(s, b.x, acc.x); acc.y = fma(s, b.y, acc.y);
}
// vector reduction into global memory (no return value): RED.E.ADD.F32x4 on sm_100a
__device__ __forceinline__ void vred_add(float* p, const float4& v) {
  AMF_DBG_WRITE(p, 16);
  atomicAdd(reinterpret_cast<float4*>(p), v);
}
__device__ __forceinline__ void vred_add(double* p, const double2& v) {
  AMF_DBG_WRITE(p, 16);
  atomicAdd(p, v.x);
  atomicAdd(p + 1, v.y);
}
// streaming (read-once) loads that do not pollute L1
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of a double, result valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double smem_part[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem_part[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem_part[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

// ---- arg-best of (value, index): better value wins, lower index on ties, NaN never wins ------
struct Best {
  double v;
  long long i;
};
template <bool MAX>
__device__ __forceinline__ bool better(double v, long long i, double bv, long long bi) {
  if (i < 0) return false;
  if (v != v) return false;
  if (bi < 0) return true;
  if (MAX ? (v > bv) : (v < bv)) return true;
  return v == bv && i < bi;
}
template <bool MAX>
__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, b.v, o);
    long long oi = __shfl_xor_sync(0xffffffffu, b.i, o);
    if (better<MAX>(ov, oi, b.v, b.i)) { b.v = ov; b.i = oi; }
  }
  return b;
}
// result valid in thread 0
template <bool MAX>
__device__ __forceinline__ Best block_best(Best b) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  b = warp_best<MAX>(b);
  if (lane == 0) { sv[w] = b.v; si[w] = b.i; }
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  if (threadIdx.x < nw) { b.v = sv[threadIdx.x]; b.i = si[threadIdx.x]; }
  else { b.v = 0; b.i = -1; }
  if (w == 0) b = warp_best<MAX>(b);
  __syncthreads();
  return b;
}

// row[p] = row of entry p of a list given by row offsets ptr[0..rows] (warp per row).  A template
// only so that every translation unit that launches it gets its own copy without -rdc.
template <typename Index>
__global__ void expand_rows_kernel(const int64_t* __restrict__ ptr, int32_t rows,
                                   Index* __restrict__ row) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps)
    for (int64_t p = ptr[r] + lane; p < ptr[r + 1]; p += 32) row[p] = (Index)r;
}

// Counts the (a, b) index pairs outside [0, a_rows) x [0, b_rows).  A template for the same reason.
template <typename Index>
__global__ void count_out_of_range_kernel(const Index* __restrict__ a, Index a_rows,
                                          const Index* __restrict__ b, Index b_rows, int64_t n,
                                          unsigned long long* __restrict__ bad) {
  unsigned int mine = 0;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x)
    mine += ((uint32_t)a[t] >= (uint32_t)a_rows) | ((uint32_t)b[t] >= (uint32_t)b_rows);
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(bad, (unsigned long long)mine);
}

// Every structure built from caller-supplied ids (rating lists, candidate pools) is checked once
// at build time -- the reference asserts the same (pmf_cy.pyx:139-140) -- so that a bad id is an
// error code, not an out-of-bounds gather inside a kernel.  Synchronises `s`.
#define AMF_CHECK_ID_RANGE(what, a_d, a_rows, b_d, b_rows, n, s)                                  \
  do {                                                                                            \
    if ((n) > 0) {                                                                                \
      unsigned long long* bad_d__ = nullptr;                                                      \
      unsigned long long bad__ = 0;                                                               \
      AMF_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&bad_d__), 8, (s)));                      \
      AMF_CUDA(cudaMemsetAsync(bad_d__, 0, 8, (s)));                                              \
      const int64_t blocks__ = ((n) + 255) / 256;                                                 \
      const int grid__ = (int)(blocks__ < (int64_t)amf::num_sms() * 8 ? blocks__                  \
                                                                       : (int64_t)amf::num_sms() * 8); \
      amf::count_out_of_range_kernel<int32_t><<<grid__, 256, 0, (s)>>>((a_d), (a_rows), (b_d),    \
                                                                       (b_rows), (n), bad_d__);   \
      AMF_LAUNCH_CHECK();                                                                         \
      AMF_CUDA(cudaMemcpyAsync(&bad__, bad_d__, 8, cudaMemcpyDeviceToHost, (s)));                 \
      AMF_CUDA(cudaFreeAsync(bad_d__, (s)));                                                      \
      AMF_CUDA(cudaStreamSynchronize((s)));                                                       \
      AMF_REQUIRE(bad__ == 0, "%s: %llu of %lld (user, item) ids lie outside %d x %d", (what),    \
                  bad__, (long long)(n), (int)(a_rows), (int)(b_rows));                           \
    }                                                                                             \
  } while (0)

// final reduction of per-block partial winners (launch with one block)
int launch_best_final(const Best* part_d, int nparts, bool maximize, amf_best_t* out_d,
                      cudaStream_t s);

}  // namespace amf

// A sparse list in the "bundled runs" layout (runs.cuh / runs.cu): entries bucketed by tile of the
// "tile side" matrix, one lane per run segment of the "own side" row, 32 segments per bundle.
struct amf_runs {
  int tile_rows, n_tiles;
  int64_t n, n_bundles, n_groups, npos;   // npos = n_groups * 256 entry slots (padding included)
  uint16_t* idx;          // [npos] local row inside the tile (padding: tile_rows)
  void* val;              // [npos] values (NULL for a candidate pool; padding: 0)
  uint32_t* orig;         // [npos] position in the caller's list (RUNS_NONE = padding / removed) or NULL
  uint32_t* pos_of;       // [n] slot of the caller's entry c, or NULL
  uint32_t* rowid;        // [n_bundles * 32] own row of every lane (RUNS_NONE = empty lane)
  uint8_t* seglen;        // [n_bundles * 32] entries of every lane's segment
  int2* binfo;            // [n_bundles + 1] {first group, longest segment}; the last one = {n_groups, 0}
  int64_t* tile_bstart;   // [n_tiles + 1] first bundle of every tile
  uint32_t* tile_ctr;     // [n_tiles] bundles of the tile handed out in the running launch (zeroed per launch)
};

struct amf_ratings {
  int32_t n_users, n_items;
  int64_t nnz;
  int dtype;
  // side 0: user-major (rows = users, idx = item); side 1: item-major
  int64_t* ptr[2];      // [rows+1]
  int32_t* idx[2];      // [nnz]
  void* val[2];         // [nnz] of dtype
  int32_t* sub_row[2];  // row containing entry s*AMF_SUB, one per sub-chunk of AMF_SUB entries
  int64_t n_sub;
  // staging for *_host entry points (grown on demand)
  void* stage[8];
  size_t stage_bytes[8];
  // [0..3) objective sums, [4] rating sum, [6] sticky Gibbs failure flag (int), [8 .. 8 + 1056)
  // moment workspace of amf_gibbs_hyper_device (d <= 32): calls on one handle are stream-ordered
  double* sums_d;
  int device;
  // side 0: users stream past item tiles (dU); side 1: items stream past user tiles (dV)
  amf_runs tiled[2];
  int tiled_row_bytes;    // padded factor-row size the tiles were cut for (0 = not built)
  int tiled_mode;         // AMF_LAYOUT_AUTO / _ROWS / _TILED
  // ratings appended since the sorted lists were built (amf_ratings_append): plain COO, folded
  // into the sorted lists by ratings_compact() once the tail is large or a consumer needs them
  int32_t *tail_i, *tail_j;
  void* tail_r;
  int64_t tail_n, tail_cap;
};

namespace amf {
// merges the appended tail into the sorted lists (no-op if there is none)
int ratings_compact(amf_ratings* h, cudaStream_t s);
// runs.cu: builds the bundled-runs layout of n entries (own[t], other[t] [, val[t]]) on the device;
// val_size = 0 / 4 / 8 bytes per value; want_orig keeps the permutation back to the caller's order
int runs_build(amf_runs* out, int64_t n, const int32_t* own_d, const int32_t* other_d,
               const void* val_d, int val_size, int32_t own_rows, int32_t other_rows, int tile_rows,
               bool want_orig, cudaStream_t s);
void runs_free(amf_runs* r);
// gibbs_hyper.cu: one Normal-Wishart draw (mu, alpha) from the factor rows `feats` (amf_gibbs_hyper_device)
int gibbs_hyper_draw(const amf_ratings* h, int dtype, int d, int64_t rows, const void* feats,
                     const double* prior, unsigned long long seed, unsigned long long stream_id,
                     void* mu, void* alpha, cudaStream_t s);
}  // namespace amf

#define AMF_SUB 32
#define AMF_SUMS_DOUBLES (8 + 32 * 32 + 32)
