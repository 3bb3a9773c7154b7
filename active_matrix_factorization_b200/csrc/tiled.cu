// Tiled copy of the rating list and the fused PMF loss + gradient that runs on it
// (pmf_cy.pyx:170-193 log_likelihood, :204-223 gradient).
//
// The row-sorted side pass (pmf.cu) gathers one factor row of the other side per rating from
// L2 and is bound by the L2->SM path.  Here every side of the gradient gets a second copy of
// the list, bucketed by TILE of the other side's matrix and sorted by its own row inside a
// tile: 4 bytes of packed index  (own row << jbits | other row % tile_rows)  + the rating.  A
// CTA keeps the tile of the other side in shared memory (TMA bulk copies), the own row and its
// gradient accumulator live in registers, and L2 is touched once per (row, tile) visit: one
// row fetch and one vector RED.ADD of the accumulated  sum_j (e_ij / sigma^2) * Other_j.
//   side 0: users stream past item tiles  -> dU and the squared error
//   side 1: items stream past user tiles  -> dV
// Memory order inside a tile is the pool's (pool.cu): chunks of TILED_CHUNK entries, a warp per
// chunk, every group of four lanes walks TILED_RUN*4 CONSECUTIVE sorted entries while each
// batch of 32 is one coalesced 128-byte read.
#include <cub/cub.cuh>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "tile_stream.cuh"

namespace amf {

constexpr int TILED_RUN = 16;                    // batches of 32 entries per chunk
constexpr int TILED_CHUNK = 32 * TILED_RUN;      // entries per chunk (one warp, one grab)
constexpr int64_t TILED_AUTO_MIN_NNZ = 1ll << 20;

static int bits_for_u64(uint64_t x) {
  int b = 1;
  while (b < 63 && (1ull << b) < x) ++b;
  return b;
}

__global__ void tiled_keys_kernel(const int32_t* __restrict__ own, const int32_t* __restrict__ other,
                                  int64_t n, int tile_rows, int jbits, int ibits,
                                  uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t i = (uint32_t)own[t], j = (uint32_t)other[t];
    const uint64_t tile = j / (uint32_t)tile_rows, jl = j % (uint32_t)tile_rows;
    keys[t] = (((tile << ibits) | i) << jbits) | jl;
    vals[t] = (uint32_t)t;
  }
}

// keys sorted ascending; start[b] = first position whose tile >= b, start[n_tiles] = n
__global__ void tiled_start_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                   int64_t n_tiles, int64_t* __restrict__ start) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p <= n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lo = (p == 0) ? -1 : (int64_t)(keys[p - 1] >> shift);
    const int64_t hi = (p == n) ? n_tiles : (int64_t)(keys[p] >> shift);
    for (int64_t b = lo + 1; b <= hi; ++b) start[b] = p;
  }
}

__global__ void tiled_count_kernel(const int64_t* __restrict__ start, int64_t n_tiles,
                                   int64_t* __restrict__ count, int64_t* __restrict__ nchunk) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b <= n_tiles;
       b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = b < n_tiles ? start[b + 1] - start[b] : 0;
    if (b < n_tiles) count[b] = c;
    nchunk[b] = (c + TILED_CHUNK - 1) / TILED_CHUNK;
  }
}

template <typename T>
__global__ void tiled_scatter_kernel(const uint64_t* __restrict__ keys,
                                     const uint32_t* __restrict__ perm, const T* __restrict__ val,
                                     int64_t n, int ibits, int jbits,
                                     const int64_t* __restrict__ start,
                                     const int64_t* __restrict__ tile_cstart,
                                     uint32_t* __restrict__ cw, T* __restrict__ rv) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[p];
    const int64_t b = (int64_t)(k >> (ibits + jbits));
    const int64_t o = p - start[b];                       // sorted offset inside the tile
    const int64_t chunk = tile_cstart[b] + o / TILED_CHUNK;
    const int oo = (int)(o % TILED_CHUNK);
    const int g = oo / (4 * TILED_RUN), r = (oo % (4 * TILED_RUN)) >> 2, q = oo & 3;
    const int64_t pos = chunk * TILED_CHUNK + r * 32 + g * 4 + q;
    cw[pos] = (uint32_t)(k & ((1ull << (ibits + jbits)) - 1));
    rv[pos] = val[perm[p]];
  }
}

static void free_tiled_side(amf_tiled_side* t) {
  cudaFree(t->cw); cudaFree(t->rv); cudaFree(t->tile_cstart); cudaFree(t->tile_count);
  memset(t, 0, sizeof(*t));
}

// side 0 is cut from the user-major list (own = user, other = item), side 1 from the item-major
template <typename T>
static int build_tiled_side(amf_ratings* h, int side, int tile_rows, cudaStream_t s) {
  amf_tiled_side* t = &h->tiled[side];
  free_tiled_side(t);
  const int64_t nnz = h->nnz;
  const int32_t own_rows = side == 0 ? h->n_users : h->n_items;
  const int32_t other_rows = side == 0 ? h->n_items : h->n_users;
  int jbits = 0;
  while ((1 << jbits) < tile_rows) ++jbits;
  const int ibits = bits_for_u64((uint64_t)own_rows);
  if (ibits + jbits > 32) {
    set_error("tiled rating list: %d rows x tiles of %d do not fit the 4-byte packed index",
              own_rows, tile_rows);
    return AMF_ERR_UNSUPPORTED;
  }
  t->tile_rows = tile_rows; t->jbits = jbits;
  t->n_tiles = (other_rows + tile_rows - 1) / tile_rows;
  const int64_t nt = t->n_tiles;
  const int tbits = bits_for_u64((uint64_t)nt + 1);
  const int grid = num_sms() * 8;
  int rc = AMF_OK;
  int32_t* own = nullptr;
  uint64_t *keys = nullptr, *keys_out = nullptr;
  uint32_t *vals = nullptr, *perm = nullptr;
  int64_t *start = nullptr, *nchunk = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0, scan_bytes = 0;
  int64_t total_chunks = 0;
#define TILED_CUDA(call)                                                                         \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));          \
      rc = AMF_ERR_CUDA;                                                                         \
      goto done;                                                                                 \
    }                                                                                            \
  } while (0)
  TILED_CUDA(cudaMalloc(&t->tile_cstart, 8 * (size_t)(nt + 1)));
  TILED_CUDA(cudaMalloc(&t->tile_count, 8 * (size_t)(nt + 1)));
  TILED_CUDA(cudaMalloc(&start, 8 * (size_t)(nt + 1)));
  TILED_CUDA(cudaMalloc(&nchunk, 8 * (size_t)(nt + 1)));
  TILED_CUDA(cudaMalloc(&own, 4 * (size_t)nnz));
  TILED_CUDA(cudaMalloc(&keys, 8 * (size_t)nnz));
  TILED_CUDA(cudaMalloc(&keys_out, 8 * (size_t)nnz));
  TILED_CUDA(cudaMalloc(&vals, 4 * (size_t)nnz));
  TILED_CUDA(cudaMalloc(&perm, 4 * (size_t)nnz));
  expand_rows_kernel<int32_t><<<grid, 256, 0, s>>>(h->ptr[side], own_rows, own);
  TILED_CUDA(cudaGetLastError());
  tiled_keys_kernel<<<grid, 256, 0, s>>>(own, h->idx[side], nnz, tile_rows, jbits, ibits, keys, vals);
  TILED_CUDA(cudaGetLastError());
  TILED_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, vals, perm, nnz,
                                             0, ibits + jbits + tbits, s));
  TILED_CUDA(cudaMalloc(&tmp, tmp_bytes > 0 ? tmp_bytes : 1));
  TILED_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, vals, perm, nnz, 0,
                                             ibits + jbits + tbits, s));
  cudaFree(tmp); tmp = nullptr;
  tiled_start_kernel<<<grid, 256, 0, s>>>(keys_out, nnz, ibits + jbits, nt, start);
  TILED_CUDA(cudaGetLastError());
  tiled_count_kernel<<<grid, 256, 0, s>>>(start, nt, t->tile_count, nchunk);
  TILED_CUDA(cudaGetLastError());
  TILED_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, nchunk, t->tile_cstart, nt + 1, s));
  TILED_CUDA(cudaMalloc(&tmp, scan_bytes > 0 ? scan_bytes : 1));
  TILED_CUDA(cub::DeviceScan::ExclusiveSum(tmp, scan_bytes, nchunk, t->tile_cstart, nt + 1, s));
  TILED_CUDA(cudaMemcpyAsync(&total_chunks, t->tile_cstart + nt, 8, cudaMemcpyDeviceToHost, s));
  TILED_CUDA(cudaStreamSynchronize(s));
  t->n_chunks = total_chunks;
  t->npad = total_chunks * TILED_CHUNK;
  TILED_CUDA(cudaMalloc(&t->cw, 4 * (size_t)t->npad));
  TILED_CUDA(cudaMalloc(&t->rv, sizeof(T) * (size_t)t->npad));
  TILED_CUDA(cudaMemsetAsync(t->cw, 0, 4 * (size_t)t->npad, s));
  TILED_CUDA(cudaMemsetAsync(t->rv, 0, sizeof(T) * (size_t)t->npad, s));
  tiled_scatter_kernel<T><<<grid, 256, 0, s>>>(keys_out, perm, (const T*)h->val[side], nnz, ibits,
                                               jbits, start, t->tile_cstart, t->cw, (T*)t->rv);
  TILED_CUDA(cudaGetLastError());
  TILED_CUDA(cudaStreamSynchronize(s));
done:
#undef TILED_CUDA
  cudaFree(own); cudaFree(keys); cudaFree(keys_out); cudaFree(vals); cudaFree(perm);
  cudaFree(start); cudaFree(nchunk); cudaFree(tmp);
  if (rc != AMF_OK) free_tiled_side(t);
  return rc;
}

// ---- accumulator update of one 4-lane group ---------------------------------------------------
template <int C>
__device__ __forceinline__ void axpy_slices(float4 (&acc)[C], float w, const float4 (&b)[C]) {
  const float2 w2 = make_float2(w, w);
#pragma unroll
  for (int t = 0; t < C; ++t) {
    const float2 lo = fma2(w2, make_float2(b[t].x, b[t].y), make_float2(acc[t].x, acc[t].y));
    const float2 hi = fma2(w2, make_float2(b[t].z, b[t].w), make_float2(acc[t].z, acc[t].w));
    acc[t] = make_float4(lo.x, lo.y, hi.x, hi.y);
  }
}
template <int C>
__device__ __forceinline__ void axpy_slices(double2 (&acc)[C], double w, const double2 (&b)[C]) {
#pragma unroll
  for (int t = 0; t < C; ++t) vfma(acc[t], w, b[t]);
}
// fire-and-forget vector reduction into global memory
__device__ __forceinline__ void red_add(float* p, const float4& v) {
  AMF_DBG_WRITE(p, 16);
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add(double* p, const double2& v) {
  AMF_DBG_WRITE(p, 16);
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p), "d"(v.x) : "memory");
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p + 1), "d"(v.y) : "memory");
}
__device__ __forceinline__ void ld4(const float* p, float (&out)[4]) {
  const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
__device__ __forceinline__ void ld4(const double* p, double (&out)[4]) {
  const double2 v0 = __ldcs(reinterpret_cast<const double2*>(p));
  const double2 v1 = __ldcs(reinterpret_cast<const double2*>(p) + 1);
  out[0] = v0.x; out[1] = v0.y; out[2] = v1.x; out[3] = v1.y;
}

// One side of the fused loss + gradient on the tiled list.  Own = the matrix whose rows stream
// (row + gradient accumulator in registers), Tile = the matrix whose tile sits in shared
// memory.  Lane layout and the conflict-free slice rotation are those of pool_pred_kernel.
template <typename T, int NVEC, int THREADS, bool GRAD>
__global__ void __launch_bounds__(THREADS, 1)
tiled_side_kernel(const uint32_t* __restrict__ cw, const T* __restrict__ rv,
                  const int64_t* __restrict__ tile_cstart, const int64_t* __restrict__ tile_count,
                  int n_tiles, int64_t n_chunks, int jbits, int tile_rows, int tile_side_rows,
                  const T* __restrict__ Own, const T* __restrict__ Tile, T inv_sigma,
                  T mean_offset, T* __restrict__ dOwn, double* __restrict__ sq_err) {
  using V = typename Vec<T>::type;
  constexpr int CPL = NVEC >= 4 ? NVEC / 4 : 1;       // 16-byte slices per lane
  constexpr uint32_t ROW_BYTES = NVEC * 16;
  constexpr uint32_t NONE = 0xffffffffu;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_v;
  __shared__ unsigned int s_ctr;
  const uint32_t smem0 = smem_u32(smem_raw);
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, l = lane & 3;
  const uint32_t jmask = (1u << jbits) - 1;
  const bool have = l < NVEC;
  // interleaved ownership: the 256-bit row load costs this kernel registers it does not have
  // (measured +6 % time)
  constexpr bool ADJ = false;
  const uint32_t off0 = slice_off0<CPL, ADJ>(l, g & 1, have);
  const uint32_t vrow0 = smem0 + off0;
  const uint64_t own_base = (uint64_t)reinterpret_cast<uintptr_t>(Own);
  const uint64_t down0 = (uint64_t)reinterpret_cast<uintptr_t>(dOwn) + off0;

  const int64_t c_lo = n_chunks * blockIdx.x / gridDim.x;
  const int64_t c_hi = n_chunks * (blockIdx.x + 1) / gridDim.x;
  if (threadIdx.x == 0) mbar_init(&bar_v, 1);
  int t_cur = 0;
  {                                                   // last tile starting at or before c_lo
    int lo = 0, hi = n_tiles;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tile_cstart[mid] <= c_lo) lo = mid; else hi = mid - 1;
    }
    t_cur = lo;
  }
  double local_sq = 0;
  uint32_t phase_v = 0;

  for (int64_t c = c_lo; c < c_hi;) {
    while (t_cur + 1 < n_tiles && tile_cstart[t_cur + 1] <= c) ++t_cur;
    const int64_t t_first = tile_cstart[t_cur];
    const int64_t seg_end = min(c_hi, tile_cstart[t_cur + 1]);
    const int64_t t_count = tile_count[t_cur];
    __syncthreads();                                  // previous tile and counter are done with
    if (threadIdx.x == 0) {
      s_ctr = 0;
      const int rows = min(tile_rows, tile_side_rows - t_cur * tile_rows);
      const uint32_t bytes = (uint32_t)rows * ROW_BYTES;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bar_v, bytes);
      const unsigned char* src =
          reinterpret_cast<const unsigned char*>(Tile) + (int64_t)t_cur * tile_rows * ROW_BYTES;
      for (uint32_t o = 0; o < bytes; o += 32768u)
        tma_load_1d(smem_raw + o, src + o, min(bytes - o, 32768u), &bar_v);
    }
    __syncthreads();
    mbar_wait(&bar_v, phase_v);
    phase_v ^= 1;

    for (;;) {
      unsigned int grab = 0;
      if (lane == 0) grab = atomicAdd(&s_ctr, 1u);
      const int64_t chunk = c + (int64_t)__shfl_sync(0xffffffffu, grab, 0);
      if (chunk >= seg_end) break;
      const int64_t cb = chunk * TILED_CHUNK;
      // entries of this chunk that are real (the last chunk of a tile is padded)
      const int64_t left = t_count - (chunk - t_first) * TILED_CHUNK;
      const int nvalid = left >= TILED_CHUNK ? TILED_CHUNK : (int)left;
      // the bounds check of padded chunks (the last chunk of a tile) is compiled out of the
      // path every other chunk takes
      auto process = [&](auto check_tag) {
        constexpr bool CHECK = decltype(check_tag)::value;
        const uint4* wp = reinterpret_cast<const uint4*>(cw + cb) + g;
        const T* rp = rv + cb + g * 4;
        // index words and ratings two batches ahead (HBM latency)
        uint4 w1 = __ldcs(wp), w2 = __ldcs(wp + 8);
        T r1[4], r2[4];
        ld4(rp, r1);
        ld4(rp + 32, r2);
        V a[CPL], acc[CPL];
  #pragma unroll
        for (int t = 0; t < CPL; ++t) { a[t] = vzero(V()); acc[t] = vzero(V()); }
        uint32_t prev_i = NONE;
        double chunk_sq = 0;
  #pragma unroll 1
        for (int r = 0; r < TILED_RUN; ++r) {
          T sq = 0;
          const uint32_t w[4] = {w1.x, w1.y, w1.z, w1.w};
          const T rs[4] = {r1[0], r1[1], r1[2], r1[3]};
          w1 = w2;
  #pragma unroll
          for (int q = 0; q < 4; ++q) r1[q] = r2[q];
          if (r + 2 < TILED_RUN) {
            w2 = __ldcs(wp + (r + 2) * 8);
            ld4(rp + (r + 2) * 32, r2);
          }
  #pragma unroll
          for (int s = 0; s < 4; ++s) {
            // sorted offset of this entry inside the chunk: group run, batch, slot
            const bool valid = !CHECK || (g * (4 * TILED_RUN) + r * 4 + s < nvalid);
            AMF_DBG_ASSERT(!valid || (int)(w[s] & jmask) < min(tile_rows, tile_side_rows - t_cur * tile_rows));
            const uint32_t row = ((w[s] & jmask) * ROW_BYTES) + vrow0;
            V b[CPL];
  #pragma unroll
            for (int t = 0; t < CPL; ++t) b[t] = lds_v(row ^ slice_xor<CPL, ADJ>(t), V());
            const uint32_t i = w[s] >> jbits;
            if (i != prev_i) {                          // next row of this lane group's run
              if (GRAD && prev_i != NONE && have) {
                const uint64_t dp = down0 + (uint64_t)prev_i * ROW_BYTES;
  #pragma unroll
                for (int t = 0; t < CPL; ++t)
                  red_add(reinterpret_cast<T*>(dp ^ (uint64_t)slice_xor<CPL, ADJ>(t)), acc[t]);
              }
              prev_i = i;
              load_row_slices<V, CPL, ADJ>(own_base + (uint64_t)i * ROW_BYTES, have ? l : 0, g & 1, off0, a);
  #pragma unroll
              for (int t = 0; t < CPL; ++t) acc[t] = vzero(V());
            }
            T dot = have ? dot_slices<CPL>(a, b) : T(0);
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            T e = (rs[s] - mean_offset) - dot;
            if (CHECK) e = valid ? e : T(0);
            sq = fma(e, e, sq);
            if (GRAD) axpy_slices<CPL>(acc, e * inv_sigma, b);
          }
          chunk_sq += (double)sq;
        }
        if (GRAD && prev_i != NONE && have) {
          const uint64_t dp = down0 + (uint64_t)prev_i * ROW_BYTES;
  #pragma unroll
          for (int t = 0; t < CPL; ++t) red_add(reinterpret_cast<T*>(dp ^ (uint64_t)slice_xor<CPL, ADJ>(t)), acc[t]);
        }
        if (l == 0) local_sq += chunk_sq;
      };
      if (nvalid < TILED_CHUNK) process(std::true_type{}); else process(std::false_type{});
    }
    c = seg_end;
  }
  if (sq_err) {
    const double s = block_sum(local_sq);
    if (threadIdx.x == 0) atomicAdd(sq_err, s);
  }
}

// tuning knob for benchmarks/tiled_variants.py: AMF_TILED_KB = shared memory given to the tile
// (default 224 KB, the most that fits next to the static buffers); clamped to [16, 224]
static int tiled_tile_kb() {
  const char* e = getenv("AMF_TILED_KB");
  const int kb = e ? atoi(e) : 224;
  return kb < 16 ? 16 : (kb > 224 ? 224 : kb);
}

template <typename T, bool GRAD>
static int launch_tiled(const amf_ratings* h, int side, int nvec, const T* Own, const T* Tile,
                        T inv_sigma, T mean_offset, T* dOwn, double* sq_err, cudaStream_t s,
                        int max_ctas) {
  const amf_tiled_side* t = &h->tiled[side];
  const int tile_side_rows = side == 0 ? h->n_items : h->n_users;
  const size_t smem = (size_t)t->tile_rows * nvec * 16;
  int64_t grid64 = (int64_t)num_sms();
  if (max_ctas > 0 && grid64 > max_ctas) grid64 = max_ctas;   // leave SMs to a concurrent collective
  if (grid64 > t->n_chunks) grid64 = t->n_chunks > 0 ? t->n_chunks : 1;
  const int grid = (int)grid64;
  if (GRAD) AMF_DBG_RANGE(0, dOwn, (size_t)(side == 0 ? h->n_users : h->n_items) * nvec * 16, s);
  // 64 registers per thread (fp32) keep 32 warps on the SM: the pass is bound by the latency of
  // the row fetch at every (row, tile) visit, so resident warps are what hides it
  constexpr int THREADS = sizeof(T) == 4 ? 1024 : 512;
#define TILED(NVEC_)                                                                              \
  do {                                                                                            \
    AMF_CUDA(cudaFuncSetAttribute(tiled_side_kernel<T, NVEC_, THREADS, GRAD>,                     \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    tiled_side_kernel<T, NVEC_, THREADS, GRAD><<<grid, THREADS, smem, s>>>(                       \
        t->cw, (const T*)t->rv, t->tile_cstart, t->tile_count, t->n_tiles, t->n_chunks, t->jbits, \
        t->tile_rows, tile_side_rows, Own, Tile, inv_sigma, mean_offset, dOwn, sq_err);           \
  } while (0)
  switch (nvec) {
    case 4: TILED(4); break;
    case 8: TILED(8); break;
    case 16: TILED(16); break;
    default:
      set_error("tiled rating list: unsupported row of %d 16-byte vectors", nvec);
      return AMF_ERR_UNSUPPORTED;
  }
#undef TILED
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

static bool tiled_row_ok(size_t row_bytes) { return row_bytes == 64 || row_bytes == 128 || row_bytes == 256; }

// Whether the fused pass for rows of `row_bytes` should run on the tiled list; builds it on
// first use.  *use is left false when the row-sorted kernels should run instead.
int tiled_prepare(amf_ratings* h, size_t row_bytes, const void* U, const void* V, const void* dU,
                  const void* dV, bool* use, cudaStream_t s) {
  *use = false;
  if (h->tiled_mode == AMF_LAYOUT_ROWS || h->nnz == 0) return AMF_OK;
  const bool aligned = (((uintptr_t)U | (uintptr_t)V | (uintptr_t)dU | (uintptr_t)dV) & (row_bytes - 1)) == 0;
  if (!tiled_row_ok(row_bytes) || !aligned) {
    if (h->tiled_mode == AMF_LAYOUT_TILED) {
      set_error("tiled rating list needs factor rows of 64, 128 or 256 bytes aligned to their size "
                "(row is %d bytes)", (int)row_bytes);
      return AMF_ERR_UNSUPPORTED;
    }
    return AMF_OK;
  }
  if (h->tiled_mode == AMF_LAYOUT_AUTO && h->nnz < TILED_AUTO_MIN_NNZ) return AMF_OK;
  if (h->tiled_row_bytes != (int)row_bytes) {
    const int tile_rows = (int)(((size_t)tiled_tile_kb() * 1024) / row_bytes);
    int rc = AMF_OK;
    for (int side = 0; side < 2 && rc == AMF_OK; ++side) {
      const int32_t tile_side_rows = side == 0 ? h->n_items : h->n_users;
      const int tr = tile_rows < tile_side_rows ? tile_rows : tile_side_rows;
      rc = h->dtype == AMF_F32 ? build_tiled_side<float>(h, side, tr, s)
                               : build_tiled_side<double>(h, side, tr, s);
    }
    if (rc != AMF_OK) {
      free_tiled_side(&h->tiled[0]); free_tiled_side(&h->tiled[1]);
      h->tiled_row_bytes = 0;
      if (rc == AMF_ERR_UNSUPPORTED && h->tiled_mode == AMF_LAYOUT_AUTO) {
        h->tiled_mode = AMF_LAYOUT_ROWS;     // do not try again at every call
        return AMF_OK;
      }
      return rc;
    }
    h->tiled_row_bytes = (int)row_bytes;
  }
  *use = true;
  return AMF_OK;
}

void tiled_free(amf_ratings* h) {
  free_tiled_side(&h->tiled[0]); free_tiled_side(&h->tiled[1]);
  h->tiled_row_bytes = 0;
}

// sides: bit 0 = users past item tiles (dU, squared error), bit 1 = items past user tiles (dV)
template <typename T>
int tiled_loss_grad(const amf_ratings* h, int ld, const T* U, const T* V, T inv_sigma,
                    T mean_offset, T* dU, T* dV, double* sq_err, cudaStream_t s,
                    cudaEvent_t dU_done, int sides, int max_ctas) {
  const int nvec = ld / Vec<T>::N;
  int rc = AMF_OK;
  if (dU) {
    if (sides & 1) {
      rc = launch_tiled<T, true>(h, 0, nvec, U, V, inv_sigma, mean_offset, dU, sq_err, s, max_ctas);
      if (rc != AMF_OK) return rc;
      if (dU_done) AMF_CUDA(cudaEventRecord(dU_done, s));   // dU is final: the caller may read it
    }
    if (sides & 2)
      rc = launch_tiled<T, true>(h, 1, nvec, V, U, inv_sigma, mean_offset, dV, nullptr, s, max_ctas);
    return rc;
  }
  if (sides & 1)
    rc = launch_tiled<T, false>(h, 0, nvec, U, V, inv_sigma, mean_offset, nullptr, sq_err, s, max_ctas);
  return rc;
}
template int tiled_loss_grad<float>(const amf_ratings*, int, const float*, const float*, float,
                                    float, float*, float*, double*, cudaStream_t, cudaEvent_t, int,
                                    int);
template int tiled_loss_grad<double>(const amf_ratings*, int, const double*, const double*, double,
                                     double, double*, double*, double*, cudaStream_t, cudaEvent_t,
                                     int, int);

}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int amf_ratings_set_layout(amf_ratings_t* h, int mode) {
  AMF_REQUIRE(h && (mode == AMF_LAYOUT_AUTO || mode == AMF_LAYOUT_ROWS || mode == AMF_LAYOUT_TILED),
              "amf_ratings_set_layout: bad arguments");
  h->tiled_mode = mode;
  if (mode == AMF_LAYOUT_ROWS) tiled_free(h);
  return AMF_OK;
}

#pragma GCC visibility pop
}  // extern "C"
