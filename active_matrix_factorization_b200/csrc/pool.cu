// Candidate pool handle: the reference's `pool` / `unrated` set (active_pmf.py:725-737) kept on
// the device in a layout built for the scoring kernel.
//
// The plain scoring kernel (scoring.cu) gathers one factor row per candidate from L2 and sits
// at the L2->SM gather ceiling.  Here the pool is bucketed once by item tile: candidates are
// sorted by (j / TJ, i, j), so a CTA keeps a TJ-row tile of V in shared memory (TMA bulk
// copies) and reads every item row from there; the user row lives in registers and is fetched
// from L2 only when the user changes (about once per TJ * density candidates).  Each candidate
// costs 4 bytes of HBM: one packed word  i << ceil(log2(TJ)) | (j % TJ).
//
// Memory order inside a tile.  The sorted list of a tile is cut into chunks of POOL_CHUNK
// candidates; a warp scores one chunk at a time in POOL_RUN batches of 32.  A group of four
// lanes owns four candidates per batch, and the chunk is stored so that the 4*POOL_RUN
// candidates a group meets over the whole chunk are CONSECUTIVE in sorted order (long runs of
// the same user -> few user-row fetches) while every batch is still one coalesced 128-byte
// read:   position = chunk*POOL_CHUNK + batch*32 + group*4 + q
//         sorted offset inside the chunk = group*(4*POOL_RUN) + batch*4 + q.
// The permutation back to the caller's order is kept, so scores and the winner are reported
// exactly as the unbucketed path would.
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"
#include "tile_stream.cuh"

namespace amf {
constexpr int POOL_RUN = 16;                    // batches of 32 candidates per chunk
constexpr int POOL_CHUNK = 32 * POOL_RUN;       // candidates per chunk (one warp, one grab)
static_assert(POOL_RUN >= 2, "the index words are prefetched two batches ahead");
constexpr int POOL_THREADS = 1024;              // one CTA per SM
constexpr uint32_t POOL_TOMBSTONE = 0xffffffffu;   // orig[] value of padding / removed candidates
}  // namespace amf

struct amf_pool {
  int64_t ncand;         // candidates given by the caller
  int64_t npad;          // entries in the bucketed arrays (every tile padded to whole chunks)
  int32_t n_users, n_items;
  int tile_rows;         // TJ: items per tile (V tile resident in shared memory)
  int jbits;             // bits of the local item index: ceil(log2(tile_rows))
  int n_tiles;
  int64_t n_chunks;      // npad / POOL_CHUNK
  uint32_t* cw;          // [npad] packed indices  i << jbits | j % TJ   (padding: 0)
  uint32_t* orig;        // [npad] position in the caller's pool (POOL_TOMBSTONE = padding/removed)
  uint32_t* pos_of;      // [ncand] bucketed position of the caller's candidate c
  int64_t* tile_cstart;  // [n_tiles+1] first chunk of every tile
  void* tmp_scores;      // scratch for scores in bucketed order
  size_t tmp_bytes;
};

namespace amf {

int acquire_partials(Best** out, cudaStream_t s);

__global__ void pool_keys_kernel(const int32_t* __restrict__ ci, const int32_t* __restrict__ cj,
                                 int64_t n, int tile_rows, int jbits, int ibits,
                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t i = (uint32_t)ci[t], j = (uint32_t)cj[t];
    const uint64_t tile = j / (uint32_t)tile_rows, jl = j % (uint32_t)tile_rows;
    keys[t] = (((tile << ibits) | i) << jbits) | jl;
    vals[t] = (uint32_t)t;
  }
}

// keys sorted ascending; start[b] = first position whose tile >= b, start[n_tiles] = n
__global__ void pool_tile_start_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                       int64_t n_tiles, int64_t* __restrict__ start) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p <= n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lo = (p == 0) ? -1 : (int64_t)(keys[p - 1] >> shift);
    const int64_t hi = (p == n) ? n_tiles : (int64_t)(keys[p] >> shift);
    for (int64_t b = lo + 1; b <= hi; ++b) start[b] = p;
  }
}

__global__ void pool_nchunk_kernel(const int64_t* __restrict__ start, int64_t n_tiles,
                                   int64_t* __restrict__ nchunk) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b <= n_tiles;
       b += (int64_t)gridDim.x * blockDim.x)
    nchunk[b] = b < n_tiles ? (start[b + 1] - start[b] + POOL_CHUNK - 1) / POOL_CHUNK : 0;
}

__global__ void pool_scatter_kernel(const uint64_t* __restrict__ keys,
                                    const uint32_t* __restrict__ perm, int64_t n, int ibits,
                                    int jbits, const int64_t* __restrict__ start,
                                    const int64_t* __restrict__ tile_cstart,
                                    uint32_t* __restrict__ cw, uint32_t* __restrict__ orig,
                                    uint32_t* __restrict__ pos_of) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[p];
    const int64_t b = (int64_t)(k >> (ibits + jbits));
    const int64_t o = p - start[b];                       // sorted offset inside the tile
    const int64_t chunk = tile_cstart[b] + o / POOL_CHUNK;
    const int oo = (int)(o % POOL_CHUNK);
    const int g = oo / (4 * POOL_RUN), r = (oo % (4 * POOL_RUN)) >> 2, q = oo & 3;
    const int64_t pos = chunk * POOL_CHUNK + r * 32 + g * 4 + q;
    cw[pos] = (uint32_t)(k & ((1ull << (ibits + jbits)) - 1));
    orig[pos] = perm[p];
    pos_of[perm[p]] = (uint32_t)pos;
  }
}

static int bits_for_count(uint64_t x) {
  int b = 1;
  while (b < 63 && (1ull << b) < x) ++b;
  return b;
}

// PRED over a bucketed pool.  One CTA per SM owns a contiguous range of chunks; for every item
// tile that range touches, the tile is brought into shared memory by TMA and the warps grab
// chunks of that tile from a shared counter.  Four lanes score one candidate: lane l holds the
// 16-byte slices l, l+4, ... of the factor rows (CPL per lane).  The two lane groups that share
// a shared-memory wavefront start at different slices ((t + group parity) mod CPL), so the
// eight lanes of a wavefront always hit eight different 16-byte bank groups whatever rows they
// read: one wavefront per 128 bytes, no conflicts.  A transpose-reduce over the four lanes
// leaves one finished dot product per lane (position base + lane): coalesced score stores and
// one compare per lane for the fused arg-best.
template <typename T, int NVEC, bool MAX>
__global__ void __launch_bounds__(POOL_THREADS, 1)
pool_pred_kernel(const uint32_t* __restrict__ cw, const uint32_t* __restrict__ orig,
                 const int64_t* __restrict__ tile_cstart, int n_tiles, int64_t n_chunks,
                 int jbits, int tile_rows, int n_items, const T* __restrict__ U,
                 const T* __restrict__ Vm, T* __restrict__ scores, int64_t index_base, Best* __restrict__ part) {
  using V = typename Vec<T>::type;
  constexpr int CPL = NVEC >= 4 ? NVEC / 4 : 1;       // 16-byte slices per lane
  constexpr uint32_t ROW_BYTES = NVEC * 16;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_v;
  __shared__ unsigned int s_ctr;
  const uint32_t smem0 = smem_u32(smem_raw);
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, l = lane & 3;
  const uint32_t jmask = (1u << jbits) - 1;
  const bool have = l < NVEC;                         // rows narrower than four slices
  constexpr bool ADJ = CPL == 2;                      // measured: -3 % time at d=32 fp32
  const uint32_t off0 = slice_off0<CPL, ADJ>(l, g & 1, have);
  const uint32_t vrow0 = smem0 + off0;                // tile base is 1024-byte aligned
  const uint64_t urow = (uint64_t)reinterpret_cast<uintptr_t>(U);

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
  T best_v = MAX ? -INFINITY : INFINITY, thr = best_v;
  int64_t best_o = -1;
  V a[CPL];
#pragma unroll
  for (int t = 0; t < CPL; ++t) a[t] = vzero(V());
  uint32_t prev_i = 0xffffffffu;
  uint32_t phase_v = 0;

  for (int64_t c = c_lo; c < c_hi;) {
    while (t_cur + 1 < n_tiles && tile_cstart[t_cur + 1] <= c) ++t_cur;
    const int64_t seg_end = min(c_hi, tile_cstart[t_cur + 1]);
    __syncthreads();                                  // previous tile and counter are done with
    if (threadIdx.x == 0) {
      s_ctr = 0;
      const int rows = min(tile_rows, n_items - t_cur * tile_rows);
      const uint32_t bytes = (uint32_t)rows * ROW_BYTES;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bar_v, bytes);
      const unsigned char* src =
          reinterpret_cast<const unsigned char*>(Vm) + (int64_t)t_cur * tile_rows * ROW_BYTES;
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
      const int64_t cb = chunk * POOL_CHUNK;
      const uint4* wp = reinterpret_cast<const uint4*>(cw + cb) + g;
      // index words two batches ahead (HBM latency)
      uint4 w1 = __ldcs(wp), w2 = __ldcs(wp + 8);
#pragma unroll 1
      for (int r = 0; r < POOL_RUN; ++r) {
        const uint4 w4 = w1;
        w1 = w2;
        if (r + 2 < POOL_RUN) w2 = __ldcs(wp + (r + 2) * 8);
        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
        T p[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          // slice t of a row sits at  off0 ^ slice_xor(t)  (rows are ROW_BYTES-aligned)
          AMF_DBG_ASSERT((int)(w[s] & jmask) < tile_rows);       // inside the shared-memory tile
          const uint32_t row = ((w[s] & jmask) * ROW_BYTES) + vrow0;
          V b[CPL];
#pragma unroll
          for (int t = 0; t < CPL; ++t) b[t] = lds_v(row ^ slice_xor<CPL, ADJ>(t), V());
          const uint32_t i = w[s] >> jbits;
          if (i != prev_i) {                          // next user of this lane group's run
            prev_i = i;
            load_row_slices<V, CPL, ADJ>(urow + (uint64_t)i * ROW_BYTES, have ? l : 0, g & 1, off0, a);
          }
          const T acc = dot_slices<CPL>(a, b);
          p[s] = have ? acc : T(0);
        }
        // transpose-reduce over the four lanes: lane l ends with candidate l of the group
        {
          const bool up2 = (l & 2) != 0;
          const T s0 = up2 ? p[0] : p[2], s1 = up2 ? p[1] : p[3];
          const T k0 = up2 ? p[2] : p[0], k1 = up2 ? p[3] : p[1];
          p[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 2);
          p[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);
          const bool up1 = (l & 1) != 0;
          const T s = up1 ? p[0] : p[1], k = up1 ? p[1] : p[0];
          p[0] = k + __shfl_xor_sync(0xffffffffu, s, 1);
        }
        const int64_t pos = cb + r * 32 + lane;
        if (scores) __stcs(scores + pos, p[0]);
        // arg-best: `thr` is the best value any lane of this warp holds; only scores that reach
        // it (ties included: a lower index may still win) look up their original position
        const bool reach = MAX ? (p[0] >= thr) : (p[0] <= thr);
        if (__any_sync(0xffffffffu, reach)) {
          if (reach) {
            const uint32_t ou = orig[pos];
            const int64_t o = (int64_t)ou;
            if (ou != POOL_TOMBSTONE &&
                (best_o < 0 || (MAX ? (p[0] > best_v) : (p[0] < best_v)) ||
                 (p[0] == best_v && o < best_o))) {
              best_v = p[0]; best_o = o;
            }
          }
          T v = best_v;                               // +-inf while the lane holds nothing
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const T other = __shfl_xor_sync(0xffffffffu, v, o);
            v = MAX ? (other > v ? other : v) : (other < v ? other : v);
          }
          thr = v;
        }
      }
    }
    c = seg_end;
  }
  Best best{(double)best_v, best_o < 0 ? -1 : best_o + index_base};
  best = block_best<MAX>(best);
  if (threadIdx.x == 0) part[blockIdx.x] = best;
}

template <typename T>
__global__ void unpermute_kernel(const T* __restrict__ in, const uint32_t* __restrict__ orig,
                                 int64_t npad, T* __restrict__ out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < npad;
       t += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t o = orig[t];
    if (o != POOL_TOMBSTONE) out[o] = in[t];
  }
}

template <typename T, bool MAX>
static int pool_pred(const amf_pool* h, int ld, const T* U, const T* V, T* scores_tmp,
                     int64_t index_base, Best* part, int grid, size_t smem, cudaStream_t s) {
  constexpr int N = Vec<T>::N;
  const int nvec = ld / N;
#define POOL(NVEC_)                                                                             \
  do {                                                                                          \
    AMF_CUDA(cudaFuncSetAttribute(pool_pred_kernel<T, NVEC_, MAX>,                              \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    pool_pred_kernel<T, NVEC_, MAX><<<grid, POOL_THREADS, smem, s>>>(                           \
        h->cw, h->orig, h->tile_cstart, h->n_tiles, h->n_chunks, h->jbits, h->tile_rows,        \
        h->n_items, U, V,                                                                       \
        scores_tmp, index_base, part);                                                          \
  } while (0)
  switch (nvec) {            // the row width is a compile-time constant of the kernel
    case 1: POOL(1); break;
    case 2: POOL(2); break;
    case 4: POOL(4); break;
    case 8: POOL(8); break;
    case 16: POOL(16); break;
    case 32: POOL(32); break;
    default:
      set_error("bucketed pool needs a padded row of 1, 2, 4, 8, 16 or 32 16-byte vectors (ld=%d)", ld);
      return AMF_ERR_UNSUPPORTED;
  }
#undef POOL
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int amf_pool_destroy(amf_pool_t* h) {
  if (!h) return AMF_OK;
  cudaFree(h->cw); cudaFree(h->orig); cudaFree(h->pos_of); cudaFree(h->tile_cstart);
  cudaFree(h->tmp_scores);
  delete h;
  return AMF_OK;
}

int amf_pool_create(amf_pool_t** out, int64_t ncand, const int32_t* ci_d, const int32_t* cj_d,
                    int32_t n_users, int32_t n_items, int tile_rows, void* stream) {
  AMF_REQUIRE(out && n_users > 0 && n_items > 0, "amf_pool_create: bad arguments");
  AMF_REQUIRE(ncand >= 0 && ncand < (1ll << 32) - 2 * POOL_CHUNK, "amf_pool_create: ncand out of range");
  AMF_REQUIRE(tile_rows >= 1 && tile_rows <= 32768, "amf_pool_create: tile_rows must be in [1, 32768]");
  int jbits = 0;
  while ((1 << jbits) < tile_rows) ++jbits;
  const int ibits = bits_for_count((uint64_t)n_users);
  AMF_REQUIRE(ibits + jbits <= 32, "amf_pool_create: %d users x tiles of %d items do not fit the "
              "4-byte packed index (use amf_score_candidates)", n_users, tile_rows);
  cudaStream_t s = (cudaStream_t)stream;
  AMF_REQUIRE(ncand == 0 || (ci_d && cj_d), "amf_pool_create: NULL candidate arrays");
  AMF_CHECK_ID_RANGE("amf_pool_create", ci_d, n_users, cj_d, n_items, ncand, s);
  amf_pool* h = new amf_pool();
  memset(h, 0, sizeof(*h));
  h->ncand = ncand; h->n_users = n_users; h->n_items = n_items;
  h->tile_rows = tile_rows; h->jbits = jbits;
  h->n_tiles = (n_items + tile_rows - 1) / tile_rows;
  int rc = AMF_OK;
  uint64_t *keys = nullptr, *keys_out = nullptr;
  uint32_t *vals = nullptr, *perm = nullptr;
  int64_t *start = nullptr, *nchunk = nullptr;
  void* tmp = nullptr;
  const size_t cnt = ncand > 0 ? (size_t)ncand : 1;
  const int grid = num_sms() * 8;
  const int64_t nt = h->n_tiles;
#define POOL_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));          \
      rc = AMF_ERR_CUDA;                                                                         \
      goto done;                                                                                 \
    }                                                                                            \
  } while (0)
  {
    const int tbits = bits_for_count((uint64_t)nt + 1);
    size_t tmp_bytes = 0, scan_bytes = 0;
    int64_t total_chunks = 0;
    POOL_CUDA(cudaMalloc(&h->tile_cstart, 8 * (size_t)(nt + 1)));
    POOL_CUDA(cudaMalloc(&start, 8 * (size_t)(nt + 1)));
    POOL_CUDA(cudaMalloc(&nchunk, 8 * (size_t)(nt + 1)));
    POOL_CUDA(cudaMalloc(&keys, 8 * cnt));
    POOL_CUDA(cudaMalloc(&keys_out, 8 * cnt));
    POOL_CUDA(cudaMalloc(&vals, 4 * cnt));
    POOL_CUDA(cudaMalloc(&perm, 4 * cnt));
    if (ncand > 0) {
      pool_keys_kernel<<<grid, 256, 0, s>>>(ci_d, cj_d, ncand, tile_rows, jbits, ibits, keys, vals);
      POOL_CUDA(cudaGetLastError());
      POOL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, vals, perm,
                                                ncand, 0, ibits + jbits + tbits, s));
      POOL_CUDA(cudaMalloc(&tmp, tmp_bytes > 0 ? tmp_bytes : 1));
      POOL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, vals, perm, ncand,
                                                0, ibits + jbits + tbits, s));
      cudaFree(tmp); tmp = nullptr;
    }
    pool_tile_start_kernel<<<grid, 256, 0, s>>>(keys_out, ncand, ibits + jbits, nt, start);
    POOL_CUDA(cudaGetLastError());
    pool_nchunk_kernel<<<grid, 256, 0, s>>>(start, nt, nchunk);
    POOL_CUDA(cudaGetLastError());
    POOL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, nchunk, h->tile_cstart, nt + 1, s));
    POOL_CUDA(cudaMalloc(&tmp, scan_bytes > 0 ? scan_bytes : 1));
    POOL_CUDA(cub::DeviceScan::ExclusiveSum(tmp, scan_bytes, nchunk, h->tile_cstart, nt + 1, s));
    POOL_CUDA(cudaMemcpyAsync(&total_chunks, h->tile_cstart + nt, 8, cudaMemcpyDeviceToHost, s));
    POOL_CUDA(cudaStreamSynchronize(s));
    h->n_chunks = total_chunks;
    h->npad = total_chunks * POOL_CHUNK;
    const size_t words = (size_t)(h->npad > 0 ? h->npad : 1);
    POOL_CUDA(cudaMalloc(&h->cw, 4 * words));
    POOL_CUDA(cudaMalloc(&h->orig, 4 * words));
    POOL_CUDA(cudaMemsetAsync(h->cw, 0, 4 * words, s));           // padding scores row 0 of U and of the tile
    POOL_CUDA(cudaMemsetAsync(h->orig, 0xff, 4 * words, s));      // ... and never competes
    POOL_CUDA(cudaMalloc(&h->pos_of, 4 * cnt));
    if (ncand > 0) {
      pool_scatter_kernel<<<grid, 256, 0, s>>>(keys_out, perm, ncand, ibits, jbits, start,
                                               h->tile_cstart, h->cw, h->orig, h->pos_of);
      POOL_CUDA(cudaGetLastError());
    }
    POOL_CUDA(cudaStreamSynchronize(s));
  }
done:
#undef POOL_CUDA
  cudaFree(keys); cudaFree(keys_out); cudaFree(vals); cudaFree(perm); cudaFree(start);
  cudaFree(nchunk); cudaFree(tmp);
  if (rc != AMF_OK) { amf_pool_destroy(h); return rc; }
  *out = h;
  return AMF_OK;
}

int64_t amf_pool_size(const amf_pool_t* h) { return h ? h->ncand : -1; }

__global__ void pool_remove_kernel(const int64_t* __restrict__ idx, int64_t n, int64_t ncand,
                                   const uint32_t* __restrict__ pos_of,
                                   uint32_t* __restrict__ orig) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx[t];
    if (c >= 0 && c < ncand) orig[pos_of[c]] = amf::POOL_TOMBSTONE;
  }
}

int amf_pool_remove(amf_pool_t* h, int64_t n, const int64_t* idx_d, void* stream) {
  AMF_REQUIRE(h && (n == 0 || idx_d), "amf_pool_remove: NULL argument");
  if (n <= 0 || h->ncand == 0) return AMF_OK;
  const int grid = (int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  pool_remove_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(idx_d, n, h->ncand, h->pos_of, h->orig);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

int amf_pool_score_pred(const amf_pool_t* hc, int dtype, int d, int ld, const void* U_d,
                        const void* V_d, void* scores_d, int maximize, int64_t index_base,
                        amf_best_t* best_d, void* stream) {
  AMF_REQUIRE(hc && U_d && V_d && best_d, "amf_pool_score_pred: NULL argument");
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_pool_score_pred: bad dtype");
  amf_pool* h = const_cast<amf_pool*>(hc);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t es = dtype == AMF_F32 ? 4 : 8;
  const int vecn = dtype == AMF_F32 ? 4 : 2;
  AMF_REQUIRE(ld >= d && ld % vecn == 0, "ld=%d must be >= d=%d and a multiple of %d", ld, d, vecn);
  AMF_REQUIRE(((uintptr_t)U_d | (uintptr_t)V_d) % ((size_t)ld * es) == 0 || ((size_t)ld * es & ((size_t)ld * es - 1)),
              "amf_pool_score_pred: U and V must be aligned to the padded row (%d bytes)", (int)(ld * es));
  Best* part = nullptr;
  int rc = acquire_partials(&part, s);
  if (rc != AMF_OK) return rc;
  void* tmp_scores = nullptr;
  if (scores_d && h->npad > 0) {
    if (h->tmp_bytes < es * (size_t)h->npad) {
      cudaFree(h->tmp_scores); h->tmp_scores = nullptr; h->tmp_bytes = 0;
      AMF_CUDA(cudaMalloc(&h->tmp_scores, es * (size_t)h->npad));
      h->tmp_bytes = es * (size_t)h->npad;
    }
    tmp_scores = h->tmp_scores;
  }
  const size_t smem = (size_t)h->tile_rows * ld * es;
  AMF_REQUIRE(smem <= 225 * 1024, "item tile (%d rows of %d) does not fit shared memory",
              h->tile_rows, ld);
  int64_t grid64 = (int64_t)num_sms();
  if (grid64 > h->n_chunks) grid64 = h->n_chunks > 0 ? h->n_chunks : 1;
  const int grid = (int)grid64;
  if (dtype == AMF_F32)
    rc = maximize ? pool_pred<float, true>(h, ld, (const float*)U_d, (const float*)V_d, (float*)tmp_scores, index_base, part, grid, smem, s)
                  : pool_pred<float, false>(h, ld, (const float*)U_d, (const float*)V_d, (float*)tmp_scores, index_base, part, grid, smem, s);
  else
    rc = maximize ? pool_pred<double, true>(h, ld, (const double*)U_d, (const double*)V_d, (double*)tmp_scores, index_base, part, grid, smem, s)
                  : pool_pred<double, false>(h, ld, (const double*)U_d, (const double*)V_d, (double*)tmp_scores, index_base, part, grid, smem, s);
  if (rc != AMF_OK) return rc;
  if (tmp_scores) {
    const int g2 = num_sms() * 8;
    if (dtype == AMF_F32) unpermute_kernel<float><<<g2, 256, 0, s>>>((const float*)tmp_scores, h->orig, h->npad, (float*)scores_d);
    else unpermute_kernel<double><<<g2, 256, 0, s>>>((const double*)tmp_scores, h->orig, h->npad, (double*)scores_d);
    AMF_LAUNCH_CHECK();
  }
  return launch_best_final(part, grid, maximize != 0, best_d, s);
}

#pragma GCC visibility pop
}  // namespace
