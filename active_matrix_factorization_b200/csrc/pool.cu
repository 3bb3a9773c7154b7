// Candidate pool handle: the reference's `pool` / `unrated` set (active_pmf.py:725-737) kept on
// the device in a layout built for the scoring kernel.
//
// The plain scoring kernel (scoring.cu) gathers one 128-byte item row per candidate from L2 and
// sits at the L2->SM gather ceiling.  Here the pool is bucketed once by item tile: candidates
// are sorted by (j / TJ, i, j) so that a CTA can keep a TJ-row tile of V in shared memory (one
// TMA bulk copy per tile), read item rows from shared memory, and touch L2 only for the user
// row of each (user, tile) run.  The permutation back to the caller's order is kept so scores
// and the winner are reported exactly as the unbucketed path would.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "common.cuh"

struct amf_pool {
  int64_t ncand;         // candidates given by the caller
  int64_t npad;          // entries in the bucketed arrays (every bucket padded to a multiple of 8)
  int32_t n_users, n_items;
  int tile_rows;         // TJ: items per tile (V tile resident in shared memory)
  int block_rows;        // TI: users per block (U blocks streamed through shared memory)
  int n_tiles, n_ublocks;
  int64_t n_buckets;     // n_tiles * n_ublocks, bucket = tile * n_ublocks + ublock
  uint32_t* cw;          // [npad] packed local indices  il | jl << 16   (0xffffffff = padding)
  uint32_t* orig;        // [npad] position in the caller's pool (POOL_TOMBSTONE = removed)
  uint32_t* pos_of;      // [ncand] bucketed position of the caller's candidate c
  int64_t* bptr;         // [n_buckets+1] padded start of every bucket
  int32_t* bcnt;         // [n_buckets]   real candidates in every bucket
  // work list: non-empty buckets cut into segments of at most POOL_WORD_CAP candidates
  int64_t n_segs;
  int32_t* seg_bucket;   // [n_segs]
  int64_t* seg_ptr;      // [n_segs+1] start in cw/orig (multiple of 8); seg_ptr[n_segs] = npad
  int32_t* seg_cnt;      // [n_segs]
  void* tmp_scores;      // scratch for scores in bucketed order
  size_t tmp_bytes;
};

namespace amf {

int acquire_partials(Best** out, cudaStream_t s);
int launch_best_final(const Best* part_d, int nparts, bool maximize, amf_best_t* out_d,
                      cudaStream_t s);

__global__ void pool_keys_kernel(const int32_t* __restrict__ ci, const int32_t* __restrict__ cj,
                                 int64_t n, int tile_rows, int block_rows, int n_ublocks,
                                 int ibits, int jbits, uint64_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t i = (uint32_t)ci[t], j = (uint32_t)cj[t];
    const uint64_t bucket = (uint64_t)(j / (uint32_t)tile_rows) * (uint64_t)n_ublocks +
                            (uint64_t)(i / (uint32_t)block_rows);
    const uint64_t il = i % (uint32_t)block_rows, jl = j % (uint32_t)tile_rows;
    keys[t] = (((bucket << ibits) | il) << jbits) | jl;
    vals[t] = (uint32_t)t;
  }
}

// keys sorted ascending; start[b] = first position whose bucket >= b, start[n_buckets] = n
__global__ void pool_bucket_start_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                         int64_t n_buckets, int64_t* __restrict__ start) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p <= n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lo = (p == 0) ? -1 : (int64_t)(keys[p - 1] >> shift);
    const int64_t hi = (p == n) ? n_buckets : (int64_t)(keys[p] >> shift);
    for (int64_t b = lo + 1; b <= hi; ++b) start[b] = p;
  }
}

__global__ void pool_count_kernel(const int64_t* __restrict__ start, int64_t n_buckets,
                                  int32_t* __restrict__ cnt, int64_t* __restrict__ padded) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < n_buckets;
       b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = start[b + 1] - start[b];
    cnt[b] = (int32_t)c;
    padded[b] = (c + 7) & ~7ll;
  }
}

__global__ void pool_scatter_kernel(const uint64_t* __restrict__ keys,
                                    const uint32_t* __restrict__ perm, int64_t n, int ibits,
                                    int jbits, const int64_t* __restrict__ start,
                                    const int64_t* __restrict__ bptr, uint32_t* __restrict__ cw,
                                    uint32_t* __restrict__ orig, uint32_t* __restrict__ pos_of) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[p];
    const int64_t b = (int64_t)(k >> (ibits + jbits));
    const uint32_t jl = (uint32_t)(k & ((1ull << jbits) - 1));
    const uint32_t il = (uint32_t)((k >> jbits) & ((1ull << ibits) - 1));
    const int64_t q = bptr[b] + (p - start[b]);
    cw[q] = il | (jl << 16);
    orig[q] = perm[p];
    pos_of[perm[p]] = (uint32_t)q;
  }
}

constexpr int POOL_WORD_CAP = 4096;   // packed index words staged per segment (16 KB)
constexpr int POOL_THREADS = 512;
constexpr uint32_t POOL_TOMBSTONE = 0xffffffffu;   // orig[] value of a removed candidate
constexpr int POOL_MAX_STAGES = 8;

static int bits_for_count(uint64_t x) {
  int b = 1;
  while (b < 63 && (1ull << b) < x) ++b;
  return b;
}

// ---- TMA (bulk async copy) + mbarrier helpers: inline PTX for sm_100a ----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (TMA, 1-D), completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}


__global__ void pool_nseg_kernel(const int32_t* __restrict__ bcnt, int64_t n_buckets, int cap,
                                 int64_t* __restrict__ nseg) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < n_buckets;
       b += (int64_t)gridDim.x * blockDim.x)
    nseg[b] = (bcnt[b] + cap - 1) / cap;
}

__global__ void pool_fill_segs_kernel(const int32_t* __restrict__ bcnt,
                                      const int64_t* __restrict__ bptr,
                                      const int64_t* __restrict__ seg_off, int64_t n_buckets,
                                      int cap, int32_t* __restrict__ seg_bucket,
                                      int64_t* __restrict__ seg_ptr, int32_t* __restrict__ seg_cnt) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < n_buckets;
       b += (int64_t)gridDim.x * blockDim.x) {
    const int cnt = bcnt[b];
    int64_t o = seg_off[b];
    for (int q = 0; q * cap < cnt; ++q, ++o) {
      seg_bucket[o] = (int32_t)b;
      seg_ptr[o] = bptr[b] + (int64_t)q * cap;
      seg_cnt[o] = min(cap, cnt - q * cap);
    }
  }
}

__device__ __forceinline__ float4 lds_v(uint32_t addr, float4) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ double2 lds_v(uint32_t addr, double2) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// PRED over a bucketed pool.  Each CTA (one per SM) owns a contiguous range of segments balanced
// by candidate count.  The item tile of the current segment stays in shared memory; a ring of
// `nstage` buffers receives, by TMA, the user block and the packed index words of the next
// segments while the current one is scored, so the scoring loop touches shared memory only.
// Lane groups of LPR lanes own LPR consecutive candidates (LPR packed words = two 16-byte
// shared loads), each lane holds one 16-byte slice (x VPL) of the rows; the user slice is
// re-read only when the user changes.  tile_rows and block_rows are powers of two: indices
// are masked, so padding words and stale tails need no branches.
template <typename T, int LPR, int VPL, bool MAX>
__global__ void __launch_bounds__(POOL_THREADS)
pool_pred_kernel(const uint32_t* __restrict__ cw, const uint32_t* __restrict__ orig,
                 const int32_t* __restrict__ seg_bucket, const int64_t* __restrict__ seg_ptr,
                 const int32_t* __restrict__ seg_cnt, int64_t n_segs, int n_ublocks,
                 int64_t npad, int tile_rows, int block_rows, int n_users, int n_items,
                 const T* __restrict__ U, const T* __restrict__ Vm, int nstage,
                 T* __restrict__ scores, int64_t index_base, Best* __restrict__ part) {
  using V = typename Vec<T>::type;
  constexpr int NVEC = LPR * VPL;                    // 16-byte vectors per factor row (ld = NVEC*N)
  constexpr int LD = NVEC * Vec<T>::N;
  constexpr uint32_t ROW_BYTES = NVEC * 16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t tile_bytes = (uint32_t)tile_rows * ROW_BYTES;
  const uint32_t blk_bytes = (uint32_t)block_rows * ROW_BYTES;
  const uint32_t stage_bytes = blk_bytes + (uint32_t)POOL_WORD_CAP * 4;
  const uint32_t smem0 = smem_u32(smem_raw);
  __shared__ __align__(8) uint64_t bar_v, bar_u[POOL_MAX_STAGES];
  __shared__ int64_t krange[2];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  constexpr int NW = POOL_THREADS / 32;
  const int g = lane / LPR, l = lane % LPR;
  const uint32_t imask = (uint32_t)block_rows - 1, jmask = (uint32_t)tile_rows - 1;
  if (threadIdx.x == 0) {
    mbar_init(&bar_v, 1);
    for (int q = 0; q < POOL_MAX_STAGES; ++q) mbar_init(&bar_u[q], 1);
    for (int e = 0; e < 2; ++e) {                    // segments starting in this CTA's share
      const int64_t target = npad / gridDim.x * (blockIdx.x + e) +
                             min((int64_t)(blockIdx.x + e), npad % (int64_t)gridDim.x);
      int64_t lo = 0, hi = n_segs;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (seg_ptr[mid] < target) lo = mid + 1; else hi = mid;
      }
      krange[e] = lo;
    }
    if (blockIdx.x == gridDim.x - 1) krange[1] = n_segs;
  }
  __syncthreads();
  const int64_t k_end = krange[1];
  int64_t k = krange[0];

  auto issue_u = [&](int64_t kk, int stage) {        // thread 0 only
    const int ub = seg_bucket[kk] % n_ublocks;
    const int rows = min(block_rows, n_users - ub * block_rows);
    const uint32_t bytes = (uint32_t)rows * ROW_BYTES;
    const uint32_t wbytes = (uint32_t)((seg_cnt[kk] + 7) & ~7) * 4u;
    unsigned char* dst = smem_raw + tile_bytes + (size_t)stage * stage_bytes;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&bar_u[stage], bytes + wbytes);
    tma_load_1d(dst, U + (int64_t)ub * block_rows * LD, bytes, &bar_u[stage]);
    tma_load_1d(dst + blk_bytes, cw + seg_ptr[kk], wbytes, &bar_u[stage]);
  };

  uint32_t phase_v = 0;
  int cur_tile = -1;
  T best_v = MAX ? -INFINITY : INFINITY;
  int64_t best_o = -1;
  int64_t kp = k;                                    // producer cursor (thread 0 only)
  if (threadIdx.x == 0)
    for (int q = 0; q < nstage && kp < k_end; ++q, ++kp) issue_u(kp, q);

  for (int it = 0; k < k_end; ++it, ++k) {
    const int stage = it % nstage;
    const uint32_t parity = (uint32_t)(it / nstage) & 1u;
    const int t = seg_bucket[k] / n_ublocks;
    if (t != cur_tile) {
      if (threadIdx.x == 0) {
        const int rows = min(tile_rows, n_items - t * tile_rows);
        const uint32_t bytes = (uint32_t)rows * ROW_BYTES;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bar_v, bytes);
        const unsigned char* src =
            reinterpret_cast<const unsigned char*>(Vm + (int64_t)t * tile_rows * LD);
        for (uint32_t off = 0; off < bytes; off += 32768u)
          tma_load_1d(smem_raw + off, src + off, min(bytes - off, 32768u), &bar_v);
      }
      mbar_wait(&bar_v, phase_v);
      phase_v ^= 1;
      cur_tile = t;
    }
    mbar_wait(&bar_u[stage], parity);
    const uint32_t u_base = smem0 + tile_bytes + (uint32_t)stage * stage_bytes;
    const uint32_t w_base = u_base + blk_bytes;
    const int64_t cb = seg_ptr[k];
    const int cnt = seg_cnt[k];

    for (int base = wib * 32; base < cnt; base += NW * 32) {
      // the group's LPR packed words (stale words past the staged length are harmless: masked)
      uint32_t w[LPR];
      const uint32_t waddr = w_base + (uint32_t)(base + g * LPR) * 4u;
      if constexpr (LPR == 8) {
        const uint4 w0 = lds_u4(waddr), w1 = lds_u4(waddr + 16);
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
        w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
      } else if constexpr (LPR == 4) {
        const uint4 w0 = lds_u4(waddr);
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
      } else {
#pragma unroll
        for (int s = 0; s < LPR; ++s)
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[s]) : "r"(waddr + 4u * s));
      }
      T p[LPR];
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const uint32_t coff = (uint32_t)(l + v * LPR) * 16u;
        // all item-row loads of the batch first (LPR independent shared loads in flight) ...
        V b[LPR];
#pragma unroll
        for (int s = 0; s < LPR; ++s)
          b[s] = lds_v(smem0 + ((w[s] >> 16) & jmask) * ROW_BYTES + coff, V());
        // ... then the user rows (re-read only when the user changes) and the products
        V a = vzero(V());
        uint32_t prev_il = 0xffffffffu;
#pragma unroll
        for (int s = 0; s < LPR; ++s) {
          const uint32_t il = w[s] & imask;
          if (il != prev_il) a = lds_v(u_base + il * ROW_BYTES + coff, V());
          prev_il = il;
          p[s] = (v == 0) ? vdot(a, b[s]) : p[s] + vdot(a, b[s]);
        }
      }
#pragma unroll
      for (int half = LPR >> 1; half >= 1; half >>= 1) {
        const bool upper = (l & half) != 0;
#pragma unroll
        for (int t2 = 0; t2 < half; ++t2) {
          const T send = upper ? p[t2] : p[t2 + half];
          const T keep = upper ? p[t2 + half] : p[t2];
          p[t2] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
      }
      const bool live = base + lane < cnt;
      if (scores && live) __stcs(scores + cb + base + lane, p[0]);
      if (live && (MAX ? (p[0] >= best_v) : (p[0] <= best_v))) {     // rare after warm-up
        const uint32_t ou = orig[cb + base + lane];
        const int64_t o = (int64_t)ou;
        if (ou != POOL_TOMBSTONE &&
            (best_o < 0 || (MAX ? (p[0] > best_v) : (p[0] < best_v)) || o < best_o)) {
          best_v = p[0]; best_o = o;
        }
      }
    }
    __syncthreads();        // segment k fully scored: its stage (and the tile) may be replaced
    if (threadIdx.x == 0 && kp < k_end) { issue_u(kp, stage); ++kp; }
  }
  Best best{(double)best_v, best_o < 0 ? -1 : best_o + index_base};
  best = block_best<MAX>(best);
  if (threadIdx.x == 0) part[blockIdx.x] = best;
}

template <typename T>
__global__ void unpermute_kernel(const T* __restrict__ in, const uint32_t* __restrict__ orig,
                                 const uint32_t* __restrict__ cw, int64_t npad,
                                 T* __restrict__ out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < npad;
       t += (int64_t)gridDim.x * blockDim.x)
    if (cw[t] != 0xffffffffu && orig[t] != POOL_TOMBSTONE) out[orig[t]] = in[t];
}

template <typename T, bool MAX>
static int pool_pred(const amf_pool* h, int ld, const T* U, const T* V, T* scores_tmp,
                     int64_t index_base, Best* part, int grid, size_t smem, int nstage,
                     cudaStream_t s) {
  constexpr int N = Vec<T>::N;
  const int nvec = ld / N;
#define POOL(LPR_, VPL_)                                                                        \
  do {                                                                                          \
    AMF_CUDA(cudaFuncSetAttribute(pool_pred_kernel<T, LPR_, VPL_, MAX>,                         \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    pool_pred_kernel<T, LPR_, VPL_, MAX><<<grid, POOL_THREADS, smem, s>>>(                      \
        h->cw, h->orig, h->seg_bucket, h->seg_ptr, h->seg_cnt, h->n_segs, h->n_ublocks,         \
        h->npad, h->tile_rows, h->block_rows, h->n_users, h->n_items, U, V, nstage,             \
        scores_tmp, index_base, part);                                                          \
  } while (0)
  switch (nvec) {            // the row width is a compile-time constant of the kernel
    case 1: POOL(1, 1); break;
    case 2: POOL(2, 1); break;
    case 4: POOL(4, 1); break;
    case 8: POOL(8, 1); break;
    case 16: POOL(8, 2); break;
    case 32: POOL(8, 4); break;
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
  cudaFree(h->cw); cudaFree(h->orig); cudaFree(h->pos_of); cudaFree(h->bptr); cudaFree(h->bcnt); cudaFree(h->tmp_scores);
  cudaFree(h->seg_bucket); cudaFree(h->seg_ptr); cudaFree(h->seg_cnt);
  delete h;
  return AMF_OK;
}

int amf_pool_create(amf_pool_t** out, int64_t ncand, const int32_t* ci_d, const int32_t* cj_d,
                    int32_t n_users, int32_t n_items, int tile_rows, int block_rows, void* stream) {
  AMF_REQUIRE(out && n_users > 0 && n_items > 0, "amf_pool_create: bad arguments");
  AMF_REQUIRE(ncand >= 0 && ncand < (1ll << 32) - 16, "amf_pool_create: ncand out of range");
  AMF_REQUIRE(tile_rows >= 1 && tile_rows <= 32768 && (tile_rows & (tile_rows - 1)) == 0 &&
              block_rows >= 1 && block_rows <= 32768 && (block_rows & (block_rows - 1)) == 0,
              "amf_pool_create: tile_rows and block_rows must be powers of two <= 32768");
  cudaStream_t s = (cudaStream_t)stream;
  amf_pool* h = new amf_pool();
  memset(h, 0, sizeof(*h));
  h->ncand = ncand; h->n_users = n_users; h->n_items = n_items;
  h->tile_rows = tile_rows; h->block_rows = block_rows;
  h->n_tiles = (n_items + tile_rows - 1) / tile_rows;
  h->n_ublocks = (n_users + block_rows - 1) / block_rows;
  h->n_buckets = (int64_t)h->n_tiles * h->n_ublocks;
  int rc = AMF_OK;
  uint64_t *keys = nullptr, *keys_out = nullptr;
  uint32_t *vals = nullptr, *perm = nullptr;
  int64_t *start = nullptr, *padded = nullptr;
  void* tmp = nullptr;
  const size_t cnt = ncand > 0 ? (size_t)ncand : 1;
  const int grid = num_sms() * 8;
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
    const int ibits = bits_for_count((uint64_t)block_rows), jbits = bits_for_count((uint64_t)tile_rows);
    const int bbits = bits_for_count((uint64_t)h->n_buckets + 1);
    size_t tmp_bytes = 0, scan_bytes = 0;
    int64_t total_pad = 0;
    POOL_CUDA(cudaMalloc(&h->bptr, 8 * (size_t)(h->n_buckets + 1)));
    POOL_CUDA(cudaMalloc(&h->bcnt, 4 * (size_t)h->n_buckets));
    POOL_CUDA(cudaMalloc(&start, 8 * (size_t)(h->n_buckets + 1)));
    POOL_CUDA(cudaMalloc(&padded, 8 * (size_t)(h->n_buckets + 1)));
    POOL_CUDA(cudaMalloc(&keys, 8 * cnt));
    POOL_CUDA(cudaMalloc(&keys_out, 8 * cnt));
    POOL_CUDA(cudaMalloc(&vals, 4 * cnt));
    POOL_CUDA(cudaMalloc(&perm, 4 * cnt));
    if (ncand > 0) {
      pool_keys_kernel<<<grid, 256, 0, s>>>(ci_d, cj_d, ncand, tile_rows, block_rows, h->n_ublocks,
                                            ibits, jbits, keys, vals);
      POOL_CUDA(cudaGetLastError());
      POOL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, vals, perm,
                                                ncand, 0, ibits + jbits + bbits, s));
      POOL_CUDA(cudaMalloc(&tmp, tmp_bytes > 0 ? tmp_bytes : 1));
      POOL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, vals, perm, ncand,
                                                0, ibits + jbits + bbits, s));
      cudaFree(tmp); tmp = nullptr;
    }
    pool_bucket_start_kernel<<<grid, 256, 0, s>>>(keys_out, ncand, ibits + jbits, h->n_buckets, start);
    POOL_CUDA(cudaGetLastError());
    pool_count_kernel<<<grid, 256, 0, s>>>(start, h->n_buckets, h->bcnt, padded);
    POOL_CUDA(cudaGetLastError());
    POOL_CUDA(cudaMemsetAsync(padded + h->n_buckets, 0, 8, s));
    POOL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, padded, h->bptr, h->n_buckets + 1, s));
    POOL_CUDA(cudaMalloc(&tmp, scan_bytes > 0 ? scan_bytes : 1));
    POOL_CUDA(cub::DeviceScan::ExclusiveSum(tmp, scan_bytes, padded, h->bptr, h->n_buckets + 1, s));
    POOL_CUDA(cudaMemcpyAsync(&total_pad, h->bptr + h->n_buckets, 8, cudaMemcpyDeviceToHost, s));
    POOL_CUDA(cudaStreamSynchronize(s));
    h->npad = total_pad;
    POOL_CUDA(cudaMalloc(&h->cw, 4 * (size_t)(total_pad > 0 ? total_pad : 1) + 256));
    POOL_CUDA(cudaMalloc(&h->orig, 4 * (size_t)(total_pad > 0 ? total_pad : 1) + 256));
    POOL_CUDA(cudaMemsetAsync(h->cw, 0xff, 4 * (size_t)(total_pad > 0 ? total_pad : 1) + 256, s));
    POOL_CUDA(cudaMalloc(&h->pos_of, 4 * cnt));
    if (ncand > 0) {
      pool_scatter_kernel<<<grid, 256, 0, s>>>(keys_out, perm, ncand, ibits, jbits, start, h->bptr,
                                               h->cw, h->orig, h->pos_of);
      POOL_CUDA(cudaGetLastError());
    }
    // segment list
    {
      int64_t *nseg = nullptr, *seg_off = nullptr;
      int64_t total = 0;
      size_t sb = 0;
      POOL_CUDA(cudaMalloc(&nseg, 8 * (size_t)(h->n_buckets + 1)));
      POOL_CUDA(cudaMalloc(&seg_off, 8 * (size_t)(h->n_buckets + 1)));
      pool_nseg_kernel<<<grid, 256, 0, s>>>(h->bcnt, h->n_buckets, POOL_WORD_CAP, nseg);
      POOL_CUDA(cudaMemsetAsync(nseg + h->n_buckets, 0, 8, s));
      POOL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, sb, nseg, seg_off, h->n_buckets + 1, s));
      cudaFree(tmp); tmp = nullptr;
      POOL_CUDA(cudaMalloc(&tmp, sb > 0 ? sb : 1));
      POOL_CUDA(cub::DeviceScan::ExclusiveSum(tmp, sb, nseg, seg_off, h->n_buckets + 1, s));
      POOL_CUDA(cudaMemcpyAsync(&total, seg_off + h->n_buckets, 8, cudaMemcpyDeviceToHost, s));
      POOL_CUDA(cudaStreamSynchronize(s));
      h->n_segs = total;
      POOL_CUDA(cudaMalloc(&h->seg_bucket, 4 * (size_t)(total + 1)));
      POOL_CUDA(cudaMalloc(&h->seg_ptr, 8 * (size_t)(total + 1)));
      POOL_CUDA(cudaMalloc(&h->seg_cnt, 4 * (size_t)(total + 1)));
      pool_fill_segs_kernel<<<grid, 256, 0, s>>>(h->bcnt, h->bptr, seg_off, h->n_buckets,
                                                 POOL_WORD_CAP, h->seg_bucket, h->seg_ptr, h->seg_cnt);
      POOL_CUDA(cudaGetLastError());
      POOL_CUDA(cudaMemcpyAsync(h->seg_ptr + total, &h->npad, 8, cudaMemcpyHostToDevice, s));
      POOL_CUDA(cudaStreamSynchronize(s));
      cudaFree(nseg); cudaFree(seg_off);
    }
  }
done:
#undef POOL_CUDA
  cudaFree(keys); cudaFree(keys_out); cudaFree(vals); cudaFree(perm); cudaFree(start);
  cudaFree(padded); cudaFree(tmp);
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
  // shared memory: the item tile + a ring of (user block + packed indices) stages.  As many
  // stages as fit (up to POOL_MAX_STAGES): the ring depth is what keeps enough bytes in flight
  // to cover the L2 latency of the user-block stream.
  const size_t tile_b = (size_t)h->tile_rows * ld * es;
  const size_t stage_b = (size_t)h->block_rows * ld * es + (size_t)POOL_WORD_CAP * 4;
  const size_t budget = 226 * 1024;
  AMF_REQUIRE(tile_b + 2 * stage_b <= budget, "item tile (%d rows) + 2 user blocks (%d rows) of "
              "%d do not fit shared memory", h->tile_rows, h->block_rows, ld);
  int nstage = (int)((budget - tile_b) / stage_b);
  if (nstage > POOL_MAX_STAGES) nstage = POOL_MAX_STAGES;
  const size_t smem = tile_b + (size_t)nstage * stage_b;
  int64_t grid64 = (int64_t)num_sms();
  const int64_t max_useful = h->npad / 512 + 1;
  if (grid64 > max_useful) grid64 = max_useful;
  const int grid = (int)grid64;
  if (dtype == AMF_F32)
    rc = maximize ? pool_pred<float, true>(h, ld, (const float*)U_d, (const float*)V_d, (float*)tmp_scores, index_base, part, grid, smem, nstage, s)
                  : pool_pred<float, false>(h, ld, (const float*)U_d, (const float*)V_d, (float*)tmp_scores, index_base, part, grid, smem, nstage, s);
  else
    rc = maximize ? pool_pred<double, true>(h, ld, (const double*)U_d, (const double*)V_d, (double*)tmp_scores, index_base, part, grid, smem, nstage, s)
                  : pool_pred<double, false>(h, ld, (const double*)U_d, (const double*)V_d, (double*)tmp_scores, index_base, part, grid, smem, nstage, s);
  if (rc != AMF_OK) return rc;
  if (tmp_scores) {
    const int g2 = num_sms() * 8;
    if (dtype == AMF_F32) unpermute_kernel<float><<<g2, 256, 0, s>>>((const float*)tmp_scores, h->orig, h->cw, h->npad, (float*)scores_d);
    else unpermute_kernel<double><<<g2, 256, 0, s>>>((const double*)tmp_scores, h->orig, h->cw, h->npad, (double*)scores_d);
    AMF_LAUNCH_CHECK();
  }
  return launch_best_final(part, grid, maximize != 0, best_d, s);
}

#pragma GCC visibility pop
}  // extern "C"
