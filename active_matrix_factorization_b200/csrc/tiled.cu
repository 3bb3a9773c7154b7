// Tiled copy of the rating list and the fused PMF loss + gradient that runs on it
// (pmf_cy.pyx:170-193 log_likelihood, :204-223 gradient).
//
// The row-sorted side pass (pmf.cu) gathers one factor row of the other side per rating from
// L2 and is bound by the L2->SM path.  Here every side of the gradient gets a second copy of
// the list in the "bundled runs" layout (runs.cuh): bucketed by TILE of the other side's matrix,
// one lane per (own row, tile) run segment, 6 bytes per rating (16-bit row inside the tile + the
// rating).  A CTA keeps the tile of the other side in shared memory (TMA bulk copies); a lane
// keeps its whole own row AND the whole gradient accumulator of that row in registers and reads
// one tile row per rating (one shared-memory wavefront: the floor of a CUDA-core formulation),
// computes the residual and the rank-one update without a single shuffle; L2 is touched once per
// bundle of 32 runs: a coalesced fetch of the 32 own rows and a coalesced vector RED.ADD of the
// 32 accumulated  sum_j (e_ij / sigma^2) * Other_j  (fetch_rows / flush_rows).
//   side 0: users stream past item tiles  -> dU and the squared error
//   side 1: items stream past user tiles  -> dV
#include <stdlib.h>

#include "common.cuh"
#include "runs.cuh"

namespace amf {

constexpr int64_t TILED_AUTO_MIN_NNZ = 1ll << 20;
constexpr int64_t TILED_BUNDLE_COST = 10;   // row fetch + flush of a bundle, in entry steps (micro_visit.cu)
constexpr size_t TILED_SMEM_BUDGET = 227 * 1024 - 1024;

#ifndef AMF_TILED_THREADS
#define AMF_TILED_THREADS 448                        // measured at C5 (benchmarks/variant_lib.sh): 256 1.094, 320 1.016,
                                                     // 352 0.971, 384 0.943, 416 1.003, 448 0.925, 480 0.993, 512 0.952 ms
#endif
template <typename T, int NVEC> constexpr int tiled_threads() {
  return NVEC * 12 > 96 ? 256 : AMF_TILED_THREADS;   // three rows of registers per lane (own, accumulator, tile row)
}
template <typename T, int NVEC> constexpr size_t tiled_stage_total() {
  return (size_t)(tiled_threads<T, NVEC>() / 32) * runs_stage_bytes<NVEC>();
}

template <int C>
__device__ __forceinline__ float dot_rows(const float4 (&a)[C], const float4 (&b)[C]) {
  float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
#pragma unroll
  for (int t = 0; t < C; ++t) {
    if (t & 1) {
      s2 = fma2(make_float2(a[t].x, a[t].y), make_float2(b[t].x, b[t].y), s2);
      s3 = fma2(make_float2(a[t].z, a[t].w), make_float2(b[t].z, b[t].w), s3);
    } else {
      s0 = fma2(make_float2(a[t].x, a[t].y), make_float2(b[t].x, b[t].y), s0);
      s1 = fma2(make_float2(a[t].z, a[t].w), make_float2(b[t].z, b[t].w), s1);
    }
  }
  return ((s0.x + s1.x) + (s2.x + s3.x)) + ((s0.y + s1.y) + (s2.y + s3.y));
}
template <int C>
__device__ __forceinline__ double dot_rows(const double2 (&a)[C], const double2 (&b)[C]) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
  for (int t = 0; t < C; ++t) {
    if (t & 1) { s2 = fma(a[t].x, b[t].x, s2); s3 = fma(a[t].y, b[t].y, s3); }
    else { s0 = fma(a[t].x, b[t].x, s0); s1 = fma(a[t].y, b[t].y, s1); }
  }
  return (s0 + s1) + (s2 + s3);
}
template <int C>
__device__ __forceinline__ void axpy_rows(float4 (&acc)[C], float w, const float4 (&b)[C]) {
  const float2 w2 = make_float2(w, w);
#pragma unroll
  for (int t = 0; t < C; ++t) {
    const float2 lo = fma2(w2, make_float2(b[t].x, b[t].y), make_float2(acc[t].x, acc[t].y));
    const float2 hi = fma2(w2, make_float2(b[t].z, b[t].w), make_float2(acc[t].z, acc[t].w));
    acc[t] = make_float4(lo.x, lo.y, hi.x, hi.y);
  }
}
template <int C>
__device__ __forceinline__ void axpy_rows(double2 (&acc)[C], double w, const double2 (&b)[C]) {
#pragma unroll
  for (int t = 0; t < C; ++t) vfma(acc[t], w, b[t]);
}
__device__ __forceinline__ void ld_vals4(const float* p, float (&out)[4]) {
  const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
__device__ __forceinline__ void ld_vals4(const double* p, double (&out)[4]) {
  const double2 v0 = __ldcs(reinterpret_cast<const double2*>(p));
  const double2 v1 = __ldcs(reinterpret_cast<const double2*>(p) + 1);
  out[0] = v0.x; out[1] = v0.y; out[2] = v1.x; out[3] = v1.y;
}

// One side of the fused loss + gradient on the bundled-runs list.  Own = the matrix whose rows
// stream (row + gradient accumulator in registers), Tile = the matrix whose tile sits in shared
// memory.  Work distribution (home tile + global per-tile counters) and tile loading are those of
// pool_pred_kernel.
template <typename T, int NVEC, bool GRAD, bool FLUSH_TMA>
__global__ void __launch_bounds__(tiled_threads<T, NVEC>(), 1)
tiled_side_kernel(const uint16_t* __restrict__ idx, const T* __restrict__ rv,
                  const uint32_t* __restrict__ rowid, const uint8_t* __restrict__ seglen,
                  const int2* __restrict__ binfo, const int64_t* __restrict__ tile_bstart,
                  uint32_t* tile_ctr, int n_tiles, int64_t n_bundles, int tile_rows, int tile_side_rows,
                  const T* __restrict__ Own, const T* __restrict__ Tile, T inv_sigma,
                  T mean_offset, T* __restrict__ dOwn, double* __restrict__ sq_err) {
  using V = typename Vec<T>::type;
  constexpr uint32_t ROW_BYTES = NVEC * 16;
  constexpr int THREADS = tiled_threads<T, NVEC>();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_v;
  __shared__ int s_next;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t stage = smem_u32(smem_raw) + warp * runs_stage_bytes<NVEC>();
  unsigned char* tile_ptr = smem_raw + tiled_stage_total<T, NVEC>();
  const uint32_t tile0 = smem_u32(tile_ptr) | runs_lane_rot<NVEC>(lane);
  const unsigned char* own_b = reinterpret_cast<const unsigned char*>(Own);
  unsigned char* down_b = reinterpret_cast<unsigned char*>(dOwn);

  const int64_t total = runs_cost(binfo, n_bundles, TILED_BUNDLE_COST);
  const int64_t b_lo = runs_split(binfo, n_bundles, TILED_BUNDLE_COST, total * blockIdx.x / gridDim.x);
  if (threadIdx.x == 0) mbar_init(&bar_v, 1);
  // the row behind the tile: what padding entries read (their residual is masked by the length)
  for (int t = threadIdx.x; t < (int)(ROW_BYTES / sizeof(T)); t += THREADS)
    reinterpret_cast<T*>(tile_ptr + (size_t)tile_rows * ROW_BYTES)[t] = T(0);
  int t_cur = 0;
  {                                                   // last tile starting at or before b_lo
    int lo = 0, hi = n_tiles;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tile_bstart[mid] <= b_lo) lo = mid; else hi = mid - 1;
    }
    t_cur = lo;
  }
  double local_sq = 0;
  uint32_t phase_v = 0;

  for (int hop = 0; hop < n_tiles;) {
    const int r_hop = runs_next_tile(tile_ctr, tile_bstart, n_tiles, t_cur + hop, n_tiles - hop, &s_next);
    if (r_hop < 0) break;
    hop += r_hop;
    const int t_now = (t_cur + hop) % n_tiles;
    ++hop;
    const int64_t c = tile_bstart[t_now], seg_end = tile_bstart[t_now + 1];
    if (threadIdx.x == 0) {
      const int rows = min(tile_rows, tile_side_rows - t_now * tile_rows);
      const uint32_t bytes = (uint32_t)rows * ROW_BYTES;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bar_v, bytes);
      const unsigned char* src =
          reinterpret_cast<const unsigned char*>(Tile) + (int64_t)t_now * tile_rows * ROW_BYTES;
      for (uint32_t o = 0; o < bytes; o += 32768u)
        tma_load_1d(tile_ptr + o, src + o, min(bytes - o, 32768u), &bar_v);
    }
    __syncthreads();
    mbar_wait(&bar_v, phase_v);
    phase_v ^= 1;

    auto grab = [&]() -> int64_t {                    // the tile's global counter: shared by all CTAs on it
      unsigned int g = 0;
      if (lane == 0) g = atomicAdd(tile_ctr + t_now, 1u);
      return c + (int64_t)__shfl_sync(0xffffffffu, g, 0);
    };
    int64_t b = grab();
    int2 info = make_int2(0, 0);
    uint32_t rid = RUNS_NONE;
    int mylen = 0;
    if (b < seg_end) { info = binfo[b]; rid = rowid[b * 32 + lane]; mylen = seglen[b * 32 + lane]; }
    while (b < seg_end) {
      const int64_t nb = grab();
      int2 ninfo = make_int2(0, 0);
      uint32_t nrid = RUNS_NONE;
      int nlen = 0;
      if (nb < seg_end) { ninfo = binfo[nb]; nrid = rowid[nb * 32 + lane]; nlen = seglen[nb * 32 + lane]; }

      V a[NVEC], acc[NVEC];
      fetch_rows<V, NVEC>(own_b, rid, stage, lane, a);
#pragma unroll
      for (int t = 0; t < NVEC; ++t) acc[t] = vzero(V());
      const int L = info.y, G = (L + 3) >> 2;
      const uint2* ip = reinterpret_cast<const uint2*>(idx) + (int64_t)info.x * 32 + lane;
      const T* rp = rv + (int64_t)info.x * 128 + lane * 4;
      uint2 w = __ldcs(ip);
      T r4[4];
      ld_vals4(rp, r4);
      T sq = 0;
#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        uint2 wn = w;
        T rn[4] = {r4[0], r4[1], r4[2], r4[3]};
        if (g + 1 < G) {
          wn = __ldcs(ip + (g + 1) * 32);
          ld_vals4(rp + (g + 1) * 128, rn);
        }
        const int ns = L - 4 * g;
        const uint32_t j4[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (s < ns) {
            AMF_DBG_ASSERT((int)j4[s] <= tile_rows);
            V bb[NVEC];
            lds_row<V, NVEC>(tile0 + j4[s] * ROW_BYTES, bb);
            T e = (r4[s] - mean_offset) - dot_rows<NVEC>(a, bb);
            e = (4 * g + s < mylen) ? e : T(0);
            sq = fma(e, e, sq);
            if (GRAD) axpy_rows<NVEC>(acc, e * inv_sigma, bb);
          }
        }
        w = wn;
#pragma unroll
        for (int q = 0; q < 4; ++q) r4[q] = rn[q];
      }
      if (GRAD) {
        if (FLUSH_TMA) flush_rows_tma<T, V, NVEC>(down_b, rid, stage, lane, acc);
        else flush_rows<T, V, NVEC>(down_b, rid, stage, lane, acc);
      }
      local_sq += (double)sq;
      b = nb; info = ninfo; rid = nrid; mylen = nlen;
    }
  }
  // bulk reductions still in flight complete before the CTA retires
  if (GRAD && FLUSH_TMA) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (sq_err) {
    const double s = block_sum(local_sq);
    if (threadIdx.x == 0) atomicAdd(sq_err, s);
  }
}

template <typename T, int NVEC> static size_t tiled_smem(int tile_rows) {
  return tiled_stage_total<T, NVEC>() + ((size_t)tile_rows + 1) * NVEC * 16;
}
// the tallest tile of rows of `row_bytes` the kernel can keep next to its staging buffers;
// AMF_TILED_KB caps it (tuning knob of benchmarks/tiled_variants.py)
template <typename T>
static int tiled_max_tile_rows(size_t row_bytes) {
  size_t stage = 0;
  switch (row_bytes) {
    case 64: stage = tiled_stage_total<T, 4>(); break;
    case 128: stage = tiled_stage_total<T, 8>(); break;
    case 256: stage = tiled_stage_total<T, 16>(); break;
    default: return 0;
  }
  size_t bytes = TILED_SMEM_BUDGET - stage - row_bytes;
  const char* e = getenv("AMF_TILED_KB");
  if (e && atoi(e) >= 16 && (size_t)atoi(e) * 1024 < bytes) bytes = (size_t)atoi(e) * 1024;
  const size_t rows = bytes / row_bytes;
  return (int)(rows > 65535 ? 65535 : rows);
}

template <typename T, bool GRAD>
static int launch_tiled(const amf_ratings* h, int side, int nvec, const T* Own, const T* Tile,
                        T inv_sigma, T mean_offset, T* dOwn, double* sq_err, cudaStream_t s,
                        int max_ctas) {
  const amf_runs* r = &h->tiled[side];
  const int tile_side_rows = side == 0 ? h->n_items : h->n_users;
  int64_t grid64 = (int64_t)num_sms();
  if (max_ctas > 0 && grid64 > max_ctas) grid64 = max_ctas;   // leave SMs to a concurrent collective
  if (grid64 > r->n_bundles) grid64 = r->n_bundles > 0 ? r->n_bundles : 1;
  const int grid = (int)grid64;
  // How the accumulated rows reach global memory: vector RED.ADD from the lanes, or bulk reductions
  // of the TMA (UBLKRED).  Measured at C5: equal in fp32 (the engine's read-back wait per quarter
  // costs what the three load-store wavefronts per row save: 0.95 ms both), 3.03 against 3.51 ms in
  // fp64, where a RED carries 8 bytes and a bulk reduction a whole 256-byte row.  AMF_TILED_FLUSH=
  // red / tma overrides.
  const char* fl = getenv("AMF_TILED_FLUSH");
  const bool tma_flush = fl ? !strcmp(fl, "tma") : sizeof(T) == 8;
  AMF_CUDA(cudaMemsetAsync(r->tile_ctr, 0, 4 * (size_t)r->n_tiles, s));   // bundles handed out: none yet
  if (GRAD) AMF_DBG_RANGE(0, dOwn, (size_t)(side == 0 ? h->n_users : h->n_items) * nvec * 16, s);
#define TILED(NVEC_)                                                                              \
  do {                                                                                            \
    const size_t smem = tiled_smem<T, NVEC_>(r->tile_rows);                                       \
    AMF_REQUIRE(smem <= TILED_SMEM_BUDGET, "tiled rating list: tile of %d rows does not fit",     \
                r->tile_rows);                                                                    \
    auto kern = tma_flush ? tiled_side_kernel<T, NVEC_, GRAD, GRAD> : tiled_side_kernel<T, NVEC_, GRAD, false>; \
    AMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, tiled_threads<T, NVEC_>(), smem, s>>>(                                           \
        r->idx, (const T*)r->val, r->rowid, r->seglen, r->binfo, r->tile_bstart, r->tile_ctr,     \
        r->n_tiles, r->n_bundles, r->tile_rows, tile_side_rows, Own, Tile, inv_sigma, mean_offset, dOwn,      \
        sq_err);                                                                                  \
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

// side 0 is cut from the user-major list (own = user, other = item), side 1 from the item-major
template <typename T>
static int build_tiled_side(amf_ratings* h, int side, int tile_rows, cudaStream_t s) {
  runs_free(&h->tiled[side]);
  const int32_t own_rows = side == 0 ? h->n_users : h->n_items;
  const int32_t other_rows = side == 0 ? h->n_items : h->n_users;
  int32_t* own = nullptr;
  AMF_CUDA(cudaMalloc(&own, 4 * (size_t)(h->nnz > 0 ? h->nnz : 1)));
  expand_rows_kernel<int32_t><<<num_sms() * 8, 256, 0, s>>>(h->ptr[side], own_rows, own);
  int rc = cudaGetLastError() == cudaSuccess ? AMF_OK : AMF_ERR_CUDA;
  if (rc == AMF_OK)
    rc = runs_build(&h->tiled[side], h->nnz, own, h->idx[side], h->val[side], (int)sizeof(T),
                    own_rows, other_rows, tile_rows, false, s);
  cudaStreamSynchronize(s);
  cudaFree(own);
  return rc;
}

static bool tiled_row_ok(size_t row_bytes) { return row_bytes == 64 || row_bytes == 128 || row_bytes == 256; }
static void free_tiled_side(amf_runs* t) { runs_free(t); }

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
    const int tile_rows = h->dtype == AMF_F32 ? tiled_max_tile_rows<float>(row_bytes)
                                              : tiled_max_tile_rows<double>(row_bytes);
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
