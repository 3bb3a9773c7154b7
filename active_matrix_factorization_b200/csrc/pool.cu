// Candidate pool handle: the reference's `pool` / `unrated` set (active_pmf.py:725-737) kept on
// the device in the "bundled runs" layout (runs.cuh) built for the scoring kernel.
//
// The plain scoring kernel (scoring.cu) gathers one factor row per candidate from L2 and sits
// at the L2->SM gather ceiling.  Here the pool is bucketed once by item tile: a CTA keeps a
// tile of V in shared memory (TMA bulk copies) and reads every item row from there; the
// candidates one user has inside one tile are a run, ONE LANE owns a run segment and keeps the
// whole user row in registers, 32 equally long segments make the bundle a warp works on.  A
// candidate costs 2 bytes of HBM (its 16-bit row inside the tile) and exactly one shared-memory
// wavefront (the 128-byte item row at d = 32 fp32) -- no shuffles, no divergent user changes.
// The permutation back to the caller's order is kept, so scores and the winner are reported
// exactly as the unbucketed path would.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "peer.cuh"
#include "runs.cuh"

struct amf_pool {
  int64_t ncand;         // candidates given by the caller
  int32_t n_users, n_items;
  amf_runs runs;         // own side = users, tile side = items; orig/pos_of kept
  void* tmp_scores;      // scratch for scores in slot order
  size_t tmp_bytes;
  unsigned int* ticket;  // CTAs of the running launch that have finished (fused winner exchange)
};

namespace amf {

int acquire_partials(Best** out, cudaStream_t s);

constexpr int64_t POOL_BUNDLE_COST = 5;      // row fetch of a bundle, in entry steps (micro_visit.cu)

// threads per CTA (512 / 640 / 768 measured equal at C5: the kernel is bound by the
// load-store pipe, not by latency; fewer warps leave more shared memory to the tile)
constexpr int POOL_THREADS_NARROW = 512, POOL_THREADS_WIDE = 256;
template <int NVEC> constexpr int pool_threads() { return NVEC >= 16 ? POOL_THREADS_WIDE : POOL_THREADS_NARROW; }
// dynamic shared memory: [per-warp staging][tile rows][one NaN row for the padding index]
template <int NVEC, int THREADS = pool_threads<NVEC>()> constexpr size_t pool_stage_total() {
  return (size_t)(THREADS / 32) * runs_stage_bytes<NVEC>();
}
constexpr size_t POOL_SMEM_BUDGET = 227 * 1024 - 1024;   // static buffers of the kernel fit the rest

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

__device__ __forceinline__ void st_scores4(float* p, const float (&v)[4]) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
}
__device__ __forceinline__ void st_scores4(double* p, const double (&v)[4]) {
  __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[1]));
  __stcs(reinterpret_cast<double2*>(p) + 1, make_double2(v[2], v[3]));
}

// what the scoring kernel needs to finish with the cross-GPU winner exchange (peers == NULL: off)
struct PoolPeer {
  unsigned char* const* peers;
  int world, rank;
  unsigned int epoch;
  unsigned int* ticket;
  amf_best_t* out;
};

// PRED over a pool in the bundled-runs layout.  One CTA per SM; it starts at the item tile an
// equal-cost split puts it on, brings the tile into shared memory by TMA, and its warps take
// bundles of that tile from the tile's global counter until the tile is dry, then it moves to
// the next tile with work left (runs_next_tile: CTAs that finish early help where work remains).
// The metadata of the next bundle is fetched while the current one is scored.  A lane scores one candidate per
// step with its user row in registers; the fused arg-best compares once per group of four steps
// and looks up the caller's position only for scores that reach the warp's running best.
template <typename T, int NVEC, bool MAX, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
pool_pred_kernel(const uint16_t* __restrict__ idx, const uint32_t* __restrict__ orig,
                 const uint32_t* __restrict__ rowid, const int2* __restrict__ binfo,
                 const int64_t* __restrict__ tile_bstart, uint32_t* tile_ctr, int n_tiles, int64_t n_bundles,
                 int tile_rows, int n_items, const T* __restrict__ U, const T* __restrict__ Vm,
                 T* __restrict__ scores, int64_t index_base, Best* __restrict__ part, PoolPeer pp) {
  using V = typename Vec<T>::type;
  constexpr uint32_t ROW_BYTES = NVEC * 16;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_v;
  __shared__ int s_next;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t stage = smem_u32(smem_raw) + warp * runs_stage_bytes<NVEC>();
  unsigned char* tile_ptr = smem_raw + pool_stage_total<NVEC, THREADS>();
  const uint32_t tile0 = smem_u32(tile_ptr) | runs_lane_rot<NVEC>(lane);
  const unsigned char* Ub = reinterpret_cast<const unsigned char*>(U);

  const int64_t total = runs_cost(binfo, n_bundles, POOL_BUNDLE_COST);
  const int64_t b_lo = runs_split(binfo, n_bundles, POOL_BUNDLE_COST, total * blockIdx.x / gridDim.x);
  if (threadIdx.x == 0) mbar_init(&bar_v, 1);
  // the row behind the tile: what padding entries read.  NaN scores never win.
  for (int t = threadIdx.x; t < (int)(ROW_BYTES / sizeof(T)); t += THREADS)
    reinterpret_cast<T*>(tile_ptr + (size_t)tile_rows * ROW_BYTES)[t] = T(NAN);
  int t_cur = 0;
  {                                                   // last tile starting at or before b_lo
    int lo = 0, hi = n_tiles;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tile_bstart[mid] <= b_lo) lo = mid; else hi = mid - 1;
    }
    t_cur = lo;
  }
  const T worst = MAX ? -INFINITY : INFINITY;
  T best_v = worst, thr = worst;
  int64_t best_o = -1;
  uint32_t phase_v = 0;

  for (int hop = 0; hop < n_tiles;) {
    const int r_hop = runs_next_tile(tile_ctr, tile_bstart, n_tiles, t_cur + hop, n_tiles - hop, &s_next);
    if (r_hop < 0) break;
    hop += r_hop;
    const int t_now = (t_cur + hop) % n_tiles;
    ++hop;
    const int64_t c = tile_bstart[t_now], seg_end = tile_bstart[t_now + 1];
    if (threadIdx.x == 0) {
      const int rows = min(tile_rows, n_items - t_now * tile_rows);
      const uint32_t bytes = (uint32_t)rows * ROW_BYTES;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bar_v, bytes);
      const unsigned char* src =
          reinterpret_cast<const unsigned char*>(Vm) + (int64_t)t_now * tile_rows * ROW_BYTES;
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
    uint2 w = make_uint2(0u, 0u);
    if (b < seg_end) {
      info = binfo[b]; rid = rowid[b * 32 + lane];
      w = __ldcs(reinterpret_cast<const uint2*>(idx) + (int64_t)info.x * 32 + lane);
    }
    while (b < seg_end) {
      // the next bundle: its metadata now, its first group of indices once the metadata is here
      const int64_t nb = grab();
      int2 ninfo = make_int2(0, 0);
      uint32_t nrid = RUNS_NONE;
      if (nb < seg_end) { ninfo = binfo[nb]; nrid = rowid[nb * 32 + lane]; }
      uint2 w_next = make_uint2(0u, 0u);

      AMF_DBG_ASSERT(rid == RUNS_NONE || rid < 0x7fffffffu);
      V a[NVEC];
      fetch_rows<V, NVEC>(Ub, rid, stage, lane, a);
      const int L = info.y, G = (L + 3) >> 2;
      const uint2* ip = reinterpret_cast<const uint2*>(idx) + (int64_t)info.x * 32 + lane;
#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        uint2 wn = w;
        if (g + 1 < G) wn = __ldcs(ip + (g + 1) * 32);
        else if (nb < seg_end) w_next = __ldcs(reinterpret_cast<const uint2*>(idx) + (int64_t)ninfo.x * 32 + lane);
        const int ns = L - 4 * g;
        const uint32_t j4[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
        T p[4] = {worst, worst, worst, worst};
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (s < ns) {
            AMF_DBG_ASSERT((int)j4[s] <= tile_rows);
            V bb[NVEC];
            lds_row<V, NVEC>(tile0 + j4[s] * ROW_BYTES, bb);
            p[s] = dot_rows<NVEC>(a, bb);
          }
        }
        const int64_t pos = ((int64_t)info.x + g) * 128 + lane * 4;
        if (scores) st_scores4(scores + pos, p);
        // arg-best: `thr` is the best value any lane of this warp holds; only scores that reach
        // it (ties included: a lower index may still win) look up their original position
        T m = p[0];
#pragma unroll
        for (int s = 1; s < 4; ++s) m = MAX ? fmax(m, p[s]) : fmin(m, p[s]);
        const bool reach = MAX ? (m >= thr) : (m <= thr);
        if (__any_sync(0xffffffffu, reach)) {
          if (reach) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              if (MAX ? (p[s] >= thr) : (p[s] <= thr)) {
                const uint32_t ou = orig[pos + s];
                const int64_t o = (int64_t)ou;
                if (ou != RUNS_NONE &&
                    (best_o < 0 || (MAX ? (p[s] > best_v) : (p[s] < best_v)) ||
                     (p[s] == best_v && o < best_o))) {
                  best_v = p[s]; best_o = o;
                }
              }
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
        w = wn;
      }
      b = nb; info = ninfo; rid = nrid; w = w_next;
    }
  }
  Best best{(double)best_v, best_o < 0 ? -1 : best_o + index_base};
  best = block_best<MAX>(best);
  if (threadIdx.x == 0) part[blockIdx.x] = best;
  // Fused tail: the LAST CTA to finish reduces the partial winners of this GPU and, for a sharded
  // pool, exchanges the result with the other GPUs over NVLink peer memory (peer.cuh) -- scoring,
  // arg-best and the cross-GPU "all-gather + chooser" are ONE kernel, nothing is launched after it.
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(pp.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  Best mine{0.0, -1};
  for (unsigned int t = threadIdx.x; t < gridDim.x; t += THREADS) {
    const volatile Best* q = part + t;
    const Best o{q->v, q->i};
    if (better<MAX>(o.v, o.i, mine.v, mine.i)) mine = o;
  }
  mine = block_best<MAX>(mine);                       // valid in thread 0
  if (threadIdx.x < 32) {
    mine.v = __shfl_sync(0xffffffffu, mine.v, 0);
    mine.i = __shfl_sync(0xffffffffu, mine.i, 0);
    Best all = mine;
    if (all.v != all.v) all.i = -1;                   // NaN never wins
    if (pp.peers)
      all = peer_exchange_warp<MAX>(pp.peers, pp.world, pp.rank, pp.epoch, mine, (int)threadIdx.x);
    if (threadIdx.x == 0) {
      pp.out->value = all.i < 0 ? 0.0 : all.v;
      pp.out->index = all.i;
      *pp.ticket = 0;                                 // ready for the next launch on this pool
    }
  }
}

template <typename T>
__global__ void unpermute_kernel(const T* __restrict__ in, const uint32_t* __restrict__ orig,
                                 int64_t npad, T* __restrict__ out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < npad;
       t += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t o = orig[t];
    if (o != RUNS_NONE) out[o] = in[t];
  }
}

template <int NVEC, int THREADS> static size_t pool_smem(int tile_rows) {
  return pool_stage_total<NVEC, THREADS>() + ((size_t)tile_rows + 1) * NVEC * 16;
}
// AMF_POOL_THREADS = 512 / 640 / 768: tuning knob of benchmarks/pool_variants.py (narrow rows only)
static int pool_threads_narrow() {
  const char* e = getenv("AMF_POOL_THREADS");
  const int t = e ? atoi(e) : POOL_THREADS_NARROW;
  return (t == 256 || t == 384 || t == 448 || t == 576 || t == 640 || t == 768) ? t : POOL_THREADS_NARROW;
}

template <typename T, bool MAX>
static int pool_pred(const amf_pool* h, int ld, const T* U, const T* V, T* scores_tmp,
                     int64_t index_base, Best* part, int grid, cudaStream_t s, PoolPeer pp) {
  constexpr int N = Vec<T>::N;
  const int nvec = ld / N;
  const amf_runs* r = &h->runs;
  AMF_CUDA(cudaMemsetAsync(r->tile_ctr, 0, 4 * (size_t)r->n_tiles, s));   // bundles handed out: none yet
#define POOL_T(NVEC_, THREADS_)                                                                 \
  do {                                                                                          \
    const size_t smem = pool_smem<NVEC_, THREADS_>(r->tile_rows);                               \
    AMF_REQUIRE(smem <= POOL_SMEM_BUDGET, "item tile (%d rows of %d) does not fit shared "      \
                "memory: build the pool with amf_pool_max_tile_rows", r->tile_rows, ld);        \
    AMF_CUDA(cudaFuncSetAttribute(pool_pred_kernel<T, NVEC_, MAX, THREADS_>,                    \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    pool_pred_kernel<T, NVEC_, MAX, THREADS_><<<grid, THREADS_, smem, s>>>(                     \
        r->idx, r->orig, r->rowid, r->binfo, r->tile_bstart, r->tile_ctr, r->n_tiles, r->n_bundles,          \
        r->tile_rows, h->n_items, U, V, scores_tmp, index_base, part, pp);                      \
  } while (0)
#define POOL(NVEC_) POOL_T(NVEC_, POOL_THREADS_NARROW)
  const int narrow = pool_threads_narrow();
  switch (nvec) {            // the row width is a compile-time constant of the kernel
    case 1: POOL(1); break;
    case 2: POOL(2); break;
    case 4: POOL(4); break;
    case 8:
      if (narrow == 768) POOL_T(8, 768); else if (narrow == 640) POOL_T(8, 640);
      else if (narrow == 384) POOL_T(8, 384); else if (narrow == 256) POOL_T(8, 256);
      else if (narrow == 448) POOL_T(8, 448); else if (narrow == 576) POOL_T(8, 576); else POOL(8);
      break;
    case 16: POOL_T(16, POOL_THREADS_WIDE); break;
    default:
      set_error("bucketed pool needs a padded row of 1, 2, 4, 8 or 16 16-byte vectors (ld=%d)", ld);
      return AMF_ERR_UNSUPPORTED;
  }
#undef POOL
#undef POOL_T
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int amf_pool_max_tile_rows(int row_bytes) {
  size_t stage = 0;
  switch (row_bytes) {
    case 16: stage = pool_stage_total<1>(); break;
    case 32: stage = pool_stage_total<2>(); break;
    case 64: stage = pool_stage_total<4>(); break;
    case 128: stage = (size_t)(pool_threads_narrow() / 32) * runs_stage_bytes<8>(); break;
    case 256: stage = pool_stage_total<16>(); break;
    default: return 0;
  }
  const size_t rows = (POOL_SMEM_BUDGET - stage) / (size_t)row_bytes - 1;
  return (int)(rows > 65535 ? 65535 : rows);
}

int amf_pool_destroy(amf_pool_t* h) {
  if (!h) return AMF_OK;
  runs_free(&h->runs);
  cudaFree(h->tmp_scores);
  cudaFree(h->ticket);
  delete h;
  return AMF_OK;
}

int amf_pool_create(amf_pool_t** out, int64_t ncand, const int32_t* ci_d, const int32_t* cj_d,
                    int32_t n_users, int32_t n_items, int tile_rows, void* stream) {
  AMF_REQUIRE(out && n_users > 0 && n_items > 0, "amf_pool_create: bad arguments");
  AMF_REQUIRE(ncand >= 0 && ncand < (1ll << 31), "amf_pool_create: ncand out of range");
  AMF_REQUIRE(tile_rows >= 1 && tile_rows <= 65535, "amf_pool_create: tile_rows must be in [1, 65535]");
  cudaStream_t s = (cudaStream_t)stream;
  AMF_REQUIRE(ncand == 0 || (ci_d && cj_d), "amf_pool_create: NULL candidate arrays");
  AMF_CHECK_ID_RANGE("amf_pool_create", ci_d, n_users, cj_d, n_items, ncand, s);
  amf_pool* h = new amf_pool();
  memset(h, 0, sizeof(*h));
  h->ncand = ncand; h->n_users = n_users; h->n_items = n_items;
  const int rc = runs_build(&h->runs, ncand, ci_d, cj_d, nullptr, 0, n_users, n_items,
                            tile_rows < n_items ? tile_rows : n_items, true, s);
  if (rc != AMF_OK) { delete h; return rc; }
  if (cudaMalloc(&h->ticket, sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(h->ticket, 0, sizeof(unsigned int)) != cudaSuccess) {
    set_error("amf_pool_create: %s", cudaGetErrorString(cudaGetLastError()));
    amf_pool_destroy(h);
    return AMF_ERR_CUDA;
  }
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
    if (c >= 0 && c < ncand) orig[pos_of[c]] = amf::RUNS_NONE;
  }
}

int amf_pool_remove(amf_pool_t* h, int64_t n, const int64_t* idx_d, void* stream) {
  AMF_REQUIRE(h && (n == 0 || idx_d), "amf_pool_remove: NULL argument");
  if (n <= 0 || h->ncand == 0) return AMF_OK;
  const int grid = (int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  pool_remove_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(idx_d, n, h->ncand, h->runs.pos_of,
                                                             h->runs.orig);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

static int pool_score_pred(const amf_pool_t* hc, int dtype, int d, int ld, const void* U_d,
                           const void* V_d, void* scores_d, int maximize, int64_t index_base,
                           amf_peer_t* peer, amf_best_t* best_d, void* stream) {
  AMF_REQUIRE(hc && U_d && V_d && best_d, "amf_pool_score_pred: NULL argument");
  AMF_REQUIRE(!peer || peer->connected, "amf_pool_score_pred_peer: amf_peer_connect has not been called");
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_pool_score_pred: bad dtype");
  amf_pool* h = const_cast<amf_pool*>(hc);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t es = dtype == AMF_F32 ? 4 : 8;
  const int vecn = dtype == AMF_F32 ? 4 : 2;
  AMF_REQUIRE(ld >= d && ld % vecn == 0, "ld=%d must be >= d=%d and a multiple of %d", ld, d, vecn);
  AMF_REQUIRE(((uintptr_t)U_d | (uintptr_t)V_d) % 16 == 0,
              "amf_pool_score_pred: U and V must be 16-byte aligned");
  const amf_runs* r = &h->runs;
  Best* part = nullptr;
  int rc = acquire_partials(&part, s);
  if (rc != AMF_OK) return rc;
  struct PartGuard {                       // released on every early return
    Best* p; cudaStream_t s; bool armed;
    ~PartGuard() { if (armed) cudaFreeAsync(p, s); }
  } guard{part, s, true};
  void* tmp_scores = nullptr;
  if (scores_d && r->npos > 0) {
    if (h->tmp_bytes < es * (size_t)r->npos) {
      cudaFree(h->tmp_scores); h->tmp_scores = nullptr; h->tmp_bytes = 0;
      AMF_CUDA(cudaMalloc(&h->tmp_scores, es * (size_t)r->npos));
      h->tmp_bytes = es * (size_t)r->npos;
    }
    tmp_scores = h->tmp_scores;
  }
  int64_t grid64 = (int64_t)num_sms();
  if (grid64 > r->n_bundles) grid64 = r->n_bundles > 0 ? r->n_bundles : 1;
  const int grid = (int)grid64;
  PoolPeer pp{nullptr, 1, 0, 0u, h->ticket, best_d};
  if (peer)       // a collective: every rank calls in the same order, so the epochs agree
    pp = PoolPeer{peer->peers_d, peer->world, peer->rank, ++peer->epoch, h->ticket, best_d};
  if (dtype == AMF_F32)
    rc = maximize ? pool_pred<float, true>(h, ld, (const float*)U_d, (const float*)V_d, (float*)tmp_scores, index_base, part, grid, s, pp)
                  : pool_pred<float, false>(h, ld, (const float*)U_d, (const float*)V_d, (float*)tmp_scores, index_base, part, grid, s, pp);
  else
    rc = maximize ? pool_pred<double, true>(h, ld, (const double*)U_d, (const double*)V_d, (double*)tmp_scores, index_base, part, grid, s, pp)
                  : pool_pred<double, false>(h, ld, (const double*)U_d, (const double*)V_d, (double*)tmp_scores, index_base, part, grid, s, pp);
  if (rc != AMF_OK) return rc;
  if (tmp_scores) {
    const int g2 = num_sms() * 8;
    if (dtype == AMF_F32) unpermute_kernel<float><<<g2, 256, 0, s>>>((const float*)tmp_scores, r->orig, r->npos, (float*)scores_d);
    else unpermute_kernel<double><<<g2, 256, 0, s>>>((const double*)tmp_scores, r->orig, r->npos, (double*)scores_d);
    AMF_LAUNCH_CHECK();
  }
  return AMF_OK;                           // the scoring kernel's last CTA wrote best_d; the guard frees the partials
}

int amf_pool_score_pred(const amf_pool_t* hc, int dtype, int d, int ld, const void* U_d,
                        const void* V_d, void* scores_d, int maximize, int64_t index_base,
                        amf_best_t* best_d, void* stream) {
  return pool_score_pred(hc, dtype, d, ld, U_d, V_d, scores_d, maximize, index_base, nullptr, best_d, stream);
}

int amf_pool_score_pred_peer(const amf_pool_t* hc, int dtype, int d, int ld, const void* U_d,
                             const void* V_d, void* scores_d, int maximize, int64_t index_base,
                             amf_peer_t* peer, amf_best_t* best_d, void* stream) {
  AMF_REQUIRE(peer, "amf_pool_score_pred_peer: NULL peer");
  return pool_score_pred(hc, dtype, d, ld, U_d, V_d, scores_d, maximize, index_base, peer, best_d, stream);
}

#pragma GCC visibility pop
}  // namespace
