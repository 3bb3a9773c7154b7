// Sample statistics of a DENSE `which` (every cell of the matrix) on the 5th-generation tensor
// cores: predict / pred_variance over S posterior samples (bayes_pmf.py:433-448) when the pool
// is all unknown cells (bayes_pmf.py:702-712), i.e. S small GEMMs P_s = U_s V_s^T with a
// running-moment epilogue.
//
//   * operands: fp32 split into TF32 hi + lo parts by a pre-pass (x = hi + lo exactly, hi has the
//     13 low mantissa bits cleared); P_s = hi.hi + hi.lo + lo.hi (3xTF32: fp32-class accuracy,
//     the dropped lo.lo term is 2^-22 relative);
//   * the variance is accumulated about the FIRST sample's prediction without keeping it in
//     registers: every accumulator tile starts as  -U_0 V_0^T  (the same three products issued
//     with the A-negate bit of the instruction descriptor) and then receives  +U_s V_s^T, so
//     TMEM holds x = P_s - P_0 and the epilogue only does  s1 += x, s2 += x*x;
//   * one CTA per 128 x 128 tile of cells: warp 0 = TMA producer (2-D tensor maps, UTMALDG),
//     warp 1 = MMA issuer (tcgen05.mma kind::tf32, cta_group::1, accumulators in TMEM, two
//     buffers of 128 columns), warps 2-9 = epilogue (tcgen05.ld 32x32b, 64 columns per thread);
//   * shared-memory operand layout: K-major, no swizzle -- core matrices of 8 rows x 16 bytes;
//     a "chunk column" (all 128 rows of one 16-byte K chunk) is 2 KB contiguous, which is exactly
//     what one row of the pre-pass output holds, so a whole operand tile (8 chunk columns: 4 hi,
//     4 lo at K = 16) is ONE 2-D TMA box {2 KB, chunks}.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "tile_stream.cuh"

namespace amf {

int acquire_partials(Best** out, cudaStream_t s);

namespace {

constexpr int TC_TILE = 128;           // cells per side of a CTA tile = UMMA M = UMMA N
// operand ring (samples s+1.. in flight while s is multiplied); K = 32 leaves room for two stages
template <int KP> struct TcStages { static constexpr int value = KP >= 32 ? 2 : 3; };
constexpr int TC_THREADS = 320;        // 1 TMA warp + 1 MMA warp + 8 epilogue warps
constexpr int TC_CHUNK_BYTES = TC_TILE * 16;   // one chunk column: 128 rows x 16 bytes

// x -> (hi, lo) TF32 parts in the chunk-column layout:
//   out[((s * chunks + kc) * rows_pad + row) * 4 + e]
// kc < kp/4: hi part of columns 4kc..4kc+3; kc >= kp/4: lo part.  Rows >= rows and columns >= d
// are zero.
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ X, int S, int rows, int d, int kp, int rows_pad,
                  float* __restrict__ out) {
  const int kq = kp / 4, chunks = 2 * kq;
  const int64_t total = (int64_t)S * chunks * rows_pad;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(t % rows_pad);
    const int kc = (int)((t / rows_pad) % chunks);
    const int s = (int)(t / ((int64_t)rows_pad * chunks));
    const bool lo = kc >= kq;
    const int c0 = 4 * (lo ? kc - kq : kc);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows) {
      float e[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float x = c0 + q < d ? X[((int64_t)s * rows + row) * d + c0 + q] : 0.f;
        const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
        e[q] = lo ? x - hi : hi;
      }
      v = make_float4(e[0], e[1], e[2], e[3]);
    }
    reinterpret_cast<float4*>(out)[t] = v;
  }
}

// ---- PTX wrappers ----------------------------------------------------------------------------
// mbarrier wait with a watchdog: a protocol error traps (an error code for the caller) instead
// of spinning for ever
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spins = 0;; ++spins) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, both K-major, TF32 in, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, no swizzle: 8-row groups 128 bytes apart (SBO), 16-byte K chunks one chunk column
// apart (LBO); descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(TC_CHUNK_BYTES >> 4) << 16) |
         ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = N = 128; bit 13 negates A
__device__ __forceinline__ constexpr uint32_t umma_idesc(bool negate_a) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((negate_a ? 1u : 0u) << 13) |
         ((uint32_t)(TC_TILE >> 3) << 17) | ((uint32_t)(TC_TILE >> 4) << 24);
}

template <int KP, bool MAX>
__global__ void __launch_bounds__(TC_THREADS, 1)
sample_stats_tc_kernel(const __grid_constant__ CUtensorMap map_u,
                       const __grid_constant__ CUtensorMap map_v, int S, int n, int m,
                       float offset, float* __restrict__ mean_out, float* __restrict__ var_out,
                       int select, int64_t index_base, Best* __restrict__ part) {
  constexpr int CHUNKS = KP / 2;                       // hi + lo chunk columns of one operand
  constexpr uint32_t OP_BYTES = CHUNKS * TC_CHUNK_BYTES;
  constexpr uint32_t STAGE_BYTES = 2 * OP_BYTES;       // U tile + V tile of one sample
  constexpr int KSTEPS = KP / 8;
  constexpr int TC_STAGES = TcStages<KP>::value;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_base, bar_full[TC_STAGES], bar_empty[TC_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_v = (m + TC_TILE - 1) / TC_TILE;
  const int row0 = (blockIdx.x / tiles_v) * TC_TILE, col0 = (blockIdx.x % tiles_v) * TC_TILE;
  unsigned char* base_ops = smem;                      // sample 0: the shift
  unsigned char* ring = smem + STAGE_BYTES;

  if (threadIdx.x == 0) {
    mbar_init(&bar_base, 1);
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_tfull[i], 1); mbar_init(&bar_tempty[i], 8 * 32); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_slot)), "r"(2 * TC_TILE) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(&bar_base, STAGE_BYTES);
      tma_load_2d(base_ops, &map_u, row0 * 2, 0, &bar_base);
      tma_load_2d(base_ops + OP_BYTES, &map_v, col0 * 2, 0, &bar_base);
      for (int s = 1; s < S; ++s) {
        const int it = s - 1, st = it % TC_STAGES;
        mbar_wait_wd(&bar_empty[st], ((it / TC_STAGES) & 1) ^ 1);
        mbar_expect_tx(&bar_full[st], STAGE_BYTES);
        unsigned char* dst = ring + (size_t)st * STAGE_BYTES;
        tma_load_2d(dst, &map_u, row0 * 2, s * CHUNKS, &bar_full[st]);
        tma_load_2d(dst + OP_BYTES, &map_v, col0 * 2, s * CHUNKS, &bar_full[st]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t IPOS = umma_idesc(false), INEG = umma_idesc(true);
      const uint32_t a0 = smem_u32(base_ops), b0 = a0 + OP_BYTES;
      mbar_wait_wd(&bar_base, 0);
      for (int it = 0; it < S; ++it) {                 // it < S-1: sample it+1 minus sample 0
        const int buf = it & 1;
        mbar_wait_wd(&bar_tempty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * TC_TILE;
        const bool last = it == S - 1;                 // the plain P_0 tile closes the run
        uint32_t acc = 0;
        // hi.hi + hi.lo + lo.hi of sample 0, negated unless this is the P_0 tile
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint32_t hi = (2 * ks) * TC_CHUNK_BYTES, lo = (KP / 4 + 2 * ks) * TC_CHUNK_BYTES;
          umma_tf32(d_tmem, umma_desc(a0 + hi), umma_desc(b0 + hi), last ? IPOS : INEG, acc); acc = 1;
          umma_tf32(d_tmem, umma_desc(a0 + hi), umma_desc(b0 + lo), last ? IPOS : INEG, 1);
          umma_tf32(d_tmem, umma_desc(a0 + lo), umma_desc(b0 + hi), last ? IPOS : INEG, 1);
        }
        if (!last) {
          const int st = it % TC_STAGES;
          mbar_wait_wd(&bar_full[st], (it / TC_STAGES) & 1);
          tc_fence_after();
          const uint32_t a1 = smem_u32(ring + (size_t)st * STAGE_BYTES), b1 = a1 + OP_BYTES;
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) {
            const uint32_t hi = (2 * ks) * TC_CHUNK_BYTES, lo = (KP / 4 + 2 * ks) * TC_CHUNK_BYTES;
            umma_tf32(d_tmem, umma_desc(a1 + hi), umma_desc(b1 + hi), IPOS, 1);
            umma_tf32(d_tmem, umma_desc(a1 + hi), umma_desc(b1 + lo), IPOS, 1);
            umma_tf32(d_tmem, umma_desc(a1 + lo), umma_desc(b1 + hi), IPOS, 1);
          }
          umma_commit(&bar_empty[st]);                 // operands of this sample may be overwritten
        }
        umma_commit(&bar_tfull[buf]);                  // accumulator tile complete
      }
    }
  } else {
    // ===== epilogue: 8 warps, TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const int col_off = half * 64;
    float s1[64], s2[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
    for (int it = 0; it < S - 1; ++it) {
      const int buf = it & 1;
      mbar_wait_wd(&bar_tfull[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + t_lane + buf * TC_TILE + col_off;
      float x[32];
      tmem_ld32(taddr, x);
#pragma unroll
      for (int c = 0; c < 32; ++c) { s1[c] += x[c]; s2[c] = fmaf(x[c], x[c], s2[c]); }
      tmem_ld32(taddr + 32, x);
      tc_fence_before();
      mbar_arrive(&bar_tempty[buf]);                   // both halves of this thread's row are read
#pragma unroll
      for (int c = 0; c < 32; ++c) { s1[32 + c] += x[c]; s2[32 + c] = fmaf(x[c], x[c], s2[32 + c]); }
    }
    // the P_0 tile: mean = P_0 + s1/S + offset, var = s2/S - (s1/S)^2 (population, np.var)
    const int it = S - 1, buf = it & 1;
    mbar_wait_wd(&bar_tfull[buf], (it >> 1) & 1);
    tc_fence_after();
    const uint32_t taddr = tmem_base + t_lane + buf * TC_TILE + col_off;
    const int row = row0 + q * 32 + lane;
    const float inv_s = 1.f / (float)S;
    Best best{0.0, -1};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float p0[32];
      tmem_ld32(taddr + 32 * h, p0);
      if (row < n) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int col = col0 + col_off + 32 * h + c;
          if (col < m) {
            const float mu = s1[32 * h + c] * inv_s;
            const float mean = p0[c] + mu + offset;
            const float var = fmaxf(fmaf(-mu, mu, s2[32 * h + c] * inv_s), 0.f);
            const int64_t cell = (int64_t)row * m + col;
            if (mean_out) mean_out[cell] = mean;
            if (var_out) var_out[cell] = var;
            const double sel = select == 0 ? (double)mean : (double)var;
            if (better<MAX>(sel, cell + index_base, best.v, best.i)) { best.v = sel; best.i = cell + index_base; }
          }
        }
      }
    }
    tc_fence_before();
    // per-warp winners to shared memory through the generic best reduction below
    best = warp_best<MAX>(best);
    __shared__ double sv[8];
    __shared__ long long si[8];
    if (lane == 0) { sv[warp - 2] = best.v; si[warp - 2] = best.i; }
    asm volatile("bar.sync 1, 256;" ::: "memory");      // the eight epilogue warps only
    if (warp == 2 && lane == 0) {
      Best b{sv[0], si[0]};
      for (int w = 1; w < 8; ++w) if (better<MAX>(sv[w], si[w], b.v, b.i)) { b.v = sv[w]; b.i = si[w]; }
      part[blockIdx.x] = b;
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * TC_TILE) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// the pre-pass output as a 2-D tensor of 8-byte elements: dim 0 = rows_pad * 2 (one chunk
// column), dim 1 = S * chunks; box = {256 elements = 2 KB = 128 rows, chunks}
int make_map(CUtensorMap* map, void* base, int S, int chunks, int rows_pad) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return AMF_ERR_UNSUPPORTED; }
  const cuuint64_t gdim[2] = {(cuuint64_t)rows_pad * 2, (cuuint64_t)S * chunks};
  const cuuint64_t gstride[1] = {(cuuint64_t)rows_pad * 16};
  const cuuint32_t box[2] = {256, (cuuint32_t)chunks};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return AMF_ERR_CUDA; }
  return AMF_OK;
}

}  // namespace

// whether the tensor-core form applies; used by amf_bayes_sample_stats to route dense calls
bool dense_tc_applicable(int dtype, int S, int d, const void* prob_d, int select) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("AMF_B200_DENSE_TC");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return enabled && dtype == AMF_F32 && S >= 2 && d >= 1 && d <= 32 && !prob_d && select != 2 &&
         encode_tiled() != nullptr;
}

int dense_tc_launch(int S, int32_t n, int32_t m, int d, const float* Us, const float* Vs,
                    float offset, float* mean_d, float* var_d, int select, int maximize,
                    int64_t index_base, amf_best_t* best_d, cudaStream_t s) {
  const int kp = (d + 7) / 8 * 8, chunks = kp / 2;
  const int n_pad = (n + TC_TILE - 1) / TC_TILE * TC_TILE, m_pad = (m + TC_TILE - 1) / TC_TILE * TC_TILE;
  float *su = nullptr, *sv = nullptr;
  const size_t bu = (size_t)S * chunks * n_pad * 16, bv = (size_t)S * chunks * m_pad * 16;
  AMF_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&su), bu, s));
  if (cudaMallocAsync(reinterpret_cast<void**>(&sv), bv, s) != cudaSuccess) {
    cudaFreeAsync(su, s);
    set_error("dense_tc: out of device memory for the split operands");
    return AMF_ERR_CUDA;
  }
  struct Free { float *a, *b; Best* p; cudaStream_t s; ~Free() { if (a) cudaFreeAsync(a, s); if (b) cudaFreeAsync(b, s); if (p) cudaFreeAsync(p, s); } } guard{su, sv, nullptr, s};
  const int sgrid = num_sms() * 8;
  split_tf32_kernel<<<sgrid, 256, 0, s>>>(Us, S, n, d, kp, n_pad, su);
  split_tf32_kernel<<<sgrid, 256, 0, s>>>(Vs, S, m, d, kp, m_pad, sv);
  AMF_LAUNCH_CHECK();
  CUtensorMap mu, mv;
  int rc = make_map(&mu, su, S, chunks, n_pad);
  if (rc == AMF_OK) rc = make_map(&mv, sv, S, chunks, m_pad);
  if (rc != AMF_OK) return rc;
  Best* part = nullptr;
  rc = acquire_partials(&part, s);
  if (rc != AMF_OK) return rc;
  guard.p = part;
  const int grid = (n_pad / TC_TILE) * (m_pad / TC_TILE);
  AMF_REQUIRE(grid <= 8192, "dense_tc: %d tiles exceed the winner scratch", grid);
  const size_t smem = (size_t)(1 + (kp >= 32 ? 2 : 3)) * 2 * chunks * TC_CHUNK_BYTES;
  int dev = 0;
  AMF_CUDA(cudaGetDevice(&dev));
#define TC(KP_, MAX_)                                                                             \
  do {                                                                                            \
    static bool attr_set[64] = {false};   /* per instantiation and device; the value is fixed */ \
    if (dev >= 64 || !attr_set[dev]) {                                                            \
      AMF_CUDA(cudaFuncSetAttribute(sample_stats_tc_kernel<KP_, MAX_>,                            \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
      if (dev < 64) attr_set[dev] = true;                                                         \
    }                                                                                             \
    sample_stats_tc_kernel<KP_, MAX_><<<grid, TC_THREADS, smem, s>>>(                             \
        mu, mv, S, n, m, offset, mean_d, var_d, select, index_base, part);                        \
  } while (0)
#define TC_K(KP_) do { if (maximize) TC(KP_, true); else TC(KP_, false); } while (0)
  switch (kp) {
    case 8: TC_K(8); break;
    case 16: TC_K(16); break;
    case 24: TC_K(24); break;
    default: TC_K(32); break;
  }
#undef TC_K
#undef TC
  AMF_LAUNCH_CHECK();
  if (best_d) {
    guard.p = nullptr;                                  // released by launch_best_final
    return launch_best_final(part, grid, maximize != 0, best_d, s);
  }
  return AMF_OK;
}

}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int amf_bayes_sample_stats_dense_tc(int S, int32_t n, int32_t m, int d, const float* Us_d,
                                    const float* Vs_d, double mean_offset, float* mean_d,
                                    float* var_d, int select, int maximize, int64_t index_base,
                                    amf_best_t* best_d, void* stream) {
  AMF_REQUIRE(S >= 2 && n >= 1 && m >= 1 && d >= 1 && d <= 32, "amf_bayes_sample_stats_dense_tc: bad sizes");
  AMF_REQUIRE(Us_d && Vs_d && (select == 0 || select == 1), "amf_bayes_sample_stats_dense_tc: bad arguments");
  return dense_tc_launch(S, n, m, d, Us_d, Vs_d, (float)mean_offset, mean_d, var_d, select, maximize,
                         index_base, best_d, (cudaStream_t)stream);
}

#pragma GCC visibility pop
}
