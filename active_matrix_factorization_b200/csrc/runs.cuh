// "Bundled runs": the layout both tile-resident kernels stream (pool.cu: candidate scoring,
// tiled.cu: the PMF loss + gradient), and the device pieces they share.
//
// A sparse list of (own row, other row [, value]) is bucketed by TILE of the other side's matrix
// (tile_rows rows: what fits shared memory) and sorted by own row inside a tile.  The entries one
// own row has inside one tile are a RUN; runs are cut into SEGMENTS of at most RUNS_MAXLEN
// entries, the segments of a tile are sorted by length (longest first) and taken 32 at a time:
// a BUNDLE is the work of one warp, ONE LANE PER SEGMENT.  The lane keeps the whole own row (and,
// for the gradient, the whole accumulator) in registers for the length of its segment and reads
// one whole tile row per entry from shared memory -- eight LDS.128 whose 16-byte slices are
// visited in an order XOR-rotated by the lane, so that every quarter-warp phase touches eight
// different bank groups whatever rows the lanes read: one shared-memory wavefront per entry,
// which is the floor of any CUDA-core formulation (benchmarks/micro_visit.cu: 1.03 clk per entry
// and SM), with no shuffles, no index broadcast and no divergent row changes.  Sorting by length
// makes the 32 segments of a bundle equally long (padding < 2 % at C5), and the own rows of a
// bundle are fetched / flushed together through a small per-warp staging buffer with coalesced
// global accesses (fetch_rows / flush_rows below).
//
// Memory order of a bundle: L = its longest segment, G = ceil(L / 4) GROUPS; group g holds
// steps 4g .. 4g+3 of all 32 lanes as  [lane][4]: a lane reads its four 16-bit local indices
// as one 8-byte word (and its four values as 16 / 32 bytes), a warp reads 256 contiguous bytes.
//   position(bundle, lane, step) = (first_group(bundle) + step / 4) * 128 + lane * 4 + step % 4
// Padding entries (lanes whose segment is shorter than L, steps past L in the last group) carry
// the local index `tile_rows`: the kernels keep one extra row behind the tile (NaN for scoring:
// never wins; zero for the gradient, where the lane's own length masks the residual).
#pragma once
#include "common.cuh"
#include "tile_stream.cuh"

namespace amf {

constexpr int RUNS_MAXLEN = 64;                    // entries per segment (8 groups)
constexpr uint32_t RUNS_NONE = 0xffffffffu;        // rowid of an empty lane / orig of padding

// per-warp staging: eight rows at a time
template <int NVEC> constexpr uint32_t runs_stage_bytes() { return 8u * NVEC * 16u; }

// which 16-byte slice of a row a lane keeps in register t, and its byte offset
template <int NVEC>
__device__ __forceinline__ uint32_t runs_lane_rot(int lane) {
  constexpr int M = NVEC >= 8 ? 7 : NVEC - 1;
  return (uint32_t)(lane & M) << 4;
}

// the lane's whole row of a tile in shared memory (row_addr = shared address of the row + the
// lane's rotation): slice t comes from  row_addr ^ 16 t
template <typename V, int NVEC>
__device__ __forceinline__ void lds_row(uint32_t row_addr, V (&b)[NVEC]) {
#pragma unroll
  for (int t = 0; t < NVEC; ++t) b[t] = lds_v(row_addr ^ (uint32_t)(t << 4), V());
}

__device__ __forceinline__ void sts_v(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void sts_v(uint32_t addr, const double2& v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

// Rows `rid` (one per lane, RUNS_NONE = none) of `table` into the lanes' registers, slice order
// rotated as lds_row expects.  Eight rows at a time: coalesced 16-byte global loads (NVEC lanes per
// row) into the warp's staging buffer, then every lane of that quarter reads its own row back.
template <typename V, int NVEC>
__device__ __forceinline__ void fetch_rows(const unsigned char* __restrict__ table, uint32_t rid,
                                           uint32_t stage, int lane, V (&a)[NVEC]) {
  constexpr uint32_t ROW_BYTES = NVEC * 16;
  constexpr int LPR = NVEC;                               // lanes per row
  constexpr int RPI = 32 / LPR > 8 ? 8 : 32 / LPR;        // rows per load instruction
  const uint32_t rr = rid == RUNS_NONE ? 0u : rid;
  const int sub = lane / LPR, ch = lane % LPR;
  const uint32_t rot = runs_lane_rot<NVEC>(lane);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int k = 0; k < 8 / RPI; ++k) {
      const int r_local = k * RPI + sub;
      const uint32_t src = __shfl_sync(0xffffffffu, rr, 8 * q + (r_local & 7));
      if (LPR * RPI == 32 || sub < RPI) {
        const V v = ldg_v(table + (uint64_t)src * ROW_BYTES + ch * 16, V());
        sts_v(stage + r_local * ROW_BYTES + ch * 16, v);
      }
    }
    __syncwarp();
    if ((lane >> 3) == q) lds_row<V, NVEC>((stage + (lane & 7) * ROW_BYTES) | rot, a);
    __syncwarp();
  }
}

__device__ __forceinline__ void red_add_v(float* p, const float4& v) {
  AMF_DBG_WRITE(p, 16);
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add_v(double* p, const double2& v) {
  AMF_DBG_WRITE(p, 16);
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p), "d"(v.x) : "memory");
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p + 1), "d"(v.y) : "memory");
}

// The reverse: the lanes' accumulated rows (rotated slice order) are added into rows `rid` of
// `table` with coalesced vector REDs, eight rows at a time through the staging buffer.
template <typename T, typename V, int NVEC>
__device__ __forceinline__ void flush_rows(unsigned char* __restrict__ table, uint32_t rid,
                                           uint32_t stage, int lane, const V (&acc)[NVEC]) {
  constexpr uint32_t ROW_BYTES = NVEC * 16;
  constexpr int LPR = NVEC;
  constexpr int RPI = 32 / LPR > 8 ? 8 : 32 / LPR;
  const int sub = lane / LPR, ch = lane % LPR;
  const uint32_t rot = runs_lane_rot<NVEC>(lane);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if ((lane >> 3) == q) {
      const uint32_t base = (stage + (lane & 7) * ROW_BYTES) | rot;
#pragma unroll
      for (int t = 0; t < NVEC; ++t) sts_v(base ^ (uint32_t)(t << 4), acc[t]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8 / RPI; ++k) {
      const int r_local = k * RPI + sub;
      const uint32_t dst = __shfl_sync(0xffffffffu, rid, 8 * q + (r_local & 7));
      if ((LPR * RPI == 32 || sub < RPI) && dst != RUNS_NONE) {
        const V v = lds_v(stage + r_local * ROW_BYTES + ch * 16, V());
        red_add_v(reinterpret_cast<T*>(table + (uint64_t)dst * ROW_BYTES + ch * 16), v);
      }
    }
    __syncwarp();
  }
}

// Same through the TMA: every lane stores its accumulated row into its staging slot and issues
// one bulk reduction (cp.reduce.async.bulk .add, UBLKRED) of that row into global memory -- the
// adds are done by the copy engine / L2, the load-store pipe only sees the eight STS.  The slot
// is free again once the engine has read it (wait_group.read), which the issuing lanes wait for
// before the next quarter overwrites the buffer.
__device__ __forceinline__ void bulk_red_add(float* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
               "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_red_add(double* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst),
               "r"(src_smem), "r"(bytes)
               : "memory");
}
template <typename T, typename V, int NVEC>
__device__ __forceinline__ void flush_rows_tma(unsigned char* __restrict__ table, uint32_t rid,
                                               uint32_t stage, int lane, const V (&acc)[NVEC]) {
  constexpr uint32_t ROW_BYTES = NVEC * 16;
  const uint32_t rot = runs_lane_rot<NVEC>(lane);
  const uint32_t slot = stage + (lane & 7) * ROW_BYTES;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if ((lane >> 3) == q) {
#pragma unroll
      for (int t = 0; t < NVEC; ++t) sts_v((slot | rot) ^ (uint32_t)(t << 4), acc[t]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (rid != RUNS_NONE) {
        AMF_DBG_WRITE(table + (uint64_t)rid * ROW_BYTES, ROW_BYTES);
        bulk_red_add(reinterpret_cast<T*>(table + (uint64_t)rid * ROW_BYTES), slot, ROW_BYTES);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
}

// first bundle whose cost prefix reaches `target`; cost(b) = 4 * first_group(b) + c0 * b
// (entries streamed + a fixed price per bundle for the row fetch / flush)
__device__ __forceinline__ int64_t runs_cost(const int2* __restrict__ binfo, int64_t b, int64_t c0) {
  return 4 * (int64_t)binfo[b].x + c0 * b;
}
__device__ __forceinline__ int64_t runs_split(const int2* __restrict__ binfo, int64_t n_bundles,
                                              int64_t c0, int64_t target) {
  int64_t lo = 0, hi = n_bundles;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (runs_cost(binfo, mid, c0) < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Dynamic work distribution of the tile kernels.  A CTA starts at its HOME tile (where an equal-cost
// static split would put it) and its warps take bundles of that tile from the tile's GLOBAL
// counter, which all CTAs on the tile share; when the tile runs dry the CTA moves on, cyclically,
// to the next tile that still has bundles nobody took -- so CTAs that finish early help on the
// tiles that are behind, at the price of one more tile load.  runs_next_tile: hops (>= 0) from
// `from` to that tile, -1 if none within `hops_left`; called by all threads of the CTA (it
// synchronises the CTA: the previous tile is done with when it returns).
__device__ __forceinline__ uint32_t runs_tile_bundles(const int64_t* __restrict__ tile_bstart, int t) {
  return (uint32_t)(tile_bstart[t + 1] - tile_bstart[t]);
}
__device__ __forceinline__ int runs_next_tile(const uint32_t* tile_ctr,
                                              const int64_t* __restrict__ tile_bstart, int n_tiles,
                                              int from, int hops_left, int* s_next) {
  for (int h0 = 0; h0 < hops_left; h0 += (int)blockDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) *s_next = 0x7fffffff;
    __syncthreads();
    const int h = h0 + (int)threadIdx.x;
    if (h < hops_left) {
      const int t = (from + h) % n_tiles;
      const uint32_t taken = *reinterpret_cast<const volatile uint32_t*>(tile_ctr + t);
      if (taken < runs_tile_bundles(tile_bstart, t)) atomicMin(s_next, h);
    }
    __syncthreads();
    const int r = *s_next;
    if (r != 0x7fffffff) return r;
  }
  return -1;
}

}  // namespace amf
