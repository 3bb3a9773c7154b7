// Pieces shared by the kernels that keep a tile of one factor matrix in shared memory and
// stream packed (row, local index) words past it (pool.cu: candidate scoring, tiled.cu: the
// PMF loss + gradient): TMA bulk copies with mbarrier completion, 16-byte shared / read-only
// global loads, packed fp32x2 arithmetic (FFMA2 on sm_100a).
#pragma once
#include "common.cuh"

namespace amf {

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
// read-only 16-byte load of a factor-row slice (L2 resident)
__device__ __forceinline__ float4 ldg_v(const unsigned char* p, float4) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ double2 ldg_v(const unsigned char* p, double2) {
  return __ldg(reinterpret_cast<const double2*>(p));
}
// read-only 32-byte load of two adjacent slices (one LDG.256 on sm_100a)
__device__ __forceinline__ void ldg_v2(const unsigned char* p, float4& lo, float4& hi) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y),
                 "=f"(hi.z), "=f"(hi.w)
               : "l"(p));
}
__device__ __forceinline__ void ldg_v2(const unsigned char* p, double2& lo, double2& hi) {
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
               : "=d"(lo.x), "=d"(lo.y), "=d"(hi.x), "=d"(hi.y)
               : "l"(p));
}
__device__ __forceinline__ float4 vsel(bool c, const float4& a, const float4& b) {
  return make_float4(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z, c ? a.w : b.w);
}
__device__ __forceinline__ double2 vsel(bool c, const double2& a, const double2& b) {
  return make_double2(c ? a.x : b.x, c ? a.y : b.y);
}

// Which 16-byte slices of a row a lane owns, for rows of NVEC slices shared by four lanes
// (CPL = NVEC / 4 per lane), and in which order it visits them.  The two lane groups that share
// a shared-memory wavefront (group parity gq) visit their slices in a different order, so the
// eight lanes of a wavefront always touch eight different 16-byte bank groups.
//   ADJ (CPL == 2 only): lane l owns the adjacent slices 2l, 2l+1 (one 256-bit global load
//             fetches both, at the price of 8 selects) and visits 2l+gq first;
//   else    : lane l owns l, l+4, l+8, ... and visits l + 4*(t ^ gq) at step t.
// Either way the t-th slice is at  off0 ^ slice_xor(t)  with off0 = slice_off0(...) the byte
// offset of the first one and slice_xor a compile-time constant (16*t or 64*t): one address per
// row, the other slices by XOR with an immediate.
template <int CPL, bool ADJ>
__device__ __forceinline__ uint32_t slice_off0(int l, int gq, bool have) {
  static_assert(!ADJ || CPL == 2, "adjacent ownership is for two slices per lane");
  static_assert((CPL & (CPL - 1)) == 0, "slices per lane must be a power of two");
  const int ll = have ? l : 0;
  if constexpr (ADJ) return (uint32_t)(2 * ll + gq) * 16u;
  return (uint32_t)(ll + 4 * (CPL > 1 ? gq : 0)) * 16u;
}
template <int CPL, bool ADJ>
__device__ __forceinline__ constexpr uint32_t slice_xor(int t) {
  return ADJ ? 16u * (uint32_t)t : 64u * (uint32_t)t;
}
// the lane's slices of the row at `row` (global, row-aligned), in visiting order
template <typename V, int CPL, bool ADJ>
__device__ __forceinline__ void load_row_slices(uint64_t row, int l, int gq, uint32_t off0,
                                                V (&a)[CPL]) {
  if constexpr (ADJ) {
    V n0, n1;
    ldg_v2(reinterpret_cast<const unsigned char*>(row + 32u * (uint32_t)l), n0, n1);
    a[0] = vsel(gq != 0, n1, n0);
    a[1] = vsel(gq != 0, n0, n1);
  } else {
#pragma unroll
    for (int t = 0; t < CPL; ++t)
      a[t] = ldg_v(reinterpret_cast<const unsigned char*>((row + off0) ^ (uint64_t)slice_xor<CPL, ADJ>(t)), V());
  }
}

__device__ __forceinline__ float vdot_acc(const float4& a, const float4& b, float acc) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}
__device__ __forceinline__ double vdot_acc(const double2& a, const double2& b, double acc) {
  return fma(a.x, b.x, fma(a.y, b.y, acc));
}


// ---- packed fp32x2 arithmetic: one FFMA2 issues two fused multiply-adds ----------------------
__device__ __forceinline__ float2 fma2(const float2& a, const float2& b, const float2& c) {
  unsigned long long ra = *reinterpret_cast<const unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<const unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<const unsigned long long*>(&c);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}


// dot product of the C 16-byte slices a lane holds of two rows (packed FFMA2 for fp32)
template <int C>
__device__ __forceinline__ float dot_slices(const float4 (&a)[C], const float4 (&b)[C]) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < C; ++t) {
    acc = fma2(make_float2(a[t].x, a[t].y), make_float2(b[t].x, b[t].y), acc);
    acc = fma2(make_float2(a[t].z, a[t].w), make_float2(b[t].z, b[t].w), acc);
  }
  return acc.x + acc.y;
}
template <int C>
__device__ __forceinline__ double dot_slices(const double2 (&a)[C], const double2 (&b)[C]) {
  double acc = 0;
#pragma unroll
  for (int t = 0; t < C; ++t) acc = fma(a[t].x, b[t].x, fma(a[t].y, b[t].y, acc));
  return acc;
}

}  // namespace amf
