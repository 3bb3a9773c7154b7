// Micro-benchmark behind the round-2 "one lane per run" kernels (pool.cu / tiled.cu): what does it
// cost, per 128-byte factor row and per SM, to (a) bring 32 random rows into the registers of 32
// lanes (one whole row per lane) and (b) add 32 register-held rows into 32 random rows of a
// table in global memory?  Variants:
//   fetch: A per-lane LDG.128 x8 (gather)        B coalesced LDG.128 -> STS -> per-lane LDS
//          C per-lane cp.async.bulk 128 B (TMA) -> mbarrier -> per-lane LDS
//   flush: D STS -> coalesced LDS -> RED.ADD.v4  E STS -> per-lane cp.reduce.async.bulk 128 B (TMA)
//          F per-lane RED.ADD.v4 x8
//   lds:   G the inner loop alone: one random tile row per lane and step (8 LDS.128 + 16 FFMA2)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_visit micro_visit.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int THREADS = 512, WARPS = THREADS / 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds4(uint32_t a) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sts4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}

// mode: 0 A, 1 B, 2 C, 3 D, 4 E, 5 F, 6 G, 7 = B + 9 steps of G + D, 8 = C + 9 steps of G + E
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
k(const float* __restrict__ tab, float* __restrict__ out, const uint32_t* __restrict__ rows,
  int rounds, int n_rows, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t stage = smem_u32(smem) + w * 4096;
  const uint32_t bar = smem_u32(&bars[w]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float4 a[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) a[t] = make_float4(1.f + lane, 2.f, 3.f + t, 4.f);
  float acc = 0.f;
  uint32_t phase = 0;
  const int gw = blockIdx.x * WARPS + w;
  for (int r = 0; r < rounds; ++r) {
    const uint32_t row = rows[((size_t)gw * rounds + r) * 32 + lane];
    constexpr bool FB = MODE == 1 || MODE == 7, FC = MODE == 2 || MODE == 8;
    constexpr bool LD = MODE == 3 || MODE == 7, LE = MODE == 4 || MODE == 8;
    constexpr bool INNER = MODE >= 6;
    const uint32_t tile0 = smem_u32(smem) + WARPS * 4096;
    if constexpr (MODE == 0) {
      const float4* p = reinterpret_cast<const float4*>(tab + (size_t)row * 32);
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float4 v = __ldg(p + ((t + lane) & 7)); acc += v.x + v.w; }
    }
    if constexpr (LE) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging free again
      __syncwarp();
    }
    if constexpr (FB) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const uint32_t rr = __shfl_sync(0xffffffffu, row, 4 * t + (lane >> 3));
        const float4 v = __ldg(reinterpret_cast<const float4*>(tab + (size_t)rr * 32) + (lane & 7));
        sts4(stage + (4 * t + (lane >> 3)) * 128 + (lane & 7) * 16, v);
      }
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 8; ++t) { a[t] = lds4(stage + lane * 128 + ((t + lane) & 7) * 16); }
      __syncwarp();
    }
    if constexpr (FC) {
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(4096u) : "memory");
      __syncwarp();
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(stage + lane * 128), "l"(tab + (size_t)row * 32), "r"(128u), "r"(bar) : "memory");
      mbar_wait(bar, phase); phase ^= 1;
#pragma unroll
      for (int t = 0; t < 8; ++t) { a[t] = lds4(stage + lane * 128 + ((t + lane) & 7) * 16); }
      __syncwarp();
    }
    if constexpr (INNER) {
      uint32_t x = row;
      const int steps = MODE == 6 ? 16 : 9;
#pragma unroll 1
      for (int s = 0; s < steps; ++s) {
        x = x * 1664525u + 1013904223u;
        const uint32_t base = tile0 + ((x >> 8) % (uint32_t)n_rows) * 128;
        float2 c0 = make_float2(0.f, 0.f), c1 = c0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float4 v = lds4(base + (((t + lane) & 7) << 4));
          c0.x = fmaf(a[t].x, v.x, c0.x); c0.y = fmaf(a[t].y, v.y, c0.y);
          c1.x = fmaf(a[t].z, v.z, c1.x); c1.y = fmaf(a[t].w, v.w, c1.y);
        }
        acc += c0.x + c0.y + c1.x + c1.y;
      }
    } else if constexpr (FB || FC) {
#pragma unroll
      for (int t = 0; t < 8; ++t) acc += a[t].x + a[t].w;
    }
    if constexpr (LD) {
#pragma unroll
      for (int t = 0; t < 8; ++t) sts4(stage + lane * 128 + ((t + lane) & 7) * 16, a[t]);
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const uint32_t rr = __shfl_sync(0xffffffffu, row, 4 * t + (lane >> 3));
        const float4 v = lds4(stage + (4 * t + (lane >> 3)) * 128 + (lane & 7) * 16);
        red4(out + (size_t)rr * 32 + (lane & 7) * 4, v);
      }
      __syncwarp();
    }
    if constexpr (LE) {
#pragma unroll
      for (int t = 0; t < 8; ++t) sts4(stage + lane * 128 + ((t + lane) & 7) * 16, a[t]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                   :: "l"(out + (size_t)row * 32), "r"(stage + lane * 128), "r"(128u) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if constexpr (MODE == 5) {
      float* p = out + (size_t)row * 32;
#pragma unroll
      for (int t = 0; t < 8; ++t) red4(p + ((t + lane) & 7) * 4, a[t]);
    }
  }
  if constexpr (MODE == 4 || MODE == 8) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 123.456f) *sink = acc;
}

// fetch H: TMA gather4 -- eight ops of four 128-byte rows each per bundle of 32 rows, then LDS.
// VERIFY: rows are compared with a direct load (err counts mismatching floats).
__global__ void __launch_bounds__(THREADS, 1)
kg4(const __grid_constant__ CUtensorMap map, const float* __restrict__ tab, const uint32_t* __restrict__ rows,
    int rounds, float* sink, unsigned* err, int verify) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t stage = smem_u32(smem) + w * 4096;
  const uint32_t bar = smem_u32(&bars[w]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  uint32_t phase = 0;
  const int gw = blockIdx.x * WARPS + w;
  for (int r = 0; r < rounds; ++r) {
    const uint32_t row = rows[((size_t)gw * rounds + r) * 32 + lane];
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(4096u) : "memory");
    __syncwarp();
    const int r0 = __shfl_sync(0xffffffffu, row, (lane & 7) * 4 + 0), r1 = __shfl_sync(0xffffffffu, row, (lane & 7) * 4 + 1);
    const int r2 = __shfl_sync(0xffffffffu, row, (lane & 7) * 4 + 2), r3 = __shfl_sync(0xffffffffu, row, (lane & 7) * 4 + 3);
    if (lane < 8)
      asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                   :: "r"(stage + lane * 512), "l"(&map), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
    mbar_wait(bar, phase); phase ^= 1;
    float4 a[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) { a[t] = lds4(stage + lane * 128 + ((t + lane) & 7) * 16); acc += a[t].x + a[t].w; }
    if (verify) {
      unsigned bad = 0;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(tab + (size_t)row * 32) + ((t + lane) & 7));
        bad += (v.x != a[t].x) + (v.y != a[t].y) + (v.z != a[t].z) + (v.w != a[t].w);
      }
      if (bad) atomicAdd(err, bad);
    }
    __syncwarp();
  }
  if (acc == 123.456f) *sink = acc;
}

__device__ __forceinline__ float2 fma2(const float2& a, const float2& b, const float2& c) {
  unsigned long long ra = *reinterpret_cast<const unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<const unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<const unsigned long long*>(&c);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
// inner loop variants: VAR 0 = scalar FFMA dot; 1 = FFMA2 dot (pool_pred_kernel's step);
// 2 = FFMA2 dot + FFMA2 rank-one update (tiled_side_kernel's step); 3 = as 2 with scalar FFMA
template <int VAR, int TH>
__global__ void __launch_bounds__(TH, 1)
kin(const uint32_t* __restrict__ rows, int rounds, int n_rows, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float4 a[8], acc4[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) { a[t] = make_float4(1.f + lane, 2.f, 3.f + t, 4.f); acc4[t] = make_float4(0.f, 0.f, 0.f, 0.f); }
  float acc = 0.f;
  const int gw = blockIdx.x * (TH / 32) + w;
  const uint32_t tile0 = smem_u32(smem) | ((lane & 7) << 4);
  for (int r = 0; r < rounds; ++r) {
    uint32_t x = rows[((size_t)gw * rounds + r) * 32 + lane];
#pragma unroll 1
    for (int s = 0; s < 16; ++s) {
      x = x * 1664525u + 1013904223u;
      const uint32_t base = tile0 + ((x >> 8) % (uint32_t)n_rows) * 128;
      float4 v[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] = lds4(base ^ (t << 4));
      float d;
      if (VAR == 0 || VAR == 3) {
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) { c0 = fmaf(a[t].x, v[t].x, c0); c1 = fmaf(a[t].y, v[t].y, c1); c2 = fmaf(a[t].z, v[t].z, c2); c3 = fmaf(a[t].w, v[t].w, c3); }
        d = (c0 + c1) + (c2 + c3);
      } else {
        float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          if (t & 1) { s2 = fma2(make_float2(a[t].x, a[t].y), make_float2(v[t].x, v[t].y), s2); s3 = fma2(make_float2(a[t].z, a[t].w), make_float2(v[t].z, v[t].w), s3); }
          else { s0 = fma2(make_float2(a[t].x, a[t].y), make_float2(v[t].x, v[t].y), s0); s1 = fma2(make_float2(a[t].z, a[t].w), make_float2(v[t].z, v[t].w), s1); }
        }
        d = ((s0.x + s1.x) + (s2.x + s3.x)) + ((s0.y + s1.y) + (s2.y + s3.y));
      }
      if (VAR >= 2) {
        const float e = 0.5f - d;
        if (VAR == 2) {
          const float2 e2 = make_float2(e, e);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float2 lo = fma2(e2, make_float2(v[t].x, v[t].y), make_float2(acc4[t].x, acc4[t].y));
            const float2 hi = fma2(e2, make_float2(v[t].z, v[t].w), make_float2(acc4[t].z, acc4[t].w));
            acc4[t] = make_float4(lo.x, lo.y, hi.x, hi.y);
          }
        } else {
#pragma unroll
          for (int t = 0; t < 8; ++t) { acc4[t].x = fmaf(e, v[t].x, acc4[t].x); acc4[t].y = fmaf(e, v[t].y, acc4[t].y); acc4[t].z = fmaf(e, v[t].z, acc4[t].z); acc4[t].w = fmaf(e, v[t].w, acc4[t].w); }
        }
      }
      acc += d;
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) acc += acc4[t].x + acc4[t].y + acc4[t].z + acc4[t].w;
  if (acc == 123.456f) *sink = acc;
}
template <int VAR, int TH>
static void run_in(const char* name, const uint32_t* rows, int rounds, int sms, int n_rows, float* sink) {
  const size_t smem = (size_t)n_rows * 128;
  CK(cudaFuncSetAttribute(kin<VAR, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  const int rr = rounds * 16 / (TH / 32);     // same number of row reads per SM for every TH
  for (int it = 0; it < 4; ++it) {
    CK(cudaEventRecord(e0));
    kin<VAR, TH><<<sms, TH, smem>>>(rows, rr, n_rows, sink);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double n = (double)(TH / 32) * rr * 32 * 16;
  printf("%-52s %8.3f ms  = %6.2f clk/row/SM @1.965GHz\n", name, best, best * 1e6 / n * 1.965);
}

template <int MODE>
static void run(const char* name, const float* tab, float* out, const uint32_t* rows, int rounds,
                int sms, double rows_per_round, int n_rows, size_t smem, float* sink) {
  CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int it = 0; it < 4; ++it) {
    CK(cudaEventRecord(e0));
    k<MODE><<<sms, THREADS, smem>>>(tab, out, rows, rounds, n_rows, sink);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double n = (double)sms * WARPS * rounds * rows_per_round;
  printf("%-44s %8.3f ms  %7.2f ns/row/SM  = %6.2f clk/row/SM @1.965GHz\n", name, best,
         best * 1e6 / (n / sms), best * 1e6 / (n / sms) * 1.965);
}

int main() {
  int dev = 0, sms = 0; CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int n_tab = 200000, rounds = 400;
  float *tab, *out, *sink; uint32_t* rows;
  CK(cudaMalloc(&tab, (size_t)n_tab * 128)); CK(cudaMalloc(&out, (size_t)n_tab * 128)); CK(cudaMalloc(&sink, 4));
  {
    std::vector<float> ht((size_t)n_tab * 32);
    for (size_t t = 0; t < ht.size(); ++t) ht[t] = (float)(t % 100003) * 0.5f;
    CK(cudaMemcpy(tab, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice));
  }
  CK(cudaMemset(out, 0, (size_t)n_tab * 128));
  const size_t nr = (size_t)sms * WARPS * rounds * 32;
  std::vector<uint32_t> h(nr);
  uint64_t s = 88172645463325252ull;
  for (size_t t = 0; t < nr; ++t) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[t] = (uint32_t)(s % n_tab); }
  CK(cudaMalloc(&rows, nr * 4)); CK(cudaMemcpy(rows, h.data(), nr * 4, cudaMemcpyHostToDevice));
  const size_t stage = WARPS * 4096;
  run<0>("fetch A: per-lane LDG.128 x8", tab, out, rows, rounds, sms, 32, 0, stage, sink);
  run<1>("fetch B: coalesced LDG -> STS -> LDS", tab, out, rows, rounds, sms, 32, 0, stage, sink);
  run<2>("fetch C: per-lane cp.async.bulk 128B -> LDS", tab, out, rows, rounds, sms, 32, 0, stage, sink);
  run<3>("flush D: STS -> LDS -> coalesced RED.v4", tab, out, rows, rounds, sms, 32, 0, stage, sink);
  run<4>("flush E: STS -> per-lane cp.reduce.async.bulk", tab, out, rows, rounds, sms, 32, 0, stage, sink);
  run<5>("flush F: per-lane RED.v4 x8", tab, out, rows, rounds, sms, 32, 0, stage, sink);
  const int tile_rows = 1200;
  const size_t big = stage + (size_t)tile_rows * 128;
  run<6>("inner G: 8 LDS.128 + 16 FFMA / lane / step", tab, out, rows, rounds, sms, 32 * 16, tile_rows, big, sink);
  run<7>("visit B + 9 steps + D (per step)", tab, out, rows, rounds, sms, 32 * 9, tile_rows, big, sink);
  run<8>("visit C + 9 steps + E, TMA (per step)", tab, out, rows, rounds, sms, 32 * 9, tile_rows, big, sink);
  run_in<0, 512>("inner: scalar FFMA dot, 512 thr", rows, rounds, sms, 1600, sink);
  run_in<1, 512>("inner: FFMA2 dot (pool step), 512 thr", rows, rounds, sms, 1600, sink);
  run_in<2, 512>("inner: FFMA2 dot + FFMA2 update (gradient step), 512 thr", rows, rounds, sms, 1600, sink);
  run_in<3, 512>("inner: FFMA dot + FFMA update, 512 thr", rows, rounds, sms, 1600, sink);
  run_in<2, 384>("inner: FFMA2 dot + FFMA2 update, 384 thr", rows, rounds, sms, 1600, sink);
  run_in<1, 1024>("inner: FFMA2 dot, 1024 thr", rows, rounds, sms, 1600, sink);
  {
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn fn = (EncodeTiledFn)p;
    CUtensorMap map;
    const cuuint64_t gdim[2] = {32, (cuuint64_t)n_tab};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {32, 1};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, tab, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode gather4 map: %d\n", (int)cr);
    if (cr == CUDA_SUCCESS) {
      unsigned* err; CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
      CK(cudaFuncSetAttribute(kg4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage));
      kg4<<<sms, THREADS, stage>>>(map, tab, rows, 4, sink, err, 1);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned herr = 0; if (e == cudaSuccess) CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
      printf("gather4 verify: %s, mismatching floats = %u\n", cudaGetErrorString(e), herr);
      if (e == cudaSuccess) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
          CK(cudaEventRecord(e0));
          kg4<<<sms, THREADS, stage>>>(map, tab, rows, rounds, sink, err, 0);
          CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it && ms < best) best = ms;
        }
        const double n = (double)WARPS * rounds * 32;
        printf("%-44s %8.3f ms  %7.2f ns/row/SM  = %6.2f clk/row/SM @1.965GHz\n", "fetch H: TMA gather4 (8 ops / 32 rows) -> LDS", best,
               best * 1e6 / n, best * 1e6 / n * 1.965);
      }
    }
  }
  return 0;
}
