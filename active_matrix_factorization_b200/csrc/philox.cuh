// Philox4x32-10 counter-based random numbers shared by the fast-mode Gibbs kernels (gibbs.cu: row
// conditionals, gibbs_hyper.cu: Normal-Wishart hyper-parameter draws).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace amf {

// Philox4x32-10 (Salmon et al., SC'11), counter = (row, component, stream lo, stream hi), key =
// seed: every normal of a chain has its own counter, so a sweep needs no generator state and the
// rows can be sampled in any order on any number of GPUs.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}
// one standard normal (Box-Muller on the first two words; uniforms in (0, 1])
__device__ __forceinline__ double philox_normal(unsigned long long seed, unsigned long long stream,
                                                uint32_t row, uint32_t comp) {
  const uint4 r = philox4x32_10(make_uint4(row, comp, (uint32_t)stream, (uint32_t)(stream >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const double u1 = ((double)r.x + 1.0) * 2.3283064365386963e-10;
  const double u2 = ((double)r.y + 1.0) * 2.3283064365386963e-10;
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// three uniforms in (0, 1] from one counter
__device__ __forceinline__ void philox_uniform3(unsigned long long seed, unsigned long long stream,
                                                uint32_t row, uint32_t comp, double& u1, double& u2,
                                                double& u3) {
  const uint4 r = philox4x32_10(make_uint4(row, comp, (uint32_t)stream, (uint32_t)(stream >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  u1 = ((double)r.x + 1.0) * 2.3283064365386963e-10;
  u2 = ((double)r.y + 1.0) * 2.3283064365386963e-10;
  u3 = ((double)r.z + 1.0) * 2.3283064365386963e-10;
}

}  // namespace amf
