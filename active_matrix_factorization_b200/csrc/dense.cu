// Dense prediction matrix and the errors taken on it (pmf_cy.pyx:25-29 rmse / rmse_on, :410-426
// predicted_matrix, rmse): out = U V' + offset as a register-blocked shared-memory product, and the
// fused form that never writes the N x M matrix: sum over the selected cells of
// (real_ij - U_i . V_j - offset)^2 and their number.  The matrices of this path are the toy /
// movielens sizes of BASELINE configs 1-4 (1.6M cells, rank <= 32): a latency-class kernel, kept
// in the compute type of the model with fp64 accumulation of the error sums.
#include "common.cuh"

namespace amf {
namespace {

constexpr int DT = 64;        // output tile is DT x DT cells, 16 x 16 threads, 4 x 4 cells per thread
constexpr int DK = 16;        // depth of one shared-memory stage

// MODE 0: write the tile to out; MODE 1: accumulate (real - pred)^2 over the selected cells
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
dense_pred_kernel(const T* __restrict__ U, const T* __restrict__ V, int n, int m, int d, int ld,
                  T offset, T* __restrict__ out, const double* __restrict__ real,
                  const unsigned char* __restrict__ mask, double* __restrict__ sums) {
  __shared__ T su[DK][DT + 1], sv[DK][DT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * DT, j0 = blockIdx.x * DT;
  T acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = T(0);
  for (int k0 = 0; k0 < d; k0 += DK) {
    __syncthreads();
    for (int t = threadIdx.x; t < DT * DK; t += 256) {
      const int r = t / DK, k = t % DK;
      su[k][r] = (i0 + r < n && k0 + k < d) ? U[(int64_t)(i0 + r) * ld + k0 + k] : T(0);
      sv[k][r] = (j0 + r < m && k0 + k < d) ? V[(int64_t)(j0 + r) * ld + k0 + k] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < DK; ++k) {
      T a4[4], b4[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { a4[a] = su[k][ty + 16 * a]; b4[a] = sv[k][tx + 16 * a]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(a4[a], b4[b], acc[a][b]);
    }
  }
  double sq = 0, cnt = 0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty + 16 * a;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = j0 + tx + 16 * b;
      if (i < n && j < m) {
        const int64_t c = (int64_t)i * m + j;
        const T p = acc[a][b] + offset;
        if (MODE == 0) {
          out[c] = p;
        } else if (!mask || mask[c]) {
          const double e = real[c] - (double)p;
          sq = fma(e, e, sq);
          cnt += 1.0;
        }
      }
    }
  }
  if (MODE == 1) {
    sq = block_sum(sq);
    cnt = block_sum(cnt);
    if (threadIdx.x == 0) { atomicAdd(sums, sq); atomicAdd(sums + 1, cnt); }
  }
}

template <typename T>
int dense_launch(int n, int m, int d, int ld, const T* U, const T* V, double offset, T* out,
                 const double* real, const unsigned char* mask, double* sums, cudaStream_t s) {
  const dim3 grid((m + DT - 1) / DT, (n + DT - 1) / DT);
  if (out) {
    dense_pred_kernel<T, 0><<<grid, 256, 0, s>>>(U, V, n, m, d, ld, (T)offset, out, nullptr, nullptr, nullptr);
  } else {
    AMF_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), s));
    dense_pred_kernel<T, 1><<<grid, 256, 0, s>>>(U, V, n, m, d, ld, (T)offset, nullptr, real, mask, sums);
  }
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

}  // namespace
}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int amf_predicted_matrix(int dtype, int32_t n, int32_t m, int d, int ld, const void* U_d,
                         const void* V_d, double offset, void* out_d, void* stream) {
  AMF_REQUIRE(U_d && V_d && out_d, "amf_predicted_matrix: NULL argument");
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_predicted_matrix: bad dtype");
  AMF_REQUIRE(n > 0 && m > 0 && d >= 1 && ld >= d, "amf_predicted_matrix: bad sizes");
  AMF_REQUIRE((n + DT - 1) / DT <= 65535, "amf_predicted_matrix: matrix too tall");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == AMF_F32)
    return dense_launch<float>(n, m, d, ld, (const float*)U_d, (const float*)V_d, offset, (float*)out_d,
                               nullptr, nullptr, nullptr, s);
  return dense_launch<double>(n, m, d, ld, (const double*)U_d, (const double*)V_d, offset,
                              (double*)out_d, nullptr, nullptr, nullptr, s);
}

int amf_sq_error_dense(int dtype, int32_t n, int32_t m, int d, int ld, const void* U_d,
                       const void* V_d, double offset, const double* real_d,
                       const unsigned char* mask_d, double* sums_d, void* stream) {
  AMF_REQUIRE(U_d && V_d && real_d && sums_d, "amf_sq_error_dense: NULL argument");
  AMF_REQUIRE(dtype == AMF_F32 || dtype == AMF_F64, "amf_sq_error_dense: bad dtype");
  AMF_REQUIRE(n > 0 && m > 0 && d >= 1 && ld >= d, "amf_sq_error_dense: bad sizes");
  AMF_REQUIRE((n + DT - 1) / DT <= 65535, "amf_sq_error_dense: matrix too tall");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == AMF_F32)
    return dense_launch<float>(n, m, d, ld, (const float*)U_d, (const float*)V_d, offset, nullptr,
                               real_d, mask_d, sums_d, s);
  return dense_launch<double>(n, m, d, ld, (const double*)U_d, (const double*)V_d, offset, nullptr,
                              real_d, mask_d, sums_d, s);
}

#pragma GCC visibility pop
}  // extern "C"
