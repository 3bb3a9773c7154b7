// Builder of the "bundled runs" layout (runs.cuh) -- all on the device: one radix sort of the
// entries by (tile, own row, local row), a scan that cuts the runs into segments, one radix sort
// of the segments by (tile, length descending), and a scatter.
#include <cub/cub.cuh>

#include "common.cuh"
#include "runs.cuh"

namespace amf {
namespace {

int bits_for(uint64_t x) {
  int b = 1;
  while (b < 63 && (1ull << b) < x) ++b;
  return b;
}

// device temporaries freed on every exit path.  Stream-ordered allocations from the device's
// default pool (its release threshold is raised by acquire_partials, scoring.cu): rebuilding a pool
// of 100M candidates needs ~3 GB of them, and cudaMalloc / cudaFree of that size cost more than
// the sorts when the device memory is fragmented.
struct Scratch {
  cudaStream_t s;
  void* p[24];
  int n = 0;
  explicit Scratch(cudaStream_t stream) : s(stream) {}
  template <typename T> cudaError_t get(T** out, size_t count) {
    *out = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(out), (count ? count : 1) * sizeof(T), s);
    if (e == cudaSuccess) p[n++] = *out;
    return e;
  }
  void drop(void* q) {
    for (int t = 0; t < n; ++t)
      if (p[t] == q) { cudaFreeAsync(q, s); p[t] = p[--n]; return; }
  }
  ~Scratch() { for (int t = 0; t < n; ++t) cudaFreeAsync(p[t], s); }
};

__global__ void runs_keys_kernel(const int32_t* __restrict__ own, const int32_t* __restrict__ other,
                                 int64_t n, int tile_rows, int jbits, int ibits,
                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t i = (uint32_t)own[t], j = (uint32_t)other[t];
    const uint64_t tile = j / (uint32_t)tile_rows, jl = j % (uint32_t)tile_rows;
    keys[t] = (((tile << ibits) | i) << jbits) | jl;
    vals[t] = (uint32_t)t;
  }
}

// keys sorted ascending; start[b] = first position whose (key >> shift) >= b, start[nb] = n
template <typename K>
__global__ void runs_start_kernel(const K* __restrict__ keys, int64_t n, int shift, int64_t nb,
                                  int64_t* __restrict__ start) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p <= n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lo = (p == 0) ? -1 : (int64_t)(keys[p - 1] >> shift);
    const int64_t hi = (p == n) ? nb : (int64_t)(keys[p] >> shift);
    for (int64_t b = lo + 1; b <= hi; ++b) start[b] = p;
  }
}

// head[p] = p where a new (tile, own row) run starts, else 0: a running maximum gives every
// entry the start of its run
__global__ void runs_head_kernel(const uint64_t* __restrict__ keys, int64_t n, int jbits,
                                 uint32_t* __restrict__ head) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x)
    head[p] = (p > 0 && (keys[p] >> jbits) != (keys[p - 1] >> jbits)) ? (uint32_t)p : 0u;
}
__global__ void runs_seghead_kernel(const uint32_t* __restrict__ run_start, int64_t n,
                                    uint32_t* __restrict__ flag) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x)
    flag[p] = (((uint32_t)p - run_start[p]) % RUNS_MAXLEN) == 0 ? 1u : 0u;
}
// seg_incl = inclusive sum of the flags: entry p belongs to segment seg_incl[p] - 1
__global__ void runs_segfirst_kernel(const uint32_t* __restrict__ seg_incl, int64_t n,
                                     uint32_t* __restrict__ seg_first) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x)
    if (p == 0 || seg_incl[p] != seg_incl[p - 1]) seg_first[seg_incl[p] - 1] = (uint32_t)p;
}
__global__ void runs_segkey_kernel(const uint64_t* __restrict__ keys,
                                   const uint32_t* __restrict__ seg_first, int64_t nseg, int64_t n,
                                   int tile_shift, uint32_t* __restrict__ key2,
                                   uint32_t* __restrict__ segv) {
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < nseg;
       g += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p0 = seg_first[g];
    const uint32_t len = (g + 1 < nseg ? seg_first[g + 1] : (uint32_t)n) - p0;
    key2[g] = ((uint32_t)(keys[p0] >> tile_shift) << 7) | (uint32_t)(RUNS_MAXLEN - len);
    segv[g] = (uint32_t)g;
  }
}
__global__ void runs_nbundle_kernel(const int64_t* __restrict__ seg_start, int64_t n_tiles,
                                    int64_t* __restrict__ nb) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b <= n_tiles;
       b += (int64_t)gridDim.x * blockDim.x)
    nb[b] = b < n_tiles ? (seg_start[b + 1] - seg_start[b] + 31) / 32 : 0;
}
// every (length-sorted) segment gets its lane of its bundle
__global__ void runs_slot_kernel(const uint32_t* __restrict__ key2, const uint32_t* __restrict__ sseg,
                                 int64_t nseg, const int64_t* __restrict__ seg_start,
                                 const int64_t* __restrict__ tile_bstart,
                                 const uint64_t* __restrict__ keys,
                                 const uint32_t* __restrict__ seg_first, int jbits, uint64_t imask,
                                 uint32_t* __restrict__ seg_slot, uint32_t* __restrict__ rowid,
                                 uint8_t* __restrict__ seglen, int64_t* __restrict__ bundle_g,
                                 int32_t* __restrict__ bundle_len) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nseg;
       q += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k2 = key2[q];
    const int64_t t = k2 >> 7;
    const int len = RUNS_MAXLEN - (int)(k2 & 127u);
    const int64_t r = q - seg_start[t];
    const int64_t b = tile_bstart[t] + (r >> 5);
    const int lane = (int)(r & 31);
    const uint32_t seg = sseg[q];
    const int64_t slot = b * 32 + lane;
    seg_slot[seg] = (uint32_t)slot;
    rowid[slot] = (uint32_t)((keys[seg_first[seg]] >> jbits) & imask);
    seglen[slot] = (uint8_t)len;
    if (lane == 0) {                        // longest segment of the bundle (sorted descending)
      bundle_g[b] = (len + 3) >> 2;
      bundle_len[b] = len;
    }
  }
}
__global__ void runs_binfo_kernel(const int64_t* __restrict__ gstart,
                                  const int32_t* __restrict__ bundle_len, int64_t n_bundles,
                                  int2* __restrict__ binfo) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b <= n_bundles;
       b += (int64_t)gridDim.x * blockDim.x)
    binfo[b] = make_int2((int)gstart[b], b < n_bundles ? bundle_len[b] : 0);
}
__global__ void runs_fill16_kernel(uint16_t* __restrict__ p, int64_t n, uint16_t v) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x)
    p[t] = v;
}
template <typename VT>
__global__ void runs_scatter_kernel(const uint64_t* __restrict__ keys,
                                    const uint32_t* __restrict__ perm, const VT* __restrict__ val,
                                    int64_t n, uint64_t jmask, const uint32_t* __restrict__ seg_incl,
                                    const uint32_t* __restrict__ seg_first,
                                    const uint32_t* __restrict__ seg_slot,
                                    const int2* __restrict__ binfo, uint16_t* __restrict__ idx,
                                    VT* __restrict__ val_out, uint32_t* __restrict__ orig,
                                    uint32_t* __restrict__ pos_of) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t seg = seg_incl[p] - 1;
    const uint32_t o = (uint32_t)p - seg_first[seg];
    const uint32_t slot = seg_slot[seg];
    const int64_t pos = ((int64_t)binfo[slot >> 5].x + (o >> 2)) * 128 + (slot & 31u) * 4 + (o & 3u);
    idx[pos] = (uint16_t)(keys[p] & jmask);
    const uint32_t src = perm[p];
    if (val_out) val_out[pos] = val[src];
    if (orig) { orig[pos] = src; pos_of[src] = (uint32_t)pos; }
  }
}

}  // namespace

void runs_free(amf_runs* r) {
  cudaFree(r->idx); cudaFree(r->val); cudaFree(r->orig); cudaFree(r->pos_of); cudaFree(r->rowid);
  cudaFree(r->seglen); cudaFree(r->binfo); cudaFree(r->tile_bstart); cudaFree(r->tile_ctr);
  memset(r, 0, sizeof(*r));
}

int runs_build(amf_runs* r, int64_t n, const int32_t* own_d, const int32_t* other_d,
               const void* val_d, int val_size, int32_t own_rows, int32_t other_rows, int tile_rows,
               bool want_orig, cudaStream_t s) {
  memset(r, 0, sizeof(*r));
  AMF_REQUIRE(tile_rows >= 1 && tile_rows <= 65535, "bundled runs: tile_rows must be in [1, 65535]");
  AMF_REQUIRE(n >= 0 && n < (1ll << 31), "bundled runs: the list has too many entries");
  AMF_REQUIRE(val_size == 0 || val_size == 4 || val_size == 8, "bundled runs: bad value size");
  int jbits = 0;
  while ((1 << jbits) < tile_rows) ++jbits;
  const int ibits = bits_for((uint64_t)own_rows);
  r->n = n; r->tile_rows = tile_rows;
  r->n_tiles = (other_rows + tile_rows - 1) / tile_rows;
  const int64_t nt = r->n_tiles;
  const int tbits = bits_for((uint64_t)nt + 1);
  AMF_REQUIRE(tbits <= 24, "bundled runs: too many tiles");
  const int grid = num_sms() * 8;
  Scratch sc(s);
  int rc = AMF_OK;
  int64_t nseg = 0, n_bundles = 0, n_groups = 0;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
#define RUNS_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));          \
      rc = AMF_ERR_CUDA;                                                                         \
      goto done;                                                                                 \
    }                                                                                            \
  } while (0)
  {
    uint64_t *keys = nullptr, *keys_s = nullptr;
    uint32_t *vals = nullptr, *perm = nullptr, *head = nullptr, *seg_incl = nullptr;
    uint32_t *seg_first = nullptr, *key2 = nullptr, *key2_s = nullptr, *segv = nullptr, *sseg = nullptr;
    uint32_t* seg_slot = nullptr;
    int64_t *seg_start = nullptr, *nb = nullptr, *bundle_g = nullptr, *gstart = nullptr;
    int32_t* bundle_len = nullptr;
    RUNS_CUDA(cudaMalloc(&r->tile_bstart, 8 * (size_t)(nt + 1)));
    RUNS_CUDA(cudaMalloc(&r->tile_ctr, 4 * (size_t)nt));
    if (n == 0) {
      RUNS_CUDA(cudaMemsetAsync(r->tile_bstart, 0, 8 * (size_t)(nt + 1), s));
      RUNS_CUDA(cudaMalloc(&r->binfo, sizeof(int2)));
      RUNS_CUDA(cudaMemsetAsync(r->binfo, 0, sizeof(int2), s));
      RUNS_CUDA(cudaStreamSynchronize(s));
      goto done;
    }
    RUNS_CUDA(sc.get(&keys, (size_t)n));
    RUNS_CUDA(sc.get(&keys_s, (size_t)n));
    RUNS_CUDA(sc.get(&vals, (size_t)n));
    RUNS_CUDA(sc.get(&perm, (size_t)n));
    runs_keys_kernel<<<grid, 256, 0, s>>>(own_d, other_d, n, tile_rows, jbits, ibits, keys, vals);
    RUNS_CUDA(cudaGetLastError());
    RUNS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_s, vals, perm, n, 0,
                                              ibits + jbits + tbits, s));
    RUNS_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, s));
    RUNS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_s, vals, perm, n, 0,
                                              ibits + jbits + tbits, s));
    RUNS_CUDA(cudaStreamSynchronize(s));
    cudaFreeAsync(tmp, s); tmp = nullptr;
    sc.drop(keys); sc.drop(vals);

    // runs -> segments
    RUNS_CUDA(sc.get(&head, (size_t)n));
    RUNS_CUDA(sc.get(&seg_incl, (size_t)n));
    runs_head_kernel<<<grid, 256, 0, s>>>(keys_s, n, jbits, head);
    RUNS_CUDA(cudaGetLastError());
    tmp_bytes = 0;
    RUNS_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tmp_bytes, head, seg_incl, cub::Max(), n, s));
    RUNS_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, s));
    RUNS_CUDA(cub::DeviceScan::InclusiveScan(tmp, tmp_bytes, head, seg_incl, cub::Max(), n, s));
    runs_seghead_kernel<<<grid, 256, 0, s>>>(seg_incl, n, head);
    RUNS_CUDA(cudaGetLastError());
    RUNS_CUDA(cudaStreamSynchronize(s));
    cudaFreeAsync(tmp, s); tmp = nullptr; tmp_bytes = 0;
    RUNS_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, head, seg_incl, n, s));
    RUNS_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, s));
    RUNS_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, head, seg_incl, n, s));
    {
      uint32_t last = 0;
      RUNS_CUDA(cudaMemcpyAsync(&last, seg_incl + (n - 1), 4, cudaMemcpyDeviceToHost, s));
      RUNS_CUDA(cudaStreamSynchronize(s));
      nseg = last;
    }
    cudaFreeAsync(tmp, s); tmp = nullptr;
    sc.drop(head);
    RUNS_CUDA(sc.get(&seg_first, (size_t)nseg));
    RUNS_CUDA(sc.get(&key2, (size_t)nseg));
    RUNS_CUDA(sc.get(&key2_s, (size_t)nseg));
    RUNS_CUDA(sc.get(&segv, (size_t)nseg));
    RUNS_CUDA(sc.get(&sseg, (size_t)nseg));
    RUNS_CUDA(sc.get(&seg_slot, (size_t)nseg));
    runs_segfirst_kernel<<<grid, 256, 0, s>>>(seg_incl, n, seg_first);
    RUNS_CUDA(cudaGetLastError());
    runs_segkey_kernel<<<grid, 256, 0, s>>>(keys_s, seg_first, nseg, n, ibits + jbits, key2, segv);
    RUNS_CUDA(cudaGetLastError());
    tmp_bytes = 0;
    RUNS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key2, key2_s, segv, sseg, nseg, 0,
                                              7 + tbits, s));
    RUNS_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, s));
    RUNS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key2, key2_s, segv, sseg, nseg, 0,
                                              7 + tbits, s));
    RUNS_CUDA(cudaStreamSynchronize(s));
    cudaFreeAsync(tmp, s); tmp = nullptr;

    // segments -> bundles
    RUNS_CUDA(sc.get(&seg_start, (size_t)(nt + 1)));
    RUNS_CUDA(sc.get(&nb, (size_t)(nt + 1)));
    runs_start_kernel<uint32_t><<<grid, 256, 0, s>>>(key2_s, nseg, 7, nt, seg_start);
    RUNS_CUDA(cudaGetLastError());
    runs_nbundle_kernel<<<grid, 256, 0, s>>>(seg_start, nt, nb);
    RUNS_CUDA(cudaGetLastError());
    tmp_bytes = 0;
    RUNS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, nb, r->tile_bstart, nt + 1, s));
    RUNS_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, s));
    RUNS_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, nb, r->tile_bstart, nt + 1, s));
    RUNS_CUDA(cudaMemcpyAsync(&n_bundles, r->tile_bstart + nt, 8, cudaMemcpyDeviceToHost, s));
    RUNS_CUDA(cudaStreamSynchronize(s));
    cudaFreeAsync(tmp, s); tmp = nullptr;
    r->n_bundles = n_bundles;
    RUNS_CUDA(cudaMalloc(&r->rowid, 4 * (size_t)n_bundles * 32));
    RUNS_CUDA(cudaMalloc(&r->seglen, (size_t)n_bundles * 32));
    RUNS_CUDA(cudaMalloc(&r->binfo, sizeof(int2) * (size_t)(n_bundles + 1)));
    RUNS_CUDA(cudaMemsetAsync(r->rowid, 0xff, 4 * (size_t)n_bundles * 32, s));
    RUNS_CUDA(cudaMemsetAsync(r->seglen, 0, (size_t)n_bundles * 32, s));
    RUNS_CUDA(sc.get(&bundle_g, (size_t)(n_bundles + 1)));
    RUNS_CUDA(sc.get(&gstart, (size_t)(n_bundles + 1)));
    RUNS_CUDA(sc.get(&bundle_len, (size_t)(n_bundles + 1)));
    RUNS_CUDA(cudaMemsetAsync(bundle_g, 0, 8 * (size_t)(n_bundles + 1), s));
    runs_slot_kernel<<<grid, 256, 0, s>>>(key2_s, sseg, nseg, seg_start, r->tile_bstart, keys_s,
                                          seg_first, jbits, (1ull << ibits) - 1, seg_slot, r->rowid,
                                          r->seglen, bundle_g, bundle_len);
    RUNS_CUDA(cudaGetLastError());
    tmp_bytes = 0;
    RUNS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, bundle_g, gstart, n_bundles + 1, s));
    RUNS_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, s));
    RUNS_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, bundle_g, gstart, n_bundles + 1, s));
    RUNS_CUDA(cudaMemcpyAsync(&n_groups, gstart + n_bundles, 8, cudaMemcpyDeviceToHost, s));
    RUNS_CUDA(cudaStreamSynchronize(s));
    cudaFreeAsync(tmp, s); tmp = nullptr;
    if (n_groups >= (1ll << 24)) {
      set_error("bundled runs: %lld groups of entries exceed the 32-bit slot index", (long long)n_groups);
      rc = AMF_ERR_UNSUPPORTED;
      goto done;
    }
    r->n_groups = n_groups;
    r->npos = n_groups * 128;
    runs_binfo_kernel<<<grid, 256, 0, s>>>(gstart, bundle_len, n_bundles, r->binfo);
    RUNS_CUDA(cudaGetLastError());

    // entries
    RUNS_CUDA(cudaMalloc(&r->idx, 2 * (size_t)r->npos));
    runs_fill16_kernel<<<grid, 256, 0, s>>>(r->idx, r->npos, (uint16_t)tile_rows);
    RUNS_CUDA(cudaGetLastError());
    if (val_size) {
      RUNS_CUDA(cudaMalloc(&r->val, (size_t)val_size * (size_t)r->npos));
      RUNS_CUDA(cudaMemsetAsync(r->val, 0, (size_t)val_size * (size_t)r->npos, s));
    }
    if (want_orig) {
      RUNS_CUDA(cudaMalloc(&r->orig, 4 * (size_t)r->npos));
      RUNS_CUDA(cudaMemsetAsync(r->orig, 0xff, 4 * (size_t)r->npos, s));
      RUNS_CUDA(cudaMalloc(&r->pos_of, 4 * (size_t)n));
    }
    if (val_size == 8)
      runs_scatter_kernel<double><<<grid, 256, 0, s>>>(keys_s, perm, (const double*)val_d, n,
                                                       (1ull << jbits) - 1, seg_incl, seg_first,
                                                       seg_slot, r->binfo, r->idx, (double*)r->val,
                                                       r->orig, r->pos_of);
    else
      runs_scatter_kernel<float><<<grid, 256, 0, s>>>(keys_s, perm, (const float*)val_d, n,
                                                      (1ull << jbits) - 1, seg_incl, seg_first,
                                                      seg_slot, r->binfo, r->idx, (float*)r->val,
                                                      r->orig, r->pos_of);
    RUNS_CUDA(cudaGetLastError());
    RUNS_CUDA(cudaStreamSynchronize(s));
  }
done:
#undef RUNS_CUDA
  if (tmp) cudaFreeAsync(tmp, s);
  if (rc != AMF_OK) runs_free(r);
  return rc;
}

}  // namespace amf
