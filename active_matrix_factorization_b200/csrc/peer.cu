// Winner exchange of a sharded candidate pool over NVLink peer memory: the "NCCL argmax /
// all-gather" of the per-GPU winners (active_pmf.py:765-770 reduces the Pool.map chunks with
// `chooser`) as ONE kernel per rank that writes its 16-byte record into every peer's
// mailbox with remote stores, raises a flag there, waits for the flags of all peers in its own
// mailbox and reduces the records with the tie-break of amf_best_reduce -- no collective library
// call, no host synchronisation.  Mailboxes are plain cudaMalloc buffers opened in the peer
// processes through CUDA IPC (one process per GPU, one node); epochs alternate between two slots so
// that a fast rank two steps ahead cannot overwrite a record a slow rank is still reading.
#include "peer.cuh"

namespace amf {
namespace {

template <bool MAX>
__global__ void __launch_bounds__(32)
peer_best_kernel(unsigned char* const* __restrict__ peers, int world, int rank, unsigned int epoch,
                 const Best* __restrict__ mine, amf_best_t* __restrict__ out) {
  const int lane = threadIdx.x;
  const Best b = peer_exchange_warp<MAX>(peers, world, rank, epoch, *mine, lane);
  if (lane == 0) { out->value = b.i < 0 ? 0.0 : b.v; out->index = b.i; }
}

}  // namespace
}  // namespace amf

using namespace amf;

extern "C" {
#pragma GCC visibility push(default)

int amf_peer_create(amf_peer_t** out, int world, int rank, unsigned char* ipc_handle_out) {
  AMF_REQUIRE(out && ipc_handle_out, "amf_peer_create: NULL argument");
  AMF_REQUIRE(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world,
              "amf_peer_create: world must be in [1, %d]", PEER_MAX);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles travel as 64 bytes");
  amf_peer* p = new amf_peer();
  memset(p, 0, sizeof(*p));
  p->world = world; p->rank = rank;
  cudaIpcMemHandle_t hnd;
  if (cudaMalloc(&p->local, PEER_BYTES) != cudaSuccess || cudaMemset(p->local, 0, PEER_BYTES) != cudaSuccess ||
      cudaIpcGetMemHandle(&hnd, p->local) != cudaSuccess) {
    set_error("amf_peer_create: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(p->local);
    delete p;
    return AMF_ERR_CUDA;
  }
  memcpy(ipc_handle_out, &hnd, 64);
  p->peers_h = new unsigned char*[world]();
  *out = p;
  return AMF_OK;
}

int amf_peer_connect(amf_peer_t* p, const unsigned char* all_handles) {
  AMF_REQUIRE(p && all_handles, "amf_peer_connect: NULL argument");
  AMF_REQUIRE(!p->connected, "amf_peer_connect: already connected");
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) { p->peers_h[r] = p->local; continue; }
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, all_handles + 64 * (size_t)r, 64);
    void* q = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&q, hnd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("amf_peer_connect: cannot map the mailbox of rank %d: %s", r, cudaGetErrorString(e));
      cudaGetLastError();
      return AMF_ERR_UNSUPPORTED;
    }
    p->peers_h[r] = static_cast<unsigned char*>(q);
  }
  AMF_CUDA(cudaMalloc(&p->peers_d, sizeof(unsigned char*) * p->world));
  AMF_CUDA(cudaMemcpy(p->peers_d, p->peers_h, sizeof(unsigned char*) * p->world, cudaMemcpyHostToDevice));
  p->connected = true;
  return AMF_OK;
}

int amf_peer_best_reduce(amf_peer_t* p, const amf_best_t* mine_d, int maximize, amf_best_t* out_d,
                         void* stream) {
  AMF_REQUIRE(p && mine_d && out_d, "amf_peer_best_reduce: NULL argument");
  AMF_REQUIRE(p->connected, "amf_peer_best_reduce: amf_peer_connect has not been called");
  const unsigned int epoch = ++p->epoch;                // every rank calls in the same order
  cudaStream_t s = (cudaStream_t)stream;
  const Best* mine = reinterpret_cast<const Best*>(mine_d);
  if (maximize) peer_best_kernel<true><<<1, 32, 0, s>>>(p->peers_d, p->world, p->rank, epoch, mine, out_d);
  else peer_best_kernel<false><<<1, 32, 0, s>>>(p->peers_d, p->world, p->rank, epoch, mine, out_d);
  AMF_LAUNCH_CHECK();
  return AMF_OK;
}

int amf_peer_destroy(amf_peer_t* p) {
  if (!p) return AMF_OK;
  if (p->peers_h) {
    for (int r = 0; r < p->world; ++r)
      if (r != p->rank && p->peers_h[r]) cudaIpcCloseMemHandle(p->peers_h[r]);
    delete[] p->peers_h;
  }
  cudaFree(p->peers_d);
  cudaFree(p->local);
  delete p;
  return AMF_OK;
}

#pragma GCC visibility pop
}  // extern "C"
