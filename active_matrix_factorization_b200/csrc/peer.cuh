// Device side of the winner exchange over NVLink peer memory (peer.cu): shared with the kernels
// that finish with it (pool.cu fuses it into the scoring kernel's last CTA).
#pragma once
#include "common.cuh"

struct amf_peer {
  int world, rank;
  unsigned char* local;        // this rank's mailbox (device memory, exported through CUDA IPC)
  unsigned char** peers_h;     // mailbox of every rank as mapped here (peers_h[rank] == local)
  unsigned char** peers_d;     // the same table on the device
  unsigned int epoch;          // number of exchanges so far (every rank counts alike)
  bool connected;
};

namespace amf {

constexpr int PEER_MAX = 32;
// mailbox: Best recs[2][PEER_MAX], then unsigned flags[2][PEER_MAX]
constexpr size_t PEER_RECS = 2 * PEER_MAX * sizeof(Best);
constexpr size_t PEER_BYTES = PEER_RECS + 2 * PEER_MAX * sizeof(unsigned int);

__device__ __forceinline__ void st_sys_v2(void* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One warp: `me` (the same record in every lane) goes into every peer's mailbox with remote stores
// and a release flag; the records of all peers are awaited (acquire) in this rank's mailbox and
// reduced with the rule of amf_best_reduce (best value, lowest index on ties, NaN / index < 0 never
// win).  The result is valid in every lane.  A peer that never arrives traps after ~1 s.
template <bool MAX>
__device__ __forceinline__ Best peer_exchange_warp(unsigned char* const* __restrict__ peers, int world,
                                                   int rank, unsigned int epoch, Best me, int lane) {
  const int slot = (int)(epoch & 1u);
  if (lane < world) {
    unsigned char* box = peers[lane];
    Best* recs = reinterpret_cast<Best*>(box) + slot * PEER_MAX;
    unsigned int* flags = reinterpret_cast<unsigned int*>(box + PEER_RECS) + slot * PEER_MAX;
    st_sys_v2(&recs[rank], (unsigned long long)__double_as_longlong(me.v), (unsigned long long)me.i);
    st_release_sys(&flags[rank], epoch);
  }
  Best b{0.0, -1};
  if (lane < world) {
    const unsigned char* box = peers[rank];
    const Best* recs = reinterpret_cast<const Best*>(box) + slot * PEER_MAX;
    const unsigned int* flags = reinterpret_cast<const unsigned int*>(box + PEER_RECS) + slot * PEER_MAX;
    unsigned int spins = 0;
    while (ld_acquire_sys(&flags[lane]) != epoch) {
      __nanosleep(40);
      if (++spins > (1u << 25)) {                       // > 1 s: a peer never arrived
        printf("amf peer exchange: rank %d timed out waiting for rank %d (epoch %u)\n", rank, lane, epoch);
        __trap();
      }
    }
    b = recs[lane];
    if (b.v != b.v) b.i = -1;                           // NaN never wins (amf_best_reduce)
  }
  return warp_best<MAX>(b);
}

}  // namespace amf
