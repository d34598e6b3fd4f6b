// Peer-memory plumbing for the Ulysses sequence-parallel exchange (one process per GPU, NVLink P2P).
//
// The exchange itself is not a collective call: the kernels that PRODUCE the data (q/k-norm + RoPE before
// attention, the attention epilogue after it) store their results straight into the destination rank's buffer
// through peer pointers (cudaIpc mappings), so the head<->sequence all-to-all of
// wan/distributed/xdit_context_parallel.py:179-184 costs no extra pass over memory and overlaps with the compute
// of the producing kernel.  Completion is signalled with one flag per (destination, source) pair:
//   producer: every CTA  stores -> __threadfence_system() -> arrives on a local counter; the last CTA to arrive
//             publishes `epoch` into flag[src] on every peer (st.release.sys);
//   consumer: comm_wait_kernel spins (ld.acquire.sys) until all P flags of its own array reach `epoch`; the next
//             kernel in the stream (attention / output projection) then reads the buffer.
// A rank only ever waits for kernels that other ranks launch unconditionally, so there is no circular wait.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kMaxPeers = 8;

struct PeerPtrs {
  void* data[kMaxPeers];            // destination buffer on each rank (own rank: local pointer)
  unsigned int* flags[kMaxPeers];   // flag array [P] on each rank
  int P, rank;
  unsigned int epoch;
  unsigned int* counter;            // local arrival counter (zero between launches)
};

DEVI void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
DEVI unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Called by every thread of a CTA after its last peer store.  `num_ctas` CTAs take part in the launch.
DEVI void peer_signal_done(const PeerPtrs& pp, unsigned int num_ctas) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(pp.counter, 1u);
    if (prev + 1 == num_ctas) {
      *pp.counter = 0;                       // ready for the next launch (stream-ordered)
      __threadfence_system();
      for (int r = 0; r < pp.P; ++r) st_release_sys(pp.flags[r] + pp.rank, pp.epoch);
    }
  }
}

__global__ void comm_wait_kernel(const unsigned int* flags, int P, unsigned int epoch) {
  if (threadIdx.x < P) {
    // epochs only grow; wrap-around safe comparison
    while (static_cast<int>(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) __nanosleep(100);
  }
}

}  // namespace b200
