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
// The spin is BOUNDED: a peer that never publishes (crashed rank, mismatched call sequence) costs `timeout_ns` once, after
// which the waiter records (source rank + 1) | epoch << 8 in a device status word (sticky: later waits return at once) and in
// a host-mapped word the Python side polls without synchronising, and the host raises instead of the GPU hanging.
// One exchange may be produced by SEVERAL launches (token chunks of the QKV projection, so that chunk i's stores overlap chunk
// i+1's GEMM): the arrival counter then counts the CTAs of all of them and the last one publishes.
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

DEVI unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}

__global__ void comm_wait_kernel(const unsigned int* flags, int P, unsigned int epoch, unsigned int* status_dev,
                                 unsigned int* status_host, unsigned long long timeout_ns) {
  if (threadIdx.x < P) {
    if (status_dev && *reinterpret_cast<volatile unsigned int*>(status_dev) != 0u) return;     // an earlier wait already failed
    const unsigned long long t0 = global_timer_ns();
    // epochs only grow; wrap-around safe comparison
    while (static_cast<int>(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
      __nanosleep(100);
      if (global_timer_ns() - t0 > timeout_ns) {
        const unsigned int code = (static_cast<unsigned int>(threadIdx.x) + 1u) | (epoch << 8);
        if (status_dev) atomicCAS(status_dev, 0u, code);
        if (status_host) { *reinterpret_cast<volatile unsigned int*>(status_host) = code; __threadfence_system(); }   // plain store: PCIe need not support atomics
        break;
      }
    }
  }
}

// All-gather by peer stores: segment b of the local tensor (nseg segments of seg_bytes, contiguous) goes to byte offset
// (b * P + rank) * seg_bytes of every rank's destination, so each destination holds [nseg, P, seg_bytes]: the token shards of
// every rank side by side, per sequence (the head output gather of xdit_context_parallel.py:142 without a collective call).
__global__ void __launch_bounds__(256) peer_allgather_kernel(const uint4* __restrict__ src, long long seg_vec, int nseg, const PeerPtrs pp) {
  const long long total = seg_vec * nseg;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const long long b = i / seg_vec, o = i - b * seg_vec;
    const uint4 v = src[i];
    const long long d = (b * pp.P + pp.rank) * seg_vec + o;
#pragma unroll 1
    for (int r = 0; r < pp.P; ++r) static_cast<uint4*>(pp.data[r])[d] = v;
  }
  peer_signal_done(pp, gridDim.x);
}

}  // namespace b200
