// The cross-GPU flag barrier of the NVLink-fused Ulysses exchange (peer_barrier_kernel, elementwise.cu) folded into the
// prologue of the kernel that CONSUMES what the peers wrote — no separate launch between producer and consumer:
//   enter (all threads, right after pdl_wait(), i.e. once the stream predecessor — the kernel that wrote this rank's
//          share into the peers' buffers — has completed): block 0 raises this rank's flag on every peer to epoch + 1,
//          every block waits until all peers have raised theirs here, then the block proceeds to read;
//   exit  (one thread per block, at the end): the LAST block to leave advances the device-side epoch, so every block of
//          the launch read the same epoch (a block reads it before it leaves) and a replayed CUDA graph keeps counting.
// Same flags, same epoch counter and same ordering argument as the stand-alone barrier kernel; the two can be mixed.
#pragma once

#include "common.cuh"
#include "ptx.cuh"

namespace ltxb {

struct PeerSync {
  int* flags[8];  // flags[i] = rank i's flag array (n_peers ints), mapped into this process
  int n_peers;    // 0: no barrier
  int my_rank;
  int* epoch_counter;  // barriers this rank has passed
  int* done_counter;   // blocks of the current launch that have left; zero between launches
};

__device__ __forceinline__ int peer_sync_enter(const PeerSync& s) {
  if (s.n_peers == 0) return 0;
  int epoch = 0;
  if (threadIdx.x < 32) {
    const int i = threadIdx.x;
    if (i < s.n_peers) {
      epoch = *reinterpret_cast<volatile int*>(s.epoch_counter) + 1;
      if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        __threadfence_system();
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(s.flags[i] + s.my_rank), "r"(epoch) : "memory");
      }
      const int* mine = s.flags[s.my_rank] + i;
      const long long t0 = clock64();
      int seen;
      do {
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        if (seen - epoch < 0 && clock64() - t0 > LTXB_WATCHDOG_CYCLES) {
          printf("ltxb: folded peer barrier watchdog: rank %d block %d waits for rank %d (epoch %d, seen %d)\n", s.my_rank, (int)blockIdx.x, i, epoch, seen);
          __trap();
        }
      } while (seen - epoch < 0);
      __threadfence_system();
    }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
  }
  __syncthreads();
  return epoch;  // valid in warp 0 (the exit is called from there)
}

// call from ONE thread of warp 0 per block, after the block's last read of peer-written data
__device__ __forceinline__ void peer_sync_exit(const PeerSync& s, int epoch) {
  if (s.n_peers == 0) return;
  const int blocks = gridDim.x * gridDim.y * gridDim.z;
  if (atomicAdd(s.done_counter, 1) == blocks - 1) {
    *s.done_counter = 0;
    *s.epoch_counter = epoch;
    __threadfence();
  }
}

}  // namespace ltxb

namespace ltxb {
// the same barrier as a kernel of its own (elementwise.cu), for consumers without the folded form
int launch_peer_barrier(const PeerSync& s, cudaStream_t stream);
// host: C-ABI struct -> kernel parameter (NULL -> n_peers 0)
inline int peer_sync_from_abi(const ltxb_peer_sync* in, PeerSync* out, const char* who) {
  *out = PeerSync{};
  if (in == nullptr) return LTXB_OK;
  if (in->n_peers < 1 || in->n_peers > 8 || in->my_rank < 0 || in->my_rank >= in->n_peers || in->epoch_counter == nullptr ||
      in->done_counter == nullptr)
    return set_error(LTXB_ERR_BAD_ARG, "%s: bad peer_sync (n_peers=%d, my_rank=%d)", who, in->n_peers, in->my_rank);
  for (int i = 0; i < in->n_peers; ++i) {
    if (in->flags[i] == nullptr) return set_error(LTXB_ERR_BAD_ARG, "%s: peer_sync flag pointer %d is null", who, i);
    out->flags[i] = in->flags[i];
  }
  out->n_peers = in->n_peers, out->my_rank = in->my_rank, out->epoch_counter = in->epoch_counter, out->done_counter = in->done_counter;
  return LTXB_OK;
}
}  // namespace ltxb
