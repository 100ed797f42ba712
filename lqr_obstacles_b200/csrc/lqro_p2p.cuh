// lqro_p2p.cuh -- the all-gather of a sharded step over NVLink / NVSwitch peer memory.
//
// The exchange of a sharded step is tiny: every rank contributes its block of the state estimates
// (n_local x 16 doubles, 1 MiB per rank at 65,536 agents on 8 GPUs) and needs everybody else's before the pair kernel starts.
// ncclAllGather takes ~0.055 ms for that, all of it latency, on the critical path of a 1 ms step.  Here every rank PUSHES its
// block straight into the state buffer of every peer (the buffers are mapped into each other's address space once, through CUDA
// IPC handles that lqro_comm_init exchanges over the NCCL communicator it has just built) and tells the peer with one flag:
//
//   flags of rank B (ints, in B's memory):  DONE[A]  last epoch for which A has finished READING its own state buffer
//                                            ARR[A]   last epoch whose block of A has arrived in B's state buffer
//   k_p2p_gather on rank A, epoch e:  wait until DONE[r] >= e - 1 for every peer r (in A's own flags: r wrote it at the end of its
//     step e - 1, so r no longer reads the slot A is about to overwrite in r's buffer); copy A's block into every peer's buffer
//     (16-byte stores over NVLink); fence (system scope); the last block sets ARR[A] = e in every peer's flags and then waits
//     until ARR[r] >= e for every r in A's own flags -- when the kernel ends, the whole state is in A's buffer.
//   k_p2p_done on rank A, epoch e (behind the last reader of the state in the step): DONE[A] = e in every peer's flags.
//
// No rank waits for anything a peer only produces after hearing from it: DONE(e - 1) is written at the end of step e - 1, ARR(e)
// right after DONE(e - 1) of everybody has been seen.  Every wait gives up after LQRO_P2P_TIMEOUT_NS (a peer that died) and
// raises the error word the host reads back with the step's results.  NCCL stays the fallback (no peer access, one process
// cannot open a handle, LQRO_NO_P2P_GATHER=1) -- the choice is agreed on by all ranks in lqro_comm_init.
#pragma once
#include "lqro_common.cuh"

#define LQRO_P2P_MAX_RANKS 16
#define LQRO_P2P_FLAG_DONE 0                         // [0, 16)
#define LQRO_P2P_FLAG_ARR LQRO_P2P_MAX_RANKS         // [16, 32)
#define LQRO_P2P_FLAG_CTR (2 * LQRO_P2P_MAX_RANKS)   // blocks of the running gather kernel that have pushed their share
#define LQRO_P2P_FLAG_ERR (2 * LQRO_P2P_MAX_RANKS + 1)
#define LQRO_P2P_FLAG_INTS 64
#ifndef LQRO_P2P_TIMEOUT_NS
#define LQRO_P2P_TIMEOUT_NS 10000000000ull           // 10 s
#endif

struct P2PPeers {
  double* x[LQRO_P2P_MAX_RANKS];   // state buffer of every rank (own entry: the local buffer)
  int* flags[LQRO_P2P_MAX_RANKS];  // flag words of every rank
};

__device__ __forceinline__ unsigned long long p2p_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void p2p_store_flag(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int p2p_load_flag(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// wait until *p >= want; false after the timeout
__device__ __forceinline__ bool p2p_wait_flag(const int* p, int want) {
  if (p2p_load_flag(p) >= want) return true;
  const unsigned long long t0 = p2p_globaltimer();
  while (p2p_load_flag(p) < want) {
    if (p2p_globaltimer() - t0 > LQRO_P2P_TIMEOUT_NS) return false;
    __nanosleep(64);
  }
  return true;
}

// grid: any number of blocks of 256 threads; cnt2 = double2 elements of one rank's block (8 * n_local)
__global__ void __launch_bounds__(256) k_p2p_gather(P2PPeers peers, int rank, int world, size_t cnt2, int epoch) {
  int* mine = peers.flags[rank];
  __shared__ int s_last;
  // (A) nobody still reads the slot this rank is about to overwrite
  if (threadIdx.x < world && threadIdx.x != rank) {
    if (!p2p_wait_flag(mine + LQRO_P2P_FLAG_DONE + threadIdx.x, epoch - 1)) atomicExch(mine + LQRO_P2P_FLAG_ERR, 1);
  }
  __syncthreads();
  // (B) push the own block into every peer's buffer, peers interleaved so that all links are busy at once
  {
    const double2* src = reinterpret_cast<const double2*>(peers.x[rank]) + (size_t)rank * cnt2;
    for (size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x; m < cnt2; m += (size_t)gridDim.x * blockDim.x) {
      const double2 v = src[m];
      for (int q = 1; q < world; ++q) {
        int r = rank + q;  // every rank starts with a different peer
        if (r >= world) r -= world;
        reinterpret_cast<double2*>(peers.x[r])[(size_t)rank * cnt2 + m] = v;
      }
    }
  }
  // (C) the block's stores are ordered before the flag of the last block
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(mine + LQRO_P2P_FLAG_CTR, 1);
    s_last = done == (int)gridDim.x - 1;
    if (s_last) mine[LQRO_P2P_FLAG_CTR] = 0;  // for the next launch (stream-ordered behind this one)
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x < world && threadIdx.x != rank) {
    p2p_store_flag(peers.flags[threadIdx.x] + LQRO_P2P_FLAG_ARR + rank, epoch);
    // (D) everybody else's block has landed here
    if (!p2p_wait_flag(mine + LQRO_P2P_FLAG_ARR + threadIdx.x, epoch)) atomicExch(mine + LQRO_P2P_FLAG_ERR, 1);
  }
  __threadfence_system();
}

// one block: this rank has finished reading its state buffer for `epoch`
__global__ void __launch_bounds__(32) k_p2p_done(P2PPeers peers, int rank, int world, int epoch) {
  __threadfence_system();
  if (threadIdx.x < world && threadIdx.x != rank) p2p_store_flag(peers.flags[threadIdx.x] + LQRO_P2P_FLAG_DONE + rank, epoch);
}
