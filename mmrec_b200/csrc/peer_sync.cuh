// Barriers between the ranks INSIDE a kernel, over peer-mapped memory (shared by peer.cu and topk.cu).
#pragma once
#include <stdio.h>

#include "common.cuh"

namespace mmrec {

constexpr int PEER_MAX = 16;

// Every rank owns a flag array `flags` of 2 * world ints in the symmetric buffer (zero at start): slot [b * world + p] is
// written by rank p for barrier b of the current call (b = 0: "my input is complete", b = 1: "my stores have landed").
// Calls are numbered by a per-rank device counter (`state[0]`, the same sequence on every rank), so the same kernel can
// be replayed from a CUDA graph: nothing about the barrier is baked into the launch.  Spins are bounded (a rank that
// never arrives must trap, not hang the box).
struct PeerFlags { int* f[PEER_MAX]; };

__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long peer_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// one thread: announce `epoch` in slot (b, rank) of every rank's flags, wait for every rank's announcement in mine
__device__ __forceinline__ void peer_barrier(const PeerFlags& F, int b, int rank, int world, int epoch) {
    __threadfence_system();
    for (int p = 0; p < world; ++p) st_release_sys(F.f[p] + b * world + rank, epoch);
    const unsigned long long t0 = peer_now_ns();
    for (int p = 0; p < world; ++p) {
        unsigned spins = 0;
        while (ld_acquire_sys(F.f[rank] + b * world + p) < epoch) {
            __nanosleep(40);
            if ((++spins & 0xfffu) == 0 && peer_now_ns() - t0 > 4000000000ull) {
                printf("mmrec: peer barrier %d timed out (rank %d waiting for rank %d, epoch %d)\n", b, rank, p, epoch);
                __trap();
            }
        }
    }
}
// kernel prologue: block 0 runs barrier 0 and releases the other blocks of this GPU; returns the call number
__device__ __forceinline__ int peer_enter(const PeerFlags& F, int* state, int rank, int world) {
    __shared__ int e_sh;
    if (threadIdx.x == 0) {
        const int e = state[0] + 1;                                  // (state[0] was stored by the previous call's last block)
        if (blockIdx.x == 0) {
            peer_barrier(F, 0, rank, world, e);
            st_release_gpu(state + 1, e);
        } else {
            const unsigned long long t0 = peer_now_ns();
            unsigned spins = 0;
            while (ld_acquire_gpu(state + 1) < e) {
                __nanosleep(200);                                    // (hundreds of blocks poll this word: keep them off the L2 port)
                if ((++spins & 0xfffu) == 0 && peer_now_ns() - t0 > 4000000000ull) { printf("mmrec: peer_enter timed out\n"); __trap(); }
            }
        }
        e_sh = e;
    }
    __syncthreads();
    return e_sh;
}
// kernel epilogue: the last block to finish closes the call (barrier 1 when `with_barrier`: every rank's stores have landed)
__device__ __forceinline__ void peer_leave(const PeerFlags& F, int* state, int rank, int world, int epoch, bool with_barrier) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(state + 2, 1);
        if (done == (int)gridDim.x - 1) {
            state[2] = 0;
            if (with_barrier) peer_barrier(F, 1, rank, world, epoch);
            state[0] = epoch;
            __threadfence();
        }
    }
}

}  // namespace mmrec
