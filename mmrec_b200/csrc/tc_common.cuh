// Thin inline-PTX layer for the sm_100a tensor-core kernels: mbarrier, bulk async copy (UBLKCP), tcgen05
// alloc / mma / commit / ld, UMMA shared-memory and instruction descriptors.  Encodings cross-checked against
// the CUTLASS 4.x headers (cute/arch/mma_sm100_desc.hpp, mma_sm100_umma.hpp, copy_sm100.hpp) -- the kernels
// themselves are hand-written.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace mmrec {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap within ~2 s of wall time, not hang the GPU (the box is shared and a hang
// is a strike).  The clock is only consulted after the fast path failed a few thousand times.
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins == 4096u) t0 = global_ns();
        if (spins > 4096u && (spins & 1023u) == 0 && global_ns() - t0 > 2000000000ull) {
            printf("mmrec: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}

// ---- bulk async copy global -> shared (non-tensor TMA, SASS UBLKCP), completion on an mbarrier ---------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- tcgen05 ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// the same with kind::f16 (fp16 or bf16 operands as the instruction descriptor says), K = 16 per instruction
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of fp32: thread i of the warp gets TMEM lane (base_lane + i), columns [col, col + 32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}

// ---- descriptors --------------------------------------------------------------------------------------------
// K-major, no swizzle ("interleave") canonical layout, in 16-byte units ((8,n),2):((1,SBO),LBO):
// a core matrix is 8 rows x 16 B stored as 128 contiguous bytes; SBO = byte distance between core matrices that
// are neighbours along M/N (next 8 rows), LBO = between neighbours along K (next 16 B of K).
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;                    // descriptor version: Blackwell
    return d;                           // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// instruction descriptor, kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// instruction descriptor, kind::f16 with fp16 operands (format 0; bf16 would be 1), fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return idesc_f16(M, N) | (1u << 7) | (1u << 10); }

// 3xTF32 split: hi = x rounded to tf32 (RN, ties away), lo = (x - hi) rounded to tf32; x = hi + lo to ~2^-22 |x|
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
    lo = __uint_as_float(l);
}

// Cheaper split for streamed operands (2 instructions instead of 9): hi = x with the low 13 mantissa bits cleared, lo = x - hi
// exactly (< 2^-10 |x|); the tensor core reads the top 19 bits of lo.  x = hi + lo to ~2^-20 |x|: enough for the 1e-4
// bar of the projections, not for ranking scores (those use split_tf32).
__device__ __forceinline__ void split_tf32_trunc(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

}  // namespace tc

// ---- tile geometry shared by the tensor-core kernels ---------------------------------------------------------
constexpr int TC_M = 128;          // users per CTA tile = TMEM lanes
constexpr int TC_N = 256;          // items per accumulator = TMEM columns
constexpr int TC_KC = 32;          // k per slab
constexpr int TC_SLAB_BYTES = TC_N * TC_KC * 4;    // 32 KB
constexpr int TC_THREADS = 192;    // warp 0 producer, warp 1 MMA, warps 2..5 epilogue

// operand packing: split into tf32 hi/lo and re-tile as [tile][kblk = KP/4][row group = R/8][8 rows][4 floats], the
// UMMA canonical K-major no-swizzle layout, so that a K slab of a tile is one contiguous byte range
template <int R>
__device__ __forceinline__ void pack_split_one(int64_t t, int64_t n_rows, const int64_t* __restrict__ idx, const float* __restrict__ E,
                                               int64_t ld, int d, int KP, float* __restrict__ hi, float* __restrict__ lo) {
    const int kblks = KP / 4;                                          // t = (padded row, kblk)
    const int64_t row = t / kblks;
    const int kb = (int)(t % kblks);
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (row < n_rows) {
        const float* src = E + (idx ? idx[row] : row) * ld;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (kb * 4 + e < d) x[e] = __ldg(src + kb * 4 + e);
    }
    float4 h, l;
    tc::split_tf32(x[0], h.x, l.x); tc::split_tf32(x[1], h.y, l.y); tc::split_tf32(x[2], h.z, l.z); tc::split_tf32(x[3], h.w, l.w);
    const int64_t tile = row / R;
    const int rr = (int)(row % R);
    const int64_t off = ((tile * kblks + kb) * (R / 8) + rr / 8) * 32 + (rr % 8) * 4;
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
}

template <int R>
__global__ void pack_split_kernel(int64_t n_rows, const int64_t* __restrict__ idx, const float* __restrict__ E, int64_t ld,
                                  int d, int KP, float* __restrict__ hi, float* __restrict__ lo, int64_t n_tiles) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // one thread per (padded row, kblk)
    if (t >= n_tiles * R * (KP / 4)) return;
    pack_split_one<R>(t, n_rows, idx, E, ld, d, KP, hi, lo);
}

}  // namespace mmrec
