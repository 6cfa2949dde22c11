// K3 on the tensor cores: S = U I^T with tcgen05.mma (kind::tf32) and the 3xTF32 operand split
//     u = u_hi + u_lo,  i = i_hi + i_lo   (each part exactly representable in tf32)
//     <u, i> ~= u_hi.i_hi + u_lo.i_hi + u_hi.i_lo          (dropped term u_lo.i_lo ~ 2^-22 relative)
// accumulated in fp32 in TMEM: fp32-level accuracy (the reference computes this GEMM in fp32 and the top-k
// order is not stable under single-pass TF32) at tensor-pipe speed.
//
// Data path
//   pack kernels   split + re-tile both operands ONCE into the UMMA canonical K-major no-swizzle layout
//                  (8 x 16 B core matrices), so a tile slab is a contiguous byte range in global memory;
//   producer warp  cp.async.bulk (UBLKCP) global -> shared, completion on mbarriers; the 128-user tile (hi+lo)
//                  stays resident, item tiles stream as 32 KB slabs (256 items x 32 k) through a 4-slot ring;
//   MMA warp       one thread issues tcgen05.mma M128 N256 K8 into one of two 256-column TMEM accumulators and
//                  frees ring slots / publishes accumulators with tcgen05.commit;
//   epilogue       4 warps (one per TMEM lane quarter) read the accumulator with tcgen05.ld 32x32b.x32 and write
//                  the scores (this file) or filter them into per-user candidate lists (score_fused.cu).
#include "tc_common.cuh"

namespace mmrec {

using namespace tc;

struct ScoreTcParams {
    const float *Uhi, *Ulo, *Ihi, *Ilo;
    int KP, n_itiles, tiles_per_split, n_splits;
    int64_t B, n_items;
    float* S;
    int64_t ldS;
};

struct TcSmemLayout {
    uint32_t u_hi, u_lo, slab0, bars, tmem_ptr, total;
    int stages;
};
__host__ __device__ inline TcSmemLayout tc_smem_layout(int KP) {
    TcSmemLayout L;
    const uint32_t u_bytes = TC_M * KP * 4;
    L.stages = KP >= 128 ? 3 : 4;
    L.u_hi = 0; L.u_lo = u_bytes; L.slab0 = 2 * u_bytes;
    L.bars = L.slab0 + L.stages * TC_SLAB_BYTES;
    L.tmem_ptr = L.bars + 16 * 8;
    L.total = L.tmem_ptr + 16;
    return L;
}
// barrier slots: 0 u_full | 1..4 full[s] | 5..8 empty[s] | 9,10 tmem_full[b] | 11,12 tmem_empty[b]

__device__ __forceinline__ void tc_producer(const ScoreTcParams& p, const TcSmemLayout& L, uint32_t sbase, int ut, int it0, int it1) {
    const uint32_t bar = sbase + L.bars;
    const uint32_t u_bytes = TC_M * p.KP * 4;
    mbar_expect_tx(bar + 0 * 8, 2 * u_bytes);
    for (uint32_t o = 0; o < u_bytes; o += 16384) {
        bulk_g2s(sbase + L.u_hi + o, (const char*)(p.Uhi + (int64_t)ut * TC_M * p.KP) + o, 16384, bar);
        bulk_g2s(sbase + L.u_lo + o, (const char*)(p.Ulo + (int64_t)ut * TC_M * p.KP) + o, 16384, bar);
    }
    const int kchunks = p.KP / TC_KC;
    uint32_t s = 0;
    for (int it = it0; it < it1; ++it) {
        for (int c = 0; c < 2 * kchunks; ++c, ++s) {
            const uint32_t slot = s % L.stages, use = s / L.stages;
            mbar_wait(bar + (5 + slot) * 8, (use & 1) ^ 1);
            mbar_expect_tx(bar + (1 + slot) * 8, TC_SLAB_BYTES);
            const float* src = (c < kchunks ? p.Ihi : p.Ilo) + ((int64_t)it * p.KP / 4 + (c % kchunks) * (TC_KC / 4)) * (TC_N / 8) * 32;
            bulk_g2s(sbase + L.slab0 + slot * TC_SLAB_BYTES, src, TC_SLAB_BYTES, bar + (1 + slot) * 8);
        }
    }
}

__device__ __forceinline__ void tc_mma_issuer(const ScoreTcParams& p, const TcSmemLayout& L, uint32_t sbase, uint32_t tmem_base,
                                              int it0, int it1) {
    const uint32_t bar = sbase + L.bars;
    constexpr uint32_t LBO_A = (TC_M / 8) * 128, LBO_B = (TC_N / 8) * 128, SBO = 128;
    const uint32_t idesc = idesc_tf32(TC_M, TC_N);
    const int kchunks = p.KP / TC_KC;
    mbar_wait(bar + 0 * 8, 0);
    fence_after_sync();
    // descriptors are built once; per instruction only the address field moves (see tools/probe_mma.py: rebuilding them
    // around every MMA makes the issuing thread, not the tensor pipe, the limit)
    const uint64_t a_hi0 = smem_desc(sbase + L.u_hi, LBO_A, SBO), a_lo0 = smem_desc(sbase + L.u_lo, LBO_A, SBO);
    const uint64_t b0 = smem_desc(sbase + L.slab0, LBO_B, SBO);
    const uint64_t ka = (2 * LBO_A) >> 4, kb = (2 * LBO_B) >> 4, slab_step = TC_SLAB_BYTES >> 4;
    uint32_t slot = 0, ph = 0;
    uint64_t bd_slot = b0;
    for (int it = it0, t = 0; it < it1; ++it, ++t) {
        const uint32_t buf = t & 1;
        mbar_wait(bar + (11 + buf) * 8, ((t >> 1) & 1) ^ 1);     // accumulator drained by the epilogue
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * TC_N;
        uint32_t acc = 0;
        uint64_t a_hi = a_hi0, a_lo = a_lo0;
        for (int c = 0; c < 2 * kchunks; ++c) {
            mbar_wait(bar + (1 + slot) * 8, ph);
            fence_after_sync();
            const bool item_lo = c >= kchunks;
            if (c == kchunks) { a_hi = a_hi0; a_lo = a_lo0; }     // second sweep over K: the items' lo parts
            uint64_t bd = bd_slot;
#pragma unroll
            for (int j = 0; j < TC_KC / 8; ++j) {
                mma_tf32(d_tmem, a_hi, bd, idesc, acc);
                acc = 1;
                if (!item_lo) mma_tf32(d_tmem, a_lo, bd, idesc, 1);
                a_hi += ka; a_lo += ka; bd += kb;
            }
            mma_commit(bar + (5 + slot) * 8);                     // slab consumed -> slot back to the producer
            bd_slot += slab_step;
            if (++slot == (uint32_t)L.stages) { slot = 0; ph ^= 1; bd_slot = b0; }
        }
        mma_commit(bar + (9 + buf) * 8);                          // accumulator complete -> epilogue
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1) score_tc_kernel(const ScoreTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const TcSmemLayout L = tc_smem_layout(p.KP);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar = sbase + L.bars;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 11; ++i) mbar_init(bar + i * 8, 1);
        mbar_init(bar + 11 * 8, 128); mbar_init(bar + 12 * 8, 128);
        mbar_fence_init();
    }
    if (warp == 1) { tmem_alloc(sbase + L.tmem_ptr, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L.tmem_ptr);

    const int ut = blockIdx.x / p.n_splits, sp = blockIdx.x % p.n_splits;
    const int it0 = sp * p.tiles_per_split;
    const int it1 = min(p.n_itiles, it0 + p.tiles_per_split);

    if (warp == 0) {
        if (lane == 0 && it0 < it1) tc_producer(p, L, sbase, ut, it0, it1);
    } else if (warp == 1) {
        if (lane == 0 && it0 < it1) tc_mma_issuer(p, L, sbase, tmem_base, it0, it1);
    } else {
        const int q = warp & 3;                                   // TMEM lane quarter this warp may read
        const int64_t row = (int64_t)ut * TC_M + q * 32 + lane;
        const bool vec = ((p.ldS & 3) == 0) && ((((uintptr_t)p.S) & 15) == 0);
        for (int it = it0, t = 0; it < it1; ++it, ++t) {
            const uint32_t buf = t & 1;
            mbar_wait(bar + (9 + buf) * 8, (t >> 1) & 1);
            fence_after_sync();
#pragma unroll 1
            for (int c8 = 0; c8 < TC_N / 32; ++c8) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * TC_N + c8 * 32, v);
                tmem_ld_wait();
                const int64_t col0 = (int64_t)it * TC_N + c8 * 32;
                if (row < p.B && col0 < p.n_items) {
                    float* dst = p.S + row * p.ldS + col0;
                    if (vec && col0 + 32 <= p.n_items) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            reinterpret_cast<float4*>(dst)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.n_items) dst[j] = __uint_as_float(v[j]);
                    }
                }
            }
            fence_before_sync();
            mbar_arrive(bar + (11 + buf) * 8);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static inline int kp_of(int d) { return d <= 32 ? 32 : (d <= 64 ? 64 : 128); }

size_t score_tc_workspace_bytes(int64_t B, int64_t n_items, int d) {
    if (d > 128 || B <= 0 || n_items <= 0) return 0;
    const int KP = kp_of(d);
    const int64_t ut = (B + TC_M - 1) / TC_M, it = (n_items + TC_N - 1) / TC_N;
    return (size_t)(2 * ut * TC_M + 2 * it * TC_N) * KP * sizeof(float) + 1024;
}

// returns 1 = done, 0 = shape not supported by this path (caller falls back to the fp32 CUDA-core kernel)
int score_tc(int64_t B, const int64_t* users, const float* Ue, int64_t ldu, int64_t n_items, const float* Ie, int64_t ldi,
             int d, float* S, int64_t ldS, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (d > 128 || !ws || ws_bytes < score_tc_workspace_bytes(B, n_items, d)) return 0;
    const int KP = kp_of(d);
    const int64_t n_ut = (B + TC_M - 1) / TC_M, n_it = (n_items + TC_N - 1) / TC_N;
    float* base = (float*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    float* Uhi = base;
    float* Ulo = Uhi + n_ut * TC_M * KP;
    float* Ihi = Ulo + n_ut * TC_M * KP;
    float* Ilo = Ihi + n_it * TC_N * KP;
    {
        const int64_t tu = n_ut * TC_M * (KP / 4), ti = n_it * TC_N * (KP / 4);
        pack_split_kernel<TC_M><<<(unsigned)((tu + 255) / 256), 256, 0, stream>>>(B, users, Ue, ldu, d, KP, Uhi, Ulo, n_ut);
        MMREC_LAUNCH_CHECK();
        pack_split_kernel<TC_N><<<(unsigned)((ti + 255) / 256), 256, 0, stream>>>(n_items, nullptr, Ie, ldi, d, KP, Ihi, Ilo, n_it);
        MMREC_LAUNCH_CHECK();
    }
    ScoreTcParams p;
    p.Uhi = Uhi; p.Ulo = Ulo; p.Ihi = Ihi; p.Ilo = Ilo; p.KP = KP; p.n_itiles = (int)n_it;
    p.B = B; p.n_items = n_items; p.S = S; p.ldS = ldS;
    // fill the machine: user tiles x item splits ~ a multiple of the SM count
    const int sms = sm_count();
    int splits = (int)(sms / n_ut);                // user tiles x item splits <= SM count: one wave, no tail
    if (splits > n_it) splits = (int)n_it;
    if (splits < 1) splits = 1;
    p.tiles_per_split = (int)((n_it + splits - 1) / splits);
    p.n_splits = (int)((n_it + p.tiles_per_split - 1) / p.tiles_per_split);
    const TcSmemLayout L = tc_smem_layout(KP);
    {   // the opt-in is per device
        static bool attr_set[64] = {false};
        int dev = 0;
        MMREC_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            MMREC_CUDA(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    score_tc_kernel<<<(unsigned)(n_ut * p.n_splits), TC_THREADS, L.total, stream>>>(p);
    MMREC_LAUNCH_CHECK();
    return 1;
}

}  // namespace mmrec

// ---- measurement aid: how many cycles does one tcgen05.mma (cta_group::1, both operands in shared memory, K-major, no swizzle
// -- the form every tensor-core kernel of this library issues) take when nothing else runs?  One CTA per SM, one thread
// issues `iters` MMAs of M = 128, the given N and K = 32 bytes (8 tf32 / 16 bf16) into one accumulator, commits and waits.
namespace mmrec {
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int kind, int N, int iters, int distinct, long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_ptr_sm;
    __shared__ __align__(8) unsigned long long bar_sm;
    const uint32_t sbase = tc::smem_u32(smem);
    const uint32_t bar = tc::smem_u32(&bar_sm);
    for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;   // 192 KB of zeros
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
    if (threadIdx.x < 32) { tc::tmem_alloc(tc::smem_u32(&tmem_ptr_sm), 512); tc::tmem_relinquish(); }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_ptr_sm;
    if (threadIdx.x == 0) {
        const uint32_t LBO_A = (128 / 8) * 128, LBO_B = (uint32_t)(N / 8) * 128;
        // instruction descriptors: tf32 (a/b format 2) or bf16 (format 1), fp32 accumulate, K-major both
        const uint32_t idesc = kind == 0 ? tc::idesc_tf32(128, N)
                                         : tc::idesc_bf16(128, N);
        const long long t0 = clock64();
        if (distinct == 0) {
            // lean issue loop: descriptors built once, only the address field moves (8 K steps, unrolled)
            const uint64_t a0 = tc::smem_desc(sbase, LBO_A, 128), b0 = tc::smem_desc(sbase + 65536, LBO_B, 128);
            const uint64_t ka = (2 * LBO_A) >> 4, kb = (2 * LBO_B) >> 4;
            for (int i = 0; i < iters; i += 8) {
                uint64_t ad = a0, bd = b0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (kind == 0) tc::mma_tf32(tmem, ad, bd, idesc, (i | j) ? 1u : 0u);
                    else tc::mma_f16(tmem, ad, bd, idesc, (i | j) ? 1u : 0u);
                    ad += ka; bd += kb;
                }
            }
        } else
        for (int i = 0; i < iters; ++i) {
            const uint32_t off = (uint32_t)(i % distinct) * 2 * LBO_A;            // walk over `distinct` K steps of the operands
            const uint64_t ad = tc::smem_desc(sbase + off, LBO_A, 128);
            const uint64_t bd = tc::smem_desc(sbase + 65536 + (uint32_t)(i % distinct) * 2 * LBO_B, LBO_B, 128);
            if (kind == 0) tc::mma_tf32(tmem, ad, bd, idesc, i > 0);
            else tc::mma_f16(tmem, ad, bd, idesc, i > 0);
        }
        tc::mma_commit(bar);
        tc::mbar_wait(bar, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}
}  // namespace mmrec

extern "C" int mmrec_debug_mma_rate(int kind, int N, int iters, int distinct, long long* cycles_per_cta, void* stream_) {
    MMREC_CHECK_ARG((kind == 0 || kind == 1) && (N == 64 || N == 128 || N == 256) && iters >= 1 && distinct >= 0 && distinct <= 8 && cycles_per_cta,
                    "mma_rate: kind 0 (tf32) | 1 (bf16), N in {64,128,256}, 1 <= distinct <= 8");
    MMREC_CUDA(cudaFuncSetAttribute(mmrec::mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    mmrec::mma_rate_kernel<<<mmrec::sm_count(), 128, 192 * 1024, (cudaStream_t)stream_>>>(kind, N, iters, distinct, cycles_per_cta);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}
