// K3 fused: scores on tcgen05 (3xTF32, same pipeline as score_tc.cu) whose epilogue never writes the score matrix.
//
//   fz_prep_kernel        one launch: split + re-tile the operands, zero the per-block scratch, and give every row a seed
//                         threshold + histogram window from a fixed 128-item sample of the catalogue (deterministic);
//   mask_csr_*_kernel     the batch's (row, item) mask as a CSR over batch rows (one pass when the rows arrive sorted);
//   score_fused_kernel    each epilogue thread owns one user row of the 128 x 256 accumulator tile (4 sets of 4 warps: a
//                         column half of one TMEM buffer each) and streams its items through a threshold filter: every
//                         value >= thr is appended to the row's candidate list and counted in the row's histogram (32
//                         linear bins in L2, shared by all item splits of the row); after each tile the threshold rises
//                         to the highest bin edge that still has `need` = k + (masked items of the row) candidates above
//                         it.  Masked train positives are NOT removed here: asking for k + m candidates guarantees k
//                         unmasked ones survive, so the mask is applied once, on the ~100 finalists, instead of on 7,000
//                         scores;
//   fused_select_kernel   per row (one warp): the finalists (value >= the row's final threshold) of all item splits,
//                         masked items dropped, top-k on (value desc, item asc) -- the contract of topk.cu.
// Anything the filter cannot certify (candidate overflow, a degenerate sample, NaNs, fewer than k finalists) raises a
// per-row flag and the row is recomputed by the exact fp32 kernels at the end of this file; no host round trip.
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cub/device/device_scan.cuh>

#include "tc_common.cuh"

namespace mmrec {

using namespace tc;

constexpr int FZ_NB = 32;               // histogram bins per row (two 16-bit counters per shared-memory word)
constexpr int FZ_CH = 16;               // accumulator columns an epilogue warp reads per tcgen05.ld
constexpr int FZ_HALF = TC_N / 2;       // columns of a tile owned by one epilogue set
constexpr int FZ_MAX_THREADS = 64 + 128 * 4;

struct FusedParams {
    const float *Uhi, *Ulo, *Ihi, *Ilo;
    int KP, n_itiles, tiles_per_split, n_splits, cap, k;
    int64_t B, n_items;
    const int32_t* mask_ptr;            // [B+1] CSR over batch rows (may be null = no mask)
    float2* cand;                       // [B][n_splits][cap]
    int32_t* cnt;                       // [B][n_splits]
    float* thr;                         // [B][n_splits]
    int32_t* flags;                     // [B] fallback flags
    uint32_t* gthr;                     // [B] highest certified threshold of any split (order-preserving key), zeroed per launch
    unsigned long long* gbins;          // [B] (lo, width) of the row's histogram window and seed threshold, from fz_prep_kernel (0 = degenerate)
    uint32_t* ghist;                    // [B][FZ_NB / 2] candidates counted per bin by ALL splits of the row (two 16-bit counters per word)
};

// Shared memory: user tile (hi | lo) | item slab ring | row state | histogram | per-warp chunk scratch | barriers.
// Epilogue sets (4 warps = 128 rows each): a set owns one column half of the tiles of one parity.  KP <= 64 runs four
// sets (both halves x both TMEM buffers in flight), KP = 128 has room for two (both halves, tiles in turn).
struct FzSmem {
    uint32_t u_hi, u_lo, slab0, rowst, hist, scratch, bars, tmem_ptr, total;
    int stages, parities, nsets;
};
__host__ __device__ inline FzSmem fz_smem(int KP) {
    FzSmem L;
    const uint32_t u_bytes = TC_M * KP * 4;
    L.stages = KP >= 128 ? 2 : 3;
    L.parities = KP >= 128 ? 1 : 2;
    L.nsets = 2 * L.parities;
    L.u_hi = 0; L.u_lo = u_bytes; L.slab0 = 2 * u_bytes;
    L.rowst = L.slab0 + L.stages * TC_SLAB_BYTES;                    // 8 words per row, field-major
    L.hist = L.rowst + 8 * TC_M * 4;                                 // (unused: the histogram is row-global, in L2)
    L.scratch = L.hist;
    L.bars = L.scratch + L.nsets * 4 * (FZ_CH * 32 * 4);
    L.tmem_ptr = L.bars + 16 * 8;
    L.total = L.tmem_ptr + 16;
    return L;
}
enum { RS_THR = 0, RS_CNT = 1, RS_LO = 2, RS_SCALE = 3, RS_WIDTH = 4, RS_BAD = 5 };

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void epi_bar_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// producer / MMA issuer: identical protocol to score_tc.cu (barrier slots: 0 u_full | 1..4 full | 5..8 empty |
// 9,10 tmem_full | 11,12 tmem_empty)
__device__ __forceinline__ void fz_producer(const FusedParams& p, const FzSmem& L, uint32_t sbase, int ut, int it0, int it1) {
    const uint32_t bar = sbase + L.bars;
    const uint32_t u_bytes = TC_M * p.KP * 4;
    mbar_expect_tx(bar, 2 * u_bytes);
    for (uint32_t o = 0; o < u_bytes; o += 16384) {
        bulk_g2s(sbase + L.u_hi + o, (const char*)(p.Uhi + (int64_t)ut * TC_M * p.KP) + o, 16384, bar);
        bulk_g2s(sbase + L.u_lo + o, (const char*)(p.Ulo + (int64_t)ut * TC_M * p.KP) + o, 16384, bar);
    }
    const int kchunks = p.KP / TC_KC;
    uint32_t s = 0;
    for (int it = it0; it < it1; ++it)
        for (int c = 0; c < 2 * kchunks; ++c, ++s) {
            const uint32_t slot = s % L.stages, use = s / L.stages;
            mbar_wait(bar + (5 + slot) * 8, (use & 1) ^ 1);
            mbar_expect_tx(bar + (1 + slot) * 8, TC_SLAB_BYTES);
            const float* src = (c < kchunks ? p.Ihi : p.Ilo) + ((int64_t)it * p.KP / 4 + (c % kchunks) * (TC_KC / 4)) * (TC_N / 8) * 32;
            bulk_g2s(sbase + L.slab0 + slot * TC_SLAB_BYTES, src, TC_SLAB_BYTES, bar + (1 + slot) * 8);
        }
}

__device__ __forceinline__ void fz_mma(const FusedParams& p, const FzSmem& L, uint32_t sbase, uint32_t tmem_base, int it0, int it1) {
    const uint32_t bar = sbase + L.bars;
    constexpr uint32_t LBO_A = (TC_M / 8) * 128, LBO_B = (TC_N / 8) * 128, SBO = 128;
    const uint32_t idesc = idesc_tf32(TC_M, TC_N);
    const int kchunks = p.KP / TC_KC;
    mbar_wait(bar, 0);
    fence_after_sync();
    uint32_t s = 0;
    for (int it = it0, t = 0; it < it1; ++it, ++t) {
        const uint32_t buf = t & 1;
        mbar_wait(bar + (11 + buf) * 8, ((t >> 1) & 1) ^ 1);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * TC_N;
        uint32_t acc = 0;
        for (int c = 0; c < 2 * kchunks; ++c, ++s) {
            const uint32_t slot = s % L.stages, use = s / L.stages;
            mbar_wait(bar + (1 + slot) * 8, use & 1);
            fence_after_sync();
            const bool item_lo = c >= kchunks;
            const int kc = c % kchunks;
            const uint32_t b_base = sbase + L.slab0 + slot * TC_SLAB_BYTES;
#pragma unroll
            for (int j = 0; j < TC_KC / 8; ++j) {
                const uint64_t bd = smem_desc(b_base + j * 2 * LBO_B, LBO_B, SBO);
                const uint32_t a_off = (kc * (TC_KC / 4) + 2 * j) * LBO_A;
                mma_tf32(d_tmem, smem_desc(sbase + L.u_hi + a_off, LBO_A, SBO), bd, idesc, acc);
                acc = 1;
                if (!item_lo) mma_tf32(d_tmem, smem_desc(sbase + L.u_lo + a_off, LBO_A, SBO), bd, idesc, 1);
            }
            mma_commit(bar + (5 + slot) * 8);
        }
        mma_commit(bar + (9 + buf) * 8);
    }
}

// Epilogue.  The four lane quarters of a set cover the 128 rows of the tile; thread (set, row) filters its set's column
// half of its set's tiles against the ROW's threshold, which all sets of the CTA share in shared memory together with
// the candidate count and the histogram (shared-memory atomics: hits are rare).  Certified thresholds are also
// published per row in global memory, so the other item splits of the row stop collecting what cannot make the top-k.
__device__ __forceinline__ void fz_epilogue(const FusedParams& p, const FzSmem& L, uint8_t* smem, uint32_t sbase, uint32_t tmem_base,
                                            int ut, int sp, int it0, int it1, int warp, int lane) {
    const uint32_t bar = sbase + L.bars;
    const int e = warp - 2;
    const int set = e >> 2;
    const int q = warp & 3;                                          // the TMEM lane quarter this warp can read
    const int half = set & 1, par = set >> 1;
    const int rl = q * 32 + lane;                                    // row inside the tile = TMEM lane
    const int64_t row = (int64_t)ut * TC_M + rl;
    const bool live = row < p.B;
    const int n_epi = 128 * L.nsets;
    const int n_tiles = it1 - it0;
    uint32_t* rs = reinterpret_cast<uint32_t*>(smem + L.rowst) + rl;           // field f of this row: rs[f * TC_M]
    uint32_t* ghist = p.ghist + (live ? row : 0) * (FZ_NB / 2);                // the row's bins, shared with the other splits
    float* sc = reinterpret_cast<float*>(smem + L.scratch) + e * (FZ_CH * 32) + lane;   // value j of the chunk: sc[j * 32]
    int need = p.k;
    if (live && p.mask_ptr) need += p.mask_ptr[row + 1] - p.mask_ptr[row];
    float2* cand = p.cand + ((int64_t)(live ? row : 0) * p.n_splits + sp) * p.cap;
    const int n_items32 = (int)p.n_items;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);

    if (set == 0) {
        // The row's histogram window (lo, width) and seed threshold were computed by fz_prep_kernel from a fixed sample of
        // the catalogue (see there): the same for every split of the row, the same in every run.
        const unsigned long long w = live ? p.gbins[row] : 0ull;
        float lo = __uint_as_float((uint32_t)(w >> 32)), width = __uint_as_float((uint32_t)w), thr = lo;
        uint32_t bad = 0;
        if (w == 0ull) { bad = live ? 1u : 0u; width = 1.f; lo = 0.f; thr = INFINITY; }   // degenerate sample (or padding row)
        rs[RS_THR * TC_M] = float_key(thr);
        rs[RS_CNT * TC_M] = 0;
        rs[RS_LO * TC_M] = __float_as_uint(lo);
        rs[RS_SCALE * TC_M] = __float_as_uint(1.0f / width);
        rs[RS_WIDTH * TC_M] = __float_as_uint(width);
        rs[RS_BAD * TC_M] = bad;
    }
    epi_bar_sync(n_epi);                                             // row state visible to every set
    const float lo = __uint_as_float(rs[RS_LO * TC_M]), scale = __uint_as_float(rs[RS_SCALE * TC_M]);
    const float width = __uint_as_float(rs[RS_WIDTH * TC_M]);
    volatile uint32_t* thr_key = rs + RS_THR * TC_M;

    for (int t = par; t < n_tiles; t += L.parities) {
        const uint32_t buf = t & 1;
        mbar_wait(bar + (9 + buf) * 8, (t >> 1) & 1);
        fence_after_sync();
        const uint32_t foreign = live ? ld_relaxed_u32(p.gthr + row) : 0u;   // consumed after the tile
        // snapshot of the row's bin counters, requested two chunks before the end of the tile and read after it (a stale
        // snapshot is still a lower bound); one of the two sets that share a tile does the walk, the other picks the threshold up from shared memory
        uint4 hw[FZ_NB / 8];
        const uint32_t tbase = lane_base + buf * TC_N + half * FZ_HALF;
        const int col_base = (it0 + t) * TC_N + half * FZ_HALF;
        const int n_valid = n_items32 - col_base;                    // columns of this half tile inside the catalogue
#pragma unroll 1
        for (int c = 0; c < FZ_HALF / FZ_CH; ++c) {
            uint32_t v[16];
            tmem_ld_32x16(tbase + c * FZ_CH, v);
            if (c == FZ_HALF / FZ_CH - 2 && half == 0) {
#pragma unroll
                for (int i = 0; i < FZ_NB / 8; ++i) hw[i] = __ldcg(reinterpret_cast<const uint4*>(ghist) + i);
            }
            tmem_ld_wait();
            if (n_valid < FZ_HALF) {                                 // last, partial tile only
#pragma unroll
                for (int j = 0; j < FZ_CH; ++j)
                    if (c * FZ_CH + j >= n_valid) v[j] = 0xff800000u;   // -inf
            }
            const float thr = key_float(*thr_key);
            // Branch-free filter: one bit per value.  Rows (= lanes) hit at different columns, so a per-value branch
            // would run its body for almost every column; instead the hits are walked per lane afterwards.
            uint32_t hits = 0;
#pragma unroll
            for (int j = 0; j < FZ_CH; ++j) hits |= (__uint_as_float(v[j]) >= thr ? 1u : 0u) << j;
            if (__any_sync(0xffffffffu, hits != 0)) {
#pragma unroll
                for (int j = 0; j < FZ_CH; ++j) sc[j * 32] = __uint_as_float(v[j]);   // own column only: no sync needed
                const int nmax = __reduce_max_sync(0xffffffffu, __popc(hits));
#pragma unroll 2
                for (int h = 0; h < nmax; ++h) {                     // warp-uniform trip count, predicated body
                    const bool on = hits != 0;
                    const int j = on ? __ffs(hits) - 1 : 0;
                    hits &= hits - 1;
                    const float x = sc[j * 32];
                    if (on) {
                        const uint32_t pos = atomicAdd(rs + RS_CNT * TC_M, 1u);
                        if (pos < (uint32_t)p.cap) {
                            cand[pos] = make_float2(x, __int_as_float(col_base + c * FZ_CH + j));
                            int b = (int)((x - lo) * scale);
                            b = b < 0 ? 0 : (b > FZ_NB - 1 ? FZ_NB - 1 : b);
                            atomicAdd(ghist + (b >> 1), (b & 1) ? 65536u : 1u);          // after the append: counted => listed
                        }
                    }
                }
            }
        }
        // accumulator drained: hand the TMEM buffer back before the (SMEM-only) threshold update
        fence_before_sync();
        mbar_arrive(bar + (11 + buf) * 8);
        // Raise the row threshold to the highest bin edge that keeps `need` counted candidates above it -- counted over the
        // whole catalogue seen so far, by every split of the row.  The counters are moving: any snapshot is a lower
        // bound, so what it certifies stays certified.
        int cnum = 0, b = -1;
        if (half == 0) {
#pragma unroll
            for (int i = FZ_NB / 8 - 1; i >= 0; --i) {
                const uint32_t w4[4] = {hw[i].x, hw[i].y, hw[i].z, hw[i].w};
#pragma unroll
                for (int j = 3; j >= 0; --j) {
#pragma unroll
                    for (int h2 = 1; h2 >= 0; --h2) {
                        if (b < 0) {
                            cnum += (int)(h2 ? (w4[j] >> 16) : (w4[j] & 0xffffu));
                            if (cnum >= need) b = (i * 4 + j) * 2 + h2;
                        }
                    }
                }
            }
        }
        uint32_t mine = 0;
        if (b >= 0 && live) {
            float cert = lo;                                         // b == 0: everything counted is >= thr0
            if (b > 0) {
                const float edge = lo + (float)b * width;
                cert = edge - 2e-6f * fmaxf(fmaxf(fabsf(edge), fabsf(lo)), width);   // slack for the bin index rounding
            }
            mine = float_key(cert);
            if (mine > *thr_key) {
                atomicMax(rs + RS_THR * TC_M, mine);
                atomicMax(p.gthr + row, mine);
            }
        }
        if (foreign > *thr_key) atomicMax(rs + RS_THR * TC_M, foreign);
    }
    epi_bar_sync(n_epi);                                             // every set is done with the row state
    if (set == 0 && live) {
        if (n_tiles > 0) {
            const uint32_t cnt = rs[RS_CNT * TC_M];
            p.cnt[row * p.n_splits + sp] = cnt < (uint32_t)p.cap ? (int32_t)cnt : p.cap;
            p.thr[row * p.n_splits + sp] = key_float(rs[RS_THR * TC_M]);
            // (a split that never collected `need` candidates keeps a low threshold; the select kernel certifies the
            // row globally, see there)
            if (rs[RS_BAD * TC_M]) atomicOr(p.flags + row, 1);                // degenerate seed (NaN / inf / constant scores)
            if (cnt > (uint32_t)p.cap) atomicOr(p.flags + row, 2);           // candidate list overflow
        } else {
            p.cnt[row * p.n_splits + sp] = 0;
            p.thr[row * p.n_splits + sp] = -INFINITY;
        }
    }
}

__global__ void __launch_bounds__(FZ_MAX_THREADS, 1) score_fused_kernel(const FusedParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const FzSmem L = fz_smem(p.KP);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar = sbase + L.bars;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 11; ++i) mbar_init(bar + i * 8, 1);
        mbar_init(bar + 11 * 8, 256); mbar_init(bar + 12 * 8, 256);   // both column halves release a TMEM buffer
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
        if (lane == 0 && it0 < it1) fz_producer(p, L, sbase, ut, it0, it1);
    } else if (warp == 1) {
        if (lane == 0 && it0 < it1) fz_mma(p, L, sbase, tmem_base, it0, it1);
    } else {
        fz_epilogue(p, L, smem, sbase, tmem_base, ut, sp, it0, it1, warp, lane);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------
// mask CSR over batch rows (input order is free)
// ------------------------------------------------------------------------------------------------------
__global__ void mask_count_kernel(int64_t nnz, const int64_t* __restrict__ rows, int64_t B, int32_t* __restrict__ counts) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < nnz && rows[j] >= 0 && rows[j] < B) atomicAdd(counts + rows[j], 1);
}
__global__ void mask_fill_kernel(int64_t nnz, const int64_t* __restrict__ rows, const int64_t* __restrict__ cols, int64_t B,
                                 int64_t item_offset, const int32_t* __restrict__ ptr, int32_t* __restrict__ cursor,
                                 int32_t* __restrict__ items) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= nnz || rows[j] < 0 || rows[j] >= B) return;
    const int pos = ptr[rows[j]] + atomicAdd(cursor + rows[j], 1);
    items[pos] = (int32_t)(cols[j] - item_offset);     // may fall outside [0, n_items): then it never matches
}

// Per-batch case (B <= MC_MAX_ROWS rows).  The reference's evaluation loader emits the mask row-major (batch row ascending:
// src/utils/dataloader.py:370-391 builds it user by user), so the common case is a sorted row array: mask_csr_sorted_kernel
// writes the row pointers from the positions where the row changes and checks the order as it goes, one fully parallel
// pass.  If any CTA saw a descent, mask_csr_small_kernel (one CTA: count in shared memory, scan, fill) redoes the job
// for arbitrary order; otherwise it exits at once.  No global atomics, no memsets, no library scan.
__global__ void __launch_bounds__(256) mask_csr_sorted_kernel(int64_t nnz, const int64_t* __restrict__ rows, const int64_t* __restrict__ cols,
                                                              int B, int64_t item_offset, int32_t* __restrict__ ptr,
                                                              int32_t* __restrict__ items, int32_t* __restrict__ unsorted) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // entry j, plus one sentinel thread j == nnz
    int bad = 0;
    if (j <= nnz) {
        const int64_t rj = j < nnz ? rows[j] : (int64_t)B;
        const int64_t rp = j > 0 ? rows[j - 1] : -1;
        bad = j < nnz && rp > rj;
        if (j < nnz) items[j] = (int32_t)(cols[j] - item_offset);
        // rows (rp, rj] start at entry j (rows outside [0, B) own no pointer; clamped so that they delimit correctly)
        const int64_t lo = rp < -1 ? -1 : (rp > B ? B : rp), hi = rj < -1 ? -1 : (rj > B ? B : rj);
        for (int64_t r = lo + 1; r <= hi; ++r) ptr[r] = (int32_t)j;
    }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) unsorted[blockIdx.x] = bad;
}


constexpr int MC_MAX_ROWS = 8192;
constexpr int MC_THREADS = 1024;
__global__ void __launch_bounds__(MC_THREADS) mask_csr_small_kernel(int64_t nnz, const int64_t* __restrict__ rows,
                                                                    const int64_t* __restrict__ cols, int B, int64_t item_offset,
                                                                    int32_t* __restrict__ ptr, int32_t* __restrict__ items,
                                                                    const int32_t* __restrict__ unsorted, int n_unsorted) {
    extern __shared__ int32_t mc_sm[];                               // count / cursor [B + 1] | warp totals [32]
    {   // runs only when mask_csr_sorted_kernel found the rows out of order (it then left garbage behind)
        int any = 0;
        for (int i = threadIdx.x; i < n_unsorted; i += MC_THREADS) any |= unsorted[i];
        if (!__syncthreads_or(any)) return;
    }
    int32_t* cnt = mc_sm;
    int32_t* wtot = mc_sm + B + 1;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int r = tid; r <= B; r += MC_THREADS) cnt[r] = 0;
    __syncthreads();
    constexpr int MC_U = 8;                                          // loads in flight per thread (the loop is latency-bound)
    for (int64_t jb = 0; jb < nnz; jb += (int64_t)MC_U * MC_THREADS) {        // warp-uniform trip count (match / shfl below)
        const int64_t j0 = jb + tid;
        int64_t r[MC_U];
#pragma unroll
        for (int u = 0; u < MC_U; ++u) {
            const int64_t j = j0 + (int64_t)u * MC_THREADS;
            r[u] = j < nnz ? __ldg(rows + j) : -1;
        }
#pragma unroll
        for (int u = 0; u < MC_U; ++u) {
            // batch rows arrive (mostly) sorted: the lanes of a warp hit a handful of counters.  One atomic per distinct
            // row of the warp (same-address shared atomics serialise a full round trip each).
            const int rr = (r[u] >= 0 && r[u] < B) ? (int)r[u] : -1;
            const unsigned peers = __match_any_sync(0xffffffffu, rr);
            if (rr >= 0 && lane == __ffs(peers) - 1) atomicAdd(cnt + rr, __popc(peers));
        }
    }
    __syncthreads();
    // exclusive scan of cnt[0..B]: each thread owns a contiguous run of rows
    const int per = (B + 1 + MC_THREADS - 1) / MC_THREADS;
    const int r0 = tid * per, r1 = min(B + 1, r0 + per);
    int local = 0;
    for (int r = r0; r < r1; ++r) local += cnt[r];
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wtot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int v = wtot[lane], sc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, sc, o);
            if (lane >= o) sc += u;
        }
        wtot[lane] = sc - v;                                         // exclusive warp offsets
    }
    __syncthreads();
    int run = wtot[wid] + incl - local;
    for (int r = r0; r < r1; ++r) {
        const int c = cnt[r];
        ptr[r] = run;
        cnt[r] = run;                                                // becomes the fill cursor
        run += c;
    }
    __syncthreads();
    for (int64_t jb = 0; jb < nnz; jb += (int64_t)MC_U * MC_THREADS) {        // warp-uniform trip count (match / shfl below)
        const int64_t j0 = jb + tid;
        int64_t r[MC_U], c[MC_U];
#pragma unroll
        for (int u = 0; u < MC_U; ++u) {
            const int64_t j = j0 + (int64_t)u * MC_THREADS;
            r[u] = j < nnz ? __ldg(rows + j) : -1;
            c[u] = j < nnz ? __ldg(cols + j) : 0;
        }
#pragma unroll
        for (int u = 0; u < MC_U; ++u) {
            const int rr = (r[u] >= 0 && r[u] < B) ? (int)r[u] : -1;
            const unsigned peers = __match_any_sync(0xffffffffu, rr);
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (rr >= 0 && lane == leader) base = atomicAdd(cnt + rr, __popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (rr >= 0) items[base + __popc(peers & ((1u << lane) - 1u))] = (int32_t)(c[u] - item_offset);   // order inside a row is free
        }
    }
}

// Seed threshold + histogram window of every row of the block, one warp per row: score the row against FZ_SAMPLE items
// spread evenly over the catalogue (fp32 on CUDA cores: 128 x d FMAs per row), take the maxima of 16 groups of 8, the r-th
// largest of them is the seed (about r / 128 of the catalogue passes; r aims at 2.5x the share the row needs, at least 8
// = 6 %), 16 bins from there to the sample maximum and 16 more above.  The seed is the one threshold nothing certifies,
// so it has to be safe -- the select kernel checks that k unmasked items clear it -- and it has to be the same whatever
// the timing: it depends on the inputs only.
constexpr int FZ_SAMPLE_FWD = 128;
__device__ __forceinline__ void fz_seed_finish(const float (&acc)[4], int lane, int64_t row, int64_t n_items, int need,
                                               unsigned long long* __restrict__ gbins) {
    // 16 groups of 8 sample scores: lanes l and l ^ 16 hold 4 each (items l + 32 e and (l ^ 16) + 32 e)
    float gm = fmaxf(fmaxf(acc[0], acc[1]), fmaxf(acc[2], acc[3]));
    gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, 16));
    float gmax = gm;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    int r = (int)ceilf(2.5f * (float)need * (float)FZ_SAMPLE_FWD / (float)n_items);
    r = r < 8 ? 8 : (r > 16 ? 16 : r);
    // rank of my group's maximum among the 16 (ties broken by group number): the group of rank r - 1 holds the seed
    int rank = 0;
    const int grp = lane & 15;
#pragma unroll
    for (int g2 = 0; g2 < 16; ++g2) {
        const float o = __shfl_sync(0xffffffffu, gm, g2);
        rank += (o > gm) || (o == gm && g2 < grp);
    }
    const unsigned who = __ballot_sync(0xffffffffu, rank == r - 1 && lane < 16);
    const float cur = __shfl_sync(0xffffffffu, gm, who ? __ffs(who) - 1 : 0);
    if (lane == 0) {
        const float width = (gmax - cur) * (1.0f / 16.f);
        unsigned long long w = 0ull;                                 // 0 = degenerate (constant / NaN / inf scores): exact kernel
        if (who && width > 0.f && cur > -INFINITY && gmax < INFINITY)
            w = ((unsigned long long)__float_as_uint(cur) << 32) | __float_as_uint(width);
        gbins[row] = w;
    }
}

constexpr int FZ_SAMPLE = 128;
constexpr int FZ_SEED_ROWS = 16;        // rows per seeding CTA (2 per warp): the sample is staged once per CTA
constexpr int FZ_SEED_DC = 32;          // embedding columns staged per pass (128 x 33 floats = 16.5 KB of shared memory)
__device__ __forceinline__ void fz_seed_rows(int64_t n_items, const float* __restrict__ Ie, int64_t ldi, int64_t nb,
                                             const int64_t* __restrict__ users, const float* __restrict__ Ue, int64_t ldu, int d,
                                             const int32_t* __restrict__ mask_ptr, int k, unsigned long long* __restrict__ gbins) {
    __shared__ float tile[FZ_SAMPLE][FZ_SEED_DC + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int RPW = FZ_SEED_ROWS / 8;                            // rows per warp (blockDim = 256)
    const int64_t row0 = (int64_t)blockIdx.x * FZ_SEED_ROWS + warp * RPW;
    const bool vec_ok = (ldi & 3) == 0 && (((uintptr_t)Ie) & 15) == 0;
    float acc[RPW][4];                                               // lane l scores sample items l, l + 32, l + 64, l + 96
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[r][e] = 0.f;
    // the user rows of this warp, all passes, requested up front (d <= 128: at most 4 passes of 32 columns)
    float ulv[RPW][4];
    int needv[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int64_t row = row0 + r;
        needv[r] = k + ((mask_ptr && row < nb) ? mask_ptr[row + 1] - mask_ptr[row] : 0);
        const float* u = row < nb ? Ue + (users ? users[row] : row) * ldu : nullptr;
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) ulv[r][ps] = (u && ps * FZ_SEED_DC + lane < d) ? __ldg(u + ps * FZ_SEED_DC + lane) : 0.f;
    }
    // ... and so is the whole sample (all passes): one memory round trip for the kernel instead of one per pass
    const bool fast = vec_ok && (d % FZ_SEED_DC) == 0;
    float4 v[4][4];
    if (fast) {
#pragma unroll
        for (int ps = 0; ps < 4; ++ps)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = (int)threadIdx.x + q * 256, si = idx >> 3, c4 = idx & 7;
                v[ps][q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ps * FZ_SEED_DC < d)
                    v[ps][q] = __ldg(reinterpret_cast<const float4*>(Ie + ((int64_t)si * n_items / FZ_SAMPLE) * ldi + ps * FZ_SEED_DC) + c4);
            }
    }
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
        const int c0 = ps * FZ_SEED_DC;
        if (c0 >= d) break;                                          // block-uniform
        __syncthreads();
        if (fast) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = (int)threadIdx.x + q * 256, si = idx >> 3, c4 = idx & 7;
                tile[si][c4 * 4 + 0] = v[ps][q].x; tile[si][c4 * 4 + 1] = v[ps][q].y;
                tile[si][c4 * 4 + 2] = v[ps][q].z; tile[si][c4 * 4 + 3] = v[ps][q].w;
            }
        } else {
            for (int t = threadIdx.x; t < FZ_SAMPLE * FZ_SEED_DC; t += blockDim.x) {
                const int si = t / FZ_SEED_DC, c = t % FZ_SEED_DC;   // coalesced along the embedding
                tile[si][c] = (c0 + c < d) ? __ldg(Ie + ((int64_t)si * n_items / FZ_SAMPLE) * ldi + c0 + c) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < FZ_SEED_DC; ++c) {                       // FZ_SEED_DC == 32: lane c holds column c0 + c of each row
            float tv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) tv[e] = tile[lane + 32 * e][c];
#pragma unroll
            for (int r = 0; r < RPW; ++r) {                          // (rows beyond the block: zeros, nothing is written for them)
                const float uc = __shfl_sync(0xffffffffu, ulv[r][ps], c);
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[r][e] = fmaf(uc, tv[e], acc[r][e]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int64_t row = row0 + r;
        if (row >= nb) continue;                                     // warp-uniform
        fz_seed_finish(acc[r], lane, row, n_items, needv[r], gbins);
    }
}

// One launch for everything the scoring kernel needs prepared: item operand (when `n_it` > 0), user operand of this row
// block, and the zeroing of the per-block scratch words (flags, shared thresholds, slots, slot counter).
__global__ void fz_prep_kernel(int64_t n_items, const float* __restrict__ Ie, int64_t ldi, int64_t n_it, float* __restrict__ Ihi,
                               float* __restrict__ Ilo, int64_t nb, const int64_t* __restrict__ users, const float* __restrict__ Ue,
                               int64_t ldu, int64_t n_ut, float* __restrict__ Uhi, float* __restrict__ Ulo, int d, int KP,
                               uint32_t* __restrict__ zero0, int64_t zero0_words, uint32_t* __restrict__ zero1, int64_t zero1_words,
                               const int32_t* __restrict__ mask_ptr, int k, unsigned long long* __restrict__ gbins, int64_t seed_blocks) {
    if ((int64_t)blockIdx.x < seed_blocks) {                         // warp-uniform role split: the first blocks seed the rows
        fz_seed_rows(n_items, Ie, ldi, nb, users, Ue, ldu, d, mask_ptr, k, gbins);
        return;
    }
    int64_t t = (blockIdx.x - seed_blocks) * (int64_t)blockDim.x + threadIdx.x;
    const int64_t ti = n_it * TC_N * (KP / 4), tu = n_ut * TC_M * (KP / 4);
    if (t < ti) { pack_split_one<TC_N>(t, n_items, nullptr, Ie, ldi, d, KP, Ihi, Ilo); return; }
    t -= ti;
    if (t < tu) { pack_split_one<TC_M>(t, nb, users, Ue, ldu, d, KP, Uhi, Ulo); return; }
    t -= tu;
    if (t < zero0_words) { zero0[t] = 0; return; }
    t -= zero0_words;
    if (t < zero1_words) zero1[t] = 0;
}

// ------------------------------------------------------------------------------------------------------
// select: finalists of all splits -> drop masked -> top-k in contract order
// ------------------------------------------------------------------------------------------------------
__device__ void fz_bitonic_desc(uint64_t* a, int n) {
    for (int size = 2; size <= n; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t x = a[lo], y = a[hi];
                if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
            }
        }
    __syncthreads();
}

constexpr int EX_SLOTS = 128;           // flagged rows served by the cached-key kernels per row block
constexpr int FZ_WFIN = 512;            // finalists one warp can rank per row (typical: 60-150)

// One warp per row, four rows per CTA.
__global__ void __launch_bounds__(128) fused_select_kernel(int64_t B, int n_splits, int cap, int k, int64_t item_offset,
                                                           const float2* __restrict__ cand, const int32_t* __restrict__ cnt,
                                                           const float* __restrict__ thr, const int32_t* __restrict__ mask_ptr,
                                                           const int32_t* __restrict__ mask_items, int32_t* __restrict__ flags,
                                                           int32_t* __restrict__ slot, int32_t* __restrict__ counter,
                                                           int32_t* __restrict__ row_of_slot,
                                                           int64_t* __restrict__ out_idx, float* __restrict__ out_val) {
    __shared__ uint64_t fin_all[4][FZ_WFIN];
    __shared__ uint32_t hist_all[4][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * 4 + warp;
    if (row >= B) return;
    // A flagged row takes a slot in the key scratch of the exact kernels (slot[] holds slot + 1, 0 = not flagged; the
    // prep kernel zeroed it together with the counter).
    auto condemn = [&]() {
        if (lane == 0) {
            const int sl = atomicAdd(counter, 1);
            slot[row] = sl + 1;
            if (sl < EX_SLOTS) row_of_slot[sl] = (int32_t)row;
        }
    };
    if (flags[row]) { condemn(); return; }                            // the scoring kernel gave up on it (seed / overflow)
    uint64_t* fin = fin_all[warp];
    const int m0 = mask_ptr ? mask_ptr[row] : 0, m1 = mask_ptr ? mask_ptr[row + 1] : 0;
    // Certificate: every split's list holds ALL of its items with value >= its own final threshold, hence all items
    // >= T = max over splits.  If at least k unmasked items clear T, the global top-k is among them.
    // (lane s holds split s's count and threshold, lane q the q-th masked item: one round trip to memory for all of them)
    const int my_cnt = lane < n_splits ? cnt[row * n_splits + lane] : 0;
    float T = lane < n_splits ? thr[row * n_splits + lane] : -INFINITY;
    for (int s = 32 + lane; s < n_splits; s += 32) T = fmaxf(T, thr[row * n_splits + s]);
    const int mlen = m1 - m0;
    const int my_mask = lane < mlen ? mask_items[m0 + lane] : INT_MIN;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) T = fmaxf(T, __shfl_xor_sync(0xffffffffu, T, o));
    const int mreg = mlen < 32 ? mlen : 32;
    int n = 0;
    bool over = false;
    // Candidates of all splits as one flat list (exclusive prefix of the counts in lane s): slot t = lane + 32 i maps to
    // (split, j).  The loads of up to SEL_U slots per lane are issued together -- the lists were written by other SMs,
    // so every load is an L2 round trip, and one round trip per row instead of one per 32 candidates is the point.
    int pre = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += v;
    }
    const int total32 = __shfl_sync(0xffffffffu, pre, 31);           // candidates in the first 32 splits
    pre -= my_cnt;                                                   // exclusive
    constexpr int SEL_U = 8;
    const int nsp = n_splits < 32 ? n_splits : 32;
    for (int t0 = 0; t0 < total32; t0 += 32 * SEL_U) {
        float2 c[SEL_U];
#pragma unroll
        for (int i = 0; i < SEL_U; ++i) {
            const int t = t0 + i * 32 + lane;
            // split of slot t: the last lane whose exclusive prefix is <= t (empty splits are skipped); warp-uniform
            // control flow, the result of lanes beyond the list is not used
            int sp = 0;
            for (int s2 = 1; s2 < nsp; ++s2)
                if (__shfl_sync(0xffffffffu, pre, s2) <= t) sp = s2;
            const int j = t - __shfl_sync(0xffffffffu, pre, sp);
            c[i] = make_float2(-INFINITY, 0.f);
            if (t < total32) c[i] = cand[(row * n_splits + sp) * cap + j];
        }
#pragma unroll
        for (int i = 0; i < SEL_U; ++i) {
            const int t = t0 + i * 32 + lane;
            bool keep = t < total32 && c[i].x >= T;
            const int item = __float_as_int(c[i].y);
            for (int q = 0; q < mreg; ++q) keep &= (__shfl_sync(0xffffffffu, my_mask, q) != item);
            if (keep)
                for (int q = m0 + 32; q < m1; ++q) keep &= (mask_items[q] != item);
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const int pos = n + __popc(bal & ((1u << lane) - 1u));
            if (keep) {
                if (pos < FZ_WFIN) fin[pos] = ((uint64_t)float_key(c[i].x) << 32) | (uint32_t)(~(uint32_t)item);
                else over = true;
            }
            n += __popc(bal);
        }
    }
    for (int s = 32; s < n_splits; ++s) {                            // (more than 32 splits: tiny batches only)
        const int ns = cnt[row * n_splits + s];
        const float2* cs = cand + (row * n_splits + s) * cap;
        for (int j0 = 0; j0 < ns; j0 += 32) {
            const int j = j0 + lane;
            float2 c = make_float2(-INFINITY, 0.f);
            if (j < ns) c = cs[j];
            bool keep = j < ns && c.x >= T;
            const int item = __float_as_int(c.y);
            for (int q = 0; q < mreg; ++q) keep &= (__shfl_sync(0xffffffffu, my_mask, q) != item);
            if (keep)
                for (int q = m0 + 32; q < m1; ++q) keep &= (mask_items[q] != item);
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const int pos = n + __popc(bal & ((1u << lane) - 1u));
            if (keep) {
                if (pos < FZ_WFIN) fin[pos] = ((uint64_t)float_key(c.x) << 32) | (uint32_t)(~(uint32_t)item);
                else over = true;
            }
            n += __popc(bal);
        }
    }
    over = __any_sync(0xffffffffu, over);
    if (over || n < k) {                                             // cannot certify this row: exact kernel takes it
        if (lane == 0) flags[row] = over ? 4 : 8;
        condemn();
        return;
    }
    __syncwarp();
    // Few finalists (the usual case): rank them all against each other.  Many: first the k-th largest 32-bit key by a warp
    // radix select (4 x 8 bits, per-warp histogram in shared memory) ...
    int m = n;
    if (n > 96) {
    uint32_t* hist = hist_all[warp];
    uint32_t prefix = 0;
    int need = k;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const uint32_t hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int b = lane; b < 256; b += 32) hist[b] = 0;
        __syncwarp();
        for (int t = lane; t < n; t += 32) {
            const uint32_t key = (uint32_t)(fin[t] >> 32);
            if ((key & hi_mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncwarp();
        // lane l owns bins [8l, 8l+8); `above` = keys in the bins owned by higher lanes (exclusive suffix sum)
        uint32_t mine[8], tot = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { mine[j] = hist[lane * 8 + j]; tot += mine[j]; }
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_down_sync(0xffffffffu, incl, o);
            if (lane + o < 32) incl += v;
        }
        uint32_t cum = incl - tot;
        int dgt = -1;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
            if (dgt < 0) {
                if (cum + mine[j] >= (uint32_t)need) dgt = lane * 8 + j;
                else cum += mine[j];
            }
        }
        // the highest lane that found a digit wins (the bins above it hold fewer than `need` keys)
        const unsigned found = __ballot_sync(0xffffffffu, dgt >= 0);
        const int win = 31 - __clz(found);
        dgt = __shfl_sync(0xffffffffu, dgt, win);
        cum = __shfl_sync(0xffffffffu, cum, win);
        prefix |= (uint32_t)dgt << shift;
        need -= (int)cum;
        __syncwarp();
    }
    // ... then only the composites with key >= that key (k of them plus ties) are ranked against each other.
    // Composites are unique (item index in the low word): rank = number of larger composites = output position.
    m = 0;
    for (int t0 = 0; t0 < n; t0 += 32) {
        const int t = t0 + lane;
        const uint64_t c = t < n ? fin[t] : 0;
        const bool keep = t < n && (uint32_t)(c >> 32) >= prefix;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) fin[m + __popc(bal & ((1u << lane) - 1u))] = c;     // in place: writes never pass the reads
        m += __popc(bal);
        __syncwarp();
    }
    }
    for (int t = lane; t < m; t += 32) {
        const uint64_t me = fin[t];
        int rank = 0;
        for (int u = 0; u < m; ++u) rank += fin[u] > me;
        if (rank < k) {
            out_idx[row * k + rank] = (int64_t)(uint32_t)(~(uint32_t)me) + item_offset;
            out_val[row * k + rank] = key_float((uint32_t)(me >> 32));
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// exact fp32 rows (flagged only).  Three small kernels, all of which exit at once for rows that are not flagged:
//   (slots in the key scratch are handed out by fused_select_kernel as it flags rows)
//   exact_keys   : EX_SPLIT CTAs per flagged row recompute its scores with the fmaf chain of the CUDA-core GEMM
//                  (k ascending, items staged through shared memory so the reads are coalesced), apply the mask and
//                  store order-preserving keys
//   exact_select : radix select + ordered tie gather + sort, the contract of mmrec_topk_rows_f32
// The scratch holds `cap_rows` rows; when more rows are flagged the host runs the pair in rounds.
// ------------------------------------------------------------------------------------------------------
constexpr int EX_SPLIT = 16;


__global__ void __launch_bounds__(256) exact_keys_kernel(const int64_t* __restrict__ users, const float* __restrict__ Ue, int64_t ldu,
                                                         int64_t n_items, const float* __restrict__ Ie, int64_t ldi, int d,
                                                         const int32_t* __restrict__ mask_ptr, const int32_t* __restrict__ mask_items,
                                                         const int32_t* __restrict__ counter, const int32_t* __restrict__ row_of_slot,
                                                         int tile_rows, unsigned* __restrict__ keys) {
    // One warp per item: the 32 lanes read the item's row coalesced, multiply by the user's row (held in registers) and
    // tree-reduce with shuffles -- fp32 throughout; the summation tree differs from the fmaf chain of the CUDA-core
    // GEMM by rounding only (these rows are compared under the near-tie rule like every other).
    (void)tile_rows;
    const int sl = blockIdx.x;
    if (sl >= *counter) return;
    const int64_t row = row_of_slot[sl];
    const int lane = threadIdx.x & 31;
    const float* u = Ue + (users ? users[row] : row) * ldu;
    float ur[8];                                                     // d <= 256
#pragma unroll
    for (int i = 0; i < 8; ++i) ur[i] = (lane + 32 * i < d) ? u[lane + 32 * i] : 0.f;
    const int m0 = mask_ptr ? mask_ptr[row] : 0, m1 = mask_ptr ? mask_ptr[row + 1] : 0;
    unsigned* out = keys + (int64_t)sl * n_items;
    const int64_t w = (int64_t)blockIdx.y * 8 + (threadIdx.x >> 5), nw = (int64_t)gridDim.y * 8;
    for (int64_t i0 = w; i0 < n_items; i0 += 4 * nw) {                // 4 items in flight per warp
        float acc[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t i = i0 + e * nw;
            acc[e] = 0.f;
            if (i < n_items) {
                const float* v = Ie + i * ldi;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (lane + 32 * q < d) acc[e] = fmaf(ur[q], __ldg(v + lane + 32 * q), acc[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t i = i0 + e * nw;
            const float r = warp_sum(acc[e]);
            if (lane == 0 && i < n_items) out[i] = float_key(r);
        }
    }
    // mask: every CTA overwrites the keys of the masked items its own warps produced (item i belongs to warp i % nw)
    __syncthreads();
    for (int q = m0 + (int)threadIdx.x; q < m1; q += (int)blockDim.x) {
        const int64_t item = mask_items[q];
        if (item >= 0 && item < n_items && (item % nw) / 8 == (int64_t)blockIdx.y) out[item] = float_key(-1e10f);
    }
}

// Warp 0 of a CTA: the highest digit whose suffix count reaches `need` (lane l owns bins [8l, 8l+8)).
__device__ __forceinline__ void pick_digit_warp(const unsigned* hist, unsigned need, unsigned prefix, int shift, unsigned* s_prefix,
                                                unsigned* s_need) {
    const int lane = threadIdx.x & 31;
    unsigned mine[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { mine[j] = hist[lane * 8 + j]; tot += mine[j]; }
    unsigned incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += v;
    }
    unsigned cum = incl - tot;
    int dgt = -1;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
        if (dgt < 0) {
            if (cum + mine[j] >= need) dgt = lane * 8 + j;
            else cum += mine[j];
        }
    }
    const unsigned found = __ballot_sync(0xffffffffu, dgt >= 0);
    if (found == 0) {                                                // fewer than `need` keys in total: digit 0 (cannot happen for k <= n)
        if (lane == 0) { *s_prefix = prefix; *s_need = need; }
        return;
    }
    const int win = 31 - __clz(found);
    if (lane == win) { *s_prefix = prefix | ((unsigned)dgt << shift); *s_need = need - cum; }
}

__global__ void __launch_bounds__(256) exact_select_kernel(int64_t n_items, int k, int64_t item_offset, const int32_t* __restrict__ counter,
                                                           const int32_t* __restrict__ row_of_slot, const unsigned* __restrict__ keys_all,
                                                           int64_t* __restrict__ out_idx, float* __restrict__ out_val) {
    __shared__ unsigned hist[256];
    __shared__ uint64_t sel[1024];
    __shared__ unsigned s_prefix, s_need, s_count, s_base;
    __shared__ unsigned warp_tot[8];
    const int sl = blockIdx.x;
    if (sl >= *counter) return;
    const int64_t row = row_of_slot[sl];
    const unsigned* keys = keys_all + (int64_t)sl * n_items;
    const int tid = threadIdx.x;
    unsigned prefix = 0, need = (unsigned)k;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const unsigned hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        hist[tid] = 0;
        __syncthreads();
        for (int64_t i = tid; i < n_items; i += 256) {
            const unsigned key = keys[i];
            if ((key & hi_mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) pick_digit_warp(hist, need, prefix, shift, &s_prefix, &s_need);
        __syncthreads();
        prefix = s_prefix; need = s_need;
        __syncthreads();
    }
    const unsigned kth = prefix;
    __shared__ unsigned tie_idx[1024];
    __shared__ unsigned n_ties;
    if (tid == 0) { s_count = 0; s_base = 0; n_ties = 0; }
    __syncthreads();
    const unsigned n_gt = (unsigned)k - need;
    // strictly greater keys in any order; the indices of the keys equal to the k-th are collected and the `need` lowest
    // of them taken (normally there is exactly one)
    for (int64_t i = tid; i < n_items; i += 256) {
        const unsigned key = keys[i];
        if (key > kth) { unsigned pos = atomicAdd(&s_count, 1u); sel[pos] = ((uint64_t)key << 32) | (uint32_t)(~(uint32_t)i); }
        else if (key == kth) { unsigned pos = atomicAdd(&n_ties, 1u); if (pos < 1024u) tie_idx[pos] = (unsigned)i; }
    }
    __syncthreads();
    if (n_ties <= 1024u) {
        const unsigned nt = n_ties;
        for (unsigned t = tid; t < nt; t += 256) {
            const unsigned me = tie_idx[t];
            unsigned rank = 0;
            for (unsigned u = 0; u < nt; ++u) rank += tie_idx[u] < me;
            if (rank < need) sel[n_gt + rank] = ((uint64_t)kth << 32) | (uint32_t)(~me);
        }
    } else {
        // degenerate row (thousands of equal scores): ordered sweep, 256 items at a time
        for (int64_t i0 = 0; i0 < n_items; i0 += 256) {
            const int64_t i = i0 + tid;
            const bool eq = i < n_items && keys[i] == kth;
            const unsigned bal = __ballot_sync(0xffffffffu, eq);
            const int lane = tid & 31, wid = tid >> 5;
            if (lane == 0) warp_tot[wid] = __popc(bal);
            __syncthreads();
            unsigned off = s_base;
            for (int w = 0; w < wid; ++w) off += warp_tot[w];
            const unsigned rank = off + __popc(bal & ((1u << lane) - 1u));
            if (eq && rank < need) sel[n_gt + rank] = ((uint64_t)kth << 32) | (uint32_t)(~(uint32_t)i);
            __syncthreads();
            if (tid == 0) { unsigned tot = 0; for (int w = 0; w < 8; ++w) tot += warp_tot[w]; s_base += tot; }
            __syncthreads();
            if (s_base >= need) break;
        }
    }
    __syncthreads();
    int n2 = 1;
    while (n2 < k) n2 <<= 1;
    for (int t = k + tid; t < n2; t += 256) sel[t] = 0;
    fz_bitonic_desc(sel, n2);
    for (int t = tid; t < k; t += 256) {
        const uint64_t c = sel[t];
        out_idx[row * k + t] = (int64_t)(uint32_t)(~(uint32_t)c) + item_offset;
        out_val[row * k + t] = key_float((uint32_t)(c >> 32));
    }
}

constexpr int EX_TILE = 128;   // items staged per tile in the overflow kernel (fits 48 KB of shared memory up to d = 90, opt-in above)

// Overflow path: more rows were flagged than the key scratch has slots (only degenerate inputs, e.g. all scores equal).
// Each CTA owns 256 consecutive rows and works through those whose slot is >= first_overflow, recomputing the keys
// in every sweep instead of caching them.
__global__ void __launch_bounds__(256) exact_overflow_kernel(int64_t B, const int64_t* __restrict__ users, const float* __restrict__ Ue,
                                                             int64_t ldu, int64_t n_items, const float* __restrict__ Ie, int64_t ldi,
                                                             int d, int k, int64_t item_offset, const int32_t* __restrict__ mask_ptr,
                                                             const int32_t* __restrict__ mask_items, const int32_t* __restrict__ slot,
                                                             int first_overflow, int64_t* __restrict__ out_idx,
                                                             float* __restrict__ out_val) {
    extern __shared__ float ex_sm[];                                 // urow[d] | tile[EX_TILE][d + 1]
    __shared__ unsigned hist[256];
    __shared__ uint64_t sel[1024];
    __shared__ unsigned s_prefix, s_need, s_count, s_base;
    __shared__ unsigned warp_tot[8];
    __shared__ int todo[256];
    __shared__ int n_todo;
    const int tid = threadIdx.x;
    if (tid == 0) n_todo = 0;
    __syncthreads();
    {
        const int64_t r = (int64_t)blockIdx.x * 256 + tid;
        if (r < B && slot[r] > first_overflow) todo[atomicAdd(&n_todo, 1)] = (int)r;   // slot[] = slot + 1
    }
    __syncthreads();
    const int n_rows_todo = n_todo;
    for (int ti = 0; ti < n_rows_todo; ++ti) {
    __syncthreads();
    const int64_t row = todo[ti];
    float* urow = ex_sm;
    float* tile = ex_sm + d;
    const int ldt = d + 1;
    const float* u = Ue + (users ? users[row] : row) * ldu;
    for (int c = tid; c < d; c += 256) urow[c] = u[c];
    const int m0 = mask_ptr ? mask_ptr[row] : 0, m1 = mask_ptr ? mask_ptr[row + 1] : 0;
    // key of item i0 + tid; the items of a tile are staged through shared memory so that the global reads are coalesced.  Returns 0 for tid beyond the catalogue.
    auto tile_key = [&](int64_t i0, bool& valid) -> unsigned {
        __syncthreads();
        const int nt = (int)((n_items - i0) < EX_TILE ? (n_items - i0) : EX_TILE);
        for (int e = tid; e < nt * d; e += 256) tile[(e / d) * ldt + (e % d)] = __ldg(Ie + (i0 + e / d) * ldi + (e % d));
        __syncthreads();
        valid = tid < nt;
        if (!valid) return 0u;
        // the arithmetic of exact_keys_kernel, one thread playing the 32 lanes: lane L sums k = L, L+32, ... with fmaf,
        // then the xor-butterfly of warp_sum (16, 8, 4, 2, 1) -- so a row gets the same bits whichever exact kernel ran it
        float part[32];
#pragma unroll
        for (int L = 0; L < 32; ++L) part[L] = 0.f;
        for (int q = 0; q * 32 < d; ++q) {
#pragma unroll
            for (int L = 0; L < 32; ++L) {
                const int c = q * 32 + L;
                if (c < d) part[L] = fmaf(urow[c], tile[tid * ldt + c], part[L]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int L = 0; L < 32; ++L)
                if (L < o) part[L] = part[L] + part[L + o];
        }
        float acc = part[0];
        const int32_t item = (int32_t)(i0 + tid);
        for (int q = m0; q < m1; ++q)
            if (mask_items[q] == item) acc = -1e10f;
        return float_key(acc);
    };
    unsigned prefix = 0, need = (unsigned)k;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const unsigned hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        hist[tid] = 0;
        for (int64_t i0 = 0; i0 < n_items; i0 += EX_TILE) {
            bool valid;
            const unsigned key = tile_key(i0, valid);
            if (valid && (key & hi_mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) pick_digit_warp(hist, need, prefix, shift, &s_prefix, &s_need);
        __syncthreads();
        prefix = s_prefix; need = s_need;
        __syncthreads();
    }
    const unsigned kth = prefix;
    if (tid == 0) { s_count = 0; s_base = 0; }
    __syncthreads();
    const unsigned n_gt = (unsigned)k - need;
    // one ordered sweep: strictly greater keys anywhere, ties on the k-th key in index order
    for (int64_t i0 = 0; i0 < n_items; i0 += EX_TILE) {
        bool valid;
        const unsigned key = tile_key(i0, valid);
        if (valid && key > kth) { unsigned pos = atomicAdd(&s_count, 1u); sel[pos] = ((uint64_t)key << 32) | (uint32_t)(~(uint32_t)(i0 + tid)); }
        const bool eq = valid && key == kth;
        const unsigned bal = __ballot_sync(0xffffffffu, eq);
        const int lane = tid & 31, wid = tid >> 5;
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        unsigned off = s_base;
        for (int w = 0; w < wid; ++w) off += warp_tot[w];
        const unsigned rank = off + __popc(bal & ((1u << lane) - 1u));
        if (eq && rank < need) sel[n_gt + rank] = ((uint64_t)kth << 32) | (uint32_t)(~(uint32_t)(i0 + tid));
        __syncthreads();
        if (tid == 0) { unsigned tot = 0; for (int w = 0; w < 8; ++w) tot += warp_tot[w]; s_base += tot; }
    }
    __syncthreads();
    int n2 = 1;
    while (n2 < k) n2 <<= 1;
    for (int t = k + tid; t < n2; t += 256) sel[t] = 0;
    fz_bitonic_desc(sel, n2);
    for (int t = tid; t < k; t += 256) {
        const uint64_t c = sel[t];
        out_idx[row * k + t] = (int64_t)(uint32_t)(~(uint32_t)c) + item_offset;
        out_val[row * k + t] = key_float((uint32_t)(c >> 32));
    }
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static inline int fz_kp(int d) { return d <= 32 ? 32 : (d <= 64 ? 64 : 128); }

struct FzPlan {
    int KP, splits, tiles_per_split, cap, cap_rows;
    int64_t n_ut, n_it, rows_blk, rows_pad;
    size_t off_uhi, off_ulo, off_ihi, off_ilo, off_cand, off_cnt, off_thr, off_flags, off_mptr, off_mcur, off_mitems, off_cub, off_slot,
        off_keys, off_gbins,
        cub_bytes, total;
};

static FzPlan fz_plan(int64_t B, int64_t n_items, int d, int k, int64_t mask_nnz) {
    FzPlan P;
    P.KP = fz_kp(d);
    P.rows_blk = B < 4096 ? B : 4096;
    P.n_ut = (P.rows_blk + TC_M - 1) / TC_M;
    P.n_it = (n_items + TC_N - 1) / TC_N;
    const int sms = sm_count();
    int splits = (int)(sms / P.n_ut);            // user tiles x item splits <= SM count: one wave, no tail
    if (splits < 1) splits = 1;
    if (splits > P.n_it) splits = (int)P.n_it;
    // every split must see enough items to fill a top-(k+m) list comfortably
    while (splits > 1 && (P.n_it / splits) * TC_N < 8 * (int64_t)k + 512) --splits;
    if (splits < 1) splits = 1;
    P.tiles_per_split = (int)((P.n_it + splits - 1) / splits);
    P.splits = (int)((P.n_it + P.tiles_per_split - 1) / P.tiles_per_split);
    P.cap = (int64_t)P.tiles_per_split * TC_N > 20000 ? 1024 : 512;
    if (P.cap < 4 * k) P.cap = 4 * k;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
    P.off_uhi = take((size_t)P.n_ut * TC_M * P.KP * 4); P.off_ulo = take((size_t)P.n_ut * TC_M * P.KP * 4);
    P.off_ihi = take((size_t)P.n_it * TC_N * P.KP * 4); P.off_ilo = take((size_t)P.n_it * TC_N * P.KP * 4);
    P.off_cand = take((size_t)P.rows_blk * P.splits * P.cap * 8);
    P.off_cnt = take((size_t)P.rows_blk * P.splits * 4); P.off_thr = take((size_t)P.rows_blk * P.splits * 4);
    P.rows_pad = (P.rows_blk + 3) & ~(int64_t)3;                     // keeps ghist 16-byte aligned behind flags | gthr
    P.off_flags = take((size_t)(2 * P.rows_pad + (FZ_NB / 2) * P.rows_blk) * 4);   // flags | gthr | ghist [rows_blk][FZ_NB / 2]  (zeroed per block)
    P.off_gbins = take((size_t)P.rows_blk * 8);                      // (lo, width) per row, written by the prep kernel
    P.off_mptr = take((size_t)(B + 2) * 4);
    P.off_mcur = take((size_t)(B + 2 > 1100 ? B + 2 : 1100) * 4);   // fill cursors, or the per-CTA order flags of the sorted-mask pass
    P.off_mitems = take((size_t)(mask_nnz > 0 ? mask_nnz : 1) * 4);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int32_t*)nullptr, (int32_t*)nullptr, (int64_t)(B + 1));
    P.cub_bytes = scan_bytes;
    P.off_cub = take(scan_bytes);
    P.off_slot = take((size_t)(P.rows_blk + 1 + EX_SLOTS) * 4);
    P.cap_rows = EX_SLOTS;
    P.off_keys = take((size_t)EX_SLOTS * n_items * 4);   // cached keys of up to EX_SLOTS flagged rows
    P.total = off + 1024;
    return P;
}

bool score_fused_supported(int64_t B, int64_t n_items, int d, int k) {
    return B > 0 && d <= 128 && k >= 1 && k <= 256 && n_items >= 8 * (int64_t)k + 512 && n_items < (1ll << 31);
}

size_t score_fused_workspace_bytes(int64_t B, int64_t n_items, int d, int k, int64_t mask_nnz) {
    if (!score_fused_supported(B, n_items, d, k)) return 0;
    return fz_plan(B, n_items, d, k, mask_nnz).total;
}

// returns 1 = done, 0 = unsupported shape / workspace (caller uses the unfused path), <0 error
int score_fused(int64_t B, const int64_t* users, const float* Ue, int64_t ldu, int64_t n_items, const float* Ie, int64_t ldi,
                int d, int64_t mask_nnz, const int64_t* mask_rows, const int64_t* mask_cols, int k, int64_t item_offset,
                int64_t* out_idx, float* out_val, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (!score_fused_supported(B, n_items, d, k) || !ws) return 0;
    const FzPlan P = fz_plan(B, n_items, d, k, mask_nnz);
    if (ws_bytes < P.total) return 0;
    char* base = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    float *Uhi = (float*)(base + P.off_uhi), *Ulo = (float*)(base + P.off_ulo);
    float *Ihi = (float*)(base + P.off_ihi), *Ilo = (float*)(base + P.off_ilo);
    int32_t *mptr = (int32_t*)(base + P.off_mptr), *mcur = (int32_t*)(base + P.off_mcur), *mitems = (int32_t*)(base + P.off_mitems);
    const int T = 256;
    static bool attr_set = false;
    if (!attr_set) {
        MMREC_CUDA(cudaFuncSetAttribute(exact_overflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        MMREC_CUDA(cudaFuncSetAttribute(score_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MMREC_CUDA(cudaFuncSetAttribute(mask_csr_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set = true;
    }
    // mask -> CSR over batch rows
    const bool has_mask = mask_nnz > 0;
    if (has_mask && B <= MC_MAX_ROWS && mask_nnz <= (1ll << 18)) {
        const int nblk = (int)((mask_nnz + 1 + T - 1) / T);          // <= 1025 words of mcur hold the per-CTA order flags
        mask_csr_sorted_kernel<<<nblk, T, 0, stream>>>(mask_nnz, mask_rows, mask_cols, (int)B, item_offset, mptr, mitems, mcur);
        MMREC_LAUNCH_CHECK();
        mask_csr_small_kernel<<<1, MC_THREADS, (size_t)(B + 1 + 32) * 4, stream>>>(mask_nnz, mask_rows, mask_cols, (int)B, item_offset, mptr,
                                                                                 mitems, mcur, nblk);
        MMREC_LAUNCH_CHECK();
    } else if (has_mask) {
        MMREC_CUDA(cudaMemsetAsync(mcur, 0, (size_t)(B + 2) * 4, stream));
        mask_count_kernel<<<(unsigned)((mask_nnz + T - 1) / T), T, 0, stream>>>(mask_nnz, mask_rows, B, mcur);
        MMREC_LAUNCH_CHECK();
        size_t tmp = P.cub_bytes;
        MMREC_CUDA(cub::DeviceScan::ExclusiveSum(base + P.off_cub, tmp, mcur, mptr, B + 1, stream));
        MMREC_CUDA(cudaMemsetAsync(mcur, 0, (size_t)(B + 2) * 4, stream));
        mask_fill_kernel<<<(unsigned)((mask_nnz + T - 1) / T), T, 0, stream>>>(mask_nnz, mask_rows, mask_cols, B, item_offset, mptr,
                                                                              mcur, mitems);
        MMREC_LAUNCH_CHECK();
    }
    const FzSmem L = fz_smem(P.KP);
    int32_t* slot = (int32_t*)(base + P.off_slot);                   // slot + 1 per row | counter | row_of_slot[EX_SLOTS]
    int32_t* counter = slot + P.rows_blk;
    int32_t* row_of_slot = counter + 1;
    unsigned* keys = (unsigned*)(base + P.off_keys);
    for (int64_t r0 = 0; r0 < B; r0 += P.rows_blk) {
        const int64_t nb = (B - r0) < P.rows_blk ? (B - r0) : P.rows_blk;
        const int64_t n_ut = (nb + TC_M - 1) / TC_M;
        // operands (items: split + re-tiled once, with the first row block) + zeroed scratch words, one launch
        const int64_t n_it_now = r0 == 0 ? P.n_it : 0;
        const int64_t zero0 = 2 * P.rows_pad + (FZ_NB / 2) * P.rows_blk, zero1 = P.rows_blk + 1;   // flags | gthr | ghist, slot | counter
        const int64_t prep_threads = n_it_now * TC_N * (P.KP / 4) + n_ut * TC_M * (P.KP / 4) + zero0 + zero1;
        const int64_t seed_blocks = (nb + FZ_SEED_ROWS - 1) / FZ_SEED_ROWS;          // T = 256 threads: 8 warps x 2 rows
        fz_prep_kernel<<<(unsigned)(seed_blocks + (prep_threads + T - 1) / T), T, 0, stream>>>(
            n_items, Ie, ldi, n_it_now, Ihi, Ilo, nb, users ? users + r0 : nullptr, users ? Ue : Ue + r0 * ldu, ldu, n_ut, Uhi, Ulo, d, P.KP,
            (uint32_t*)(base + P.off_flags), zero0, (uint32_t*)slot, zero1, has_mask ? mptr + r0 : nullptr, k,
            (unsigned long long*)(base + P.off_gbins), seed_blocks);
        MMREC_LAUNCH_CHECK();
        FusedParams p;
        p.Uhi = Uhi; p.Ulo = Ulo; p.Ihi = Ihi; p.Ilo = Ilo; p.KP = P.KP; p.n_itiles = (int)P.n_it;
        p.tiles_per_split = P.tiles_per_split; p.n_splits = P.splits; p.cap = P.cap; p.k = k; p.B = nb; p.n_items = n_items;
        p.mask_ptr = has_mask ? mptr + r0 : nullptr;
        p.cand = (float2*)(base + P.off_cand); p.cnt = (int32_t*)(base + P.off_cnt); p.thr = (float*)(base + P.off_thr);
        p.flags = (int32_t*)(base + P.off_flags);
        p.gthr = (uint32_t*)(base + P.off_flags) + P.rows_pad;
        p.gbins = (unsigned long long*)(base + P.off_gbins);
        p.ghist = (uint32_t*)(base + P.off_flags) + 2 * P.rows_pad;
        score_fused_kernel<<<(unsigned)(n_ut * P.splits), 64 + 128 * L.nsets, L.total, stream>>>(p);
        MMREC_LAUNCH_CHECK();
        fused_select_kernel<<<(unsigned)((nb + 3) / 4), 128, 0, stream>>>(nb, P.splits, P.cap, k, item_offset, p.cand, p.cnt, p.thr, p.mask_ptr,
                                                              mitems, p.flags, slot, counter, row_of_slot, out_idx + r0 * k,
                                                              out_val + r0 * k);
        MMREC_LAUNCH_CHECK();
        // rows the filter could not certify: exact fp32 recompute (normally none; the kernels exit at once then)
        {
            const int tile_rows = d <= 64 ? 256 : 128;
            const int64_t* ub = users ? users + r0 : nullptr;
            const float* ue = users ? Ue : Ue + r0 * ldu;
            exact_keys_kernel<<<dim3(EX_SLOTS, EX_SPLIT), 256, 0, stream>>>(ub, ue, ldu, n_items, Ie, ldi, d, p.mask_ptr, mitems,
                                                                                 counter, row_of_slot, tile_rows, keys);
            MMREC_LAUNCH_CHECK();
            exact_select_kernel<<<EX_SLOTS, 256, 0, stream>>>(n_items, k, item_offset, counter, row_of_slot, keys, out_idx + r0 * k,
                                                             out_val + r0 * k);
            MMREC_LAUNCH_CHECK();
            exact_overflow_kernel<<<(unsigned)((nb + 255) / 256), 256, (size_t)(d + EX_TILE * (d + 1)) * sizeof(float), stream>>>(
                nb, ub, ue, ldu, n_items, Ie, ldi, d, k, item_offset, p.mask_ptr, mitems, slot, EX_SLOTS, out_idx + r0 * k, out_val + r0 * k);
            MMREC_LAUNCH_CHECK();
        }
    }
    return 1;
}

}  // namespace mmrec

// Diagnostic (synchronises the device): rows of the LAST row block of the last mmrec_score_topk_f32 call on this
// workspace that were handed to the exact kernel.  -1 if the fused path does not apply to this shape.
extern "C" int64_t mmrec_debug_fused_fallback_rows(const void* ws, int64_t B, int64_t n_items, int d, int k, int64_t mask_nnz) {
    using namespace mmrec;
    if (!score_fused_supported(B, n_items, d, k) || !ws) return -1;
    const FzPlan P = fz_plan(B, n_items, d, k, mask_nnz);
    const char* base = (const char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    const int64_t nb = B % P.rows_blk ? B % P.rows_blk : P.rows_blk;
    int32_t* h = (int32_t*)malloc((size_t)nb * 4);
    if (!h || cudaMemcpy(h, base + P.off_flags, (size_t)nb * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { free(h); return -1; }
    int64_t n = 0, why[4] = {0, 0, 0, 0};
    for (int64_t i = 0; i < nb; ++i) {
        n += h[i] != 0;
        for (int b = 0; b < 4; ++b) why[b] += (h[i] >> b) & 1;
    }
    free(h);
    if (getenv("MMREC_DEBUG"))
        fprintf(stderr, "mmrec: fused fallback rows %lld of %lld (degenerate seed %lld, list overflow %lld, > %d finalists %lld, < k certified %lld)\n",
                (long long)n, (long long)nb, (long long)why[0], (long long)why[1], FZ_WFIN, (long long)why[2], (long long)why[3]);
    return n;
}
