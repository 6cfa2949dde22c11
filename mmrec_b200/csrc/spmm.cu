// K1: CSR SpMM  y = A X  with the layer-combination epilogue fused (running sum / mean, "+h", LayerGCN's
// cosine gate).  L2/HBM-bound gather kernel: no tensor cores (0.25-0.5 flop/byte).
//
// Mapping (D = embedding width, a multiple of 32 floats):
//   * a task is a whole row or, for rows longer than the plan's segment length, one segment of it
//     (mmrec_spmm_plan; tasks arrive sorted longest-first) -- the power-law item rows would otherwise
//     serialise on one warp;
//   * T lanes (default D/4: one float4 per lane) own one task, a warp runs 32/T tasks side by side; each lane
//     holds V = D/(4T) float4 of the row, so one row of X is V coalesced T*16-byte requests; UNR rows of X are
//     in flight per lane group before the first FMA;
//   * every lane loads the (column, value) pairs it needs itself -- the T lanes of a group read the same
//     address, one broadcast transaction, no shuffles in the gather loop; the next batch's indices, the next
//     task's descriptor and the running-sum row of the epilogue are requested while the gather is in flight;
//   * the longest tasks (plan: n_cta_tasks) are run by a whole CTA each, reduced through shared memory in
//     lane-group order; the rest one lane group each, warps walking the sorted list boustrophedon;
//   * split rows: each segment writes its partial sum to scratch, the LAST segment to arrive (per-row
//     counter) adds the partials in segment order -> the summation order never depends on scheduling, so
//     results are bit-reproducible.
// Grid: persistent, (SM count x resident CTAs) CTAs of 8 warps striding over the task list.
#include <cooperative_groups.h>

#include "common.cuh"

namespace mmrec {

struct SpmmParams {
    int64_t n_rows, n_cols;
    const int32_t* rowptr; const int32_t* colidx; const float* vals;
    const int4* tasks; int64_t n_tasks, n_heavy; const int4* split_rows; int32_t* counters; float* partial;
    const float* X; int64_t ldx;
    float* Y; int64_t ldy;
    const float* acc_in; float* acc_out; int64_t ldacc; float acc_div;
    const float* gate_ref; int64_t ldgate;
    int y_acc;                         // Y[r,:] += y[r,:] instead of = (column-panelled products: the panels of one matrix add up in Y)
    const float* post; int64_t post_row0; uint32_t spost;   // acc_out[r,:] += post[r - post_row0, :] for r >= post_row0 (after the division)
    int d;
    uint32_t sx, sy, sacc, sgate;      // the leading dimensions as byte strides (vector kernel: one IMAD.WIDE per row address)
};

// T lanes cooperate on one task (row or row segment); a warp runs 32/T tasks at once.  Fewer lanes per row
// means more rows in flight per SM -- the kernel is bound by dependent-load latency (task descriptor -> column
// indices -> rows of X), not by bytes, so rows in flight is what buys throughput at Amazon-scale graphs.

// Row `r`, float4 number `f4` of a matrix with byte stride `stride`: 32 x 32 -> 64-bit multiply-add, one instruction.
__device__ __forceinline__ const float4* row_f4(const float* base, uint32_t stride, int r, int f4) {
    return reinterpret_cast<const float4*>(reinterpret_cast<const char*>(base) + (uint64_t)(uint32_t)r * stride) + f4;
}
__device__ __forceinline__ float4* row_f4(float* base, uint32_t stride, int r, int f4) {
    return reinterpret_cast<float4*>(reinterpret_cast<char*>(base) + (uint64_t)(uint32_t)r * stride) + f4;
}

template <int D, int T>
struct VecCfg {
    static_assert(D % (4 * T) == 0, "row must split into float4 per lane");
    static constexpr int V = D / (4 * T);              // float4 per lane per row of X
    static constexpr int GPW = 32 / T;                 // tasks per warp
    static constexpr int UNR = (V >= 4) ? 2 : (V == 2 ? 4 : 8);   // rows of X in flight per lane group
};

template <int D, int T>
__device__ __forceinline__ void spmm_epilogue(const SpmmParams& p, bool on, int row, int l, float4 (&y)[VecCfg<D, T>::V],
                                              const float4 (&accin)[VecCfg<D, T>::V]) {
    using C = VecCfg<D, T>;
    // Called by ALL 32 lanes (warp-uniform control flow); `on` predicates the memory traffic of this lane group.
    // Lane l of the group owns floats [(v*T + l)*4, +4) of the row: each v is one coalesced T*16-byte request.
    if (p.gate_ref) {
        float dot = 0.f, ny = 0.f, nr = 0.f;
        if (on) {
#pragma unroll
            for (int v = 0; v < C::V; ++v) {
                float4 r = __ldg(row_f4(p.gate_ref, p.sgate, row, v * T + l));
                dot += y[v].x * r.x + y[v].y * r.y + y[v].z * r.z + y[v].w * r.w;
                ny += y[v].x * y[v].x + y[v].y * y[v].y + y[v].z * y[v].z + y[v].w * y[v].w;
                nr += r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w;
            }
        }
#pragma unroll
        for (int o = T / 2; o > 0; o >>= 1) {       // xor distances < T stay inside the lane group
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            ny += __shfl_xor_sync(0xffffffffu, ny, o);
            nr += __shfl_xor_sync(0xffffffffu, nr, o);
        }
        // F.cosine_similarity(eps=1e-8): <x/max(|x|,eps), y/max(|y|,eps)>  (layergcn.py:132)
        float c = dot / (fmaxf(sqrtf(ny), 1e-8f) * fmaxf(sqrtf(nr), 1e-8f));
#pragma unroll
        for (int v = 0; v < C::V; ++v) { y[v].x *= c; y[v].y *= c; y[v].z *= c; y[v].w *= c; }
    }
    if (!on) return;
    if (p.Y) {
#pragma unroll
        for (int v = 0; v < C::V; ++v) {
            float4 o = y[v];
            if (p.y_acc) {
                const float4 q = *row_f4(p.Y, p.sy, row, v * T + l);
                o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
            }
            *row_f4(p.Y, p.sy, row, v * T + l) = o;
        }
    }
    if (p.acc_out) {
#pragma unroll
        for (int v = 0; v < C::V; ++v) {
            float4 a = y[v];
            if (p.acc_in) { a.x += accin[v].x; a.y += accin[v].y; a.z += accin[v].z; a.w += accin[v].w; }
            if (p.acc_div != 1.0f) {
                a.x = __fdiv_rn(a.x, p.acc_div); a.y = __fdiv_rn(a.y, p.acc_div);
                a.z = __fdiv_rn(a.z, p.acc_div); a.w = __fdiv_rn(a.w, p.acc_div);
            }
            if (p.post && row >= p.post_row0) {
                const float4 q = *row_f4(p.post, p.spost, row - (int)p.post_row0, v * T + l);
                a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
            }
            *row_f4(p.acc_out, p.sacc, row, v * T + l) = a;
        }
    }
}

// Gather of one task range [b, b + len) by one lane group: UNR rows of X in flight, the next batch's indices
// travel while they are.  `maxlen` is the trip bound shared by every group that runs in lock step with this one.
template <int D, int T>
__device__ __forceinline__ void spmm_gather(const SpmmParams& p, int b, int len, int maxlen, int l, float4 (&acc)[VecCfg<D, T>::V]) {
    using C = VecCfg<D, T>;
    int cj[C::UNR]; float wj[C::UNR];
    const int32_t* cp = p.colidx + b;                                // walked with immediate offsets: no per-load address math
    const float* vp = p.vals + b;
    const float* xl = p.X + l * 4;
#pragma unroll
    for (int u = 0; u < C::UNR; ++u) {
        const bool ok = u < len;
        cj[u] = ok ? __ldg(cp + u) : 0;
        wj[u] = ok ? __ldg(vp + u) : 0.f;
    }
    for (int j0 = 0; j0 < maxlen; j0 += C::UNR, cp += C::UNR, vp += C::UNR) {
        float4 x[C::UNR][C::V];
#pragma unroll
        for (int u = 0; u < C::UNR; ++u) {
            const bool ok = j0 + u < len;
#pragma unroll
            for (int v = 0; v < C::V; ++v)
                x[u][v] = ok ? __ldg(row_f4(xl, p.sx, cj[u], v * T)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float wc[C::UNR];
#pragma unroll
        for (int u = 0; u < C::UNR; ++u) wc[u] = wj[u];
#pragma unroll
        for (int u = 0; u < C::UNR; ++u) {
            const bool ok = j0 + C::UNR + u < len;
            cj[u] = ok ? __ldg(cp + C::UNR + u) : 0;
            wj[u] = ok ? __ldg(vp + C::UNR + u) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < C::UNR; ++u) {
#pragma unroll
            for (int v = 0; v < C::V; ++v) {
                acc[v].x = fmaf(wc[u], x[u][v].x, acc[v].x);
                acc[v].y = fmaf(wc[u], x[u][v].y, acc[v].y);
                acc[v].z = fmaf(wc[u], x[u][v].z, acc[v].z);
                acc[v].w = fmaf(wc[u], x[u][v].w, acc[v].w);
            }
        }
    }
}

// A finished task of a split row: publish the partial, and if this is the last segment of the row to arrive, add
// the partials in segment order and return true (the caller then runs the epilogue).  Warp-collective (full mask);
// `split` marks the lane groups that hold such a task.
template <int D, int T>
__device__ __forceinline__ bool spmm_split_finish(const SpmmParams& p, bool split, int row, int b, int sid, int l,
                                                  float4 (&acc)[VecCfg<D, T>::V], float4 (&accin)[VecCfg<D, T>::V]) {
    using C = VecCfg<D, T>;
    int4 sr = make_int4(0, 1, 0, 1);
    int old = -1;
    bool last = false;
    if (split) {
        sr = __ldg(p.split_rows + sid);                     // {first_slot, n_seg, row_begin, seg_len}
        const int seg = (b - sr.z) / sr.w;
        float* slot = p.partial + ((int64_t)sr.x + seg) * D;
#pragma unroll
        for (int v = 0; v < C::V; ++v) *reinterpret_cast<float4*>(slot + (v * T + l) * 4) = acc[v];
        __threadfence();
    }
    __syncwarp();
    if (split && l == 0) old = atomicAdd(p.counters + sid, 1);
    old = __shfl_sync(0xffffffffu, old, 0, T);
    if (split && old == sr.y - 1) {
        __threadfence();
#pragma unroll
        for (int v = 0; v < C::V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        constexpr int RB = 8;                               // partials in flight; the order of the adds stays fixed
        for (int s2 = 0; s2 < sr.y; s2 += RB) {
            float4 q[RB][C::V];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const float* ps = p.partial + ((int64_t)sr.x + s2 + r) * D;
#pragma unroll
                for (int v = 0; v < C::V; ++v)
                    q[r][v] = (s2 + r < sr.y) ? __ldcg(reinterpret_cast<const float4*>(ps + (v * T + l) * 4))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int r = 0; r < RB; ++r) {
#pragma unroll
                for (int v = 0; v < C::V; ++v) {
                    acc[v].x += q[r][v].x; acc[v].y += q[r][v].y; acc[v].z += q[r][v].z; acc[v].w += q[r][v].w;
                }
            }
        }
        if (p.acc_in) {
#pragma unroll
            for (int v = 0; v < C::V; ++v)
                accin[v] = *row_f4(p.acc_in, p.sacc, row, v * T + l);
        }
        if (l == 0) p.counters[sid] = 0;                    // self-cleaning for the next launch
        last = true;
    }
    __syncwarp();
    return last;
}

// Phase 1: the n_heavy longest tasks, one CTA each -- its 256/T lane groups split the task evenly, reduce through
// shared memory in group order, warp 0 finishes the row.  A 5,000-nnz item row becomes ~10 CTA tasks of one or two
// load batches per group instead of a chain of dozens of dependent batches on one warp.
// Phase 2: the remaining (short) tasks, one lane group each, 32/T per warp in lock step: warp-uniform control flow
// (trip count = longest task of the warp; the plan sorts by length so neighbours are alike), everything per-group
// predicated, no shuffles in the gather loop (the T lanes of a group read the same (col, val) address = one
// broadcast transaction).  Warps walk the sorted list boustrophedon, so whoever got the longest tasks in one sweep
// gets the shortest in the next.
template <int D, int T>
__device__ __forceinline__ void spmm_vec_body(const SpmmParams& p, float* __restrict__ red) {
    using C = VecCfg<D, T>;
    constexpr int G = 256 / T;                              // lane groups per CTA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / T, l = lane % T;
    const int64_t n_work = p.tasks ? p.n_tasks : p.n_rows;
    const int64_t n_heavy = p.tasks ? p.n_heavy : 0;

    // ---------------- phase 1: CTA-cooperative tasks
    for (int64_t ct = blockIdx.x; ct < n_heavy; ct += gridDim.x) {
        const int4 tk = __ldg(p.tasks + ct);
        const int row = tk.x, sid = tk.w;
        const int len = tk.z - tk.y;
        const int chunk = (len + G - 1) / G;
        const int gi = warp * C::GPW + g;
        const int mb = tk.y + gi * chunk;
        int mylen = tk.z - mb;
        mylen = mylen < 0 ? 0 : (mylen > chunk ? chunk : mylen);
        float4 accin[C::V];
#pragma unroll
        for (int v = 0; v < C::V; ++v) accin[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (warp == 0 && g == 0 && p.acc_in && sid < 0) {
#pragma unroll
            for (int v = 0; v < C::V; ++v)
                accin[v] = *row_f4(p.acc_in, p.sacc, row, v * T + l);
        }
        float4 acc[C::V];
#pragma unroll
        for (int v = 0; v < C::V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        spmm_gather<D, T>(p, mb, mylen, chunk, l, acc);
#pragma unroll
        for (int v = 0; v < C::V; ++v) *reinterpret_cast<float4*>(red + gi * D + (v * T + l) * 4) = acc[v];
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int v = 0; v < C::V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g == 0) {
                for (int q = 0; q < G; ++q) {               // group order: fixed summation order
#pragma unroll
                    for (int v = 0; v < C::V; ++v) {
                        const float4 r4 = *reinterpret_cast<const float4*>(red + q * D + (v * T + l) * 4);
                        acc[v].x += r4.x; acc[v].y += r4.y; acc[v].z += r4.z; acc[v].w += r4.w;
                    }
                }
            }
            bool do_epi = g == 0 && sid < 0;
            if (sid >= 0) do_epi = spmm_split_finish<D, T>(p, g == 0, row, tk.y, sid, l, acc, accin);
            spmm_epilogue<D, T>(p, do_epi, row, l, acc, accin);
        }
        __syncthreads();
    }

    // ---------------- phase 2: one lane group per task
    const int64_t W = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t n_light = n_work - n_heavy;
    auto slot_of = [&](int64_t it) -> int64_t { return it * W + ((it & 1) ? (W - 1 - w) : w); };   // boustrophedon
    auto fetch = [&](int64_t slot) -> int4 {
        const int64_t t = slot * C::GPW + g;
        if (t >= n_light) return make_int4(-1, 0, 0, -1);
        if (p.tasks) return __ldg(p.tasks + n_heavy + t);
        return make_int4((int)t, __ldg(p.rowptr + t), __ldg(p.rowptr + t + 1), -1);
    };
    const int64_t n_slots = (n_light + C::GPW - 1) / C::GPW;
    const int64_t n_iter = (n_slots + W - 1) / W;                   // the same for every warp of the grid
    int4 nxt = fetch(slot_of(0));
    for (int64_t it = 0; it < n_iter; ++it) {
        const int row = nxt.x, b = nxt.y, sid = nxt.w;
        const int len = nxt.z - nxt.y;
        const bool valid = row >= 0;
        nxt = fetch(slot_of(it + 1));                               // next task's descriptor: in flight during the gather
        const int maxlen = __reduce_max_sync(0xffffffffu, valid ? len : 0);
        float4 accin[C::V];
#pragma unroll
        for (int v = 0; v < C::V; ++v) accin[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && p.acc_in && sid < 0) {                         // epilogue operand does not depend on the gather
#pragma unroll
            for (int v = 0; v < C::V; ++v)
                accin[v] = *row_f4(p.acc_in, p.sacc, row, v * T + l);
        }
        float4 acc[C::V];
#pragma unroll
        for (int v = 0; v < C::V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        spmm_gather<D, T>(p, b, valid ? len : 0, maxlen, l, acc);
        bool do_epi = valid && sid < 0;
        const bool split = valid && sid >= 0;
        if (__any_sync(0xffffffffu, split)) {
            if (spmm_split_finish<D, T>(p, split, row, b, sid, l, acc, accin)) do_epi = true;
        }
        spmm_epilogue<D, T>(p, do_epi, row, l, acc, accin);
    }
}

template <int D, int T>
__global__ void __launch_bounds__(256) spmm_vec_kernel(const SpmmParams p) {
    __shared__ __align__(16) float red[(256 / T) * D];
    spmm_vec_body<D, T>(p, red);
}

// A chain of SpMMs in ONE persistent cooperative launch: the steps run in order on the same resident grid, a grid-wide
// barrier wherever a step reads what an earlier one wrote (`sync_mask`).  The propagation of one `forward` -- L layers
// on A_hat plus the item-item layer -- is one launch instead of L + 1: at Amazon-scale graphs a layer is a handful of
// dependent L2 round trips, and launch ramp + tail of every layer were a third of its time.
constexpr int SPMM_CHAIN_MAX = 8;
struct SpmmChain { SpmmParams step[SPMM_CHAIN_MAX]; int n; unsigned sync_mask; };

template <int D, int T>
__global__ void __launch_bounds__(256) spmm_chain_kernel(const SpmmChain c) {
    __shared__ __align__(16) float red[(256 / T) * D];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    for (int i = 0; i < c.n; ++i) {
        if ((c.sync_mask >> i) & 1u) { __threadfence(); grid.sync(); }
        spmm_vec_body<D, T>(c.step[i], red);
    }
}

// Any d: 32 columns at a time, scalar loads.  Correctness path for odd widths, not tuned.
__global__ void __launch_bounds__(256) spmm_generic_kernel(const SpmmParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = warp0; row < p.n_rows; row += nwarps) {
        const int b = p.rowptr[row], e = p.rowptr[row + 1];
        float dot = 0.f, ny = 0.f, nr = 0.f;
        if (p.gate_ref) {
            for (int k0 = 0; k0 < p.d; k0 += 32) {
                const int k = k0 + lane;
                float acc = 0.f;
                if (k < p.d)
                    for (int j = b; j < e; ++j) acc = fmaf(p.vals[j], p.X[(int64_t)p.colidx[j] * p.ldx + k], acc);
                float r = k < p.d ? p.gate_ref[row * p.ldgate + k] : 0.f;
                dot += acc * r; ny += acc * acc; nr += r * r;
            }
            dot = warp_sum(dot); ny = warp_sum(ny); nr = warp_sum(nr);
        }
        const float c = p.gate_ref ? dot / (fmaxf(sqrtf(ny), 1e-8f) * fmaxf(sqrtf(nr), 1e-8f)) : 1.f;
        for (int k0 = 0; k0 < p.d; k0 += 32) {
            const int k = k0 + lane;
            if (k >= p.d) continue;
            float acc = 0.f;
            for (int j = b; j < e; ++j) acc = fmaf(p.vals[j], p.X[(int64_t)p.colidx[j] * p.ldx + k], acc);
            if (p.gate_ref) acc *= c;
            if (p.Y) p.Y[row * p.ldy + k] = p.y_acc ? p.Y[row * p.ldy + k] + acc : acc;
            if (p.acc_out) {
                float a = acc + (p.acc_in ? p.acc_in[row * p.ldacc + k] : 0.f);
                if (p.acc_div != 1.0f) a = __fdiv_rn(a, p.acc_div);
                if (p.post && row >= p.post_row0) a += p.post[(row - p.post_row0) * (int64_t)(p.spost / 4) + k];
                p.acc_out[row * p.ldacc + k] = a;
            }
        }
    }
}

static thread_local int g_spmm_y_acc = 0;   // set by mmrec_spmm_acc_f32 around its call of mmrec_spmm_f32
int g_spmm_lanes = 0;   // 0 = default lanes per task for the width; set by mmrec_spmm_set_lanes (tuning knob)

template <int D, int T>
static int launch_vec(const SpmmParams& p, cudaStream_t stream) {
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        MMREC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, spmm_vec_kernel<D, T>, 256, 0));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    const int64_t n_work = p.tasks ? p.n_tasks : p.n_rows;
    const int per_block = 8 * (32 / T);
    int64_t grid = (n_work - p.n_heavy + per_block - 1) / per_block;
    if (grid < p.n_heavy) grid = p.n_heavy;
    const int64_t cap = (int64_t)sm_count() * blocks_per_sm;
    if (grid > cap) grid = cap;
    if (grid < 1) return MMREC_OK;
    spmm_vec_kernel<D, T><<<(unsigned)grid, 256, 0, stream>>>(p);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}

template <int D>
static int launch_vec_d(const SpmmParams& p, cudaStream_t stream) {
    constexpr int T4 = D / 4 > 32 ? 32 : D / 4;      // 1 float4 per lane
    constexpr int T8 = D / 8 > 32 ? 32 : D / 8;      // 2 float4 per lane
    constexpr int T16 = D / 16 > 32 ? 32 : D / 16;   // 4 float4 per lane
    int T = g_spmm_lanes ? g_spmm_lanes : T4;       // default: one float4 per lane (measured best at d = 64)
    if (T == T4) return launch_vec<D, T4>(p, stream);
    if (T == T16) return launch_vec<D, T16>(p, stream);
    return launch_vec<D, T8>(p, stream);
}

}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_spmm_set_lanes(int lanes_per_row) {
    MMREC_CHECK_ARG(lanes_per_row == 0 || lanes_per_row == 2 || lanes_per_row == 4 || lanes_per_row == 8 ||
                    lanes_per_row == 16 || lanes_per_row == 32, "spmm_set_lanes: 0 (default) or a power of two <= 32");
    g_spmm_lanes = lanes_per_row;
    return MMREC_OK;
}

extern "C" int mmrec_spmm_f32(int64_t n_rows, int64_t n_cols, int d, const int32_t* rowptr, const int32_t* colidx,
                              const float* vals, const int32_t* tasks, int64_t n_tasks, int64_t n_cta_tasks,
                              const int32_t* split_rows,
                              int32_t* counters, float* partial, const float* X, int64_t ldx, float* Y, int64_t ldy,
                              const float* acc_in, float* acc_out, int64_t ldacc, float acc_div, const float* gate_ref,
                              int64_t ldgate, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMREC_CHECK_ARG(n_rows >= 0 && n_cols >= 0 && d >= 1, "spmm: bad sizes");
    if (n_rows == 0) return MMREC_OK;
    MMREC_CHECK_ARG(rowptr && X && (Y || acc_out), "spmm: null pointer");
    MMREC_CHECK_ARG(ldx >= d && (!Y || ldy >= d) && (!acc_out || ldacc >= d) && (!gate_ref || ldgate >= d), "spmm: leading dimension < d");
    MMREC_CHECK_ARG(acc_div != 0.0f, "spmm: acc_div == 0");
    MMREC_CHECK_ARG(!tasks || (n_tasks >= 0 && n_cta_tasks >= 0 && n_cta_tasks <= n_tasks && split_rows && counters && partial),
                    "spmm: plan pointers missing");
    SpmmParams p;
    p.n_rows = n_rows; p.n_cols = n_cols; p.rowptr = rowptr; p.colidx = colidx; p.vals = vals;
    p.tasks = (const int4*)tasks; p.n_tasks = n_tasks; p.n_heavy = tasks ? n_cta_tasks : 0; p.split_rows = (const int4*)split_rows;
    p.counters = counters; p.partial = partial; p.X = X; p.ldx = ldx; p.Y = Y; p.ldy = ldy;
    p.acc_in = acc_in; p.acc_out = acc_out; p.ldacc = ldacc; p.acc_div = acc_div; p.gate_ref = gate_ref;
    p.ldgate = ldgate; p.d = d;
    p.post = nullptr; p.post_row0 = 0; p.spost = 0; p.y_acc = g_spmm_y_acc;
    MMREC_CHECK_ARG(!(p.y_acc && gate_ref), "spmm: y_accumulate does not combine with the cosine gate");
    MMREC_CHECK_ARG(ldx < (1ll << 30) && ldy < (1ll << 30) && ldacc < (1ll << 30) && ldgate < (1ll << 30) && n_cols < (1ll << 31) &&
                    n_rows < (1ll << 31), "spmm: leading dimension / size out of range");
    p.sx = (uint32_t)(ldx * 4); p.sy = (uint32_t)(ldy * 4); p.sacc = (uint32_t)(ldacc * 4); p.sgate = (uint32_t)(ldgate * 4);
    auto al16 = [](const void* q, int64_t ld) { return q == nullptr || ((((uintptr_t)q) & 15) == 0 && (ld & 3) == 0); };
    const bool vec_ok = al16(X, ldx) && al16(Y, ldy) && al16(acc_in, ldacc) && al16(acc_out, ldacc) &&
                        al16(gate_ref, ldgate) && al16(partial, 4);
    if (vec_ok) {
        switch (d) {
            case 32: return launch_vec_d<32>(p, stream);
            case 64: return launch_vec_d<64>(p, stream);
            case 128: return launch_vec_d<128>(p, stream);
            case 256: return launch_vec_d<256>(p, stream);
            default: break;
        }
    }
    p.tasks = nullptr;   // the generic kernel walks whole rows
    int64_t grid = (n_rows + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    spmm_generic_kernel<<<(unsigned)grid, 256, 0, stream>>>(p);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}

// mmrec_spmm_f32 whose Y output ACCUMULATES (Y += A X): the column panels of one matrix are multiplied one after the other
// so that the gathered rows of X stay L2-resident (graphs whose dense operand does not fit the L2).
extern "C" int mmrec_spmm_acc_f32(int64_t n_rows, int64_t n_cols, int d, const int32_t* rowptr, const int32_t* colidx,
                                  const float* vals, const int32_t* tasks, int64_t n_tasks, int64_t n_cta_tasks,
                                  const int32_t* split_rows, int32_t* counters, float* partial, const float* X, int64_t ldx, float* Y,
                                  int64_t ldy, const float* acc_in, float* acc_out, int64_t ldacc, float acc_div, int y_accumulate,
                                  void* stream_) {
    g_spmm_y_acc = y_accumulate ? 1 : 0;
    const int rc = mmrec_spmm_f32(n_rows, n_cols, d, rowptr, colidx, vals, tasks, n_tasks, n_cta_tasks, split_rows, counters, partial, X, ldx,
                                  Y, ldy, acc_in, acc_out, ldacc, acc_div, nullptr, 0, stream_);
    g_spmm_y_acc = 0;
    return rc;
}

namespace mmrec {
template <int D>
static int launch_chain(const SpmmChain& c, int64_t max_work, cudaStream_t stream) {
    constexpr int T = D / 4 > 32 ? 32 : D / 4;                       // one float4 per lane, as the single-step default
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        MMREC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, spmm_chain_kernel<D, T>, 256, 0));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    const int per_block = 8 * (32 / T);
    int64_t grid = (max_work + per_block - 1) / per_block;
    const int64_t cap = (int64_t)sm_count() * blocks_per_sm;         // every block resident: the grid barrier needs it
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    void* args[] = {(void*)&c};
    MMREC_CUDA(cudaLaunchCooperativeKernel((const void*)spmm_chain_kernel<D, T>, dim3((unsigned)grid), dim3(256), args, 0, stream));
    ++g_launches;
    return MMREC_OK;
}
}  // namespace mmrec

extern "C" int mmrec_spmm_chain_f32(int d, int n_steps, const mmrec_spmm_step* steps, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMREC_CHECK_ARG(n_steps >= 1 && n_steps <= SPMM_CHAIN_MAX && steps, "spmm_chain: 1 <= n_steps <= %d", SPMM_CHAIN_MAX);
    if (!(d == 32 || d == 64 || d == 128 || d == 256) || g_spmm_lanes) {
        set_error("spmm_chain: d = %d (or a lane override) has no chained kernel; launch the steps one by one", d);
        return MMREC_EUNSUPPORTED;
    }
    SpmmChain c;
    c.n = n_steps; c.sync_mask = 0;
    int64_t max_work = 1;
    auto al16 = [](const void* q, int64_t ld) { return q == nullptr || ((((uintptr_t)q) & 15) == 0 && (ld & 3) == 0); };
    for (int i = 0; i < n_steps; ++i) {
        const mmrec_spmm_step& t = steps[i];
        MMREC_CHECK_ARG(t.n_rows >= 0 && t.n_cols >= 0 && t.rowptr && t.X && (t.Y || t.acc_out), "spmm_chain: step %d: null pointer / bad sizes", i);
        MMREC_CHECK_ARG(t.tasks && t.n_tasks >= 0 && t.n_cta_tasks >= 0 && t.n_cta_tasks <= t.n_tasks && t.split_rows && t.counters && t.partial,
                        "spmm_chain: step %d: the chained kernel needs the work plan (mmrec_spmm_plan)", i);
        MMREC_CHECK_ARG(t.ldx >= d && (!t.Y || t.ldy >= d) && (!t.acc_out || t.ldacc >= d) && (!t.post || t.ldpost >= d) && t.acc_div != 0.0f,
                        "spmm_chain: step %d: leading dimension < d or acc_div == 0", i);
        MMREC_CHECK_ARG(t.ldx < (1ll << 30) && t.ldy < (1ll << 30) && t.ldacc < (1ll << 30) && t.ldpost < (1ll << 30) && t.n_cols < (1ll << 31) &&
                        t.n_rows < (1ll << 31), "spmm_chain: step %d: size out of range", i);
        if (!(al16(t.X, t.ldx) && al16(t.Y, t.ldy) && al16(t.acc_in, t.ldacc) && al16(t.acc_out, t.ldacc) && al16(t.post, t.ldpost) && al16(t.partial, 4))) {
            set_error("spmm_chain: step %d: operands not 16-byte aligned; launch the steps one by one", i);
            return MMREC_EUNSUPPORTED;
        }
        SpmmParams& p = c.step[i];
        p.n_rows = t.n_rows; p.n_cols = t.n_cols; p.rowptr = t.rowptr; p.colidx = t.colidx; p.vals = t.vals;
        p.tasks = (const int4*)t.tasks; p.n_tasks = t.n_tasks; p.n_heavy = t.n_cta_tasks; p.split_rows = (const int4*)t.split_rows;
        p.counters = t.counters; p.partial = t.partial; p.X = t.X; p.ldx = t.ldx; p.Y = t.Y; p.ldy = t.ldy;
        p.acc_in = t.acc_in; p.acc_out = t.acc_out; p.ldacc = t.ldacc; p.acc_div = t.acc_div; p.gate_ref = nullptr; p.ldgate = 0;
        p.post = t.post; p.post_row0 = t.post_row0; p.d = d; p.y_acc = 0;
        p.sx = (uint32_t)(t.ldx * 4); p.sy = (uint32_t)(t.ldy * 4); p.sacc = (uint32_t)(t.ldacc * 4); p.sgate = 0; p.spost = (uint32_t)(t.ldpost * 4);
        if (t.sync_before && i > 0) c.sync_mask |= 1u << i;
        const int64_t work = t.n_tasks > t.n_cta_tasks ? t.n_tasks : t.n_cta_tasks * 8;
        if (work > max_work) max_work = work;
    }
    switch (d) {
        case 32: return launch_chain<32>(c, max_work, stream);
        case 64: return launch_chain<64>(c, max_work, stream);
        case 128: return launch_chain<128>(c, max_work, stream);
        default: return launch_chain<256>(c, max_work, stream);
    }
}
