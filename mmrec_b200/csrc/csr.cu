// K1c: COO -> CSR (coalesce semantics), SpMM work plan, bipartite edge normalisation.
// Per-graph / per-epoch preprocessing; the sort and scans use CUB device primitives, everything
// else is hand-written.  Reference behaviour replaced: the coalesce()+COO->CSR that ATen runs inside
// every torch.sparse.mm call (see include/mmrec_b200.h).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace mmrec {

__global__ void make_keys_kernel(int64_t nnz, const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                 const float* __restrict__ val, uint64_t n_cols, uint64_t* __restrict__ keys,
                                 float* __restrict__ vals) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    keys[i] = (uint64_t)row[i] * n_cols + (uint64_t)col[i];
    vals[i] = val ? val[i] : 1.0f;
}

__global__ void head_flags_kernel(int64_t nnz, const uint64_t* __restrict__ keys, int sum_dup, int32_t* __restrict__ head) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    head[i] = (!sum_dup || i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void compact_kernel(int64_t nnz, const uint64_t* __restrict__ keys, const float* __restrict__ vals,
                               const int32_t* __restrict__ pos, int sum_dup, uint64_t n_cols,
                               int32_t* __restrict__ colidx, float* __restrict__ out_vals, int64_t* __restrict__ nnz_out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i == 0) nnz_out[0] = nnz > 0 ? (int64_t)pos[nnz - 1] : 0;
    if (i >= nnz) return;
    uint64_t k = keys[i];
    bool head = !sum_dup || i == 0 || k != keys[i - 1];
    if (!head) return;
    float s = vals[i];
    if (sum_dup)
        for (int64_t j = i + 1; j < nnz && keys[j] == k; ++j) s += vals[j];   // input order (stable sort)
    int32_t p = pos[i] - 1;
    colidx[p] = (int32_t)(k % n_cols);
    out_vals[p] = s;
}

__global__ void rowptr_kernel(int64_t n_rows, int64_t nnz, const uint64_t* __restrict__ keys,
                              const int32_t* __restrict__ pos, uint64_t n_cols, int32_t* __restrict__ rowptr) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    uint64_t target = (uint64_t)r * n_cols;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    rowptr[r] = (lo == nnz) ? (nnz > 0 ? pos[nnz - 1] : 0) : pos[lo] - 1;
}

struct CsrWs {
    uint64_t *keys_a, *keys_b;
    float *vals_a, *vals_b;
    int32_t *head, *pos;
    void* cub_tmp;
    size_t cub_bytes, total;
};

static CsrWs csr_ws_layout(int64_t nnz, void* base) {
    CsrWs w;
    size_t n = (size_t)(nnz > 0 ? nnz : 1);
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (float*)nullptr,
                                    (float*)nullptr, (int64_t)n);
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (int32_t*)nullptr, (int32_t*)nullptr, (int64_t)n);
    w.cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* q = p ? p + off : nullptr; off += align_up(bytes, 256); return q; };
    w.keys_a = (uint64_t*)take(n * 8);
    w.keys_b = (uint64_t*)take(n * 8);
    w.vals_a = (float*)take(n * 4);
    w.vals_b = (float*)take(n * 4);
    w.head = (int32_t*)take(n * 4);
    w.pos = (int32_t*)take(n * 4);
    w.cub_tmp = take(w.cub_bytes);
    w.total = off;
    return w;
}

// ------------------------------------------------------------------------------------------------
// SpMM plan
// ------------------------------------------------------------------------------------------------
__global__ void plan_count_kernel(int64_t n_rows, const int32_t* __restrict__ rowptr, int seg, int32_t* __restrict__ nt,
                                  int32_t* __restrict__ ns, int32_t* __restrict__ nl, int* __restrict__ longest) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    int t = 0, s = 0, l = 0;
    if (r < n_rows) {
        int len = rowptr[r + 1] - rowptr[r];
        t = len <= seg ? 1 : (len + seg - 1) / seg;
        s = t > 1 ? 1 : 0;
        l = t > 1 ? t : 0;
        atomicMax(longest, len);
    }
    nt[r] = t; ns[r] = s; nl[r] = l;
}

__global__ void plan_fill_kernel(int64_t n_rows, const int32_t* __restrict__ rowptr, int seg, int light_max,
                                 const int32_t* __restrict__ t_off, const int32_t* __restrict__ s_off,
                                 const int32_t* __restrict__ l_off, const int* __restrict__ longest,
                                 int4* __restrict__ tasks, uint32_t* __restrict__ keys, int4* __restrict__ split_rows,
                                 int64_t* __restrict__ counts) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r == 0) {
        counts[0] = t_off[n_rows]; counts[1] = s_off[n_rows]; counts[2] = l_off[n_rows]; counts[3] = *longest;
    }
    if (r >= n_rows) return;
    int b = rowptr[r], e = rowptr[r + 1];
    int nt = t_off[r + 1] - t_off[r];
    int t0 = t_off[r];
    // tasks longer than light_max are run by a whole CTA; sorted longest-first they form a prefix of the list
    {
        int heavy = 0;
        for (int j = 0; j < nt; ++j) {
            int sb = b + j * seg;
            int se = (nt == 1) ? e : (sb + seg < e ? sb + seg : e);
            heavy += (se - sb) > light_max;
        }
        if (heavy) atomicAdd((unsigned long long*)(counts + 4), (unsigned long long)heavy);
    }
    if (nt == 1) {
        tasks[t0] = make_int4((int)r, b, e, -1);
        keys[t0] = (uint32_t)(seg - (e - b));          // ascending key = longest task first
    } else {
        int sid = s_off[r];
        split_rows[sid] = make_int4(l_off[r], nt, b, seg);
        for (int j = 0; j < nt; ++j) {
            int sb = b + j * seg;
            int se = sb + seg < e ? sb + seg : e;
            tasks[t0 + j] = make_int4((int)r, sb, se, sid);
            keys[t0 + j] = (uint32_t)(seg - (se - sb));
        }
    }
}

__global__ void plan_pad_kernel(int64_t max_tasks, uint32_t* __restrict__ keys, int4* __restrict__ tasks) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= max_tasks) return;
    keys[t] = 4095u;                                   // unused slots sort behind every real task
    tasks[t] = make_int4(-1, 0, 0, -1);
}

// ------------------------------------------------------------------------------------------------
// bipartite degree normalisation
// ------------------------------------------------------------------------------------------------
__global__ void degree_kernel(int64_t n_edges, const int64_t* __restrict__ users, const int64_t* __restrict__ items,
                              int32_t* __restrict__ du, int32_t* __restrict__ di) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    atomicAdd(&du[users[e]], 1);
    atomicAdd(&di[items[e]], 1);
}

__global__ void norm_vals_kernel(int64_t n_edges, const int64_t* __restrict__ users, const int64_t* __restrict__ items,
                                 const int32_t* __restrict__ du, const int32_t* __restrict__ di, float eps,
                                 float* __restrict__ vals) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    // torch.pow(x, -0.5) == 1 / sqrt(x) with IEEE sqrt and divide (freedom.py:149-153)
    float ru = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn((float)du[users[e]], eps)));
    float ri = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn((float)di[items[e]], eps)));
    vals[e] = __fmul_rn(ru, ri);
}

}  // namespace mmrec

using namespace mmrec;

extern "C" size_t mmrec_csr_from_coo_workspace_bytes(int64_t nnz, int64_t /*n_rows*/) {
    return csr_ws_layout(nnz, nullptr).total;
}

extern "C" int mmrec_csr_from_coo(int64_t nnz, const int64_t* row, const int64_t* col, const float* val,
                                  int64_t n_rows, int64_t n_cols, int sum_duplicates, int32_t* rowptr,
                                  int32_t* colidx, float* vals, int64_t* nnz_out, void* ws, size_t ws_bytes,
                                  void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMREC_CHECK_ARG(nnz >= 0 && n_rows >= 0 && n_cols > 0, "csr_from_coo: bad sizes nnz=%lld rows=%lld cols=%lld",
                    (long long)nnz, (long long)n_rows, (long long)n_cols);
    MMREC_CHECK_ARG(nnz < (1ll << 31) && n_rows < (1ll << 31) && n_cols < (1ll << 31), "csr_from_coo: int32 CSR overflow");
    MMREC_CHECK_ARG(rowptr && nnz_out && (nnz == 0 || (row && col && colidx && vals)), "csr_from_coo: null pointer");
    CsrWs w = csr_ws_layout(nnz, ws);
    if (ws_bytes < w.total || !ws) {
        set_error("csr_from_coo: workspace %zu < %zu", ws_bytes, w.total);
        return MMREC_EWORKSPACE;
    }
    const int T = 256;
    if (nnz > 0) {
        int64_t nb = (nnz + T - 1) / T;
        make_keys_kernel<<<(unsigned)nb, T, 0, stream>>>(nnz, row, col, val, (uint64_t)n_cols, w.keys_a, w.vals_a);
        MMREC_LAUNCH_CHECK();
        int end_bit = 1;
        unsigned __int128 span = (unsigned __int128)n_rows * (unsigned __int128)n_cols;
        while (end_bit < 64 && ((unsigned __int128)1 << end_bit) < span) ++end_bit;
        size_t tmp = w.cub_bytes;
        MMREC_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.keys_a, w.keys_b, w.vals_a, w.vals_b, nnz, 0,
                                                   end_bit, stream));
        head_flags_kernel<<<(unsigned)nb, T, 0, stream>>>(nnz, w.keys_b, sum_duplicates, w.head);
        MMREC_LAUNCH_CHECK();
        tmp = w.cub_bytes;
        MMREC_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tmp, w.head, w.pos, nnz, stream));
        compact_kernel<<<(unsigned)nb, T, 0, stream>>>(nnz, w.keys_b, w.vals_b, w.pos, sum_duplicates, (uint64_t)n_cols,
                                                      colidx, vals, nnz_out);
        MMREC_LAUNCH_CHECK();
    } else {
        MMREC_CUDA(cudaMemsetAsync(nnz_out, 0, sizeof(int64_t), stream));
    }
    int64_t nbr = (n_rows + 1 + T - 1) / T;
    rowptr_kernel<<<(unsigned)nbr, T, 0, stream>>>(n_rows, nnz, w.keys_b, w.pos, (uint64_t)n_cols, rowptr);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}

namespace {
struct PlanWs {
    int32_t *nt, *ns, *nl, *t_off, *s_off, *l_off;
    int* longest;
    uint32_t *keys_a, *keys_b;
    int4* tasks_tmp;
    void* cub_tmp;
    size_t cub_bytes, total;
};
PlanWs plan_ws_layout(int64_t n_rows, int64_t max_tasks, void* base) {
    PlanWs w;
    size_t n = (size_t)n_rows + 1;
    size_t m = (size_t)(max_tasks > 0 ? max_tasks : 1);
    size_t scan_bytes = 0, sort_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int32_t*)nullptr, (int32_t*)nullptr, (int64_t)n);
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int4*)nullptr,
                                    (int4*)nullptr, (int64_t)m, 0, 12);
    w.cub_bytes = scan_bytes > sort_bytes ? scan_bytes : sort_bytes;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* q = p ? p + off : nullptr; off += mmrec::align_up(bytes, 256); return q; };
    w.nt = (int32_t*)take(n * 4); w.ns = (int32_t*)take(n * 4); w.nl = (int32_t*)take(n * 4);
    w.t_off = (int32_t*)take(n * 4); w.s_off = (int32_t*)take(n * 4); w.l_off = (int32_t*)take(n * 4);
    w.longest = (int*)take(256);
    w.keys_a = (uint32_t*)take(m * 4); w.keys_b = (uint32_t*)take(m * 4);
    w.tasks_tmp = (int4*)take(m * 16);
    w.cub_tmp = take(w.cub_bytes);
    w.total = off;
    return w;
}
}  // namespace

extern "C" size_t mmrec_spmm_plan_workspace_bytes(int64_t n_rows, int64_t max_tasks) {
    return plan_ws_layout(n_rows, max_tasks, nullptr).total;
}

extern "C" int mmrec_spmm_plan(int64_t n_rows, const int32_t* rowptr, int seg, int light_max, int64_t max_tasks,
                               int32_t* tasks, int32_t* split_rows, int64_t* counts, void* ws, size_t ws_bytes,
                               void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMREC_CHECK_ARG(n_rows >= 0 && seg >= 32 && seg <= 4000 && light_max >= 1 && max_tasks >= n_rows && rowptr && tasks &&
                        split_rows && counts,
                    "spmm_plan: bad argument (need 32 <= seg <= 4000, light_max >= 1, max_tasks >= n_rows)");
    MMREC_CUDA(cudaMemsetAsync(counts, 0, 8 * sizeof(int64_t), stream));
    PlanWs w = plan_ws_layout(n_rows, max_tasks, ws);
    if (ws_bytes < w.total || !ws) {
        set_error("spmm_plan: workspace %zu < %zu", ws_bytes, w.total);
        return MMREC_EWORKSPACE;
    }
    const int T = 256;
    int64_t nb = (n_rows + 1 + T - 1) / T;
    MMREC_CUDA(cudaMemsetAsync(w.longest, 0, sizeof(int), stream));
    plan_count_kernel<<<(unsigned)nb, T, 0, stream>>>(n_rows, rowptr, seg, w.nt, w.ns, w.nl, w.longest);
    MMREC_LAUNCH_CHECK();
    size_t tmp = w.cub_bytes;
    MMREC_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tmp, w.nt, w.t_off, n_rows + 1, stream));
    tmp = w.cub_bytes;
    MMREC_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tmp, w.ns, w.s_off, n_rows + 1, stream));
    tmp = w.cub_bytes;
    MMREC_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tmp, w.nl, w.l_off, n_rows + 1, stream));
    if (max_tasks > 0) {
        plan_pad_kernel<<<(unsigned)((max_tasks + T - 1) / T), T, 0, stream>>>(max_tasks, w.keys_a, w.tasks_tmp);
        MMREC_LAUNCH_CHECK();
    }
    plan_fill_kernel<<<(unsigned)nb, T, 0, stream>>>(n_rows, rowptr, seg, light_max, w.t_off, w.s_off, w.l_off, w.longest,
                                                    w.tasks_tmp, w.keys_a, (int4*)split_rows, counts);
    MMREC_LAUNCH_CHECK();
    if (max_tasks > 0) {   // longest tasks first: homogeneous work inside a warp, heavy rows never in the tail
        tmp = w.cub_bytes;
        MMREC_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.keys_a, w.keys_b, w.tasks_tmp, (int4*)tasks, max_tasks, 0, 12,
                                                   stream));
    }
    return MMREC_OK;
}

extern "C" size_t mmrec_bipartite_norm_workspace_bytes(int64_t n_users, int64_t n_items) {
    return (size_t)(n_users + n_items) * sizeof(int32_t);
}

extern "C" int mmrec_bipartite_norm_f32(int64_t n_edges, const int64_t* users, const int64_t* items, int64_t n_users,
                                        int64_t n_items, float eps, float* vals, void* ws, size_t ws_bytes,
                                        void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMREC_CHECK_ARG(n_edges >= 0 && n_users > 0 && n_items > 0, "bipartite_norm: bad sizes");
    MMREC_CHECK_ARG(n_edges == 0 || (users && items && vals), "bipartite_norm: null pointer");
    size_t need = mmrec_bipartite_norm_workspace_bytes(n_users, n_items);
    if (!ws || ws_bytes < need) {
        set_error("bipartite_norm: workspace %zu < %zu", ws_bytes, need);
        return MMREC_EWORKSPACE;
    }
    int32_t* du = (int32_t*)ws;
    int32_t* di = du + n_users;
    MMREC_CUDA(cudaMemsetAsync(ws, 0, need, stream));
    if (n_edges == 0) return MMREC_OK;
    const int T = 256;
    int64_t nb = (n_edges + T - 1) / T;
    degree_kernel<<<(unsigned)nb, T, 0, stream>>>(n_edges, users, items, du, di);
    MMREC_LAUNCH_CHECK();
    norm_vals_kernel<<<(unsigned)nb, T, 0, stream>>>(n_edges, users, items, du, di, eps, vals);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}
