// f2: the evaluator's hit matrix and per-k metric sums on the device.
//
// Replaces `TopKEvaluator.evaluate` -> `calculate_metrics` (src/utils/topk_evaluator.py:70-102: a Python loop over users and
// top-k positions building the hit matrix on the host from `.cpu().numpy()` of the index matrix) and the metric functions
// of src/utils/metrics.py:12-105.  The [n_users, K] index matrix never leaves the device; what goes back to the host is
// 4 x K float64 sums.
//
// One warp per user: lane j tests position j (j + 32, ...) of the user's top-K list against the user's ground-truth
// items (sorted; binary search), a warp scan turns hits into the cumulative quantities every metric is defined on
//     recall[j]    = hits[..j] / pos_len                      (metrics.py:12-24)
//     ndcg[j]      = dcg[..j] / idcg[min(j, min(pos_len, K) - 1)]   (metrics.py:40-71)
//     precision[j] = hits[..j] / (j + 1)                      (metrics.py:96-105)
//     map[j]       = sum_{i<=j} precision[i] hit[i] / min(j + 1, pos_len)   (metrics.py:74-93)
// all in float64 like numpy; the CTA adds its users' values in shared memory and issues one atomic per (metric, j).
// The host divides by the number of users and rounds to 4 decimals as the reference does.
#include "common.cuh"

namespace mmrec {

constexpr int EV_MAXK = 128;
constexpr int EV_WARPS = 8;

__global__ void __launch_bounds__(32 * EV_WARPS) topk_metrics_kernel(int64_t n_users, int K, const int64_t* __restrict__ topk_idx,
                                                                     const int64_t* __restrict__ pos_ptr, const int64_t* __restrict__ pos_items,
                                                                     const double* __restrict__ disc, const double* __restrict__ idcg_all,
                                                                     double* __restrict__ sums /* [4][K]: recall, ndcg, precision, map */) {
    __shared__ double acc[4][EV_MAXK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < 4 * EV_MAXK; t += blockDim.x) (&acc[0][0])[t] = 0.0;
    __syncthreads();
    const int64_t u = (int64_t)blockIdx.x * EV_WARPS + warp;
    if (u < n_users) {
        const int64_t p0 = pos_ptr[u], p1 = pos_ptr[u + 1];
        const double pos_len = (double)(p1 - p0);
        const int cap = (int)((p1 - p0) < K ? (p1 - p0) : K);       // min(pos_len, K)
        double hits_before = 0.0, dcg_before = 0.0, sp_before = 0.0;
        for (int j0 = 0; j0 < K; j0 += 32) {                        // warp-uniform trip count
            const int j = j0 + lane;
            int hit = 0;
            if (j < K) {
                const int64_t item = topk_idx[u * K + j];
                int64_t lo = p0, hi = p1;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (pos_items[mid] < item) lo = mid + 1; else hi = mid;
                }
                hit = lo < p1 && pos_items[lo] == item;
            }
            // inclusive scans over the 32 positions of this round
            double ch = (double)hit, cd = (hit && j < K) ? disc[j] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double a = __shfl_up_sync(0xffffffffu, ch, o), b = __shfl_up_sync(0xffffffffu, cd, o);
                if (lane >= o) { ch += a; cd += b; }
            }
            ch += hits_before; cd += dcg_before;
            double pre_hit = (hit && j < K) ? ch / (double)(j + 1) : 0.0;       // precision at j, counted where there is a hit
            double cs = pre_hit;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double a = __shfl_up_sync(0xffffffffu, cs, o);
                if (lane >= o) cs += a;
            }
            cs += sp_before;
            if (j < K && cap > 0) {
                atomicAdd(&acc[0][j], ch / pos_len);
                atomicAdd(&acc[1][j], cd / idcg_all[j < cap - 1 ? j : cap - 1]);
                atomicAdd(&acc[2][j], ch / (double)(j + 1));
                atomicAdd(&acc[3][j], cs / (double)((j + 1) < cap ? (j + 1) : cap));
            }
            hits_before = __shfl_sync(0xffffffffu, ch, 31);
            dcg_before = __shfl_sync(0xffffffffu, cd, 31);
            sp_before = __shfl_sync(0xffffffffu, cs, 31);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 4 * K; t += blockDim.x) {
        const int m = t / K, j = t - m * K;
        const double v = acc[m][j];
        if (v != 0.0) atomicAdd(sums + (int64_t)m * K + j, v);
    }
}

}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_topk_metrics_f64(int64_t n_users, int K, const int64_t* topk_idx, const int64_t* pos_ptr, const int64_t* pos_items,
                                      const double* disc, const double* idcg_all, double* sums, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMREC_CHECK_ARG(n_users >= 0 && K >= 1 && K <= EV_MAXK, "topk_metrics: need 1 <= K <= %d", EV_MAXK);
    if (n_users == 0) return MMREC_OK;
    MMREC_CHECK_ARG(topk_idx && pos_ptr && pos_items && disc && idcg_all && sums, "topk_metrics: null pointer");
    topk_metrics_kernel<<<(unsigned)((n_users + EV_WARPS - 1) / EV_WARPS), 32 * EV_WARPS, 0, stream>>>(n_users, K, topk_idx, pos_ptr, pos_items,
                                                                                                  disc, idcg_all, sums);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}
