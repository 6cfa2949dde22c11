// K4: the user-embedding exchange of the item-sharded propagation as ONE kernel over peer memory.
//
// Rank g's SpMM writes its partial user sums R_g E_Ig straight into a buffer that every peer has mapped (CUDA IPC /
// symmetric memory over NVLink).  After a cross-GPU barrier this kernel, on every rank, streams all `world` partials
// with 16-byte P2P loads, adds them IN RANK ORDER (every rank gets the same bits, run to run) and applies the layer
// mean epilogue (`acc_out = (acc_in + sum) / div`) in the same pass -- the all-reduce, the `acc + part` pass and
// the final division of the NCCL formulation collapse into one read of (world) x [U, d] and two writes.
// HBM/NVLink-bound: 4 n (world + 1 + writes) bytes per launch, no reuse.
#include "common.cuh"

namespace mmrec {

constexpr int PEER_MAX = 16;
struct PeerParts { const float4* p[PEER_MAX]; };

__global__ void __launch_bounds__(256) peer_sum_kernel(int64_t n4, int world, const PeerParts parts, const float4* __restrict__ acc_in,
                                                       float4* __restrict__ acc_out, float div, float4* __restrict__ sum_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 s = parts.p[0][i];
        for (int r = 1; r < world; ++r) {                            // fixed order: bit-identical on every rank
            const float4 v = parts.p[r][i];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        if (sum_out) sum_out[i] = s;
        if (acc_out) {
            float4 a = s;
            if (acc_in) { const float4 c = acc_in[i]; a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w; }
            if (div != 1.0f) { a.x = __fdiv_rn(a.x, div); a.y = __fdiv_rn(a.y, div); a.z = __fdiv_rn(a.z, div); a.w = __fdiv_rn(a.w, div); }
            acc_out[i] = a;
        }
    }
}

}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_peer_sum_f32(int64_t n, int world, const void* const* parts, const float* acc_in, float* acc_out, float acc_div,
                                  float* sum_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMREC_CHECK_ARG(n >= 0 && world >= 1 && world <= PEER_MAX && parts, "peer_sum: bad sizes (1 <= world <= 16)");
    MMREC_CHECK_ARG((n & 3) == 0, "peer_sum: n must be a multiple of 4 floats");
    MMREC_CHECK_ARG(acc_out || sum_out, "peer_sum: no output");
    MMREC_CHECK_ARG(acc_div != 0.0f, "peer_sum: acc_div == 0");
    if (n == 0) return MMREC_OK;
    PeerParts P;
    for (int r = 0; r < PEER_MAX; ++r) P.p[r] = nullptr;
    for (int r = 0; r < world; ++r) {
        MMREC_CHECK_ARG(parts[r] && (((uintptr_t)parts[r]) & 15) == 0, "peer_sum: partial pointers must be non-null and 16-byte aligned");
        P.p[r] = (const float4*)parts[r];
    }
    MMREC_CHECK_ARG(((((uintptr_t)acc_in) | ((uintptr_t)acc_out) | ((uintptr_t)sum_out)) & 15) == 0, "peer_sum: 16-byte alignment");
    const int64_t n4 = n / 4;
    int64_t grid = (n4 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    peer_sum_kernel<<<(unsigned)grid, 256, 0, stream>>>(n4, world, P, (const float4*)acc_in, (float4*)acc_out, acc_div, (float4*)sum_out);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}
