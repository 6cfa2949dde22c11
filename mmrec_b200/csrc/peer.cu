// K4: the user-embedding exchange of the item-sharded propagation as ONE kernel over peer memory.
//
// Rank g's SpMM writes its partial user sums R_g E_Ig straight into a buffer that every peer has mapped (CUDA IPC /
// symmetric memory over NVLink).  After a cross-GPU barrier this kernel, on every rank, streams all `world` partials
// with 16-byte P2P loads, adds them IN RANK ORDER (every rank gets the same bits, run to run) and applies the layer
// mean epilogue (`acc_out = (acc_in + sum) / div`) in the same pass -- the all-reduce, the `acc + part` pass and
// the final division of the NCCL formulation collapse into one read of (world) x [U, d] and two writes.
// HBM/NVLink-bound: 4 n (world + 1 + writes) bytes per launch, no reuse.
#include "peer_sync.cuh"

namespace mmrec {

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


// Reduce-scatter + all-gather form of the same exchange (what scales: every rank READS (world-1)/world of one partial
// and WRITES (world-1)/world of the result over NVLink, instead of reading world-1 whole partials).  Rank `rank` owns
// the slice [lo4, hi4) of the float4 range: it sums that slice of all partials in rank order -- each element is summed
// exactly once, by its owner, so every rank ends up with identical bits -- and pushes the result into the same slice of
// every rank's `dst` buffer (its own included).  The layer-mean accumulator lives sliced as well: `acc` is this rank's
// slice only; a non-final layer does acc += sum and pushes the sum (the next layer's E_U), the final layer pushes
// (acc + sum) / div (the propagated user table).  The caller synchronises the ranks before (partials complete) and
// after (pushes visible) the call.
template <int W, int U>
__global__ void __launch_bounds__(256) peer_reduce_push_kernel(int64_t lo4, int64_t hi4, const PeerParts parts, const PeerParts dst,
                                                               const float4* __restrict__ acc_in, float4* __restrict__ acc_out, float div,
                                                               int final_layer) {
    // W ranks, U elements per thread and trip: W * U sixteen-byte loads are in flight per thread before the first add (a peer
    // load is a ~2-3 us round trip over NVLink; with one element per thread the kernel ran at ~150 GB/s per direction)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = lo4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < hi4; i0 += stride * U) {
        float4 v[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
#pragma unroll
            for (int r = 0; r < W; ++r) v[u][r] = i < hi4 ? parts.p[r][i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i >= hi4) continue;
            float4 s = v[u][0];
#pragma unroll
            for (int r = 1; r < W; ++r) { s.x += v[u][r].x; s.y += v[u][r].y; s.z += v[u][r].z; s.w += v[u][r].w; }   // rank order
            float4 out = s;
            if (acc_in) {
                float4 a = acc_in[i - lo4];
                a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
                if (final_layer) {
                    if (div != 1.0f) { a.x = __fdiv_rn(a.x, div); a.y = __fdiv_rn(a.y, div); a.z = __fdiv_rn(a.z, div); a.w = __fdiv_rn(a.w, div); }
                    out = a;
                } else if (acc_out) {
                    acc_out[i - lo4] = a;
                }
            }
#pragma unroll
            for (int r = 0; r < W; ++r) const_cast<float4*>(dst.p[r])[i] = out;
        }
    }
}

// any world size up to PEER_MAX (the common ones have their own instantiation above)
__global__ void __launch_bounds__(256) peer_reduce_push_any_kernel(int64_t lo4, int64_t hi4, int world, const PeerParts parts, const PeerParts dst,
                                                                   const float4* __restrict__ acc_in, float4* __restrict__ acc_out, float div,
                                                                   int final_layer) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = lo4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hi4; i += stride) {
        float4 s = parts.p[0][i];
        for (int r = 1; r < world; ++r) { const float4 v = parts.p[r][i]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
        float4 out = s;
        if (acc_in) {
            float4 a = acc_in[i - lo4];
            a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
            if (final_layer) {
                if (div != 1.0f) { a.x = __fdiv_rn(a.x, div); a.y = __fdiv_rn(a.y, div); a.z = __fdiv_rn(a.z, div); a.w = __fdiv_rn(a.w, div); }
                out = a;
            } else if (acc_out) {
                acc_out[i - lo4] = a;
            }
        }
        for (int r = 0; r < world; ++r) const_cast<float4*>(dst.p[r])[i] = out;
    }
}

// mmrec_peer_reduce_push_f32 with both barriers inside: ONE launch per layer does "wait until every rank's partial is
// complete -> reduce my slice -> store it to every rank -> wait until every rank's slices have landed".
__global__ void __launch_bounds__(256) peer_exchange_kernel(int64_t lo4, int64_t hi4, int world, int rank, const PeerParts parts,
                                                            const PeerParts dst, const PeerFlags flags, int* __restrict__ state,
                                                            const float4* __restrict__ acc_in, float4* __restrict__ acc_out, float div,
                                                            int final_layer) {
    const int epoch = peer_enter(flags, state, rank, world);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = lo4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hi4; i += stride) {
        float4 v[PEER_MAX];
#pragma unroll
        for (int r = 0; r < PEER_MAX; ++r)
            if (r < world) v[r] = __ldcv(parts.p[r] + i);            // (volatile: never a stale line of an earlier call)
        float4 s = v[0];
#pragma unroll
        for (int r = 1; r < PEER_MAX; ++r)
            if (r < world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
        float4 out = s;
        if (acc_in) {
            float4 a = acc_in[i - lo4];
            a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
            if (final_layer) {
                if (div != 1.0f) { a.x = __fdiv_rn(a.x, div); a.y = __fdiv_rn(a.y, div); a.z = __fdiv_rn(a.z, div); a.w = __fdiv_rn(a.w, div); }
                out = a;
            } else if (acc_out) {
                acc_out[i - lo4] = a;
            }
        }
#pragma unroll
        for (int r = 0; r < PEER_MAX; ++r)
            if (r < world) const_cast<float4*>(dst.p[r])[i] = out;
    }
    peer_leave(flags, state, rank, world, epoch, true);
}

// barrier-only kernel on the same flags (e.g. before gathers / merges that read what the peers wrote)
__global__ void peer_barrier_kernel(int world, int rank, const PeerFlags flags, int* __restrict__ state) {
    const int epoch = peer_enter(flags, state, rank, world);
    peer_leave(flags, state, rank, world, epoch, false);
}

// dst[p * n4 + i] = src[p][i]: the shards of a table read straight from the peers that own them (all-gather by P2P loads)
__global__ void __launch_bounds__(256) peer_gather_kernel(int64_t n4, int world, const PeerParts src, float4* __restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total = n4 * world;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {   // 4 peer loads in flight per thread
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < total) { const int p = (int)(i / n4); v[u] = src.p[p][i - (int64_t)p * n4]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < total) dst[i] = v[u];
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

extern "C" int mmrec_peer_reduce_push_f32(int64_t n, int world, int rank, const void* const* parts, void* const* dst, const float* acc_in,
                                          float* acc_out, float acc_div, int final_layer, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMREC_CHECK_ARG(n >= 0 && world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world && parts && dst,
                    "peer_reduce_push: bad sizes (1 <= world <= 16, 0 <= rank < world)");
    MMREC_CHECK_ARG((n & 3) == 0, "peer_reduce_push: n must be a multiple of 4 floats");
    MMREC_CHECK_ARG(acc_div != 0.0f, "peer_reduce_push: acc_div == 0");
    if (n == 0) return MMREC_OK;
    PeerParts P, D;
    for (int r = 0; r < PEER_MAX; ++r) { P.p[r] = nullptr; D.p[r] = nullptr; }
    for (int r = 0; r < world; ++r) {
        MMREC_CHECK_ARG(parts[r] && dst[r] && ((((uintptr_t)parts[r]) | ((uintptr_t)dst[r])) & 15) == 0,
                        "peer_reduce_push: buffer pointers must be non-null and 16-byte aligned");
        P.p[r] = (const float4*)parts[r]; D.p[r] = (const float4*)dst[r];
    }
    MMREC_CHECK_ARG(((((uintptr_t)acc_in) | ((uintptr_t)acc_out)) & 15) == 0, "peer_reduce_push: 16-byte alignment");
    const int64_t n4 = n / 4;
    const int64_t per = (n4 + world - 1) / world;                     // slice of this rank (the accumulator holds `per` float4)
    const int64_t lo4 = per * rank < n4 ? per * rank : n4, hi4 = lo4 + per < n4 ? lo4 + per : n4;
    if (hi4 <= lo4) return MMREC_OK;
    const int U = world == 2 ? 4 : (world == 4 ? 2 : 1);
    int64_t grid = (hi4 - lo4 + 256 * U - 1) / (256 * U);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    const float4 *ai = (const float4*)acc_in;
    float4* ao = (float4*)acc_out;
    switch (world) {
        case 2: peer_reduce_push_kernel<2, 4><<<(unsigned)grid, 256, 0, stream>>>(lo4, hi4, P, D, ai, ao, acc_div, final_layer); break;
        case 4: peer_reduce_push_kernel<4, 2><<<(unsigned)grid, 256, 0, stream>>>(lo4, hi4, P, D, ai, ao, acc_div, final_layer); break;
        case 8: peer_reduce_push_kernel<8, 1><<<(unsigned)grid, 256, 0, stream>>>(lo4, hi4, P, D, ai, ao, acc_div, final_layer); break;
        default: peer_reduce_push_any_kernel<<<(unsigned)grid, 256, 0, stream>>>(lo4, hi4, world, P, D, ai, ao, acc_div, final_layer); break;
    }
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}

extern "C" int mmrec_peer_gather_f32(int64_t n_each, int world, const void* const* src, float* dst, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMREC_CHECK_ARG(n_each >= 0 && world >= 1 && world <= PEER_MAX && src && dst, "peer_gather: bad sizes (1 <= world <= 16)");
    MMREC_CHECK_ARG((n_each & 3) == 0 && (((uintptr_t)dst) & 15) == 0, "peer_gather: n_each must be a multiple of 4 floats, dst 16-byte aligned");
    if (n_each == 0) return MMREC_OK;
    PeerParts S;
    for (int r = 0; r < PEER_MAX; ++r) S.p[r] = nullptr;
    for (int r = 0; r < world; ++r) {
        MMREC_CHECK_ARG(src[r] && (((uintptr_t)src[r]) & 15) == 0, "peer_gather: source pointers must be non-null and 16-byte aligned");
        S.p[r] = (const float4*)src[r];
    }
    const int64_t n4 = n_each / 4;
    int64_t grid = (n4 * world + 1023) / 1024;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    peer_gather_kernel<<<(unsigned)grid, 256, 0, stream>>>(n4, world, S, (float4*)dst);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}

extern "C" int mmrec_peer_exchange_f32(int64_t n, int world, int rank, const void* const* parts, void* const* dst, void* const* flags,
                                       int32_t* state, const float* acc_in, float* acc_out, float acc_div, int final_layer, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMREC_CHECK_ARG(n >= 0 && world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world && parts && dst && flags && state,
                    "peer_exchange: bad sizes (1 <= world <= 16, 0 <= rank < world) or null pointer");
    MMREC_CHECK_ARG((n & 3) == 0, "peer_exchange: n must be a multiple of 4 floats");
    MMREC_CHECK_ARG(acc_div != 0.0f, "peer_exchange: acc_div == 0");
    PeerParts P, D;
    PeerFlags F;
    for (int r = 0; r < PEER_MAX; ++r) { P.p[r] = nullptr; D.p[r] = nullptr; F.f[r] = nullptr; }
    for (int r = 0; r < world; ++r) {
        MMREC_CHECK_ARG(parts[r] && dst[r] && flags[r] && ((((uintptr_t)parts[r]) | ((uintptr_t)dst[r])) & 15) == 0 && (((uintptr_t)flags[r]) & 3) == 0,
                        "peer_exchange: buffer pointers must be non-null and aligned (16 bytes data, 4 bytes flags)");
        P.p[r] = (const float4*)parts[r]; D.p[r] = (const float4*)dst[r]; F.f[r] = (int*)flags[r];
    }
    MMREC_CHECK_ARG(((((uintptr_t)acc_in) | ((uintptr_t)acc_out)) & 15) == 0, "peer_exchange: 16-byte alignment");
    const int64_t n4 = n / 4;
    const int64_t per = (n4 + world - 1) / world;
    const int64_t lo4 = per * rank < n4 ? per * rank : n4, hi4 = lo4 + per < n4 ? lo4 + per : n4;
    // The blocks wait for block 0's handshake with the peers, i.e. they sit on their SMs while the item-side SpMM of the
    // same layer runs on the other stream: a few dozen blocks saturate one NVLink direction (256 threads x 16 B x the
    // loads in flight each) and leave the rest of the GPU to that SpMM.
    int64_t grid = (hi4 - lo4 + 255) / 256;
    const int64_t cap = (hi4 - lo4) * 16 > (64ll << 20) ? (int64_t)sm_count() : 48;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    peer_exchange_kernel<<<(unsigned)grid, 256, 0, stream>>>(lo4, hi4, world, rank, P, D, F, state, (const float4*)acc_in, (float4*)acc_out,
                                                             acc_div, final_layer);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}

extern "C" int mmrec_peer_barrier(int world, int rank, void* const* flags, int32_t* state, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMREC_CHECK_ARG(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world && flags && state, "peer_barrier: bad arguments");
    PeerFlags F;
    for (int r = 0; r < PEER_MAX; ++r) F.f[r] = nullptr;
    for (int r = 0; r < world; ++r) {
        MMREC_CHECK_ARG(flags[r], "peer_barrier: null flag pointer");
        F.f[r] = (int*)flags[r];
    }
    peer_barrier_kernel<<<1, 32, 0, stream>>>(world, rank, F, state);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}
