// K2: fused gather -> linear(+bias) -> optional L2 normalise (exact-fp32 CUDA-core path).
// The tensor-core (tcgen05, 3xTF32) variant lives in project_tc.cu; this file is the bit-faithful
// fp32 path and the fallback for widths the tensor-core tiles do not cover.
#include "gemm_simt.cuh"

#include <stdlib.h>
#include <string.h>

using namespace mmrec;

namespace mmrec {
// project_tc.cu
int project_tc(int64_t n_out, const int64_t* idx, const float* table, int64_t F, const float* W, const float* bias, int d,
               int l2_normalize, float* Y, int64_t ldy, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t project_tc_workspace_bytes(int64_t n_out, int64_t F, int d);

static int g_project_path = -1;   // 0 = exact fp32 CUDA cores, 1 = tcgen05 3xTF32 (default)
static int project_path() {
    if (g_project_path < 0) {
        const char* e = getenv("MMREC_PROJECT_PATH");
        g_project_path = (e && strcmp(e, "simt") == 0) ? 0 : 1;
    }
    return g_project_path;
}

// any d: one CTA per output row, threads stride over the d outputs (correctness path for odd widths)
__global__ void project_generic_kernel(int64_t n_out, const int64_t* __restrict__ idx, const float* __restrict__ table,
                                       int64_t F, const float* __restrict__ W, const float* __restrict__ bias, int d,
                                       int l2, float* __restrict__ Y, int64_t ldy) {
    extern __shared__ float sh[];   // d floats
    __shared__ float red[32];
    const int64_t n = blockIdx.x;
    const float* x = table + (idx ? idx[n] : n) * F;
    float ss = 0.f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const float* w = W + (int64_t)c * F;
        float acc = 0.f;
        for (int64_t k = 0; k < F; ++k) acc = fmaf(x[k], w[k], acc);
        if (bias) acc += bias[c];
        sh[c] = acc;
        ss += acc * acc;
    }
    if (l2) {
        ss = warp_sum(ss);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
            v = warp_sum(v);
            if (threadIdx.x == 0) red[0] = v;
        }
        __syncthreads();
    } else {
        __syncthreads();
    }
    const float inv = l2 ? 1.0f / fmaxf(sqrtf(red[0]), 1e-12f) : 1.0f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) Y[n * ldy + c] = l2 ? sh[c] * inv : sh[c];
}
}  // namespace mmrec

extern "C" int mmrec_project_set_path(int tensor_core) { g_project_path = tensor_core ? 1 : 0; return MMREC_OK; }

extern "C" size_t mmrec_project_workspace_bytes(int64_t n_out, int64_t F, int d) {
    return project_tc_workspace_bytes(n_out, F, d) + 256;
}

extern "C" int mmrec_project_f32(int64_t n_out, const int64_t* idx, const float* table, int64_t n_table, int64_t F,
                                 const float* W, const float* bias, int d, int l2_normalize, float* Y, int64_t ldy,
                                 void* ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMREC_CHECK_ARG(n_out >= 0 && n_table >= 0 && F >= 1 && d >= 1, "project: bad sizes");
    if (n_out == 0) return MMREC_OK;
    MMREC_CHECK_ARG(table && W && Y && ldy >= d, "project: null pointer or ldy < d");
    if (project_path() == 1 && ws) {
        int r = project_tc(n_out, idx, table, F, W, bias, d, l2_normalize, Y, ldy, ws, ws_bytes, stream);
        if (r != 0) return r < 0 ? r : MMREC_OK;
    }
    GemmNT p;
    p.A = table; p.lda = F; p.a_idx = idx; p.M = n_out;
    p.B = W; p.ldb = F; p.N = d; p.K = F; p.bias = bias; p.C = Y; p.ldc = ldy; p.l2_normalize = l2_normalize;
    if (d <= 32) return launch_gemm_nt<64, 32, 4, 2>(p, stream);
    if (d <= 64) return launch_gemm_nt<32, 64, 2, 4>(p, stream);
    if (d <= 128) return launch_gemm_nt<32, 128, 2, 8>(p, stream);
    if (d <= 256) return launch_gemm_nt<32, 256, 2, 16>(p, stream);
    project_generic_kernel<<<(unsigned)n_out, 128, (size_t)d * sizeof(float), stream>>>(n_out, idx, table, F, W, bias, d,
                                                                                   l2_normalize, Y, ldy);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}
