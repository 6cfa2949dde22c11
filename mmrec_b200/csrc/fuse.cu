// a5b -- MGCN's row-wise fusion (src/models/mgcn.py:153-154,187-201) as two kernels instead of ~15 eager element-wise /
// small-GEMM launches over [n_users + n_items, d] tensors:
//
//   gate_rows   out[n,:] = mul[n,:] * sigmoid(X[n,:] W^T + b)              -- `item_id_embedding.weight * gate_v(image_feats)`
//   mgcn_fuse   per row: attention of the two modality views (query_common: Linear -> Tanh -> Linear(d,1), softmax over the
//               two logits), common = w0 img + w1 txt, preference gates on the content embedding, side = (sep_img + sep_txt +
//               common) / 3, out = content + side.
//
// Every row is independent; the d x d weights live in shared memory (transposed, padded), a warp owns 4 rows at a time, a
// lane owns d/32 output features.  Bandwidth-trivial (3 reads + 1 write of [N, d]); the point is the launch count and the
// intermediate tensors that no longer exist.  fp32 fmaf chains, tanhf / expf of the CUDA math library, IEEE division.
#include <cuda_runtime.h>

#include "common.cuh"

namespace mmrec {

constexpr int FR = 4;                 // rows per warp iteration

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Wt[k * (D + 1) + j] = W[j * D + k]
template <int D>
__device__ __forceinline__ void load_weight_t(float* Wt, const float* __restrict__ W) {
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int j = e / D, k = e - j * D;
        Wt[k * (D + 1) + j] = __ldg(W + e);
    }
}

// rows row0 .. row0+3 of X (leading dimension D) -> xs[k * FR + r]; rows beyond n read as zero
template <int D>
__device__ __forceinline__ void stage_rows(float* xs, const float* __restrict__ X, int64_t row0, int64_t n, int lane) {
#pragma unroll
    for (int r = 0; r < FR; ++r)
#pragma unroll
        for (int jt = 0; jt < D / 32; ++jt) {
            const int k = lane + 32 * jt;
            xs[k * FR + r] = (row0 + r < n) ? __ldg(X + (row0 + r) * D + k) : 0.f;
        }
}

template <int D>
__global__ void __launch_bounds__(256) gate_rows_kernel(int64_t n, const float* __restrict__ X, const float* __restrict__ W,
                                                        const float* __restrict__ b, const float* __restrict__ mul, float* __restrict__ out) {
    constexpr int JT = D / 32, LD = D + 1;
    extern __shared__ __align__(16) float fsm[];
    float* Wt = fsm;                               // [D][LD]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    load_weight_t<D>(Wt, W);
    float* xs = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(Wt + D * LD) + 15) & ~uintptr_t(15)) + warp * D * FR;   // 16-byte aligned row tiles
    float bj[JT];
#pragma unroll
    for (int jt = 0; jt < JT; ++jt) bj[jt] = b ? __ldg(b + lane + 32 * jt) : 0.f;
    __syncthreads();
    for (int64_t row0 = ((int64_t)blockIdx.x * n_warps + warp) * FR; row0 < n; row0 += (int64_t)gridDim.x * n_warps * FR) {
        stage_rows<D>(xs, X, row0, n, lane);
        __syncwarp();
        float acc[JT][FR];
#pragma unroll
        for (int jt = 0; jt < JT; ++jt)
#pragma unroll
            for (int r = 0; r < FR; ++r) acc[jt][r] = bj[jt];
#pragma unroll 8
        for (int k = 0; k < D; ++k) {
            const float4 x = *reinterpret_cast<const float4*>(xs + k * FR);
#pragma unroll
            for (int jt = 0; jt < JT; ++jt) {
                const float w = Wt[k * LD + lane + 32 * jt];
                acc[jt][0] = fmaf(w, x.x, acc[jt][0]);
                acc[jt][1] = fmaf(w, x.y, acc[jt][1]);
                acc[jt][2] = fmaf(w, x.z, acc[jt][2]);
                acc[jt][3] = fmaf(w, x.w, acc[jt][3]);
            }
        }
#pragma unroll
        for (int r = 0; r < FR; ++r) {
            if (row0 + r >= n) break;
#pragma unroll
            for (int jt = 0; jt < JT; ++jt) {
                const int64_t o = (row0 + r) * D + lane + 32 * jt;
                const float g = sigmoidf_(acc[jt][r]);
                out[o] = mul ? __ldg(mul + o) * g : g;
            }
        }
        __syncwarp();
    }
}

struct MgcnFuseParams {
    int64_t n;
    const float *img, *txt, *content;
    const float *Wq, *bq, *wq2;      // query_common: Linear(d,d) + Tanh + Linear(d,1,bias=False)
    const float *Wgi, *bgi, *Wgt, *bgt;   // gate_image_prefer / gate_text_prefer: Linear(d,d) + Sigmoid
    float *out, *side;               // side nullable
};

template <int D>
__global__ void __launch_bounds__(256) mgcn_fuse_kernel(const MgcnFuseParams p) {
    constexpr int JT = D / 32, LD = D + 1;
    extern __shared__ __align__(16) float fsm[];
    float* Wq = fsm;
    float* Wi = Wq + D * LD;
    float* Wt = Wi + D * LD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    float* xs = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(Wt + D * LD) + 15) & ~uintptr_t(15)) + warp * 3 * D * FR;
    float* xi = xs;
    float* xt = xs + D * FR;
    float* xc = xs + 2 * D * FR;
    load_weight_t<D>(Wq, p.Wq);
    load_weight_t<D>(Wi, p.Wgi);
    load_weight_t<D>(Wt, p.Wgt);
    float bq[JT], bi[JT], bt[JT], w2[JT];
#pragma unroll
    for (int jt = 0; jt < JT; ++jt) {
        const int j = lane + 32 * jt;
        bq[jt] = p.bq ? __ldg(p.bq + j) : 0.f;
        bi[jt] = p.bgi ? __ldg(p.bgi + j) : 0.f;
        bt[jt] = p.bgt ? __ldg(p.bgt + j) : 0.f;
        w2[jt] = __ldg(p.wq2 + j);
    }
    __syncthreads();
    for (int64_t row0 = ((int64_t)blockIdx.x * n_warps + warp) * FR; row0 < p.n; row0 += (int64_t)gridDim.x * n_warps * FR) {
        stage_rows<D>(xi, p.img, row0, p.n, lane);
        stage_rows<D>(xt, p.txt, row0, p.n, lane);
        stage_rows<D>(xc, p.content, row0, p.n, lane);
        __syncwarp();
        float hi[JT][FR], ht[JT][FR], gi[JT][FR], gt[JT][FR];
#pragma unroll
        for (int jt = 0; jt < JT; ++jt)
#pragma unroll
            for (int r = 0; r < FR; ++r) { hi[jt][r] = bq[jt]; ht[jt][r] = bq[jt]; gi[jt][r] = bi[jt]; gt[jt][r] = bt[jt]; }
#pragma unroll 4
        for (int k = 0; k < D; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(xi + k * FR);
            const float4 b = *reinterpret_cast<const float4*>(xt + k * FR);
            const float4 c = *reinterpret_cast<const float4*>(xc + k * FR);
            const float av[FR] = {a.x, a.y, a.z, a.w}, bv[FR] = {b.x, b.y, b.z, b.w}, cv[FR] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int jt = 0; jt < JT; ++jt) {
                const int j = lane + 32 * jt;
                const float wq = Wq[k * LD + j], wi = Wi[k * LD + j], wt = Wt[k * LD + j];
#pragma unroll
                for (int r = 0; r < FR; ++r) {
                    hi[jt][r] = fmaf(wq, av[r], hi[jt][r]);
                    ht[jt][r] = fmaf(wq, bv[r], ht[jt][r]);
                    gi[jt][r] = fmaf(wi, cv[r], gi[jt][r]);
                    gt[jt][r] = fmaf(wt, cv[r], gt[jt][r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < FR; ++r) {
            float si = 0.f, st = 0.f;
#pragma unroll
            for (int jt = 0; jt < JT; ++jt) {
                si = fmaf(w2[jt], tanhf(hi[jt][r]), si);
                st = fmaf(w2[jt], tanhf(ht[jt][r]), st);
            }
            si = warp_sum(si);
            st = warp_sum(st);
            const float m = fmaxf(si, st);                              // softmax over the two logits (mgcn.py:189-190)
            const float ei = expf(si - m), et = expf(st - m);
            const float w0 = ei / (ei + et), w1 = et / (ei + et);
            if (row0 + r >= p.n) continue;
#pragma unroll
            for (int jt = 0; jt < JT; ++jt) {
                const int j = lane + 32 * jt;
                const float ximg = xi[j * FR + r], xtxt = xt[j * FR + r], cont = xc[j * FR + r];
                const float common = w0 * ximg + w1 * xtxt;                // mgcn.py:191-192
                const float sep_i = sigmoidf_(gi[jt][r]) * (ximg - common);   // :193-198
                const float sep_t = sigmoidf_(gt[jt][r]) * (xtxt - common);
                const float side = (sep_i + sep_t + common) / 3.f;          // :199
                const int64_t o = (row0 + r) * D + j;
                if (p.side) p.side[o] = side;
                p.out[o] = cont + side;                                     // :201
            }
        }
        __syncwarp();
    }
}

template <int D>
static constexpr int fuse_warps() { return D >= 128 ? 4 : 8; }            // D = 128: three weight matrices take 198 KB of the 227 KB

template <int D>
static size_t fuse_smem(int n_weights, int n_inputs) {
    return (size_t)n_weights * D * (D + 1) * sizeof(float) + 16 + (size_t)fuse_warps<D>() * n_inputs * D * FR * sizeof(float) + 16;
}

static unsigned fuse_grid(int64_t n, int warps) {
    int64_t g = (n + warps * FR - 1) / (warps * FR);
    const int64_t cap = sm_count();                                       // persistent: the weights are loaded once per CTA
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <int D>
static int launch_gate(int64_t n, const float* X, const float* W, const float* b, const float* mul, float* out, cudaStream_t stream) {
    const size_t smem = fuse_smem<D>(1, 1);
    MMREC_CUDA(cudaFuncSetAttribute(gate_rows_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gate_rows_kernel<D><<<fuse_grid(n, fuse_warps<D>()), 32 * fuse_warps<D>(), smem, stream>>>(n, X, W, b, mul, out);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}

template <int D>
static int launch_fuse(const MgcnFuseParams& p, cudaStream_t stream) {
    const size_t smem = fuse_smem<D>(3, 3);
    MMREC_CUDA(cudaFuncSetAttribute(mgcn_fuse_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mgcn_fuse_kernel<D><<<fuse_grid(p.n, fuse_warps<D>()), 32 * fuse_warps<D>(), smem, stream>>>(p);
    MMREC_LAUNCH_CHECK();
    return MMREC_OK;
}

}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_gate_rows_f32(int64_t n, int d, const float* X, const float* W, const float* b, const float* mul, float* out,
                                   void* stream_) {
    MMREC_CHECK_ARG(n >= 0 && d >= 1, "gate_rows: bad sizes");
    if (d != 32 && d != 64 && d != 128) {
        set_error("gate_rows: d = %d has no kernel (32, 64, 128)", d);
        return MMREC_EUNSUPPORTED;
    }
    if (n == 0) return MMREC_OK;
    MMREC_CHECK_ARG(X && W && out, "gate_rows: null pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (d == 32) return launch_gate<32>(n, X, W, b, mul, out, stream);
    if (d == 64) return launch_gate<64>(n, X, W, b, mul, out, stream);
    return launch_gate<128>(n, X, W, b, mul, out, stream);
}

extern "C" int mmrec_mgcn_fuse_f32(int64_t n, int d, const float* img, const float* txt, const float* content, const float* Wq,
                                   const float* bq, const float* wq2, const float* Wgi, const float* bgi, const float* Wgt,
                                   const float* bgt, float* out, float* side, void* stream_) {
    MMREC_CHECK_ARG(n >= 0 && d >= 1, "mgcn_fuse: bad sizes");
    if (d != 32 && d != 64 && d != 128) {
        set_error("mgcn_fuse: d = %d has no kernel (32, 64, 128)", d);
        return MMREC_EUNSUPPORTED;
    }
    if (n == 0) return MMREC_OK;
    MMREC_CHECK_ARG(img && txt && content && Wq && wq2 && Wgi && Wgt && out, "mgcn_fuse: null pointer");
    MgcnFuseParams p{n, img, txt, content, Wq, bq, wq2, Wgi, bgi, Wgt, bgt, out, side};
    cudaStream_t stream = (cudaStream_t)stream_;
    if (d == 32) return launch_fuse<32>(p, stream);
    if (d == 64) return launch_fuse<64>(p, stream);
    return launch_fuse<128>(p, stream);
}
