// K3 tensor-core path (tcgen05, 3xTF32 split).  Placeholder until the kernel lands: reports "unsupported"
// so that score.cu uses the exact fp32 CUDA-core path.
#include "common.cuh"
namespace mmrec {
int score_tc(int64_t, const int64_t*, const float*, int64_t, int64_t, const float*, int64_t, int, float*, int64_t,
             cudaStream_t) {
    return 0;
}
}  // namespace mmrec
