// debug.cu — measurement helpers that are not on the product path.
//
// ia2c_debug_fp32_peak: the FP32 issue peak of this GPU, measured the way the MLP kernels use the pipe.  The actor /
// critic networks are 6-wide (ac_nets.py:24), so no op is a dense contraction and tensor cores do not apply
// (SURVEY.md §7.3): the bound for rollout_fused_kernel / actor_grad_kernel is the CUDA-core fma pipe.  BASELINE.md §3
// asks for a MEASURED denominator for those kernels — this kernel issues nothing but independent fma chains from
// registers at full occupancy: `packed` != 0 uses fma.rn.f32x2 (SASS FFMA2, two IEEE FMAs per lane per issue, the
// instruction mlp_f2.cuh is written in), 0 uses scalar fma.rn.f32 (FFMA).  bench.py turns the event-timed duration
// into TFLOP/s and reports the two MLP-bound kernels as a fraction of it.
#include "common.cuh"

namespace ia2c {
namespace {

constexpr int kPeakThreads = 256, kPeakChains = 8;

template <bool PACKED>
__global__ void __launch_bounds__(kPeakThreads) fp32_peak_kernel(float* __restrict__ out, int iters, float a, float b) {
    float2 acc[kPeakChains];
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) acc[k] = make_float2((float)(threadIdx.x + k), (float)(blockIdx.x - k));
    const float2 a2 = make_float2(a, a * 0.5f), b2 = make_float2(b, -b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int k = 0; k < kPeakChains; ++k) {
                if (PACKED) {
                    acc[k] = __ffma2_rn(acc[k], a2, b2);
                } else {
                    acc[k].x = fmaf(acc[k].x, a2.x, b2.x);
                    acc[k].y = fmaf(acc[k].y, a2.y, b2.y);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) s += acc[k].x + acc[k].y;
    if (s == 12345.678f) out[0] = s;   // never true in practice: keeps the chains alive
}

}  // namespace
}  // namespace ia2c

using namespace ia2c;

// flops_out (host, may be NULL) receives the FLOPs one launch performs: blocks * 256 threads * iters * 8 * 8 chains * 2 lanes * 2.
extern "C" int ia2c_debug_fp32_peak(float* out, int32_t iters, int32_t packed, int32_t blocks_per_sm, double* host_flops_out,
                                    void* stream) {
    IA2C_REQUIRE(out && iters > 0 && blocks_per_sm > 0 && blocks_per_sm <= 8, "ia2c_debug_fp32_peak: bad arguments");
    const int blocks = kSMs * blocks_per_sm;
    if (packed) fp32_peak_kernel<true><<<blocks, kPeakThreads, 0, as_stream(stream)>>>(out, iters, 0.999f, 0.001f);
    else fp32_peak_kernel<false><<<blocks, kPeakThreads, 0, as_stream(stream)>>>(out, iters, 0.999f, 0.001f);
    if (host_flops_out) *host_flops_out = (double)blocks * kPeakThreads * (double)iters * 8.0 * kPeakChains * 2.0 * 2.0;
    return check_launch("fp32_peak_kernel");
}
