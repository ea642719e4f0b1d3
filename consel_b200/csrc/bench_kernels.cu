// Roofline denominators that MEASURED_PEAKS.json does not carry: the FP64 FMA pipe.
// (bench.py reports the contraction kernel against this number, measured in the
// same process on the same GPU; HBM and bf16 peaks come from the driver's file.)
#include "cgpu_internal.h"

namespace cgpu {

__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (double)(threadIdx.x + i) * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = __fma_rn(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    if (s == 123.456) out[0] = s;  // keep the loop alive
}

int bench_i8_peak(double *tops, cudaStream_t st);  // i8_kernels.cu
int bench_i8_peak_sustained(double seconds, double *tops, cudaStream_t st);

}  // namespace cgpu

// dense int8 tensor-core issue rate (TOP/s) of tcgen05.mma kind::i8 from shared-memory-resident
// operands, measured on this device: the denominator of the GEMM's roofline fraction
extern "C" int cgpu_bench_i8_peak(double *tops) {
    using namespace cgpu;
    CGPU_TRY(ensure_init());
    if (!tops) return CGPU_EINVAL;
    return bench_i8_peak(tops, G().stream);
}

extern "C" int cgpu_bench_i8_peak_sustained(double seconds, double *tops) {
    using namespace cgpu;
    CGPU_TRY(ensure_init());
    if (!tops || !(seconds > 0) || seconds > 30) return CGPU_EINVAL;
    return bench_i8_peak_sustained(seconds, tops, G().stream);
}

extern "C" int cgpu_bench_fp64_peak(double *tflops) {
    using namespace cgpu;
    CGPU_TRY(ensure_init());
    cudaStream_t st = G().stream;
    double *d = nullptr;
    CGPU_CUDA(dev_alloc(&d, 8));
    const int iters = 4096, blocks = G().sm_count * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a, st);
        fp64_peak_kernel<<<blocks, 256, 0, st>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(b, st);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    CGPU_LAUNCH_CHECK();
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    dev_free(d);
    double flops = 2.0 * 16 * (double)iters * 256.0 * blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    return CGPU_OK;
}
