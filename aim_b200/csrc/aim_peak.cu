// Integer-pipe ceiling, measured on the device the aligner runs on (SURVEY.md section 8d asks for a
// measured INT32 peak, not a datasheet one).  A dependent-free mix of the three instruction kinds
// the alignment kernels are made of - add (IADD3), logic (LOP3) and min/max (VIMNMX) - over eight
// independent accumulator chains per thread, every SM filled.
#include <cuda_runtime.h>

#include <string>

#include "aim_internal.h"

namespace aim {
namespace {
__global__ void __launch_bounds__(256) int_peak_kernel(int *out, int iters, int c0, int c1)
{
    int a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (int)threadIdx.x * (j + 1) + c1;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a[j] = a[j] + c0;          // add
                a[j] = a[j] ^ (c1 + r);    // logic
                a[j] = max(a[j], c1 - j);  // min/max
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 0x7fffffff) out[0] = s;  // keep the chains alive
}
}  // namespace
}  // namespace aim

// Returns AIM_OK and the sustained rate in int32 operations per second (source-level ops: 3 per
// statement above), timed with CUDA events after a warm-up launch.
extern "C" int aim_measure_int_peak(int device, double *ops_per_s)
{
    using namespace aim;
    if (!ops_per_s) return AIM_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return AIM_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { set_error("cudaGetDeviceProperties failed"); return AIM_ERR_CUDA; }
    int *out = nullptr;
    if (cudaMalloc(&out, 4) != cudaSuccess) return AIM_ERR_NOMEM;
    const int iters = 4096, grid = prop.multiProcessorCount * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int_peak_kernel<<<grid, block>>>(out, iters, 3, 5);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        int_peak_kernel<<<grid, block>>>(out, iters, 3 + rep, 5);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (err != cudaSuccess) { set_error(std::string("int peak: ") + cudaGetErrorString(err)); return AIM_ERR_CUDA; }
    const double ops = (double)grid * block * (double)iters * 4.0 * 8.0 * 3.0;
    *ops_per_s = ops / ((double)best * 1e-3);
    return AIM_OK;
}
