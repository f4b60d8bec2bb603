// Roofline denominators measured on the box the bench runs on ("derive the FP32 peak ... and confirm with an FMA
// micro-benchmark in the same run").  Two probes, timed with CUDA events on the caller's stream:
//   * FP32: every thread runs 16 independent FFMA chains (no memory traffic), grid = 148 SMs x 8 CTAs of 256 threads;
//   * issue: the same loop shape as the tick kernel's argmin (packed FFMA2), to show what two FMAs per issue slot reach.
// MEASURED_PEAKS.json holds HBM and bf16 tensor peaks only; the analytic-dynamics kernels are FP32-issue-bound.
#include <cuda_runtime.h>

#include "../../include/mppi_b200.h"

namespace {

template <bool PACKED>
__global__ void __launch_bounds__(256) fma_probe_kernel(float *out, int iters, float a, float b) {
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
        if (PACKED) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long x = ((unsigned long long)__float_as_uint(r[i + 1]) << 32) | __float_as_uint(r[i]);
                const unsigned long long aa = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
                const unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(aa), "l"(bb));
                r[i] = __uint_as_float((unsigned)x); r[i + 1] = __uint_as_float((unsigned)(x >> 32));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(a), "f"(b));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    if (s == 123.456f) out[0] = s;                    // keeps the chains alive; never true
}

// the hardware tanh the learned-dynamics kernel applies (MUFU.TANH via tanh.approx.f32), exposed so the parity tests can
// state how far it is from the exact function (PTX: max relative error 2^-11)
__global__ void tanh_probe_kernel(const float *in, float *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(in[i]));
    out[i] = y;
}

}  // namespace

extern "C" int mppi_probe_tanh(int32_t device, const float *in, float *out, int32_t n) {
    if (!in || !out || n < 1) return MPPI_E_BADARG;
    if (cudaSetDevice(device) != cudaSuccess) return MPPI_E_CUDA;
    float *d = nullptr;
    if (cudaMalloc(&d, sizeof(float) * 2 * (size_t)n) != cudaSuccess) return MPPI_E_NOMEM;
    cudaError_t e = cudaMemcpy(d, in, sizeof(float) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        tanh_probe_kernel<<<(n + 255) / 256, 256>>>(d, d + n, n);
        e = cudaMemcpy(out, d + n, sizeof(float) * n, cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    return e == cudaSuccess ? MPPI_OK : MPPI_E_CUDA;
}

extern "C" int mppi_probe_fp32_peak(int32_t device, int32_t packed, double *tflops_out) {
    if (!tflops_out) return MPPI_E_BADARG;
    if (cudaSetDevice(device) != cudaSuccess) return MPPI_E_CUDA;
    int n_sm = 0;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return MPPI_E_CUDA;
    float *d = nullptr;
    if (cudaMalloc(&d, sizeof(float)) != cudaSuccess) return MPPI_E_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = n_sm * 8, iters = 1 << 15;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {                // first repetition is the warm-up
        cudaEventRecord(e0);
        if (packed) fma_probe_kernel<true><<<grid, 256>>>(d, iters, 0.999f, 1e-3f);
        else fma_probe_kernel<false><<<grid, 256>>>(d, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return MPPI_E_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 16.0 * (double)iters * 256.0 * grid / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *tflops_out = best;
    return cudaGetLastError() == cudaSuccess ? MPPI_OK : MPPI_E_CUDA;
}
