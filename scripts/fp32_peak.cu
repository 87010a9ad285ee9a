// FP32 FMA micro-benchmark: the denominator of the energy / sweep kernels' roofline (SURVEY.md 8d asks for a measured
// figure; MEASURED_PEAKS.json has none for FP32).  Every thread runs 16 independent dependent-FMA chains; variant 0
// issues scalar FFMA, variant 1 the packed fma.rn.f32x2 (FFMA2) the kernels use.  Built by __graft_entry__.build() into
// scripts/libfp32_peak.so and called by bench.py through ctypes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o libfp32_peak.so fp32_peak.cu
#include <cuda_runtime.h>

template <int VARIANT>
__global__ void __launch_bounds__(256) fma_kernel(float* out, int iters, float b, float c) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
    if (VARIANT == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
        }
    } else {
        unsigned long long p[8], bb, cc;
        asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
        asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p[i]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;      // never true: keeps the chains alive
}

// Returns 0 on success; *tflops = 2 * FMAs / best-of-5 kernel time (CUDA events), *ms that time.
extern "C" int fp32_peak(int device, int variant, double* tflops, double* ms) {
    if (cudaSetDevice(device) != cudaSuccess) return 1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 1;
    float* out = nullptr;
    if (cudaMalloc(&out, 64) != cudaSuccess) return 1;
    const int iters = 1 << 14, threads = 256, blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        if (variant == 0) fma_kernel<0><<<blocks, threads>>>(out, iters, 0.999f, 1e-4f);
        else fma_kernel<1><<<blocks, threads>>>(out, iters, 0.999f, 1e-4f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return 2;
        float t;
        cudaEventElapsedTime(&t, e0, e1);
        if (rep > 0 && t < best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    const double fmas = (double)blocks * threads * 16.0 * iters;
    *tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    *ms = best;
    return 0;
}
