// Micro-benchmark: cycles per tcgen05.mma for TS / SS operand modes, tf32 / bf16, N = 64..256.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// mode 0: TS tf32, 1: SS tf32, 2: TS bf16(f16 kind), 3: SS bf16
template <int MODE>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_t, uint64_t a_d, uint64_t b_d, uint32_t idesc, uint32_t acc) {
    if (MODE == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(b_d), "r"(idesc), "r"(acc) : "memory");
    else if (MODE == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_d), "l"(b_d), "r"(idesc), "r"(acc) : "memory");
    else if (MODE == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(b_d), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_d), "l"(b_d), "r"(idesc), "r"(acc) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int same_d, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const bool bf = MODE >= 2;
        const uint32_t fmt = bf ? 1u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 32 * 1024);
        const int kcols = bf ? 8 : 8;   // TMEM columns per k-step of A (bf16: 16 elems packed 2 per column)
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int k = it & 3;
            const uint32_t dcol = same_d ? 0u : (uint32_t)(((it >> 5) & 1) * 256);
            mma<MODE>(tmem + dcol, tmem + 256 + (uint32_t)((it & 15) * kcols), make_desc(a_s + 32 * k),
                      make_desc(b_s + 32 * k), idesc, (it & 31) ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        out[2 * blockIdx.x] = t1 - t0;
        out[2 * blockIdx.x + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int MODE>
void run(const char* name, int N, int same_d, int grid) {
    long long* d;
    cudaMalloc(&d, sizeof(long long) * 2 * grid);
    const int iters = 4096;
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    bench<MODE><<<grid, 128, 100 * 1024>>>(N, iters, same_d, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double flop = 2.0 * 128 * N * (MODE >= 2 ? 16 : 8);
    printf("%-10s N=%3d same_d=%d grid=%3d: issue %.1f clk/mma, complete %.1f clk/mma -> %.0f flop/clk/SM  (%s)\n", name, N,
           same_d, grid, (double)h[0] / iters, (double)h[1] / iters, flop * iters / h[1], cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int N : {64, 128, 256}) {
        run<0>("TS tf32", N, 1, 1);
        run<1>("SS tf32", N, 1, 1);
        run<2>("TS bf16", N, 1, 1);
        run<3>("SS bf16", N, 1, 1);
    }
    run<0>("TS tf32", 128, 0, 1);
    run<0>("TS tf32", 128, 1, 148);
    run<1>("SS tf32", 256, 1, 148);
    return 0;
}
