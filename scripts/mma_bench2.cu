// Micro-benchmark 2: cost of the per-tile bookkeeping around tcgen05.mma (commit / fence / try_wait).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFF) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_t, uint64_t b_d, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(b_d), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// flags: 1 commit every 4, 2 try_wait (already complete) every 4, 4 fence every 4, 8 rotate B stage + D like the kernel,
//        16 wait for the commit barrier of the tile issued 5 tiles ago (stage recycling)
__global__ void __launch_bounds__(128, 1) bench(int flags, int tiles, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[16];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
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
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | (8u << 24);
        const uint32_t b_s = smem_u32(smem);
        // bars[15] is completed once so try_wait(parity 0) succeeds immediately
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[15])) : "memory");
        uint32_t stage = 0, phase = 0;
        long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            if (flags & 16) {
                // recycle: wait until the MMAs that used this stage 5 tiles ago have completed
                if (t >= 5) while (!mbar_try(smem_u32(&bars[stage]), phase ^ 1)) {}
            }
            if (flags & 2) while (!mbar_try(smem_u32(&bars[15]), 0)) {}
            if (flags & 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sb = (flags & 8) ? b_s + stage * 16384 : b_s;
            const uint32_t dcol = (flags & 8) ? (uint32_t)(((t >> 3) & 3) * 128) : 0u;
            const uint32_t acol = (flags & 8) ? (uint32_t)((((t >> 3) & 2) ^ 2) * 128 + (t & 7) * 32) & 511u : 256u + (t & 3) * 32;
            for (int j = 0; j < 4; ++j) mma_ts(tmem + dcol, tmem + acol + 8 * j, make_desc(sb + 32 * j), idesc, (t & 7) || j ? 1u : 0u);
            if (flags & (1 | 16))
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[stage])) : "memory");
            if (++stage == 5) { stage = 0; phase ^= 1; }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[14])) : "memory");
        long long t1 = clock64();
        while (!mbar_try(smem_u32(&bars[14]), 0)) {}
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int tiles = 1024;
    for (int flags : {0, 1, 2, 4, 8, 1 | 2 | 4, 1 | 2 | 4 | 8, 16, 16 | 2 | 4 | 8}) {
        bench<<<1, 128, 100 * 1024>>>(flags, tiles, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("flags=%2d: issue %.1f clk/tile, complete %.1f clk/tile (ideal 256)  %s\n", flags, (double)h[0] / tiles,
               (double)h[1] / tiles, cudaGetErrorString(e));
    }
    return 0;
}
