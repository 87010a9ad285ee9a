// Micro-benchmark 3: tcgen05.mma issue rate from a converged warp with elect.sync (vs a divergent lane-0 branch).
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
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// variant 0: divergent lane 0; 1: converged warp + elect; N = 128 or 256
__global__ void __launch_bounds__(128, 1) bench(int variant, int N, int tiles, long long* out) {
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
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t b_s = smem_u32(smem);
    if (threadIdx.x < 32 && (variant == 1 || threadIdx.x == 0)) {
        uint32_t stage = 0;
        long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            const uint32_t sb = b_s + stage * 16384;
            const uint32_t acol = 256u + (t & 3) * 32;
            if (variant == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_ts(tmem, tmem + acol + 8 * j, make_desc(sb + 32 * j), idesc, (t & 7) || j ? 1u : 0u);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[stage])) : "memory");
            } else {
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_ts(tmem, tmem + acol + 8 * j, make_desc(sb + 32 * j), idesc, (t & 7) || j ? 1u : 0u);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[stage])) : "memory");
                }
                __syncwarp();
            }
            if (++stage == 5) stage = 0;
        }
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[14])) : "memory");
            long long t1 = clock64();
            while (!mbar_try(smem_u32(&bars[14]), 0)) {}
            long long t2 = clock64();
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
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
    for (int variant : {0, 1})
        for (int N : {128, 256}) {
            bench<<<1, 128, 100 * 1024>>>(variant, N, tiles, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2];
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("variant=%d N=%d: issue %.1f clk/tile, complete %.1f clk/tile (pipe ideal %d)  %s\n", variant, N,
                   (double)h[0] / tiles, (double)h[1] / tiles, N * 2, cudaGetErrorString(e));
        }
    return 0;
}
