// Micro-benchmark: latency / throughput of cp.async.bulk global->shared (16 KB tiles) from L2.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__global__ void __launch_bounds__(32, 1) lat(const uint8_t* src, int tile_bytes, int depth, int ntiles, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint8_t* base = src + (size_t)blockIdx.x * 0;   // all CTAs stream the same bytes
        long long t0 = clock64();
        int issued = 0, done = 0;
        uint32_t phase[16] = {0};
        while (done < ntiles) {
            while (issued < ntiles && issued - done < depth) {
                int s = issued % depth;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(tile_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)s * tile_bytes)), "l"(base + (size_t)issued * tile_bytes), "r"(tile_bytes), "r"(smem_u32(&bars[s])) : "memory");
                ++issued;
            }
            int s = done % depth;
            while (!mbar_try(smem_u32(&bars[s]), phase[s])) {}
            phase[s] ^= 1;
            ++done;
        }
        out[blockIdx.x] = clock64() - t0;
    }
}
int main() {
    const int tile = 16384, ntiles = 1024;
    uint8_t* src;
    cudaMalloc(&src, (size_t)tile * ntiles);
    cudaMemset(src, 1, (size_t)tile * ntiles);
    long long* d;
    cudaMalloc(&d, 8 * 148);
    cudaFuncSetAttribute(lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int grid : {1, 32, 64, 128})
        for (int depth : {1, 2, 5, 8, 12}) {
            for (int rep = 0; rep < 2; ++rep) lat<<<grid, 32, 200 * 1024>>>(src, tile, depth, ntiles, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("grid=%3d depth=%2d: %.0f clk per 16 KB tile (max over CTAs) -> %.1f B/clk/SM  %s\n", grid, depth,
                   (double)mx / ntiles, tile * (double)ntiles / mx, cudaGetErrorString(e));
        }
    return 0;
}
