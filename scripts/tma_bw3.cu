// Micro-benchmark: cp.async.bulk throughput per SM vs copy size, depth and number of issuing warps.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// each of `nw` warps streams its own contiguous share with copies of `csize` bytes, `depth` in flight
__global__ void __launch_bounds__(256, 1) k(const uint8_t* src, size_t bytes, int csize, int depth, int nw, long long* out, int lanes) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[8][8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 64; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0][0] + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long t0 = clock64();
    const int w = lanes ? (int)threadIdx.x : (int)(threadIdx.x >> 5);
    if (w < nw && (lanes ? threadIdx.x < 32 : (threadIdx.x & 31) == 0)) {
        const size_t share = bytes / nw;
        const uint8_t* base = src + w * share;
        const size_t n = share / csize;
        uint8_t* sm = smem + (size_t)w * depth * csize;
        size_t issued = 0, done = 0;
        uint32_t ph = 0;
        while (done < n) {
            while (issued < n && issued - done < (size_t)depth) {
                int s = issued % depth;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[w][s])), "r"(csize) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + (size_t)s * csize)), "l"(base + issued * csize), "r"(csize), "r"(smem_u32(&bars[w][s])) : "memory");
                ++issued;
            }
            int s = done % depth;
            while (!mbar_try(smem_u32(&bars[w][s]), (ph >> s) & 1)) {}
            ph ^= 1u << s;
            ++done;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
int main() {
    const size_t bytes = 16u << 20;
    uint8_t* src;
    cudaMalloc(&src, bytes);
    cudaMemset(src, 1, bytes);
    long long* d;
    cudaMalloc(&d, 8 * 148);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct C { int csize, depth, nw; } cs[] = {{16384, 4, 1}, {32768, 4, 1}, {8192, 8, 1}, {4096, 8, 1}, {2048, 8, 1},
                                               {16384, 4, 2}, {8192, 4, 2}, {8192, 4, 4}, {4096, 4, 8}, {16384, 2, 4}, {2048, 8, 8}};
    for (int grid : {1, 64})
        for (auto c : cs) {
            for (int rep = 0; rep < 2; ++rep) k<<<grid, 256, 200 * 1024>>>(src, bytes, c.csize, c.depth, c.nw, d, 1);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148], mx = 0;
            cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
            for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("grid=%3d copy=%5d B depth=%d warps=%d: %.1f B/clk/SM  (%s)\n", grid, c.csize, c.depth, c.nw, (double)bytes / mx, cudaGetErrorString(e));
        }
    return 0;
}
