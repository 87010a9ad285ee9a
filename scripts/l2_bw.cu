// Micro-benchmark: per-SM L2->SM bandwidth by access path (all CTAs read the same 16 MB, L2-resident).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// mode 0: LDG.128 (ld.global.nc) by all threads; 1: cp.async 16 B into smem; 2: TMA bulk 16 KB x depth 4 from warp 0 + LDG by the others
__global__ void __launch_bounds__(256, 1) bw(const uint4* src, size_t n16, int mode, long long* out, uint4* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[4];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long t0 = clock64();
    uint4 acc = make_uint4(0, 0, 0, 0);
    if (mode == 0 || (mode == 2 && threadIdx.x >= 32)) {
        const int nt = mode == 0 ? 256 : 224, t = mode == 0 ? threadIdx.x : threadIdx.x - 32;
        const size_t lim = mode == 0 ? n16 : n16 / 2;
        const uint4* p = mode == 0 ? src : src + n16 / 2;
        for (size_t i = t; i + 7 * nt < lim; i += 8 * nt) {
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(p + i + j * nt);
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc.x ^= v[j].x; acc.y ^= v[j].y; acc.z ^= v[j].z; acc.w ^= v[j].w; }
        }
    } else if (mode == 1) {
        for (size_t i = threadIdx.x; i + 7 * 256 < n16; i += 8 * 256) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem + (((i / 256) & 7) * 8 + j) * 4096 + threadIdx.x * 16)), "l"(src + i + j * 256) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 2;" ::: "memory");
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (mode == 2 && threadIdx.x == 0) {
        const int tile = 16384;
        const size_t ntiles = (n16 / 2) * 16 / tile;
        size_t issued = 0, done = 0;
        uint32_t phase[4] = {0, 0, 0, 0};
        while (done < ntiles) {
            while (issued < ntiles && issued - done < 4) {
                int s = issued & 3;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(tile) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + s * tile)), "l"((const uint8_t*)src + issued * tile), "r"(tile), "r"(smem_u32(&bars[s])) : "memory");
                ++issued;
            }
            int s = done & 3;
            while (!mbar_try(smem_u32(&bars[s]), phase[s])) {}
            phase[s] ^= 1;
            ++done;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc.x == 0x12345678) sink[0] = acc;
}
int main() {
    const size_t bytes = 16u << 20, n16 = bytes / 16;
    uint4 *src, *sink;
    cudaMalloc(&src, bytes);
    cudaMalloc(&sink, 64);
    cudaMemset(src, 1, bytes);
    long long* d;
    cudaMalloc(&d, 8 * 148);
    cudaFuncSetAttribute(bw, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    const char* names[] = {"LDG.128", "cp.async16", "TMA+LDG"};
    for (int grid : {1, 64, 148})
        for (int mode = 0; mode < 3; ++mode) {
            for (int rep = 0; rep < 2; ++rep) bw<<<grid, 256, 128 * 1024>>>(src, n16, mode, d, sink);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148], mx = 0;
            cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
            for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("grid=%3d %-10s: %.1f B/clk/SM  (%s)\n", grid, names[mode], (double)bytes / mx, cudaGetErrorString(e));
        }
    return 0;
}
