// Micro-benchmark: TMEM load/store throughput per SM (tcgen05.ld / tcgen05.st 32x32b), by number of warps.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>   // 0: ld.x16  1: ld.x32  2: st.x16  3: ld.x16 + st.x16  4: two ld.x16 in flight
__global__ void __launch_bounds__(512, 1) bw(int iters, long long* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = slot;
    const uint32_t addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 64;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 3 || MODE == 4) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(addr + (it & 1) * 16));
            if (MODE == 4)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                               "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                             : "r"(addr + 32));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] + v[15] + (MODE == 4 ? v[16] : 0);
        }
        if (MODE == 1) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] + v[31];
        }
        if (MODE == 2 || MODE == 3) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                         ::"r"(addr + 32), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                           "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
            if (MODE == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x * 2] = t1 - t0;
    if (acc == 0x12345678) out[1] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
template <int MODE>
void run(const char* name, int bytes_per_iter_per_warp) {
    long long* d;
    cudaMalloc(&d, 64);
    const int iters = 2000;
    for (int warps : {1, 4, 8, 16}) {
        bw<MODE><<<1, warps * 32>>>(iters, d);
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("%-22s warps %2d: %7.1f clk/iter/warp  -> %6.1f B/clk/SM  %s\n", name, warps, (double)h[0] / iters,
               (double)bytes_per_iter_per_warp * warps * iters / h[0], e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    cudaFree(d);
}
int main() {
    run<0>("ld.x16 + wait", 2048);
    run<1>("ld.x32 + wait", 4096);
    run<4>("2 x ld.x16 + wait", 4096);
    run<2>("st.x16 + wait", 2048);
    run<3>("ld.x16,wait,st.x16", 4096);
    return 0;
}
