// On-device observables of sampled configurations.
//
// fs_classify_wells  <- classify_particles + the per-configuration reductions of calculate_well_statistics
//                       (hybrid_NF_MCMC/utils.py:61-141): particle in well A / B / outside, "all in A" / "all in
//                       B" flag and mean x per configuration.
// fs_pair_histogram  <- the per-configuration histogram of calculate_pair_correlation
//                       (hybrid_NF_MCMC/utils.py:530-556): minimum-image pair distances in the reference's
//                       float32 arithmetic, np.histogram bins arange(0, bound + dr, dr).
//
// Both are HBM-bound streaming kernels (8N bytes in, N + 6 or 4 nbins bytes out per configuration) with an
// O(N^2) in-register/shared-memory part for the histogram; one block per configuration.
#include "common.cuh"

namespace fs {

// utils.py:107-141.  Circle of radius 1.1 r0 around (L/4, L/2) -> 1 (A), around (3L/4, L/2) -> 2 (B), else 0;
// float64 like the reference (positions are promoted by the Python float centre), np.round = rint.
__global__ void __launch_bounds__(128) classify_wells_kernel(const float* __restrict__ pos, int B, int N, double Lx,
                                                             double Ly, double radius2,
                                                             unsigned char* __restrict__ cls,
                                                             unsigned char* __restrict__ state,
                                                             double* __restrict__ avg_x) {
    const int b = blockIdx.x;
    if (b >= B) return;
    const float2* src = reinterpret_cast<const float2*>(pos) + (size_t)b * N;
    const double cxa = Lx / 4.0, cxb = 3.0 * Lx / 4.0, cy = Ly / 2.0;
    int n_a = 0, n_b = 0;
    double sx = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float2 p = __ldg(src + i);
        const double x = (double)p.x, y = (double)p.y;
        double dy = y - cy;
        dy -= Ly * rint(dy / Ly);
        double dxa = x - cxa;
        dxa -= Lx * rint(dxa / Lx);
        double dxb = x - cxb;
        dxb -= Lx * rint(dxb / Lx);
        const bool in_a = dxa * dxa + dy * dy <= radius2;
        const bool in_b = dxb * dxb + dy * dy <= radius2;
        const unsigned char c = in_a ? 1 : (in_b ? 2 : 0);
        if (cls) cls[(size_t)b * N + i] = c;
        n_a += (c == 1);
        n_b += (c == 2);
        sx += x;
    }
    __shared__ int sh_a[4], sh_b[4];
    __shared__ double sh_x[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_a += __shfl_xor_sync(0xffffffffu, n_a, o);
        n_b += __shfl_xor_sync(0xffffffffu, n_b, o);
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        sh_a[w] = n_a;
        sh_b[w] = n_b;
        sh_x[w] = sx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int ta = 0, tb = 0;
        double tx = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            ta += sh_a[i];
            tb += sh_b[i];
            tx += sh_x[i];
        }
        if (state) state[b] = (ta == N) ? 1 : ((tb == N) ? 2 : 0);
        if (avg_x) avg_x[b] = tx / (double)N;
    }
}

// utils.py:544-552 in the reference's float32 arithmetic (samples are float32, the Python float 2*bound does not
// promote them): diff - (2b) * round(diff / (2b)) with separate multiply and subtract, sqrt(dx^2 + dy^2);
// distances equal to 0 are dropped (utils.py:550); np.histogram over the float64 edges k * dr, last bin closed.
// Every unordered pair is counted twice, like the flattened full distance matrix.
__global__ void __launch_bounds__(256) pair_histogram_kernel(const float* __restrict__ cfg, int B, int N, float box,
                                                             double dr, int nbins, double last_edge,
                                                             unsigned int* __restrict__ counts) {
    extern __shared__ unsigned int hist[];
    float2* sp = reinterpret_cast<float2*>(hist + nbins);
    const int b = blockIdx.x;
    if (b >= B) return;
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) hist[k] = 0u;
    const float2* src = reinterpret_cast<const float2*>(cfg) + (size_t)b * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) sp[i] = __ldg(src + i);
    __syncthreads();
    const long long npairs = (long long)N * (N - 1) / 2;
    for (long long t = threadIdx.x; t < npairs; t += blockDim.x) {
        // unordered pair (i, j), i < j, from the linear index t (row-major upper triangle)
        int i = (int)((2.0 * N - 1.0 - sqrt((2.0 * N - 1.0) * (2.0 * N - 1.0) - 8.0 * (double)t)) * 0.5);
        long long row0 = (long long)i * (2 * N - i - 1) / 2;
        while (row0 > t) { --i; row0 = (long long)i * (2 * N - i - 1) / 2; }
        while (row0 + (N - 1 - i) <= t) { row0 += N - 1 - i; ++i; }
        const int j = i + 1 + (int)(t - row0);
        const float2 a = sp[i], c = sp[j];
        float dx = __fsub_rn(a.x, c.x), dy = __fsub_rn(a.y, c.y);
        dx = __fsub_rn(dx, __fmul_rn(box, rintf(__fdiv_rn(dx, box))));
        dy = __fsub_rn(dy, __fmul_rn(box, rintf(__fdiv_rn(dy, box))));
        const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        if (d == 0.0f) continue;
        const double v = (double)d;
        if (v > last_edge) continue;
        int k = (int)floor(v / dr);
        if (k >= nbins) k = nbins - 1;
        while (k > 0 && v < (double)k * dr) --k;                       // edges are exactly k * dr (np.arange)
        while (k + 1 < nbins && v >= (double)(k + 1) * dr) ++k;
        atomicAdd(&hist[k], 2u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) counts[(size_t)b * nbins + k] = hist[k];
}

}  // namespace fs

extern "C" int fs_classify_wells(const float* pos, int B, int N, double Lx, double Ly, double r0, unsigned char* cls,
                                 unsigned char* state, double* avg_x, void* stream) {
    if (!pos || B < 0 || N < 1 || !(Lx > 0) || !(Ly > 0) || !(r0 > 0)) {
        fs::set_error("fs_classify_wells: invalid argument");
        return FS_ERR_INVALID;
    }
    if (B == 0) return FS_OK;
    const double radius = r0 * 1.1;
    fs::classify_wells_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(pos, B, N, Lx, Ly, radius * radius, cls, state, avg_x);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "classify_wells_kernel");
}

extern "C" int fs_pair_histogram(const float* cfg, int B, int N, double bound, double dr, int nbins,
                                 unsigned int* counts, void* stream) {
    if (!cfg || !counts || B < 0 || N < 2 || !(bound > 0) || !(dr > 0) || nbins < 1) {
        fs::set_error("fs_pair_histogram: invalid argument");
        return FS_ERR_INVALID;
    }
    if (B == 0) return FS_OK;
    const size_t smem = (size_t)nbins * sizeof(unsigned int) + (size_t)N * sizeof(float2);
    if (smem > 200 * 1024) {
        fs::set_error("fs_pair_histogram: N=%d, nbins=%d do not fit in shared memory", N, nbins);
        return FS_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        FS_CUDA(cudaFuncSetAttribute(fs::pair_histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float box = (float)(2.0 * bound);
    fs::pair_histogram_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(cfg, B, N, box, dr, nbins, (double)nbins * dr,
                                                                      counts);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "pair_histogram_kernel");
}
