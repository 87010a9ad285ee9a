// Total and single-particle pair energy for B independent configurations.
//
// fs_energy_total   <- EnergyCalculator.calculate_total_energy_virial (MCMC/energy_calculator.py:121-203)
// fs_energy_particle<- EnergyCalculator.calculate_particle_energy_virial (:48-108)
//
// Layout: a group of G threads (32..256, power of two) owns one configuration;
// its 2N floats are staged once into shared memory (stored twice, so the cyclic
// neighbour index i+k never needs a modulo).  Pairs are enumerated cyclically:
// particle i meets i+1 .. i+floor((N-1)/2), plus i+N/2 for the first half when N
// is even - every unordered pair exactly once and every thread the same trip
// count.  FP32 on the CUDA cores; cross-thread sums are finished in FP64.
#include "common.cuh"

namespace fs {

template <int G>
__global__ void __launch_bounds__(256) energy_total_kernel(const float* __restrict__ pos, int B, int N,
                                                           PotDev P, float* __restrict__ E,
                                                           float* __restrict__ W,
                                                           unsigned char* __restrict__ overlap) {
    extern __shared__ float2 smem[];
    constexpr int GROUPS = 256 / G;
    const int g = threadIdx.x / G;
    const int t = threadIdx.x % G;
    const int b = blockIdx.x * GROUPS + g;
    float2* sp = smem + (size_t)g * 2 * N;
    __shared__ double red_e[8], red_w[8];
    __shared__ float red_m[8];

    const bool live = b < B;
    if (live) {
        const float2* src = reinterpret_cast<const float2*>(pos) + (size_t)b * N;
        if ((N & 1) == 0) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            for (int i = t; i < N / 2; i += G) {
                float4 v = __ldg(s4 + i);
                sp[2 * i] = make_float2(v.x, v.y);
                sp[2 * i + 1] = make_float2(v.z, v.w);
                sp[N + 2 * i] = make_float2(v.x, v.y);
                sp[N + 2 * i + 1] = make_float2(v.z, v.w);
            }
        } else {
            for (int i = t; i < N; i += G) {
                float2 v = __ldg(src + i);
                sp[i] = v;
                sp[N + i] = v;
            }
        }
    }
    if (G > 32) __syncthreads(); else __syncwarp();

    float e = 0.f, w = 0.f, r2min = 3.0e38f;
    if (live) {
        const int half = (N - 1) / 2;
        for (int i = t; i < N; i += G) {
            const float2 pi = sp[i];
            float e0 = 0.f, w0 = 0.f, e1 = 0.f, w1 = 0.f;
            int k = 1;
            for (; k + 1 <= half; k += 2) {
                float2 a = sp[i + k], c = sp[i + k + 1];
                pair_accum(pi.x - a.x, pi.y - a.y, P, e0, w0, r2min);
                pair_accum(pi.x - c.x, pi.y - c.y, P, e1, w1, r2min);
            }
            if (k <= half) {
                float2 a = sp[i + k];
                pair_accum(pi.x - a.x, pi.y - a.y, P, e0, w0, r2min);
            }
            if ((N & 1) == 0 && i < N / 2) {
                float2 a = sp[i + N / 2];
                pair_accum(pi.x - a.x, pi.y - a.y, P, e1, w1, r2min);
            }
            e += e0 + e1 + wells(pi.x, pi.y, P);
            w += w0 + w1;
        }
    }
    // reduce over the group: shuffles inside a warp, shared memory across warps
    double de = e, dw = w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        de += __shfl_xor_sync(0xffffffffu, de, o);
        dw += __shfl_xor_sync(0xffffffffu, dw, o);
    }
    r2min = warp_min(r2min);
    if (G > 32) {
        const int wid = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) {
            red_e[wid] = de;
            red_w[wid] = dw;
            red_m[wid] = r2min;
        }
        __syncthreads();
        if (t == 0) {
            constexpr int WPG = G / 32;
            de = 0; dw = 0; r2min = 3.0e38f;
            for (int i = 0; i < WPG; ++i) {
                de += red_e[g * WPG + i];
                dw += red_w[g * WPG + i];
                r2min = fminf(r2min, red_m[g * WPG + i]);
            }
        }
    }
    if (live && t == 0) {
        const bool ov = r2min < P.rcore2;
        const float inf = __int_as_float(0x7f800000);
        E[b] = ov ? inf : (float)de;
        W[b] = ov ? inf : (float)dw;
        if (overlap) overlap[b] = ov ? 1 : 0;
    }
}

// One warp per configuration: energy of particle idx[b] against all others.
__global__ void __launch_bounds__(256) energy_particle_kernel(const float* __restrict__ pos,
                                                              const int* __restrict__ idx,
                                                              const float* __restrict__ new_xy, int B, int N,
                                                              PotDev P, float* __restrict__ e_out,
                                                              float* __restrict__ w_out,
                                                              unsigned char* __restrict__ overlap) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const float2* src = reinterpret_cast<const float2*>(pos) + (size_t)b * N;
    const int p = idx[b];
    float2 pp = new_xy ? make_float2(new_xy[2 * b], new_xy[2 * b + 1]) : __ldg(src + p);
    float e = 0.f, w = 0.f, r2min = 3.0e38f;
    for (int j = lane; j < N; j += 32) {
        if (j == p) continue;
        float2 q = __ldg(src + j);
        pair_accum(pp.x - q.x, pp.y - q.y, P, e, w, r2min);
    }
    double de = e, dw = w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        de += __shfl_xor_sync(0xffffffffu, de, o);
        dw += __shfl_xor_sync(0xffffffffu, dw, o);
    }
    r2min = warp_min(r2min);
    if (lane == 0) {
        const bool ov = r2min < P.rcore2;
        const float inf = __int_as_float(0x7f800000);
        de += (double)wells(pp.x, pp.y, P);
        e_out[b] = ov ? inf : (float)de;
        w_out[b] = ov ? inf : (float)dw;
        if (overlap) overlap[b] = ov ? 1 : 0;
    }
}

template <int G>
static int launch_total(const float* pos, int B, int N, const PotDev& P, float* E, float* W,
                        unsigned char* ov, cudaStream_t s) {
    constexpr int GROUPS = 256 / G;
    size_t smem = (size_t)GROUPS * 2 * N * sizeof(float2);
    if (smem > 48 * 1024)
        FS_CUDA(cudaFuncSetAttribute(energy_total_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (B + GROUPS - 1) / GROUPS;
    energy_total_kernel<G><<<grid, 256, smem, s>>>(pos, B, N, P, E, W, ov);
    fs::count_launch();
    return cuda_check(cudaGetLastError(), "energy_total_kernel");
}

}  // namespace fs

extern "C" int fs_energy_total(const float* pos, int B, int N, float Lx, float Ly, const fs_pot* pot,
                               float* E, float* W, unsigned char* overlap, void* stream) {
    if (!pos || !pot || !E || !W || B < 0 || N < 1 || !(Lx > 0) || !(Ly > 0)) {
        fs::set_error("fs_energy_total: invalid argument");
        return FS_ERR_INVALID;
    }
    if (N > 12288) {
        fs::set_error("fs_energy_total: N=%d exceeds the shared-memory tile (max 12288)", N);
        return FS_ERR_UNSUPPORTED;
    }
    if (B == 0) return FS_OK;
    fs::PotDev P = fs::make_pot(pot, Lx, Ly);
    cudaStream_t s = (cudaStream_t)stream;
    // group size: about half a particle per thread keeps the cyclic loops long enough
    if (N <= 48) return fs::launch_total<32>(pos, B, N, P, E, W, overlap, s);
    if (N <= 96) return fs::launch_total<64>(pos, B, N, P, E, W, overlap, s);
    if (N <= 192) return fs::launch_total<128>(pos, B, N, P, E, W, overlap, s);
    return fs::launch_total<256>(pos, B, N, P, E, W, overlap, s);
}

extern "C" int fs_energy_particle(const float* pos, const int* idx, const float* new_xy, int B, int N,
                                  float Lx, float Ly, const fs_pot* pot, float* e, float* w,
                                  unsigned char* overlap, void* stream) {
    if (!pos || !idx || !pot || !e || !w || B < 0 || N < 1 || !(Lx > 0) || !(Ly > 0)) {
        fs::set_error("fs_energy_particle: invalid argument");
        return FS_ERR_INVALID;
    }
    if (B == 0) return FS_OK;
    fs::PotDev P = fs::make_pot(pot, Lx, Ly);
    int grid = (B + 7) / 8;
    fs::energy_particle_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pos, idx, new_xy, B, N, P, e, w, overlap);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "energy_particle_kernel");
}
