// Training-target energy of Algorithm 2's reverse-KL term, batched, with its gradient.
//
// fs_target_energy <- SimpleLJ._energy + DoubleWellLJ.double_well_potential / _energy
//                     (NF/normflows/Energy/SimpleLJ.py:15-39, 61-128), called by NormalizingFlow.reverse_kld
//                     (NF/normflows/core.py:139-141) from hybrid_NF_MCMC/main_algorithm_2.py:319, 446.
//
// This is NOT the sampler's energy (energy.cu): coordinates are wrapped into the box (x - 2b round(x / 2b)) but pair
// distances take no minimum image, an extra particle sits at the origin, the pair term is the plain LJ 4 (r^-12 - r^-6)
// without cut-off above r = 0.82 and the linear soft core -80 (r - 0.82) + 30 below it, the sum is divided by the
// temperature; the double well (centres (-b/2, 0), (b/2, 0), minimum image, tanh wall) is added undivided.
// One block per configuration: thread i owns particle i, walks all partners (full sums give its gradient directly),
// energies are block-reduced.  dE/dx is produced in the same pass for autograd (d wrap / dx = 1 almost everywhere).
#include "common.cuh"

namespace fs {

struct TargetDev {
    int n, num_wells;
    float two_b, inv_two_b, inv_T;
    float V0[2], r0, k, cx[2];
};

__global__ void __launch_bounds__(128) target_energy_kernel(const float* __restrict__ x, int B, TargetDev T,
                                                            float* __restrict__ E, float* __restrict__ dEdx) {
    extern __shared__ float2 sm[];                 // wrapped coordinates, slot 0 = the particle at the origin
    __shared__ float red[4];
    const int b = blockIdx.x;
    const int n = T.n;
    const float2* xb = reinterpret_cast<const float2*>(x) + (size_t)b * n;
    if (threadIdx.x == 0) sm[0] = make_float2(0.f, 0.f);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float2 p = xb[i];
        // x - 2b round(x / 2b), torch.round = half to even (SimpleLJ.py:19-20)
        sm[i + 1] = make_float2(p.x - T.two_b * rintf(p.x / T.two_b), p.y - T.two_b * rintf(p.y / T.two_b));
    }
    __syncthreads();
    float e_acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float2 pi = sm[i + 1];
        float ei = 0.f, gx = 0.f, gy = 0.f;
        for (int j = 0; j <= n; ++j) {
            if (j == i + 1) continue;
            const float2 pj = sm[j];
            const float dx = pi.x - pj.x, dy = pi.y - pj.y;
            const float r = sqrtf(dx * dx + dy * dy);                      // torch.norm, no minimum image (:25-27)
            float e, dedr;
            if (r <= 0.82f) {                                              // soft core (:31-34)
                e = -80.0f * (r - 0.82f) + 30.0f;
                dedr = -80.0f;
            } else {
                const float inv = 1.0f / r;
                const float i2 = inv * inv, i6 = i2 * i2 * i2;
                e = 4.0f * (i6 * i6 - i6);
                dedr = 4.0f * (6.0f * i6 - 12.0f * i6 * i6) * inv;
            }
            ei += (j == 0) ? e : 0.5f * e;                                 // pairs of two real particles are met twice
            if (r > 0.f) {
                const float s = dedr / r;
                gx += s * dx;
                gy += s * dy;
            }
        }
        ei *= T.inv_T;
        gx *= T.inv_T;
        gy *= T.inv_T;
        // double well on the UNWRAPPED coordinates with minimum image (:84-112), not divided by T (:114-128)
        const float2 p = xb[i];
        for (int wi = 0; wi < T.num_wells; ++wi) {
            float dx = p.x - T.cx[wi], dy = p.y;
            dx -= T.two_b * rintf(dx / T.two_b);
            dy -= T.two_b * rintf(dy / T.two_b);
            const float r = sqrtf(dx * dx + dy * dy);
            const float t = tanhf(T.k * (r - T.r0));
            ei += T.V0[wi] * (1.0f - 0.5f * (1.0f + t));
            if (r > 0.f) {
                const float s = T.V0[wi] * (-0.5f) * T.k * (1.0f - t * t) / r;
                gx += s * dx;
                gy += s * dy;
            }
        }
        e_acc += ei;
        if (dEdx) reinterpret_cast<float2*>(dEdx)[(size_t)b * n + i] = make_float2(gx, gy);
    }
    e_acc = warp_sum(e_acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        E[b] = s;
    }
}

}  // namespace fs

extern "C" int fs_target_energy(const float* x, int B, int n_particles, double bound, double temperature,
                                const fs_pot* pot, float* E, float* dEdx, void* stream) {
    if (!x || !E || B < 0 || n_particles < 1 || !(bound > 0) || !(temperature > 0)) {
        fs::set_error("fs_target_energy: invalid argument");
        return FS_ERR_INVALID;
    }
    if (B == 0) return FS_OK;
    fs::TargetDev T;
    T.n = n_particles;
    T.num_wells = pot ? pot->num_wells : 0;
    T.two_b = (float)(2.0 * bound);
    T.inv_two_b = 1.0f / T.two_b;
    T.inv_T = (float)(1.0 / temperature);
    T.V0[0] = pot ? (float)pot->V0[0] : 0.f;
    T.V0[1] = pot ? (float)pot->V0[1] : 0.f;
    T.r0 = pot ? (float)pot->r0 : 0.f;
    T.k = pot ? (float)pot->k : 0.f;
    T.cx[0] = (float)(-bound / 2);
    T.cx[1] = (float)(bound / 2);
    const size_t smem = (size_t)(n_particles + 1) * sizeof(float2);
    if (smem > 48 * 1024) {
        fs::set_error("fs_target_energy: n_particles=%d does not fit in shared memory", n_particles);
        return FS_ERR_UNSUPPORTED;
    }
    fs::target_energy_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(x, B, T, E, dEdx);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "target_energy_kernel");
}
