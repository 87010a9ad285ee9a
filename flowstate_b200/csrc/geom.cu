// Element-wise building blocks of the energy path, exposed so the reference's
// small public helpers keep working (and can be checked against its known answers):
//   fs_apply_pbc      <- SimulationBox.apply_pbc          (MCMC/simulation_box.py:19-29)
//   fs_distances      <- SimulationBox.compute_distances  (MCMC/simulation_box.py:31-65)
//   fs_lj_pair        <- lennard_jones_energy_virial      (MCMC/potential.py:3-29)
//   fs_double_well    <- double_well_potential            (MCMC/potential.py:55-116)
#include "common.cuh"

namespace fs {

__global__ void apply_pbc_kernel(float* pos, size_t n, float Lx, float Ly) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pos[2 * i] = np_mod(pos[2 * i], Lx);
    pos[2 * i + 1] = np_mod(pos[2 * i + 1], Ly);
}

__global__ void distances_kernel(const float* __restrict__ p1, int p1_stride, const float* __restrict__ p2,
                                 size_t n, PotDev P, float* __restrict__ r) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float dx = min_image(p1[i * p1_stride] - p2[2 * i], P.Lx, P.inv_Lx);
    const float dy = min_image(p1[i * p1_stride + 1] - p2[2 * i + 1], P.Ly, P.inv_Ly);
    r[i] = sqrtf(__fmaf_rn(dy, dy, dx * dx));
}

__global__ void lj_pair_kernel(const float* __restrict__ r, size_t n, PotDev P, float* __restrict__ e,
                               float* __restrict__ w) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float rr = r[i];
    float ee = 0.f, ww = 0.f;
    if (rr * rr <= P.rc2) {
        const float inv = __frcp_rn(rr * rr);
        const float s6 = inv * inv * inv;
        ee = __fmaf_rn(4.0f * s6, s6 - 1.0f, -P.e_cut);
        ww = 48.0f * s6 * (s6 - 0.5f);
    }
    e[i] = ee;
    w[i] = ww;
}

__global__ void double_well_kernel(const float* __restrict__ pos, size_t n, PotDev P, float* __restrict__ v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    v[i] = wells(pos[2 * i], pos[2 * i + 1], P);
}

}  // namespace fs

#define FS_GRID(n) (unsigned)(((n) + 255) / 256), 256, 0, (cudaStream_t)stream

extern "C" int fs_apply_pbc(float* pos, long long n, float Lx, float Ly, void* stream) {
    if (!pos || n < 0 || !(Lx > 0) || !(Ly > 0)) { fs::set_error("fs_apply_pbc: invalid argument"); return FS_ERR_INVALID; }
    if (n == 0) return FS_OK;
    fs::apply_pbc_kernel<<<FS_GRID(n)>>>(pos, (size_t)n, Lx, Ly);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "apply_pbc_kernel");
}

extern "C" int fs_distances(const float* p1, int p1_is_single, const float* p2, long long n, float Lx, float Ly,
                            float* r, void* stream) {
    if (!p1 || !p2 || !r || n < 0 || !(Lx > 0) || !(Ly > 0)) { fs::set_error("fs_distances: invalid argument"); return FS_ERR_INVALID; }
    if (n == 0) return FS_OK;
    fs_pot pot = {0, {0, 0}, 1, 1, 2.5f, 0.5f};
    fs::PotDev P = fs::make_pot(&pot, Lx, Ly);
    fs::distances_kernel<<<FS_GRID(n)>>>(p1, p1_is_single ? 0 : 2, p2, (size_t)n, P, r);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "distances_kernel");
}

extern "C" int fs_lj_pair(const float* r, long long n, const fs_pot* pot, float* e, float* w, void* stream) {
    if (!r || !pot || !e || !w || n < 0) { fs::set_error("fs_lj_pair: invalid argument"); return FS_ERR_INVALID; }
    if (n == 0) return FS_OK;
    fs::PotDev P = fs::make_pot(pot, 1.f, 1.f);
    fs::lj_pair_kernel<<<FS_GRID(n)>>>(r, (size_t)n, P, e, w);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "lj_pair_kernel");
}

extern "C" int fs_double_well(const float* pos, long long n, float Lx, float Ly, const fs_pot* pot, float* v,
                              void* stream) {
    if (!pos || !pot || !v || n < 0 || !(Lx > 0) || !(Ly > 0)) { fs::set_error("fs_double_well: invalid argument"); return FS_ERR_INVALID; }
    if (n == 0) return FS_OK;
    fs::PotDev P = fs::make_pot(pot, Lx, Ly);
    fs::double_well_kernel<<<FS_GRID(n)>>>(pos, (size_t)n, P, v);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "double_well_kernel");
}
