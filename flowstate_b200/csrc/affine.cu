// Affine (RealNVP) coupling transforms and periodic shifts / wraps.
//
// fs_affine_coupling <- MaskedAffineFlow.forward / inverse (NF/normflows/flows/affine/coupling.py:163-229) and the
//                       element-wise half of AffineCoupling.forward / inverse (:99-160): the scale and shift parameters
//                       come from the caller's conditioner network (plain library GEMMs); the kernel applies
//                       z' = z exp(s) + t (or its inverse, or the sigmoid scale maps), the mask, the "non-finite
//                       parameter -> NaN" rule of MaskedAffineFlow and the per-row log-determinant in one pass.
// fs_periodic_shift  <- PeriodicShift.forward / inverse and PeriodicWrap.inverse (NF/normflows/flows/periodic.py:6-73):
//                       torch.remainder(z + shift + bound, 2 bound) - bound on the selected columns.
// One warp per row, lanes over the features, shuffle reduction of the log-determinant.
#include "common.cuh"

namespace fs {

struct AffineArgs {
    const float* z; long long z_rs;
    const float* mask;
    const float* scale; long long s_rs; int s_es;
    const float* shift; long long t_rs; int t_es;
    float* out; long long o_rs;
    float* logdet;
    int rows, n, scale_map, inverse, nan_rule;
};

__global__ void __launch_bounds__(256) affine_coupling_kernel(AffineArgs A) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= A.rows) return;
    const float qnan = __int_as_float(0x7fc00000);
    float ld = 0.f;
    for (int j = lane; j < A.n; j += 32) {
        const float z = A.z[(size_t)row * A.z_rs + j];
        const float b = A.mask ? A.mask[j] : 0.f;                       // b = 1: the feature is left unchanged
        float s = A.scale ? A.scale[(size_t)row * A.s_rs + (size_t)j * A.s_es] : 0.f;
        float t = A.shift ? A.shift[(size_t)row * A.t_rs + (size_t)j * A.t_es] : 0.f;
        if (A.nan_rule) {                                              // coupling.py:199-202, 211-214
            if (!isfinite(s)) s = qnan;
            if (!isfinite(t)) t = qnan;
        }
        float zt, l;
        if (A.scale_map == 0) {                                        // exp (RealNVP)
            zt = A.inverse ? (z - t) * expf(-s) : z * expf(s) + t;
            l = A.inverse ? -s : s;
        } else {                                                       // sigmoid (Glow) / sigmoid_inv, coupling.py:133-152
            const float sg = 1.0f / (1.0f + expf(-(s + 2.0f)));
            const bool mul = (A.scale_map == 2) != (A.inverse != 0);   // sigmoid_inv forward and sigmoid inverse multiply
            zt = A.inverse ? (mul ? (z - t) * sg : (z - t) / sg) : (mul ? z * sg + t : z / sg + t);
            l = mul ? logf(sg) : -logf(sg);
        }
        if (A.mask) {                                                  // f(z) = b z + (1 - b) (...), log-det sum((1 - b) s)
            zt = b * z + (1.0f - b) * zt;
            l = (1.0f - b) * l;
        }
        A.out[(size_t)row * A.o_rs + j] = zt;
        ld += l;
    }
    ld = warp_sum(ld);
    if (lane == 0 && A.logdet) A.logdet[row] = ld;
}

__global__ void periodic_shift_kernel(const float* __restrict__ z, float* __restrict__ out, long long total, int D,
                                      const int* __restrict__ col_slot, const float* __restrict__ bound,
                                      const float* __restrict__ shift) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int slot = col_slot[(int)(e % D)];                           // -1: column not selected
    float v = z[e];
    if (slot >= 0) {
        const float b = bound[slot];
        v = np_mod(v + shift[slot] + b, 2.0f * b) - b;                 // torch.remainder = floor-mod
    }
    out[e] = v;
}

}  // namespace fs

extern "C" int fs_affine_coupling(const float* z, long long z_row_stride, const float* mask, const float* scale,
                                  long long scale_row_stride, int scale_elem_stride, const float* shift,
                                  long long shift_row_stride, int shift_elem_stride, int rows, int n, int scale_map,
                                  int inverse, int nan_rule, float* out, long long out_row_stride, float* logdet,
                                  void* stream) {
    if (!z || !out || rows < 0 || n < 1 || scale_map < 0 || scale_map > 2) {
        fs::set_error("fs_affine_coupling: invalid argument");
        return FS_ERR_INVALID;
    }
    if (rows == 0) return FS_OK;
    fs::AffineArgs A{z, z_row_stride, mask, scale, scale_row_stride, scale_elem_stride, shift, shift_row_stride,
                     shift_elem_stride, out, out_row_stride, logdet, rows, n, scale_map, inverse, nan_rule};
    fs::affine_coupling_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(A);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "affine_coupling_kernel");
}

extern "C" int fs_periodic_shift(const float* z, int rows, int D, const int* col_slot, const float* bound,
                                 const float* shift, float* out, void* stream) {
    if (!z || !out || !col_slot || !bound || !shift || rows < 0 || D < 1) {
        fs::set_error("fs_periodic_shift: invalid argument");
        return FS_ERR_INVALID;
    }
    if (rows == 0) return FS_OK;
    const long long total = (long long)rows * D;
    fs::periodic_shift_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z, out, total, D,
                                                                                                 col_slot, bound, shift);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "periodic_shift_kernel");
}
