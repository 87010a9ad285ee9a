// Rational-quadratic spline for TRAINING (Algorithm 2's per-cycle updates, Algorithm 1's pre-training): value + log-det
// in one kernel, and their hand-written reverse-mode derivative in another.
//
// fs_spline_train_fwd / fs_spline_train_bwd <- unconstrained_rational_quadratic_spline + rational_quadratic_spline in
//     the density direction (NF/normflows/utils/splines.py:16-88, 91-161, 203-222) as autograd differentiates them when
//     the drivers call NormalizingFlow.forward_kld (NF/normflows/core.py:88-108; main_algorithm_1.py:306,
//     main_algorithm_2.py:447).  Eager autograd runs ~80 element-wise kernels per spline and direction; these two replace
//     them for the conditional spline of the transformed half AND the unconditional spline of the identity half.
//
// One thread per (row, coordinate).  Parameters of a coordinate: P = 3 nb + 1 consecutive floats
// [nb widths | nb heights | nb + 1 derivatives] (coupling.py:166, 335-342), addressed as
// theta + row * row_stride + coord * P  (row_stride = 0: parameters shared by all rows - the unconditional spline,
// coupling.py:208-238).  Widths / heights logits are multiplied by `scale` (1 / sqrt(hidden) for the conditional spline,
// coupling.py:340-342; 1 for the unconditional one).  Fork quirks as in the inference kernels: last knot
// + 1e-6 in the bin search, independent boundary derivatives, inputs outside [-bound, bound] pass through with log-det 0.
// The backward kernel recomputes the forward from (x, theta) - nothing is saved - and writes dL/dx and dL/dtheta
// [rows, N, P] (the caller sums over rows when the parameters are shared).
#include "spline_train.cuh"

namespace fs {

template <bool BWD>
__global__ void __launch_bounds__(128) spline_train_kernel(const float* __restrict__ x, const float* __restrict__ theta,
                                                           long long row_stride, int rows, int N, int nb, float bound,
                                                           float scale, float* __restrict__ y, float* __restrict__ ld,
                                                           const float* __restrict__ gy, const float* __restrict__ gld,
                                                           float* __restrict__ gx, float* __restrict__ gtheta) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)rows * N) return;
    const int r = (int)(e / N), j = (int)(e % N);
    const int P = 3 * nb + 1;
    const float* u = theta + (size_t)r * row_stride + (size_t)j * P;
    float yv = 0.f, lv = 0.f, gxv = 0.f;
    spline_point<BWD>(x[e], u, nb, bound, scale, yv, lv, BWD ? gy[e] : 0.f, BWD ? gld[e] : 0.f, gxv,
                      BWD ? gtheta + (size_t)e * P : nullptr);
    if (BWD) {
        gx[e] = gxv;
    } else {
        y[e] = yv;
        ld[e] = lv;
    }
}

}  // namespace fs

static int spline_train_check(const void* x, const void* theta, int rows, int N, int nb, double bound, const char* who) {
    if (!x || !theta || rows < 0 || N < 1 || nb < 1 || !(bound > 0)) {
        fs::set_error("%s: invalid argument", who);
        return FS_ERR_INVALID;
    }
    return FS_OK;
}

extern "C" int fs_spline_train_fwd(const float* x, const float* theta, long long theta_row_stride, int rows, int N, int nb,
                                   double bound, double scale, float* y, float* logdet, void* stream) {
    if (int r = spline_train_check(x, theta, rows, N, nb, bound, "fs_spline_train_fwd")) return r;
    if (!y || !logdet) { fs::set_error("fs_spline_train_fwd: null output"); return FS_ERR_INVALID; }
    if (rows == 0) return FS_OK;
    const long long n = (long long)rows * N;
    fs::spline_train_kernel<false><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        x, theta, theta_row_stride, rows, N, nb, (float)bound, (float)scale, y, logdet, nullptr, nullptr, nullptr, nullptr);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "spline_train_kernel<fwd>");
}

extern "C" int fs_spline_train_bwd(const float* x, const float* theta, long long theta_row_stride, int rows, int N, int nb,
                                   double bound, double scale, const float* grad_y, const float* grad_logdet,
                                   float* grad_x, float* grad_theta, void* stream) {
    if (int r = spline_train_check(x, theta, rows, N, nb, bound, "fs_spline_train_bwd")) return r;
    if (!grad_y || !grad_logdet || !grad_x || !grad_theta) {
        fs::set_error("fs_spline_train_bwd: null gradient buffer");
        return FS_ERR_INVALID;
    }
    if (rows == 0) return FS_OK;
    const long long n = (long long)rows * N;
    fs::spline_train_kernel<true><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        x, theta, theta_row_stride, rows, N, nb, (float)bound, (float)scale, nullptr, nullptr, grad_y, grad_logdet, grad_x,
        grad_theta);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "spline_train_kernel<bwd>");
}
