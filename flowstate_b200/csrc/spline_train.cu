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
#include "common.cuh"

namespace fs {

static constexpr float kTMin = 1e-3f;   // min bin width / height / derivative (utils/splines.py:6-8)

struct SplineBin {
    int k;              // selected bin
    float c0, c1;       // cumulative sizes at knots k and k + 1 (before the affine map to [-bound, bound])
    float p0, p1;       // sums of the softmax probabilities below knots k and k + 1
    float lo, hi;       // knots k and k + 1
};

// softmax over `nb` logits u[i] * scale -> normaliser; returns max and 1 / sum
__device__ __forceinline__ void softmax_norm(const float* __restrict__ u, int nb, float scale, float& m, float& rz) {
    m = -3.0e38f;
    for (int i = 0; i < nb; ++i) m = fmaxf(m, u[i] * scale);
    float z = 0.f;
    for (int i = 0; i < nb; ++i) z += expf(u[i] * scale - m);
    rz = 1.0f / z;
}

// knots of one axis around bin k (k given), utils/splines.py:117-127: cumsum of MIN + (1 - MIN nb) softmax, affine to
// [-bound, bound], end knots forced
__device__ __forceinline__ void axis_knots(const float* __restrict__ u, int nb, float scale, float m, float rz, float bound,
                                           int k, SplineBin& b) {
    const float c = 1.0f - kTMin * nb;
    float cum = 0.f, ps = 0.f;
    for (int i = 0; i < k; ++i) {
        const float p = expf(u[i] * scale - m) * rz;
        cum += kTMin + c * p;
        ps += p;
    }
    const float pk = expf(u[k] * scale - m) * rz;
    b.c0 = cum;
    b.c1 = cum + kTMin + c * pk;
    b.p0 = ps;
    b.p1 = ps + pk;
    b.k = k;
    b.lo = (k == 0) ? -bound : 2.0f * bound * b.c0 - bound;
    b.hi = (k == nb - 1) ? bound : 2.0f * bound * b.c1 - bound;
}

// bin search on the width axis: bin = #(x >= knot_j) - 1 with the last knot + 1e-6 (utils/splines.py:11-13)
__device__ __forceinline__ int search_bin(const float* __restrict__ u, int nb, float scale, float m, float rz, float bound,
                                          float x) {
    const float c = 1.0f - kTMin * nb;
    float cum = 0.f;
    int k = 0;
    for (int i = 0; i < nb - 1; ++i) {
        cum += kTMin + c * expf(u[i] * scale - m) * rz;
        const float knot = 2.0f * bound * cum - bound;            // knot i + 1 (interior)
        if (x >= knot) k = i + 1;
    }
    return k;
}

__device__ __forceinline__ float softplus_acc(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

struct RqFwd {
    float wk, hk, s, th, omt, tt, t, den, numA, dn, d0, d1;
};

__device__ __forceinline__ void rq_forward(float x, float x0, float x1, float y0, float y1, float d0, float d1, RqFwd& f,
                                           float& y, float& ld) {
    f.wk = x1 - x0;
    f.hk = y1 - y0;
    f.s = f.hk / f.wk;
    f.th = (x - x0) / f.wk;
    f.omt = 1.0f - f.th;
    f.tt = f.th * f.omt;
    f.t = d0 + d1 - 2.0f * f.s;
    f.den = f.s + f.t * f.tt;
    f.numA = f.s * f.th * f.th + d0 * f.tt;
    f.dn = d1 * f.th * f.th + 2.0f * f.s * f.tt + d0 * f.omt * f.omt;
    f.d0 = d0;
    f.d1 = d1;
    y = y0 + f.hk * f.numA / f.den;
    ld = logf(f.s * f.s * f.dn) - 2.0f * logf(f.den);             // utils/splines.py:214-222
}

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
    const float xv = x[e];
    float* gt = BWD ? gtheta + (size_t)e * P : nullptr;
    if (!(xv >= -bound && xv <= bound)) {                          // tails: identity, log-det 0 (utils/splines.py:24-39)
        if (BWD) {
            gx[e] = gy[e];
            for (int i = 0; i < P; ++i) gt[i] = 0.f;
        } else {
            y[e] = xv;
            ld[e] = 0.f;
        }
        return;
    }
    float mw, rzw, mh, rzh;
    softmax_norm(u, nb, scale, mw, rzw);
    softmax_norm(u + nb, nb, scale, mh, rzh);
    const int k = search_bin(u, nb, scale, mw, rzw, bound, xv);
    SplineBin bw, bh;
    axis_knots(u, nb, scale, mw, rzw, bound, k, bw);
    axis_knots(u + nb, nb, scale, mh, rzh, bound, k, bh);
    const float ud0 = u[2 * nb + k], ud1 = u[2 * nb + k + 1];
    const float d0 = kTMin + softplus_acc(ud0), d1 = kTMin + softplus_acc(ud1);
    RqFwd f;
    float yv, lv;
    rq_forward(xv, bw.lo, bw.hi, bh.lo, bh.hi, d0, d1, f, yv, lv);
    if (!BWD) {
        y[e] = yv;
        ld[e] = lv;
        return;
    }
    // ---- reverse mode through the rational-quadratic formula ----
    const float gyv = gy[e], gl = gld[e];
    const float q = f.numA / f.den;
    float hk_b = gyv * q;
    const float q_b = gyv * f.hk;
    const float numA_b = q_b / f.den;
    float den_b = -q_b * q / f.den - 2.0f * gl / f.den;
    float s_b = 2.0f * gl / f.s;
    const float dn_b = gl / f.dn;
    float d1_b = dn_b * f.th * f.th;
    s_b += dn_b * 2.0f * f.tt;
    float tt_b = dn_b * 2.0f * f.s;
    float d0_b = dn_b * f.omt * f.omt;
    float th_b = dn_b * 2.0f * f.d1 * f.th;
    float omt_b = dn_b * 2.0f * f.d0 * f.omt;
    s_b += numA_b * f.th * f.th;
    th_b += numA_b * 2.0f * f.s * f.th;
    d0_b += numA_b * f.tt;
    tt_b += numA_b * f.d0;
    s_b += den_b;
    const float t_b = den_b * f.tt;
    tt_b += den_b * f.t;
    d0_b += t_b;
    d1_b += t_b;
    s_b -= 2.0f * t_b;
    th_b += tt_b * f.omt;
    omt_b += tt_b * f.th;
    th_b -= omt_b;
    const float x_b = th_b / f.wk;
    float x0_b = -x_b;
    float wk_b = -th_b * f.th / f.wk;
    hk_b += s_b / f.wk;
    wk_b += -s_b * f.s / f.wk;
    const float x1_b = wk_b;
    x0_b -= wk_b;
    const float y1_b = hk_b;
    const float y0_b = gyv - hk_b;
    gx[e] = x_b;
    // ---- knots -> cumulative sizes -> softmax logits (end knots are constants: no gradient through them) ----
    const float c = 1.0f - kTMin * nb, two_b = 2.0f * bound;
    {
        const float G0 = (k == 0) ? 0.f : two_b * x0_b, G1 = (k == nb - 1) ? 0.f : two_b * x1_b;
        const float dot = G0 * bw.p0 + G1 * bw.p1;
        for (int i = 0; i < nb; ++i) {
            const float p = expf(u[i] * scale - mw) * rzw;
            gt[i] = scale * c * p * ((i < k ? G0 : 0.f) + (i < k + 1 ? G1 : 0.f) - dot);
        }
    }
    {
        const float G0 = (k == 0) ? 0.f : two_b * y0_b, G1 = (k == nb - 1) ? 0.f : two_b * y1_b;
        const float dot = G0 * bh.p0 + G1 * bh.p1;
        for (int i = 0; i < nb; ++i) {
            const float p = expf(u[nb + i] * scale - mh) * rzh;
            gt[nb + i] = scale * c * p * ((i < k ? G0 : 0.f) + (i < k + 1 ? G1 : 0.f) - dot);
        }
    }
    for (int i = 0; i <= nb; ++i) gt[2 * nb + i] = 0.f;
    gt[2 * nb + k] = d0_b / (1.0f + expf(-ud0));                  // d softplus = sigmoid
    gt[2 * nb + k + 1] = d1_b / (1.0f + expf(-ud1));
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
