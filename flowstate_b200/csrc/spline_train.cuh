// Device side of the training spline (see spline_train.cu for the reference citations).
#pragma once
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

// One (row, coordinate) element: forward value + log-det, or (BWD) the reverse-mode derivative written to gx and
// gt[0 .. P) (P = 3 nb + 1), recomputing the forward from (xv, u).  Shared by spline_train_kernel and the whole-flow
// chain kernels of train.cu.
template <bool BWD>
__device__ __forceinline__ void spline_point(float xv, const float* __restrict__ u, int nb, float bound, float scale,
                                             float& y_out, float& ld_out, float gyv, float gl, float& gx_out,
                                             float* __restrict__ gt) {
    const int P = 3 * nb + 1;
    if (!(xv >= -bound && xv <= bound)) {                          // tails: identity, log-det 0 (utils/splines.py:24-39)
        if (BWD) {
            gx_out = gyv;
            for (int i = 0; i < P; ++i) gt[i] = 0.f;
        } else {
            y_out = xv;
            ld_out = 0.f;
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
        y_out = yv;
        ld_out = lv;
        return;
    }
    // ---- reverse mode through the rational-quadratic formula ----
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
    gx_out = x_b;
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
