// Packed flow (device-resident, inference form) shared by the FP32 and the
// tensor-core conditioner paths.
#pragma once
#include <vector>

#include "common.cuh"

struct fs_flow {
    int K, N, D, H, n_blocks, nb, P;   // P = 3 nb + 1 parameters per transformed coordinate
    double bound;
    float bound_f, pf_scale, base_logc, inv_sqrt_h;
    int* idf;   // [N] identity feature indices
    int* trf;   // [N] transformed feature indices
    // per layer, FP32 row-major, BatchNorm folded (see flow.cu: pack_layer)
    struct Layer {
        float* init_w;   // [H, 2N]
        float* init_b;   // [H]
        float* bn0_s;    // [n_blocks, H]   scale  of the first BN of each block
        float* bn0_o;    // [n_blocks, H]   offset
        float* w0;       // [n_blocks, H, H]  first linear with the second BN folded in
        float* b0;       // [n_blocks, H]
        float* w1;       // [n_blocks, H, H]
        float* b1;       // [n_blocks, H]
        float* final_w;  // [P N, H]  parameter-major rows (k*N + j)
        float* final_b;  // [P N]
        float* u_x;      // [nb+1, N] knots of the unconditional spline (x), knot-major
        float* u_y;      // [nb+1, N] knots (y)
        float* u_d;      // [nb+1, N] derivatives (already 1e-3 + softplus)
    };
    std::vector<Layer> layers;
    std::vector<void*> allocs;
    void* tc;   // tensor-core pack (flow_tc.cu), or nullptr
    int* tc_err;   // device error word written by the tensor kernel's watchdog
    int sm_count;
    int lp_mode = 0;                   // fs_flow_set_layer_parallel: 0 auto, 1 prefer (the caller's passes run alone), 2 never
    void* repack_tab = nullptr;        // fs_flow_update: device table of per-layer pointers (repack.cu)
    double* repack_scratch = nullptr;  // ... and its float64 scratch
};

namespace fs {
static constexpr float kMinW = 1e-3f, kMinH = 1e-3f, kMinD = 1e-3f;   // utils/splines.py:6-8

// ---------------------------------------------------------------------------
// spline device code
// ---------------------------------------------------------------------------
__device__ __forceinline__ float softplus_t(float x) {   // F.softplus, threshold 20
    return x > 20.0f ? x : log1pf(expf(x));
}

// Rational-quadratic bin evaluation (utils/splines.py:163-222) given the selected bin.
__device__ __forceinline__ void rq_eval(float x, float xk, float wk, float yk, float hk, float dk, float dk1,
                                        bool inverse, float& y, float& ld) {
    const float sk = hk / wk;
    const float t = dk + dk1 - 2.0f * sk;
    if (inverse) {
        const float dy = x - yk;
        const float a = dy * t + hk * (sk - dk);
        const float b = hk * dk - dy * t;
        const float c = -sk * dy;
        const float disc = fabsf(b * b - 4.0f * a * c);
        const float root = (2.0f * c) / (-b - sqrtf(disc));
        y = root * wk + xk;
        const float tt = root * (1.0f - root);
        const float den = sk + t * tt;
        const float omr = 1.0f - root;
        const float num = (sk * sk) * (dk1 * (root * root) + 2.0f * sk * tt + dk * (omr * omr));
        ld = -(logf(num) - 2.0f * logf(den));
    } else {
        const float th = (x - xk) / wk;
        const float tt = th * (1.0f - th);
        const float num = hk * (sk * (th * th) + dk * tt);
        const float den = sk + t * tt;
        y = yk + num / den;
        const float omt = 1.0f - th;
        const float dnum = (sk * sk) * (dk1 * (th * th) + 2.0f * sk * tt + dk * (omt * omt));
        ld = logf(dnum) - 2.0f * logf(den);
    }
}

// Fast-math variants for the tensor-core epilogue: raw MUFU exp2 / log2 / reciprocal / square root (~1e-6 relative, far
// inside the tensor path's error budget; flush-to-zero forms, so none of the range-scaling code the C intrinsics add
// around them; the arguments here are O(1e-9 .. 1e9)).  The FP32 path keeps the accurate ones above.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float softplus_fast(float x) {
    return fmaxf(x, 0.0f) + 0.6931471805599453f * lg2_approx(1.0f + ex2_approx(-1.4426950408889634f * fabsf(x)));
}
__device__ __forceinline__ void rq_eval_fast(float x, float xk, float wk, float yk, float hk, float dk, float dk1,
                                             bool inverse, float& y, float& ld) {
    const float sk = hk * rcp_approx(wk);
    const float t = dk + dk1 - 2.0f * sk;
    if (inverse) {
        const float dy = x - yk;
        const float a = dy * t + hk * (sk - dk);
        const float b = hk * dk - dy * t;
        const float c = -sk * dy;
        const float disc = fabsf(b * b - 4.0f * a * c);
        const float root = (2.0f * c) * rcp_approx(-b - sqrt_approx(disc));
        y = root * wk + xk;
        const float tt = root * (1.0f - root);
        const float den = sk + t * tt;
        const float omr = 1.0f - root;
        const float num = (sk * sk) * (dk1 * (root * root) + 2.0f * sk * tt + dk * (omr * omr));
        const float rden = rcp_approx(den);
        ld = -0.6931471805599453f * lg2_approx(num * rden * rden);
    } else {
        const float th = (x - xk) * rcp_approx(wk);
        const float tt = th * (1.0f - th);
        const float num = hk * (sk * (th * th) + dk * tt);
        const float den = sk + t * tt;
        const float rden = rcp_approx(den);
        y = yk + num * rden;
        const float omt = 1.0f - th;
        const float dnum = (sk * sk) * (dk1 * (th * th) + 2.0f * sk * tt + dk * (omt * omt));
        ld = 0.6931471805599453f * lg2_approx(dnum * rden * rden);
    }
}

// ---- tensor-core pack (flow_tc.cu fills it, repack.cu refreshes it on the device) ----
static constexpr int TC_KB = 32;          // K elements per TF32 weight tile (one 128-byte swizzle atom per row)
struct TcLayer {
    float* wstream;    // all weight tiles of the layer in consumption order
    float* bn0_s;      // [n_blocks, H]
    float* bn0_o;      // [n_blocks, H]  BatchNorm offset with the running bias folded in
    float* b0;         // [n_blocks, H]
    float* b_final;    // [n_chunks * 128]  b_f + W_f c
    float* psets;      // [n_blocks + 1][3][H]: set 0 = {-, s_0, o'_0}; set b+1 = {b0'_b, s_{b+1}, o'_{b+1}}
    void* wfused;      // fused-spline final layer, FP16 operands: per coordinate (H/64)/KPS stages of KPS tiles of
                       // [chn rows x 64 k] halves (half the bytes and twice the MMA rate of TF32, same 11-bit significand)
    float* b_fused;    // [N][chn] (+ 32 floats of padding), folded with the FP16-rounded weights
};

struct TcPack {
    int H, NH, Kp0, n_pieces, n_chunks, nstage;
    int chn;           // columns of a fused final-layer chunk (0: fused path unavailable)
    size_t tiles_per_layer, tiles_before_final;
    size_t smem_bytes;
    int* xcols;        // device [4][N]: input / output columns of the transformed coordinates, density then sampling
    std::vector<TcLayer> layers;
    TcLayer* layers_dev = nullptr;   // the same table on the device (layer-parallel launches index it by CTA)
    bool lp_ok = false;              // identity set closed under the roll: all conditioner inputs known up front
};

// Feature matrix of the tensor path: element (row b, feature k) of a [rows][K0] matrix, stored as 128-row tiles of
// quads: [b / 128][k / 4][b % 128][k % 4].  The conditioner's epilogue threads own one row each (TMEM lane = row), so a
// warp's float4 load of quad k/4 touches 32 consecutive rows = 512 contiguous bytes; the row-major layout made every
// one of those loads 32 separate sectors (117 k clk of LSU wavefronts per tile at K0 = 512, measured in-kernel).
__host__ __device__ inline size_t a0_tiled(int b, int k, int K0) {
    return ((((size_t)(b >> 7) * (size_t)((K0 + 3) >> 2) + (size_t)(k >> 2)) * 128 + (size_t)(b & 127)) << 2) + (size_t)(k & 3);
}

// final-layer rows/bias permuted to parameter-major order (row k*N + j), see flow.cu
void permute_final(const fs_layer_params* p, int N, int P, int H, std::vector<float>& w, std::vector<float>& b);
// tensor-core path (flow_tc.cu)
int tc_pack(fs_flow* f, const fs_flow_desc* d);
void tc_free(fs_flow* f);
size_t tc_workspace_bytes(const fs_flow* f, int B);
int tc_conditioner(fs_flow* f, int layer, const float* A0, bool tiled, int rows, float* theta, void* ws,
                   size_t ws_bytes, int* nan_flag, cudaStream_t s);
bool tc_has_fused(const fs_flow* f);
int tc_conditioner_spline(fs_flow* f, int layer, const float* A0, bool tiled, int rows, int direction, const float* xin,
                          float* xout, float* logdet, int* nan_flag, cudaStream_t s);
bool tc_layer_parallel_ok(const fs_flow* f);
size_t tc_lp_flag_ints(const fs_flow* f, int rows);
int tc_conditioner_spline_all(fs_flow* f, int direction, int rows, const float* A0, size_t a0_stride, float* buf0,
                              float* buf1, float* ldp, size_t ld_stride, int* flags, int* nan_flag, cudaStream_t s);
}  // namespace fs
