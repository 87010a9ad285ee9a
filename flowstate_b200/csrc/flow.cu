// Circular coupled rational-quadratic-spline flow, inference (eval-mode) path.
//
// Reference (paths relative to <ref>/NF/normflows):
//   flows/neural_spline/wrapper.py:98-275   layer wrapper (forward = prqct.inverse, inverse = prqct.forward)
//   flows/neural_spline/coupling.py:71-134  split / conditioner / splines / scatter / roll by D/2
//   flows/neural_spline/coupling.py:156-170,335-368  (B,N,3nb+1) parameters, widths/heights / sqrt(H)
//   flows/neural_spline/coupling.py:176-265 unconditional spline of the identity half
//   utils/splines.py:16-222                 unconstrained / rational-quadratic spline
//   utils/nn.py:120-137                     cat[cos(s x), sin(s x)]
//   nets/resnet.py:7-104                    residual conditioner (BatchNorm eps 1e-3, eval mode)
//   Energy/Uniform.py:50-74                 base log-probability
//   core.py:28-86,178-214                   layer loops
//
// This file holds the pack (BatchNorm folding, unconditional knot tables), the
// CUDA-core FP32 conditioner (FS_PREC_FP32: reference arithmetic, used as the
// precision yardstick for the tensor path), the feature / identity-half kernels
// (one warp per row, coordinates one per lane) and the stand-alone conditional
// spline kernel of the paths that materialise theta; the tensor path with the
// spline fused into the conditioner's epilogue lives in flow_tc.cu.
#include <math.h>
#include <string.h>

#include <vector>

#include "flow.cuh"

namespace fs {

struct FlowDev {
    int N, D, H, nb, P;
    float bound, pf_scale, inv_sqrt_h;
    const int* idf;
    const int* trf;
};

// v <- x - shift (float64 subtraction stored float32, MCMC/monte_carlo.py:251-258) or x + shift
__global__ void shift_kernel(const float* __restrict__ x, float* __restrict__ v, size_t n, double shift) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (float)((double)x[i] + shift);
}

// ---------------------------------------------------------------------------
// Lane-per-coordinate spline kernels (the ones the flow passes launch).
// A warp walks rows with a grid stride; a chunk of 32 coordinates sits one per lane.  theta is
// parameter-major ([3nb+1][N] per row, see pack_layer), so the 2nb softmax logits of a chunk are 2nb
// row segments of 128 bytes: they are copied to shared memory with cp.async (sm[k*32 + lane], bank-
// conflict free for the per-lane walks), double buffered so the next chunk's copy is in flight while
// this one is evaluated.  Each lane then runs softmax -> inclusive prefix sums (in place) -> binary
// search for its bin -> rational-quadratic evaluation: no shuffles, exps only for the two softmaxes,
// softplus only for the two derivatives of the selected bin (fetched from global memory; their rows
// are prefetched into L2 together with the copy).
// ---------------------------------------------------------------------------
#define FS_SPLINE_WARPS 4   // prep_forward_v2: rows per block
#define FS_SPLINE_MAXW 4    // spline_kernel: warps (= chunks of a row in flight) per block

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Unconditional spline of the identity half from the packed knot tables.  Tables are knot-major ([nb+1][N]: entry k of
// coordinate j at k*N + j), so the 32 lanes of a warp read consecutive addresses.
// U coordinates of one lane at once: the binary searches (same trip count for every coordinate) and the
// table loads of the U coordinates interleave, which hides the L1/L2 latency of the dependent knot loads.
template <int U, bool FAST = false>
__device__ __forceinline__ void rqs_table_multi(const float (&x)[U], const bool (&valid)[U], const int (&j)[U],
                                                const float* __restrict__ ux, const float* __restrict__ uy,
                                                const float* __restrict__ ud, int N, int nb, float bound, bool inverse,
                                                float (&y)[U], float (&ld)[U]) {
    const float* ks = inverse ? uy : ux;
    int lo[U], hi[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { lo[u] = 0; hi[u] = nb; }
    for (int span = nb; span > 1; span = (span + 1) >> 1) {           // ceil-halving covers every hi - lo
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (valid[u] && hi[u] - lo[u] > 1) {
                const int mid = (lo[u] + hi[u]) >> 1;
                if (x[u] >= __ldg(ks + (size_t)mid * N + j[u])) lo[u] = mid; else hi[u] = mid;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        y[u] = x[u];
        ld[u] = 0.f;
        if (valid[u] && x[u] >= -bound && x[u] <= bound) {
            const int sel = lo[u];
            const float* cx = ux + j[u];
            const float* cy = uy + j[u];
            const float* cd = ud + j[u];
            const float xk = __ldg(cx + (size_t)sel * N), xk1 = __ldg(cx + (size_t)(sel + 1) * N);
            const float yk = __ldg(cy + (size_t)sel * N), yk1 = __ldg(cy + (size_t)(sel + 1) * N);
            if (FAST)
                rq_eval_fast(x[u], xk, xk1 - xk, yk, yk1 - yk, __ldg(cd + (size_t)sel * N),
                             __ldg(cd + (size_t)(sel + 1) * N), inverse, y[u], ld[u]);
            else
                rq_eval(x[u], xk, xk1 - xk, yk, yk1 - yk, __ldg(cd + (size_t)sel * N),
                        __ldg(cd + (size_t)(sel + 1) * N), inverse, y[u], ld[u]);
        }
    }
}

// Conditional spline of ONE coordinate.  p[k * 32], k < 2nb: the staged width and height logits of this
// lane's coordinate (overwritten by the inclusive prefix sums of the softmax numerators); gp: this
// coordinate's column of theta in global memory (stride N) for the two derivatives of the selected bin.
// Knot k of an axis is 2b (g S[k-1] + min k) - b with g = (1 - min nb) / S[nb-1]  (utils/splines.py:84-100
// restated on unnormalised sums); knot 0 = -b, knot nb = b.
__device__ __forceinline__ void rqs_cond_lane(float x, float* p, const float* __restrict__ gp, int N, int nb,
                                              float bound, float inv_sqrt_h, bool inverse, float& y, float& ld) {
    if (!(x >= -bound && x <= bound)) {   // utils/splines.py:24,38-39
        y = x;
        ld = 0.0f;
        return;
    }
    const float c2 = inv_sqrt_h * 1.4426950408889634f;      // softmax(u / sqrt(H)) via exp2
    float* q = p + nb * 32;
    float mw = -3.0e38f, mh = -3.0e38f;
#pragma unroll 8
    for (int k = 0; k < nb; ++k) {
        mw = fmaxf(mw, p[k * 32]);
        mh = fmaxf(mh, q[k * 32]);
    }
    const float ow = -mw * c2, oh = -mh * c2;
    float sw = 0.f, sh = 0.f;
#pragma unroll 8
    for (int k = 0; k < nb; ++k) {
        sw += ex2_fast(__fmaf_rn(p[k * 32], c2, ow));
        sh += ex2_fast(__fmaf_rn(q[k * 32], c2, oh));
        p[k * 32] = sw;
        q[k * 32] = sh;
    }
    const float gw = (1.0f - kMinW * (float)nb) / sw, gh = (1.0f - kMinH * (float)nb) / sh;
    const float two_b = 2.0f * bound;
    const float* srch = inverse ? q : p;
    const float gs = inverse ? gh : gw, ms = inverse ? kMinH : kMinW;
    int lo = 0, hi = nb;                 // last knot <= x == #(x >= knots) - 1 (utils/splines.py:11-13)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        const float knot = __fmaf_rn(two_b, __fmaf_rn(gs, srch[(mid - 1) * 32], ms * (float)mid), -bound);
        if (x >= knot) lo = mid; else hi = mid;
    }
    const int sel = lo;
    const float w0 = sel ? p[(sel - 1) * 32] : 0.f, h0 = sel ? q[(sel - 1) * 32] : 0.f;
    const float xl = __fmaf_rn(two_b, __fmaf_rn(gw, w0, kMinW * (float)sel), -bound);
    const float yl = __fmaf_rn(two_b, __fmaf_rn(gh, h0, kMinH * (float)sel), -bound);
    const float xr = (sel == nb - 1) ? bound : __fmaf_rn(two_b, __fmaf_rn(gw, p[sel * 32], kMinW * (float)(sel + 1)), -bound);
    const float yr = (sel == nb - 1) ? bound : __fmaf_rn(two_b, __fmaf_rn(gh, q[sel * 32], kMinH * (float)(sel + 1)), -bound);
    const float dk = kMinD + softplus_t(__ldg(gp + (size_t)(2 * nb + sel) * N));
    const float dk1 = kMinD + softplus_t(__ldg(gp + (size_t)(2 * nb + sel + 1) * N));
    rq_eval(x, xl, xr - xl, yl, yr - yl, dk, dk1, inverse, y, ld);
}

// Start the copy of the 2nb logit rows of coordinates [j0, j0+nc) of one row of theta into sm[k*32 + lane]
// (one cp.async group) and pull the nb+1 derivative rows of the same chunk into L2.
__device__ __forceinline__ void stage_issue(const float* __restrict__ th, float* sm, int j0, int nc, int N, int nb,
                                            int lane) {
    const int nb2 = 2 * nb;
    if (nc == 32 && (N & 3) == 0) {       // 16-byte copies: 8 lanes per 128-byte row segment, 4 rows per pass
        const int seg = (lane & 7) * 4;
        const float* src = th + j0 + seg;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm + seg);
        for (int k = lane >> 3; k < nb2; k += 4)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + k * 128), "l"(src + (size_t)k * N) : "memory");
    } else if (lane < nc) {
        const float* src = th + j0 + lane;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm + lane);
        for (int k = 0; k < nb2; ++k)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + k * 128), "l"(src + (size_t)k * N) : "memory");
    }
    for (int k = nb2 + lane; k <= 3 * nb; k += 32)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(th + (size_t)k * N + j0));
}

// DENSITY (step 2 of the density direction): conditional spline on the transformed half, scatter + roll by
// D/2 (coupling.py:86-102).  !DENSITY (step 2 of sampling): inverse conditional spline on the transformed
// half (coupling.py:126-135).  The identity half is written by prep_inverse_v2 / prep_forward_v2.
// A block owns a row at a time (grid stride) and its warps take the row's 32-coordinate chunks, so the
// block reads each parameter row of theta as one contiguous run; the per-row log-det is summed in a
// fixed order (deterministic).  The operand of the next chunk is loaded together with its theta copy.
template <bool DENSITY>
__global__ void __launch_bounds__(32 * FS_SPLINE_MAXW) spline_kernel(
    const float* __restrict__ v, const float* __restrict__ theta, float* __restrict__ out,
    float* __restrict__ logdet, int rows, FlowDev F, int* nan_flag) {
    extern __shared__ __align__(16) float sp_smem[];
    __shared__ float red[FS_SPLINE_MAXW];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int nb = F.nb, h = F.D / 2, nchunk = (F.N + 31) >> 5;
    const int bufsz = 32 * 2 * nb;
    float* sm = sp_smem + (size_t)wib * 2 * bufsz;
    const size_t trow = (size_t)F.N * F.P;
    int buf = 0;
    int ft_n = 0;
    float xt_n = 0.f;
    if ((int)blockIdx.x < rows && wib < nchunk) {
        stage_issue(theta + (size_t)blockIdx.x * trow, sm, wib * 32, min(32, F.N - wib * 32), F.N, nb, lane);
        if (wib * 32 + lane < F.N) {
            ft_n = F.trf[wib * 32 + lane];
            xt_n = v[(size_t)blockIdx.x * F.D + (DENSITY ? ft_n : (ft_n + h) % F.D)];
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    bool bad = false;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const float* th = theta + (size_t)r * trow;
        float acc = 0.f;
        for (int c = wib; c < nchunk; c += W) {
            const int ft = ft_n;
            const float xt = xt_n;
            int rn = r, cn = c + W;
            if (cn >= nchunk) {
                cn = wib;
                rn = r + gridDim.x;
            }
            if (rn < rows) {
                stage_issue(theta + (size_t)rn * trow, sm + (buf ^ 1) * bufsz, cn * 32, min(32, F.N - cn * 32), F.N, nb,
                            lane);
                if (cn * 32 + lane < F.N) {
                    ft_n = F.trf[cn * 32 + lane];
                    xt_n = v[(size_t)rn * F.D + (DENSITY ? ft_n : (ft_n + h) % F.D)];
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            const int j = c * 32 + lane;
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            if (j < F.N) {
                float y, ld;
                rqs_cond_lane(xt, sm + buf * bufsz + lane, th + j, F.N, nb, F.bound, F.inv_sqrt_h, !DENSITY, y, ld);
                bad = bad || (y != y) || (ld != ld);
                acc += ld;
                out[(size_t)r * F.D + (DENSITY ? (ft + h) % F.D : ft)] = y;
            }
            __syncwarp();
            buf ^= 1;
        }
        if (logdet) {
            acc = warp_sum_f(acc);
            if (W == 1) {
                if (lane == 0) logdet[r] += acc;
            } else {
                if (lane == 0) red[wib] = acc;
                __syncthreads();
                if (threadIdx.x == 0) {
                    float t = 0.f;
                    for (int i = 0; i < W; ++i) t += red[i];
                    logdet[r] += t;
                }
                __syncthreads();
            }
        }
    }
    if (bad && nan_flag) atomicOr(nan_flag, 1);
}

// sampling direction, step 1: roll, inverse unconditional spline on the identity half, periodic
// features of the NEW identity values (coupling.py:113-124)
// FAST (tensor path): MUFU-based sin / cos / log / reciprocal, ~1e-6, far inside that path's error budget
template <int U, bool FAST>
__global__ void __launch_bounds__(32 * FS_SPLINE_WARPS) prep_forward_v2(
    const float* __restrict__ v, float* __restrict__ out, float* __restrict__ A0, float* __restrict__ logdet, int rows,
    FlowDev F, const float* __restrict__ ux, const float* __restrict__ uy, const float* __restrict__ ud,
    int* nan_flag) {
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * FS_SPLINE_WARPS + wib;
    if (b >= rows) return;
    const float* vr = v + (size_t)b * F.D;
    const int h = F.D / 2, nb = F.nb;
    float acc = 0.f;
    bool bad = false;
    for (int j0 = lane; j0 < F.N; j0 += 32 * U) {
        float x[U], y[U], ld[U];
        int j[U], fi[U];
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            j[u] = j0 + 32 * u;
            valid[u] = j[u] < F.N;
            fi[u] = valid[u] ? F.idf[j[u]] : 0;
            x[u] = valid[u] ? vr[(fi[u] + h) % F.D] : 0.f;
        }
        rqs_table_multi<U, FAST>(x, valid, j, ux, uy, ud, F.N, nb, F.bound, true, y, ld);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!valid[u]) continue;
            out[(size_t)b * F.D + fi[u]] = y[u];
            float sn, cs;
            if (FAST) __sincosf(F.pf_scale * y[u], &sn, &cs); else sincosf(F.pf_scale * y[u], &sn, &cs);
            if (FAST) {           // tensor path: row-tiled feature layout (a0_tiled), coalesced for its lane = row loads
                A0[a0_tiled(b, j[u], 2 * F.N)] = cs;
                A0[a0_tiled(b, F.N + j[u], 2 * F.N)] = sn;
            } else {
                A0[(size_t)b * 2 * F.N + j[u]] = cs;
                A0[(size_t)b * 2 * F.N + F.N + j[u]] = sn;
            }
            acc += ld[u];
            bad = bad || (y[u] != y[u]) || (ld[u] != ld[u]);
        }
    }
    acc = warp_sum_f(acc);
    if (lane == 0 && logdet) logdet[b] += acc;
    if (bad && nan_flag) atomicOr(nan_flag, 1);
}

// density direction, step 1: periodic features of the identity half (utils/nn.py:125-127) and the
// unconditional spline on the identity half, scattered + rolled by D/2 (coupling.py:86-102).  One warp per
// row; the knot tables stay L1-resident here (no shared-memory carve-out).
template <int U, bool FAST>
__global__ void __launch_bounds__(32 * FS_SPLINE_WARPS) prep_inverse_v2(
    const float* __restrict__ v, float* __restrict__ out, float* __restrict__ A0, float* __restrict__ logdet, int rows,
    FlowDev F, const float* __restrict__ ux, const float* __restrict__ uy, const float* __restrict__ ud,
    int* nan_flag) {
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * FS_SPLINE_WARPS + wib;
    if (b >= rows) return;
    const float* vr = v + (size_t)b * F.D;
    const int h = F.D / 2, nb = F.nb;
    float acc = 0.f;
    bool bad = false;
    for (int j0 = lane; j0 < F.N; j0 += 32 * U) {
        float x[U], y[U], ld[U];
        int j[U], fi[U];
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            j[u] = j0 + 32 * u;
            valid[u] = j[u] < F.N;
            fi[u] = valid[u] ? F.idf[j[u]] : 0;
            x[u] = valid[u] ? vr[fi[u]] : 0.f;
        }
        rqs_table_multi<U, FAST>(x, valid, j, ux, uy, ud, F.N, nb, F.bound, false, y, ld);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!valid[u]) continue;
            float sn, cs;
            if (FAST) __sincosf(F.pf_scale * x[u], &sn, &cs); else sincosf(F.pf_scale * x[u], &sn, &cs);
            if (FAST) {           // tensor path: row-tiled feature layout (a0_tiled), coalesced for its lane = row loads
                A0[a0_tiled(b, j[u], 2 * F.N)] = cs;
                A0[a0_tiled(b, F.N + j[u], 2 * F.N)] = sn;
            } else {
                A0[(size_t)b * 2 * F.N + j[u]] = cs;
                A0[(size_t)b * 2 * F.N + F.N + j[u]] = sn;
            }
            out[(size_t)b * F.D + (fi[u] + h) % F.D] = y[u];
            acc += ld[u];
            bad = bad || (y[u] != y[u]) || (ld[u] != ld[u]);
        }
    }
    acc = warp_sum_f(acc);
    if (lane == 0 && logdet) logdet[b] += acc;
    if (bad && nan_flag) atomicOr(nan_flag, 1);
}


// Both prep steps with the knot tables staged in shared memory (the kernel the flow passes launch when the tables of
// a 32-coordinate slice fit: nb <= 62).  prep_*_v2 gather the knots straight from global memory: after the first
// probe of the bin search every lane of a warp asks for a different table row, ~20 sectors per request, and the
// kernel sits on the LSU's wavefront rate (ncu r02b: 8 wavefronts per element, 40 us per launch at 8192 x 256).
// Here a block of 8 warps owns 32 rows and walks the identity coordinates in slices of 32 (lane = coordinate, every
// lane four rows at once: the U independent chains of rqs_table_multi).  The slice's tables [x | y | d][nb+1][32]
// are copied by cp.async into one of two shared buffers while the previous slice is computed; knot k of the lane's
// coordinate is word 32 k + lane: bank = lane whatever k, no conflicts.  Sums run over the same elements in the same
// order as prep_*_v2 (lane-strided coordinates, then the warp tree): results are bit-identical.
// DENSITY: true = prep_inverse (coupling.py:86-102), false = prep_forward (coupling.py:113-124).
#ifndef FS_PREP3_WARPS
#define FS_PREP3_WARPS 8
#endif
#ifndef FS_PREP3_ROWS
#define FS_PREP3_ROWS 4
#endif
__device__ __forceinline__ void cp_async4(unsigned dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
template <bool FAST, bool DENSITY>
__global__ void __launch_bounds__(32 * FS_PREP3_WARPS) prep_v3(
    const float* __restrict__ v, float* __restrict__ out, float* __restrict__ A0, float* __restrict__ logdet, int rows,
    FlowDev F, const float* __restrict__ ux, const float* __restrict__ uy, const float* __restrict__ ud,
    int* nan_flag) {
    extern __shared__ float prep_tab[];
    constexpr int U = FS_PREP3_ROWS;
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = F.nb, nk = nb + 1, N = F.N, D = F.D, h = D / 2;
    const int tsz = 3 * nk * 32;                                   // floats per buffer
    const int n_slices = (N + 31) >> 5;
    const int b0 = (blockIdx.x * FS_PREP3_WARPS + wib) * U;
    const unsigned tab_s = (unsigned)__cvta_generic_to_shared(prep_tab) + 4u * (unsigned)lane;
    auto stage = [&](int sl, int buf) {
        if (32 * sl + lane < N) {
#pragma unroll
            for (int t = 0; t < 3; ++t) {                           // knot row k of the slice: 32 consecutive floats
                const float* src = (t == 0 ? ux : (t == 1 ? uy : ud)) + 32 * sl + lane + wib * N;
                unsigned dst = tab_s + 4u * (unsigned)(buf * tsz + (t * nk + wib) * 32);
                for (int k = wib; k < nk; k += FS_PREP3_WARPS, src += FS_PREP3_WARPS * N, dst += 128u * FS_PREP3_WARPS)
                    cp_async4(dst, src);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // branch-free bin search: the largest index <= nb - 1 whose knot is <= x, bit by bit from the top
    // (same answer as the halving search of rqs_table_multi on monotone knots; NaN -> 0 in both)
    int top = 1;
    while (2 * top <= nb - 1) top *= 2;
    stage(0, 0);
    float acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.f;
    bool bad = false;
    // the layer input of a slice is a dependent load (index table -> row): both run one slice ahead of the arithmetic
    auto column = [&](int sl) {
        const int j = 32 * sl + lane;
        return (sl < n_slices && j < N) ? __ldg(F.idf + j) : -1;
    };
    const float* vrow = v + (size_t)b0 * D;
    float* orow = out + (size_t)b0 * D;
    // feature (row b, column k) of the tensor path sits at tile_base(b) + 512 (k / 4) + k % 4 (a0_tiled, flow.cuh)
    float* arow = FAST ? A0 + ((size_t)(b0 >> 7) * (size_t)((2 * N + 3) >> 2) * 512 + (size_t)(b0 & 127) * 4)
                       : A0 + (size_t)b0 * 2 * N;                  // the 4 rows of a warp never straddle a 128-row tile
    auto rolled = [&](int fi) { const int c = fi + h; return c >= D ? c - D : c; };
    auto fetch = [&](int fi, float (&xn)[U]) {
        const int c = DENSITY ? fi : rolled(fi);
#pragma unroll
        for (int u = 0; u < U; ++u) xn[u] = (fi >= 0 && b0 + u < rows) ? vrow[u * D + c] : 0.f;
    };
    int fi_cur = column(0), fi_nxt = column(1);
    float x_nxt[U];
    fetch(fi_cur, x_nxt);
    for (int sl = 0; sl < n_slices; ++sl) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");       // slice sl has landed (this thread's copies)
        __syncthreads();                                           // ... everybody's; and slice sl - 1 is no longer read
        if (sl + 1 < n_slices) stage(sl + 1, (sl + 1) & 1);        // overlaps the arithmetic below
        const float* tx = prep_tab + (sl & 1) * tsz + lane;
        const float* ty = tx + nk * 32;
        const float* td = ty + nk * 32;
        const float* ks = DENSITY ? tx : ty;                      // density: spline forward (search x knots); sampling: inverse
        const int j = 32 * sl + lane;
        const int fi = fi_cur;
        float x[U], y[U], ld[U];
        bool valid[U];
        int lo[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            valid[u] = fi >= 0 && (b0 + u < rows);
            x[u] = x_nxt[u];
            lo[u] = 0;
        }
        fi_cur = fi_nxt;
        fetch(fi_cur, x_nxt);                                      // slice sl + 1
        fi_nxt = column(sl + 2);
        for (int step = top; step >= 1; step >>= 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int cand = lo[u] + step;                     // <= 2 top - 1 <= nb: inside the nb + 1 rows staged
                const bool ok = cand <= nb - 1 && x[u] >= ks[cand * 32];
                lo[u] = ok ? cand : lo[u];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {                              // evaluated for every lane, selected afterwards
            const int sel = lo[u];
            const float xk = tx[sel * 32], xk1 = tx[(sel + 1) * 32];
            const float yk = ty[sel * 32], yk1 = ty[(sel + 1) * 32];
            float yy, ll;
            if (FAST)
                rq_eval_fast(x[u], xk, xk1 - xk, yk, yk1 - yk, td[sel * 32], td[(sel + 1) * 32], !DENSITY, yy, ll);
            else
                rq_eval(x[u], xk, xk1 - xk, yk, yk1 - yk, td[sel * 32], td[(sel + 1) * 32], !DENSITY, yy, ll);
            const bool inside = x[u] >= -F.bound && x[u] <= F.bound;   // utils/splines.py:24, 38-39
            y[u] = inside ? yy : x[u];
            ld[u] = inside ? ll : 0.f;
        }
        const int oc = DENSITY ? rolled(fi) : fi;
        const int kc = FAST ? ((j >> 2) * 512 + (j & 3)) : j;                         // cos feature j, sin feature N + j
        const int ksn = FAST ? (((N + j) >> 2) * 512 + ((N + j) & 3)) : N + j;
        const int ustep = FAST ? 4 : 2 * N;                                           // next row
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!valid[u]) continue;
            const float arg = F.pf_scale * (DENSITY ? x[u] : y[u]);   // features of the identity VALUES the conditioner sees
            float sn, cs;
            if (FAST) __sincosf(arg, &sn, &cs); else sincosf(arg, &sn, &cs);
            arow[u * ustep + kc] = cs;
            arow[u * ustep + ksn] = sn;
            orow[u * D + oc] = y[u];
            acc[u] += ld[u];
            bad = bad || (y[u] != y[u]) || (ld[u] != ld[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float a = warp_sum_f(acc[u]);
        if (lane == 0 && logdet && b0 + u < rows) logdet[b0 + u] += a;
    }
    if (bad && nan_flag) atomicOr(nan_flag, 1);
}
static inline size_t prep3_smem(const fs_flow* f) { return (size_t)2 * 3 * (f->nb + 1) * 32 * 4; }
// measured on B200 (scripts/kernel_times.py, us per launch, v3 / v2): N = 256: 44 / 52 (16384 rows, density), 32 / 36
// (8192 rows, sampling); N = 32: 8.1 / 7.2 -> one slice has nothing to overlap the staging with.  FS_PREP_V2=1 / FS_PREP_V3=1
// force either (the tests compare them bit for bit).
static inline bool use_prep3(const fs_flow* f) {
    if (prep3_smem(f) > 48 * 1024 || getenv("FS_PREP_V2")) return false;
    return f->N > 32 || getenv("FS_PREP_V3");
}
static inline unsigned prep3_grid(int rows) {
    return (unsigned)((rows + FS_PREP3_WARPS * FS_PREP3_ROWS - 1) / (FS_PREP3_WARPS * FS_PREP3_ROWS));
}


// warps per block of spline_kernel: one per 32-coordinate chunk of a row, at most FS_SPLINE_MAXW
static int spline_warps(const fs_flow* f) {
    const int nchunk = (f->N + 31) / 32;
    return nchunk < FS_SPLINE_MAXW ? nchunk : FS_SPLINE_MAXW;
}

static size_t spline_smem_bytes(const fs_flow* f) {
    return (size_t)spline_warps(f) * 2 * 32 * 2 * f->nb * sizeof(float);   // two buffers per warp
}

// one wave of resident blocks; blocks stride over rows
static unsigned spline_grid(const fs_flow* f, int rows) {
    const size_t per_block = spline_smem_bytes(f) + 1024;
    int resident = (int)((227 * 1024) / per_block);
    const int by_warps = 48 / spline_warps(f);
    resident = resident > by_warps ? by_warps : resident;
    resident = resident > 32 ? 32 : (resident < 1 ? 1 : resident);
    const long long cap = (long long)f->sm_count * resident;
    return (unsigned)(rows < cap ? rows : cap);
}

// out <- z (+ shift); logq <- logdet + UniformParticle.log_prob(z)  (Energy/Uniform.py:50-74)
__global__ void __launch_bounds__(256) finish_kernel(const float* __restrict__ v, float* __restrict__ out,
                                                     const float* __restrict__ logdet,
                                                     float* __restrict__ logdet_out, float* __restrict__ logq,
                                                     int rows, int D, float bound, float base_logc,
                                                     double shift, const float* __restrict__ ld_part = nullptr,
                                                     int n_part = 0, size_t part_stride = 0) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= rows) return;
    bool inb = true;
    for (int i = lane; i < D; i += 32) {
        const float z = v[(size_t)b * D + i];
        inb = inb && (z >= -bound) && (z <= bound);
        if (out) out[(size_t)b * D + i] = (shift != 0.0) ? (float)((double)z + shift) : z;
    }
    inb = __all_sync(0xffffffffu, inb);
    if (lane == 0) {
        float ld = logdet[b];
        for (int i = 0; i < n_part; ++i) ld += ld_part[(size_t)i * part_stride + b];   // layer-parallel pass: fixed order
        if (logdet_out) logdet_out[b] = ld;
        if (logq) logq[b] = inb ? ld + base_logc : -__int_as_float(0x7f800000);
    }
}

// ---------------------------------------------------------------------------
// FP32 conditioner GEMM: C[M,Nout] = pro(A)[M,K] . W[Nout,K]^T + bias (+ relu | + residual)
// 64x64x16 tiles, 256 threads, 4x4 micro-tiles.
// ---------------------------------------------------------------------------
enum { PRO_NONE = 0, PRO_BNRELU = 1 };
enum { EPI_NONE = 0, EPI_RELU = 1, EPI_RESIDUAL = 2 };

template <int PRO, int EPI>
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ A, const float* __restrict__ Wt,
                                                     const float* __restrict__ bias,
                                                     const float* __restrict__ ps, const float* __restrict__ po,
                                                     float* C, int M, int Nout, int K) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = tid % 16, ty = tid / 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid / 4;          // 0..63 row inside the tile
    const int lk = (tid % 4) * 4;    // 0,4,8,12
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lk + q;
            float a = 0.f, w = 0.f;
            if (k < K) {
                if (m0 + lr < M) {
                    a = A[(size_t)(m0 + lr) * K + k];
                    if (PRO == PRO_BNRELU) a = fmaxf(__fmaf_rn(a, ps[k], po[k]), 0.f);
                }
                if (n0 + lr < Nout) w = __ldg(Wt + (size_t)(n0 + lr) * K + k);
            }
            As[lk + q][lr] = a;
            Ws[lk + q][lr] = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = Ws[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= Nout) continue;
            float v = acc[i][j] + bias[n];
            if (EPI == EPI_RELU) v = fmaxf(v, 0.f);
            if (EPI == EPI_RESIDUAL) v += C[(size_t)m * Nout + n];
            C[(size_t)m * Nout + n] = v;
        }
    }
}

template <int PRO, int EPI>
static int linear(const float* A, const float* W, const float* bias, const float* ps, const float* po, float* C,
                  int M, int Nout, int K, cudaStream_t s) {
    dim3 grid((Nout + 63) / 64, (M + 63) / 64);
    linear_kernel<PRO, EPI><<<grid, 256, 0, s>>>(A, W, bias, ps, po, C, M, Nout, K);
    fs::count_launch();
    return cuda_check(cudaGetLastError(), "linear_kernel");
}

// resnet.py:92-104 in eval mode with BatchNorm folded at pack time
static int conditioner_fp32(const fs_flow* f, int li, const float* A0, int rows, float* hbuf, float* tbuf,
                            float* theta, cudaStream_t s) {
    const fs_flow::Layer& L = f->layers[li];
    const int H = f->H;
    int r = linear<PRO_NONE, EPI_NONE>(A0, L.init_w, L.init_b, nullptr, nullptr, hbuf, rows, H, 2 * f->N, s);
    if (r) return r;
    for (int b = 0; b < f->n_blocks; ++b) {
        r = linear<PRO_BNRELU, EPI_RELU>(hbuf, L.w0 + (size_t)b * H * H, L.b0 + (size_t)b * H,
                                         L.bn0_s + (size_t)b * H, L.bn0_o + (size_t)b * H, tbuf, rows, H, H, s);
        if (r) return r;
        r = linear<PRO_NONE, EPI_RESIDUAL>(tbuf, L.w1 + (size_t)b * H * H, L.b1 + (size_t)b * H, nullptr, nullptr,
                                           hbuf, rows, H, H, s);
        if (r) return r;
    }
    return linear<PRO_NONE, EPI_NONE>(hbuf, L.final_w, L.final_b, nullptr, nullptr, theta, rows, f->N * f->P, H, s);
}

// ---------------------------------------------------------------------------
// pack
// ---------------------------------------------------------------------------
template <typename T>
static int upload(fs_flow* f, const std::vector<T>& h, T** out) {
    void* d = nullptr;
    FS_CUDA(cudaMalloc(&d, h.size() * sizeof(T) + 16));
    f->allocs.push_back(d);
    FS_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (T*)d;
    return FS_OK;
}

static void host_knots(const float* un, int nb, double bound, double minsz, std::vector<double>& k) {
    // utils/splines.py:117-127 in float64
    double m = un[0];
    for (int i = 1; i < nb; ++i) m = un[i] > m ? un[i] : m;
    double sum = 0;
    std::vector<double> e(nb);
    for (int i = 0; i < nb; ++i) { e[i] = exp((double)un[i] - m); sum += e[i]; }
    k.assign(nb + 1, 0.0);
    double c = 0;
    for (int i = 0; i < nb; ++i) {
        c += minsz + (1 - minsz * nb) * (e[i] / sum);
        k[i + 1] = 2 * bound * c - bound;
    }
    k[0] = -bound;
    k[nb] = bound;
}

void permute_final(const fs_layer_params* p, int N, int P, int H, std::vector<float>& w, std::vector<float>& b) {
    w.resize((size_t)N * P * H);
    b.resize((size_t)N * P);
    for (int j = 0; j < N; ++j)
        for (int k = 0; k < P; ++k) {
            const size_t src = (size_t)j * P + k, dst = (size_t)k * N + j;
            memcpy(&w[dst * H], p->final_w + src * H, sizeof(float) * H);
            b[dst] = p->final_b[src];
        }
}

static int pack_layer(fs_flow* f, const fs_flow_desc* d, const fs_layer_params* p, fs_flow::Layer* L) {
    const int H = f->H, N = f->N, nb = f->nb, nB = f->n_blocks, P = f->P;
    std::vector<float> v;
    v.assign(p->init_w, p->init_w + (size_t)H * 2 * N);
    if (int r = upload(f, v, &L->init_w)) return r;
    v.assign(p->init_b, p->init_b + H);
    if (int r = upload(f, v, &L->init_b)) return r;
    std::vector<float> s0((size_t)nB * H), o0((size_t)nB * H), w0((size_t)nB * H * H), b0((size_t)nB * H);
    for (int b = 0; b < nB; ++b) {
        for (int j = 0; j < 2; ++j) {
            const size_t o = ((size_t)b * 2 + j) * H;
            for (int c = 0; c < H; ++c) {
                // nn.BatchNorm1d(eps=1e-3) in eval mode: y = (x - mean) / sqrt(var + eps) * w + b
                const double sc = (double)p->bn_w[o + c] / sqrt((double)p->bn_var[o + c] + (double)d->bn_eps);
                const double of = (double)p->bn_b[o + c] - (double)p->bn_mean[o + c] * sc;
                if (j == 0) {
                    s0[(size_t)b * H + c] = (float)sc;
                    o0[(size_t)b * H + c] = (float)of;
                } else {
                    // second BN follows linear 0 directly: fold into its rows
                    const float* wr = p->lin_w + (((size_t)b * 2 + 0) * H + c) * H;
                    for (int k = 0; k < H; ++k) w0[((size_t)b * H + c) * H + k] = (float)(sc * (double)wr[k]);
                    b0[(size_t)b * H + c] = (float)(sc * (double)p->lin_b[((size_t)b * 2 + 0) * H + c] + of);
                }
            }
        }
    }
    if (int r = upload(f, s0, &L->bn0_s)) return r;
    if (int r = upload(f, o0, &L->bn0_o)) return r;
    if (int r = upload(f, w0, &L->w0)) return r;
    if (int r = upload(f, b0, &L->b0)) return r;
    std::vector<float> w1((size_t)nB * H * H), b1((size_t)nB * H);
    for (int b = 0; b < nB; ++b) {
        memcpy(&w1[(size_t)b * H * H], p->lin_w + ((size_t)b * 2 + 1) * H * H, sizeof(float) * H * H);
        memcpy(&b1[(size_t)b * H], p->lin_b + ((size_t)b * 2 + 1) * H, sizeof(float) * H);
    }
    if (int r = upload(f, w1, &L->w1)) return r;
    if (int r = upload(f, b1, &L->b1)) return r;
    // The final layer's rows are permuted from the reference's coordinate-major order (row j*P + k,
    // coupling.py:166) to parameter-major (row k*N + j): theta comes out as [P][N] per sample, and the 32
    // coordinates a spline warp works on are contiguous for every parameter index.
    std::vector<float> fb;
    permute_final(p, N, P, H, v, fb);
    if (int r = upload(f, v, &L->final_w)) return r;
    if (int r = upload(f, fb, &L->final_b)) return r;
    std::vector<float> ux((size_t)N * (nb + 1)), uy((size_t)N * (nb + 1)), ud((size_t)N * (nb + 1));
    std::vector<double> k;
    for (int j = 0; j < N; ++j) {
        host_knots(p->un_w + (size_t)j * nb, nb, f->bound, 1e-3, k);
        for (int i = 0; i <= nb; ++i) ux[(size_t)i * N + j] = (float)k[i];
        host_knots(p->un_h + (size_t)j * nb, nb, f->bound, 1e-3, k);
        for (int i = 0; i <= nb; ++i) uy[(size_t)i * N + j] = (float)k[i];
        for (int i = 0; i <= nb; ++i) {
            const double x = p->un_d[(size_t)j * (nb + 1) + i];
            ud[(size_t)i * N + j] = (float)(1e-3 + (x > 20 ? x : log1p(exp(x))));
        }
    }
    if (int r = upload(f, ux, &L->u_x)) return r;
    if (int r = upload(f, uy, &L->u_y)) return r;
    return upload(f, ud, &L->u_d);
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// rows processed per pass so that the parameter buffer theta stays below 4 GiB (180 GB of HBM)
// tensor path with the spline applied in the conditioner's epilogue: theta is never materialised
// (FS_NO_FUSE=1 keeps theta + the spline kernel; read per call so tests can switch it)
static bool fused_path(const fs_flow* f, int precision) {
    return precision == FS_PREC_TF32 && tc_has_fused(f) && !getenv("FS_NO_FUSE");
}

static int chunk_rows(const fs_flow* f, int B) {
    size_t per_row = (size_t)f->N * f->P * sizeof(float);
    size_t rows = (4096ull << 20) / per_row;
    if (rows < 128) rows = 128;
    rows = rows / 128 * 128;
    return (int)(rows < (size_t)B ? rows : (size_t)B);
}

struct Workspace {
    float *v0, *v1, *A0, *h, *t, *theta, *ld;
    void* tc;
    size_t tc_bytes;
    // layer-parallel pass (layer_parallel()): K feature matrices a0_stride floats apart, K log-det partials of Bc
    // floats, the chunk counters of the launch
    size_t a0_stride;
    float* ldp;
    int* flags;
};

// One launch for the conditioners of all K layers of a pass (tc_conditioner_spline_all): the fused tensor path on a
// flow whose identity coordinates stay identity coordinates under the roll (every even-N configuration, SURVEY.md
// A.4-Q2).  FS_NO_LP=1 keeps one launch per layer (read per call so tests can compare the two).
static bool layer_parallel(const fs_flow* f, int precision) {
    return fused_path(f, precision) && tc_layer_parallel_ok(f);
}
// ... used for a pass of `rows` rows when it pays.  A step's CTA runs its GEMM stack at once but starts its spline
// chunks only when the previous step of its row tile has finished, so a CTA may idle on its SM for up to one
// final-layer time.  That is free while every CTA of the launch is resident anyway (K x tiles <= SMs); beyond that
// the wait disappears when the previous step's CTA was dispatched early enough: tiles / SMs of a kernel time earlier,
// against the fraction of a kernel the trunk (GEMM0 + residual blocks) takes.  Measured (profiles/r02_layer_parallel.md):
// alg1_n32 passes 1.5 x faster, alg1_n256 (final layer 61 % of the flops, 64 / 128 tiles) no gain -> one launch per layer.
// FS_LP_MAX_TILES overrides the tile limit (development).
static bool layer_parallel_rows(const fs_flow* f, int precision, int rows) {
    if (!layer_parallel(f, precision)) return false;
    const int tiles = (rows + 127) / 128;
    if (const char* e = getenv("FS_LP_MAX_TILES")) return tiles <= atoi(e);
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    if (f->lp_mode == 2) return false;
    if (tiles > 96) return false;
    if (f->K * tiles <= sms) return true;
    // "prefer" (fs_flow_set_layer_parallel): the caller runs its passes alone on the GPU, so thread blocks that wait
    // for their predecessor step take nothing from anybody; with two whole steps resident the trunks of step s + 1
    // overlap the spline chain of step s (Alg-2 flow, 23 layers x 32 / 64 tiles: passes 11 % faster).  Not the default:
    // next to other streams (the hybrid round samples its next proposals beside the sweep) waiting blocks hold SMs
    // the other kernels need, and the measured gain was nil.
    if (f->lp_mode == 1 && 2 * tiles <= sms) return true;
    const double trunk = 2.0 * (2.0 * f->N) * f->H + 4.0 * f->n_blocks * (double)f->H * f->H;
    const double fin = 2.0 * f->H * (double)f->N * (3.0 * f->nb + 1.0);
    return (double)tiles / sms + trunk / (trunk + fin) >= 1.0;
}

static size_t carve(const fs_flow* f, int B, int precision, void* base, Workspace* w) {
    const int Bc = chunk_rows(f, B);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return base ? (void*)((char*)base + o) : nullptr;
    };
    float* v0 = (float*)take((size_t)Bc * f->D * 4);
    float* v1 = (float*)take((size_t)Bc * f->D * 4);
    // features: row-major [Bc][2N] on the FP32 path, row-tiled (a0_tiled: 128-row tiles, quads of 4 features) on the tensor path
    const bool lp = layer_parallel(f, precision);
    const size_t a0_floats = (size_t)((Bc + 127) / 128 * 128) * ((2 * f->N + 3) / 4 * 4);
    float* A0 = (float*)take(a0_floats * 4 * (lp ? f->K : 1));
    const bool fp32 = precision != FS_PREC_TF32;                   // hidden activations live in TMEM on the tensor path
    float* h = (float*)take(fp32 ? (size_t)Bc * f->H * 4 : 0);
    float* t = (float*)take(fp32 ? (size_t)Bc * f->H * 4 : 0);
    float* th = (float*)take(fused_path(f, precision) ? 0 : (size_t)Bc * f->N * f->P * 4);
    float* ld = (float*)take((size_t)Bc * 4);
    size_t tcb = precision == FS_PREC_TF32 ? tc_workspace_bytes(f, Bc) : 0;
    void* tc = take(tcb);
    float* ldp = (float*)take(lp ? (size_t)f->K * Bc * 4 : 0);
    int* flags = (int*)take(lp ? tc_lp_flag_ints(f, Bc) * 4 : 0);
    if (w) *w = Workspace{v0, v1, A0, h, t, th, ld, tc, tcb, a0_floats, ldp, flags};
    return off;
}

static FlowDev flow_dev(const fs_flow* f) {
    FlowDev F;
    F.N = f->N; F.D = f->D; F.H = f->H; F.nb = f->nb; F.P = f->P;
    F.bound = f->bound_f; F.pf_scale = f->pf_scale; F.inv_sqrt_h = f->inv_sqrt_h;
    F.idf = f->idf; F.trf = f->trf;
    return F;
}

static int run_conditioner(fs_flow* f, int li, const Workspace& w, int rows, int precision, int* nan_flag,
                           cudaStream_t s) {
    if (precision == FS_PREC_TF32) return tc_conditioner(f, li, w.A0, true, rows, w.theta, w.tc, w.tc_bytes, nan_flag, s);
    return conditioner_fp32(f, li, w.A0, rows, w.h, w.t, w.theta, s);
}

}  // namespace fs

using namespace fs;

extern "C" int fs_flow_create(const fs_flow_desc* d, fs_flow** out) {
    if (!d || !out || d->K < 1 || d->N < 1 || d->H < 1 || d->n_blocks < 0 || d->nb < 1 || !d->layers ||
        !d->identity_features || !d->transform_features || !(d->bound > 0)) {
        set_error("fs_flow_create: invalid descriptor");
        return FS_ERR_INVALID;
    }
    if (d->nb > 32) {   // one spline bin per lane
        set_error("fs_flow_create: num_bins=%d > 32 is not supported", d->nb);
        return FS_ERR_UNSUPPORTED;
    }
    fs_flow* f = new fs_flow();
    f->K = d->K; f->N = d->N; f->D = 2 * d->N; f->H = d->H; f->n_blocks = d->n_blocks; f->nb = d->nb;
    f->P = 3 * d->nb + 1;
    f->bound = d->bound;
    f->bound_f = (float)d->bound;
    f->pf_scale = (float)(M_PI / d->bound);                       // wrapper.py:151-154
    f->base_logc = (float)(-f->D) * logf((float)(2.0 * d->bound));  // Energy/Uniform.py:67-68
    f->inv_sqrt_h = (float)(1.0 / sqrt((double)d->H));             // coupling.py:340-342
    f->tc = nullptr;
    f->tc_err = nullptr;
    std::vector<int> idf(d->identity_features, d->identity_features + d->N);
    std::vector<int> trf(d->transform_features, d->transform_features + d->N);
    for (int j = 0; j < d->N; ++j) {
        if (idf[j] < 0 || idf[j] >= f->D || trf[j] < 0 || trf[j] >= f->D) {
            delete f;
            set_error("fs_flow_create: feature index out of range");
            return FS_ERR_INVALID;
        }
    }
    int r = upload(f, idf, &f->idf);
    if (!r) r = upload(f, trf, &f->trf);
    f->layers.resize(d->K);
    for (int i = 0; i < d->K && !r; ++i) r = pack_layer(f, d, &d->layers[i], &f->layers[i]);
    if (!r && spline_smem_bytes(f) > 226 * 1024) {
        set_error("fs_flow_create: num_bins too large for the spline kernel's shared-memory staging");
        r = FS_ERR_INVALID;
    }
    if (!r) {   // opt in to the device maximum once per create: the attribute is per function, not per flow
        r = cuda_check(cudaFuncSetAttribute(spline_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            226 * 1024), "spline smem");
        if (!r) r = cuda_check(cudaFuncSetAttribute(spline_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    226 * 1024), "spline smem");
    }
    if (!r) {
        int dev = 0;
        r = cuda_check(cudaGetDevice(&dev), "cudaGetDevice");
        if (!r) r = cuda_check(cudaDeviceGetAttribute(&f->sm_count, cudaDevAttrMultiProcessorCount, dev), "SM count");
    }
    if (!r) r = tc_pack(f, d);
    if (r) {
        fs_flow_destroy(f);
        return r;
    }
    *out = f;
    return FS_OK;
}

extern "C" void fs_flow_destroy(fs_flow* f) {
    if (!f) return;
    tc_free(f);
    for (void* p : f->allocs) cudaFree(p);
    delete f;
}

extern "C" size_t fs_flow_workspace_bytes(const fs_flow* f, int B, int precision) {
    if (!f || B < 1) return 0;
    return carve(f, B, precision, nullptr, nullptr);
}

static int check_ws(const fs_flow* f, int B, int precision, void* ws, size_t bytes, const char* who) {
    if (precision != FS_PREC_FP32 && precision != FS_PREC_TF32) {
        set_error("%s: unknown precision %d", who, precision);
        return FS_ERR_INVALID;
    }
    if (precision == FS_PREC_TF32 && !f->tc) {
        set_error("%s: the tensor-core path does not support this shape (H=%d); use FS_PREC_FP32", who, f->H);
        return FS_ERR_UNSUPPORTED;
    }
    size_t need = carve(f, B, precision, nullptr, nullptr);
    if (!ws || bytes < need) {
        set_error("%s: workspace too small (%zu < %zu)", who, bytes, need);
        return FS_ERR_INVALID;
    }
    return FS_OK;
}

// ResidualNet.forward of one layer's conditioner (nets/resnet.py:92-104, eval mode) on ready-made
// periodic features: features [rows, 2N] -> theta [rows, 3nb+1, N] (the library's parameter-major order).  rows must not exceed the chunk
// the workspace was sized for.
extern "C" int fs_flow_conditioner(fs_flow* f, int layer, const float* features, int rows, float* theta,
                                   void* workspace, size_t workspace_bytes, int precision, void* stream) {
    if (!f || !features || !theta || rows < 0 || layer < 0 || layer >= f->K) {
        set_error("fs_flow_conditioner: invalid argument");
        return FS_ERR_INVALID;
    }
    if (rows == 0) return FS_OK;
    if (int r = check_ws(f, rows, precision, workspace, workspace_bytes, "fs_flow_conditioner")) return r;
    if (rows > chunk_rows(f, rows)) {
        set_error("fs_flow_conditioner: rows=%d exceeds one workspace chunk (%d)", rows, chunk_rows(f, rows));
        return FS_ERR_INVALID;
    }
    Workspace w;
    carve(f, rows, precision, workspace, &w);
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == FS_PREC_TF32) return tc_conditioner(f, layer, features, false, rows, theta, w.tc, w.tc_bytes, nullptr, s);
    return conditioner_fp32(f, layer, features, rows, w.h, w.t, theta, s);
}

extern "C" int fs_flow_has_tensor_path(const fs_flow* f) { return (f && f->tc) ? 1 : 0; }

extern "C" int fs_flow_coupling(fs_flow* f, int layer, int direction, const float* features, const float* xin,
                                float* xout, float* logdet, int rows, int* nan_flag, void* stream) {
    const bool tiled = (direction & FS_FEATURES_TILED) != 0;
    direction &= ~FS_FEATURES_TILED;
    if (!f || !features || !xin || !xout || rows < 0 || layer < 0 || layer >= f->K || (direction != 1 && direction != 2)) {
        set_error("fs_flow_coupling: invalid argument");
        return FS_ERR_INVALID;
    }
    if (rows == 0) return FS_OK;
    return tc_conditioner_spline(f, layer, features, tiled, rows, direction, xin, xout, logdet, nan_flag,
                                 (cudaStream_t)stream);
}

// The conditioners + conditional splines of ALL K layers of a pass in one launch (what fs_flow_inverse / fs_flow_forward
// do after their K feature kernels; exposed for benchmarks and tests).  features: K row-tiled matrices in step order
// (density: layers K-1 .. 0; sampling: 0 .. K-1), fs_flow_tiled_features_bytes(rows, 2N) bytes apart; buf0 [rows, D] holds
// the input (its transformed columns are read), the steps ping-pong between buf0 and buf1 and step K-1 writes
// buf[K & 1]; logdet_parts [K, rows] receives one partial per step; scratch: K * ceil(rows / 128) * 8 ints.
extern "C" int fs_flow_coupling_all(fs_flow* f, int direction, const float* features, float* buf0, float* buf1,
                                    float* logdet_parts, int* scratch, int rows, int* nan_flag, void* stream) {
    if (!f || !features || !buf0 || !buf1 || !logdet_parts || !scratch || rows < 0 || (direction != 1 && direction != 2)) {
        set_error("fs_flow_coupling_all: invalid argument");
        return FS_ERR_INVALID;
    }
    if (rows == 0) return FS_OK;
    if (!tc_layer_parallel_ok(f)) {
        set_error("fs_flow_coupling_all: this flow has no layer-parallel path (fused tensor path, identity set closed under the roll)");
        return FS_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    FS_CUDA(cudaMemsetAsync(logdet_parts, 0, (size_t)f->K * rows * 4, s));
    FS_CUDA(cudaMemsetAsync(scratch, 0, tc_lp_flag_ints(f, rows) * 4, s));
    const size_t stride = (size_t)((rows + 127) / 128 * 128) * ((2 * f->N + 3) / 4 * 4);
    return tc_conditioner_spline_all(f, direction, rows, features, stride, buf0, buf1, logdet_parts, (size_t)rows, scratch,
                                     nan_flag, s);
}

// [rows, K0] row-major -> the row-tiled layout the tensor path reads (a0_tiled); the full passes write that layout
// directly from their feature kernels, this is for callers of fs_flow_coupling that hold a row-major matrix.
__global__ void tile_features_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int K0) {
    const size_t n = (size_t)rows * K0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / K0), k = (int)(i % K0);
        out[a0_tiled(b, k, K0)] = in[i];
    }
}

extern "C" size_t fs_flow_tiled_features_bytes(int rows, int K0) {
    if (rows <= 0 || K0 <= 0) return 0;
    return (size_t)((rows + 127) / 128 * 128) * (size_t)((K0 + 3) / 4 * 4) * 4;
}

extern "C" int fs_flow_tile_features(const float* features, int rows, int K0, float* tiled, void* stream) {
    if (rows < 0 || K0 <= 0 || (rows > 0 && (!features || !tiled))) {
        set_error("fs_flow_tile_features: invalid argument");
        return FS_ERR_INVALID;
    }
    if (rows == 0) return FS_OK;
    const size_t n = (size_t)rows * K0;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
    tile_features_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(features, tiled, rows, K0);
    fs::count_launch();
    return cuda_check(cudaGetLastError(), "tile_features_kernel");
}

extern "C" int fs_flow_inverse(fs_flow* f, const float* x, int B, double in_shift, float* z, float* logdet,
                               float* logq, int* nan_flag, void* workspace, size_t workspace_bytes, int precision,
                               void* stream) {
    if (!f || !x || B < 0) { set_error("fs_flow_inverse: invalid argument"); return FS_ERR_INVALID; }
    if (B == 0) return FS_OK;
    if (int r = check_ws(f, B, precision, workspace, workspace_bytes, "fs_flow_inverse")) return r;
    cudaStream_t s = (cudaStream_t)stream;
    Workspace w;
    carve(f, B, precision, workspace, &w);
    const int Bc = chunk_rows(f, B);
    const FlowDev F = flow_dev(f);
    const bool fused = fused_path(f, precision);
    for (int r0 = 0; r0 < B; r0 += Bc) {
        const int rows = (B - r0 < Bc) ? B - r0 : Bc;
        const size_t n = (size_t)rows * f->D;
        shift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x + (size_t)r0 * f->D, w.v0, n, -in_shift);
    fs::count_launch();
        FS_CUDA(cudaMemsetAsync(w.ld, 0, (size_t)rows * 4, s));
        float* cur = w.v0;
        float* nxt = w.v1;
        const bool lp = layer_parallel_rows(f, precision, rows);
        if (lp) {
            FS_CUDA(cudaMemsetAsync(w.ldp, 0, (size_t)f->K * Bc * 4, s));
            FS_CUDA(cudaMemsetAsync(w.flags, 0, tc_lp_flag_ints(f, Bc) * 4, s));
        }
        for (int li = f->K - 1; li >= 0; --li) {                        // core.py:82-85
            const fs_flow::Layer& L = f->layers[li];
            float* const a0 = w.A0 + (lp ? (size_t)(f->K - 1 - li) * w.a0_stride : 0);
            {
                const unsigned pg = (rows + FS_SPLINE_WARPS - 1) / FS_SPLINE_WARPS, pb = 32 * FS_SPLINE_WARPS;
                const bool fast = precision == FS_PREC_TF32;
#define FS_PREP(U, FAST) prep_inverse_v2<U, FAST><<<pg, pb, 0, s>>>(cur, nxt, a0, w.ld, rows, F, L.u_x, L.u_y, L.u_d, nan_flag)
#define FS_PREP3(FAST) prep_v3<FAST, true><<<prep3_grid(rows), 32 * FS_PREP3_WARPS, prep3_smem(f), s>>>(cur, nxt, a0, w.ld, rows, F, L.u_x, L.u_y, L.u_d, nan_flag)
                if (use_prep3(f)) { if (fast) FS_PREP3(true); else FS_PREP3(false); }
                else if (f->N > 32) { if (fast) FS_PREP(4, true); else FS_PREP(4, false); }
                else { if (fast) FS_PREP(1, true); else FS_PREP(1, false); }
#undef FS_PREP3
#undef FS_PREP
            }
    fs::count_launch();
            if (lp) {
                // identity columns only: the conditioners of all layers follow in one launch below
            } else if (fused) {
                if (int r = tc_conditioner_spline(f, li, w.A0, true, rows, 1, cur, nxt, w.ld, nan_flag, s)) return r;
            } else {
                if (int r = run_conditioner(f, li, w, rows, precision, nan_flag, s)) return r;
                spline_kernel<true><<<spline_grid(f, rows), 32 * spline_warps(f), spline_smem_bytes(f), s>>>(
                    cur, w.theta, nxt, w.ld, rows, F, nan_flag);
                fs::count_launch();
            }
            float* tmp = cur; cur = nxt; nxt = tmp;
        }
        if (lp)
            if (int r = tc_conditioner_spline_all(f, 1, rows, w.A0, w.a0_stride, w.v0, w.v1, w.ldp, (size_t)Bc, w.flags,
                                                  nan_flag, s))
                return r;
        finish_kernel<<<(rows + 7) / 8, 256, 0, s>>>(cur, z ? z + (size_t)r0 * f->D : nullptr, w.ld,
                                                     logdet ? logdet + r0 : nullptr, logq ? logq + r0 : nullptr,
                                                     rows, f->D, f->bound_f, f->base_logc, 0.0, w.ldp,
                                                     lp ? f->K : 0, (size_t)Bc);
    fs::count_launch();
        FS_CUDA(cudaGetLastError());
    }
    return FS_OK;
}

extern "C" int fs_flow_set_layer_parallel(fs_flow* f, int mode) {
    if (!f || mode < 0 || mode > 2) { set_error("fs_flow_set_layer_parallel: invalid argument"); return FS_ERR_INVALID; }
    f->lp_mode = mode;
    return FS_OK;
}

extern "C" int fs_flow_uses_layer_parallel(const fs_flow* f, int rows, int precision) {
    if (!f || rows <= 0) return 0;
    const int Bc = chunk_rows(f, rows);
    return layer_parallel_rows(f, precision, rows < Bc ? rows : Bc) ? 1 : 0;
}

extern "C" int fs_flow_forward(fs_flow* f, const float* zin, int B, double out_shift, float* x, float* logdet,
                               int* nan_flag, void* workspace, size_t workspace_bytes, int precision,
                               void* stream) {
    if (!f || !zin || !x || B < 0) { set_error("fs_flow_forward: invalid argument"); return FS_ERR_INVALID; }
    if (B == 0) return FS_OK;
    if (int r = check_ws(f, B, precision, workspace, workspace_bytes, "fs_flow_forward")) return r;
    cudaStream_t s = (cudaStream_t)stream;
    Workspace w;
    carve(f, B, precision, workspace, &w);
    const int Bc = chunk_rows(f, B);
    const FlowDev F = flow_dev(f);
    const bool fused = fused_path(f, precision);
    for (int r0 = 0; r0 < B; r0 += Bc) {
        const int rows = (B - r0 < Bc) ? B - r0 : Bc;
        const size_t n = (size_t)rows * f->D;
        FS_CUDA(cudaMemcpyAsync(w.v0, zin + (size_t)r0 * f->D, n * 4, cudaMemcpyDeviceToDevice, s));
        FS_CUDA(cudaMemsetAsync(w.ld, 0, (size_t)rows * 4, s));
        float* cur = w.v0;
        float* nxt = w.v1;
        const bool lp = layer_parallel_rows(f, precision, rows);
        if (lp) {
            FS_CUDA(cudaMemsetAsync(w.ldp, 0, (size_t)f->K * Bc * 4, s));
            FS_CUDA(cudaMemsetAsync(w.flags, 0, tc_lp_flag_ints(f, Bc) * 4, s));
        }
        for (int li = 0; li < f->K; ++li) {                              // core.py:52-55
            const fs_flow::Layer& L = f->layers[li];
            float* const a0 = w.A0 + (lp ? (size_t)li * w.a0_stride : 0);
            {
                const unsigned pg = (rows + FS_SPLINE_WARPS - 1) / FS_SPLINE_WARPS, pb = 32 * FS_SPLINE_WARPS;
                const bool fast = precision == FS_PREC_TF32;
#define FS_PREP(U, FAST) prep_forward_v2<U, FAST><<<pg, pb, 0, s>>>(cur, nxt, a0, w.ld, rows, F, L.u_x, L.u_y, L.u_d, nan_flag)
#define FS_PREP3(FAST) prep_v3<FAST, false><<<prep3_grid(rows), 32 * FS_PREP3_WARPS, prep3_smem(f), s>>>(cur, nxt, a0, w.ld, rows, F, L.u_x, L.u_y, L.u_d, nan_flag)
                if (use_prep3(f)) { if (fast) FS_PREP3(true); else FS_PREP3(false); }
                else if (f->N > 32) { if (fast) FS_PREP(4, true); else FS_PREP(4, false); }
                else { if (fast) FS_PREP(1, true); else FS_PREP(1, false); }
#undef FS_PREP3
#undef FS_PREP
            }
    fs::count_launch();
            if (lp) {
                // identity columns only: the conditioners of all layers follow in one launch below
            } else if (fused) {
                if (int r = tc_conditioner_spline(f, li, w.A0, true, rows, 2, cur, nxt, w.ld, nan_flag, s)) return r;
            } else {
                if (int r = run_conditioner(f, li, w, rows, precision, nan_flag, s)) return r;
                spline_kernel<false><<<spline_grid(f, rows), 32 * spline_warps(f), spline_smem_bytes(f), s>>>(
                    cur, w.theta, nxt, w.ld, rows, F, nan_flag);
                fs::count_launch();
            }
            float* tmp = cur; cur = nxt; nxt = tmp;
        }
        if (lp)
            if (int r = tc_conditioner_spline_all(f, 2, rows, w.A0, w.a0_stride, w.v0, w.v1, w.ldp, (size_t)Bc, w.flags,
                                                  nan_flag, s))
                return r;
        finish_kernel<<<(rows + 7) / 8, 256, 0, s>>>(cur, x + (size_t)r0 * f->D, w.ld,
                                                     logdet ? logdet + r0 : nullptr, nullptr, rows, f->D,
                                                     f->bound_f, f->base_logc, out_shift, w.ldp, lp ? f->K : 0,
                                                     (size_t)Bc);
    fs::count_launch();
        FS_CUDA(cudaGetLastError());
    }
    return FS_OK;
}
