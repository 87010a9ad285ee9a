// One forward-KL training step of the whole flow as ~36 kernel launches (forward + backward), gradients written
// straight into the caller's gradient tensors.
//
// fs_train_forward_kld <- NormalizingFlow.forward_kld + loss.backward() as the drivers run them every minibatch
//     (NF/normflows/core.py:88-108; hybrid_NF_MCMC/main_algorithm_1.py:297-320, main_algorithm_2.py:437-452):
//         loss = -mean_rows( sum_layers log|det J_layer| ),  layers K-1 .. 0 in the density direction,
//     each layer = CircularCoupledRationalQuadraticSpline.inverse (flows/neural_spline/wrapper.py:16-93 ->
//     coupling.py:86-102): unconditional spline on the identity features, ResidualNet(periodic features of the identity
//     features) -> conditional spline on the transformed features, scatter, roll by D/2.
//
// Eager autograd issues ~110 small kernels per layer and direction (5 k per step for the K = 23 flow of Algorithm 2),
// launch-bound even inside a CUDA graph (15 ms per step).  The structure used here is the one of the layer-parallel
// inference pass (flow_tc.cu): the identity set is closed under the roll, so
//   1. the identity values of ALL layers follow from the unconditional splines alone: one elementwise kernel walks a
//      (row, coordinate) through the K layers (chain_kernel<false, SHARED>) and leaves the conditioner inputs of every layer;
//   2. the K conditioners are independent: every stage of the ResidualNet (nets/resnet.py:7-104: Linear, then per block
//      BatchNorm1d(eps 1e-3, batch statistics) -> ReLU -> Linear twice with the residual, final Linear) is ONE launch
//      with the layer in blockIdx.z - FP32 SIMT tile GEMMs with the BatchNorm + ReLU of the operand recomputed in the
//      loader (gemm_nt / gemm_nn / gemm_tn), batch statistics and their backward in bn_stats / bn_bwd;
//   3. with all spline parameters known, the transformed coordinates are again an elementwise chain over the layers.
// The backward pass runs the same three stages in reverse (chain_kernel<true>, the transposed GEMMs, chain_kernel<true>
// with the conditioners' input gradients injected), parameters and gradients addressed through per-layer pointer
// tables, so nothing is flattened, copied or accumulated by autograd.  Row reductions (BatchNorm statistics, bias and
// shared-parameter gradients, the loss) are fixed-order: the step is repeatable bit for bit.
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "spline_train.cuh"

namespace fs {

enum { T_W0 = 0, T_B0 = 1, T_BLK = 2 };                      // per block: bn0.w bn0.b lin0.W lin0.b bn1.w bn1.b lin1.W lin1.b
static inline int t_per(int nbk) { return 2 + 8 * nbk + 5; }
static inline int t_wf(int nbk) { return 2 + 8 * nbk; }      // then bf, uw, uh, ud

// ---------------------------------------------------------------------------
// Tile GEMMs, 64 x 64 x 16, 256 threads, 4 x 4 outputs per thread; blockIdx.z = layer step.
// ---------------------------------------------------------------------------
#define FS_TG_ACC()                                                                                   \
    _Pragma("unroll") for (int k = 0; k < 16; ++k) {                                                  \
        float av[4], bv[4];                                                                           \
        _Pragma("unroll") for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];                      \
        _Pragma("unroll") for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];                      \
        _Pragma("unroll") for (int i = 0; i < 4; ++i)                                                 \
            _Pragma("unroll") for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]); \
    }

// C[m, n] = sum_k pro(A[m, k]) W[n, k] + bias[n] (+ R[m, n]);  PRO = 1: pro(a) = relu(a sc[k] + of[k]) (BatchNorm + ReLU)
template <int PRO, int RES>
__global__ void __launch_bounds__(256) gemm_nt(const float* __restrict__ A, int M, int Kd, int Nout, float* const* ptab,
                                               int per, int wi, int bi, const float* __restrict__ sc,
                                               const float* __restrict__ of, const float* __restrict__ R,
                                               float* __restrict__ C) {
    __shared__ float As[16][68], Bs[16][68];
    const int z = blockIdx.z, tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    A += (size_t)z * M * Kd;
    C += (size_t)z * M * Nout;
    if (RES) R += (size_t)z * M * Nout;
    const float* W = ptab[(size_t)z * per + wi];
    const float* bias = ptab[(size_t)z * per + bi];
    if (PRO) { sc += (size_t)z * Kd; of += (size_t)z * Kd; }
    float acc[4][4] = {};
    const int lr = tid / 4, lk = (tid % 4) * 4;
    for (int k0 = 0; k0 < Kd; k0 += 16) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = a;
        const int k = k0 + lk;
        if (k < Kd) {
            if (m0 + lr < M) {
                a = *reinterpret_cast<const float4*>(A + (size_t)(m0 + lr) * Kd + k);
                if (PRO) {
                    const float4 s4 = *reinterpret_cast<const float4*>(sc + k), o4 = *reinterpret_cast<const float4*>(of + k);
                    a.x = fmaxf(__fmaf_rn(a.x, s4.x, o4.x), 0.f); a.y = fmaxf(__fmaf_rn(a.y, s4.y, o4.y), 0.f);
                    a.z = fmaxf(__fmaf_rn(a.z, s4.z, o4.z), 0.f); a.w = fmaxf(__fmaf_rn(a.w, s4.w, o4.w), 0.f);
                }
            }
            if (n0 + lr < Nout) w = *reinterpret_cast<const float4*>(W + (size_t)(n0 + lr) * Kd + k);
        }
        As[lk][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
        Bs[lk][lr] = w.x; Bs[lk + 1][lr] = w.y; Bs[lk + 2][lr] = w.z; Bs[lk + 3][lr] = w.w;
        __syncthreads();
        FS_TG_ACC()
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
            if (RES) v += R[(size_t)m * Nout + n];
            C[(size_t)m * Nout + n] = v;
        }
    }
}

// dX[m, k] = sum_n dY[m, n] W[n, k];  MASK = 1: times [X[m, k] sc[k] + of[k] > 0] (the ReLU behind the BatchNorm of X)
// splits > 1 (MASK = 0 only): blockIdx.x = column tile * splits + part; part p sums its share of the n range and writes
// dX + p * part_stride (partials, added up in fixed order by sum_parts: the step stays bit-repeatable).  The final
// layer's dX has a reduction length of N (3 nb + 1) ~ 3 k against a 256 x 128 output per layer: without the split the
// launch is 184 blocks of 184 dependent load -> compute iterations.
template <int MASK>
__global__ void __launch_bounds__(256) gemm_nn(const float* __restrict__ dY, int M, int Nout, int Kd, float* const* ptab,
                                               int per, int wi, const float* __restrict__ X, const float* __restrict__ sc,
                                               const float* __restrict__ of, float* __restrict__ dX, int splits = 1,
                                               size_t part_stride = 0) {
    __shared__ float As[16][68], Bs[16][68];
    const int z = blockIdx.z, tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int part = (int)blockIdx.x % splits;
    const int m0 = blockIdx.y * 64, c0 = ((int)blockIdx.x / splits) * 64;
    dY += (size_t)z * M * Nout;
    dX += (size_t)z * M * Kd + (size_t)part * part_stride;
    const float* W = ptab[(size_t)z * per + wi];
    float acc[4][4] = {};
    const int lr = tid / 4, ln = (tid % 4) * 4;          // dY tile: row lr, 4 consecutive n
    const int wr = tid / 16, wc = (tid % 16) * 4;        // W tile: n-row wr, 4 consecutive columns
    const int nper = ((Nout + splits - 1) / splits + 15) / 16 * 16;
    const int n_begin = part * nper, n_end = min(Nout, n_begin + nper);
    for (int n0 = n_begin; n0 < n_end; n0 += 16) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = a;
        if (m0 + lr < M && n0 + ln < n_end) a = *reinterpret_cast<const float4*>(dY + (size_t)(m0 + lr) * Nout + n0 + ln);
        if (n0 + wr < n_end && c0 + wc < Kd) w = *reinterpret_cast<const float4*>(W + (size_t)(n0 + wr) * Kd + c0 + wc);
        As[ln][lr] = a.x; As[ln + 1][lr] = a.y; As[ln + 2][lr] = a.z; As[ln + 3][lr] = a.w;
        *reinterpret_cast<float4*>(&Bs[wr][wc]) = w;
        __syncthreads();
        FS_TG_ACC()
        __syncthreads();
    }
    if (MASK) { X += (size_t)z * M * Kd; sc += (size_t)z * Kd; of += (size_t)z * Kd; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + tx * 4 + j;
            if (c >= Kd) continue;
            float v = acc[i][j];
            if (MASK && !(__fmaf_rn(X[(size_t)m * Kd + c], sc[c], of[c]) > 0.f)) v = 0.f;
            dX[(size_t)m * Kd + c] = v;
        }
    }
}

// out[i] = parts[0][i] + parts[1][i] + ... in that order (float4 lanes)
__global__ void __launch_bounds__(256) sum_parts(const float* __restrict__ parts, size_t part_stride, int splits, size_t n4,
                                                 float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = reinterpret_cast<const float4*>(parts)[i];
    for (int p = 1; p < splits; ++p) {
        const float4 b = reinterpret_cast<const float4*>(parts + (size_t)p * part_stride)[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    reinterpret_cast<float4*>(out)[i] = a;
}

// dW[n, k] = sum_m dY[m, n] pro(A[m, k]),  db[n] = sum_m dY[m, n]  (written, not accumulated)
template <int PRO>
__global__ void __launch_bounds__(256) gemm_tn(const float* __restrict__ dY, int M, int Nout, int Kd,
                                               const float* __restrict__ A, const float* __restrict__ sc,
                                               const float* __restrict__ of, float* const* gtab, int per, int wi, int bi) {
    __shared__ float As[16][68], Bs[16][68];
    const int z = blockIdx.z, tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int n0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    dY += (size_t)z * M * Nout;
    A += (size_t)z * M * Kd;
    if (PRO) { sc += (size_t)z * Kd; of += (size_t)z * Kd; }
    float* dW = gtab[(size_t)z * per + wi];
    float* db = gtab[(size_t)z * per + bi];
    float acc[4][4] = {};
    float dbv[4] = {};
    const int r = tid / 16, c4 = (tid % 16) * 4;         // both tiles: m-row r, 4 consecutive columns
    float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), o4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (PRO && c0 + c4 < Kd) {
        s4 = *reinterpret_cast<const float4*>(sc + c0 + c4);
        o4 = *reinterpret_cast<const float4*>(of + c0 + c4);
    }
    for (int m0 = 0; m0 < M; m0 += 16) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (m0 + r < M) {
            if (n0 + c4 < Nout) a = *reinterpret_cast<const float4*>(dY + (size_t)(m0 + r) * Nout + n0 + c4);
            if (c0 + c4 < Kd) {
                b = *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * Kd + c0 + c4);
                if (PRO) {
                    b.x = fmaxf(__fmaf_rn(b.x, s4.x, o4.x), 0.f); b.y = fmaxf(__fmaf_rn(b.y, s4.y, o4.y), 0.f);
                    b.z = fmaxf(__fmaf_rn(b.z, s4.z, o4.z), 0.f); b.w = fmaxf(__fmaf_rn(b.w, s4.w, o4.w), 0.f);
                }
            }
        }
        *reinterpret_cast<float4*>(&As[r][c4]) = a;
        *reinterpret_cast<float4*>(&Bs[r][c4]) = b;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dbv[i] += av[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= Nout) continue;
        if (blockIdx.x == 0 && tx == 0) db[n] = dbv[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + tx * 4 + j;
            if (c < Kd) dW[(size_t)n * Kd + c] = acc[i][j];
        }
    }
}

// ---------------------------------------------------------------------------
// BatchNorm1d in training mode (torch.nn.BatchNorm1d(H, eps = 1e-3), nets/resnet.py:25-27): statistics over the rows of
// X[z] [M, H].  Block = 32 features x 8 row groups, blockIdx.y = z.  Writes sc = gamma rstd, of = beta - mean sc (what the
// GEMM loaders apply), mean, rstd, and updates running_mean / running_var (momentum, unbiased variance) in place.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_stats(const float* __restrict__ X, int M, int H, float* const* ptab, int per,
                                                int gi, int bti, float* const* rtab, int rper, int rmi, float eps,
                                                float momentum, int update, float* __restrict__ sc,
                                                float* __restrict__ of, float* __restrict__ mean_o,
                                                float* __restrict__ rstd_o) {
    __shared__ float red[8][33];
    const int z = blockIdx.y, f = threadIdx.x % 32, g = threadIdx.x / 32, c = blockIdx.x * 32 + f;
    X += (size_t)z * M * H;
    float s = 0.f;
    if (c < H) for (int m = g; m < M; m += 8) s += X[(size_t)m * H + c];
    red[g][f] = s;
    __syncthreads();
    float mean = 0.f;
    for (int i = 0; i < 8; ++i) mean += red[i][f];
    mean /= (float)M;
    __syncthreads();
    float v = 0.f;
    if (c < H) for (int m = g; m < M; m += 8) { const float d = X[(size_t)m * H + c] - mean; v = __fmaf_rn(d, d, v); }
    red[g][f] = v;
    __syncthreads();
    if (g == 0 && c < H) {
        float var = 0.f;
        for (int i = 0; i < 8; ++i) var += red[i][f];
        var /= (float)M;
        const float rstd = rsqrtf(var + eps);
        const float gam = ptab[(size_t)z * per + gi][c], bet = ptab[(size_t)z * per + bti][c];
        const size_t o = (size_t)z * H + c;
        sc[o] = gam * rstd;
        of[o] = bet - mean * gam * rstd;
        mean_o[o] = mean;
        rstd_o[o] = rstd;
        if (update) {
            float* rm = rtab[(size_t)z * rper + rmi];
            float* rv = rtab[(size_t)z * rper + rmi + 1];
            rm[c] = (1.0f - momentum) * rm[c] + momentum * mean;
            rv[c] = (1.0f - momentum) * rv[c] + momentum * var * ((float)M / (float)(M - 1));
        }
    }
}

// Backward of y = gamma (x - mean) rstd + beta over the batch: dA = dL/dy (already masked by the ReLU).
//   dgamma = sum dA xhat, dbeta = sum dA, dX = gamma rstd (dA - dbeta / M - xhat dgamma / M) (+ R)
__global__ void __launch_bounds__(256) bn_bwd(const float* __restrict__ dA, const float* __restrict__ X, int M, int H,
                                              float* const* ptab, float* const* gtab, int per, int gi, int bti,
                                              const float* __restrict__ mean_i, const float* __restrict__ rstd_i,
                                              const float* __restrict__ R, float* __restrict__ dX) {
    __shared__ float r1[8][33], r2[8][33];
    const int z = blockIdx.y, f = threadIdx.x % 32, g = threadIdx.x / 32, c = blockIdx.x * 32 + f;
    const size_t zo = (size_t)z * M * H;
    dA += zo; X += zo; dX += zo;
    if (R) R += zo;
    const float mean = c < H ? mean_i[(size_t)z * H + c] : 0.f, rstd = c < H ? rstd_i[(size_t)z * H + c] : 0.f;
    float s1 = 0.f, s2 = 0.f;
    if (c < H)
        for (int m = g; m < M; m += 8) {
            const float d = dA[(size_t)m * H + c];
            s1 += d;
            s2 = __fmaf_rn(d, (X[(size_t)m * H + c] - mean) * rstd, s2);
        }
    r1[g][f] = s1;
    r2[g][f] = s2;
    __syncthreads();
    s1 = 0.f; s2 = 0.f;
    for (int i = 0; i < 8; ++i) { s1 += r1[i][f]; s2 += r2[i][f]; }
    if (c >= H) return;
    if (g == 0) {
        gtab[(size_t)z * per + gi][c] = s2;
        gtab[(size_t)z * per + bti][c] = s1;
    }
    const float gr = ptab[(size_t)z * per + gi][c] * rstd, i1 = s1 / (float)M, i2 = s2 / (float)M;
    for (int m = g; m < M; m += 8) {
        const size_t o = (size_t)m * H + c;
        float v = gr * (dA[o] - i1 - (X[o] - mean) * rstd * i2);
        if (R) v += R[o];
        dX[o] = v;
    }
}

// ---------------------------------------------------------------------------
// Periodic features (utils/nn.py:65-137 as the fork uses it): cat[cos(s x), sin(s x)], and their backward.
// ---------------------------------------------------------------------------
__global__ void features_fwd(const float* __restrict__ x, long long rows, int N, float s, float* __restrict__ f) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * N) return;
    const long long r = e / N;
    const int j = (int)(e % N);
    float sn, cs;
    sincosf(s * x[e], &sn, &cs);
    f[r * 2 * N + j] = cs;
    f[r * 2 * N + N + j] = sn;
}
__global__ void features_bwd(const float* __restrict__ x, const float* __restrict__ df, long long rows, int N, float s,
                             float* __restrict__ dx) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * N) return;
    const long long r = e / N;
    const int j = (int)(e % N);
    float sn, cs;
    sincosf(s * x[e], &sn, &cs);
    dx[e] = s * (cs * df[r * 2 * N + N + j] - sn * df[r * 2 * N + j]);
}

// ---------------------------------------------------------------------------
// A coordinate's way through the K layers.  In the space of the transformed (or identity) features, index j of layer
// step z holds after the scatter + roll (coupling.py:100-101) index tau[j] of step z + 1.  Forward: thread (row, j0)
// starts from x0[row, cols[j0]], records the input of every step (xs[z][row, j]: the backward pass and - for the identity
// chain - the conditioners read it) and the sum of its log-dets.  Backward: thread (row, j_K) walks back through
// tau^-1, gld = the same constant for every element (-1 / rows for the forward-KL loss), the gradient of the value
// starts at 0 (the loss does not see z) and picks up inject[z][row, j] at every step (identity chain: dL/d(conditioner
// input)).  SHARED: the parameters of a step are shared by all rows (unconditional spline).
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// chain_lanes: the same coordinate chains with LPE (16 or 32) lanes per (row, coordinate) element - lane i owns bin i of
// the element's spline (its width, height and derivative logits): the three parameter segments are coalesced loads, the
// softmax normalisers are segmented shuffle reductions, the knots a segmented inclusive scan, the bin search a ballot,
// and in the backward pass lane i writes the gradient of "its" three parameters (coalesced stores).  The thread-per-
// element kernel above walks 46 parameters per layer through strided loads with 16 k threads in flight (3.5 warps per
// SM: latency-bound, 45 % of the training step); here the same pass has 16 / 32 x the threads and no strided access.
// The arithmetic per element is that of spline_point (spline_train.cuh) with the prefix sums taken in scan order.
// nb <= LPE; the (nb + 1)-th derivative logit is read by every lane (broadcast).
// ---------------------------------------------------------------------------
// rq_forward / softplus of spline_train.cuh with approximate reciprocals (1-2 ulp): the chain kernels below are bound by
// instruction issue, an IEEE division is ~8 instructions and log1pf ~30.  The log-det keeps logf.
// FS_CHAIN_FAST bits (development switch, -DFS_CHAIN_FAST=n through FS_NVCC_FLAGS): 1 softplus, 2 rational-quadratic
// forward, 4 softmax numerators with ex2.approx.  Bit 4 stays off: the gradient of the shared (unconditional) width
// logits is a sum over the rows of terms that cancel to ~1e-3 of their size, and the same-sign 1e-7 errors of the
// approximate exponential showed up as 1.8 % there (tests/test_gpu_train.py, N = 64 case); the others measure
// <= 1.4e-4 like the exact forms.
#ifndef FS_CHAIN_FAST
#define FS_CHAIN_FAST 3
#endif
__device__ __forceinline__ float softplus_fast_t(float x) {
    if (!(FS_CHAIN_FAST & 1)) return softplus_acc(x);
    return fmaxf(x, 0.f) + __logf(1.0f + __expf(-fabsf(x)));
}
__device__ __forceinline__ void rq_forward_fast(float x, float x0, float x1, float y0, float y1, float d0, float d1, RqFwd& f,
                                                float& y, float& ld) {
    if (!(FS_CHAIN_FAST & 2)) { rq_forward(x, x0, x1, y0, y1, d0, d1, f, y, ld); return; }
    f.wk = x1 - x0;
    f.hk = y1 - y0;
    const float rwk = __frcp_rn(f.wk);
    f.s = f.hk * rwk;
    f.th = (x - x0) * rwk;
    f.omt = 1.0f - f.th;
    f.tt = f.th * f.omt;
    f.t = d0 + d1 - 2.0f * f.s;
    f.den = f.s + f.t * f.tt;
    f.numA = f.s * f.th * f.th + d0 * f.tt;
    f.dn = d1 * f.th * f.th + 2.0f * f.s * f.tt + d0 * f.omt * f.omt;
    f.d0 = d0;
    f.d1 = d1;
    y = y0 + f.hk * f.numA * __frcp_rn(f.den);
    ld = logf(f.s * f.s * f.dn) - 2.0f * logf(f.den);             // utils/splines.py:214-222
}

template <int LPE>
__device__ __forceinline__ float seg_max(float v) {
#pragma unroll
    for (int o = LPE / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, LPE));
    return v;
}
template <int LPE>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
    for (int o = LPE / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, LPE);
    return v;
}
template <int LPE>
__device__ __forceinline__ float seg_scan(float v, int li) {      // inclusive prefix sum over the LPE lanes of a segment
#pragma unroll
    for (int o = 1; o < LPE; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, v, o, LPE);
        if (li >= o) v += t;
    }
    return v;
}

// One axis (widths or heights) of one element: softmax probabilities p (0 on pad lanes), prefix sums ps, and for bin k
// the knots lo / hi with the probability sums below them (p0, p1).  `k` < 0: the bin is searched on this axis (x given).
template <int LPE>
__device__ __forceinline__ void lane_axis(float logit, bool live, int li, int nb, float bound, float x, int& k, float& p,
                                          float& lo, float& hi, float& p0, float& p1) {
    const float c = 1.0f - kTMin * nb;
    const float m = seg_max<LPE>(live ? logit : -3.0e38f);
    const float e = live ? ((FS_CHAIN_FAST & 4) ? __expf(logit - m) : expf(logit - m)) : 0.f;   // logit - m <= 0
    const float rz = __frcp_rn(seg_sum<LPE>(e));
    p = e * rz;
    const float ps = seg_scan<LPE>(p, li);                                 // sum of p[0 .. li]
    const float cum = __fmaf_rn(c, ps, kTMin * (float)(li + 1));          // cumulative size at knot li + 1
    if (k < 0) {                                                           // utils/splines.py:11-13 on the interior knots
        const bool ge = li < nb - 1 && x >= 2.0f * bound * cum - bound;
        const unsigned seg = LPE == 32 ? 0xffffffffu : (0xffffu << (threadIdx.x & 16));
        k = __popc(__ballot_sync(0xffffffffu, ge) & seg);
    }
    const int base = (threadIdx.x & 31) & ~(LPE - 1);
    const float c1 = __shfl_sync(0xffffffffu, cum, base + k);
    const float q1 = __shfl_sync(0xffffffffu, ps, base + k);
    const float c0 = __shfl_sync(0xffffffffu, cum, base + (k > 0 ? k - 1 : 0));
    const float q0 = __shfl_sync(0xffffffffu, ps, base + (k > 0 ? k - 1 : 0));
    p0 = k > 0 ? q0 : 0.f;
    p1 = q1;
    lo = (k == 0) ? -bound : 2.0f * bound * (k > 0 ? c0 : 0.f) - bound;
    hi = (k == nb - 1) ? bound : 2.0f * bound * c1 - bound;
}

template <bool BWD, bool SHARED, int LPE>
__global__ void __launch_bounds__(128) chain_lanes(const float* __restrict__ x0, int D, const int* __restrict__ cols,
                                                   const int* __restrict__ tau, int rows, int N, int K, int nb, float bound,
                                                   float scale, const float* __restrict__ theta, float* __restrict__ xs,
                                                   float* __restrict__ ldsum, float gld, const float* __restrict__ inject,
                                                   float* __restrict__ gtheta) {
    const long long total = (long long)rows * N;
    long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPE;
    const bool valid = e < total;                      // whole segments: the grid covers total * LPE threads exactly or pads
    if (!valid) e = total - 1;                         // padding segments shadow the last element (no stores)
    const int li = threadIdx.x & (LPE - 1);
    const int base = (threadIdx.x & 31) & ~(LPE - 1);
    const int r = (int)(e / N);
    int j = (int)(e % N);
    const int P = 3 * nb + 1;
    const size_t step = (size_t)rows * N;
    const bool live = li < nb;
    const float c = 1.0f - kTMin * nb, two_b = 2.0f * bound;
    if (!BWD) {
        float x = x0[(size_t)r * D + cols[j]], lds = 0.f;
        for (int z = 0; z < K; ++z) {
            const size_t o = z * step + (size_t)r * N + j;
            if (valid && li == 0) xs[o] = x;
            const float* u = SHARED ? theta + ((size_t)z * N + j) * P : theta + o * P;
            const float uw = live ? u[li] * scale : 0.f, uh = live ? u[nb + li] * scale : 0.f;
            const float ud = live ? u[2 * nb + li] : 0.f, ud_last = u[3 * nb];
            const bool in = x >= -bound && x <= bound;                    // tails: identity (utils/splines.py:24-39)
            const float xc = in ? x : 0.f;
            int k = -1;
            float pw, ph, xlo, xhi, ylo, yhi, a0, a1;
            lane_axis<LPE>(uw, live, li, nb, bound, xc, k, pw, xlo, xhi, a0, a1);
            lane_axis<LPE>(uh, live, li, nb, bound, xc, k, ph, ylo, yhi, a0, a1);
            const float ud0 = __shfl_sync(0xffffffffu, ud, base + k);
            const float udn = __shfl_sync(0xffffffffu, ud, base + (k + 1 < LPE ? k + 1 : k));
            const float ud1 = (k + 1 < nb) ? udn : ud_last;
            RqFwd f;
            float y, ld;
            rq_forward_fast(xc, xlo, xhi, ylo, yhi, kTMin + softplus_fast_t(ud0), kTMin + softplus_fast_t(ud1), f, y, ld);
            lds += in ? ld : 0.f;
            x = in ? y : x;
            j = tau[j];
        }
        if (valid && li == 0) ldsum[e] = lds;
    } else {
        float gy = 0.f;
        for (int z = K - 1; z >= 0; --z) {
            j = tau[j];                                            // tau here is the inverse table
            const size_t o = z * step + (size_t)r * N + j;
            const float* u = SHARED ? theta + ((size_t)z * N + j) * P : theta + o * P;
            float* gt = gtheta + o * P;
            const float x = xs[o];
            const float uw = live ? u[li] * scale : 0.f, uh = live ? u[nb + li] * scale : 0.f;
            const float ud = live ? u[2 * nb + li] : 0.f, ud_last = u[3 * nb];
            const bool in = x >= -bound && x <= bound;
            const float xc = in ? x : 0.f;
            int k = -1;
            float pw, ph, xlo, xhi, ylo, yhi, wp0, wp1, hp0, hp1;
            lane_axis<LPE>(uw, live, li, nb, bound, xc, k, pw, xlo, xhi, wp0, wp1);
            lane_axis<LPE>(uh, live, li, nb, bound, xc, k, ph, ylo, yhi, hp0, hp1);
            const float ud0 = __shfl_sync(0xffffffffu, ud, base + k);
            const float udn = __shfl_sync(0xffffffffu, ud, base + (k + 1 < LPE ? k + 1 : k));
            const float ud1 = (k + 1 < nb) ? udn : ud_last;
            RqFwd f;
            float yv, lv;
            rq_forward_fast(xc, xlo, xhi, ylo, yhi, kTMin + softplus_fast_t(ud0), kTMin + softplus_fast_t(ud1), f, yv, lv);
            // ---- reverse mode through the rational-quadratic formula (as spline_point<true>) ----
            const float gyv = gy, gl = gld;
            const float rden = __frcp_rn(f.den), rwk = __frcp_rn(f.wk);
            const float q = f.numA * rden;
            float hk_b = gyv * q;
            const float q_b = gyv * f.hk;
            const float numA_b = q_b * rden;
            float den_b = (-q_b * q - 2.0f * gl) * rden;
            float s_b = __fdividef(2.0f * gl, f.s);
            const float dn_b = __fdividef(gl, f.dn);
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
            const float x_b = th_b * rwk;
            float x0_b = -x_b;
            float wk_b = -th_b * f.th * rwk;
            hk_b += s_b * rwk;
            wk_b += -s_b * f.s * rwk;
            const float x1_b = wk_b;
            x0_b -= wk_b;
            const float y1_b = hk_b;
            const float y0_b = gyv - hk_b;
            // ---- knots -> cumulative sizes -> softmax logits: lane li writes the gradients of its three logits ----
            float gw = 0.f, gh = 0.f, gd = 0.f, gd_last = 0.f;
            if (in) {
                {
                    const float G0 = (k == 0) ? 0.f : two_b * x0_b, G1 = (k == nb - 1) ? 0.f : two_b * x1_b;
                    const float dot = G0 * wp0 + G1 * wp1;
                    gw = scale * c * pw * ((li < k ? G0 : 0.f) + (li < k + 1 ? G1 : 0.f) - dot);
                }
                {
                    const float G0 = (k == 0) ? 0.f : two_b * y0_b, G1 = (k == nb - 1) ? 0.f : two_b * y1_b;
                    const float dot = G0 * hp0 + G1 * hp1;
                    gh = scale * c * ph * ((li < k ? G0 : 0.f) + (li < k + 1 ? G1 : 0.f) - dot);
                }
                const float g0 = __fdividef(d0_b, 1.0f + __expf(-ud0)), g1 = __fdividef(d1_b, 1.0f + __expf(-ud1));   // d softplus = sigmoid
                gd = (li == k) ? g0 : ((li == k + 1) ? g1 : 0.f);
                gd_last = (k + 1 == nb) ? g1 : 0.f;
            }
            if (valid) {
                if (live) {
                    gt[li] = gw;
                    gt[nb + li] = gh;
                    gt[2 * nb + li] = gd;
                }
                if (li == 0) gt[3 * nb] = gd_last;
            }
            gy = (in ? x_b : gyv) + (inject ? inject[o] : 0.f);
        }
    }
}

template <bool BWD, bool SHARED>
__global__ void __launch_bounds__(128) chain_kernel(const float* __restrict__ x0, int D, const int* __restrict__ cols,
                                                    const int* __restrict__ tau, int rows, int N, int K, int nb, float bound,
                                                    float scale, const float* __restrict__ theta, float* __restrict__ xs,
                                                    float* __restrict__ ldsum, float gld, const float* __restrict__ inject,
                                                    float* __restrict__ gtheta) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)rows * N) return;
    const int r = (int)(e / N);
    int j = (int)(e % N);
    const int P = 3 * nb + 1;
    const size_t step = (size_t)rows * N;
    if (!BWD) {
        float x = x0[(size_t)r * D + cols[j]], lds = 0.f;
        for (int z = 0; z < K; ++z) {
            xs[z * step + (size_t)r * N + j] = x;
            const float* u = SHARED ? theta + ((size_t)z * N + j) * P : theta + (z * step + (size_t)r * N + j) * P;
            float y, ld, gx;
            spline_point<false>(x, u, nb, bound, scale, y, ld, 0.f, 0.f, gx, nullptr);
            lds += ld;
            x = y;
            j = tau[j];
        }
        ldsum[e] = lds;
    } else {
        float gy = 0.f;
        for (int z = K - 1; z >= 0; --z) {
            j = tau[j];                                            // tau here is the inverse table
            const size_t o = z * step + (size_t)r * N + j;
            const float* u = SHARED ? theta + ((size_t)z * N + j) * P : theta + o * P;
            float y, ld, gx;
            spline_point<true>(xs[o], u, nb, bound, scale, y, ld, gy, gld, gx, gtheta + o * P);
            gy = gx + (inject ? inject[o] : 0.f);
        }
    }
}

// theta_shared[z][j] = [uw[j] | uh[j] | ud[j]] of step z (coupling.py:208-238: the unconditional transform's parameters)
__global__ void gather_uncond(float* const* ptab, int per, int ui, int K, int N, int nb, float* __restrict__ th) {
    const int P = 3 * nb + 1;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)K * N * P) return;
    const int p = (int)(e % P), j = (int)((e / P) % N), z = (int)(e / ((long long)P * N));
    float v;
    if (p < nb) v = ptab[(size_t)z * per + ui][j * nb + p];
    else if (p < 2 * nb) v = ptab[(size_t)z * per + ui + 1][j * nb + p - nb];
    else v = ptab[(size_t)z * per + ui + 2][j * (nb + 1) + p - 2 * nb];
    th[e] = v;
}
// ... and the gradient: sum over the rows of gth[z][row][j][p], in row order
__global__ void scatter_uncond(const float* __restrict__ gth, int rows, float* const* gtab, int per, int ui, int K, int N,
                               int nb) {
    const int P = 3 * nb + 1;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)K * N * P) return;
    const int p = (int)(e % P), j = (int)((e / P) % N), z = (int)(e / ((long long)P * N));
    const float* src = gth + ((size_t)z * rows * N + j) * P + p;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += src[(size_t)r * N * P];
    if (p < nb) gtab[(size_t)z * per + ui][j * nb + p] = s;
    else if (p < 2 * nb) gtab[(size_t)z * per + ui + 1][j * nb + p - nb] = s;
    else gtab[(size_t)z * per + ui + 2][j * (nb + 1) + p - 2 * nb] = s;
}

// loss = -(sum a + sum b) / rows, one block, fixed order
__global__ void __launch_bounds__(1024) loss_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                    int rows, float* __restrict__ loss) {
    __shared__ double red[1024];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) s += (double)a[i] + (double)b[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = (float)(-red[0] / (double)rows);
}

}  // namespace fs

using namespace fs;

struct fs_train {
    int K, N, H, nbk, nb, P, D, per, rper;
    float bound, fscale, eps, momentum;
    float** ptab = nullptr;      // device: parameters, step-major (step z = layer K - 1 - z)
    float** gtab = nullptr;      // device: gradients
    float** rtab = nullptr;      // device: BatchNorm running statistics
    int* tabs = nullptr;         // device: colsT | colsI | tauT | tauI | tauT^-1 | tauI^-1, N ints each
    float* ws = nullptr;         // workspace for `cap` rows
    int cap = 0;
    size_t ws_bytes = 0;
};

namespace {
static constexpr int FINAL_DX_SPLITS = 4;
struct Carve {
    float *thI, *xsI, *xsT, *ldI, *ldT, *feat, *hs, *ts, *st, *theta, *dtheta, *dhA, *dhB, *d1, *d2, *dfeat, *dident, *parts;
};
size_t carve_train(const fs_train* t, int B, float* base, Carve* c) {
    size_t off = 0;
    auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 63) / 64 * 64; return p; };
    const size_t K = t->K, N = t->N, H = t->H, P = t->P, b = B;
    Carve v;
    v.thI = take(K * N * P);
    v.xsI = take(K * b * N);
    v.xsT = take(K * b * N);
    v.ldI = take(b * N);
    v.ldT = take(b * N);
    v.feat = take(K * b * 2 * N);
    v.hs = take((size_t)(t->nbk + 1) * K * b * H);
    v.ts = take((size_t)t->nbk * K * b * H);
    v.st = take((size_t)2 * t->nbk * 4 * K * H);             // per BatchNorm site: sc | of | mean | rstd
    v.theta = take(K * b * N * P);
    v.dtheta = take(K * b * N * P);
    v.dhA = take(K * b * H);
    v.dhB = take(K * b * H);
    v.d1 = take(K * b * H);
    v.d2 = take(K * b * H);
    v.dfeat = take(K * b * 2 * N);
    v.dident = take(K * b * N);
    v.parts = take((size_t)FINAL_DX_SPLITS * K * b * H);     // partials of the final layer's dX (split reduction)
    if (c) *c = v;
    return off * sizeof(float);
}
}  // namespace

extern "C" int fs_train_create(const fs_train_desc* d, fs_train** out) {
    if (!d || !out || d->K < 1 || d->N < 2 || d->H < 4 || d->n_blocks < 0 || d->nb < 1 || !d->params || !d->grads ||
        !d->transform_features || !d->identity_features || !(d->bound > 0) || (d->n_blocks > 0 && !d->bn_running)) {
        set_error("fs_train_create: invalid descriptor");
        return FS_ERR_INVALID;
    }
    if ((d->N & 1) || (d->H & 3) || ((d->N * (3 * d->nb + 1)) & 3)) {
        set_error("fs_train_create: N must be even, H and N (3 nb + 1) multiples of 4 (N=%d, H=%d, nb=%d)", d->N, d->H, d->nb);
        return FS_ERR_UNSUPPORTED;
    }
    const int N = d->N, D = 2 * N, h = D / 2, K = d->K;
    // the chains need both feature sets to be closed under the roll by D / 2
    std::vector<int> posT(D, -1), posI(D, -1), tabs(6 * N);
    for (int j = 0; j < N; ++j) {
        const int a = d->transform_features[j], b = d->identity_features[j];
        if (a < 0 || a >= D || b < 0 || b >= D || posT[a] >= 0 || posI[b] >= 0) {
            set_error("fs_train_create: invalid feature lists");
            return FS_ERR_INVALID;
        }
        posT[a] = j;
        posI[b] = j;
    }
    for (int j = 0; j < N; ++j) {
        const int a = d->transform_features[j], b = d->identity_features[j];
        // out[:, a] moves to column (a + h) % D of the rolled layer output (torch.cat([out[:, h:], out[:, :h]]))
        const int ta = posT[(a + h) % D], tb = posI[(b + h) % D];
        if (ta < 0 || tb < 0 || posI[a] >= 0) {
            set_error("fs_train_create: the identity features are not closed under the roll by D/2");
            return FS_ERR_UNSUPPORTED;
        }
        tabs[j] = a;
        tabs[N + j] = b;
        tabs[2 * N + j] = ta;
        tabs[3 * N + j] = tb;
        tabs[4 * N + ta] = j;
        tabs[5 * N + tb] = j;
    }
    fs_train* t = new fs_train();
    t->K = K; t->N = N; t->H = d->H; t->nbk = d->n_blocks; t->nb = d->nb; t->P = 3 * d->nb + 1; t->D = D;
    t->per = t_per(d->n_blocks);
    t->rper = 4 * d->n_blocks;
    t->bound = (float)d->bound; t->fscale = (float)d->feature_scale; t->eps = (float)d->bn_eps;
    t->momentum = (float)d->bn_momentum;
    // tables in step order: step z of the density pass is layer K - 1 - z (core.py:88-93)
    std::vector<float*> p((size_t)K * t->per), g((size_t)K * t->per), r((size_t)K * (t->rper ? t->rper : 1), nullptr);
    for (int z = 0; z < K; ++z) {
        const int li = K - 1 - z;
        for (int i = 0; i < t->per; ++i) {
            p[(size_t)z * t->per + i] = d->params[(size_t)li * t->per + i];
            g[(size_t)z * t->per + i] = d->grads[(size_t)li * t->per + i];
            if (!p[(size_t)z * t->per + i] || !g[(size_t)z * t->per + i]) {
                delete t;
                set_error("fs_train_create: null parameter / gradient pointer (layer %d, entry %d)", li, i);
                return FS_ERR_INVALID;
            }
        }
        for (int i = 0; i < t->rper; ++i) r[(size_t)z * t->rper + i] = d->bn_running[(size_t)li * t->rper + i];
    }
    auto up = [&](const void* src, size_t bytes, void** dst) {
        if (cudaMalloc(dst, bytes) != cudaSuccess) return FS_ERR_CUDA;
        return cuda_check(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice), "cudaMemcpy");
    };
    int rc = up(p.data(), p.size() * sizeof(float*), (void**)&t->ptab);
    if (!rc) rc = up(g.data(), g.size() * sizeof(float*), (void**)&t->gtab);
    if (!rc) rc = up(r.data(), r.size() * sizeof(float*), (void**)&t->rtab);
    if (!rc) rc = up(tabs.data(), tabs.size() * sizeof(int), (void**)&t->tabs);
    if (rc) { fs_train_destroy(t); set_error("fs_train_create: device allocation failed"); return rc; }
    *out = t;
    return FS_OK;
}

extern "C" void fs_train_destroy(fs_train* t) {
    if (!t) return;
    cudaFree(t->ptab); cudaFree(t->gtab); cudaFree(t->rtab); cudaFree(t->tabs); cudaFree(t->ws);
    delete t;
}

extern "C" int fs_train_forward_kld(fs_train* t, const float* x, int B, float* loss, int update_running, void* stream) {
    if (!t || !x || !loss || B < 2) {
        set_error("fs_train_forward_kld: invalid argument (BatchNorm needs at least two rows)");
        return FS_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (B > t->cap) {                                              // grow the workspace (synchronises: rare)
        FS_CUDA(cudaStreamSynchronize(s));
        if (t->ws) cudaFree(t->ws);
        t->ws = nullptr;
        t->cap = 0;
        const size_t bytes = carve_train(t, B, nullptr, nullptr);
        if (cudaMalloc((void**)&t->ws, bytes) != cudaSuccess) {
            set_error("fs_train_forward_kld: cannot allocate %zu bytes of workspace", bytes);
            return FS_ERR_CUDA;
        }
        t->cap = B;
        t->ws_bytes = bytes;
    }
    Carve c;
    carve_train(t, B, t->ws, &c);
    const int K = t->K, N = t->N, H = t->H, P = t->P, nb = t->nb, nbk = t->nbk, per = t->per, NP = N * P;
    const int *colsT = t->tabs, *colsI = t->tabs + N, *tauT = t->tabs + 2 * N, *tauI = t->tabs + 3 * N,
              *itauT = t->tabs + 4 * N, *itauI = t->tabs + 5 * N;
    const size_t KBH = (size_t)K * B * H, KH = (size_t)K * H;
    const unsigned ce = (unsigned)(((long long)B * N + 127) / 128);
    const long long kbn = (long long)K * B * N;
    auto site = [&](int b, int j, int what) { return c.st + ((size_t)(2 * b + j) * 4 + what) * KH; };   // sc of mean rstd
    auto grid2 = [&](int cols, int rows_) { return dim3((cols + 63) / 64, (rows_ + 63) / 64, K); };
    const dim3 bng((H + 31) / 32, K);
    int launches = 0;
    // coordinate chains: lane-per-bin kernel for nb <= 32 (16 lanes per element up to nb = 16), else one thread per element
    auto chain = [&](auto bwd, auto shared, const float* x0, const int* cols, const int* tau, float scale_, const float* th,
                     float* xs_, float* lds_, float gld_, const float* inj, float* gth) {
        constexpr bool BW = decltype(bwd)::value, SH = decltype(shared)::value;
        static const bool scalar_chain = getenv("FS_TRAIN_CHAIN_SCALAR") != nullptr;   // development: thread-per-element kernel
        if (nb <= 16 && !scalar_chain) {
            const unsigned g16 = (unsigned)(((long long)B * N * 16 + 127) / 128);
            chain_lanes<BW, SH, 16><<<g16, 128, 0, s>>>(x0, t->D, cols, tau, B, N, K, nb, t->bound, scale_, th, xs_, lds_, gld_, inj, gth);
        } else if (nb <= 32 && !scalar_chain) {
            const unsigned g32 = (unsigned)(((long long)B * N * 32 + 127) / 128);
            chain_lanes<BW, SH, 32><<<g32, 128, 0, s>>>(x0, t->D, cols, tau, B, N, K, nb, t->bound, scale_, th, xs_, lds_, gld_, inj, gth);
        } else {
            chain_kernel<BW, SH><<<ce, 128, 0, s>>>(x0, t->D, cols, tau, B, N, K, nb, t->bound, scale_, th, xs_, lds_, gld_, inj, gth);
        }
    };
    using std::true_type;
    using std::false_type;

    // ---- forward ----
    gather_uncond<<<(unsigned)(((long long)K * NP + 255) / 256), 256, 0, s>>>(t->ptab, per, t_wf(nbk) + 2, K, N, nb, c.thI);
    chain(false_type{}, true_type{}, x, colsI, tauI, 1.0f, c.thI, c.xsI, c.ldI, 0.f, nullptr, nullptr);
    features_fwd<<<(unsigned)((kbn + 255) / 256), 256, 0, s>>>(c.xsI, (long long)K * B, N, t->fscale, c.feat);
    gemm_nt<0, 0><<<grid2(H, B), 256, 0, s>>>(c.feat, B, 2 * N, H, t->ptab, per, T_W0, T_B0, nullptr, nullptr, nullptr, c.hs);
    launches += 4;
    for (int b = 0; b < nbk; ++b) {
        const int pb = T_BLK + 8 * b;
        float* hb = c.hs + (size_t)b * KBH;
        float* tb = c.ts + (size_t)b * KBH;
        bn_stats<<<bng, 256, 0, s>>>(hb, B, H, t->ptab, per, pb + 0, pb + 1, t->rtab, t->rper, 4 * b, t->eps, t->momentum,
                                     update_running, site(b, 0, 0), site(b, 0, 1), site(b, 0, 2), site(b, 0, 3));
        gemm_nt<1, 0><<<grid2(H, B), 256, 0, s>>>(hb, B, H, H, t->ptab, per, pb + 2, pb + 3, site(b, 0, 0), site(b, 0, 1),
                                                  nullptr, tb);
        bn_stats<<<bng, 256, 0, s>>>(tb, B, H, t->ptab, per, pb + 4, pb + 5, t->rtab, t->rper, 4 * b + 2, t->eps,
                                     t->momentum, update_running, site(b, 1, 0), site(b, 1, 1), site(b, 1, 2),
                                     site(b, 1, 3));
        gemm_nt<1, 1><<<grid2(H, B), 256, 0, s>>>(tb, B, H, H, t->ptab, per, pb + 6, pb + 7, site(b, 1, 0), site(b, 1, 1), hb,
                                                  hb + KBH);
        launches += 4;
    }
    float* hlast = c.hs + (size_t)nbk * KBH;
    gemm_nt<0, 0><<<grid2(NP, B), 256, 0, s>>>(hlast, B, H, NP, t->ptab, per, t_wf(nbk), t_wf(nbk) + 1, nullptr, nullptr,
                                               nullptr, c.theta);
    chain(false_type{}, false_type{}, x, colsT, tauT, 1.0f / sqrtf((float)H), c.theta, c.xsT, c.ldT, 0.f, nullptr, nullptr);
    loss_kernel<<<1, 1024, 0, s>>>(c.ldI, c.ldT, (long long)B * N, B, loss);
    launches += 3;

    // ---- backward: d loss / d log-det = -1 / B for every element ----
    const float gld = -1.0f / (float)B;
    chain(true_type{}, false_type{}, nullptr, nullptr, itauT, 1.0f / sqrtf((float)H), c.theta, c.xsT, nullptr, gld, nullptr,
          c.dtheta);
    gemm_tn<0><<<dim3((H + 63) / 64, (NP + 63) / 64, K), 256, 0, s>>>(c.dtheta, B, NP, H, hlast, nullptr, nullptr, t->gtab, per,
                                                                      t_wf(nbk), t_wf(nbk) + 1);
    if (NP >= 1024 && (KBH & 3) == 0) {                   // long reduction, small output: split it four ways
        gemm_nn<0><<<dim3(((H + 63) / 64) * FINAL_DX_SPLITS, (B + 63) / 64, K), 256, 0, s>>>(
            c.dtheta, B, NP, H, t->ptab, per, t_wf(nbk), nullptr, nullptr, nullptr, c.parts, FINAL_DX_SPLITS, KBH);
        sum_parts<<<(unsigned)((KBH / 4 + 255) / 256), 256, 0, s>>>(c.parts, KBH, FINAL_DX_SPLITS, KBH / 4, c.dhA);
        ++launches;
    } else {
        gemm_nn<0><<<grid2(H, B), 256, 0, s>>>(c.dtheta, B, NP, H, t->ptab, per, t_wf(nbk), nullptr, nullptr, nullptr, c.dhA);
    }
    launches += 3;
    float *dA = c.dhA, *dB = c.dhB;
    for (int b = nbk - 1; b >= 0; --b) {
        const int pb = T_BLK + 8 * b;
        float* hb = c.hs + (size_t)b * KBH;
        float* tb = c.ts + (size_t)b * KBH;
        gemm_tn<1><<<dim3((H + 63) / 64, (H + 63) / 64, K), 256, 0, s>>>(dA, B, H, H, tb, site(b, 1, 0), site(b, 1, 1), t->gtab,
                                                                         per, pb + 6, pb + 7);
        gemm_nn<1><<<grid2(H, B), 256, 0, s>>>(dA, B, H, H, t->ptab, per, pb + 6, tb, site(b, 1, 0), site(b, 1, 1), c.d1);
        bn_bwd<<<bng, 256, 0, s>>>(c.d1, tb, B, H, t->ptab, t->gtab, per, pb + 4, pb + 5, site(b, 1, 2), site(b, 1, 3), nullptr,
                                   c.d2);
        gemm_tn<1><<<dim3((H + 63) / 64, (H + 63) / 64, K), 256, 0, s>>>(c.d2, B, H, H, hb, site(b, 0, 0), site(b, 0, 1),
                                                                         t->gtab, per, pb + 2, pb + 3);
        gemm_nn<1><<<grid2(H, B), 256, 0, s>>>(c.d2, B, H, H, t->ptab, per, pb + 2, hb, site(b, 0, 0), site(b, 0, 1), c.d1);
        bn_bwd<<<bng, 256, 0, s>>>(c.d1, hb, B, H, t->ptab, t->gtab, per, pb + 0, pb + 1, site(b, 0, 2), site(b, 0, 3), dA, dB);
        float* tmp = dA; dA = dB; dB = tmp;
        launches += 6;
    }
    gemm_tn<0><<<dim3((2 * N + 63) / 64, (H + 63) / 64, K), 256, 0, s>>>(dA, B, H, 2 * N, c.feat, nullptr, nullptr, t->gtab, per,
                                                                         T_W0, T_B0);
    gemm_nn<0><<<grid2(2 * N, B), 256, 0, s>>>(dA, B, H, 2 * N, t->ptab, per, T_W0, nullptr, nullptr, nullptr, c.dfeat);
    features_bwd<<<(unsigned)((kbn + 255) / 256), 256, 0, s>>>(c.xsI, c.dfeat, (long long)K * B, N, t->fscale, c.dident);
    // identity chain: the per-row parameter gradients reuse the dtheta buffer (the conditioners are done with it)
    chain(true_type{}, true_type{}, nullptr, nullptr, itauI, 1.0f, c.thI, c.xsI, nullptr, gld, c.dident, c.dtheta);
    scatter_uncond<<<(unsigned)(((long long)K * NP + 255) / 256), 256, 0, s>>>(c.dtheta, B, t->gtab, per, t_wf(nbk) + 2, K, N, nb);
    launches += 5;
    count_launch(launches);
    return cuda_check(cudaGetLastError(), "fs_train_forward_kld");
}
