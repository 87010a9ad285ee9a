// Device-side re-pack of a flow whose parameters changed (Algorithm 2 trains the flow every cycle).
//
// fs_flow_update <- the analogue of `model.eval()` after an optimizer step
//                   (hybrid_NF_MCMC/main_algorithm_2.py:450-451 -> :476): same result as fs_flow_create on the new
//                   parameters, but reading them where they live (device memory) and writing every packed buffer in place:
//                   BatchNorm folding, bias folding of the residual stream, the swizzled TF32 / FP16 weight streams of
//                   the tensor path, the FP32-path matrices, the unconditional-spline knot tables.  No device-to-host
//                   copy, no host loop, no allocation: a handful of kernels over all K layers (grid.y = layer).
//
// The arithmetic mirrors the host pack (flow.cu: pack_layer, flow_tc.cu: tc_pack) step by step - float64 folding, the
// same rounding points - so an updated flow agrees with a freshly created one to the last bit except where libm's and
// CUDA's float64 exp differ in the final ulp (knot tables) and where a dot product is summed in a different order
// (folded final-layer biases).
#include <cuda_fp16.h>

#include <vector>

#include "flow.cuh"

namespace fs {

struct RepackLayer {
    // sources: reference state_dict layout (fs_layer_params), DEVICE pointers
    const float *init_w, *init_b, *bn_w, *bn_b, *bn_mean, *bn_var, *lin_w, *lin_b, *final_w, *final_b, *un_w, *un_h, *un_d;
    // FP32-path destinations (fs_flow::Layer)
    float *d_init_w, *d_init_b, *d_bn0_s, *d_bn0_o, *d_w0, *d_b0, *d_w1, *d_b1, *d_final_w, *d_final_b, *d_ux, *d_uy, *d_ud;
    // tensor-path destinations (TcLayer), null without the tensor path
    float *t_wstream, *t_psets, *t_bn0_s, *t_bn0_o, *t_b0, *t_bfinal, *t_bfused;
    uint16_t* t_wfused;
    // scratch: scale of the second BatchNorm of every block [nB, H], folded bias of the residual stream [H]
    double *sc1, *cfin;
};

struct RepackDims {
    int K, N, H, nB, nb, P, K0, Kp0, n_chunks, chn, NP;
    double bound, eps;
};

__device__ __forceinline__ float tf32_round_dev(float x) {
    uint32_t u = __float_as_uint(x);
    if ((u & 0x7F800000u) == 0x7F800000u) return x;
    u += 0x00001000u;            // round to nearest, ties away (cvt.rna), as tf32_round in flow_tc.cu
    u &= 0xFFFFE000u;
    return __uint_as_float(u);
}
__device__ __forceinline__ float half_round_dev(float x) { return __half2float(__float2half_rn(x)); }

// ---- 1. BatchNorm folding (flow.cu: pack_layer; nn.BatchNorm1d(eps) in eval mode) ----
__global__ void repack_fold_bn(const RepackLayer* __restrict__ Ls, RepackDims D) {
    const RepackLayer L = Ls[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // (block b, channel c)
    if (i >= D.nB * D.H) return;
    const int b = i / D.H, c = i % D.H;
    const size_t o0 = ((size_t)b * 2 + 0) * D.H + c, o1 = ((size_t)b * 2 + 1) * D.H + c;
    const double s0 = (double)L.bn_w[o0] / sqrt((double)L.bn_var[o0] + D.eps);
    const double f0 = (double)L.bn_b[o0] - (double)L.bn_mean[o0] * s0;
    const double s1 = (double)L.bn_w[o1] / sqrt((double)L.bn_var[o1] + D.eps);
    const double f1 = (double)L.bn_b[o1] - (double)L.bn_mean[o1] * s1;
    L.d_bn0_s[i] = (float)s0;
    L.d_bn0_o[i] = (float)f0;
    const float b0 = (float)(s1 * (double)L.lin_b[o0] + f1);      // second BN folded into linear 0
    L.d_b0[i] = b0;
    L.d_b1[i] = L.lin_b[o1];
    L.sc1[i] = s1;
    if (L.t_bn0_s) {
        L.t_bn0_s[i] = (float)s0;
        L.t_b0[i] = b0;
    }
}

// ---- 2. bias folding of the residual stream (flow_tc.cu: the kernel carries u = h - c) + parameter sets ----
__global__ void repack_bias_chain(const RepackLayer* __restrict__ Ls, RepackDims D) {
    const RepackLayer L = Ls[blockIdx.y];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= D.H) return;
    L.d_init_b[k] = L.init_b[k];
    double c = (double)L.init_b[k];
    for (int b = 0; b < D.nB; ++b) {
        const size_t i = (size_t)b * D.H + k;
        const float s0 = L.d_bn0_s[i], o0 = L.d_bn0_o[i];
        const float o0f = (float)((double)o0 + (double)s0 * c);
        if (L.t_psets) {
            L.t_bn0_o[i] = o0f;
            float* ps = L.t_psets + (size_t)b * 3 * D.H;
            ps[D.H + k] = s0;
            ps[2 * D.H + k] = o0f;
            if (b == 0) ps[k] = 0.f;
            L.t_psets[(size_t)(b + 1) * 3 * D.H + k] = L.d_b0[i];
        }
        c += (double)L.lin_b[((size_t)b * 2 + 1) * D.H + k];
    }
    if (L.t_psets) {
        L.t_psets[(size_t)D.nB * 3 * D.H + D.H + k] = 0.f;
        L.t_psets[(size_t)D.nB * 3 * D.H + 2 * D.H + k] = 0.f;
    }
    L.cfin[k] = c;
}

// ---- 3. FP32-path matrices ----
__global__ void repack_fp32_weights(const RepackLayer* __restrict__ Ls, RepackDims D) {
    const RepackLayer L = Ls[blockIdx.y];
    const size_t n_init = (size_t)D.H * D.K0, n_blk = (size_t)D.nB * D.H * D.H, n_fin = (size_t)D.NP * D.H;
    const size_t total = n_init + 2 * n_blk + n_fin + D.NP;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        if (i < n_init) {
            L.d_init_w[i] = L.init_w[i];
        } else if (i < n_init + n_blk) {                           // linear 0 with the second BatchNorm folded in
            const size_t j = i - n_init;
            const size_t b = j / ((size_t)D.H * D.H), r = (j / D.H) % D.H, k = j % D.H;
            L.d_w0[j] = (float)(L.sc1[b * D.H + r] * (double)L.lin_w[((b * 2 + 0) * D.H + r) * D.H + k]);
        } else if (i < n_init + 2 * n_blk) {
            const size_t j = i - n_init - n_blk;
            const size_t b = j / ((size_t)D.H * D.H), rk = j % ((size_t)D.H * D.H);
            L.d_w1[j] = L.lin_w[(b * 2 + 1) * D.H * D.H + rk];
        } else if (i < n_init + 2 * n_blk + n_fin) {               // parameter-major rows: dst k N + j <- src j P + k
            const size_t j = i - n_init - 2 * n_blk;
            const size_t dst = j / D.H, kk = j % D.H;
            const size_t src = (dst % D.N) * D.P + dst / D.N;
            L.d_final_w[j] = L.final_w[src * D.H + kk];
        } else {
            const size_t dst = i - n_init - 2 * n_blk - n_fin;
            L.d_final_b[dst] = L.final_b[(dst % D.N) * D.P + dst / D.N];
        }
    }
}

// ---- 4. tensor path: the weight stream in consumption order (flow_tc.cu: tc_pack), one 32-bit word per thread ----
__global__ void repack_tc_stream(const RepackLayer* __restrict__ Ls, RepackDims D) {
    const RepackLayer L = Ls[blockIdx.y];
    if (!L.t_wstream) return;
    const int H = D.H;
    const size_t wA = (size_t)(D.Kp0 / TC_KB) * H * 32;             // GEMM0: TF32 tiles [H x 32]
    const size_t wG = (size_t)H * H / 2;                            // one H x H GEMM in FP16 = H*H/2 words
    const size_t wB = (size_t)D.nB * 2 * wG;
    const size_t wC = (size_t)D.n_chunks * (H / TC_KB) * 128 * 32;  // theta-path final layer: TF32 tiles [128 x 32]
    uint32_t* out = reinterpret_cast<uint32_t*>(L.t_wstream);
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < wA + wB + wC; w += (size_t)gridDim.x * blockDim.x) {
        if (w < wA) {
            const size_t t = w / ((size_t)H * 32), r = w % ((size_t)H * 32);
            const int i = (int)(r / 32), pc = (int)(r % 32) / 4, e = (int)(r % 4);
            const int k = (int)t * 32 + 4 * (pc ^ (i & 7)) + e;
            out[w] = k < D.K0 ? __float_as_uint(tf32_round_dev(L.init_w[(size_t)i * D.K0 + k])) : 0u;
        } else if (w < wA + wB) {
            const size_t g = (w - wA) / wG;                         // GEMM index: block = g / 2, linear = g % 2
            const size_t h0 = ((w - wA) % wG) * 2;                  // first of the two halves of this word
            const int b = (int)(g / 2), lin = (int)(g % 2);
            int kt, row_off, rows;
            size_t rem;
            if (H == 256) {
                if (h0 < 2 * 256 * 64) { kt = (int)(h0 / (256 * 64)); rem = h0 % (256 * 64); row_off = 0; rows = 256; }
                else {
                    const size_t h2 = h0 - 2 * 256 * 64;
                    const int nh = (int)(h2 / (2 * 128 * 64));
                    const size_t r2 = h2 % (2 * 128 * 64);
                    kt = 2 + (int)(r2 / (128 * 64)); rem = r2 % (128 * 64); row_off = nh * 128; rows = 128;
                }
            } else { kt = (int)(h0 / ((size_t)H * 64)); rem = h0 % ((size_t)H * 64); row_off = 0; rows = H; }
            (void)rows;
            const int i = (int)(rem / 64), pos = (int)(rem % 64), pc = pos / 8, e = pos % 8;   // e even
            const int n = row_off + i, k = kt * 64 + 8 * (pc ^ (i & 7)) + e;
            const float* W = L.lin_w + ((size_t)b * 2 + lin) * H * H + (size_t)n * H + k;
            float v0 = W[0], v1 = W[1];
            if (lin == 0) {
                const double sc = L.sc1[(size_t)b * H + n];
                v0 = (float)(sc * (double)v0);
                v1 = (float)(sc * (double)v1);
            }
            const __half2 hh = __halves2half2(__float2half_rn(v0), __float2half_rn(v1));
            out[w] = *reinterpret_cast<const uint32_t*>(&hh);
        } else {
            const size_t r = w - wA - wB;
            const size_t per_chunk = (size_t)(H / TC_KB) * 128 * 32;
            const int c = (int)(r / per_chunk);
            const size_t rc = r % per_chunk;
            const int tk = (int)(rc / (128 * 32)), i = (int)(rc % (128 * 32)) / 32, pc = (int)(rc % 32) / 4, e = (int)(rc % 4);
            const int n = c * 128 + i, k = tk * 32 + 4 * (pc ^ (i & 7)) + e;
            uint32_t v = 0u;
            if (n < D.NP) {
                const size_t src = (size_t)(n % D.N) * D.P + n / D.N;
                v = __float_as_uint(tf32_round_dev(L.final_w[src * H + k]));
            }
            out[w] = v;
        }
    }
}

__device__ __forceinline__ int fused_row(int cidx, int nb) {       // chunk column -> parameter index (or -1 / -2 = pad)
    if (cidx < 32) return cidx < nb ? cidx : -2;
    if (cidx < 64) return cidx - 32 < nb ? nb + cidx - 32 : -2;
    return cidx - 64 <= nb ? 2 * nb + cidx - 64 : -1;
}

// fused-spline final layer: per coordinate H/64 tiles of [chn x 64] halves
__global__ void repack_tc_fused(const RepackLayer* __restrict__ Ls, RepackDims D) {
    const RepackLayer L = Ls[blockIdx.y];
    if (!L.t_wfused) return;
    const int H = D.H, chn = D.chn;
    const size_t per_tile = (size_t)chn * 32;                        // words
    const size_t total = (size_t)D.N * (H / 64) * per_tile;
    uint32_t* out = reinterpret_cast<uint32_t*>(L.t_wfused);
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const size_t tile = w / per_tile, r = (w % per_tile) * 2;
        const int j = (int)(tile / (H / 64)), kt = (int)(tile % (H / 64));
        const int i = (int)(r / 64), pos = (int)(r % 64), pc = pos / 8, e = pos % 8;
        const int k = kt * 64 + 8 * (pc ^ (i & 7)) + e;
        const int fr = fused_row(i, D.nb);
        uint32_t v = 0u;
        if (fr >= 0) {
            const float* W = L.final_w + ((size_t)j * D.P + fr) * H + k;
            const __half2 hh = __halves2half2(__float2half_rn(W[0]), __float2half_rn(W[1]));
            v = *reinterpret_cast<const uint32_t*>(&hh);
        }
        out[w] = v;
    }
}

// ---- 5. folded final-layer biases: one warp per output row, float64 dot product with the rounded weights ----
__global__ void repack_tc_bias(const RepackLayer* __restrict__ Ls, RepackDims D) {
    const RepackLayer L = Ls[blockIdx.y];
    if (!L.t_bfinal) return;
    const int lane = threadIdx.x & 31;
    const int n_fin = D.n_chunks * 128, n_fus = D.chn ? D.N * D.chn : 0;
    for (int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); o < n_fin + n_fus; o += gridDim.x * (blockDim.x >> 5)) {
        long long src = -1;
        float padv = 0.f;
        bool f16 = false;
        if (o < n_fin) {
            if (o < D.NP) src = (long long)(o % D.N) * D.P + o / D.N;
        } else {
            const int j = (o - n_fin) / D.chn, cidx = (o - n_fin) % D.chn;
            const int fr = fused_row(cidx, D.nb);
            f16 = true;
            if (fr >= 0) src = (long long)j * D.P + fr;
            else if (fr == -2) padv = -3.0e38f;
        }
        double acc = 0.0;
        if (src >= 0) {
            const float* W = L.final_w + (size_t)src * D.H;
            for (int k = lane; k < D.H; k += 32)
                acc += (double)(f16 ? half_round_dev(W[k]) : tf32_round_dev(W[k])) * L.cfin[k];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
            acc += (double)L.final_b[src];
        }
        if (lane == 0) {
            const float v = src >= 0 ? (float)acc : padv;
            if (o < n_fin) L.t_bfinal[o] = v;
            else L.t_bfused[o - n_fin] = v;
        }
    }
}

// ---- 6. unconditional-spline knot tables (flow.cu: host_knots; utils/splines.py:117-129 in float64), knot-major ----
__global__ void repack_knots(const RepackLayer* __restrict__ Ls, RepackDims D) {
    const RepackLayer L = Ls[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = D.nb, N = D.N;
    if (t < 2 * N) {
        const int j = t % N;
        const float* un = (t < N ? L.un_w : L.un_h) + (size_t)j * nb;
        float* dst = t < N ? L.d_ux : L.d_uy;
        double m = un[0];
        for (int i = 1; i < nb; ++i) m = un[i] > m ? (double)un[i] : m;
        double sum = 0;
        for (int i = 0; i < nb; ++i) sum += exp((double)un[i] - m);
        double c = 0;
        dst[j] = (float)(-D.bound);
        for (int i = 0; i < nb; ++i) {
            c += 1e-3 + (1 - 1e-3 * nb) * (exp((double)un[i] - m) / sum);
            dst[(size_t)(i + 1) * N + j] = (float)(i + 1 == nb ? D.bound : 2 * D.bound * c - D.bound);
        }
    } else if (t < 2 * N + N * (nb + 1)) {
        const int q = t - 2 * N, j = q / (nb + 1), i = q % (nb + 1);
        const double x = L.un_d[(size_t)j * (nb + 1) + i];
        L.d_ud[(size_t)i * N + j] = (float)(1e-3 + (x > 20 ? x : log1p(exp(x))));
    }
}

}  // namespace fs

using namespace fs;

// allocates the per-flow scratch of fs_flow_update on first use
static int repack_prepare(fs_flow* f) {
    if (f->repack_tab) return FS_OK;
    void* tab = nullptr;
    FS_CUDA(cudaMalloc(&tab, sizeof(RepackLayer) * f->K));
    f->allocs.push_back(tab);
    void* sc = nullptr;
    FS_CUDA(cudaMalloc(&sc, sizeof(double) * (size_t)f->K * ((size_t)f->n_blocks * f->H + f->H) + 16));
    f->allocs.push_back(sc);
    f->repack_tab = tab;
    f->repack_scratch = (double*)sc;
    return FS_OK;
}

extern "C" int fs_flow_update(fs_flow* f, const fs_flow_desc* d, void* stream) {
    if (!f || !d || !d->layers || d->K != f->K || d->N != f->N || d->H != f->H || d->n_blocks != f->n_blocks ||
        d->nb != f->nb || d->bound != f->bound) {
        set_error("fs_flow_update: the descriptor does not match the packed flow");
        return FS_ERR_INVALID;
    }
    if (int r = repack_prepare(f)) return r;
    cudaStream_t s = (cudaStream_t)stream;
    TcPack* P = (TcPack*)f->tc;
    std::vector<RepackLayer> tab(f->K);
    const size_t per = (size_t)f->n_blocks * f->H + f->H;
    for (int i = 0; i < f->K; ++i) {
        const fs_layer_params& p = d->layers[i];
        if (!p.init_w || !p.init_b || !p.final_w || !p.final_b || !p.un_w || !p.un_h || !p.un_d ||
            (f->n_blocks && (!p.bn_w || !p.bn_b || !p.bn_mean || !p.bn_var || !p.lin_w || !p.lin_b))) {
            set_error("fs_flow_update: null parameter pointer in layer %d", i);
            return FS_ERR_INVALID;
        }
        RepackLayer& R = tab[i];
        R.init_w = p.init_w; R.init_b = p.init_b; R.bn_w = p.bn_w; R.bn_b = p.bn_b; R.bn_mean = p.bn_mean;
        R.bn_var = p.bn_var; R.lin_w = p.lin_w; R.lin_b = p.lin_b; R.final_w = p.final_w; R.final_b = p.final_b;
        R.un_w = p.un_w; R.un_h = p.un_h; R.un_d = p.un_d;
        const fs_flow::Layer& L = f->layers[i];
        R.d_init_w = L.init_w; R.d_init_b = L.init_b; R.d_bn0_s = L.bn0_s; R.d_bn0_o = L.bn0_o; R.d_w0 = L.w0;
        R.d_b0 = L.b0; R.d_w1 = L.w1; R.d_b1 = L.b1; R.d_final_w = L.final_w; R.d_final_b = L.final_b;
        R.d_ux = L.u_x; R.d_uy = L.u_y; R.d_ud = L.u_d;
        if (P) {
            const TcLayer& T = P->layers[i];
            R.t_wstream = T.wstream; R.t_psets = T.psets; R.t_bn0_s = T.bn0_s; R.t_bn0_o = T.bn0_o; R.t_b0 = T.b0;
            R.t_bfinal = T.b_final; R.t_bfused = T.b_fused; R.t_wfused = (uint16_t*)T.wfused;
        } else {
            R.t_wstream = R.t_psets = R.t_bn0_s = R.t_bn0_o = R.t_b0 = R.t_bfinal = R.t_bfused = nullptr;
            R.t_wfused = nullptr;
        }
        R.sc1 = f->repack_scratch + (size_t)i * per;
        R.cfin = R.sc1 + (size_t)f->n_blocks * f->H;
    }
    // the pointer table is tiny (K x 35 pointers); staged through the stream so the update stays ordered with the passes
    FS_CUDA(cudaMemcpyAsync(f->repack_tab, tab.data(), sizeof(RepackLayer) * f->K, cudaMemcpyHostToDevice, s));
    RepackDims D;
    D.K = f->K; D.N = f->N; D.H = f->H; D.nB = f->n_blocks; D.nb = f->nb; D.P = f->P; D.K0 = 2 * f->N;
    D.NP = f->N * f->P;
    D.Kp0 = P ? P->Kp0 : 0; D.n_chunks = P ? P->n_chunks : 0; D.chn = P ? P->chn : 0;
    D.bound = f->bound; D.eps = (double)d->bn_eps;
    const RepackLayer* T = (const RepackLayer*)f->repack_tab;
    const unsigned K = (unsigned)f->K;
    if (f->n_blocks) {
        repack_fold_bn<<<dim3((f->n_blocks * f->H + 255) / 256, K), 256, 0, s>>>(T, D);
        count_launch();
    }
    repack_bias_chain<<<dim3((f->H + 127) / 128, K), 128, 0, s>>>(T, D);
    repack_fp32_weights<<<dim3(256, K), 256, 0, s>>>(T, D);
    repack_knots<<<dim3((2 * f->N + f->N * (f->nb + 1) + 127) / 128, K), 128, 0, s>>>(T, D);
    count_launch(3);
    if (P) {
        repack_tc_stream<<<dim3(512, K), 256, 0, s>>>(T, D);
        repack_tc_bias<<<dim3(256, K), 256, 0, s>>>(T, D);
        count_launch(2);
        if (P->chn) {
            repack_tc_fused<<<dim3(256, K), 256, 0, s>>>(T, D);
            count_launch();
        }
    }
    return cuda_check(cudaGetLastError(), "fs_flow_update");
}
