// Total and single-particle pair energy for B independent configurations.
//
// fs_energy_total   <- EnergyCalculator.calculate_total_energy_virial (MCMC/energy_calculator.py:121-203)
// fs_energy_particle<- EnergyCalculator.calculate_particle_energy_virial (:48-108)
//
// Layout: a group of G threads (32..256, power of two) owns one configuration;
// its 2N floats are staged once into shared memory (stored twice, so the cyclic
// neighbour index i+k never needs a modulo).  Pairs are enumerated cyclically:
// particle i meets i+1 .. i+floor((N-1)/2), plus i+N/2 for the first half when N
// is even - every unordered pair exactly once and every thread the same trip
// count.  FP32 on the CUDA cores; cross-thread sums are finished in FP64.
#include <stdlib.h>

#include "common.cuh"

namespace fs {

// Pair accumulation of the total-energy kernel.  Instead of the energy and the virial, the three sums
//   A12 = sum r^-12,  A6 = sum r^-6,  cnt = #pairs   over pairs with r <= r_cut
// are kept: E = 4 (A12 - A6) - cnt e_cut, W = 48 A12 - 24 A6 (potential.py:11-27) - three predicated
// instructions per pair for both observables.  FASTWRAP: every coordinate of the configuration lies in
// [0, L], so |d| <= L and the minimum image is one conditional shift by +-L (same result as
// d - L rint(d / L), including the tie |d| = L/2 which both leave unshifted).
template <bool FASTWRAP>
__device__ __forceinline__ float wrap1(float d, float L, float halfL, float invL) {
    if (FASTWRAP) return fabsf(d) > halfL ? d - copysignf(L, d) : d;
    return min_image(d, L, invL);
}

template <bool FASTWRAP>
__device__ __forceinline__ void pair_sums(float dx, float dy, const PotDev& P, float hx, float hy, float& a12,
                                          float& a6, float& cnt, float& r2min) {
    dx = wrap1<FASTWRAP>(dx, P.Lx, hx, P.inv_Lx);
    dy = wrap1<FASTWRAP>(dy, P.Ly, hy, P.inv_Ly);
    const float r2 = __fmaf_rn(dy, dy, dx * dx);
    r2min = fminf(r2min, r2);
    float inv;                                       // single MUFU.RCP (1 ulp; r^2 is never denormal here)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(r2));
    const float s6 = inv * inv * inv;
    if (r2 <= P.rc2) {
        a12 = __fmaf_rn(s6, s6, a12);
        a6 += s6;
        cnt += 1.0f;
    }
}

template <int G, bool FASTWRAP>
__device__ __forceinline__ void total_pairs(const float2* sp, int N, int t, const PotDev& P, float& a12, float& a6,
                                            float& cnt, float& r2min, float& ew) {
    const float hx = 0.5f * P.Lx, hy = 0.5f * P.Ly;
    const int half = (N - 1) / 2;
    for (int i = t; i < N; i += G) {
        const float2 pi = sp[i];
        float b12 = 0.f, b6 = 0.f, bc = 0.f, c12 = 0.f, c6 = 0.f, cc = 0.f;
        int k = 1;
        for (; k + 1 <= half; k += 2) {
            const float2 a = sp[i + k], c = sp[i + k + 1];
            pair_sums<FASTWRAP>(pi.x - a.x, pi.y - a.y, P, hx, hy, b12, b6, bc, r2min);
            pair_sums<FASTWRAP>(pi.x - c.x, pi.y - c.y, P, hx, hy, c12, c6, cc, r2min);
        }
        if (k <= half) {
            const float2 a = sp[i + k];
            pair_sums<FASTWRAP>(pi.x - a.x, pi.y - a.y, P, hx, hy, b12, b6, bc, r2min);
        }
        if ((N & 1) == 0 && i < N / 2) {
            const float2 a = sp[i + N / 2];
            pair_sums<FASTWRAP>(pi.x - a.x, pi.y - a.y, P, hx, hy, c12, c6, cc, r2min);
        }
        a12 += b12 + c12;
        a6 += b6 + c6;
        cnt += bc + cc;
        ew += wells(pi.x, pi.y, P);
    }
}

template <int G>
__global__ void __launch_bounds__(256) energy_total_kernel(const float* __restrict__ pos, int B, int N,
                                                           PotDev P, float* __restrict__ E,
                                                           float* __restrict__ W,
                                                           unsigned char* __restrict__ overlap) {
    extern __shared__ float2 smem[];
    constexpr int GROUPS = 256 / G;
    const int g = threadIdx.x / G;
    const int t = threadIdx.x % G;
    const int b = blockIdx.x * GROUPS + g;
    float2* sp = smem + (size_t)g * 2 * N;
    __shared__ double red_e[8], red_w[8];
    __shared__ float red_m[8];

    const bool live = b < B;
    bool inbox = true;
    if (live) {
        const float2* src = reinterpret_cast<const float2*>(pos) + (size_t)b * N;
        for (int i = t; i < N; i += G) {
            const float2 v = __ldg(src + i);
            sp[i] = v;
            sp[N + i] = v;
            inbox = inbox && v.x >= 0.f && v.x <= P.Lx && v.y >= 0.f && v.y <= P.Ly;
        }
    }
    const int all_in = __syncthreads_and(inbox ? 1 : 0);     // also orders the smem writes before the reads

    float a12 = 0.f, a6 = 0.f, cnt = 0.f, ew = 0.f, r2min = 3.0e38f;
    if (live) {
        if (all_in) total_pairs<G, true>(sp, N, t, P, a12, a6, cnt, r2min, ew);
        else total_pairs<G, false>(sp, N, t, P, a12, a6, cnt, r2min, ew);
    }
    // E = 4 (A12 - A6) - cnt e_cut + wells, W = 48 A12 - 24 A6; cross-thread sums in float64
    double de = 4.0 * ((double)a12 - (double)a6) - (double)cnt * (double)P.e_cut + (double)ew;
    double dw = 48.0 * (double)a12 - 24.0 * (double)a6;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        de += __shfl_xor_sync(0xffffffffu, de, o);
        dw += __shfl_xor_sync(0xffffffffu, dw, o);
    }
    r2min = warp_min(r2min);
    if (G > 32) {
        const int wid = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) {
            red_e[wid] = de;
            red_w[wid] = dw;
            red_m[wid] = r2min;
        }
        __syncthreads();
        if (t == 0) {
            constexpr int WPG = G / 32;
            de = 0; dw = 0; r2min = 3.0e38f;
            for (int i = 0; i < WPG; ++i) {
                de += red_e[g * WPG + i];
                dw += red_w[g * WPG + i];
                r2min = fminf(r2min, red_m[g * WPG + i]);
            }
        }
    }
    if (live && t == 0) {
        const bool ov = r2min < P.rcore2;
        const float inf = __int_as_float(0x7f800000);
        E[b] = ov ? inf : (float)de;
        W[b] = ov ? inf : (float)dw;
        if (overlap) overlap[b] = ov ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------
// Packed variant (shared-memory tile <= 112 KB, i.e. N <= 3584 at one configuration per block): the same enumeration, two pairs per step in Blackwell's packed FP32 pairs
// (add/mul/fma.f32x2 -> FADD2 / FMUL2 / FFMA2).  Shared memory holds the configuration as separate x and y
// arrays, each stored cyclically (2N entries) and once more shifted by one element, so the coordinates of two
// consecutive partners (x[j], x[j+1]) are ONE aligned 64-bit load whatever the parity of j.  Minimum image for
// both pairs and both axes with d - L rint(d / L) (rint by the 1.5 * 2^23 trick, round-half-even like np.round);
// the cut-off mask is a 0/1 float multiplied into r^-6.  ~12 instructions per pair instead of ~19.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pk2f(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2f(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2f(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long add2f(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long sub2f(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long mul2f(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

struct Pk2Consts {
    unsigned long long invLx, invLy, nLx, nLy, magic, nmagic;
};

// two pairs: X = (dx_a, dx_c), Y = (dy_a, dy_c) before the minimum image.  The FP32 pipe is what bounds the kernel
// (ncu: 68 % busy, every packed instruction holds it for two cycles), so whatever can run elsewhere does: the cut-off
// is a select on r^-6 and the in-range count an integer add (ALU pipe) instead of a 0/1 factor and a packed add -
// 14 packed instructions (7 per axis pair) per two pairs instead of 16.
__device__ __forceinline__ void pair_sums_x2(unsigned long long X, unsigned long long Y, const Pk2Consts& C, float rc2,
                                             unsigned long long& a12, unsigned long long& a6, int& cnt, float& r2min) {
    unsigned long long t = add2f(add2f(mul2f(X, C.invLx), C.magic), C.nmagic);
    X = fma2f(t, C.nLx, X);
    t = add2f(add2f(mul2f(Y, C.invLy), C.magic), C.nmagic);
    Y = fma2f(t, C.nLy, Y);
    const unsigned long long r2 = fma2f(Y, Y, mul2f(X, X));
    float ra, rc;
    upk2f(r2, ra, rc);
    r2min = fminf(r2min, fminf(ra, rc));
    float ia, ic;                                    // MUFU.RCP (1 ulp; r^2 is never denormal here)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ia) : "f"(ra));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ic) : "f"(rc));
    const unsigned long long inv = pk2f(ia, ic);
    const unsigned long long s6 = mul2f(mul2f(inv, inv), inv);
    float sa, sc, ma, mc;
    upk2f(s6, sa, sc);
    // select + predicated integer increments (written out: the compiler's own form is an add and a predicated move
    // per pair, and the move runs on the FP32 pipe)
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.le.f32 p, %5, %7;\n\t"
        "setp.le.f32 q, %6, %7;\n\t"
        "selp.f32 %0, %3, 0f00000000, p;\n\t"
        "selp.f32 %1, %4, 0f00000000, q;\n\t"
        "@p add.s32 %2, %2, 1;\n\t"
        "@q add.s32 %2, %2, 1;\n\t}"
        : "=f"(ma), "=f"(mc), "+r"(cnt)
        : "f"(sa), "f"(sc), "f"(ra), "f"(rc), "f"(rc2));
    const unsigned long long s6m = pk2f(ma, mc);
    a12 = fma2f(s6m, s6, a12);
    a6 = add2f(a6, s6m);
}

template <int G>
__global__ void __launch_bounds__(256) energy_total_kernel_v2(const float* __restrict__ pos, int B, int N, PotDev P,
                                                              float* __restrict__ E, float* __restrict__ W,
                                                              unsigned char* __restrict__ overlap) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int GROUPS = 256 / G;
    const int g = threadIdx.x / G;
    const int t = threadIdx.x % G;
    const int b = blockIdx.x * GROUPS + g;
    // per configuration: X0[2N] Y0[2N] X1[2N] Y1[2N]; X1[m] = x[m + 1] (cyclic)
    float* X0 = smem_f + (size_t)g * 8 * N;
    float* Y0 = X0 + 2 * N;
    float* X1 = Y0 + 2 * N;
    float* Y1 = X1 + 2 * N;
    __shared__ double red_e[8], red_w[8];
    __shared__ float red_m[8];

    const bool live = b < B;
    if (live) {
        const float2* src = reinterpret_cast<const float2*>(pos) + (size_t)b * N;
        for (int i = t; i < N; i += G) {
            const float2 v = __ldg(src + i);
            X0[i] = v.x; X0[N + i] = v.x;
            Y0[i] = v.y; Y0[N + i] = v.y;
            const int m = (i == 0) ? N - 1 : i - 1;          // X1[m] = x[m + 1]
            X1[m] = v.x; X1[N + m] = v.x;
            Y1[m] = v.y; Y1[N + m] = v.y;
        }
    }
    __syncthreads();

    float a12 = 0.f, a6 = 0.f, cnt = 0.f, ew = 0.f, r2min = 3.0e38f;
    if (live) {
        Pk2Consts C;
        C.invLx = pk2f(P.inv_Lx, P.inv_Lx); C.invLy = pk2f(P.inv_Ly, P.inv_Ly);
        C.nLx = pk2f(-P.Lx, -P.Lx); C.nLy = pk2f(-P.Ly, -P.Ly);
        C.magic = pk2f(12582912.0f, 12582912.0f); C.nmagic = pk2f(-12582912.0f, -12582912.0f);
        const float hx = 0.5f * P.Lx, hy = 0.5f * P.Ly;
        const int half = (N - 1) / 2;
        for (int i = t; i < N; i += G) {
            const float pix = X0[i], piy = Y0[i];
            const unsigned long long PX = pk2f(pix, pix), PY = pk2f(piy, piy);
            unsigned long long b12 = 0ull, b6 = 0ull, c12 = 0ull, c6 = 0ull;   // +0.0f pairs
            int bc = 0, cc = 0;
            // partners j = i+1 .. i+half; (x[j], x[j+1]) is an aligned pair in X0 for even j, in X1 - 1 for odd j
            const bool odd = ((i + 1) & 1) != 0;
            const float* bx = odd ? X1 - 1 : X0;
            const float* by = odd ? Y1 - 1 : Y0;
            int k = 1;
            for (; k + 3 <= half; k += 4) {
                const unsigned long long xa = *reinterpret_cast<const unsigned long long*>(bx + i + k);
                const unsigned long long ya = *reinterpret_cast<const unsigned long long*>(by + i + k);
                const unsigned long long xc = *reinterpret_cast<const unsigned long long*>(bx + i + k + 2);
                const unsigned long long yc = *reinterpret_cast<const unsigned long long*>(by + i + k + 2);
                pair_sums_x2(sub2f(PX, xa), sub2f(PY, ya), C, P.rc2, b12, b6, bc, r2min);
                pair_sums_x2(sub2f(PX, xc), sub2f(PY, yc), C, P.rc2, c12, c6, cc, r2min);
            }
            for (; k + 1 <= half; k += 2) {
                const unsigned long long xa = *reinterpret_cast<const unsigned long long*>(bx + i + k);
                const unsigned long long ya = *reinterpret_cast<const unsigned long long*>(by + i + k);
                pair_sums_x2(sub2f(PX, xa), sub2f(PY, ya), C, P.rc2, b12, b6, bc, r2min);
            }
            float s12, s6, sc, u, v;
            upk2f(add2f(b12, c12), u, v); s12 = u + v;
            upk2f(add2f(b6, c6), u, v); s6 = u + v;
            sc = (float)(bc + cc);
            if (k <= half)
                pair_sums<false>(pix - X0[i + k], piy - Y0[i + k], P, hx, hy, s12, s6, sc, r2min);
            if ((N & 1) == 0 && i < N / 2)
                pair_sums<false>(pix - X0[i + N / 2], piy - Y0[i + N / 2], P, hx, hy, s12, s6, sc, r2min);
            a12 += s12;
            a6 += s6;
            cnt += sc;
            ew += wells(pix, piy, P);
        }
    }
    double de = 4.0 * ((double)a12 - (double)a6) - (double)cnt * (double)P.e_cut + (double)ew;
    double dw = 48.0 * (double)a12 - 24.0 * (double)a6;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        de += __shfl_xor_sync(0xffffffffu, de, o);
        dw += __shfl_xor_sync(0xffffffffu, dw, o);
    }
    r2min = warp_min(r2min);
    if (G > 32) {
        const int wid = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) {
            red_e[wid] = de;
            red_w[wid] = dw;
            red_m[wid] = r2min;
        }
        __syncthreads();
        if (t == 0) {
            constexpr int WPG = G / 32;
            de = 0; dw = 0; r2min = 3.0e38f;
            for (int i = 0; i < WPG; ++i) {
                de += red_e[g * WPG + i];
                dw += red_w[g * WPG + i];
                r2min = fminf(r2min, red_m[g * WPG + i]);
            }
        }
    }
    if (live && t == 0) {
        const bool ov = r2min < P.rcore2;
        const float inf = __int_as_float(0x7f800000);
        E[b] = ov ? inf : (float)de;
        W[b] = ov ? inf : (float)dw;
        if (overlap) overlap[b] = ov ? 1 : 0;
    }
}

// One warp per configuration: energy of particle idx[b] against all others.
__global__ void __launch_bounds__(256) energy_particle_kernel(const float* __restrict__ pos,
                                                              const int* __restrict__ idx,
                                                              const float* __restrict__ new_xy, int B, int N,
                                                              PotDev P, float* __restrict__ e_out,
                                                              float* __restrict__ w_out,
                                                              unsigned char* __restrict__ overlap) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const float2* src = reinterpret_cast<const float2*>(pos) + (size_t)b * N;
    const int p = idx[b];
    float2 pp = new_xy ? make_float2(new_xy[2 * b], new_xy[2 * b + 1]) : __ldg(src + p);
    float e = 0.f, w = 0.f, r2min = 3.0e38f;
    for (int j = lane; j < N; j += 32) {
        if (j == p) continue;
        float2 q = __ldg(src + j);
        pair_accum(pp.x - q.x, pp.y - q.y, P, e, w, r2min);
    }
    double de = e, dw = w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        de += __shfl_xor_sync(0xffffffffu, de, o);
        dw += __shfl_xor_sync(0xffffffffu, dw, o);
    }
    r2min = warp_min(r2min);
    if (lane == 0) {
        const bool ov = r2min < P.rcore2;
        const float inf = __int_as_float(0x7f800000);
        de += (double)wells(pp.x, pp.y, P);
        e_out[b] = ov ? inf : (float)de;
        w_out[b] = ov ? inf : (float)dw;
        if (overlap) overlap[b] = ov ? 1 : 0;
    }
}

template <int G>
static int launch_total(const float* pos, int B, int N, const PotDev& P, float* E, float* W,
                        unsigned char* ov, cudaStream_t s) {
    constexpr int GROUPS = 256 / G;
    int grid = (B + GROUPS - 1) / GROUPS;
    const size_t smem2 = (size_t)GROUPS * 8 * N * sizeof(float);
    static long v2max = -1;                              // tuning knob: largest tile (bytes) of the packed variant
    if (v2max < 0) { const char* e = getenv("FS_ENERGY_V2MAX"); v2max = e ? atol(e) : 112 * 1024; }
    if ((long)smem2 <= v2max) {                          // packed variant while two blocks per SM still fit
        if (smem2 > 48 * 1024)
            FS_CUDA(cudaFuncSetAttribute(energy_total_kernel_v2<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        energy_total_kernel_v2<G><<<grid, 256, smem2, s>>>(pos, B, N, P, E, W, ov);
        fs::count_launch();
        return cuda_check(cudaGetLastError(), "energy_total_kernel_v2");
    }
    size_t smem = (size_t)GROUPS * 2 * N * sizeof(float2);
    if (smem > 48 * 1024)
        FS_CUDA(cudaFuncSetAttribute(energy_total_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    energy_total_kernel<G><<<grid, 256, smem, s>>>(pos, B, N, P, E, W, ov);
    fs::count_launch();
    return cuda_check(cudaGetLastError(), "energy_total_kernel");
}

}  // namespace fs

extern "C" int fs_energy_total(const float* pos, int B, int N, float Lx, float Ly, const fs_pot* pot,
                               float* E, float* W, unsigned char* overlap, void* stream) {
    if (!pos || !pot || !E || !W || B < 0 || N < 1 || !(Lx > 0) || !(Ly > 0)) {
        fs::set_error("fs_energy_total: invalid argument");
        return FS_ERR_INVALID;
    }
    if (N > 12288) {
        fs::set_error("fs_energy_total: N=%d exceeds the shared-memory tile (max 12288)", N);
        return FS_ERR_UNSUPPORTED;
    }
    if (B == 0) return FS_OK;
    fs::PotDev P = fs::make_pot(pot, Lx, Ly);
    cudaStream_t s = (cudaStream_t)stream;
    // group size
    static int g_forced = -1;                            // tuning knob (32 / 64 / 128 / 256)
    if (g_forced < 0) { const char* e = getenv("FS_ENERGY_G"); g_forced = e ? atoi(e) : 0; }
    // (measured, scripts/energy_sweep.py with FS_ENERGY_G: about four particles per thread is best - several
    // configurations per block hide each other's load / reduction phases, and the smaller groups need fewer registers)
    int G = N <= 160 ? 32 : (N <= 320 ? 64 : (N <= 768 ? 128 : 256));
    if (g_forced) G = g_forced;
    if (G == 32) return fs::launch_total<32>(pos, B, N, P, E, W, overlap, s);
    if (G == 64) return fs::launch_total<64>(pos, B, N, P, E, W, overlap, s);
    if (G == 128) return fs::launch_total<128>(pos, B, N, P, E, W, overlap, s);
    return fs::launch_total<256>(pos, B, N, P, E, W, overlap, s);
}

extern "C" int fs_energy_particle(const float* pos, const int* idx, const float* new_xy, int B, int N,
                                  float Lx, float Ly, const fs_pot* pot, float* e, float* w,
                                  unsigned char* overlap, void* stream) {
    if (!pos || !idx || !pot || !e || !w || B < 0 || N < 1 || !(Lx > 0) || !(Ly > 0)) {
        fs::set_error("fs_energy_particle: invalid argument");
        return FS_ERR_INVALID;
    }
    if (B == 0) return FS_OK;
    fs::PotDev P = fs::make_pot(pot, Lx, Ly);
    int grid = (B + 7) / 8;
    fs::energy_particle_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pos, idx, new_xy, B, N, P, e, w, overlap);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "energy_particle_kernel");
}
