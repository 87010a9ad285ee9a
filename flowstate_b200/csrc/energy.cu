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

#include "accept.cuh"

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

// FOLD variant of pair_sums_x2 for configurations that lie in the box (every coordinate in [0, L], checked while the
// tile is staged): |d| <= L, so the minimum-image magnitude is min(|d|, L - |d|) - the same value d - L rint(d / L)
// has (one rounding of L - |d| either way; both keep |d| at the tie |d| = L/2) - in one FP32 add and one min (ALU
// pipe, |.| as operand modifiers) per component instead of four packed FP32 operations per axis pair.  The in-range
// sums are predicated scalar operations (no select): per two pairs 10 FP32-pipe instructions (20 pipe cycles, the
// rint form holds the pipe for 28) and 9 ALU instructions.
__device__ __forceinline__ void pair_sums_x2_fold(unsigned long long X, unsigned long long Y, float Lx, float Ly, float rc2,
                                                  float& a12a, float& a12c, float& a6a, float& a6c, int& cnt,
                                                  float& r2min) {
    float xa, xc, ya, yc;
    upk2f(X, xa, xc);
    upk2f(Y, ya, yc);
    xa = fminf(fabsf(xa), Lx - fabsf(xa));
    xc = fminf(fabsf(xc), Lx - fabsf(xc));
    ya = fminf(fabsf(ya), Ly - fabsf(ya));
    yc = fminf(fabsf(yc), Ly - fabsf(yc));
    const unsigned long long Xm = pk2f(xa, xc), Ym = pk2f(ya, yc);
    const unsigned long long r2 = fma2f(Ym, Ym, mul2f(Xm, Xm));
    float ra, rc;
    upk2f(r2, ra, rc);
    r2min = fminf(r2min, fminf(ra, rc));
    float ia, ic;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ia) : "f"(ra));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ic) : "f"(rc));
    const unsigned long long inv = pk2f(ia, ic);
    const unsigned long long s6 = mul2f(mul2f(inv, inv), inv);
    float sa, sc;
    upk2f(s6, sa, sc);
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.le.f32 p, %7, %9;\n\t"
        "setp.le.f32 q, %8, %9;\n\t"
        "@p fma.rn.f32 %0, %5, %5, %0;\n\t"
        "@q fma.rn.f32 %1, %6, %6, %1;\n\t"
        "@p add.rn.f32 %2, %2, %5;\n\t"
        "@q add.rn.f32 %3, %3, %6;\n\t"
        "@p add.s32 %4, %4, 1;\n\t"
        "@q add.s32 %4, %4, 1;\n\t}"
        : "+f"(a12a), "+f"(a12c), "+f"(a6a), "+f"(a6c), "+r"(cnt)
        : "f"(sa), "f"(sc), "f"(ra), "f"(rc), "f"(rc2));
}

// The pair walk of one thread of energy_total_kernel_v2 (particles t, t + G, ...).
template <int G, bool FOLD>
__device__ __forceinline__ void total_pairs_v2(const float* X0, const float* Y0, const float* X1, const float* Y1, int N,
                                               int t, const PotDev& P, float& a12, float& a6, float& cnt, float& ew,
                                               float& r2min) {
    Pk2Consts C;
    C.invLx = pk2f(P.inv_Lx, P.inv_Lx); C.invLy = pk2f(P.inv_Ly, P.inv_Ly);
    C.nLx = pk2f(-P.Lx, -P.Lx); C.nLy = pk2f(-P.Ly, -P.Ly);
    C.magic = pk2f(12582912.0f, 12582912.0f); C.nmagic = pk2f(-12582912.0f, -12582912.0f);
    const float hx = 0.5f * P.Lx, hy = 0.5f * P.Ly;
    const int half = (N - 1) / 2;
    for (int i = t; i < N; i += G) {
        const float pix = X0[i], piy = Y0[i];
        const unsigned long long PX = pk2f(pix, pix), PY = pk2f(piy, piy);
        // partners j = i+1 .. i+half; (x[j], x[j+1]) is an aligned pair in X0 for even j, in X1 - 1 for odd j
        const bool odd = ((i + 1) & 1) != 0;
        const float* bx = odd ? X1 - 1 : X0;
        const float* by = odd ? Y1 - 1 : Y0;
        float s12, s6, sc;
        int k = 1;
        if (FOLD) {
            float f12[4] = {0.f, 0.f, 0.f, 0.f}, f6[4] = {0.f, 0.f, 0.f, 0.f};
            int bc = 0, cc = 0;
            for (; k + 3 <= half; k += 4) {
                const unsigned long long xa = *reinterpret_cast<const unsigned long long*>(bx + i + k);
                const unsigned long long ya = *reinterpret_cast<const unsigned long long*>(by + i + k);
                const unsigned long long xc = *reinterpret_cast<const unsigned long long*>(bx + i + k + 2);
                const unsigned long long yc = *reinterpret_cast<const unsigned long long*>(by + i + k + 2);
                pair_sums_x2_fold(sub2f(PX, xa), sub2f(PY, ya), P.Lx, P.Ly, P.rc2, f12[0], f12[1], f6[0], f6[1], bc, r2min);
                pair_sums_x2_fold(sub2f(PX, xc), sub2f(PY, yc), P.Lx, P.Ly, P.rc2, f12[2], f12[3], f6[2], f6[3], cc, r2min);
            }
            for (; k + 1 <= half; k += 2) {
                const unsigned long long xa = *reinterpret_cast<const unsigned long long*>(bx + i + k);
                const unsigned long long ya = *reinterpret_cast<const unsigned long long*>(by + i + k);
                pair_sums_x2_fold(sub2f(PX, xa), sub2f(PY, ya), P.Lx, P.Ly, P.rc2, f12[0], f12[1], f6[0], f6[1], bc, r2min);
            }
            s12 = (f12[0] + f12[2]) + (f12[1] + f12[3]);
            s6 = (f6[0] + f6[2]) + (f6[1] + f6[3]);
            sc = (float)(bc + cc);
        } else {
            unsigned long long b12 = 0ull, b6 = 0ull, c12 = 0ull, c6 = 0ull;   // +0.0f pairs
            int bc = 0, cc = 0;
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
            float u, v;
            upk2f(add2f(b12, c12), u, v); s12 = u + v;
            upk2f(add2f(b6, c6), u, v); s6 = u + v;
            sc = (float)(bc + cc);
        }
        if (k <= half)
            pair_sums<FOLD>(pix - X0[i + k], piy - Y0[i + k], P, hx, hy, s12, s6, sc, r2min);
        if ((N & 1) == 0 && i < N / 2)
            pair_sums<FOLD>(pix - X0[i + N / 2], piy - Y0[i + N / 2], P, hx, hy, s12, s6, sc, r2min);
        a12 += s12;
        a6 += s6;
        cnt += sc;
        ew += wells(pix, piy, P);
    }
}

// ACCEPT: the fused global move (fs_accept_global_fused) - `pos` holds the proposals; after the energy of a proposal
// is known, thread 0 of its group applies the acceptance rule (accept.cuh) and the group copies an accepted proposal
// from the shared-memory tile over the chain's state `state` - the proposal is read from HBM once and its energy never
// leaves the chip before the decision.
template <int G, bool NOFOLD, bool ACCEPT, int BS>
__global__ void __launch_bounds__(BS) energy_total_kernel_v2(const float* __restrict__ pos, int B, int N, PotDev P,
                                                              float* __restrict__ E, float* __restrict__ W,
                                                              unsigned char* __restrict__ overlap,
                                                              float* __restrict__ state, AcceptArgs A) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int GROUPS = BS / G;
    const int g = threadIdx.x / G;
    const int t = threadIdx.x % G;
    const int b = blockIdx.x * GROUPS + g;
    // per configuration: X0[2N] Y0[2N] X1[2N] Y1[2N]; X1[m] = x[m + 1] (cyclic)
    float* X0 = smem_f + (size_t)g * 8 * N;
    float* Y0 = X0 + 2 * N;
    float* X1 = Y0 + 2 * N;
    float* Y1 = X1 + 2 * N;
    __shared__ double red_e[BS / 32], red_w[BS / 32];
    __shared__ float red_m[BS / 32];

    const bool live = b < B;
    bool inbox = true;
    if (live) {
        const float2* src = reinterpret_cast<const float2*>(pos) + (size_t)b * N;
        for (int i = t; i < N; i += G) {
            const float2 v = __ldg(src + i);
            X0[i] = v.x; X0[N + i] = v.x;
            Y0[i] = v.y; Y0[N + i] = v.y;
            const int m = (i == 0) ? N - 1 : i - 1;          // X1[m] = x[m + 1]
            X1[m] = v.x; X1[N + m] = v.x;
            Y1[m] = v.y; Y1[N + m] = v.y;
            inbox = inbox && v.x >= 0.f && v.x <= P.Lx && v.y >= 0.f && v.y <= P.Ly;
        }
    }
    const int all_in = __syncthreads_and(inbox ? 1 : 0);     // block-uniform; also orders the smem writes before the reads

    float a12 = 0.f, a6 = 0.f, cnt = 0.f, ew = 0.f, r2min = 3.0e38f;
    if (live) {
        if (all_in && !NOFOLD) total_pairs_v2<G, true>(X0, Y0, X1, Y1, N, t, P, a12, a6, cnt, ew, r2min);
        else total_pairs_v2<G, false>(X0, Y0, X1, Y1, N, t, P, a12, a6, cnt, ew, r2min);
    }
    double de = 4.0 * ((double)a12 - (double)a6) - (double)cnt * (double)P.e_cut + (double)ew;
    double dw = 48.0 * (double)a12 - 24.0 * (double)a6;
#pragma unroll
    for (int o = (G < 32 ? G : 32) / 2; o > 0; o >>= 1) {      // G = 16: two configurations per warp
        de += __shfl_xor_sync(0xffffffffu, de, o);
        dw += __shfl_xor_sync(0xffffffffu, dw, o);
        r2min = fminf(r2min, __shfl_xor_sync(0xffffffffu, r2min, o));
    }
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
    __shared__ int s_ok[GROUPS];
    if (live && t == 0) {
        const bool ov = r2min < P.rcore2;
        const float inf = __int_as_float(0x7f800000);
        const float e_out = ov ? inf : (float)de, w_out = ov ? inf : (float)dw;
        E[b] = e_out;
        W[b] = w_out;
        if (overlap) overlap[b] = ov ? 1 : 0;
        if (ACCEPT) s_ok[g] = accept_decide(A, b, e_out, w_out);
    }
    if (ACCEPT) {
        __syncthreads();
        if (live && s_ok[g]) {
            float2* dst = reinterpret_cast<float2*>(state) + (size_t)b * N;
            for (int i = t; i < N; i += G) dst[i] = make_float2(X0[i], Y0[i]);
        }
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

// Returns FS_OK, an error, or (fused request only) 1 when the tile is too large for the packed kernel and nothing was
// launched - the caller then runs the two-kernel sequence.
template <int G, int BS = 256>
static int launch_total(const float* pos, int B, int N, const PotDev& P, float* E, float* W,
                        unsigned char* ov, cudaStream_t s, float* state = nullptr, const AcceptArgs* acc = nullptr) {
    constexpr int GROUPS = BS / G;
    int grid = (B + GROUPS - 1) / GROUPS;
    const size_t smem2 = (size_t)GROUPS * 8 * N * sizeof(float);
    static long v2max = -1;                              // tuning knob: largest tile (bytes) of the packed variant
    if (v2max < 0) { const char* e = getenv("FS_ENERGY_V2MAX"); v2max = e ? atol(e) : 112 * 1024; }
    if ((long)smem2 <= v2max || BS == 1024) {            // 256 threads: while two blocks per SM still fit
        static int nofold = -1;                          // FS_ENERGY_NOFOLD=1: rint minimum image for in-box tiles too (A/B runs)
        if (nofold < 0) { const char* e = getenv("FS_ENERGY_NOFOLD"); nofold = e ? atoi(e) : 0; }
#define FS_LAUNCH_V2(NF, AC)                                                                                          \
    do {                                                                                                              \
        if (smem2 > 48 * 1024)                                                                                        \
            FS_CUDA(cudaFuncSetAttribute(energy_total_kernel_v2<G, NF, AC, BS>,                                       \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));                   \
        energy_total_kernel_v2<G, NF, AC, BS><<<grid, BS, smem2, s>>>(pos, B, N, P, E, W, ov, state,                  \
                                                                      acc ? *acc : AcceptArgs());                     \
    } while (0)
        if (acc) FS_LAUNCH_V2(false, true);
        else if (nofold) FS_LAUNCH_V2(true, false);
        else FS_LAUNCH_V2(false, false);
#undef FS_LAUNCH_V2
        fs::count_launch();
        return cuda_check(cudaGetLastError(), "energy_total_kernel_v2");
    }
    if (acc) return 1;
    if constexpr (G >= 32 && G <= 256) {
        size_t smem = (size_t)GROUPS * 2 * N * sizeof(float2);
        if (smem > 48 * 1024)
            FS_CUDA(cudaFuncSetAttribute(energy_total_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        energy_total_kernel<G><<<grid, 256, smem, s>>>(pos, B, N, P, E, W, ov);
        fs::count_launch();
        return cuda_check(cudaGetLastError(), "energy_total_kernel");
    } else {
        set_error("fs_energy_total: group size %d has no scalar kernel (N = %d)", G, N);
        return FS_ERR_UNSUPPORTED;
    }
}

static int group_size(int N) {
    static int g_forced = -1;                            // tuning knob (16 / 32 / 64 / 128 / 256 / 1024)
    if (g_forced < 0) { const char* e = getenv("FS_ENERGY_G"); g_forced = e ? atoi(e) : 0; }
    // (measured, scripts/energy_sweep.py with FS_ENERGY_G: about four particles per thread is best - several
    // configurations per block hide each other's load / reduction phases, and the smaller groups need fewer registers)
    if (g_forced) return g_forced;
    // beyond the two-blocks-per-SM tile (N > 3584) one block of 1024 threads owns the SM: the packed kernel up to
    // N = 7168 (224 KB tile) instead of the scalar one
    if (N > 3584 && (size_t)N * 32 <= 224 * 1024) return 1024;
    return N <= 80 ? 16 : (N <= 160 ? 32 : (N <= 320 ? 64 : (N <= 768 ? 128 : 256)));
}

static int total_dispatch(const float* pos, int B, int N, const PotDev& P, float* E, float* W, unsigned char* ov,
                          cudaStream_t s, float* state = nullptr, const AcceptArgs* acc = nullptr) {
    const int G = group_size(N);
    if (G == 1024 && (size_t)N * 32 <= 224 * 1024) return launch_total<1024, 1024>(pos, B, N, P, E, W, ov, s, state, acc);
    if (G == 16) return launch_total<16>(pos, B, N, P, E, W, ov, s, state, acc);
    if (G == 32) return launch_total<32>(pos, B, N, P, E, W, ov, s, state, acc);
    if (G == 64) return launch_total<64>(pos, B, N, P, E, W, ov, s, state, acc);
    if (G == 128) return launch_total<128>(pos, B, N, P, E, W, ov, s, state, acc);
    return launch_total<256>(pos, B, N, P, E, W, ov, s, state, acc);
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
    return fs::total_dispatch(pos, B, N, P, E, W, overlap, s);
}

extern "C" int fs_accept_global_fused(float* pos, const float* prop, double* E, double* W, float* E_new, float* W_new,
                                      const float* logq_old, const float* logq_new, const double* u,
                                      const fs_rng* rng, double beta, long long* attempts, long long* accepted,
                                      unsigned char* accept_mask, int B, int N, float Lx, float Ly, const fs_pot* pot,
                                      void* stream) {
    if (!pos || !prop || !pot || !E_new || !W_new || B < 0 || N < 1 || !(Lx > 0) || !(Ly > 0) || pos == prop) {
        fs::set_error("fs_accept_global_fused: invalid argument");
        return FS_ERR_INVALID;
    }
    if (N > 12288) {
        fs::set_error("fs_accept_global_fused: N=%d exceeds the shared-memory tile (max 12288)", N);
        return FS_ERR_UNSUPPORTED;
    }
    fs::AcceptArgs A;
    int r = fs::make_accept_args("fs_accept_global_fused", E, W, logq_old, logq_new, u, rng, beta, attempts, accepted,
                                 accept_mask, &A);
    if (r != FS_OK) return r;
    if (B == 0) return FS_OK;
    fs::PotDev P = fs::make_pot(pot, Lx, Ly);
    cudaStream_t s = (cudaStream_t)stream;
    r = fs::total_dispatch(prop, B, N, P, E_new, W_new, nullptr, s, pos, &A);
    if (r != 1) return r;
    // tile too large for the packed kernel: energy, then the stand-alone accept kernel
    r = fs::total_dispatch(prop, B, N, P, E_new, W_new, nullptr, s);
    if (r != FS_OK) return r;
    return fs_accept_global(pos, prop, E, W, E_new, W_new, logq_old, logq_new, u, rng, beta, attempts, accepted,
                            accept_mask, B, N, stream);
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
