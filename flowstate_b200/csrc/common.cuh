// Shared device helpers: pair potential, wells, RNG, error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/flowstate_b200.h"

namespace fs {

void set_error(const char* fmt, ...);
int cuda_check(cudaError_t e, const char* what);
void count_launch(int n = 1);
#define FS_CUDA(x)                                       \
    do {                                                 \
        int _r = fs::cuda_check((x), #x);                \
        if (_r) return _r;                               \
    } while (0)

// Potential constants handed to kernels by value.
struct PotDev {
    int num_wells;
    float V0[2];
    float r0, k;
    float rc2, rcore2, e_cut;
    float Lx, Ly, inv_Lx, inv_Ly;
    // wells: geometry in float64 (the wall k (r - r0) is steep: k = 15 turns a float32 rounding of
    // r, r0 or the centre into > 1e-5 of energy), potential.py:89-112
    double cxd[2], cyd, Lxd, Lyd, r0d, k2d;
    // float32 pre-test: outside r0 + m the term is 0, inside r0 - m it is V0, with m = 12 / k: |2k(r - r0)| > 24 there
    // and V0 e^-24 = 4e-10 is five orders below the 1e-5 tolerance.  well_shift_ok: the box is wide enough that a
    // particle passing the pre-test has an unambiguous periodic image (the float32 image shift can be reused).
    float cxf[2], cyf, well_out2, well_in;
    int well_shift_ok;
};

inline PotDev make_pot(const fs_pot* p, float Lx, float Ly) {
    PotDev d;
    d.num_wells = p->num_wells;
    d.V0[0] = (float)p->V0[0];
    d.V0[1] = (float)p->V0[1];
    d.r0 = (float)p->r0;
    d.k = (float)p->k;
    d.rc2 = (float)(p->r_cut * p->r_cut);
    d.rcore2 = (float)(p->r_core * p->r_core);
    double s6 = pow(1.0 / (double)p->r_cut, 6.0);
    d.e_cut = (float)(4.0 * (s6 * s6 - s6));          // potential.py:21-26
    d.Lx = Lx;
    d.Ly = Ly;
    d.inv_Lx = 1.0f / Lx;
    d.inv_Ly = 1.0f / Ly;
    d.Lxd = (double)Lx;
    d.Lyd = (double)Ly;
    d.cxd[0] = d.Lxd / 4;
    d.cxd[1] = 3 * d.Lxd / 4;
    d.cyd = d.Lyd / 2;
    d.r0d = (double)p->r0;
    d.k2d = 2.0 * (double)p->k;
    d.cxf[0] = (float)d.cxd[0];
    d.cxf[1] = (float)d.cxd[1];
    d.cyf = (float)d.cyd;
    {
        const double m = p->k > 0 ? 12.0 / (double)p->k : 1e30;
        const double ro = (double)p->r0 + m;
        d.well_out2 = (float)(ro * ro);
        d.well_in = (float)((double)p->r0 - m);                    // may be negative: shortcut never taken
        d.well_shift_ok = (0.5 * (double)Lx > 1.01 * ro + 1e-3 && 0.5 * (double)Ly > 1.01 * ro + 1e-3) ? 1 : 0;
    }
    return d;
}

// round-half-even of |t| < 2^22 with two full-rate adds (np.round, simulation_box.py:38-39)
__device__ __forceinline__ float rint_fast(float t) {
    const float magic = 12582912.0f;   // 1.5 * 2^23
    return __fsub_rn(__fadd_rn(t, magic), magic);
}

__device__ __forceinline__ float min_image(float d, float L, float invL) {
    return __fmaf_rn(-L, rint_fast(d * invL), d);
}

// One pair: accumulates shifted LJ energy and virial when r <= r_cut and tracks the
// smallest r^2 seen (hard-core test is done once on the minimum).
// potential.py:11-27, energy_calculator.py:73-81.
__device__ __forceinline__ void pair_accum(float dx, float dy, const PotDev& P,
                                           float& e, float& w, float& r2min) {
    dx = min_image(dx, P.Lx, P.inv_Lx);
    dy = min_image(dy, P.Ly, P.inv_Ly);
    float r2 = __fmaf_rn(dy, dy, dx * dx);
    r2min = fminf(r2min, r2);
    float inv = __frcp_rn(r2);
    float s6 = inv * inv * inv;
    if (r2 <= P.rc2) {
        e += __fmaf_rn(4.0f * s6, s6 - 1.0f, -P.e_cut);
        w += 48.0f * s6 * (s6 - 0.5f);
    }
}

// External double well of one particle (potential.py:95-112).
// V0 (1 - 0.5 (1 + tanh a)) == V0 / (1 + exp(2a)), evaluated in the stable form; the distance
// to the well centre is formed in float64 (float32 seed + one Newton step for the root).
__device__ __forceinline__ float well_term(float x, float y, int wi, const PotDev& P) {
    {   // float32 pre-test with a wide margin: far outside the wall the term is exactly 0, deep inside exactly V0
        const float fx = min_image(x - P.cxf[wi], P.Lx, P.inv_Lx);
        const float fy = min_image(y - P.cyf, P.Ly, P.inv_Ly);
        const float r2f = __fmaf_rn(fy, fy, fx * fx);
        if (r2f > P.well_out2) return 0.0f;
        if (P.well_in > 0.0f && r2f < P.well_in * P.well_in) return P.V0[wi];
    }
    double dx = (double)x - P.cxd[wi];
    double dy = (double)y - P.cyd;
    dx -= P.Lxd * rint(dx / P.Lxd);
    dy -= P.Lyd * rint(dy / P.Lyd);
    const double r2 = dx * dx + dy * dy;
    const float rf = sqrtf((float)r2);
    double r = 0.0;
    if (rf > 0.0f) r = (double)rf + (r2 - (double)rf * (double)rf) * (double)(0.5f / rf);
    const float a2 = (float)(P.k2d * (r - P.r0d));
    return P.V0[wi] / (1.0f + expf(a2));
}


// The same term for the throughput sweep: parameters picked by selects (a dynamically indexed kernel parameter lives
// in local memory), and on the wall the float64 geometry reuses the float32 image shift (no float64 division) when the
// box is wide enough for it to be unambiguous (well_shift_ok).
__device__ __forceinline__ float well_term_sel(float x, float y, int wi, const PotDev& P) {
    const float cxf = wi ? P.cxf[1] : P.cxf[0];
    const float v0 = wi ? P.V0[1] : P.V0[0];
    float fx = x - cxf, fy = y - P.cyf;
    const float nx = rint_fast(fx * P.inv_Lx), ny = rint_fast(fy * P.inv_Ly);
    fx = __fmaf_rn(-P.Lx, nx, fx);
    fy = __fmaf_rn(-P.Ly, ny, fy);
    const float r2f = __fmaf_rn(fy, fy, fx * fx);
    if (r2f > P.well_out2) return 0.0f;
    if (P.well_in > 0.0f && r2f < P.well_in * P.well_in) return v0;
    const double cxd = wi ? P.cxd[1] : P.cxd[0];
    double dx = (double)x - cxd;
    double dy = (double)y - P.cyd;
    if (P.well_shift_ok) {
        dx -= P.Lxd * (double)nx;
        dy -= P.Lyd * (double)ny;
    } else {
        dx -= P.Lxd * rint(dx / P.Lxd);
        dy -= P.Lyd * rint(dy / P.Lyd);
    }
    const double r2 = dx * dx + dy * dy;
    const float rf = sqrtf((float)r2);
    double r = 0.0;
    if (rf > 0.0f) r = (double)rf + (r2 - (double)rf * (double)rf) * (double)(0.5f / rf);
    const float a2 = (float)(P.k2d * (r - P.r0d));
    return v0 / (1.0f + expf(a2));
}

// Sum over the wells (potential.py:95-112).  well_term_sel: same value as well_term (the float32 image shift equals
// rint(dx / L) whenever the pre-test lets a particle through in a box wide enough - well_shift_ok), no float64
// division on the wall.
__device__ __forceinline__ float wells(float x, float y, const PotDev& P) {
    float v = 0.f;
    if (P.num_wells >= 1) v += well_term_sel(x, y, 0, P);
    if (P.num_wells == 2) v += well_term_sel(x, y, 1, P);
    return v;
}

// numpy float32 floor-mod (npy_divmodf), simulation_box.py:23-26 on a float32 state.
__device__ __forceinline__ float np_mod(float a, float b) {
    float m = fmodf(a, b);
    if (m != 0.0f) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysignf(0.0f, b);
    }
    return m;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------
// RNG
// ---------------------------------------------------------------------------
struct RngDev {
    int kind;
    unsigned long long* pcg_state;
    unsigned long long philox_seed;
    long long chain_id0;
    const int* replay_idx;
    const double* replay_u;
    int idx_stride, u_stride;
    int* replay_cursor;
};

inline RngDev make_rng(const fs_rng* r) {
    RngDev d;
    d.kind = r->kind;
    d.pcg_state = r->pcg_state;
    d.philox_seed = r->philox_seed;
    d.chain_id0 = r->chain_id0;
    d.replay_idx = r->replay_idx;
    d.replay_u = r->replay_u;
    d.idx_stride = r->idx_stride;
    d.u_stride = r->u_stride;
    d.replay_cursor = r->replay_cursor;
    return d;
}

// PCG64 (XSL-RR 128/64) exactly as numpy drives it: step, then output.
struct Pcg64 {
    uint64_t hi, lo, inc_hi, inc_lo;
    uint32_t has32, buf32;

    __device__ __forceinline__ uint64_t next64() {
        const uint64_t MH = 0x2360ED051FC65DA4ull, ML = 0x4385DF649FCCF645ull;
        uint64_t nlo = lo * ML;
        uint64_t nhi = __umul64hi(lo, ML) + hi * ML + lo * MH;
        uint64_t slo = nlo + inc_lo;
        nhi += inc_hi + (slo < nlo ? 1ull : 0ull);
        lo = slo;
        hi = nhi;
        uint64_t x = hi ^ lo;
        unsigned rot = (unsigned)(hi >> 58);
        return (x >> rot) | (x << ((64u - rot) & 63u));
    }
    __device__ __forceinline__ uint32_t next32() {   // pcg64_next32: low half first, high half buffered
        if (has32) {
            has32 = 0;
            return buf32;
        }
        uint64_t n = next64();
        has32 = 1;
        buf32 = (uint32_t)(n >> 32);
        return (uint32_t)n;
    }
    __device__ __forceinline__ double next_double() {
        return (double)(next64() >> 11) * (1.0 / 9007199254740992.0);
    }
    // Generator.integers(n), n < 2^32: buffered_bounded_lemire_uint32
    __device__ __forceinline__ uint32_t bounded(uint32_t n) {
        if (n <= 1) return 0;
        uint32_t rng = n - 1;
        uint64_t m = (uint64_t)next32() * (uint64_t)n;
        uint32_t left = (uint32_t)m;
        if (left < n) {
            uint32_t thr = (0xFFFFFFFFu - rng) % n;
            while (left < thr) {
                m = (uint64_t)next32() * (uint64_t)n;
                left = (uint32_t)m;
            }
        }
        return (uint32_t)(m >> 32);
    }
    __device__ __forceinline__ void load(const unsigned long long* s) {
        hi = s[0]; lo = s[1]; inc_hi = s[2]; inc_lo = s[3];
        has32 = (uint32_t)s[4]; buf32 = (uint32_t)s[5];
    }
    __device__ __forceinline__ void store(unsigned long long* s) const {
        s[0] = hi; s[1] = lo; s[4] = has32; s[5] = buf32;
    }
};

// Philox4x32-10
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t h0 = __umulhi(M0, ctr.x), l0 = M0 * ctr.x;
        uint32_t h1 = __umulhi(M1, ctr.z), l1 = M1 * ctr.z;
        ctr = make_uint4(h1 ^ ctr.y ^ key.x, l1, h0 ^ ctr.w ^ key.y, l0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ double u32x2_to_double(uint32_t a, uint32_t b) {
    uint64_t v = ((uint64_t)a << 32) | b;
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace fs
