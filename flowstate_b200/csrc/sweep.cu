// Local-displacement Metropolis sweep: `steps` single-particle moves on each of B
// independent chains in one launch.
//
// fs_local_sweep <- MonteCarlo.particle_displacement + metropolis_acceptance_particle_move
//                   (MCMC/monte_carlo.py:146-223) with the two
//                   calculate_particle_energy_virial calls (MCMC/energy_calculator.py:48-108)
// fs_adjust_displacement <- MonteCarlo.adjust_displacement (MCMC/monte_carlo.py:375-403)
//
// One warp owns one chain for the whole launch: the chain's positions live in
// shared memory, each lane evaluates the old and the new pair energy of the
// moved particle against its slice of the other particles from a single load,
// and six warp-shuffle reductions finish the step.  Draw order per step follows
// the reference: particle index, two displacement uniforms, and a third uniform
// only for finite uphill moves.
#include <stdlib.h>

#include "common.cuh"

namespace fs {

struct StepDraw {
    int p;
    double u1, u2;
};

template <int KIND>
struct ChainRng {
    Pcg64 pcg;
    uint2 key;
    uint4 ctr;
    uint4 blk;
    const int* ridx;
    const double* ru;
    int ci, cu;

    __device__ __forceinline__ void init(const RngDev& R, int b, long long attempts) {
        if (KIND == FS_RNG_PCG64) {
            pcg.load(R.pcg_state + (size_t)b * 6);
        } else if (KIND == FS_RNG_PHILOX) {
            key = make_uint2((uint32_t)R.philox_seed, (uint32_t)(R.philox_seed >> 32));
            long long cid = R.chain_id0 + b;
            ctr = make_uint4(0u, 0u, (uint32_t)cid, (uint32_t)((unsigned long long)cid >> 32));
        } else {
            ridx = R.replay_idx + (size_t)b * R.idx_stride;
            ru = R.replay_u + (size_t)b * R.u_stride;
            ci = R.replay_cursor[2 * b];
            cu = R.replay_cursor[2 * b + 1];
        }
    }
    // step_id = value of the attempts counter before this step
    __device__ __forceinline__ StepDraw draw(int N, long long step_id) {
        StepDraw d;
        if (KIND == FS_RNG_PCG64) {
            d.p = (int)pcg.bounded((uint32_t)N);
            d.u1 = pcg.next_double();
            d.u2 = pcg.next_double();
        } else if (KIND == FS_RNG_PHILOX) {
            // one Philox block per step id: {particle index, u1, u2, accept uniform}, 32-bit each
            ctr.x = (uint32_t)step_id;
            ctr.y = (uint32_t)((unsigned long long)step_id >> 32);
            blk = philox4x32(ctr, key);
            d.p = (int)__umulhi(blk.x, (uint32_t)N);
            d.u1 = (double)blk.y * (1.0 / 4294967296.0);
            d.u2 = (double)blk.z * (1.0 / 4294967296.0);
        } else {
            d.p = ridx[ci++];
            d.u1 = ru[cu++];
            d.u2 = ru[cu++];
        }
        return d;
    }
    __device__ __forceinline__ double accept_uniform() {
        if (KIND == FS_RNG_PCG64) return pcg.next_double();
        if (KIND == FS_RNG_PHILOX) return (double)blk.w * (1.0 / 4294967296.0);
        return ru[cu++];
    }
    __device__ __forceinline__ void finish(const RngDev& R, int b) {
        if (KIND == FS_RNG_PCG64) {
            pcg.store(R.pcg_state + (size_t)b * 6);
        } else if (KIND == FS_RNG_REPLAY) {
            R.replay_cursor[2 * b] = ci;
            R.replay_cursor[2 * b + 1] = cu;
        }
    }
};

template <int KIND>
__global__ void __launch_bounds__(128) local_sweep_kernel(float* __restrict__ pos, double* __restrict__ E,
                                                          double* __restrict__ W,
                                                          const double* __restrict__ max_disp,
                                                          long long* __restrict__ attempts,
                                                          long long* __restrict__ accepted, int B, int N,
                                                          int steps, PotDev P, double beta, RngDev R,
                                                          unsigned char* __restrict__ trace_accept,
                                                          int* __restrict__ trace_idx,
                                                          float* __restrict__ trace_e) {
    extern __shared__ float2 smem[];
    const int wib = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + wib;
    if (b >= B) return;
    float2* sp = smem + (size_t)wib * N;
    float2* gp = reinterpret_cast<float2*>(pos) + (size_t)b * N;
    for (int i = lane; i < N; i += 32) sp[i] = gp[i];
    __syncwarp();

    const double md = max_disp[b];
    long long att = attempts[b];
    long long acc = accepted[b];
    double Eb = E[b], Wb = W[b];
    ChainRng<KIND> rng;
    rng.init(R, b, att);
    const float inf = __int_as_float(0x7f800000);

    for (int s = 0; s < steps; ++s) {
        StepDraw d = rng.draw(N, att);
        att += 1;
        const int p = d.p;
        const float2 old = sp[p];
        // new_positions[p] += displacement (float64 add, stored float32), then % L
        // (monte_carlo.py:161-166 on a float32 state)
        float nx = (float)((double)old.x + (d.u1 - 0.5) * md);
        float ny = (float)((double)old.y + (d.u2 - 0.5) * md);
        nx = np_mod(nx, P.Lx);
        ny = np_mod(ny, P.Ly);

        float eo = 0.f, wo = 0.f, en = 0.f, wn = 0.f, mo = 3.0e38f, mn = 3.0e38f;
        for (int j = lane; j < N; j += 32) {
            if (j == p) continue;
            const float2 q = sp[j];
            pair_accum(old.x - q.x, old.y - q.y, P, eo, wo, mo);
            pair_accum(nx - q.x, ny - q.y, P, en, wn, mn);
        }
        // wells of the old / new position: lanes 0..3 take one tanh each
        if (P.num_wells == 2) {
            if (lane < 2) eo += well_term(old.x, old.y, lane, P);
            else if (lane < 4) en += well_term(nx, ny, lane - 2, P);
        } else if (P.num_wells == 1) {
            if (lane == 0) eo += well_term(old.x, old.y, 0, P);
            else if (lane == 1) en += well_term(nx, ny, 0, P);
        }
        eo = warp_sum(eo);
        en = warp_sum(en);
        wo = warp_sum(wo);
        wn = warp_sum(wn);
        mo = warp_min(mo);
        mn = warp_min(mn);
        if (mo < P.rcore2) { eo = inf; wo = inf; }
        if (mn < P.rcore2) { en = inf; wn = inf; }

        bool ok;
        if (en <= eo) {
            ok = true;
        } else if (en == inf) {
            ok = false;
        } else {
            const double factor = exp(-beta * ((double)en - (double)eo));
            const double u = rng.accept_uniform();
            ok = u < factor;
        }
        if (ok) {
            if (lane == 0) sp[p] = make_float2(nx, ny);
            acc += 1;
            Eb += (double)en - (double)eo;
            Wb += (double)wn - (double)wo;
        }
        if (lane == 0) {
            const size_t o = (size_t)b * steps + s;
            if (trace_accept) trace_accept[o] = ok ? 1 : 0;
            if (trace_idx) trace_idx[o] = p;
            if (trace_e) {
                trace_e[2 * o] = eo;
                trace_e[2 * o + 1] = en;
            }
        }
        __syncwarp();
    }
    for (int i = lane; i < N; i += 32) gp[i] = sp[i];
    if (lane == 0) {
        attempts[b] = att;
        accepted[b] = acc;
        E[b] = Eb;
        W[b] = Wb;
        rng.finish(R, b);
    }
}


// Throughput kernel (Philox streams; what the drivers and the bench run): the same move with the per-step overheads
// trimmed - one Philox block per step generated LPC steps at a time (lane `sub` of a chain's lane group prepares step
// base + sub, the step reads it with four shuffles), energy / virial DIFFERENCES reduced instead of four separate
// sums, hard-core tests by ballot, wells behind a float32 pre-test, several chains per warp, an exact minimum image
// for close pairs.
// packed FP32 pairs (add/sub/mul/fma.f32x2)
__device__ __forceinline__ unsigned long long pk2s(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2s(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2s(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long add2s(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long sub2s(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long mul2s(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// numpy float32 floor-mod for the common case -L <= a < 2L (one shift, no branch); anything else takes np_mod.
__device__ __noinline__ float np_mod_far(float a, float L) { return np_mod(a, L); }
__device__ __forceinline__ float np_mod_near(float a, float L) {
    // a >= L: a - L is exact (Sterbenz), what fmodf returns; a < 0: fmodf keeps a and numpy adds the divisor (the sum
    // may round to L); 0 <= a < L: unchanged, with -0.0 -> +0.0 like copysignf(0, L)
    const float shift = (a >= L) ? -L : ((a < 0.0f) ? L : 0.0f);
    float r = a + shift;
    if (__builtin_expect(!(a >= -L && a < 2.0f * L), 0)) r = np_mod_far(a, L);   // rare: more than one box length away
    return r;
}

// LPC lanes of a warp own one chain (32 / LPC chains per warp): the per-step scalar work (random numbers, the
// displacement, the floor-mod, the decision) is issued once per warp-instruction for all chains of the warp, each lane
// walks ceil(N / LPC) partners of the moved particle, and the step finishes with a log2(LPC)-level segmented shuffle
// sum of the energy difference and two ballots for the hard-core tests.  Chain slots in shared memory are `stride2`
// float2 apart (stride2 * 8 bytes = 64 mod 128, so the groups of a warp read disjoint banks).
// TRACE adds the outputs of the parity tests (accept flag, particle index, e_old / e_new reduced separately); the
// decision arithmetic is the same code either way, so a traced run follows the untraced trajectory bit for bit.
// SKIP (dilute boxes) keeps ONLY (x, y) in shared memory and derives the centred copy of the few partners that pass the
// cut-off pre-test on the fly (two selects + two adds): 8 bytes per particle instead of 16, so that - with the register
// budget of seven 128-thread blocks per SM - all 1024 blocks of the N = 256 hybrid configuration (8192 chains) are
// resident at once.  With both copies stored the launch was 1.38 waves of equally long blocks, i.e. two full block
// times, the second one at a third of the occupancy and spread thinly over every SM.
template <int LPC, bool TRACE, bool SKIP>
__global__ void __launch_bounds__(128, SKIP ? 7 : 5) local_sweep_fast_kernel(float* __restrict__ pos, double* __restrict__ E,
                                                               double* __restrict__ W,
                                                               const double* __restrict__ max_disp,
                                                               long long* __restrict__ attempts,
                                                               long long* __restrict__ accepted, int B, int N,
                                                               int steps, PotDev P, double beta,
                                                               unsigned long long seed, long long chain_id0,
                                                               int stride2, int slot2,
                                                               unsigned char* __restrict__ trace_accept,
                                                               int* __restrict__ trace_idx,
                                                               float* __restrict__ trace_e) {
    constexpr int CPW = 32 / LPC;                 // chains per warp
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ float4 smem4[];
    const int wib = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int grp = lane / LPC, sub = lane % LPC, gl0 = grp * LPC;
    const int b_first = (blockIdx.x * (blockDim.x >> 5) + wib) * CPW;
    if (b_first >= B) return;
    const bool live = b_first + grp < B;          // groups past the last chain shadow it and store nothing
    const int b = live ? b_first + grp : B - 1;
    // Chain state in shared memory: per particle (x, y, cx, cy) with c = x - L [x > L / 2], the same point seen from the
    // box centre.  The minimum image of a difference is whichever of x_i - x_j and c_i - c_j is smaller in magnitude
    // (both lie in {d, d +- L}, and one of them has |.| <= L / 2); for a close pair the smaller one is a difference of
    // nearby float32 numbers and therefore exact - the rint-based form subtracts at magnitude L first and loses up to
    // ulp(L) / r of relative accuracy on pairs that straddle the periodic boundary (1e-5 of r^-12 at L = 11).
    // (two float2 arrays per chain slot: the dilute-box variant reads the centred copy only for warp trips that hold a
    // pair inside the cut-off)
    const float hLx = 0.5f * P.Lx, hLy = 0.5f * P.Ly;
    constexpr bool CEN = !SKIP;                  // centred copy stored (slot2 = 2 * stride2) or derived on the fly
    float2* sp = reinterpret_cast<float2*>(smem4) + (size_t)(wib * CPW + grp) * slot2;
    float2* sc = sp + stride2;                   // CEN only
    auto centred = [&](float2 v) { return make_float2(v.x - (v.x > hLx ? P.Lx : 0.f), v.y - (v.y > hLy ? P.Ly : 0.f)); };
    float2* gp = reinterpret_cast<float2*>(pos) + (size_t)b * N;
    const int iters = (N + LPC - 1) / LPC;
    const float qnan = __int_as_float(0x7fc00000);
    // Slots are padded to iters * LPC entries with NaN positions: a NaN r^2 fails the cut-off test (every term 0) and is
    // ignored by fminf, so neither the padding nor the moved particle itself (overwritten with NaN while its partners
    // are walked) needs a mask inside the pair loop.
    for (int i = sub; i < iters * LPC; i += LPC) {
        float2 v = make_float2(qnan, qnan), c = v;
        if (i < N) {
            v = gp[i];
            c = centred(v);
        }
        sp[i] = v;
        if (CEN) sc[i] = c;
    }
    __syncwarp();

    const double md = max_disp[b];
    const long long att0 = attempts[b];
    int acc = 0;
    double Eb = E[b];
    double Wl = 0.0;                         // this lane's share of the accepted virial differences (summed at the end)
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const long long cid = chain_id0 + b;
    const uint32_t cz = (uint32_t)cid, cw = (uint32_t)((unsigned long long)cid >> 32);
    const float inf = __int_as_float(0x7f800000);
    const float Lx = P.Lx, Ly = P.Ly, rc2 = P.rc2, ecut = P.e_cut, rcore2 = P.rcore2;
    const float nbeta = -(float)beta;
    // wells: lanes 0..3 of a group evaluate (old, well 0), (old, well 1), (new, well 0), (new, well 1) with one code path
    const int nw = P.num_wells;
    const bool well_lane = (nw == 2) ? sub < 4 : (nw == 1 ? (sub == 0 || sub == 2) : false);
    const int well_idx = sub & 1;
    const bool well_new = (sub & 2) != 0;
    const unsigned gmask = (LPC == 32) ? FULL : (((1u << LPC) - 1u) << gl0);
    const uint32_t a0 = (uint32_t)att0;      // low bits of the step id: slot inside the LPC-step block of random numbers
    uint4 blk = make_uint4(0, 0, 0, 0);
    struct {
        unsigned long long iLx, iLy, nLx, nLy, magic, nmagic, four, mone, mhalf;
    } K;
    K.iLx = pk2s(P.inv_Lx, P.inv_Lx); K.iLy = pk2s(P.inv_Ly, P.inv_Ly); K.nLx = pk2s(-Lx, -Lx); K.nLy = pk2s(-Ly, -Ly);
    K.magic = pk2s(12582912.0f, 12582912.0f); K.nmagic = pk2s(-12582912.0f, -12582912.0f);
    K.four = pk2s(4.0f, 4.0f); K.mone = pk2s(-1.0f, -1.0f); K.mhalf = pk2s(-0.5f, -0.5f);

    // One Philox block per step id {particle index, u1, u2, accept uniform}; a group prepares LPC steps at a time
    // (lane `sub` the step  base + sub) and reads them back by shuffles.  The draws of step s + 1 are fetched while
    // step s runs (they do not depend on it), which takes them off the step's dependent instruction chain.
    int p_n;
    double dx_n, dy_n;
    uint32_t u3_n;
    auto fetch = [&](int s) {
        const uint32_t slot = (a0 + (uint32_t)s) & (LPC - 1);
        if (__any_sync(FULL, slot == 0 || s == 0)) {          // groups of a warp whose counters are not aligned refill
            const unsigned long long sid = (unsigned long long)(att0 + s - (long long)slot + sub);   // their own block again
            blk = philox4x32(make_uint4((uint32_t)sid, (uint32_t)(sid >> 32), cz, cw), key);
        }
        const int src = gl0 + (int)slot;
        const uint32_t r_idx = __shfl_sync(FULL, blk.x, src);
        const uint32_t r_u1 = __shfl_sync(FULL, blk.y, src);
        const uint32_t r_u2 = __shfl_sync(FULL, blk.z, src);
        u3_n = __shfl_sync(FULL, blk.w, src);
        p_n = (int)__umulhi(r_idx, (uint32_t)N);
        dx_n = ((double)r_u1 * (1.0 / 4294967296.0) - 0.5) * md;
        dy_n = ((double)r_u2 * (1.0 / 4294967296.0) - 0.5) * md;
    };
    fetch(0);

    for (int s = 0; s < steps; ++s) {
        const int p = p_n;
        const double ddx = dx_n, ddy = dy_n;
        const uint32_t r_u3 = u3_n;
        const float2 old = sp[p], oldc = CEN ? sc[p] : centred(old);
        __syncwarp();
        if (sub == 0) {                                        // the moved particle is not its own partner
            sp[p] = make_float2(qnan, qnan);
            if (CEN) sc[p] = make_float2(qnan, qnan);
        }
        // new_positions[p] += displacement (float64 add, stored float32), then % L (monte_carlo.py:161-166)
        float nx = (float)((double)old.x + ddx);
        float ny = (float)((double)old.y + ddy);
        nx = np_mod_near(nx, Lx);
        ny = np_mod_near(ny, Ly);
        __syncwarp();
        if (s + 1 < steps) fetch(s + 1);

        // pair terms of the moved particle at its old and new position (energy_calculator.py:48-108); (old, new) ride
        // in the two halves of Blackwell's packed FP32 pairs (FADD2 / FMUL2 / FFMA2).  SKIP (dilute boxes): a trip of
        // the warp whose 64 pairs all lie outside the cut-off contributes nothing and ends after r^2.
        float mo = 3.0e38f, mn = 3.0e38f;
        unsigned long long e2 = 0ull, w2 = 0ull;                     // (sum e_old, sum e_new), (sum w_old, sum w_new)
        const float cnx = nx - (nx > hLx ? Lx : 0.f), cny = ny - (ny > hLy ? Ly : 0.f);
        const unsigned long long PX = pk2s(old.x, nx), PY = pk2s(old.y, ny), PCX = pk2s(oldc.x, cnx),
                                 PCY = pk2s(oldc.y, cny);
#pragma unroll 4
        for (int it = 0; it < iters; ++it) {
            const float2 q = sp[sub + it * LPC];
            const unsigned long long X0 = sub2s(PX, pk2s(q.x, q.x)), Y0 = sub2s(PY, pk2s(q.y, q.y));
            if (SKIP) {
                // dilute boxes: the cut-off test of the whole warp trip on the rint-based image d - L rint(d / L) (its
                // float32 error is irrelevant at r = r_cut, where the shifted energy is 0); most trips end here
                const unsigned long long X = fma2s(add2s(fma2s(X0, K.iLx, K.magic), K.nmagic), K.nLx, X0);
                const unsigned long long Y = fma2s(add2s(fma2s(Y0, K.iLy, K.magic), K.nmagic), K.nLy, Y0);
                float ta, tb;
                upk2s(fma2s(Y, Y, mul2s(X, X)), ta, tb);
                if (!__any_sync(FULL, fminf(ta, tb) <= 1.0001f * rc2)) continue;
            }
            // both candidates of the minimum image per axis, squared; the smaller one is the image (see above)
            const float2 qc = CEN ? sc[sub + it * LPC] : centred(q);     // NaN stays NaN (the comparison is false)
            const unsigned long long X1 = sub2s(PCX, pk2s(qc.x, qc.x)), Y1 = sub2s(PCY, pk2s(qc.y, qc.y));
            float ax0, bx0, ax1, bx1, ay0, by0, ay1, by1;
            upk2s(mul2s(X0, X0), ax0, bx0);
            upk2s(mul2s(X1, X1), ax1, bx1);
            upk2s(mul2s(Y0, Y0), ay0, by0);
            upk2s(mul2s(Y1, Y1), ay1, by1);
            const float r2o = fminf(ax0, ax1) + fminf(ay0, ay1), r2n = fminf(bx0, bx1) + fminf(by0, by1);
            mo = fminf(mo, r2o);
            mn = fminf(mn, r2n);
            float io, in_;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(io) : "f"(r2o));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(in_) : "f"(r2n));
            const bool co = r2o <= rc2, cn = r2n <= rc2;          // outside the cut-off (and self): every term is 0
            const unsigned long long inv = pk2s(co ? io : 0.f, cn ? in_ : 0.f);
            const unsigned long long s6 = mul2s(mul2s(inv, inv), inv);
            // e = 4 s6 (s6 - 1) - e_cut, w = 48 s6 (s6 - 1/2)   (potential.py:11-27)
            e2 = add2s(e2, fma2s(mul2s(s6, K.four), add2s(s6, K.mone), pk2s(co ? -ecut : 0.f, cn ? -ecut : 0.f)));
            w2 = fma2s(s6, add2s(s6, K.mhalf), w2);
        }
        float eo_s, en_s, wo_s, wn_s;
        upk2s(e2, eo_s, en_s);
        upk2s(w2, wo_s, wn_s);
        float de = en_s - eo_s;
        const float dw = 48.0f * (wn_s - wo_s);
        float wv = 0.f;
        if (well_lane) wv = well_term_sel(well_new ? nx : old.x, well_new ? ny : old.y, well_idx, P);
        de += well_new ? wv : -wv;
#pragma unroll
        for (int o = LPC / 2; o > 0; o >>= 1) de += __shfl_xor_sync(FULL, de, o);
        const bool ov_o = (__ballot_sync(FULL, mo < rcore2) & gmask) != 0u;
        const bool ov_n = (__ballot_sync(FULL, mn < rcore2) & gmask) != 0u;

        // metropolis_acceptance_particle_move (monte_carlo.py:191-223): e_old = inf makes `new <= old` true for any
        // e_new; otherwise an overlapping new position is rejected; downhill is accepted; uphill draws
        // u < exp(-beta dE).  A float exponential decides unless u falls within 1e-4 (relative) of the threshold;
        // only then is the float64 expression evaluated, so the decision is always the float64 one.
        const float uf = (float)r_u3 * (1.0f / 4294967296.0f);
        const float pf = __expf(nbeta * de);
        bool ok = ov_o || (!ov_n && (de <= 0.f || uf < pf * 0.9999f));
        if (!ov_o && !ov_n && de > 0.f && uf >= pf * 0.9999f && uf <= pf * 1.0001f)
            ok = (double)r_u3 * (1.0 / 4294967296.0) < exp(-beta * (double)de);
        if (TRACE) {
            // e_old / e_new of the moved particle, reduced separately (outputs only: the decision above used `de`)
            float eo_t = eo_s + (well_new ? 0.f : wv), en_t = en_s + (well_new ? wv : 0.f);
#pragma unroll
            for (int o = LPC / 2; o > 0; o >>= 1) {
                eo_t += __shfl_xor_sync(FULL, eo_t, o);
                en_t += __shfl_xor_sync(FULL, en_t, o);
            }
            if (sub == 0 && live) {
                const size_t o = (size_t)b * steps + s;
                if (trace_accept) trace_accept[o] = ok ? 1 : 0;
                if (trace_idx) trace_idx[o] = p;
                if (trace_e) {
                    trace_e[2 * o] = ov_o ? inf : eo_t;
                    trace_e[2 * o + 1] = ov_n ? inf : en_t;
                }
            }
        }
        if (sub == 0) {
            sp[p] = ok ? make_float2(nx, ny) : old;
            if (CEN) sc[p] = ok ? make_float2(cnx, cny) : oldc;
        }
        if (ok) {
            acc += 1;
            if (__builtin_expect(ov_o || ov_n, 0)) {           // leaving (or, from an overlap, entering) the hard core
                const float big_o = ov_o ? inf : 0.f, big_n = ov_n ? inf : 0.f;
                Eb += (double)de + ((double)big_n - (double)big_o);
                if (sub == 0) Wl += (double)big_n - (double)big_o;
            } else {
                Eb += (double)de;
            }
            Wl += (double)dw;                                  // float64 per step: W does not depend on launch splitting
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = LPC / 2; o > 0; o >>= 1) Wl += __shfl_xor_sync(FULL, Wl, o);
    if (live) {
        for (int i = sub; i < N; i += LPC) gp[i] = sp[i];
        if (sub == 0) {
            attempts[b] = att0 + steps;
            accepted[b] += acc;
            E[b] = Eb;
            W[b] += Wl;
        }
    }
}

__global__ void adjust_displacement_kernel(double* max_disp, const long long* att, const long long* acc,
                                           long long* patt, long long* pacc, double target, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (att[b] > patt[b]) {
        const long long da = att[b] - patt[b];
        const long long dc = acc[b] - pacc[b];
        const double frac = da > 0 ? (double)dc / (double)da : 0.0;
        const double md = max_disp[b];
        double nm = md * (frac / target);
        const double ratio = nm / md;
        if (ratio > 1.5) nm = md * 1.5;
        else if (ratio < 0.5) nm = md * 0.5;
        max_disp[b] = nm;
        patt[b] = att[b];
        pacc[b] = acc[b];
    }
}

template <int KIND>
static int launch_sweep(float* pos, double* E, double* W, const double* md, long long* att, long long* acc,
                        int B, int N, int steps, const PotDev& P, double beta, const RngDev& R,
                        unsigned char* ta, int* ti, float* te, cudaStream_t s) {
    // warps per CTA: 4 unless the chain state is large
    int wpc = 4;
    while (wpc > 1 && (size_t)wpc * N * sizeof(float2) > 200 * 1024) wpc >>= 1;
    size_t smem = (size_t)wpc * N * sizeof(float2);
    if (smem > 227 * 1024) {
        set_error("fs_local_sweep: N=%d does not fit in shared memory", N);
        return FS_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        FS_CUDA(cudaFuncSetAttribute(local_sweep_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (B + wpc - 1) / wpc;
    local_sweep_kernel<KIND><<<grid, wpc * 32, smem, s>>>(pos, E, W, md, att, acc, B, N, steps, P, beta, R, ta, ti, te);
    fs::count_launch();
    return cuda_check(cudaGetLastError(), "local_sweep_kernel");
}

template <int LPC, bool TRACE, bool SKIP>
static int launch_fast_t(float* pos, double* E, double* W, const double* md, long long* att, long long* acc, int B,
                         int N, int steps, const PotDev& P, double beta, unsigned long long seed, long long chain_id0,
                         unsigned char* ta, int* ti, float* te, cudaStream_t s) {
    constexpr int CPW = 32 / LPC;
    // slot stride (float4 units): N rounded up to whole trips of the lane group (room for the NaN padding)
    const int stride2 = (N + LPC - 1) / LPC * LPC;
    // float2 units between two chains' slots: (x, y) and the centred copy, or (SKIP) the positions alone - then padded
    // to 8 mod 16 for the 8-lane groups so that the two groups of a half-warp read different banks
    const int slot2 = SKIP ? (LPC == 8 ? stride2 + ((8 - stride2 % 16) + 16) % 16 : stride2) : 2 * stride2;
    int wpc = 4;
    while (wpc > 1 && (size_t)wpc * CPW * slot2 * sizeof(float2) > 200 * 1024) wpc >>= 1;
    const size_t smem = (size_t)wpc * CPW * slot2 * sizeof(float2);
    if (smem > 227 * 1024) {
        set_error("fs_local_sweep: N=%d does not fit in shared memory", N);
        return FS_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        FS_CUDA(cudaFuncSetAttribute(local_sweep_fast_kernel<LPC, TRACE, SKIP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    const int cpc = wpc * CPW;
    local_sweep_fast_kernel<LPC, TRACE, SKIP><<<(B + cpc - 1) / cpc, wpc * 32, smem, s>>>(
        pos, E, W, md, att, acc, B, N, steps, P, beta, seed, chain_id0, stride2, slot2, ta, ti, te);
    fs::count_launch();
    return cuda_check(cudaGetLastError(), "local_sweep_fast_kernel");
}

// lanes per chain: 8 (four chains per warp) for small systems, 16 for medium ones, else 32
static int launch_fast(float* pos, double* E, double* W, const double* md, long long* att, long long* acc, int B, int N,
                       int steps, const PotDev& P, double beta, unsigned long long seed, long long chain_id0,
                       unsigned char* ta, int* ti, float* te, cudaStream_t s) {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("FS_SWEEP_LPC"); forced = e ? atoi(e) : 0; }   // tuning knob (8 / 16 / 32)
    int lpc = forced ? forced : (N <= 128 ? 8 : (N <= 768 ? 16 : 32));   // measured: N = 32 / 64: 8, N = 256: 16
    const bool tr = ta || ti || te;
    // dilute boxes: most warp trips of the pair loop see no partner inside the cut-off (probability of a pair inside it
    // is pi rc^2 / (Lx Ly); a trip holds 64 pairs) -> early-out variant
    static int skip_forced = -1;
    if (skip_forced < 0) { const char* e = getenv("FS_SWEEP_SKIP"); skip_forced = e ? (atoi(e) ? 1 : 0) : 2; }
    const bool skip = skip_forced == 2 ? (64.0 * 3.14159 * P.rc2 < 0.75 * (double)P.Lx * (double)P.Ly) : skip_forced == 1;
#define FS_FAST2(L, T, S) \
    return launch_fast_t<L, T, S>(pos, E, W, md, att, acc, B, N, steps, P, beta, seed, chain_id0, ta, ti, te, s)
#define FS_FAST(L)                                     \
    do {                                               \
        if (tr) { if (skip) FS_FAST2(L, true, true); else FS_FAST2(L, true, false); }     \
        else { if (skip) FS_FAST2(L, false, true); else FS_FAST2(L, false, false); }      \
    } while (0)
    if (lpc == 8) { FS_FAST(8); }
    if (lpc == 16) { FS_FAST(16); }
    FS_FAST(32);
#undef FS_FAST2
#undef FS_FAST
}

}  // namespace fs

extern "C" int fs_local_sweep(float* pos, double* E, double* W, const double* max_disp, long long* attempts,
                              long long* accepted, int B, int N, int steps, float Lx, float Ly, double beta,
                              const fs_pot* pot, const fs_rng* rng, unsigned char* trace_accept,
                              int* trace_idx, float* trace_e, void* stream) {
    if (!pos || !E || !W || !max_disp || !attempts || !accepted || !pot || !rng || B < 0 || N < 1 || steps < 0 ||
        !(Lx > 0) || !(Ly > 0)) {
        fs::set_error("fs_local_sweep: invalid argument");
        return FS_ERR_INVALID;
    }
    if (B == 0 || steps == 0) return FS_OK;
    fs::PotDev P = fs::make_pot(pot, Lx, Ly);
    fs::RngDev R = fs::make_rng(rng);
    cudaStream_t s = (cudaStream_t)stream;
    switch (rng->kind) {
        case FS_RNG_PCG64:
            if (!rng->pcg_state) break;
            return fs::launch_sweep<FS_RNG_PCG64>(pos, E, W, max_disp, attempts, accepted, B, N, steps, P, beta, R,
                                                  trace_accept, trace_idx, trace_e, s);
        case FS_RNG_PHILOX:
            return fs::launch_fast(pos, E, W, max_disp, attempts, accepted, B, N, steps, P, beta, rng->philox_seed,
                                   rng->chain_id0, trace_accept, trace_idx, trace_e, s);
        case FS_RNG_PHILOX_REF:      // the same Philox draws through the reference-order kernel (parity tests)
            R.kind = FS_RNG_PHILOX;
            return fs::launch_sweep<FS_RNG_PHILOX>(pos, E, W, max_disp, attempts, accepted, B, N, steps, P, beta, R,
                                                   trace_accept, trace_idx, trace_e, s);
        case FS_RNG_REPLAY:
            if (!rng->replay_idx || !rng->replay_u || !rng->replay_cursor) break;
            return fs::launch_sweep<FS_RNG_REPLAY>(pos, E, W, max_disp, attempts, accepted, B, N, steps, P, beta, R,
                                                   trace_accept, trace_idx, trace_e, s);
    }
    fs::set_error("fs_local_sweep: bad rng descriptor (kind=%d)", rng->kind);
    return FS_ERR_INVALID;
}

extern "C" int fs_adjust_displacement(double* max_disp, const long long* attempts, const long long* accepted,
                                      long long* prev_attempts, long long* prev_accepted, double target, int B,
                                      void* stream) {
    if (!max_disp || !attempts || !accepted || !prev_attempts || !prev_accepted || B < 0 || !(target > 0)) {
        fs::set_error("fs_adjust_displacement: invalid argument");
        return FS_ERR_INVALID;
    }
    if (B == 0) return FS_OK;
    fs::adjust_displacement_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        max_disp, attempts, accepted, prev_attempts, prev_accepted, target, B);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "adjust_displacement_kernel");
}
