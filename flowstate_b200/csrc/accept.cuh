// The flow-augmented Metropolis rule of one chain, shared by accept_global_kernel (accept.cu) and the fused
// energy + accept kernel (energy.cu).  MonteCarlo.nf_big_move steps 3-5 (MCMC/monte_carlo.py:264-303):
//   ratio_log = -beta (E_new - E_old) - (nll_new - nll_old); ratio = exp(ratio_log)   (float64 like the reference)
//   accept if ratio >= 1, else draw ONE uniform and accept if u < ratio
//   accept: accepted += 1, cached energy / virial <- the proposal's; attempts += 1 in every case (:240).
#pragma once
#include "common.cuh"

namespace fs {

struct AcceptArgs {
    double* E;
    double* W;
    const float* lq_old;
    const float* lq_new;
    const double* u_in;      // optional uniforms (replay); nullptr draws from R
    RngDev R;
    int kind;                // FS_RNG_* of R (FS_RNG_PHILOX_REF draws like FS_RNG_PHILOX)
    double beta;
    long long* attempts;
    long long* accepted;
    unsigned char* mask;     // optional
};

// Called by ONE thread of chain b.  Returns 1 when the proposal is accepted (scalars already updated).
__device__ __forceinline__ int accept_decide(const AcceptArgs& A, int b, float e_new, float w_new) {
    const long long att = A.attempts[b];
    const double eno = A.E[b];
    const double enn = (double)e_new;
    const double nll_old = -(double)A.lq_old[b];
    const double nll_new = -(double)A.lq_new[b];
    const double ratio_log = -A.beta * (enn - eno) - (nll_new - nll_old);
    const double ratio = exp(ratio_log);          // NaN compares false on both tests below, like numpy
    int ok;
    if (ratio >= 1.0) {
        ok = 1;
    } else {
        double u;
        if (A.u_in) {
            u = A.u_in[b];
        } else if (A.kind == FS_RNG_PCG64) {
            Pcg64 g;
            g.load(A.R.pcg_state + (size_t)b * 6);
            u = g.next_double();
            g.store(A.R.pcg_state + (size_t)b * 6);
        } else if (A.kind == FS_RNG_PHILOX || A.kind == FS_RNG_PHILOX_REF) {
            uint2 key = make_uint2((uint32_t)A.R.philox_seed, (uint32_t)(A.R.philox_seed >> 32));
            long long cid = A.R.chain_id0 + b;
            uint4 ctr = make_uint4((uint32_t)att, (uint32_t)((unsigned long long)att >> 32), (uint32_t)cid,
                                   (uint32_t)((unsigned long long)cid >> 32));
            uint4 r = philox4x32(ctr, key);
            u = (double)r.w * (1.0 / 4294967296.0);
        } else {
            int cu = A.R.replay_cursor[2 * b + 1];
            u = A.R.replay_u[(size_t)b * A.R.u_stride + cu];
            A.R.replay_cursor[2 * b + 1] = cu + 1;
        }
        ok = u < ratio ? 1 : 0;
    }
    A.attempts[b] = att + 1;
    if (ok) {
        A.accepted[b] += 1;
        A.E[b] = enn;
        A.W[b] = (double)w_new;
    }
    if (A.mask) A.mask[b] = (unsigned char)ok;
    return ok;
}

// Host side: argument checks + AcceptArgs shared by fs_accept_global and fs_accept_global_fused.
// Returns FS_OK or an error code (message set).
inline int make_accept_args(const char* who, double* E, double* W, const float* lq_old, const float* lq_new,
                            const double* u, const fs_rng* rng, double beta, long long* attempts,
                            long long* accepted, unsigned char* mask, AcceptArgs* A) {
    if (!E || !W || !lq_old || !lq_new || !attempts || !accepted || (!u && !rng)) {
        set_error("%s: invalid argument", who);
        return FS_ERR_INVALID;
    }
    int kind = FS_RNG_REPLAY;
    if (rng) {
        A->R = make_rng(rng);
        kind = rng->kind;
    } else {
        A->R = RngDev();
        A->R.kind = FS_RNG_REPLAY;
    }
    if (!u) {
        if (kind == FS_RNG_PCG64 && !rng->pcg_state) { set_error("%s: pcg_state is NULL", who); return FS_ERR_INVALID; }
        if (kind == FS_RNG_REPLAY && (!rng->replay_u || !rng->replay_cursor)) { set_error("%s: replay buffers are NULL", who); return FS_ERR_INVALID; }
    }
    A->E = E; A->W = W; A->lq_old = lq_old; A->lq_new = lq_new; A->u_in = u; A->kind = kind; A->beta = beta;
    A->attempts = attempts; A->accepted = accepted; A->mask = mask;
    return FS_OK;
}

}  // namespace fs
