// Flow-augmented Metropolis accept/reject of NF-proposed global moves.
//
// fs_accept_global <- steps 3-5 of MonteCarlo.nf_big_move (MCMC/monte_carlo.py:264-303):
//   ratio_log = -beta (E_new - E_old) - (nll_new - nll_old); ratio = exp(ratio_log)
//   accept if ratio >= 1, else draw ONE uniform and accept if u < ratio
//   accept: particles <- proposal, accepted += 1, cached energy <- E_new
//   attempts += 1 in every case (:240).
// One warp per chain: lane 0 decides, all lanes copy the proposal (HBM-bound:
// 8N bytes read + 8N written per accepted chain).
#include "common.cuh"

namespace fs {

template <int KIND>
__global__ void __launch_bounds__(256) accept_global_kernel(
    float* __restrict__ pos, const float* __restrict__ prop, double* __restrict__ E, double* __restrict__ W,
    const float* __restrict__ E_new, const float* __restrict__ W_new, const float* __restrict__ lq_old,
    const float* __restrict__ lq_new, const double* __restrict__ u_in, RngDev R, double beta,
    long long* __restrict__ attempts, long long* __restrict__ accepted, unsigned char* __restrict__ mask, int B,
    int N) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    int ok = 0;
    if (lane == 0) {
        const long long att = attempts[b];
        const double eno = E[b];
        const double enn = (double)E_new[b];
        const double nll_old = -(double)lq_old[b];
        const double nll_new = -(double)lq_new[b];
        const double ratio_log = -beta * (enn - eno) - (nll_new - nll_old);
        const double ratio = exp(ratio_log);      // NaN compares false on both tests below, like numpy
        if (ratio >= 1.0) {
            ok = 1;
        } else {
            double u;
            if (u_in) {
                u = u_in[b];
            } else if (KIND == FS_RNG_PCG64) {
                Pcg64 g;
                g.load(R.pcg_state + (size_t)b * 6);
                u = g.next_double();
                g.store(R.pcg_state + (size_t)b * 6);
            } else if (KIND == FS_RNG_PHILOX) {
                uint2 key = make_uint2((uint32_t)R.philox_seed, (uint32_t)(R.philox_seed >> 32));
                long long cid = R.chain_id0 + b;
                uint4 ctr = make_uint4((uint32_t)att, (uint32_t)((unsigned long long)att >> 32), (uint32_t)cid,
                                       (uint32_t)((unsigned long long)cid >> 32));
                uint4 r = philox4x32(ctr, key);
                u = (double)r.w * (1.0 / 4294967296.0);
            } else {
                int cu = R.replay_cursor[2 * b + 1];
                u = R.replay_u[(size_t)b * R.u_stride + cu];
                R.replay_cursor[2 * b + 1] = cu + 1;
            }
            ok = u < ratio ? 1 : 0;
        }
        attempts[b] = att + 1;
        if (ok) {
            accepted[b] += 1;
            E[b] = enn;
            W[b] = (double)W_new[b];
        }
        if (mask) mask[b] = (unsigned char)ok;
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (ok) {
        const float2* src = reinterpret_cast<const float2*>(prop) + (size_t)b * N;
        float2* dst = reinterpret_cast<float2*>(pos) + (size_t)b * N;
        for (int i = lane; i < N; i += 32) dst[i] = __ldg(src + i);
    }
}

}  // namespace fs

extern "C" int fs_accept_global(float* pos, const float* prop, double* E, double* W, const float* E_new,
                                const float* W_new, const float* logq_old, const float* logq_new, const double* u,
                                const fs_rng* rng, double beta, long long* attempts, long long* accepted,
                                unsigned char* accept_mask, int B, int N, void* stream) {
    if (!pos || !prop || !E || !W || !E_new || !W_new || !logq_old || !logq_new || !attempts || !accepted ||
        B < 0 || N < 1 || (!u && !rng)) {
        fs::set_error("fs_accept_global: invalid argument");
        return FS_ERR_INVALID;
    }
    if (B == 0) return FS_OK;
    fs::RngDev R;
    int kind = FS_RNG_REPLAY;
    if (rng) {
        R = fs::make_rng(rng);
        kind = rng->kind;
    } else {
        R = fs::RngDev();
        R.kind = FS_RNG_REPLAY;
    }
    if (!u) {
        if (kind == FS_RNG_PCG64 && !rng->pcg_state) { fs::set_error("fs_accept_global: pcg_state is NULL"); return FS_ERR_INVALID; }
        if (kind == FS_RNG_REPLAY && (!rng->replay_u || !rng->replay_cursor)) { fs::set_error("fs_accept_global: replay buffers are NULL"); return FS_ERR_INVALID; }
    }
    cudaStream_t s = (cudaStream_t)stream;
    int grid = (B + 7) / 8;
#define FS_LAUNCH(K)                                                                                          \
    fs::accept_global_kernel<K><<<grid, 256, 0, s>>>(pos, prop, E, W, E_new, W_new, logq_old, logq_new, u, R, \
                                                     beta, attempts, accepted, accept_mask, B, N)
    if (kind == FS_RNG_PCG64) FS_LAUNCH(FS_RNG_PCG64);
    else if (kind == FS_RNG_PHILOX || kind == FS_RNG_PHILOX_REF) FS_LAUNCH(FS_RNG_PHILOX);
    else FS_LAUNCH(FS_RNG_REPLAY);
#undef FS_LAUNCH
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "accept_global_kernel");
}
