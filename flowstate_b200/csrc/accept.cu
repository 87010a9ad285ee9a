// Flow-augmented Metropolis accept/reject of NF-proposed global moves.
//
// fs_accept_global <- steps 3-5 of MonteCarlo.nf_big_move (MCMC/monte_carlo.py:264-303), decision rule in accept.cuh.
// One warp per chain: every lane first issues the loads of its share of the proposal (up to PRE float2 per lane in
// registers), so the 8N bytes are in flight while lane 0 walks the decision chain (5 scalar loads -> float64 exp ->
// uniform); accepted chains then only store.  HBM-bound: 8N bytes read + 8N written per accepted chain.
// (fs_accept_global_fused in energy.cu evaluates the proposal's energy in the same kernel.)
#include "accept.cuh"

namespace fs {

constexpr int PRE = 8;    // float2 per lane held in registers: N <= 256 needs no second pass over the proposal

__global__ void __launch_bounds__(256) accept_global_kernel(float* __restrict__ pos, const float* __restrict__ prop,
                                                            const float* __restrict__ E_new,
                                                            const float* __restrict__ W_new, AcceptArgs A, int B,
                                                            int N) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const float2* src = reinterpret_cast<const float2*>(prop) + (size_t)b * N;
    float2* dst = reinterpret_cast<float2*>(pos) + (size_t)b * N;
    float2 buf[PRE];
#pragma unroll
    for (int j = 0; j < PRE; ++j) {
        const int i = lane + 32 * j;
        if (i < N) buf[j] = __ldg(src + i);
    }
    int ok = 0;
    if (lane == 0) ok = accept_decide(A, b, E_new[b], W_new[b]);
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (ok) {
#pragma unroll
        for (int j = 0; j < PRE; ++j) {
            const int i = lane + 32 * j;
            if (i < N) dst[i] = buf[j];
        }
        for (int i = lane + 32 * PRE; i < N; i += 32) dst[i] = __ldg(src + i);
    }
}

}  // namespace fs

extern "C" int fs_accept_global(float* pos, const float* prop, double* E, double* W, const float* E_new,
                                const float* W_new, const float* logq_old, const float* logq_new, const double* u,
                                const fs_rng* rng, double beta, long long* attempts, long long* accepted,
                                unsigned char* accept_mask, int B, int N, void* stream) {
    if (!pos || !prop || !E_new || !W_new || B < 0 || N < 1) {
        fs::set_error("fs_accept_global: invalid argument");
        return FS_ERR_INVALID;
    }
    fs::AcceptArgs A;
    int r = fs::make_accept_args("fs_accept_global", E, W, logq_old, logq_new, u, rng, beta, attempts, accepted,
                                 accept_mask, &A);
    if (r != FS_OK) return r;
    if (B == 0) return FS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int grid = (B + 7) / 8;
    fs::accept_global_kernel<<<grid, 256, 0, s>>>(pos, prop, E_new, W_new, A, B, N);
    fs::count_launch();
    return fs::cuda_check(cudaGetLastError(), "accept_global_kernel");
}
