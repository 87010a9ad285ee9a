// Adam over ONE flat parameter buffer.
//
// fs_adam_step <- torch.optim.Adam(model.parameters(), lr, weight_decay).step() as the drivers call it after every
// minibatch (hybrid_NF_MCMC/main_algorithm_1.py:297-320, main_algorithm_2.py:440-451): L2 weight decay added to the
// gradient, exponential moving averages, bias corrections, p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps).
// drivers/training.FlowTrainer keeps every trainable parameter, its gradient and both moments as views of four flat
// float32 buffers, so the whole optimizer step is two launches (torch's multi-tensor Adam walks ~700 tensors on the
// host first: ~1.5 ms of Python per step for the Algorithm-2 flow, as long as its forward + backward kernels).
// The reference's "skip the step when the loss is NaN / Inf" (main_algorithm_2.py:449) is taken ON THE DEVICE: the
// prepare kernel reads the loss (and an optional flag all-reduced over the ranks) and the apply kernel returns at once
// when the step is skipped - the host never waits for the loss.
// state[4] (device): [0] steps applied so far, [1] 1 if the last call applied its step, [2] lr / bc1, [3] 1 / sqrt(bc2).
#include "common.cuh"

namespace fs {

__global__ void adam_prepare_kernel(float* __restrict__ st, const float* __restrict__ skip, const float* __restrict__ loss,
                                    float lr, float beta1, float beta2) {
    bool bad = false;
    if (skip) bad = bad || !(*skip == 0.0f);
    if (loss) bad = bad || !isfinite(*loss);
    if (bad) {
        st[1] = 0.0f;
        return;
    }
    const float step = st[0] + 1.0f;
    st[0] = step;
    st[1] = 1.0f;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    st[2] = (float)((double)lr / bc1);
    st[3] = (float)(1.0 / sqrt(bc2));
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float step_size, float inv_bc2_sqrt,
                                         float beta1, float beta2, float eps, float wd) {
    g = __fmaf_rn(wd, p, g);
    m = __fmaf_rn(1.0f - beta1, g - m, m);                   // exp_avg.lerp_(grad, 1 - beta1)
    v = __fmaf_rn(beta2, v, (1.0f - beta2) * g * g);
    const float denom = __fmaf_rn(sqrtf(v), inv_bc2_sqrt, eps);
    p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_apply_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v, long long n,
                                                         const float* __restrict__ st, float beta1, float beta2,
                                                         float eps, float wd) {
    if (st[1] == 0.0f) return;
    const float step_size = st[2], inv_bc2_sqrt = st[3];
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        const float4 G = reinterpret_cast<const float4*>(g)[i];
        adam_one(P.x, G.x, M.x, V.x, step_size, inv_bc2_sqrt, beta1, beta2, eps, wd);
        adam_one(P.y, G.y, M.y, V.y, step_size, inv_bc2_sqrt, beta1, beta2, eps, wd);
        adam_one(P.z, G.z, M.z, V.z, step_size, inv_bc2_sqrt, beta1, beta2, eps, wd);
        adam_one(P.w, G.w, M.w, V.w, step_size, inv_bc2_sqrt, beta1, beta2, eps, wd);
        reinterpret_cast<float4*>(p)[i] = P;
        reinterpret_cast<float4*>(m)[i] = M;
        reinterpret_cast<float4*>(v)[i] = V;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        adam_one(p[i], g[i], m[i], v[i], step_size, inv_bc2_sqrt, beta1, beta2, eps, wd);
}

}  // namespace fs

extern "C" int fs_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                            float* state, const float* skip, const float* loss, float lr, float beta1, float beta2,
                            float eps, float weight_decay, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !state || n < 0 ||
        ((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) {
        fs::set_error("fs_adam_step: invalid argument (buffers must be 16-byte aligned)");
        return FS_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    fs::adam_prepare_kernel<<<1, 1, 0, s>>>(state, skip, loss, lr, beta1, beta2);
    if (n > 0) {
        long long blocks = (n / 4 + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (blocks < 1) blocks = 1;
        fs::adam_apply_kernel<<<(int)blocks, 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n, state, beta1, beta2, eps,
                                                          weight_decay);
    }
    fs::count_launch(n > 0 ? 2 : 1);
    return fs::cuda_check(cudaGetLastError(), "adam kernels");
}
