// vmgym_optim.cu — the optimiser step of PPOAgent.update (src/agents/ppo.py:143,284-287) on flat fp32 buffers:
//   nn.utils.clip_grad_norm_(parameters, max_grad_norm)      -> global 2-norm, coefficient min(1, max_norm / (norm + 1e-6))
//   torch.optim.AdamW(lr, betas, eps, weight_decay).step()    -> decoupled decay, bias-corrected moments
// Everything the host would have to read back to drive it lives on the device: the gradient norm, the step counter and a
// SKIP flag (the KL early stop of ppo.py:263-264 decided on the device: a skipped step changes nothing, not even the
// moments or the step counter), so a minibatch loop enqueues without a host synchronisation.
// Deterministic: the norm is reduced in a fixed order (per-block partials, then the same tree in every block).
// Compiled with -fmad=false like the rest of the library.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "vmgym.h"

extern "C" void vmgym_internal_set_error(const char* msg);

namespace vmgym_optim {

constexpr int NORM_BLOCKS = 1024;      // partial sums of the squared gradient norm
constexpr int THREADS = 256;

__device__ __forceinline__ double block_sum(double x, double* sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = x;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) sh[0] = t;
    }
    __syncthreads();
    return sh[0];
}

// partials[b] = sum over the block's grid-stride share of (grad * grad_scale)^2, in fp64
__global__ void __launch_bounds__(THREADS) sqnorm_kernel(const float* __restrict__ grad, long long n, float grad_scale,
                                                         double* __restrict__ partials)
{
    __shared__ double sh[32];
    double acc = 0.0;
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 g = g4[i];
        const double a = (double)g.x * grad_scale, b = (double)g.y * grad_scale, c = (double)g.z * grad_scale, d = (double)g.w * grad_scale;
        acc += a * a + b * b + c * c + d * d;
    }
    if (blockIdx.x == 0) {
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
            const double a = (double)grad[i] * grad_scale;
            acc += a * a;
        }
    }
    const double s = block_sum(acc, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

struct AdamArgs {
    float* param;
    const float* grad;
    float* m;
    float* v;
    long long n;
    float lr, beta1, beta2, eps, weight_decay, max_grad_norm, grad_scale;
    const double* partials;
    const int32_t* skip;          // nullptr = never skip
    int32_t* step;                // optimiser step counter (device), incremented by finish_kernel when not skipped
    float* grad_norm_out;         // optional: the (scaled) total norm of this call
};

__global__ void __launch_bounds__(THREADS) adamw_kernel(const AdamArgs a)
{
    __shared__ double sh[32];
    if (a.skip && *a.skip != 0) return;
    // total norm: every block reduces the same partials in the same order -> the same clip coefficient everywhere
    double acc = 0.0;
    for (int i = threadIdx.x; i < NORM_BLOCKS; i += blockDim.x) acc += a.partials[i];
    const double total = sqrt(block_sum(acc, sh));
    float coef = 1.0f;
    if (a.max_grad_norm > 0.0f) {
        const double c = (double)a.max_grad_norm / (total + 1e-6);          // torch.nn.utils.clip_grad_norm_
        coef = c < 1.0 ? (float)c : 1.0f;
    }
    const int t = *a.step + 1;
    const double bc1 = 1.0 - pow((double)a.beta1, (double)t), bc2 = 1.0 - pow((double)a.beta2, (double)t);
    const float step_size = (float)((double)a.lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float decay = 1.0f - a.lr * a.weight_decay;
    const float gs = a.grad_scale * coef;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        const float g = a.grad[i] * gs;
        float p = a.param[i] * decay;                                        // param.mul_(1 - lr * weight_decay)
        float m = a.m[i];
        m = m + (g - m) * (1.0f - a.beta1);                                  // exp_avg.lerp_(grad, 1 - beta1)
        const float v = a.v[i] * a.beta2 + (g * g) * (1.0f - a.beta2);       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(v) / bc2_sqrt + a.eps;
        p = p - step_size * (m / denom);                                     // param.addcdiv_(exp_avg, denom, value=-step_size)
        a.param[i] = p; a.m[i] = m; a.v[i] = v;
    }
}

__global__ void finish_kernel(const AdamArgs a)
{
    double acc = 0.0;
    for (int i = 0; i < NORM_BLOCKS; i++) acc += a.partials[i];
    if (a.grad_norm_out) *a.grad_norm_out = (float)sqrt(acc);
    if (!(a.skip && *a.skip != 0)) *a.step += 1;
}

}  // namespace vmgym_optim

extern "C" int vmgym_adamw_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, float max_grad_norm, float grad_scale,
                                double* d_workspace, const int32_t* d_skip, int32_t* d_step, float* d_grad_norm_out, void* stream)
{
    using namespace vmgym_optim;
    if (!d_param || !d_grad || !d_exp_avg || !d_exp_avg_sq || !d_workspace || !d_step || n < 0) {
        vmgym_internal_set_error("vmgym_adamw_step: null operand");
        return VMGYM_EINVAL;
    }
    if (((uintptr_t)d_grad & 15) != 0) {
        vmgym_internal_set_error("vmgym_adamw_step: the gradient buffer must be 16-byte aligned");
        return VMGYM_EINVAL;
    }
    if (n == 0) return VMGYM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    sqnorm_kernel<<<NORM_BLOCKS, THREADS, 0, st>>>(d_grad, n, grad_scale, d_workspace);
    AdamArgs a;
    a.param = d_param; a.grad = d_grad; a.m = d_exp_avg; a.v = d_exp_avg_sq; a.n = n;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_grad_norm = max_grad_norm;
    a.grad_scale = grad_scale; a.partials = d_workspace; a.skip = d_skip; a.step = d_step; a.grad_norm_out = d_grad_norm_out;
    long long blocks = (n + THREADS * 4 - 1) / (THREADS * 4);
    if (blocks > 148 * 8) blocks = 148 * 8;
    adamw_kernel<<<(unsigned)blocks, THREADS, 0, st>>>(a);
    finish_kernel<<<1, 1, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { vmgym_internal_set_error(cudaGetErrorString(e)); return VMGYM_ECUDA; }
    return VMGYM_OK;
}
