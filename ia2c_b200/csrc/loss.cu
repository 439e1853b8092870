// loss.cu — critic / actor losses on network outputs (class API) and the Adam step.
//
// Replaces, for the drop-in classes:
//   CriticNetwork.batch_update's one-hot select + nn.MSELoss + its backward   (ac_nets.py:64-71)
//   ActorNetwork.batch_update's Categorical log-prob / entropy loss + backward  (ac_nets.py:113-118)
//   torch.optim.Adam.step, single-tensor path, default betas/eps                (ac_nets.py:50,72,89,119)
// Gradients are the closed forms of SURVEY.md Appendix A.4 (checked against autograd in the oracle).
#include "common.cuh"

namespace ia2c {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = kSMs * 4;
constexpr int kMaxO = 32;
constexpr float kEpsClamp = 1.1920928955078125e-07f;  // torch.finfo(float32).eps used by clamp_probs

__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;  // valid on thread 0
}

__global__ void __launch_bounds__(kThreads)
critic_loss_kernel(const float* __restrict__ Q, const int32_t* __restrict__ act, const float* __restrict__ target,
                   float* __restrict__ dQ, float* __restrict__ dtarget, float* __restrict__ partial, int64_t B, int O) {
    __shared__ float red[kThreads / 32];
    const float inv_b = 1.f / (float)B;
    float acc = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
        const int a = act[r];
        const float delta = target[r] - Q[r * O + a];
        acc = fmaf(delta, delta, acc);
        const float g = 2.f * delta * inv_b;
        for (int o = 0; o < O; ++o) dQ[r * O + o] = (o == a) ? -g : 0.f;
        if (dtarget) dtarget[r] = g;
    }
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kThreads)
actor_loss_kernel(const float* __restrict__ probs, const int32_t* __restrict__ act, const float* __restrict__ adv,
                  float beta, float* __restrict__ dprobs, float* __restrict__ dadv, int32_t* __restrict__ status,
                  float* __restrict__ partial, int64_t B, int O) {
    __shared__ float red[kThreads / 32];
    const float inv_b = 1.f / (float)B;
    float acc = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
        float p[kMaxO];
        float s = 0.f;
        bool bad = false;
        for (int o = 0; o < O; ++o) {
            p[o] = probs[r * O + o];
            s += p[o];
            bad |= !(p[o] >= 0.f);
        }
        bad |= !(fabsf(s - 1.f) < 1e-6f);  // Categorical's simplex validation (raises in the reference)
        if (bad && status) *status = 1;
        const int a = act[r];
        const float ad = adv[r];
        // q = p/s; logit = log(clamp(q)); H = -sum(logit*q); g = dL/dq
        float ent = 0.f, qg = 0.f, neglogp = 0.f;
        float g[kMaxO];
        for (int o = 0; o < O; ++o) {
            const float q = p[o] / s;
            const bool inside = (q >= kEpsClamp) && (q <= 1.f - kEpsClamp);
            const float logit = logf(fminf(fmaxf(q, kEpsClamp), 1.f - kEpsClamp));
            ent -= logit * q;
            float go = beta * (logit + (inside ? 1.f : 0.f));
            if (o == a) {
                neglogp = -logit;
                if (inside) go -= ad / q;
            }
            g[o] = go;
            qg = fmaf(q, go, qg);
            p[o] = q;
        }
        acc += ad * neglogp - beta * ent;
        // dL/dp_k = (g_k - sum_j q_j g_j) / s   (normalisation q = p/s), times 1/B of the mean
        for (int o = 0; o < O; ++o) dprobs[r * O + o] = (g[o] - qg) / s * inv_b;
        if (dadv) dadv[r] = neglogp * inv_b;
    }
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void finish_mean_kernel(const float* __restrict__ partial, int n, float inv_b, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += partial[i];
        *out = s * inv_b;
    }
}

// Adam, one thread per parameter.  Arithmetic order follows torch's _single_tensor_adam:
//   m = m + (g - m)*(1-b1);  v = v*b2 + (1-b2)*g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(kThreads)
adam_kernel(float* __restrict__ params, const float* __restrict__ grad, float* __restrict__ grad_accum,
            float* __restrict__ m, float* __restrict__ v, const int32_t* __restrict__ step_count, double lr, double b1,
            double b2, double eps_d, int P) {
    const int net = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int t = step_count[net] + 1;
    const double bc1 = 1.0 - pow(b1, (double)t);
    const double bc2 = 1.0 - pow(b2, (double)t);
    const float step_size = (float)(lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float w1 = (float)(1.0 - b1), w2 = (float)(1.0 - b2), b2f = (float)b2, eps = (float)eps_d;
    const int64_t k = (int64_t)net * P + i;
    float g = grad[k];
    if (grad_accum) {
        g += grad_accum[k];
        grad_accum[k] = g;
    }
    const float mi = m[k] + (g - m[k]) * w1;
    const float vi = v[k] * b2f + w2 * g * g;
    m[k] = mi;
    v[k] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    params[k] = params[k] - step_size * (mi / denom);
}
__global__ void adam_bump_kernel(int32_t* step_count, int nets) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nets) step_count[i] += 1;
}

}  // namespace
}  // namespace ia2c

using namespace ia2c;

extern "C" size_t ia2c_loss_workspace(int64_t B) {
    (void)B;
    return kMaxBlocks;
}

static int loss_blocks(int64_t B) {
    int64_t b = (B + kThreads - 1) / kThreads;
    return (int)(b < 1 ? 1 : (b > kMaxBlocks ? kMaxBlocks : b));
}

extern "C" int ia2c_critic_loss(const float* Q, const int32_t* act, const float* target, float* loss_out, float* dQ,
                                float* dtarget, float* workspace, int64_t B, int32_t O, void* stream) {
    IA2C_REQUIRE(Q && act && target && loss_out && dQ && workspace && B > 0, "ia2c_critic_loss: null pointer or B=%lld", (long long)B);
    IA2C_REQUIRE(O >= 1 && O <= kMaxO, "ia2c_critic_loss: O=%d outside 1..%d", O, kMaxO);
    const int blocks = loss_blocks(B);
    cudaStream_t s = as_stream(stream);
    critic_loss_kernel<<<blocks, kThreads, 0, s>>>(Q, act, target, dQ, dtarget, workspace, B, O);
    int rc = check_launch("critic_loss_kernel");
    if (rc) return rc;
    finish_mean_kernel<<<1, 32, 0, s>>>(workspace, blocks, 1.f / (float)B, loss_out);
    return check_launch("finish_mean_kernel");
}

extern "C" int ia2c_actor_loss(const float* probs, const int32_t* act, const float* adv, float beta, float* loss_out,
                               float* dprobs, float* dadv, int32_t* status_out, float* workspace, int64_t B,
                               int32_t O, void* stream) {
    IA2C_REQUIRE(probs && act && adv && loss_out && dprobs && workspace && B > 0, "ia2c_actor_loss: null pointer or B=%lld", (long long)B);
    IA2C_REQUIRE(O >= 1 && O <= kMaxO, "ia2c_actor_loss: O=%d outside 1..%d", O, kMaxO);
    const int blocks = loss_blocks(B);
    cudaStream_t s = as_stream(stream);
    actor_loss_kernel<<<blocks, kThreads, 0, s>>>(probs, act, adv, beta, dprobs, dadv, status_out, workspace, B, O);
    int rc = check_launch("actor_loss_kernel");
    if (rc) return rc;
    finish_mean_kernel<<<1, 32, 0, s>>>(workspace, blocks, 1.f / (float)B, loss_out);
    return check_launch("finish_mean_kernel");
}

extern "C" int ia2c_adam_step(float* params, const float* grad, float* grad_accum, float* exp_avg, float* exp_avg_sq,
                              int32_t* step_count, double lr, double beta1, double beta2, double eps, int32_t nets,
                              int32_t P, void* stream) {
    IA2C_REQUIRE(params && grad && exp_avg && exp_avg_sq && step_count, "ia2c_adam_step: null pointer");
    IA2C_REQUIRE(nets >= 1 && P >= 1, "ia2c_adam_step: nets=%d P=%d", nets, P);
    cudaStream_t s = as_stream(stream);
    dim3 grid(ceil_div(P, kThreads), nets);
    adam_kernel<<<grid, kThreads, 0, s>>>(params, grad, grad_accum, exp_avg, exp_avg_sq, step_count, lr, beta1, beta2, eps, P);
    int rc = check_launch("adam_kernel");
    if (rc) return rc;
    adam_bump_kernel<<<ceil_div(nets, 256), 256, 0, s>>>(step_count, nets);
    return check_launch("adam_bump_kernel");
}
