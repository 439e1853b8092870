// actor_pipe.cu — the actor-phase gradient (ia2c.py:116-129 + ActorNetwork.batch_update, ac_nets.py:104-119)
// as a warp-specialised pipeline over time.
//
// A block owns 32 consecutive envs of ONE agent (blockIdx.y); lane = env.  next_obs[t] IS obs[t+1], so the
// UPDATED critic is evaluated once per observation.  Like the rollout (rollout_fused.cu) the work of one row is
// a dependent-latency chain (~450 instructions through critic forward -> advantage -> actor forward -> Categorical
// loss -> backward); a single warp issues one instruction every ~3.5 cycles, so the chain is cut into four stages
// that run on the four warps (= the four schedulers of the SM), each one time step behind its producer, meeting
// through small shared-memory rings and ONE __syncthreads per iteration:
//   warp 0  L  "loader"   (step it)    obs / own action / predicted partner / reward of step t: global -> rings
//                                      (register-prefetched one iteration ahead);
//           W  "W1 grad"  (row  it-4)  owns the W1/b1 gradient (dz1 handed over by warp 3).
//   warp 1  Cv "critic"   (obs  it-1)  UPDATED critic forward, weights in registers; publishes Q(obs_t)[ja_t], which is
//                                      both the baseline of row t and the bootstrap of row t-1 (ia2c.py:120-127).
//   warp 2  Pf "policy"   (obs  it-1)  actor forward + softmax + Categorical normalisation/logits; parks activations.
//   warp 3  Lb "loss/bwd" (row  it-3)  advantage, loss, output gradient, backward through W3/W2 (rows in registers);
//                                      owns the W2/b2/W3/b3 gradient.
// All 105 gradient accumulators stay in registers for the whole episode; one warp butterfly at the end, then lane 0
// writes the block's partial row.  Deterministic: fixed summation order, no atomics.
#include "common.cuh"
#include "mlp_f2.cuh"

namespace ia2c {
namespace {

constexpr int A = IA2C_AGENT_ACTIONS, J = IA2C_JOINT_ACTIONS;
constexpr int kPipeBlock = 128;
constexpr int kRing = 8;
constexpr float kEpsClamp = 1.1920928955078125e-07f;   // torch.finfo(float32).eps: Categorical clamps probs to [eps, 1-eps]

__device__ __forceinline__ int joint_index(int i, int n, int own, int other) {
    return (i < (i + 1) % n) ? own * A + other : other * A + own;
}

template <int C>
__global__ void __launch_bounds__(kPipeBlock) actor_pipe_kernel(ia2c_episode_desc d, float* __restrict__ partials) {
    constexpr int P = kActorP, W = 32 * C;      // W env columns per block; lane owns columns lane, lane+32, ...
    __shared__ float x_s[kRing][kIn][W];        // observation of step t                         (slot t & 7)
    __shared__ int ja_s[kRing][W];              // joint index (own action, predicted partner) of step t
    __shared__ int own_s[kRing][W];             // own action of step t
    __shared__ float rew_s[kRing][W];           // reward of row t
    __shared__ float v_s[kRing][W];             // Q_new(obs_t)[ja_t]
    __shared__ float2 hst_s[4][6][W];           // actor activations (h1 | h2 pairs) of obs t       (slot t & 3)
    __shared__ float ql_s[4][2 * A][W];         // normalised probs q[0..A) | clamped logits [A..2A) (slot t & 3)
    __shared__ float2 dz1_s[2][3][W];           // dL/dz1 of row t                                  (slot t & 1)

    pdl_prologue();
    const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.y, N = d.N, T = d.T;
    const int64_t E = d.E;
    int64_t ec[C];                              // dead columns read a valid one and contribute nothing
    bool live[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int64_t e = (int64_t)blockIdx.x * W + c * 32 + lane;
        live[c] = e < E;
        ec[c] = live[c] ? e : E - 1;
    }
    const int n_iter = T + 4;                   // rows 0..T-1 leave the last stage (W) at iteration T+3
    float* out = partials + ((int64_t)n * gridDim.x + blockIdx.x) * (P + 1);

    if (role == 0) {
        // ================================================================ L: loader (step it) + W: W1/b1 gradient (row it-4)
        float2 gB[21];
#pragma unroll
        for (int k = 0; k < 21; ++k) gB[k] = make_float2(0.f, 0.f);
        // two register sets, each fetched TWO iterations ahead of its hand-over: a load has two full iterations
        // (> the L2 latency) to land, so the loader never holds the block's barrier.
        struct Fetched { float x[C][kIn]; float r[C]; int a[C], pp[C]; } f0, f1;
        auto fetch = [&](Fetched& f, int t) {   // t <= T
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int64_t r = (int64_t)t * E + ec[c];
                load_obs6(d.obs + r * kIn, f.x[c]);
                f.a[c] = __ldg(d.act + r * N + n);
                f.pp[c] = __ldg(d.partner_pred + r * N + n);
                f.r[c] = t < T ? __ldg(d.reward + r) : 0.f;
            }
        };
        auto publish = [&](Fetched& f, int t) {
            const int slot = t & (kRing - 1);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int col = c * 32 + lane;
#pragma unroll
                for (int k = 0; k < kIn; ++k) x_s[slot][k][col] = f.x[c][k];
                own_s[slot][col] = f.a[c];
                ja_s[slot][col] = joint_index(n, N, f.a[c], f.pp[c]);   // ia2c.py:120-121 (and the bootstrap index of row t-1)
                rew_s[slot][col] = f.r[c];
            }
            if (t + 2 <= T) fetch(f, t + 2);
        };
        fetch(f0, 0);
        if (T >= 1) fetch(f1, 1);
        if (blockIdx.x == 0 && lane == 0 && !(d.flags & IA2C_FLAG_SKIP_ADAM)) d.actor_step[n] += 1;
        for (int it = 0; it < n_iter; ++it) {
            if (it <= T) {
                if (it & 1) publish(f1, it); else publish(f0, it);
            }
            const int tb = it - 4;
            if (tb >= 0 && tb < T) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int col = c * 32 + lane;
                    float xb[kIn];
                    float2 dz1[3];
#pragma unroll
                    for (int k = 0; k < kIn; ++k) xb[k] = x_s[tb & (kRing - 1)][k][col];
#pragma unroll
                    for (int k = 0; k < 3; ++k) dz1[k] = dz1_s[tb & 1][k][col];
                    accumulate_w1(xb, dz1, gB);
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < 21; ++k) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                gB[k].x += __shfl_xor_sync(0xffffffffu, gB[k].x, off);
                gB[k].y += __shfl_xor_sync(0xffffffffu, gB[k].y, off);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 21; ++k) { out[2 * k] = gB[k].x; out[2 * k + 1] = gB[k].y; }
        }
    } else if (role == 1) {
        // ================================================================ Cv: UPDATED critic forward, observation t = it-1
        RegNet<J> cnet;
        load_regnet<J>(cnet, d.critic_params + (int64_t)n * kCriticP);
        for (int it = 0; it < n_iter; ++it) {
            const int t = it - 1;
            if (t >= 0 && t <= T) {
                const int slot = t & (kRing - 1);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int col = c * 32 + lane;
                    float x[kIn], q[J];
#pragma unroll
                    for (int k = 0; k < kIn; ++k) x[k] = x_s[slot][k][col];
                    forward_regnet<J>(cnet, x, q);
                    v_s[slot][col] = select_out<J>(q, ja_s[slot][col]);
                }
            }
            __syncthreads();
        }
    } else if (role == 2) {
        // ================================================================ Pf: actor forward + Categorical(probs), observation t = it-1
        RegNet<A> anet;
        load_regnet<A>(anet, d.actor_params + (int64_t)n * kActorP);
        for (int it = 0; it < n_iter; ++it) {
            const int t = it - 1;
            if (t >= 0 && t < T) {
                const int slot = t & 3;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int col = c * 32 + lane;
                    float x[kIn], p[A];
                    float2 h1[3], h2[3];
#pragma unroll
                    for (int k = 0; k < kIn; ++k) x[k] = x_s[t & (kRing - 1)][k][col];
                    forward_regnet<A>(anet, x, h1, h2, p);
                    softmax_inplace<A>(p);
                    float s = 0.f;
#pragma unroll
                    for (int o = 0; o < A; ++o) s += p[o];
#pragma unroll
                    for (int o = 0; o < A; ++o) {
                        const float qv = p[o] / s;                 // Categorical(probs=p) renormalises
                        ql_s[slot][o][col] = qv;
                        ql_s[slot][A + o][col] = logf(fminf(fmaxf(qv, kEpsClamp), 1.f - kEpsClamp));
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) { hst_s[slot][k][col] = h1[k]; hst_s[slot][3 + k][col] = h2[k]; }
                }
            }
            __syncthreads();
        }
    } else {
    // ==================================================================== Lb: advantage, loss, backward, row t = it-3
    constexpr int GA = F2<A>::G2 - 21;           // float2 accumulators for flat entries [42, 106): W2 | b2 | W3 | b3 | loss slot
    RegBack<A> back;
    load_regback<A>(back, d.actor_params + (int64_t)n * kActorP);
    float2 gA[GA];
#pragma unroll
    for (int k = 0; k < GA; ++k) gA[k] = make_float2(0.f, 0.f);
    const float inv_T = 1.f / (float)((int64_t)T * d.E_total);
    const float gamma = d.gamma, beta = d.beta;
    float loss = 0.f;
    for (int it = 0; it < n_iter; ++it) {
        const int t = it - 3;
        if (t >= 0 && t < T) {
            const int s0 = t & (kRing - 1), s1 = (t + 1) & (kRing - 1), slot = t & 3;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int col = c * 32 + lane;
                const float inv_b = live[c] ? inv_T : 0.f;
                const float adv = (rew_s[s0][col] + gamma * v_s[s1][col]) - v_s[s0][col];   // ia2c.py:127
                if (d.adv_dump && live[c]) d.adv_dump[((int64_t)n * T + t) * E + ec[c]] = adv;
                const int own = own_s[s0][col];
                // loss_row = adv * (-logit[a]) - beta * H  (ac_nets.py:113-117), differentiated through clamp, normalise, softmax
                float ent = 0.f, qg = 0.f, neglogp = 0.f, gq[A], qq[A];
#pragma unroll
                for (int o = 0; o < A; ++o) qq[o] = ql_s[slot][o][col];
                const float adv_over_q = adv / select_out<A>(qq, own);   // ONE division (the taken action's), not one per unrolled o
#pragma unroll
                for (int o = 0; o < A; ++o) {
                    const float qv = qq[o], logit = ql_s[slot][A + o][col];
                    const bool inside = (qv >= kEpsClamp) && (qv <= 1.f - kEpsClamp);
                    ent -= logit * qv;
                    float go = beta * (logit + (inside ? 1.f : 0.f));
                    if (o == own) {
                        neglogp = -logit;
                        if (inside) go -= adv_over_q;
                    }
                    gq[o] = go;
                    qg = fmaf(qv, go, qg);
                }
                if (live[c]) loss += adv * neglogp - beta * ent;
                float dy[A];
#pragma unroll
                for (int o = 0; o < A; ++o) dy[o] = qq[o] * (gq[o] - qg) * inv_b;
                float2 h1[3], h2[3], dz1[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) { h1[k] = hst_s[slot][k][col]; h2[k] = hst_s[slot][3 + k][col]; }
                bwd_regback<A>(back, h1, h2, [&](int o) { return dy[o]; }, gA, dz1);
#pragma unroll
                for (int k = 0; k < 3; ++k) dz1_s[t & 1][k][col] = dz1[k];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < GA; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            gA[k].x += __shfl_xor_sync(0xffffffffu, gA[k].x, off);
            gA[k].y += __shfl_xor_sync(0xffffffffu, gA[k].y, off);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, off);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < GA; ++k) {
            if (42 + 2 * k < P) out[42 + 2 * k] = gA[k].x;
            if (42 + 2 * k + 1 < P) out[42 + 2 * k + 1] = gA[k].y;
        }
        out[P] = loss;
    }
    }
}

}  // namespace

// env columns per lane: C = 1.  Two columns per lane were measured SLOWER at every size tried (36.8 vs 26.0 us at 4096 envs x
// 2 agents): ptxas keeps the two chains sequential under the register budget of the critic-forward warp, and two co-resident
// blocks per SM overlap better than one block with twice the work.  The template parameter stays for the next attempt.
int64_t actor_pipe_blocks(int64_t E, int) { return (E + 31) / 32; }

int actor_pipe_launch(const ia2c_episode_desc* d, cudaStream_t s) {
    dim3 grid((unsigned)actor_pipe_blocks(d->E, d->N), (unsigned)d->N);
    return launch_pdl("actor_pipe_kernel", actor_pipe_kernel<1>, grid, dim3(kPipeBlock), 0, s, *d, d->partials);
}

}  // namespace ia2c
