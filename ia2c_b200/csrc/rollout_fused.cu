// rollout_fused.cu — persistent one-launch rollout for small agent counts (N <= 8).
//
// Replaces the rollout part of the episode loop, ia2c.py:72-102, in ONE kernel launch: each thread owns
// one (env, agent) pair for the whole episode and loops over the T+1 time steps with everything
// on-chip — its actor's 105 weights in registers, the env state replicated in the lanes of the env's
// group, its K belief vectors as integer hundredths in registers.  Per step:
//   Org transition (action counts reduced by warp shuffles inside the env's lane group, fp64 reward
//   recurrence) -> observation -> actor forward + softmax + sample -> the other agents' actions fetched
//   by shuffles -> K fp64 belief updates -> partner modes; the trajectory row is streamed to HBM.
// Parameters are frozen during a rollout and envs are independent, so no inter-block communication is
// needed.  The per-step kernels in trainer.cu / belief.cu remain the general path (any N) and the
// reference for this kernel's parity tests: both must produce identical bytes.
#include "common.cuh"

namespace ia2c {
namespace {

constexpr int F = IA2C_OBS_FEATURES, A = IA2C_AGENT_ACTIONS;
constexpr int kThreads = 64;

__device__ __forceinline__ uint32_t pack_count(int a) { return a == 0 ? 1u : (a == 1 ? (1u << 10) : (1u << 20)); }
__device__ __forceinline__ int mode3(int c0, int c1, int c2) {
    int best = 0, bc = c0;
    if (c1 > bc) { best = 1; bc = c1; }
    if (c2 > bc) { best = 2; }
    return best;
}

template <int N, int M>
__global__ void __launch_bounds__(kThreads) rollout_fused_kernel(ia2c_episode_desc d) {
    constexpr int K = N - 1;
    constexpr int G = N <= 2 ? 2 : (N <= 4 ? 4 : 8);     // lanes per env
    constexpr int EPW = 32 / G;
    __shared__ double tab[101];                          // k / 100, correctly rounded
    __shared__ double fa_s[N * M * A];
    for (int k = threadIdx.x; k <= 100; k += blockDim.x) tab[k] = __ddiv_rn((double)k, 100.0);
    for (int k = threadIdx.x; k < N * M * A; k += blockDim.x) fa_s[k] = d.filter_action[k];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int sub = lane & (G - 1);
    const int leader = lane & ~(G - 1);
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t e = warp_global * EPW + (lane / G);
    const int64_t E = d.E;
    const bool live = e < E;            // uniform over the env's lane group
    const bool agent = live && sub < N; // this lane owns agent `sub`
    const int i = sub < N ? sub : 0;

    float w[kActorP];
#pragma unroll
    for (int k = 0; k < kActorP; ++k) w[k] = d.actor_params[i * kActorP + k];
    const double* fa = fa_s + i * M * A;

    // env state, replicated in every lane of the group
    int s = 2, prev_cls = 1, cur_cls = 1, elapsed = 0;
    double hist = 0.0, ep_ret = 0.0;
    int bel[K][M];
    const int prior_k = (int)rint(100.0 / M);            // round(1/M, 2) in hundredths (Q12)
#pragma unroll
    for (int jj = 0; jj < K; ++jj)
#pragma unroll
        for (int m = 0; m < M; ++m) bel[jj][m] = prior_k;
    int a = 0;
    int last_pred[K];
#pragma unroll
    for (int jj = 0; jj < K; ++jj) last_pred[jj] = 0;

    for (int t = 0; t <= d.T; ++t) {
        // ---- Org step from the actions of t-1 (all lanes of the group compute it redundantly)
        if (t > 0) {
            uint32_t packed = (sub < N) ? pack_count(a) : 0u;
#pragma unroll
            for (int off = G >> 1; off > 0; off >>= 1) packed += __shfl_xor_sync(0xffffffffu, packed, off);
            int s2;
            double base;
            org_transition(s, packed & 1023, (packed >> 10) & 1023, (packed >> 20) & 1023, N, s2, base);
            double r = org_reward(base, hist);
            prev_cls = cur_cls;
            cur_cls = org_obs_class(s2);
            ep_ret += r;
            if (live && sub == 0) {
                const int64_t o = (int64_t)(t - 1) * E + e;
                d.reward[o] = (float)r;
                if (d.state_trace) d.state_trace[o] = s2;
                if (d.reward_f64) d.reward_f64[o] = r;
            }
            ++elapsed;
            if (d.max_episode_steps > 0 && elapsed >= d.max_episode_steps) {  // same-step autoreset (Q14)
                s2 = 2; r = 0.0; prev_cls = 1; cur_cls = 1; elapsed = 0;
            }
            s = s2;
            hist = r;
        }
        float x[F];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            x[k] = (k == prev_cls) ? 1.f : 0.f;
            x[3 + k] = (k == cur_cls) ? 1.f : 0.f;
        }
        if (live) {
            float* o = d.obs + ((int64_t)t * E + e) * F;
#pragma unroll
            for (int k = 0; k < F; ++k)
                if ((k % G) == sub) o[k] = x[k];
        }
        // ---- own actor: forward, softmax, sample
        const int64_t row = ((int64_t)t * E + e) * N + i;
        if (agent && d.inj_actions) {
            a = d.inj_actions[row];
        } else {
            float h1[H], h2[H], y[A];
            mlp_forward<F, A>(w, x, h1, h2, y);
            softmax_inplace<A>(y);
            float u = 0.f;
            if (agent)
                u = d.inj_u_action ? d.inj_u_action[row]
                                   : philox_uniform_f32(d.seed, kStreamAction, d.episode, (uint32_t)t,
                                                        (uint64_t)((d.env_offset + e) * N + i));
            a = sample_inverse_cdf<A>(y, u);
        }
        if (agent) d.act[row] = (uint8_t)a;
        uint32_t packed = (sub < N) ? pack_count(a) : 0u;
#pragma unroll
        for (int off = G >> 1; off > 0; off >>= 1) packed += __shfl_xor_sync(0xffffffffu, packed, off);
        // ---- beliefs over the K modelled others (ascending agent order, skipping self)
        int pc0 = 0, pc1 = 0, pc2 = 0;
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
            const int j = jj + (jj >= i);
            const int seen = __shfl_sync(0xffffffffu, a, leader + j);
            double prev[M], lik[A], b[M], pred[A];
#pragma unroll
            for (int m = 0; m < M; ++m) prev[m] = tab[bel[jj][m]];
#pragma unroll
            for (int k = 0; k < A; ++k) lik[k] = (k == seen) ? 0.8 : 0.1;   // ia2c.py:53-58
            const int64_t rec = (((int64_t)t * E + e) * N + i) * K + jj;
            double u = 0.0;
            if (agent)
                u = d.inj_u_belief ? d.inj_u_belief[rec]
                                   : philox_uniform_f64(d.seed, kStreamBelief, d.episode, (uint32_t)t,
                                                        (uint64_t)(((d.env_offset + e) * N + i) * (int64_t)K + jj));
            // belief update, operation order of SURVEY.md Appendix A.2 (kept in sync with belief.cu)
            double bp[M];
#pragma unroll
            for (int m = 0; m < M; ++m) {
                double acc = __dmul_rn(lik[0], __dmul_rn(fa[m * A + 0], prev[m]));
#pragma unroll
                for (int k = 1; k < A; ++k) acc = __dadd_rn(acc, __dmul_rn(lik[k], __dmul_rn(fa[m * A + k], prev[m])));
                bp[m] = acc;
            }
            double S = bp[0];
#pragma unroll
            for (int m = 1; m < M; ++m) S = __dadd_rn(S, bp[m]);
            const double rS = drcp_seq(S);
#pragma unroll
            for (int m = 0; m < M; ++m) b[m] = ddiv_with(bp[m], S, rS);
#pragma unroll
            for (int k = 0; k < A; ++k) {
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < M; ++m) acc = __dadd_rn(acc, __dmul_rn(b[m], fa[m * A + k]));
                pred[k] = acc;
            }
            double c = pred[0];
            int ap = 0;
            bool found = u < c;
#pragma unroll
            for (int k = 1; k < A; ++k) {
                c = __dadd_rn(c, pred[k]);
                if (!found && u < c) { ap = k; found = true; }
            }
#pragma unroll
            for (int m = 0; m < M; ++m) bel[jj][m] = __double2int_rn(__dmul_rn(b[m], 100.0));
            last_pred[jj] = ap;
            pc0 += (ap == 0); pc1 += (ap == 1); pc2 += (ap == 2);
            if (agent && d.pred_dump) d.pred_dump[rec] = (uint8_t)ap;
            if (agent && d.belief_dump) {
#pragma unroll
                for (int m = 0; m < M; ++m) d.belief_dump[rec * M + m] = (uint8_t)bel[jj][m];
            }
        }
        if (agent) {
            const int c0 = packed & 1023, c1 = (packed >> 10) & 1023, c2 = (packed >> 20) & 1023;
            d.partner_true[row] = (uint8_t)mode3(c0 - (a == 0), c1 - (a == 1), c2 - (a == 2));
            d.partner_pred[row] = (uint8_t)mode3(pc0, pc1, pc2);
        }
    }
    // ---- persist the final env / belief state exactly as the per-step path leaves it
    if (live && sub == 0) {
        d.env_state[e] = s;
        d.env_hist[e] = hist;
        d.env_elapsed[e] = elapsed;
        d.ep_return[e] = ep_ret;
        *reinterpret_cast<uchar2*>(d.env_cls + 2 * e) = make_uchar2((unsigned char)prev_cls, (unsigned char)cur_cls);
    }
    if (agent) {
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                if (m < 4) lo |= (uint32_t)bel[jj][m] << (8 * m); else hi |= (uint32_t)bel[jj][m] << (8 * (m - 4));
            }
            hi |= (uint32_t)last_pred[jj] << 16;
            *reinterpret_cast<uint2*>(d.belief_records + ((e * N + i) * (int64_t)K + jj) * IA2C_BELIEF_RECORD) = make_uint2(lo, hi);
        }
    }
}

template <int N, int M>
int launch(const ia2c_episode_desc* d, cudaStream_t s) {
    constexpr int G = N <= 2 ? 2 : (N <= 4 ? 4 : 8);
    const int64_t threads = ((d->E + (32 / G) - 1) / (32 / G)) * 32;
    rollout_fused_kernel<N, M><<<ceil_div(threads, kThreads), kThreads, 0, s>>>(*d);
    return check_launch("rollout_fused_kernel");
}

}  // namespace

// Returns 1 if (N, M) has a fused instantiation.
int rollout_fused_supported(int N, int M) { return (M == 5 && N >= 2 && N <= 8) || (M == 3 && N == 2); }

int rollout_fused_launch(const ia2c_episode_desc* d, cudaStream_t s) {
    if (d->M == 3 && d->N == 2) return launch<2, 3>(d, s);
    switch (d->N) {
        case 2: return launch<2, 5>(d, s);
        case 3: return launch<3, 5>(d, s);
        case 4: return launch<4, 5>(d, s);
        case 5: return launch<5, 5>(d, s);
        case 6: return launch<6, 5>(d, s);
        case 7: return launch<7, 5>(d, s);
        case 8: return launch<8, 5>(d, s);
    }
    set_error("rollout_fused: no instantiation for N=%d M=%d", d->N, d->M);
    return IA2C_ERR_UNSUPPORTED;
}

}  // namespace ia2c
