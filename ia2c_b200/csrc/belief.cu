// belief.cu — Bayesian belief filter over the other agents' candidate models.
//
// Replaces BeliefFilter.update (belief_filter_deprecated.py:45-59) and noisy_private_obs
// (ia2c.py:53-58).  All arithmetic is fp64 with explicit round-to-nearest intrinsics so that ptxas
// cannot contract mul+add: the posterior is rounded to 2 decimals before it becomes the next prior
// (belief_filter_deprecated.py:58), so a single flipped rounding is a 0.01 error.
// Operation order per row:
//   t[a][m] = Fm[m][a]*prev[m];  bp[m] = (lik0*t0 + lik1*t1) + lik2*t2 ...;  S = ((bp0+bp1)+bp2)+...
//   b[m] = bp[m]/S;  pred[a] = sum_m b[m]*Fm[m][a];  ap = first a with u < cumsum(pred), none -> 0
//   out[m] = rint(b[m]*100)/100
//
// Two layouts:
//   dense  — the reference's fp64 [R,M]/[R,A] arrays (class API, 88 B per update at M=5,A=3);
//   pairs  — the trainer's packed records: posteriors are k/100 with integer k, stored as uint8
//            (lossless), 8 B per (env, agent, modelled other); likelihood synthesised from the other
//            agent's action; stages the models, actions and a k/100 table in shared memory and
//            reduces the predicted actions per agent (mode) with shared-memory counters.
#include <algorithm>

#include "common.cuh"

namespace ia2c {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxDense = 8;

// Core update with compile-time M, A.  fa is [M][A] (row = model).
template <int M, int A>
__device__ __forceinline__ int belief_core(const double* __restrict__ fa, const double (&lik)[A],
                                           const double (&prev)[M], double u, double (&b)[M], double (&pred)[A]) {
    double bp[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double acc = __dmul_rn(lik[0], __dmul_rn(fa[m * A + 0], prev[m]));
#pragma unroll
        for (int a = 1; a < A; ++a) acc = __dadd_rn(acc, __dmul_rn(lik[a], __dmul_rn(fa[m * A + a], prev[m])));
        bp[m] = acc;
    }
    double S = bp[0];
#pragma unroll
    for (int m = 1; m < M; ++m) S = __dadd_rn(S, bp[m]);
    const double rS = drcp_seq(S);   // one refined reciprocal shared by the M divisions (common.cuh)
#pragma unroll
    for (int m = 0; m < M; ++m) b[m] = ddiv_with(bp[m], S, rS);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < M; ++m) acc = __dadd_rn(acc, __dmul_rn(b[m], fa[m * A + a]));
        pred[a] = acc;
    }
    double c = pred[0];
    int ap = 0;
    bool found = u < c;
#pragma unroll
    for (int a = 1; a < A; ++a) {
        c = __dadd_rn(c, pred[a]);
        if (!found && u < c) {
            ap = a;
            found = true;
        }
    }
    return ap;  // falls through to 0 when u >= cumsum[-1]
}

template <int M, int A>
__global__ void __launch_bounds__(kThreads)
belief_dense_kernel(const double* __restrict__ filter_action, const double* __restrict__ lik_in,
                    const double* __restrict__ prev_in, const double* __restrict__ u_in, int64_t* __restrict__ ap_out,
                    double* __restrict__ bprime_out, double* __restrict__ pred_out, int64_t R) {
    __shared__ double fa[M * A];
    for (int i = threadIdx.x; i < M * A; i += blockDim.x) fa[i] = filter_action[i];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (r >= R) return;
    double lik[A], prev[M], b[M], pred[A];
#pragma unroll
    for (int a = 0; a < A; ++a) lik[a] = lik_in[r * A + a];
#pragma unroll
    for (int m = 0; m < M; ++m) prev[m] = prev_in[r * M + m];
    const int ap = belief_core<M, A>(fa, lik, prev, u_in[r], b, pred);
    ap_out[r] = ap;
#pragma unroll
    for (int m = 0; m < M; ++m) bprime_out[r * M + m] = ddiv_seq(rint(__dmul_rn(b[m], 100.0)), 100.0);
    if (pred_out) {
#pragma unroll
        for (int a = 0; a < A; ++a) pred_out[r * A + a] = pred[a];
    }
}

// Runtime-dimension fallback (M, A <= 8), same operation order.
__global__ void __launch_bounds__(kThreads)
belief_dense_generic_kernel(const double* __restrict__ filter_action, const double* __restrict__ lik_in,
                            const double* __restrict__ prev_in, const double* __restrict__ u_in,
                            int64_t* __restrict__ ap_out, double* __restrict__ bprime_out,
                            double* __restrict__ pred_out, int64_t R, int M, int A) {
    __shared__ double fa[kMaxDense * kMaxDense];
    for (int i = threadIdx.x; i < M * A; i += blockDim.x) fa[i] = filter_action[i];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (r >= R) return;
    double bp[kMaxDense], b[kMaxDense];
    double S = 0.0;
    for (int m = 0; m < M; ++m) {
        const double p = prev_in[r * M + m];
        double acc = __dmul_rn(lik_in[r * A], __dmul_rn(fa[m * A], p));
        for (int a = 1; a < A; ++a) acc = __dadd_rn(acc, __dmul_rn(lik_in[r * A + a], __dmul_rn(fa[m * A + a], p)));
        bp[m] = acc;
        S = (m == 0) ? acc : __dadd_rn(S, acc);
    }
    for (int m = 0; m < M; ++m) {
        b[m] = ddiv_seq(bp[m], S);
        bprime_out[r * M + m] = ddiv_seq(rint(__dmul_rn(b[m], 100.0)), 100.0);
    }
    const double u = u_in[r];
    double c = 0.0;
    int ap = 0;
    bool found = false;
    for (int a = 0; a < A; ++a) {
        double acc = 0.0;
        for (int m = 0; m < M; ++m) acc = __dadd_rn(acc, __dmul_rn(b[m], fa[m * A + a]));
        if (pred_out) pred_out[r * A + a] = acc;
        c = (a == 0) ? acc : __dadd_rn(c, acc);
        if (!found && u < c) {
            ap = a;
            found = true;
        }
    }
    ap_out[r] = ap;
}

// ------------------------------------------------------------------------------------------------
// Packed pairwise update.  A block owns `n_envs` whole envs (small N*K) or one env's agent chunk
// [i0, i1) (large N*K); within the block consecutive threads take consecutive records (coalesced
// 8-byte loads/stores).
struct PairsArgs {
    uint8_t* records;
    const double* filter_action;   // [N,M,3]
    const uint8_t* actions;        // [E,N]
    const double* u_injected;      // [E,N,K] or null
    uint8_t* pred_out;             // [E,N,K] or null
    uint8_t* belief_out;           // [E,N,K,M] or null
    uint8_t* pred_partner_out;     // [E,N] or null
    int64_t E, env_offset;
    int N, K, envs_per_block, agents_per_block, chunks_per_env, reset_prior;
    uint64_t seed;
    uint32_t episode, t;
};

template <int M>
__global__ void __launch_bounds__(kThreads) belief_pairs_kernel(PairsArgs P) {
    constexpr int A = IA2C_AGENT_ACTIONS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = P.N, K = P.K;
    const int64_t e0 = (int64_t)(blockIdx.x / P.chunks_per_env) * P.envs_per_block;
    const int chunk = blockIdx.x % P.chunks_per_env;
    const int i0 = chunk * P.agents_per_block;
    const int i1 = min(N, i0 + P.agents_per_block);
    const int n_agents = i1 - i0;
    const int n_envs = (int)min((int64_t)P.envs_per_block, P.E - e0);

    pdl_prologue();
    double* tab = reinterpret_cast<double*>(smem_raw);             // [101] k/100
    double* fa = tab + 104;                                        // [n_agents][M*A]
    int* counts = reinterpret_cast<int*>(fa + P.agents_per_block * M * A);   // [envs_per_block][agents_per_block][A]
    uint8_t* act = reinterpret_cast<uint8_t*>(counts + P.envs_per_block * P.agents_per_block * A);  // [envs_per_block][N]

    for (int k = threadIdx.x; k <= 100; k += blockDim.x) tab[k] = __ddiv_rn((double)k, 100.0);
    for (int i = threadIdx.x; i < n_agents * M * A; i += blockDim.x) fa[i] = P.filter_action[(int64_t)i0 * M * A + i];
    for (int i = threadIdx.x; i < n_envs * n_agents * A; i += blockDim.x) counts[i] = 0;
    for (int i = threadIdx.x; i < n_envs * N; i += blockDim.x) act[i] = P.actions[e0 * N + i];
    __syncthreads();

    const int prior_k = (int)rint(100.0 * (rint(100.0 / M) / 100.0));  // round(1/M, 2) in hundredths
    const int per_env = n_agents * K;
    const int total = n_envs * per_env;
    for (int q = threadIdx.x; q < total; q += blockDim.x) {
        const int el = q / per_env;
        const int rem = q - el * per_env;
        const int il = rem / K;          // local agent
        const int jj = rem - il * K;     // modelled-other slot
        const int i = i0 + il;
        const int j = jj + (jj >= i);    // others in ascending order, skipping i
        const int64_t e = e0 + el;
        const int64_t rec = ((e * N + i) * (int64_t)K + jj);
        uint2 raw = make_uint2(0u, 0u);
        if (!P.reset_prior) raw = *reinterpret_cast<const uint2*>(P.records + rec * IA2C_BELIEF_RECORD);
        double prev[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint32_t word = m < 4 ? raw.x : raw.y;
            const int k = P.reset_prior ? prior_k : (int)((word >> (8 * (m & 3))) & 0xFFu);
            prev[m] = tab[k];
        }
        const int seen = act[el * N + j];
        double lik[A];
#pragma unroll
        for (int a = 0; a < A; ++a) lik[a] = (a == seen) ? 0.8 : 0.1;   // ia2c.py:53-58
        const double u = P.u_injected ? P.u_injected[rec]
                                      : philox_belief_uniform(P.seed, P.episode, P.t, (uint64_t)((P.env_offset + e) * N + i), K, jj);
        double b[M], pred[A];
        const int ap = belief_core<M, A>(fa + il * M * A, lik, prev, u, b, pred);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint32_t k = (uint32_t)__double2int_rn(__dmul_rn(b[m], 100.0));  // rint, half-to-even
            if (m < 4) lo |= k << (8 * m); else hi |= k << (8 * (m - 4));
            if (P.belief_out) P.belief_out[rec * M + m] = (uint8_t)k;
        }
        hi |= (uint32_t)ap << 16;  // byte 6
        *reinterpret_cast<uint2*>(P.records + rec * IA2C_BELIEF_RECORD) = make_uint2(lo, hi);
        if (P.pred_out) P.pred_out[rec] = (uint8_t)ap;
        if (P.pred_partner_out) atomicAdd(&counts[(el * n_agents + il) * A + ap], 1);
    }
    if (P.pred_partner_out) {
        __syncthreads();
        for (int x = threadIdx.x; x < n_envs * n_agents; x += blockDim.x) {
            const int el = x / n_agents, il = x - el * n_agents;
            const int* c = counts + x * A;
            int best = 0;
#pragma unroll
            for (int a = 1; a < A; ++a) best = c[a] > c[best] ? a : best;   // ties -> lowest action
            P.pred_partner_out[(e0 + el) * N + i0 + il] = (uint8_t)best;
        }
    }
}

__global__ void debug_divide_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ q_seq,
                                    double* __restrict__ q_ieee, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    q_seq[i] = ddiv_seq(a[i], b[i]);
    q_ieee[i] = __ddiv_rn(a[i], b[i]);
}

// ------------------------------------------------------------------------------------------------
// Table-driven pairwise update for many modelled others (K >= 32).
//
// belief_pairs_kernel is bound by the fp64 pipe (~100 DP instructions per record against 16 B of traffic).  Two
// observations remove two thirds of them without changing a single output bit:
//  (1) the un-normalised posterior bp[m] = (l0*(F[m][0]*p) + l1*(F[m][1]*p)) + l2*(F[m][2]*p) depends only on the
//      agent's model m, the SEEN action (3 likelihood rows) and the prior, which is k/100 with integer k in 0..100.
//      A block that owns ONE agent therefore tabulates all 3*M*101 values once (same fp64 operations, same
//      order) and each record needs M table lookups instead of 6M multiply/adds;
//  (2) the mixture prediction is only COMPARED with u.  It is evaluated in fp32 (a different pipe); when u lies
//      within 1e-5 of a decision boundary — the fp32 error is below 1e-6 — the record falls back to the exact
//      fp64 sequence.  The fallback rate is ~6e-5 per record;
//  (3) the posterior is stored rounded to hundredths, so the M IEEE divisions are replaced by multiplications with the
//      shared reciprocal plus a fixed-point test that proves the rounding agrees (exact fallback otherwise).
// Per record that leaves S, one shared reciprocal, M corrected divisions and M roundings on the fp64 pipe.
// Block = (agent i, chunk of envs): the agent's K records of one env are 8*K contiguous bytes.
// FAST = the steady-state call of the rollout (device Philox, no dumps, priors from the records): the optional
// pointers are compiled out instead of costing a predicated-off instruction each per record.
template <int M, bool FAST>
__global__ void __launch_bounds__(kThreads, 4) belief_pairs_table_kernel(PairsArgs P) {
    constexpr int A = IA2C_AGENT_ACTIONS;
    const bool reset_prior = !FAST && P.reset_prior;
    const double* const u_injected = FAST ? nullptr : P.u_injected;
    uint8_t* const belief_out = FAST ? nullptr : P.belief_out;
    uint8_t* const pred_out = FAST ? nullptr : P.pred_out;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = P.N, K = P.K, i = blockIdx.y;
    const int EC = P.envs_per_block;
    const int64_t e0 = (int64_t)blockIdx.x * EC;
    const int n_envs = (int)min((int64_t)EC, P.E - e0);
    double* tab = reinterpret_cast<double*>(smem_raw);              // [104]   k/100
    double* bpt = tab + 104;                                        // [A][M][101]
    double* fa = bpt + A * M * 101;                                 // [M][A]
    float* fa32 = reinterpret_cast<float*>(fa + M * A);             // [M][A]
    int* counts = reinterpret_cast<int*>(fa32 + ((M * A + 3) & ~3)); // [EC][A]
    uint8_t* act = reinterpret_cast<uint8_t*>(counts + EC * A);     // [EC][N]
    pdl_release();
    // the table depends only on the agent's models: it is built while the kernel that samples this step's actions
    // may still be running (programmatic dependent launch); pdl_wait() below orders the reads of its output
    for (int k = threadIdx.x; k <= 100; k += blockDim.x) tab[k] = __ddiv_rn((double)k, 100.0);
    for (int k = threadIdx.x; k < M * A; k += blockDim.x) {
        fa[k] = P.filter_action[(int64_t)i * M * A + k];
        fa32[k] = (float)fa[k];
    }
    for (int k = threadIdx.x; k < n_envs * A; k += blockDim.x) counts[k] = 0;
    __syncthreads();
    for (int x = threadIdx.x; x < A * M * 101; x += blockDim.x) {
        const int seen = x / (M * 101), m = (x / 101) % M, k = x % 101;
        const double p = tab[k];
        double acc = __dmul_rn(seen == 0 ? 0.8 : 0.1, __dmul_rn(fa[m * A + 0], p));
#pragma unroll
        for (int a = 1; a < A; ++a) acc = __dadd_rn(acc, __dmul_rn(seen == a ? 0.8 : 0.1, __dmul_rn(fa[m * A + a], p)));
        bpt[x] = acc;
    }
    pdl_wait();
    for (int k = threadIdx.x; k < n_envs * N; k += blockDim.x) act[k] = P.actions[e0 * N + k];
    __syncthreads();
    const int prior_k = (int)rint(100.0 / M);
    // One thread = one PAIR of modelled-other slots (2s, 2s+1) of one env: the two records share a Philox block and
    // give the scheduler two independent chains.
    const int KP = (K + 1) >> 1;
    const float inv_kp = 1.f / (float)KP;
    const int total = n_envs * KP;
    auto one_record = [&](int64_t rec, int el, int jj, double u, uint2 raw) -> int {
        const int j = jj + (jj >= i);
        const double* row = bpt + act[el * N + j] * (M * 101);
        double bp[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint32_t word = m < 4 ? raw.x : raw.y;
            const int k = reset_prior ? prior_k : (int)((word >> (8 * (m & 3))) & 0xFFu);
            bp[m] = row[m * 101 + k];
        }
        double S = bp[0];
#pragma unroll
        for (int m = 1; m < M; ++m) S = __dadd_rn(S, bp[m]);
        const double rS = drcp_seq(S);
        // (3) screened quotients: q~ = bp * (1/S) is within 4e-16 of the IEEE quotient b = bp / S, and the posterior is
        //     only used ROUNDED to hundredths: z = rint(q~ * 100 * 2^20) is 100*b in fixed point, exact to 1e-6; unless
        //     its fraction is within 2^-19 of one half, rint(100*q~) == rint(100*b) and the M divisions are skipped.
        uint32_t kq[M];
        float bf[M];
        bool near_half = false;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const double qa = __dmul_rn(bp[m], rS);
            const int z = __double2int_rn(__dmul_rn(qa, 104857600.0));   // 100 * 2^20, exact scaling
            bf[m] = (float)qa;
            kq[m] = (uint32_t)(z + (1 << 19)) >> 20;
            const int frac = z & ((1 << 20) - 1);
            near_half |= (unsigned)(frac - (1 << 19) + 2) <= 4u;
        }
        if (near_half) {   // exact path (identical to belief_core): ~1e-5 of the records
#pragma unroll
            for (int m = 0; m < M; ++m) kq[m] = (uint32_t)__double2int_rn(__dmul_rn(ddiv_with(bp[m], S, rS), 100.0));
        }
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            if (m < 4) lo |= kq[m] << (8 * m); else hi |= kq[m] << (8 * (m - 4));
            if (belief_out) belief_out[rec * M + m] = (uint8_t)kq[m];
        }
        // fp32 screening of the inverse-CDF decision
        float pf[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m < M; ++m) acc = fmaf(bf[m], fa32[m * A + a], acc);
            pf[a] = acc;
        }
        const float uf = (float)u;
        float cf = pf[0];
        int ap = 0;
        bool found = uf < cf, risky = fabsf(uf - cf) < 1e-5f;
#pragma unroll
        for (int a = 1; a < A; ++a) {
            cf += pf[a];
            risky |= fabsf(uf - cf) < 1e-5f;
            if (!found && uf < cf) { ap = a; found = true; }
        }
        if (risky) {   // exact sequence, identical to belief_core
            double b[M];
#pragma unroll
            for (int m = 0; m < M; ++m) b[m] = ddiv_with(bp[m], S, rS);
            double c = 0.0;
            ap = 0;
            found = false;
#pragma unroll
            for (int a = 0; a < A; ++a) {
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < M; ++m) acc = __dadd_rn(acc, __dmul_rn(b[m], fa[m * A + a]));
                c = (a == 0) ? acc : __dadd_rn(c, acc);
                if (!found && u < c) { ap = a; found = true; }
            }
        }
        hi |= (uint32_t)ap << 16;  // byte 6
        *reinterpret_cast<uint2*>(P.records + rec * IA2C_BELIEF_RECORD) = make_uint2(lo, hi);
        if (pred_out) pred_out[rec] = (uint8_t)ap;
        if (FAST || P.pred_partner_out) atomicAdd(&counts[el * A + ap], 1);   // ptxas aggregates same-address lanes (REDUX)
        return ap;
    };
    // the two records of the NEXT iteration are fetched before the current pair is processed (software prefetch: the
    // kernel is bound by the latency of these loads at 4 blocks per SM).
    struct Slot { int el, sl; int64_t rec; uint2 raw0, raw1; };
    auto locate = [&](int q, Slot& s) {
        s.el = (int)(((float)q + 0.5f) * inv_kp);   // q / KP (exact: q < 2^16, margin 0.5/KP)
        s.sl = q - s.el * KP;
        s.rec = ((e0 + s.el) * N + i) * (int64_t)K + 2 * s.sl;
        s.raw0 = s.raw1 = make_uint2(0u, 0u);
        if (!reset_prior) {
            const uint2* rp = reinterpret_cast<const uint2*>(P.records + s.rec * IA2C_BELIEF_RECORD);
            s.raw0 = rp[0];
            if (2 * s.sl + 1 < K) s.raw1 = rp[1];
        }
    };
    Slot cur;
    if ((int)threadIdx.x < total) locate(threadIdx.x, cur);
    for (int q = threadIdx.x; q < total; q += blockDim.x) {
        Slot nxt = cur;
        if (q + (int)blockDim.x < total) locate(q + blockDim.x, nxt);
        const int el = cur.el, sl = cur.sl, jj = 2 * sl;
        const bool two = jj + 1 < K;
        const int64_t e = e0 + el, rec = cur.rec;
        double u0, u1 = 0.0;
        if (u_injected) {
            u0 = u_injected[rec];
            if (two) u1 = u_injected[rec + 1];
        } else {
            philox_belief_pair(P.seed, P.episode, P.t, (uint64_t)((P.env_offset + e) * N + i), K, sl, u0, u1);
        }
        one_record(rec, el, jj, u0, cur.raw0);
        if (two) one_record(rec + 1, el, jj + 1, u1, cur.raw1);
        cur = nxt;
    }
    if (P.pred_partner_out) {
        __syncthreads();
        for (int el = threadIdx.x; el < n_envs; el += blockDim.x) {
            const int* c = counts + el * A;
            int best = 0;
#pragma unroll
            for (int a = 1; a < A; ++a) best = c[a] > c[best] ? a : best;   // ties -> lowest action
            P.pred_partner_out[(e0 + el) * N + i] = (uint8_t)best;
        }
    }
}

template <int M>
int launch_pairs_table(PairsArgs& P, cudaStream_t stream) {
    P.envs_per_block = std::max(1, std::min(64, 16384 / P.K));   // amortise the table build (~15 % of the instructions at 4096)
    const int64_t env_blocks = (P.E + P.envs_per_block - 1) / P.envs_per_block;
    size_t smem = (104 + 3 * M * 101 + M * 3) * sizeof(double) + ((M * 3 + 3) & ~3) * sizeof(float) +
                  (size_t)P.envs_per_block * 3 * sizeof(int) + (size_t)P.envs_per_block * P.N;
    smem = (smem + 15) & ~size_t(15);
    dim3 grid((unsigned)env_blocks, P.N);
    const bool fast = !P.reset_prior && !P.u_injected && !P.belief_out && !P.pred_out && P.pred_partner_out;
    if (fast) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(belief_pairs_table_kernel<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        return launch_pdl("belief_pairs_table_kernel", belief_pairs_table_kernel<M, true>, grid, dim3(kThreads), smem, stream, P);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(belief_pairs_table_kernel<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        return launch_pdl("belief_pairs_table_kernel", belief_pairs_table_kernel<M, false>, grid, dim3(kThreads), smem, stream, P);
    }
}

template <int M>
int launch_pairs(PairsArgs& P, cudaStream_t stream) {
    if (P.K >= 32 && P.N <= 65535) return launch_pairs_table<M>(P, stream);
    const int64_t per_env = (int64_t)P.N * P.K;
    const int target = 4096;  // records per block
    if (per_env <= target) {
        P.envs_per_block = (int)(target / per_env);
        if (P.envs_per_block > 256) P.envs_per_block = 256;
        P.agents_per_block = P.N;
        P.chunks_per_env = 1;
    } else {
        P.envs_per_block = 1;
        const int chunks = (int)((per_env + target - 1) / target);
        P.agents_per_block = (P.N + chunks - 1) / chunks;
        P.chunks_per_env = (P.N + P.agents_per_block - 1) / P.agents_per_block;
    }
    const int64_t env_blocks = (P.E + P.envs_per_block - 1) / P.envs_per_block;
    const int64_t blocks = env_blocks * P.chunks_per_env;
    size_t smem = 104 * sizeof(double) + (size_t)P.agents_per_block * M * 3 * sizeof(double) +
                  (size_t)P.envs_per_block * P.agents_per_block * 3 * sizeof(int) + (size_t)P.envs_per_block * P.N;
    smem = (smem + 15) & ~size_t(15);
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(belief_pairs_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    return launch_pdl("belief_pairs_kernel", belief_pairs_kernel<M>, dim3((unsigned)blocks), dim3(kThreads), smem, stream, P);
}

}  // namespace
}  // namespace ia2c

using namespace ia2c;

extern "C" int ia2c_belief_update_dense(const double* filter_action, const double* lik, const double* prev,
                                        const double* u, int64_t* ap, double* bprime, double* prediction,
                                        int64_t R, int32_t M, int32_t A, void* stream) {
    IA2C_REQUIRE(R > 0 && filter_action && lik && prev && u && ap && bprime, "ia2c_belief_update_dense: R=%lld or null arrays", (long long)R);
    IA2C_REQUIRE(M >= 1 && M <= kMaxDense && A >= 1 && A <= kMaxDense, "ia2c_belief_update_dense: M=%d A=%d outside 1..8", M, A);
    const int blocks = ceil_div(R, kThreads);
    cudaStream_t s = as_stream(stream);
    if (M == 5 && A == 3) belief_dense_kernel<5, 3><<<blocks, kThreads, 0, s>>>(filter_action, lik, prev, u, ap, bprime, prediction, R);
    else if (M == 3 && A == 3) belief_dense_kernel<3, 3><<<blocks, kThreads, 0, s>>>(filter_action, lik, prev, u, ap, bprime, prediction, R);
    else if (M == 5 && A == 5) belief_dense_kernel<5, 5><<<blocks, kThreads, 0, s>>>(filter_action, lik, prev, u, ap, bprime, prediction, R);
    else belief_dense_generic_kernel<<<blocks, kThreads, 0, s>>>(filter_action, lik, prev, u, ap, bprime, prediction, R, M, A);
    return check_launch("belief_dense_kernel");
}

extern "C" int ia2c_debug_divide(const double* a, const double* b, double* q_seq, double* q_ieee, int64_t n, void* stream) {
    IA2C_REQUIRE(a && b && q_seq && q_ieee && n > 0, "ia2c_debug_divide: null pointer or n=%lld", (long long)n);
    debug_divide_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(a, b, q_seq, q_ieee, n);
    return check_launch("debug_divide_kernel");
}

extern "C" int ia2c_belief_update_pairs(uint8_t* records, const double* filter_action, const uint8_t* actions,
                                        const double* u_injected, uint8_t* pred_out, uint8_t* belief_out,
                                        uint8_t* pred_partner_out, int64_t E, int32_t N, int32_t M,
                                        int32_t reset_prior, uint64_t seed, uint32_t episode, uint32_t t,
                                        int64_t env_offset, void* stream) {
    IA2C_REQUIRE(E > 0 && records && filter_action && actions, "ia2c_belief_update_pairs: E=%lld or null arrays", (long long)E);
    IA2C_REQUIRE(N >= 2 && N <= 1023, "ia2c_belief_update_pairs: N=%d outside 2..1023", N);
    IA2C_REQUIRE(M >= 2 && M <= IA2C_MAX_MODELS, "ia2c_belief_update_pairs: M=%d outside 2..%d", M, IA2C_MAX_MODELS);
    PairsArgs P{records, filter_action, actions, u_injected, pred_out, belief_out, pred_partner_out,
                E, env_offset, N, N - 1, 0, 0, 0, reset_prior, seed, episode, t};
    cudaStream_t s = as_stream(stream);
    switch (M) {
        case 2: return launch_pairs<2>(P, s);
        case 3: return launch_pairs<3>(P, s);
        case 4: return launch_pairs<4>(P, s);
        case 5: return launch_pairs<5>(P, s);
        default: return launch_pairs<6>(P, s);
    }
}
