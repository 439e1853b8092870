// org_env.cu — batched Org environment: reset / step for E independent envs (x N agents).
//
// Replaces Org.reset (Org.py:128-148) and Org.step (Org.py:51-126) of the reference, plus the
// TimeLimit + same-step autoreset that gym.make_vec adds around it (ia2c.py:34-42).
// Arithmetic: integer state machine + one fp64 divide and add per env-step (never contracted).
// Layout: structure of arrays over envs; one thread per env for N <= 8 agents (actions are a few
// bytes per env), one warp per env with agents in lanes for larger N (ballot-free packed counts
// reduced by warp shuffles).  HBM-bound: ~60 B per env-step at N=2.
#include "common.cuh"

namespace ia2c {
namespace {

constexpr int kThreads = 256;

struct OrgArrays {
    int32_t* state;
    double* hist;
    uint8_t* cls;       // [E,2] {previous, current}
    int32_t* elapsed;
    float* obs_out;     // [E,6]
    double* reward_out;
    float* reward_f32_out;
    int32_t* state_trace;
    uint8_t* truncated_out;
    int32_t max_episode_steps;
};

// Coalesced store of 32 envs x 6 floats from one warp: staged through shared memory so that the
// warp writes 768 contiguous bytes as float4s instead of 32 strided 24-byte rows.
__device__ __forceinline__ void store_obs_warp(float* __restrict__ obs_out, int64_t warp_first_env, int64_t E,
                                               int prev_cls, int cur_cls, bool active, float* stage /*[192]*/) {
    const int lane = threadIdx.x & 31;
    if (obs_out == nullptr) return;
    float* mine = stage + lane * 6;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        mine[k] = (k == prev_cls) ? 1.f : 0.f;
        mine[3 + k] = (k == cur_cls) ? 1.f : 0.f;
    }
    __syncwarp();
    const int64_t remaining = E - warp_first_env;
    const int n_env = remaining < 32 ? (int)remaining : 32;
    float* dst = obs_out + warp_first_env * 6;
    if (n_env == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {  // 192 floats = 48 float4
        float4* d4 = reinterpret_cast<float4*>(dst);
        const float4* s4 = reinterpret_cast<const float4*>(stage);
        d4[lane] = s4[lane];
        if (lane < 16) d4[32 + lane] = s4[32 + lane];
    } else {
        for (int i = lane; i < n_env * 6; i += 32) dst[i] = stage[i];
    }
    (void)active;
    __syncwarp();
}

// Everything after the joint decision: reward recurrence, observation memory shift, TimeLimit.
__device__ __forceinline__ void org_finish(const OrgArrays& A, int64_t e, int s, int s2, double base, bool valid,
                                           int& prev_cls_out, int& cur_cls_out) {
    double r = A.hist[e];
    if (valid) r = org_reward(base, r);  // unknown action codes leave state and reward untouched (Q16)
    const int old_cur = A.cls[2 * e + 1];
    int prev_cls = old_cur;               // memory shift (Org.py:112-113)
    int cur_cls = org_obs_class(valid ? s2 : s);
    int new_state = valid ? s2 : s;
    if (A.state_trace) A.state_trace[e] = new_state;
    if (A.reward_out) A.reward_out[e] = r;
    if (A.reward_f32_out) A.reward_f32_out[e] = (float)r;
    bool trunc = false;
    if (A.max_episode_steps > 0) {
        const int el = A.elapsed[e] + 1;
        trunc = el >= A.max_episode_steps;
        A.elapsed[e] = trunc ? 0 : el;
    }
    if (trunc) {  // same-step autoreset: reset observation returned, reward above is the real one (Q14)
        new_state = 2;
        r = 0.0;
        prev_cls = 1;
        cur_cls = 1;
    }
    if (A.truncated_out) A.truncated_out[e] = trunc ? 1 : 0;
    A.state[e] = new_state;
    A.hist[e] = r;
    *reinterpret_cast<uchar2*>(A.cls + 2 * e) = make_uchar2((unsigned char)prev_cls, (unsigned char)cur_cls);
    prev_cls_out = prev_cls;
    cur_cls_out = cur_cls;
}

__global__ void __launch_bounds__(kThreads) org_reset_kernel(OrgArrays A, int64_t E) {
    __shared__ float stage[kThreads / 32][192];
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool active = e < E;
    if (active) {
        A.state[e] = 2;
        A.hist[e] = 0.0;
        *reinterpret_cast<uchar2*>(A.cls + 2 * e) = make_uchar2(1, 1);
        if (A.elapsed) A.elapsed[e] = 0;
    }
    store_obs_warp(A.obs_out, e - (threadIdx.x & 31), E, 1, 1, active, stage[threadIdx.x >> 5]);
}

// One thread per env.  JOINT: the reference's joint code; otherwise N <= 8 per-agent actions.
template <bool JOINT>
__global__ void __launch_bounds__(kThreads)
org_step_thread_kernel(OrgArrays A, const int32_t* __restrict__ joint, const uint8_t* __restrict__ actions, int N,
                       int64_t E) {
    __shared__ float stage[kThreads / 32][192];
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool active = e < E;
    int prev_cls = 1, cur_cls = 1;
    if (active) {
        int n_s = 0, n_b = 0, n_g = 0, n_agents = N;
        bool valid = true;
        if (JOINT) {
            const int j = joint[e];
            valid = (j >= 0) && (j <= 8);
            const int a1 = valid ? j / 3 : 1, a2 = valid ? j % 3 : 1;
            n_s = (a1 == 0) + (a2 == 0);
            n_b = (a1 == 1) + (a2 == 1);
            n_g = (a1 == 2) + (a2 == 2);
            n_agents = 2;
        } else {
            const uint8_t* a = actions + e * N;
            for (int i = 0; i < N; ++i) {
                const int v = a[i];
                n_s += (v == 0);
                n_b += (v == 1);
                n_g += (v == 2);
            }
        }
        const int s = A.state[e];
        int s2;
        double base;
        org_transition(s, n_s, n_b, n_g, n_agents, s2, base);
        org_finish(A, e, s, s2, base, valid, prev_cls, cur_cls);
    }
    store_obs_warp(A.obs_out, e - (threadIdx.x & 31), E, prev_cls, cur_cls, active, stage[threadIdx.x >> 5]);
}

// Counts of the action codes 0 / 1 / 2 among the four bytes of a word, packed 3 x 10 bits; any other byte counts nothing.
__device__ __forceinline__ uint32_t count_actions4(uint32_t v) {
    const uint32_t hi = (v >> 2) & 0x3F3F3F3Fu;                                    // the six high bits of every byte
    const uint32_t ok = ~((((hi + 0x7F7F7F7Fu) | hi) & 0x80808080u) >> 7) & 0x01010101u;   // 1 where the byte is < 4
    const uint32_t b0 = v, b1 = v >> 1;
    return (uint32_t)__popc(~b0 & ~b1 & ok) | ((uint32_t)__popc(b0 & ~b1 & ok) << 10) | ((uint32_t)__popc(~b0 & b1 & ok) << 20);
}

// Org-N step for N > 8: a warp owns 32 envs.  Phase 1 counts the actions of one env (or of several, for N <= 64) per load
// instruction with the agents' bytes in the lanes — 32-bit words, 4 actions each, counts packed 3 x 10 bits, one warp
// reduction per env — and hands env j's total to lane j.  Phase 2 is org_step_thread_kernel's body with all 32 lanes busy:
// coalesced state / reward-history loads and stores, observation rows staged through shared memory into 128-bit stores.
// (The first version spent a whole warp per env and left the transition to lane 0: 0.8 TB/s at N = 256.)
// Requires N % 4 == 0 (rows are then word-aligned); other N take org_step_warp_kernel below.
template <int WPL>   // words per lane and env: >= ceil(N / 128), a power of two; 1 also covers the several-envs-per-load case
__global__ void __launch_bounds__(kThreads)
org_step_warp32_kernel(OrgArrays A, const uint8_t* __restrict__ actions, int N, int64_t E) {
    __shared__ float stage[kThreads / 32][192];
    const int lane = threadIdx.x & 31;
    const int64_t e0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x - lane);   // the warp's first env
    if (e0 >= E) return;
    const int NW = N >> 2;                                   // words per env
    const uint32_t* a4 = reinterpret_cast<const uint32_t*>(actions);
    uint32_t mine = 0;                                       // lane j: packed counts of env e0 + j
    if (WPL == 1 && NW <= 16) {
        // several envs per load instruction: a group of GW = pow2 >= NW lanes per env
        int GW = 1;
        while (GW < NW) GW <<= 1;
        const int EPI = 32 / GW, g = lane / GW, w = lane % GW;
        for (int it = 0; it < GW; ++it) {                    // 32 / EPI iterations
            const int64_t e = e0 + it * EPI + g;
            uint32_t packed = (w < NW && e < E) ? count_actions4(a4[e * NW + w]) : 0u;
            for (int off = GW >> 1; off > 0; off >>= 1) packed += __shfl_xor_sync(0xffffffffu, packed, off);
            const uint32_t got = __shfl_sync(0xffffffffu, packed, ((lane - it * EPI) & (EPI - 1)) * GW);
            if (lane >= it * EPI && lane < (it + 1) * EPI) mine = got;
        }
    } else {
        // one env per warp reduction; WPL words per lane and env, EPC envs (8 loads per lane) in flight
        constexpr int EPC = WPL >= 8 ? 1 : 8 / WPL;
        for (int j0 = 0; j0 < 32; j0 += EPC) {
            uint32_t v[EPC][WPL];
#pragma unroll
            for (int u = 0; u < EPC; ++u) {
                const int64_t e = e0 + j0 + u;
#pragma unroll
                for (int k = 0; k < WPL; ++k) {
                    const int w = lane + 32 * k;
                    v[u][k] = (w < NW && e < E) ? a4[e * NW + w] : 0xFFFFFFFFu;   // 0xFF bytes count nothing
                }
            }
#pragma unroll
            for (int u = 0; u < EPC; ++u) {
                uint32_t packed = 0u;
#pragma unroll
                for (int k = 0; k < WPL; ++k) packed += count_actions4(v[u][k]);
                const uint32_t total = __reduce_add_sync(0xffffffffu, packed);
                if (lane == j0 + u) mine = total;
            }
        }
    }
    const int64_t e = e0 + lane;
    const bool active = e < E;
    int prev_cls = 1, cur_cls = 1;
    if (active) {
        const int s = A.state[e];
        int s2;
        double base;
        org_transition(s, mine & 1023, (mine >> 10) & 1023, (mine >> 20) & 1023, N, s2, base);
        org_finish(A, e, s, s2, base, true, prev_cls, cur_cls);
    }
    store_obs_warp(A.obs_out, e0, E, prev_cls, cur_cls, active, stage[threadIdx.x >> 5]);
}

// One warp per env, agents in lanes: the general form (any N > 8, unaligned rows).  Counts are packed 3 x 10 bits and reduced by shuffles.
__global__ void __launch_bounds__(kThreads)
org_step_warp_kernel(OrgArrays A, const uint8_t* __restrict__ actions, int N, int64_t E) {
    const int lane = threadIdx.x & 31;
    const int64_t e = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    if (e >= E) return;
    const uint8_t* a = actions + e * N;
    uint32_t packed = 0;  // n_s | n_b << 10 | n_g << 20   (N <= 1023)
    if ((N & 3) == 0 && ((e * N) & 3) == 0) {
        const uint32_t* a4 = reinterpret_cast<const uint32_t*>(a);
        for (int w = lane; w < (N >> 2); w += 32) {
            const uint32_t v = a4[w];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t x = (v >> (8 * b)) & 0xFFu;
                packed += (x == 0) ? 1u : (x == 1 ? (1u << 10) : (x == 2 ? (1u << 20) : 0u));
            }
        }
    } else {
        for (int i = lane; i < N; i += 32) {
            const uint32_t x = a[i];
            packed += (x == 0) ? 1u : (x == 1 ? (1u << 10) : (x == 2 ? (1u << 20) : 0u));
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) packed += __shfl_xor_sync(0xffffffffu, packed, off);
    int prev_cls = 1, cur_cls = 1;
    if (lane == 0) {
        const int s = A.state[e];
        int s2;
        double base;
        org_transition(s, packed & 1023, (packed >> 10) & 1023, (packed >> 20) & 1023, N, s2, base);
        org_finish(A, e, s, s2, base, true, prev_cls, cur_cls);
    }
    prev_cls = __shfl_sync(0xffffffffu, prev_cls, 0);
    cur_cls = __shfl_sync(0xffffffffu, cur_cls, 0);
    if (A.obs_out && lane < 6) A.obs_out[e * 6 + lane] = (lane < 3 ? (lane == prev_cls) : (lane - 3 == cur_cls)) ? 1.f : 0.f;
}

}  // namespace
}  // namespace ia2c

using namespace ia2c;

extern "C" int ia2c_org_reset(int32_t* state, double* hist, uint8_t* cls, int32_t* elapsed, float* obs_out,
                              int64_t E, void* stream) {
    IA2C_REQUIRE(E > 0 && state && hist && cls, "ia2c_org_reset: E=%lld or null state arrays", (long long)E);
    OrgArrays A{state, hist, cls, elapsed, obs_out, nullptr, nullptr, nullptr, nullptr, 0};
    org_reset_kernel<<<ceil_div(E, kThreads), kThreads, 0, as_stream(stream)>>>(A, E);
    return check_launch("org_reset_kernel");
}

extern "C" int ia2c_org_step_joint(int32_t* state, double* hist, uint8_t* cls, int32_t* elapsed,
                                   const int32_t* joint, float* obs_out, double* reward_out, float* reward_f32_out,
                                   int32_t* state_trace, uint8_t* truncated_out, int64_t E,
                                   int32_t max_episode_steps, void* stream) {
    IA2C_REQUIRE(E > 0 && state && hist && cls && joint, "ia2c_org_step_joint: E=%lld or null arrays", (long long)E);
    IA2C_REQUIRE(max_episode_steps <= 0 || elapsed, "ia2c_org_step_joint: TimeLimit needs the elapsed array");
    OrgArrays A{state, hist, cls, elapsed, obs_out, reward_out, reward_f32_out, state_trace, truncated_out,
                max_episode_steps};
    org_step_thread_kernel<true><<<ceil_div(E, kThreads), kThreads, 0, as_stream(stream)>>>(A, joint, nullptr, 2, E);
    return check_launch("org_step_thread_kernel<joint>");
}

extern "C" int ia2c_org_step_agents(int32_t* state, double* hist, uint8_t* cls, int32_t* elapsed,
                                    const uint8_t* actions, float* obs_out, double* reward_out,
                                    float* reward_f32_out, int32_t* state_trace, uint8_t* truncated_out, int64_t E,
                                    int32_t N, int32_t max_episode_steps, void* stream) {
    IA2C_REQUIRE(E > 0 && state && hist && cls && actions, "ia2c_org_step_agents: E=%lld or null arrays", (long long)E);
    IA2C_REQUIRE(N >= 1 && N <= 1023, "ia2c_org_step_agents: N=%d outside 1..1023", N);
    IA2C_REQUIRE(max_episode_steps <= 0 || elapsed, "ia2c_org_step_agents: TimeLimit needs the elapsed array");
    OrgArrays A{state, hist, cls, elapsed, obs_out, reward_out, reward_f32_out, state_trace, truncated_out,
                max_episode_steps};
    if (N <= 8) {
        org_step_thread_kernel<false><<<ceil_div(E, kThreads), kThreads, 0, as_stream(stream)>>>(A, nullptr, actions, N, E);
        return check_launch("org_step_thread_kernel<agents>");
    }
    if ((N & 3) == 0 && (reinterpret_cast<uintptr_t>(actions) & 3) == 0) {
        const dim3 grid((unsigned)ceil_div(E, kThreads));
        const cudaStream_t s = as_stream(stream);
        if (N <= 128) org_step_warp32_kernel<1><<<grid, kThreads, 0, s>>>(A, actions, N, E);
        else if (N <= 256) org_step_warp32_kernel<2><<<grid, kThreads, 0, s>>>(A, actions, N, E);
        else if (N <= 512) org_step_warp32_kernel<4><<<grid, kThreads, 0, s>>>(A, actions, N, E);
        else org_step_warp32_kernel<8><<<grid, kThreads, 0, s>>>(A, actions, N, E);
        return check_launch("org_step_warp32_kernel");
    }
    org_step_warp_kernel<<<ceil_div(E * 32, kThreads), kThreads, 0, as_stream(stream)>>>(A, actions, N, E);
    return check_launch("org_step_warp_kernel");
}
