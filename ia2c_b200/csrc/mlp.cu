// mlp.cu — the reference's 3-layer actor/critic MLP (in -> 6 -> 6 -> out) for arbitrary input width.
//
// Replaces NeuralNet.forward (ac_nets.py:34-41), autograd through it (ac_nets.py:71,118), and
// ActorNetwork.sample_action (ac_nets.py:94-102) behind the class API.  (The fused trainer has its own
// fully-templated row kernels in update.cu / rollout.cu; these are the general-shape versions.)
//
// hidden_size = 6 makes every GEMM a K=6 or N=6 sliver: CUDA-core FFMA with the weights staged in
// shared memory, never tensor cores.  Two row mappings:
//   F <  64 : one thread per row (the Org shape, F=6)
//   F >= 64 : one warp per row, features strided over lanes so the [rows,F] read is coalesced, partial
//             layer-1 sums reduced by shuffles (the a2c_test.py shape, F=500: an HBM stream of X)
// Backward: B1 recomputes the forward per row, forms dz2/dz1, reduces the small gradients per block
// and stores dz1[rows,6]; B2 forms dW1 = dz1^T X with one thread per feature column (coalesced over
// F) and row chunks over grid.y; B3 adds the partials in a fixed order (deterministic).
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace ia2c {
namespace {

constexpr int kThreads = 256;
constexpr int kWideF = 64;
constexpr int kMaxBlocks = kSMs * 4;

template <int OMAX>
struct Acts {
    float h1[H], h2[H], y[OMAX];
};

// layers 2 and 3 (+ optional softmax) from h1; w in shared memory.
template <int OMAX>
__device__ __forceinline__ void tail_forward(const float* w, const MlpLayout& L, int O, bool softmax, Acts<OMAX>& a) {
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float acc = w[L.b2 + j];
#pragma unroll
        for (int k = 0; k < H; ++k) acc = fmaf(w[L.w2 + j * H + k], a.h1[k], acc);
        a.h2[j] = fmaxf(acc, 0.f);
    }
#pragma unroll
    for (int o = 0; o < OMAX; ++o) {
        if (o < O) {
            float acc = w[L.b3 + o];
#pragma unroll
            for (int k = 0; k < H; ++k) acc = fmaf(w[L.w3 + o * H + k], a.h2[k], acc);
            a.y[o] = acc;
        }
    }
    if (softmax) {
        float m = a.y[0];
#pragma unroll
        for (int o = 1; o < OMAX; ++o) if (o < O) m = fmaxf(m, a.y[o]);
        float s = 0.f;
#pragma unroll
        for (int o = 0; o < OMAX; ++o) if (o < O) { a.y[o] = expf(a.y[o] - m); s += a.y[o]; }
        const float inv = 1.f / s;
#pragma unroll
        for (int o = 0; o < OMAX; ++o) if (o < O) a.y[o] *= inv;
    }
}

// layer 1 for one row, thread-per-row (all lanes distinct rows).
__device__ __forceinline__ void layer1_thread(const float* w, const MlpLayout& L, const float* __restrict__ xr, float (&h1)[H]) {
    float z[H];
#pragma unroll
    for (int j = 0; j < H; ++j) z[j] = w[L.b1 + j];
    for (int f = 0; f < L.F; ++f) {
        const float xv = __ldg(xr + f);
#pragma unroll
        for (int j = 0; j < H; ++j) z[j] = fmaf(w[L.w1 + j * L.F + f], xv, z[j]);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) h1[j] = fmaxf(z[j], 0.f);
}

// layer 1 for one row, warp-per-row (features over lanes).  Every lane ends with the full h1.
__device__ __forceinline__ void layer1_warp(const float* w, const MlpLayout& L, const float* __restrict__ xr, float (&h1)[H]) {
    const int lane = threadIdx.x & 31;
    float z[H];
#pragma unroll
    for (int j = 0; j < H; ++j) z[j] = 0.f;
    // 32-bit loads, 128 contiguous bytes per warp instruction, eight in flight per lane (a 128-bit variant with
    // 128-bit shared-memory weight reads measured slower: 136 us vs 104 us at F=500, rows=65536)
#pragma unroll 8
    for (int f = lane; f < L.F; f += 32) {
        const float xv = __ldg(xr + f);
#pragma unroll
        for (int j = 0; j < H; ++j) z[j] = fmaf(w[L.w1 + j * L.F + f], xv, z[j]);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) h1[j] = fmaxf(warp_sum(z[j]) + w[L.b1 + j], 0.f);
}

template <int OMAX, bool WIDE>
__global__ void __launch_bounds__(kThreads)
mlp_forward_kernel(const float* __restrict__ params, const float* __restrict__ x, float* __restrict__ y,
                   float* __restrict__ h1_out, int64_t rows, int F, int O, int softmax) {
    extern __shared__ float w[];
    const MlpLayout L(F, O);
    const float* p = params + (int64_t)blockIdx.y * L.P;
    for (int i = threadIdx.x; i < L.P; i += blockDim.x) w[i] = p[i];
    __syncthreads();
    float* yn = y + (int64_t)blockIdx.y * rows * O;
    Acts<OMAX> a;
    if (!WIDE) {
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
            layer1_thread(w, L, x + r * F, a.h1);
            if (h1_out) {
#pragma unroll
                for (int j = 0; j < H; ++j) h1_out[r * H + j] = a.h1[j];
            }
            tail_forward<OMAX>(w, L, O, softmax != 0, a);
#pragma unroll
            for (int o = 0; o < OMAX; ++o) if (o < O) yn[r * O + o] = a.y[o];
        }
    } else {
        const int lane = threadIdx.x & 31;
        const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
        for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
            layer1_warp(w, L, x + r * F, a.h1);
            if (h1_out) {   // one store instruction: lane j < 6 writes h1[j]
                float v = a.h1[0];
#pragma unroll
                for (int j = 1; j < H; ++j) v = (lane == j) ? a.h1[j] : v;
                if (lane < H) h1_out[r * H + lane] = v;
            }
            tail_forward<OMAX>(w, L, O, softmax != 0, a);
#pragma unroll
            for (int o = 0; o < OMAX; ++o) if (o < O && lane == (o & 31)) yn[r * O + o] = a.y[o];
        }
    }
}

template <int OMAX, bool WIDE>
__global__ void __launch_bounds__(kThreads)
actor_sample_kernel(const float* __restrict__ params, const float* __restrict__ x, const float* __restrict__ u,
                    int64_t* __restrict__ actions, float* __restrict__ probs_out, int64_t rows, int F, int O,
                    uint64_t seed, uint64_t counter) {
    extern __shared__ float w[];
    const MlpLayout L(F, O);
    for (int i = threadIdx.x; i < L.P; i += blockDim.x) w[i] = params[i];
    __syncthreads();
    Acts<OMAX> a;
    const int lane = threadIdx.x & 31;
    const int64_t stride = WIDE ? (((int64_t)gridDim.x * blockDim.x) >> 5) : (int64_t)gridDim.x * blockDim.x;
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (WIDE) r >>= 5;
    for (; r < rows; r += stride) {
        if (WIDE) layer1_warp(w, L, x + r * F, a.h1); else layer1_thread(w, L, x + r * F, a.h1);
        tail_forward<OMAX>(w, L, O, true, a);
        const float uu = u ? u[r] : philox_uniform_f32(seed, kStreamAction, (uint32_t)(counter >> 16), (uint32_t)(counter & 0xFFFF), (uint64_t)r);
        float s = 0.f;
#pragma unroll
        for (int o = 0; o < OMAX; ++o) if (o < O) s += a.y[o];
        const float us = uu * s;
        float c = 0.f;
        int act = O - 1;
        bool found = false;
#pragma unroll
        for (int o = 0; o < OMAX; ++o) {
            if (o < O) {
                c += a.y[o];
                if (!found && us < c) { act = o; found = true; }
            }
        }
        if (!WIDE || lane == 0) {
            actions[r] = act;
            if (probs_out) {
#pragma unroll
                for (int o = 0; o < OMAX; ++o) if (o < O) probs_out[r * O + o] = a.y[o];
            }
        }
    }
}

// ---- backward B1: per-row forward recompute, dz2/dz1, small-gradient block partials, dz1 store, dx.
// small gradient vector layout in the partials: [b1(6) | W2(36) | b2(6) | W3(O*6) | b3(O)]

template <int OMAX, bool WIDE>
__global__ void __launch_bounds__(kThreads)
mlp_backward_rows_kernel(const float* __restrict__ params, const float* __restrict__ x, const float* __restrict__ dy_in,
                         const float* __restrict__ h1_saved, float* __restrict__ dz1_out, float* __restrict__ dx,
                         float* __restrict__ partials, int64_t rows, int F, int O, int softmax) {
    extern __shared__ float w[];
    const MlpLayout L(F, O);
    for (int i = threadIdx.x; i < L.P; i += blockDim.x) w[i] = params[i];
    __syncthreads();
    const int n_small = H + H * H + H + O * H + O;
    float gb1[H], gw2[H * H], gb2[H], gw3[OMAX * H], gb3[OMAX];  // compile-time indexed -> registers
#pragma unroll
    for (int i = 0; i < H; ++i) gb1[i] = gb2[i] = 0.f;
#pragma unroll
    for (int i = 0; i < H * H; ++i) gw2[i] = 0.f;
#pragma unroll
    for (int i = 0; i < OMAX * H; ++i) gw3[i] = 0.f;
#pragma unroll
    for (int i = 0; i < OMAX; ++i) gb3[i] = 0.f;
    Acts<OMAX> a;
    const int lane = threadIdx.x & 31;
    const int64_t stride = WIDE ? (((int64_t)gridDim.x * blockDim.x) >> 5) : (int64_t)gridDim.x * blockDim.x;
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (WIDE) r >>= 5;
    for (; r < rows; r += stride) {
        if (h1_saved) {   // layer-1 activations kept by the forward pass: X is not read again here
#pragma unroll
            for (int j = 0; j < H; ++j) a.h1[j] = __ldg(h1_saved + r * H + j);
        } else if (WIDE) {
            layer1_warp(w, L, x + r * F, a.h1);
        } else {
            layer1_thread(w, L, x + r * F, a.h1);
        }
        tail_forward<OMAX>(w, L, O, softmax != 0, a);
        float dy[OMAX];
#pragma unroll
        for (int o = 0; o < OMAX; ++o) dy[o] = (o < O) ? dy_in[r * O + o] : 0.f;
        if (softmax) {  // dL/dy_k = p_k (G_k - sum_j p_j G_j)
            float dot = 0.f;
#pragma unroll
            for (int o = 0; o < OMAX; ++o) if (o < O) dot = fmaf(a.y[o], dy[o], dot);
#pragma unroll
            for (int o = 0; o < OMAX; ++o) if (o < O) dy[o] = a.y[o] * (dy[o] - dot);
        }
        const bool owner = !WIDE || lane == 0;  // in warp-per-row mode every lane holds the same values
        float dh2[H], dh1[H], dz1[H];
#pragma unroll
        for (int k = 0; k < H; ++k) dh2[k] = 0.f;
#pragma unroll
        for (int o = 0; o < OMAX; ++o) {
            if (o < O) {
                if (owner) gb3[o] += dy[o];
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    if (owner) gw3[o * H + k] = fmaf(dy[o], a.h2[k], gw3[o * H + k]);
                    dh2[k] = fmaf(dy[o], w[L.w3 + o * H + k], dh2[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < H; ++k) dh1[k] = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const float dz = a.h2[j] > 0.f ? dh2[j] : 0.f;
            if (owner) gb2[j] += dz;
#pragma unroll
            for (int k = 0; k < H; ++k) {
                if (owner) gw2[j * H + k] = fmaf(dz, a.h1[k], gw2[j * H + k]);
                dh1[k] = fmaf(dz, w[L.w2 + j * H + k], dh1[k]);
            }
        }
#pragma unroll
        for (int j = 0; j < H; ++j) {
            dz1[j] = a.h1[j] > 0.f ? dh1[j] : 0.f;
            if (owner) gb1[j] += dz1[j];
        }
        if (owner) {
#pragma unroll
            for (int j = 0; j < H; ++j) dz1_out[r * H + j] = dz1[j];
        }
        if (dx) {
            for (int f = WIDE ? lane : 0; f < F; f += WIDE ? 32 : 1) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < H; ++j) acc = fmaf(dz1[j], w[L.w1 + j * F + f], acc);
                dx[r * F + f] = acc;
            }
        }
    }
    // block reduce the small gradients (fixed order)
    __syncthreads();
    float* red = w;  // reuse: needs nwarps * n_small floats (host guarantees the allocation)
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float* mine = red + warp * n_small;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const float s1 = warp_sum(gb1[i]), s2 = warp_sum(gb2[i]);
        if (lane == 0) { mine[i] = s1; mine[H + H * H + i] = s2; }
    }
#pragma unroll
    for (int i = 0; i < H * H; ++i) {
        const float s2 = warp_sum(gw2[i]);
        if (lane == 0) mine[H + i] = s2;
    }
#pragma unroll
    for (int i = 0; i < OMAX * H; ++i) {
        const float s3 = warp_sum(gw3[i]);
        if (lane == 0 && i < O * H) mine[2 * H + H * H + i] = s3;
    }
#pragma unroll
    for (int i = 0; i < OMAX; ++i) {
        const float s3 = warp_sum(gb3[i]);
        if (lane == 0 && i < O) mine[2 * H + H * H + O * H + i] = s3;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_small; i += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < nwarps; ++k) s += red[k * n_small + i];
        partials[(int64_t)blockIdx.x * n_small + i] = s;
    }
}

// ---- backward B2: dW1[j][f] = sum_r dz1[r][j] * x[r][f]; thread per feature, rows chunked over grid.y.
__global__ void __launch_bounds__(kThreads)
mlp_backward_w1_kernel(const float* __restrict__ x, const float* __restrict__ dz1, float* __restrict__ partials_w1,
                       int64_t rows, int F, int rows_per_chunk) {
    __shared__ float dz[64][H];
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t r1 = (r0 + rows_per_chunk < rows) ? r0 + rows_per_chunk : rows;
    float acc[H];
#pragma unroll
    for (int j = 0; j < H; ++j) acc[j] = 0.f;
    for (int64_t base = r0; base < r1; base += 64) {
        const int n = (r1 - base) < 64 ? (int)(r1 - base) : 64;
        __syncthreads();
        for (int i = threadIdx.x; i < n * H; i += blockDim.x) dz[i / H][i % H] = dz1[base * H + i];
        __syncthreads();
        if (f < F) {
#pragma unroll 8
            for (int rr = 0; rr < n; ++rr) {
                const float xv = __ldg(x + (base + rr) * F + f);
#pragma unroll
                for (int j = 0; j < H; ++j) acc[j] = fmaf(dz[rr][j], xv, acc[j]);
            }
        }
    }
    if (f < F) {
#pragma unroll
        for (int j = 0; j < H; ++j) partials_w1[((int64_t)blockIdx.y * H + j) * F + f] = acc[j];
    }
}

// ---- index input: the row's observation is one_hot(idx, F), as a2c_test.py:57,67 builds it.  Layer 1
// is then b1 + W1[:, idx] — the dense kernels compute exactly that (every other term is fma(w, 0, acc) = acc), so these
// kernels return the dense kernels' bits without the F-wide row: no [rows, F] tensor is ever materialised or read.
// Only the F-independent tail of the parameters is staged in shared memory; the W1 column is gathered through L1.
__device__ __forceinline__ void layer1_index(const float* __restrict__ params, const float* ws, int F, int64_t idx, float (&h1)[H]) {
    idx = idx < 0 ? 0 : (idx >= F ? F - 1 : idx);   // never read outside W1; callers validate the range (Python) or flag it (ia2c_net_update)
#pragma unroll
    for (int j = 0; j < H; ++j) h1[j] = fmaxf(__fadd_rn(__ldg(params + (int64_t)j * F + idx), ws[j]), 0.f);
}

template <int OMAX, bool SAMPLE>
__global__ void __launch_bounds__(kThreads)
mlp_forward_index_kernel(const float* __restrict__ params, const int64_t* __restrict__ idx, float* __restrict__ y,
                         float* __restrict__ h1_out, const float* __restrict__ u, int64_t* __restrict__ actions,
                         int64_t rows, int F, int O, int softmax, uint64_t seed, uint64_t counter) {
    extern __shared__ float ws[];                 // parameters after W1: b1 | W2 | b2 | W3 | b3
    const MlpLayout L0(0, O);                     // the same layout with the W1 block removed
    for (int i = threadIdx.x; i < L0.P; i += blockDim.x) ws[i] = params[(int64_t)H * F + i];
    __syncthreads();
    Acts<OMAX> a;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        layer1_index(params, ws, F, idx[r], a.h1);
        if (h1_out) {
#pragma unroll
            for (int j = 0; j < H; ++j) h1_out[r * H + j] = a.h1[j];
        }
        tail_forward<OMAX>(ws, L0, O, SAMPLE || softmax != 0, a);
        if (y) {
#pragma unroll
            for (int o = 0; o < OMAX; ++o) if (o < O) y[r * O + o] = a.y[o];
        }
        if (SAMPLE) {   // same sampler as actor_sample_kernel
            const float uu = u ? u[r] : philox_uniform_f32(seed, kStreamAction, (uint32_t)(counter >> 16), (uint32_t)(counter & 0xFFFF), (uint64_t)r);
            float s = 0.f;
#pragma unroll
            for (int o = 0; o < OMAX; ++o) if (o < O) s += a.y[o];
            const float us = uu * s;
            float c = 0.f;
            int act = O - 1;
            bool found = false;
#pragma unroll
            for (int o = 0; o < OMAX; ++o) {
                if (o < O) {
                    c += a.y[o];
                    if (!found && us < c) { act = o; found = true; }
                }
            }
            actions[r] = act;
        }
    }
}

// dW1[j][f] = sum over the rows with idx == f of dz1[r][j], rows in ascending order: the dense B2 kernel's sums exactly.
__global__ void __launch_bounds__(kThreads)
mlp_backward_w1_index_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dz1, float* __restrict__ partials_w1,
                             int64_t rows, int F, int rows_per_chunk) {
    __shared__ float dz[64][H];
    __shared__ int id_s[64];
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t r1 = (r0 + rows_per_chunk < rows) ? r0 + rows_per_chunk : rows;
    float acc[H];
#pragma unroll
    for (int j = 0; j < H; ++j) acc[j] = 0.f;
    for (int64_t base = r0; base < r1; base += 64) {
        const int n = (r1 - base) < 64 ? (int)(r1 - base) : 64;
        __syncthreads();
        for (int i = threadIdx.x; i < n * H; i += blockDim.x) dz[i / H][i % H] = dz1[base * H + i];
        for (int i = threadIdx.x; i < n; i += blockDim.x) id_s[i] = (int)idx[base + i];
        __syncthreads();
        if (f < F) {
#pragma unroll 8
            for (int rr = 0; rr < n; ++rr) {
                if (id_s[rr] == f) {
#pragma unroll
                    for (int j = 0; j < H; ++j) acc[j] = __fadd_rn(dz[rr][j], acc[j]);
                }
            }
        }
    }
    if (f < F) {
#pragma unroll
        for (int j = 0; j < H; ++j) partials_w1[((int64_t)blockIdx.y * H + j) * F + f] = acc[j];
    }
}

// ---- backward B3: fixed-order sum of partials into the flat gradient.  Block = 32 gradient entries x 32 slices: slice k
// adds the partial rows k, k+32, ... in ascending order (loads independent, issued back to back), then the 32 slices are
// added in ascending order.  (One thread per entry walking all <= 1024 partial rows was a chain of dependent L2 round
// trips: ~60 us of the 99 us backward at 65536 x 500.)
constexpr int kB3Slices = 32;
__global__ void __launch_bounds__(32 * kB3Slices)
mlp_backward_reduce_kernel(const float* __restrict__ partials_small, int n_blocks, const float* __restrict__ partials_w1,
                           int n_chunks, float* __restrict__ grad, int F, int O, int accumulate) {
    __shared__ float part[kB3Slices][33];
    const MlpLayout L(F, O);
    const int n_small = L.P - H * F;
    const int col = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + col;
    float s = 0.f;
    if (i < L.P) {
        const bool w1 = i < H * F;
        const float* src = w1 ? partials_w1 + i : partials_small + (i - H * F);   // flat order after W1 == small layout
        const int64_t stride = w1 ? (int64_t)H * F : n_small;
        const int n = w1 ? n_chunks : n_blocks;
        for (int c0 = slice; c0 < n; c0 += 8 * kB3Slices) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = c0 + u * kB3Slices;
                v[u] = c < n ? src[(int64_t)c * stride] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
    }
    part[slice][col] = s;
    __syncthreads();
    if (slice != 0 || i >= L.P) return;
#pragma unroll
    for (int k = 1; k < kB3Slices; ++k) s += part[k][col];
    grad[i] = accumulate ? grad[i] + s : s;
}

struct BackwardPlan {
    int blocks_rows, chunks, rows_per_chunk, n_small;
    size_t off_dz1, off_small, off_w1, total;
};
BackwardPlan plan_backward(int64_t rows, int F, int O) {
    BackwardPlan p;
    const bool wide = F >= kWideF;
    const int64_t threads = wide ? rows * 32 : rows;
    p.blocks_rows = (int)std::min<int64_t>(kMaxBlocks, (threads + kThreads - 1) / kThreads);
    if (p.blocks_rows < 1) p.blocks_rows = 1;
    p.chunks = (int)std::min<int64_t>(1024, (rows + 63) / 64);   // >= 64 rows per chunk, enough blocks to fill the chip
    if (p.chunks < 1) p.chunks = 1;
    p.rows_per_chunk = (int)((rows + p.chunks - 1) / p.chunks);
    p.n_small = H + H * H + H + O * H + O;
    p.off_dz1 = 0;
    p.off_small = (size_t)rows * H;
    p.off_w1 = p.off_small + (size_t)p.blocks_rows * p.n_small;
    p.total = p.off_w1 + (size_t)p.chunks * H * F;
    return p;
}

template <typename Fn>
int dispatch_omax(int O, Fn&& fn) {
    if (O <= 4) return fn(std::integral_constant<int, 4>());
    if (O <= 9) return fn(std::integral_constant<int, 9>());
    return fn(std::integral_constant<int, 32>());
}

}  // namespace
}  // namespace ia2c

using namespace ia2c;

extern "C" int ia2c_mlp_forward(const float* params, const float* x, float* y, float* h1_out, int64_t rows, int32_t F,
                                int32_t O, int32_t nets, int32_t softmax, void* stream) {
    IA2C_REQUIRE(params && x && y && rows > 0, "ia2c_mlp_forward: null pointer or rows=%lld", (long long)rows);
    IA2C_REQUIRE(F >= 1 && F <= 8192 && O >= 1 && O <= 32 && nets >= 1, "ia2c_mlp_forward: F=%d O=%d nets=%d unsupported", F, O, nets);
    IA2C_REQUIRE(h1_out == nullptr || nets == 1, "ia2c_mlp_forward: h1_out is only defined for nets == 1");
    const MlpLayout L(F, O);
    const bool wide = F >= kWideF;
    const int64_t threads = wide ? rows * 32 : rows;
    dim3 grid((unsigned)std::min<int64_t>(kMaxBlocks, (threads + kThreads - 1) / kThreads), nets);
    const size_t smem = L.P * sizeof(float);
    cudaStream_t s = as_stream(stream);
    return dispatch_omax(O, [&](auto om) {
        constexpr int OM = decltype(om)::value;
        if (wide) {
            if (smem > 48 * 1024) cudaFuncSetAttribute(mlp_forward_kernel<OM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            mlp_forward_kernel<OM, true><<<grid, kThreads, smem, s>>>(params, x, y, h1_out, rows, F, O, softmax);
        } else {
            mlp_forward_kernel<OM, false><<<grid, kThreads, smem, s>>>(params, x, y, h1_out, rows, F, O, softmax);
        }
        return check_launch("mlp_forward_kernel");
    });
}

extern "C" int ia2c_actor_sample(const float* params, const float* x, const float* u, int64_t* actions_out,
                                 float* probs_out, int64_t rows, int32_t F, int32_t O, uint64_t seed,
                                 uint64_t counter, void* stream) {
    IA2C_REQUIRE(params && x && actions_out && rows > 0, "ia2c_actor_sample: null pointer or rows=%lld", (long long)rows);
    IA2C_REQUIRE(F >= 1 && F <= 8192 && O >= 1 && O <= 32, "ia2c_actor_sample: F=%d O=%d unsupported", F, O);
    const MlpLayout L(F, O);
    const bool wide = F >= kWideF;
    const int64_t threads = wide ? rows * 32 : rows;
    const int grid = (int)std::min<int64_t>(kMaxBlocks, (threads + kThreads - 1) / kThreads);
    const size_t smem = L.P * sizeof(float);
    cudaStream_t s = as_stream(stream);
    return dispatch_omax(O, [&](auto om) {
        constexpr int OM = decltype(om)::value;
        if (wide) {
            if (smem > 48 * 1024) cudaFuncSetAttribute(actor_sample_kernel<OM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            actor_sample_kernel<OM, true><<<grid, kThreads, smem, s>>>(params, x, u, actions_out, probs_out, rows, F, O, seed, counter);
        } else {
            actor_sample_kernel<OM, false><<<grid, kThreads, smem, s>>>(params, x, u, actions_out, probs_out, rows, F, O, seed, counter);
        }
        return check_launch("actor_sample_kernel");
    });
}

extern "C" size_t ia2c_mlp_backward_workspace(int64_t rows, int32_t F, int32_t O) {
    if (rows <= 0 || F < 1 || O < 1) return 0;
    return plan_backward(rows, F, O).total;
}

extern "C" int ia2c_mlp_backward(const float* params, const float* x, const float* dy, const float* h1_saved, float* grad,
                                 float* dx, float* workspace, int64_t rows, int32_t F, int32_t O, int32_t softmax,
                                 int32_t accumulate, void* stream) {
    IA2C_REQUIRE(params && x && dy && grad && workspace && rows > 0, "ia2c_mlp_backward: null pointer or rows=%lld", (long long)rows);
    IA2C_REQUIRE(F >= 1 && F <= 8192 && O >= 1 && O <= 32, "ia2c_mlp_backward: F=%d O=%d unsupported", F, O);
    const MlpLayout L(F, O);
    const BackwardPlan p = plan_backward(rows, F, O);
    // with the layer-1 activations saved there is no row of X to stream here: one thread per row is enough
    const bool wide = F >= kWideF && h1_saved == nullptr;
    const int blocks_rows = wide ? p.blocks_rows : (int)std::min<int64_t>(p.blocks_rows, (rows + kThreads - 1) / kThreads);
    cudaStream_t s = as_stream(stream);
    float* dz1 = workspace + p.off_dz1;
    float* psmall = workspace + p.off_small;
    float* pw1 = workspace + p.off_w1;
    // shared memory: weights, later reused for the block reduction of the small gradients
    const size_t smem = std::max<size_t>(L.P, (size_t)(kThreads / 32) * p.n_small) * sizeof(float);
    int rc = dispatch_omax(O, [&](auto om) {
        constexpr int OM = decltype(om)::value;
        if (wide) {
            if (smem > 48 * 1024) cudaFuncSetAttribute(mlp_backward_rows_kernel<OM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            mlp_backward_rows_kernel<OM, true><<<blocks_rows, kThreads, smem, s>>>(params, x, dy, h1_saved, dz1, dx, psmall, rows, F, O, softmax);
        } else {
            if (smem > 48 * 1024) cudaFuncSetAttribute(mlp_backward_rows_kernel<OM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            mlp_backward_rows_kernel<OM, false><<<blocks_rows, kThreads, smem, s>>>(params, x, dy, h1_saved, dz1, dx, psmall, rows, F, O, softmax);
        }
        return check_launch("mlp_backward_rows_kernel");
    });
    if (rc) return rc;
    dim3 g2(ceil_div(F, kThreads), p.chunks);
    mlp_backward_w1_kernel<<<g2, kThreads, 0, s>>>(x, dz1, pw1, rows, F, p.rows_per_chunk);
    rc = check_launch("mlp_backward_w1_kernel");
    if (rc) return rc;
    mlp_backward_reduce_kernel<<<ceil_div(L.P, 32), 32 * kB3Slices, 0, s>>>(psmall, blocks_rows, pw1, p.chunks, grad, F, O, accumulate);
    return check_launch("mlp_backward_reduce_kernel");
}

// ---- index-input entry points (one_hot(idx, F) observations; see the kernels above)
static int launch_forward_index(const float* params, const int64_t* idx, float* y, float* h1_out, const float* u,
                                int64_t* actions, int64_t rows, int F, int O, int softmax, uint64_t seed, uint64_t counter,
                                bool sample, cudaStream_t s) {
    const MlpLayout L0(0, O);
    const int grid = (int)std::min<int64_t>(kMaxBlocks, (rows + kThreads - 1) / kThreads);
    const size_t smem = L0.P * sizeof(float);
    return dispatch_omax(O, [&](auto om) {
        constexpr int OM = decltype(om)::value;
        if (sample) mlp_forward_index_kernel<OM, true><<<grid, kThreads, smem, s>>>(params, idx, y, h1_out, u, actions, rows, F, O, 1, seed, counter);
        else mlp_forward_index_kernel<OM, false><<<grid, kThreads, smem, s>>>(params, idx, y, h1_out, nullptr, nullptr, rows, F, O, softmax, 0, 0);
        return check_launch("mlp_forward_index_kernel");
    });
}

extern "C" int ia2c_mlp_forward_index(const float* params, const int64_t* idx, float* y, float* h1_out, int64_t rows,
                                      int32_t F, int32_t O, int32_t softmax, void* stream) {
    IA2C_REQUIRE(params && idx && y && rows > 0, "ia2c_mlp_forward_index: null pointer or rows=%lld", (long long)rows);
    IA2C_REQUIRE(F >= 1 && O >= 1 && O <= 32, "ia2c_mlp_forward_index: F=%d O=%d unsupported", F, O);
    return launch_forward_index(params, idx, y, h1_out, nullptr, nullptr, rows, F, O, softmax, 0, 0, false, as_stream(stream));
}

extern "C" int ia2c_actor_sample_index(const float* params, const int64_t* idx, const float* u, int64_t* actions_out,
                                       float* probs_out, int64_t rows, int32_t F, int32_t O, uint64_t seed,
                                       uint64_t counter, void* stream) {
    IA2C_REQUIRE(params && idx && actions_out && rows > 0, "ia2c_actor_sample_index: null pointer or rows=%lld", (long long)rows);
    IA2C_REQUIRE(F >= 1 && O >= 1 && O <= 32, "ia2c_actor_sample_index: F=%d O=%d unsupported", F, O);
    return launch_forward_index(params, idx, probs_out, nullptr, u, actions_out, rows, F, O, 1, seed, counter, true, as_stream(stream));
}

extern "C" int ia2c_mlp_backward_index(const float* params, const int64_t* idx, const float* dy, const float* h1_saved,
                                       float* grad, float* workspace, int64_t rows, int32_t F, int32_t O, int32_t softmax,
                                       int32_t accumulate, void* stream) {
    IA2C_REQUIRE(params && idx && dy && h1_saved && grad && workspace && rows > 0,
                 "ia2c_mlp_backward_index: null pointer or rows=%lld (h1_saved from ia2c_mlp_forward_index is required)", (long long)rows);
    IA2C_REQUIRE(F >= 1 && F <= 8192 && O >= 1 && O <= 32, "ia2c_mlp_backward_index: F=%d O=%d unsupported", F, O);
    const MlpLayout L(F, O);
    const BackwardPlan p = plan_backward(rows, F, O);   // same workspace layout as the dense backward
    const int blocks_rows = (int)std::min<int64_t>(p.blocks_rows, (rows + kThreads - 1) / kThreads);
    cudaStream_t s = as_stream(stream);
    float* dz1 = workspace + p.off_dz1;
    float* psmall = workspace + p.off_small;
    float* pw1 = workspace + p.off_w1;
    const size_t smem = std::max<size_t>(L.P, (size_t)(kThreads / 32) * p.n_small) * sizeof(float);
    int rc = dispatch_omax(O, [&](auto om) {
        constexpr int OM = decltype(om)::value;
        if (smem > 48 * 1024) cudaFuncSetAttribute(mlp_backward_rows_kernel<OM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mlp_backward_rows_kernel<OM, false><<<blocks_rows, kThreads, smem, s>>>(params, nullptr, dy, h1_saved, dz1, nullptr, psmall, rows, F, O, softmax);
        return check_launch("mlp_backward_rows_kernel");
    });
    if (rc) return rc;
    dim3 g2(ceil_div(F, kThreads), p.chunks);
    mlp_backward_w1_index_kernel<<<g2, kThreads, 0, s>>>(idx, dz1, pw1, rows, F, p.rows_per_chunk);
    rc = check_launch("mlp_backward_w1_index_kernel");
    if (rc) return rc;
    mlp_backward_reduce_kernel<<<ceil_div(L.P, 32), 32 * kB3Slices, 0, s>>>(psmall, blocks_rows, pw1, p.chunks, grad, F, O, accumulate);
    return check_launch("mlp_backward_reduce_kernel");
}
