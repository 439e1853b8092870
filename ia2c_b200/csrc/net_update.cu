// net_update.cu — CriticNetwork.batch_update / ActorNetwork.batch_update as ONE call.
//
// Replaces the body of ac_nets.py:62-72 (critic: zero_grad, forward, MSE on the selected output, backward, Adam) and
// ac_nets.py:112-119 (actor: forward, Categorical log-prob / entropy loss, backward WITHOUT zero_grad, Adam) for the case
// where the target / advantage carries no autograd graph (a2c_test.py, a2c_org_test.py's critic, ia2c.py's actors): no
// framework graph, no per-call allocations — the caller hands in one workspace.
//
// Two implementations behind ia2c_net_update:
//   * dense inputs with a wide first layer (the a2c_test.py shape, 500 one-hot features): net_update_dense_kernel, ONE
//     pass over X.  X is the only large operand (rows x F x 4 bytes; 131 MB at 65536 x 500) and the layer-1 weight
//     gradient dW1 = dz1^T X needs it again after the whole forward/backward of the row, so each 32-row tile of X is
//     brought into shared memory ONCE by a bulk async copy (cp.async.bulk + mbarrier, 3 tiles in flight per SM) and is
//     used twice from there: forward dot products (lane owns features f = lane mod 32, its 6 x F/32 slice of W1 in
//     registers, packed FFMA2) and, after the rows' tail (layers 2-3, loss, backward to dz1, lanes over the inner
//     dimensions), the rank-1 updates of dW1 (thread owns feature columns, accumulators in registers).  HBM-bound:
//     the roofline is rows*F*4 bytes / measured copy bandwidth.
//   * everything else (index inputs, narrow first layers): the existing forward / loss / backward / Adam kernels in
//     sequence inside this call.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

using namespace ia2c;

namespace ia2c {
namespace {

// ------------------------------------------------------------------------------------------------ single-pass dense kernel
constexpr int kUThreads = 256, kUWarps = kUThreads / 32;
constexpr int kURows = 32;             // rows per tile: 4 per warp, 8 lanes per row in the tail
constexpr int kUStages = 3;            // tiles in shared memory (kUStages - 1 in flight while one is computed on)
constexpr int kUTF = 16;               // lane-owned feature steps: F <= 32 * kUTF = 512
constexpr int kUOMax = 8;              // outputs (8 lanes per row in the tail)
constexpr int kRec = 36;               // floats per row record: dz1[0..7] h1[8..13] h2[14..19] dz2[20..25] one[26] loss[27] dz3[28..35]
constexpr float kEpsClampU = 1.1920928955078125e-07f;  // torch.finfo(float32).eps used by clamp_probs

struct DenseArgs {
    const float* x;
    const float* params;
    const int32_t* act;
    const float* signal;
    float* partials;       // [grid][P + 1]: this block's gradient sums, then its loss sum
    int32_t* status;
    int64_t rows;
    int F, O, n_tiles;
    float beta, inv_b;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// bulk async copy global -> shared (TMA engine, no tensor map: one contiguous run), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar) : "memory");
}

// sum over the 8 lanes of an aligned group (xor offsets 4, 2, 1 stay inside it)
__device__ __forceinline__ float group8_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
__device__ __forceinline__ float group8_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return v;
}

// KIND 0: critic (MSE on the selected output), 1: actor (softmax + Categorical log-prob / entropy loss)
template <int KIND>
__global__ void __launch_bounds__(kUThreads, 1) net_update_dense_kernel(const __grid_constant__ DenseArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int F = A.F, O = A.O;
    const MlpLayout L(F, O);
    const int n_small = L.P - H * F;                               // b1, W2, b2, W3, b3
    const uint32_t tile_bytes = (uint32_t)kURows * (uint32_t)F * 4u;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);            // [kUStages]
    float* xs = reinterpret_cast<float*>(smem + 128);              // [kUStages][kURows][F]
    float* wsm = xs + (size_t)kUStages * kURows * F;               // [n_small] small parameters
    float* rec0 = wsm + ((n_small + 3) & ~3);                      // [2][kURows][kRec]: row records, double-buffered
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = lane & 7, grp = lane & ~7;                       // tail: 8 lanes per row, c = inner index
    pdl_release();                                                 // the reduce + Adam kernel may be scheduled; it waits for this grid
    if (tid == 0) {
        for (int s = 0; s < kUStages; ++s) mbar_init(smem_u32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int k = tid; k < n_small; k += kUThreads) wsm[k] = A.params[H * F + k];
    __syncthreads();
    auto tile_rows = [&](int tile) -> int { return (int)min((int64_t)kURows, A.rows - (int64_t)tile * kURows); };
    auto load_tile = [&](int tile, int stage) {   // thread 0 only
        const uint32_t bytes = (uint32_t)tile_rows(tile) * (uint32_t)F * 4u;
        const uint32_t bar = smem_u32(bars + stage);
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(xs) + (uint32_t)stage * tile_bytes, A.x + (int64_t)tile * kURows * F, bytes, bar);
    };
    if (tid == 0) {
        for (int s = 0; s < kUStages; ++s) {
            const int tile = blockIdx.x + s * gridDim.x;
            if (tile < A.n_tiles) load_tile(tile, s);
        }
    }
    // ---- per-thread constants
    // phase 1: lane owns features f = lane + 32 t; its slice of W1 as output pairs (j, j+1)
    float2 w1r[kUTF][3];
    uint32_t fmask = 0u;
#pragma unroll
    for (int t = 0; t < kUTF; ++t) {
        const int f = lane + 32 * t;
        const bool ok = f < F;
        fmask |= ok ? (1u << t) : 0u;
#pragma unroll
        for (int pz = 0; pz < 3; ++pz)
            w1r[t][pz] = ok ? make_float2(A.params[(2 * pz) * F + f], A.params[(2 * pz + 1) * F + f]) : make_float2(0.f, 0.f);
    }
    // tail: lane c of a group holds row c of W2 / W3 (forward) and column c (backward)
    const int o_b1 = 0, o_w2 = H, o_b2 = H + H * H, o_w3 = 2 * H + H * H, o_b3 = o_w3 + O * H;
    float w2row[H], w2col[H], w3row[H], w3col[kUOMax];
#pragma unroll
    for (int k = 0; k < H; ++k) {
        w2row[k] = c < H ? wsm[o_w2 + c * H + k] : 0.f;
        w2col[k] = c < H ? wsm[o_w2 + k * H + c] : 0.f;
        w3row[k] = c < O ? wsm[o_w3 + c * H + k] : 0.f;
    }
#pragma unroll
    for (int o = 0; o < kUOMax; ++o) w3col[o] = (c < H && o < O) ? wsm[o_w3 + o * H + c] : 0.f;
    const float b1c = c < H ? wsm[o_b1 + c] : 0.f, b2c = c < H ? wsm[o_b2 + c] : 0.f, b3c = c < O ? wsm[o_b3 + c] : 0.f;
    // phase 2: thread owns feature columns tid and tid + 256; and up to two of the n_small + 1 small sums
    float2 gacc[2][3];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int pz = 0; pz < 3; ++pz) gacc[q][pz] = make_float2(0.f, 0.f);
    const bool fa_ok = tid < F, fb_ok = tid + kUThreads < F;
    auto entry = [&](int e, int& a_off, int& b_off) {   // small sum e = sum_r rec[r][a_off] * rec[r][b_off]
        if (e < H) { a_off = e; b_off = 26; }                                                  // gb1[j]   = sum dz1[j]
        else if (e < o_b2) { const int i = e - o_w2; a_off = 20 + i / H; b_off = 8 + i % H; }    // gW2[j][k] = sum dz2[j] h1[k]
        else if (e < o_w3) { a_off = 20 + (e - o_b2); b_off = 26; }                             // gb2[j]
        else if (e < o_b3) { const int i = e - o_w3; a_off = 28 + i / H; b_off = 14 + i % H; }  // gW3[o][k] = sum dz3[o] h2[k]
        else if (e < n_small) { a_off = 28 + (e - o_b3); b_off = 26; }                          // gb3[o]
        else { a_off = 27; b_off = 26; }                                                        // loss
    };
    int ea0 = 26, eb0 = 26, ea1 = 26, eb1 = 26;
    const bool e0_ok = tid <= n_small, e1_ok = tid + kUThreads <= n_small;
    if (e0_ok) entry(tid, ea0, eb0);
    if (e1_ok) entry(tid + kUThreads, ea1, eb1);
    float sacc0 = 0.f, sacc1 = 0.f;

    int use = 0;
    for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++use) {
        const int stage = use % kUStages;
        const uint32_t parity = (uint32_t)(use / kUStages) & 1u;
        const int nr = tile_rows(tile);
        const float* xt = xs + (size_t)stage * kURows * F;
        float* rec = rec0 + (use & 1) * (kURows * kRec);
        mbar_wait(smem_u32(bars + stage), parity);
        // ---------------- phase 1: z1 partial sums of the warp's four rows over the lane's features
        float2 acc[4][3];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int pz = 0; pz < 3; ++pz) acc[rr][pz] = make_float2(0.f, 0.f);
        const float* xw = xt + (size_t)(warp * 4) * F + lane;
#pragma unroll
        for (int t = 0; t < kUTF; ++t) {
            if (32 * t < F) {   // block-uniform
                const bool ok = (fmask >> t) & 1u;
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const float xv = ok ? xw[(size_t)rr * F + 32 * t] : 0.f;
#pragma unroll
                    for (int pz = 0; pz < 3; ++pz) acc[rr][pz] = __ffma2_rn(w1r[t][pz], make_float2(xv, xv), acc[rr][pz]);
                }
            }
        }
        // transposed warp reduction of 4 rows x 8 slots (6 used): afterwards lane l holds the total of slot l = 8 * row + j
        float v[32];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
            for (int pz = 0; pz < 3; ++pz) { v[8 * rr + 2 * pz] = acc[rr][pz].x; v[8 * rr + 2 * pz + 1] = acc[rr][pz].y; }
            v[8 * rr + 6] = 0.f;
            v[8 * rr + 7] = 0.f;
        }
#pragma unroll
        for (int off = 16, n = 32; off > 0; off >>= 1, n >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (k < n / 2) {
                    const float keep = upper ? v[k + n / 2] : v[k];
                    const float send = upper ? v[k] : v[k + n / 2];
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
        }
        const float z1 = v[0];
        // ---------------- tail: layers 2-3, loss, backward to dz1 — 8 lanes per row, lane c = inner index
        const int row_t = warp * 4 + (lane >> 3);                  // row within the tile
        const bool row_ok = row_t < nr;
        const int64_t row = (int64_t)tile * kURows + row_t;
        const float h1 = c < H ? fmaxf(z1 + b1c, 0.f) : 0.f;
        float z2 = b2c;
#pragma unroll
        for (int j = 0; j < H; ++j) z2 = fmaf(w2row[j], __shfl_sync(0xffffffffu, h1, grp + j), z2);
        const float h2 = c < H ? fmaxf(z2, 0.f) : 0.f;
        float y = b3c;
#pragma unroll
        for (int k = 0; k < H; ++k) y = fmaf(w3row[k], __shfl_sync(0xffffffffu, h2, grp + k), y);
        const int a = row_ok ? A.act[row] : 0;
        const float sig = row_ok ? A.signal[row] : 0.f;
        float dz3, loss_row;
        if (KIND == 0) {
            const float ya = __shfl_sync(0xffffffffu, y, grp + (a & 7));
            const float delta = sig - ya;                               // ia2c_critic_loss
            loss_row = delta * delta;
            dz3 = (c == a) ? -(2.f * delta * A.inv_b) : 0.f;
        } else {
            // softmax (tail_forward), then Categorical(probs=p) loss and its gradient (ia2c_actor_loss), then back through the softmax
            const float m = group8_max(c < O ? y : -INFINITY);
            const float e = c < O ? expf(y - m) : 0.f;
            const float inv = 1.f / group8_sum(e);
            const float pr = e * inv;
            const float sp = group8_sum(pr);
            const bool bad = (c < O && !(pr >= 0.f)) || !(fabsf(sp - 1.f) < 1e-6f);
            if (bad && row_ok) *A.status = 1;
            const float q = pr / sp;
            const bool inside = (q >= kEpsClampU) && (q <= 1.f - kEpsClampU);
            const float logit = logf(fminf(fmaxf(q, kEpsClampU), 1.f - kEpsClampU));
            const float ent = -group8_sum(c < O ? logit * q : 0.f);
            float go = A.beta * (logit + (inside ? 1.f : 0.f));
            if (c == a && inside) go -= sig / q;
            const float neglogp = -__shfl_sync(0xffffffffu, logit, grp + (a & 7));
            const float qg = group8_sum(c < O ? q * go : 0.f);
            loss_row = sig * neglogp - A.beta * ent;
            const float dpr = c < O ? (go - qg) / sp * A.inv_b : 0.f;   // dL/dprobs
            const float dot = group8_sum(pr * dpr);
            dz3 = pr * (dpr - dot);                                     // through the softmax
        }
        if (!row_ok) { dz3 = 0.f; loss_row = 0.f; }
        float dh2 = 0.f;
#pragma unroll
        for (int o = 0; o < kUOMax; ++o) dh2 = fmaf(w3col[o], __shfl_sync(0xffffffffu, dz3, grp + o), dh2);
        const float dz2 = (c < H && h2 > 0.f) ? dh2 : 0.f;
        float dh1 = 0.f;
#pragma unroll
        for (int k = 0; k < H; ++k) dh1 = fmaf(w2col[k], __shfl_sync(0xffffffffu, dz2, grp + k), dh1);
        const float dz1 = (c < H && h1 > 0.f) ? dh1 : 0.f;
        {
            float* rr = rec + row_t * kRec;
            rr[c] = dz1;                                   // slots 6, 7 = 0
            rr[28 + c] = dz3;
            if (c < H) { rr[8 + c] = row_ok ? h1 : 0.f; rr[14 + c] = row_ok ? h2 : 0.f; rr[20 + c] = dz2; }
            if (c == 6) rr[26] = 1.f;
            if (c == 7) rr[27] = loss_row;
        }
        __syncthreads();   // the ONLY block barrier per tile: records visible; every thread has finished phase 2 of the previous tile
        if (tid == 0 && use > 0) {
            const int next = tile + (kUStages - 1) * gridDim.x;       // goes into the stage of the previous tile
            if (next < A.n_tiles) load_tile(next, (use - 1) % kUStages);
        }
        // ---------------- phase 2: rank-1 updates of dW1 from the staged tile (thread = feature column) + the small sums
#pragma unroll 4
        for (int r = 0; r < nr; ++r) {
            const float* rr = rec + r * kRec;
            const float4 d03 = *reinterpret_cast<const float4*>(rr);
            const float2 d45 = *reinterpret_cast<const float2*>(rr + 4);
            const float xa = fa_ok ? xt[(size_t)r * F + tid] : 0.f;
            const float xb = fb_ok ? xt[(size_t)r * F + tid + kUThreads] : 0.f;
            gacc[0][0] = __ffma2_rn(make_float2(d03.x, d03.y), make_float2(xa, xa), gacc[0][0]);
            gacc[0][1] = __ffma2_rn(make_float2(d03.z, d03.w), make_float2(xa, xa), gacc[0][1]);
            gacc[0][2] = __ffma2_rn(d45, make_float2(xa, xa), gacc[0][2]);
            gacc[1][0] = __ffma2_rn(make_float2(d03.x, d03.y), make_float2(xb, xb), gacc[1][0]);
            gacc[1][1] = __ffma2_rn(make_float2(d03.z, d03.w), make_float2(xb, xb), gacc[1][1]);
            gacc[1][2] = __ffma2_rn(d45, make_float2(xb, xb), gacc[1][2]);
        }
        if (warp <= n_small / 32) {   // warp-uniform: the warps that own small sums (n_small + 1 <= 2 * 256 entries)
#pragma unroll 4
            for (int r = 0; r < nr; ++r) {
                const float* rr = rec + r * kRec;
                sacc0 = fmaf(rr[ea0], rr[eb0], sacc0);
                if (n_small >= kUThreads) sacc1 = fmaf(rr[ea1], rr[eb1], sacc1);
            }
        }
    }
    // ---- this block's sums
    float* out = A.partials + (size_t)blockIdx.x * (L.P + 1);
#pragma unroll
    for (int pz = 0; pz < 3; ++pz) {
        if (fa_ok) { out[(2 * pz) * F + tid] = gacc[0][pz].x; out[(2 * pz + 1) * F + tid] = gacc[0][pz].y; }
        if (fb_ok) { out[(2 * pz) * F + tid + kUThreads] = gacc[1][pz].x; out[(2 * pz + 1) * F + tid + kUThreads] = gacc[1][pz].y; }
    }
    if (e0_ok) out[H * F + tid] = sacc0;
    if (e1_ok) out[H * F + tid + kUThreads] = sacc1;
}

// grad[i] (+)= sum over blocks (fixed order) of partials[b][i], then the Adam step on that entry (arithmetic order of
// torch's _single_tensor_adam, as adam_kernel in loss.cu); entry P is the loss sum -> loss_out = sum / rows.  Launched with
// programmatic dependent launch: the Adam constants are computed while the gradient kernel drains.
__global__ void __launch_bounds__(256) net_update_reduce_adam_kernel(const float* __restrict__ partials, int n_blocks, int P,
                                                                     float* __restrict__ grad, int accumulate, float inv_b,
                                                                     float* __restrict__ loss_out, float* __restrict__ params,
                                                                     float* __restrict__ m, float* __restrict__ v,
                                                                     int32_t* __restrict__ step_count, double lr) {
    pdl_release();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double b1 = 0.9, b2 = 0.999;
    pdl_wait();
    if (i > P) return;
    float s = 0.f;
    for (int b = 0; b < n_blocks; ++b) s += partials[(size_t)b * (P + 1) + i];
    if (i == P) {
        *loss_out = s * inv_b;
        return;
    }
    const float g = accumulate ? grad[i] + s : s;
    grad[i] = g;
    const int t = *step_count + 1;                         // bumped by net_update_bump_kernel afterwards
    const double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
    const float step_size = (float)(lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    const float w1 = (float)(1.0 - b1), w2 = (float)(1.0 - b2), b2f = (float)b2, eps = 1e-8f;
    const float mi = m[i] + (g - m[i]) * w1;
    const float vi = v[i] * b2f + w2 * g * g;
    m[i] = mi;
    v[i] = vi;
    params[i] = params[i] - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
}
__global__ void net_update_bump_kernel(int32_t* step_count) {
    pdl_prologue();
    if (threadIdx.x == 0 && blockIdx.x == 0) *step_count += 1;
}

// index inputs: flag class values outside [0, F) (bit 1 of *status) — the gather kernels clamp, so nothing is read out of bounds
__global__ void index_check_kernel(const int64_t* __restrict__ idx, int64_t rows, int F, int32_t* __restrict__ status) {
    bool bad = false;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = idx[r];
        bad |= v < 0 || v >= F;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(status, 2);
}

size_t dense_smem_bytes(int F, int O) {
    const MlpLayout L(F, O);
    const int n_small = L.P - H * F;
    return 128 + (size_t)kUStages * kURows * F * 4 + (size_t)((n_small + 3) & ~3) * 4 + (size_t)2 * kURows * kRec * 4;
}
bool dense_applicable(const float* x, int64_t rows, int F, int O) {
    return x && rows >= 1024 && F >= 64 && F <= 32 * kUTF && (F & 3) == 0 && O <= kUOMax && ((uintptr_t)x & 15) == 0 &&
           dense_smem_bytes(F, O) <= 227 * 1024;
}
int dense_grid(int64_t rows) { return (int)std::min<int64_t>(kSMs, (rows + kURows - 1) / kURows); }

}  // namespace
}  // namespace ia2c

namespace {
struct Ws {
    size_t y, h1, dy, back, loss, partials, total;   // float offsets
};
Ws plan_ws(int64_t rows, int F, int O) {
    Ws w;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 3) & ~size_t(3); return o; };
    w.y = take((size_t)rows * O);
    w.h1 = take((size_t)rows * IA2C_HIDDEN);
    w.dy = take((size_t)rows * O);
    w.back = take(ia2c_mlp_backward_workspace(rows, F, O));
    w.loss = take(ia2c_loss_workspace(rows));
    w.partials = take((size_t)kSMs * ((size_t)IA2C_HIDDEN * F + 64 + 7 * (size_t)O));   // single-pass kernel: [blocks][P + 1]
    w.total = off;
    return w;
}
}  // namespace

extern "C" size_t ia2c_net_update_workspace(int64_t rows, int32_t F, int32_t O) {
    if (rows <= 0 || F < 1 || O < 1) return 0;
    return plan_ws(rows, F, O).total;
}

extern "C" int ia2c_net_update(int32_t kind, float* params, float* grad, float* exp_avg, float* exp_avg_sq, int32_t* step_count,
                               const float* x, const int64_t* idx, const int32_t* act, const float* signal, float beta, double lr,
                               float* loss_out, int32_t* status_out, float* workspace, int64_t rows, int32_t F, int32_t O,
                               void* stream) {
    IA2C_REQUIRE(kind == 0 || kind == 1, "ia2c_net_update: kind=%d (0 critic, 1 actor)", kind);
    IA2C_REQUIRE(params && grad && exp_avg && exp_avg_sq && step_count && act && signal && loss_out && workspace && rows > 0,
                 "ia2c_net_update: null pointer or rows=%lld", (long long)rows);
    IA2C_REQUIRE((x != nullptr) != (idx != nullptr), "ia2c_net_update: exactly one of x (dense rows) and idx (one-hot class indices)");
    IA2C_REQUIRE(kind == 0 || status_out, "ia2c_net_update: the actor update needs status_out");
    IA2C_REQUIRE(F >= 1 && F <= 8192 && O >= 1 && O <= 32, "ia2c_net_update: F=%d O=%d unsupported", F, O);
    const Ws w = plan_ws(rows, F, O);
    float* y = workspace + w.y;
    float* h1 = workspace + w.h1;
    float* dy = workspace + w.dy;
    const int softmax = kind;                 // the actor's output layer is a softmax (ac_nets.py:41)
    const int accumulate = kind;              // the actor never zeroes its gradient (ac_nets.py:112-119)
    const int P = IA2C_HIDDEN * F + IA2C_HIDDEN + IA2C_HIDDEN * IA2C_HIDDEN + IA2C_HIDDEN + O * IA2C_HIDDEN + O;
    int rc;
    if (dense_applicable(x, rows, F, O) && !getenv("IA2C_NO_SINGLE_PASS")) {   // wide dense rows: ONE pass over x
        DenseArgs A;
        A.x = x; A.params = params; A.act = act; A.signal = signal;
        A.partials = workspace + w.partials;
        A.status = status_out;
        A.rows = rows; A.F = F; A.O = O;
        A.n_tiles = (int)((rows + kURows - 1) / kURows);
        A.beta = beta;
        A.inv_b = 1.f / (float)rows;
        const int grid = dense_grid(rows);
        const size_t smem = dense_smem_bytes(F, O);
        cudaStream_t s = as_stream(stream);
        if (kind == 0) {
            cudaFuncSetAttribute(net_update_dense_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            net_update_dense_kernel<0><<<grid, kUThreads, smem, s>>>(A);
        } else {
            cudaFuncSetAttribute(net_update_dense_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            net_update_dense_kernel<1><<<grid, kUThreads, smem, s>>>(A);
        }
        if ((rc = check_launch("net_update_dense_kernel"))) return rc;
        if ((rc = launch_pdl("net_update_reduce_adam_kernel", net_update_reduce_adam_kernel, dim3(ceil_div(P + 1, 256)), dim3(256), 0, s,
                             A.partials, grid, P, grad, accumulate, A.inv_b, loss_out, params, exp_avg, exp_avg_sq, step_count, lr)))
            return rc;
        return launch_pdl("net_update_bump_kernel", net_update_bump_kernel, dim3(1), dim3(32), 0, s, step_count);
    }
    if (idx && status_out) {
        index_check_kernel<<<(int)std::min<int64_t>(kSMs * 4, (rows + 255) / 256), 256, 0, as_stream(stream)>>>(idx, rows, F, status_out);
        if ((rc = check_launch("index_check_kernel"))) return rc;
    }
    if (x) rc = ia2c_mlp_forward(params, x, y, h1, rows, F, O, 1, softmax, stream);
    else rc = ia2c_mlp_forward_index(params, idx, y, h1, rows, F, O, softmax, stream);
    if (rc) return rc;
    if (kind == 0) rc = ia2c_critic_loss(y, act, signal, loss_out, dy, nullptr, workspace + w.loss, rows, O, stream);
    else rc = ia2c_actor_loss(y, act, signal, beta, loss_out, dy, nullptr, status_out, workspace + w.loss, rows, O, stream);
    if (rc) return rc;
    if (x) rc = ia2c_mlp_backward(params, x, dy, h1, grad, nullptr, workspace + w.back, rows, F, O, softmax, accumulate, stream);
    else rc = ia2c_mlp_backward_index(params, idx, dy, h1, grad, workspace + w.back, rows, F, O, softmax, accumulate, stream);
    if (rc) return rc;
    return ia2c_adam_step(params, grad, nullptr, exp_avg, exp_avg_sq, step_count, lr, 0.9, 0.999, 1e-8, 1, P, stream);
}
