// coach_b200/csrc/learn.cu -- element-wise and reduction kernels of learn_from_batch (TD targets, head losses,
// global-norm clipping, TF-semantics Adam, polyak target update).  Reference lines are cited in include/coach_b200.h.
#include <math.h>

#include "common.cuh"

namespace cb200 {

// ---- DQN / DDQN TD targets ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dqn_td_targets_kernel(const float* __restrict__ q_next,
                                                             const float* __restrict__ q_select,
                                                             const float* __restrict__ q_online,
                                                             const int64_t* __restrict__ actions,
                                                             const double* __restrict__ rewards,
                                                             const uint8_t* __restrict__ game_overs, double discount,
                                                             int64_t B, int64_t A, float* __restrict__ targets,
                                                             double* __restrict__ td_err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    // np.argmax: first maximum
    int64_t best = 0;
    float bv = q_select[i * A];
    for (int64_t a = 1; a < A; ++a) {
        const float v = q_select[i * A + a];
        if (v > bv) {
            bv = v;
            best = a;
        }
    }
    const int64_t act = actions[i];
    for (int64_t a = 0; a < A; ++a) targets[i * A + a] = q_online[i * A + a];
    // new_target = r + (1.0 - done) * discount * q'[a*]     (dqn_agent.py:100-101; left-to-right, fp64)
    const double not_done = __dsub_rn(1.0, game_overs[i] ? 1.0 : 0.0);
    const double t1 = __dmul_rn(__dmul_rn(not_done, discount), (double)q_next[i * A + best]);
    const double y = __dadd_rn(rewards[i], t1);
    if (act >= 0 && act < A) {
        td_err[i] = fabs(__dsub_rn(y, (double)q_online[i * A + act]));   // :102
        targets[i * A + act] = (float)y;                                 // :103 (fp32 array element assignment)
    } else {
        td_err[i] = 0.0;
    }
}

// ---- regression head loss (Huber / MSE) -----------------------------------------------------------------------------
// one CTA; fixed-order reduction => run-to-run bit-stable loss
__global__ void __launch_bounds__(1024) regression_head_kernel(const float* __restrict__ out,
                                                               const float* __restrict__ target,
                                                               const float* __restrict__ weights, int64_t B, int64_t W,
                                                               int huber, float loss_weight, float* __restrict__ d_out,
                                                               float* __restrict__ loss_out) {
    __shared__ float red[1024];
    float local = 0.f;
    const float inv_b = 1.0f / (float)B;
    for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
        const float w = loss_weight * (weights ? weights[b] : 1.0f);
        float row = 0.f;
        for (int64_t a = 0; a < W; ++a) {
            const float e = out[b * W + a] - target[b * W + a];     // predictions - labels
            float l, g;
            if (huber) {
                const float ae = fabsf(e);
                const float q = fminf(ae, 1.0f);                    // delta = 1
                const float lin = ae - q;
                l = 0.5f * q * q + lin;
                g = (ae <= 1.0f) ? e : (e > 0.f ? 1.0f : -1.0f);
            } else {
                l = e * e;
                g = 2.0f * e;
            }
            row += l;
            d_out[b * W + a] = w * inv_b * g;
        }
        local += w * row;
    }
    red[threadIdx.x] = local;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0 && loss_out) *loss_out = red[0] * inv_b;
}

// ---- fused DQN / DDQN Q-head step -----------------------------------------------------------------------------------
// Everything between the feature layer and the loss in ONE launch (+ one small reduction launch):
//   Q(s') of the target head, Q(s) of the online head [, Q(s') of the online head: DDQN action selection]
//   (q_head.py:52-54: Dense(num_actions) on the middleware output)  ->  TD targets / errors (dqn_agent.py:92-103,
//   fp64, bit-exact given the Q values)  ->  Huber / MSE head loss and dL/dQ (head.py:165-177)  ->  the head's own
//   backward: dL/dW = h^T dQ, dL/db = sum_b dQ, and the gradient w.r.t. the feature layer's pre-activation
//   dL/dz = (dQ W^T) * relu'(h), written as fp32 and as operand planes for the feature layer's tensor-core GEMMs.
// Replaces 8-9 launches of latency-bound kernels (three 512x512x6 skinny GEMMs, TD targets, loss, two backward GEMMs, a
// transpose) whose arithmetic is 12 MFLOP in total.
// One warp = kHeadRows batch rows; lane l owns features [l K/32, (l+1) K/32): its slice of a row of h is contiguous, so
// is its slice of W [K, A] (row-major), and the dL/dz planes get whole 16-byte core rows.  Dot products are reduced with
// an xor butterfly (every lane ends with the same bits); batch-wise sums (dW, db, loss) go through per-warp partials in
// global memory and a fixed-order second pass: run-to-run identical bits.
// exact 3-way bf16 truncation split of 8 fp32 values into three 16-byte groups (same arithmetic as gemm::split8 of
// nn_gemm_tc.cuh, restated here because that header defines kernels) and the core-tiled element index of nn_gemm.cuh
struct HeadSplit8 {
    uint4 h, m, l;
};
__device__ __forceinline__ HeadSplit8 head_split8(const float (&x)[8]) {
    uint32_t hb[8], mb[8], lb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        hb[j] = __float_as_uint(x[j]) & 0xffff0000u;
        const float r1 = x[j] - __uint_as_float(hb[j]);
        mb[j] = __float_as_uint(r1) & 0xffff0000u;
        lb[j] = __float_as_uint(r1 - __uint_as_float(mb[j]));
    }
    HeadSplit8 s;
    s.h = make_uint4(__byte_perm(hb[0], hb[1], 0x7632), __byte_perm(hb[2], hb[3], 0x7632),
                     __byte_perm(hb[4], hb[5], 0x7632), __byte_perm(hb[6], hb[7], 0x7632));
    s.m = make_uint4(__byte_perm(mb[0], mb[1], 0x7632), __byte_perm(mb[2], mb[3], 0x7632),
                     __byte_perm(mb[4], mb[5], 0x7632), __byte_perm(mb[6], mb[7], 0x7632));
    s.l = make_uint4(__byte_perm(lb[0], lb[1], 0x7632), __byte_perm(lb[2], lb[3], 0x7632),
                     __byte_perm(lb[4], lb[5], 0x7632), __byte_perm(lb[6], lb[7], 0x7632));
    return s;
}
__device__ __forceinline__ size_t head_tiled_elem(size_t prow, int col, int pcols) {
    return ((prow >> 3) * (size_t)(pcols >> 3) + (size_t)(col >> 3)) * 64 + (prow & 7) * 8 + (col & 7);
}

constexpr int kHeadMaxA = 8;
constexpr int kHeadRows = 2;
constexpr int kHeadWarps = 8;

struct HeadParams {
    const float *h_next, *h_online, *h_select;
    const float *w_target, *b_target, *w_online, *b_online;
    const int64_t* actions;
    const double* rewards;
    const uint8_t* game_overs;
    const float* weights;
    double discount;
    int huber, B, K, A;
    float *q_online, *q_next, *targets;
    double* td_err;
    float *dq, *loss, *dh;
    uint16_t* dh_planes;
    int64_t dh_plane_stride;
    float *dw, *db, *workspace;
};

// Lane l owns features k = l + 32 j (j < KPL): a row of h is read with coalesced 128-byte loads and the head kernels,
// staged TRANSPOSED in shared memory (Wt[a][k]), are read conflict-free.  The row's dL/dz goes through a per-warp
// shared-memory row so that it leaves as whole 16-byte core rows / float4s.
template <int KPL>
__device__ __forceinline__ void head_dot(const float* __restrict__ hrow, const float* __restrict__ wt /* smem [A][K] */,
                                         const float* __restrict__ bias, int A, int K, int lane, float (&hv)[KPL],
                                         float (&q)[kHeadMaxA]) {
#pragma unroll
    for (int j = 0; j < KPL; ++j) hv[j] = __ldg(hrow + lane + 32 * j);
#pragma unroll
    for (int a = 0; a < kHeadMaxA; ++a) {
        q[a] = 0.f;
        if (a < A) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < KPL; ++j) s = fmaf(hv[j], wt[a * K + lane + 32 * j], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            q[a] = s + __ldg(bias + a);
        }
    }
}

template <int KPL>
__global__ void __launch_bounds__(32 * kHeadWarps) dqn_head_fused_kernel(HeadParams p) {
    extern __shared__ __align__(16) float head_smem[];     // Wt_online [A][K] | Wt_target [A][K] | row buffers [warps][K]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kHeadWarps + warp;                        // global warp
    const int A = p.A, K = p.K;
    float* wt_on = head_smem;
    float* wt_tg = head_smem + A * K;
    float* rowbuf = head_smem + 2 * A * K + warp * K;
    for (int i = threadIdx.x; i < A * K; i += blockDim.x) {               // W [K, A] row-major -> Wt [A][K]
        const int k = i / A, a = i - k * A;
        wt_on[a * K + k] = __ldg(p.w_online + i);
        wt_tg[a * K + k] = __ldg(p.w_target + i);
    }
    __syncthreads();
    float acc_w[KPL][kHeadMaxA];                                          // this lane's slice of dW, over the warp's rows
    float acc_b[kHeadMaxA], acc_loss = 0.f;
#pragma unroll
    for (int j = 0; j < KPL; ++j)
#pragma unroll
        for (int a = 0; a < kHeadMaxA; ++a) acc_w[j][a] = 0.f;
#pragma unroll
    for (int a = 0; a < kHeadMaxA; ++a) acc_b[a] = 0.f;
    const float inv_b = 1.0f / (float)p.B;
    const bool use_sel = p.h_select != nullptr;
    for (int rr = 0; rr < kHeadRows; ++rr) {
        const int r = gw * kHeadRows + rr;
        if (r >= p.B) break;
        float hv[KPL], qn[kHeadMaxA], qs[kHeadMaxA], qo[kHeadMaxA];
        head_dot<KPL>(p.h_next + (size_t)r * K, wt_tg, p.b_target, A, K, lane, hv, qn);
        if (use_sel) head_dot<KPL>(p.h_select + (size_t)r * K, wt_on, p.b_online, A, K, lane, hv, qs);
        head_dot<KPL>(p.h_online + (size_t)r * K, wt_on, p.b_online, A, K, lane, hv, qo);    // hv = h_online slice
        // ---- TD target (every lane, identical values) -- dqn_agent.py:92-103 ------------------------------------------
        int best = 0;
        float bv = use_sel ? qs[0] : qn[0];                               // ddqn_agent.py:42-43 / dqn_agent.py:78-79
        float q_best = qn[0];
#pragma unroll
        for (int a = 1; a < kHeadMaxA; ++a) {
            const float sv = use_sel ? qs[a] : qn[a];
            if (a < A && sv > bv) {                                       // np.argmax: first maximum
                bv = sv;
                best = a;
                q_best = qn[a];
            }
        }
        (void)best;
        const int64_t act = p.actions[r];
        const double not_done = __dsub_rn(1.0, p.game_overs[r] ? 1.0 : 0.0);
        const double y = __dadd_rn(p.rewards[r], __dmul_rn(__dmul_rn(not_done, p.discount), (double)q_best));
        float tgt[kHeadMaxA], dq[kHeadMaxA];
        double td = 0.0;
#pragma unroll
        for (int a = 0; a < kHeadMaxA; ++a) {
            tgt[a] = qo[a];
            if (a < A && a == act) {
                td = fabs(__dsub_rn(y, (double)qo[a]));
                tgt[a] = (float)y;
            }
        }
        // ---- head loss and dL/dQ (head.py:165-177; tf.losses.huber_loss delta = 1 / mean_squared_error) ---------------
        const float w = p.weights ? p.weights[r] : 1.0f;
        float row = 0.f;
#pragma unroll
        for (int a = 0; a < kHeadMaxA; ++a) {
            dq[a] = 0.f;
            if (a < A) {
                const float e = qo[a] - tgt[a];
                float l, g;
                if (p.huber) {
                    const float ae = fabsf(e);
                    const float qq = fminf(ae, 1.0f);
                    l = 0.5f * qq * qq + (ae - qq);
                    g = (ae <= 1.0f) ? e : (e > 0.f ? 1.0f : -1.0f);
                } else {
                    l = e * e;
                    g = 2.0f * e;
                }
                row += l;
                dq[a] = w * inv_b * g;
            }
        }
        acc_loss += w * row;
        if (lane == 0) {
#pragma unroll
            for (int a = 0; a < kHeadMaxA; ++a) {
                if (a < A) {
                    p.q_online[(size_t)r * A + a] = qo[a];
                    if (p.q_next) p.q_next[(size_t)r * A + a] = qn[a];
                    p.targets[(size_t)r * A + a] = tgt[a];
                    p.dq[(size_t)r * A + a] = dq[a];
                }
            }
            p.td_err[r] = td;
        }
        // ---- backward of the head for this row ---------------------------------------------------------------------------
        __syncwarp();                                                     // the previous row's readers are done
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            float s = 0.f;
#pragma unroll
            for (int a = 0; a < kHeadMaxA; ++a)
                if (a < A) {
                    s = fmaf(dq[a], wt_on[a * K + lane + 32 * j], s);
                    acc_w[j][a] = fmaf(hv[j], dq[a], acc_w[j][a]);
                }
            rowbuf[lane + 32 * j] = hv[j] > 0.f ? s : 0.f;                // relu'(h) on the post-activation value
        }
#pragma unroll
        for (int a = 0; a < kHeadMaxA; ++a) acc_b[a] += dq[a];
        __syncwarp();
        // lane l now takes the KPL consecutive features [l KPL, (l + 1) KPL) of the row
        float dz[KPL];
#pragma unroll
        for (int j = 0; j < KPL / 4; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(rowbuf + lane * KPL + 4 * j);
            dz[4 * j] = v.x; dz[4 * j + 1] = v.y; dz[4 * j + 2] = v.z; dz[4 * j + 3] = v.w;
        }
        if (p.dh) {
            float4* o = reinterpret_cast<float4*>(p.dh + (size_t)r * K + lane * KPL);
#pragma unroll
            for (int j = 0; j < KPL / 4; ++j) o[j] = make_float4(dz[4 * j], dz[4 * j + 1], dz[4 * j + 2], dz[4 * j + 3]);
        }
        if (p.dh_planes) {
#pragma unroll
            for (int c8 = 0; c8 < KPL / 8; ++c8) {
                const float x8[8] = {dz[8 * c8], dz[8 * c8 + 1], dz[8 * c8 + 2], dz[8 * c8 + 3],
                                     dz[8 * c8 + 4], dz[8 * c8 + 5], dz[8 * c8 + 6], dz[8 * c8 + 7]};
                const HeadSplit8 sp = head_split8(x8);
                uint16_t* d = p.dh_planes + head_tiled_elem((size_t)r, lane * KPL + 8 * c8, K);
                *reinterpret_cast<uint4*>(d) = sp.h;
                *reinterpret_cast<uint4*>(d + p.dh_plane_stride) = sp.m;
                *reinterpret_cast<uint4*>(d + 2 * p.dh_plane_stride) = sp.l;
            }
        }
    }
    // ---- per-warp partials: [dW (K * A) | db (A) | loss (1)] --------------------------------------------------------------
    float* part = p.workspace + (size_t)gw * ((size_t)K * A + A + 1);
#pragma unroll
    for (int j = 0; j < KPL; ++j)
#pragma unroll
        for (int a = 0; a < kHeadMaxA; ++a)
            if (a < A) part[((size_t)lane + 32 * j) * A + a] = acc_w[j][a];
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < kHeadMaxA; ++a)
            if (a < A) part[(size_t)K * A + a] = acc_b[a];
        part[(size_t)K * A + A] = acc_loss;
    }
}

// [nparts][n_out] partials -> dW | db | loss: a block owns 32 consecutive outputs, its 8 warps walk the partials 8
// apart (coalesced 128-byte reads) and the 8 sums are folded in a fixed order
__global__ void __launch_bounds__(256) dqn_head_reduce_kernel(const float* __restrict__ ws, int nparts, int n_out,
                                                              int KA, int A, float inv_b, float* __restrict__ dw,
                                                              float* __restrict__ db, float* __restrict__ loss) {
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (i < n_out)
        for (int q = w; q < nparts; q += 8) s += ws[(size_t)q * n_out + i];
    red[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < n_out) {
        for (int k = 1; k < 8; ++k) s += red[k][lane];
        if (i < KA) dw[i] = s;
        else if (i < KA + A) db[i - KA] = s;
        else if (loss) *loss = s * inv_b;
    }
}

__global__ void dueling_fwd_kernel(const float* __restrict__ v, const float* __restrict__ adv, int64_t B, int64_t A,
                                   float* __restrict__ q) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    float s = 0.f;
    for (int64_t a = 0; a < A; ++a) s += adv[i * A + a];
    const float mean = s / (float)A;
    for (int64_t a = 0; a < A; ++a) q[i * A + a] = v[i] + (adv[i * A + a] - mean);
}
__global__ void dueling_bwd_kernel(const float* __restrict__ dq, int64_t B, int64_t A, float* __restrict__ d_v,
                                   float* __restrict__ d_adv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    float s = 0.f;
    for (int64_t a = 0; a < A; ++a) s += dq[i * A + a];
    d_v[i] = s;
    const float mean = s / (float)A;
    for (int64_t a = 0; a < A; ++a) d_adv[i * A + a] = dq[i * A + a] - mean;
}

// ---- flat-buffer reductions / updates -------------------------------------------------------------------------------
constexpr int kRedBlocks = 1024;
__global__ void __launch_bounds__(256) sumsq_stage1(const float* __restrict__ x, int64_t n, float* __restrict__ part) {
    __shared__ float red[256];
    // contiguous slab per block, strided inside the block: fixed association order for a given n
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n, lo + per);
    float s = 0.f;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) s = fmaf(x[i], x[i], s);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(1024) sumsq_stage2(const float* __restrict__ part, int nparts,
                                                     float* __restrict__ out) {
    __shared__ float red[1024];
    red[threadIdx.x] = (threadIdx.x < nparts) ? part[threadIdx.x] : 0.f;
    __syncthreads();
    for (int k = 512; k > 0; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

__global__ void __launch_bounds__(256) clip_kernel(float* __restrict__ g, int64_t n, const float* __restrict__ sumsq,
                                                   float clip) {
    const float norm = sqrtf(*sumsq);
    const float scale = clip / fmaxf(norm, clip);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        g[i] *= scale;
}
__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ g, int64_t n, float s) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        g[i] *= s;
}

__global__ void __launch_bounds__(256) adam_tf_kernel(float* __restrict__ theta, float* __restrict__ m,
                                                      float* __restrict__ v, const float* __restrict__ g, int64_t n,
                                                      float alpha, float one_minus_b1, float one_minus_b2, float eps) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = __fadd_rn(m[i], __fmul_rn(__fsub_rn(gi, m[i]), one_minus_b1));
        const float vi = __fadd_rn(v[i], __fmul_rn(__fsub_rn(__fmul_rn(gi, gi), v[i]), one_minus_b2));
        m[i] = mi;
        v[i] = vi;
        theta[i] = __fsub_rn(theta[i], __fdiv_rn(__fmul_rn(mi, alpha), __fadd_rn(__fsqrt_rn(vi), eps)));
    }
}

// Adam with its step state on the device: state = {beta1_power, beta2_power}.  Launch-parameter-constant, so a whole
// training step can be captured in a CUDA graph and replayed; adam_state_advance_kernel multiplies the powers.
__global__ void __launch_bounds__(256) adam_tf_dev_kernel(float* __restrict__ theta, float* __restrict__ m,
                                                          float* __restrict__ v, const float* __restrict__ g,
                                                          int64_t n, float lr, float one_minus_b1, float one_minus_b2,
                                                          float eps, const float* __restrict__ state) {
    const float alpha = __fdiv_rn(__fmul_rn(lr, __fsqrt_rn(__fsub_rn(1.0f, state[1]))), __fsub_rn(1.0f, state[0]));
    // four parameters per thread through 128-bit accesses (the flat buffers are 32-byte aligned and padded to a multiple
    // of 8 elements: architectures/network.py ParamStore); same per-element operations, same bits
    const bool vec = (n % 4 == 0) && (((uintptr_t)theta | (uintptr_t)m | (uintptr_t)v | (uintptr_t)g) % 16 == 0);
    auto upd = [&](float& th, float& mm, float& vv, float gi) {
        mm = __fadd_rn(mm, __fmul_rn(__fsub_rn(gi, mm), one_minus_b1));
        vv = __fadd_rn(vv, __fmul_rn(__fsub_rn(__fmul_rn(gi, gi), vv), one_minus_b2));
        th = __fsub_rn(th, __fdiv_rn(__fmul_rn(mm, alpha), __fadd_rn(__fsqrt_rn(vv), eps)));
    };
    if (vec) {
        const int64_t n4 = n / 4;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
            const float4 g4 = reinterpret_cast<const float4*>(g)[i];
            float4 t4 = reinterpret_cast<float4*>(theta)[i], m4 = reinterpret_cast<float4*>(m)[i],
                   v4 = reinterpret_cast<float4*>(v)[i];
            upd(t4.x, m4.x, v4.x, g4.x);
            upd(t4.y, m4.y, v4.y, g4.y);
            upd(t4.z, m4.z, v4.z, g4.z);
            upd(t4.w, m4.w, v4.w, g4.w);
            reinterpret_cast<float4*>(m)[i] = m4;
            reinterpret_cast<float4*>(v)[i] = v4;
            reinterpret_cast<float4*>(theta)[i] = t4;
        }
        return;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float th = theta[i], mm = m[i], vv = v[i];
        upd(th, mm, vv, g[i]);
        m[i] = mm;
        v[i] = vv;
        theta[i] = th;
    }
}
__global__ void adam_state_advance_kernel(float* state, float beta1, float beta2) {
    state[0] = __fmul_rn(state[0], beta1);
    state[1] = __fmul_rn(state[1], beta2);
}
__global__ void add_i64_kernel(int64_t* x, int64_t delta) { *x += delta; }

__global__ void __launch_bounds__(256) polyak_kernel(float* __restrict__ target, const float* __restrict__ online,
                                                     int64_t n, float rate, float one_minus_rate) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        target[i] = __fadd_rn(__fmul_rn(rate, online[i]), __fmul_rn(one_minus_rate, target[i]));
}

static unsigned flat_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_dqn_td_targets(const float* q_next, const float* q_select, const float* q_online, const int64_t* actions,
                         const double* rewards, const uint8_t* game_overs, double discount, int64_t batch,
                         int64_t n_actions, float* targets_out, double* td_err_out, void* stream) {
    CB200_CHECK_ARG(q_next && q_select && q_online && actions && rewards && game_overs && targets_out && td_err_out,
                    "null pointer");
    CB200_CHECK_ARG(batch > 0 && n_actions > 0, "bad shape");
    CB200_LAUNCH(dqn_td_targets_kernel, (unsigned)((batch + 255) / 256), 256, 0, as_stream(stream), q_next, q_select,
                 q_online, actions, rewards, game_overs, discount, batch, n_actions, targets_out, td_err_out);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_regression_head_loss_grad(const float* out, const float* target, const float* weights, int64_t batch,
                                    int64_t width, int huber, float loss_weight, float* d_out, float* loss_out,
                                    void* stream) {
    CB200_CHECK_ARG(out && target && d_out && batch > 0 && width > 0, "bad arguments");
    int threads = 32;
    while (threads < batch && threads < 1024) threads *= 2;
    CB200_LAUNCH(regression_head_kernel, 1, threads, 0, as_stream(stream), out, target, weights, batch, width, huber,
                 loss_weight, d_out, loss_out);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_dqn_head_fused(const cb200_dqn_head_desc* d, void* stream) {
    CB200_CHECK_ARG(d != nullptr, "null descriptor");
    CB200_CHECK_ARG(d->h_next && d->h_online && d->w_target && d->b_target && d->w_online && d->b_online && d->actions &&
                        d->rewards && d->game_overs && d->q_online && d->targets && d->td_err && d->dq && d->dw && d->db &&
                        d->workspace,
                    "null pointer");
    CB200_CHECK_ARG(d->batch > 0 && d->n_actions > 0 && d->n_actions <= kHeadMaxA, "1 <= n_actions <= 8");
    CB200_CHECK_ARG(d->features == 256 || d->features == 512, "features must be 256 or 512");
    CB200_CHECK_ARG(!d->dh_planes || (d->dh_plane_stride % 8 == 0 && d->batch % 8 == 0), "planes: batch % 8, stride % 8");
    HeadParams p;
    p.h_next = d->h_next; p.h_online = d->h_online; p.h_select = d->h_select;
    p.w_target = d->w_target; p.b_target = d->b_target; p.w_online = d->w_online; p.b_online = d->b_online;
    p.actions = d->actions; p.rewards = d->rewards; p.game_overs = d->game_overs; p.weights = d->weights;
    p.discount = d->discount; p.huber = d->huber; p.B = (int)d->batch; p.K = d->features; p.A = d->n_actions;
    p.q_online = d->q_online; p.q_next = d->q_next; p.targets = d->targets; p.td_err = d->td_err;
    p.dq = d->dq; p.loss = d->loss; p.dh = d->dh;
    p.dh_planes = static_cast<uint16_t*>(d->dh_planes); p.dh_plane_stride = d->dh_plane_stride;
    p.dw = d->dw; p.db = d->db; p.workspace = d->workspace;
    const int warps = (p.B + kHeadRows - 1) / kHeadRows;
    const unsigned grid = (unsigned)((warps + kHeadWarps - 1) / kHeadWarps);
    const int nparts = (int)grid * kHeadWarps;                 // idle warps of the last block write zero partials
    cudaStream_t st = as_stream(stream);
    const size_t smem = (size_t)(2 * p.A * p.K + kHeadWarps * p.K) * sizeof(float);      // <= 48 KB (A <= 8, K <= 512)
    if (p.K == 512) {
        CB200_LAUNCH(dqn_head_fused_kernel<16>, grid, 32 * kHeadWarps, smem, st, p);
    } else {
        CB200_LAUNCH(dqn_head_fused_kernel<8>, grid, 32 * kHeadWarps, smem, st, p);
    }
    const int n_out = p.K * p.A + p.A + 1;
    CB200_LAUNCH(dqn_head_reduce_kernel, (unsigned)((n_out + 31) / 32), 256, 0, st, p.workspace, nparts, n_out,
                 p.K * p.A, p.A, 1.0f / (float)p.B, p.dw, p.db, p.loss);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_dueling_combine_fwd(const float* v, const float* adv, int64_t batch, int64_t n_actions, float* q,
                              void* stream) {
    CB200_CHECK_ARG(v && adv && q && batch > 0 && n_actions > 0, "bad arguments");
    CB200_LAUNCH(dueling_fwd_kernel, (unsigned)((batch + 255) / 256), 256, 0, as_stream(stream), v, adv, batch,
                 n_actions, q);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_dueling_combine_bwd(const float* dq, int64_t batch, int64_t n_actions, float* d_v, float* d_adv,
                              void* stream) {
    CB200_CHECK_ARG(dq && d_v && d_adv && batch > 0 && n_actions > 0, "bad arguments");
    CB200_LAUNCH(dueling_bwd_kernel, (unsigned)((batch + 255) / 256), 256, 0, as_stream(stream), dq, batch, n_actions,
                 d_v, d_adv);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_sumsq(const float* x, int64_t n, float* out, float* workspace, void* stream) {
    CB200_CHECK_ARG(x && out && workspace && n > 0, "bad arguments");
    int blocks = (int)((n + 4095) / 4096);
    if (blocks > kRedBlocks) blocks = kRedBlocks;
    CB200_LAUNCH(sumsq_stage1, blocks, 256, 0, as_stream(stream), x, n, workspace);
    CB200_LAUNCH(sumsq_stage2, 1, 1024, 0, as_stream(stream), workspace, blocks, out);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_clip_by_global_norm(float* g, int64_t n, const float* sumsq, float clip, void* stream) {
    CB200_CHECK_ARG(g && sumsq && n > 0 && clip > 0, "bad arguments");
    CB200_LAUNCH(clip_kernel, flat_grid(n), 256, 0, as_stream(stream), g, n, sumsq, clip);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_scale(float* g, int64_t n, float s, void* stream) {
    CB200_CHECK_ARG(g && n > 0, "bad arguments");
    CB200_LAUNCH(scale_kernel, flat_grid(n), 256, 0, as_stream(stream), g, n, s);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_adam_tf(float* theta, float* m, float* v, const float* g, int64_t n, float lr, float beta1, float beta2,
                  float epsilon, float beta1_power, float beta2_power, void* stream) {
    CB200_CHECK_ARG(theta && m && v && g && n > 0, "bad arguments");
    const float alpha = lr * sqrtf(1.0f - beta2_power) / (1.0f - beta1_power);
    CB200_LAUNCH(adam_tf_kernel, flat_grid(n), 256, 0, as_stream(stream), theta, m, v, g, n, alpha, 1.0f - beta1,
                 1.0f - beta2, epsilon);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_adam_tf_dev(float* theta, float* m, float* v, const float* g, int64_t n, float lr, float beta1, float beta2,
                      float epsilon, float* state, void* stream) {
    CB200_CHECK_ARG(theta && m && v && g && state && n > 0, "bad arguments");
    CB200_LAUNCH(adam_tf_dev_kernel, flat_grid(n), 256, 0, as_stream(stream), theta, m, v, g, n, lr, 1.0f - beta1,
                 1.0f - beta2, epsilon, state);
    CB200_LAUNCH(adam_state_advance_kernel, 1, 1, 0, as_stream(stream), state, beta1, beta2);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_add_i64(int64_t* x, int64_t delta, void* stream) {
    CB200_CHECK_ARG(x != nullptr, "null pointer");
    CB200_LAUNCH(add_i64_kernel, 1, 1, 0, as_stream(stream), x, delta);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_polyak(float* target, const float* online, int64_t n, double rate, void* stream) {
    CB200_CHECK_ARG(target && online && n > 0, "bad arguments");
    CB200_LAUNCH(polyak_kernel, flat_grid(n), 256, 0, as_stream(stream), target, online, n, (float)rate,
                 (float)(1.0 - rate));
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

}  // extern "C"
