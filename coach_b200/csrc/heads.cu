// coach_b200/csrc/heads.cu -- policy / value head losses of the actor-critic agents (ClippedPPO, DDPG/TD3, SAC) and
// the minibatch row gather used by the epoch loops.  Reference lines are cited in include/coach_b200.h.
#include <math.h>

#include "common.cuh"

namespace cb200 {

constexpr int kMaxActionDim = 32;
constexpr float kLog2Pi = 1.8378770664093453f;
constexpr float kTfEps = 1e-15f;          // `eps` of heads/ppo_head.py (std + eps)

// fixed-order block reduction of one float per thread (blockDim.x a power of two <= 1024)
__device__ __forceinline__ float block_sum(float v, float* red) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    const float out = red[0];
    __syncthreads();
    return out;
}

// =====================================================================================================================
// PPOHead, continuous actions (heads/ppo_head.py:52-144): diagonal Gaussian with state-independent log-std.
//   sigma_j   = exp(logstd_j) + eps
//   logp_i    = -0.5 * sum_j ((a_ij - mu_ij)/sigma_j)^2 - sum_j log sigma_j - 0.5*k*log(2 pi)
//   ratio_i   = exp(logp_i - logp_old_i);  clipped_i = clip(ratio_i, 1 - e, 1 + e),  e = clip_eps * rescaler
//   L         = -mean_i min(ratio_i * A_i, clipped_i * A_i)  -  beta * mean entropy
// Outputs the gradients wrt mu and logstd and the scalars the reference logs (loss, KL(old||new), entropy, mean
// ratio, mean clipped ratio).  One block; fixed reduction order.
// =====================================================================================================================
__global__ void __launch_bounds__(256) ppo_continuous_head_kernel(
    const float* __restrict__ mu, const float* __restrict__ logstd, const float* __restrict__ actions,
    const float* __restrict__ old_mu, const float* __restrict__ old_logstd, const float* __restrict__ advantages,
    int64_t B, int A, float clip_eps, float beta_entropy, float* __restrict__ d_mu, float* __restrict__ d_logstd,
    float* __restrict__ scalars /* [loss, kl, entropy, mean ratio, mean clipped ratio] */) {
    __shared__ float red[256];
    __shared__ float sig[kMaxActionDim], osig[kMaxActionDim], dls_part[kMaxActionDim];
    if (threadIdx.x < A) {
        sig[threadIdx.x] = expf(logstd[threadIdx.x]) + kTfEps;
        osig[threadIdx.x] = expf(old_logstd[threadIdx.x]) + kTfEps;
        dls_part[threadIdx.x] = 0.f;
    }
    __syncthreads();
    float sum_log_sig = 0.f, sum_log_osig = 0.f;
    for (int j = 0; j < A; ++j) {
        sum_log_sig += logf(sig[j]);
        sum_log_osig += logf(osig[j]);
    }
    const float inv_b = 1.0f / (float)B;
    const float lo = 1.0f - clip_eps, hi = 1.0f + clip_eps;
    float loss_acc = 0.f, kl_acc = 0.f, ratio_acc = 0.f, cratio_acc = 0.f;
    float dls_local[kMaxActionDim];
#pragma unroll
    for (int j = 0; j < kMaxActionDim; ++j) dls_local[j] = 0.f;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
        float q = 0.f, qo = 0.f, kl = 0.f;
        for (int j = 0; j < A; ++j) {
            const float a = actions[i * A + j];
            const float z = (a - mu[i * A + j]) / sig[j];
            const float zo = (a - old_mu[i * A + j]) / osig[j];
            q += z * z;
            qo += zo * zo;
            // KL(old || new) of diagonal Gaussians
            const float dm = (mu[i * A + j] - old_mu[i * A + j]) / sig[j];
            const float rs = osig[j] / sig[j];
            kl += 0.5f * (rs * rs + dm * dm - 1.0f) - logf(rs);
        }
        const float logp = -0.5f * q - sum_log_sig - 0.5f * A * kLog2Pi;
        const float logp_old = -0.5f * qo - sum_log_osig - 0.5f * A * kLog2Pi;
        const float ratio = expf(logp - logp_old);
        const float cl = fminf(fmaxf(ratio, lo), hi);
        const float adv = advantages[i];
        const float s1 = ratio * adv, s2 = cl * adv;
        // tf.minimum passes the gradient to its first argument when s1 <= s2; clip_by_value passes it inside [lo, hi]
        float ds_dratio;
        if (s1 <= s2) ds_dratio = adv;
        else ds_dratio = (ratio >= lo && ratio <= hi) ? adv : 0.f;
        loss_acc += fminf(s1, s2);
        kl_acc += kl;
        ratio_acc += ratio;
        cratio_acc += cl;
        const float dlogp = -inv_b * ds_dratio * ratio;       // dL/dlogp_i
        for (int j = 0; j < A; ++j) {
            const float z = (actions[i * A + j] - mu[i * A + j]) / sig[j];
            d_mu[i * A + j] = dlogp * z / sig[j];
            // d logp / d logstd_j = (z^2 - 1) * exp(logstd_j) / sigma_j
            dls_local[j] += dlogp * (z * z - 1.0f) * ((sig[j] - kTfEps) / sig[j]);
        }
    }
    const float loss_sum = block_sum(loss_acc, red);
    const float kl_sum = block_sum(kl_acc, red);
    const float ratio_sum = block_sum(ratio_acc, red);
    const float cratio_sum = block_sum(cratio_acc, red);
    for (int j = 0; j < A; ++j) {
        const float s = block_sum(dls_local[j], red);
        if (threadIdx.x == 0) dls_part[j] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float entropy = 0.5f * A * (1.0f + kLog2Pi) + sum_log_sig;     // state independent
        for (int j = 0; j < A; ++j) {
            // entropy regulariser -beta * H:  dH/dlogstd_j = exp(logstd_j) / sigma_j
            d_logstd[j] = dls_part[j] - beta_entropy * ((sig[j] - kTfEps) / sig[j]);
        }
        if (scalars) {
            scalars[0] = -loss_sum * inv_b - beta_entropy * entropy;
            scalars[1] = kl_sum * inv_b;
            scalars[2] = entropy;
            scalars[3] = ratio_sum * inv_b;
            scalars[4] = cratio_sum * inv_b;
        }
    }
}

// =====================================================================================================================
// dst[c][i, :] = src[c][idx[*offset + i], :] -- minibatch extraction inside a (CUDA-graph captured) epoch loop: the
// launch parameters stay constant, only the device scalar *offset changes between replays.
// =====================================================================================================================
struct AtColumns {
    const uint8_t* src[CB200_MAX_COLUMNS];
    uint8_t* dst[CB200_MAX_COLUMNS];
    int64_t row_bytes[CB200_MAX_COLUMNS];
    int n;
};
__global__ void __launch_bounds__(256) gather_at_kernel(AtColumns cols, const int64_t* __restrict__ idx,
                                                        const int64_t* __restrict__ offset_ptr, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t off = offset_ptr ? *offset_ptr : 0;
    for (int64_t w = warp; w < n * cols.n; w += nwarps) {
        const int64_t i = w / cols.n;
        const int c = (int)(w - i * cols.n);
        const int64_t row = idx ? idx[off + i] : (off + i);
        const uint8_t* s = cols.src[c] + row * cols.row_bytes[c];
        uint8_t* d = cols.dst[c] + i * cols.row_bytes[c];
        const int64_t bytes = cols.row_bytes[c];
        const uintptr_t al = reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d) | (uintptr_t)bytes;
        if ((al & 3) == 0) {
            for (int64_t o = (int64_t)lane * 4; o < bytes; o += 128)
                *reinterpret_cast<uint32_t*>(d + o) = *reinterpret_cast<const uint32_t*>(s + o);
        } else {
            for (int64_t o = lane; o < bytes; o += 32) d[o] = s[o];
        }
    }
}


// dz[r, c] = dy[r, c] * act'(y[r, c])  with independent leading dimensions (activation backward on a column block of a
// wider buffer, e.g. the embedder part of a concatenated critic input)
__global__ void __launch_bounds__(256) act_backward_kernel(const float* __restrict__ dy, int ld_dy,
                                                           const float* __restrict__ y, int ld_y, int64_t rows,
                                                           int cols, int act, float* __restrict__ dz, int ld_dz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    const float yy = y[r * ld_y + c];
    float g = 1.f;
    if (act == CB200_ACT_RELU) g = yy > 0.f ? 1.f : 0.f;
    else if (act == CB200_ACT_TANH) g = 1.f - yy * yy;
    dz[r * ld_dz + c] = dy[r * ld_dy + c] * g;
}

// dst[r, c] = alpha * src[r, c] + beta * dst[r, c]   (strided 2-D; beta == 0 never reads dst)
__global__ void __launch_bounds__(256) axpby_2d_kernel(const float* __restrict__ src, int ld_src, int64_t rows, int cols,
                                                       float alpha, float beta, float* __restrict__ dst, int ld_dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    const float v = alpha * src[r * ld_src + c];
    dst[r * ld_dst + c] = (beta == 0.f) ? v : (v + beta * dst[r * ld_dst + c]);
}

// DDPG / TD3 / SAC bootstrapped targets, evaluated like the numpy expression (fp64 rewards, fp32 network output):
//   y = r + (1 - done) * discount * q_next      [done ignored when use_non_zero_discount_for_terminal_states]
//   optional clip (ddpg_agent.py:163-164); written as fp32 (the TF placeholder dtype)
__global__ void __launch_bounds__(256) ac_td_targets_kernel(const double* __restrict__ rewards,
                                                            const uint8_t* __restrict__ dones,
                                                            const float* __restrict__ q_next, int ld_q, int64_t B,
                                                            double discount, int ignore_done, int use_clip,
                                                            double clip_lo, double clip_hi, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double y;
    if (ignore_done) {
        // ddpg_agent.py:155: rewards + discount * q -- a Python float times a float32 array stays float32 in numpy, so
        // this product is rounded to fp32 before the fp64 add (pinned by tests/golden/agent_prologues.npz, "ddpg2")
        y = __dadd_rn(rewards[i], (double)__fmul_rn((float)discount, q_next[i * ld_q]));
    } else {
        // :157-158: (1.0 - game_overs) is a float64 array, the whole product is fp64
        const double nd = __dsub_rn(1.0, dones[i] ? 1.0 : 0.0);
        y = __dadd_rn(rewards[i], __dmul_rn(__dmul_rn(nd, discount), (double)q_next[i * ld_q]));
    }
    if (use_clip) y = fmin(fmax(y, clip_lo), clip_hi);
    out[i] = (float)y;
}

// out[i] = min(a[i], b[i])  (TD3 / SAC clipped double-Q)
__global__ void __launch_bounds__(256) min2_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                   float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fminf(a[i], b[i]);
}

// TD3 target policy smoothing (td3_agent.py:162-164): a = clip(a + clip(noise, -c, c), lo, hi), in place.  The
// reference adds the fp64 numpy draw to the fp32 network output in fp64, clips in fp64 and the result is rounded to
// fp32 once when it is fed to the critic: same operations here (pinned by tests/golden/agent_prologues.npz).
__global__ void __launch_bounds__(256) td3_smooth_kernel(float* __restrict__ actions, const double* __restrict__ noise,
                                                         int64_t n, double noise_clip, double lo, double hi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double nz = fmin(fmax(noise[i], -noise_clip), noise_clip);
    actions[i] = (float)fmin(fmax(__dadd_rn((double)actions[i], nz), lo), hi);
}

// =====================================================================================================================
// CategoricalQHead + the distributional TD target of CategoricalDQNAgent / RainbowDQNAgent
// (agents/categorical_dqn_agent.py:105-165, rainbow_dqn_agent.py:93-140, heads/categorical_q_head.py:41-57).
// One warp per sample:
//   p_next[a, :]  = softmax(next logits[a, :])  (fp32)      Q[a] = sum_j (double)p[a, j] * z[j]   (np.dot with fp64 z)
//   a*            = argmax_a Q[a] of the target prediction  (of the `select` prediction for Rainbow's double-Q rule)
//   m[:]          = the projection of  r + boot * gamma_n * z_j  onto the support, accumulated in fp64 in the
//                   reference's order (j ascending; first the floor bin, then the ceil bin; an integral b_j adds
//                   nothing to either bin -- the reference's arithmetic, kept)
//   labels[a, :]  = online softmax for a != action (TD_targets starts as the online prediction), (float)m for the action
//   loss[a]       = sum_j labels_j * (log sum_k exp(x_k - max) - (x_j - max))         (tf.nn.softmax_cross_entropy)
//   dlogits[a, :] = softmax - labels for the taken action, exactly 0 elsewhere (labels ARE the softmax there)
// next_is_prob: the "next" / "select" inputs already hold probabilities (parity tests pin the projection bit for bit).
// =====================================================================================================================
struct C51Params {
    const float* next; const float* online; const float* select;
    const int64_t* actions; const double* rewards; const uint8_t* game_overs; const double* bootstrap;
    const double* z;
    double gamma_n;
    int B, A, N, next_is_prob;
    float* labels; float* dlogits; float* loss_rows; double* td_err; double* q_online; int64_t* target_actions;
};

constexpr int kC51Warps = 4;

__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// softmax of one row of N logits into dst (shared); returns (max, sum of exp) to every lane
__device__ __forceinline__ void warp_softmax(const float* __restrict__ x, int N, int lane, float* dst, float& mx,
                                             float& sum) {
    float m = -INFINITY;
    for (int j = lane; j < N; j += 32) m = fmaxf(m, x[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < N; j += 32) {
        const float e = expf(x[j] - m);
        dst[j] = e;
        s += e;
    }
    s = warp_sum(s);
    for (int j = lane; j < N; j += 32) dst[j] = dst[j] / s;
    mx = m;
    sum = s;
    __syncwarp();
}

__global__ void __launch_bounds__(kC51Warps * 32) c51_head_kernel(C51Params p) {
    extern __shared__ double c51_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kC51Warps + warp;
    if (b >= p.B) return;
    const int A = p.A, N = p.N;
    double* s_m = c51_smem + (size_t)warp * N;                                         // [N] fp64 projection
    float* s_p = reinterpret_cast<float*>(c51_smem + (size_t)kC51Warps * N) + (size_t)warp * 2 * N;   // [2][N]
    float* s_sel = s_p + N;
    const int64_t row = (int64_t)b * A * N;

    // ---- target action ------------------------------------------------------------------------------------------
    int best = 0;
    double best_q = 0.0;
    const float* sel_src = p.select ? p.select : p.next;
    for (int a = 0; a < A; ++a) {
        float mx, sum;
        if (p.next_is_prob) {
            for (int j = lane; j < N; j += 32) s_sel[j] = sel_src[row + (int64_t)a * N + j];
            __syncwarp();
        } else {
            warp_softmax(sel_src + row + (int64_t)a * N, N, lane, s_sel, mx, sum);
        }
        double q = 0.0;
        for (int j = lane; j < N; j += 32) q += (double)s_sel[j] * p.z[j];
        q = warp_sum(q);
        if (a == 0 || q > best_q) { best_q = q; best = a; }                            // np.argmax: first maximum
        __syncwarp();
    }
    if (p.target_actions && lane == 0) p.target_actions[b] = best;
    {   // distribution of the target network for the chosen action
        float mx, sum;
        if (p.next_is_prob) {
            for (int j = lane; j < N; j += 32) s_p[j] = p.next[row + (int64_t)best * N + j];
            __syncwarp();
        } else {
            warp_softmax(p.next + row + (int64_t)best * N, N, lane, s_p, mx, sum);
        }
    }
    // ---- projection, sequential in j like the reference loop ---------------------------------------------------------
    for (int j = lane; j < N; j += 32) s_m[j] = 0.0;
    __syncwarp();
    if (lane == 0) {
        const double boot = p.bootstrap ? p.bootstrap[b] : __dsub_rn(1.0, p.game_overs[b] ? 1.0 : 0.0);
        const double coef = __dmul_rn(boot, p.gamma_n);
        const double r = p.rewards[b], z0 = p.z[0], zl = p.z[N - 1];
        const double dz = __dsub_rn(p.z[1], z0);
        for (int j = 0; j < N; ++j) {
            const double tz = fmax(fmin(__dadd_rn(r, __dmul_rn(coef, p.z[j])), zl), z0);
            const double bj = __ddiv_rn(__dsub_rn(tz, z0), dz);
            const double u = ceil(bj), l = floor(bj);
            const double pj = (double)s_p[j];
            s_m[(int)l] = __dadd_rn(s_m[(int)l], __dmul_rn(pj, __dsub_rn(u, bj)));
            s_m[(int)u] = __dadd_rn(s_m[(int)u], __dmul_rn(pj, __dsub_rn(bj, l)));
        }
    }
    __syncwarp();
    // ---- online head: labels, cross entropy, gradient ---------------------------------------------------------------
    const int act = (int)p.actions[b];
    for (int a = 0; a < A; ++a) {
        const float* x = p.online + row + (int64_t)a * N;
        float mx, sum;
        warp_softmax(x, N, lane, s_sel, mx, sum);
        const float lse = logf(sum);
        float loss = 0.f;
        double q = 0.0;
        for (int j = lane; j < N; j += 32) {
            const float sm = s_sel[j];
            const float lab = (a == act) ? (float)s_m[j] : sm;
            loss += lab * (lse - (x[j] - mx));
            p.labels[row + (int64_t)a * N + j] = lab;
            p.dlogits[row + (int64_t)a * N + j] = (a == act) ? (sm - lab) : 0.f;
            q += (double)sm * p.z[j];
        }
        loss = warp_sum(loss);
        q = warp_sum(q);
        if (lane == 0) {
            p.loss_rows[(int64_t)b * A + a] = loss;
            if (p.q_online) p.q_online[(int64_t)b * A + a] = q;
            if (a == act) p.td_err[b] = (double)loss;
        }
        __syncwarp();
    }
}

// q_values of the CategoricalQHead (categorical_q_head.py:56): tensordot(cast(softmax, fp64), z), one warp per (b, a) row
__global__ void __launch_bounds__(128) c51_q_values_kernel(const float* __restrict__ logits, const double* __restrict__ z,
                                                           int64_t rows, int N, double* __restrict__ q) {
    const int64_t r = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* x = logits + r * N;
    float m = -INFINITY;
    for (int j = lane; j < N; j += 32) m = fmaxf(m, x[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < N; j += 32) s += expf(x[j] - m);
    s = warp_sum(s);
    double acc = 0.0;
    for (int j = lane; j < N; j += 32) acc += (double)(expf(x[j] - m) / s) * z[j];
    acc = warp_sum(acc);
    if (lane == 0) q[r] = acc;
}

// total loss = tf.reduce_sum over the [B, A] loss tensor (general_network.py:360), one block, fixed order
__global__ void __launch_bounds__(256) c51_loss_sum_kernel(const float* __restrict__ loss_rows, int64_t n,
                                                           float* __restrict__ total) {
    __shared__ float red[256];
    float v = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += 256) v += loss_rows[i];
    const float t = block_sum(v, red);
    if (threadIdx.x == 0) *total = t;
}

// =====================================================================================================================
// SACPolicyHead (heads/sac_head.py:60-97): head output z = [mu | log_sigma_raw]  (2A columns),
//   log_sigma = clip(log_sigma_raw, -20, 2);  u = mu + exp(log_sigma) * eps;  a = tanh(u);
//   log pi(a|s) = sum_j [ -0.5 eps_j^2 - log_sigma_j - 0.5 log(2 pi) ] - sum_j log(1 - tanh(u_j)^2 + 1e-6)
// (MultivariateNormalDiag.log_prob of the reparameterised sample, minus the squash correction :49-58).
// =====================================================================================================================
constexpr float kSacLogSigMin = -20.f, kSacLogSigMax = 2.f, kSacEps = 1e-6f;

__global__ void __launch_bounds__(256) sac_policy_sample_kernel(const float* __restrict__ z, const float* __restrict__ eps,
                                                                int64_t B, int A, float* __restrict__ raw,
                                                                float* __restrict__ act, float* __restrict__ logp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    float lp = 0.f;
    for (int j = 0; j < A; ++j) {
        const float mu = z[i * 2 * A + j];
        const float ls = fminf(fmaxf(z[i * 2 * A + A + j], kSacLogSigMin), kSacLogSigMax);
        const float e = eps[i * A + j];
        const float u = mu + expf(ls) * e;
        const float t = tanhf(u);
        if (raw) raw[i * A + j] = u;
        if (act) act[i * A + j] = t;
        lp += -0.5f * e * e - ls - 0.5f * kLog2Pi - logf(1.0f - t * t + kSacEps);
    }
    if (logp) logp[i] = lp;
}

// gradient of  mean_b log pi(a~|s)  (noise eps_lp)  minus  sum_b <dq_da_b, a~_b>  (noise eps_q)  wrt the head output z:
// the two terms of policy_grads = dlogp_dphi - dq_dphi (soft_actor_critic_agent.py:213-232); each term was evaluated
// by its own sess.run and therefore with its own noise sample.
__global__ void __launch_bounds__(256) sac_policy_grad_kernel(const float* __restrict__ z, const float* __restrict__ eps_lp,
                                                              const float* __restrict__ eps_q,
                                                              const float* __restrict__ dq_da, int64_t B, int A,
                                                              float* __restrict__ dz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float inv_b = 1.0f / (float)B;
    for (int j = 0; j < A; ++j) {
        const float mu = z[i * 2 * A + j];
        const float lsr = z[i * 2 * A + A + j];
        const float ls = fminf(fmaxf(lsr, kSacLogSigMin), kSacLogSigMax);
        const float in_range = (lsr >= kSacLogSigMin && lsr <= kSacLogSigMax) ? 1.f : 0.f;   // clip_by_value gradient
        const float sig = expf(ls);
        // --- d mean(log pi) ---
        const float e2 = eps_lp[i * A + j];
        const float u2 = mu + sig * e2, t2 = tanhf(u2);
        const float gprime = (-2.0f * t2 * (1.0f - t2 * t2)) / (1.0f - t2 * t2 + kSacEps);   // d/du log(1 - t^2 + eps)
        float d_mu = inv_b * (-gprime);
        float d_ls = inv_b * (-1.0f - gprime * sig * e2) * in_range;
        // --- minus sum <dq_da, tanh(mu + sig * eps3)> ---
        const float e3 = eps_q[i * A + j];
        const float u3 = mu + sig * e3, t3 = tanhf(u3);
        const float w = dq_da[i * A + j] * (1.0f - t3 * t3);
        d_mu -= w;
        d_ls -= w * sig * e3 * in_range;
        dz[i * 2 * A + j] = d_mu;
        dz[i * 2 * A + A + j] = d_ls;
    }
}

// seeds of d mean_b(min(q1, q2)) / dq_k (sac_q_head.py:84-86; tf.minimum sends the gradient to its first argument on
// ties), optionally also out = min(q1, q2)
__global__ void __launch_bounds__(256) sac_min_seed_kernel(const float* __restrict__ q1, const float* __restrict__ q2,
                                                           int64_t B, float* __restrict__ d1, float* __restrict__ d2,
                                                           float* __restrict__ qmin) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const bool first = q1[i] <= q2[i];
    const float s = 1.0f / (float)B;
    if (d1) d1[i] = first ? s : 0.f;
    if (d2) d2[i] = first ? 0.f : s;
    if (qmin) qmin[i] = first ? q1[i] : q2[i];
}

// out[i] = a[i] - b[i]   (value targets = min Q - log pi, soft_actor_critic_agent.py:244)
__global__ void __launch_bounds__(256) sub_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                  float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}

// out[i] = (float) in[i]   (fp64 advantages / value targets -> fp32 network feeds)
__global__ void __launch_bounds__(256) f64_to_f32_kernel(const double* __restrict__ in, int64_t n,
                                                         float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_ppo_continuous_head(const float* mu, const float* logstd, const float* actions, const float* old_mu,
                              const float* old_logstd, const float* advantages, int64_t batch, int32_t action_dim,
                              float clip_eps, float beta_entropy, float* d_mu, float* d_logstd, float* scalars,
                              void* stream) {
    CB200_CHECK_ARG(mu && logstd && actions && old_mu && old_logstd && advantages && d_mu && d_logstd, "null pointer");
    CB200_CHECK_ARG(batch > 0 && action_dim > 0 && action_dim <= kMaxActionDim, "bad shape (action_dim <= 32)");
    CB200_LAUNCH(ppo_continuous_head_kernel, 1, 256, 0, as_stream(stream), mu, logstd, actions, old_mu, old_logstd,
                 advantages, batch, action_dim, clip_eps, beta_entropy, d_mu, d_logstd, scalars);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_gather_at(const cb200_column* h_columns, int n_columns, const int64_t* idx, const int64_t* offset,
                    int64_t n, void* stream) {
    CB200_CHECK_ARG(h_columns && n_columns > 0 && n_columns <= CB200_MAX_COLUMNS && n > 0, "bad arguments");
    AtColumns cols;
    cols.n = n_columns;
    for (int c = 0; c < n_columns; ++c) {
        CB200_CHECK_ARG(h_columns[c].src && h_columns[c].dst && h_columns[c].row_bytes > 0, "bad column entry");
        cols.src[c] = static_cast<const uint8_t*>(h_columns[c].src);
        cols.dst[c] = static_cast<uint8_t*>(h_columns[c].dst);
        cols.row_bytes[c] = h_columns[c].row_bytes;
    }
    int64_t grid = (n * n_columns + 7) / 8;
    if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
    CB200_LAUNCH(gather_at_kernel, (unsigned)grid, 256, 0, as_stream(stream), cols, idx, offset, n);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}


int cb200_act_backward(const float* dy, int32_t ld_dy, const float* y, int32_t ld_y, int64_t rows, int32_t cols,
                       int32_t act, float* dz, int32_t ld_dz, void* stream) {
    CB200_CHECK_ARG(dy && y && dz && rows > 0 && cols > 0 && ld_dy >= cols && ld_y >= cols && ld_dz >= cols,
                    "bad arguments");
    const int64_t n = rows * cols;
    CB200_LAUNCH(act_backward_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(stream), dy, ld_dy, y, ld_y, rows,
                 cols, act, dz, ld_dz);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_axpby_2d(const float* src, int32_t ld_src, int64_t rows, int32_t cols, float alpha, float beta, float* dst,
                   int32_t ld_dst, void* stream) {
    CB200_CHECK_ARG(src && dst && rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= cols, "bad arguments");
    const int64_t n = rows * cols;
    CB200_LAUNCH(axpby_2d_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(stream), src, ld_src, rows, cols,
                 alpha, beta, dst, ld_dst);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_ac_td_targets(const double* rewards, const uint8_t* game_overs, const float* q_next, int32_t ld_q,
                        int64_t batch, double discount, int32_t use_non_zero_discount_for_terminal_states,
                        int32_t use_clip, double clip_lo, double clip_hi, float* targets_out, void* stream) {
    CB200_CHECK_ARG(rewards && game_overs && q_next && targets_out && batch > 0 && ld_q >= 1, "bad arguments");
    CB200_LAUNCH(ac_td_targets_kernel, (unsigned)((batch + 255) / 256), 256, 0, as_stream(stream), rewards, game_overs,
                 q_next, ld_q, batch, discount, use_non_zero_discount_for_terminal_states, use_clip, clip_lo, clip_hi,
                 targets_out);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_min2(const float* a, const float* b, int64_t n, float* out, void* stream) {
    CB200_CHECK_ARG(a && b && out && n > 0, "bad arguments");
    CB200_LAUNCH(min2_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(stream), a, b, n, out);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_td3_smooth_actions(float* actions, const double* noise, int64_t n, double noise_clip, double lo, double hi,
                             void* stream) {
    CB200_CHECK_ARG(actions && noise && n > 0, "bad arguments");
    CB200_LAUNCH(td3_smooth_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(stream), actions, noise, n,
                 noise_clip, lo, hi);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}


int cb200_c51_head(const float* next, const float* online, const float* select, const int64_t* actions,
                   const double* rewards, const uint8_t* game_overs, const double* bootstrap, const double* z,
                   double gamma_n, int32_t batch, int32_t n_actions, int32_t n_atoms, int32_t next_is_prob,
                   float* labels, float* dlogits, float* loss_rows, float* total_loss, double* td_err,
                   double* q_online, int64_t* target_actions, void* stream) {
    CB200_CHECK_ARG(next && online && actions && rewards && (game_overs || bootstrap) && z, "bad arguments");
    CB200_CHECK_ARG(labels && dlogits && loss_rows && total_loss && td_err, "bad output arguments");
    CB200_CHECK_ARG(batch > 0 && n_actions > 0 && n_atoms >= 2 && n_atoms <= 1024, "bad shape");
    C51Params p;
    p.next = next; p.online = online; p.select = select; p.actions = actions; p.rewards = rewards;
    p.game_overs = game_overs; p.bootstrap = bootstrap; p.z = z; p.gamma_n = gamma_n;
    p.B = batch; p.A = n_actions; p.N = n_atoms; p.next_is_prob = next_is_prob;
    p.labels = labels; p.dlogits = dlogits; p.loss_rows = loss_rows; p.td_err = td_err; p.q_online = q_online;
    p.target_actions = target_actions;
    const size_t smem = (size_t)kC51Warps * n_atoms * (sizeof(double) + 2 * sizeof(float));
    CB200_LAUNCH(c51_head_kernel, (unsigned)((batch + kC51Warps - 1) / kC51Warps), kC51Warps * 32, smem,
                 as_stream(stream), p);
    CB200_CHECK_LAUNCH();
    CB200_LAUNCH(c51_loss_sum_kernel, 1, 256, 0, as_stream(stream), loss_rows, (int64_t)batch * n_actions, total_loss);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}


int cb200_c51_q_values(const float* logits, const double* z, int64_t rows, int32_t n_atoms, double* q_out, void* stream) {
    CB200_CHECK_ARG(logits && z && q_out && rows > 0 && n_atoms >= 2, "bad arguments");
    CB200_LAUNCH(c51_q_values_kernel, (unsigned)((rows + 3) / 4), 128, 0, as_stream(stream), logits, z, rows, n_atoms,
                 q_out);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_sac_policy_sample(const float* head_out, const float* eps, int64_t batch, int32_t action_dim, float* raw_out,
                            float* actions_out, float* logp_out, void* stream) {
    CB200_CHECK_ARG(head_out && eps && batch > 0 && action_dim > 0, "bad arguments");
    CB200_LAUNCH(sac_policy_sample_kernel, (unsigned)((batch + 255) / 256), 256, 0, as_stream(stream), head_out, eps,
                 batch, action_dim, raw_out, actions_out, logp_out);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_sac_policy_grad(const float* head_out, const float* eps_logp, const float* eps_q, const float* dq_da,
                          int64_t batch, int32_t action_dim, float* d_head_out, void* stream) {
    CB200_CHECK_ARG(head_out && eps_logp && eps_q && dq_da && d_head_out && batch > 0 && action_dim > 0,
                    "bad arguments");
    CB200_LAUNCH(sac_policy_grad_kernel, (unsigned)((batch + 255) / 256), 256, 0, as_stream(stream), head_out, eps_logp,
                 eps_q, dq_da, batch, action_dim, d_head_out);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_sac_min_seed(const float* q1, const float* q2, int64_t batch, float* d1, float* d2, float* qmin,
                       void* stream) {
    CB200_CHECK_ARG(q1 && q2 && batch > 0, "bad arguments");
    CB200_LAUNCH(sac_min_seed_kernel, (unsigned)((batch + 255) / 256), 256, 0, as_stream(stream), q1, q2, batch, d1, d2,
                 qmin);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_sub(const float* a, const float* b, int64_t n, float* out, void* stream) {
    CB200_CHECK_ARG(a && b && out && n > 0, "bad arguments");
    CB200_LAUNCH(sub_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(stream), a, b, n, out);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

int cb200_f64_to_f32(const double* in, int64_t n, float* out, void* stream) {
    CB200_CHECK_ARG(in && out && n > 0, "bad arguments");
    CB200_LAUNCH(f64_to_f32_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(stream), in, n, out);
    CB200_CHECK_LAUNCH();
    return CB200_OK;
}

}  // extern "C"
