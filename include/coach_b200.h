/*
 * include/coach_b200.h -- C ABI of libcoach_b200.so (hand-written sm_100a CUDA behind Coach's
 * replay-sample -> learn_from_batch hot path).
 *
 * The reference (IntelLabs/coach, rl-coach 1.0.1) is pure Python and has no FFI of its own; its plugin boundary is
 * class substitution through `Parameters.path` strings (rl_coach/memories/memory.py:36-38,
 * rl_coach/agents/dqn_agent.py:64-66).  The Python classes in `coach_b200/` mirror those reference classes and bind
 * to the entry points below with ctypes.  Every entry point:
 *   - takes plain device/host pointers, sizes and a CUDA stream handle (`void* stream` = cudaStream_t; NULL = the
 *     legacy default stream); no torch types appear anywhere in this ABI;
 *   - is asynchronous on `stream` unless stated otherwise and never synchronises the device;
 *   - returns CB200_OK (0) or a negative CB200_ERR_* code; `cb200_last_error()` returns a thread-local message.
 * Each declaration cites the reference code (file:line under /root/reference/rl_coach/) whose arithmetic it replaces.
 *
 * Pointers are DEVICE pointers unless the parameter name starts with `h_`.
 */
#ifndef COACH_B200_H
#define COACH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_ABI_VERSION 1

#define CB200_OK 0
#define CB200_ERR_INVALID_ARGUMENT (-1)
#define CB200_ERR_CUDA (-2)
#define CB200_ERR_UNSUPPORTED (-3)

int cb200_abi_version(void);
const char* cb200_last_error(void);
/* Number of kernels launched through this library by the calling process so far (bench.py's `gpu_launches`). */
int64_t cb200_launch_count(void);
/* Multiprocessor count and compute capability of the current device (any pointer may be NULL). */
int cb200_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Runtime tuning knobs (benchmark sweeps): "gather_stages" (smem stages per CTA, default 6),
 * "gather_ctas_per_sm" (persistent gather grid = SMs x this, default 2). Unknown keys are stored and ignored. */
int cb200_tune(const char* key, int value);

/* =====================================================================================================================
 * Segment trees (prioritized replay).  Layout = the reference's: one implicit binary heap per tree, float64,
 * 2*size-1 entries, root at [0], children of p at 2p+1 / 2p+2, leaves at [size-1, 2*size-1); `size` a power of two.
 * memories/non_episodic/prioritized_experience_replay.py:43-156 (SegmentTree).
 * ===================================================================================================================*/

/* SegmentTree.__init__ :54-61 -- sum tree <- 0, min tree <- +inf, max tree <- -inf. `winner` (int32[size]) is the
 * scratch array used by cb200_per_update for last-writer-wins duplicate resolution; it is set to -1. */
int cb200_per_init(double* sum_tree, double* min_tree, double* max_tree, int32_t* winner, int64_t size, void* stream);

/* PrioritizedExperienceReplay.update_priorities :203-217 + _update_priority :188-201 + SegmentTree.update/_propagate
 * :116-129,:63-74, for a whole batch at once.  Leaf idx[i] of the sum and min trees receives p_alpha[i], of the max
 * tree p_raw[i]; duplicates resolve last-writer-wins in batch order (the reference applies them sequentially); every
 * ancestor is then recomputed as op(left, right) level by level, which is bit-identical to the sequential reference
 * because parents are recomputed, never incrementally adjusted.  `max_priority_out` (device double, may be NULL)
 * receives max_tree[0] (= self.maximal_priority, :201).  n <= 1024 runs as one CTA; larger n uses one launch per
 * tree level. */
int cb200_per_update(double* sum_tree, double* min_tree, double* max_tree, int32_t* winner, int64_t size,
                     const int64_t* idx, const double* p_alpha, const double* p_raw, int64_t n,
                     double* max_priority_out, void* stream);

/* priority = error + epsilon; p_raw = priority; p_alpha = priority ** alpha  (:197-200) computed ON DEVICE with CUDA's
 * pow (<= 2 ulp; glibc's pow, which the reference uses, is not correctly rounded either, so device and reference
 * can differ in the last bit of ~0.1% of leaves -- see cb200_host_priorities for the libm-exact route).
 * *neg_flag (device int32, may be NULL) is set to 1 if any err[i] < 0 (reference raises ValueError :195). */
int cb200_per_priorities_device(const double* err, int64_t n, double epsilon, double alpha, double* p_alpha,
                                double* p_raw, int32_t* neg_flag, void* stream);

/* Same arithmetic with the host's libm `pow` -- bit-identical to the reference on the same machine.  Synchronous,
 * HOST pointers; the agent overlaps it with the network backward pass.  Returns CB200_ERR_INVALID_ARGUMENT (and
 * processes nothing) if an error value is negative. */
int cb200_host_priorities(const double* h_err, int64_t n, double epsilon, double alpha, double* h_p_alpha,
                          double* h_p_raw);

/* PrioritizedExperienceReplay.store :264-283 + SegmentTree.add :102-114 for n consecutive transitions starting at
 * ring cursor `cursor` (wraps at size): every new leaf gets p_alpha (sum, min) / p_raw (max) where the caller passes
 * p_raw = maximal_priority and p_alpha = maximal_priority ** alpha. */
int cb200_per_store(double* sum_tree, double* min_tree, double* max_tree, int32_t* winner, int64_t size,
                    int64_t cursor, int64_t n, double p_alpha, double p_raw, void* stream);

/* PrioritizedExperienceReplay.sample :229-253 + SegmentTree._retrieve :76-92 for `n` samples.
 *   u[i]        raw random.random() draws (the host keeps Python's MT19937 stream; random.uniform(a,b)=a+(b-a)*u)
 *   nt          num_transitions() -- the reference's doubled count (store appends twice, :271/:280)
 *   idx_out     int64[n]  leaf indices (bit-exact)
 *   w_out       double[n] normalised importance weights ((nt*P)^-beta / max_w); w32_out float[n] the same rounded
 *               once to fp32 (what the TF placeholder receives); either may be NULL
 * One warp per sample; each round fetches a 5-level sub-tree (62 nodes) with two coalesced loads per lane and walks
 * it with shuffles, so a 2^20-leaf descent costs 4 dependent memory round trips instead of 20. */
int cb200_per_sample(const double* sum_tree, const double* min_tree, int64_t size, const double* u, int64_t n,
                     int64_t nt, double beta, int64_t* idx_out, double* w_out, float* w32_out, void* stream);

/* =====================================================================================================================
 * Transition columns.  The replay ring is struct-of-arrays in HBM: one row-major [capacity, row_bytes] byte matrix
 * per transition field (state, next_state, action, reward, game_over, ...).  Replaces the AoS->SoA gather of
 * core_types.py:488-623 (Batch.states/next_states/actions/rewards/game_overs) and the list indexing of
 * experience_replay.py:90.
 * ===================================================================================================================*/
#define CB200_MAX_COLUMNS 8

typedef struct cb200_column {
    const void* src;     /* [capacity, row_bytes] ring column                                         */
    void* dst;           /* [n, row_bytes] staged minibatch column                                    */
    int64_t row_bytes;   /* bytes per transition in this column                                       */
} cb200_column;

/* dst[c][i, :] = src[c][idx[i], :] for every column c.  Rows that are 16-byte multiples and >= 2 KiB move through
 * shared memory with 1-D bulk async copies (TMA: cp.async.bulk global->shared->global, mbarrier completion) from a
 * persistent grid; smaller / unaligned rows use vectorised LSU copies. */
int cb200_gather(const cb200_column* h_columns, int n_columns, const int64_t* idx, int64_t n, void* stream);

/* Fused PER sample + gather (one launch): cb200_per_sample followed by cb200_gather on the freshly drawn indices
 * without leaving the kernel. */
int cb200_per_sample_gather(const double* sum_tree, const double* min_tree, int64_t size, const double* u, int64_t n,
                            int64_t nt, double beta, int64_t* idx_out, double* w_out, float* w32_out,
                            const cb200_column* h_columns, int n_columns, void* stream);

/* Ring append: src[c][(cursor + i) % capacity, :] = staged[c][i, :] (experience_replay.py:131-150 store; the
 * `cb200_column.src` member is the ring (written), `.dst` the staged rows (read)). */
int cb200_scatter_ring(const cb200_column* h_columns, int n_columns, int64_t cursor, int64_t capacity, int64_t n,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COACH_B200_H */
