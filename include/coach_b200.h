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
/* Runtime switches for benchmark A/B runs (unknown keys are stored and ignored):
 *   "gather_ctas_per_sm" (persistent gather grid = SMs x this, default 4), "gather_stages" (0 = automatic)
 *   "gemm_tc" (1)          tcgen05 path of cb200_gemm; 0 = fp32 CUDA-core kernels only
 *   "gemm_skinny" (1)      dedicated kernels for products with n <= 8 or k <= 8
 *   "gemm_persistent" (0)  persistent schedule of cb200_gemm_tiled (two TMEM accumulator sets, 8 epilogue warps) */
int cb200_tune(const char* key, int value);

/* =====================================================================================================================
 * Segment trees (prioritized replay).  Layout = the reference's: one implicit binary heap per tree, float64,
 * 2*size-1 entries, root at [0], children of p at 2p+1 / 2p+2, leaves at [size-1, 2*size-1); `size` a power of two.
 * memories/non_episodic/prioritized_experience_replay.py:43-156 (SegmentTree).
 * ===================================================================================================================*/

/* Marks [ptr, ptr+bytes) as L2-persisting for kernels subsequently launched on `stream` (stream access-policy window
 * + persisting-L2 carve-out).  Used for the top levels of the sum tree (the first 2^k entries of the heap array are
 * its top k levels), which every sample's descent re-reads while ~100 MB of minibatch traffic per step would
 * otherwise evict them from the 126 MB L2.  ptr == NULL clears the window.  CB200_ERR_UNSUPPORTED if the device has
 * no persisting-L2 support. */
int cb200_l2_persist(const void* ptr, int64_t bytes, void* stream);

/* SegmentTree.__init__ :54-61 -- sum tree <- 0, min tree <- +inf, max tree <- -inf. `winner` (int32[size]) is the
 * scratch array used by cb200_per_update for last-writer-wins duplicate resolution; it is set to -1. */
int cb200_per_init(double* sum_tree, double* min_tree, double* max_tree, int32_t* winner, int64_t size, void* stream);

/* PrioritizedExperienceReplay.update_priorities :203-217 + _update_priority :188-201 + SegmentTree.update/_propagate
 * :116-129,:63-74, for a whole batch at once.  Leaf idx[i] of the sum and min trees receives p_alpha[i], of the max
 * tree p_raw[i]; duplicates resolve last-writer-wins in batch order (the reference applies them sequentially); every
 * ancestor is then recomputed as op(left, right) level by level, which is bit-identical to the sequential reference
 * because parents are recomputed, never incrementally adjusted.  `max_priority_out` (device double, may be NULL)
 * receives max_tree[0] (= self.maximal_priority, :201).  n <= 1024 runs as ONE launch of one CTA that sorts the batch's
 * leaves in shared memory and walks all paths bottom-up without global round trips between levels; larger n uses one
 * launch per tree level.  Entries that must not be applied are skipped: a leaf outside [0, size) (the reference raises
 * ValueError :124-126; here bit 1 of *error_flags, device int32, may be NULL, is set) and entries whose p_alpha is
 * negative (the marker cb200_per_priorities_device leaves for a negative / NaN error). */
int cb200_per_update(double* sum_tree, double* min_tree, double* max_tree, int32_t* winner, int64_t size,
                     const int64_t* idx, const double* p_alpha, const double* p_raw, int64_t n,
                     double* max_priority_out, int32_t* error_flags, void* stream);

/* priority = error + epsilon; p_raw = priority; p_alpha = priority ** alpha  (:197-200) computed ON DEVICE with CUDA's
 * pow (<= 2 ulp; glibc's pow, which the reference uses, is not correctly rounded either, so device and reference
 * can differ in the last bit of ~0.1% of leaves -- see cb200_host_priorities for the libm-exact route).
 * Bit 0 of *neg_flag (device int32, may be NULL) is set if any err[i] is negative or NaN (the reference raises
 * ValueError :195); such entries get p_alpha = p_raw = -1, which cb200_per_update skips, so an invalid error can never
 * reach the trees. */
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
 * One warp per sample; each round fetches a 7-level sub-tree (254 nodes, 8 coalesced loads per lane) into a per-warp
 * shared-memory scratch and replays the reference's comparisons from it, so a 2^20-leaf descent costs 3 dependent
 * memory round trips instead of 20. */
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

/* Fused input path of the image agents (one launch): PER sample -> gather the uint8 frames of the drawn slots -> bf16
 * plane of their space-to-depth(s) view, i.e. the operand of the first convolution (cb200_u8_s2d_planes documents the
 * view and the plane format).  Replaces, for the image columns, the staged uint8 copy of cb200_per_sample_gather
 * (memories/non_episodic/prioritized_experience_replay.py:219-262 + core_types.py:488-511) AND the conversion passes
 * over it (embedders/embedder.py:103 input rescale is applied by the GEMM, see cb200_tgemm_desc.a_u8_div).
 *   image_columns[k]: src = ring column (uint8 [capacity, h*w*c]), dst = plane (bf16 [(h/s)*(w/s)*n, s*s*c],
 *                     core-tiled), row_bytes = h*w*c; 1 or 2 columns (state, next_state)
 *   small_columns   : the remaining columns, copied row by row into their staged [n, row_bytes] buffers
 * n must be a multiple of 8 (whole 8-row groups of the plane matrix).  idx_out / w_out / w32_out as cb200_per_sample.
 * Frame-deduplicated ring (`frames` != NULL, s == c == 4): the ring stores every h x w frame ONCE in `frames`
 * (uint8 [frame slots, h*w]) -- the reference shares them between s, s' and neighbouring transitions through LazyStack
 * (filters/observation/observation_stacking_filter.py:27-41, agents/agent.py:905-973); image_columns[k].src is then the
 * int32 [capacity, c] table of the frame slots of each transition's stack (row_bytes = 4*c) and the kernel assembles
 * the last-axis stack while converting; frame_slots = rows of `frames` (a stack in four consecutive slots is fetched
 * by one 2-D TMA box per chunk). */
int cb200_per_sample_gather_s2d(const double* sum_tree, const double* min_tree, int64_t size, const double* u, int64_t n,
                                int64_t num_transitions, double beta, int64_t* idx_out, double* w_out, float* w32_out,
                                const cb200_column* image_columns, int n_image, int32_t h, int32_t w, int32_t c,
                                int32_t s, const cb200_column* small_columns, int n_small, const void* frames,
                                int64_t frame_slots, void* stream);

/* The same for given slot indices (uniform ExperienceReplay.sample, experience_replay.py:71-93). */
int cb200_gather_s2d(const int64_t* idx, int64_t n, const cb200_column* image_columns, int n_image, int32_t h, int32_t w,
                     int32_t c, int32_t s, const cb200_column* small_columns, int n_small, const void* frames,
                     int64_t frame_slots, void* stream);

/* Frame-deduplicated ring, un-fused readers (Batch.states() of the slots idx, core_types.py:488-511):
 * out[i, pix, c] = frames[frame_index[idx[i], c], pix], i.e. np.stack(frames, axis=-1) of observation_stacking_filter.py:
 * 37-41 as a uint8 [n, frame_bytes, stack] array. */
int cb200_gather_stack(const void* frames, int64_t frame_bytes, const int32_t* frame_index, int32_t stack, const int64_t* idx,
                       int64_t n, void* out, void* stream);

/* Ring append: src[c][(cursor + i) % capacity, :] = staged[c][i, :] (experience_replay.py:131-150 store; the
 * `cb200_column.src` member is the ring (written), `.dst` the staged rows (read)). */
int cb200_scatter_ring(const cb200_column* h_columns, int n_columns, int64_t cursor, int64_t capacity, int64_t n,
                       void* stream);

/* Same append for a PACKED staging area, all columns in one launch: column c of staged record i is at
 * h_columns[c].dst + i * staged_stride (the host staging area of DeviceRing.store is one pinned record per transition,
 * so a flush is one H2D copy and one kernel). */
int cb200_scatter_ring_packed(const cb200_column* h_columns, int n_columns, int64_t staged_stride, int64_t cursor,
                              int64_t capacity, int64_t n, void* stream);


/* =====================================================================================================================
 * Learn step: dense contractions.  One primitive ("gather-GEMM") covers conv forward (implicit im2col over NHWC),
 * conv weight / data gradients and dense forward / backward:
 *
 *      C[m, n] = epilogue( sum_r A(m, r) * B(r, n) ),      A(m, r) = a_src[ a_rowoff[m] + a_coloff[r] ]   (elements)
 *
 * Replaces the TensorFlow ops behind architectures/tensorflow_components/layers.py:108-183 (Conv2d / Dense),
 * embedders/embedder.py:95-124 (x / 255 input rescale, via a_lut) and tf.gradients over them
 * (architectures/tensorflow_components/architecture.py:193).  fp32 FFMA, deterministic (fixed reduction order).
 * ===================================================================================================================*/
#define CB200_ACT_NONE 0
#define CB200_ACT_RELU 1
#define CB200_ACT_TANH 2

typedef struct cb200_gemm_desc {
    /* A operand */
    const void* a_src;          /* fp32 (a_lut == NULL) or uint8 (value = a_lut[byte]) element array                */
    const float* a_lut;         /* 256-entry table, e.g. lut[v] = (float)v / 255.0f                                  */
    const int32_t* a_rowoff;    /* [a_rows] element offset contributed by the row index                              */
    const int32_t* a_coloff;    /* [a_cols] element offset contributed by the column (reduction) index               */
    const int32_t* a_rowinfo;   /* optional (i << 16 | j) per row   } A(m, r) = 0 unless 0 <= i - a < a_oh and       */
    const int32_t* a_colinfo;   /* optional (a << 16 | b) per col   }                    0 <= j - b < a_ow           */
    int32_t a_oh, a_ow;
    int32_t a_rows, a_cols;     /* logical extent of A                                                              */
    int32_t a_transposed;       /* 0: C[a_rows, N] = A * B   (reduction over a_cols)                                */
                                /* 1: C[a_cols, N] = A^T * B (reduction over a_rows; weight gradients)              */
    /* B operand: row-major [R, N] */
    const float* b;
    int32_t ldb;
    int32_t n;
    /* output / epilogue */
    float* c;                   /* [rows, ldc]                                                                       */
    int32_t ldc;
    const float* bias;          /* [n] or NULL                                                                       */
    int32_t act;                /* CB200_ACT_* applied after the bias                                                */
    const float* mask_y;        /* optional, indexed like c: c = value * act'(mask_y) with act' from the activation  */
    int32_t mask_act;           /*   OUTPUT (relu: y > 0, tanh: 1 - y*y) -- fuses the activation backward            */
    const int32_t* c_rowmap;    /* optional output row remap (transposed-conv stride classes)                        */
    int32_t accumulate;         /* c += value instead of c = value                                                   */
    /* split reduction */
    float* workspace;           /* >= splits * rows * n floats when splits > 1                                       */
    int32_t splits;             /* 0 / 1 = no split; k > 1 = k partial sums reduced in fixed order                   */
    /* fast-path hints (the tables are built by the caller, who knows their structure) */
    int32_t a_vec4;             /* 1: every aligned group of 4 column indices is contiguous in memory, a_cols % 4 == 0 */
                                /*    and every a_rowoff % 4 == 0  => 128-bit (fp32) / 32-bit (uint8) operand loads   */
    int32_t a_ones_col;         /* a_transposed only: 1 = append an output row a_cols holding sum_m B[m, :] (the bias  */
                                /*    gradient lands in c[a_cols, :], i.e. right behind the kernel gradient)          */
    float a_u8_div;             /* uint8 A only, 0 = not declared: the caller states a_lut[v] == (float)v / a_u8_div.  */
                                /*    The tensor-core path then contracts the raw integers (exact in bf16) and divides */
                                /*    each accumulated sum by a_u8_div once; other paths read a_lut and ignore this.   */
    /* bf16 operand planes (hi, mid, lo; x == hi + mid + lo exactly) in the 8x8 core-tiled format: a [rows, cols]    */
    /* matrix stores element (r, c) of plane p at                                                                    */
    /*     planes + p * plane_stride + ((r / 8) * (cols / 8) + c / 8) * 64 + (r % 8) * 8 + c % 8                     */
    /* Activations / gradients use rows = pixel * batch + b.  Consumed natively by cb200_gemm_tiled; this entry point */
    /* can read B planes (conv1: weights, dY) and write the planes of its result.                                    */
    const void* b_planes;       /* planes of b [R, n] (ldb == n)                                                     */
    int64_t b_plane_stride;
    int32_t b_prow_npix;        /* > 0: reduction row r = b * npix + q (NHWC) is plane row q * b_prow_batch + b       */
    int32_t b_prow_batch;
    void* c_planes;             /* if set, the epilogue also writes the planes of the final c values                */
    int64_t c_plane_stride;
    int32_t c_plane_cols;       /* columns of the plane matrix (= n)                                                */
    int32_t c_prow_npix;        /* > 0: output row m = b * npix + q is plane row q * c_prow_batch + b; 0: row m       */
    int32_t c_prow_batch;
    int32_t a_lda;              /* > 0: A is a plain row-major fp32 matrix with this leading dimension (the tables   */
                                /*    say the same); lets products with n <= 8 or a_cols <= 8 take the skinny kernels */
} cb200_gemm_desc;

int cb200_gemm(const cb200_gemm_desc* h_desc, void* stream);

/* =====================================================================================================================
 * Tensor-core GEMMs on pre-split operands (csrc/nn_gemm_tiled.cuh).  Convolutions and dense layers as "multi-tap"
 * contractions over plane matrices (rows = pixel * batch + b, cols = channels; weights = stacks of [a_cols, n] blocks):
 *   mode 0: C[q * B + b, :] = sum over the tap list of output pixel q, entries (a_pix, w_blk):
 *                                A[a_pix * B + b, :] * W[w_blk]                (forward, data gradient)
 *   mode 1: C[t * a_cols + c, :] = sum_q sum_b A[a_pix[t * num_q + q] * B + b, c] * G[q * B + b, :]   (weight gradient)
 * Replaces, for layers whose input already lives on the device as planes, the same reference code as cb200_gemm
 * (layers.py:108-183 forward, tf.gradients backward).  Operands move by 1-D bulk copies (TMA); 3xBF16 products with
 * fp32 TMEM accumulation, at most 32 reduction chunks of 32 per launch slice (use `splits`).
 * ===================================================================================================================*/
typedef struct cb200_tgemm_desc {
    int32_t mode;
    int32_t batch;              /* B, multiple of 32                                                                 */
    const void* a_planes;       /* planes of A [a_pixels * B, a_cols]                                                */
    int64_t a_plane_stride;
    int32_t a_cols;             /* 32, 64, 128 or a multiple of 128                                                  */
    const void* b_planes;       /* mode 0: weight blocks [blocks][a_cols, n];  mode 1: G [num_q * B, n]               */
    int64_t b_plane_stride;
    int32_t n;                  /* 32 or a multiple of 64                                                            */
    const int32_t* list_ptr;    /* mode 0: [num_q + 1] offsets into list                                             */
    const int32_t* list;        /* mode 0: pairs (a_pix, w_blk)                                                      */
    int32_t max_list_len;       /* mode 0: longest tap list                                                          */
    const int32_t* a_pix;       /* mode 1: [taps * num_q] input pixel under tap t at output pixel q                  */
    int32_t num_q;
    int32_t taps;
    /* output / epilogue: as in cb200_gemm_desc; rows of C are q * B + b (mode 0) or t * a_cols + c (mode 1)          */
    /* c may be NULL when c_planes is set: only the planes of the result are produced (forward-only networks)         */
    float* c;
    int32_t ldc;
    const float* bias;
    int32_t act;
    const float* mask_y;
    int32_t mask_act;
    const int32_t* c_rowmap;    /* e.g. q * B + b -> b * num_q + q to store NHWC                                      */
    float* workspace;
    int32_t splits;
    void* c_planes;             /* tiled planes of C with plane row = C row, c_plane_cols == n                       */
    int64_t c_plane_stride;
    int32_t c_plane_cols;
    const void* mask_planes;    /* optional, instead of mask_y: the masking activation as tiled planes with the geometry */
    int64_t mask_plane_stride;  /*    of the result (rows q * B + b, n columns); c_plane_cols must be set to n            */
    int32_t bias_row;           /* mode 1: 1 = also produce row taps * a_cols = sum over all rows of G (the bias gradient, */
                                /*    stored right behind the kernel gradient); c / workspace / c_rowmap have one more row  */
    int32_t a_num_planes;       /* 3 (0 = 3): fp32 split;  1: A holds raw uint8 values as ONE exact bf16 plane, every    */
    float a_u8_div;             /*    accumulated sum is divided by a_u8_div (x / 255 input rescale, embedder.py:103)  */
    int64_t a_rows;             /* rows of the A plane matrix (a_pixels * B) and of the B operand's plane matrix      */
    int64_t b_rows;             /*   (mode 0: blocks * a_cols, mode 1: num_q * B): bounds of the TMA tensor maps      */
    int32_t b_interleaved;      /* mode 0, n <= 64 (or n % 128 != 0): the B planes are "row-group interleaved" --        */
                                /*   (row group | plane | column core | 64), written with plane_stride -1 by              */
                                /*   cb200_split_planes (segment layout 1) / cb200_permute_f32 -- and the 3xBF16 product  */
                                /*   set is issued as three wide tcgen05.mma on the [b1|b2|b3] operand                    */
    const int32_t* a_pix_host;  /* mode 1, optional: HOST copy of a_pix.  When the taps that share a 128-row tile of the   */
                                /*   result sit a constant number of pixels apart (at most three distinct strides over   */
                                /*   the tiles: every convolution of the path), the A^T operand of a reduction chunk is  */
                                /*   ONE 5-D TMA box (64 | cores | taps | row groups | planes) instead of one 1-D bulk   */
                                /*   copy per (plane, tap, row group); NULL: bulk copies                                 */
    /* private: TMA tensor maps of the operands, built by the first call with this descriptor (keep the descriptor    */
    /* alive and unchanged between calls; zero-initialise)                                                            */
    uint64_t tmap_key;
    int32_t a_tma;              /* private: number of A^T tensor maps in use (0: bulk copies)                           */
    uint8_t a_tile_class[64];   /* private: tensor map of each 128-row tile                                             */
    uint8_t tmap_storage[4 * 128 + 64];
} cb200_tgemm_desc;

int cb200_gemm_tiled(const cb200_tgemm_desc* h_desc, void* stream);

/* uint8 NHWC frames [batch, h, w, c] -> one exact bf16 plane of the space-to-depth(s) view: plane row
 * ((y / s) * (w / s) + x / s) * batch + b, column ((y % s) * s + x % s) * c + ch; [h/s * w/s * batch, s*s*c] tiled.
 * Turns the strided first convolution (Atari: 8x8 stride 4 on 84x84x4) into a 2x2 stride-1 one on 64 channels. */
int cb200_u8_s2d_planes(const void* x, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t s, void* plane,
                        void* stream);

/* fp32 row-major matrices -> tiled planes, one launch for a list of matrices inside one fp32 buffer (the parameter
 * buffer, once per step): d_segments[k] = {src offset, rows, cols, plane offset, layout} in elements (device memory);
 * layout 0: three planes `plane_stride` apart; layout 1: row-group interleaved planes (cb200_tgemm_desc.b_interleaved),
 * the segment occupies 3 * rows * cols elements from its plane offset. */
int cb200_split_planes(const float* src, void* planes, int64_t plane_stride, const int64_t* d_segments,
                       int32_t num_segments, int64_t max_segment_elems, void* stream);

/* out[j] = sum_i x[i, j] for x [rows, cols] (bias gradients: tf.gradients wrt the bias of Dense / Conv2d), reduced in a
 * fixed order (two deterministic stages; `workspace` >= 1024 * cols floats). */
int cb200_colsum(const float* x, int64_t rows, int64_t cols, float* out, float* workspace, void* stream);

/* dst[i] = src[table[i]], i < n  (fp32; static permutations of weight tensors for the data-gradient GEMMs, e.g. the
 * per-stride-class [taps*N, Cin] matrices of the transposed convolution) */
/* (plane_stride -1: row-group interleaved planes, see cb200_tgemm_desc.b_interleaved) */
int cb200_permute_f32(const float* src, const int32_t* table, int64_t n, float* dst, void* dst_planes,
                      int64_t plane_stride, int32_t plane_cols, void* stream);
                      /* dst_planes optional (NULL): also write dst, seen as [n / plane_cols, plane_cols], as planes */

/* dst[c, r] = src[r, c]  (fp32; pre-transposition of weight matrices for the data-gradient GEMMs) */
int cb200_transpose(const float* src, int64_t rows, int64_t cols, float* dst, void* dst_planes, int64_t plane_stride,
                    void* stream);


/* =====================================================================================================================
 * Learn step: element-wise / reduction kernels.
 * ===================================================================================================================*/

/* DQNAgent.learn_from_batch, agents/dqn_agent.py:92-103 (+ ddqn_agent.py:42-43 action selection):
 *   a*_i   = argmax_a q_select[i, a]                     (first maximum, np.argmax)
 *   y_i    = r_i + (1.0 - done_i) * discount * q_next[i, a*_i]        evaluated in fp64 like the Python loop
 *   err_i  = |y_i - q_online[i, act_i]|  (fp64)          -> td_err_out (the new priorities' input)
 *   targets = copy of q_online with targets[i, act_i] = (float) y_i
 * q_select = q_next for DQN, Q_online(s') for DDQN. */
int cb200_dqn_td_targets(const float* q_next, const float* q_select, const float* q_online, const int64_t* actions,
                         const double* rewards, const uint8_t* game_overs, double discount, int64_t batch,
                         int64_t n_actions, float* targets_out, double* td_err_out, void* stream);

/* Generic head loss of heads/head.py:165-177 for a Q / V style regression head:
 *   loss = mean_b( loss_weight * w_b * sum_a l(target_ba, out_ba) ),  l = Huber(delta=1) (tf.losses.huber_loss,
 *   q_head.py:44-45) or squared error (tf.losses.mean_squared_error);  w = importance weights (NULL => ones).
 *   d_out[b, a] = loss_weight * w_b / batch * l'(out_ba - target_ba)
 * loss_out: device float (fixed-order reduction). */
int cb200_regression_head_loss_grad(const float* out, const float* target, const float* weights, int64_t batch,
                                    int64_t width, int huber, float loss_weight, float* d_out, float* loss_out,
                                    void* stream);

/* Fused DQN / DDQN Q-head step: Q(s') of the target head, Q(s) [and Q(s'), DDQN] of the online head (q_head.py:52-54),
 * TD targets / errors (agents/dqn_agent.py:92-103, ddqn_agent.py:42-43; fp64, bit-exact given the Q values), Huber / MSE
 * head loss and dL/dQ (heads/head.py:165-177), and the head's backward pass: dL/dW, dL/db and the gradient w.r.t. the
 * feature layer's pre-activation (dQ W^T masked with relu'(h)), as fp32 and / or as operand planes.  Two launches
 * instead of the eight or nine of cb200_gemm x 5 + cb200_dqn_td_targets + cb200_regression_head_loss_grad. */
typedef struct cb200_dqn_head_desc {
    const float* h_next;        /* [batch, features] post-ReLU features of s' from the TARGET network                      */
    const float* h_online;      /* [batch, features] features of s from the online network                                */
    const float* h_select;      /* DDQN: features of s' from the ONLINE network (action selection); NULL for DQN          */
    const float* w_target;      /* target head kernel [features, n_actions] and bias                                      */
    const float* b_target;
    const float* w_online;
    const float* b_online;
    const int64_t* actions;     /* [batch]                                                                                 */
    const double* rewards;
    const uint8_t* game_overs;
    const float* weights;       /* importance weights [batch] or NULL                                                     */
    double discount;
    int32_t huber;              /* 1: tf.losses.huber_loss(delta 1), 0: mean squared error                                */
    int64_t batch;
    int32_t features;           /* 256 or 512                                                                              */
    int32_t n_actions;          /* <= 8                                                                                    */
    float* q_online;            /* out [batch, n_actions]                                                                  */
    float* q_next;              /* out, optional                                                                           */
    float* targets;             /* out [batch, n_actions]: Q(s) with the taken action's entry replaced by the TD target    */
    double* td_err;             /* out [batch]: |target - Q(s, a)| (the PER priorities' errors)                           */
    float* dq;                  /* out [batch, n_actions]: dL/dQ                                                           */
    float* loss;                /* out scalar, optional                                                                    */
    float* dh;                  /* out, optional: [batch, features] dL/d(pre-activation of the feature layer)             */
    void* dh_planes;            /* out, optional: the same as tiled bf16 hi / mid / lo planes                              */
    int64_t dh_plane_stride;
    float* dw;                  /* out [features, n_actions]: gradient of the online head kernel                          */
    float* db;                  /* out [n_actions]                                                                         */
    float* workspace;           /* ceil(batch / 16) * 8 * (features * n_actions + n_actions + 1) floats                    */
} cb200_dqn_head_desc;

int cb200_dqn_head_fused(const cb200_dqn_head_desc* h_desc, void* stream);

/* DuelingQHead (heads/dueling_q_head.py:33-47): q = v + (adv - mean_a adv); backward: d_v = sum_a dq,
 * d_adv = dq - mean_a dq. */
int cb200_dueling_combine_fwd(const float* v, const float* adv, int64_t batch, int64_t n_actions, float* q,
                              void* stream);
int cb200_dueling_combine_bwd(const float* dq, int64_t batch, int64_t n_actions, float* d_v, float* d_adv,
                              void* stream);

/* sum of squares of a flat fp32 buffer in a fixed order -> *out (device float); tf.global_norm =
 * sqrt(sum_t sum(t^2)) (architecture.py:194).  workspace >= 1024 floats. */
int cb200_sumsq(const float* x, int64_t n, float* out, float* workspace, void* stream);

/* tf.clip_by_global_norm (architecture.py:239-240): g *= clip / max(sqrt(*sumsq), clip)  -- in place. */
int cb200_clip_by_global_norm(float* g, int64_t n, const float* sumsq, float clip, void* stream);

/* g *= s (apply_gradients `scaler`, architecture.py:485-493: 1/num_workers for sync training) */
int cb200_scale(float* g, int64_t n, float s, void* stream);

/* tf.train.AdamOptimizer step with TF-1.x semantics (general_network.py:390-394; kernel form of
 * tensorflow/core/kernels/training_ops.cc ApplyAdam):
 *   alpha = lr * sqrt(1 - beta2_power) / (1 - beta1_power)       (fp32; the powers are the fp32 running products)
 *   m += (g - m) * (1 - beta1);  v += (g*g - v) * (1 - beta2);  theta -= (m * alpha) / (sqrt(v) + epsilon)
 * over the whole flat parameter buffer in one launch. */
int cb200_adam_tf(float* theta, float* m, float* v, const float* g, int64_t n, float lr, float beta1, float beta2,
                  float epsilon, float beta1_power, float beta2_power, void* stream);

/* Same optimizer step with the running powers kept in DEVICE memory (state = {beta1_power, beta2_power}, initialised
 * to {beta1, beta2}); the step multiplies them afterwards.  All launch parameters are constant from step to step, so
 * a complete training step can be captured in a CUDA graph and replayed. */
int cb200_adam_tf_dev(float* theta, float* m, float* v, const float* g, int64_t n, float lr, float beta1, float beta2,
                      float epsilon, float* state, void* stream);

/* *x += delta on the device (minibatch cursor of a captured epoch loop, see cb200_gather_at) */
int cb200_add_i64(int64_t* x, int64_t delta, void* stream);

/* NetworkWrapper.update_target_network -> set_weights (architecture.py:598-607):
 *   target = rate * online + (1 - rate) * target   in fp32, rate and (1 - rate) rounded to fp32 first (numpy). */
int cb200_polyak(float* target, const float* online, int64_t n, double rate, void* stream);

/* PPOHead for continuous actions (heads/ppo_head.py:52-98,118-144): diagonal Gaussian policy with a state-independent
 * log-std variable, sigma = exp(logstd) + 1e-15; likelihood ratio exp(logp - logp_old) (:78), clipped to
 * 1 +- clip_eps (clip_eps = clip_likelihood_ratio_using_epsilon * clipping_decay rescaler, :80-84), surrogate
 * L = -mean(min(ratio*A, clip(ratio)*A)) (:85-90), entropy regulariser -beta*H (:93-95).  The old policy is given by
 * its mean per sample and its log-std vector (the frozen target network, clipped_ppo_agent.py:240).
 * Outputs d(L)/d(mu) [batch, action_dim], d(L)/d(logstd) [action_dim] and scalars[5] = {loss, KL(old||new), entropy,
 * mean ratio, mean clipped ratio} (the signals clipped_ppo_agent.py:227-230 fetches).  action_dim <= 32. */
int cb200_ppo_continuous_head(const float* mu, const float* logstd, const float* actions, const float* old_mu,
                              const float* old_logstd, const float* advantages, int64_t batch, int32_t action_dim,
                              float clip_eps, float beta_entropy, float* d_mu, float* d_logstd, float* scalars,
                              void* stream);

/* dst[c][i, :] = src[c][idx[*offset + i], :] for i < n (idx == NULL: rows *offset + i).  `offset` is a DEVICE scalar
 * so that the launch is identical for every minibatch of an epoch (CUDA-graph replay; only *offset changes).
 * Minibatch slicing of clipped_ppo_agent.py:232-265 after batch.shuffle(). */
int cb200_gather_at(const cb200_column* h_columns, int n_columns, const int64_t* idx, const int64_t* offset,
                    int64_t n, void* stream);

/* dz[r, c] = dy[r, c] * act'(y[r, c]) with independent leading dimensions: activation backward on a column block of a
 * wider buffer (the embedder part of a critic's concatenated [action, embedding] input, general_network.py:272-277). */
int cb200_act_backward(const float* dy, int32_t ld_dy, const float* y, int32_t ld_y, int64_t rows, int32_t cols,
                       int32_t act, float* dz, int32_t ld_dz, void* stream);

/* dst[r, c] = alpha * src[r, c] + beta * dst[r, c] on strided 2-D fp32 blocks (beta == 0: dst is not read).  Used for
 * the embedding merger concat, the actor's output scale and the -(1/B) * dQ/da seed of the actor update
 * (ddpg_agent.py:171-186). */
int cb200_axpby_2d(const float* src, int32_t ld_src, int64_t rows, int32_t cols, float alpha, float beta, float* dst,
                   int32_t ld_dst, void* stream);

/* Bootstrapped critic targets of DDPG / TD3 / SAC (ddpg_agent.py:156-164, td3_agent.py:172-181,
 * soft_actor_critic_agent.py:265-266): y = r + (1 - done) * discount * q_next in fp64 (numpy), optional clip, stored
 * as fp32.  q_next is read with stride ld_q. */
int cb200_ac_td_targets(const double* rewards, const uint8_t* game_overs, const float* q_next, int32_t ld_q,
                        int64_t batch, double discount, int32_t use_non_zero_discount_for_terminal_states,
                        int32_t use_clip, double clip_lo, double clip_hi, float* targets_out, void* stream);

/* out = min(a, b) element-wise (clipped double-Q: td3_v_head.py:61, sac_q_head.py:84-86) */
int cb200_min2(const float* a, const float* b, int64_t n, float* out, void* stream);

/* TD3 target policy smoothing (td3_agent.py:162-164): a = clip(a + clip(noise, -noise_clip, noise_clip), lo, hi);
 * noise is the fp64 np.random.normal draw, the sum and both clips are evaluated in fp64 like numpy does and rounded
 * to fp32 once (bit-exact with the reference, tests/golden/agent_prologues.npz) */
int cb200_td3_smooth_actions(float* actions, const double* noise, int64_t n, double noise_clip, double lo, double hi,
                             void* stream);

/* CategoricalQHead + distributional TD targets (agents/categorical_dqn_agent.py:105-165, rainbow_dqn_agent.py:93-140,
 * architectures/tensorflow_components/heads/categorical_q_head.py:41-57).  Inputs are the [batch, n_actions, n_atoms]
 * head logits of target(s'), online(s) and -- for the double-Q rule of Rainbow, else NULL -- online(s').  z = the fp64
 * support (np.linspace); gamma_n = discount (** n_step); bootstrap (fp64 per sample, NULL -> 1 - game_over) is
 * info['should_bootstrap_next_state'].  Writes: labels = TD_targets fed to the train op (online softmax, projected
 * distribution m on the taken action's row; projection accumulated in fp64 in the reference's loop order, bit-exact
 * given equal probabilities), dlogits = d(total loss)/d(online logits) (softmax - labels on the taken row, 0 elsewhere),
 * loss_rows [batch, n_actions] = tf.nn.softmax_cross_entropy_with_logits, total_loss = their sum
 * (general_network.py:360), td_err = loss_rows[b, action[b]] (what update_priorities is handed, :160-163), optional
 * q_online [batch, n_actions] fp64 (distribution_prediction_to_q_values) and target_actions.  next_is_prob != 0: `next`
 * / `select` already hold probabilities (parity tests feed the fixture's network outputs). */
int cb200_c51_head(const float* next, const float* online, const float* select, const int64_t* actions,
                   const double* rewards, const uint8_t* game_overs, const double* bootstrap, const double* z,
                   double gamma_n, int32_t batch, int32_t n_actions, int32_t n_atoms, int32_t next_is_prob,
                   float* labels, float* dlogits, float* loss_rows, float* total_loss, double* td_err,
                   double* q_online, int64_t* target_actions, void* stream);

/* q_values output of the CategoricalQHead (categorical_q_head.py:56): q[r] = sum_j (double)softmax(logits[r, :])_j * z[j]
 * for rows = batch * n_actions rows of n_atoms logits; z = the fp32-rounded support cast back to fp64 (:36-37) */
int cb200_c51_q_values(const float* logits, const double* z, int64_t rows, int32_t n_atoms, double* q_out, void* stream);

/* SACPolicyHead (heads/sac_head.py:60-97).  head_out [batch, 2*action_dim] = [mu | raw log-sigma]; log-sigma is clipped
 * to [-20, 2]; u = mu + exp(log_sigma) * eps; a = tanh(u); logp = MVN-diag log-prob of u minus the tanh squash
 * correction sum_j log(1 - a_j^2 + 1e-6).  Any output may be NULL. */
int cb200_sac_policy_sample(const float* head_out, const float* eps, int64_t batch, int32_t action_dim, float* raw_out,
                            float* actions_out, float* logp_out, void* stream);

/* d/d(head_out) of  mean_b logp(eps_logp)  -  sum_b <dq_da_b, tanh(mu + sigma * eps_q)_b>  -- the combination
 * policy_grads = dlogp_dphi - dq_dphi of soft_actor_critic_agent.py:213-232, each term with its own noise sample
 * (the reference evaluates them in separate sess.run calls which re-sample, sac_head.py:80). */
int cb200_sac_policy_grad(const float* head_out, const float* eps_logp, const float* eps_q, const float* dq_da,
                          int64_t batch, int32_t action_dim, float* d_head_out, void* stream);

/* qmin = min(q1, q2) and the seeds d mean_b(qmin) / dq1, dq2 (sac_q_head.py:84-86); outputs may be NULL. */
int cb200_sac_min_seed(const float* q1, const float* q2, int64_t batch, float* d1, float* d2, float* qmin,
                       void* stream);

/* out = a - b  (fp32) */
int cb200_sub(const float* a, const float* b, int64_t n, float* out, void* stream);

/* out[i] = (float) in[i] */
int cb200_f64_to_f32(const double* in, int64_t n, float* out, void* stream);

/* =====================================================================================================================
 * Scalar RL recurrences (fp64, as the reference computes them).
 * ===================================================================================================================*/

/* Generalised advantage estimation over a whole rollout of n transitions laid out episode after episode:
 * ActorCriticAgent.get_general_advantage_estimation_values (agents/actor_critic_agent.py:111-125: deltas, then
 * scipy.signal.lfilter([1],[1,-gamma*lambda]) on the reversed deltas) applied per episode as ClippedPPOAgent.
 * fill_advantages does (agents/clipped_ppo_agent.py:170-207): episodes end at game_over flags, the bootstrap value
 * appended at an episode end is 0 (:188), the value target is advantage + V(s_t) (:121), transitions after the last
 * game_over receive nothing (*n_valid = index of the last game_over + 1; later entries are still written but must be
 * ignored, cf. zip() truncation :203).  values[t] = V(s_t) (fp32 network output).  One block-wide scan of affine maps
 * (thread chunks -> warp shuffles -> shared memory). */
int cb200_gae_scan(const double* rewards, const float* values, const uint8_t* game_overs, int64_t n, double discount,
                   double gae_lambda, double* advantages, double* value_targets, int64_t* n_valid, void* stream);

/* x[:n_valid] = (x - mean) / std with the population std (clipped_ppo_agent.py:201), in place; x[n_valid:] = NaN.
 * n_valid may be NULL (= n).  mean_std_out (device double[2], may be NULL) receives mean and std. */
int cb200_standardize(double* x, int64_t n, const int64_t* n_valid, double* mean_std_out, void* stream);

/* Episode.update_discounted_rewards (core_types.py:771-790): out[t] = sum_{k < n_step} discount^k * r[t+k] inside t's
 * episode [ep_start[t], ep_end[t]), n_step == -1 => to the end of the episode; accumulated in the reference's order
 * (k ascending, running power of the discount), bit-identical to the numpy loop. */
int cb200_nstep_returns(const double* rewards, const int64_t* ep_start, const int64_t* ep_end, int64_t n,
                        double discount, int64_t n_step, double* out, void* stream);

/* NumpySharedRunningStats (utilities/shared_running_stats.py:115-164), used by ObservationNormalizationFilter
 * (filters/observation/observation_normalization_filter.py:71-78):
 *   push      : sum += sum_rows x, sumsq += sum_rows x^2   (fp64; the host adds `rows` to its count)
 *   finalize  : mean = sum/count; std = sqrt(max((sumsq - count*mean^2) / max(count-1, 1), epsilon))
 *   normalize : clip((x - mean) / (std + 1e-15), lo, hi) -> fp32 (network feed) and/or fp64 */
int cb200_running_stats_push(const float* x, int64_t rows, int64_t cols, double* sum, double* sumsq, void* stream);
int cb200_running_stats_finalize(const double* sum, const double* sumsq, double count, double epsilon, int64_t cols,
                                 double* mean, double* std_out, void* stream);
int cb200_running_stats_normalize(const float* x, int64_t rows, int64_t cols, const double* mean, const double* std_in,
                                  double clip_lo, double clip_hi, float* out32, double* out64, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COACH_B200_H */
