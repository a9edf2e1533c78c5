// coach_b200/csrc/nn_gemm_tiled.cuh -- tensor-core GEMMs whose operands are pre-split bf16 planes in the 8x8
// core-tiled format (nn_gemm.cuh: tiled_elem): convolutions and dense layers as "multi-tap" contractions.
//
// Activations / gradients are plane matrices with rows = pixel * batch + b (pixel-major, batch-inner) and cols =
// channels; weights are stacks of [rows, n] blocks.  Two contraction shapes cover the whole learn step:
//
//   mode 0 (forward / data gradient; A is K-major):
//       C[q * B + b, n] = sum over the tap list of output pixel q, entries (a_pix, w_blk):
//                            sum_c  A[a_pix * B + b, c] * W[w_blk][c, n]
//     conv forward   : q = output pixel, one entry per kernel tap (a_pix = the input pixel under the tap)
//     conv data grad : q = INPUT pixel, entries = the taps whose output pixel exists (gather form: no zero padding, no
//                      stride-class decomposition -- invalid taps are simply not in the list), W = per-tap W^T
//     dense          : one pixel, one entry; a dense layer on a flattened conv map = one entry per pixel
//   mode 1 (weight gradient; A^T is MN-major):
//       C[t * Ca + c, n] = sum_q sum_b  A[a_pix(t, q) * B + b, c] * G[q * B + b, n]
//
// A CTA owns one 128 x BN output tile: 128 consecutive batch rows of one pixel (mode 0) or 128 rows of the stacked
// per-tap weight matrix (mode 1).  Because of the plane format every operand chunk (32 reduction indices) is a few
// contiguous runs of 128-byte core matrices, so the PRODUCER warp moves it with 1-D bulk copies (cp.async.bulk,
// completion counted in bytes on an mbarrier -- the TMA engine, no registers, no shared-memory stores by threads), the
// MMA warp (one elected thread) issues the 3xBF16 product set of nn_gemm_tc.cuh as soon as the "full" barrier of a
// stage flips and releases the stage through tcgen05.commit -> "empty" barrier, and the four EPILOGUE warps drain
// the TMEM accumulators at the end (tc_epilogue: bias / activation / derivative mask, fp32 result + planes).
//
// Tensor maps.  A plane set is a 4-D bf16 tensor (64 elements of a core | cores per 8-row group | row groups | 3
// planes); one TMA tile operation with box (64, cores, row groups, 3) fetches the hi / mid / lo planes of a whole
// operand chunk and lands them densely -- exactly the layouts below.  Mode 0 needs TWO operations per chunk (A, B),
// mode 1 one for G plus 1-D bulk copies for the A^T runs (one per tap and 8-row group: taps sit at unrelated pixels
// and the M direction must stay uniformly strided in shared memory).  A TMA operation costs ~60 cycles of issue on
// an SM whatever its size (measured, profiles/), which is what makes few large operations matter.
//
// Shared-memory operand layouts (no swizzle; LBO = step between core matrices along K, SBO = along M / N):
//   A  K-major  [128 rows, 32 k] : core (rg, kg) at rg * 512 + kg * 128          LBO = 128,  SBO = 512
//   A^T MN-major [128 m,   32 k] : core (kg, mg) at kg * 2048 + mg * 128         LBO = 2048, SBO = 128
//   B  MN-major [32 k, BN n]     : core (kg, ng) at kg * (BN/8) * 128 + ng * 128 LBO = BN * 16, SBO = 128
#pragma once
#include <cuda.h>      // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint in nn.cu)

#include "nn_gemm_tc.cuh"

namespace cb200 {
namespace gemm {

#ifdef CB200_TC_PROF
// build-time instrumentation (CB200_EXTRA_NVCC_FLAGS=-DCB200_TC_PROF python -m coach_b200.build --force): cycles the
// producer lane 0 / the MMA thread / epilogue thread 0 of CTA (0,0,0) spend per phase; read with cb200_tc_prof_read
// (tools/tc_phase_probe.py)
__device__ unsigned long long g_tc_prof[16];
#define TC_PROF_T(var) const long long var = clock64()
#define TC_PROF_ADD(i, a, b) \
    if (prof_on) g_tc_prof[i] += (unsigned long long)((b) - (a))
#else
#define TC_PROF_T(var)
#define TC_PROF_ADD(i, a, b)
#endif

struct TiledParams {
    int mode;
    int batch;                 // B, multiple of 32
    const uint16_t* a;         // planes of the A matrix [a_pixels * B, a_cols]
    long long a_stride;
    int a_cols;                // Ca, multiple of 32
    const uint16_t* b;         // mode 0: weight blocks [blocks][Ca, n];  mode 1: G [q_pixels * B, n]
    long long b_stride;
    int n;
    const int32_t* list_ptr;   // mode 0: [num_q + 1]
    const int2* list;          // mode 0: (a_pix, w_blk)
    const int32_t* a_pix;      // mode 1: [taps * num_q]
    int num_q, taps;
    int chunks_per_split;
    float a_u8_div;            // NA == 1: A holds raw uint8 values (exact in bf16); every sum is divided by this
    int bias_row;              // mode 1: also produce row M = sum over all rows of G (the bias gradient)
    int a_tma;                 // mode 1: the A^T operand of a chunk is one 5-D TMA box (tensor map tile_class[tile])
    uint8_t tile_class[64];
};

// NA = planes of the A operand: 3 (fp32 split) or 1 (uint8 values, exact: 3 products instead of 6)
// warps 0-7: epilogue (two per TMEM lane quadrant, half of the tile's columns each; in mode 1 they also help issuing
// bulk copies during the main loop), warp 8: producer, warp 9: MMA issuer
constexpr int kTlThreads = 320;
constexpr int kTlProducerWarp = 8, kTlMmaWarp = 9;

template <int BN, int NA>
struct TiledCfg {
    static constexpr int kStages = BN == 128 ? 4 : 3;
    static constexpr size_t kSmemBytes = (size_t)kStages * (NA * kTcBM * kTcBK * 2 + 3 * BN * kTcBK * 2) + 128 + 4096;
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], "
        "[%7];" ::"r"(smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], "
        "[%6];" ::"r"(smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

// tmA: planes of A with box (64, 4, 16, NA) (mode 0 only); tmB: planes of the B operand with box (64, BN / 8, 4, 3)
// kCat (mode 0, BN <= 64, row-group interleaved B planes): the B box lands as [k-group][plane][column core], one
// MN-major operand [32 k, 3 BN] = [b1 | b2 | b3].  The product set becomes THREE tcgen05.mma per k16 step,
//     a1 x [b1|b2|b3] (N = 3 BN) -> columns MAIN | CA | CB       a2 x [b1|b2] (N = 2 BN) -> CA | CB
//     a3 x [b1]       (N = BN)   -> CA                            value = MAIN + (CA + CB)
// i.e. the same six products with each A plane read from shared memory once per k-step instead of 3 / 2 / 1 times: an
// SS-mode MMA of width 64 reads (128 + 64) x 16 x 2 B = 6 KB per 32 tensor cycles, 192 B/cycle against the 128 B/cycle
// shared memory delivers -- the six narrow MMAs were bound by shared-memory reads, not by the tensor pipe.
template <int BN, bool kTransA, int NA, bool kCat = false>
__global__ void __launch_bounds__(kTlThreads) gemm_tc_tiled_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            const __grid_constant__ CUtensorMap tmA1,
                                                            const __grid_constant__ CUtensorMap tmA2, TiledParams tp,
                                                            EpiParams ep, int M) {
    constexpr int S = TiledCfg<BN, NA>::kStages;
    constexpr int A_SPLIT = kTcBM * kTcBK * 2;            // 8 KB per plane
    constexpr int B_SPLIT = BN * kTcBK * 2;
    constexpr int STAGE = NA * A_SPLIT + 3 * B_SPLIT;
    constexpr int B_KG = (BN / 8) * 128;                  // bytes of one k-group (8 reduction rows) of the B tile
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * STAGE);     // [S]
    uint64_t* empty_bar = full_bar + S;                                    // [S]
    uint64_t* done_bar = empty_bar + S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
    uint8_t* ones_tile = smem + S * STAGE + 128;          // 4 KB: a [128 x 16] A^T operand of bf16 1.0 (bias row)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Bias gradient (mode 1): the CTAs of the first M tile also accumulate ones^T * G in two more TMEM accumulators --
    // every row of that product is sum_rows G[row, :], the epilogue keeps row 0.  Three extra MMAs per k16 step in
    // 1 / gridDim.x of the CTAs replace a separate column-sum pass over dY.
    const bool bias_cta = kTransA && tp.bias_row != 0 && blockIdx.x == 0;
    static_assert(!kCat || (!kTransA && BN <= 64), "kCat: mode 0, BN <= 64");
    const uint32_t tmem_cols_needed = kCat ? 3u * BN : (bias_cta ? 4u : 2u) * BN;
    const uint32_t TMEM_COLS = tmem_cols_needed <= 32 ? 32u : (tmem_cols_needed <= 64 ? 64u : (tmem_cols_needed <= 128 ? 128u : (tmem_cols_needed <= 256 ? 256u : 512u)));
    const int B = tp.batch, Ca = tp.a_cols, N = tp.n;
    const int n0 = blockIdx.y * BN;
    const int split = blockIdx.z;

    // ---- tile decode ----------------------------------------------------------------------------------------------
    int q = 0, b0 = 0, m0, m_end, total, list_lo = 0;
    const int kc_per = Ca / kTcBK;                        // mode 0: chunks per tap
    const int bc_per = B / kTcBK;                         // mode 1: chunks per pixel
    if (!kTransA) {
        const int tiles_per_q = (B + kTcBM - 1) / kTcBM;
        q = blockIdx.x / tiles_per_q;
        b0 = (blockIdx.x % tiles_per_q) * kTcBM;
        m0 = q * B + b0;
        m_end = q * B + B;
        list_lo = __ldg(tp.list_ptr + q);
        total = (__ldg(tp.list_ptr + q + 1) - list_lo) * kc_per;
    } else {
        m0 = blockIdx.x * kTcBM;
        m_end = M - (tp.bias_row ? 1 : 0);       // the bias row is not part of any tile
        total = tp.num_q * bc_per;
    }
    const int c_lo = split * tp.chunks_per_split;
    const int c_hi = min(total, c_lo + tp.chunks_per_split);
    const int nchunks = max(0, c_hi - c_lo);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    if (bias_cta && tid < 128) {     // (any 128 threads)
        // 2048 bf16 ones, 16 per thread; generic-proxy stores made visible to the tensor core (async proxy)
        uint4* o = reinterpret_cast<uint4*>(ones_tile) + 2 * tid;
        o[0] = o[1] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        fence_proxy_async_smem();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_main = *tmem_slot;
    const uint32_t tmem_corr = tmem_main + BN;
    const uint32_t tmem_bias_main = tmem_main + 2 * BN, tmem_bias_corr = tmem_main + 3 * BN;
#ifdef CB200_TC_PROF
    const bool cta0 = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#endif

    if (warp == kTlProducerWarp) {
        // ================= producer: bulk copies of the operand cores into the stage ring =========================
#ifdef CB200_TC_PROF
        const bool prof_on = cta0 && lane == 0;
#endif
        // mode 1: taps covered by this M tile, and the channel range inside a tap
        const int taps_in_tile = kTransA ? (Ca >= kTcBM ? 1 : min(kTcBM / Ca, tp.taps - m0 / Ca)) : 0;
        const int t0 = kTransA ? m0 / Ca : 0;
        const int cw = kTransA ? min(Ca, kTcBM) : 0;                           // channels per tap inside the tile
        const int c0 = kTransA ? m0 % Ca : 0;
        // a tensor-map box always delivers (and counts) its full size, rows past the batch included
        const bool a_tma = kTransA && tp.a_tma != 0;
        const int a_cls = a_tma ? tp.tile_class[blockIdx.x & 63] : 0;
        const CUtensorMap* a_map = a_cls == 0 ? &tmA : (a_cls == 1 ? &tmA1 : &tmA2);
        const uint32_t a_bytes = (kTransA && !a_tma) ? (uint32_t)(taps_in_tile * 4 * (cw / 8) * 128) : (uint32_t)A_SPLIT;
        const uint32_t tx_bytes = (uint32_t)NA * a_bytes + 3u * (uint32_t)B_SPLIT;
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            if (!kTransA || a_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(a_map) : "memory");
        }
        for (int j = 0; j < nchunks; ++j) {
            const int s = j % S, u = j / S;
            TC_PROF_T(t0c);
            if (u > 0) mbar_wait(empty_bar + s, (uint32_t)((u - 1) & 1));      // MMAs that read this stage are done
            TC_PROF_T(t1c);
            uint8_t* sA = smem + s * STAGE;
            uint8_t* sB = sA + NA * A_SPLIT;
            uint64_t* bar = full_bar + s;
            if (lane == 0) mbar_expect_tx(bar, tx_bytes);
            __syncwarp();
            const int cj = c_lo + j;
            if (!kTransA) {
                if (lane == 0) {
                    const int e = cj / kc_per, kc = cj % kc_per;
                    const int2 ent = __ldg(tp.list + list_lo + e);
                    tma_load_4d(sA, &tmA, 0, kc * 4, (int)(((size_t)ent.x * B + b0) >> 3), 0, bar);
                    if (kCat) tma_load_4d(sB, &tmB, 0, n0 >> 3, 0, (ent.y * Ca + kc * kTcBK) >> 3, bar);
                    else tma_load_4d(sB, &tmB, 0, n0 >> 3, (ent.y * Ca + kc * kTcBK) >> 3, 0, bar);
                }
            } else {
                const int qq = cj / bc_per, bc = cj % bc_per;
                if (lane == 0) tma_load_4d(sB, &tmB, 0, n0 >> 3, (int)(((size_t)qq * B + (size_t)bc * kTcBK) >> 3), 0, bar);
                if (a_tma) {
                    // A^T: the taps of the tile sit a constant number of pixels apart -- one 5-D box (64 | cores | taps
                    // | 4 row groups | planes) lands as [plane][k-group][tap][core], the layout of the bulk path below
                    if (lane == 0) {
                        const int apix = __ldg(tp.a_pix + (size_t)t0 * tp.num_q + qq);
                        tma_load_5d(sA, a_map, 0, c0 >> 3, 0, (int)(((size_t)apix * B + (size_t)bc * kTcBK) >> 3), 0, bar);
                    }
                    continue;
                }
                // A^T: per tap of the tile, 4 k-groups (8 batch rows each) x a run of cw / 8 cores.  A bulk copy costs
                // ~60 cycles of issue in the issuing warp whatever its size, so the runs are dealt out over nine
                // warps: this one and the eight epilogue warps, which are idle until the accumulators are complete.
                for (int idx = lane; idx < NA * 4 * taps_in_tile; idx += 9 * 32) {
                    const int p = idx / (4 * taps_in_tile), r = idx % (4 * taps_in_tile);
                    const int tt = r >> 2, kg = r & 3;
                    const int apix = __ldg(tp.a_pix + (size_t)(t0 + tt) * tp.num_q + qq);
                    const size_t rg = (((size_t)apix * B + (size_t)bc * kTcBK) >> 3) + kg;
                    bulk_g2s(sA + p * A_SPLIT + kg * 2048 + tt * (Ca >> 3) * 128,
                             tp.a + p * tp.a_stride + (rg * (size_t)(Ca >> 3) + (size_t)(c0 >> 3)) * 64,
                             (uint32_t)((cw >> 3) * 128), bar);
                }
            }
            TC_PROF_T(t2c);
            TC_PROF_ADD(0, t0c, t1c);      // producer: wait for a free stage
            TC_PROF_ADD(1, t1c, t2c);      // producer: issue the copies
            TC_PROF_ADD(2, t0c - 1, t0c);  // chunk count
        }
    } else if (warp == kTlMmaWarp) {
        // ================= MMA issuer ===============================================================================
#ifdef CB200_TC_PROF
        const bool prof_on = cta0 && lane == 0;
#endif
        const uint32_t idesc = umma_instr_desc_bf16(BN, kTransA ? 1 : 0, 1);
        constexpr uint32_t A_LBO = kTransA ? 2048u : 128u, A_SBO = kTransA ? 128u : 512u;
        constexpr uint32_t A_KS = kTransA ? 2u * 2048u : 256u;                 // bytes per k16 step
        constexpr uint32_t B_LBO = (uint32_t)B_KG, B_SBO = 128u, B_KS = 2u * (uint32_t)B_KG;
        const uint64_t a_hi = umma_smem_desc(0u, A_LBO, A_SBO), b_hi = umma_smem_desc(0u, B_LBO, B_SBO);
        const uint32_t smem_base = smem_u32(smem);
        for (int j = 0; j < nchunks; ++j) {
            const int s = j % S, u = j / S;
            TC_PROF_T(t0m);
            mbar_wait(full_bar + s, (uint32_t)(u & 1));                        // the bulk copies of this stage landed
            TC_PROF_T(t1m);
            if (lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_base = smem_base + s * STAGE, b_base = a_base + NA * A_SPLIT;
#pragma unroll
                for (int ks = 0; ks < kTcBK / 16; ++ks) {
                    const uint64_t a0 = a_hi | (uint64_t)((a_base + ks * A_KS) >> 4);
                    const uint64_t a1 = a0 + (A_SPLIT >> 4), a2 = a0 + 2 * (A_SPLIT >> 4);
                    const uint32_t first = (j == 0 && ks == 0) ? 0u : 1u;
                    if (kCat) {
                        // B stage = [k-group][plane][column core]: k-group stride 3 * B_KG, column-core stride 128
                        const uint64_t bd = umma_smem_desc(b_base + ks * 2u * 3u * (uint32_t)B_KG, 3u * (uint32_t)B_KG, 128u);
                        umma_bf16(tmem_main, a0, bd, umma_instr_desc_bf16(3 * BN, 0, 1), first);   // a1 [b1|b2|b3]
                        if (NA == 3) {
                            umma_bf16(tmem_main + BN, a1, bd, umma_instr_desc_bf16(2 * BN, 0, 1), 1u);   // a2 [b1|b2]
                            umma_bf16(tmem_main + BN, a2, bd, umma_instr_desc_bf16(BN, 0, 1), 1u);       // a3 [b1]
                        }
                        continue;
                    }
                    const uint64_t b0d = b_hi | (uint64_t)((b_base + ks * B_KS) >> 4);
                    const uint64_t b1 = b0d + (B_SPLIT >> 4), b2 = b0d + 2 * (B_SPLIT >> 4);
                    umma_bf16(tmem_main, a0, b0d, idesc, first);       // a1 b1
                    umma_bf16(tmem_corr, a0, b2, idesc, first);        // a1 b3
                    if (NA == 3) {
                        umma_bf16(tmem_corr, a2, b0d, idesc, 1u);      // a3 b1
                        umma_bf16(tmem_corr, a1, b1, idesc, 1u);       // a2 b2
                    }
                    umma_bf16(tmem_corr, a0, b1, idesc, 1u);           // a1 b2
                    if (NA == 3) umma_bf16(tmem_corr, a1, b0d, idesc, 1u);   // a2 b1
                    if (bias_cta) {
                        const uint64_t od = umma_smem_desc(smem_u32(ones_tile), 2048u, 128u);
                        umma_bf16(tmem_bias_main, od, b0d, idesc, first);      // 1 g1
                        umma_bf16(tmem_bias_corr, od, b2, idesc, first);       // 1 g3
                        umma_bf16(tmem_bias_corr, od, b1, idesc, 1u);          // 1 g2
                    }
                }
                umma_commit(empty_bar + s);
                if (j == nchunks - 1) umma_commit(done_bar);
            }
            __syncwarp();
            TC_PROF_T(t2m);
            TC_PROF_ADD(3, t0m, t1m);      // MMA thread: wait for data
            TC_PROF_ADD(4, t1m, t2m);      // MMA thread: issue
        }
    } else {
        // ================= epilogue warps (TMEM lanes 0..127) =====================================================
#ifdef CB200_TC_PROF
        const bool prof_on = cta0 && tid == 0;
#endif
        if (kTransA) {
            // main loop: help the producer with the A^T bulk copies (slots 1..8 of the nine-way deal)
            const int taps_in_tile = Ca >= kTcBM ? 1 : min(kTcBM / Ca, tp.taps - m0 / Ca);
            const int t0 = m0 / Ca, cw = min(Ca, kTcBM), c0 = m0 % Ca;
            for (int j = 0; j < nchunks; ++j) {
                const int s = j % S, u = j / S;
                if (tp.a_tma != 0 || (warp + 1) * 32 >= NA * 4 * taps_in_tile) break;   // nothing dealt to this warp
                if (u > 0) mbar_wait(empty_bar + s, (uint32_t)((u - 1) & 1));
                uint8_t* sA = smem + s * STAGE;
                uint64_t* bar = full_bar + s;
                const int cj = c_lo + j;
                const int qq = cj / bc_per, bc = cj % bc_per;
                for (int idx = (warp + 1) * 32 + lane; idx < NA * 4 * taps_in_tile; idx += 9 * 32) {
                    const int p = idx / (4 * taps_in_tile), r = idx % (4 * taps_in_tile);
                    const int tt = r >> 2, kg = r & 3;
                    const int apix = __ldg(tp.a_pix + (size_t)(t0 + tt) * tp.num_q + qq);
                    const size_t rg = (((size_t)apix * B + (size_t)bc * kTcBK) >> 3) + kg;
                    bulk_g2s(sA + p * A_SPLIT + kg * 2048 + tt * (Ca >> 3) * 128,
                             tp.a + p * tp.a_stride + (rg * (size_t)(Ca >> 3) + (size_t)(c0 >> 3)) * 64,
                             (uint32_t)((cw >> 3) * 128), bar);
                }
                __syncwarp();
            }
        }
        TC_PROF_T(t0e);
        if (nchunks > 0) {
            mbar_wait(done_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        TC_PROF_T(t1e);
        {
            const int half = warp >> 2;                        // warps 0-3: low half of the columns, 4-7: high half
            tc_epilogue<BN>(ep, tmem_main, tmem_corr, nchunks > 0, m0, n0, M, m_end, N, split, NA == 1, tp.a_u8_div, -1,
                            half * (BN / 2), half * (BN / 2) + BN / 2, kCat ? tmem_main + 2 * BN : 0xffffffffu);
        }
        if (bias_cta && warp == 0) {
            // row `m_end` (= taps * Ca) of the result: lane 0 owns TMEM lane 0 of the bias accumulators
#pragma unroll 1
            for (int col = 0; col < BN; col += 16) {
                uint32_t vm[16], vc[16];
                if (nchunks > 0) {
                    CB200_TMEM_LD16(vm, tmem_bias_main + (uint32_t)col);
                    CB200_TMEM_LD16(vc, tmem_bias_corr + (uint32_t)col);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) vm[j] = vc[j] = 0u;
                }
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = n0 + col + j;
                        if (n >= N) continue;
                        const float v = __uint_as_float(vm[j]) + __uint_as_float(vc[j]);
                        if (ep.splits > 1)
                            ep.partial[((size_t)split * M + m_end) * N + n] = v;
                        else
                            epilogue_store(ep, m_end, n, v);
                    }
                }
            }
        }
        TC_PROF_T(t2e);
        TC_PROF_ADD(5, t0e, t1e);          // main loop as seen by the epilogue warps
        TC_PROF_ADD(6, t1e, t2e);          // epilogue
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_main), "r"(TMEM_COLS));
    }
}

}  // namespace gemm
}  // namespace cb200
