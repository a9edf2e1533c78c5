// coach_b200/csrc/nn_gemm_tiled_persist.cuh -- persistent form of the multi-tap tcgen05 GEMM (nn_gemm_tiled.cuh).
//
// Same contraction, operands, shared-memory layouts and per-unit operation order (=> identical bits); what changes
// is the schedule.  In the one-tile-per-CTA kernel a tile's epilogue (TMEM -> registers -> fp32 result + three planes)
// runs with the tensor core and the TMA path idle, all co-resident CTAs reach their epilogue together (a write burst),
// and a launch of 1.1 waves pays for two.  Here ONE CTA per SM walks the work units (output tile x reduction slice)
// round-robin with
//   * the whole shared memory as operand ring (5-8 stages), filled across unit boundaries by FOUR producer warps:
//     mode 0 -- one warp per operand (A box, B box), each keeping the unit's tap-list entries in registers
//     (loaded one unit ahead, broadcast by shuffle), so no dependent global load sits in front of a TMA issue;
//     mode 1 -- the G box plus the A^T bulk copies dealt over all 128 producer lanes;
//   * TWO TMEM accumulator sets: the MMA thread starts unit i+1 in the other set while
//   * EIGHT epilogue warps (two per TMEM lane quadrant, half of the columns each) drain unit i.
// Every role looks one unit ahead for the per-unit metadata (tap-list bounds), so unit boundaries cost no round trip.
// mbarriers: full / empty per stage (producers <-> MMA), acc_full / acc_empty per accumulator set (MMA <-> epilogue).
#pragma once
#include "nn_gemm_tiled.cuh"

namespace cb200 {
namespace gemm {

constexpr int kPsMaxStages = 8;
constexpr int kPsEpiWarps = 8;                       // warps 0-7
constexpr int kPsMmaWarp = 8;
constexpr int kPsProdWarp0 = 9, kPsProdWarps = 4;    // warps 9-12
constexpr int kPsThreads = 32 * (kPsEpiWarps + 1 + kPsProdWarps);

template <int BN, int NA>
struct PersistCfg {
    static constexpr int kStageBytes = NA * kTcBM * kTcBK * 2 + 3 * BN * kTcBK * 2;
    static constexpr int kBudget = 200 * 1024;
    static constexpr int kStagesRaw = (kBudget - 256 - 4096) / kStageBytes;
    static constexpr int kStages = kStagesRaw > kPsMaxStages ? kPsMaxStages : kStagesRaw;
    static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 256 + 4096;
};

struct UnitGrid {
    int gx, gy, gz;            // tiles along M (mode 0: pixel x batch tiles), tiles along N, reduction slices
};

template <int BN, bool kTransA, int NA, bool kCat = false>
__global__ void __launch_bounds__(kPsThreads, 1) gemm_tc_tiled_persist_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                              const __grid_constant__ CUtensorMap tmB,
                                                                              TiledParams tp, EpiParams ep, int M,
                                                                              UnitGrid ug) {
    constexpr int S = PersistCfg<BN, NA>::kStages;
    constexpr int A_SPLIT = kTcBM * kTcBK * 2;
    constexpr int B_SPLIT = BN * kTcBK * 2;
    constexpr int STAGE = NA * A_SPLIT + 3 * B_SPLIT;
    constexpr int B_KG = (BN / 8) * 128;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * STAGE);     // [S]
    uint64_t* empty_bar = full_bar + kPsMaxStages;                         // [S]
    uint64_t* acc_full = empty_bar + kPsMaxStages;                         // [2]
    uint64_t* acc_empty = acc_full + 2;                                    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    uint8_t* ones_tile = smem + S * STAGE + 256;                           // 4 KB of bf16 1.0 (bias row operand)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = tp.batch, Ca = tp.a_cols, N = tp.n;
    const bool bias = kTransA && tp.bias_row != 0;
    static_assert(!kCat || (!kTransA && BN <= 64), "kCat: mode 0, BN <= 64");
    const uint32_t set_cols = kCat ? 3u * BN : (bias ? 4u : 2u) * BN;      // one accumulator set
    const uint32_t need = 2u * set_cols;
    const uint32_t TMEM_COLS = need <= 32 ? 32u : (need <= 64 ? 64u : (need <= 128 ? 128u : (need <= 256 ? 256u : 512u)));
    const int kc_per = Ca / kTcBK, bc_per = B / kTcBK;
    const int total_units = ug.gx * ug.gy * ug.gz;
    const int m_limit = M - (bias ? 1 : 0);
    const int tiles_per_q = (B + kTcBM - 1) / kTcBM;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 32) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar + s, kTransA ? 1u : 2u);   // mode 0: the A producer and the B producer both arrive
            mbar_init(empty_bar + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(acc_full + a, 1);
            mbar_init(acc_empty + a, kPsEpiWarps);        // one elected lane of each epilogue warp
        }
        fence_mbar_init();
    }
    if (bias && tid >= 64 && tid < 192) {
        uint4* o = reinterpret_cast<uint4*>(ones_tile) + 2 * (tid - 64);
        o[0] = o[1] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        fence_proxy_async_smem();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // unit -> (tile along M, tile along N, slice); every role walks the same sequence
    auto decode = [&](int u, int& ux, int& n0, int& split) {
        ux = u % ug.gx;
        n0 = ((u / ug.gx) % ug.gy) * BN;
        split = u / (ug.gx * ug.gy);
    };
    // tap-list bounds of a unit's output pixel (mode 0): the only per-unit metadata that lives in global memory;
    // each role loads the pair of the NEXT unit while it works on the current one
    auto load_lp = [&](int u) -> int2 {
        if (kTransA || u >= total_units) return make_int2(0, 0);
        const int q = (u % ug.gx) / tiles_per_q;
        return make_int2(__ldg(tp.list_ptr + q), __ldg(tp.list_ptr + q + 1));
    };
    // reduction chunks [c_lo, c_lo + nchunks) of a unit
    auto chunk_range = [&](int2 lp, int split, int& c_lo) -> int {
        const int total = kTransA ? tp.num_q * bc_per : (lp.y - lp.x) * kc_per;
        c_lo = split * tp.chunks_per_split;
        return max(0, min(total, c_lo + tp.chunks_per_split) - c_lo);
    };

    if (warp >= kPsProdWarp0) {
        // ================= producers ================================================================================
        const int pw = warp - kPsProdWarp0;                       // 0 .. 3
        if (!kTransA && pw >= 2) {
            // mode 0 needs two producer warps only
        } else {
            if (lane == 0) {
                if (kTransA || pw == 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
                if (!kTransA && pw == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            }
            uint32_t g = 0;                                        // chunks issued so far (stage ring position)
            // mode 0 metadata pipeline: list bounds two units ahead, the unit's list entries one unit ahead
            int2 lp_cur = load_lp(blockIdx.x);
            int2 lp_nxt = load_lp(blockIdx.x + gridDim.x);
            auto load_ents = [&](int u, int2 lp) -> int2 {
                if (kTransA || u >= total_units) return make_int2(0, 0);
                int ux, n0, split, c_lo;
                decode(u, ux, n0, split);
                const int nch = chunk_range(lp, split, c_lo);
                const int e_lo = c_lo / kc_per, e_hi = nch > 0 ? (c_lo + nch - 1) / kc_per : e_lo - 1;
                return (e_lo + lane <= e_hi) ? __ldg(tp.list + lp.x + e_lo + lane) : make_int2(0, 0);
            };
            int2 ent_cur = load_ents(blockIdx.x, lp_cur);
            for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                int ux, n0, split, c_lo;
                decode(u, ux, n0, split);
                const int2 lp = lp_cur;
                const int nchunks = chunk_range(lp, split, c_lo);
                const int2 ent = ent_cur;
                // look ahead: entries of the next unit (its bounds arrived an iteration ago), bounds of the one after
                ent_cur = load_ents(u + gridDim.x, lp_nxt);
                lp_cur = lp_nxt;
                lp_nxt = load_lp(u + 2 * gridDim.x);
                if (!kTransA) {
                    const int b0 = (ux % tiles_per_q) * kTcBM;
                    const int e_lo = c_lo / kc_per;
                    for (int j = 0; j < nchunks; ++j, ++g) {
                        const uint32_t s = g % S, use = g / S;
                        if (use > 0) mbar_wait(empty_bar + s, (use - 1) & 1u);
                        const int cj = c_lo + j;
                        const int e = cj / kc_per - e_lo, kc = cj % kc_per;
                        const int ex = __shfl_sync(0xffffffffu, ent.x, e), ey = __shfl_sync(0xffffffffu, ent.y, e);
                        if (lane == 0) {
                            uint8_t* sA = smem + s * STAGE;
                            uint64_t* bar = full_bar + s;
                            if (pw == 0) {
                                mbar_expect_tx(bar, (uint32_t)(NA * A_SPLIT));
                                tma_load_4d(sA, &tmA, 0, kc * 4, (int)(((size_t)ex * B + b0) >> 3), 0, bar);
                            } else {
                                mbar_expect_tx(bar, 3u * (uint32_t)B_SPLIT);
                                if (kCat)
                                    tma_load_4d(sA + NA * A_SPLIT, &tmB, 0, n0 >> 3, 0, (ey * Ca + kc * kTcBK) >> 3, bar);
                                else
                                    tma_load_4d(sA + NA * A_SPLIT, &tmB, 0, n0 >> 3, (ey * Ca + kc * kTcBK) >> 3, 0, bar);
                            }
                        }
                        __syncwarp();
                    }
                } else {
                    const int m0 = ux * kTcBM;
                    const int taps_in_tile = Ca >= kTcBM ? 1 : min(kTcBM / Ca, tp.taps - m0 / Ca);
                    const int t0 = m0 / Ca;
                    const int cw = min(Ca, kTcBM);
                    const int c0 = m0 % Ca;
                    const uint32_t a_bytes = (uint32_t)(taps_in_tile * 4 * (cw / 8) * 128);
                    const uint32_t tx_bytes = (uint32_t)NA * a_bytes + 3u * (uint32_t)B_SPLIT;
                    const int ncopies = NA * 4 * taps_in_tile;
                    if (pw * 32 >= ncopies && pw > 0) {            // nothing dealt to this warp in this unit
                        g += (uint32_t)nchunks;
                        continue;
                    }
                    for (int j = 0; j < nchunks; ++j, ++g) {
                        const uint32_t s = g % S, use = g / S;
                        if (use > 0) mbar_wait(empty_bar + s, (use - 1) & 1u);
                        uint8_t* sA = smem + s * STAGE;
                        uint8_t* sB = sA + NA * A_SPLIT;
                        uint64_t* bar = full_bar + s;
                        const int cj = c_lo + j;
                        const int qq = cj / bc_per, bc = cj % bc_per;
                        if (pw == 0 && lane == 0) {
                            mbar_expect_tx(bar, tx_bytes);
                            tma_load_4d(sB, &tmB, 0, n0 >> 3, (int)(((size_t)qq * B + (size_t)bc * kTcBK) >> 3), 0, bar);
                        }
                        for (int idx = pw * 32 + lane; idx < ncopies; idx += kPsProdWarps * 32) {
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
            }
        }
    } else if (warp == kPsMmaWarp) {
        // ================= MMA issuer ===============================================================================
        const uint32_t idesc = umma_instr_desc_bf16(BN, kTransA ? 1 : 0, 1);
        constexpr uint32_t A_LBO = kTransA ? 2048u : 128u, A_SBO = kTransA ? 128u : 512u;
        constexpr uint32_t A_KS = kTransA ? 2u * 2048u : 256u;
        constexpr uint32_t B_LBO = (uint32_t)B_KG, B_SBO = 128u, B_KS = 2u * (uint32_t)B_KG;
        const uint64_t a_hi = umma_smem_desc(0u, A_LBO, A_SBO), b_hi = umma_smem_desc(0u, B_LBO, B_SBO);
        const uint64_t ones_desc = umma_smem_desc(smem_u32(ones_tile), 2048u, 128u);
        const uint32_t smem_base = smem_u32(smem);
        uint32_t g = 0;
        int i = 0;                                             // units done by this CTA
        int2 lp_nxt = load_lp(blockIdx.x);
        for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++i) {
            int ux, n0, split, c_lo;
            decode(u, ux, n0, split);
            const int2 lp = lp_nxt;
            lp_nxt = load_lp(u + gridDim.x);
            const int nchunks = chunk_range(lp, split, c_lo);
            const int set = i & 1;
            const uint32_t k = (uint32_t)(i >> 1);             // k-th use of this accumulator set
            if (k > 0) mbar_wait(acc_empty + set, (k - 1) & 1u);   // the epilogue of its previous unit is done
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_main = tmem_base + set * set_cols, tmem_corr = tmem_main + BN;
            const bool bias_unit = bias && ux == 0;
            for (int j = 0; j < nchunks; ++j, ++g) {
                const uint32_t s = g % S, use = g / S;
                mbar_wait(full_bar + s, use & 1u);
                if (lane == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_base = smem_base + s * STAGE, b_base = a_base + NA * A_SPLIT;
#pragma unroll
                    for (int ks = 0; ks < kTcBK / 16; ++ks) {
                        const uint64_t a0 = a_hi | (uint64_t)((a_base + ks * A_KS) >> 4);
                        const uint64_t a1 = a0 + (A_SPLIT >> 4), a2 = a0 + 2 * (A_SPLIT >> 4);
                        const uint32_t first = (j == 0 && ks == 0) ? 0u : 1u;
                        if (kCat) {      // three wide MMAs on the [b1|b2|b3] operand (nn_gemm_tiled.cuh)
                            const uint64_t bd = umma_smem_desc(b_base + ks * 2u * 3u * (uint32_t)B_KG, 3u * (uint32_t)B_KG, 128u);
                            umma_bf16(tmem_main, a0, bd, umma_instr_desc_bf16(3 * BN, 0, 1), first);
                            if (NA == 3) {
                                umma_bf16(tmem_main + BN, a1, bd, umma_instr_desc_bf16(2 * BN, 0, 1), 1u);
                                umma_bf16(tmem_main + BN, a2, bd, umma_instr_desc_bf16(BN, 0, 1), 1u);
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
                        if (bias_unit) {
                            umma_bf16(tmem_main + 2 * BN, ones_desc, b0d, idesc, first);   // 1 g1
                            umma_bf16(tmem_main + 3 * BN, ones_desc, b2, idesc, first);    // 1 g3
                            umma_bf16(tmem_main + 3 * BN, ones_desc, b1, idesc, 1u);       // 1 g2
                        }
                    }
                    umma_commit(empty_bar + s);
                    if (j == nchunks - 1) umma_commit(acc_full + set);
                }
                __syncwarp();
            }
            if (nchunks == 0 && lane == 0) {
                // nothing to accumulate: hand the (unused) set to the epilogue, which writes zeros
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc_full + set)) : "memory");
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue: warps 0-3 take the low half of the columns, warps 4-7 the high half ===========
        const int half = warp >> 2;
        const int col_lo = BN >= 32 ? half * (BN / 2) : 0, col_hi = BN >= 32 ? col_lo + BN / 2 : (half == 0 ? BN : 0);
        int i = 0;
        int2 lp_nxt = load_lp(blockIdx.x);
        for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++i) {
            int ux, n0, split, c_lo;
            decode(u, ux, n0, split);
            const int2 lp = lp_nxt;
            lp_nxt = load_lp(u + gridDim.x);
            const int nchunks = chunk_range(lp, split, c_lo);
            int m0, m_end;
            if (!kTransA) {
                const int q = ux / tiles_per_q;
                m0 = q * B + (ux % tiles_per_q) * kTcBM;
                m_end = q * B + B;
            } else {
                m0 = ux * kTcBM;
                m_end = m_limit;
            }
            const int set = i & 1;
            const uint32_t k = (uint32_t)(i >> 1);
            mbar_wait(acc_full + set, k & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_main = tmem_base + set * set_cols, tmem_corr = tmem_main + BN;
            tc_epilogue<BN>(ep, tmem_main, tmem_corr, nchunks > 0, m0, n0, M, m_end, N, split, NA == 1, tp.a_u8_div, -1,
                            col_lo, col_hi, kCat ? tmem_main + 2 * BN : 0xffffffffu);
            if (bias && ux == 0 && warp == 0) {
                // row m_limit (= taps * Ca) of the result: lane 0 owns TMEM lane 0 of the bias accumulators
#pragma unroll 1
                for (int col = 0; col < BN; col += 16) {
                    uint32_t vm[16], vc[16];
                    if (nchunks > 0) {
                        CB200_TMEM_LD16(vm, tmem_main + 2 * BN + (uint32_t)col);
                        CB200_TMEM_LD16(vc, tmem_main + 3 * BN + (uint32_t)col);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) vm[jj] = vc[jj] = 0u;
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) {
                            const int n = n0 + col + jj;
                            if (n >= N) continue;
                            const float v = __uint_as_float(vm[jj]) + __uint_as_float(vc[jj]);
                            if (ep.splits > 1)
                                ep.partial[((size_t)split * M + m_limit) * N + n] = v;
                            else
                                epilogue_store(ep, m_limit, n, v);
                        }
                    }
                }
            }
            // all TMEM reads of this warp are complete (tcgen05.wait::ld inside): release the accumulator set
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc_empty + set)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

}  // namespace gemm
}  // namespace cb200
