// coach_b200/csrc/nn_gemm_fast.cuh -- vectorised variant of the gather-GEMM (see nn_gemm.cuh for the operand model).
//
// Same contraction, same tables, same epilogue; usable when the operands can be moved 4 elements at a time:
//   * A: every aligned group of 4 consecutive column indices r..r+3 is contiguous in memory (a_coloff[r+i] =
//     a_coloff[r]+i), a_cols % 4 == 0, every a_rowoff is a multiple of 4 (fp32: 16-byte aligned float4; uint8: one
//     32-bit word).  True for NHWC im2col with C % 4 == 0, dense matrices with K % 4 == 0 and the transposed-conv
//     gather (groups never straddle a tap because N % 4 == 0).
//   * B / C: n % 4 == 0 and 16-byte aligned rows.
// Thread tile 8x8 (two 4-wide fragments half a tile apart in each direction), BK = 16, double-buffered shared tiles,
// register prefetch of the next chunk.  8x8 is what balances the shared-memory pipe (4 LDS.128 per k) against the
// FMA pipe (64 FFMA per k) on an SM.
#pragma once
#include "nn_gemm.cuh"

namespace cb200 {
namespace gemm {

template <int BM_, int BN_>
struct FastCfg {
    static constexpr int BM = BM_, BN = BN_, BK = 16, TM = 8, TN = 8;
    static constexpr int TX = BN / 8, TY = BM / 8, T = TX * TY;
    static constexpr int GA = BM * 4 / T;     // float4 groups of the A tile per thread
    static constexpr int GB = BN * 4 / T;     // float4 groups of the B tile per thread
    static_assert(BM * 4 % T == 0 && BN * 4 % T == 0, "loader shape");
};

struct FastA {
    const void* src;
    const float* lut;
    const int32_t* rowoff;
    const int32_t* coloff;
    const int32_t* rowinfo;
    const int32_t* colinfo;
    int oh, ow;
    int rows, cols;
    int ones_col;      // transposed mode only: logical column index `cols` yields 1.0 (bias-gradient row), else -1
};

__device__ __forceinline__ float4 load4(const FastA& a, int off) {
    if (a.lut) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(a.src) + off));
        return make_float4(a.lut[w & 255], a.lut[(w >> 8) & 255], a.lut[(w >> 16) & 255], a.lut[w >> 24]);
    }
    return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(a.src) + off));
}

__device__ __forceinline__ bool tap_valid(const FastA& a, int m, int r) {
    if (!a.rowinfo) return true;
    const int ri = __ldg(a.rowinfo + m), ci = __ldg(a.colinfo + r);
    const int y = (ri >> 16) - (ci >> 16), x = (ri & 0xffff) - (ci & 0xffff);
    return y >= 0 && y < a.oh && x >= 0 && x < a.ow;
}

template <class C, bool kTransA>
__global__ void __launch_bounds__(C::T) gemm_fast_kernel(FastA a, const float* __restrict__ b, int ldb, EpiParams ep,
                                                         int M, int N, int R, int r_per_split) {
    __shared__ __align__(16) float As[2][C::BK * C::BM];
    __shared__ __align__(16) float Bs[2][C::BK * C::BN];
    __shared__ float lut_s[256];
    const int tid = threadIdx.x;
    const int tx = tid % C::TX, ty = tid / C::TX;
    const int m0 = blockIdx.x * C::BM, n0 = blockIdx.y * C::BN;
    const int split = blockIdx.z;
    const int r_lo = split * r_per_split;
    const int r_hi = min(R, r_lo + r_per_split);
    if (a.lut) {
        for (int i = tid; i < 256; i += C::T) lut_s[i] = a.lut[i];
        __syncthreads();
        a.lut = lut_s;
    }

    // ---- per-thread constant parts of the A addressing ----------------------------------------------------------
    // non-transposed: group g -> tile row mm = g % BM, k-quad kq = g / BM        (lanes walk rows: conflict-free STS)
    // transposed    : group g -> row-quad  mq = g % (BM/4), reduction kk = g / (BM/4)   (lanes walk memory: coalesced)
    int fixoff[C::GA];     // rowoff[m] (non-transposed) or coloff[k] (transposed) of the group's fixed coordinate
#pragma unroll
    for (int i = 0; i < C::GA; ++i) {
        const int g = tid + i * C::T;
        if (!kTransA) {
            const int m = m0 + g % C::BM;
            fixoff[i] = (m < a.rows) ? __ldg(a.rowoff + m) : -1;
        } else {
            const int k = m0 + 4 * (g % (C::BM / 4));
            fixoff[i] = (k < a.cols) ? __ldg(a.coloff + k) : (k == a.ones_col ? -2 : -1);
        }
    }

    float4 ra[C::GA], rb[C::GB];
    auto fetch = [&](int r0) {
#pragma unroll
        for (int i = 0; i < C::GA; ++i) {
            const int g = tid + i * C::T;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!kTransA) {
                const int r = r0 + 4 * (g / C::BM);
                if (fixoff[i] >= 0 && r < r_hi) {
                    const int m = m0 + g % C::BM;
                    if (tap_valid(a, m, r)) v = load4(a, fixoff[i] + __ldg(a.coloff + r));
                }
            } else {
                const int mrow = r0 + g / (C::BM / 4);          // reduction index = logical A row
                if (mrow < r_hi) {
                    if (fixoff[i] >= 0) {
                        const int k = m0 + 4 * (g % (C::BM / 4));
                        if (tap_valid(a, mrow, k)) v = load4(a, __ldg(a.rowoff + mrow) + fixoff[i]);
                    } else if (fixoff[i] == -2) {
                        v.x = 1.f;                               // bias-gradient row: sum_m 1 * dY[m, n]
                    }
                }
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < C::GB; ++i) {
            const int g = tid + i * C::T;
            const int nq = g % (C::BN / 4), kk = g / (C::BN / 4);
            const int r = r0 + kk, n = n0 + 4 * nq;
            rb[i] = (r < r_hi && n < N) ? __ldg(reinterpret_cast<const float4*>(b + (size_t)r * ldb + n))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto commit = [&](float* As_, float* Bs_) {
#pragma unroll
        for (int i = 0; i < C::GA; ++i) {
            const int g = tid + i * C::T;
            if (!kTransA) {
                const int mm = g % C::BM, kq = g / C::BM;
                As_[(4 * kq + 0) * C::BM + mm] = ra[i].x;
                As_[(4 * kq + 1) * C::BM + mm] = ra[i].y;
                As_[(4 * kq + 2) * C::BM + mm] = ra[i].z;
                As_[(4 * kq + 3) * C::BM + mm] = ra[i].w;
            } else {
                const int mq = g % (C::BM / 4), kk = g / (C::BM / 4);
                *reinterpret_cast<float4*>(As_ + kk * C::BM + 4 * mq) = ra[i];
            }
        }
#pragma unroll
        for (int i = 0; i < C::GB; ++i) {
            const int g = tid + i * C::T;
            *reinterpret_cast<float4*>(Bs_ + (g / (C::BN / 4)) * C::BN + 4 * (g % (C::BN / 4))) = rb[i];
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nchunks = (r_hi - r_lo + C::BK - 1) / C::BK;
    if (nchunks > 0) {
        fetch(r_lo);
        commit(As[0], Bs[0]);
    }
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        const int cur = c & 1;
        if (c + 1 < nchunks) fetch(r_lo + (c + 1) * C::BK);
        mma_chunk<C>(As[cur], Bs[cur], acc, ty, tx);
        if (c + 1 < nchunks) commit(As[cur ^ 1], Bs[cur ^ 1]);
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + frag_index<C::BM, 8>(ty, i);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + frag_index<C::BN, 8>(tx, j);
            if (n >= N) continue;
            if (ep.splits > 1)
                ep.partial[((size_t)split * M + m) * N + n] = acc[i][j];
            else
                epilogue_store(ep, m, n, acc[i][j]);
        }
    }
}

}  // namespace gemm
}  // namespace cb200
