// coach_b200/csrc/nn_gemm_skinny.cuh -- the skinny dense contractions of the heads (Q head: 512 -> num_actions).
//
// A [rows, lda] row-major fp32 (cb200_gemm_desc.a_lda > 0), one side of the product at most 8 wide.  A 128 x 32 tile
// kernel spends its time on 4 CTAs for these shapes; here the parallelism follows the long dimensions instead:
//   skinny_n : C[m, n] = sum_r A[m, r] B[r, n],       n <= 8          one warp per output row (forward)
//   skinny_r : C[m, n] = sum_{r <= 8} A[m, r] B[r, n]                 four outputs per thread (data gradient)
//   skinny_tn: C[k, n] = sum_m A[m, k] B[m, n],       n <= 8          32 columns x 32 row lanes per block (weight
//              gradient; the bias-gradient row sum_m B[m, n] rides along as row a_cols)
// fp32 FFMA with fixed summation orders (deterministic); same epilogue as every other path (epilogue_store).
#pragma once
#include "nn_gemm.cuh"

namespace cb200 {
namespace gemm {

__global__ void __launch_bounds__(256) skinny_n_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b,
                                                       int ldb, EpiParams ep, int M, int N, int R) {
    const int lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const float* arow = a + (size_t)m * lda;
    for (int r = 4 * lane; r < R; r += 128) {                 // R % 4 == 0, rows 16-byte aligned
        const float4 v = __ldg(reinterpret_cast<const float4*>(arow + r));
        const float av[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float* brow = b + (size_t)(r + i) * ldb;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < N) acc[j] = fmaf(av[i], __ldg(brow + j), acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < N && lane == j) epilogue_store(ep, m, j, acc[j]);
}

__global__ void __launch_bounds__(256) skinny_r_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b,
                                                       int ldb, EpiParams ep, int M, int N, int R) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (m, group of 4 columns)
    const int ng = N >> 2;
    if (g >= (int64_t)M * ng) return;
    const int m = (int)(g / ng), n = (int)(g % ng) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < R; ++r) {
        const float av = __ldg(a + (size_t)m * lda + r);
        const float4 bv = __ldg(reinterpret_cast<const float4*>(b + (size_t)r * ldb + n));
        acc.x = fmaf(av, bv.x, acc.x);
        acc.y = fmaf(av, bv.y, acc.y);
        acc.z = fmaf(av, bv.z, acc.z);
        acc.w = fmaf(av, bv.w, acc.w);
    }
    epilogue_store(ep, m, n, acc.x);
    epilogue_store(ep, m, n + 1, acc.y);
    epilogue_store(ep, m, n + 2, acc.z);
    epilogue_store(ep, m, n + 3, acc.w);
}

__global__ void __launch_bounds__(1024) skinny_tn_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b,
                                                        int ldb, EpiParams ep, int rows, int K, int N, int ones_row) {
    __shared__ float red[8][32][33];                          // [n][row lane][column]
    const int kx = threadIdx.x & 31, ml = threadIdx.x >> 5;   // 32 columns x 32 row lanes
    const int kblocks = (K + 31) / 32;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if ((int)blockIdx.x < kblocks) {
        const int k = blockIdx.x * 32 + kx;
        for (int m = ml; m < rows; m += 32) {
            const float av = k < K ? __ldg(a + (size_t)m * lda + k) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < N) acc[j] = fmaf(av, __ldg(b + (size_t)m * ldb + j), acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[j][ml][kx] = acc[j];
        __syncthreads();
        if (ml == 0 && k < K) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j >= N) continue;
                float v = red[j][0][kx];
                for (int l = 1; l < 32; ++l) v += red[j][l][kx];
                epilogue_store(ep, k, j, v);
            }
        }
    } else {
        // the bias-gradient row: thread t sums rows t, t + 256, ... of every column, lanes folded in a fixed order
        for (int m = threadIdx.x; m < rows; m += 1024) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < N) acc[j] += __ldg(b + (size_t)m * ldb + j);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[j][ml][kx] = acc[j];
        __syncthreads();
        if (threadIdx.x < N) {
            float v = 0.f;
            for (int l = 0; l < 32; ++l)
                for (int x = 0; x < 32; ++x) v += red[threadIdx.x][l][x];
            epilogue_store(ep, ones_row, threadIdx.x, v);
        }
    }
}

}  // namespace gemm
}  // namespace cb200
