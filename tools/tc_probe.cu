// tools/tc_probe.cu -- accuracy / correctness probe for the tcgen05 (5th-gen tensor core) path:
//   C[M, N] = A[M, K] * B[K, N]   (fp32 in / fp32 out)
// computed as a 3-way BF16 split (x = x1 + x2 + x3, 6 cross products) with fp32 accumulation in TMEM.
// A row-major [M, K]  -> K-major operand;  B row-major [K, N] -> MN-major operand.  Operands are staged by the threads
// into the canonical no-swizzle UMMA shared-memory layouts (cute/atom/mma_traits_sm100.hpp, make_umma_desc).
// One CTA = 128 threads = one 128 x N output tile; single-stage (the probe measures correctness, not speed).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
// | layout_type [61,64) (0 = no swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// InstrDescriptor: c_format F32 (1) [4,6) | a_format BF16 (1) [7,10) | b_format BF16 (1) [10,13) | a_major [15] |
// b_major [16] | n>>3 [17,23) | m>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= 1u << 7;
    d |= 1u << 10;
    d |= (uint32_t)a_mn_major << 15;
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(BM >> 4) << 24;
    return d;
}

__device__ __forceinline__ void split3(float x, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
    h = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h);
    m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);
    l = __float2bfloat16_rn(r2);
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}

__global__ void __launch_bounds__(128) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ C, int M, int N, int K, int n_products) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // layout: [A split 0..2][B split 0..2] then mbarrier + tmem pointer
    const int a_bytes = BM * BK * 2;          // one split of the A chunk
    const int b_bytes = N * BK * 2;
    uint8_t* sA = smem;
    uint8_t* sB = smem + 3 * a_bytes;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 3 * a_bytes + 3 * b_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * BM;
    int ncols = 32;
    while (ncols < N) ncols <<= 1;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = make_idesc(N, /*a MN-major*/ 0, /*b MN-major*/ 1);

    const int nchunks = (K + BK - 1) / BK;
    uint32_t phase = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int k0 = c * BK;
        if (c > 0) {
            // previous chunk's MMAs must have finished reading shared memory
            asm volatile(
                "{\n\t.reg .pred p;\n\tWAIT_PREV:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra DONE_PREV;\n\tbra WAIT_PREV;\n\tDONE_PREV:\n\t}\n" ::"r"(smem_u32(mbar)),
                "r"(phase)
                : "memory");
            phase ^= 1;
        }
        // ---- A: thread = row, 32 k-values -> 4 k-groups of 8, 16 bytes per split per group --------------------------
        {
            const int r = tid;
            const int gm = m0 + r;
#pragma unroll
            for (int kg = 0; kg < BK / 8; ++kg) {
                __nv_bfloat16 h[8], m[8], l[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = k0 + kg * 8 + j;
                    const float x = (gm < M && k < K) ? A[(size_t)gm * K + k] : 0.f;
                    split3(x, h[j], m[j], l[j]);
                }
                const int off = kg * (BM / 8) * 128 + (r / 8) * 128 + (r % 8) * 16;
                *reinterpret_cast<uint4*>(sA + 0 * a_bytes + off) = *reinterpret_cast<uint4*>(h);
                *reinterpret_cast<uint4*>(sA + 1 * a_bytes + off) = *reinterpret_cast<uint4*>(m);
                *reinterpret_cast<uint4*>(sA + 2 * a_bytes + off) = *reinterpret_cast<uint4*>(l);
            }
        }
        // ---- B: items (k, n-group of 8) -------------------------------------------------------------------------------
        {
            const int ngroups = N / 8;
            for (int it = tid; it < BK * ngroups; it += 128) {
                const int kk = it / ngroups, ng = it % ngroups;
                const int k = k0 + kk;
                __nv_bfloat16 h[8], m[8], l[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float x = (k < K) ? B[(size_t)k * N + ng * 8 + j] : 0.f;
                    split3(x, h[j], m[j], l[j]);
                }
                const int off = (kk / 8) * ngroups * 128 + ng * 128 + (kk % 8) * 16;
                *reinterpret_cast<uint4*>(sB + 0 * b_bytes + off) = *reinterpret_cast<uint4*>(h);
                *reinterpret_cast<uint4*>(sB + 1 * b_bytes + off) = *reinterpret_cast<uint4*>(m);
                *reinterpret_cast<uint4*>(sB + 2 * b_bytes + off) = *reinterpret_cast<uint4*>(l);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_lbo = (BM / 8) * 128, a_sbo = 128;          // K-major: LBO = k-group stride
            const uint32_t b_lbo = (N / 8) * 128, b_sbo = 128;           // MN-major: LBO = k-group stride
            // products ordered small to large: (a1 b3) (a3 b1) (a2 b2) (a1 b2) (a2 b1) (a1 b1)
            const int pa[6] = {0, 2, 1, 0, 1, 0};
            const int pb[6] = {2, 0, 1, 1, 0, 0};
            for (int ks = 0; ks < BK / 16; ++ks) {
                for (int p = 6 - n_products; p < 6; ++p) {
                    const uint64_t da = make_desc(smem_u32(sA + pa[p] * a_bytes) + ks * 2 * a_lbo, a_lbo, a_sbo);
                    const uint64_t db = make_desc(smem_u32(sB + pb[p] * b_bytes) + ks * 2 * b_lbo, b_lbo, b_sbo);
                    const uint32_t accum = (c > 0 || ks > 0 || p > 6 - n_products) ? 1u : 0u;
                    mma_bf16(tmem_base, da, db, idesc, accum);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_u32(mbar))
                         : "memory");
        }
    }
    // wait for the last commit
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LAST:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_LAST;\n\tbra WAIT_LAST;\n\tDONE_LAST:\n\t}\n" ::"r"(smem_u32(mbar)),
        "r"(phase)
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w owns TMEM lanes 32w .. 32w+31 (= rows), 8 columns per load
    const int row = m0 + warp * 32 + lane;
    for (int col = 0; col < N; col += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)col;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < M) {
#pragma unroll
            for (int j = 0; j < 8; ++j) C[(size_t)row * N + col + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols));
    }
}

}  // namespace

extern "C" int tc_probe(const float* A, const float* B, float* C, int M, int N, int K, int n_products, void* stream) {
    if (N % 16 != 0 || N < 16 || N > 256 || n_products < 1 || n_products > 6) return -1;
    const size_t smem = 3 * BM * BK * 2 + 3 * (size_t)N * BK * 2 + 64;
    cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_probe_kernel<<<(M + BM - 1) / BM, 128, smem, (cudaStream_t)stream>>>(A, B, C, M, N, K, n_products);
    return (int)cudaGetLastError();
}
