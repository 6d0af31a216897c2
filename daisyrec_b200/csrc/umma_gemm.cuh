// umma_gemm.cuh -- bf16 tensor-core GEMM for the NeuMF tower on sm_100a: tcgen05.mma with the accumulator in TMEM.
//
//   C[M,N] (op)= opA(A)[M,K] * opB(B)[K,N]       A, B, C fp32 in global memory; operands are rounded to bf16
//   while they are staged into shared memory, products accumulate in fp32 in tensor memory.
//
// Same call signature and epilogues as the fp32 CUDA-core sgemm_kernel in neumf.cu, so the three GEMM call sites of
// the tower (forward NT + bias + ReLU, input-gradient NN + ReLU mask, weight-gradient TN split-K) switch by dtype.
//
// One CTA (128 threads) owns a 128-row tile of C and the full N (<= 256):
//   * TMEM: `cols` columns (power of two >= 32) x 128 lanes hold the fp32 accumulator (tcgen05.alloc by warp 0);
//   * per 32-deep K chunk all threads stage A[128x32] and B[Npad x 32] into shared memory in the canonical
//     no-swizzle core-matrix layouts (8 x 16-byte core matrices; K-major for operands that are contiguous along K,
//     MN-major for the transposed operands of the backward GEMMs; LBO / SBO padded so the 16-byte staging stores are
//     bank-conflict free), fence.proxy.async, then ONE thread issues two tcgen05.mma (K = 16 each, M = 128,
//     N = Npad) and tcgen05.commit's an mbarrier that releases the buffers;
//   * epilogue: warp w reads TMEM lanes [32w, 32w+32) with tcgen05.ld.32x32b.x32 (lane == output row), transposes the
//     32x32 block through shared memory and writes row-contiguous fp32 with bias / ReLU / mask fused, or atomically
//     accumulates (optionally transposed) for the split-K weight gradient.
// The tower GEMMs are skinny (N, K <= 128 against M ~ 10^6): they are bound by streaming A from HBM, not by the
// tensor pipe, so the kernel relies on several resident CTAs per SM for overlap rather than on an intra-CTA pipeline.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace drb {

constexpr int kUmmaBK = 32;        // K elements staged per chunk (2 MMAs of K=16)
constexpr int kUmmaMaxN = 256;

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    // cute::UMMA::SmemDescriptor: start [0,14) >>4, LBO [16,30) >>4, SBO [32,46) >>4, version [46,48) = 1,
    // base_offset [49,52) = 0, lbo_mode [52] = 0, layout_type [61,64) = 0 (SWIZZLE_NONE / interleave)
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ uint32_t umma_idesc_bf16_f32(int M, int N, bool a_mn = false, bool b_mn = false)
{
    // cute::UMMA::InstrDescriptor: c_format [4,6) = 1 (F32), a_format [7,10) = 1 (BF16), b_format [10,13) = 1 (BF16),
    // a_major [15] = 0, b_major [16] = 0 (K-major), n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}

// Operand tiles in shared memory (no swizzle, 8 x 16-byte core matrices), padded so that the 16-byte staging stores of a
// quarter warp fall into distinct bank groups:
//   K-major  (source contiguous along K):   elem(r, k) at (k/8)*LBO + (r/8)*128 + (r%8)*16 + (k%8)*2,  LBO = rows/8*128 + 32, SBO = 128
//   MN-major (source contiguous along rows): elem(r, k) at (k/8)*LBO + (r/8)*144 + (k%8)*16 + (r%8)*2,  LBO = rows/8*144,      SBO = 144
// Either way one work item converts 8 consecutive source floats to bf16 and issues ONE 16-byte shared store.
__host__ __device__ __forceinline__ uint32_t umma_lbo(bool mn_major, int rows)
{
    return mn_major ? (uint32_t)(rows / 8) * 144u : (uint32_t)(rows / 8) * 128u + 32u;
}
__host__ __device__ __forceinline__ uint32_t umma_sbo(bool mn_major) { return mn_major ? 144u : 128u; }

//   MN = false: src(r, k) = S[(r0 + r) * ld + k]       MN = true: src(r, k) = S[k * ld + (r0 + r)]
template <bool MN>
__device__ __forceinline__ void umma_stage_tile(unsigned char *smem, int rows, const float *__restrict__ S, long long ld,
                                                long long r0, long long r_lim, int k0, int k_lim, int tid, int nthreads)
{
    const uint32_t lbo = umma_lbo(MN, rows);
    const int items = MN ? (rows / 8) * kUmmaBK : rows * (kUmmaBK / 8);
    for (int it = tid; it < items; it += nthreads) {
        float v[8];
        uint32_t off;
        if (!MN) {
            const int r = it / (kUmmaBK / 8), c1 = it % (kUmmaBK / 8);    // 4 lanes read 128 contiguous bytes of a row
            const long long gr = r0 + r;
            const int k = k0 + c1 * 8;
            if (gr < r_lim && k + 8 <= k_lim && ((ld & 3) == 0)) {
                const float4 *p = reinterpret_cast<const float4 *>(S + gr * ld + k);
                float4 a = __ldg(p), b = __ldg(p + 1);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = (gr < r_lim && k + e < k_lim) ? __ldg(S + gr * ld + k + e) : 0.f;
            }
            off = (uint32_t)c1 * lbo + (uint32_t)(r / 8) * 128u + (uint32_t)(r % 8) * 16u;
        } else {
            const int rg = it % (rows / 8), c = it / (rows / 8);          // consecutive lanes read consecutive row groups
            const long long gr = r0 + (long long)rg * 8;
            const int k = k0 + c;
            if (k < k_lim && gr + 8 <= r_lim && ((ld & 3) == 0) && ((gr & 3) == 0)) {
                const float4 *p = reinterpret_cast<const float4 *>(S + (long long)k * ld + gr);
                float4 a = __ldg(p), b = __ldg(p + 1);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = (k < k_lim && gr + e < r_lim) ? __ldg(S + (long long)k * ld + gr + e) : 0.f;
            }
            off = (uint32_t)(c / 8) * lbo + (uint32_t)rg * 144u + (uint32_t)(c % 8) * 16u;
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4 *>(smem + off) = o;
    }
}

// EPI 0: C = acc   1: C = relu(acc + bias[n])   2: C = acc * (ref(m,n) > 0)   3: atomicAdd(C, acc) (split-K over grid.z)
// EPI 4: atomicAdd(C^T, acc) -- transposed accumulate C[n*ldc + m] (split-K)
template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(128) umma_gemm_kernel(int M, int N, int K, const float *__restrict__ A, long long lda,
                                                        const float *__restrict__ B, long long ldb, float *__restrict__ C,
                                                        long long ldc, const float *__restrict__ bias,
                                                        const float *__restrict__ ref, long long ldref, int k_chunk, int Npad,
                                                        int tmem_cols, float alpha)
{
    constexpr int kABytes = (kUmmaBK / 8) * (128 / 8) * 144;                 // worst case (MN-major) A tile
    constexpr int kBBytes = (kUmmaBK / 8) * (kUmmaMaxN / 8) * 144;
    __shared__ __align__(128) unsigned char s_all[kABytes + kBBytes];         // A tile | B tile; reused by the epilogue
    unsigned char *sA = s_all, *sB = s_all + kABytes;
    // A(m,k) = A[k*lda + m] (TA) and B(k,n) = B[k*ldb + n] (!TB) are contiguous along the MN dimension -> MN-major tiles
    constexpr bool A_MN = TA, B_MN = !TB;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m0 = (long long)blockIdx.x * 128;
    const int kb = (EPI >= 3) ? blockIdx.z * k_chunk : 0;
    const int ke = (EPI >= 3) ? min(K, kb + k_chunk) : K;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc_bf16_f32(128, Npad, A_MN, B_MN);
    const uint32_t lboA = umma_lbo(A_MN, 128), lboB = umma_lbo(B_MN, Npad);
    uint32_t phase = 0;
    bool first = true;
    for (int k0 = kb; k0 < ke; k0 += kUmmaBK) {
        umma_stage_tile<A_MN>(sA, 128, A, lda, m0, M, k0, ke, tid, 128);
        umma_stage_tile<B_MN>(sB, Npad, B, ldb, 0, N, k0, ke, tid, 128);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < kUmmaBK / 16; ++kk) {
                uint64_t da = umma_smem_desc(smem_u32(sA) + kk * 2 * lboA, lboA, umma_sbo(A_MN));
                uint64_t db = umma_smem_desc(smem_u32(sB) + kk * 2 * lboB, lboB, umma_sbo(B_MN));
                uint32_t acc = (first && kk == 0) ? 0u : 1u;
                asm volatile(
                    "{\n\t"
                    ".reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                    "}\n" ::"r"(tmem),
                    "l"(da), "l"(db), "r"(idesc), "r"(acc)
                    : "memory");
            }
            // commit: arrives on the mbarrier once every MMA issued so far has finished reading smem / writing TMEM
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar))
                         : "memory");
        }
        first = false;
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        tc_fence_after();
    }
    // epilogue: lane == output row inside this warp's 32-lane quarter of TMEM.  32 columns at a time are pulled out with
    // one tcgen05.ld.32x32b.x32, transposed through a padded shared-memory tile (the operand buffers are free now) and
    // written / accumulated with row-contiguous, fully coalesced accesses (lane == column).
    float *tile = reinterpret_cast<float *>(s_all) + warp * (32 * 33);
    const long long mrow0 = m0 + warp * 32;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const bool any = kb < ke;
    for (int c = 0; c < Npad; c += 32) {
        uint32_t r[32];
        if (c + 32 <= Npad) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr + (uint32_t)c));
        } else {   // Npad is a multiple of 16: a trailing half chunk
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr + (uint32_t)c));
#pragma unroll
            for (int e = 16; e < 32; ++e) r[e] = 0u;
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 32; ++e) tile[lane * 33 + e] = __uint_as_float(r[e]);
        __syncwarp();
        const int n = c + lane;
        if (any && n < N) {
            const float bn = (EPI == 1) ? bias[n] : 0.f;
            for (int rr = 0; rr < 32; ++rr) {
                const long long m = mrow0 + rr;
                if (m >= M) break;
                float v = tile[rr * 33 + lane];
                if (EPI == 1) { v += bn; v = v > 0.f ? v : 0.f; }
                if (EPI == 2) { v = (ref[m * ldref + n] > 0.f) ? v * alpha : 0.f; }
                if (EPI == 3) atomicAdd(C + m * ldc + n, v);
                else if (EPI == 4) atomicAdd(C + (long long)n * ldc + m, v);
                else C[m * ldc + n] = v;
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
    }
}

template <bool TA, bool TB, int EPI>
static int launch_umma_gemm(long long M, int N, int K, const float *A, long long lda, const float *B, long long ldb, float *C,
                            long long ldc, const float *bias, const float *ref, long long ldref, cudaStream_t st,
                            float alpha = 1.f)
{
    if (M <= 0 || N <= 0 || K <= 0) return DRB_OK;
    DRB_REQUIRE(N <= kUmmaMaxN, "umma_gemm: N=%d exceeds %d", N, kUmmaMaxN);
    int Npad = (N + 15) / 16 * 16;
    int cols = 32;
    while (cols < Npad) cols <<= 1;
    dim3 grid((unsigned)((M + 127) / 128), 1, 1);
    int k_chunk = K;
    if (EPI >= 3) {   // split-K: the [out x in] result is one tile, parallelism comes from the K (row) dimension.
        // Each CTA runs its K steps back to back (stage -> mma -> wait), so latency is hidden by CTA count: aim at
        // ~8 resident CTAs per SM, but keep >= 512 rows per chunk so the atomic epilogue stays a small fraction.
        long long want = (long long)sm_count() * 8;
        long long max_chunks = (K + 511) / 512;
        long long chunks = want < max_chunks ? want : max_chunks;
        if (chunks < 1) chunks = 1;
        k_chunk = (int)(((K + chunks - 1) / chunks + kUmmaBK - 1) / kUmmaBK * kUmmaBK);
        grid.z = (unsigned)((K + k_chunk - 1) / k_chunk);
    }
    umma_gemm_kernel<TA, TB, EPI><<<grid, 128, 0, st>>>((int)M, N, K, A, lda, B, ldb, C, ldc, bias, ref, ldref, k_chunk, Npad, cols,
                                                        alpha);
    DRB_CUDA(cudaGetLastError());
    return DRB_OK;
}

}  // namespace drb
