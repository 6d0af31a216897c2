// neumf_fused.cuh -- the whole NeuMF tower step of a 64-triple tile inside ONE CTA: activations never leave the SM.
//
// Stands behind NeuMF.forward / calc_loss / backward (daisy/model/NeuMFRecommender.py:118-169) for model_name 'NeuMF',
// num_layers = 2, factors = 32, dropout 0 (BASELINE config 3: F = 32, tower 128 -> 64 -> 32).  The layer-wise
// path (neumf.cu: gather, 2 forward GEMMs, head, 4 backward GEMMs, 2 column sums, scatter) streams fp32 activations through
// HBM between ~12 launches (profiles/r01c: 10 % of the HBM roofline).  Here one persistent CTA per SM walks tiles of 64
// triples = 128 rows (rows 0..63 the pos items, 64..127 the neg items of the same triples):
//
//   gather   A0 = cat(UM[u], IM[item]) rounded to bf16 straight into the K-major core-matrix image the tensor core reads
//            (lane group per row, 128-bit loads; the user row is loaded once and stored for both of its tile rows)
//   MMA      Z1 = A0 W1^T  -> TMEM            tcgen05.mma kind::f16, M = 128, fp32 accumulate
//   epilogue A1 = relu(Z1 + b1) -> bf16 image (tcgen05.ld, thread = tile row)
//   MMA      Z2 = A1 W2^T  -> TMEM
//   head     h = relu(Z2 + b2);  pred = wp . cat(UG[u] * IG[item], h) + bp;  x = pred_pos - pred_neg (rows r and r + 64 meet
//            through shared memory);  c = BPR coefficient;  loss / regulariser norms;  GMF-table gradients by RED.128;
//            dZ2 = +-c wp_h [h > 0] -> bf16 image
//   MMA      dA1 = dZ2 W2  -> TMEM;   gW2^T += A1^T dZ2  -> TMEM (accumulated over ALL tiles of the CTA)
//   epilogue dZ1 = dA1 [A1 > 0] -> bf16 image
//   MMA      dA0 = dZ1 W1  -> TMEM;   gW1^T += A0^T dZ1  -> TMEM (accumulated over all tiles)
//   epilogue dA0 -> RED.128 into gUM[u] / gIM[item]
//   finally  the two weight gradients leave TMEM once per CTA; bias / predict-layer gradients, loss and norms are carried in
//            registers across tiles and reduced once.
// Every activation / gradient tile is written ONCE as a K-major operand image (element (row, k) at
// (k/8) LBO + (row/8) 128 + (row%8) 16 + (k%8) 2).  The same bytes are the MN-major image of the TRANSPOSED tile when the
// descriptor's two strides are swapped (K-group stride 128, MN-group stride LBO), which is how A^T dZ and dZ W are fed
// without a second copy.  HBM traffic per triple: the gather and scatter of the 96-float rows (2.3 KB).
#pragma once
#include "umma_gemm.cuh"

namespace drb {

constexpr int kFusedThreads = 512;      // 16 warps: TMEM lane quarter = warp % 4, column quarter = warp / 4
constexpr int kFusedTile = 64;          // triples per tile (128 rows)

struct FusedParams {
    const float *UG, *IG, *UM, *IM;     // tables
    const float *W;                     // tower block: W1 [N1, N0], b1 [N1], W2 [N2, N1], b2 [N2], wp [2F], bp
    const int32_t *bu, *bi, *bj;
    long long B;                        // triples in this step
    float *gUG, *gIG, *gUM, *gIM, *gW;  // gradient accumulators (table-shaped; gW like W)
    unsigned *cntU;
    unsigned long long *cntI;
    double *red;                        // [11] bpr, l1[5], s2[5]  (UG_u, UM_u, IG_i, IM_i, IG_j)
    int has_reg, apply;
};

__host__ __device__ constexpr uint32_t fused_lbo(int rows) { return (uint32_t)(rows / 8) * 128u + 32u; }

template <int F>
struct FusedLayout {
    static constexpr int D = 2 * F, N0 = 4 * F, N1 = 2 * F, N2 = F;
    static constexpr uint32_t LBO_T = fused_lbo(128);                 // images with 128 tile rows
    static constexpr uint32_t LBO_W1 = fused_lbo(N1), LBO_W2 = fused_lbo(N2);
    static constexpr uint32_t A0 = 0;
    static constexpr uint32_t A1 = A0 + (N0 / 8) * LBO_T;
    static constexpr uint32_t DZ1 = A1 + (N1 / 8) * LBO_T;
    static constexpr uint32_t DZ2 = DZ1 + (N1 / 8) * LBO_T;
    static constexpr uint32_t W1 = DZ2 + (N2 / 8) * LBO_T;
    static constexpr uint32_t W2 = W1 + (N0 / 8) * LBO_W1;
    static constexpr uint32_t W_END = W2 + (N1 / 8) * LBO_W2;
    // fp32 staging of the NEXT tile's gathered rows (cp.async): MLP rows [3][64][D], GMF rows [3][64][F + 4] (padded: the head
    // reads one row per lane with 128-bit loads)
    static constexpr uint32_t SM = (W_END + 127) / 128 * 128;
    static constexpr uint32_t SG = SM + 3 * 64 * D * 4;
    static constexpr int GROW = F + 4;
    static constexpr uint32_t TAIL = SG + 3 * 64 * GROW * 4;
    // the gW2 product reads A1^T as an M = 128 operand although only N1 <= 96 feature rows exist: MN-groups beyond N1/8 fall
    // into the images behind A1 (finite bf16 data, rows of the result that nobody reads) -- keep that window inside the buffer
    static constexpr uint32_t SPAN = (A1 + 16 * LBO_T + 256 > TAIL) ? (A1 + 16 * LBO_T + 256) : TAIL;
    static constexpr uint32_t BYTES = (SPAN + 127) / 128 * 128;
    // TMEM columns
    static constexpr int C_Z1 = 0, C_Z2 = N1, C_DA0 = 128, C_GW2 = 128 + N0, C_GW1 = 128 + N0 + 32 * ((N2 + 31) / 32);
    static constexpr int C_END = C_GW1 + N1;
    static_assert(N1 + N2 <= 128 && C_END <= 512, "TMEM budget");
};

__device__ __forceinline__ void fused_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void fused_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 / 8 consecutive fp32 columns of this warp's 32 TMEM lanes (thread = lane = tile row); the caller waits (tmem_wait)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 8 bf16 (k0 .. k0+7, k0 % 8 == 0) of tile row `row` of a K-major image with 128 rows: one 16-byte store
__device__ __forceinline__ void image_store8(unsigned char *img, uint32_t lbo, int row, int k0, const float *v)
{
    uint4 a;
    a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]); a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4 *>(img + (uint32_t)(k0 >> 3) * lbo + (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u) = a;
}

template <int F>
__global__ void __launch_bounds__(kFusedThreads, 1) neumf_fused_kernel(FusedParams p)
{
    using L = FusedLayout<F>;
    constexpr int D = L::D, N0 = L::N0, N1 = L::N1, N2 = L::N2;
    constexpr int NP = 4;                         // column parts: every TMEM tile is split over warp / 4
    constexpr int C1 = N1 / NP, C2 = N2 / NP, CG = F / NP, C0 = N0 / NP;   // columns per thread: Z1/dA1, Z2, GMF, dA0
    static_assert(C1 == 16 && C2 == 8 && CG == 8 && C0 == 32, "the epilogues are written for factors = 32");
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_b1[N1], s_b2[N2], s_wp[2 * F + 1];
    __shared__ float s_pred[NP][128];
    __shared__ int s_idx[2][3][kFusedTile];       // [buffer][u | i | j][triple] of the current and the next tile
    __shared__ float s_colsum[N1 + N2 + 2 * F];   // final cross-thread reduction of the register column sums
    __shared__ double s_red[11];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, part = warp >> 2;     // TMEM lane quarter of this warp, column quarter it works on
    const int row = q * 32 + lane;                // tile row owned in every epilogue (4 threads share it)
    const bool pos_row = row < kFusedTile;
    const float sign = pos_row ? 1.f : -1.f;

    // ---- one-off: TMEM, barrier, weights as bf16 operand images, biases
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    const float *W1 = p.W, *b1 = W1 + (size_t)N1 * N0, *W2 = b1 + N1, *b2 = W2 + (size_t)N2 * N1, *wp = b2 + N2;
    for (int it = tid; it < N1 * (N0 / 8); it += kFusedThreads) {            // W1 [N1 rows, N0 k] K-major image
        const int r = it / (N0 / 8), kg = it % (N0 / 8);
        const float4 *s4 = reinterpret_cast<const float4 *>(W1 + (size_t)r * N0 + kg * 8);
        const float4 a = __ldg(s4), b = __ldg(s4 + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4 *>(smem + L::W1 + (uint32_t)kg * L::LBO_W1 + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u) = o;
    }
    for (int it = tid; it < N2 * (N1 / 8); it += kFusedThreads) {            // W2 [N2 rows, N1 k]
        const int r = it / (N1 / 8), kg = it % (N1 / 8);
        const float4 *s4 = reinterpret_cast<const float4 *>(W2 + (size_t)r * N1 + kg * 8);
        const float4 a = __ldg(s4), b = __ldg(s4 + 1);
        uint4 o;
        o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w); o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
        *reinterpret_cast<uint4 *>(smem + L::W2 + (uint32_t)kg * L::LBO_W2 + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u) = o;
    }
    for (int k = tid; k < N1; k += kFusedThreads) s_b1[k] = b1[k];
    for (int k = tid; k < N2; k += kFusedThreads) s_b2[k] = b2[k];
    for (int k = tid; k < 2 * F + 1; k += kFusedThreads) s_wp[k] = wp[k];
    if (tid < 11) s_red[tid] = 0.0;
    for (int k = tid; k < N1 + N2 + 2 * F; k += kFusedThreads) s_colsum[k] = 0.f;
    // the windows the padded M = 128 views may touch must hold finite numbers before the first product reads them
    for (uint32_t o = L::DZ1 + tid * 16u; o < L::W1; o += kFusedThreads * 16u) *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);          // this warp's TMEM lanes
    uint32_t phase = 0;

    // register accumulators carried across tiles (reduced once at the end)
    float acc_loss = 0.f, acc_l1[5] = {0, 0, 0, 0, 0}, acc_s2[5] = {0, 0, 0, 0, 0};
    float gb1[C1], gb2[C2], gwg[CG], gwh[C2];
#pragma unroll
    for (int k = 0; k < C1; ++k) gb1[k] = 0.f;
#pragma unroll
    for (int k = 0; k < C2; ++k) { gb2[k] = 0.f; gwh[k] = 0.f; }
#pragma unroll
    for (int k = 0; k < CG; ++k) gwg[k] = 0.f;

    auto idesc = [&](int N, bool a_mn, bool b_mn) { return umma_idesc_bf16_f32(128, N, a_mn, b_mn); };

    const long long ntiles = (p.B + kFusedTile - 1) / kFusedTile;
    bool first_tile = true;

    // a tile's index lists: threads 0..191 hold one index each (u | i | j of triple tid % 64)
    auto fetch_index = [&](long long tile_) {
        int v = 0;
        if (tid < 3 * kFusedTile) {
            const int kind = tid / kFusedTile, r = tid % kFusedTile;
            const long long t = tile_ * kFusedTile + r;
            if (t < p.B) v = __ldg((kind == 0 ? p.bu : kind == 1 ? p.bi : p.bj) + t);
        }
        return v;
    };
    auto store_index = [&](int v, int buf) {
        if (tid < 3 * kFusedTile) s_idx[buf][tid / kFusedTile][tid % kFusedTile] = v;
    };
    // asynchronous global -> shared copies (cp.async, 16 bytes per lane) of a tile's gathered rows, straight from the tables
    constexpr int G = D / 4;                               // lanes per D-float MLP row
    constexpr int GROUPS = kFusedThreads / G;
    constexpr int GG = F / 4;                              // lanes per F-float GMF row
    constexpr int GGROUPS = kFusedThreads / GG;
    auto prefetch_mlp = [&](long long tile_, int buf) {
        const int nt_ = (int)min((long long)kFusedTile, p.B - tile_ * kFusedTile);
        const int gl = tid % G, grp = tid / G;
#pragma unroll
        for (int w = grp; w < 3 * kFusedTile; w += GROUPS) {
            const int kind = w / kFusedTile, r = w % kFusedTile;            // 0: UM[u], 1: IM[i], 2: IM[j]
            const uint32_t off = L::SM + (uint32_t)(w * D + gl * 4) * 4u;
            if (r < nt_) {
                const float *src = (kind == 0 ? p.UM : p.IM) + (size_t)s_idx[buf][kind][r] * D + gl * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + off), "l"(src) : "memory");
            } else {
                *reinterpret_cast<float4 *>(smem + off) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto prefetch_gmf = [&](long long tile_, int buf) {
        const int nt_ = (int)min((long long)kFusedTile, p.B - tile_ * kFusedTile);
        const int gl = tid % GG, grp = tid / GG;
#pragma unroll
        for (int w = grp; w < 3 * kFusedTile; w += GGROUPS) {
            const int kind = w / kFusedTile, r = w % kFusedTile;            // 0: UG[u], 1: IG[i], 2: IG[j]
            const uint32_t off = L::SG + (uint32_t)(w * L::GROW + gl * 4) * 4u;
            if (r < nt_) {
                const float *src = (kind == 0 ? p.UG : p.IG) + (size_t)s_idx[buf][kind][r] * F + gl * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + off), "l"(src) : "memory");
            } else {
                *reinterpret_cast<float4 *>(smem + off) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int cur = 0;
    if ((long long)blockIdx.x < ntiles) {
        store_index(fetch_index(blockIdx.x), 0);
        __syncthreads();
        prefetch_mlp(blockIdx.x, 0);
        prefetch_gmf(blockIdx.x, 0);
    }
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, cur ^= 1) {
        const long long t0 = tile * kFusedTile;
        const int nt = (int)min((long long)kFusedTile, p.B - t0);           // valid triples in this tile
        const int tr = row & (kFusedTile - 1);                              // triple of this thread's row
        const bool ok = tr < nt;
        const long long next_tile = tile + gridDim.x;
        const bool has_next = next_tile < ntiles;
        const int idx_next = has_next ? fetch_index(next_tile) : 0;        // lands while this tile's rows are converted
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                                    // staged rows of this tile visible to everyone
        const int u = s_idx[cur][0][tr];
        const int item = s_idx[cur][pos_row ? 1 : 2][tr];

        // ---------------------------------------------------------------- A0: staged fp32 rows -> bf16 K-major image
        {
            const int gl = tid % G, grp = tid / G;
#pragma unroll
            for (int w = grp; w < 3 * kFusedTile; w += GROUPS) {
                const int kind = w / kFusedTile, r = w % kFusedTile;        // 0: UM[u] -> rows r and r+64; 1: IM[i]; 2: IM[j]
                const float4 v = *reinterpret_cast<const float4 *>(smem + L::SM + (uint32_t)(w * D + gl * 4) * 4u);
                if (p.has_reg && kind < 2 && r < nt) {     // UM_u and IM_i rows enter the regulariser once per triple
                    const float a = fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w);
                    const float s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
                    if (kind == 0) { acc_l1[1] += a; acc_s2[1] += s2; } else { acc_l1[3] += a; acc_s2[3] += s2; }
                }
                uint2 o;
                o.x = pack_bf16x2(v.x, v.y);
                o.y = pack_bf16x2(v.z, v.w);
                const int k = (kind == 0 ? 0 : D) + gl * 4;
                const int trow = kind == 2 ? r + kFusedTile : r;
                const uint32_t off = L::A0 + (uint32_t)(k >> 3) * L::LBO_T + (uint32_t)(trow >> 3) * 128u + (uint32_t)(trow & 7) * 16u +
                                     (uint32_t)(k & 7) * 2u;
                *reinterpret_cast<uint2 *>(smem + off) = o;
                if (kind == 0) *reinterpret_cast<uint2 *>(smem + off + (kFusedTile >> 3) * 128u) = o;   // the neg row of the triple
            }
        }
        if (has_next) store_index(idx_next, cur ^ 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();

        // ---------------------------------------------------------------- Z1 = A0 W1^T
        if (tid == 0) {
            tc_fence_after();
            const uint32_t id = idesc(N1, false, false);
#pragma unroll
            for (int kk = 0; kk < N0 / 16; ++kk)
                fused_mma(tmem + L::C_Z1, umma_smem_desc(sbase + L::A0 + kk * 2 * L::LBO_T, L::LBO_T, 128),
                          umma_smem_desc(sbase + L::W1 + kk * 2 * L::LBO_W1, L::LBO_W1, 128), id, kk > 0);
            fused_commit(&s_bar);
        }
        if (has_next) prefetch_mlp(next_tile, cur ^ 1);      // the MLP staging has been consumed: refill it under the MMAs
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        tc_fence_after();

        // ---------------------------------------------------------------- A1 = relu(Z1 + b1) -> image
        {
            const int k0 = part * C1;
            uint32_t r16[16];
            tmem_ld16_nowait(taddr + (uint32_t)(L::C_Z1 + k0), r16);
            tmem_wait();
            float v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float z = __uint_as_float(r16[e]) + s_b1[k0 + e];
                v[e] = (ok && z > 0.f) ? z : 0.f;
            }
            image_store8(smem + L::A1, L::LBO_T, row, k0, v);
            image_store8(smem + L::A1, L::LBO_T, row, k0 + 8, v + 8);
        }
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();

        // ---------------------------------------------------------------- Z2 = A1 W2^T
        if (tid == 0) {
            tc_fence_after();
            const uint32_t id = idesc(N2, false, false);
#pragma unroll
            for (int kk = 0; kk < N1 / 16; ++kk)
                fused_mma(tmem + L::C_Z2, umma_smem_desc(sbase + L::A1 + kk * 2 * L::LBO_T, L::LBO_T, 128),
                          umma_smem_desc(sbase + L::W2 + kk * 2 * L::LBO_W2, L::LBO_W2, 128), id, kk > 0);
            fused_commit(&s_bar);
        }
        // GMF rows of this thread's 8 columns (shared memory staging; rows beyond the batch were staged as zeros)
        float gu[CG], gi[CG];
        {
            const int k0 = part * CG;
            const float4 *ug4 = reinterpret_cast<const float4 *>(smem + L::SG + (uint32_t)((0 * kFusedTile + tr) * L::GROW + k0) * 4u);
            const float4 *ig4 =
                reinterpret_cast<const float4 *>(smem + L::SG + (uint32_t)(((pos_row ? 1 : 2) * kFusedTile + tr) * L::GROW + k0) * 4u);
#pragma unroll
            for (int c = 0; c < CG / 4; ++c) {
                const float4 a = ug4[c], b = ig4[c];
                gu[4 * c] = a.x; gu[4 * c + 1] = a.y; gu[4 * c + 2] = a.z; gu[4 * c + 3] = a.w;
                gi[4 * c] = b.x; gi[4 * c + 1] = b.y; gi[4 * c + 2] = b.z; gi[4 * c + 3] = b.w;
            }
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        tc_fence_after();

        // ---------------------------------------------------------------- head: h, prediction, BPR coefficient, dZ2, GMF gradients
        float hval[C2];
        {
            const int k0 = part * C2;
            uint32_t r8[8];
            tmem_ld8_nowait(taddr + (uint32_t)(L::C_Z2 + k0), r8);
            tmem_wait();
            float part_sum = 0.f;
#pragma unroll
            for (int e = 0; e < C2; ++e) {
                const float z = __uint_as_float(r8[e]) + s_b2[k0 + e];
                hval[e] = (ok && z > 0.f) ? z : 0.f;
                part_sum = fmaf(s_wp[F + k0 + e], hval[e], part_sum);
            }
#pragma unroll
            for (int e = 0; e < CG; ++e) part_sum = fmaf(s_wp[part * CG + e], gu[e] * gi[e], part_sum);
            s_pred[part][row] = part_sum;
            if (p.has_reg && ok) {
                float a1 = 0.f, q1 = 0.f, a2 = 0.f, q2 = 0.f;
#pragma unroll
                for (int e = 0; e < CG; ++e) {
                    a1 += fabsf(gu[e]); q1 = fmaf(gu[e], gu[e], q1);
                    a2 += fabsf(gi[e]); q2 = fmaf(gi[e], gi[e], q2);
                }
                if (pos_row) { acc_l1[0] += a1; acc_s2[0] += q1; acc_l1[2] += a2; acc_s2[2] += q2; }   // UG_u once, IG_i
                else { acc_l1[4] += a2; acc_s2[4] += q2; }                                               // IG_j
            }
        }
        __syncthreads();
        {
            const float pp = (s_pred[0][tr] + s_pred[1][tr]) + (s_pred[2][tr] + s_pred[3][tr]);
            const float pn = (s_pred[0][tr + kFusedTile] + s_pred[1][tr + kFusedTile]) +
                             (s_pred[2][tr + kFusedTile] + s_pred[3][tr + kFusedTile]);
            const float x = pp - pn;                                    // the predict bias cancels in the pair
            const float sg = 1.f / (1.f + expf(-x));
            if (ok && pos_row && part == 0) acc_loss += -logf(1e-10f + sg);
            const float cbpr = -(sg * (1.f - sg)) / (1e-10f + sg);
            const float dp = ok ? sign * cbpr : 0.f;                    // d loss / d pred of THIS row
            float dz[C2];
#pragma unroll
            for (int e = 0; e < C2; ++e) {
                dz[e] = hval[e] > 0.f ? dp * s_wp[F + part * C2 + e] : 0.f;
                gb2[e] += dz[e];
                gwh[e] += dp * hval[e];
            }
#pragma unroll
            for (int e = 0; e < CG; ++e) gwg[e] += dp * (gu[e] * gi[e]);
            image_store8(smem + L::DZ2, L::LBO_T, row, part * C2, dz);
            if (p.apply && ok) {
                const int k0 = part * CG;
#pragma unroll
                for (int c = 0; c < CG / 4; ++c) {
                    Vec<4> g1, g2;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float w = dp * s_wp[k0 + 4 * c + e];
                        g1.v[e] = w * gi[4 * c + e];                    // d / d UG[u]
                        g2.v[e] = w * gu[4 * c + e];                    // d / d IG[item]
                    }
                    red_row<4>(p.gUG + (size_t)u * F + k0 + 4 * c, g1);
                    red_row<4>(p.gIG + (size_t)item * F + k0 + 4 * c, g2);
                }
                if (part == 0) {
                    if (pos_row) {
                        red_add_u32(p.cntU + u, 1u);
                        red_add_u64(p.cntI + item, 1ull);
                    } else {
                        red_add_u64(p.cntI + item, 1ull << 32);
                    }
                }
            }
        }
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (!p.apply) {                                      // loss only
            if (has_next) prefetch_gmf(next_tile, cur ^ 1);
            first_tile = false;
            continue;
        }

        // ---------------------------------------------------------------- dA1 = dZ2 W2 ;  gW2^T += A1^T dZ2
        if (tid == 0) {
            tc_fence_after();
            {   // B = W2 read transposed (MN-major view of its K-major image): mn = in (N1), k = out (N2)
                const uint32_t id = idesc(N1, false, true);
#pragma unroll
                for (int kk = 0; kk < N2 / 16; ++kk)
                    fused_mma(tmem + L::C_Z1, umma_smem_desc(sbase + L::DZ2 + kk * 2 * L::LBO_T, L::LBO_T, 128),
                              umma_smem_desc(sbase + L::W2 + kk * 256, 128, L::LBO_W2), id, kk > 0);
            }
            {   // A = A1^T (features on the M side, tile rows as K), B = dZ2 read transposed: both MN-major views
                const uint32_t id = idesc(N2, true, true);
#pragma unroll
                for (int kk = 0; kk < 128 / 16; ++kk)
                    fused_mma(tmem + L::C_GW2, umma_smem_desc(sbase + L::A1 + kk * 256, 128, L::LBO_T),
                              umma_smem_desc(sbase + L::DZ2 + kk * 256, 128, L::LBO_T), id, (!first_tile) || kk > 0);
            }
            fused_commit(&s_bar);
        }
        if (has_next) prefetch_gmf(next_tile, cur ^ 1);      // the GMF staging has been consumed by the head
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        tc_fence_after();

        // ---------------------------------------------------------------- dZ1 = dA1 [A1 > 0] -> image
        {
            const int k0 = part * C1;
            uint32_t r16[16];
            tmem_ld16_nowait(taddr + (uint32_t)(L::C_Z1 + k0), r16);
            const uint32_t base = (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
            const uint4 m0 = *reinterpret_cast<const uint4 *>(smem + L::A1 + (uint32_t)(k0 >> 3) * L::LBO_T + base);
            const uint4 m1 = *reinterpret_cast<const uint4 *>(smem + L::A1 + (uint32_t)((k0 >> 3) + 1) * L::LBO_T + base);
            const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
            tmem_wait();
            float v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const uint32_t hbits = (e & 1) ? (mw[e >> 1] >> 16) : (mw[e >> 1] & 0xffffu);   // bf16 of A1[row][k0 + e]
                const bool on = (hbits & 0x7fffu) != 0u && (hbits & 0x8000u) == 0u;              // > 0
                v[e] = on ? __uint_as_float(r16[e]) : 0.f;
                gb1[e] += v[e];
            }
            image_store8(smem + L::DZ1, L::LBO_T, row, k0, v);
            image_store8(smem + L::DZ1, L::LBO_T, row, k0 + 8, v + 8);
        }
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();

        // ---------------------------------------------------------------- dA0 = dZ1 W1 ;  gW1^T += A0^T dZ1
        if (tid == 0) {
            tc_fence_after();
            {
                const uint32_t id = idesc(N0, false, true);
#pragma unroll
                for (int kk = 0; kk < N1 / 16; ++kk)
                    fused_mma(tmem + L::C_DA0, umma_smem_desc(sbase + L::DZ1 + kk * 2 * L::LBO_T, L::LBO_T, 128),
                              umma_smem_desc(sbase + L::W1 + kk * 256, 128, L::LBO_W1), id, kk > 0);
            }
            {
                const uint32_t id = idesc(N1, true, true);
#pragma unroll
                for (int kk = 0; kk < 128 / 16; ++kk)
                    fused_mma(tmem + L::C_GW1, umma_smem_desc(sbase + L::A0 + kk * 256, 128, L::LBO_T),
                              umma_smem_desc(sbase + L::DZ1 + kk * 256, 128, L::LBO_T), id, (!first_tile) || kk > 0);
            }
            fused_commit(&s_bar);
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        tc_fence_after();

        // ---------------------------------------------------------------- scatter dA0: user half -> gUM[u], item half -> gIM[item]
        {
            // parts 0,1: columns [0,64) = the user half of the MLP row; parts 2,3: columns [64,128) = the item half
            float *dst = (part < 2 ? p.gUM + (size_t)u * D : p.gIM + (size_t)item * D) + (part & 1) * C0;
            uint32_t ra[16], rb[16];
            tmem_ld16_nowait(taddr + (uint32_t)(L::C_DA0 + part * C0), ra);
            tmem_ld16_nowait(taddr + (uint32_t)(L::C_DA0 + part * C0 + 16), rb);
            tmem_wait();
            if (ok) {
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                    Vec<4> g;
                    g.v[0] = __uint_as_float(ra[4 * e4]); g.v[1] = __uint_as_float(ra[4 * e4 + 1]);
                    g.v[2] = __uint_as_float(ra[4 * e4 + 2]); g.v[3] = __uint_as_float(ra[4 * e4 + 3]);
                    red_row<4>(dst + 4 * e4, g);
                }
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                    Vec<4> g;
                    g.v[0] = __uint_as_float(rb[4 * e4]); g.v[1] = __uint_as_float(rb[4 * e4 + 1]);
                    g.v[2] = __uint_as_float(rb[4 * e4 + 2]); g.v[3] = __uint_as_float(rb[4 * e4 + 3]);
                    red_row<4>(dst + 16 + 4 * e4, g);
                }
            }
        }
        first_tile = false;
        tc_fence_before();
        __syncthreads();                                     // A0 / TMEM free for the next tile
    }

    // ---------------------------------------------------------------- once per CTA: weight gradients out of TMEM
    tc_fence_after();
    if (p.apply && !first_tile) {
        float *gW1 = p.gW, *gb1g = gW1 + (size_t)N1 * N0, *gW2 = gb1g + N1, *gb2g = gW2 + (size_t)N2 * N1, *gwp = gb2g + N2;
        {   // gW1^T: lane = input feature m (0..N0-1 = 128 rows), column = output n
            const int n0 = part * C1;
            uint32_t r16[16];
            tmem_ld16_nowait(taddr + (uint32_t)(L::C_GW1 + n0), r16);
            tmem_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) atomicAdd(gW1 + (size_t)(n0 + e) * N0 + row, __uint_as_float(r16[e]));
        }
        {   // gW2^T: lanes 0..N1-1 valid
            const int n0 = part * C2;
            uint32_t r8[8];
            tmem_ld8_nowait(taddr + (uint32_t)(L::C_GW2 + n0), r8);       // whole warp executes the collective load
            tmem_wait();
            if (row < N1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) atomicAdd(gW2 + (size_t)(n0 + e) * N1 + row, __uint_as_float(r8[e]));
            }
        }
        // register column sums: warp shuffle over the 32 rows of the warp, then shared, then one global atomic per column
#pragma unroll
        for (int k = 0; k < C1; ++k) {
            float v = gb1[k];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) atomicAdd(&s_colsum[part * C1 + k], v);
        }
#pragma unroll
        for (int k = 0; k < C2; ++k) {
            float v = gb2[k], g = gwg[k], h = gwh[k];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                v += __shfl_xor_sync(0xffffffffu, v, off);
                g += __shfl_xor_sync(0xffffffffu, g, off);
                h += __shfl_xor_sync(0xffffffffu, h, off);
            }
            if (lane == 0) {
                atomicAdd(&s_colsum[N1 + part * C2 + k], v);
                atomicAdd(&s_colsum[N1 + N2 + part * CG + k], g);
                atomicAdd(&s_colsum[N1 + N2 + F + part * C2 + k], h);
            }
        }
        __syncthreads();
        for (int k = tid; k < N1; k += kFusedThreads) if (s_colsum[k] != 0.f) atomicAdd(gb1g + k, s_colsum[k]);
        for (int k = tid; k < N2; k += kFusedThreads) if (s_colsum[N1 + k] != 0.f) atomicAdd(gb2g + k, s_colsum[N1 + k]);
        for (int k = tid; k < 2 * F; k += kFusedThreads) if (s_colsum[N1 + N2 + k] != 0.f) atomicAdd(gwp + k, s_colsum[N1 + N2 + k]);
    }
    // loss and regulariser norms
    {
        const int nv = p.has_reg ? 11 : 1;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            if (k >= nv) break;
            float v = k == 0 ? acc_loss : (k <= 5 ? acc_l1[k - 1] : acc_s2[k - 6]);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) atomicAdd(&s_red[k], (double)v);
        }
        __syncthreads();
        if (tid < nv && s_red[tid] != 0.0) atomicAdd(p.red + tid, s_red[tid]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int F>
static int launch_neumf_fused_f(const FusedParams &p, cudaStream_t st)
{
    using L = FusedLayout<F>;
    static bool attr_set = false;
    if (!attr_set) {
        DRB_CUDA(cudaFuncSetAttribute(neumf_fused_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
        attr_set = true;
    }
    long long tiles = (p.B + kFusedTile - 1) / kFusedTile;
    int grid = (int)(tiles < (long long)sm_count() ? tiles : (long long)sm_count());
    if (grid < 1) grid = 1;
    neumf_fused_kernel<F><<<grid, kFusedThreads, L::BYTES, st>>>(p);
    DRB_CUDA(cudaGetLastError());
    return DRB_OK;
}

// The epilogues split every TMEM tile into two column halves of 16-column chunks: factors must be a multiple of 32, and
// 10 F accumulator columns must fit the 512 of TMEM -> factors = 32 (BASELINE config 3).  Other shapes use the layer-wise path.
static bool neumf_fused_supported(int F, int L, int mode, float dropout)
{
    return L == 2 && mode == 0 && dropout == 0.f && F == 32;
}

static int launch_neumf_fused(int F, const FusedParams &p, cudaStream_t st)
{
    if (F == 32) return launch_neumf_fused_f<32>(p, st);
    DRB_REQUIRE(false, "neumf fused tower: unsupported factors=%d", F);
}

}  // namespace drb
