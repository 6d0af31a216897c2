// p2p.cu -- the user-sharded multi-GPU step as ONE persistent kernel per rank with the exchange inside.
//
// The reference has no multi-device path; the single-GPU step kernel is the oracle.  The NCCL form of the sharded step
// (comm.cu) runs phase 1 -> grouped all-reduce of {gQ 6.85 MB, counters, scalars} -> phase 2 as three stream operations per
// step, fully serialised.  Here each rank launches the SAME persistent cooperative kernel as on one GPU (step_kernel.cuh)
// with an exchange policy that talks to the other ranks through peer-mapped memory (CUDA IPC over NVLink / NVSwitch):
//
//   phase 1     local triples; the item-side gradient, the item counters and the 8 loss/norm partial sums accumulate into
//               THIS rank's exchange buffer X[b] (b = step parity; double-buffered so nobody waits for a buffer to drain)
//   rendezvous A  CTA 0 publishes a system-scope release flag in every peer's buffer and waits for the peers' flags
//               (bounded spin -> DRB_ERR_PEER, never a hung GPU); it then adds the ranks' 8 scalars in rank order
//   item slice  rank r owns items [I r/N, I (r+1)/N): it adds the N ranks' accumulator rows and counters straight out of peer
//               memory (rank order), applies the regulariser + SGD / Adam rule once, and stores the new rows into EVERY
//               rank's replica of Q (peer stores).  6.85 MB (N-1)/N read + written per rank and step instead of an
//               all-reduce of the whole table; replicas are bit-identical by construction (one writer per row).
//   rendezvous B  "my slice has landed everywhere" flags; meanwhile the local user rows are swept and X[b^1] is cleared
//   next step.
// Two cross-GPU rendezvous per step, no NCCL, no relaunch, no host round trip.
#include <stdlib.h>
#include <string.h>

#include "step_kernel.cuh"

namespace drb {

constexpr int kMaxPeers = 8;
constexpr size_t kCtrlBytes = 4096;

// control block at the start of every rank's exchange buffer
struct P2PCtrl {
    unsigned long long flag_a[64];   // [q] written by peer q: last step for which q finished phase 1
    unsigned long long flag_b[64];   // [q] written by peer q: last step whose item slice q has stored everywhere
    unsigned long long go;           // local: verdict of CTA 0's wait, read by the grid after the barrier
    unsigned long long pad[7];
    double totals[8];                // local: the ranks' scalars added in rank order
};
static_assert(sizeof(P2PCtrl) <= kCtrlBytes, "control block too large");

struct P2PLayout {
    size_t acc[2], gq[2], cnt[2], q, total;
};
static inline size_t p2p_align(size_t x) { return (x + 255) & ~(size_t)255; }
static P2PLayout p2p_layout(int I, int F)
{
    P2PLayout L;
    size_t off = kCtrlBytes;
    for (int b = 0; b < 2; ++b) { L.acc[b] = off; off += 256; }
    for (int b = 0; b < 2; ++b) { L.gq[b] = off; off += p2p_align(sizeof(float) * (size_t)I * F); }
    for (int b = 0; b < 2; ++b) { L.cnt[b] = off; off += p2p_align(sizeof(unsigned long long) * (size_t)I); }
    L.q = off; off += p2p_align(sizeof(float) * (size_t)I * F);
    L.total = off;
    return L;
}

struct P2PParams {
    char *peer[kMaxPeers];           // base of every rank's exchange buffer as mapped HERE (peer[rank] = own buffer)
    int rank, world;
    int item_lo[kMaxPeers + 1];
    unsigned long long seq0;         // global steps finished before this launch (flags carry absolute step numbers)
    unsigned long long timeout_ns;
    size_t off_acc[2], off_gq[2], off_cnt[2], off_q;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// peer memory is never served from this SM's L1: system-scope relaxed accesses
template <int VEC>
__device__ __forceinline__ Vec<VEC> ld_peer(const float *p)
{
    Vec<VEC> r;
    if constexpr (VEC == 4) {
        asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p) : "memory");
    } else if constexpr (VEC == 2) {
        asm volatile("ld.relaxed.sys.global.v2.f32 {%0, %1}, [%2];" : "=f"(r.v[0]), "=f"(r.v[1]) : "l"(p) : "memory");
    } else {
        asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(r.v[0]) : "l"(p) : "memory");
    }
    return r;
}
template <int VEC>
__device__ __forceinline__ void st_peer(float *p, const Vec<VEC> &r)
{
    if constexpr (VEC == 4) {
        asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]),
                     "f"(r.v[3]) : "memory");
    } else if constexpr (VEC == 2) {
        asm volatile("st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]) : "memory");
    } else {
        asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(r.v[0]) : "memory");
    }
}

struct P2PExchange {
    static constexpr bool kActive = true;
    P2PParams x;
    int b;                           // parity of the running step

    __device__ __forceinline__ P2PCtrl *ctrl(int q) const { return (P2PCtrl *)x.peer[q]; }

    __device__ __forceinline__ void begin_step(StepParams &p, long long s, double *&acc)
    {
        b = (int)((x.seq0 + (unsigned long long)s) & 1ull);
        char *self = x.peer[x.rank];
        p.ws.gQ = (float *)(self + x.off_gq[b]);
        p.ws.cntI = (unsigned long long *)(self + x.off_cnt[b]);
        acc = (double *)(self + x.off_acc[b]);
    }

    // wait until every peer's flag (kind 0: A, 1: B) has reached seq; false on time-out
    __device__ __forceinline__ bool wait_flags(int kind, unsigned long long seq) const
    {
        const P2PCtrl *me = ctrl(x.rank);
        const unsigned long long t0 = globaltimer_ns();
        for (int q = 0; q < x.world; ++q) {
            if (q == x.rank) continue;
            const unsigned long long *f = kind == 0 ? &me->flag_a[q] : &me->flag_b[q];
            while (ld_acquire_sys(f) < seq) {
                if (globaltimer_ns() - t0 > x.timeout_ns) return false;
            }
        }
        return true;
    }

    // after the local grid barrier that ends phase 1: rendezvous A + the global scalars
    __device__ __forceinline__ bool after_phase1(const StepParams &p, long long s, double *&acc, unsigned long long &epoch)
    {
        const unsigned long long seq = x.seq0 + (unsigned long long)s + 1ull;
        P2PCtrl *me = ctrl(x.rank);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            __threadfence_system();                          // this rank's accumulators (published by the grid barrier)
            for (int q = 0; q < x.world; ++q)
                if (q != x.rank) st_release_sys(&ctrl(q)->flag_a[x.rank], seq);
            const bool ok = wait_flags(0, seq);
            if (ok) {
                double tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int q = 0; q < x.world; ++q) {              // rank order: every rank computes the same doubles
                    const double *a = (const double *)(x.peer[q] + x.off_acc[b]);
                    for (int k = 0; k < 8; ++k) {
                        double v;
                        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(a + k) : "memory");
                        tot[k] += v;
                    }
                }
                for (int k = 0; k < 8; ++k) me->totals[k] = tot[k];
            }
            __threadfence();
            *(volatile unsigned long long *)&me->go = ok ? seq : ~0ull;
        }
        grid_barrier(&p.ws.hdr->barrier, epoch);
        const unsigned long long go = *(volatile unsigned long long *)&me->go;
        if (go != seq) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                p.ws.hdr->status = DRB_ERR_PEER;
                p.ws.hdr->nan_step = p.first_step + s;
            }
            return false;
        }
        acc = me->totals;
        return true;
    }

    // phase 2, item half: this rank's slice of the item table
    template <int VEC, int W, int NCH>
    __device__ __forceinline__ void item_slice(const StepParams &p, long long s, const Norms &nm, const AdamCoef &ac, int gl,
                                               int group, int groups_per_cta, int chunks, unsigned long long &epoch)
    {
        const int F = p.F;
        const int lo = x.item_lo[x.rank], hi = x.item_lo[x.rank + 1];
        const long long tg = (long long)gridDim.x * groups_per_cta;
        const bool adam = p.opt == DRB_OPT_ADAM;
        for (long long row = lo + (long long)blockIdx.x * groups_per_cta + group; row < hi; row += tg) {
            unsigned long long cnt = 0;
            if (gl == 0) {
                for (int q = 0; q < x.world; ++q) {
                    unsigned long long c;
                    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];"
                                 : "=l"(c) : "l"((const unsigned long long *)(x.peer[q] + x.off_cnt[b]) + row) : "memory");
                    cnt += c;                                    // pos | neg << 32: the halves never carry below 2^32 occurrences
                }
            }
            cnt = __shfl_sync(0xffffffffu, cnt, (threadIdx.x & 31) - gl);
            const bool touched = cnt != 0;
            if (!touched && !adam) continue;                     // SGD: an untouched row has no gradient and no regulariser
            const float ca = (float)(unsigned)(cnt & 0xffffffffull), cb = p.neg_mult * (float)(unsigned)(cnt >> 32);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int c = gl + ch * W;
                if (c >= chunks) continue;
                const size_t o = (size_t)row * F + (size_t)c * VEC;
                Vec<VEC> g;
#pragma unroll
                for (int e = 0; e < VEC; ++e) g.v[e] = 0.f;
                if (touched) {
                    for (int q = 0; q < x.world; ++q) {          // rank order
                        const Vec<VEC> t = ld_peer<VEC>((const float *)(x.peer[q] + x.off_gq[b]) + o);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) g.v[e] += t.v[e];
                    }
                }
                Vec<VEC> th = ld_row<VEC>(p.Q + o), m, v;
                if (adam) {
                    m = ld_row<VEC>(p.ws.mQ + o);
                    v = ld_row<VEC>(p.ws.vQ + o);
                }
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const float t = th.v[e];
                    float gg = p.gscale * g.v[e];
                    if (touched) {
                        const float sg = p.reg1 * sgnf(t);
                        gg += ca * (sg + p.reg2 * t * nm.inv_i) + cb * (sg + p.reg2 * t * nm.inv_j);
                    }
                    if (!adam) {
                        th.v[e] = t - p.lr * gg;
                    } else {
                        float mm = m.v[e], vv = v.v[e];
                        mm = mm + (gg - mm) * (1.f - p.beta1);
                        vv = vv * p.beta2 + (1.f - p.beta2) * gg * gg;
                        th.v[e] = t - ac.step_size * (mm / (sqrtf(vv) / ac.bc2_sqrt + p.eps));
                        m.v[e] = mm;
                        v.v[e] = vv;
                    }
                }
                if (adam) {
                    st_row<VEC>(p.ws.mQ + o, m);
                    st_row<VEC>(p.ws.vQ + o, v);
                }
                for (int q = 0; q < x.world; ++q) {              // one writer per row: replicas stay bit-identical
                    float *dst = (float *)(x.peer[q] + x.off_q) + o;
                    if (q == x.rank) st_row<VEC>(dst, th); else st_peer<VEC>(dst, th);
                }
            }
        }
        __threadfence_system();                                  // the peer stores of this thread, before the B flag
        grid_barrier(&p.ws.hdr->barrier, epoch);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            __threadfence_system();
            const unsigned long long seq = x.seq0 + (unsigned long long)s + 1ull;
            for (int q = 0; q < x.world; ++q)
                if (q != x.rank) st_release_sys(&ctrl(q)->flag_b[x.rank], seq);
        }
        // clear the OTHER parity's accumulators: every peer finished reading them before it raised this step's A flag
        {
            char *self = x.peer[x.rank];
            float4 *gz = (float4 *)(self + x.off_gq[b ^ 1]);
            const long long n4 = ((long long)p.I * F) >> 2, gsz = (long long)gridDim.x * kThreads;
            const long long gt = (long long)blockIdx.x * kThreads + threadIdx.x;
            for (long long k = gt; k < n4; k += gsz) __stcg(gz + k, make_float4(0.f, 0.f, 0.f, 0.f));
            float *gtail = (float *)(self + x.off_gq[b ^ 1]);
            for (long long k = (n4 << 2) + gt; k < (long long)p.I * F; k += gsz) __stcg(gtail + k, 0.f);
            unsigned long long *cz = (unsigned long long *)(self + x.off_cnt[b ^ 1]);
            for (long long k = gt; k < p.I; k += gsz) __stcg(cz + k, 0ull);
            if (gt < 8) __stcg((double *)(self + x.off_acc[b ^ 1]) + gt, 0.0);
        }
    }

    // end of the step: every rank's slice is in this rank's replica of Q
    __device__ __forceinline__ bool end_step(const StepParams &p, long long s, unsigned long long &epoch)
    {
        const unsigned long long seq = x.seq0 + (unsigned long long)s + 1ull;
        P2PCtrl *me = ctrl(x.rank);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const bool ok = wait_flags(1, seq);
            __threadfence();
            *(volatile unsigned long long *)&me->go = ok ? (seq | (1ull << 62)) : ~0ull;
        }
        grid_barrier(&p.ws.hdr->barrier, epoch);
        const unsigned long long go = *(volatile unsigned long long *)&me->go;
        if (go != (seq | (1ull << 62))) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                p.ws.hdr->status = DRB_ERR_PEER;
                p.ws.hdr->nan_step = p.first_step + s;
            }
            return false;
        }
        return true;
    }
};

template <int VEC, int W, int NCH, bool LEAN>
__global__ void __launch_bounds__(kThreads, DRB_MINB) mf_bpr_p2p_steps_kernel(StepParams p, P2PParams xp)
{
    P2PExchange x;
    x.x = xp;
    x.b = 0;
    bpr_steps_body<VEC, W, NCH, false, P2PExchange, LEAN>(p, x);
}

typedef void (*P2PKernel)(StepParams, P2PParams);

// mf_bpr.cu: the geometry / index-tile cap chosen on the device for this factor count and table size (W == 0: none)
void lean_geom(int F, long long table_rows, int &W, int &NCH);
bool lean_enabled(int F, long long table_rows);
int lean_tile_cap(int F, long long table_rows);

// lean = the MF hot body (32-bit row offsets, its own lane geometry; step_params_lean); the exchange policy is the same
static P2PKernel pick_p2p(int F, bool lean = false, long long table_rows = 0)
{
    RowGeom g = row_geom(F);
    if (g.vec != 4 || g.nch != 1 || g.width < 4) return nullptr;
    if (lean) {
        int W, NCH;
        lean_geom(F, table_rows, W, NCH);
#define DRB_P2P(w, n) \
    if (W == w && NCH == n) return mf_bpr_p2p_steps_kernel<4, w, n, true>;
        DRB_P2P(4, 1) DRB_P2P(8, 1) DRB_P2P(16, 1) DRB_P2P(32, 1)
        DRB_P2P(2, 2) DRB_P2P(4, 2) DRB_P2P(8, 2) DRB_P2P(16, 2)
        DRB_P2P(1, 4) DRB_P2P(2, 4) DRB_P2P(4, 4) DRB_P2P(8, 4)
#undef DRB_P2P
    }
    switch (g.width) {
        case 4: return mf_bpr_p2p_steps_kernel<4, 4, 1, false>;
        case 8: return mf_bpr_p2p_steps_kernel<4, 8, 1, false>;
        case 16: return mf_bpr_p2p_steps_kernel<4, 16, 1, false>;
        case 32: return mf_bpr_p2p_steps_kernel<4, 32, 1, false>;
        default: return nullptr;
    }
}

}  // namespace drb

using namespace drb;

extern "C" size_t drb_p2p_buffer_bytes(int32_t I, int32_t F) { return p2p_layout(I, F).total; }

// offset (bytes) of the item-table replica inside an exchange buffer
extern "C" size_t drb_p2p_q_offset(int32_t I, int32_t F) { return p2p_layout(I, F).q; }

// cudaMalloc + zero an exchange buffer on the current device and export it (64-byte CUDA IPC handle)
extern "C" int drb_p2p_alloc(size_t bytes, void **d_ptr, uint8_t *h_handle64)
{
    DRB_REQUIRE(d_ptr && h_handle64 && bytes > 0, "p2p_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void *p = nullptr;
    DRB_CUDA(cudaMalloc(&p, bytes));
    DRB_CUDA(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e, "cudaIpcGetMemHandle", __FILE__, __LINE__);
    }
    memcpy(h_handle64, &h, 64);
    *d_ptr = p;
    return DRB_OK;
}

extern "C" int drb_p2p_open(const uint8_t *h_handle64, void **d_ptr)
{
    DRB_REQUIRE(h_handle64 && d_ptr, "p2p_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, 64);
    DRB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DRB_OK;
}

extern "C" int drb_p2p_close(void *d_ptr)
{
    if (d_ptr) DRB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return DRB_OK;
}

extern "C" int drb_p2p_free(void *d_ptr)
{
    if (d_ptr) DRB_CUDA(cudaFree(d_ptr));
    return DRB_OK;
}

// n_steps synchronous GLOBAL steps on this rank's shard in ONE persistent launch; step s trains the local triples
// [d_step_offsets[s], d_step_offsets[s+1]).  h_peer_bufs[q] = rank q's exchange buffer as mapped here (own buffer at [rank]);
// the item table replica lives inside the own buffer (drb_p2p_q_offset).  steps_done = global steps already run on these
// buffers (the rendezvous flags carry absolute step numbers).
extern "C" int drb_mf_bpr_train_steps_p2p(float *d_P_local, void *d_ws, int32_t U_local, int32_t I, int32_t F,
                                          void *const *h_peer_bufs, int32_t rank, int32_t world, const int32_t *d_bu,
                                          const int32_t *d_bi, const int32_t *d_bj, const int64_t *d_step_offsets,
                                          int64_t n_local, int64_t batch_per_rank, int64_t first_step, int64_t n_steps,
                                          const drb_hyper *hyper, int64_t steps_done, double *d_step_loss,
                                          double peer_timeout_s, int32_t sync_and_check, int64_t *bad_step, void *stream)
{
    DRB_REQUIRE(h_peer_bufs && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world && d_step_offsets && hyper &&
                    n_steps >= 0 && batch_per_rank > 0,
                "train_steps_p2p: bad arguments");
    DRB_REQUIRE(hyper->opt == DRB_OPT_SGD || hyper->opt == DRB_OPT_ADAM, "train_steps_p2p: SGD and Adam only");
    DRB_REQUIRE(hyper->loss == DRB_LOSS_BPR, "train_steps_p2p: BPR only (the other pair-wise losses use the NCCL step)");
    DRB_REQUIRE(pick_p2p(F) != nullptr, "train_steps_p2p: factors=%d unsupported (multiple of 4, at most 128)", F);
    if (n_steps == 0) return DRB_OK;
    const P2PLayout L = p2p_layout(I, F);
    char *self = (char *)h_peer_bufs[rank];
    StepParams p;
    // local batches are ragged (d_step_offsets): fill_params' uniform-batch range check does not apply, p.n is set below
    int rc = fill_params(p, d_P_local, (float *)(self + L.q), d_ws, U_local, I, F, d_bu, d_bi, d_bj,
                         (first_step + n_steps) * batch_per_rank + 1, batch_per_rank, first_step, n_steps, hyper, steps_done,
                         d_step_loss, 1);
    if (rc != DRB_OK) return rc;
    p.n = n_local;
    p.step_offsets = (const long long *)d_step_offsets;
    p.dense_hint = 1;
    const long long table_rows = (long long)U_local + I;
    const bool lean_k = step_params_lean(p) && lean_enabled(F, table_rows);
    P2PKernel k = pick_p2p(F, lean_k, table_rows);
    P2PParams x;
    for (int q = 0; q < kMaxPeers; ++q) x.peer[q] = q < world ? (char *)h_peer_bufs[q] : nullptr;
    x.rank = rank;
    x.world = world;
    for (int q = 0; q <= kMaxPeers; ++q) x.item_lo[q] = q <= world ? (int)((long long)I * q / world) : I;
    x.seq0 = (unsigned long long)steps_done;
    x.timeout_ns = (unsigned long long)((peer_timeout_s > 0 ? peer_timeout_s : 20.0) * 1e9);
    for (int b = 0; b < 2; ++b) { x.off_acc[b] = L.acc[b]; x.off_gq[b] = L.gq[b]; x.off_cnt[b] = L.cnt[b]; }
    x.off_q = L.q;

    static thread_local P2PKernel cached_k = nullptr;
    static thread_local int cached_per_sm = 0;
    if (cached_k != k) {
        int q = 0;
        DRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k, kThreads, 0));
        cached_k = k;
        cached_per_sm = q;
    }
    DRB_REQUIRE(cached_per_sm > 0, "p2p step kernel does not fit on an SM");
    const int max_grid = cached_per_sm * sm_count();
    const int tile = pick_tile((batch_per_rank + max_grid - 1) / max_grid, lean_k ? lean_tile_cap(F, table_rows) : kTileDefault);
    p.tile = tile;
    cudaStream_t st = (cudaStream_t)stream;
    DRB_CUDA(cudaMemsetAsync(p.ws.hdr, 0, sizeof(WsHeader), st));
    void *args[] = {&p, &x};
    DRB_CUDA(cudaLaunchCooperativeKernel((void *)k, dim3(max_grid), dim3(kThreads), args, 0, st));
    if (sync_and_check) {
        WsHeader h;
        DRB_CUDA(cudaMemcpyAsync(&h, d_ws, sizeof(WsHeader), cudaMemcpyDeviceToHost, st));
        DRB_CUDA(cudaStreamSynchronize(st));
        if (bad_step) *bad_step = h.status != 0 ? h.nan_step : -1;
        if (h.status == DRB_ERR_NAN_LOSS) {
            set_error("Loss=Nan or Infinity at step %lld: current settings does not fit the recommender", h.nan_step);
            return DRB_ERR_NAN_LOSS;
        }
        if (h.status == DRB_ERR_PEER) {
            set_error("peer exchange timed out at step %lld (a rank did not reach the rendezvous)", h.nan_step);
            return DRB_ERR_PEER;
        }
    }
    return DRB_OK;
}
