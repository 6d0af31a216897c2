// mf_bpr.cu -- the BPR-MF training step as ONE persistent cooperative sm_100a kernel.
//
// Stands behind GeneralRecommender.fit's step loop (daisy/model/AbstractRecommender.py:112-128)
// with MF.calc_loss (daisy/model/MFRecommender.py:70-97), BPRLoss (daisy/utils/loss.py:11),
// autograd's embedding backward (:125) and optim.SGD/Adam.step (:126, :53-56).
//
// Synchronous-step semantics (every gradient of a step is taken at the PRE-step weights, the
// Frobenius norms couple the whole batch) are kept exactly, without ever materialising the
// reference's table-sized dense gradient:
//
//   phase 1 (read-only on P,Q)   per triple (u,i,j): index tile staged by TMA (cp.async.bulk) into
//       shared memory; a group of W lanes gathers the three factor rows with 128-bit L2 loads,
//       reduces the two dot products with xor-shuffles (canonical order), evaluates
//       s = sigmoid(x), loss = -log(1e-10+s), c = -s(1-s)/(1e-10+s) and issues vector
//       RED.ADD.F32x4 reductions of the BPR part of the gradient into the L2-resident
//       accumulators gP/gQ:  gP[u] += c(q_i-q_j), gQ[i] += c p_u, gQ[j] -= c p_u; it also counts how
//       often each row occurs (cntU, cntI = pos | neg<<32) and accumulates the six batch norms.
//   -- grid barrier --           (norms and loss are now final; nobody reads P,Q any more)
//   phase 2                      every touched row is applied exactly once:
//       g = gP[r] + cnt * (reg_1 sgn(theta) + reg_2 theta / ||.||_F);  theta -= lr g  (or Adam);
//       the accumulator row and its counter are reset for the next step.  Rows are found either
//       by a dense sweep (large batches: every row is touched) or by claiming the counter with
//       atomicExch from the triple that touched it (small batches).
//   -- grid barrier --           next step.
//
// A NaN loss (ValueError in the reference, :122-123) stops the loop before the update of that step.
//
// Variants of the same kernel: GEN = false is the BPR-only hot instantiation, GEN = true selects HingeLoss / TOP1Loss
// (daisy/utils/loss.py:16-33) or the point-wise CL / SL branch (MFRecommender.py:75-81: the third plane holds the label,
// only P_u and Q_i take part) at run time; `phases` splits it into phase-1 / phase-2 launches (multi-GPU exchange,
// LightGCN, NeuMF); `neg_row_ptr` switches on the fused sampler (a fresh negative per triple drawn inside phase 1).
#include <math.h>

#include "step.cuh"

namespace drb {

constexpr int kThreads = 256;
constexpr int kTileMax = 512;  // triples per staged index tile
#ifndef DRB_MINB
#define DRB_MINB 2             // resident CTAs per SM the register allocator must allow
#endif
#ifndef DRB_UNR
#define DRB_UNR 2              // triples in flight per lane group (memory-level parallelism)
#endif

// ------------------------------------------------------------------ device pieces
__device__ __forceinline__ float sgnf(float x) { return (float)((x > 0.f) - (x < 0.f)); }

struct Norms {
    float inv_u, inv_i, inv_j;  // 1/||.||_F, 0 when the norm is 0 (zero subgradient)
};

struct AdamCoef {
    float step_size, bc2_sqrt;
};

// Apply the accumulated gradient of ONE table row (all W lanes of the group cooperate).
// cnt_a / cnt_b: occurrences weighted by inv_a / inv_b (user rows: cnt_b = 0).
template <int VEC, int W, int NCH, int OPT>
__device__ __forceinline__ void apply_row(float *theta_row, float *g_row, float *m_row, float *v_row, int gl,
                                          int chunks, float cnt_a, float inv_a, float cnt_b, float inv_b,
                                          const StepParams &p, const AdamCoef &ac, bool touched)
{
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        int c = gl + ch * W;
        if (c >= chunks) continue;
        float *tp = theta_row + c * VEC;
        Vec<VEC> th = ld_row<VEC>(tp);
        Vec<VEC> g;
        if (touched) {
            g = ld_row<VEC>(g_row + c * VEC);
            Vec<VEC> z;
#pragma unroll
            for (int e = 0; e < VEC; ++e) z.v[e] = 0.f;
            st_row<VEC>(g_row + c * VEC, z);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) g.v[e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            float t = th.v[e];
            float gg = g.v[e];
            if (touched) {
                float sg = p.reg1 * sgnf(t);
                gg += cnt_a * (sg + p.reg2 * t * inv_a) + cnt_b * (sg + p.reg2 * t * inv_b);
            }
            g.v[e] = gg;
        }
        if constexpr (OPT == DRB_OPT_SGD) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) th.v[e] = th.v[e] - p.lr * g.v[e];
        } else {
            Vec<VEC> m = ld_row<VEC>(m_row + c * VEC), v = ld_row<VEC>(v_row + c * VEC);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                float gk = g.v[e];
                m.v[e] = m.v[e] + (gk - m.v[e]) * (1.f - p.beta1);
                v.v[e] = v.v[e] * p.beta2 + (1.f - p.beta2) * gk * gk;
                float denom = sqrtf(v.v[e]) / ac.bc2_sqrt + p.eps;
                th.v[e] = th.v[e] - ac.step_size * (m.v[e] / denom);
            }
            st_row<VEC>(m_row + c * VEC, m);
            st_row<VEC>(v_row + c * VEC, v);
        }
        st_row<VEC>(tp, th);
    }
}

// Dense phase-2 sweep: lane groups walk ALL rows of P then Q, R rows in flight each.  Counter, theta
// and gradient accumulator of the R rows are loaded unconditionally and up front (one memory round
// trip instead of three dependent ones); an untouched SGD row has cnt == 0 and g == 0, so nothing is
// written for it.  Adam moves every row (dense optimiser semantics of the reference).  Adagrad / RMSprop
// (AbstractRecommender.py:57-60, torch defaults) keep ONE state row in the m slot: Adagrad leaves an untouched row
// alone (g = 0 adds nothing), RMSprop's running square of an untouched row still decays by alpha.
template <int VEC, int W, int NCH, int OPT>
__device__ __forceinline__ void dense_sweep(const StepParams &p, const Norms &nm, const AdamCoef &ac, int gl, int group,
                                            int groups_per_cta, int chunks)
{
    constexpr int R = (OPT == DRB_OPT_SGD) ? ((NCH * VEC <= 4) ? 4 : 2) : ((NCH * VEC <= 4) ? 2 : 1);
    const long long rows = (long long)p.U + p.I;
    const long long tg = (long long)gridDim.x * groups_per_cta;
    const int F = p.F;
    for (long long r0 = (long long)blockIdx.x * groups_per_cta + group; r0 < rows; r0 += tg * R) {
        float *th_p[R], *g_p[R], *m_p[R], *v_p[R];
        unsigned long long cnt[R];
        bool act[R], is_user[R];
        Row<VEC, W, NCH> th[R], g[R], m[R], v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            long long r = r0 + (long long)k * tg;
            act[k] = r < rows;
            is_user[k] = r < p.U;
            long long it = is_user[k] ? r : r - p.U;
            size_t o = (size_t)(act[k] ? it : 0) * F;
            th_p[k] = (is_user[k] ? p.P : p.Q) + o;
            g_p[k] = (is_user[k] ? p.ws.gP : p.ws.gQ) + o;
            cnt[k] = 0;
            if (act[k]) cnt[k] = is_user[k] ? (unsigned long long)__ldcg(p.ws.cntU + it) : __ldcg(p.ws.cntI + it);
            th[k] = load_row<VEC, W, NCH>(th_p[k], gl, chunks, act[k]);
            g[k] = load_row<VEC, W, NCH>(g_p[k], gl, chunks, act[k]);
            if constexpr (OPT != DRB_OPT_SGD) {
                m_p[k] = (is_user[k] ? p.ws.mP : p.ws.mQ) + o;
                m[k] = load_row<VEC, W, NCH>(m_p[k], gl, chunks, act[k]);
            }
            if constexpr (OPT == DRB_OPT_ADAM) {
                v_p[k] = (is_user[k] ? p.ws.vP : p.ws.vQ) + o;
                v[k] = load_row<VEC, W, NCH>(v_p[k], gl, chunks, act[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const bool touched = cnt[k] != 0;
            if (!act[k] || ((OPT == DRB_OPT_SGD || OPT == DRB_OPT_ADAGRAD) && !touched && !p.dense_grad)) continue;
            const float ca = (float)(unsigned)(cnt[k] & 0xffffffffull), cb = p.neg_mult * (float)(unsigned)(cnt[k] >> 32);
            const float ia = is_user[k] ? nm.inv_u : nm.inv_i, ib = nm.inv_j;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                int c = gl + ch * W;
                if (c >= chunks) continue;
                Vec<VEC> &t = th[k].c[ch];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    float x = t.v[e], gg = p.gscale * g[k].c[ch].v[e];
                    if (touched) {
                        float sg = p.reg1 * sgnf(x);
                        gg += ca * (sg + p.reg2 * x * ia) + cb * (sg + p.reg2 * x * ib);
                    }
                    if constexpr (OPT == DRB_OPT_SGD) {
                        t.v[e] = x - p.lr * gg;
                    } else if constexpr (OPT == DRB_OPT_ADAGRAD) {   // sum += g^2; theta -= lr g / (sqrt(sum) + 1e-10)
                        float ss = m[k].c[ch].v[e] + gg * gg;
                        t.v[e] = x - p.lr * (gg / (sqrtf(ss) + 1e-10f));
                        m[k].c[ch].v[e] = ss;
                    } else if constexpr (OPT == DRB_OPT_RMSPROP) {   // sq = .99 sq + .01 g^2; theta -= lr g / (sqrt(sq) + 1e-8)
                        float sq = m[k].c[ch].v[e] * 0.99f + (1.f - 0.99f) * gg * gg;
                        t.v[e] = x - p.lr * (gg / (sqrtf(sq) + 1e-8f));
                        m[k].c[ch].v[e] = sq;
                    } else {
                        float mm = m[k].c[ch].v[e], vv = v[k].c[ch].v[e];
                        mm = mm + (gg - mm) * (1.f - p.beta1);
                        vv = vv * p.beta2 + (1.f - p.beta2) * gg * gg;
                        float denom = sqrtf(vv) / ac.bc2_sqrt + p.eps;
                        t.v[e] = x - ac.step_size * (mm / denom);
                        m[k].c[ch].v[e] = mm;
                        v[k].c[ch].v[e] = vv;
                    }
                }
                st_row<VEC>(th_p[k] + c * VEC, t);
                if constexpr (OPT != DRB_OPT_SGD) st_row<VEC>(m_p[k] + c * VEC, m[k].c[ch]);
                if constexpr (OPT == DRB_OPT_ADAM) st_row<VEC>(v_p[k] + c * VEC, v[k].c[ch]);
                if (touched) {
                    Vec<VEC> z;
#pragma unroll
                    for (int e = 0; e < VEC; ++e) z.v[e] = 0.f;
                    st_row<VEC>(g_p[k] + c * VEC, z);
                }
            }
            if (touched && gl == 0 && !p.keep_counts) {
                long long r = r0 + (long long)k * tg;
                if (is_user[k]) p.ws.cntU[r] = 0u; else p.ws.cntI[r - p.U] = 0ull;
            }
        }
    }
}

// Fresh uniform negative for (user u, global triple index gt, step): a Philox word scaled to [0, n_comp) by multiply-high,
// then the k-th item missing from the user's sorted row: item = k + #{s : col[s] - s <= k} (one binary search).
__device__ __forceinline__ int draw_negative(const StepParams &p, int u, unsigned long long gt, unsigned long long step)
{
    const long long rb = p.neg_row_ptr[u], re = p.neg_row_ptr[u + 1];
    const unsigned n_comp = (unsigned)((long long)p.I - (re - rb));
    uint32_t c[4] = {(uint32_t)gt, (uint32_t)(gt >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
    philox4x32(c, (uint32_t)p.neg_seed, (uint32_t)(p.neg_seed >> 32));
    const int k = (int)__umulhi(c[0], n_comp);
    long long lo = 0, hi = re - rb;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if ((long long)__ldg(p.neg_col + rb + mid) - mid <= (long long)k) lo = mid + 1; else hi = mid;
    }
    const int item = k + (int)lo;
    return item < p.I ? item : p.I - 1;   // only reachable for a user who interacted with every item (rejected by the host)
}

template <int VEC, int W, int NCH, bool GEN>
__global__ void __launch_bounds__(kThreads, DRB_MINB) mf_bpr_steps_kernel(StepParams p)
{
    constexpr int GPW = 32 / W;                  // lane groups per warp
    constexpr int GROUPS = (kThreads / 32) * GPW;  // lane groups per CTA
    constexpr int UNR = (NCH * VEC <= 4) ? DRB_UNR : 1;  // triples in flight per group

    __shared__ __align__(128) int32_t s_idx[2][3][kTileMax];
    __shared__ uint64_t s_bar[2];
    __shared__ double s_red[8][kThreads / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gl = lane % W, gw = lane / W;
    const int group = warp * GPW + gw;
    const int chunks = p.F / VEC;
    const int F = p.F;
    WsHeader *hdr = p.ws.hdr;

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (*(volatile int *)&hdr->status != 0) return;   // split mode: a previous step already raised NaN
    uint32_t par0 = 0, par1 = 0;
    unsigned long long epoch = 0;
    const int tile = p.tile;
    const bool pw = GEN && p.loss >= DRB_LOSS_CL;   // point-wise: bj is the label plane, no negative row

    // stage one index tile: TMA bulk copy when full and 16-byte aligned, plain loads otherwise
    auto stage = [&](long long tbase, int cnt, int b) {
        const int32_t *su = p.bu + tbase, *si = p.bi + tbase, *sj = p.bj + tbase;
        bool bulk = (cnt % 4 == 0) && ((((uintptr_t)su | (uintptr_t)si | (uintptr_t)sj) & 15) == 0);
        if (bulk) {
            if (tid == 0) {
                uint32_t bytes = (uint32_t)cnt * 4u;
                mbar_expect_tx(&s_bar[b], 3u * bytes);
                tma_load_1d(&s_idx[b][0][0], su, bytes, &s_bar[b]);
                tma_load_1d(&s_idx[b][1][0], si, bytes, &s_bar[b]);
                tma_load_1d(&s_idx[b][2][0], sj, bytes, &s_bar[b]);
            }
        } else {
            for (int k = tid; k < cnt; k += kThreads) {
                s_idx[b][0][k] = __ldg(su + k);
                s_idx[b][1][k] = __ldg(si + k);
                s_idx[b][2][k] = __ldg(sj + k);
            }
            __syncthreads();
            if (tid == 0) mbar_arrive(&s_bar[b]);
        }
    };

    for (long long s = 0; s < p.n_steps; ++s) {
        const long long step = p.first_step + s;
        const long long base = step * p.batch;
        const long long nb = min(p.batch, p.n - base);
        const long long ntiles = (nb + tile - 1) / tile;
        double *acc = hdr->acc[s & 1];
        const bool has_reg = (p.reg1 != 0.f) || (p.reg2 != 0.f);

        // ------------------------------------------------------------ phase 1
        if (p.phases & 1) {
        if (tid < 8 * (kThreads / 32)) (&s_red[0][0])[tid] = 0.0;   // per-warp fp64 accumulators of this step
        __syncthreads();
        int buf = 0;
        long long t_i = blockIdx.x;
        if (t_i < ntiles) stage(base + t_i * tile, (int)min((long long)tile, nb - t_i * tile), 0);
        for (; t_i < ntiles; t_i += gridDim.x) {
            long long t_n = t_i + gridDim.x;
            if (t_n < ntiles) stage(base + t_n * tile, (int)min((long long)tile, nb - t_n * tile), buf ^ 1);
            if (buf == 0) { mbar_wait(&s_bar[0], par0); par0 ^= 1; } else { mbar_wait(&s_bar[1], par1); par1 ^= 1; }
            const int cnt = (int)min((long long)tile, nb - t_i * tile);
            const int32_t *xu = s_idx[buf][0], *xi = s_idx[buf][1], *xj = s_idx[buf][2];
            float t_loss = 0.f, t_l1u = 0.f, t_l1i = 0.f, t_l1j = 0.f, t_s2u = 0.f, t_s2i = 0.f, t_s2j = 0.f, t_gb0 = 0.f;

            for (int tb = 0; tb < cnt; tb += GROUPS * UNR) {
                Row<VEC, W, NCH> rp[UNR], rqi[UNR], rqj[UNR];
                int iu[UNR], ii[UNR], ij[UNR];
                float lab[UNR];
                bool ok[UNR];
#pragma unroll
                for (int r = 0; r < UNR; ++r) {
                    int t = tb + r * GROUPS + group;
                    ok[r] = t < cnt;
                    iu[r] = ok[r] ? xu[t] : 0;
                    ii[r] = ok[r] ? xi[t] : 0;
                    ij[r] = ok[r] ? xj[t] : 0;
                    lab[r] = 0.f;
                    if (pw) {                       // label = batch[2].float() (MFRecommender.py:76); the j row stays zero
                        lab[r] = (float)ij[r];
                        ij[r] = 0;
                    }
                    if (p.neg_row_ptr != nullptr && ok[r]) {
                        const long long gt = base + t_i * tile + t;            // position of the triple in the planes
                        ij[r] = draw_negative(p, iu[r], (unsigned long long)gt, (unsigned long long)step);
                        if (p.neg_out != nullptr && gl == 0) p.neg_out[gt] = ij[r];
                    }
                    rp[r] = load_row<VEC, W, NCH>(p.P + (size_t)iu[r] * F, gl, chunks, ok[r]);
                    rqi[r] = load_row<VEC, W, NCH>(p.Q + (size_t)ii[r] * F, gl, chunks, ok[r]);
                    rqj[r] = load_row<VEC, W, NCH>(p.Q + (size_t)ij[r] * F, gl, chunks, ok[r] && !pw);
                }
                // scores of the UNR triples of this group (every lane of the group ends up with the same values)
                float ps[UNR], ns[UNR], cs[UNR], cn[UNR];
#pragma unroll
                for (int r = 0; r < UNR; ++r) {
                    ps[r] = dot_rows<VEC, W, NCH>(rp[r], rqi[r]);
                    ns[r] = dot_rows<VEC, W, NCH>(rp[r], rqj[r]);
                    if (GEN && p.bias != nullptr) {   // FM: pred += (u_bias(user) + i_bias(item)) + bias_  (FMRecommender.py:66-67)
                        const float ub = __ldcg(p.bias + iu[r]), b0 = __ldcg(p.bias + p.U + p.I);
                        ps[r] += (ub + __ldcg(p.bias + p.U + ii[r])) + b0;
                        ns[r] += (ub + __ldcg(p.bias + p.U + ij[r])) + b0;
                    }
                    if (pw) ns[r] = lab[r];         // pair_loss receives the label in place of the negative score
                }
                // The scalar chain (sigmoid -> log -> coefficient, ~40 instructions) would be replayed by all W lanes for
                // each of the UNR triples; instead lane gl evaluates it ONCE, for triple (gl % UNR) of its group, and the
                // coefficients d(loss)/d(pos), d(loss)/d(neg) are handed round with shuffles.
                auto pair_loss = [&](float pos, float neg, float &c_pos, float &c_neg) -> float {
                    if (GEN && p.loss == DRB_LOSS_CL) {     // BCEWithLogitsLoss(sum): (1-y) x - log_sigmoid(x), neg = y
                        const float z = expf(-fabsf(pos));
                        const float logsig = fminf(pos, 0.f) - log1pf(z);
                        const float dls = pos < 0.f ? 1.f - z / (1.f + z) : z / (1.f + z);
                        c_pos = (1.f - neg) - dls;
                        c_neg = 0.f;
                        return (1.f - neg) * pos - logsig;
                    }
                    if (GEN && p.loss == DRB_LOSS_SL) {     // MSELoss(sum): (x - y)^2, neg = y
                        const float d = pos - neg;
                        c_pos = 2.f * d;
                        c_neg = 0.f;
                        return d * d;
                    }
                    if (GEN && p.loss == DRB_LOSS_HL) {     // clamp(1 - (pos - neg), min=0); clamp's backward passes at equality
                        const float m = 1.f - (pos - neg);
                        c_pos = (m >= 0.f) ? -1.f : 0.f;
                        c_neg = -c_pos;
                        return m > 0.f ? m : 0.f;
                    }
                    if (GEN && p.loss == DRB_LOSS_TL) {     // sigmoid(neg - pos) + sigmoid(neg^2)
                        const float s1 = 1.f / (1.f + expf(-(neg - pos))), s2 = 1.f / (1.f + expf(-(neg * neg)));
                        c_pos = -(s1 * (1.f - s1));
                        c_neg = s1 * (1.f - s1) + s2 * (1.f - s2) * 2.f * neg;
                        return s1 + s2;
                    }
                    const float x = pos - neg;
                    const float sg = 1.f / (1.f + expf(-x));
                    c_pos = -(sg * (1.f - sg)) / (1e-10f + sg);
                    c_neg = -c_pos;
                    return -logf(1e-10f + sg);
                };
                if constexpr (W >= UNR) {
                    float p_own = ps[0], n_own = ns[0];
                    bool ok_own = ok[0];
#pragma unroll
                    for (int r = 1; r < UNR; ++r)
                        if ((gl % UNR) == r) { p_own = ps[r]; n_own = ns[r]; ok_own = ok[r]; }
                    float cp_own, cn_own;
                    const float l_own = pair_loss(p_own, n_own, cp_own, cn_own);
                    if (gl < UNR && ok_own) t_loss += l_own;
#pragma unroll
                    for (int r = 0; r < UNR; ++r) {
                        cs[r] = __shfl_sync(0xffffffffu, cp_own, (lane - gl) + r);
                        cn[r] = GEN ? __shfl_sync(0xffffffffu, cn_own, (lane - gl) + r) : -cs[r];
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < UNR; ++r) {
                        const float l = pair_loss(ps[r], ns[r], cs[r], cn[r]);
                        if (gl == 0 && ok[r]) t_loss += l;
                    }
                }
#pragma unroll
                for (int r = 0; r < UNR; ++r) {
                    if (!ok[r]) continue;
                    const float c = cs[r];
                    if (has_reg) {
                        float l1u = 0, l1i = 0, l1j = 0, s2u = 0, s2i = 0, s2j = 0;
                        Row<VEC, W, NCH> nu_ = rp[r], ni_ = rqi[r], nj_ = rqj[r];
                        if (p.Pn != nullptr) {   // regulariser on the ego rows (LightGCNRecommender.py:145-146,159)
                            nu_ = load_row<VEC, W, NCH>(p.Pn + (size_t)iu[r] * F, gl, chunks, true);
                            ni_ = load_row<VEC, W, NCH>(p.Qn + (size_t)ii[r] * F, gl, chunks, true);
                            nj_ = load_row<VEC, W, NCH>(p.Qn + (size_t)ij[r] * F, gl, chunks, true);
                        }
#pragma unroll
                        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                            for (int e = 0; e < VEC; ++e) {
                                float a = nu_.c[ch].v[e], b = ni_.c[ch].v[e], d = nj_.c[ch].v[e];
                                l1u += fabsf(a); s2u = fmaf(a, a, s2u);
                                l1i += fabsf(b); s2i = fmaf(b, b, s2i);
                                l1j += fabsf(d); s2j = fmaf(d, d, s2j);
                            }
                        t_l1u += l1u; t_l1i += l1i; t_l1j += l1j;
                        t_s2u += s2u; t_s2i += s2i; t_s2j += s2j;
                    }
                    if (p.apply) {
#pragma unroll
                        for (int ch = 0; ch < NCH; ++ch) {
                            int cc = gl + ch * W;
                            if (cc >= chunks) continue;
                            Vec<VEC> gu, gi, gj;
#pragma unroll
                            for (int e = 0; e < VEC; ++e) {
                                if (!GEN || p.loss == DRB_LOSS_BPR) {   // c_neg == -c_pos: the reference's BPR arithmetic
                                    gu.v[e] = c * (rqi[r].c[ch].v[e] - rqj[r].c[ch].v[e]);
                                    gi.v[e] = c * rp[r].c[ch].v[e];
                                    gj.v[e] = -gi.v[e];
                                } else {
                                    gu.v[e] = c * rqi[r].c[ch].v[e] + cn[r] * rqj[r].c[ch].v[e];
                                    gi.v[e] = c * rp[r].c[ch].v[e];
                                    gj.v[e] = cn[r] * rp[r].c[ch].v[e];
                                }
                            }
                            red_row<VEC>(p.ws.gP + (size_t)iu[r] * F + cc * VEC, gu);
                            red_row<VEC>(p.ws.gQ + (size_t)ii[r] * F + cc * VEC, gi);
                            if (!pw) red_row<VEC>(p.ws.gQ + (size_t)ij[r] * F + cc * VEC, gj);
                        }
                        if (gl == 0) {
                            red_add_u32(p.ws.cntU + iu[r], 1u);
                            red_add_u64(p.ws.cntI + ii[r], 1ull);
                            if (!pw) red_add_u64(p.ws.cntI + ij[r], 1ull << 32);
                            if (GEN && p.bias != nullptr) {   // d loss / d (u_bias, i_bias, bias_): no regulariser (:76-95)
                                const float cboth = pw ? c : c + cn[r];
                                asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p.ws.gB + iu[r]), "f"(cboth) : "memory");
                                asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p.ws.gB + p.U + ii[r]), "f"(c) : "memory");
                                if (!pw)
                                    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p.ws.gB + p.U + ij[r]), "f"(cn[r]) : "memory");
                                t_gb0 += cboth;
                            }
                        }
                    }
                }
            }
            // per-thread fp32 partials cover <= tile/GROUPS triples: warp-reduce, widen to fp64 in smem
            {
                float tv[8] = {t_loss, t_l1u, t_l1i, t_l1j, t_s2u, t_s2i, t_s2j, t_gb0};
                const int nv = has_reg ? 7 : 1;
                for (int k = 0; k < 8; ++k) {
                    if (k >= nv && !(GEN && k == 7 && p.bias != nullptr)) continue;
                    float v = tv[k];
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    if (lane == 0) s_red[k][warp] += (double)v;
                }
            }
            __syncthreads();  // tile buffer free for re-staging
            buf ^= 1;
        }
        // CTA reduction of the 7 partial sums -> one fp64 atomic each
        __syncthreads();
        if (tid < (has_reg ? 7 : 1) || (GEN && tid == 7 && p.bias != nullptr)) {
            double v = 0;
            for (int w = 0; w < kThreads / 32; ++w) v += s_red[tid][w];
            if (v != 0.0) atomicAdd(&acc[tid], v);
        }
        }  // phase 1
        if (p.phases == 3) grid_barrier(&hdr->barrier, epoch);
        if (!(p.phases & 2)) break;   // split mode: the host reduces gQ / counters / acc across ranks now

        // ------------------------------------------------------------ phase 2
        double bpr, l1u, l1i, l1j, s2u, s2i, s2j;
        {
            const volatile double *va = acc;
            bpr = va[0]; l1u = va[1]; l1i = va[2]; l1j = va[3]; s2u = va[4]; s2i = va[5]; s2j = va[6];
        }
        double nu = sqrt(s2u), ni = sqrt(s2i), nj = sqrt(s2j);
        // fp32 assembly of the scalar loss, in the reference's order (MFRecommender.py:88-95)
        float loss = (float)bpr;
        loss += p.reg1 * ((float)l1i + (float)l1j);
        loss += p.reg2 * ((float)ni + (float)nj);
        loss += p.reg1 * (float)l1u;
        loss += p.reg2 * (float)nu;
        if (blockIdx.x == 0 && tid == 0) p.step_loss[s] = (double)loss;
        if (blockIdx.x == 0 && tid < 8) hdr->acc[(s + 1) & 1][tid] = 0.0;  // recycle the other accumulator
        if (isnan(loss)) {
            if (blockIdx.x == 0 && tid == 0) {
                hdr->status = DRB_ERR_NAN_LOSS;
                hdr->nan_step = step;
            }
            break;  // uniform across the grid: every CTA computed the same loss
        }
        if (p.apply) {
            Norms nm;
            nm.inv_u = nu > 0 ? (float)(1.0 / nu) : 0.f;
            nm.inv_i = ni > 0 ? (float)(1.0 / ni) : 0.f;
            nm.inv_j = nj > 0 ? (float)(1.0 / nj) : 0.f;
            AdamCoef ac;
            ac.step_size = 0.f;
            ac.bc2_sqrt = 1.f;
            if (p.opt == DRB_OPT_ADAM) {
                double t = (double)(p.adam_step0 + s + 1);
                ac.step_size = (float)((double)p.lr / (1.0 - pow((double)p.beta1, t)));
                ac.bc2_sqrt = (float)sqrt(1.0 - pow((double)p.beta2, t));
            }
            const bool dense = p.dense_hint >= 0 ? (p.dense_hint != 0)
                                                 : ((p.opt != DRB_OPT_SGD) || (3 * nb >= ((long long)p.U + p.I) / 4));
            if (dense || p.opt != DRB_OPT_SGD) {   // stateful optimisers always sweep (claim mode is SGD only)
                if (p.opt == DRB_OPT_SGD)
                    dense_sweep<VEC, W, NCH, DRB_OPT_SGD>(p, nm, ac, gl, group, GROUPS, chunks);
                else if (p.opt == DRB_OPT_ADAM)
                    dense_sweep<VEC, W, NCH, DRB_OPT_ADAM>(p, nm, ac, gl, group, GROUPS, chunks);
                else if constexpr (GEN) {          // launch_steps routes these two to the GEN instantiation
                    if (p.opt == DRB_OPT_ADAGRAD)
                        dense_sweep<VEC, W, NCH, DRB_OPT_ADAGRAD>(p, nm, ac, gl, group, GROUPS, chunks);
                    else
                        dense_sweep<VEC, W, NCH, DRB_OPT_RMSPROP>(p, nm, ac, gl, group, GROUPS, chunks);
                }
            } else {
                // claim mode (SGD only): the first group to swap a row's counter to zero applies it
                for (long long t0 = (long long)blockIdx.x * tile; t0 < nb; t0 += (long long)gridDim.x * tile) {
                    const int cnt = (int)min((long long)tile, nb - t0);
                    for (int tb = 0; tb < cnt; tb += GROUPS) {
                        int t = tb + group;
                        bool ok = t < cnt;
                        int u = 0, i = 0, j = 0;
                        if (ok) {
                            u = __ldg(p.bu + base + t0 + t);
                            i = __ldg(p.bi + base + t0 + t);
                            j = pw ? i : __ldg(p.bj + base + t0 + t);   // point-wise: that plane holds labels
                        }
                        unsigned cu = 0;
                        unsigned long long ci = 0, cj = 0;
                        if (ok && gl == 0) {
                            cu = atomicExch(p.ws.cntU + u, 0u);
                            ci = atomicExch(p.ws.cntI + i, 0ull);
                            cj = atomicExch(p.ws.cntI + j, 0ull);
                        }
                        cu = __shfl_sync(0xffffffffu, cu, gw * W);
                        ci = __shfl_sync(0xffffffffu, ci, gw * W);
                        cj = __shfl_sync(0xffffffffu, cj, gw * W);
                        if (cu != 0) {
                            size_t o = (size_t)u * F;
                            apply_row<VEC, W, NCH, DRB_OPT_SGD>(p.P + o, p.ws.gP + o, nullptr, nullptr, gl, chunks, (float)cu,
                                                                nm.inv_u, 0.f, 0.f, p, ac, true);
                        }
                        if (ci != 0) {
                            size_t o = (size_t)i * F;
                            apply_row<VEC, W, NCH, DRB_OPT_SGD>(p.Q + o, p.ws.gQ + o, nullptr, nullptr, gl, chunks,
                                                                (float)(unsigned)(ci & 0xffffffffull), nm.inv_i,
                                                                (float)(unsigned)(ci >> 32), nm.inv_j, p, ac, true);
                        }
                        if (cj != 0) {
                            size_t o = (size_t)j * F;
                            apply_row<VEC, W, NCH, DRB_OPT_SGD>(p.Q + o, p.ws.gQ + o, nullptr, nullptr, gl, chunks,
                                                                (float)(unsigned)(cj & 0xffffffffull), nm.inv_i,
                                                                (float)(unsigned)(cj >> 32), nm.inv_j, p, ac, true);
                        }
                    }
                }
            }
        }
        if (GEN && p.apply && p.bias != nullptr) {
            // FM's U + I + 1 first-order scalars: the same optimiser switch, no regulariser; the accumulator is cleared
            const double gb0 = ((const volatile double *)acc)[7];
            float step_size = 0.f, bc2_sqrt = 1.f;
            if (p.opt == DRB_OPT_ADAM) {
                double t = (double)(p.adam_step0 + s + 1);
                step_size = (float)((double)p.lr / (1.0 - pow((double)p.beta1, t)));
                bc2_sqrt = (float)sqrt(1.0 - pow((double)p.beta2, t));
            }
            const long long nbias = (long long)p.U + p.I + 1;
            for (long long k = (long long)blockIdx.x * kThreads + tid; k < nbias; k += (long long)gridDim.x * kThreads) {
                const float g = (k == nbias - 1) ? (float)gb0 : __ldcg(p.ws.gB + k);
                float th = __ldcg(p.bias + k);
                if (p.opt == DRB_OPT_SGD) {
                    th = th - p.lr * g;
                } else if (p.opt == DRB_OPT_ADAGRAD) {
                    const float ss = __ldcg(p.ws.mB + k) + g * g;
                    th = th - p.lr * (g / (sqrtf(ss) + 1e-10f));
                    __stcg(p.ws.mB + k, ss);
                } else if (p.opt == DRB_OPT_RMSPROP) {
                    const float sq = __ldcg(p.ws.mB + k) * 0.99f + (1.f - 0.99f) * g * g;
                    th = th - p.lr * (g / (sqrtf(sq) + 1e-8f));
                    __stcg(p.ws.mB + k, sq);
                } else {
                    float mm = __ldcg(p.ws.mB + k), vv = __ldcg(p.ws.vB + k);
                    mm = mm + (g - mm) * (1.f - p.beta1);
                    vv = vv * p.beta2 + (1.f - p.beta2) * g * g;
                    th = th - step_size * (mm / (sqrtf(vv) / bc2_sqrt + p.eps));
                    __stcg(p.ws.mB + k, mm);
                    __stcg(p.ws.vB + k, vv);
                }
                __stcg(p.bias + k, th);
                if (k != nbias - 1 && g != 0.f) __stcg(p.ws.gB + k, 0.f);
            }
        }
        if (s + 1 < p.n_steps) grid_barrier(&hdr->barrier, epoch);
    }
}

// b?[k] = triples[perm[k], ?]
__global__ void gather_triples_kernel(const int32_t *__restrict__ triples, const int64_t *__restrict__ perm, long long n,
                                      int32_t *__restrict__ bu, int32_t *__restrict__ bi, int32_t *__restrict__ bj)
{
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        long long src = perm ? perm[k] : k;
        const int32_t *t = triples + 3 * src;
        bu[k] = __ldg(t);
        bi[k] = __ldg(t + 1);
        bj[k] = __ldg(t + 2);
    }
}

// ------------------------------------------------------------------ host dispatch
typedef void (*StepKernel)(StepParams);

template <int VEC, bool GEN>
static StepKernel pick_kernel_v(int W, int NCH)
{
#define DRB_CASE(w, n) \
    if (W == w && NCH == n) return mf_bpr_steps_kernel<VEC, w, n, GEN>;
    DRB_CASE(1, 1) DRB_CASE(2, 1) DRB_CASE(4, 1) DRB_CASE(8, 1) DRB_CASE(16, 1) DRB_CASE(32, 1)
    DRB_CASE(32, 2) DRB_CASE(32, 4) DRB_CASE(32, 8)
#undef DRB_CASE
    return nullptr;
}

// GEN = false: BPR only (the hot instantiation, no loss-kind branches); GEN = true: HL / TL selected at run time
static StepKernel pick_kernel(int F, bool gen)
{
    if (F <= 0) return nullptr;
    RowGeom g = row_geom(F);
    if (gen) {
        if (g.vec == 4) return pick_kernel_v<4, true>(g.width, g.nch);
        if (g.vec == 2) return pick_kernel_v<2, true>(g.width, g.nch);
        return pick_kernel_v<1, true>(g.width, g.nch);
    }
    if (g.vec == 4) return pick_kernel_v<4, false>(g.width, g.nch);
    if (g.vec == 2) return pick_kernel_v<2, false>(g.width, g.nch);
    return pick_kernel_v<1, false>(g.width, g.nch);
}

int launch_steps(StepParams &p, cudaStream_t st, bool keep_status)
{
    // GEN instantiation: any loss but BPR, and the Adagrad / RMSprop sweeps (kept out of the hot BPR + SGD/Adam kernel)
    StepKernel k = pick_kernel(p.F, p.loss != DRB_LOSS_BPR || p.opt > DRB_OPT_ADAM || p.bias != nullptr);
    DRB_REQUIRE(k != nullptr, "unsupported factors=%d (row too long for 32 lanes x 8 chunks)", p.F);
    // occupancy of the chosen instantiation, cached (the query costs microseconds and this runs once per step in the
    // split multi-GPU / LightGCN / NeuMF paths)
    static thread_local StepKernel cached_k = nullptr;
    static thread_local int cached_per_sm = 0;
    if (cached_k != k) {
        int q = 0;
        DRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k, kThreads, 0));
        cached_k = k;
        cached_per_sm = q;
    }
    const int per_sm = cached_per_sm;
    DRB_REQUIRE(per_sm > 0, "step kernel does not fit on an SM");
    const int max_grid = per_sm * sm_count();
    // tile: as large as possible (<= kTileMax) while still giving every CTA work
    long long want = (p.batch + max_grid - 1) / max_grid;
    int tile = (int)((want + 15) / 16 * 16);
    if (tile < 16) tile = 16;
    if (tile > kTileMax) tile = kTileMax;
    p.tile = tile;
    long long tiles = (p.batch + tile - 1) / tile;
    long long rows_work = ((long long)p.U + p.I + 63) / 64;
    bool dense = p.dense_hint >= 0 ? (p.dense_hint != 0)
                                   : ((p.opt != DRB_OPT_SGD) || (3 * p.batch >= ((long long)p.U + p.I) / 4));
    long long want_grid = (dense && p.apply) ? (tiles > rows_work ? tiles : rows_work) : tiles;
    int grid = (int)(want_grid < 1 ? 1 : (want_grid > max_grid ? max_grid : want_grid));
    if (p.phases & 1)
        DRB_CUDA(cudaMemsetAsync(p.ws.hdr, 0, (p.phases == 3 && !keep_status) ? sizeof(WsHeader) : kHdrResetBytes, st));
    void *args[] = {&p};
    DRB_CUDA(cudaLaunchCooperativeKernel((void *)k, dim3(grid), dim3(kThreads), args, 0, st));
    return DRB_OK;
}

int check_nan(void *d_ws, cudaStream_t st, int64_t *nan_step)
{
    WsHeader h;
    DRB_CUDA(cudaMemcpyAsync(&h, d_ws, sizeof(WsHeader), cudaMemcpyDeviceToHost, st));
    DRB_CUDA(cudaStreamSynchronize(st));
    if (h.status == DRB_ERR_NAN_LOSS) {
        if (nan_step) *nan_step = h.nan_step;
        set_error("Loss=Nan or Infinity at step %lld: current settings does not fit the recommender", h.nan_step);
        return DRB_ERR_NAN_LOSS;
    }
    if (nan_step) *nan_step = -1;
    return DRB_OK;
}

}  // namespace drb

using namespace drb;

extern "C" size_t drb_mf_workspace_bytes(int32_t U, int32_t I, int32_t F, int32_t opt)
{
    return carve(nullptr, U, I, F, opt, nullptr);
}

extern "C" int drb_mf_workspace_init(void *d_ws, int32_t U, int32_t I, int32_t F, int32_t opt, void *stream)
{
    DRB_REQUIRE(d_ws != nullptr && U > 0 && I > 0 && F > 0, "workspace_init: bad arguments");
    size_t bytes = carve(nullptr, U, I, F, opt, nullptr);
    DRB_CUDA(cudaMemsetAsync(d_ws, 0, bytes, (cudaStream_t)stream));
    return DRB_OK;
}

static int fill_params(StepParams &p, float *P, float *Q, void *d_ws, int U, int I, int F, const int32_t *bu,
                       const int32_t *bi, const int32_t *bj, long long n, long long batch, long long first, long long nsteps,
                       const drb_hyper *h, long long adam_step0, double *d_step_loss, int apply, float *d_bias = nullptr)
{
    DRB_REQUIRE(P && Q && d_ws && bu && bi && bj && h && d_step_loss, "null pointer argument");
    DRB_REQUIRE(U > 0 && I > 0 && F > 0 && batch > 0 && n >= 0 && first >= 0 && nsteps >= 0, "bad sizes");
    DRB_REQUIRE(h->opt >= DRB_OPT_SGD && h->opt <= DRB_OPT_RMSPROP, "unknown optimizer id %d", h->opt);
    DRB_REQUIRE(h->loss >= DRB_LOSS_BPR && h->loss <= DRB_LOSS_SL, "unknown loss id %d", h->loss);
    DRB_REQUIRE((first + nsteps - 1) * batch < n || nsteps == 0 || n == 0, "steps [%lld,%lld) exceed %lld triples", first,
                first + nsteps, n);
    p.P = P; p.Q = Q;
    carve(d_ws, U, I, F, h->opt, &p.ws, d_bias != nullptr);
    p.bu = bu; p.bi = bi; p.bj = bj;
    p.n = n; p.batch = batch; p.first_step = first; p.n_steps = nsteps;
    p.U = U; p.I = I; p.F = F; p.tile = kTileMax;
    p.lr = h->lr; p.reg1 = h->reg_1; p.reg2 = h->reg_2; p.opt = h->opt;
    p.beta1 = h->beta1; p.beta2 = h->beta2; p.eps = h->eps;
    p.adam_step0 = adam_step0;
    p.step_loss = d_step_loss;
    p.apply = apply;
    p.phases = 3;
    p.dense_hint = -1;
    p.Pn = nullptr;
    p.Qn = nullptr;
    p.gscale = 1.f;
    p.dense_grad = 0;
    p.neg_mult = 1.f;
    p.keep_counts = 0;
    p.neg_row_ptr = nullptr;
    p.neg_col = nullptr;
    p.neg_out = nullptr;
    p.neg_seed = 0ull;
    p.loss = h->loss;
    p.bias = d_bias;
    return DRB_OK;
}

extern "C" int drb_mf_bpr_train_steps(float *d_P, float *d_Q, void *d_ws, int32_t U, int32_t I, int32_t F,
                                      const int32_t *d_bu, const int32_t *d_bi, const int32_t *d_bj, int64_t n,
                                      int64_t batch, int64_t first_step, int64_t n_steps, const drb_hyper *hyper,
                                      int64_t adam_step0, double *d_step_loss, int32_t sync_and_check,
                                      int64_t *nan_step, void *stream)
{
    StepParams p;
    int rc = fill_params(p, d_P, d_Q, d_ws, U, I, F, d_bu, d_bi, d_bj, n, batch, first_step, n_steps, hyper, adam_step0,
                         d_step_loss, 1);
    if (rc != DRB_OK) return rc;
    if (n_steps == 0) return DRB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_steps(p, st);
    if (rc != DRB_OK) return rc;
    if (sync_and_check) return check_nan(d_ws, st, nan_step);
    return DRB_OK;
}

extern "C" int drb_mf_bpr_train_steps_fused_neg(float *d_P, float *d_Q, void *d_ws, int32_t U, int32_t I, int32_t F,
                                                const int32_t *d_bu, const int32_t *d_bi, const int64_t *d_row_ptr,
                                                const int32_t *d_col, uint64_t seed, int32_t *d_neg_out, int64_t n,
                                                int64_t batch, int64_t first_step, int64_t n_steps, const drb_hyper *hyper,
                                                int64_t adam_step0, double *d_step_loss, int32_t sync_and_check,
                                                int64_t *nan_step, void *stream)
{
    DRB_REQUIRE(d_row_ptr && d_col, "train_steps_fused_neg: the user->item CSR is required");
    DRB_REQUIRE(hyper && hyper->loss < DRB_LOSS_CL, "train_steps_fused_neg: pair-wise losses only");
    StepParams p;
    // the negative plane is unused in this mode (bi stands in so that the TMA staging code stays uniform)
    int rc = fill_params(p, d_P, d_Q, d_ws, U, I, F, d_bu, d_bi, d_bi, n, batch, first_step, n_steps, hyper, adam_step0,
                         d_step_loss, 1);
    if (rc != DRB_OK) return rc;
    if (n_steps == 0) return DRB_OK;
    p.neg_row_ptr = d_row_ptr;
    p.neg_col = d_col;
    p.neg_out = d_neg_out;
    p.neg_seed = seed;
    p.dense_hint = 1;   // phase 2 must not re-derive negatives: dense sweep only
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_steps(p, st);
    if (rc != DRB_OK) return rc;
    if (sync_and_check) return check_nan(d_ws, st, nan_step);
    return DRB_OK;
}

extern "C" int drb_mf_bpr_loss(const float *d_P, const float *d_Q, void *d_ws, int32_t U, int32_t I, int32_t F,
                               const int32_t *d_bu, const int32_t *d_bi, const int32_t *d_bj, int64_t batch,
                               const drb_hyper *hyper, double *d_loss, void *stream)
{
    StepParams p;
    int rc = fill_params(p, (float *)d_P, (float *)d_Q, d_ws, U, I, F, d_bu, d_bi, d_bj, batch, batch, 0, 1, hyper, 0,
                         d_loss, 0);
    if (rc != DRB_OK) return rc;
    return launch_steps(p, (cudaStream_t)stream);
}

// ---- FM (daisy/model/FMRecommender.py:61-97): the MF step with first-order terms; d_bias = [u_bias (U), i_bias (I), bias_]
extern "C" size_t drb_fm_workspace_bytes(int32_t U, int32_t I, int32_t F, int32_t opt)
{
    return carve(nullptr, U, I, F, opt, nullptr, 1);
}

extern "C" int drb_fm_workspace_init(void *d_ws, int32_t U, int32_t I, int32_t F, int32_t opt, void *stream)
{
    DRB_REQUIRE(d_ws != nullptr && U > 0 && I > 0 && F > 0, "fm_workspace_init: bad arguments");
    DRB_CUDA(cudaMemsetAsync(d_ws, 0, carve(nullptr, U, I, F, opt, nullptr, 1), (cudaStream_t)stream));
    return DRB_OK;
}

extern "C" int drb_fm_train_steps(float *d_P, float *d_Q, float *d_bias, void *d_ws, int32_t U, int32_t I, int32_t F,
                                  const int32_t *d_bu, const int32_t *d_bi, const int32_t *d_bj, int64_t n, int64_t batch,
                                  int64_t first_step, int64_t n_steps, const drb_hyper *hyper, int64_t adam_step0,
                                  int32_t apply, double *d_step_loss, int32_t sync_and_check, int64_t *nan_step, void *stream)
{
    DRB_REQUIRE(d_bias != nullptr, "fm_train_steps: the bias vector is required");
    StepParams p;
    int rc = fill_params(p, d_P, d_Q, d_ws, U, I, F, d_bu, d_bi, d_bj, n, batch, first_step, n_steps, hyper, adam_step0,
                         d_step_loss, apply ? 1 : 0, d_bias);
    if (rc != DRB_OK) return rc;
    if (n_steps == 0) return DRB_OK;
    DRB_REQUIRE(apply || n_steps == 1, "fm_train_steps: apply=0 evaluates the loss of ONE batch");
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_steps(p, st);
    if (rc != DRB_OK) return rc;
    if (sync_and_check) return check_nan(d_ws, st, nan_step);
    return DRB_OK;
}

extern "C" int drb_mf_bpr_train_step_host(float *d_P, float *d_Q, void *d_ws, int32_t U, int32_t I, int32_t F,
                                          const int32_t *h_bu, const int32_t *h_bi, const int32_t *h_bj, int64_t batch,
                                          const drb_hyper *hyper, int64_t adam_step0, int32_t *d_stage, double *h_loss,
                                          void *stream)
{
    DRB_REQUIRE(h_bu && h_bi && h_bj && d_stage && h_loss && batch > 0, "train_step_host: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    size_t stride = (size_t)((batch + 3) / 4 * 4);  // keep each array 16-byte aligned for the TMA path
    size_t bytes = sizeof(int32_t) * (size_t)batch;
    DRB_CUDA(cudaMemcpyAsync(d_stage, h_bu, bytes, cudaMemcpyHostToDevice, st));
    DRB_CUDA(cudaMemcpyAsync(d_stage + stride, h_bi, bytes, cudaMemcpyHostToDevice, st));
    DRB_CUDA(cudaMemcpyAsync(d_stage + 2 * stride, h_bj, bytes, cudaMemcpyHostToDevice, st));
    double *d_loss = (double *)(d_stage + 3 * stride);
    StepParams p;
    int rc = fill_params(p, d_P, d_Q, d_ws, U, I, F, d_stage, d_stage + stride, d_stage + 2 * stride, batch, batch, 0, 1,
                         hyper, adam_step0, d_loss, 1);
    if (rc != DRB_OK) return rc;
    rc = launch_steps(p, st);
    if (rc != DRB_OK) return rc;
    DRB_CUDA(cudaMemcpyAsync(h_loss, d_loss, sizeof(double), cudaMemcpyDeviceToHost, st));
    int64_t nan_step = -1;
    return check_nan(d_ws, st, &nan_step);
}

// Pipelined end-to-end steps from HOST index planes: the H2D copy of step s+1 (copy stream) overlaps the
// kernel of step s (compute stream); every step's loss is read back to the host asynchronously.
extern "C" int drb_mf_bpr_train_steps_host(float *d_P, float *d_Q, void *d_ws, int32_t U, int32_t I, int32_t F,
                                           const int32_t *h_bu, const int32_t *h_bi, const int32_t *h_bj, int64_t n,
                                           int64_t batch, int64_t n_steps, const drb_hyper *hyper, int64_t adam_step0,
                                           int32_t *d_stage, double *d_loss, double *h_loss, int64_t *nan_step,
                                           void *stream)
{
    DRB_REQUIRE(h_bu && h_bi && h_bj && d_stage && d_loss && h_loss && batch > 0 && n_steps >= 0 && n >= 0,
                "train_steps_host: bad arguments");
    DRB_REQUIRE(n_steps == 0 || (n_steps - 1) * batch < n, "train_steps_host: %lld steps exceed %lld triples",
                (long long)n_steps, (long long)n);
    if (n_steps == 0) return DRB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    static thread_local cudaStream_t copy_st = nullptr;
    if (!copy_st) DRB_CUDA(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
    cudaEvent_t ready[2], freed[2], start;
    for (int k = 0; k < 2; ++k) {
        DRB_CUDA(cudaEventCreateWithFlags(&ready[k], cudaEventDisableTiming));
        DRB_CUDA(cudaEventCreateWithFlags(&freed[k], cudaEventDisableTiming));
    }
    DRB_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    const size_t stride = (size_t)((batch + 3) / 4 * 4);
    DRB_CUDA(cudaMemsetAsync(d_ws, 0, sizeof(WsHeader), st));   // clear a stale NaN flag once; sticky afterwards
    DRB_CUDA(cudaEventRecord(start, st));
    DRB_CUDA(cudaStreamWaitEvent(copy_st, start, 0));            // staging slots may still be in use upstream
    int rc = DRB_OK;
    for (int64_t s = 0; s < n_steps && rc == DRB_OK; ++s) {
        const int slot = (int)(s & 1);
        int32_t *sb = d_stage + (size_t)slot * 3 * stride;
        const int64_t base = s * batch, nb = (n - base < batch) ? n - base : batch;
        const size_t bytes = sizeof(int32_t) * (size_t)nb;
        if (s >= 2) DRB_CUDA(cudaStreamWaitEvent(copy_st, freed[slot], 0));
        DRB_CUDA(cudaMemcpyAsync(sb, h_bu + base, bytes, cudaMemcpyHostToDevice, copy_st));
        DRB_CUDA(cudaMemcpyAsync(sb + stride, h_bi + base, bytes, cudaMemcpyHostToDevice, copy_st));
        DRB_CUDA(cudaMemcpyAsync(sb + 2 * stride, h_bj + base, bytes, cudaMemcpyHostToDevice, copy_st));
        DRB_CUDA(cudaEventRecord(ready[slot], copy_st));
        DRB_CUDA(cudaStreamWaitEvent(st, ready[slot], 0));
        StepParams p;
        rc = fill_params(p, d_P, d_Q, d_ws, U, I, F, sb, sb + stride, sb + 2 * stride, nb, nb, 0, 1, hyper, adam_step0 + s,
                         d_loss + s, 1);
        if (rc == DRB_OK) rc = launch_steps(p, st, /*keep_status=*/true);
        if (rc != DRB_OK) break;
        DRB_CUDA(cudaEventRecord(freed[slot], st));
        DRB_CUDA(cudaMemcpyAsync(h_loss + s, d_loss + s, sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    int rc2 = (rc == DRB_OK) ? check_nan(d_ws, st, nan_step) : rc;
    cudaStreamSynchronize(copy_st);
    for (int k = 0; k < 2; ++k) {
        cudaEventDestroy(ready[k]);
        cudaEventDestroy(freed[k]);
    }
    cudaEventDestroy(start);
    return rc2;
}

extern "C" int drb_gather_triples(const int32_t *d_triples, const int64_t *d_perm, int64_t n, int32_t *d_bu,
                                  int32_t *d_bi, int32_t *d_bj, void *stream)
{
    DRB_REQUIRE(d_triples && d_bu && d_bi && d_bj && n >= 0, "gather_triples: bad arguments");
    if (n == 0) return DRB_OK;
    long long blocks = (n + 255) / 256;
    long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    gather_triples_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(d_triples, d_perm, n, d_bu, d_bi, d_bj);
    DRB_CUDA(cudaGetLastError());
    return DRB_OK;
}

extern "C" int drb_mf_workspace_layout(int32_t U, int32_t I, int32_t F, int32_t opt, int64_t *out8)
{
    DRB_REQUIRE(out8 && U > 0 && I > 0 && F > 0, "workspace_layout: bad arguments");
    Workspace w;
    carve((void *)(uintptr_t)256, U, I, F, opt, &w);   // fake non-null base: pointers become offsets + 256
    auto off = [](const void *p) { return (int64_t)((uintptr_t)p - 256); };
    out8[0] = off(&w.hdr->acc[0][0]);  out8[1] = 8 * sizeof(double);
    out8[2] = off(w.gQ);               out8[3] = (int64_t)sizeof(float) * I * F;
    out8[4] = off(w.cntI);             out8[5] = (int64_t)sizeof(unsigned long long) * I;
    out8[6] = off(w.gP);               out8[7] = off(w.cntU);
    return DRB_OK;
}

extern "C" int drb_mf_bpr_phase(float *d_P, float *d_Q, void *d_ws, int32_t U, int32_t I, int32_t F, const int32_t *d_bu,
                                const int32_t *d_bi, const int32_t *d_bj, int64_t begin, int64_t count, int32_t phase,
                                const drb_hyper *hyper, int64_t adam_step0, double *d_loss, void *stream)
{
    DRB_REQUIRE(phase == 1 || phase == 2, "mf_bpr_phase: phase must be 1 or 2");
    DRB_REQUIRE(begin >= 0 && count >= 0, "mf_bpr_phase: bad range");
    DRB_REQUIRE(hyper && hyper->loss < DRB_LOSS_CL, "mf_bpr_phase: the sharded step covers the pair-wise losses only");
    StepParams p;
    // count may be 0 on a rank (its users have no triple in this global batch): phases still run (loss, sweep)
    int rc = fill_params(p, d_P, d_Q, d_ws, U, I, F, d_bu + begin, d_bi + begin, d_bj + begin, count, count > 0 ? count : 1,
                         0, 1, hyper, adam_step0, d_loss, 1);
    if (rc != DRB_OK) return rc;
    p.phases = phase;
    p.dense_hint = 1;
    return launch_steps(p, (cudaStream_t)stream);
}
