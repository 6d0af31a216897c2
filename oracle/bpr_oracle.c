/*
 * oracle/bpr_oracle.c -- CPU restatement of daisyRec's pair-wise BPR hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA path in
 * daisyrec_b200/csrc; nothing in the product may call, link or import it.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * load the shared object built from it (oracle/_build/libbpr_oracle.so).
 *
 * Parity pinning: the reference (AmazingDD/daisyRec v2.3.0) ships no tests and no
 * golden vectors.  This restatement is pinned instead against
 * outputs of the reference itself, generated in the build container by
 * oracle/gen_golden.py (imports /root/reference through oracle/ref_harness.py) and
 * committed under tests/golden/ (npz files); tests/test_oracle_golden.py replays them.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference root).  Third-party arithmetic restated here:
 *   - numpy (pinned ">=1.18.0", requirements.txt:2; 2.3.5 installed) legacy
 *     RandomState: MT19937 (init_genrand seeding, genrand_int32 tempering) and
 *     randint's masked-rejection bounded draw on 32-bit words; np.setdiff1d.
 *   - PyTorch (pinned ">=1.1.0", requirements.txt:1; 2.11.0 installed): nn.Embedding
 *     gather, autograd of sum/sigmoid/log/norm, optim.SGD / optim.Adam update rules,
 *     argsort(descending).
 *
 * Arithmetic conventions of the oracle: element-wise math in fp32 exactly as written
 * below; batch reductions (loss, norms, gradient sums over duplicate rows) accumulate in
 * fp64 and are rounded to fp32 once.  That is the "exact sum of fp32 terms" both the
 * reference (fp32 pairwise / index_add order) and the CUDA path (fp32 atomics) approximate.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * MT19937, numpy legacy flavour.
 * numpy/random/src/mt19937/mt19937.c: mt19937_seed (Knuth init_genrand, pos=624),
 * mt19937_gen (regenerate 624 words, tempering).  Call site in the reference:
 * np.random.seed via daisy/utils/config.py:34, consumed by daisy/utils/sampler.py:86 and
 * daisy/utils/utils.py:75,79.
 * State layout handed across ctypes: uint32[625] = key[624] + pos.
 * ---------------------------------------------------------------------------------- */
#define MT_N 624
#define MT_M 397

void orc_mt_seed(uint32_t *st, uint32_t seed)
{
    for (int pos = 0; pos < MT_N; pos++) {
        st[pos] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)pos + 1u;
    }
    st[MT_N] = MT_N;
}

static void mt_regen(uint32_t *mt)
{
    int kk;
    uint32_t y;
    for (kk = 0; kk < MT_N - MT_M; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < MT_N - 1; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

uint32_t orc_mt_next(uint32_t *st)
{
    if (st[MT_N] >= MT_N) {
        mt_regen(st);
        st[MT_N] = 0;
    }
    uint32_t y = st[st[MT_N]++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* RandomState.randint(0, n) for n-1 <= 0xffffffff: legacy rk_interval /
 * _bounded_uint masked rejection (numpy/random/src/legacy + _bounded_integers):
 * max = n-1; max==0 returns 0 WITHOUT consuming a word; else mask = smallest 2^k-1 >= max,
 * draw 32-bit words until (word & mask) <= max. */
static uint32_t bounded_draw(uint32_t *st, uint32_t n)
{
    uint32_t mx = n - 1u;
    if (mx == 0u)
        return 0u;
    uint32_t mask = mx;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    uint32_t v;
    do {
        v = orc_mt_next(st) & mask;
    } while (v > mx);
    return v;
}

/* ------------------------------------------------------------------------------------
 * daisy/utils/sampler.py:63,84-89  -- BasicNegtiveSampler.sampling(), uniform branch.
 *   js = zeros((user_num, num_ng), int32)
 *   for u: js[u] = np.random.choice(np.setdiff1d(np.arange(item_num), past_inter), size=num_ng)
 * Restated literally: materialise the sorted complement of the user's positives, then
 * index it with num_ng bounded draws (np.random.choice(a, size) == a[randint(0, len(a), size)]).
 * row_ptr/col: user->item CSR of config['train_ur']; duplicates / unsorted columns are
 * tolerated (set semantics, like setdiff1d).  Returns 0, or -(u+1) if user u has no
 * complement (numpy raises "a cannot be empty" there).
 * ---------------------------------------------------------------------------------- */
int orc_sample_negatives(uint32_t *mt_state, const int64_t *row_ptr, const int32_t *col, int32_t user_num,
                         int32_t item_num, int32_t num_ng, int32_t *js)
{
    uint8_t *seen = (uint8_t *)malloc((size_t)item_num);
    int32_t *comp = (int32_t *)malloc(sizeof(int32_t) * (size_t)item_num);
    int rc = 0;
    for (int32_t u = 0; u < user_num && rc == 0; u++) {
        memset(seen, 0, (size_t)item_num);
        for (int64_t e = row_ptr[u]; e < row_ptr[u + 1]; e++)
            if (col[e] >= 0 && col[e] < item_num)
                seen[col[e]] = 1;
        int32_t n = 0;
        for (int32_t it = 0; it < item_num; it++)
            if (!seen[it])
                comp[n++] = it;
        if (n == 0) {
            rc = -(u + 1);
            break;
        }
        for (int32_t g = 0; g < num_ng; g++)
            js[(int64_t)u * num_ng + g] = comp[bounded_draw(mt_state, (uint32_t)n)];
    }
    free(seen);
    free(comp);
    return rc;
}

/* daisy/utils/sampler.py:91,99-101 -- neg_set = js[user]; explode -> int32 [N*G, 3] rows
 * (u, i, j); row order = DataFrame row order, the G negatives of a row consecutive. */
void orc_explode_triples(const int32_t *coo_u, const int32_t *coo_i, int64_t nnz, const int32_t *js, int32_t num_ng,
                         int32_t *triples)
{
    for (int64_t r = 0; r < nnz; r++)
        for (int32_t g = 0; g < num_ng; g++) {
            int32_t *t = triples + 3 * (r * num_ng + g);
            t[0] = coo_u[r];
            t[1] = coo_i[r];
            t[2] = js[(int64_t)coo_u[r] * num_ng + g];
        }
}

/* ------------------------------------------------------------------------------------
 * daisy/utils/utils.py:53-85 -- build_candidates_set, one user.
 *   sample_num = cand_num - len(r) if len(r) <= cand_num else 0
 *   if sample_num == 0: samples = np.random.choice(list(r), cand_num)
 *   else: neg = setdiff1d(arange(item_num), list(r)+list(train_ur[u]));
 *         samples = concat(np.random.choice(neg, size=sample_num), list(r))
 * gt[] is list(r) in the caller's (Python set iteration) order; train[] the user's train
 * positives.  Writes cand_num ids.  Returns 0 or -1 (empty complement).
 * ---------------------------------------------------------------------------------- */
int orc_build_candidates_user(uint32_t *mt_state, const int32_t *gt, int32_t n_gt, const int32_t *train,
                              int32_t n_train, int32_t item_num, int32_t cand_num, int64_t *out)
{
    int32_t sample_num = (n_gt <= cand_num) ? cand_num - n_gt : 0;
    if (sample_num == 0) {
        for (int32_t c = 0; c < cand_num; c++)
            out[c] = gt[bounded_draw(mt_state, (uint32_t)n_gt)];
        return 0;
    }
    uint8_t *seen = (uint8_t *)calloc((size_t)item_num, 1);
    int32_t *comp = (int32_t *)malloc(sizeof(int32_t) * (size_t)item_num);
    for (int32_t e = 0; e < n_gt; e++)
        if (gt[e] >= 0 && gt[e] < item_num)
            seen[gt[e]] = 1;
    for (int32_t e = 0; e < n_train; e++)
        if (train[e] >= 0 && train[e] < item_num)
            seen[train[e]] = 1;
    int32_t n = 0;
    for (int32_t it = 0; it < item_num; it++)
        if (!seen[it])
            comp[n++] = it;
    int rc = 0;
    if (n == 0) {
        rc = -1;
    } else {
        for (int32_t c = 0; c < sample_num; c++)
            out[c] = comp[bounded_draw(mt_state, (uint32_t)n)];
        for (int32_t e = 0; e < n_gt; e++)
            out[sample_num + e] = gt[e];
    }
    free(seen);
    free(comp);
    return rc;
}

/* ------------------------------------------------------------------------------------
 * Canonical fp32 dot product of the framework (DESIGN.md "score order").
 * The reference scores with library GEMV/bmm kernels whose summation order is not part
 * of its contract (daisy/model/MFRecommender.py:66,115,131).  The CUDA path fixes ONE
 * order so that scores -- and therefore top-K index lists -- are reproducible bit for bit:
 *   vec    = 4 if F%4==0, 2 if F%2==0, else 1
 *   chunks = F / vec;  W = min(32, next_pow2(chunks))   (lanes cooperating on one row)
 *   lane l: acc = 0; for chunk c = l, l+W, ...: for e in 0..vec-1: acc = fmaf(a, b, acc)
 *   then a xor-butterfly over the W lanes: for off = W/2 .. 1: acc[l] += acc[l ^ off]
 * (every lane ends with the same value because fp addition is commutative).
 * ---------------------------------------------------------------------------------- */
static int dot_vec(int F) { return (F % 4 == 0) ? 4 : (F % 2 == 0) ? 2 : 1; }
static int dot_width(int F)
{
    int chunks = F / dot_vec(F), w = 1;
    while (w < chunks && w < 32)
        w <<= 1;
    return w;
}

float orc_dot(const float *a, const float *b, int32_t F)
{
    int vec = dot_vec(F), W = dot_width(F), chunks = F / vec;
    float acc[32];
    for (int l = 0; l < W; l++) {
        float s = 0.f;
        for (int c = l; c < chunks; c += W)
            for (int e = 0; e < vec; e++)
                s = fmaf(a[c * vec + e], b[c * vec + e], s);
        acc[l] = s;
    }
    for (int off = W >> 1; off >= 1; off >>= 1) {
        float nxt[32];
        for (int l = 0; l < W; l++)
            nxt[l] = acc[l] + acc[l ^ off];
        memcpy(acc, nxt, sizeof(float) * (size_t)W);
    }
    return acc[0];
}

/* daisy/model/MFRecommender.py:63-68 (forward) / :99-104 (predict): y = sum_f P[u,f]*Q[i,f]. */
void orc_mf_predict(const float *P, const float *Q, int32_t F, const int32_t *u, const int32_t *i, int64_t n,
                    float *out)
{
    for (int64_t t = 0; t < n; t++)
        out[t] = orc_dot(P + (int64_t)u[t] * F, Q + (int64_t)i[t] * F, F);
}

/* ------------------------------------------------------------------------------------
 * One BPR-MF training step == calc_loss + backward + optimizer.step on one batch.
 *   forward      daisy/model/MFRecommender.py:63-68,71-73,83-85
 *   BPR loss     daisy/utils/loss.py:11            -(1e-10 + sigmoid(pos-neg)).log().sum()
 *   regulariser  daisy/model/MFRecommender.py:88-89,94-95 (un-squared L1 / Frobenius norms of
 *                the gathered [B,F] matrices, duplicates counted)
 *   backward     autograd (daisy/model/AbstractRecommender.py:125): with s = sigmoid(x),
 *                c = -s(1-s)/(1e-10+s):
 *                  g_u += c (q_i - q_j) + reg_1 sgn(p_u) + reg_2 p_u/||P_u||_F
 *                  g_i += c p_u        + reg_1 sgn(q_i) + reg_2 q_i/||Q_i||_F
 *                  g_j += -c p_u       + reg_1 sgn(q_j) + reg_2 q_j/||Q_j||_F
 *                (norm == 0 -> zero subgradient; sgn(0) = 0)
 *   update       optim.SGD, no momentum / weight decay (AbstractRecommender.py:55-56,126):
 *                theta -= lr * g     -- or optim.Adam defaults (:53-54): dense, every row moves.
 * opt: 0 = SGD, 1 = Adam (state m,v are table-sized, step_count is the 1-based step).
 * apply: 0 = loss only (calc_loss), 1 = also update.
 * A NaN loss leaves the tables untouched and returns NaN (AbstractRecommender.py:122-123).
 * parts[8] (optional): bpr, l1_u, l1_i, l1_j, fro_u, fro_i, fro_j, total.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    float lr, reg_1, reg_2;
    int32_t opt;
    float beta1, beta2, eps;
} orc_hyper;

static void adam_dense(float *theta, float *m, float *v, const double *g, int64_t n, const orc_hyper *h,
                       int64_t step_count)
{
    /* torch.optim.Adam (single tensor path): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
     * step_size = lr / (1-b1^t); denom = sqrt(v)/sqrt(1-b2^t) + eps; theta -= step_size * m/denom */
    double bc1 = 1.0 - pow((double)h->beta1, (double)step_count);
    double bc2 = 1.0 - pow((double)h->beta2, (double)step_count);
    float step_size = (float)((double)h->lr / bc1);
    float bc2_sqrt = (float)sqrt(bc2);
    for (int64_t k = 0; k < n; k++) {
        float gk = (float)g[k];
        m[k] = m[k] + (gk - m[k]) * (1.f - h->beta1);              /* lerp form used by torch */
        v[k] = v[k] * h->beta2 + (1.f - h->beta2) * gk * gk;
        float denom = sqrtf(v[k]) / bc2_sqrt + h->eps;
        theta[k] = theta[k] - step_size * (m[k] / denom);
    }
}

double orc_mf_bpr_step(float *P, float *Q, int32_t U, int32_t I, int32_t F, const int32_t *bu, const int32_t *bi,
                       const int32_t *bj, int64_t B, const orc_hyper *h, int32_t apply, float *mP, float *vP,
                       float *mQ, float *vQ, int64_t step_count, double *parts)
{
    const float gamma = 1e-10f;
    float *coef = (float *)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
    double bpr = 0.0, l1u = 0, l1i = 0, l1j = 0, s2u = 0, s2i = 0, s2j = 0;
    for (int64_t t = 0; t < B; t++) {
        const float *p = P + (int64_t)bu[t] * F, *qi = Q + (int64_t)bi[t] * F, *qj = Q + (int64_t)bj[t] * F;
        float pos = orc_dot(p, qi, F), neg = orc_dot(p, qj, F);
        float x = pos - neg;
        float s = 1.f / (1.f + expf(-x));
        bpr += (double)(-logf(gamma + s));
        coef[t] = -(s * (1.f - s)) / (gamma + s);
        for (int f = 0; f < F; f++) {
            l1u += fabsf(p[f]);
            s2u += (double)(p[f] * p[f]);
            l1i += fabsf(qi[f]);
            s2i += (double)(qi[f] * qi[f]);
            l1j += fabsf(qj[f]);
            s2j += (double)(qj[f] * qj[f]);
        }
    }
    double nu = sqrt(s2u), ni = sqrt(s2i), nj = sqrt(s2j);
    /* fp32 loss assembly, as the reference adds fp32 scalars (MFRecommender.py:88-95) */
    float loss = (float)bpr;
    loss += h->reg_1 * ((float)l1i + (float)l1j);
    loss += h->reg_2 * ((float)ni + (float)nj);
    loss += h->reg_1 * (float)l1u;
    loss += h->reg_2 * (float)nu;
    if (parts) {
        parts[0] = bpr; parts[1] = l1u; parts[2] = l1i; parts[3] = l1j;
        parts[4] = nu;  parts[5] = ni;  parts[6] = nj;  parts[7] = (double)loss;
    }
    if (!apply || isnan(loss)) {
        free(coef);
        return (double)loss;
    }
    double *gP = (double *)calloc((size_t)U * F, sizeof(double));
    double *gQ = (double *)calloc((size_t)I * F, sizeof(double));
    float inu = nu > 0 ? (float)(1.0 / nu) : 0.f, ini = ni > 0 ? (float)(1.0 / ni) : 0.f,
          inj = nj > 0 ? (float)(1.0 / nj) : 0.f;
    for (int64_t t = 0; t < B; t++) {
        const float *p = P + (int64_t)bu[t] * F, *qi = Q + (int64_t)bi[t] * F, *qj = Q + (int64_t)bj[t] * F;
        double *gu = gP + (int64_t)bu[t] * F, *gi = gQ + (int64_t)bi[t] * F, *gj = gQ + (int64_t)bj[t] * F;
        float c = coef[t];
        for (int f = 0; f < F; f++) {
            float sp = (p[f] > 0) - (p[f] < 0), si = (qi[f] > 0) - (qi[f] < 0), sj = (qj[f] > 0) - (qj[f] < 0);
            gu[f] += (double)(c * (qi[f] - qj[f])) + (double)(h->reg_1 * sp) + (double)(h->reg_2 * p[f] * inu);
            gi[f] += (double)(c * p[f]) + (double)(h->reg_1 * si) + (double)(h->reg_2 * qi[f] * ini);
            gj[f] += (double)(-c * p[f]) + (double)(h->reg_1 * sj) + (double)(h->reg_2 * qj[f] * inj);
        }
    }
    if (h->opt == 0) {
        for (int64_t k = 0; k < (int64_t)U * F; k++)
            P[k] = P[k] - h->lr * (float)gP[k];
        for (int64_t k = 0; k < (int64_t)I * F; k++)
            Q[k] = Q[k] - h->lr * (float)gQ[k];
    } else {
        adam_dense(P, mP, vP, gP, (int64_t)U * F, h, step_count);
        adam_dense(Q, mQ, vQ, gQ, (int64_t)I * F, h, step_count);
    }
    free(gP);
    free(gQ);
    free(coef);
    return (double)loss;
}

/* daisy/model/AbstractRecommender.py:112-128 -- the step loop of one epoch over a given
 * permutation of triple indices (DataLoader(shuffle=True) order, drop_last=False).
 * triples: int32 [T,3] (sampler output).  perm: int64 [T] or NULL (identity).
 * step_loss[ceil(T/B)] receives each step's loss; returns the epoch sum (current_loss, :128)
 * or NaN at the first NaN step (remaining steps are not run, like the ValueError). */
double orc_mf_bpr_epoch(float *P, float *Q, int32_t U, int32_t I, int32_t F, const int32_t *triples, int64_t T,
                        const int64_t *perm, int64_t batch, const orc_hyper *h, float *mP, float *vP, float *mQ,
                        float *vQ, int64_t first_step_count, double *step_loss)
{
    int32_t *bu = (int32_t *)malloc(sizeof(int32_t) * (size_t)batch * 3);
    int32_t *bi = bu + batch, *bj = bi + batch;
    double total = 0.0;
    int64_t s = 0;
    for (int64_t base = 0; base < T; base += batch, s++) {
        int64_t nb = (T - base < batch) ? T - base : batch;
        for (int64_t k = 0; k < nb; k++) {
            int64_t p = perm ? perm[base + k] : base + k;
            bu[k] = triples[3 * p];
            bi[k] = triples[3 * p + 1];
            bj[k] = triples[3 * p + 2];
        }
        double l = orc_mf_bpr_step(P, Q, U, I, F, bu, bi, bj, nb, h, 1, mP, vP, mQ, vQ, first_step_count + s, NULL);
        if (step_loss)
            step_loss[s] = l;
        if (isnan(l)) {
            total = l;
            break;
        }
        total += l;
    }
    free(bu);
    return total;
}

/* ------------------------------------------------------------------------------------
 * daisy/model/MFRecommender.py:106-123 -- MF.rank: per user, score its cand_num candidates,
 * argsort descending, gather ids, keep topk; output float32 (the reference torch.cat's onto
 * an empty float tensor, :107,:121).  Ties (equal fp32 scores): lower candidate position
 * first (torch leaves it unspecified; duplicates of one id make it immaterial).
 * ---------------------------------------------------------------------------------- */
typedef struct {
    float s;
    int32_t pos;
} orc_sc;

static int sc_desc(const void *a, const void *b)
{
    const orc_sc *x = (const orc_sc *)a, *y = (const orc_sc *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->pos > y->pos) - (x->pos < y->pos);
}

void orc_mf_rank(const float *P, const float *Q, int32_t F, const int64_t *users, int64_t n_users,
                 const int64_t *cands, int32_t cand_num, int32_t topk, float *out)
{
    orc_sc *sc = (orc_sc *)malloc(sizeof(orc_sc) * (size_t)cand_num);
    for (int64_t r = 0; r < n_users; r++) {
        const float *p = P + users[r] * F;
        const int64_t *c = cands + r * cand_num;
        for (int32_t k = 0; k < cand_num; k++) {
            sc[k].s = orc_dot(p, Q + c[k] * F, F);
            sc[k].pos = k;
        }
        qsort(sc, (size_t)cand_num, sizeof(orc_sc), sc_desc);
        for (int32_t k = 0; k < topk && k < cand_num; k++)
            out[r * topk + k] = (float)c[sc[k].pos];
    }
    free(sc);
}

/* daisy/model/MFRecommender.py:126-133 -- MF.full_rank(u): P[u] @ Q.T, argsort descending,
 * first topk item ids (int64); no masking of train items.  Ties: lower item id first. */
void orc_mf_full_rank(const float *P, const float *Q, int32_t F, int32_t item_num, const int64_t *users,
                      int64_t n_users, int32_t topk, int64_t *out)
{
    orc_sc *sc = (orc_sc *)malloc(sizeof(orc_sc) * (size_t)item_num);
    for (int64_t r = 0; r < n_users; r++) {
        const float *p = P + users[r] * F;
        for (int32_t k = 0; k < item_num; k++) {
            sc[k].s = orc_dot(p, Q + (int64_t)k * F, F);
            sc[k].pos = k;
        }
        qsort(sc, (size_t)item_num, sizeof(orc_sc), sc_desc);
        for (int32_t k = 0; k < topk && k < item_num; k++)
            out[r * topk + k] = sc[k].pos;
    }
    free(sc);
}

/* ====================================================================================
 * LightGCN (daisy/model/LightGCNRecommender.py)
 * ------------------------------------------------------------------------------------
 * orc_lgcn_propagate: forward() :117-129 -- E_0 = cat(P, Q); E_l = A_hat E_{l-1} (torch.sparse.mm, :122);
 *   output = mean over the L+1 layers (:125-126).  A_hat is given as CSR over the U+I nodes
 *   (row_ptr int64, col int32 ascending, val fp32) -- the symmetric normalised adjacency of
 *   get_norm_adj_mat :73-107 (built by oracle.lgcn_norm_adj in Python, pinned against the reference).
 *   fp32 products, fp64 row accumulation rounded once (the "exact sum of fp32 terms" convention).
 * ================================================================================== */
static void spmm_csr(const int64_t *row_ptr, const int32_t *col, const float *val, int64_t n, int32_t F, const float *X,
                     float *Y)
{
    double *acc = (double *)malloc(sizeof(double) * (size_t)F);
    for (int64_t r = 0; r < n; r++) {
        for (int f = 0; f < F; f++) acc[f] = 0.0;
        for (int64_t e = row_ptr[r]; e < row_ptr[r + 1]; e++) {
            const float *x = X + (int64_t)col[e] * F;
            float v = val[e];
            for (int f = 0; f < F; f++) acc[f] += (double)(v * x[f]);
        }
        for (int f = 0; f < F; f++) Y[r * F + f] = (float)acc[f];
    }
    free(acc);
}

/* out[n*F] = mean_{l=0..L} A^l X0 */
void orc_lgcn_propagate(const int64_t *row_ptr, const int32_t *col, const float *val, int64_t n, int32_t F, int32_t L,
                        const float *X0, float *out)
{
    size_t sz = (size_t)n * F;
    float *a = (float *)malloc(sizeof(float) * sz), *b = (float *)malloc(sizeof(float) * sz);
    double *sum = (double *)malloc(sizeof(double) * sz);
    memcpy(a, X0, sizeof(float) * sz);
    for (size_t k = 0; k < sz; k++) sum[k] = (double)X0[k];
    for (int l = 0; l < L; l++) {
        spmm_csr(row_ptr, col, val, n, F, a, b);
        for (size_t k = 0; k < sz; k++) sum[k] += (double)b[k];
        float *t = a; a = b; b = t;
    }
    for (size_t k = 0; k < sz; k++) out[k] = (float)(sum[k] / (double)(L + 1));
    free(a); free(b); free(sum);
}

/* One LightGCN BPR step == calc_loss :131-169 + backward + optimizer.step.
 *   scores on the PROPAGATED rows (:141-143,:157-158), BPR loss (daisy/utils/loss.py:11),
 *   regulariser on the EGO (layer-0) rows (:145-146,:159,:163-164), un-squared L1 / Frobenius norms.
 *   backward: dL/dE_mean is non-zero on the batch rows only; since E_mean = 1/(L+1) sum_l A^l E_0 and A is
 *   symmetric, dL/dE_0 = 1/(L+1) sum_l A^l (dL/dE_mean)  -> the same propagation applied to the gradient;
 *   plus the ego-row regulariser gradient.  E0 = cat(P,Q) [(U+I),F] updated in place; opt as in orc_mf_bpr_step.
 */
double orc_lgcn_bpr_step(float *E0, int32_t U, int32_t I, int32_t F, int32_t L, const int64_t *row_ptr, const int32_t *col,
                         const float *val, const int32_t *bu, const int32_t *bi, const int32_t *bj, int64_t B,
                         const orc_hyper *h, int32_t apply, float *m, float *v, int64_t step_count)
{
    const float gamma = 1e-10f;
    int64_t n = (int64_t)U + I;
    size_t sz = (size_t)n * F;
    float *Em = (float *)malloc(sizeof(float) * sz);
    orc_lgcn_propagate(row_ptr, col, val, n, F, L, E0, Em);
    const float *Pm = Em, *Qm = Em + (size_t)U * F, *P = E0, *Q = E0 + (size_t)U * F;
    float *coef = (float *)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
    double bpr = 0, l1u = 0, l1i = 0, l1j = 0, s2u = 0, s2i = 0, s2j = 0;
    for (int64_t t = 0; t < B; t++) {
        const float *p = Pm + (int64_t)bu[t] * F, *qi = Qm + (int64_t)bi[t] * F, *qj = Qm + (int64_t)bj[t] * F;
        float x = orc_dot(p, qi, F) - orc_dot(p, qj, F);
        float s = 1.f / (1.f + expf(-x));
        bpr += (double)(-logf(gamma + s));
        coef[t] = -(s * (1.f - s)) / (gamma + s);
        const float *pe = P + (int64_t)bu[t] * F, *qie = Q + (int64_t)bi[t] * F, *qje = Q + (int64_t)bj[t] * F;
        for (int f = 0; f < F; f++) {
            l1u += fabsf(pe[f]); s2u += (double)(pe[f] * pe[f]);
            l1i += fabsf(qie[f]); s2i += (double)(qie[f] * qie[f]);
            l1j += fabsf(qje[f]); s2j += (double)(qje[f] * qje[f]);
        }
    }
    double nu = sqrt(s2u), ni = sqrt(s2i), nj = sqrt(s2j);
    float loss = (float)bpr;
    loss += h->reg_1 * ((float)l1u + (float)l1i + (float)l1j);      /* :163 */
    loss += h->reg_2 * ((float)nu + (float)ni + (float)nj);         /* :164 */
    if (!apply || isnan(loss)) {
        free(Em); free(coef);
        return (double)loss;
    }
    /* dL/dE_mean (dense buffer, non-zero on batch rows) */
    double *Gd = (double *)calloc(sz, sizeof(double));
    for (int64_t t = 0; t < B; t++) {
        const float *p = Pm + (int64_t)bu[t] * F, *qi = Qm + (int64_t)bi[t] * F, *qj = Qm + (int64_t)bj[t] * F;
        double *gu = Gd + (int64_t)bu[t] * F, *gi = Gd + ((int64_t)U + bi[t]) * F, *gj = Gd + ((int64_t)U + bj[t]) * F;
        float c = coef[t];
        for (int f = 0; f < F; f++) {
            gu[f] += (double)(c * (qi[f] - qj[f]));
            gi[f] += (double)(c * p[f]);
            gj[f] += (double)(-c * p[f]);
        }
    }
    float *G = (float *)malloc(sizeof(float) * sz), *Gp = (float *)malloc(sizeof(float) * sz);
    for (size_t k = 0; k < sz; k++) G[k] = (float)Gd[k];
    orc_lgcn_propagate(row_ptr, col, val, n, F, L, G, Gp);          /* 1/(L+1) sum_l A^l G */
    /* ego regulariser gradient */
    float inu = nu > 0 ? (float)(1.0 / nu) : 0.f, ini = ni > 0 ? (float)(1.0 / ni) : 0.f, inj = nj > 0 ? (float)(1.0 / nj) : 0.f;
    for (size_t k = 0; k < sz; k++) Gd[k] = (double)Gp[k];
    for (int64_t t = 0; t < B; t++) {
        const float *pe = P + (int64_t)bu[t] * F, *qie = Q + (int64_t)bi[t] * F, *qje = Q + (int64_t)bj[t] * F;
        double *gu = Gd + (int64_t)bu[t] * F, *gi = Gd + ((int64_t)U + bi[t]) * F, *gj = Gd + ((int64_t)U + bj[t]) * F;
        for (int f = 0; f < F; f++) {
            float sp = (pe[f] > 0) - (pe[f] < 0), si = (qie[f] > 0) - (qie[f] < 0), sj = (qje[f] > 0) - (qje[f] < 0);
            gu[f] += (double)(h->reg_1 * sp) + (double)(h->reg_2 * pe[f] * inu);
            gi[f] += (double)(h->reg_1 * si) + (double)(h->reg_2 * qie[f] * ini);
            gj[f] += (double)(h->reg_1 * sj) + (double)(h->reg_2 * qje[f] * inj);
        }
    }
    if (h->opt == 0) {
        for (size_t k = 0; k < sz; k++) E0[k] = E0[k] - h->lr * (float)Gd[k];
    } else {
        adam_dense(E0, m, v, Gd, (int64_t)sz, h, step_count);
    }
    free(Em); free(coef); free(Gd); free(G); free(Gp);
    return (double)loss;
}
