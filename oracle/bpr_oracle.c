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
 * opt: 0 = SGD, 1 = Adam (state m,v are table-sized, step_count is the 1-based step), 2 = Adagrad, 3 = RMSprop
 *      (MF only; one table-sized state each, kept in m).
 * apply: 0 = loss only (calc_loss), 1 = also update.
 * A NaN loss leaves the tables untouched and returns NaN (AbstractRecommender.py:122-123).
 * parts[8] (optional): bpr, l1_u, l1_i, l1_j, fro_u, fro_i, fro_j, total.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    float lr, reg_1, reg_2;
    int32_t opt;
    float beta1, beta2, eps;
    int32_t loss; /* 0 BPR (loss.py:5-13), 1 HL = HingeLoss (loss.py:16-23), 2 TL = TOP1Loss (loss.py:26-33);
                   * point-wise branch of MF.calc_loss (MFRecommender.py:75-81), batch[2] = label:
                   * 3 CL = nn.BCEWithLogitsLoss(reduction='sum'), 4 SL = nn.MSELoss(reduction='sum')
                   * (AbstractRecommender.py:79-82) */
} orc_hyper;

/* pairwise loss term and its derivatives w.r.t. the positive / negative score */
static float pair_loss(int kind, float pos, float neg, float *cp, float *cn)
{
    const float gamma = 1e-10f;
    if (kind == 3) {                       /* point-wise, neg carries the label y.  ATen binary_cross_entropy_with_logits:
                                            * (1 - y) x - log_sigmoid(x), log_sigmoid(x) = min(x,0) - log1p(exp(-|x|));
                                            * log_sigmoid_backward: x < 0 ? 1 - z/(1+z) : z/(1+z), z = exp(-|x|) */
        float x = pos, y = neg, z = expf(-fabsf(x));
        float logsig = fminf(x, 0.f) - log1pf(z);
        float dls = x < 0.f ? 1.f - z / (1.f + z) : z / (1.f + z);
        *cp = (1.f - y) - dls;
        *cn = 0.f;
        return (1.f - y) * x - logsig;
    }
    if (kind == 4) {                       /* point-wise MSE: (x - y)^2 */
        float d = pos - neg;
        *cp = 2.f * d;
        *cn = 0.f;
        return d * d;
    }
    if (kind == 1) {                       /* clamp(1 - (pos - neg), min=0); clamp's backward passes at equality */
        float m = 1.f - (pos - neg);
        *cp = (m >= 0.f) ? -1.f : 0.f;
        *cn = -*cp;
        return m > 0.f ? m : 0.f;
    }
    if (kind == 2) {                       /* sigmoid(neg - pos) + sigmoid(neg^2) */
        float s1 = 1.f / (1.f + expf(-(neg - pos))), s2 = 1.f / (1.f + expf(-(neg * neg)));
        *cp = -(s1 * (1.f - s1));
        *cn = s1 * (1.f - s1) + s2 * (1.f - s2) * 2.f * neg;
        return s1 + s2;
    }
    float x = pos - neg;
    float s = 1.f / (1.f + expf(-x));
    *cp = -(s * (1.f - s)) / (gamma + s);
    *cn = -*cp;
    return -logf(gamma + s);
}

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

/* torch.optim.Adagrad defaults (lr_decay 0, eps 1e-10, initial accumulator 0; AbstractRecommender.py:57-58):
 *   sum += g*g; theta -= lr * g / (sqrt(sum) + eps)            -- state in m
 * torch.optim.RMSprop defaults (alpha 0.99, eps 1e-8, no momentum, not centered; :59-60):
 *   sq = alpha*sq + (1-alpha) g*g; theta -= lr * g / (sqrt(sq) + eps)   -- state in m; dense (sq of untouched rows decays) */
static void adagrad_dense(float *theta, float *sum, const double *g, int64_t n, const orc_hyper *h)
{
    for (int64_t k = 0; k < n; k++) {
        float gk = (float)g[k];
        sum[k] = sum[k] + gk * gk;
        float std = sqrtf(sum[k]) + 1e-10f;
        theta[k] = theta[k] - h->lr * (gk / std);
    }
}

static void rmsprop_dense(float *theta, float *sq, const double *g, int64_t n, const orc_hyper *h)
{
    const float alpha = 0.99f;
    for (int64_t k = 0; k < n; k++) {
        float gk = (float)g[k];
        sq[k] = sq[k] * alpha + (1.f - alpha) * gk * gk;
        float avg = sqrtf(sq[k]) + 1e-8f;
        theta[k] = theta[k] - h->lr * (gk / avg);
    }
}

/* FM's first-order terms (daisy/model/FMRecommender.py:46-49,66-67): u_bias [U], i_bias [I], bias_ [1] and their
 * optimiser state (m / v like the tables).  NULL = plain MF. */
typedef struct {
    float *ub, *ib, *b0;
    float *m_ub, *v_ub, *m_ib, *v_ib, *m_b0, *v_b0;
} orc_fm_bias;

static void dense_update(float *theta, float *m, float *v, const double *g, int64_t n, const orc_hyper *h, int64_t step_count);

static double mf_step_impl(float *P, float *Q, int32_t U, int32_t I, int32_t F, const int32_t *bu, const int32_t *bi,
                           const int32_t *bj, int64_t B, const orc_hyper *h, int32_t apply, float *mP, float *vP,
                           float *mQ, float *vQ, int64_t step_count, double *parts, const orc_fm_bias *fm)
{
    float *coef = (float *)malloc(sizeof(float) * 2 * (size_t)(B > 0 ? B : 1));
    double bpr = 0.0, l1u = 0, l1i = 0, l1j = 0, s2u = 0, s2i = 0, s2j = 0;
    const int pw = h->loss >= 3;           /* point-wise: bj holds labels, only P_u and Q_i are regularised (:79-80,:94-95) */
    for (int64_t t = 0; t < B; t++) {
        const float *p = P + (int64_t)bu[t] * F, *qi = Q + (int64_t)bi[t] * F, *qj = pw ? NULL : Q + (int64_t)bj[t] * F;
        float pos = orc_dot(p, qi, F), neg = pw ? (float)bj[t] : orc_dot(p, qj, F);
        if (fm) {                          /* pred += u_bias(user) + i_bias(item) + bias_  (FMRecommender.py:66-67) */
            pos += (fm->ub[bu[t]] + fm->ib[bi[t]]) + fm->b0[0];
            if (!pw)
                neg += (fm->ub[bu[t]] + fm->ib[bj[t]]) + fm->b0[0];
        }
        bpr += (double)pair_loss(h->loss, pos, neg, &coef[2 * t], &coef[2 * t + 1]);
        for (int f = 0; f < F; f++) {
            l1u += fabsf(p[f]);
            s2u += (double)(p[f] * p[f]);
            l1i += fabsf(qi[f]);
            s2i += (double)(qi[f] * qi[f]);
            if (!pw) {
                l1j += fabsf(qj[f]);
                s2j += (double)(qj[f] * qj[f]);
            }
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
    double *gub = fm ? (double *)calloc((size_t)U, sizeof(double)) : NULL;
    double *gib = fm ? (double *)calloc((size_t)I, sizeof(double)) : NULL;
    double gb0 = 0.0;
    float inu = nu > 0 ? (float)(1.0 / nu) : 0.f, ini = ni > 0 ? (float)(1.0 / ni) : 0.f,
          inj = nj > 0 ? (float)(1.0 / nj) : 0.f;
    for (int64_t t = 0; t < B; t++) {
        const float *p = P + (int64_t)bu[t] * F, *qi = Q + (int64_t)bi[t] * F;
        double *gu = gP + (int64_t)bu[t] * F, *gi = gQ + (int64_t)bi[t] * F;
        float cp = coef[2 * t], cn = coef[2 * t + 1];
        if (fm) {                          /* biases are not regularised (FMRecommender.py:76-95) */
            gub[bu[t]] += (double)cp + (double)cn;
            gib[bi[t]] += (double)cp;
            if (!pw)
                gib[bj[t]] += (double)cn;
            gb0 += (double)cp + (double)cn;
        }
        if (pw) {
            for (int f = 0; f < F; f++) {
                float sp = (p[f] > 0) - (p[f] < 0), si = (qi[f] > 0) - (qi[f] < 0);
                gu[f] += (double)(cp * qi[f]) + (double)(h->reg_1 * sp) + (double)(h->reg_2 * p[f] * inu);
                gi[f] += (double)(cp * p[f]) + (double)(h->reg_1 * si) + (double)(h->reg_2 * qi[f] * ini);
            }
            continue;
        }
        const float *qj = Q + (int64_t)bj[t] * F;
        double *gj = gQ + (int64_t)bj[t] * F;
        for (int f = 0; f < F; f++) {
            float sp = (p[f] > 0) - (p[f] < 0), si = (qi[f] > 0) - (qi[f] < 0), sj = (qj[f] > 0) - (qj[f] < 0);
            double du = (h->loss == 0) ? (double)(cp * (qi[f] - qj[f])) : (double)(cp * qi[f]) + (double)(cn * qj[f]);
            gu[f] += du + (double)(h->reg_1 * sp) + (double)(h->reg_2 * p[f] * inu);
            gi[f] += (double)(cp * p[f]) + (double)(h->reg_1 * si) + (double)(h->reg_2 * qi[f] * ini);
            gj[f] += (double)(cn * p[f]) + (double)(h->reg_1 * sj) + (double)(h->reg_2 * qj[f] * inj);
        }
    }
    dense_update(P, mP, vP, gP, (int64_t)U * F, h, step_count);
    dense_update(Q, mQ, vQ, gQ, (int64_t)I * F, h, step_count);
    if (fm) {
        dense_update(fm->ub, fm->m_ub, fm->v_ub, gub, U, h, step_count);
        dense_update(fm->ib, fm->m_ib, fm->v_ib, gib, I, h, step_count);
        dense_update(fm->b0, fm->m_b0, fm->v_b0, &gb0, 1, h, step_count);
        free(gub);
        free(gib);
    }
    free(gP);
    free(gQ);
    free(coef);
    return (double)loss;
}

static void dense_update(float *theta, float *m, float *v, const double *g, int64_t n, const orc_hyper *h, int64_t step_count)
{
    if (h->opt == 0) {
        for (int64_t k = 0; k < n; k++)
            theta[k] = theta[k] - h->lr * (float)g[k];
    } else if (h->opt == 2) {
        adagrad_dense(theta, m, g, n, h);
    } else if (h->opt == 3) {
        rmsprop_dense(theta, m, g, n, h);
    } else {
        adam_dense(theta, m, v, g, n, h, step_count);
    }
}

/* the optimiser switch (SGD / Adam / Adagrad / RMSprop, torch defaults) for the other oracle files */
void orc_dense_update(float *theta, float *m, float *v, const double *g, int64_t n, const orc_hyper *h, int64_t step_count)
{
    dense_update(theta, m, v, g, n, h, step_count);
}

double orc_mf_bpr_step(float *P, float *Q, int32_t U, int32_t I, int32_t F, const int32_t *bu, const int32_t *bi,
                       const int32_t *bj, int64_t B, const orc_hyper *h, int32_t apply, float *mP, float *vP,
                       float *mQ, float *vQ, int64_t step_count, double *parts)
{
    return mf_step_impl(P, Q, U, I, F, bu, bi, bj, B, h, apply, mP, vP, mQ, vQ, step_count, parts, NULL);
}

/* ------------------------------------------------------------------------------------
 * daisy/model/FMRecommender.py:61-97 -- FM.forward / calc_loss + backward + optimizer.step: the MF step above with
 * pred = <p_u, q_i> + ((u_bias[u] + i_bias[i]) + bias_) in fp32 (:66-67: the three first-order terms are summed first,
 * then added to the factor product); the regulariser is MF's (factor rows only).  bias [U + I + 1] = u_bias, i_bias,
 * bias_ packed; bias_state [2 * (U + I + 1)] = their m then v (Adam) / state (Adagrad, RMSprop), or NULL for SGD.
 * ---------------------------------------------------------------------------------- */
double orc_fm_step(float *P, float *Q, float *bias, int32_t U, int32_t I, int32_t F, const int32_t *bu, const int32_t *bi,
                   const int32_t *bj, int64_t B, const orc_hyper *h, int32_t apply, float *mP, float *vP, float *mQ,
                   float *vQ, float *bias_state, int64_t step_count, double *parts)
{
    orc_fm_bias fm;
    int64_t nb = (int64_t)U + I + 1;
    fm.ub = bias;
    fm.ib = bias + U;
    fm.b0 = bias + U + I;
    fm.m_ub = bias_state ? bias_state : NULL;
    fm.m_ib = bias_state ? bias_state + U : NULL;
    fm.m_b0 = bias_state ? bias_state + U + I : NULL;
    fm.v_ub = bias_state ? bias_state + nb : NULL;
    fm.v_ib = bias_state ? bias_state + nb + U : NULL;
    fm.v_b0 = bias_state ? bias_state + nb + U + I : NULL;
    return mf_step_impl(P, Q, U, I, F, bu, bi, bj, B, h, apply, mP, vP, mQ, vQ, step_count, parts, &fm);
}

/* FM.rank / full_rank / predict scores (FMRecommender.py:99-131): <p_u, q_c> + ((u_bias[u] + i_bias[c]) + bias_), then
 * argsort(descending) -- stable by position like orc_mf_rank.  items NULL: all item ids (full_rank). */
void orc_fm_scores(const float *P, const float *Q, const float *bias, int32_t U, int32_t I, int32_t F, int64_t user,
                   const int64_t *items, int64_t n, float *scores)
{
    for (int64_t k = 0; k < n; k++) {
        int64_t c = items ? items[k] : k;
        scores[k] = orc_dot(P + user * F, Q + c * F, F) + ((bias[user] + bias[U + c]) + bias[U + I]);
    }
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

/* ====================================================================================
 * NeuMF (daisy/model/NeuMFRecommender.py), model_name == 'NeuMF', dropout == 0.
 * ------------------------------------------------------------------------------------
 * forward :118-137   GMF = UG[u] * IG[i];  x0 = cat(UM[u], IM[i]);  for each layer: x = relu(W x + b) (:58-64, the
 *                    Dropout in front of every Linear is the identity at p = 0);  pred = wp . cat(GMF, x) + bp (:66-71).
 * calc_loss :139-169 BPR (loss.py:11) + the regulariser EXACTLY as written, including the reference's quirk that
 *                    lines :158 and :160 use embed_item_GMF(neg_item) where the MLP table was meant:
 *                      reg_1*(|IG_i|_1+|IG_j|_1) + reg_1*(|IM_i|_1+|IG_j|_1) + reg_2*(|IG_i|+|IG_j|) + reg_2*(|IM_i|+|IG_j|)
 *                      + reg_1*|UG_u|_1 + reg_1*|UM_u|_1 + reg_2*|UG_u| + reg_2*|UM_u|
 * backward + optimizer.step: closed form below; dense Adam (torch defaults) or SGD on every parameter.
 *
 * Parameter block layout (one flat fp32 buffer `W`, same order as module registration):
 *   layer l = 0..L-1: weight [out_l, in_l] row-major, then bias [out_l], with in_l = 2D/2^l, out_l = in_l/2;
 *   then predict weight [2F], predict bias [1].         D = F * 2^(L-1).
 * ================================================================================== */
/* mode = config['model_name'] (NeuMFRecommender.py:48-50,63-68): 0 'NeuMF' / 'NeuMF-pre' predict over cat(GMF, tower) [2F],
 * 1 'GMF' predict over the GMF product [F] (the tower is never called), 2 'MLP' predict over the tower output [F]. */
static int64_t neumf_param_count_m(int F, int L, int mode)
{
    int64_t D = (int64_t)F << (L - 1), n = 0, in = 2 * D;
    for (int l = 0; l < L; l++) { n += in * (in / 2) + in / 2; in /= 2; }
    return n + (mode == 0 ? 2 : 1) * F + 1;
}
static int64_t neumf_param_count(int F, int L) { return neumf_param_count_m(F, L, 0); }

int64_t orc_neumf_param_count(int32_t F, int32_t L) { return neumf_param_count(F, L); }
int64_t orc_neumf_param_count_ex(int32_t F, int32_t L, int32_t mode) { return neumf_param_count_m(F, L, mode); }

/* forward of one (u, item) pair; acts[] receives the activations of every layer (concatenated), returns pred */
/* keep (train mode, nn.Dropout in front of every Linear, :61): this row's keep factors (0 or 1/(1-p)), the L layer inputs
 * concatenated [2D, D, ..., 2F]; NULL = no dropout.  x0 and the hidden activations are stored AFTER the mask (the next
 * Linear's input); a dropped or non-positive unit is 0 there, which is all relu' needs in the backward pass. */
static float neumf_forward_ex(const float *UG, const float *IG, const float *UM, const float *IM, const float *W, int F,
                              int L, int u, int it, float *x0, float *acts, float *gmf, int mode, const float *keep)
{
    int D = F << (L - 1);
    for (int f = 0; f < F; f++) gmf[f] = UG[(int64_t)u * F + f] * IG[(int64_t)it * F + f];
    for (int d = 0; d < D; d++) { x0[d] = UM[(int64_t)u * D + d]; x0[D + d] = IM[(int64_t)it * D + d]; }
    if (keep) for (int k = 0; k < 2 * D; k++) x0[k] *= keep[k];
    const float *in = x0;
    int n_in = 2 * D;
    const float *w = W;
    float *out = acts;
    const float *kp = keep ? keep + 2 * D : NULL;
    for (int l = 0; l < L; l++) {
        int n_out = n_in / 2;
        const float *b = w + (int64_t)n_in * n_out;
        for (int o = 0; o < n_out; o++) {
            double acc = 0.0;
            for (int k = 0; k < n_in; k++) acc += (double)(w[(int64_t)o * n_in + k] * in[k]);
            float z = (float)acc + b[o];
            out[o] = z > 0.f ? z : 0.f;
            if (kp && l + 1 < L) out[o] *= kp[o];
        }
        if (kp) kp += n_out;
        w = b + n_out;
        in = out;
        out += n_out;
        n_in = n_out;
    }
    /* predict layer (:126-137): over cat(GMF, tower out) / GMF / tower out, then bias */
    double acc = 0.0;
    const int hoff = mode == 0 ? F : 0;
    if (mode != 2) for (int f = 0; f < F; f++) acc += (double)(w[f] * gmf[f]);
    if (mode != 1) for (int f = 0; f < F; f++) acc += (double)(w[hoff + f] * in[f]);
    return (float)acc + w[(mode == 0 ? 2 : 1) * F];
}

static float neumf_forward_one(const float *UG, const float *IG, const float *UM, const float *IM, const float *W, int F,
                               int L, int u, int it, float *x0, float *acts, float *gmf)
{
    return neumf_forward_ex(UG, IG, UM, IM, W, F, L, u, it, x0, acts, gmf, 0, NULL);
}

void orc_neumf_predict(const float *UG, const float *IG, const float *UM, const float *IM, const float *W, int32_t F,
                       int32_t L, const int32_t *u, const int32_t *it, int64_t n, float *out)
{
    int D = F << (L - 1);
    float *x0 = (float *)malloc(sizeof(float) * (size_t)(2 * D + 2 * D + F));
    for (int64_t t = 0; t < n; t++) out[t] = neumf_forward_one(UG, IG, UM, IM, W, F, L, u[t], it[t], x0, x0 + 2 * D, x0 + 4 * D);
    free(x0);
}

/* Adam / SGD on a flat block given fp64 gradients */
static void update_block(float *theta, float *m, float *v, const double *g, int64_t n, const orc_hyper *h, int64_t step_count)
{
    if (h->opt == 0) {
        for (int64_t k = 0; k < n; k++) theta[k] = theta[k] - h->lr * (float)g[k];
    } else {
        adam_dense(theta, m, v, g, n, h, step_count);
    }
}

/* One NeuMF BPR step.  tables: UG[U,F] IG[I,F] UM[U,D] IM[I,D]; W flat tower block.  state (Adam): m,v blocks in the
 * order UG, IG, UM, IM, W (each table-sized), may be NULL for SGD.  Returns the loss. */
double orc_neumf_bpr_step_ex(float *UG, float *IG, float *UM, float *IM, float *W, int32_t U, int32_t I, int32_t F, int32_t L,
                             const int32_t *bu, const int32_t *bi, const int32_t *bj, int64_t B, const orc_hyper *h,
                             int32_t apply, float **m, float **v, int64_t step_count, int32_t mode, const float *keep);

double orc_neumf_bpr_step(float *UG, float *IG, float *UM, float *IM, float *W, int32_t U, int32_t I, int32_t F, int32_t L,
                          const int32_t *bu, const int32_t *bi, const int32_t *bj, int64_t B, const orc_hyper *h,
                          int32_t apply, float **m, float **v, int64_t step_count)
{
    return orc_neumf_bpr_step_ex(UG, IG, UM, IM, W, U, I, F, L, bu, bi, bj, B, h, apply, m, v, step_count, 0, NULL);
}

/* mode as above; keep: [2B, n_keep] keep factors (0 or 1/(1-p)), row r = side * B + t (side 0 = the pos forward, which
 * draws its masks first, 1 = the neg forward), n_keep = sum of the L layer input widths; NULL = eval / dropout 0. */
double orc_neumf_bpr_step_ex(float *UG, float *IG, float *UM, float *IM, float *W, int32_t U, int32_t I, int32_t F, int32_t L,
                             const int32_t *bu, const int32_t *bi, const int32_t *bj, int64_t B, const orc_hyper *h,
                             int32_t apply, float **m, float **v, int64_t step_count, int32_t mode, const float *keep)
{
    const float gamma = 1e-10f;
    const int D = F << (L - 1);
    int n_act = 0, n_keep = 0;
    { int in = 2 * D; for (int l = 0; l < L; l++) { n_act += in / 2; n_keep += in; in /= 2; } }
    const int use_g = mode != 2, use_h = mode != 1, hoff = mode == 0 ? F : 0, pw = (mode == 0 ? 2 : 1) * F;
    const int64_t nW = neumf_param_count_m(F, L, mode);
    /* per-row storage of x0 / activations / gmf for pos and neg */
    size_t per = (size_t)(2 * D + n_act + F);
    float *buf = (float *)malloc(sizeof(float) * per * 2 * (size_t)(B > 0 ? B : 1));
    float *coef = (float *)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
    double bpr = 0, l1[5] = {0, 0, 0, 0, 0}, s2[5] = {0, 0, 0, 0, 0};   /* UG_u, UM_u, IG_i, IM_i, IG_j */
    for (int64_t t = 0; t < B; t++) {
        float *rp = buf + per * (size_t)(2 * t), *rn = rp + per;
        const float *kpos = (keep && use_h) ? keep + (int64_t)t * n_keep : NULL;
        const float *kneg = (keep && use_h) ? keep + (int64_t)(B + t) * n_keep : NULL;
        float pos = neumf_forward_ex(UG, IG, UM, IM, W, F, L, bu[t], bi[t], rp, rp + 2 * D, rp + 2 * D + n_act, mode, kpos);
        float neg = neumf_forward_ex(UG, IG, UM, IM, W, F, L, bu[t], bj[t], rn, rn + 2 * D, rn + 2 * D + n_act, mode, kneg);
        float x = pos - neg;
        float s = 1.f / (1.f + expf(-x));
        bpr += (double)(-logf(gamma + s));
        coef[t] = -(s * (1.f - s)) / (gamma + s);
        const float *rows[5] = {UG + (int64_t)bu[t] * F, UM + (int64_t)bu[t] * D, IG + (int64_t)bi[t] * F,
                                IM + (int64_t)bi[t] * D, IG + (int64_t)bj[t] * F};
        const int len[5] = {F, D, F, D, F};
        for (int q = 0; q < 5; q++)
            for (int f = 0; f < len[q]; f++) { l1[q] += fabsf(rows[q][f]); s2[q] += (double)(rows[q][f] * rows[q][f]); }
    }
    double nr[5];
    for (int q = 0; q < 5; q++) nr[q] = sqrt(s2[q]);
    float loss = (float)bpr;                                             /* :154-167, fp32 scalar adds in order */
    loss += h->reg_1 * ((float)l1[2] + (float)l1[4]);
    loss += h->reg_1 * ((float)l1[3] + (float)l1[4]);
    loss += h->reg_2 * ((float)nr[2] + (float)nr[4]);
    loss += h->reg_2 * ((float)nr[3] + (float)nr[4]);
    loss += h->reg_1 * (float)l1[0];
    loss += h->reg_1 * (float)l1[1];
    loss += h->reg_2 * (float)nr[0];
    loss += h->reg_2 * (float)nr[1];
    if (!apply || isnan(loss)) { free(buf); free(coef); return (double)loss; }

    double *gUG = (double *)calloc((size_t)U * F, sizeof(double)), *gIG = (double *)calloc((size_t)I * F, sizeof(double));
    double *gUM = (double *)calloc((size_t)U * D, sizeof(double)), *gIM = (double *)calloc((size_t)I * D, sizeof(double));
    double *gW = (double *)calloc((size_t)nW, sizeof(double));
    float *dcur = (float *)malloc(sizeof(float) * (size_t)(2 * D)), *dprev = (float *)malloc(sizeof(float) * (size_t)(2 * D));
    float inv[5];
    for (int q = 0; q < 5; q++) inv[q] = nr[q] > 0 ? (float)(1.0 / nr[q]) : 0.f;
    /* offsets of each layer's weight inside W */
    int64_t woff[16]; int nin[16];
    { int64_t o = 0; int in = 2 * D; for (int l = 0; l < L; l++) { woff[l] = o; nin[l] = in; o += (int64_t)in * (in / 2) + in / 2; in /= 2; } woff[L] = o; }
    const float *wp = W + woff[L];
    for (int64_t t = 0; t < B; t++) {
        for (int side = 0; side < 2; side++) {
            const float *r = buf + per * (size_t)(2 * t + side);
            const float *x0 = r, *acts = r + 2 * D, *gmf = r + 2 * D + n_act;
            int u = bu[t], it = side == 0 ? bi[t] : bj[t];
            float dp = side == 0 ? coef[t] : -coef[t];                     /* dL/dpred */
            /* predict layer */
            const float *hL = acts + (n_act - F);
            const float *kr = (keep && use_h) ? keep + (int64_t)(side * B + t) * n_keep : NULL;
            for (int f = 0; f < F; f++) {
                if (use_g) {
                    gW[woff[L] + f] += (double)(dp * gmf[f]);
                    gUG[(int64_t)u * F + f] += (double)(dp * wp[f] * IG[(int64_t)it * F + f]);
                    gIG[(int64_t)it * F + f] += (double)(dp * wp[f] * UG[(int64_t)u * F + f]);
                }
                if (use_h) gW[woff[L] + hoff + f] += (double)(dp * hL[f]);
            }
            gW[woff[L] + pw] += (double)dp;
            if (!use_h) continue;                                         /* 'GMF': the tower is not part of the graph */
            /* tower backward */
            for (int f = 0; f < F; f++) dcur[f] = dp * wp[hoff + f];
            int a_off = n_act;                                            /* end of layer l's activations */
            int k_off = n_keep;                                           /* end of layer l's keep factors */
            for (int l = L - 1; l >= 0; l--) {
                int n_in = nin[l], n_out = n_in / 2;
                a_off -= n_out;
                const float *out = acts + a_off;
                const float *in = (l == 0) ? x0 : acts + (a_off - n_in);
                const float *w = W + woff[l];
                for (int o = 0; o < n_out; o++) if (!(out[o] > 0.f)) dcur[o] = 0.f;   /* relu' */
                for (int k = 0; k < n_in; k++) dprev[k] = 0.f;
                for (int o = 0; o < n_out; o++) {
                    float dz = dcur[o];
                    gW[woff[l] + (int64_t)n_in * n_out + o] += (double)dz;
                    for (int k = 0; k < n_in; k++) {
                        gW[woff[l] + (int64_t)o * n_in + k] += (double)(dz * in[k]);
                        dprev[k] += dz * w[(int64_t)o * n_in + k];
                    }
                }
                k_off -= n_in;
                if (kr) for (int k = 0; k < n_in; k++) dprev[k] *= kr[k_off + k];   /* through this layer's Dropout */
                float *tmp = dcur; dcur = dprev; dprev = tmp;
            }
            for (int d = 0; d < D; d++) {
                gUM[(int64_t)u * D + d] += (double)dcur[d];
                gIM[(int64_t)it * D + d] += (double)dcur[D + d];
            }
        }
        /* regulariser gradients (per occurrence), with the :158/:160 quirk: IG_j twice, IM_j never */
        int u = bu[t], i = bi[t], j = bj[t];
        for (int f = 0; f < F; f++) {
            float a = UG[(int64_t)u * F + f], b = IG[(int64_t)i * F + f], c = IG[(int64_t)j * F + f];
            gUG[(int64_t)u * F + f] += (double)(h->reg_1 * ((a > 0) - (a < 0))) + (double)(h->reg_2 * a * inv[0]);
            gIG[(int64_t)i * F + f] += (double)(h->reg_1 * ((b > 0) - (b < 0))) + (double)(h->reg_2 * b * inv[2]);
            gIG[(int64_t)j * F + f] += 2.0 * ((double)(h->reg_1 * ((c > 0) - (c < 0))) + (double)(h->reg_2 * c * inv[4]));
        }
        for (int d = 0; d < D; d++) {
            float a = UM[(int64_t)u * D + d], b = IM[(int64_t)i * D + d];
            gUM[(int64_t)u * D + d] += (double)(h->reg_1 * ((a > 0) - (a < 0))) + (double)(h->reg_2 * a * inv[1]);
            gIM[(int64_t)i * D + d] += (double)(h->reg_1 * ((b > 0) - (b < 0))) + (double)(h->reg_2 * b * inv[3]);
        }
    }
    update_block(UG, m ? m[0] : NULL, v ? v[0] : NULL, gUG, (int64_t)U * F, h, step_count);
    update_block(IG, m ? m[1] : NULL, v ? v[1] : NULL, gIG, (int64_t)I * F, h, step_count);
    update_block(UM, m ? m[2] : NULL, v ? v[2] : NULL, gUM, (int64_t)U * D, h, step_count);
    update_block(IM, m ? m[3] : NULL, v ? v[3] : NULL, gIM, (int64_t)I * D, h, step_count);
    update_block(W, m ? m[4] : NULL, v ? v[4] : NULL, gW, nW, h, step_count);
    free(buf); free(coef); free(gUG); free(gIG); free(gUM); free(gIM); free(gW); free(dcur); free(dprev);
    return (double)loss;
}

/* NeuMF.rank :178-209 / full_rank :211-232: scores through the full tower; ties by lower position / item id */
void orc_neumf_rank_ex(const float *UG, const float *IG, const float *UM, const float *IM, const float *W, int32_t F,
                       int32_t L, const int64_t *users, int64_t n_users, const int64_t *cands, int32_t cand_num,
                       int32_t item_num, int32_t topk, float *out_f, int64_t *out_i, int32_t mode);

void orc_neumf_rank(const float *UG, const float *IG, const float *UM, const float *IM, const float *W, int32_t F, int32_t L,
                    const int64_t *users, int64_t n_users, const int64_t *cands, int32_t cand_num, int32_t item_num,
                    int32_t topk, float *out_f, int64_t *out_i)
{
    orc_neumf_rank_ex(UG, IG, UM, IM, W, F, L, users, n_users, cands, cand_num, item_num, topk, out_f, out_i, 0);
}

void orc_neumf_rank_ex(const float *UG, const float *IG, const float *UM, const float *IM, const float *W, int32_t F,
                       int32_t L, const int64_t *users, int64_t n_users, const int64_t *cands, int32_t cand_num,
                       int32_t item_num, int32_t topk, float *out_f, int64_t *out_i, int32_t mode)
{
    int D = F << (L - 1);
    int cnt = cands ? cand_num : item_num;
    orc_sc *sc = (orc_sc *)malloc(sizeof(orc_sc) * (size_t)cnt);
    float *x0 = (float *)malloc(sizeof(float) * (size_t)(4 * D + F));
    for (int64_t r = 0; r < n_users; r++) {
        for (int k = 0; k < cnt; k++) {
            int it = cands ? (int)cands[r * cand_num + k] : k;
            sc[k].s = neumf_forward_ex(UG, IG, UM, IM, W, F, L, (int)users[r], it, x0, x0 + 2 * D, x0 + 4 * D, mode, NULL);
            sc[k].pos = k;
        }
        qsort(sc, (size_t)cnt, sizeof(orc_sc), sc_desc);
        for (int k = 0; k < topk && k < cnt; k++) {
            if (cands) out_f[r * topk + k] = (float)cands[r * cand_num + sc[k].pos];
            else out_i[r * topk + k] = sc[k].pos;
        }
    }
    free(sc); free(x0);
}
