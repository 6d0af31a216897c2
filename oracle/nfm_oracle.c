/*
 * oracle/nfm_oracle.c -- CPU restatement of NFM (daisy/model/NFMRecommender.py), with NGCF the rank-4 row of SURVEY.md 8(f).
 *
 * TEST INFRASTRUCTURE ONLY (same rules as bpr_oracle.c).  Pinned by tests/golden/nfm.npz, generated from the reference by
 * oracle/gen_golden.py with dropout = 0 (the reference's masks come from torch's RNG).
 *
 * Model (:110-123): e = P[u] * Q[i]  ->  [BatchNorm1d(F)]  ->  L x { Linear(F, F) -> [BatchNorm1d(F)] -> act }  ->
 *   fm = h_L + (u_bias[u] + i_bias[i] + bias_)   (the scalar is broadcast over all F columns, :120)  ->  pred = <wp, fm>
 *   (prediction: Linear(F, 1, bias=False), :90).
 * Training mode (calc_loss, :125-151): the positive and the negative batch are two separate forward calls, so every
 *   BatchNorm uses the statistics of ITS call (biased variance, eps 1e-5) and updates its running statistics twice per step
 *   (momentum 0.1, unbiased variance), positive call first.  Loss = BPR(sum) + the FM-style regulariser on the factor rows.
 * Eval mode (rank / full_rank / predict, :153-209): BatchNorm uses the running statistics.
 *
 * Parameter block N (flat fp32, module registration order :64-90): [gamma0, beta0 (FM_layers BN, if batch_norm)],
 *   per layer: W [F, F] (out, in), b [F], [gamma, beta], then wp [F].   Running statistics R: per BatchNorm mean [F], var [F].
 * act: 0 relu, 1 sigmoid, 2 tanh.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float lr, reg_1, reg_2;
    int32_t opt;
    float beta1, beta2, eps;
    int32_t loss;
} orc_hyper;

float orc_dot(const float *a, const float *b, int32_t F);
void orc_dense_update(float *theta, float *m, float *v, const double *g, int64_t n, const orc_hyper *h, int64_t step_count);

#define NFM_MAXL 8
#define BN_EPS 1e-5f

int64_t orc_nfm_param_count(int32_t F, int32_t L, int32_t bn)
{
    return (bn ? 2 * F : 0) + (int64_t)L * ((int64_t)F * F + F + (bn ? 2 * F : 0)) + F;
}

static float act_f(int act, float z)
{
    if (act == 0) return z > 0.f ? z : 0.f;
    if (act == 1) return 1.f / (1.f + expf(-z));
    return tanhf(z);
}

static float act_grad(int act, float z, float h)
{
    if (act == 0) return z > 0.f ? 1.f : 0.f;
    if (act == 1) return h * (1.f - h);
    return 1.f - h * h;
}

/* BatchNorm1d forward.  train: batch statistics (+ running update); eval: running statistics.
 * xhat and inv_std are kept for the backward pass (train only). */
static void bn_forward(const float *x, int64_t B, int F, const float *gamma, const float *beta, float *rm, float *rv, int train,
                       float *y, float *xhat, float *inv_std)
{
    for (int f = 0; f < F; f++) {
        float mean, var;
        if (train) {
            double s = 0.0, ss = 0.0;
            for (int64_t r = 0; r < B; r++) s += (double)x[r * F + f];
            mean = (float)(s / (double)B);
            for (int64_t r = 0; r < B; r++) {
                double d = (double)x[r * F + f] - (double)mean;
                ss += d * d;
            }
            var = (float)(ss / (double)B);
            float unbiased = B > 1 ? (float)(ss / (double)(B - 1)) : var;
            rm[f] = (1.f - 0.1f) * rm[f] + 0.1f * mean;
            rv[f] = (1.f - 0.1f) * rv[f] + 0.1f * unbiased;
        } else {
            mean = rm[f];
            var = rv[f];
        }
        float is = 1.f / sqrtf(var + BN_EPS);
        if (inv_std) inv_std[f] = is;
        for (int64_t r = 0; r < B; r++) {
            float xh = (x[r * F + f] - mean) * is;
            if (xhat) xhat[r * F + f] = xh;
            y[r * F + f] = xh * gamma[f] + beta[f];
        }
    }
}

/* dx = inv_std / B * (B dxh - sum(dxh) - xhat sum(dxh xhat)),  dxh = dy gamma */
static void bn_backward(const float *dy, const float *xhat, const float *inv_std, const float *gamma, int64_t B, int F, float *dx,
                        double *dgamma, double *dbeta)
{
    for (int f = 0; f < F; f++) {
        double s1 = 0.0, s2 = 0.0, dg = 0.0, db = 0.0;
        for (int64_t r = 0; r < B; r++) {
            double d = (double)dy[r * F + f];
            db += d;
            dg += d * (double)xhat[r * F + f];
            s1 += d * (double)gamma[f];
            s2 += d * (double)gamma[f] * (double)xhat[r * F + f];
        }
        dgamma[f] += dg;
        dbeta[f] += db;
        for (int64_t r = 0; r < B; r++) {
            double dxh = (double)dy[r * F + f] * (double)gamma[f];
            dx[r * F + f] = (float)((double)inv_std[f] / (double)B * ((double)B * dxh - s1 - (double)xhat[r * F + f] * s2));
        }
    }
}

typedef struct {
    float *e, *h0;                         /* e = p*q, h0 = FM_layers(e) before its Dropout */
    float *h0d, *hd[NFM_MAXL];             /* the Dropout outputs (what the next module reads); == h0 / h[l] without dropout */
    float *xh0, is0[512];
    float *zpre[NFM_MAXL], *z[NFM_MAXL], *h[NFM_MAXL], *xh[NFM_MAXL];   /* Linear out, (BN out =) act input, act out, BN xhat */
    float *is[NFM_MAXL];
    float *fm, *pred;
} nfm_pass;

/* keep (optional, training only): the factors nn.Dropout multiplies with, 0 or 1/(1-p), layout [site][B][F] with site 0 = the
 * Dropout of FM_layers (:67) and site 1+l = the Dropout behind activation l (:88) */
static void nfm_forward(const float *P, const float *Q, const float *bias, const float *N, float *R, int32_t U, int32_t I, int32_t F,
                        int32_t L, int32_t bn, int32_t act, const int32_t *bu, const int32_t *bi, int64_t B, int train, nfm_pass *a,
                        const float *keep)
{
    size_t sz = (size_t)B * F;
    const float *w = N;
    float *r = R;
    a->e = (float *)malloc(sizeof(float) * sz);
    a->h0 = (float *)malloc(sizeof(float) * sz);
    a->xh0 = (float *)malloc(sizeof(float) * sz);
    for (int64_t t = 0; t < B; t++)
        for (int f = 0; f < F; f++) a->e[t * F + f] = P[(int64_t)bu[t] * F + f] * Q[(int64_t)bi[t] * F + f];
    if (bn) {
        bn_forward(a->e, B, F, w, w + F, r, r + F, train, a->h0, a->xh0, a->is0);
        w += 2 * F;
        r += 2 * F;
    } else {
        memcpy(a->h0, a->e, sizeof(float) * sz);
    }
    a->h0d = (float *)malloc(sizeof(float) * sz);
    for (size_t k = 0; k < sz; k++) a->h0d[k] = keep ? a->h0[k] * keep[k] : a->h0[k];
    const float *hin = a->h0d;
    for (int l = 0; l < L; l++) {
        const float *W = w, *b = w + (size_t)F * F;
        w = b + F;
        a->zpre[l] = (float *)malloc(sizeof(float) * sz);
        a->z[l] = (float *)malloc(sizeof(float) * sz);
        a->h[l] = (float *)malloc(sizeof(float) * sz);
        a->xh[l] = (float *)malloc(sizeof(float) * sz);
        a->is[l] = (float *)malloc(sizeof(float) * (size_t)F);
        for (int64_t t = 0; t < B; t++)
            for (int o = 0; o < F; o++) {
                double acc = 0.0;
                for (int k = 0; k < F; k++) acc += (double)(hin[t * F + k] * W[(size_t)o * F + k]);
                a->zpre[l][t * F + o] = (float)acc + b[o];
            }
        if (bn) {
            bn_forward(a->zpre[l], B, F, w, w + F, r, r + F, train, a->z[l], a->xh[l], a->is[l]);
            w += 2 * F;
            r += 2 * F;
        } else {
            memcpy(a->z[l], a->zpre[l], sizeof(float) * sz);
        }
        for (size_t k = 0; k < sz; k++) a->h[l][k] = act_f(act, a->z[l][k]);
        a->hd[l] = (float *)malloc(sizeof(float) * sz);
        for (size_t k = 0; k < sz; k++) a->hd[l][k] = keep ? a->h[l][k] * keep[(size_t)(1 + l) * sz + k] : a->h[l][k];
        hin = a->hd[l];
    }
    const float *wp = w;
    a->fm = (float *)malloc(sizeof(float) * sz);
    a->pred = (float *)malloc(sizeof(float) * (size_t)B);
    for (int64_t t = 0; t < B; t++) {
        const float bsum = (bias[bu[t]] + bias[U + bi[t]]) + bias[U + I];       /* :120 */
        double acc = 0.0;
        for (int f = 0; f < F; f++) {
            a->fm[t * F + f] = hin[t * F + f] + bsum;
            acc += (double)(a->fm[t * F + f] * wp[f]);
        }
        a->pred[t] = (float)acc;
    }
}

static void nfm_free(nfm_pass *a, int L)
{
    free(a->e); free(a->h0); free(a->h0d); free(a->xh0); free(a->fm); free(a->pred);
    for (int l = 0; l < L; l++) { free(a->zpre[l]); free(a->z[l]); free(a->h[l]); free(a->hd[l]); free(a->xh[l]); free(a->is[l]); }
}

/* eval-mode scores of (users[k], items[k]) pairs: forward() under model.eval() (:153-209) */
void orc_nfm_scores(const float *P, const float *Q, const float *bias, const float *N, const float *R, int32_t U, int32_t I,
                    int32_t F, int32_t L, int32_t bn, int32_t act, const int32_t *users, const int32_t *items, int64_t n,
                    float *scores)
{
    nfm_pass a;
    float *Rc = NULL;
    if (bn) {
        size_t nr = (size_t)(1 + L) * 2 * F;
        Rc = (float *)malloc(sizeof(float) * nr);
        memcpy(Rc, R, sizeof(float) * nr);
    }
    nfm_forward(P, Q, bias, N, Rc, U, I, F, L, bn, act, users, items, n, 0, &a, NULL);
    memcpy(scores, a.pred, sizeof(float) * (size_t)n);
    nfm_free(&a, L);
    free(Rc);
}

static void nfm_backward(const float *P, const float *Q, const float *N, int32_t U, int32_t F, int32_t L, int32_t bn, int32_t act,
                         const int32_t *bu, const int32_t *bi, int64_t B, const nfm_pass *a, const float *dpred, double *gP,
                         double *gQ, double *gbias, int32_t I, double *gN, const float *keep)
{
    size_t sz = (size_t)B * F;
    /* offsets inside N */
    int64_t o_bn0 = 0, o = bn ? 2 * F : 0, oW[NFM_MAXL], oBN[NFM_MAXL];
    for (int l = 0; l < L; l++) { oW[l] = o; o += (int64_t)F * F + F; oBN[l] = o; if (bn) o += 2 * F; }
    const int64_t o_wp = o;
    const float *wp = N + o_wp;
    float *dh = (float *)malloc(sizeof(float) * sz), *tmp = (float *)malloc(sizeof(float) * sz);
    for (int64_t t = 0; t < B; t++) {
        double bs = 0.0;
        for (int f = 0; f < F; f++) {
            float d = dpred[t] * wp[f];
            dh[t * F + f] = d;
            bs += (double)d;
            gN[o_wp + f] += (double)(dpred[t] * a->fm[t * F + f]);
        }
        gbias[bu[t]] += bs;
        gbias[U + bi[t]] += bs;
        gbias[U + I] += bs;
    }
    for (int l = L - 1; l >= 0; l--) {
        const float *W = N + oW[l];
        const float *hin = l == 0 ? a->h0d : a->hd[l - 1];
        if (keep)                                                                                    /* Dropout backward */
            for (size_t k = 0; k < sz; k++) dh[k] = dh[k] * keep[(size_t)(1 + l) * sz + k];
        for (size_t k = 0; k < sz; k++) tmp[k] = dh[k] * act_grad(act, a->z[l][k], a->h[l][k]);      /* d act input */
        if (bn) {
            bn_backward(tmp, a->xh[l], a->is[l], N + oBN[l], B, F, dh, gN + oBN[l], gN + oBN[l] + F);   /* dh := d Linear out */
        } else {
            memcpy(dh, tmp, sizeof(float) * sz);
        }
        for (int64_t t = 0; t < B; t++)
            for (int oo = 0; oo < F; oo++) {
                float d = dh[t * F + oo];
                if (d == 0.f) continue;
                gN[oW[l] + (int64_t)F * F + oo] += (double)d;
                for (int k = 0; k < F; k++) gN[oW[l] + (int64_t)oo * F + k] += (double)(d * hin[t * F + k]);
            }
        for (int64_t t = 0; t < B; t++)
            for (int k = 0; k < F; k++) {
                double acc = 0.0;
                for (int oo = 0; oo < F; oo++) acc += (double)(dh[t * F + oo] * W[(size_t)oo * F + k]);
                tmp[t * F + k] = (float)acc;
            }
        memcpy(dh, tmp, sizeof(float) * sz);
    }
    if (keep)
        for (size_t k = 0; k < sz; k++) dh[k] = dh[k] * keep[k];
    if (bn) {
        bn_backward(dh, a->xh0, a->is0, N + o_bn0, B, F, tmp, gN + o_bn0, gN + o_bn0 + F);
        memcpy(dh, tmp, sizeof(float) * sz);
    }
    for (int64_t t = 0; t < B; t++)
        for (int f = 0; f < F; f++) {
            gP[(int64_t)bu[t] * F + f] += (double)(dh[t * F + f] * Q[(int64_t)bi[t] * F + f]);
            gQ[(int64_t)bi[t] * F + f] += (double)(dh[t * F + f] * P[(int64_t)bu[t] * F + f]);
        }
    free(dh);
    free(tmp);
}

/* One NFM BPR step == calc_loss (:125-151) + backward + optimizer.step.  state (optional): m then v over [P | Q | bias | N].
 * R (running statistics) is updated by the two training-mode forward calls.  Returns the fp32 loss. */
double orc_nfm_bpr_step_ex(float *P, float *Q, float *bias, float *N, float *R, int32_t U, int32_t I, int32_t F, int32_t L,
                           int32_t bn, int32_t act, const int32_t *bu, const int32_t *bi, const int32_t *bj, int64_t B,
                           const orc_hyper *h, int32_t apply, float *state, int64_t step_count, const float *keep_pos,
                           const float *keep_neg);

double orc_nfm_bpr_step(float *P, float *Q, float *bias, float *N, float *R, int32_t U, int32_t I, int32_t F, int32_t L,
                        int32_t bn, int32_t act, const int32_t *bu, const int32_t *bi, const int32_t *bj, int64_t B,
                        const orc_hyper *h, int32_t apply, float *state, int64_t step_count)
{
    return orc_nfm_bpr_step_ex(P, Q, bias, N, R, U, I, F, L, bn, act, bu, bi, bj, B, h, apply, state, step_count, NULL, NULL);
}

/* keep_pos / keep_neg (optional): the Dropout factors of the positive / the negative forward call ([1 + L][B][F], 0 or 1/(1-p)),
 * drawn by the caller in the reference's order (dropout = config['dropout'] > 0, :67,:88) */
double orc_nfm_bpr_step_ex(float *P, float *Q, float *bias, float *N, float *R, int32_t U, int32_t I, int32_t F, int32_t L,
                           int32_t bn, int32_t act, const int32_t *bu, const int32_t *bi, const int32_t *bj, int64_t B,
                           const orc_hyper *h, int32_t apply, float *state, int64_t step_count, const float *keep_pos,
                           const float *keep_neg)
{
    const float gamma = 1e-10f;
    if (F > 512) return NAN;
    nfm_pass pa, na;
    nfm_forward(P, Q, bias, N, R, U, I, F, L, bn, act, bu, bi, B, 1, &pa, keep_pos);
    nfm_forward(P, Q, bias, N, bn ? R : NULL, U, I, F, L, bn, act, bu, bj, B, 1, &na, keep_neg);
    float *c = (float *)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1)), *cn = (float *)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
    double bpr = 0, l1u = 0, l1i = 0, l1j = 0, s2u = 0, s2i = 0, s2j = 0;
    for (int64_t t = 0; t < B; t++) {
        float x = pa.pred[t] - na.pred[t];
        float s = 1.f / (1.f + expf(-x));
        bpr += (double)(-logf(gamma + s));
        c[t] = -(s * (1.f - s)) / (gamma + s);
        cn[t] = -c[t];
        const float *p = P + (int64_t)bu[t] * F, *qi = Q + (int64_t)bi[t] * F, *qj = Q + (int64_t)bj[t] * F;
        for (int f = 0; f < F; f++) {
            l1u += fabsf(p[f]); s2u += (double)(p[f] * p[f]);
            l1i += fabsf(qi[f]); s2i += (double)(qi[f] * qi[f]);
            l1j += fabsf(qj[f]); s2j += (double)(qj[f] * qj[f]);
        }
    }
    double nu = sqrt(s2u), ni = sqrt(s2i), nj = sqrt(s2j);
    float loss = (float)bpr;
    loss += h->reg_1 * ((float)l1i + (float)l1j);      /* :141 */
    loss += h->reg_2 * ((float)ni + (float)nj);        /* :142 */
    loss += h->reg_1 * (float)l1u;                     /* :148 */
    loss += h->reg_2 * (float)nu;                      /* :149 */
    if (!apply || isnan(loss)) {
        free(c); free(cn); nfm_free(&pa, L); nfm_free(&na, L);
        return (double)loss;
    }
    const int64_t nP = (int64_t)U * F, nQ = (int64_t)I * F, nB = (int64_t)U + I + 1, nN = orc_nfm_param_count(F, L, bn);
    double *gP = (double *)calloc((size_t)nP, sizeof(double)), *gQ = (double *)calloc((size_t)nQ, sizeof(double));
    double *gb = (double *)calloc((size_t)nB, sizeof(double)), *gN = (double *)calloc((size_t)nN, sizeof(double));
    nfm_backward(P, Q, N, U, F, L, bn, act, bu, bi, B, &pa, c, gP, gQ, gb, I, gN, keep_pos);
    nfm_backward(P, Q, N, U, F, L, bn, act, bu, bj, B, &na, cn, gP, gQ, gb, I, gN, keep_neg);
    float inu = nu > 0 ? (float)(1.0 / nu) : 0.f, ini = ni > 0 ? (float)(1.0 / ni) : 0.f, inj = nj > 0 ? (float)(1.0 / nj) : 0.f;
    for (int64_t t = 0; t < B; t++) {
        const float *p = P + (int64_t)bu[t] * F, *qi = Q + (int64_t)bi[t] * F, *qj = Q + (int64_t)bj[t] * F;
        double *gu = gP + (int64_t)bu[t] * F, *gi = gQ + (int64_t)bi[t] * F, *gj = gQ + (int64_t)bj[t] * F;
        for (int f = 0; f < F; f++) {
            float sp = (float)((p[f] > 0) - (p[f] < 0)), si = (float)((qi[f] > 0) - (qi[f] < 0)), sj = (float)((qj[f] > 0) - (qj[f] < 0));
            gu[f] += (double)(h->reg_1 * sp) + (double)(h->reg_2 * p[f] * inu);
            gi[f] += (double)(h->reg_1 * si) + (double)(h->reg_2 * qi[f] * ini);
            gj[f] += (double)(h->reg_1 * sj) + (double)(h->reg_2 * qj[f] * inj);
        }
    }
    const int64_t tot = nP + nQ + nB + nN;
    float *m = state, *v = state ? state + tot : NULL;
    orc_dense_update(P, m, v, gP, nP, h, step_count);
    orc_dense_update(Q, m ? m + nP : NULL, v ? v + nP : NULL, gQ, nQ, h, step_count);
    orc_dense_update(bias, m ? m + nP + nQ : NULL, v ? v + nP + nQ : NULL, gb, nB, h, step_count);
    orc_dense_update(N, m ? m + nP + nQ + nB : NULL, v ? v + nP + nQ + nB : NULL, gN, nN, h, step_count);
    free(gP); free(gQ); free(gb); free(gN); free(c); free(cn);
    nfm_free(&pa, L);
    nfm_free(&na, L);
    return (double)loss;
}
