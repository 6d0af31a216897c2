/*
 * oracle/ngcf_oracle.c -- CPU restatement of NGCF (daisy/model/NGCFRecommender.py), the rank-4 row of SURVEY.md 8(f).
 *
 * TEST INFRASTRUCTURE ONLY (same rules as bpr_oracle.c).  Pinned by tests/golden/ngcf.npz, generated from the reference
 * by oracle/gen_golden.py (node_dropout = 0, mess_dropout = 0: the reference's dropout masks come from torch's RNG and,
 * for the message dropout, are drawn even at rank() time -- :164 builds a fresh nn.Dropout in training mode).
 *
 * Model (:157-172): E_0 = cat(embed_user, embed_item) [n, F]; per layer l (BiGNN, :38-59)
 *     X = A_hat E_l;  Y = (E_l + X) W1^T + b1 + (X * E_l) W2^T + b2;  Z = LeakyReLU_0.2(Y);  E_{l+1} = Z / max(||Z||_2, 1e-12)
 * (row-wise F.normalize, :165); the representation is the CONCATENATION of E_0 .. E_L (:167).  Scores are dot products of
 * concatenated rows; loss = BPR(sum) + reg on the EGO rows (:197-198), un-squared norms.  Optimiser: Adam (default) / SGD
 * on E_0 and on every W1, b1, W2, b2.
 *
 * Parameter block W (flat fp32), per layer: W1 [out, in], b1 [out], W2 [out, in], b2 [out]  (module registration order,
 * :106-108 / :46-47).  dims[0] = F, dims[l+1] = hidden size of layer l.
 * Arithmetic convention as in bpr_oracle.c: fp32 element-wise, fp64 accumulation of sums rounded once.
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

#define NGCF_MAXL 8

static void spmm(const int64_t *row_ptr, const int32_t *col, const float *val, int64_t n, int32_t F, const float *X, float *Y)
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

int64_t orc_ngcf_param_count(const int32_t *dims, int32_t L)
{
    int64_t n = 0;
    for (int l = 0; l < L; l++) n += 2 * ((int64_t)dims[l] * dims[l + 1] + dims[l + 1]);
    return n;
}

typedef struct {
    float *E[NGCF_MAXL + 1];     /* E_0 .. E_L (E_0 aliases the ego table) */
    float *X[NGCF_MAXL], *Y[NGCF_MAXL], *Z[NGCF_MAXL];
    double *rn[NGCF_MAXL];       /* max(||Z row||, 1e-12) */
} ngcf_acts;

/* keep (optional): the factors of nn.Dropout(mess_dropout) (:164), 0 or 1/(1-p), layers concatenated: [n, dims[1]], [n, dims[2]], ... */
static void ngcf_forward(const float *E0, const float *W, int64_t n, const int32_t *dims, int32_t L, const int64_t *row_ptr,
                         const int32_t *col, const float *val, ngcf_acts *a, const float *keep)
{
    const float *w = W;
    const float *kp = keep;
    a->E[0] = (float *)E0;
    for (int l = 0; l < L; l++) {
        const int in = dims[l], out = dims[l + 1];
        const float *W1 = w, *b1 = w + (size_t)in * out, *W2 = b1 + out, *b2 = W2 + (size_t)in * out;
        w = b2 + out;
        a->X[l] = (float *)malloc(sizeof(float) * (size_t)n * in);
        a->Y[l] = (float *)malloc(sizeof(float) * (size_t)n * out);
        a->Z[l] = (float *)malloc(sizeof(float) * (size_t)n * out);
        a->E[l + 1] = (float *)malloc(sizeof(float) * (size_t)n * out);
        a->rn[l] = (double *)malloc(sizeof(double) * (size_t)n);
        spmm(row_ptr, col, val, n, in, a->E[l], a->X[l]);
        for (int64_t r = 0; r < n; r++) {
            const float *e = a->E[l] + r * in, *x = a->X[l] + r * in;
            double ss = 0.0;
            for (int o = 0; o < out; o++) {
                double acc1 = 0.0, acc2 = 0.0;
                for (int k = 0; k < in; k++) {
                    acc1 += (double)((e[k] + x[k]) * W1[(size_t)o * in + k]);
                    acc2 += (double)((x[k] * e[k]) * W2[(size_t)o * in + k]);
                }
                float y = ((float)acc1 + b1[o]) + ((float)acc2 + b2[o]);       /* inter_part1 + inter_part2 (:59) */
                float z = y > 0.f ? y : 0.2f * y;
                if (kp) z = z * kp[r * out + o];                                /* message dropout (:164) */
                a->Y[l][r * out + o] = y;
                a->Z[l][r * out + o] = z;
                ss += (double)(z * z);
            }
            double nr = sqrt(ss);
            if (nr < 1e-12) nr = 1e-12;
            a->rn[l][r] = nr;
            for (int o = 0; o < out; o++) a->E[l + 1][r * out + o] = (float)((double)a->Z[l][r * out + o] / nr);
        }
        if (kp) kp += (size_t)n * out;
    }
}

static void ngcf_free(ngcf_acts *a, int32_t L)
{
    for (int l = 0; l < L; l++) {
        free(a->X[l]); free(a->Y[l]); free(a->Z[l]); free(a->E[l + 1]); free(a->rn[l]);
    }
}

/* NGCF.forward (:157-172): out [n, sum(dims)] = cat(E_0 .. E_L, dim=1) */
void orc_ngcf_forward_ex(const float *E0, const float *W, int32_t U, int32_t I, const int32_t *dims, int32_t L,
                         const int64_t *row_ptr, const int32_t *col, const float *val, float *out, const float *keep);

void orc_ngcf_forward(const float *E0, const float *W, int32_t U, int32_t I, const int32_t *dims, int32_t L,
                      const int64_t *row_ptr, const int32_t *col, const float *val, float *out)
{
    orc_ngcf_forward_ex(E0, W, U, I, dims, L, row_ptr, col, val, out, NULL);
}

/* keep: message-dropout factors (see ngcf_forward); the reference's forward() applies them whenever it runs, rank() included */
void orc_ngcf_forward_ex(const float *E0, const float *W, int32_t U, int32_t I, const int32_t *dims, int32_t L,
                         const int64_t *row_ptr, const int32_t *col, const float *val, float *out, const float *keep)
{
    int64_t n = (int64_t)U + I;
    int C = 0;
    for (int l = 0; l <= L; l++) C += dims[l];
    ngcf_acts a;
    ngcf_forward(E0, W, n, dims, L, row_ptr, col, val, &a, keep);
    for (int64_t r = 0; r < n; r++) {
        int o = 0;
        for (int l = 0; l <= L; l++) {
            memcpy(out + r * C + o, a.E[l] + r * dims[l], sizeof(float) * (size_t)dims[l]);
            o += dims[l];
        }
    }
    ngcf_free(&a, L);
}

/* One NGCF BPR step == calc_loss (:174-205) + backward + optimizer.step.  state (optional): m then v, each the size of
 * [E0 | W] (Adam), in that order.  Returns the fp32 loss. */
double orc_ngcf_bpr_step_ex(float *E0, float *W, int32_t U, int32_t I, const int32_t *dims, int32_t L, const int64_t *row_ptr,
                            const int32_t *col, const float *val, const int32_t *bu, const int32_t *bi, const int32_t *bj,
                            int64_t B, const orc_hyper *h, int32_t apply, float *state, int64_t step_count, const float *keep);

double orc_ngcf_bpr_step(float *E0, float *W, int32_t U, int32_t I, const int32_t *dims, int32_t L, const int64_t *row_ptr,
                         const int32_t *col, const float *val, const int32_t *bu, const int32_t *bi, const int32_t *bj,
                         int64_t B, const orc_hyper *h, int32_t apply, float *state, int64_t step_count)
{
    return orc_ngcf_bpr_step_ex(E0, W, U, I, dims, L, row_ptr, col, val, bu, bi, bj, B, h, apply, state, step_count, NULL);
}

double orc_ngcf_bpr_step_ex(float *E0, float *W, int32_t U, int32_t I, const int32_t *dims, int32_t L, const int64_t *row_ptr,
                            const int32_t *col, const float *val, const int32_t *bu, const int32_t *bi, const int32_t *bj,
                            int64_t B, const orc_hyper *h, int32_t apply, float *state, int64_t step_count, const float *keep)
{
    const float gamma = 1e-10f;
    const int64_t n = (int64_t)U + I;
    const int F = dims[0];
    int C = 0, off[NGCF_MAXL + 2];
    for (int l = 0; l <= L; l++) { off[l] = C; C += dims[l]; }
    ngcf_acts a;
    ngcf_forward(E0, W, n, dims, L, row_ptr, col, val, &a, keep);
    float *all = (float *)malloc(sizeof(float) * (size_t)n * C);
    for (int64_t r = 0; r < n; r++)
        for (int l = 0; l <= L; l++) memcpy(all + r * C + off[l], a.E[l] + r * dims[l], sizeof(float) * (size_t)dims[l]);
    float *coef = (float *)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
    double bpr = 0, l1u = 0, l1i = 0, l1j = 0, s2u = 0, s2i = 0, s2j = 0;
    for (int64_t t = 0; t < B; t++) {
        const float *p = all + (int64_t)bu[t] * C, *qi = all + ((int64_t)U + bi[t]) * C, *qj = all + ((int64_t)U + bj[t]) * C;
        float x = orc_dot(p, qi, C) - orc_dot(p, qj, C);
        float s = 1.f / (1.f + expf(-x));
        bpr += (double)(-logf(gamma + s));
        coef[t] = -(s * (1.f - s)) / (gamma + s);
        const float *pe = E0 + (int64_t)bu[t] * F, *qie = E0 + ((int64_t)U + bi[t]) * F, *qje = E0 + ((int64_t)U + bj[t]) * F;
        for (int f = 0; f < F; f++) {
            l1u += fabsf(pe[f]); s2u += (double)(pe[f] * pe[f]);
            l1i += fabsf(qie[f]); s2i += (double)(qie[f] * qie[f]);
            l1j += fabsf(qje[f]); s2j += (double)(qje[f] * qje[f]);
        }
    }
    double nu = sqrt(s2u), ni = sqrt(s2i), nj = sqrt(s2j);
    float loss = (float)bpr;
    loss += h->reg_1 * (((float)l1u + (float)l1i) + (float)l1j);     /* :197 */
    loss += h->reg_2 * (((float)nu + (float)ni) + (float)nj);        /* :198 */
    if (!apply || isnan(loss)) {
        free(all); free(coef); ngcf_free(&a, L);
        return (double)loss;
    }
    /* d loss / d all (dense, non-zero on batch rows) */
    double *G = (double *)calloc((size_t)n * C, sizeof(double));
    for (int64_t t = 0; t < B; t++) {
        const float *p = all + (int64_t)bu[t] * C, *qi = all + ((int64_t)U + bi[t]) * C, *qj = all + ((int64_t)U + bj[t]) * C;
        double *gu = G + (int64_t)bu[t] * C, *gi = G + ((int64_t)U + bi[t]) * C, *gj = G + ((int64_t)U + bj[t]) * C;
        float c = coef[t];
        for (int k = 0; k < C; k++) {
            gu[k] += (double)(c * (qi[k] - qj[k]));
            gi[k] += (double)(c * p[k]);
            gj[k] -= (double)(c * p[k]);
        }
    }
    const int64_t nW = orc_ngcf_param_count(dims, L);
    double *gW = (double *)calloc((size_t)nW, sizeof(double));
    /* walk the layers backwards; dE = gradient w.r.t. E_{l+1} coming from the layer above (none for the last one) */
    float *dE = NULL;
    int64_t woff[NGCF_MAXL];
    { int64_t o = 0; for (int l = 0; l < L; l++) { woff[l] = o; o += 2 * ((int64_t)dims[l] * dims[l + 1] + dims[l + 1]); } }
    int64_t koff[NGCF_MAXL];
    { int64_t o = 0; for (int l = 0; l < L; l++) { koff[l] = o; o += n * dims[l + 1]; } }
    for (int l = L - 1; l >= 0; l--) {
        const int in = dims[l], out = dims[l + 1];
        const float *kl = keep ? keep + koff[l] : NULL;
        const float *W1 = W + woff[l], *W2 = W1 + (size_t)in * out + out;
        double *gW1 = gW + woff[l], *gb1 = gW1 + (size_t)in * out, *gW2 = gb1 + out, *gb2 = gW2 + (size_t)in * out;
        float *dX = (float *)malloc(sizeof(float) * (size_t)n * in);
        float *dEl = (float *)malloc(sizeof(float) * (size_t)n * in);
        float *dY = (float *)malloc(sizeof(float) * (size_t)out);
        for (int64_t r = 0; r < n; r++) {
            const float *e = a.E[l] + r * in, *x = a.X[l] + r * in, *y = a.Y[l] + r * out, *z = a.Z[l] + r * out,
                        *nn = a.E[l + 1] + r * out;
            /* dN = block l+1 of G (+ the gradient from the layer above) */
            double dot = 0.0;
            for (int o = 0; o < out; o++) {
                double dn = G[r * C + off[l + 1] + o] + (dE ? (double)dE[r * out + o] : 0.0);
                dot += dn * (double)nn[o];
            }
            const double nr = a.rn[l][r];
            int any = 0;
            for (int o = 0; o < out; o++) {
                double dn = G[r * C + off[l + 1] + o] + (dE ? (double)dE[r * out + o] : 0.0);
                /* normalize backward: (dN - N <N, dN>) / ||Z||  (||Z|| clamped at 1e-12: the clamp branch has no N-term) */
                double dz = (nr > 1e-12) ? (dn - (double)nn[o] * dot) / nr : dn / nr;
                (void)z;
                if (kl) dz = (double)((float)dz * kl[r * out + o]);               /* Dropout backward */
                float dy = (float)dz * (y[o] > 0.f ? 1.f : 0.2f);
                dY[o] = dy;
                if (dy != 0.f) any = 1;
            }
            for (int k = 0; k < in; k++) { dX[r * in + k] = 0.f; dEl[r * in + k] = 0.f; }
            if (!any) continue;
            for (int o = 0; o < out; o++) {
                const float dy = dY[o];
                if (dy == 0.f) continue;
                gb1[o] += (double)dy;
                gb2[o] += (double)dy;
                for (int k = 0; k < in; k++) {
                    gW1[(size_t)o * in + k] += (double)(dy * (e[k] + x[k]));
                    gW2[(size_t)o * in + k] += (double)(dy * (x[k] * e[k]));
                }
            }
            for (int k = 0; k < in; k++) {
                double dS = 0.0, dT = 0.0;
                for (int o = 0; o < out; o++) {
                    dS += (double)(dY[o] * W1[(size_t)o * in + k]);
                    dT += (double)(dY[o] * W2[(size_t)o * in + k]);
                }
                dEl[r * in + k] = (float)(dS + dT * (double)x[k]);     /* through (E + X) and X * E, E side */
                dX[r * in + k] = (float)(dS + dT * (double)e[k]);      /* X side */
            }
        }
        /* X = A_hat E_l, A_hat symmetric: dE_l += A_hat dX */
        float *AdX = (float *)malloc(sizeof(float) * (size_t)n * in);
        spmm(row_ptr, col, val, n, in, dX, AdX);
        for (size_t k = 0; k < (size_t)n * in; k++) dEl[k] += AdX[k];
        free(AdX); free(dX); free(dY);
        if (dE) free(dE);
        dE = dEl;
    }
    /* gradient of the ego table: block 0 of G + the chain through layer 0 + the regulariser on the batch rows */
    double *gE = (double *)calloc((size_t)n * F, sizeof(double));
    for (int64_t r = 0; r < n; r++)
        for (int f = 0; f < F; f++) gE[r * F + f] = G[r * C + f] + (dE ? (double)dE[r * F + f] : 0.0);
    if (dE) free(dE);
    float inu = nu > 0 ? (float)(1.0 / nu) : 0.f, ini = ni > 0 ? (float)(1.0 / ni) : 0.f, inj = nj > 0 ? (float)(1.0 / nj) : 0.f;
    for (int64_t t = 0; t < B; t++) {
        int64_t rows[3] = {bu[t], (int64_t)U + bi[t], (int64_t)U + bj[t]};
        float inv[3] = {inu, ini, inj};
        for (int q = 0; q < 3; q++) {
            const float *e = E0 + rows[q] * F;
            double *g = gE + rows[q] * F;
            for (int f = 0; f < F; f++) {
                float sg = (float)((e[f] > 0) - (e[f] < 0));
                g[f] += (double)(h->reg_1 * sg) + (double)(h->reg_2 * e[f] * inv[q]);
            }
        }
    }
    const int64_t nE = n * F, tot = nE + nW;
    float *m = state, *v = state ? state + tot : NULL;
    orc_dense_update(E0, m, v, gE, nE, h, step_count);
    orc_dense_update(W, m ? m + nE : NULL, v ? v + nE : NULL, gW, nW, h, step_count);
    free(gE); free(gW); free(G); free(all); free(coef);
    ngcf_free(&a, L);
    return (double)loss;
}
