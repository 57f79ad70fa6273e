/* TEST INFRASTRUCTURE ONLY — see pst_oracle.h.  CPU restatement (not a copy) of the
 * reference's hot-path arithmetic; every routine cites the reference lines it follows.
 * Float/double placement mirrors C's implicit conversions in the reference because they
 * decide bit-exactness.  Build: gcc -O2 -ffp-contract=off (x86-64
 * baseline, no FMA), like the reference's distutils default. */
#include "pst_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXTAP 16

static float *falloc(size_t n) { return (float *)calloc(n ? n : 1, sizeof(float)); }

/* ------------------------------------------------------------------ B-filter taps */

/* dip_cfuns.c:835-855 apfilt_init: binomial-like table in double */
static void bfilt_table(int nw, double *b)
{
    int nf = 2 * nw;
    for (int k = 0; k <= nf; k++) {
        double bk = 1.0;
        for (int j = 0; j < nf; j++) {
            if (j < nf - k) bk *= (k + j + 1.0) / (2 * (2 * j + 1) * (j + 1));
            else            bk *= 1.0 / (2 * (2 * j + 1));
        }
        b[k] = bk;
    }
}

/* the j-th linear factor of tap k; evaluated in FLOAT like the reference's
 * `(nf-j-p)` / `(p+j+1)` with int nf,j and float p (dip_cfuns.c:874-878) */
static inline float bracket(int nf, int j, int k, float p)
{
    if (j < nf - k) return (float)(nf - j) - p;
    return (p + (float)j) + 1.0f;
}

/* dip_cfuns.c:863-881 */
void pso_passfilter(int nw, float p, float *a)
{
    double b[MAXTAP];
    int nf = 2 * nw;
    bfilt_table(nw, b);
    for (int k = 0; k <= nf; k++) {
        double ak = b[k];
        for (int j = 0; j < nf; j++) ak *= bracket(nf, j, k, p);
        a[k] = (float)ak;
    }
}

/* dip_cfuns.c:883-910 */
void pso_aderfilter(int nw, float p, float *a)
{
    double b[MAXTAP];
    int nf = 2 * nw;
    bfilt_table(nw, b);
    for (int k = 0; k <= nf; k++) {
        double ak = 0.;
        for (int i = 0; i < nf; i++) {
            double ai = -1.0;
            for (int j = 0; j < nf; j++) {
                if (j != i) ai *= bracket(nf, j, k, p);
                else if (j < nf - k) ai *= (-1);
            }
            ak += ai;
        }
        a[k] = (float)(ak * b[k]);
    }
}

/* ------------------------------------------------------------------ PWD stencil */

/* dip_cfuns.c:1135-1200 (allpass1) and :1399-1465 (allpass2) with left=false,
 * drift=false, nj=1: y[i] = sum_w (x[i+(w-nw)+ip] - x[i-(w-nw)]) * flt_w(p[i]) */
void pso_allpass(const float *u, const float *sg, int n1, int n2, int n3, int nw,
                 int xline, int der, float *y)
{
    size_t n = (size_t)n1 * n2 * n3;
    long ip = xline ? (long)n1 * n2 : n1;
    int e2 = xline ? n2 : n2 - 1, e3 = xline ? n3 - 1 : n3;
    float flt[MAXTAP];
    memset(y, 0, n * sizeof(float));
    for (int i3 = 0; i3 < e3; i3++)
        for (int i2 = 0; i2 < e2; i2++)
            for (int i1 = nw; i1 < n1 - nw; i1++) {
                long i = i1 + (long)n1 * (i2 + (long)n2 * i3);
                if (der) pso_aderfilter(nw, sg[i], flt);
                else     pso_passfilter(nw, sg[i], flt);
                float acc = 0.f;
                for (int w = 0; w <= 2 * nw; w++) {
                    int s = w - nw;
                    acc += (u[i + s + ip] - u[i - s]) * flt[w];
                }
                y[i] = acc;
            }
}

/* dip_cfuns.c:914-997 mask32 with both=false, nj=1: footprints touching zero samples */
static void footprint_mask(const float *um, int n1, int n2, int n3, int nw,
                           unsigned char *m_in, unsigned char *m_x)
{
    size_t n = (size_t)n1 * n2 * n3;
    unsigned char *z = (unsigned char *)malloc(n);
    for (size_t i = 0; i < n; i++) { z[i] = (um[i] == 0.f); m_in[i] = 0; m_x[i] = 0; }
    for (int i3 = 0; i3 < n3; i3++)
        for (int i2 = 0; i2 < n2 - 1; i2++)
            for (int i1 = nw; i1 < n1 - nw; i1++) {
                long i = i1 + (long)n1 * (i2 + (long)n2 * i3);
                for (int s = -nw; s <= nw; s++)
                    m_in[i] = (unsigned char)(m_in[i] || z[i - s] || z[i + n1 + s]);
            }
    for (int i3 = 0; i3 < n3 - 1; i3++)
        for (int i2 = 0; i2 < n2; i2++)
            for (int i1 = nw; i1 < n1 - nw; i1++) {
                long i = i1 + (long)n1 * (i2 + (long)n2 * i3);
                for (int s = -nw; s <= nw; s++)
                    m_x[i] = (unsigned char)(m_x[i] || z[i - s] || z[i + (long)n1 * n2 + s]);
            }
    free(z);
}

/* ------------------------------------------------------------------ triangle smoothing */

/* One line x[o + i*d], i<nx, radius nb: dip_cfuns.c:616-625 (ps_smooth2) =
 * triple2 (:564-580) -> doubint2 (:508-529) -> fold2 (:458-484).  t has nx+2nb floats. */
static void tri_line(float *x, long o, long d, int nx, int nb, float *t)
{
    int np = nx + 2 * nb;
    float wt = (float)(1.0 / (nb * nb));          /* ps_triangle_init :421 (double -> float) */
    float w2 = (float)(2. * wt);                  /* `2.*wt` passed as float a (:574) */
    float wm = -wt;
    for (int i = 0; i < np; i++) t[i] = 0;
    for (int i = 0; i < nx; i++) t[i]          += wm * x[o + i * d];   /* three saxpy sweeps */
    for (int i = 0; i < nx; i++) t[i + nb]     += w2 * x[o + i * d];
    for (int i = 0; i < nx; i++) t[i + 2 * nb] += wm * x[o + i * d];
    float s = 0.f;
    for (int i = 0; i < np; i++) { s += t[i]; t[i] = s; }             /* forward running sum */
    s = 0.f;
    for (int i = np - 1; i >= 0; i--) { s += t[i]; t[i] = s; }        /* backward running sum */
    for (int i = 0; i < nx; i++) x[o + i * d] = t[i + nb];            /* fold2: middle */
    for (int j = nb + nx; j < np; j += nx) {                          /* right reflections */
        for (int i = 0; i < nx && i < np - j; i++) x[o + (nx - 1 - i) * d] += t[j + i];
        j += nx;
        for (int i = 0; i < nx && i < np - j; i++) x[o + i * d] += t[j + i];
    }
    for (int j = nb; j >= 0; j -= nx) {                               /* left reflections */
        for (int i = 0; i < nx && i < j; i++) x[o + i * d] += t[j - 1 - i];
        j -= nx;
        for (int i = 0; i < nx && i < j; i++) x[o + (nx - 1 - i) * d] += t[j - 1 - i];
    }
}

/* dip_cfuns.c:672-712 ps_trianglen_lop body: axes 1,2,3 in turn, each line in place */
void pso_smooth3(float *x, int n1, int n2, int n3, int r1, int r2, int r3)
{
    int nmax = n1 > n2 ? n1 : n2; if (n3 > nmax) nmax = n3;
    int rmax = r1 > r2 ? r1 : r2; if (r3 > rmax) rmax = r3;
    float *t = falloc((size_t)nmax + 2 * (size_t)rmax + 2);
    if (r1 > 1)
        for (long l = 0; l < (long)n2 * n3; l++) tri_line(x, l * n1, 1, n1, r1, t);
    if (r2 > 1)
        for (int i3 = 0; i3 < n3; i3++)
            for (int i1 = 0; i1 < n1; i1++) tri_line(x, i1 + (long)n1 * n2 * i3, n1, n2, r2, t);
    if (r3 > 1)
        for (long l = 0; l < (long)n1 * n2; l++) tri_line(x, l, (long)n1 * n2, n3, r3, t);
    free(t);
}

/* One line of smoothcf (dip_cfuns.c:2084-2098) with every option:
 *   adj = 0: ps_smooth2 :616-625 = triple2 :560-577 (box: +wt at 1, -wt at 2nb), doubint2 :508-529 (forward sum,
 *            then backward unless box || der), fold2 :458-484
 *   adj = 1: ps_smooth :591-603 = fold :439-456, doubint :487-505 (backward sum, then forward unless box || der),
 *            triple :531-547 (box: (tmp[i+1] - tmp[i+2nb]) * wt in float; triangle: 2.*tmp1 - tmp - tmp2 in double)
 *   wt = 1/(2nb-1) for a box, 1/nb^2 for a triangle (ps_triangle_init :415-424) */
static void tri_line_any(float *x, long o, long d, int nx, int nb, float *t, int adj, int box, int der)
{
    int np = nx + 2 * nb;
    float wt = box ? (float)(1.0 / (2 * nb - 1)) : (float)(1.0 / (nb * nb));
    int single = box || der;
    if (!adj) {
        for (int i = 0; i < np; i++) t[i] = 0;
        if (box) {
            float wp = +wt, wm = -wt;
            for (int i = 0; i < nx; i++) t[i + 1] += wp * x[o + i * d];
            for (int i = 0; i < nx; i++) t[i + 2 * nb] += wm * x[o + i * d];
        } else {
            float w2 = (float)(2. * wt), wm = -wt;
            for (int i = 0; i < nx; i++) t[i] += wm * x[o + i * d];
            for (int i = 0; i < nx; i++) t[i + nb] += w2 * x[o + i * d];
            for (int i = 0; i < nx; i++) t[i + 2 * nb] += wm * x[o + i * d];
        }
        float s = 0.f;
        for (int i = 0; i < np; i++) { s += t[i]; t[i] = s; }
        if (!single) { s = 0.f; for (int i = np - 1; i >= 0; i--) { s += t[i]; t[i] = s; } }
        for (int i = 0; i < nx; i++) x[o + i * d] = t[i + nb];
        for (int j = nb + nx; j < np; j += nx) {
            for (int i = 0; i < nx && i < np - j; i++) x[o + (nx - 1 - i) * d] += t[j + i];
            j += nx;
            for (int i = 0; i < nx && i < np - j; i++) x[o + i * d] += t[j + i];
        }
        for (int j = nb; j >= 0; j -= nx) {
            for (int i = 0; i < nx && i < j; i++) x[o + i * d] += t[j - 1 - i];
            j -= nx;
            for (int i = 0; i < nx && i < j; i++) x[o + (nx - 1 - i) * d] += t[j - 1 - i];
        }
    } else {
        for (int i = 0; i < nx; i++) t[i + nb] = x[o + i * d];
        for (int j = nb + nx; j < np; j += nx) {
            for (int i = 0; i < nx && i < np - j; i++) t[j + i] = x[o + (nx - 1 - i) * d];
            j += nx;
            for (int i = 0; i < nx && i < np - j; i++) t[j + i] = x[o + i * d];
        }
        for (int j = nb; j >= 0; j -= nx) {
            for (int i = 0; i < nx && i < j; i++) t[j - 1 - i] = x[o + i * d];
            j -= nx;
            for (int i = 0; i < nx && i < j; i++) t[j - 1 - i] = x[o + (nx - 1 - i) * d];
        }
        float s = 0.f;
        for (int i = np - 1; i >= 0; i--) { s += t[i]; t[i] = s; }
        if (!single) { s = 0.f; for (int i = 0; i < np; i++) { s += t[i]; t[i] = s; } }
        if (box) for (int i = 0; i < nx; i++) x[o + i * d] = (t[i + 1] - t[i + 2 * nb]) * wt;
        else for (int i = 0; i < nx; i++) x[o + i * d] = (float)((2. * t[i + nb] - t[i] - t[i + 2 * nb]) * wt);
    }
}

void pso_smoothcf(float *x, int n1, int n2, int n3, int repeat, int adj, const int *rect, const int *diff, const int *box)
{
    int nn[3] = {n1, n2, n3};
    int nmax = n1 > n2 ? n1 : n2; if (n3 > nmax) nmax = n3;
    int rmax = rect[0] > rect[1] ? rect[0] : rect[1]; if (rect[2] > rmax) rmax = rect[2];
    float *t = falloc((size_t)nmax + 2 * (size_t)rmax + 2);
    for (int a = 0; a < 3; a++) {
        if (rect[a] <= 1) continue;
        long nl = (long)n1 * n2 * n3 / nn[a];
        for (long l = 0; l < nl; l++) {
            long o, d;
            if (a == 0) { o = l * n1; d = 1; }
            else if (a == 1) { o = (l % n1) + (long)n1 * n2 * (l / n1); d = n1; }
            else { o = l; d = (long)n1 * n2; }
            for (int q = 0; q < repeat; q++) tri_line_any(x, o, d, nn[a], rect[a], t, adj, box[a], diff[a]);
        }
    }
    free(t);
}

void pso_smooth3_fwd(float *x, int n1, int n2, int n3, int r1, int r2, int r3, int repeat)
{
    const int rect[3] = {r1, r2, r3}, z[3] = {0, 0, 0};
    pso_smoothcf(x, n1, n2, n3, repeat, 1, rect, z, z);
}

/* smoothcf (dip_cfuns.c:2006-2123) with adj = 0, no diff / box: every line of an axis is smoothed `repeat` times
 * in a row (ps_smooth2), axes in turn */
void pso_smooth3_rep(float *x, int n1, int n2, int n3, int r1, int r2, int r3, int repeat)
{
    int nmax = n1 > n2 ? n1 : n2; if (n3 > nmax) nmax = n3;
    int rmax = r1 > r2 ? r1 : r2; if (r3 > rmax) rmax = r3;
    float *t = falloc((size_t)nmax + 2 * (size_t)rmax + 2);
    if (r1 > 1)
        for (long l = 0; l < (long)n2 * n3; l++) for (int q = 0; q < repeat; q++) tri_line(x, l * n1, 1, n1, r1, t);
    if (r2 > 1)
        for (int i3 = 0; i3 < n3; i3++)
            for (int i1 = 0; i1 < n1; i1++) for (int q = 0; q < repeat; q++) tri_line(x, i1 + (long)n1 * n2 * i3, n1, n2, r2, t);
    if (r3 > 1)
        for (long l = 0; l < (long)n1 * n2; l++) for (int q = 0; q < repeat; q++) tri_line(x, l, (long)n1 * n2, n3, r3, t);
    free(t);
}

/* ------------------------------------------------------------------ shaping CG + divne */

/* Sensitivity probe (not reference behaviour): mode 1 sums the same exact double products in
 * blocks of 4096 and then the block sums — a different ASSOCIATION of the same double additions,
 * like any parallel reduction.  Used only to measure how much the reference's OWN result moves
 * when the last bits of its dot products move (tests/test_oracle.py). */
static int g_dot_mode = 0;
void pso_set_dot_mode(int mode) { g_dot_mode = mode; }

static double ddot(size_t n, const float *a, const float *b)   /* ps_cblas_dsdot :122-137 */
{
    double s = 0.;
    if (g_dot_mode == 0) {
        for (size_t i = 0; i < n; i++) s += (double)a[i] * b[i];
        return s;
    }
    for (size_t i0 = 0; i0 < n; i0 += 4096) {
        double t = 0.;
        size_t i1 = i0 + 4096 < n ? i0 + 4096 : n;
        for (size_t i = i0; i < i1; i++) t += (double)a[i] * b[i];
        s += t;
    }
    return s;
}

typedef struct { int n1, n2, n3, r1, r2, r3; size_t n; float *tmp; } shaper;

/* y (+)= S x with S = pso_smooth3 (ps_trianglen_lop :672-712; adjoint == forward) */
static void shape_apply(shaper *sh, int add, const float *x, float *y)
{
    memcpy(sh->tmp, x, sh->n * sizeof(float));
    pso_smooth3(sh->tmp, sh->n1, sh->n2, sh->n3, sh->r1, sh->r2, sh->r3);
    if (add) for (size_t i = 0; i < sh->n; i++) y[i] += sh->tmp[i];
    else     for (size_t i = 0; i < sh->n; i++) y[i] = 0.f + sh->tmp[i];
}

/* ps_conjgrad :257-383 with prec=NULL, hasp0=false, L = diag(w) (ps_weight_lop :739-758),
 * eps := eps1*eps1 and tol as set by ps_conjgrad_init (:218-240).  Returns iterations run. */
static int shaping_cg(shaper *sh, const float *w, float *p, float *x, const float *dat,
                      int niter, float eps, float tol)
{
    size_t n = sh->n;
    float *r = falloc(n), *sp = falloc(n), *sx = falloc(n), *sr = falloc(n);
    float *gp = falloc(n), *gx = falloc(n), *gr = falloc(n);
    double gn, gnp = 0., alpha, beta, g0 = 0., dg;
    int iter = 0;
    for (size_t i = 0; i < n; i++) { r[i] = -dat[i]; p[i] = 0.f; x[i] = 0.f; }
    if (ddot(n, r, r) == 0.) goto done;
    for (iter = 0; iter < niter; iter++) {
        for (size_t i = 0; i < n; i++) { gp[i] = eps * p[i]; gx[i] = -eps * x[i]; }
        for (size_t i = 0; i < n; i++) gx[i] += r[i] * w[i];          /* L' r, add */
        shape_apply(sh, 1, gx, gp);                                   /* gp += S gx */
        shape_apply(sh, 0, gp, gx);                                   /* gx  = S gp */
        for (size_t i = 0; i < n; i++) gr[i] = 0.f + gx[i] * w[i];    /* gr = L gx */
        gn = ddot(n, gp, gp);
        if (iter == 0) {
            g0 = gn;
            memcpy(sp, gp, n * sizeof(float));
            memcpy(sx, gx, n * sizeof(float));
            memcpy(sr, gr, n * sizeof(float));
        } else {
            alpha = gn / gnp;
            dg = gn / g0;
            if (alpha < tol || dg < tol) break;
            float a = (float)alpha;                                   /* saxpy takes float a */
            for (size_t i = 0; i < n; i++) {                          /* g += a*s; swap(s,g) */
                float t;
                t = gp[i] + a * sp[i]; gp[i] = sp[i]; sp[i] = t;
                t = gx[i] + a * sx[i]; gx[i] = sx[i]; sx[i] = t;
                t = gr[i] + a * sr[i]; gr[i] = sr[i]; sr[i] = t;
            }
        }
        beta = ddot(n, sr, sr) + eps * (ddot(n, sp, sp) - ddot(n, sx, sx));
        alpha = -gn / beta;
        float a = (float)alpha;
        for (size_t i = 0; i < n; i++) { p[i] += a * sp[i]; x[i] += a * sx[i]; r[i] += a * sr[i]; }
        gnp = gn;
    }
done:
    free(r); free(sp); free(sx); free(sr); free(gp); free(gx); free(gr);
    return iter;
}

/* ps_divne :796-827.  CG constants fixed by ps_divn_init -> ps_conjgrad_init(...,1.,1.e-6,...)
 * (:776-777): eps_cg = 1*1, tol = 1e-6. */
int pso_divne(float *num, float *den, float *rat, int n1, int n2, int n3,
              int r1, int r2, int r3, int liter, float eps)
{
    shaper sh = { n1, n2, n3, r1, r2, r3, (size_t)n1 * n2 * n3, NULL };
    size_t n = sh.n;
    double norm;
    if (eps > 0.0f)
        for (size_t i = 0; i < n; i++) {
            norm = 1.0 / hypot(den[i], eps);
            num[i] *= norm;
            den[i] *= norm;
        }
    norm = ddot(n, den, den);
    if (norm == 0.0) { memset(rat, 0, n * sizeof(float)); return 0; }
    norm = sqrt(n / norm);
    for (size_t i = 0; i < n; i++) { num[i] *= norm; den[i] *= norm; }
    sh.tmp = falloc(n);
    float *p = falloc(n);
    int it = shaping_cg(&sh, den, p, rat, num, liter, 1.f * 1.f, 1.e-6f);
    free(p); free(sh.tmp);
    return it;
}

/* ------------------------------------------------------------------ Gauss-Newton dip */

/* dip3 :1619-1691 for one direction.  eps handed to divne is 1.0: the file-scope `eps`
 * shared by the CG and dip3 sections is overwritten by ps_conjgrad_init.
 * pmin/pmax are -/+FLT_MAX in dipc (:1779-1782) so the clip never acts on finite values. */
/* executed-iteration counters of the last pso_dip call (test aid: the data-dependent branches -- CG early exit,
 * line-search halvings -- must go the same way on the GPU): {CG iterations, line-search evaluations, GN iterations} */
static long long g_counts[3];
void pso_get_counts(long long *out3) { out3[0] = g_counts[0]; out3[1] = g_counts[1]; out3[2] = g_counts[2]; }

static void gauss_newton(const float *u, float *p, const unsigned char *mask, int xline,
                         int n1, int n2, int n3, int niter, int liter, int nw,
                         int r1, int r2, int r3)
{
    size_t n = (size_t)n1 * n2 * n3;
    float *u1 = falloc(n), *u2 = falloc(n), *dp = falloc(n), *p0 = falloc(n);
    const float pmin = -3.402823466e+38F, pmax = 3.402823466e+38F;
    pso_allpass(u, p, n1, n2, n3, nw, xline, 0, u2);
    for (int iter = 0; iter < niter; iter++) {
        pso_allpass(u, p, n1, n2, n3, nw, xline, 1, u1);
        float usum = 0.0f;
        for (size_t i = 0; i < n; i++) { p0[i] = p[i]; usum += u2[i] * u2[i]; }
        if (mask)
            for (size_t i = 0; i < n; i++) if (mask[i]) { u1[i] = 0.f; u2[i] = 0.f; }
        g_counts[0] += pso_divne(u2, u1, dp, n1, n2, n3, r1, r2, r3, liter, 1.0f);
        g_counts[2]++;
        float lam = 1.f;
        for (int k = 0; k < 8; k++) {
            g_counts[1]++;
            for (size_t i = 0; i < n; i++) {
                float pi = p0[i] + lam * dp[i];
                if (pi < pmin) pi = pmin;
                if (pi > pmax) pi = pmax;
                p[i] = pi;
            }
            pso_allpass(u, p, n1, n2, n3, nw, xline, 0, u2);
            float usum2 = 0.f;
            for (size_t i = 0; i < n; i++) usum2 += u2[i] * u2[i];
            if (usum2 < usum) break;
            lam *= 0.5f;
        }
    }
    free(u1); free(u2); free(dp); free(p0);
}

/* dipc :1694-1989: inline dip fully, then xline dip (only when n3 > 1) */
int pso_dip(const float *din, const float *mask, int n1, int n2, int n3, int niter, int liter,
            int order, int r1, int r2, int r3, float *dip_out)
{
    size_t n = (size_t)n1 * n2 * n3;
    unsigned char *m_in = NULL, *m_x = NULL;
    if (mask) {
        m_in = (unsigned char *)malloc(n); m_x = (unsigned char *)malloc(n);
        footprint_mask(mask, n1, n2, n3, order, m_in, m_x);
    }
    memset(dip_out, 0, (n3 == 1 ? n : 2 * n) * sizeof(float));
    g_counts[0] = g_counts[1] = g_counts[2] = 0;
    gauss_newton(din, dip_out, m_in, 0, n1, n2, n3, niter, liter, order, r1, r2, r3);
    if (n3 != 1)
        gauss_newton(din, dip_out + n, m_x, 1, n1, n2, n3, niter, liter, order, r1, r2, r3);
    free(m_in); free(m_x);
    return 0;
}

/* ------------------------------------------------------------------ plane-wave prediction */

typedef struct {
    int n1, nw, nb;
    float eps;
    float *diag, *offd[2 * MAXTAP], *d, *o[2 * MAXTAP];
    float *a1[MAXTAP], *a2[MAXTAP], *t1, *t2, *r2;
} predictor;

static predictor *predictor_new(int n1, int nw, float eps)
{
    predictor *P = (predictor *)calloc(1, sizeof(predictor));
    P->n1 = n1; P->nw = nw; P->nb = 2 * nw; P->eps = eps;
    P->diag = falloc(n1); P->d = falloc(n1); P->t1 = falloc(n1); P->t2 = falloc(n1); P->r2 = falloc(n1);
    for (int m = 0; m < P->nb; m++) { P->offd[m] = falloc(n1); P->o[m] = falloc(n1); }
    for (int j = 0; j <= 2 * nw; j++) { P->a1[j] = falloc(n1); P->a2[j] = falloc(n1); }
    return P;
}

static void predictor_free(predictor *P)
{
    free(P->diag); free(P->d); free(P->t1); free(P->t2); free(P->r2);
    for (int m = 0; m < P->nb; m++) { free(P->offd[m]); free(P->o[m]); }
    for (int j = 0; j <= 2 * P->nw; j++) { free(P->a1[j]); free(P->a2[j]); }
    free(P);
}

/* regularization() sof3d_cfuns.c:548-565: eps*D2'D2-like pentadiagonal, eps2 == eps */
static void reg_fill(predictor *P)
{
    int n1 = P->n1;
    float eps = P->eps, eps2 = P->eps;
    for (int i = 0; i < n1; i++) {
        P->diag[i] = 6. * eps;
        P->offd[0][i] = -4. * eps;
        P->offd[1][i] = eps;
        for (int m = 2; m < P->nb; m++) P->offd[m][i] = 0.0;
    }
    P->diag[0] = P->diag[n1 - 1] = eps2 + eps;
    P->diag[1] = P->diag[n1 - 2] = eps2 + 5. * eps;
    P->offd[0][0] = P->offd[0][n1 - 2] = -2. * eps;
}

/* pwd_define :401-445: taps per sample (reversed when forw), accumulate W'W bands */
static void wtw_add(predictor *P, int forw, const float *sg, float **a)
{
    int n = P->n1, nw = P->nw, na = 2 * nw + 1;
    float b[MAXTAP];
    for (int i = 0; i < n; i++) {
        pso_passfilter(nw, sg[i], b);
        for (int j = 0; j < na; j++) a[j][i] = forw ? b[na - 1 - j] : b[j];
    }
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < na; j++) {
            int k = i + j - nw;
            if (k >= nw && k < n - nw) { float aj = a[j][k]; P->diag[i] += aj * aj; }
        }
        for (int m = 0; m < 2 * nw; m++)
            for (int j = m + 1; j < na; j++) {
                int k = i + j - nw;
                if (k >= nw && k < n - nw) { float aj = a[j][k], am = a[j - m - 1][k]; P->offd[m][i] += am * aj; }
            }
    }
}

/* pwd_set(adj=false) :447-498: out = W'(W inp), rows of W limited to [nw, n-nw) */
static void wtw_apply(predictor *P, float **a, const float *inp, float *out, float *tmp)
{
    int n = P->n1, nw = P->nw, na = 2 * nw + 1;
    for (int i = 0; i < n; i++) tmp[i] = 0.f;
    for (int i = nw; i < n - nw; i++)
        for (int j = 0; j < na; j++) tmp[i] += a[j][i] * inp[i + j - nw];
    for (int i = 0; i < n; i++) {
        out[i] = 0.f;
        for (int j = 0; j < na; j++) {
            int k = i + j - nw;
            if (k >= nw && k < n - nw) out[i] += a[j][k] * tmp[k];
        }
    }
}

/* sf_banded_define :159-185: LDL' of the symmetric banded matrix */
static void ldl_factor(predictor *P)
{
    int n = P->n1, band = P->nb;
    for (int k = 0; k < n; k++) {
        float t = P->diag[k];
        int m1 = k < band ? k : band;
        for (int m = 0; m < m1; m++) t -= (P->o[m][k - m - 1]) * (P->o[m][k - m - 1]) * (P->d[k - m - 1]);
        P->d[k] = t;
        int q1 = (n - k - 1) < band ? (n - k - 1) : band;
        for (int q = 0; q < q1; q++) {
            t = P->offd[q][k];
            m1 = k < (band - q - 1) ? k : (band - q - 1);
            for (int m = 0; m < m1; m++)
                t -= (P->o[m][k - m - 1]) * (P->o[q + m + 1][k - m - 1]) * (P->d[k - m - 1]);
            P->o[q][k] = t / P->d[k];
        }
    }
}

/* sf_banded_solve :245-264 */
static void ldl_solve(predictor *P, float *b)
{
    int n = P->n1, band = P->nb;
    for (int k = 1; k < n; k++) {
        float t = b[k];
        int m1 = k < band ? k : band;
        for (int m = 0; m < m1; m++) t -= (P->o[m][k - m - 1]) * b[k - m - 1];
        b[k] = t;
    }
    for (int k = n - 1; k >= 0; k--) {
        float t = b[k] / P->d[k];
        int m1 = (n - k - 1) < band ? (n - k - 1) : band;
        for (int m = 0; m < m1; m++) t -= P->o[m][k] * b[k + m + 1];
        b[k] = t;
    }
}

/* predict1_step :596-622 (also predict_step(adj=false) :567-594 when out aliases in1) */
static void predict_one(predictor *P, int forw, const float *in1, const float *sg, float *out)
{
    int n1 = P->n1;
    float eps2 = P->eps;
    reg_fill(P);
    wtw_add(P, forw, sg, P->a1);
    ldl_factor(P);
    float t0 = in1[0], t1 = in1[1], t2 = in1[n1 - 2], t3 = in1[n1 - 1];
    wtw_apply(P, P->a1, in1, P->r2, P->t1);
    memcpy(out, P->r2, n1 * sizeof(float));
    out[0] += eps2 * t0; out[1] += eps2 * t1; out[n1 - 2] += eps2 * t2; out[n1 - 1] += eps2 * t3;
    ldl_solve(P, out);
}

/* predict2_step :624-661 */
static void predict_two(predictor *P, int forw1, int forw2, const float *in1, const float *in2,
                        const float *sg1, const float *sg2, float *out)
{
    int n1 = P->n1;
    float eps2 = P->eps;
    reg_fill(P);
    wtw_add(P, forw1, sg1, P->a1);
    wtw_add(P, forw2, sg2, P->a2);
    ldl_factor(P);
    float t0 = 0.5 * (in1[0] + in2[0]), t1 = 0.5 * (in1[1] + in2[1]);
    float t2 = 0.5 * (in1[n1 - 2] + in2[n1 - 2]), t3 = 0.5 * (in1[n1 - 1] + in2[n1 - 1]);
    wtw_apply(P, P->a1, in1, P->t1, P->r2);
    wtw_apply(P, P->a2, in2, P->t2, P->r2);
    for (int i = 0; i < n1; i++) out[i] = P->t1[i] + P->t2[i];
    out[0] += eps2 * t0; out[1] += eps2 * t1; out[n1 - 2] += eps2 * t2; out[n1 - 1] += eps2 * t3;
    ldl_solve(P, out);
}

void pso_predict(int n1, int nw, float eps, int two, int forw1, int forw2, const float *trace1,
                 const float *trace2, const float *sig1, const float *sig2, float *out)
{
    predictor *P = predictor_new(n1, nw, eps);
    if (two) predict_two(P, forw1, forw2, trace1, trace2, sig1, sig2, out);
    else     predict_one(P, forw1, trace1, sig1, out);
    predictor_free(P);
}

/* ------------------------------------------------------------------ 3-D spray */

static const float *g_cost;
static int by_cost(const void *a, const void *b)          /* fermat() :984-996 */
{
    float ta = g_cost[*(const int *)a], tb = g_cost[*(const int *)b];
    if (ta > tb) return 1;
    if (ta == tb) return 0;
    return -1;
}

/* get_update :1022-1054 for slot j of an m1 x m2 cost table: which neighbours feed it */
static int slot_parents(const float *cost, int m1, int m2, int j, int *up1, int *up2)
{
    int i1 = j % m1, i2 = j / m1, upd = 0;
    float t1 = cost[j];
    g_cost = cost;
    *up1 = *up2 = 0;
    if (m1 > 1) {
        int a = j - 1, b = j + 1;
        *up1 = (i1 && (i1 == m1 - 1 || 1 != by_cost(&a, &b)));
        if (t1 > cost[*up1 ? a : b]) upd |= 1;
    }
    if (m2 > 1) {
        int a = j - m1, b = j + m1;
        *up2 = (i2 && (i2 == m2 - 1 || 1 != by_cost(&a, &b)));
        if (t1 > cost[*up2 ? a : b]) upd |= 2;
    }
    return upd;
}

/* spray loop shared by csomean3d :1455-1521 and csomf3d :1639-1705.
 * u is [n2*n3][np][n1], zero-initialised; returns it (caller frees). */
static float *spray3(const float *din, const float *dipi, const float *dipx, int n1, int n2, int n3,
                     int ns2, int ns3, int order)
{
    int np2 = 2 * ns2 + 1, np3 = 2 * ns3 + 1, np = np2 * np3, n23 = n2 * n3;
    float eps = 0.01;                                   /* caller's eps overridden (:1402,:1586) */
    float *cost = falloc(np);
    int *visit = (int *)malloc(np * sizeof(int));
    for (int k3 = 0; k3 < np3; k3++)
        for (int k2 = 0; k2 < np2; k2++) cost[k3 * np2 + k2] = hypotf(k2 - ns2, k3 - ns3);
    for (int i = 0; i < np; i++) visit[i] = i;
    g_cost = cost;
    qsort(visit, np, sizeof(int), by_cost);             /* update_init :999-1014 */
    predictor *P = predictor_new(n1, order, eps * eps);
    float *u = falloc((size_t)n23 * np * n1);
#define U(loc, slot) (u + ((size_t)(loc) * np + (slot)) * n1)
    for (int i = 0; i < n23; i++) {
        memcpy(U(i, ns3 * np2 + ns2), din + (size_t)i * n1, n1 * sizeof(float));
        int i2 = i % n2, i3 = i / n2;
        for (int ip = 0; ip < np; ip++) {
            int jp = visit[ip], up2, up3;
            int upd = slot_parents(cost, np2, np3, jp, &up2, &up3);
            int j2 = i2 + jp % np2 - ns2, j3 = i3 + jp / np2 - ns3;
            if (j2 < 0 || j2 >= n2 || j3 < 0 || j3 >= n3) continue;
            int j = j2 + j3 * n2, l2 = 0, l3 = 0, k2 = 0, k3 = 0;
            const float *q2 = NULL, *q3 = NULL;
            if (upd & 1) {
                if (up2) { if (j2 == 0) continue;      l2 = j - 1; q2 = dipi + (size_t)l2 * n1; k2 = jp - 1; }
                else     { if (j2 == n2 - 1) continue; l2 = j + 1; q2 = dipi + (size_t)j * n1;  k2 = jp + 1; }
            }
            if (upd & 2) {
                if (up3) { if (j3 == 0) continue;      l3 = j - n2; q3 = dipx + (size_t)l3 * n1; k3 = jp - np2; }
                else     { if (j3 == n3 - 1) continue; l3 = j + n2; q3 = dipx + (size_t)j * n1;  k3 = jp + np2; }
            }
            if (upd == 1)      predict_one(P, up2, U(l2, k2), q2, U(j, jp));
            else if (upd == 2) predict_one(P, up3, U(l3, k3), q3, U(j, jp));
            else if (upd == 3) predict_two(P, up2, up3, U(l2, k2), U(l3, k3), q2, q3, U(j, jp));
        }
    }
#undef U
    predictor_free(P); free(cost); free(visit);
    return u;
}

/* csomean3d :1523-1537: mean over ALL np slots, edge zeros included */
int pso_somean3d(const float *din, const float *dipi, const float *dipx, int n1, int n2, int n3,
                 int ns2, int ns3, int order, float *out)
{
    int np = (2 * ns2 + 1) * (2 * ns3 + 1), n23 = n2 * n3;
    float *u = spray3(din, dipi, dipx, n1, n2, n3, ns2, ns3, order);
    for (int i = 0; i < n23; i++)
        for (int k = 0; k < n1; k++) {
            float sum = 0;
            for (int j = 0; j < np; j++) sum = sum + u[((size_t)i * np + j) * n1 + k];
            sum = sum / np;
            out[(size_t)i * n1 + k] = sum;
        }
    free(u);
    return 0;
}

/* sf_quantile :1058-1084 (Hoare's FIND) */
static float kth(int q, int n, float *a)
{
    float *low = a, *hi = a + n - 1, *k = a + q;
    while (low < hi) {
        float ak = *k, *i = low, *j = hi;
        do {
            while (*i < ak) i++;
            while (*j > ak) j--;
            if (i <= j) { float b = *i; *i++ = *j; *j-- = b; }
        } while (i <= j);
        if (j < k) low = i;
        if (k < i) hi = j;
    }
    return *k;
}

/* mf(axis=2, ifbound=1) :1086-1108,:1138-1252 on a [np][n1] panel, keeping only the row
 * `keep`: running median of length nfw along the slot axis with edge replication */
static void slot_median_row(const float *panel, int n1, int np, int nfw, int keep, float *row)
{
    int m = (nfw - 1) / 2;
    float win[256];
    for (int k = 0; k < n1; k++) {
        for (int q = 0; q < nfw; q++) {
            int s = keep - m + q;
            if (s < 0) s = 0;
            if (s > np - 1) s = np - 1;
            win[q] = panel[(size_t)s * n1 + k];
        }
        row[k] = kth(m, nfw, win);
    }
}

/* svmf(axis=2, ifbound=1) :1110-1136,:1254-1352 on a [np][n1] panel, keeping only the row `keep` (space-varying
 * median filter, `option=2`): a first median of length nfw+2 over EVERY slot row (window rows i-m2 .. i+m2+2 of the
 * edge-replicated panel), the panel average of its magnitudes (sequential float sum, row-major), a window length per
 * sample from |x| against avg/2, avg, 2 avg (nfw+2, nfw, nfw-2, nfw-4), and the median of that length around the sample.
 * DEFINED-BEHAVIOUR VARIANT: for the last slot row the reference's first pass reads one row past its `extendt` buffer
 * (:1283: extended row m+i+k-m2 reaches np+2m for i = np-1, k = nfilter-1), i.e. heap contents enter the panel
 * average; here that row is the edge replica like the rows before it.  Everything else follows the reference. */
static int svmf_row(const float *panel, int n1, int np, int nfw, int keep, float *row)
{
    const int nfilter = nfw + 2, m = (nfilter - 1) / 2, m2 = (nfw - 1) / 2;
    float win[260];
    if (nfw - 4 < 1 || nfilter > 258) return -1;          /* the reference calls sf_quantile(-1, -1, ..) below that */
    float sum = 0.f;
    for (int i = 0; i < np; i++)
        for (int j = 0; j < n1; j++) {
            for (int k = 0; k < nfilter; k++) {
                int s = (m + i + k - m2) - m;                  /* extended row -> panel row, replicated outside */
                if (s < 0) s = 0;
                if (s > np - 1) s = np - 1;
                win[k] = panel[(size_t)s * n1 + j];
            }
            sum = sum + fabsf(kth(m, nfilter, win));
        }
    const float avg = sum / (n1 * np);
    for (int j = 0; j < n1; j++) {
        const float x = fabsf(panel[(size_t)keep * n1 + j]);
        int wl;
        if (x < avg) wl = (x < avg / 2) ? nfw + 2 : nfw + 0;
        else         wl = (x > avg * 2) ? nfw - 4 : nfw - 2;
        const int h = (wl - 1) / 2;
        for (int k = 0; k < wl; k++) {
            int s = (keep + m + k - h) - m;
            if (s < 0) s = 0;
            if (s > np - 1) s = np - 1;
            win[k] = panel[(size_t)s * n1 + j];
        }
        row[j] = kth(h, wl, win);
    }
    return 0;
}

/* csomf3d :1707-1736 (option 1; option 2 in the defined-behaviour variant of svmf_row) */
int pso_somf3d(const float *din, const float *dipi, const float *dipx, int n1, int n2, int n3,
               int ns2, int ns3, int nmf, int option, int order, float *out)
{
    int np = (2 * ns2 + 1) * (2 * ns3 + 1), n23 = n2 * n3;
    if ((option != 1 && option != 2) || nmf > 255) return -1;
    if (option == 2 && nmf - 4 < 1) return -1;
    float *u = spray3(din, dipi, dipx, n1, n2, n3, ns2, ns3, order);
    for (int i = 0; i < n23; i++) {
        if (option == 1) slot_median_row(u + (size_t)i * np * n1, n1, np, nmf, (np - 1) / 2, out + (size_t)i * n1);
        else             svmf_row(u + (size_t)i * np * n1, n1, np, nmf, (np - 1) / 2, out + (size_t)i * n1);
    }
    free(u);
    return 0;
}

/* ------------------------------------------------------------------ 2-D spray */

/* pwspray_lop(adj=false) sof_cfuns.c:948-1031 on one [n2][n1] panel: u is [n2][2ns+1][n1] */
static void spray2(predictor *P, const float *d, const float *dip, int n1, int n2, int ns, float *u)
{
    int ns2 = 2 * ns + 1;
    float *tr = falloc(n1);
    memset(u, 0, (size_t)n1 * n2 * ns2 * sizeof(float));
    for (int i = 0; i < n2; i++) {
        for (int k = 0; k < n1; k++) { tr[k] = d[(size_t)i * n1 + k]; u[((size_t)i * ns2 + ns) * n1 + k] += tr[k]; }
        for (int is = 0; is < ns; is++) {
            int ip = i - is - 1;
            if (ip < 0) break;
            predict_one(P, 0, tr, dip + (size_t)ip * n1, tr);
            float *dst = u + ((size_t)ip * ns2 + ns - is - 1) * n1;
            for (int k = 0; k < n1; k++) dst[k] += tr[k];
        }
        memcpy(tr, d + (size_t)i * n1, n1 * sizeof(float));
        for (int is = 0; is < ns; is++) {
            int ip = i + is + 1;
            if (ip >= n2) break;
            predict_one(P, 1, tr, dip + (size_t)(ip - 1) * n1, tr);
            float *dst = u + ((size_t)ip * ns2 + ns + is + 1) * n1;
            for (int k = 0; k < n1; k++) dst[k] += tr[k];
        }
    }
    free(tr);
}

/* pwsmooth_lop(adj=false) :1076-1110 given the normalisation w1 */
static void smooth2_apply(const float *u, const float *w1, int n1, int n2, int ns, float *out)
{
    int ns2 = 2 * ns + 1;
    for (int i2 = 0; i2 < n2; i2++)
        for (int k = 0; k < n1; k++) {
            float ws = w1[(size_t)i2 * n1 + k], acc = 0.f;
            for (int is = 0; is < ns2; is++) {
                float w = (float)(ns + 1 - abs(is - ns));
                acc += u[((size_t)i2 * ns2 + is) * n1 + k] * w * ws;
            }
            out[(size_t)i2 * n1 + k] = acc;
        }
}

/* csomean2d :1433-1532 with adj=0: per slice pwsmooth_set (:1113-1132) then pwsmooth_lop */
int pso_somean2d(const float *din, const float *dip, int n1, int n2, int n3, int ns, int order,
                 float eps, float *out)
{
    size_t n12 = (size_t)n1 * n2;
    int ns2 = 2 * ns + 1;
    predictor *P = predictor_new(n1, order, eps * eps);
    float *u = falloc(n12 * ns2), *w1 = falloc(n12), *t = falloc(n12);
    for (int i3 = 0; i3 < n3; i3++) {
        const float *sl = dip + i3 * n12;
        for (size_t i = 0; i < n12; i++) w1[i] = 1.0f;
        spray2(P, w1, sl, n1, n2, ns, u);
        smooth2_apply(u, w1, n1, n2, ns, t);
        for (size_t i = 0; i < n12; i++) w1[i] = (0.0f != t[i]) ? (float)(1.0 / t[i]) : 0.0f;
        spray2(P, din + i3 * n12, sl, n1, n2, ns, u);
        smooth2_apply(u, w1, n1, n2, ns, out + i3 * n12);
    }
    free(u); free(w1); free(t); predictor_free(P);
    return 0;
}

/* csomf2d :1534-1672 (option 1): spray, then median over the slot axis, centre row kept */
int pso_somf2d(const float *din, const float *dip, int n1, int n2, int n3, int ns, int nmf,
               int option, int order, float eps, float *out)
{
    size_t n12 = (size_t)n1 * n2;
    int np = 2 * ns + 1;
    if ((option != 1 && option != 2) || nmf > 255) return -1;
    if (option == 2 && nmf - 4 < 1) return -1;
    predictor *P = predictor_new(n1, order, eps * eps);
    float *u = falloc(n12 * np);
    for (int i3 = 0; i3 < n3; i3++) {
        spray2(P, din + i3 * n12, dip + i3 * n12, n1, n2, ns, u);
        for (int i = 0; i < n2; i++) {
            if (option == 1) slot_median_row(u + (size_t)i * np * n1, n1, np, nmf, (np - 1) / 2, out + i3 * n12 + (size_t)i * n1);
            else             svmf_row(u + (size_t)i * np * n1, n1, np, nmf, (np - 1) / 2, out + i3 * n12 + (size_t)i * n1);
        }
    }
    free(u); predictor_free(P);
    return 0;
}

/* cpaint2d / cpaint3d paint_cfuns.c:1861-2024 (the two entries are the same code): plane-wave painting -- the seed
 * trace sits at trace i0 and is predicted outwards trace by trace, leftwards with the target-location slope
 * (predict_step(false, false, trace, pp[i2])), rightwards with the parent-location slope and reversed taps
 * (predict_step(false, true, trace, pp[i2-1])). */
int pso_paint2d(const float *dip, const float *seed, int n1, int n2, int order, int i0, float eps, float *out)
{
    if (i0 < 0 || i0 >= n2) return -1;
    predictor *P = predictor_new(n1, order, eps * eps);
    float *tr = falloc(n1);
    memcpy(tr, seed, n1 * sizeof(float));
    memcpy(out + (size_t)i0 * n1, seed, n1 * sizeof(float));
    for (int i2 = i0 - 1; i2 >= 0; i2--) {
        predict_one(P, 0, tr, dip + (size_t)i2 * n1, tr);
        memcpy(out + (size_t)i2 * n1, tr, n1 * sizeof(float));
    }
    memcpy(tr, seed, n1 * sizeof(float));
    for (int i2 = i0 + 1; i2 < n2; i2++) {
        predict_one(P, 1, tr, dip + (size_t)(i2 - 1) * n1, tr);
        memcpy(out + (size_t)i2 * n1, tr, n1 * sizeof(float));
    }
    free(tr); predictor_free(P);
    return 0;
}

/* ------------------------------------------------------------------ PWD-residual interpolation */

/* allpass3_lop :625-729 (drift=false, nj=1): y[0:N] = inline PWD of x, y[N:2N] = xline PWD;
 * adjoint scatters in the same loop order. */
static void pwd3_lop(int adj, int add, int n1, int n2, int n3, int nw, int nj1, int nj2, const float *pp,
                     const float *qq, float *xx, float *yy)
{
    /* allpass3_lop soint3d_cfuns.c:625-729 (drift = false): shifts (iw - nw) * nj, rows [nw*nj, n1 - nw*nj) */
    size_t n = (size_t)n1 * n2 * n3;
    float flt[MAXTAP];
    if (!add) {
        if (adj) memset(xx, 0, n * sizeof(float));
        else     memset(yy, 0, 2 * n * sizeof(float));
    }
    for (int iz = 0; iz < n3; iz++)
        for (int iy = 0; iy < n2 - 1; iy++)
            for (int ix = nw * nj1; ix < n1 - nw * nj1; ix++) {
                long i = ix + (long)n1 * (iy + (long)n2 * iz);
                pso_passfilter(nw, pp[i], flt);
                for (int iw = 0; iw <= 2 * nw; iw++) {
                    int is = (iw - nw) * nj1;
                    if (adj) { xx[i + n1 + is] += yy[i] * flt[iw]; xx[i - is] -= yy[i] * flt[iw]; }
                    else     yy[i] += (xx[i + n1 + is] - xx[i - is]) * flt[iw];
                }
            }
    long pl = (long)n1 * n2;
    for (int iz = 0; iz < n3 - 1; iz++)
        for (int iy = 0; iy < n2; iy++)
            for (int ix = nw * nj2; ix < n1 - nw * nj2; ix++) {
                long i = ix + (long)n1 * (iy + (long)n2 * iz);
                pso_passfilter(nw, qq[i], flt);
                for (int iw = 0; iw <= 2 * nw; iw++) {
                    int is = (iw - nw) * nj2;
                    if (adj) { xx[i + pl + is] += yy[i + n] * flt[iw]; xx[i - is] -= yy[i + n] * flt[iw]; }
                    else     yy[i + n] += (xx[i + pl + is] - xx[i - is]) * flt[iw];
                }
            }
}

static float sumsq_f(size_t n, const float *x)       /* ps_cblas_snrm2 :805-818: float sum of squares */
{
    float s = 0.0;
    for (size_t i = 0; i < n; i++) s += x[i] * x[i];
    return s;
}

/* MT19937 (init_genrand / genrand_int32 / genrand_real1, soint3d_cfuns.c:2304-2370) and the Box-Muller pair
 * generator ps_randn_one_bm (:2372-2402): the second value of each pair is returned first, the first one is kept
 * for the next call. */
typedef struct { unsigned long mt[624]; int mti; int have; float kept; } noise_gen;
static void noise_seed(noise_gen *g, unsigned long s)
{
    g->mt[0] = s & 0xffffffffUL;
    for (int i = 1; i < 624; i++) g->mt[i] = (1812433253UL * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + i) & 0xffffffffUL;
    g->mti = 624; g->have = 0; g->kept = 0.f;
}
static unsigned long noise_u32(noise_gen *g)
{
    if (g->mti >= 624) {
        for (int k = 0; k < 624; k++) {
            unsigned long y = (g->mt[k] & 0x80000000UL) | (g->mt[(k + 1) % 624] & 0x7fffffffUL);
            g->mt[k] = g->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        g->mti = 0;
    }
    unsigned long y = g->mt[g->mti++];
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680UL; y ^= (y << 15) & 0xefc60000UL; y ^= (y >> 18);
    return y & 0xffffffffUL;
}
static float noise_normal(noise_gen *g)
{
    if (g->have) { g->have = 0; return g->kept; }
    double x1, x2;
    do { x1 = noise_u32(g) * (1.0 / 4294967295.0); } while (x1 == 0.0);
    x2 = noise_u32(g) * (1.0 / 4294967295.0);
    double z1 = sqrt(-2.0 * log(x1)), z2 = 2.0 * 3.14159265358979323846264338328 * x2;
    double y1 = z1 * cos(z2), y2 = z1 * sin(z2);
    g->have = 1; g->kept = (float)y1;
    return (float)y2;
}

/* csoint3d :2405-2508; var > 0 puts a*randn (a = sqrtf(var), MT19937 seeded with `seed`) on the right-hand side */
int pso_soint3d_noise(const float *din, const float *mask, const float *dipi, const float *dipx,
                      int n1, int n2, int n3, int order, int niter, int seed, float var, float *out)
{
    return pso_soint3d_full(din, mask, dipi, dipx, n1, n2, n3, order, 1, 1, niter, seed, var, out);
}

int pso_soint3d_full(const float *din, const float *mask, const float *dipi, const float *dipx,
                     int n1, int n2, int n3, int order, int nj1, int nj2, int niter, int seed, float var, float *out)
{
    size_t n = (size_t)n1 * n2 * n3, ny = 2 * n;
    float *x = out, *g = falloc(n), *rr = falloc(ny), *gg = falloc(ny), *S = falloc(n), *Ss = falloc(ny);
    unsigned char *known = (unsigned char *)malloc(n);
    for (size_t i = 0; i < n; i++) known[i] = mask ? (mask[i] != 0.f) : (din[i] != 0.f);
    /* ps_solver :1018-1040: rr = -dat (dat = 0 when var = 0); x = x0; rr += L x */
    {
        noise_gen G;
        noise_seed(&G, (unsigned long)seed);
        const float a = sqrtf(var);
        for (size_t i = 0; i < ny; i++) { float d = a * noise_normal(&G); rr[i] = -d; }
    }
    memcpy(x, din, n * sizeof(float));
    pwd3_lop(0, 1, n1, n2, n3, order, nj1, nj2, dipi, dipx, x, rr);
    float dpr0 = sumsq_f(ny, rr), dpg0 = 1.f, dpr, dpg;
    int first = 1;
    for (int iter = 0; iter < niter; iter++) {
        pwd3_lop(1, 0, n1, n2, n3, order, nj1, nj2, dipi, dipx, g, rr);
        for (size_t i = 0; i < n; i++) if (known[i]) g[i] = 0.0;
        pwd3_lop(0, 0, n1, n2, n3, order, nj1, nj2, dipi, dipx, g, gg);
        if (iter == 0) { dpg0 = sumsq_f(n, g); dpr = 1.; dpg = 1.; }
        else { dpr = sumsq_f(ny, rr) / dpr0; dpg = sumsq_f(n, g) / dpg0; }
        if (dpr < 1.e-12f || dpg < 1.e-12f) break;
        /* ps_cgstep :826-877 */
        double alfa, beta;
        if (first) {
            first = 0;
            memset(S, 0, n * sizeof(float)); memset(Ss, 0, ny * sizeof(float));
            beta = 0.0;
            alfa = ddot(ny, gg, gg);
            if (alfa <= 0.) continue;
            alfa = -ddot(ny, gg, rr) / alfa;
        } else {
            double gdg = ddot(ny, gg, gg), sds = ddot(ny, Ss, Ss), gds = ddot(ny, gg, Ss);
            if (gdg == 0. || sds == 0.) continue;
            double determ = 1.0 - (gds / gdg) * (gds / sds);
            if (determ > 1.e-12f) determ *= gdg * sds; else determ = gdg * sds * 1.e-12f;
            double gdr = -ddot(ny, gg, rr), sdr = -ddot(ny, Ss, rr);
            alfa = (sds * gdr - gds * sdr) / determ;
            beta = (-gds * gdr + gdg * sdr) / determ;
        }
        float fb = (float)beta, fa = (float)alfa;
        for (size_t i = 0; i < n; i++) S[i] *= fb;
        for (size_t i = 0; i < n; i++) S[i] += fa * g[i];
        for (size_t i = 0; i < ny; i++) Ss[i] *= fb;
        for (size_t i = 0; i < ny; i++) Ss[i] += fa * gg[i];
        for (size_t i = 0; i < n; i++) x[i] += S[i];
        for (size_t i = 0; i < ny; i++) rr[i] += Ss[i];
    }
    free(g); free(rr); free(gg); free(S); free(Ss); free(known);
    return 0;
}

int pso_soint3d(const float *din, const float *mask, const float *dipi, const float *dipx,
                int n1, int n2, int n3, int order, int niter, float *out)
{
    return pso_soint3d_noise(din, mask, dipi, dipx, n1, n2, n3, order, niter, 202223, 0.f, out);
}

/* ------------------------------------------------------------------ spray-operator interpolation */

/* pwd_set(adj=true) soint3d_cfuns.c (same as sof3d_cfuns.c:460-478): inp = W'(W ... )' applied in
 * the scatter order of the reference */
static void wtw_apply_adj(predictor *P, float **a, float *io, float *tmp)
{
    int n = P->n1, nw = P->nw, na = 2 * nw + 1;
    for (int i = 0; i < n; i++) tmp[i] = 0.f;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < na; j++) {
            int k = i + j - nw;
            if (k >= nw && k < n - nw) tmp[k] += a[j][k] * io[i];
        }
    for (int i = 0; i < n; i++) io[i] = 0.f;
    for (int i = nw; i < n - nw; i++)
        for (int j = 0; j < na; j++) io[i + j - nw] += a[j][i] * tmp[i];
}

/* predict_step(adj=true) soint3d_cfuns.c:1777-1804: solve first, then the adjoint of W'W */
static void predict_adj(predictor *P, int forw, float *trace, const float *sg)
{
    int n1 = P->n1;
    float eps2 = P->eps;
    reg_fill(P);
    wtw_add(P, forw, sg, P->a1);
    ldl_factor(P);
    ldl_solve(P, trace);
    float t0 = trace[0], t1 = trace[1], t2 = trace[n1 - 2], t3 = trace[n1 - 1];
    wtw_apply_adj(P, P->a1, trace, P->t1);
    trace[0] += eps2 * t0; trace[1] += eps2 * t1; trace[n1 - 2] += eps2 * t2; trace[n1 - 1] += eps2 * t3;
}

void pso_predict_adj(int n1, int nw, float eps, int forw, float *trace, const float *sig)
{
    predictor *P = predictor_new(n1, nw, eps);
    predict_adj(P, forw, trace, sig);
    predictor_free(P);
}

/* pwspray_lop(adj=true, add=true) :1963-2003 on one [n2][n1] panel: u1 += spray'(u) */
static void spray2_adj(predictor *P, const float *u, const float *dip, int n1, int n2, int ns, float *u1)
{
    int ns2 = 2 * ns + 1;
    float *tr = falloc(n1);
    for (int i = 0; i < n2; i++) {
        for (int k = 0; k < n1; k++) tr[k] = 0.0f;
        for (int is = ns - 1; is >= 0; is--) {
            int ip = i + is + 1;
            if (ip >= n2) continue;
            const float *src = u + ((size_t)ip * ns2 + ns + is + 1) * n1;
            for (int k = 0; k < n1; k++) tr[k] += src[k];
            predict_adj(P, 1, tr, dip + (size_t)(ip - 1) * n1);
        }
        for (int k = 0; k < n1; k++) { u1[(size_t)i * n1 + k] += tr[k]; tr[k] = 0.0f; }
        for (int is = ns - 1; is >= 0; is--) {
            int ip = i - is - 1;
            if (ip < 0) continue;
            const float *src = u + ((size_t)ip * ns2 + ns - is - 1) * n1;
            for (int k = 0; k < n1; k++) tr[k] += src[k];
            predict_adj(P, 0, tr, dip + (size_t)ip * n1);
        }
        for (int k = 0; k < n1; k++) {
            u1[(size_t)i * n1 + k] += tr[k];
            tr[k] = u[((size_t)i * ns2 + ns) * n1 + k];
            u1[(size_t)i * n1 + k] += tr[k];
        }
    }
    free(tr);
}

typedef struct { predictor *P; int n1, n2, ns; float *u, *w1, *t; } smoother2;

/* pwsmooth_set :2117-2132 (normalisation by the smooth of ones) */
static void smoother2_set(smoother2 *s, const float *dip)
{
    size_t n12 = (size_t)s->n1 * s->n2;
    for (size_t i = 0; i < n12; i++) s->w1[i] = 1.0f;
    spray2(s->P, s->w1, dip, s->n1, s->n2, s->ns, s->u);
    smooth2_apply(s->u, s->w1, s->n1, s->n2, s->ns, s->t);
    for (size_t i = 0; i < n12; i++) s->w1[i] = (0.0f != s->t[i]) ? (float)(1.0 / s->t[i]) : 0.0f;
}

/* pwsmooth_lop :2079-2110: forward writes `out` (no add); adjoint ACCUMULATES into `in` when add */
static void smoother2_fwd(smoother2 *s, const float *dip, const float *in, float *out)
{
    spray2(s->P, in, dip, s->n1, s->n2, s->ns, s->u);
    smooth2_apply(s->u, s->w1, s->n1, s->n2, s->ns, out);
}
static void smoother2_adj(smoother2 *s, const float *dip, int add, float *in, const float *out)
{
    int ns2 = 2 * s->ns + 1;
    size_t n12 = (size_t)s->n1 * s->n2;
    if (!add) memset(in, 0, n12 * sizeof(float));
    for (int i2 = 0; i2 < s->n2; i2++)
        for (int k = 0; k < s->n1; k++) {
            float ws = s->w1[(size_t)i2 * s->n1 + k];
            for (int is = 0; is < ns2; is++) {
                float w = (float)(s->ns + 1 - abs(is - s->ns));
                s->u[((size_t)i2 * ns2 + is) * s->n1 + k] = out[(size_t)i2 * s->n1 + k] * w * ws;
            }
        }
    spray2_adj(s->P, s->u, dip, s->n1, s->n2, s->ns, in);
}

/* csomean2d with adj=1 (sof_cfuns.c:1503-1508): per slice pwsmooth_set, then smooth = S' input */
int pso_somean2d_adj(const float *din, const float *dip, int n1, int n2, int n3, int ns, int order,
                     float eps, float *out)
{
    size_t n12 = (size_t)n1 * n2;
    smoother2 s = { predictor_new(n1, order, eps * eps), n1, n2, ns, falloc(n12 * (2 * ns + 1)), falloc(n12), falloc(n12) };
    for (int i3 = 0; i3 < n3; i3++) {
        smoother2_set(&s, dip + i3 * n12);
        smoother2_adj(&s, dip + i3 * n12, 0, out + i3 * n12, din + i3 * n12);
    }
    predictor_free(s.P); free(s.u); free(s.w1); free(s.t);
    return 0;
}

typedef struct {
    int n1, n2, n3, ns1, ns2, o1, o2; float eps;
    const float *idip; float *xdipT;              /* xline slopes as [n2][n3][n1] */
    float *itmp, *itmp2, *xtmp;
} smoother3;

/* pwsmooth3_lop :2231-2300 */
static void smoother3_apply(smoother3 *S, int adj, int add, float *trace, float *smooth)
{
    int n1 = S->n1, n2 = S->n2, n3 = S->n3;
    size_t n = (size_t)n1 * n2 * n3, n12 = (size_t)n1 * n2, n13 = (size_t)n1 * n3;
    if (!add) { if (adj) memset(trace, 0, n * sizeof(float)); else memset(smooth, 0, n * sizeof(float)); }
    smoother2 sa = { predictor_new(n1, S->o1, S->eps * S->eps), n1, n2, S->ns1, falloc(n12 * (2 * S->ns1 + 1)), falloc(n12), falloc(n12) };
    smoother2 sb = { predictor_new(n1, S->o2, S->eps * S->eps), n1, n3, S->ns2, falloc(n13 * (2 * S->ns2 + 1)), falloc(n13), falloc(n13) };
    if (adj) {
        for (int i3 = 0; i3 < n3; i3++) for (int i2 = 0; i2 < n2; i2++)
            memcpy(S->xtmp + ((size_t)i2 * n3 + i3) * n1, smooth + ((size_t)i3 * n2 + i2) * n1, n1 * sizeof(float));
        for (int i2 = 0; i2 < n2; i2++) {
            smoother2_set(&sb, S->xdipT + i2 * n13);
            smoother2_adj(&sb, S->xdipT + i2 * n13, 0, S->itmp2 + i2 * n13, S->xtmp + i2 * n13);
        }
        for (int i3 = 0; i3 < n3; i3++) for (int i2 = 0; i2 < n2; i2++)
            memcpy(S->itmp + ((size_t)i3 * n2 + i2) * n1, S->itmp2 + ((size_t)i2 * n3 + i3) * n1, n1 * sizeof(float));
        for (int i3 = 0; i3 < n3; i3++) {
            smoother2_set(&sa, S->idip + i3 * n12);
            smoother2_adj(&sa, S->idip + i3 * n12, 1, trace + i3 * n12, S->itmp + i3 * n12);
        }
    } else {
        for (int i3 = 0; i3 < n3; i3++) {
            smoother2_set(&sa, S->idip + i3 * n12);
            smoother2_fwd(&sa, S->idip + i3 * n12, trace + i3 * n12, S->itmp + i3 * n12);
        }
        for (int i3 = 0; i3 < n3; i3++) for (int i2 = 0; i2 < n2; i2++)
            memcpy(S->itmp2 + ((size_t)i2 * n3 + i3) * n1, S->itmp + ((size_t)i3 * n2 + i2) * n1, n1 * sizeof(float));
        for (int i2 = 0; i2 < n2; i2++) {
            smoother2_set(&sb, S->xdipT + i2 * n13);
            smoother2_fwd(&sb, S->xdipT + i2 * n13, S->itmp2 + i2 * n13, S->xtmp + i2 * n13);
        }
        for (int i3 = 0; i3 < n3; i3++) for (int i2 = 0; i2 < n2; i2++) for (int k = 0; k < n1; k++)
            smooth[k + (size_t)n1 * (i2 + (size_t)n2 * i3)] += S->xtmp[((size_t)i2 * n3 + i3) * n1 + k];
    }
    predictor_free(sa.P); free(sa.u); free(sa.w1); free(sa.t);
    predictor_free(sb.P); free(sb.u); free(sb.w1); free(sb.t);
}

/* ps_conjgrad (hasp0 = true, prec = NULL, L = known-data mask, eps = lam^2, tol = 10*FLT_EPSILON) as csint3d
 * :2590-2600 and csint2d (soint2d_cfuns.c:2508-2520) run it; S is applied through `apply(S, adj, add, trace, smooth)` */
typedef void (*shape_fn)(void *S, int adj, int add, float *trace, float *smooth);

static void sint_cg(size_t n, const float *din, const float *mask, int niter, shape_fn apply, void *S, float *out)
{
    unsigned char *known = (unsigned char *)malloc(n);
    float lam = 0.;
    for (size_t i = 0; i < n; i++) { if (mask[i] != 0.) { known[i] = 1; lam += 1.; } else known[i] = 0; }
    lam = sqrtf(lam / n);
    const float ceps = lam * lam, tol = 10 * 1.19209290e-07F;
    float *p = falloc(n), *x = out, *r = falloc(n), *sp = falloc(n), *sx = falloc(n), *sr = falloc(n);
    float *gp = falloc(n), *gx = falloc(n), *gr = falloc(n);
    memcpy(p, din, n * sizeof(float));
    for (size_t i = 0; i < n; i++) r[i] = -din[i];
    apply(S, 0, 0, p, x);                                              /* x = S p */
    for (size_t i = 0; i < n; i++) if (known[i]) r[i] += x[i];         /* r += L x */
    double gn, gnp = 0., alpha, beta, g0 = 0., dg;
    if (ddot(n, r, r) != 0.) {
        for (int iter = 0; iter < niter; iter++) {
            for (size_t i = 0; i < n; i++) { gp[i] = ceps * p[i]; gx[i] = -ceps * x[i]; }
            for (size_t i = 0; i < n; i++) if (known[i]) gx[i] += r[i];       /* L' r, add */
            apply(S, 1, 1, gp, gx);                                            /* gp += S' gx */
            apply(S, 0, 0, gp, gx);                                            /* gx  = S gp */
            for (size_t i = 0; i < n; i++) { gr[i] = 0.f; if (known[i]) gr[i] += gx[i]; }
            gn = ddot(n, gp, gp);
            if (iter == 0) {
                g0 = gn;
                memcpy(sp, gp, n * sizeof(float)); memcpy(sx, gx, n * sizeof(float)); memcpy(sr, gr, n * sizeof(float));
            } else {
                alpha = gn / gnp; dg = gn / g0;
                if (alpha < tol || dg < tol) break;
                float a = (float)alpha;
                for (size_t i = 0; i < n; i++) {
                    float t;
                    t = gp[i] + a * sp[i]; gp[i] = sp[i]; sp[i] = t;
                    t = gx[i] + a * sx[i]; gx[i] = sx[i]; sx[i] = t;
                    t = gr[i] + a * sr[i]; gr[i] = sr[i]; sr[i] = t;
                }
            }
            beta = ddot(n, sr, sr) + ceps * (ddot(n, sp, sp) - ddot(n, sx, sx));
            alpha = -gn / beta;
            float a = (float)alpha;
            for (size_t i = 0; i < n; i++) { p[i] += a * sp[i]; x[i] += a * sx[i]; r[i] += a * sr[i]; }
            gnp = gn;
        }
    }
    free(p); free(r); free(sp); free(sx); free(sr); free(gp); free(gx); free(gr); free(known);
}

static void smoother3_shape(void *S, int adj, int add, float *trace, float *smooth) { smoother3_apply((smoother3 *)S, adj, add, trace, smooth); }

/* csint3d :2510-2640 */
int pso_sint3d(const float *din, const float *dipi, const float *dipx, const float *mask,
               int n1, int n2, int n3, int niter, int ns1, int ns2, int order1, int order2, float eps,
               float *out)
{
    size_t n = (size_t)n1 * n2 * n3;
    smoother3 S = { n1, n2, n3, ns1, ns2, order1, order2, eps, dipi, falloc(n), falloc(n), falloc(n), falloc(n) };
    for (int i3 = 0; i3 < n3; i3++) for (int i2 = 0; i2 < n2; i2++)
        memcpy(S.xdipT + ((size_t)i2 * n3 + i3) * n1, dipx + ((size_t)i3 * n2 + i2) * n1, n1 * sizeof(float));
    sint_cg(n, din, mask, niter, smoother3_shape, &S, out);
    free(S.xdipT); free(S.itmp); free(S.itmp2); free(S.xtmp);
    return 0;
}

/* csint2d soint2d_cfuns.c:2421-2530 (§8f rank 3, oracle groundwork): the same shaping CG with the single 2-D plane-wave
 * smoother as S (pwsmooth_set once, pwsmooth_lop forward / adjoint) */
typedef struct { smoother2 sm; const float *dip; } shape2;
static void smoother2_shape(void *Sv, int adj, int add, float *trace, float *smooth)
{
    shape2 *S = (shape2 *)Sv;
    if (adj) smoother2_adj(&S->sm, S->dip, add, trace, smooth);
    else {
        size_t n12 = (size_t)S->sm.n1 * S->sm.n2;
        if (add) {
            float *tmp = falloc(n12);
            smoother2_fwd(&S->sm, S->dip, trace, tmp);
            for (size_t i = 0; i < n12; i++) smooth[i] += tmp[i];
            free(tmp);
        } else smoother2_fwd(&S->sm, S->dip, trace, smooth);
    }
}

int pso_sint2d(const float *din, const float *dip, const float *mask, int n1, int n2, int niter, int ns, int order,
               float eps, float *out)
{
    size_t n12 = (size_t)n1 * n2;
    shape2 S = { { predictor_new(n1, order, eps * eps), n1, n2, ns, falloc(n12 * (2 * ns + 1)), falloc(n12), falloc(n12) }, dip };
    smoother2_set(&S.sm, dip);
    sint_cg(n12, din, mask, niter, smoother2_shape, &S, out);
    predictor_free(S.sm.P); free(S.sm.u); free(S.sm.w1); free(S.sm.t);
    return 0;
}
