/* TEST INFRASTRUCTURE ONLY.  Plain-C CPU restatement of the reference's hot path
 * (aaspip/pyseistr: dip_cfuns.c, sof3d_cfuns.c, sof_cfuns.c).  Never linked, loaded
 * or called by the product path (pyseistr_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may use it.  Parity is PINNED: tests/test_oracle.py
 * checks every function below bit-for-bit against oracle/_ref (the unmodified
 * reference C compiled from /root/reference) and against tests/golden/ fixtures
 * generated from it (the reference ships no golden vectors of its own).
 *
 * All volumes are float32 in the reference's flattened Fortran order
 * i = i1 + n1*(i2 + n2*i3).
 */
#ifndef PST_ORACLE_H
#define PST_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* sensitivity probe: 0 = the reference's sequential double dot products (default), 1 = the same
 * products summed in blocks (see pst_oracle.c) */
void pso_set_dot_mode(int mode);

/* B-filter taps: dip_cfuns.c:835-910 */
void pso_passfilter(int nw, float sigma, float *taps /*[2nw+1]*/);
void pso_aderfilter(int nw, float sigma, float *taps /*[2nw+1]*/);

/* PWD stencil, left=false, drift=false, nj=1: dip_cfuns.c:1135-1200 (inline), :1399-1465 (xline) */
void pso_allpass(const float *u, const float *sigma, int n1, int n2, int n3, int nw,
                 int xline, int der, float *y);

/* N-D triangle smoothing in place, axes with rect<=1 skipped: dip_cfuns.c:458-727 (ps_smooth2) */
void pso_smooth3(float *x, int n1, int n2, int n3, int r1, int r2, int r3);
/* smoothcf dip_cfuns.c:2006-2123, adj = 0: each axis pass repeated `repeat` times per line */
void pso_smooth3_rep(float *x, int n1, int n2, int n3, int r1, int r2, int r3, int repeat);
/* smoothcf with adj = 1: ps_smooth (fold, doubint, triple) dip_cfuns.c:439-456,487-505,531-547,591-603 */
void pso_smooth3_fwd(float *x, int n1, int n2, int n3, int r1, int r2, int r3, int repeat);
/* smoothcf with every option (repeat, adj, rect[3], diff[3], box[3]) dip_cfuns.c:2006-2123 */
void pso_smoothcf(float *x, int n1, int n2, int n3, int repeat, int adj, const int *rect, const int *diff, const int *box);

/* Smooth division rat ~ num/den (num, den are overwritten): dip_cfuns.c:796-827 + :257-383.
 * Returns the number of CG iterations executed. */
int pso_divne(float *num, float *den, float *rat, int n1, int n2, int n3,
              int r1, int r2, int r3, int liter, float eps);

/* dipc: dip_cfuns.c:1694-1989.  mask may be NULL.  dip_out holds N floats when n3==1,
 * else 2N (inline then xline).  eps_dv/eps_cg/tol_cg are not parameters because the
 * reference ignores them. */
int pso_dip(const float *din, const float *mask, int n1, int n2, int n3, int niter, int liter,
            int order, int r1, int r2, int r3, float *dip_out);

/* one plane-wave prediction (sof3d_cfuns.c:596-661); two==0: predict1_step from trace1/sig1,
 * two!=0: predict2_step */
void pso_predict(int n1, int nw, float eps, int two, int forw1, int forw2, const float *trace1,
                 const float *trace2, const float *sig1, const float *sig2, float *out);

/* csomean3d / csomf3d: sof3d_cfuns.c:1355-1552, :1554-1752 (option 1 = MF only) */
int pso_somean3d(const float *din, const float *dipi, const float *dipx, int n1, int n2, int n3,
                 int ns2, int ns3, int order, float *out);
int pso_somf3d(const float *din, const float *dipi, const float *dipx, int n1, int n2, int n3,
               int ns2, int ns3, int nmf, int option, int order, float *out);

/* csomean2d / csomf2d: sof_cfuns.c:1433-1532, :1534-1672 (option 1 = MF only; adj=0) */
int pso_somean2d(const float *din, const float *dip, int n1, int n2, int n3, int ns, int order,
                 float eps, float *out);
int pso_somf2d(const float *din, const float *dip, int n1, int n2, int n3, int ns, int nmf,
               int option, int order, float eps, float *out);

/* csoint3d: soint3d_cfuns.c:2405-2508 with var=0, drift=0, nj1=nj2=1 (PWD-residual CG interpolation:
 * allpass3_lop :625-729, ps_solver :894-1174 with known mask and x0, ps_cgstep :826-877).
 * mask may be NULL (hasmask=0: known = data != 0). */
int pso_soint3d(const float *din, const float *mask, const float *dipi, const float *dipx,
                int n1, int n2, int n3, int order, int niter, float *out);

/* same with the noise right-hand side of var > 0 (MT19937 + Box-Muller, soint3d_cfuns.c:2304-2402,2431-2479) */
int pso_soint3d_noise(const float *din, const float *mask, const float *dipi, const float *dipx,
                      int n1, int n2, int n3, int order, int niter, int seed, float var, float *out);

/* same with dealiasing strides njs = (nj1, nj2) of the inline / xline stencils (allpass_init :…, shifts (iw-nw)*nj) */
int pso_soint3d_full(const float *din, const float *mask, const float *dipi, const float *dipx,
                     int n1, int n2, int n3, int order, int nj1, int nj2, int niter, int seed, float var, float *out);

/* csint3d: soint3d_cfuns.c:2510-2640 (shaping CG, L = known-data mask :1350-1372, S = pwsmooth3_lop
 * :2231-2300 = inline 2-D pwsmooth o transpose o xline 2-D pwsmooth, adjoint spray :1963-2003,
 * predict_step(adj) :1777-1804; ps_conjgrad with hasp0 = true, eps = lam^2, tol = 10*FLT_EPSILON). */
int pso_sint3d(const float *din, const float *dipi, const float *dipx, const float *mask,
               int n1, int n2, int n3, int niter, int ns1, int ns2, int order1, int order2, float eps,
               float *out);
/* csomean2d with adj = 1: sof_cfuns.c:1503-1508 (pwsmooth_lop(adj) :1064-1100, pwspray_lop(adj) :948-1031) */
int pso_somean2d_adj(const float *din, const float *dip, int n1, int n2, int n3, int ns, int order,
                     float eps, float *out);
/* csint2d soint2d_cfuns.c:2421-2530 (oracle groundwork, no GPU counterpart yet) */
int pso_sint2d(const float *din, const float *dip, const float *mask, int n1, int n2, int niter, int ns, int order,
               float eps, float *out);
/* cpaint2d / cpaint3d paint_cfuns.c:1861-2024: plane-wave painting of a seed trace */
int pso_paint2d(const float *dip, const float *seed, int n1, int n2, int order, int i0, float eps, float *out);
/* one adjoint prediction step in place (unit-test hook): predict_step(adj=true) :1777-1804 */
void pso_predict_adj(int n1, int nw, float eps, int forw, float *trace, const float *sig);

#ifdef __cplusplus
}
#endif
#endif
