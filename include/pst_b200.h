/* pst_b200 — C-ABI of the B200-native (sm_100a) implementation of pyseistr's 3-D/2-D
 * structure-oriented filtering hot path.  This is the drop-in boundary: each entry point
 * replaces one function of the reference's CPython extension modules (the FFI the
 * reference's Python `*c` wrappers bind), with plain pointers and sizes.
 *
 * Conventions
 *  - every volume is float32 in the reference's flattened Fortran order
 *    i = i1 + n1*(i2 + n2*i3)  (axis 1 = time/depth is contiguous);
 *  - `pst_*` entry points take HOST pointers (copies H2D/D2H inside the call, like the
 *    reference's copy-in/copy-out); `pst_*_dev` take DEVICE pointers on the context's GPU;
 *  - return 0 on success, <0 on error (PST_E*), message via pst_last_error();
 *  - there is NO CPU fallback: without a CUDA device every call fails with PST_ENODEV.
 *  - a context is not re-entrant (one call at a time per context; use one context per
 *    thread), unlike the reference which is not re-entrant per PROCESS (file-scope statics).
 */
#ifndef PST_B200_H
#define PST_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PST_OK        0
#define PST_EINVAL   -1   /* bad argument (the reference prints and returns NULL) */
#define PST_ENODEV   -2   /* no usable CUDA device / wrong architecture */
#define PST_ENOMEM   -3   /* device or host allocation failed */
#define PST_ECUDA    -4   /* CUDA runtime error */
#define PST_EUNSUP   -5   /* parameter combination not implemented on the GPU path */
#define PST_ECOMM    -6   /* multi-GPU communicator error */

typedef struct pst_ctx pst_ctx;

/* Statistics of the last call on a context (roofline accounting uses EXECUTED counts). */
typedef struct pst_stats {
    long long kernel_launches;   /* kernels of this library launched by the last call */
    long long cg_iterations;     /* shaping-CG iterations executed (all GN iterations, both dips) */
    long long linesearch_evals;  /* allpass evaluations inside the line search */
    long long gn_iterations;     /* Gauss-Newton iterations executed */
    long long smooth_passes;     /* single-axis triangle smoothing passes executed */
    long long predictions;       /* plane-wave predictions (trace solves) executed */
    double    device_ms;         /* CUDA-event time of the device part of the last call */
    double    h2d_bytes, d2h_bytes;
    /* per kernel class (PST_K_*), filled only while profiling is on (pst_ctx_set_profile):
     * CUDA-event time on the launching stream around every launch, and launch counts */
    double    class_ms[12];
    long long class_launches[12];
    double    class_bytes[12];   /* ALGORITHMIC bytes (compulsory reads + writes) of those launches */
    double    class_flops[12];   /* ALGORITHMIC flops of those launches (ALU-bound classes: PST_K_PREDICT) */
} pst_stats;

#define PST_K_ALLPASS   0   /* PWD stencil (+ fused line-search update, sum of squares) */
#define PST_K_TRI1      1   /* triangle smoothing, axis 1 (contiguous lines, shared-memory tile) */
#define PST_K_TRI2      2   /* triangle smoothing, axis 2 (strided lines) */
#define PST_K_TRI3      3   /* triangle smoothing, axis 3 (strided lines) */
#define PST_K_CGVEC     4   /* divne pre-scaling / CG set-up vector kernels with double reductions */
#define PST_K_PREDICT   5   /* plane-wave prediction: banded LDL' factor + solve per trace */
#define PST_K_SLOTRED   6   /* mean / median / weighted sum over the sprayed slots */
#define PST_K_OTHER     7   /* transposes, fills, final reductions */
#define PST_K_CGHEAD    8   /* CG: p,x,r += a*s of the previous iteration + gradient head gx = -eps x + w r */
#define PST_K_CGGP      9   /* CG: gp = eps p + S(gx), sum gp^2 */
#define PST_K_CGDIR    10   /* CG: gr = w gx, direction update s = g + alpha s, three dots */
#define PST_K_TRI3BWD  11   /* distributed axis 3 with PST_TRI3_SPLIT=1: the backward (carry-down) kernels; else part of PST_K_TRI3 */
#define PST_K_NCLASS   12

const char *pst_last_error(void);
const char *pst_version(void);
int  pst_device_count(void);

/* One context per GPU (device ordinal).  */
int  pst_ctx_create(int device, pst_ctx **ctx);
void pst_ctx_destroy(pst_ctx *ctx);
int  pst_ctx_stats(pst_ctx *ctx, pst_stats *out);
/* The host-pointer entry points reset the statistics themselves; the *_dev entry points
 * accumulate, so reset explicitly around a device-resident sequence. */
int  pst_ctx_reset_stats(pst_ctx *ctx);
/* Per-launch CUDA-event timing by kernel class (off by default; costs two event records per launch). */
int  pst_ctx_set_profile(pst_ctx *ctx, int on);
/* CUDA-event stopwatch on the context's stream (torch.cuda.Event cannot see this stream). */
int  pst_timer_start(pst_ctx *ctx);
int  pst_timer_stop(pst_ctx *ctx, double *elapsed_ms);
/* Multi-GPU: slab decomposition along n3 over `nranks` processes (one per GPU).  `nccl_id`
 * is the 128-byte ncclUniqueId created by rank 0 (pst_comm_unique_id) and distributed by
 * the host side (torch.distributed / MPI / files). */
int  pst_comm_unique_id(void *id128);
int  pst_ctx_create_dist(int device, int rank, int nranks, const void *nccl_id128, pst_ctx **ctx);
/* The library's slab rule: rank r of G owns global planes [n3*r/G, n3*(r+1)/G).  In a distributed
 * context every entry point takes the GLOBAL n3 and pointers to this rank's slab (for pst_dip the
 * output is the slab of the inline dip followed by the slab of the xline dip). */
int  pst_ctx_slab(pst_ctx *ctx, int n3, int *z0, int *z1);

/* ---- dip estimation.  Replaces dipcfun.dipc (reference pyseistr/src/dip_cfuns.c:1694-1989,
 * "Oiiiiiifffiiiii"); called by dip3dc (pyseistr/dip3d.py:59-116) and dip2dc
 * (pyseistr/dip2d.py:115-221, n3=1).  mask may be NULL (hasmask=0).  eps_dv, eps_cg and
 * tol_cg are accepted and IGNORED exactly as the reference's C does (divne runs
 * with eps=1, CG with eps=1, tol=1e-6).  dip_out: n1*n2*n3 floats when n3==1, else
 * 2*n1*n2*n3 (inline dip then xline dip). */
int pst_dip(pst_ctx *ctx, const float *din, const float *mask, int n1, int n2, int n3,
            int niter, int liter, int order, float eps_dv, float eps_cg, float tol_cg,
            int r1, int r2, int r3, int verb, float *dip_out);
int pst_dip_dev(pst_ctx *ctx, const float *d_din, const float *d_mask, int n1, int n2, int n3,
                int niter, int liter, int order, int r1, int r2, int r3, int verb,
                float *d_dip_out);

/* ---- 3-D structure-oriented mean / median.  Replace sof3dcfun.csomean3d
 * (sof3d_cfuns.c:1355-1552, "OOOiiiiiifi") and sof3dcfun.csomf3d (:1554-1752,
 * "OOOiiiiiiiifi"); called by somean3dc (pyseistr/somean3d.py:36-73) and somf3dc
 * (pyseistr/somf3d.py:54-97).  eps is accepted and IGNORED (the reference overwrites it
 * with 0.01).  option: 1 = median filter (MF). */
int pst_somean3d(pst_ctx *ctx, const float *din, const float *dipi, const float *dipx,
                 int n1, int n2, int n3, int ns2, int ns3, int order, float eps, int verb,
                 float *out);
int pst_somf3d(pst_ctx *ctx, const float *din, const float *dipi, const float *dipx,
               int n1, int n2, int n3, int ns2, int ns3, int nmf, int option, int order,
               float eps, int verb, float *out);
int pst_somean3d_dev(pst_ctx *ctx, const float *d_din, const float *d_dipi, const float *d_dipx,
                     int n1, int n2, int n3, int ns2, int ns3, int order, float *d_out);
int pst_somf3d_dev(pst_ctx *ctx, const float *d_din, const float *d_dipi, const float *d_dipx,
                   int n1, int n2, int n3, int ns2, int ns3, int nmf, int option, int order,
                   float *d_out);

/* ---- 2-D structure-oriented mean / median.  Replace sofcfun.csomean2d
 * (pyseistr/src/sof_cfuns.c:1433-1532, "OOiiiiiifi") and sofcfun.csomf2d (:1534-1672,
 * "OOiiiiiiifi"); called by somean2dc (pyseistr/somean2d.py:36-74) and somf2dc
 * (pyseistr/somf2d.py:60-105).  Here eps IS honoured (regularisation eps*eps). */
int pst_somean2d(pst_ctx *ctx, const float *din, const float *dip, int n1, int n2, int n3,
                 int ns, int order, int adj, float eps, int verb, float *out);
int pst_somf2d(pst_ctx *ctx, const float *din, const float *dip, int n1, int n2, int n3,
               int ns, int nmf, int option, int order, float eps, int verb, float *out);
/* device-pointer variants (forward operator; adj = 1 of csomean2d is host-pointer only) */
int pst_somean2d_dev(pst_ctx *ctx, const float *d_din, const float *d_dip, int n1, int n2, int n3,
                     int ns, int order, float eps, float *d_out);
int pst_somf2d_dev(pst_ctx *ctx, const float *d_din, const float *d_dip, int n1, int n2, int n3,
                   int ns, int nmf, int option, int order, float eps, float *d_out);

/* ---- 3-D structure-oriented interpolation (PWD-residual CG).  Replaces soint3dcfun.csoint3d
 * (pyseistr/src/soint3d_cfuns.c:2405-2508, "OOOOiiiiiiiiiifi"); called by soint3dc
 * (pyseistr/soint3d.py:65-108).  GPU path: any njs >= 1; drift = 0 only (drift != 0: PST_EUNSUP).
 * hasmask=1: known samples are mask != 0, else din != 0.  var > 0 puts sqrt(var)*N(0,1) on the right-hand
 * side, drawn on the host from MT19937(seed) + Box-Muller exactly like the reference (:2304-2402).
 * Distributed contexts (var = 0): global n3, slab pointers; one halo plane per operator application. */
int pst_soint3d(pst_ctx *ctx, const float *din, const float *mask, const float *dipi, const float *dipx,
                int n1, int n2, int n3, int nw, int nj1, int nj2, int niter, int drift, int seed,
                int hasmask, float var, int verb, float *out);
int pst_soint3d_dev(pst_ctx *ctx, const float *d_din, const float *d_mask, const float *d_dipi,
                    const float *d_dipx, int n1, int n2, int n3, int nw, int nj1, int nj2, int niter,
                    int drift, int seed, int hasmask, float var, int verb, float *d_out);

/* ---- 3-D interpolation by shaping-regularised CG with the plane-wave smoother as shaping operator.
 * Replaces soint3dcfun.csint3d (pyseistr/src/soint3d_cfuns.c:2510-2640, "OOOOiiiiiiiiif": din, dipi,
 * dipx, mask, n1, n2, n3, niter, ns1, ns2, order1, order2, verb, eps); called by sint3dc
 * (pyseistr/sint.py:97-131).  Known samples are mask != 0; eps is the regularisation of the
 * plane-wave predictions (squared internally, like the reference). */
int pst_sint3d(pst_ctx *ctx, const float *din, const float *dipi, const float *dipx, const float *mask,
               int n1, int n2, int n3, int niter, int ns1, int ns2, int order1, int order2, int verb,
               float eps, float *out);
int pst_sint3d_dev(pst_ctx *ctx, const float *d_din, const float *d_dipi, const float *d_dipx,
                   const float *d_mask, int n1, int n2, int n3, int niter, int ns1, int ns2,
                   int order1, int order2, int verb, float eps, float *d_out);

/* ---- plane-wave painting.  Replaces paint2dcfun.cpaint2d / cpaint3d (pyseistr/src/paint_cfuns.c:1861-2024,
 * "OOiiiifi": dip, trace, n1, n2, order, i0, eps, verb; the two reference entries are the same code); called by pwpaintc
 * and rgt (pyseistr/rgt.py).  The seed trace (n1 floats) sits at trace i0 and is predicted outwards trace by trace with
 * regularisation eps*eps: one dependent chain of n2 - 1 predictions.  out: n1*n2 floats.  Single-GPU contexts. */
int pst_paint2d(pst_ctx *ctx, const float *dip, const float *seed, int n1, int n2, int order, int i0, float eps,
                int verb, float *out);
int pst_paint2d_dev(pst_ctx *ctx, const float *d_dip, const float *d_seed, int n1, int n2, int order, int i0,
                    float eps, float *d_out);

/* ---- building blocks exposed for parity tests and for the smoothing wrapper (SURVEY §8f
 * rank 1: dipcfun.smoothcf, dip_cfuns.c:2006-2123 with adj=0).  Device pointers. */
int pst_allpass_dev(pst_ctx *ctx, const float *d_u, const float *d_sigma, int n1, int n2, int n3,
                    int order, int xline, int der, float *d_y);
int pst_smooth3_dev(pst_ctx *ctx, float *d_x, int n1, int n2, int n3, int r1, int r2, int r3, int repeat, int adj);
/* Test hook: the kernels of the MULTI-GPU axis-3 smoothing pass (pst_ctx_create_dist contexts, short n3-slabs) run rank
 * after rank on one GPU over the nranks n3-slabs of d_x, in place; equals pst_smooth3_dev(.., 1, 1, r3, 1, 0) bit for bit.
 * PST_EUNSUP outside those kernels' geometry (radius 2-6 or 8, nranks >= 2 equal slabs of 32, 64, 96 or k x 128 planes). */
int pst_selftest_axis3_slabs(pst_ctx *ctx, float *d_x, int n1, int n2, int n3, int r3, int nranks);
int pst_divne_dev(pst_ctx *ctx, float *d_num, float *d_den, float *d_rat, int n1, int n2, int n3,
                  int r1, int r2, int r3, int liter, int *iters_run);
int pst_smooth3(pst_ctx *ctx, const float *x, int n1, int n2, int n3, int r1, int r2, int r3,
                int repeat, int adj, float *out);

/* ---- N-D triangle / box smoothing with every option.  Replaces dipcfun.smoothcf (pyseistr/src/dip_cfuns.c:2006-2123,
 * "Oiiiiiiiiiiiiii": din, n1, n2, n3, repeat, adj, r1, r2, r3, diff1, diff2, diff3, box1, box2, box3); called by smoothc
 * (pyseistr/smooth.py:115-183, whose default is adj = 1).  adj = 0: ps_smooth2, adj = 1: ps_smooth. */
int pst_smoothcf(pst_ctx *ctx, const float *x, int n1, int n2, int n3, int repeat, int adj, int r1, int r2, int r3,
                 int diff1, int diff2, int diff3, int box1, int box2, int box3, float *out);
int pst_smoothcf_dev(pst_ctx *ctx, float *d_x, int n1, int n2, int n3, int repeat, int adj, int r1, int r2, int r3,
                     int diff1, int diff2, int diff3, int box1, int box2, int box3);

/* ---- raw device memory helpers so that a host language without a CUDA binding can keep
 * volumes resident between calls (bench `value` leg, pipelines dip -> somf). */
int pst_dev_alloc(pst_ctx *ctx, size_t bytes, void **d_ptr);
int pst_dev_free(pst_ctx *ctx, void *d_ptr);
int pst_h2d(pst_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int pst_d2h(pst_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
int pst_host_alloc_pinned(size_t bytes, void **h_ptr);
int pst_host_free_pinned(void *h_ptr);
int pst_sync(pst_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
