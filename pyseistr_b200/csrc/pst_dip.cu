// dip3d / dip2d local-slope estimation on sm_100a.
//
// Replaces (reference pyseistr/src/dip_cfuns.c): apfilt/passfilter/aderfilter :835-910,
// allpass1/allpass2 :1135-1200,:1399-1465, mask32 :914-997, ps_smooth2/ps_trianglen_lop
// :458-727, ps_weight_lop :739-758, ps_conjgrad :257-383, ps_divne :796-827, dip3 :1619-1691
// and the dipc driver :1694-1989.
//
// Arithmetic contract: every float operation is performed in the reference's order with the
// reference's float/double placement and without FMA contraction (-fmad=false), so the
// vectors are bit-identical to the reference's; only the global sums (double tree
// reductions here, sequential sums there) may differ in the last bits of a double.
#include "pst_common.cuh"
#include "pst_tri3_reg_core.h"
#include "pst_tri_stream.cuh"
#include "pst_tri_rc.cuh"
#include "pst_tri_sys.cuh"
#include "pst_tri_l2.cuh"
#ifndef PST_TRI_L2_DEFAULT
#define PST_TRI_L2_DEFAULT 1      /* axis 1 (contiguous lines): measured fastest there */
#endif

#include <math.h>
#include <stdlib.h>
#include <algorithm>

struct BTab { double b[PST_MAXTAP]; };

static BTab make_btab(int nw)      // apfilt_init, dip_cfuns.c:838-855 (host, IEEE double)
{
    BTab t{};
    const int nf = 2 * nw;
    for (int k = 0; k <= nf; k++) {
        double bk = 1.0;
        for (int j = 0; j < nf; j++) {
            if (j < nf - k) bk *= (k + j + 1.0) / (2 * (2 * j + 1) * (j + 1));
            else            bk *= 1.0 / (2 * (2 * j + 1));
        }
        t.b[k] = bk;
    }
    return t;
}

// linear factor j of tap k, evaluated in float as the reference's (nf-j-p) / (p+j+1)
template <int NF>
__device__ __forceinline__ float pst_bracket(int j, int k, float p)
{
    return (j < NF - k) ? ((float)(NF - j) - p) : ((p + (float)j) + 1.0f);
}

template <int NW>
__device__ __forceinline__ void pst_passfilter(const BTab &tb, float p, float (&a)[2 * NW + 1])
{
    constexpr int NF = 2 * NW;
#pragma unroll
    for (int k = 0; k <= NF; k++) {
        double ak = tb.b[k];
#pragma unroll
        for (int j = 0; j < NF; j++) ak *= (double)pst_bracket<NF>(j, k, p);
        a[k] = (float)ak;
    }
}

template <int NW>
__device__ __forceinline__ void pst_aderfilter(const BTab &tb, float p, float (&a)[2 * NW + 1])
{
    constexpr int NF = 2 * NW;
#pragma unroll
    for (int k = 0; k <= NF; k++) {
        double ak = 0.;
#pragma unroll
        for (int i = 0; i < NF; i++) {
            double ai = -1.0;
#pragma unroll
            for (int j = 0; j < NF; j++) {
                if (j != i) ai *= (double)pst_bracket<NF>(j, k, p);
                else if (j < NF - k) ai = -ai;
            }
            ak += ai;
        }
        a[k] = (float)(ak * tb.b[k]);
    }
}

// ---------------------------------------------------------------------------------------
// G1: PWD stencil.  y[i] = sum_w (u[i+(w-nw)+ip] - u[i-(w-nw)]) * flt_w(p[i]),  ip = n1
// (inline) or n1*n2 (xline); zero on the nw border rows and on the last trace / plane.
// LS variant fuses the line-search update p = p0 + lam*dp (dip3 :1669-1675) in front.
// Always emits the block partial of sum(y^2) (usum / usum2 of dip3 :1650-1654,:1681-1685).
// One block walks a group of traces, threads walk i1: fully coalesced.
template <int NW, bool DER, bool LS>
__global__ void __launch_bounds__(256)
allpass_kernel(const float *__restrict__ u, const float *__restrict__ p_in,
               const float *__restrict__ dp, float lam, float *__restrict__ p_out,
               float *__restrict__ y, int n1, int n2, int n3, int xline, int n3_live, BTab tb,
               unsigned gpp, int tg, double *__restrict__ partial)
{
    const long ip = xline ? (long)n1 * n2 : (long)n1;
    double acc2[1] = {0.0};
    // block = group (blockIdx.x % gpp) of tg consecutive traces of plane blockIdx.x / gpp: the pieces of the
    // canonical sums (pst_common.cuh); tg depends on n1 only
    const int i3 = (int)(blockIdx.x / gpp), t0 = (int)(blockIdx.x - (unsigned)i3 * gpp) * tg;
    const int t1 = t0 + tg < n2 ? t0 + tg : n2;
    for (int i2 = t0; i2 < t1; i2++) {
        const long tr = (long)i3 * n2 + i2;
        const bool live_tr = xline ? (i3 < n3_live) : (i2 < n2 - 1);
        const long base = tr * n1;
        for (int i1 = threadIdx.x; i1 < n1; i1 += blockDim.x) {
            const long i = base + i1;
            float sg;
            if (LS) {
                float pi = p_in[i] + lam * dp[i];
                const float pmax = 3.402823466e+38F;      // dipc :1779-1782 (+-FLT_MAX)
                if (pi < -pmax) pi = -pmax;
                if (pi > pmax) pi = pmax;
                p_out[i] = pi;
                sg = pi;
            } else {
                sg = p_in[i];
            }
            float out = 0.f;
            if (live_tr && i1 >= NW && i1 < n1 - NW) {
                float flt[2 * NW + 1];
                if (DER) pst_aderfilter<NW>(tb, sg, flt);
                else     pst_passfilter<NW>(tb, sg, flt);
#pragma unroll
                for (int w = 0; w <= 2 * NW; w++) {
                    const int s = w - NW;
                    out += (u[i + s + ip] - u[i - s]) * flt[w];
                }
            }
            y[i] = out;
            acc2[0] += (double)out * (double)out;
        }
    }
    pst_block_reduce<1>(acc2, partial);
}

// G5: mask32 (both=false, nj=1): footprint of either stencil touches a zero sample
template <int NW>
__global__ void mask_kernel(const float *__restrict__ um, unsigned char *__restrict__ m_in,
                            unsigned char *__restrict__ m_x, int n1, int n2, int n3, int n3_live)
{
    const long ntr = (long)n2 * n3;
    const long pl = (long)n1 * n2;
    for (long tr = blockIdx.x; tr < ntr; tr += gridDim.x) {
        const int i2 = (int)(tr % n2), i3 = (int)(tr / n2);
        for (int i1 = threadIdx.x; i1 < n1; i1 += blockDim.x) {
            const long i = tr * n1 + i1;
            bool a = false, b = false;
            if (i1 >= NW && i1 < n1 - NW) {
                if (i2 < n2 - 1)
                    for (int s = -NW; s <= NW; s++) a = a || (um[i - s] == 0.f) || (um[i + n1 + s] == 0.f);
                if (i3 < n3_live)
                    for (int s = -NW; s <= NW; s++) b = b || (um[i - s] == 0.f) || (um[i + pl + s] == 0.f);
            }
            m_in[i] = a;
            m_x[i] = b;
        }
    }
}

// ---------------------------------------------------------------------------------------
// G2: triangle smoothing of every line of one axis, reference arithmetic (ps_smooth2):
//   t_k = ((-wt x_k) + 2wt x_{k-nb}) + (-wt x_{k-2nb})      k in [0, nx+2nb)
//   F_k = F_{k-1} + t_k   (forward running sum, float, sequential)
//   B_k = B_{k+1} + F_k   (backward running sum)
//   y_i = (B_{i+nb} + B_{nb+nx+(nx-1-i)}[i >= nx-nb]) + B_{nb-1-i}[i < nb]     (fold2)
// The two running sums are inherently serial per line (bit-exact float rounding order);
// the parallelism is across lines.

__device__ __forceinline__ float tri_spread(const float *xl, long d, int k, int nx,
                                            int nb, float wm, float w2)
{
    float t = 0.f;
    if (k < nx) t = t + wm * xl[(long)k * d];
    if (k >= nb && k - nb < nx) t = t + w2 * xl[(long)(k - nb) * d];
    if (k >= 2 * nb && k - 2 * nb < nx) t = t + wm * xl[(long)(k - 2 * nb) * d];
    return t;
}

// Strided axes (2 and 3; also the slow-but-correct fallback for axis 1): one thread per
// line, adjacent threads on adjacent i1 so every access is coalesced.  F is staged in a
// global scratch volume laid out [.. k ..][ia] (extended axis), read back in the backward
// sweep which also performs the fold.  Requires nb <= nx.
__global__ void __launch_bounds__(256)
tri_lines_kernel(float *x, float *scr, long nlines, long na, long sa,
                 long sb, long d, int nx, int nb, float wt, float w2)
{
    const long l = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    const long ia = l % na, ib = l / na;
    float *xl = x + ia * sa + ib * sb;
    const int np = nx + 2 * nb;
    float *sl = scr + ia + na * ((long)np * ib);
    const float wm = -wt;
    float s = 0.f;
    for (int k = 0; k < np; k++) {
        s += tri_spread(xl, d, k, nx, nb, wm, w2);
        sl[(long)k * na] = s;
    }
    s = 0.f;
    int k = np - 1;
    for (; k >= nb + nx; k--) {                    // right tail: stash into its target
        s += sl[(long)k * na];
        xl[(long)(nx - 1 - (k - nb - nx)) * d] = s;
    }
    for (; k >= nb; k--) {                         // middle (+ stashed right reflection)
        s += sl[(long)k * na];
        const int i = k - nb;
        float v = s;
        if (i >= nx - nb) v = v + xl[(long)i * d];
        xl[(long)i * d] = v;
    }
    for (; k >= 0; k--) {                          // left tail
        s += sl[(long)k * na];
        xl[(long)(nb - 1 - k) * d] += s;
    }
}

// Literal fold for nb > nx (multiple reflections; tiny axes): scratch holds B after the
// sweeps, then the reference's fold2 loops (:458-484) are followed verbatim.
__global__ void tri_lines_literal_kernel(float *x, float *scr,
                                         long nlines, long na, long sa, long sb, long d, int nx,
                                         int nb, float wt, float w2)
{
    const long l = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    const long ia = l % na, ib = l / na;
    float *xl = x + ia * sa + ib * sb;
    const int np = nx + 2 * nb;
    float *sl = scr + ia + na * ((long)np * ib);
    const float wm = -wt;
    float s = 0.f;
    for (int k = 0; k < np; k++) { s += tri_spread(xl, d, k, nx, nb, wm, w2); sl[(long)k * na] = s; }
    s = 0.f;
    for (int k = np - 1; k >= 0; k--) { s += sl[(long)k * na]; sl[(long)k * na] = s; }
    for (int i = 0; i < nx; i++) xl[(long)i * d] = sl[(long)(i + nb) * na];
    for (int j = nb + nx; j < np; j += nx) {
        for (int i = 0; i < nx && i < np - j; i++) xl[(long)(nx - 1 - i) * d] += sl[(long)(j + i) * na];
        j += nx;
        for (int i = 0; i < nx && i < np - j; i++) xl[(long)i * d] += sl[(long)(j + i) * na];
    }
    for (int j = nb; j >= 0; j -= nx) {
        for (int i = 0; i < nx && i < j; i++) xl[(long)i * d] += sl[(long)(j - 1 - i) * na];
        j -= nx;
        for (int i = 0; i < nx && i < j; i++) xl[(long)(nx - 1 - i) * d] += sl[(long)(j - 1 - i) * na];
    }
}

// One line of smoothcf (dip_cfuns.c:2084-2098) with every option, one thread per line, the extended line in global
// scratch (user-facing options only; dip3d's shaping operator goes through the streaming kernel):
//   adj = 0: ps_smooth2 :616-625 = triple2 :560-577 (box: +wt at 1, -wt at 2nb), doubint2 :508-529 (forward sum, then
//            backward unless box || der), fold2 :458-484
//   adj = 1: ps_smooth :591-603 = fold :439-456, doubint :487-505 (backward sum, then forward unless box || der),
//            triple :531-547 (box: (tmp[i+1] - tmp[i+2nb]) * wt in float; triangle: 2.*tmp1 - tmp - tmp2 in double)
__global__ void tri_lines_any_kernel(float *x, float *scr, long nlines, long na, long sa, long sb, long d, int nx,
                                     int nb, float wt, int adj, int box, int der)
{
    const long l = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    const long ia = l % na, ib = l / na;
    float *xl = x + ia * sa + ib * sb;
    const int np = nx + 2 * nb;
    float *sl = scr + ia + na * ((long)np * ib);
    const bool single = box || der;
#define T_(k) sl[(long)(k) * na]
#define X_(i) xl[(long)(i) * d]
    if (!adj) {
        for (int k = 0; k < np; k++) T_(k) = 0.f;
        if (box) {
            const float wp = +wt, wm = -wt;
            for (int i = 0; i < nx; i++) T_(i + 1) += wp * X_(i);
            for (int i = 0; i < nx; i++) T_(i + 2 * nb) += wm * X_(i);
        } else {
            const float w2 = (float)(2. * wt), wm = -wt;
            for (int i = 0; i < nx; i++) T_(i) += wm * X_(i);
            for (int i = 0; i < nx; i++) T_(i + nb) += w2 * X_(i);
            for (int i = 0; i < nx; i++) T_(i + 2 * nb) += wm * X_(i);
        }
        float s = 0.f;
        for (int k = 0; k < np; k++) { s += T_(k); T_(k) = s; }
        if (!single) { s = 0.f; for (int k = np - 1; k >= 0; k--) { s += T_(k); T_(k) = s; } }
        for (int i = 0; i < nx; i++) X_(i) = T_(i + nb);
        for (int j = nb + nx; j < np; j += nx) {
            for (int i = 0; i < nx && i < np - j; i++) X_(nx - 1 - i) += T_(j + i);
            j += nx;
            for (int i = 0; i < nx && i < np - j; i++) X_(i) += T_(j + i);
        }
        for (int j = nb; j >= 0; j -= nx) {
            for (int i = 0; i < nx && i < j; i++) X_(i) += T_(j - 1 - i);
            j -= nx;
            for (int i = 0; i < nx && i < j; i++) X_(nx - 1 - i) += T_(j - 1 - i);
        }
    } else {
        for (int i = 0; i < nx; i++) T_(i + nb) = X_(i);
        for (int j = nb + nx; j < np; j += nx) {
            for (int i = 0; i < nx && i < np - j; i++) T_(j + i) = X_(nx - 1 - i);
            j += nx;
            for (int i = 0; i < nx && i < np - j; i++) T_(j + i) = X_(i);
        }
        for (int j = nb; j >= 0; j -= nx) {
            for (int i = 0; i < nx && i < j; i++) T_(j - 1 - i) = X_(i);
            j -= nx;
            for (int i = 0; i < nx && i < j; i++) T_(j - 1 - i) = X_(nx - 1 - i);
        }
        float s = 0.f;
        for (int k = np - 1; k >= 0; k--) { s += T_(k); T_(k) = s; }
        if (!single) { s = 0.f; for (int k = 0; k < np; k++) { s += T_(k); T_(k) = s; } }
        if (box) {
            for (int i = 0; i < nx; i++) X_(i) = (T_(i + 1) - T_(i + 2 * nb)) * wt;
        } else {
            for (int i = 0; i < nx; i++) {
                const double t0 = T_(i), t1 = T_(i + nb), t2 = T_(i + 2 * nb);
                X_(i) = (float)((2. * t1 - t0 - t2) * (double)wt);
            }
        }
    }
#undef T_
#undef X_
}

// Axis 1 (contiguous lines): a CTA stages LPC whole lines in shared memory.  Phase 1: all
// threads build t_k with coalesced loads.  Phase 2: LPC threads run the two serial running
// sums in shared memory (row pitch odd => conflict-free).  Phase 3: all threads fold and
// store coalesced.  HBM traffic is the compulsory 8 B/voxel.  Requires nb <= nx.
template <int LPC>
__global__ void __launch_bounds__(128)
tri_axis1_kernel(float *x, long nlines, int nx, int nb, float wt, float w2, int pitch)
{
    extern __shared__ float tile[];
    const int np = nx + 2 * nb;
    const float wm = -wt;
    for (long l0 = (long)blockIdx.x * LPC; l0 < nlines; l0 += (long)gridDim.x * LPC) {
        const int nl = (int)min((long)LPC, nlines - l0);
        for (int r = 0; r < nl; r++) {
            const float *xl = x + (l0 + r) * nx;
            float *tr = tile + (size_t)r * pitch;
            for (int k = threadIdx.x; k < np; k += blockDim.x) tr[k] = tri_spread(xl, 1, k, nx, nb, wm, w2);
        }
        __syncthreads();
        if (threadIdx.x < nl) {
            float *tr = tile + (size_t)threadIdx.x * pitch;
            float s = 0.f;
            int k = 0;
            for (; k + 8 <= np; k += 8) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = tr[k + q];
#pragma unroll
                for (int q = 0; q < 8; q++) { s += v[q]; tr[k + q] = s; }
            }
            for (; k < np; k++) { s += tr[k]; tr[k] = s; }
            s = 0.f;
            k = np - 1;
            for (; k - 7 >= 0; k -= 8) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = tr[k - q];
#pragma unroll
                for (int q = 0; q < 8; q++) { s += v[q]; tr[k - q] = s; }
            }
            for (; k >= 0; k--) { s += tr[k]; tr[k] = s; }
        }
        __syncthreads();
        for (int r = 0; r < nl; r++) {
            float *xl = x + (l0 + r) * nx;
            const float *tr = tile + (size_t)r * pitch;
            for (int i = threadIdx.x; i < nx; i += blockDim.x) {
                float v = tr[i + nb];
                if (i >= nx - nb) v = v + tr[nb + nx + (nx - 1 - i)];
                if (i < nb) v = v + tr[nb - 1 - i];
                xl[i] = v;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// G2, main path: shared-memory tile kernel for any axis.  A CTA owns W whole lines:
//   phase 1 (all threads)  tile <- t_k, built from coalesced global loads with deep MLP;
//   phase 2 (W threads)    the two serial running sums, in shared memory (the only serial part);
//   phase 3 (all threads)  fold, optional fused CG epilogue, coalesced store.
// HBM traffic is the compulsory read + write (8 B/voxel); several CTAs per SM overlap the
// phases.  CONTIG: axis 1, tile[w][pitch] (pitch = 4 mod 32 words: conflict-free 128-bit
// accesses for the serial threads).  Strided axes: tile[k][16], lanes on adjacent lines.
// The epilogue of the LAST smoothed axis absorbs the CG vector work that follows the shaping
// operator in ps_conjgrad (:318-361), so gx/gr never touch HBM.
enum { EPI_NONE = 0, EPI_GP = 1, EPI_DIR_FIRST = 2, EPI_DIR = 3 };

struct TriArgs {
    const float *src;
    float *dst;
    long ngroups;          // tiles
    long na, sb, d;        // strided: fastest line-index extent, outer line stride, element stride
    long nlines;           // contiguous: number of lines
    int nx, nb, W, pitch;
    float wt, w2;
    double bytes;          // algorithmic bytes of this launch (profiling only)
    // epilogue operands
    const float *p, *w;
    float *gp, *sp, *sx, *sr;
    float eps, alpha;
    double *partial;
};

__device__ __forceinline__ float tri_fold(const float *col, int stride, int i, int nx, int nb)
{
    float v = col[(size_t)(i + nb) * stride];
    if (i >= nx - nb) v = v + col[(size_t)(nb + nx + (nx - 1 - i)) * stride];
    if (i < nb) v = v + col[(size_t)(nb - 1 - i) * stride];
    return v;
}

// epilogue operands of one element, loaded ahead of use so that U elements' loads are in flight
struct EpiIn { float p, gp, w, sp, sx, sr; };

template <int EPI>
__device__ __forceinline__ void tri_epi_load(const TriArgs &A, long gi, EpiIn &e)
{
    if (EPI == EPI_GP) e.p = A.p[gi];
    if (EPI == EPI_DIR_FIRST || EPI == EPI_DIR) { e.gp = A.gp[gi]; e.w = A.w[gi]; }
    if (EPI == EPI_DIR) { e.sp = A.sp[gi]; e.sx = A.sx[gi]; e.sr = A.sr[gi]; }
}

template <int EPI>
__device__ __forceinline__ void tri_epi_apply(const TriArgs &A, long gi, float v, const EpiIn &e, double (&acc)[3])
{
    if (EPI == EPI_NONE) {
        A.dst[gi] = v;
    } else if (EPI == EPI_GP) {                 // gp = eps*p + S(gx)   (:303,:318)
        float g = A.eps * e.p;
        g += v;
        A.gp[gi] = g;
        acc[0] += (double)g * (double)g;
    } else {                                    // gx = S(gp); gr = gx*w; s = g (+ alpha*s)  (:319-361)
        const float gxi = 0.f + v;
        const float gri = 0.f + gxi * e.w;
        float a, b, c;
        if (EPI == EPI_DIR_FIRST) { a = e.gp; b = gxi; c = gri; }
        else {
            a = e.gp + A.alpha * e.sp;
            b = gxi + A.alpha * e.sx;
            c = gri + A.alpha * e.sr;
        }
        A.sp[gi] = a; A.sx[gi] = b; A.sr[gi] = c;
        acc[0] += (double)c * (double)c;
        acc[1] += (double)a * (double)a;
        acc[2] += (double)b * (double)b;
    }
}

struct EpiIn4 { float4 p, gp, w, sp, sx, sr; };

template <int EPI>
__device__ __forceinline__ void tri_epi_load4(const TriArgs &A, long gi, EpiIn4 &e)
{
    if (EPI == EPI_GP) e.p = *reinterpret_cast<const float4 *>(A.p + gi);
    if (EPI == EPI_DIR_FIRST || EPI == EPI_DIR) {
        e.gp = *reinterpret_cast<const float4 *>(A.gp + gi);
        e.w = *reinterpret_cast<const float4 *>(A.w + gi);
    }
    if (EPI == EPI_DIR) {
        e.sp = *reinterpret_cast<const float4 *>(A.sp + gi);
        e.sx = *reinterpret_cast<const float4 *>(A.sx + gi);
        e.sr = *reinterpret_cast<const float4 *>(A.sr + gi);
    }
}

template <int EPI>
__device__ __forceinline__ void tri_epi_lane(const TriArgs &A, float v, float p, float gp, float w, float sp,
                                             float sx, float sr, float &o0, float &o1, float &o2, double (&acc)[3])
{
    if (EPI == EPI_NONE) {
        o0 = v;
    } else if (EPI == EPI_GP) {
        float g = A.eps * p;
        g += v;
        o0 = g;
        acc[0] += (double)g * (double)g;
    } else {
        const float gxi = 0.f + v;
        const float gri = 0.f + gxi * w;
        float a, b, c;
        if (EPI == EPI_DIR_FIRST) { a = gp; b = gxi; c = gri; }
        else { a = gp + A.alpha * sp; b = gxi + A.alpha * sx; c = gri + A.alpha * sr; }
        o0 = a; o1 = b; o2 = c;
        acc[0] += (double)c * (double)c;
        acc[1] += (double)a * (double)a;
        acc[2] += (double)b * (double)b;
    }
}

template <int EPI>
__device__ __forceinline__ void tri_epi_apply4(const TriArgs &A, long gi, float4 v, const EpiIn4 &e, double (&acc)[3])
{
    float4 o0, o1, o2;
    tri_epi_lane<EPI>(A, v.x, e.p.x, e.gp.x, e.w.x, e.sp.x, e.sx.x, e.sr.x, o0.x, o1.x, o2.x, acc);
    tri_epi_lane<EPI>(A, v.y, e.p.y, e.gp.y, e.w.y, e.sp.y, e.sx.y, e.sr.y, o0.y, o1.y, o2.y, acc);
    tri_epi_lane<EPI>(A, v.z, e.p.z, e.gp.z, e.w.z, e.sp.z, e.sx.z, e.sr.z, o0.z, o1.z, o2.z, acc);
    tri_epi_lane<EPI>(A, v.w, e.p.w, e.gp.w, e.w.w, e.sp.w, e.sx.w, e.sr.w, o0.w, o1.w, o2.w, acc);
    if (EPI == EPI_NONE) *reinterpret_cast<float4 *>(A.dst + gi) = o0;
    else if (EPI == EPI_GP) *reinterpret_cast<float4 *>(A.gp + gi) = o0;
    else {
        *reinterpret_cast<float4 *>(A.sp + gi) = o0;
        *reinterpret_cast<float4 *>(A.sx + gi) = o1;
        *reinterpret_cast<float4 *>(A.sr + gi) = o2;
    }
}

// t_k for U rows/samples at once: all 3U loads are unconditional (clamped index) and issued
// before any use; out-of-range taps are masked to zero afterwards (adding +-0 does not change
// the value the reference computes by skipping the tap).
template <int U>
__device__ __forceinline__ void tri_spread_batch(const float *xl, long d, int k0, int kstep, int nx, int nb,
                                                 float wm, float w2, float (&t)[U])
{
    float xa[U], xb[U], xc[U];
#pragma unroll
    for (int q = 0; q < U; q++) {
        const int k = k0 + q * kstep;
        const int ia = min(k, nx - 1);
        const int ib = min(max(k - nb, 0), nx - 1);
        const int ic = min(max(k - 2 * nb, 0), nx - 1);
        xa[q] = xl[(long)ia * d];
        xb[q] = xl[(long)ib * d];
        xc[q] = xl[(long)ic * d];
    }
#pragma unroll
    for (int q = 0; q < U; q++) {
        const int k = k0 + q * kstep;
        const float va = (k < nx) ? xa[q] : 0.f;
        const float vb = (k >= nb && k - nb < nx) ? xb[q] : 0.f;
        const float vc = (k >= 2 * nb && k - 2 * nb < nx) ? xc[q] : 0.f;
        float v = 0.f;
        v = v + wm * va;
        v = v + w2 * vb;
        v = v + wm * vc;
        t[q] = v;
    }
}

__device__ __forceinline__ float tri_t(float va, float vb, float vc, float wm, float w2)
{
    float v = 0.f;
    v = v + wm * va;
    v = v + w2 * vb;
    v = v + wm * vc;
    return v;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ float4 f4sel(bool c, float4 a)
{
    return c ? a : make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- generic (scalar) tile kernel: any n1, any alignment ---------------------------------
template <bool CONTIG, int EPI>
__global__ void __launch_bounds__(128)
tri_tile_kernel(const TriArgs A)
{
    extern __shared__ __align__(16) float tile[];
    const int nx = A.nx, nb = A.nb, np = nx + 2 * nb, W = A.W, tid = threadIdx.x;
    const float wm = -A.wt, w2 = A.w2;
    double acc[3] = {0.0, 0.0, 0.0};
    const long gpb = CONTIG ? 1 : (A.na + W - 1) / W;
    constexpr int U = 8;
    for (long g = blockIdx.x; g < A.ngroups; g += gridDim.x) {
        long base;
        int nw;
        if (CONTIG) {
            const long l0 = g * W;
            nw = (int)min((long)W, A.nlines - l0);
            base = l0 * nx;
        } else {
            const long ib = g / gpb, ia0 = (g % gpb) * W;
            nw = (int)min((long)W, A.na - ia0);
            base = ia0 + ib * A.sb;
        }
        // ---- phase 1: tile <- t_k
        if (CONTIG) {
            for (int w = 0; w < nw; w++) {
                const float *xl = A.src + base + (long)w * nx;
                float *row = tile + (size_t)w * A.pitch;
                for (int k0 = tid; k0 < np; k0 += 128 * U) {
                    float t[U];
                    tri_spread_batch<U>(xl, 1, k0, 128, nx, nb, wm, w2, t);
#pragma unroll
                    for (int q = 0; q < U; q++) if (k0 + q * 128 < np) row[k0 + q * 128] = t[q];
                }
            }
        } else {
            const int w = tid & 15, kr = tid >> 4;
            const float *xl = A.src + base + min(w, nw - 1);
            for (int k0 = kr; k0 < np; k0 += 8 * U) {
                float t[U];
                tri_spread_batch<U>(xl, A.d, k0, 8, nx, nb, wm, w2, t);
                if (w < nw) {
#pragma unroll
                    for (int q = 0; q < U; q++) if (k0 + q * 8 < np) tile[(size_t)(k0 + q * 8) * 16 + w] = t[q];
                }
            }
        }
        __syncthreads();
        // ---- phase 2: serial running sums, one thread per line
        if (tid < nw) {
            float *col = CONTIG ? tile + (size_t)tid * A.pitch : tile + tid;
            const int st = CONTIG ? 1 : 16;
            float s = 0.f;
            int k = 0;
            for (; k + 8 <= np; k += 8) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = col[(size_t)(k + q) * st];
#pragma unroll
                for (int q = 0; q < 8; q++) { s += v[q]; col[(size_t)(k + q) * st] = s; }
            }
            for (; k < np; k++) { s += col[(size_t)k * st]; col[(size_t)k * st] = s; }
            s = 0.f;
            k = np - 1;
            for (; k - 7 >= 0; k -= 8) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = col[(size_t)(k - q) * st];
#pragma unroll
                for (int q = 0; q < 8; q++) { s += v[q]; col[(size_t)(k - q) * st] = s; }
            }
            for (; k >= 0; k--) { s += col[(size_t)k * st]; col[(size_t)k * st] = s; }
        }
        __syncthreads();
        // ---- phase 3: fold + epilogue + store (epilogue operands loaded U4 elements ahead)
        constexpr int U4 = 4;
        if (CONTIG) {
            for (int w = 0; w < nw; w++) {
                const float *row = tile + (size_t)w * A.pitch;
                const long lb = base + (long)w * nx;
                for (int i0 = tid; i0 < nx; i0 += 128 * U4) {
                    EpiIn e[U4];
#pragma unroll
                    for (int q = 0; q < U4; q++) tri_epi_load<EPI>(A, lb + min(i0 + q * 128, nx - 1), e[q]);
#pragma unroll
                    for (int q = 0; q < U4; q++) {
                        const int i = i0 + q * 128;
                        if (i < nx) tri_epi_apply<EPI>(A, lb + i, tri_fold(row, 1, i, nx, nb), e[q], acc);
                    }
                }
            }
        } else {
            const int w = tid & 15, ir = tid >> 4;
            const int wc = min(w, nw - 1);
            const float *col = tile + wc;
            for (int i0 = ir; i0 < nx; i0 += 8 * U4) {
                EpiIn e[U4];
#pragma unroll
                for (int q = 0; q < U4; q++) tri_epi_load<EPI>(A, base + wc + (long)min(i0 + q * 8, nx - 1) * A.d, e[q]);
                if (w < nw) {
#pragma unroll
                    for (int q = 0; q < U4; q++) {
                        const int i = i0 + q * 8;
                        if (i < nx) tri_epi_apply<EPI>(A, base + w + (long)i * A.d, tri_fold(col, 16, i, nx, nb), e[q], acc);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (EPI != EPI_NONE) pst_block_reduce<3>(acc, A.partial);
}

// serial running sums over NG float4 groups spaced `gs` float4 apart (forward then backward);
// hand-scheduled so the next group's load and the previous group's store overlap the FADD chain.
// One readable pad group must exist before group 0 and after group ng-1.  Rows >= np in the
// last group hold t = 0 (forward) and are skipped by the backward sum.
__device__ __forceinline__ void tri_serial_v4(float4 *col, int gs, int np)
{
    const int ng = (np + 3) >> 2, nfull = np >> 2;
    float s = 0.f;
    {
        float4 cur = col[0], done = cur;
        for (int g = 0; g < ng; g++) {
            const float4 nxt = col[(size_t)(g + 1) * gs];
            s += cur.x; cur.x = s;
            if (g > 0) col[(size_t)(g - 1) * gs] = done;
            s += cur.y; cur.y = s;
            s += cur.z; cur.z = s;
            s += cur.w; cur.w = s;
            done = cur;
            cur = nxt;
        }
        col[(size_t)(ng - 1) * gs] = done;
    }
    s = 0.f;
    if (ng > nfull) {
        float4 v = col[(size_t)nfull * gs];
        const int rem = np & 3;
        if (rem > 2) { s += v.z; v.z = s; }
        if (rem > 1) { s += v.y; v.y = s; }
        s += v.x; v.x = s;
        col[(size_t)nfull * gs] = v;
    }
    if (nfull > 0) {
        float4 cur = col[(size_t)(nfull - 1) * gs], done = cur;
        for (int g = nfull - 1; g >= 0; g--) {
            const float4 nxt = col[((long)g - 1) * gs];
            s += cur.w; cur.w = s;
            if (g < nfull - 1) col[(size_t)(g + 1) * gs] = done;
            s += cur.z; cur.z = s;
            s += cur.y; cur.y = s;
            s += cur.x; cur.x = s;
            done = cur;
            cur = nxt;
        }
        col[0] = done;
    }
}

// ---- strided axes, 16-byte path (n1 % 4 == 0, 16-byte aligned volumes) -------------------
// Shared-memory tile of 16 lines in groups of 4 rows; one group = 64 floats + 4 pad (pitch 68:
// consecutive groups start 4 banks apart).  Layout A (row-major inside a group, elem(k,c) at
// 68(k>>2) + 16(k&3) + c) receives x via cp.async; layout B (k-minor, elem(k,c) at
// 68(k>>2) + 4c + (k&3)) holds t / F / B.  Both layouts keep a group in the same 64 floats, so
// t can be formed in place batch by batch in ascending k.  Lane mapping for the transposing
// phases: a quarter-warp = 8 consecutive groups of one 4-column block => every 128-bit shared
// access is conflict-free.
#define TRI_GP 68
template <int EPI>
__global__ void __launch_bounds__(128)
tri_tile_strided_v4_kernel(const TriArgs A)
{
    extern __shared__ __align__(16) float tile_raw[];
    float *const tile = tile_raw + TRI_GP;                    // one pad group in front
    const int nx = A.nx, nb = A.nb, np = nx + 2 * nb, tid = threadIdx.x;
    const int ng = (np + 3) >> 2;
    const float wm = -A.wt, w2 = A.w2;
    double acc[3] = {0.0, 0.0, 0.0};
    const long gpb = (A.na + 15) / 16;
    const int lane = tid & 31, wq = tid >> 5;
    const int tc4 = (lane >> 3) * 4, tgl = 8 * wq + (lane & 7);   // transposing phases: columns, group-in-batch
    for (long g0 = blockIdx.x; g0 < A.ngroups; g0 += gridDim.x) {
        const long ib = g0 / gpb, ia0 = (g0 % gpb) * 16;
        const int nw = (int)min((long)16, A.na - ia0);
        const long base = ia0 + ib * A.sb;
        // ---- phase 1a: x rows -> layout A at rows [2nb, 2nb+nx)
        {
            const int c4 = (tid & 3) * 4, r = tid >> 2;
            if (c4 < nw) {
                const float *xl = A.src + base + c4;
                for (int j = r; j < nx; j += 32) {
                    const int row = j + 2 * nb;
                    cp_async16(tile + (size_t)(row >> 2) * TRI_GP + (row & 3) * 16 + c4, xl + (long)j * A.d);
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();
        // ---- phase 1b: t_k in place, layout A -> layout B, 32 groups per batch
        for (int gb = 0; gb < ng; gb += 32) {
            const int g = gb + tgl;
            float4 t[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int k = 4 * g + e;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                const int ra = k + 2 * nb, rb = k + nb;
                const float4 xa = (k < nx) ? *reinterpret_cast<const float4 *>(tile + (size_t)(ra >> 2) * TRI_GP + (ra & 3) * 16 + tc4) : z;
                const float4 xb = (k >= nb && k - nb < nx) ? *reinterpret_cast<const float4 *>(tile + (size_t)(rb >> 2) * TRI_GP + (rb & 3) * 16 + tc4) : z;
                const float4 xc = (k >= 2 * nb && k - 2 * nb < nx) ? *reinterpret_cast<const float4 *>(tile + (size_t)g * TRI_GP + e * 16 + tc4) : z;
                t[e].x = tri_t(xa.x, xb.x, xc.x, wm, w2);
                t[e].y = tri_t(xa.y, xb.y, xc.y, wm, w2);
                t[e].z = tri_t(xa.z, xb.z, xc.z, wm, w2);
                t[e].w = tri_t(xa.w, xb.w, xc.w, wm, w2);
            }
            __syncthreads();
            if (g < ng) {
                float *q = tile + (size_t)g * TRI_GP + tc4 * 4;
                *reinterpret_cast<float4 *>(q + 0) = make_float4(t[0].x, t[1].x, t[2].x, t[3].x);
                *reinterpret_cast<float4 *>(q + 4) = make_float4(t[0].y, t[1].y, t[2].y, t[3].y);
                *reinterpret_cast<float4 *>(q + 8) = make_float4(t[0].z, t[1].z, t[2].z, t[3].z);
                *reinterpret_cast<float4 *>(q + 12) = make_float4(t[0].w, t[1].w, t[2].w, t[3].w);
            }
            __syncthreads();
        }
        // ---- phase 2
        if (tid < nw) tri_serial_v4(reinterpret_cast<float4 *>(tile) + tid, TRI_GP / 4, np);
        __syncthreads();
        // ---- phase 3: layout B -> rows of 4 columns, fold, epilogue, 16-byte stores
        {
            const bool live = tc4 < nw;
            const int cc = live ? tc4 : 0;
            auto rowB = [&](int k) {                          // scalar gather, only for reflections
                const float *q = tile + (size_t)(k >> 2) * TRI_GP + tc4 * 4 + (k & 3);
                return make_float4(q[0], q[4], q[8], q[12]);
            };
            for (int gb = 0; gb < ng; gb += 32) {
                const int g = gb + tgl;
                float4 cB[4];
                {
                    const float *q = tile + (size_t)min(g, ng - 1) * TRI_GP + tc4 * 4;
#pragma unroll
                    for (int e2 = 0; e2 < 4; e2++) cB[e2] = *reinterpret_cast<const float4 *>(q + 4 * e2);
                }
                const float4 rw[4] = {make_float4(cB[0].x, cB[1].x, cB[2].x, cB[3].x), make_float4(cB[0].y, cB[1].y, cB[2].y, cB[3].y),
                                      make_float4(cB[0].z, cB[1].z, cB[2].z, cB[3].z), make_float4(cB[0].w, cB[1].w, cB[2].w, cB[3].w)};
#pragma unroll
                for (int h = 0; h < 2; h++) {                 // two rows at a time: bounded registers
                    EpiIn4 ein[2];
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int i = min(max(4 * g + 2 * h + e - nb, 0), nx - 1);
                        tri_epi_load4<EPI>(A, base + cc + (long)i * A.d, ein[e]);
                    }
                    if (live) {
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int i = 4 * g + 2 * h + e - nb;
                            if (i >= 0 && i < nx) {
                                float4 v = rw[2 * h + e];
                                if (i >= nx - nb) {
                                    const float4 u = rowB(nb + nx + (nx - 1 - i));
                                    v.x = v.x + u.x; v.y = v.y + u.y; v.z = v.z + u.z; v.w = v.w + u.w;
                                }
                                if (i < nb) {
                                    const float4 u = rowB(nb - 1 - i);
                                    v.x = v.x + u.x; v.y = v.y + u.y; v.z = v.z + u.z; v.w = v.w + u.w;
                                }
                                tri_epi_apply4<EPI>(A, base + tc4 + (long)i * A.d, v, ein[e], acc);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    if (EPI != EPI_NONE) pst_block_reduce<3>(acc, A.partial);
}

// ---- contiguous axis, 16-byte path ---------------------------------------------------------
// unaligned 4-float read from shared memory as two aligned 16-byte loads + a uniform select
__device__ __forceinline__ float4 lds_f4_unaligned(const float *p)
{
    const int m = (int)((((uintptr_t)p) >> 2) & 3);
    const float4 *q = reinterpret_cast<const float4 *>(p - m);
    const float4 lo = q[0];
    if (m == 0) return lo;
    const float4 hi = q[1];
    if (m == 1) return make_float4(lo.y, lo.z, lo.w, hi.x);
    if (m == 2) return make_float4(lo.z, lo.w, hi.x, hi.y);
    return make_float4(lo.w, hi.x, hi.y, hi.z);
}

template <int EPI>
__global__ void __launch_bounds__(128)
tri_tile_contig_v4_kernel(const TriArgs A)
{
    extern __shared__ __align__(16) float tile_raw[];
    float *const tile = tile_raw + 4;                         // 4 floats of pad before row 0
    const int nx = A.nx, nb = A.nb, np = nx + 2 * nb, W = A.W, tid = threadIdx.x;
    const float wm = -A.wt, w2 = A.w2;
    double acc[3] = {0.0, 0.0, 0.0};
    const int lane = tid & 31, warp = tid >> 5;
    // x_j is parked at rowbase[xo + j] (16-byte aligned for cp.async); t_k / F_k / B_k live at
    // rowbase[sh + k] with sh = xo - 2nb, i.e. R[k] with R = rowbase + sh and x_k = R[k + 2nb].
    const int xo = (2 * nb + 3) & ~3, sh = xo - 2 * nb;
    for (long g = blockIdx.x; g < A.ngroups; g += gridDim.x) {
        const long l0 = g * W;
        const int nw = (int)min((long)W, A.nlines - l0);
        const long base = l0 * nx;
        // ---- phase 1: each warp owns lines w = warp, warp+4, ...
        for (int w = warp; w < nw; w += 4) {
            const float *xl = A.src + base + (long)w * nx;
            float *row = tile + (size_t)w * A.pitch;
            for (int j = 4 * lane; j < nx; j += 128) cp_async16(row + xo + j, xl + j);
        }
        cp_async_wait_all();
        __syncwarp();
        for (int w = warp; w < nw; w += 4) {
            float *R = tile + (size_t)w * A.pitch + sh;
            for (int kb = 0; kb < np; kb += 256) {            // ascending batches, reads before writes
                float va[8], vb[8], vc[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int k = kb + lane + 32 * q;
                    va[q] = (k < nx) ? R[k + 2 * nb] : 0.f;
                    vb[q] = (k >= nb && k - nb < nx) ? R[k + nb] : 0.f;
                    vc[q] = (k >= 2 * nb && k - 2 * nb < nx) ? R[k] : 0.f;
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int k = kb + lane + 32 * q;
                    if (k < np) R[k] = tri_t(va[q], vb[q], vc[q], wm, w2);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- phase 2: the running sums start at R[0]; peel to a 16-byte boundary
        if (tid < nw) {
            float *R = tile + (size_t)tid * A.pitch + sh;
            const int head = sh ? min(4 - sh, np) : 0;        // rowbase is 16-byte aligned
            float s = 0.f;
            for (int k = 0; k < head; k++) { s += R[k]; R[k] = s; }
            const int nq = (np - head) >> 2;
            float4 *r4 = reinterpret_cast<float4 *>(R + head);
            if (nq > 0) {
                float4 cur = r4[0], done = cur;
                for (int q = 0; q < nq; q++) {
                    const float4 nxt = r4[q + 1];
                    s += cur.x; cur.x = s;
                    if (q > 0) r4[q - 1] = done;
                    s += cur.y; cur.y = s;
                    s += cur.z; cur.z = s;
                    s += cur.w; cur.w = s;
                    done = cur;
                    cur = nxt;
                }
                r4[nq - 1] = done;
            }
            for (int k = head + 4 * nq; k < np; k++) { s += R[k]; R[k] = s; }
            s = 0.f;
            for (int k = np - 1; k >= head + 4 * nq; k--) { s += R[k]; R[k] = s; }
            if (nq > 0) {
                float4 cur = r4[nq - 1], done = cur;
                for (int q = nq - 1; q >= 0; q--) {
                    const float4 nxt = r4[q - 1];
                    s += cur.w; cur.w = s;
                    if (q < nq - 1) r4[q + 1] = done;
                    s += cur.z; cur.z = s;
                    s += cur.y; cur.y = s;
                    s += cur.x; cur.x = s;
                    done = cur;
                    cur = nxt;
                }
                r4[0] = done;
            }
            for (int k = head - 1; k >= 0; k--) { s += R[k]; R[k] = s; }
        }
        __syncthreads();
        // ---- phase 3: 4 outputs per thread, 16-byte global accesses
        for (int w = 0; w < nw; w++) {
            const float *R = tile + (size_t)w * A.pitch + sh;
            const long lb = base + (long)w * nx;
            for (int i0 = 4 * tid; i0 < nx; i0 += 4 * 128 * 2) {
                EpiIn4 ein[2];
#pragma unroll
                for (int q = 0; q < 2; q++) tri_epi_load4<EPI>(A, lb + min(i0 + 512 * q, nx - 4), ein[q]);
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int i = i0 + 512 * q;
                    if (i < nx) {
                        float4 v = lds_f4_unaligned(R + nb + i);
                        if (i + 3 >= nx - nb || i < nb) {      // reflections touch this quad
                            float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const int ii = i + e;
                                if (ii >= nx - nb) vv[e] = vv[e] + R[nb + nx + (nx - 1 - ii)];
                                if (ii < nb) vv[e] = vv[e] + R[nb - 1 - ii];
                            }
                            v = make_float4(vv[0], vv[1], vv[2], vv[3]);
                        }
                        tri_epi_apply4<EPI>(A, lb + i, v, ein[q], acc);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (EPI != EPI_NONE) pst_block_reduce<3>(acc, A.partial);
}

// ---------------------------------------------------------------------------------------
// G3/G4: streaming kernels of ps_divne and ps_conjgrad with fused double reductions.

// divne step 1 (:802-809): optional mask zeroing (dip3 :1656-1663), num,den *= 1/hypot(den,eps)
// in double; partial of sum(den^2) (:811).
__global__ void __launch_bounds__(256)
divne_prescale_kernel(float *__restrict__ num, float *__restrict__ den,
                      const unsigned char *__restrict__ mask, float eps, Span S,
                      double *__restrict__ partial)
{
    double acc[1] = {0.0};
    size_t i0, i1, step;
    pst_span(S, 1, i0, i1, step);
    for (size_t i = i0; i < i1; i += step) {
        float a = num[i], b = den[i];
        if (mask && mask[i]) { a = 0.f; b = 0.f; }
        if (eps > 0.0f) {
            const double norm = 1.0 / hypot((double)b, (double)eps);
            a = (float)((double)a * norm);
            b = (float)((double)b * norm);
        }
        num[i] = a;
        den[i] = b;
        acc[0] += (double)b * (double)b;
    }
    pst_block_reduce<1>(acc, partial);
}

// divne step 2 (:817-823) fused with the CG set-up (ps_conjgrad :279-297):
// w = den*norm; r = -(num*norm); p = x = 0; partial of r.r
__global__ void __launch_bounds__(256)
divne_scale_init_kernel(const float *__restrict__ num, float *__restrict__ den, double norm,
                        float *__restrict__ r, float *__restrict__ p, float *__restrict__ x,
                        Span S, double *__restrict__ partial)
{
    double acc[1] = {0.0};
    size_t i0, i1, step;
    pst_span(S, 1, i0, i1, step);
    for (size_t i = i0; i < i1; i += step) {
        const float a = (float)((double)num[i] * norm);
        den[i] = (float)((double)den[i] * norm);
        const float ri = -a;
        r[i] = ri;
        p[i] = 0.f;
        x[i] = 0.f;
        acc[0] += (double)ri * (double)ri;
    }
    pst_block_reduce<1>(acc, partial);
}

// CG step "update + gradient head":  (optional) p,x,r += a*s   (:372-374 of the previous
// iteration), then tmp = (-eps*x) + r*w   (:306-316: gx = -eps x; gx += L' r)
template <bool UPDATE>
__global__ void __launch_bounds__(256)
cg_head_kernel(float *__restrict__ p, float *__restrict__ x, float *__restrict__ r,
               const float *__restrict__ sp, const float *__restrict__ sx,
               const float *__restrict__ sr, const float *__restrict__ w, float a, float eps,
               float *__restrict__ tmp, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float xi = x[i], ri = r[i];
        if (UPDATE) {
            p[i] += a * sp[i];
            xi += a * sx[i];
            ri += a * sr[i];
            x[i] = xi;
            r[i] = ri;
        }
        float g = -eps * xi;
        g += ri * w[i];
        tmp[i] = g;
    }
}

// final model update only (after the last iteration / on early exit nothing is pending)
__global__ void __launch_bounds__(256)
cg_tail_kernel(float *__restrict__ x, const float *__restrict__ sx, float a, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] += a * sx[i];
}

// gp = eps*p + S(gx)  (:303,:318); partial gp.gp.  (The second shaping call reads gp out of place.)
__global__ void __launch_bounds__(256)
cg_gp_kernel(const float *__restrict__ p, const float *__restrict__ tmp, float *__restrict__ gp,
             float eps, Span S, double *__restrict__ partial)
{
    double acc[1] = {0.0};
    size_t i0, i1, step;
    pst_span(S, 1, i0, i1, step);
    for (size_t i = i0; i < i1; i += step) {
        float g = eps * p[i];
        g += tmp[i];
        gp[i] = g;
        acc[0] += (double)g * (double)g;
    }
    pst_block_reduce<1>(acc, partial);
}

// direction update (:333-361): gx = 0 + S(gp) (in tmp), gr = 0 + gx*w,
// s = g (first) or s = g + alpha*s; partials of sr.sr, sp.sp, sx.sx (:363)
template <bool FIRST>
__global__ void __launch_bounds__(256)
cg_dir_kernel(const float *__restrict__ gp, const float *__restrict__ tmp,
              const float *__restrict__ w, float *__restrict__ sp, float *__restrict__ sx,
              float *__restrict__ sr, float alpha, size_t n, double *__restrict__ partial)
{
    double acc[3] = {0.0, 0.0, 0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float gpi = gp[i];
        const float gxi = 0.f + tmp[i];
        const float gri = 0.f + gxi * w[i];
        float a, b, c;
        if (FIRST) { a = gpi; b = gxi; c = gri; }
        else {
            a = gpi + alpha * sp[i];
            b = gxi + alpha * sx[i];
            c = gri + alpha * sr[i];
        }
        sp[i] = a; sx[i] = b; sr[i] = c;
        acc[0] += (double)c * (double)c;
        acc[1] += (double)a * (double)a;
        acc[2] += (double)b * (double)b;
    }
    pst_block_reduce<3>(acc, partial);
}


// ---- 16-byte variants of the three CG streaming kernels (n % 4 == 0, 16-byte aligned vectors): one
// thread moves 4 consecutive samples per array per step, i.e. 4x the bytes in flight per thread.
// Elementwise arithmetic is unchanged; the double sums run over a different (still fixed) partition.
__device__ __forceinline__ float4 ld4(const float *p, size_t i) { return *reinterpret_cast<const float4 *>(p + i); }
__device__ __forceinline__ void st4(float *p, size_t i, float4 v) { *reinterpret_cast<float4 *>(p + i) = v; }

template <bool UPDATE>
__global__ void __launch_bounds__(256)
cg_head4_kernel(float *__restrict__ p, float *__restrict__ x, float *__restrict__ r,
                const float *__restrict__ sp, const float *__restrict__ sx,
                const float *__restrict__ sr, const float *__restrict__ w, float a, float eps,
                float *__restrict__ tmp, size_t n)
{
    for (size_t i = 4 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 4 * (size_t)gridDim.x * blockDim.x) {
        float4 xv = ld4(x, i), rv = ld4(r, i);
        const float4 wv = ld4(w, i);
        if (UPDATE) {
            float4 pv = ld4(p, i);
            const float4 spv = ld4(sp, i), sxv = ld4(sx, i), srv = ld4(sr, i);
            pv.x += a * spv.x; pv.y += a * spv.y; pv.z += a * spv.z; pv.w += a * spv.w;
            xv.x += a * sxv.x; xv.y += a * sxv.y; xv.z += a * sxv.z; xv.w += a * sxv.w;
            rv.x += a * srv.x; rv.y += a * srv.y; rv.z += a * srv.z; rv.w += a * srv.w;
            st4(p, i, pv); st4(x, i, xv); st4(r, i, rv);
        }
        float4 g;
        g.x = -eps * xv.x; g.x += rv.x * wv.x;
        g.y = -eps * xv.y; g.y += rv.y * wv.y;
        g.z = -eps * xv.z; g.z += rv.z * wv.z;
        g.w = -eps * xv.w; g.w += rv.w * wv.w;
        st4(tmp, i, g);
    }
}

__global__ void __launch_bounds__(256)
cg_gp4_kernel(const float *__restrict__ p, const float *__restrict__ tmp, float *__restrict__ gp,
              float eps, Span S, double *__restrict__ partial)
{
    double acc[1] = {0.0};
    size_t i0, i1, step;
    pst_span(S, 4, i0, i1, step);
    for (size_t i = i0; i < i1; i += step) {
        const float4 pv = ld4(p, i), tv = ld4(tmp, i);
        float4 g;
        g.x = eps * pv.x; g.x += tv.x;
        g.y = eps * pv.y; g.y += tv.y;
        g.z = eps * pv.z; g.z += tv.z;
        g.w = eps * pv.w; g.w += tv.w;
        st4(gp, i, g);
        acc[0] += (double)g.x * (double)g.x;
        acc[0] += (double)g.y * (double)g.y;
        acc[0] += (double)g.z * (double)g.z;
        acc[0] += (double)g.w * (double)g.w;
    }
    pst_block_reduce<1>(acc, partial);
}

template <bool FIRST>
__device__ __forceinline__ void cg_dir_one(float gpi, float t, float wi, float spi, float sxi, float sri, float alpha,
                                           float &a, float &b, float &c, double (&acc)[3])
{
    const float gxi = 0.f + t;
    const float gri = 0.f + gxi * wi;
    if (FIRST) { a = gpi; b = gxi; c = gri; }
    else { a = gpi + alpha * spi; b = gxi + alpha * sxi; c = gri + alpha * sri; }
    acc[0] += (double)c * (double)c;
    acc[1] += (double)a * (double)a;
    acc[2] += (double)b * (double)b;
}

template <bool FIRST>
__global__ void __launch_bounds__(256)
cg_dir4_kernel(const float *__restrict__ gp, const float *__restrict__ tmp,
               const float *__restrict__ w, float *__restrict__ sp, float *__restrict__ sx,
               float *__restrict__ sr, float alpha, size_t n, double *__restrict__ partial)
{
    double acc[3] = {0.0, 0.0, 0.0};
    for (size_t i = 4 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 4 * (size_t)gridDim.x * blockDim.x) {
        const float4 g = ld4(gp, i), t = ld4(tmp, i), wv = ld4(w, i);
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1, s3 = s1;
        if (!FIRST) { s1 = ld4(sp, i); s2 = ld4(sx, i); s3 = ld4(sr, i); }
        float4 a, b, c;
        cg_dir_one<FIRST>(g.x, t.x, wv.x, s1.x, s2.x, s3.x, alpha, a.x, b.x, c.x, acc);
        cg_dir_one<FIRST>(g.y, t.y, wv.y, s1.y, s2.y, s3.y, alpha, a.y, b.y, c.y, acc);
        cg_dir_one<FIRST>(g.z, t.z, wv.z, s1.z, s2.z, s3.z, alpha, a.z, b.z, c.z, acc);
        cg_dir_one<FIRST>(g.w, t.w, wv.w, s1.w, s2.w, s3.w, alpha, a.w, b.w, c.w, acc);
        st4(sp, i, a); st4(sx, i, b); st4(sr, i, c);
    }
    pst_block_reduce<3>(acc, partial);
}

// ---- device-resident CG scalars ---------------------------------------------------------------------------------
// The step lengths of ps_conjgrad (:330-371) are formed ON THE DEVICE from the reduced (and, across ranks, all-reduced)
// dot products, in the same double arithmetic the host used, and the vector kernels read them from this block: one CG
// solve is enqueued without a single host synchronisation.  The early exit (:349) raises `stop`; every later kernel of
// the solve returns at once, and the host -- which polls a pinned copy of the flag without blocking -- stops launching.
struct CgCtl {
    double gn, gnp, g0;
    float alpha_dir;       // gn / gnp of this iteration: s = g + alpha s
    float a_pending;       // -gn / beta of the last finished iteration: p, x, r += a s (applied by the next head / the tail)
    int stop, iters;
};

__global__ void cg_ctl_init_kernel(CgCtl *ctl) { ctl->gn = ctl->gnp = ctl->g0 = 0.; ctl->alpha_dir = ctl->a_pending = 0.f; ctl->stop = 0; ctl->iters = 0; }

// after gn = gp.gp is reduced (:322): g0 on the first iteration, else alpha, dg and the exit test (:331-349)
__global__ void cg_ctl_gn_kernel(CgCtl *ctl, const double *__restrict__ rec, int iter, float tol)
{
    if (ctl->stop) return;
    const double gn = rec[0];
    ctl->gn = gn;
    if (iter == 0) { ctl->g0 = gn; return; }
    const double alpha = gn / ctl->gnp, dg = gn / ctl->g0;
    if (alpha < tol || dg < tol) { ctl->stop = 1; return; }
    ctl->alpha_dir = (float)alpha;
}

// after sr.sr, sp.sp, sx.sx are reduced (:363-371): beta, the step length, gnp
__global__ void cg_ctl_beta_kernel(CgCtl *ctl, const double *__restrict__ rec, float eps)
{
    if (ctl->stop) return;
    const double beta = rec[0] + (double)eps * (rec[1] - rec[2]);
    const double alpha = -ctl->gn / beta;
    ctl->a_pending = (float)alpha;
    ctl->gnp = ctl->gn;
    ctl->iters++;
}

template <bool UPDATE>
__global__ void __launch_bounds__(256)
cg_head4d_kernel(float *__restrict__ p, float *__restrict__ x, float *__restrict__ r,
                 const float *__restrict__ sp, const float *__restrict__ sx,
                 const float *__restrict__ sr, const float *__restrict__ w, const CgCtl *__restrict__ ctl, float eps,
                 float *__restrict__ tmp, Span S)
{
    if (ctl->stop) return;
    const float a = ctl->a_pending;
    // block = one piece of one plane, like the kernels with sums: contiguous pieces stream faster than a grid-stride walk
    // (measured on gp / direction when they moved to pieces: 201 -> 186 and 577 -> 547 ms per step)
    size_t i0, i1, step;
    pst_span(S, 4, i0, i1, step);
    for (size_t i = i0; i < i1; i += step) {
        float4 xv = ld4(x, i), rv = ld4(r, i);
        const float4 wv = ld4(w, i);
        if (UPDATE) {
            float4 pv = ld4(p, i);
            const float4 spv = ld4(sp, i), sxv = ld4(sx, i), srv = ld4(sr, i);
            pv.x += a * spv.x; pv.y += a * spv.y; pv.z += a * spv.z; pv.w += a * spv.w;
            xv.x += a * sxv.x; xv.y += a * sxv.y; xv.z += a * sxv.z; xv.w += a * sxv.w;
            rv.x += a * srv.x; rv.y += a * srv.y; rv.z += a * srv.z; rv.w += a * srv.w;
            st4(p, i, pv); st4(x, i, xv); st4(r, i, rv);
        }
        float4 g;
        g.x = -eps * xv.x; g.x += rv.x * wv.x;
        g.y = -eps * xv.y; g.y += rv.y * wv.y;
        g.z = -eps * xv.z; g.z += rv.z * wv.z;
        g.w = -eps * xv.w; g.w += rv.w * wv.w;
        st4(tmp, i, g);
    }
}

// scalar fall-back of the same (n % 4 != 0 or unaligned vectors)
template <bool UPDATE>
__global__ void __launch_bounds__(256)
cg_headd_kernel(float *__restrict__ p, float *__restrict__ x, float *__restrict__ r,
                const float *__restrict__ sp, const float *__restrict__ sx,
                const float *__restrict__ sr, const float *__restrict__ w, const CgCtl *__restrict__ ctl, float eps,
                float *__restrict__ tmp, Span S)
{
    if (ctl->stop) return;
    const float a = ctl->a_pending;
    size_t i0, i1, step;
    pst_span(S, 1, i0, i1, step);
    for (size_t i = i0; i < i1; i += step) {
        float xi = x[i], ri = r[i];
        if (UPDATE) {
            p[i] += a * sp[i];
            xi += a * sx[i];
            ri += a * sr[i];
            x[i] = xi;
            r[i] = ri;
        }
        float g = -eps * xi;
        g += ri * w[i];
        tmp[i] = g;
    }
}

template <bool FIRST, bool VEC>
__global__ void __launch_bounds__(256)
cg_dird_kernel(const float *__restrict__ gp, const float *__restrict__ tmp,
               const float *__restrict__ w, float *__restrict__ sp, float *__restrict__ sx,
               float *__restrict__ sr, const CgCtl *__restrict__ ctl, Span S, double *__restrict__ partial)
{
    double acc[3] = {0.0, 0.0, 0.0};
    if (!ctl->stop) {
        const float alpha = ctl->alpha_dir;
        size_t i0, i1, step;
        pst_span(S, VEC ? 4 : 1, i0, i1, step);
        if (VEC) {
            for (size_t i = i0; i < i1; i += step) {
                const float4 g = ld4(gp, i), t = ld4(tmp, i), wv = ld4(w, i);
                float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1, s3 = s1;
                if (!FIRST) { s1 = ld4(sp, i); s2 = ld4(sx, i); s3 = ld4(sr, i); }
                float4 a, b, c;
                cg_dir_one<FIRST>(g.x, t.x, wv.x, s1.x, s2.x, s3.x, alpha, a.x, b.x, c.x, acc);
                cg_dir_one<FIRST>(g.y, t.y, wv.y, s1.y, s2.y, s3.y, alpha, a.y, b.y, c.y, acc);
                cg_dir_one<FIRST>(g.z, t.z, wv.z, s1.z, s2.z, s3.z, alpha, a.z, b.z, c.z, acc);
                cg_dir_one<FIRST>(g.w, t.w, wv.w, s1.w, s2.w, s3.w, alpha, a.w, b.w, c.w, acc);
                st4(sp, i, a); st4(sx, i, b); st4(sr, i, c);
            }
        } else {
            for (size_t i = i0; i < i1; i += step) {
                float a, b, c;
                cg_dir_one<FIRST>(gp[i], tmp[i], w[i], FIRST ? 0.f : sp[i], FIRST ? 0.f : sx[i], FIRST ? 0.f : sr[i], alpha, a, b, c, acc);
                sp[i] = a; sx[i] = b; sr[i] = c;
            }
        }
    }
    pst_block_reduce<3>(acc, partial);
}

__global__ void __launch_bounds__(256)
cg_taild_kernel(float *__restrict__ x, const float *__restrict__ sx, const CgCtl *__restrict__ ctl, size_t n)
{
    if (ctl->stop) return;                      // early exit: the last step was applied by the head of the stopped iteration
    const float a = ctl->a_pending;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] += a * sx[i];
}

__global__ void fill_kernel(float *__restrict__ x, float v, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = v;
}

// =======================================================================================
// host side
// =======================================================================================

// n3 = planes held by this rank (whole cube on a single GPU), n = n1*n2*n3.  In a distributed
// context the cube has n3g planes in total and this rank owns global planes [z0, z0+n3).
struct DipGeom {
    int n1, n2, n3, r1, r2, r3; size_t n;
    int n3g = 0, z0 = 0; double nglob = 0.0;
    bool dist = false;
    // work buffers of the distributed axis-3 pass (arena): halo planes before / after the slab,
    // carry planes (in / out), see smooth_axis3_dist
    float *hb = nullptr, *ha = nullptr, *cin = nullptr, *cout = nullptr;
    // neighbours' arenas mapped with CUDA IPC (equal slabs only: identical arena layouts), see smooth_axis3_dist
    char *peer_prev = nullptr, *peer_next = nullptr;
};
static DipGeom make_geom(int n1, int n2, int n3, int r1, int r2, int r3)
{
    DipGeom g;
    g.n1 = n1; g.n2 = n2; g.n3 = n3; g.r1 = r1; g.r2 = r2; g.r3 = r3; g.n = (size_t)n1 * n2 * n3;
    g.n3g = n3; g.z0 = 0; g.nglob = (double)g.n;
    return g;
}

static size_t tri_scratch_floats(const DipGeom &g)
{
    // largest extended volume of the strided-axis sweeps (+ axis 1 when it falls back)
    size_t a1 = (size_t)(g.n1 + 2 * (size_t)g.r1) * g.n2 * g.n3;
    size_t a2 = (size_t)g.n1 * (g.n2 + 2 * (size_t)g.r2) * g.n3;
    size_t a3 = (size_t)g.n1 * g.n2 * (g.n3 + 2 * (size_t)g.r3);
    size_t m = a1 > a2 ? a1 : a2;
    return m > a3 ? m : a3;
}

// Geometry of this rank's part of a cube with n3 GLOBAL planes: the whole cube on a single-GPU context, the
// rank's n3-slab (pst_ctx_slab rule) in a distributed one.  *scr = floats of smoothing scratch the geometry needs.
static int slab_geom(pst_ctx *c, const char *who, int n1, int n2, int n3, int r1, int r2, int r3, DipGeom *g, size_t *scr)
{
    const bool dist = c->comm != nullptr && c->nranks > 1;
    int z0 = 0, z1 = n3;
    if (dist) {
        z0 = (int)(((long)n3 * c->rank) / c->nranks);
        z1 = (int)(((long)n3 * (c->rank + 1)) / c->nranks);
        const int need_planes = std::max(2 * (r3 > 1 ? r3 : 0), 1);
        if ((n3 / c->nranks) < need_planes) {
            pst_set_error("%s: %d planes over %d ranks leaves slabs thinner than 2*r3 = %d", who, n3, c->nranks, 2 * r3);
            return PST_EUNSUP;
        }
    }
    *g = make_geom(n1, n2, z1 - z0, r1, r2, r3);
    g->n3g = n3; g->z0 = z0; g->nglob = (double)n1 * n2 * n3; g->dist = dist;
    *scr = tri_scratch_floats(*g);
    if (dist) *scr = std::max(*scr, (size_t)(z1 - z0 + 2 * r3) * (size_t)n1 * n2);
    return PST_OK;
}

int pst_comm_map_arenas(pst_ctx *c, char **prev, char **next);                       // pst_comm.cu

// floats slab_halos() takes from the arena
static size_t slab_halo_floats(const DipGeom &g)
{
    return g.dist ? (size_t)g.n1 * g.n2 * (size_t)(2 * std::max(g.r3, 1) + 1 + (g.n3 / 128 + 3)) + 4 * 64 : 0;
}

// work planes of the distributed axis-3 pass (smooth_axis3_dist): r3-plane halos either side + carry planes
static int slab_halos(pst_ctx *c, DipGeom *g)
{
    if (!g->dist) return PST_OK;
    const size_t plane = (size_t)g->n1 * g->n2;
    PST_TRY(pst_arena_get(c, plane * (size_t)std::max(g->r3, 1), &g->hb));
    PST_TRY(pst_arena_get(c, plane * (size_t)std::max(g->r3, 1), &g->ha));
    PST_TRY(pst_arena_get(c, plane * (size_t)(g->n3 / 128 + 3), &g->cin));   // carry plane in / the running sum before every chunk (register kernels)
    PST_TRY(pst_arena_get(c, plane, &g->cout));
    if (g->r3 > 1 && g->n3g % c->nranks == 0) PST_TRY(pst_comm_map_arenas(c, &g->peer_prev, &g->peer_next));   // collective
    return PST_OK;
}

static int tri_lines_launch(pst_ctx *c, int cls, float *x, float *scr, long nlines, long na, long sa, long sb,
                            long d, int nx, int nb)
{
    const float wt = (float)(1.0 / ((double)nb * nb));       // ps_triangle_init :421
    const float w2 = (float)(2. * wt);
    const int threads = 128;
    const long blocks = (nlines + threads - 1) / threads;
    PST_LAUNCHB(c, cls, 8.0 * (double)nlines * nx,
        if (nb <= nx)
            tri_lines_kernel<<<(unsigned)blocks, threads, 0, c->stream>>>(x, scr, nlines, na, sa, sb, d, nx, nb, wt, w2);
        else
            tri_lines_literal_kernel<<<(unsigned)blocks, threads, 0, c->stream>>>(x, scr, nlines, na, sa, sb, d, nx, nb, wt, w2));
    c->stats.smooth_passes++;
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

// ---------------------------------------------------------------------------------------
// Distributed axis-3 smoothing (n3-slabs over ranks): the two running sums run ACROSS ranks in the
// reference's order, so results stay bit-identical to the single-GPU / reference result.  Rank r
// owns the extended positions k in [K0, K1) of every line (K0 = z0 + nb, or 0 on rank 0;
// K1 = z1 + nb, or n3g + 2nb on the last rank).  The running value at the slab edge is handed to
// the neighbour as a carry plane; lines are processed in chunks so that ranks work concurrently
// (software pipeline over chunks).  x at global plane j comes from the slab, or from the nb-plane
// halos received before the pass.
struct Tri3Args {
    const float *x;        // slab [nz][L]
    const float *hb, *ha;  // halos: planes [z0-nb, z0) and [z1, z1+nb)
    float *F;              // scratch [K1-K0][L]
    // carry pipeline: incoming carries + per-CTA flag in MY mailbox (nullptr = edge rank, carry 0);
    // outgoing carries + flag in the NEIGHBOUR's mailbox (peer memory over NVLink; nullptr = edge)
    const float *cin; const unsigned *fin;
    float *cout; unsigned *fout;
    const uint2 *pin; uint2 *pout;   // tile kernels: {carry, epoch} pairs (mine, incoming) / (neighbour's, outgoing)
    uint2 *pself;                    // register kernels, last rank: its own (otherwise unused) backward mailbox, see tri3_reg_bwd_kernel
    unsigned *err; unsigned epoch;
    // peer-memory halos: hb / ha point INTO the neighbours' slabs (CUDA IPC); the kernel first waits for the flags the
    // neighbours raise (in my mailbox) once their current input is complete.  Null = halos were copied by NCCL.
    const unsigned *hr_prev, *hr_next;
    // recompute variant of the tile kernels: the forward kernel stores no F, only each line's incoming carry (csave) and
    // -- with peer halos -- a local copy of the planes it read from the NEXT rank (ha_keep): that rank overwrites them
    // in its backward kernel, which runs before mine
    float *csave, *ha_keep;
    float *dst;            // fold output [nz][L]
    long L, l0, l1;        // lines per plane, chunk [l0, l1)
    int n3g, z0, nz, nb, K0, K1;
    float wt, w2;
    int rev;               // tile kernels, backward: visit the tiles in descending order (the F tiles written last are still in L2)
};

__device__ __forceinline__ float tri3_x(const Tri3Args &A, int j, long l)
{
    if (j >= A.z0 && j < A.z0 + A.nz) return A.x[(long)(j - A.z0) * A.L + l];
    if (j < A.z0) return A.hb[(long)(j - (A.z0 - A.nb)) * A.L + l];
    return A.ha[(long)(j - (A.z0 + A.nz)) * A.L + l];
}

// Carry hand-off of the distributed axis-3 kernels: every line's carry travels as ONE 8-byte store {carry bits, epoch}
// into the neighbour's mailbox and is polled by the consuming thread itself (the same trick as NCCL's
// LL protocol: an aligned 8-byte store is delivered atomically, so the epoch validates the payload).
// No system-scope fence, no separate flag: measured, a fence + flag per CTA doubled the kernel time.
__device__ __forceinline__ float tri3_pair_recv(const uint2 *p, unsigned epoch, unsigned *err)
{
    const long long t0 = clock64();
    unsigned x, y;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "l"(p) : "memory");
        if (y == epoch) break;
        if (clock64() - t0 > 20000000000LL) { *err = 1u; break; }     // ~10 s
    }
    return __uint_as_float(x);
}
__device__ __forceinline__ void tri3_pair_send(uint2 *p, float v, unsigned epoch)
{
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}

// peer-memory halos: one thread waits until both neighbours have announced that their current input is complete
__device__ __forceinline__ void tri3_wait_halos(const Tri3Args &A)
{
    if (A.hr_prev == nullptr && A.hr_next == nullptr) return;        // uniform
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        const unsigned *fl[2] = {A.hr_prev, A.hr_next};
        for (int q = 0; q < 2; q++) {
            if (!fl[q]) continue;
            unsigned v;
            for (;;) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(fl[q]) : "memory");
                if (v == A.epoch) break;
                if (clock64() - t0 > 20000000000LL) { *A.err = 1u; break; }
                __nanosleep(100);
            }
        }
    }
    __syncthreads();
}
// raised on the stream right after the kernel that produced the input of an axis-3 pass
__global__ void tri3_halo_ready_kernel(unsigned *to_prev, unsigned *to_next, unsigned epoch)
{
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (to_prev) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(to_prev), "r"(epoch) : "memory");
        if (to_next) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(to_next), "r"(epoch) : "memory");
    }
}

// line kernels (tall slabs): one flag per CTA (one polling thread); measured at 2 GPUs the per-thread pair
// polling of the tile kernels costs more here (303K resident pollers per GPU)
// consumer side of the carry hand-off: one thread spins (acquire, system scope) on this CTA's
// flag, with a generous timeout so that a failed neighbour cannot hang the GPU
__device__ __forceinline__ void tri3_wait(const unsigned *flag, unsigned epoch, unsigned *err)
{
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v == epoch) break;
            if (clock64() - t0 > 20000000000LL) { *err = 1u; break; }     // ~10 s
            __nanosleep(200);
        }
    }
    __syncthreads();
}
// producer side: carries were stored to the neighbour's mailbox by all threads
__device__ __forceinline__ void tri3_signal(unsigned *flag, unsigned epoch)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}

__global__ void __launch_bounds__(128)
tri3_dist_fwd_kernel(const Tri3Args A)
{
    const long l = (long)blockIdx.x * blockDim.x + threadIdx.x;
    tri3_wait_halos(A);
    if (A.cin) tri3_wait(A.fin + blockIdx.x, A.epoch, A.err);
    const bool live = l < A.L;
    const float wm = -A.wt;
    float s = 0.f;
    if (live) {
        if (A.cin) s = __ldcv(A.cin + l);
        // loads of U steps are issued together, ahead of the dependent FADD chain: the latency of a
        // CTA is what every downstream rank waits for (pipeline fill), so it must be short
        constexpr int U = 8;
        for (int k0 = A.K0; k0 < A.K1; k0 += U) {
            float va[U], vb[U], vc[U];
#pragma unroll
            for (int q = 0; q < U; q++) {
                const int k = k0 + q;
                const bool in = k < A.K1;
                va[q] = (in && k < A.n3g) ? tri3_x(A, k, l) : 0.f;
                vb[q] = (in && k >= A.nb && k - A.nb < A.n3g) ? tri3_x(A, k - A.nb, l) : 0.f;
                vc[q] = (in && k >= 2 * A.nb && k - 2 * A.nb < A.n3g) ? tri3_x(A, k - 2 * A.nb, l) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < U; q++) {
                const int k = k0 + q;
                if (k < A.K1) {
                    float t = 0.f;
                    t = t + wm * va[q];
                    t = t + A.w2 * vb[q];
                    t = t + wm * vc[q];
                    s += t;
                    A.F[(long)(k - A.K0) * A.L + l] = s;
                }
            }
        }
        if (A.cout) A.cout[l] = s;
    }
    if (A.cout) tri3_signal(A.fout + blockIdx.x, A.epoch);
}

// The same forward kernel with the (-1, 2, -1) window of every line held in registers (radius as a template parameter):
// one global load per step instead of three -- the two delayed taps of the kernel above are re-reads of planes that are
// nb and 2 nb planes (tens of MB) behind, i.e. mostly HBM traffic again (measured at 2 GPUs, 512-plane slabs: 2.8 ms per
// pass pair = 3.0 TB/s of the 16 B/voxel the pair should move).
template <int NB>
__global__ void __launch_bounds__(128)
tri3_dist_fwd_win_kernel(const Tri3Args A)
{
    const long l = (long)blockIdx.x * blockDim.x + threadIdx.x;
    tri3_wait_halos(A);
    if (A.cin) tri3_wait(A.fin + blockIdx.x, A.epoch, A.err);
    const bool live = l < A.L;
    const float wm = -A.wt;
    float s = 0.f;
    if (live) {
        if (A.cin) s = __ldcv(A.cin + l);
        constexpr int U = 8;
        float win[2 * NB + U];                                   // x of planes k0 - 2nb .. k0 + U - 1 (0 outside the cube)
#pragma unroll
        for (int i = 0; i < 2 * NB; i++) {
            const int j = A.K0 - 2 * NB + i;
            win[i] = (j >= 0 && j < A.n3g) ? tri3_x(A, j, l) : 0.f;
        }
        float *Fo = A.F + l;
        for (int k0 = A.K0; k0 < A.K1; k0 += U) {
#pragma unroll
            for (int q = 0; q < U; q++) {
                const int k = k0 + q;
                win[2 * NB + q] = (k < A.K1 && k < A.n3g) ? tri3_x(A, k, l) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < U; q++) {
                const int k = k0 + q;
                if (k < A.K1) {
                    float t = 0.f;
                    t = t + wm * win[2 * NB + q];
                    t = t + A.w2 * win[NB + q];
                    t = t + wm * win[q];
                    s += t;
                    Fo[(long)(k - A.K0) * A.L] = s;
                }
            }
#pragma unroll
            for (int i = 0; i < 2 * NB; i++) win[i] = win[i + U];
        }
        if (A.cout) A.cout[l] = s;
    }
    if (A.cout) tri3_signal(A.fout + blockIdx.x, A.epoch);
}

// backward sum fused with fold2 (:458-484): y_i = (B_{i+nb} + B_{nb+n3g+(n3g-1-i)}[i >= n3g-nb])
// + B_{nb-1-i}[i < nb].  Every term lives on the owning rank (slab height >= 2nb): the right-pad
// values (last rank, visited first) are parked in their target rows, the left-pad values
// (rank 0, visited last) are added in place.
__global__ void __launch_bounds__(128)
tri3_dist_bwd_kernel(const Tri3Args A)
{
    const long l = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (A.cin) tri3_wait(A.fin + blockIdx.x, A.epoch, A.err);
    const bool live = l < A.L;
    const int nb = A.nb, n3g = A.n3g;
    if (live) {
        float s = A.cin ? __ldcv(A.cin + l) : 0.f;
        int k = A.K1 - 1;
        for (; k >= nb + n3g && k >= A.K0; k--) {               // right pad (last rank only)
            s += A.F[(long)(k - A.K0) * A.L + l];
            const int gi = n3g - 1 - (k - nb - n3g);
            A.dst[(long)(gi - A.z0) * A.L + l] = s;
        }
        const int kmid = max(nb, A.K0);                        // middle: global plane gi = k - nb
        constexpr int U = 8;
        for (; k - (U - 1) >= kmid; k -= U) {
            float f[U];
#pragma unroll
            for (int q = 0; q < U; q++) f[q] = A.F[(long)(k - q - A.K0) * A.L + l];
#pragma unroll
            for (int q = 0; q < U; q++) {
                s += f[q];
                const int gi = k - q - nb;
                float v = s;
                if (gi >= n3g - nb) v = v + A.dst[(long)(gi - A.z0) * A.L + l];
                A.dst[(long)(gi - A.z0) * A.L + l] = v;
            }
        }
        for (; k >= kmid; k--) {
            s += A.F[(long)(k - A.K0) * A.L + l];
            const int gi = k - nb;
            float v = s;
            if (gi >= n3g - nb) v = v + A.dst[(long)(gi - A.z0) * A.L + l];
            A.dst[(long)(gi - A.z0) * A.L + l] = v;
        }
        for (; k >= A.K0; k--) {                                // left pad (rank 0 only)
            s += A.F[(long)(k - A.K0) * A.L + l];
            A.dst[(long)(nb - 1 - k - A.z0) * A.L + l] += s;
        }
        if (A.cout) A.cout[l] = s;
    }
    if (A.cout) tri3_signal(A.fout + blockIdx.x, A.epoch);
}


// ---- distributed axis 3, tile kernels -------------------------------------------------------
// What every downstream rank waits for is the LATENCY of one CTA of the upstream rank (7 hops at 8
// GPUs), so a CTA must be short: it first pulls its whole tile (W lines x all local rows) into
// shared memory with one burst of 16-byte cp.async (all requests in flight at once, issued BEFORE
// the carry flag is awaited), then each line's running sum is a pure shared-memory/FADD loop
// (~10 instructions per sample) instead of a chain of dependent global loads.
//   forward : rows = x planes [K0-2nb, K1) (zero outside the cube) -> F rows [K0, K1) to scratch
//   backward: rows = F rows [K0, K1) -> fold2 -> dst; parked reflections live in the consumed rows
template <int W, bool STOREF>
__global__ void __launch_bounds__(128)
tri3_tile_fwd_kernel(const Tri3Args A)
{
    extern __shared__ __align__(16) float t3s[];
    const int nb = A.nb, R = A.K1 - A.K0 + 2 * nb, tid = threadIdx.x;
    const long l0 = (long)blockIdx.x * W;
    const long l = l0 + tid;
    const bool live = tid < W && l < A.L;
    // ---- burst load: row r <-> global plane j = K0 - 2nb + r.  The rows of my own slab go first; the flags of the
    // neighbours' planes ("my input is complete") are awaited while those loads are in flight, then the halo rows follow.
    constexpr int CPR = W / 4;                                 // 16-byte chunks per row
    float s = 0.f;
    {
        auto rows = [&](bool halo) {
            for (int idx = tid; idx < R * CPR; idx += blockDim.x) {
                const int r = idx / CPR, ch = idx - r * CPR;
                const int j = A.K0 - 2 * nb + r;
                const long lc = l0 + 4 * ch;
                float *dsts = t3s + (size_t)r * W + 4 * ch;
                const bool zero = j < 0 || j >= A.n3g || lc >= A.L;
                const bool mine = j >= A.z0 && j < A.z0 + A.nz;
                if (halo ? (zero || mine) : !(zero || mine)) continue;
                if (zero) {
                    *reinterpret_cast<float4 *>(dsts) = make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    const float *srcp;
                    if (mine) srcp = A.x + (long)(j - A.z0) * A.L + lc;
                    else if (j < A.z0) srcp = A.hb + (long)(j - (A.z0 - nb)) * A.L + lc;
                    else srcp = A.ha + (long)(j - (A.z0 + A.nz)) * A.L + lc;
                    cp_async16(dsts, srcp);
                }
            }
        };
        rows(false);
        asm volatile("cp.async.commit_group;" ::: "memory");
        tri3_wait_halos(A);
        rows(true);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (!STOREF && A.ha_keep) {
        // keep the planes read from the next rank (rows of planes [z1, K1)) for my backward kernel
        const int r0 = A.z0 + A.nz - (A.K0 - 2 * nb);
        for (int idx = tid; idx < (R - r0) * CPR; idx += blockDim.x) {
            const int r = r0 + idx / CPR, ch = idx % CPR;
            const long lc = l0 + 4 * ch;
            if (lc < A.L) *reinterpret_cast<float4 *>(A.ha_keep + (long)(r - r0) * A.L + lc) = *reinterpret_cast<const float4 *>(t3s + (size_t)r * W + 4 * ch);
        }
        __syncthreads();                                       // (the stencil values below overwrite rows in place)
    }
    if (live) {
        const float wm = -A.wt, w2 = A.w2;
        float *xc = t3s + tid;                                 // x_{k-2nb} of step k = K0 + i at row i
        float *Fo = A.F + l;
        const int n = A.K1 - A.K0;
        // every downstream rank waits for the latency of this CTA AFTER its carry has arrived: the stencil values
        // t_k = ((0 + wm x_k) + w2 x_{k-nb}) + wm x_{k-2nb} do not depend on the carry and are formed first (in place: row i
        // is read by steps i, i - nb and i - 2nb only), so that one addition per sample is left behind the wait
#pragma unroll 4
        for (int i = 0; i < n; i++) {
            float t = wm * xc[(size_t)(i + 2 * nb) * W];
            t = t + w2 * xc[(size_t)(i + nb) * W];
            t = t + wm * xc[(size_t)i * W];
            xc[(size_t)i * W] = t;
        }
        if (A.pin) s = tri3_pair_recv(A.pin + l, A.epoch, A.err);
        if (!STOREF) A.csave[l] = s;
#pragma unroll 8
        for (int i = 0; i < n; i++) {
            s += xc[(size_t)i * W];
            if (STOREF) Fo[(long)i * A.L] = s;
        }
        if (A.pout) tri3_pair_send(A.pout + l, s, A.epoch);
    }
}

// RCMP: the tile holds x (rows of planes [K0 - 2nb, K1), like the forward kernel's) and F is recomputed in place from the
// saved incoming carry -- the same additions in the same order -- before the backward carry is needed: the pass moves
// 4 (forward read) + 4 + 4 bytes per voxel instead of 16, and the forward kernel stores nothing.
template <int W, bool RCMP>
__global__ void __launch_bounds__(128)
tri3_tile_bwd_kernel(const Tri3Args A)
{
    extern __shared__ __align__(16) float t3s[];
    const int nb = A.nb, n3g = A.n3g, R = A.K1 - A.K0, tid = threadIdx.x;
    const long l0 = (long)(A.rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * W;
    const long l = l0 + tid;
    const bool live = tid < W && l < A.L;
    constexpr int CPR = W / 4;
    float s = 0.f;
    if (!RCMP) {
        for (int idx = tid; idx < R * CPR; idx += blockDim.x) {
            const int r = idx / CPR, ch = idx - r * CPR;
            const long lc = l0 + 4 * ch;
            if (lc < A.L) cp_async16(t3s + (size_t)r * W + 4 * ch, A.F + (long)r * A.L + lc);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (A.pin && live) s = tri3_pair_recv(A.pin + l, A.epoch, A.err);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    } else {
        // (no halo flags to wait for: the forward kernel of this pass already did, and the planes of the previous rank
        // stay untouched until its backward kernel, which waits for my carries)
        const int RX = R + 2 * nb;
        for (int idx = tid; idx < RX * CPR; idx += blockDim.x) {
            const int r = idx / CPR, ch = idx - r * CPR;
            const int j = A.K0 - 2 * nb + r;
            const long lc = l0 + 4 * ch;
            float *dsts = t3s + (size_t)r * W + 4 * ch;
            if (j < 0 || j >= n3g || lc >= A.L) {
                *reinterpret_cast<float4 *>(dsts) = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                const float *srcp;
                if (j >= A.z0 && j < A.z0 + A.nz) srcp = A.x + (long)(j - A.z0) * A.L + lc;
                else if (j < A.z0) srcp = A.hb + (long)(j - (A.z0 - nb)) * A.L + lc;
                else srcp = (A.ha_keep ? A.ha_keep : A.ha) + (long)(j - (A.z0 + A.nz)) * A.L + lc;
                cp_async16(dsts, srcp);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        float sf = live ? A.csave[l] : 0.f;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (live) {
            const float wm = -A.wt, w2 = A.w2;
            float *xc = t3s + tid;                             // x_{k-2nb} of step k = K0 + i at row i; F_k replaces it
#pragma unroll 4
            for (int i = 0; i < R; i++) {
                float t = wm * xc[(size_t)(i + 2 * nb) * W];
                t = t + w2 * xc[(size_t)(i + nb) * W];
                t = t + wm * xc[(size_t)i * W];
                sf += t;
                xc[(size_t)i * W] = sf;
            }
            if (A.pin) s = tri3_pair_recv(A.pin + l, A.epoch, A.err);
        }
    }
    if (live) {
        float *Fc = t3s + tid;                                 // row (k - K0); consumed rows are reused as parking slots
        float *dl = A.dst + l;
        int k = A.K1 - 1;
        for (; k >= nb + n3g && k >= A.K0; k--) {               // right pad (last rank): park B_k in its own row
            s += Fc[(size_t)(k - A.K0) * W];
            Fc[(size_t)(k - A.K0) * W] = s;
        }
        const int kmid = max(nb, A.K0);
#pragma unroll 4
        for (; k >= kmid; k--) {                               // samples gi = k - nb
            s += Fc[(size_t)(k - A.K0) * W];
            const int gi = k - nb;
            float v = s;
            if (gi >= n3g - nb) v = v + Fc[(size_t)(nb + n3g + (n3g - 1 - gi) - A.K0) * W];
            if (gi < nb) Fc[(size_t)(k - A.K0) * W] = v;        // head: completed by the left pad (rank 0)
            else dl[(long)(gi - A.z0) * A.L] = v;
        }
        for (; k >= A.K0; k--) {                                // left pad (rank 0): y_i = head_i + B_{nb-1-i}
            s += Fc[(size_t)(k - A.K0) * W];
            const int gi = nb - 1 - k;
            dl[(long)(gi - A.z0) * A.L] = Fc[(size_t)(gi + nb - A.K0) * W] + s;
        }
        if (A.pout) tri3_pair_send(A.pout + l, s, A.epoch);
    }
}

// ---- distributed axis 3, register kernels (equal slabs of v x NZ planes: 2, 4, 8 ranks at n3 = 1024) --------------------
// One thread owns one line and walks the line's local rows in chunks of NZ planes held in REGISTERS: no shared memory, no
// block synchronisation after the halo flags, ~140 independent coalesced loads in flight per thread and chunk, 11 (forward)
// and 18 (backward) instructions per sample; the pass moves 12 B per voxel instead of 16 (F is recomputed, not staged).
// The per-line arithmetic is pst_tri3_reg_core.h (host-tested); here: the mailboxes and the launch.
struct Tri3RegIO {
    const Tri3Args &A;
    __device__ __forceinline__ void wait_halos() { tri3_wait_halos(A); }
    __device__ __forceinline__ float recv(long l) { return tri3_pair_recv(A.pin + l, A.epoch, A.err); }
};

template <int NB, int NZ>
__global__ void __launch_bounds__(64)
tri3_reg_fwd_kernel(const Tri3Args A)
{
    const long l = (long)blockIdx.x * 64 + threadIdx.x;
    Tri3RegIO io{A};
    float s;
    if (!tri3_reg::fwd_line<NB, NZ>(A, io, l, l < A.L, &s)) return;
    if (A.K1 != A.n3g + 2 * NB) tri3_pair_send(A.pout + l, s, A.epoch);
    else tri3_pair_send(A.pself + l, 0.f, A.epoch);            // the last rank's backward carry: +0, through its own mailbox
}

template <int NB, int NZ>
__global__ void __launch_bounds__(64)
tri3_reg_bwd_kernel(const Tri3Args A)
{
    const long l = (long)(A.rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * 64 + threadIdx.x;
    if (l >= A.L) return;
    Tri3RegIO io{A};
    const float s = tri3_reg::bwd_line<NB, NZ>(A, io, l);
    if (A.K0 != 0) tri3_pair_send(A.pout + l, s, A.epoch);
}

// register kernels: radii and chunk heights instantiated; equal slabs of whole chunks
static int tri3_reg_chunk(int nb, int n3g, int nranks)
{
    static const bool on = []() { const char *e = getenv("PST_TRI3_REG"); return !(e && e[0] == '0'); }();
    if (!on || nranks < 2 || n3g % nranks != 0) return 0;
    if (!(nb == 2 || nb == 3 || nb == 4 || nb == 5 || nb == 6 || nb == 8)) return 0;
    const int nz = n3g / nranks;
    const int NZ = nz % 128 == 0 ? 128 : ((nz % 32 == 0 && nz <= 96) ? 32 : 0);
    return (NZ && NZ >= 2 * nb) ? NZ : 0;
}
static bool tri3_reg_ok(int nb, int n3g, int nranks) { return tri3_reg_chunk(nb, n3g, nranks) != 0; }
template <int NB>
static void tri3_reg_launch_nb(const Tri3Args &A, bool fwd, int NZ, unsigned blocks, cudaStream_t st)
{
    if (fwd) {
        if (NZ == 128) tri3_reg_fwd_kernel<NB, 128><<<blocks, 64, 0, st>>>(A);
        else tri3_reg_fwd_kernel<NB, 32><<<blocks, 64, 0, st>>>(A);
    } else {
        if (NZ == 128) tri3_reg_bwd_kernel<NB, 128><<<blocks, 64, 0, st>>>(A);
        else tri3_reg_bwd_kernel<NB, 32><<<blocks, 64, 0, st>>>(A);
    }
}
// the forward kernel reads 4 B per voxel, the backward kernel reads 4 and writes 4
static int tri3_reg_launch(pst_ctx *c, const Tri3Args &A, bool fwd, int NZ, size_t nvox, int cls)
{
    const unsigned blocks = (unsigned)((A.L + 63) / 64);
    PST_LAUNCHB(c, cls, (fwd ? 4.0 : 8.0) * (double)nvox,
        switch (A.nb) {
            case 2: tri3_reg_launch_nb<2>(A, fwd, NZ, blocks, c->stream); break;
            case 3: tri3_reg_launch_nb<3>(A, fwd, NZ, blocks, c->stream); break;
            case 4: tri3_reg_launch_nb<4>(A, fwd, NZ, blocks, c->stream); break;
            case 5: tri3_reg_launch_nb<5>(A, fwd, NZ, blocks, c->stream); break;
            case 6: tri3_reg_launch_nb<6>(A, fwd, NZ, blocks, c->stream); break;
            default: tri3_reg_launch_nb<8>(A, fwd, NZ, blocks, c->stream); break;
        });
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

// widest tile (lines per CTA) whose rows fit in shared memory; 0 = use the line kernels
static int tri3_tile_width(int rows_max, long L, const void *a, const void *b, const void *c2, const void *d, const void *e)
{
    auto a16 = [](const void *q) { return (((uintptr_t)q) & 15) == 0; };
    if (L % 4 != 0 || !a16(a) || !a16(b) || !a16(c2) || !a16(d) || !a16(e)) return 0;
    static const bool on = []() { const char *v = getenv("PST_TRI3_TILE"); return !(v && v[0] == '0'); }();
    if (!on) return 0;
    // narrower tiles leave too few chain threads per SM: measured at 2 GPUs (527 rows, W = 32) the tile
    // kernels take 2.6 ms per launch against 1.4 ms for the line kernels, so tall slabs keep the latter
    // widest tile tried: 64 lines (measured on the 128-plane slab, 128 threads: 0.50 / 0.44 / 0.42 ms per pass at 128 / 64 / 32
    // lines; 32-line tiles lose on tall slabs).  PST_TRI3_WMAX overrides.
    static const int wmax = []() { const char *v = getenv("PST_TRI3_WMAX"); return v ? atoi(v) : 64; }();
    const int ws[3] = {128, 64, 32};                          // 32 only on request (PST_TRI3_WMAX=32)
    for (int w : ws) if (w <= wmax && (size_t)rows_max * w * 4 <= 75 * 1024) return w;
    return 0;
}

// ---- shaping operator driver ------------------------------------------------------------
struct EpiSpec {
    int kind = EPI_NONE;
    const float *p = nullptr, *w = nullptr;
    float *gp = nullptr, *sp = nullptr, *sx = nullptr, *sr = nullptr;
    float eps = 0.f, alpha = 0.f;
    int rec = -1;                 // reduction record receiving the fused sums
    bool stream_only = false;     // fuse only if the streaming kernel takes it (else: plain pass, *fused = false)
};
// called right before the last axis is launched: lets the caller fetch scalars the epilogue
// needs (alpha) and veto the launch (CG early exit).  Return <0 error, 0 go on, 1 skip.
typedef int (*late_bind_fn)(void *user, EpiSpec *epi);

struct TilePlan { bool ok; int W, pitch; size_t smem; int ctas_per_sm; };

static TilePlan tile_plan(bool contig, bool vec, int nx, int nb)
{
    TilePlan t{false, 0, 0, 0, 1};
    if (nb > nx) return t;                       // multiple reflections: literal fallback
    const int np = nx + 2 * nb;
    const size_t soft = 74 * 1024, hard = 220 * 1024;
    if (contig) {
        int pitch = np + 8;                       // alignment shift of the 16-byte path (<=3) + prefetch pad
        while (pitch % 32 != 4) pitch++;
        int W = 16;
        while (W > 1 && ((size_t)W * pitch + 8) * 4 > soft) W >>= 1;
        if (((size_t)W * pitch + 8) * 4 > soft) {   // even one line is big: allow one CTA per SM
            W = 16;
            while (W > 1 && ((size_t)W * pitch + 8) * 4 > hard) W >>= 1;
            if (((size_t)W * pitch + 8) * 4 > hard) return t;
        }
        t.W = W; t.pitch = pitch; t.smem = ((size_t)W * pitch + 8) * 4;
    } else if (vec) {
        t.W = 16; t.pitch = 16;
        t.smem = (size_t)(((np + 3) >> 2) + 2) * TRI_GP * 4;
        if (t.smem > hard) return t;
    } else {
        t.W = 16; t.pitch = 16; t.smem = (size_t)np * 16 * 4;
        if (t.smem > hard) return t;
    }
    t.ok = true;
    t.ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (227 * 1024) / (t.smem + 1024)));
    return t;
}

template <typename K>
static int tile_launch_k(pst_ctx *c, int cls, K kern, bool *attr_done, const TriArgs &A, size_t smem, int grid)
{
    const double bytes = A.bytes;
    if (!*attr_done) {
        PST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
        *attr_done = true;
    }
    PST_LAUNCHB(c, cls, bytes, (kern<<<grid, 128, smem, c->stream>>>(A)));
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

template <int EPI>
static int tile_launch_epi(pst_ctx *c, int cls, bool contig, bool vec, const TriArgs &A, size_t smem, int grid)
{
    // cudaFuncSetAttribute is per device: one flag per (kernel, device)
    static bool done[4][64] = {};
    bool *d = &done[0][c->device & 63];
    if (vec) return contig ? tile_launch_k(c, cls, tri_tile_contig_v4_kernel<EPI>, d, A, smem, grid)
                           : tile_launch_k(c, cls, tri_tile_strided_v4_kernel<EPI>, d + 64, A, smem, grid);
    return contig ? tile_launch_k(c, cls, tri_tile_kernel<true, EPI>, d + 128, A, smem, grid)
                  : tile_launch_k(c, cls, tri_tile_kernel<false, EPI>, d + 192, A, smem, grid);
}

static int tile_launch(pst_ctx *c, int cls, int epi, bool contig, bool vec, const TriArgs &A, size_t smem, int grid)
{
    switch (epi) {
        case EPI_GP:        return tile_launch_epi<EPI_GP>(c, cls, contig, vec, A, smem, grid);
        case EPI_DIR_FIRST: return tile_launch_epi<EPI_DIR_FIRST>(c, cls, contig, vec, A, smem, grid);
        case EPI_DIR:       return tile_launch_epi<EPI_DIR>(c, cls, contig, vec, A, smem, grid);
        default:            return tile_launch_epi<EPI_NONE>(c, cls, contig, vec, A, smem, grid);
    }
}

int pst_comm_send(pst_ctx *c, const float *d_buf, size_t count, int peer);            // pst_comm.cu
int pst_comm_recv(pst_ctx *c, float *d_buf, size_t count, int peer);
int pst_comm_sendrecv(pst_ctx *c, const float *send, size_t nsend, int peer_out, float *recv, size_t nrecv, int peer_in);
int pst_comm_halo_exchange(pst_ctx *c, const float *send_lo, const float *send_hi, float *recv_lo,
                           float *recv_hi, size_t count);

int pst_comm_mailbox(pst_ctx *c, size_t L, pst_mailbox_view *v);                      // pst_comm.cu
int pst_comm_check(pst_ctx *c);

// launches of the tile kernels (shared by the distributed pass and the single-GPU short-axis path)
static int tri3_tiles_attr(pst_ctx *c)
{
    static bool attr_dev[64] = {};                       // per-device function attribute
    bool &attr_done = attr_dev[c->device & 63];
    if (!attr_done) {
#define PST_T3_ATTR(K) PST_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024))
        PST_T3_ATTR((tri3_tile_fwd_kernel<128, true>)); PST_T3_ATTR((tri3_tile_fwd_kernel<64, true>)); PST_T3_ATTR((tri3_tile_fwd_kernel<32, true>));
        PST_T3_ATTR((tri3_tile_fwd_kernel<128, false>)); PST_T3_ATTR((tri3_tile_fwd_kernel<64, false>)); PST_T3_ATTR((tri3_tile_fwd_kernel<32, false>));
        PST_T3_ATTR((tri3_tile_bwd_kernel<128, true>)); PST_T3_ATTR((tri3_tile_bwd_kernel<64, true>)); PST_T3_ATTR((tri3_tile_bwd_kernel<32, true>));
        PST_T3_ATTR((tri3_tile_bwd_kernel<128, false>)); PST_T3_ATTR((tri3_tile_bwd_kernel<64, false>)); PST_T3_ATTR((tri3_tile_bwd_kernel<32, false>));
#undef PST_T3_ATTR
        attr_done = true;
    }
    return PST_OK;
}
static int tri3_tiles_fwd(pst_ctx *c, const Tri3Args &A, int W, unsigned blocks, bool rcmp, size_t nvox)
{
    PST_TRY(tri3_tiles_attr(c));
    const size_t smem_f = (size_t)(A.K1 - A.K0 + 2 * A.nb) * W * 4;
    // 128 threads whatever the tile width: the narrower tiles keep their burst-load issue rate (measured: 64-line tiles
    // with 64 threads 110 ms per step against 95 with 128, 2 ranks x 128-plane slabs)
    static const int T_env = []() { const char *e = getenv("PST_TRI3_THREADS"); const int v = e ? atoi(e) : 128; return v == 64 || v == 32 ? v : 128; }();
    const int T = T_env > W ? T_env : W;
    PST_LAUNCHB(c, PST_K_TRI3, 8.0 * (double)nvox,
        if (rcmp) {
            if (W == 128) tri3_tile_fwd_kernel<128, false><<<blocks, T, smem_f, c->stream>>>(A);
            else if (W == 64) tri3_tile_fwd_kernel<64, false><<<blocks, T, smem_f, c->stream>>>(A);
            else tri3_tile_fwd_kernel<32, false><<<blocks, T, smem_f, c->stream>>>(A);
        } else {
            if (W == 128) tri3_tile_fwd_kernel<128, true><<<blocks, T, smem_f, c->stream>>>(A);
            else if (W == 64) tri3_tile_fwd_kernel<64, true><<<blocks, T, smem_f, c->stream>>>(A);
            else tri3_tile_fwd_kernel<32, true><<<blocks, T, smem_f, c->stream>>>(A);
        });
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}
static int tri3_tiles_bwd(pst_ctx *c, const Tri3Args &A, int W, unsigned blocks, bool rcmp, size_t nvox, int cls)
{
    const size_t smem_f = (size_t)(A.K1 - A.K0 + 2 * A.nb) * W * 4, smem_b = (size_t)(A.K1 - A.K0) * W * 4;
    static const int T_env = []() { const char *e = getenv("PST_TRI3_THREADS"); const int v = e ? atoi(e) : 128; return v == 64 || v == 32 ? v : 128; }();
    const int T = T_env > W ? T_env : W;
    PST_LAUNCHB(c, cls, 8.0 * (double)nvox,
        if (rcmp) {
            if (W == 128) tri3_tile_bwd_kernel<128, true><<<blocks, T, smem_f, c->stream>>>(A);
            else if (W == 64) tri3_tile_bwd_kernel<64, true><<<blocks, T, smem_f, c->stream>>>(A);
            else tri3_tile_bwd_kernel<32, true><<<blocks, T, smem_f, c->stream>>>(A);
        } else {
            if (W == 128) tri3_tile_bwd_kernel<128, false><<<blocks, T, smem_b, c->stream>>>(A);
            else if (W == 64) tri3_tile_bwd_kernel<64, false><<<blocks, T, smem_b, c->stream>>>(A);
            else tri3_tile_bwd_kernel<32, false><<<blocks, T, smem_b, c->stream>>>(A);
        });
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

// Single GPU, short third axis (PST_TRI3_SOLO=1): the tile kernels of the distributed pass on the whole axis (first and
// last rank at once: no carries, no halos).  Exists to measure and profile those kernels without neighbours.
static int smooth_axis3_solo(pst_ctx *c, const DipGeom &g, const float *src, float *dst, float *scr, int W)
{
    Tri3Args A{};
    const int nb = g.r3;
    A.x = src; A.F = scr; A.dst = dst; A.L = (long)g.n1 * g.n2; A.l0 = 0; A.l1 = A.L;
    A.n3g = g.n3; A.z0 = 0; A.nz = g.n3; A.nb = nb; A.K0 = 0; A.K1 = g.n3 + 2 * nb;
    A.wt = (float)(1.0 / ((double)nb * nb));
    A.w2 = (float)(2. * A.wt);
    A.err = nullptr; A.epoch = 1;                        // (no waits: nothing can time out)
    static const bool rev_on = []() { const char *e = getenv("PST_TRI3_REV"); return !(e && e[0] == '0'); }();
    A.rev = rev_on ? 1 : 0;
    static const bool fake_reg = []() { const char *e = getenv("PST_TRI3_SOLO"); return e && e[0] == '2'; }();
    const int fakeNZ = tri3_reg_chunk(nb, 2 * g.n3, 2);
    if (fake_reg && fakeNZ) {
        // TIMING ONLY (wrong numbers): the register kernels as an interior rank would run them, halo rows read from the
        // volume itself, no carries
        A.K0 = nb; A.K1 = g.n3 + nb; A.n3g = g.n3 + 1000; A.hb = src; A.ha = src; A.csave = scr;
        // carries: a mailbox whose every word already reads 0x01010101 (= the epoch), and one to send into
        uint2 *in = (uint2 *)(scr + A.L), *out = in + A.L;
        A.epoch = 0x01010101u; A.err = (unsigned *)(out + A.L);
        PST_CUDA(cudaMemsetAsync(in, 0x01, (size_t)A.L * sizeof(uint2), c->stream));
        PST_CUDA(cudaMemsetAsync(A.err, 0, 64, c->stream));
        A.pin = in; A.pout = out;
        PST_TRY(tri3_reg_launch(c, A, true, fakeNZ, g.n, PST_K_TRI3));
        PST_TRY(tri3_reg_launch(c, A, false, fakeNZ, g.n, PST_K_TRI3));
        c->stats.smooth_passes++;
        return PST_OK;
    }
    const unsigned blocks = (unsigned)((A.L + W - 1) / W);
    PST_TRY(tri3_tiles_fwd(c, A, W, blocks, false, g.n));
    PST_TRY(tri3_tiles_bwd(c, A, W, blocks, false, g.n, PST_K_TRI3));
    c->stats.smooth_passes++;
    return PST_OK;
}

static int smooth_axis3_dist(pst_ctx *c, const DipGeom &g, const float *src, float *dst, float *scr)
{
    const int nb = g.r3, nz = g.n3, n3g = g.n3g;
    const long L = (long)g.n1 * g.n2;
    const bool first = c->rank == 0, last = c->rank == c->nranks - 1;
    pst_mailbox_view mb;
    PST_TRY(pst_comm_mailbox(c, (size_t)L, &mb));
    Tri3Args A{};
    // nb-plane halos of the CURRENT input (it changes every pass).  Equal slabs give every rank the same arena layout:
    // the kernels then read the halo planes straight from the neighbours' slabs over NVLink (peer memory), after
    // each rank has announced "my input is complete" with one flag store per neighbour -- no copy, no NCCL call.
    // WAR safety: a neighbour overwrites those planes only in or after ITS backward kernel, which cannot start before
    // its forward kernel has received the carries of all my forward CTAs, i.e. after they have read their halos.
    const char *sp = (const char *)src;
    const bool in_arena = sp >= c->arena && sp + g.n * sizeof(float) <= c->arena + c->arena_size;
    static const bool peer_on = []() { const char *e = getenv("PST_TRI3_PEERHALO"); return !(e && e[0] == '0'); }();
    const bool peer = peer_on && in_arena && (first || g.peer_prev) && (last || g.peer_next) && (n3g % c->nranks == 0);
    if (peer) {
        const size_t off = (size_t)(sp - c->arena);
        A.hb = first ? nullptr : (const float *)(g.peer_prev + off) + (size_t)(nz - nb) * L;
        A.ha = last ? nullptr : (const float *)(g.peer_next + off);
        A.hr_prev = first ? nullptr : mb.hr_in_prev;
        A.hr_next = last ? nullptr : mb.hr_in_next;
        PST_LAUNCH(c, PST_K_OTHER, (tri3_halo_ready_kernel<<<1, 32, 0, c->stream>>>(mb.hr_out_prev, mb.hr_out_next, mb.epoch)));
    } else {
        PST_TRY(pst_comm_halo_exchange(c, src, src + (size_t)(nz - nb) * L, g.hb, g.ha, (size_t)nb * L));
        A.hb = g.hb; A.ha = g.ha;
    }
    A.x = src; A.F = scr; A.dst = dst; A.L = L; A.l0 = 0; A.l1 = L;
    A.n3g = n3g; A.z0 = g.z0; A.nz = nz; A.nb = nb;
    A.K0 = first ? 0 : g.z0 + nb;
    A.K1 = last ? n3g + 2 * nb : g.z0 + nz + nb;
    A.wt = (float)(1.0 / ((double)nb * nb));
    A.w2 = (float)(2. * A.wt);
    A.err = mb.err; A.epoch = mb.epoch;
    // tile kernels (burst-loaded shared-memory tiles, short CTA latency) when they fit; the tile width
    // must be the same on every rank (the carry flags are per CTA): derived from the tallest slab
    const int nz_max = (n3g + c->nranks - 1) / c->nranks;
    const int W = tri3_tile_width(nz_max + 3 * nb, L, src, dst, scr, A.hb, A.ha);
    const unsigned blocks = (unsigned)((L + (W ? W : 128) - 1) / (W ? W : 128));
    // tile kernels, PST_TRI3_RC=1: recompute F in the backward kernel instead of staging it through HBM (12 instead of 16 B
    // per voxel).  Measured SLOWER on B200 (2 ranks, 128-plane slabs: 0.55 against 0.45 ms per pass; 256-plane slabs: 1.54
    // against 1.16 ms): these kernels are bound by the serial per-thread chains of their CTAs, not by HBM, and the
    // recomputation lengthens exactly that.  Off by default.
    static const bool rcmp_on = []() { const char *e = getenv("PST_TRI3_RC"); return e && e[0] == '1'; }();
    // register kernels (short slabs, first choice): always the recompute scheme
    const int regNZ = tri3_reg_chunk(nb, n3g, c->nranks);      // chunk height of the register kernels, 0 = not applicable
    const bool reg = regNZ != 0;
    const bool rcmp = (rcmp_on && W > 0) || reg;
    A.csave = g.cin;
    A.ha_keep = (rcmp && peer && !last) ? g.ha : nullptr;
    // forward sums: carries flow rank -> rank+1, CTA by CTA, through the neighbour's mailbox
    A.cin = first ? nullptr : mb.cf_in;  A.fin = mb.ff_in;
    A.cout = last ? nullptr : mb.cf_out; A.fout = mb.ff_out;
    A.pin = first ? nullptr : mb.pf_in; A.pout = last ? nullptr : mb.pf_out;
    A.pself = (uint2 *)mb.pb_in;
    if (reg) {
        PST_TRY(tri3_reg_launch(c, A, true, regNZ, g.n, PST_K_TRI3));
    } else if (W) {
        PST_TRY(tri3_tiles_fwd(c, A, W, blocks, rcmp, g.n));
    } else {
        static const bool win_on = []() { const char *e = getenv("PST_TRI3_WIN"); return !(e && e[0] == '0'); }();
        PST_LAUNCHB(c, PST_K_TRI3, 8.0 * (double)g.n,
            if (!win_on) tri3_dist_fwd_kernel<<<blocks, 128, 0, c->stream>>>(A);
            else switch (nb) {
                case 2: tri3_dist_fwd_win_kernel<2><<<blocks, 128, 0, c->stream>>>(A); break;
                case 3: tri3_dist_fwd_win_kernel<3><<<blocks, 128, 0, c->stream>>>(A); break;
                case 4: tri3_dist_fwd_win_kernel<4><<<blocks, 128, 0, c->stream>>>(A); break;
                case 5: tri3_dist_fwd_win_kernel<5><<<blocks, 128, 0, c->stream>>>(A); break;
                case 6: tri3_dist_fwd_win_kernel<6><<<blocks, 128, 0, c->stream>>>(A); break;
                case 7: tri3_dist_fwd_win_kernel<7><<<blocks, 128, 0, c->stream>>>(A); break;
                case 8: tri3_dist_fwd_win_kernel<8><<<blocks, 128, 0, c->stream>>>(A); break;
                case 10: tri3_dist_fwd_win_kernel<10><<<blocks, 128, 0, c->stream>>>(A); break;
                default: tri3_dist_fwd_kernel<<<blocks, 128, 0, c->stream>>>(A); break;
            });
    }
    // backward sums (+ fold): carries flow rank -> rank-1
    A.cin = last ? nullptr : mb.cb_in;    A.fin = mb.fb_in;
    A.cout = first ? nullptr : mb.cb_out; A.fout = mb.fb_out;
    A.pin = (last && !reg) ? nullptr : mb.pb_in; A.pout = first ? nullptr : mb.pb_out;
    static const bool rev_on = []() { const char *e = getenv("PST_TRI3_REV"); return !(e && e[0] == '0'); }();
    A.rev = rev_on ? 1 : 0;
    static const int bwd_cls = []() { const char *e = getenv("PST_TRI3_SPLIT"); return (e && e[0] == '1') ? PST_K_TRI3BWD : PST_K_TRI3; }();
    if (reg) {
        PST_TRY(tri3_reg_launch(c, A, false, regNZ, g.n, bwd_cls));
    } else if (W) {
        PST_TRY(tri3_tiles_bwd(c, A, W, blocks, rcmp, g.n, bwd_cls));
    } else {
        PST_LAUNCHB(c, PST_K_TRI3, 8.0 * (double)g.n, (tri3_dist_bwd_kernel<<<blocks, 128, 0, c->stream>>>(A)));
    }
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

// one axis: src -> dst (dst may alias src).  axis 0/1/2.  epi (may be null) is fused only on
// the tile path; *fused reports it.
static int smooth_axis(pst_ctx *c, const DipGeom &g, int axis, const float *src, float *dst, float *scr,
                       const EpiSpec *epi, bool *fused)
{
    const int nn[3] = {g.n1, g.n2, g.n3}, rr[3] = {g.r1, g.r2, g.r3};
    const int nx = nn[axis], nb = rr[axis];
    const int cls = axis == 0 ? PST_K_TRI1 : (axis == 1 ? PST_K_TRI2 : PST_K_TRI3);
    if (fused) *fused = false;
    static const bool solo3 = []() { const char *e = getenv("PST_TRI3_SOLO"); return e && (e[0] == '1' || e[0] == '2'); }();
    if (axis == 2 && !g.dist && solo3 && nb > 1 && 2 * nb <= g.n3 && !(epi && epi->kind != EPI_NONE)) {
        const int W = tri3_tile_width(g.n3 + 4 * nb, (long)g.n1 * g.n2, src, dst, scr, nullptr, nullptr);
        if (W) return smooth_axis3_solo(c, g, src, dst, scr, W);
    }
    if (axis == 2 && g.dist) {
        c->stats.smooth_passes++;
        return smooth_axis3_dist(c, g, src, dst, scr);
    }
    // 16-byte path: every row/line start must be 16-byte aligned
    auto al16 = [](const void *q) { return q == nullptr || (((uintptr_t)q) & 15) == 0; };
    const bool has_epi = epi && epi->kind != EPI_NONE;
    const bool vec = (g.n1 % 4 == 0) && al16(src) && al16(dst) &&
                     (!has_epi || (al16(epi->p) && al16(epi->w) && al16(epi->gp) && al16(epi->sp) && al16(epi->sx) && al16(epi->sr)));
    c->stats.smooth_passes++;
    // main path: the streaming (persistent, TMA-staged, ping-pong chain) kernel, pst_tri_stream.cu
    static const bool stream_on = []() { const char *e = getenv("PST_TRI_STREAM"); return !(e && e[0] == '0'); }();
    const bool stream_ok = stream_on && pst_tri_stream_ok(axis, g.n1, g.n2, g.n3, nb, src, dst);
    if (stream_ok && has_epi && epi->kind == EPI_GP && al16(epi->gp) && al16(epi->p)) {
        // gp = eps*p + S(src) and sum gp^2 in the store side of the streaming kernel
        pst_tri_stream_epi e{epi->p, epi->eps, c->d_partial, PST_RED_SLOTS};
        int rc = 0, grid = 0;
        PST_LAUNCHB(c, cls, 12.0 * (double)g.n,
                    rc = pst_tri_stream_launch(c->stream, c->sm_count, axis, src, epi->gp, g.n1, g.n2, g.n3, nb, nullptr, &e, &grid));
        if (rc != 0) { pst_set_error("pst_tri_stream_launch failed (%d)", rc); return PST_ECUDA; }
        PST_TRY(pst_finish_reduce(c, grid, 1, epi->rec));
        if (fused) *fused = true;
        return PST_OK;
    }
    if (has_epi && epi->stream_only) { epi = nullptr; }
    const bool has_epi2 = epi && epi->kind != EPI_NONE;
    // the L2-resident checkpoint + recompute smoother, pst_tri_l2.cu.  PST_TRI_L2 = bit mask of the axes it takes
    // (bit 0: axis 1 / contiguous, bit 1: axis 2, bit 2: axis 3)
    static const int l2_axes = []() { const char *e = getenv("PST_TRI_L2"); return e ? atoi(e) : PST_TRI_L2_DEFAULT; }();
    if (((l2_axes >> axis) & 1) && !has_epi2 && pst_tri_l2_ok(axis, g.n1, g.n2, g.n3, nb, src, dst)) {
        int rc = 0;
        PST_LAUNCHB(c, cls, 8.0 * (double)g.n, rc = pst_tri_l2_launch(c->stream, c->sm_count, axis, src, dst, g.n1, g.n2, g.n3, nb));
        if (rc == 0) return PST_OK;
        if (rc == -5) { pst_set_error("pst_tri_l2_launch: kernel launch failed"); return PST_ECUDA; }
        // set-up refused (tensor map encoding, attribute): nothing was launched, fall through
    }
    // the systolic register-resident smoother, pst_tri_sys.cu.  PST_TRI_SYS=1: every axis, =2: strided axes only, 0: off
    // default 2 (measured on B200, 1000x1024x1024: 2.06 - 2.10 ms per strided pass against 2.37 - 2.49 for the streaming kernel)
    static const int sys_mode = []() { const char *e = getenv("PST_TRI_SYS"); return e ? atoi(e) : 2; }();
    if (sys_mode > 0 && (axis != 0 || sys_mode == 1) && !has_epi2 && pst_tri_sys_ok(axis, g.n1, g.n2, g.n3, nb, src, dst)) {
        int rc = 0;
        PST_LAUNCHB(c, cls, 8.0 * (double)g.n, rc = pst_tri_sys_launch(c->stream, c->sm_count, axis, src, dst, g.n1, g.n2, g.n3, nb, nullptr));
        if (rc == 0) return PST_OK;
        if (rc == -5) { pst_set_error("pst_tri_sys_launch: kernel launch failed"); return PST_ECUDA; }
    }
    // experiment (off by default): the checkpoint + recompute smoother, pst_tri_rc.cu.  PST_TRI_RC=1: strided axes,
    // =2: every axis; PST_TRI_RC_BLOCK=16|32: block length
    static const int rc_mode = []() { const char *e = getenv("PST_TRI_RC"); return e ? atoi(e) : 0; }();
    static const int rc_block = []() { const char *e = getenv("PST_TRI_RC_BLOCK"); return e ? atoi(e) : 32; }();
    if (rc_mode > 0 && !has_epi2 && (axis != 0 || rc_mode >= 2) && pst_tri_rc_ok(axis, g.n1, g.n2, g.n3, nb, rc_block)) {
        int rc = 0;
        PST_LAUNCHB(c, cls, 8.0 * (double)g.n, rc = pst_tri_rc_launch(c->stream, axis, src, dst, g.n1, g.n2, g.n3, nb, rc_block));
        if (rc == 0) return PST_OK;
        if (rc != -1) { pst_set_error("pst_tri_rc_launch failed (%d)", rc); return PST_ECUDA; }
    }
    if (stream_ok && !has_epi2) {
        int rc = 0;
        PST_LAUNCHB(c, cls, 8.0 * (double)g.n,
                    rc = pst_tri_stream_launch(c->stream, c->sm_count, axis, src, dst, g.n1, g.n2, g.n3, nb, nullptr));
        if (rc == 0) return PST_OK;
        if (rc == -5) { pst_set_error("pst_tri_stream_launch: kernel launch failed"); return PST_ECUDA; }
        // set-up refused (tensor map encoding, shared-memory attribute): nothing was launched, use the tile kernels
    }
    const TilePlan tp = tile_plan(axis == 0, vec, nx, nb);
    if (tp.ok) {
        TriArgs A{};
        A.src = src; A.dst = dst; A.nx = nx; A.nb = nb; A.W = tp.W; A.pitch = tp.pitch;
        A.wt = (float)(1.0 / ((double)nb * nb));            // ps_triangle_init :421
        A.w2 = (float)(2. * A.wt);
        if (axis == 0) {
            A.nlines = (long)g.n2 * g.n3;
            A.ngroups = (A.nlines + tp.W - 1) / tp.W;
        } else if (axis == 1) {
            A.na = g.n1; A.sb = (long)g.n1 * g.n2; A.d = g.n1;
            A.ngroups = ((A.na + 15) / 16) * (long)g.n3;
        } else {
            A.na = (long)g.n1 * g.n2; A.sb = 0; A.d = (long)g.n1 * g.n2;
            A.ngroups = (A.na + 15) / 16;
        }
        A.bytes = 8.0 * (double)g.n;
        int kind = EPI_NONE;
        if (epi && epi->kind != EPI_NONE) {
            A.bytes += (epi->kind == EPI_GP ? 4.0 : (epi->kind == EPI_DIR ? 28.0 : 16.0)) * (double)g.n;
            kind = epi->kind;
            A.p = epi->p; A.w = epi->w; A.gp = epi->gp; A.sp = epi->sp; A.sx = epi->sx; A.sr = epi->sr;
            A.eps = epi->eps; A.alpha = epi->alpha; A.partial = c->d_partial;
        }
        long grid = std::min<long>(A.ngroups, (long)c->sm_count * tp.ctas_per_sm);
        PST_TRY(tile_launch(c, cls, kind, axis == 0, vec, A, tp.smem, (int)grid));
        if (kind != EPI_NONE) {
            PST_TRY(pst_finish_reduce(c, (int)grid, 3, epi->rec));
            if (fused) *fused = true;
        }
        return PST_OK;
    }
    // fallback: in-place line kernels with F staged through global scratch
    if (dst != src) PST_CUDA(cudaMemcpyAsync(dst, src, g.n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    if (axis == 0) return tri_lines_launch(c, cls, dst, scr, (long)g.n2 * g.n3, (long)g.n2 * g.n3, g.n1, 0, 1, nx, nb);
    if (axis == 1) return tri_lines_launch(c, cls, dst, scr, (long)g.n1 * g.n3, g.n1, 1, (long)g.n1 * g.n2, g.n1, nx, nb);
    return tri_lines_launch(c, cls, dst, scr, (long)g.n1 * g.n2, (long)g.n1 * g.n2, 1, 0, (long)g.n1 * g.n2, nx, nb);
}

// ps_trianglen_lop body (:692-700): tmp <- S(src), axes 1, 2, 3 in turn.  With `epi`, the last
// smoothed axis applies the epilogue instead of storing to tmp (then *fused = true and tmp is
// not the result); `late` is invoked just before that last launch.  Returns 1 when `late`
// vetoed the last axis.
int pst_shape_apply(pst_ctx *c, const DipGeom &g, const float *src, float *tmp, float *scr,
                    EpiSpec *epi, late_bind_fn late, void *user, bool *fused)
{
    const int rr[3] = {g.r1, g.r2, g.r3};
    int act[3], na = 0;
    for (int a = 0; a < 3; a++) if (rr[a] > 1) act[na++] = a;
    if (fused) *fused = false;
    if (na == 0) {                                   // S = identity
        if (late) { int v = late(user, epi); if (v) return v; }
        if (tmp != src) PST_CUDA(cudaMemcpyAsync(tmp, src, g.n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        return PST_OK;
    }
    const float *in = src;
    for (int j = 0; j < na; j++) {
        const bool last = (j == na - 1);
        if (last && late) { int v = late(user, epi); if (v) return v; }
        PST_TRY(smooth_axis(c, g, act[j], in, tmp, scr, last ? epi : nullptr, last ? fused : nullptr));
        in = tmp;
    }
    return PST_OK;
}

int pst_smooth3_inplace(pst_ctx *c, float *x, float *scr, int n1, int n2, int n3, int r1, int r2, int r3)
{
    DipGeom g = make_geom(n1, n2, n3, r1, r2, r3);
    return pst_shape_apply(c, g, x, x, scr, nullptr, nullptr, nullptr, nullptr);
}

template <int NW>
static int allpass_launch_nw(pst_ctx *c, const float *u, const float *p_in, const float *dp, float lam,
                             float *p_out, float *y, int n1, int n2, int n3, int xline, bool der,
                             bool ls, int rec, int n3_live, int z0, int n3g)
{
    static const BTab tb = make_btab(NW);
    const unsigned ch = pst_red_ch((size_t)n1 * n2, n3g);
    const int tg = n1 >= (int)ch ? 1 : (int)(ch / (unsigned)n1);     // traces per piece: a function of the global shape only
    const unsigned gpp = (unsigned)((n2 + tg - 1) / tg);
    const long blocks = (long)gpp * n3;
    PST_TRY(pst_reserve_partials(c, (size_t)blocks, n3g));
    const int threads = n1 >= 256 ? 256 : (n1 >= 128 ? 128 : 64);
    PST_LAUNCHB(c, PST_K_ALLPASS, (ls ? 20.0 : 12.0) * (double)n1 * n2 * n3,
        if (ls)
            allpass_kernel<NW, false, true><<<(unsigned)blocks, threads, 0, c->stream>>>(u, p_in, dp, lam, p_out, y, n1, n2, n3, xline, n3_live, tb, gpp, tg, c->d_partial);
        else if (der)
            allpass_kernel<NW, true, false><<<(unsigned)blocks, threads, 0, c->stream>>>(u, p_in, nullptr, 0.f, nullptr, y, n1, n2, n3, xline, n3_live, tb, gpp, tg, c->d_partial);
        else
            allpass_kernel<NW, false, false><<<(unsigned)blocks, threads, 0, c->stream>>>(u, p_in, nullptr, 0.f, nullptr, y, n1, n2, n3, xline, n3_live, tb, gpp, tg, c->d_partial));
    PST_CUDA(cudaGetLastError());
    return pst_finish_reduce_canon(c, (int)gpp, n3, z0, n3g, 1, rec);
}

// n3_live: planes whose next plane exists (xline stencil); -1 = n3 - 1 (whole cube on this GPU).
// In a distributed context u carries one halo plane behind the slab and n3_live = n3 except on
// the last rank.
int pst_allpass_launch(pst_ctx *c, const float *u, const float *p_in, const float *dp, float lam,
                       float *p_out, float *y, int n1, int n2, int n3, int nw, int xline, bool der,
                       bool ls, int rec, int n3_live = -1, int z0 = 0, int n3g = -1)
{
    if (n3_live < 0) n3_live = n3 - 1;
    if (n3g < 0) n3g = n3;
    if (nw == 1) return allpass_launch_nw<1>(c, u, p_in, dp, lam, p_out, y, n1, n2, n3, xline, der, ls, rec, n3_live, z0, n3g);
    if (nw == 2) return allpass_launch_nw<2>(c, u, p_in, dp, lam, p_out, y, n1, n2, n3, xline, der, ls, rec, n3_live, z0, n3g);
    pst_set_error("order=%d unsupported (1 or 2)", nw);
    return PST_EUNSUP;
}

struct CgWork {
    float *p, *r, *sp, *sx, *sr, *gp, *tmp, *scr;
};

// scalars of ps_conjgrad that gate the direction update (:330-349): fetched as late as possible
struct CgLate { pst_ctx *c; double gn, gnp, g0; float tol; int iter; bool stop; float alpha; };

static int cg_late_bind(void *user, EpiSpec *epi)
{
    CgLate *L = (CgLate *)user;
    double h[PST_RED_SLOTS];
    PST_TRY(pst_fetch_record(L->c, 1, 1, h));
    L->gn = h[0];
    if (L->iter == 0) { L->g0 = L->gn; return 0; }
    const double alpha = L->gn / L->gnp, dg = L->gn / L->g0;
    if (alpha < L->tol || dg < L->tol) { L->stop = true; return 1; }
    L->alpha = (float)alpha;
    if (epi) epi->alpha = L->alpha;
    return 0;
}

// ps_divne (:796-827) + ps_conjgrad (:257-383, prec=NULL, hasp0=false, eps=1*1, tol=1e-6).
// num/den are overwritten (den becomes the weight); rat receives the CG model x.
int pst_divne_run(pst_ctx *c, const DipGeom &g, float *num, float *den, float *rat,
                  const unsigned char *mask, const CgWork &w, int liter, float eps_div, int *iters_run)
{
    const size_t n = g.n;
    const float eps = 1.f * 1.f, tol = 1.e-6f;
    const int threads = 256;
    const int grid = pst_grid_for(c, n, threads);
    // every sum of the solve is canonical (pst_common.cuh): the scalars do not depend on the slab decomposition
    const Span S = pst_span_canon((size_t)g.n1 * g.n2, g.n3, g.n3g);
    const unsigned gridc = S.ppp * (unsigned)g.n3;
    PST_TRY(pst_reserve_partials(c, gridc, g.n3g));
    auto finish = [&](int nv, int rec) { return pst_finish_reduce_canon(c, (int)S.ppp, g.n3, g.z0, g.n3g, nv, rec); };
    double h[PST_RED_SLOTS];
    if (iters_run) *iters_run = 0;

    PST_LAUNCHB(c, PST_K_CGVEC, 16.0 * (double)n, (divne_prescale_kernel<<<gridc, threads, 0, c->stream>>>(num, den, mask, eps_div, S, c->d_partial)));
    PST_TRY(finish(1, 0));
    PST_TRY(pst_fetch_record(c, 0, 1, h));
    if (h[0] == 0.0) {
        PST_LAUNCH(c, PST_K_OTHER, (fill_kernel<<<grid, threads, 0, c->stream>>>(rat, 0.f, n)));
        return PST_OK;
    }
    const double norm = sqrt(g.nglob / h[0]);
    PST_LAUNCHB(c, PST_K_CGVEC, 24.0 * (double)n, (divne_scale_init_kernel<<<gridc, threads, 0, c->stream>>>(num, den, norm, w.r, w.p, rat, S, c->d_partial)));
    PST_TRY(finish(1, 0));
    PST_TRY(pst_fetch_record(c, 0, 1, h));
    if (h[0] == 0.0) return PST_OK;               // zero residual: p = x = 0 (:299-303)

    // A/B switch.  Measured on B200 at 1000x1024x1024 (round 1): fusing the CG vector work into the
    // epilogue of the last smoothing pass costs more than it saves (4.83 s unfused vs 5.15 s fused per
    // step): the streaming kernels run at ~5.5 TB/s, the tile kernel's phase 3 does not.  Default: off.
    static const bool fuse = []() { const char *e = getenv("PST_FUSE_EPILOGUE"); return e && e[0] == '1'; }();
    static const bool fuse_gp = fuse || []() { const char *e = getenv("PST_FUSE_GP"); return e && e[0] == '1'; }();   // measured: 4.5 ms fused pass vs 2.5 + 2.7 ms separate -> off
    // 16-byte kernels when every vector is 16-byte aligned and n % 4 == 0
    auto a16 = [](const void *q) { return (((uintptr_t)q) & 15) == 0; };
    static const bool vec_on = []() { const char *e = getenv("PST_CG_VEC4"); return !(e && e[0] == '0'); }();
    const bool vec4 = vec_on && (S.n12 % 4 == 0) && a16(w.p) && a16(rat) && a16(w.r) && a16(w.sp) && a16(w.sx) && a16(w.sr) &&
                      a16(den) && a16(w.tmp) && a16(w.gp);
    const int grid4 = pst_grid_for(c, n / 4 + 1, threads, 2);
    // default: CG scalars stay on the device (no host synchronisation inside the solve).  PST_CG_DEVSCALARS=0 or one
    // of the fused-epilogue experiments selects the host-scalar loop below.
    static const bool dev_scalars = []() { const char *e = getenv("PST_CG_DEVSCALARS"); return !(e && e[0] == '0'); }();
    if (dev_scalars && !fuse && !fuse_gp) {
        CgCtl *ctl = (CgCtl *)c->d_cgctl;
        const double *rec1 = c->d_red + (size_t)1 * PST_RED_SLOTS, *rec2 = c->d_red + (size_t)2 * PST_RED_SLOTS;
        *c->h_cgstop = 0;
        PST_LAUNCH(c, PST_K_OTHER, (cg_ctl_init_kernel<<<1, 1, 0, c->stream>>>(ctl)));
        int launched = 0;
        for (int iter = 0; iter < liter; iter++) {
            if (*c->h_cgstop) break;                 // the device left the solve (pinned copy, refreshed every iteration)
            PST_LAUNCHB(c, PST_K_CGHEAD, (iter ? 44.0 : 16.0) * (double)n,
                if (vec4 && iter) cg_head4d_kernel<true><<<gridc, threads, 0, c->stream>>>(w.p, rat, w.r, w.sp, w.sx, w.sr, den, ctl, eps, w.tmp, S);
                else if (vec4)    cg_head4d_kernel<false><<<gridc, threads, 0, c->stream>>>(w.p, rat, w.r, w.sp, w.sx, w.sr, den, ctl, eps, w.tmp, S);
                else if (iter)    cg_headd_kernel<true><<<gridc, threads, 0, c->stream>>>(w.p, rat, w.r, w.sp, w.sx, w.sr, den, ctl, eps, w.tmp, S);
                else              cg_headd_kernel<false><<<gridc, threads, 0, c->stream>>>(w.p, rat, w.r, w.sp, w.sx, w.sr, den, ctl, eps, w.tmp, S));
            PST_TRY(pst_shape_apply(c, g, w.tmp, w.tmp, w.scr, nullptr, nullptr, nullptr, nullptr));
            PST_LAUNCHB(c, PST_K_CGGP, 12.0 * (double)n,
                if (vec4) cg_gp4_kernel<<<gridc, threads, 0, c->stream>>>(w.p, w.tmp, w.gp, eps, S, c->d_partial);
                else cg_gp_kernel<<<gridc, threads, 0, c->stream>>>(w.p, w.tmp, w.gp, eps, S, c->d_partial));
            PST_TRY(finish(1, 1));
            PST_LAUNCH(c, PST_K_OTHER, (cg_ctl_gn_kernel<<<1, 1, 0, c->stream>>>(ctl, rec1, iter, tol)));
            PST_CUDA(cudaMemcpyAsync((void *)c->h_cgstop, &ctl->stop, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            PST_TRY(pst_shape_apply(c, g, w.gp, w.tmp, w.scr, nullptr, nullptr, nullptr, nullptr));
            PST_LAUNCHB(c, PST_K_CGDIR, (iter ? 36.0 : 24.0) * (double)n,
                if (iter == 0 && vec4) cg_dird_kernel<true, true><<<gridc, threads, 0, c->stream>>>(w.gp, w.tmp, den, w.sp, w.sx, w.sr, ctl, S, c->d_partial);
                else if (iter == 0)    cg_dird_kernel<true, false><<<gridc, threads, 0, c->stream>>>(w.gp, w.tmp, den, w.sp, w.sx, w.sr, ctl, S, c->d_partial);
                else if (vec4)         cg_dird_kernel<false, true><<<gridc, threads, 0, c->stream>>>(w.gp, w.tmp, den, w.sp, w.sx, w.sr, ctl, S, c->d_partial);
                else                   cg_dird_kernel<false, false><<<gridc, threads, 0, c->stream>>>(w.gp, w.tmp, den, w.sp, w.sx, w.sr, ctl, S, c->d_partial));
            PST_TRY(finish(3, 2));
            PST_LAUNCH(c, PST_K_OTHER, (cg_ctl_beta_kernel<<<1, 1, 0, c->stream>>>(ctl, rec2, eps)));
            launched++;
        }
        if (launched > 0)      // only the model x (= rat) is consumed after the last iteration
            PST_LAUNCHB(c, PST_K_CGHEAD, 12.0 * (double)n, (cg_taild_kernel<<<grid, threads, 0, c->stream>>>(rat, w.sx, ctl, n)));
        // executed iterations (statistics, iteration-count parity tests): the one host synchronisation of the solve
        PST_CUDA(cudaMemcpyAsync((void *)(c->h_cgstop + 1), &ctl->iters, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        PST_CUDA(cudaStreamSynchronize(c->stream));
        const int done = c->h_cgstop[1];
        c->stats.cg_iterations += done;
        if (iters_run) *iters_run = done;
        PST_CUDA(cudaGetLastError());
        return PST_OK;
    }
    CgLate L{c, 0., 0., 0., tol, 0, false, 0.f};
    float a_pending = 0.f;
    bool pending = false;
    int iter;
    for (iter = 0; iter < liter; iter++) {
        PST_LAUNCHB(c, PST_K_CGHEAD, (pending ? 44.0 : 16.0) * (double)n,
            if (vec4 && pending)
                cg_head4_kernel<true><<<grid4, threads, 0, c->stream>>>(w.p, rat, w.r, w.sp, w.sx, w.sr, den, a_pending, eps, w.tmp, n);
            else if (vec4)
                cg_head4_kernel<false><<<grid4, threads, 0, c->stream>>>(w.p, rat, w.r, w.sp, w.sx, w.sr, den, 0.f, eps, w.tmp, n);
            else if (pending)
                cg_head_kernel<true><<<grid, threads, 0, c->stream>>>(w.p, rat, w.r, w.sp, w.sx, w.sr, den, a_pending, eps, w.tmp, n);
            else
                cg_head_kernel<false><<<grid, threads, 0, c->stream>>>(w.p, rat, w.r, w.sp, w.sx, w.sr, den, 0.f, eps, w.tmp, n));
        pending = false;
        // gp = eps*p + S(gx), sum gp^2 -> record 1
        EpiSpec e1;
        e1.kind = EPI_GP; e1.p = w.p; e1.gp = w.gp; e1.eps = eps; e1.rec = 1;
        e1.stream_only = !fuse;      // default: fuse gp only where the streaming kernel does it (chain-bound pass, free bytes)
        bool fused = false;
        PST_TRY(pst_shape_apply(c, g, w.tmp, w.tmp, w.scr, fuse_gp ? &e1 : nullptr, nullptr, nullptr, &fused));
        if (!fused) {
            PST_LAUNCHB(c, PST_K_CGGP, 12.0 * (double)n,
                if (vec4) cg_gp4_kernel<<<gridc, threads, 0, c->stream>>>(w.p, w.tmp, w.gp, eps, S, c->d_partial);
                else cg_gp_kernel<<<gridc, threads, 0, c->stream>>>(w.p, w.tmp, w.gp, eps, S, c->d_partial));
            PST_TRY(finish(1, 1));
        }
        // gx = S(gp); direction update fused into the last axis once alpha is known
        EpiSpec e2;
        e2.kind = (iter == 0) ? EPI_DIR_FIRST : EPI_DIR;
        e2.w = den; e2.gp = w.gp; e2.sp = w.sp; e2.sx = w.sx; e2.sr = w.sr; e2.rec = 2;
        L.iter = iter; L.stop = false;
        int v = pst_shape_apply(c, g, w.gp, w.tmp, w.scr, fuse ? &e2 : nullptr, cg_late_bind, &L, &fused);
        if (v < 0) return v;
        if (L.stop) break;
        if (!fused) {
            if (iter == 0)
                PST_LAUNCHB(c, PST_K_CGDIR, 24.0 * (double)n,
                    if (vec4) cg_dir4_kernel<true><<<grid4, threads, 0, c->stream>>>(w.gp, w.tmp, den, w.sp, w.sx, w.sr, 0.f, n, c->d_partial);
                    else cg_dir_kernel<true><<<grid, threads, 0, c->stream>>>(w.gp, w.tmp, den, w.sp, w.sx, w.sr, 0.f, n, c->d_partial));
            else
                PST_LAUNCHB(c, PST_K_CGDIR, 36.0 * (double)n,
                    if (vec4) cg_dir4_kernel<false><<<grid4, threads, 0, c->stream>>>(w.gp, w.tmp, den, w.sp, w.sx, w.sr, L.alpha, n, c->d_partial);
                    else cg_dir_kernel<false><<<grid, threads, 0, c->stream>>>(w.gp, w.tmp, den, w.sp, w.sx, w.sr, L.alpha, n, c->d_partial));
            PST_TRY(pst_finish_reduce(c, vec4 ? grid4 : grid, 3, 2));
        }
        PST_TRY(pst_fetch_record(c, 2, 3, h));
        const double beta = h[0] + (double)eps * (h[1] - h[2]);
        const double alpha = -L.gn / beta;
        a_pending = (float)alpha;
        pending = true;
        L.gnp = L.gn;
        c->stats.cg_iterations++;
    }
    if (pending) {      // only the model x (= rat) is consumed after the last iteration
        PST_LAUNCHB(c, PST_K_CGHEAD, 12.0 * (double)n, (cg_tail_kernel<<<grid, threads, 0, c->stream>>>(rat, w.sx, a_pending, n)));
    }
    if (iters_run) *iters_run = iter;
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

// dip3 (:1619-1691) for one direction.  p holds the initial dip (zeros) and receives the result.
static int gauss_newton(pst_ctx *c, const DipGeom &g, const float *u, float *p, const unsigned char *mask,
                        int xline, int niter, int liter, int nw, float *u1, float *u2, float *dp,
                        float *ptrial, const CgWork &w, int verb, int n3_live = -1)
{
    double h[PST_RED_SLOTS];
    float *pcur = p, *pnext = ptrial;
    PST_TRY(pst_allpass_launch(c, u, pcur, nullptr, 0.f, nullptr, u2, g.n1, g.n2, g.n3, nw, xline, false, false, 3, n3_live, g.z0, g.n3g));
    PST_TRY(pst_fetch_record(c, 3, 1, h));
    double usum = h[0];
    for (int iter = 0; iter < niter; iter++) {
        PST_TRY(pst_allpass_launch(c, u, pcur, nullptr, 0.f, nullptr, u1, g.n1, g.n2, g.n3, nw, xline, true, false, 4, n3_live, g.z0, g.n3g));
        int its = 0;
        PST_TRY(pst_divne_run(c, g, u2, u1, dp, mask, w, liter, 1.0f, &its));
        float lam = 1.f;
        double usum2 = 0.;
        int k;
        for (k = 0; k < 8; k++) {
            PST_TRY(pst_allpass_launch(c, u, pcur, dp, lam, pnext, u2, g.n1, g.n2, g.n3, nw, xline, false, true, 3, n3_live, g.z0, g.n3g));
            PST_TRY(pst_fetch_record(c, 3, 1, h));
            c->stats.linesearch_evals++;
            usum2 = h[0];
            if (usum2 < usum) break;
            lam *= 0.5f;
        }
        if (verb) printf("[pst] dip%d iter %d: cg=%d ls=%d usum %.9g -> %.9g\n", xline + 1, iter, its, k < 8 ? k + 1 : 8, usum, usum2);
        usum = usum2;                  // next iteration's usum is sum(u2^2) of the kept trial
        float *t = pcur; pcur = pnext; pnext = t;
        c->stats.gn_iterations++;
    }
    if (pcur != p) PST_CUDA(cudaMemcpyAsync(p, pcur, g.n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    return PST_OK;
}

static int check_dip_args(int n1, int n2, int n3, int niter, int liter, int order, int r1, int r2, int r3)
{
    if (n1 < 1 || n2 < 1 || n3 < 1 || niter < 0 || liter < 0) { pst_set_error("dip: bad dimensions/iterations"); return PST_EINVAL; }
    if (order != 1 && order != 2) { pst_set_error("dip: order=%d unsupported (1 or 2)", order); return PST_EUNSUP; }
    if (n1 < 2 * order + 1) { pst_set_error("dip: n1=%d too short for order %d", n1, order); return PST_EINVAL; }
    if (r1 < 1 || r2 < 1 || r3 < 1) { pst_set_error("dip: rect must be >= 1"); return PST_EINVAL; }
    return PST_OK;
}

// In a distributed context (pst_ctx_create_dist, nranks > 1) n3 is the GLOBAL number of planes and
// d_din / d_mask / d_dip_out are this rank's slab (pst_ctx_slab): nz = z1 - z0 planes; d_dip_out
// holds the slab of the inline dip followed by the slab of the xline dip.
extern "C" int pst_dip_dev(pst_ctx *c, const float *d_din, const float *d_mask, int n1, int n2, int n3,
                           int niter, int liter, int order, int r1, int r2, int r3, int verb,
                           float *d_dip_out)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_TRY(check_dip_args(n1, n2, n3, niter, liter, order, r1, r2, r3));
    PST_CUDA(cudaSetDevice(c->device));
    DipGeom g;
    size_t scr = 0;
    PST_TRY(slab_geom(c, "dip", n1, n2, n3, r1, r2, r3, &g, &scr));
    const bool dist = g.dist;
    const int nz = g.n3;
    const size_t n = g.n, plane = (size_t)n1 * n2;
    // slabs: halo / carry planes of the axis-3 pass, the data (and mask) slab with the neighbour's first plane appended
    const size_t extra = dist ? (slab_halo_floats(g) + (n + 2 * plane) + (d_mask ? n + 2 * plane : 0) + 16 * 64) : 0;
    const size_t need = (11 * n + scr + extra) * sizeof(float) + 2 * n + 32 * 256;
    PST_TRY(pst_arena_reserve(c, need));
    pst_arena_reset(c);
    float *u1, *u2, *dp, *ptrial;
    CgWork w{};
    unsigned char *m_in = nullptr, *m_x = nullptr;
    PST_TRY(pst_arena_get(c, n, &u1));
    PST_TRY(pst_arena_get(c, n, &u2));
    PST_TRY(pst_arena_get(c, n, &dp));
    PST_TRY(pst_arena_get(c, n, &ptrial));
    PST_TRY(pst_arena_get(c, n, &w.p));
    PST_TRY(pst_arena_get(c, n, &w.r));
    PST_TRY(pst_arena_get(c, n, &w.sp));
    PST_TRY(pst_arena_get(c, n, &w.sx));
    PST_TRY(pst_arena_get(c, n, &w.sr));
    PST_TRY(pst_arena_get(c, n, &w.gp));
    PST_TRY(pst_arena_get(c, n, &w.tmp));
    PST_TRY(pst_arena_get(c, scr, &w.scr));
    const float *u = d_din, *um = d_mask;
    int n3_live = nz - 1;
    PST_TRY(pst_pipe_wait_planes(c, nz));                 // host-pointer entry: the uploads run on the copy stream
    if (dist) {
        // the xline stencil reads plane i3+1: keep a copy of the slab with the neighbour's first
        // plane appended (static data: exchanged once per call)
        float *ue, *dummy;
        PST_TRY(pst_arena_get(c, n + plane, &ue));
        PST_TRY(pst_arena_get(c, plane, &dummy));
        PST_CUDA(cudaMemcpyAsync(ue, d_din, n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        PST_TRY(pst_comm_halo_exchange(c, d_din, d_din, dummy, ue + n, plane));
        u = ue;
        n3_live = (c->rank == c->nranks - 1) ? nz - 1 : nz;
        PST_TRY(slab_halos(c, &g));
    }
    if (d_mask) {
        if (dist) {
            // mask32 (dip_cfuns.c:914-997) looks at the xline stencil footprint, i.e. at plane i3+1 of the mask
            // volume: same one-plane halo as the data
            float *me, *dummy;
            PST_TRY(pst_arena_get(c, n + plane, &me));
            PST_TRY(pst_arena_get(c, plane, &dummy));
            PST_CUDA(cudaMemcpyAsync(me, d_mask, n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
            PST_TRY(pst_comm_halo_exchange(c, d_mask, d_mask, dummy, me + n, plane));
            um = me;
        }
        PST_TRY(pst_arena_get(c, n, &m_in));
        PST_TRY(pst_arena_get(c, n, &m_x));
        const int grid = (int)min((long)n2 * nz, (long)c->sm_count * 8);
        PST_LAUNCH(c, PST_K_OTHER,
            if (order == 1) mask_kernel<1><<<grid, 128, 0, c->stream>>>(um, m_in, m_x, n1, n2, nz, n3_live);
            else            mask_kernel<2><<<grid, 128, 0, c->stream>>>(um, m_in, m_x, n1, n2, nz, n3_live));
    }
    const int ndip = (n3 == 1) ? 1 : 2;
    PST_CUDA(cudaMemsetAsync(d_dip_out, 0, ndip * n * sizeof(float), c->stream));
    PST_TRY(gauss_newton(c, g, u, d_dip_out, m_in, 0, niter, liter, order, u1, u2, dp, ptrial, w, verb, n3_live));
    if (ndip == 2) PST_TRY(pst_pipe_emit(c, d_dip_out, 0, nz));   // the inline dip leaves while the xline dip is estimated
    if (ndip == 2)
        PST_TRY(gauss_newton(c, g, u, d_dip_out + n, m_x, 1, niter, liter, order, u1, u2, dp, ptrial, w, verb, n3_live));
    if (dist) PST_TRY(pst_comm_check(c));
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

extern "C" int pst_allpass_dev(pst_ctx *c, const float *d_u, const float *d_sigma, int n1, int n2, int n3,
                               int order, int xline, int der, float *d_y)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    if (n1 < 2 * order + 1) { pst_set_error("allpass: n1 too short"); return PST_EINVAL; }
    if (c->comm && c->nranks > 1) { pst_set_error("allpass: single-GPU contexts only (test hook; pst_dip runs the stencil on slabs)"); return PST_EUNSUP; }
    PST_TRY(pst_allpass_launch(c, d_u, d_sigma, nullptr, 0.f, nullptr, d_y, n1, n2, n3, order, xline, der != 0, false, 5));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    return PST_OK;
}

static int smooth_axis_any(pst_ctx *c, float *x, float *scr, int n1, int n2, int n3, int a, int nb, int adj, int box, int der)
{
    const int nn[3] = {n1, n2, n3};
    const int nx = nn[a];
    const float wt = box ? (float)(1.0 / (double)(2 * nb - 1)) : (float)(1.0 / ((double)nb * nb));   // ps_triangle_init :415-424
    long nlines, na, sa, sb, d;
    if (a == 0) { nlines = (long)n2 * n3; na = nlines; sa = n1; sb = 0; d = 1; }
    else if (a == 1) { nlines = (long)n1 * n3; na = n1; sa = 1; sb = (long)n1 * n2; d = n1; }
    else { nlines = (long)n1 * n2; na = nlines; sa = 1; sb = 0; d = (long)n1 * n2; }
    const int threads = 128;
    const long blocks = (nlines + threads - 1) / threads;
    PST_LAUNCHB(c, a == 0 ? PST_K_TRI1 : (a == 1 ? PST_K_TRI2 : PST_K_TRI3), 8.0 * (double)nlines * nx,
        (tri_lines_any_kernel<<<(unsigned)blocks, threads, 0, c->stream>>>(x, scr, nlines, na, sa, sb, d, nx, nb, wt, adj, box, der)));
    c->stats.smooth_passes++;
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

// smoothcf (dip_cfuns.c:2006-2123): axes in turn, every line of an axis smoothed `repeat` times in a row
extern "C" int pst_smoothcf_dev(pst_ctx *c, float *d_x, int n1, int n2, int n3, int repeat, int adj, int r1, int r2, int r3,
                                int diff1, int diff2, int diff3, int box1, int box2, int box3)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    if (r1 < 1 || r2 < 1 || r3 < 1 || n1 < 1 || n2 < 1 || n3 < 1 || repeat < 1) { pst_set_error("smoothcf: bad arguments"); return PST_EINVAL; }
    if (c->comm && c->nranks > 1 && (adj || diff1 || diff2 || diff3 || box1 || box2 || box3 || repeat != 1)) {
        pst_set_error("smoothcf: options other than adj=0, repeat=1 are single-GPU only"); return PST_EUNSUP;
    }
    PST_CUDA(cudaSetDevice(c->device));
    // distributed contexts: n3 is the GLOBAL plane count and d_x this rank's slab (the plain triangle smoother only)
    DipGeom g;
    size_t scr = 0;
    PST_TRY(slab_geom(c, "smoothcf", n1, n2, n3, r1, r2, r3, &g, &scr));
    PST_TRY(pst_arena_reserve(c, (scr + slab_halo_floats(g)) * sizeof(float) + 4096));
    pst_arena_reset(c);
    float *s;
    PST_TRY(pst_arena_get(c, scr, &s));
    PST_TRY(slab_halos(c, &g));
    const int rr[3] = {r1, r2, r3}, df[3] = {diff1, diff2, diff3}, bx[3] = {box1, box2, box3};
    const bool plain = !adj && !diff1 && !diff2 && !diff3 && !box1 && !box2 && !box3;
    if (plain && repeat == 1) {
        PST_TRY(pst_shape_apply(c, g, d_x, d_x, s, nullptr, nullptr, nullptr, nullptr));
    } else {
        for (int a = 0; a < 3; a++) {
            if (rr[a] <= 1) continue;
            for (int q = 0; q < repeat; q++) {
                if (!adj && !df[a] && !bx[a])       // ps_smooth2 without options: the streaming kernel
                    PST_TRY(pst_smooth3_inplace(c, d_x, s, n1, n2, n3, a == 0 ? r1 : 1, a == 1 ? r2 : 1, a == 2 ? r3 : 1));
                else
                    PST_TRY(smooth_axis_any(c, d_x, s, n1, n2, n3, a, rr[a], adj ? 1 : 0, bx[a] ? 1 : 0, df[a] ? 1 : 0));
            }
        }
    }
    if (g.dist) PST_TRY(pst_comm_check(c));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    return PST_OK;
}

// Test hook: the register kernels of the distributed axis-3 pass, run RANK AFTER RANK on one GPU over the n3-slabs
// of d_x (in place), with device buffers standing in for the neighbours' mailboxes -- every kernel variant (first,
// interior, last rank) is exercised without a second GPU.  The result must equal pst_smooth3_dev(1, 1, r3) bit for bit.
extern "C" int pst_selftest_axis3_slabs(pst_ctx *c, float *d_x, int n1, int n2, int n3, int r3, int nranks)
{
    if (!c || !d_x) { pst_set_error("selftest_axis3_slabs: null pointer"); return PST_EINVAL; }
    const int NZ = tri3_reg_chunk(r3, n3, nranks);
    if (!NZ) {
        pst_set_error("selftest_axis3_slabs: geometry outside the register kernels (radius 2-6 or 8, nranks >= 2 equal slabs of 32, 64, 96 or k x 128 planes)");
        return PST_EUNSUP;
    }
    const int nch = n3 / nranks / NZ;
    PST_CUDA(cudaSetDevice(c->device));
    const long L = (long)n1 * n2;
    const int nb = r3;
    struct Bufs { uint2 *pf = nullptr, *pb = nullptr; float *csave = nullptr, *keep = nullptr; };
    std::vector<Bufs> B(nranks);
    unsigned *d_err = nullptr;
    int rc = PST_OK;
    auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == PST_OK) { pst_set_error("selftest_axis3_slabs: %s", cudaGetErrorString(e)); rc = PST_ECUDA; } };
    fail(cudaMalloc((void **)&d_err, 64));
    if (rc == PST_OK) fail(cudaMemsetAsync(d_err, 0, 64, c->stream));
    for (int r = 0; r < nranks && rc == PST_OK; r++) {
        fail(cudaMalloc((void **)&B[r].pf, L * sizeof(uint2)));
        fail(cudaMalloc((void **)&B[r].pb, L * sizeof(uint2)));
        fail(cudaMalloc((void **)&B[r].csave, (size_t)nch * L * sizeof(float)));
        fail(cudaMalloc((void **)&B[r].keep, (size_t)nb * L * sizeof(float)));
        if (rc == PST_OK) { fail(cudaMemsetAsync(B[r].pf, 0, L * sizeof(uint2), c->stream)); fail(cudaMemsetAsync(B[r].pb, 0, L * sizeof(uint2), c->stream)); }
    }
    auto args = [&](int r) {
        Tri3Args A{};
        const int z0 = (int)(((long)n3 * r) / nranks), z1 = (int)(((long)n3 * (r + 1)) / nranks);
        const bool first = r == 0, last = r == nranks - 1;
        A.x = d_x + (size_t)z0 * L; A.dst = d_x + (size_t)z0 * L; A.F = nullptr;
        A.hb = first ? nullptr : d_x + (size_t)(z0 - nb) * L;
        A.ha = last ? nullptr : d_x + (size_t)z1 * L;
        A.csave = B[r].csave; A.ha_keep = last ? nullptr : B[r].keep;
        A.L = L; A.l0 = 0; A.l1 = L; A.n3g = n3; A.z0 = z0; A.nz = z1 - z0; A.nb = nb;
        A.K0 = first ? 0 : z0 + nb; A.K1 = last ? n3 + 2 * nb : z1 + nb;
        A.wt = (float)(1.0 / ((double)nb * nb)); A.w2 = (float)(2. * A.wt);
        A.err = d_err; A.epoch = 1; A.rev = (r & 1);           // both tile orders
        return A;
    };
    for (int r = 0; r < nranks && rc == PST_OK; r++) {
        Tri3Args A = args(r);
        A.pin = r == 0 ? nullptr : B[r].pf; A.pout = r == nranks - 1 ? nullptr : B[r + 1].pf; A.pself = B[r].pb;
        rc = tri3_reg_launch(c, A, true, NZ, (size_t)A.nz * L, PST_K_TRI3);
    }
    for (int r = nranks - 1; r >= 0 && rc == PST_OK; r--) {
        Tri3Args A = args(r);
        A.pin = B[r].pb; A.pout = r == 0 ? nullptr : B[r - 1].pb;
        rc = tri3_reg_launch(c, A, false, NZ, (size_t)A.nz * L, PST_K_TRI3);
    }
    unsigned h_err = 0;
    if (rc == PST_OK) fail(cudaMemcpyAsync(&h_err, d_err, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    fail(cudaStreamSynchronize(c->stream));
    if (rc == PST_OK && h_err) { pst_set_error("selftest_axis3_slabs: a carry never arrived"); rc = PST_ECOMM; }
    for (auto &b : B) { cudaFree(b.pf); cudaFree(b.pb); cudaFree(b.csave); cudaFree(b.keep); }
    cudaFree(d_err);
    return rc;
}

extern "C" int pst_smooth3_dev(pst_ctx *c, float *d_x, int n1, int n2, int n3, int r1, int r2, int r3, int repeat, int adj)
{
    return pst_smoothcf_dev(c, d_x, n1, n2, n3, repeat, adj, r1, r2, r3, 0, 0, 0, 0, 0, 0);
}

extern "C" int pst_divne_dev(pst_ctx *c, float *d_num, float *d_den, float *d_rat, int n1, int n2, int n3,
                             int r1, int r2, int r3, int liter, int *iters_run)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    if (r1 < 1 || r2 < 1 || r3 < 1 || n1 < 1 || n2 < 1 || n3 < 1) { pst_set_error("divne: bad arguments"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    // distributed contexts: n3 is the GLOBAL plane count, the pointers are this rank's slab, the sums are all-reduced
    DipGeom g;
    size_t scr = 0;
    PST_TRY(slab_geom(c, "divne", n1, n2, n3, r1, r2, r3, &g, &scr));
    const size_t n = g.n;
    PST_TRY(pst_arena_reserve(c, (7 * n + scr + slab_halo_floats(g)) * sizeof(float) + 16 * 256));
    pst_arena_reset(c);
    PST_TRY(slab_halos(c, &g));
    CgWork w{};
    PST_TRY(pst_arena_get(c, n, &w.p));
    PST_TRY(pst_arena_get(c, n, &w.r));
    PST_TRY(pst_arena_get(c, n, &w.sp));
    PST_TRY(pst_arena_get(c, n, &w.sx));
    PST_TRY(pst_arena_get(c, n, &w.sr));
    PST_TRY(pst_arena_get(c, n, &w.gp));
    PST_TRY(pst_arena_get(c, n, &w.tmp));
    PST_TRY(pst_arena_get(c, scr, &w.scr));
    PST_TRY(pst_divne_run(c, g, d_num, d_den, d_rat, nullptr, w, liter, 1.0f, iters_run));
    if (g.dist) PST_TRY(pst_comm_check(c));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    return PST_OK;
}

// =======================================================================================
// soint3dc: structure-oriented interpolation by minimising the 3-D PWD residual with CG.
// Replaces (reference pyseistr/src/soint3d_cfuns.c) allpass3_lop :625-729, ps_cgstep :826-877,
// ps_solver :894-1174 (options actually used: "known", "x0") and the csoint3d driver :2405-2508.
// The adjoint, a scatter in the reference, is evaluated as a gather that adds the contributions
// of every target in the reference's own order (source index ascending, inline operator before
// xline), so the vectors stay bit-identical; the five CG dots are double tree reductions.
// =======================================================================================

// taps of the inline / xline B-filters at every voxel, stored tap-major ([w][N]) : the slopes are
// fixed during the solve, so passfilter (20 double multiplies) runs once per voxel, not per iteration
template <int NW>
__global__ void __launch_bounds__(256)
pwd3_taps_kernel(const float *__restrict__ pp, const float *__restrict__ qq, float *__restrict__ fi,
                 float *__restrict__ fx, size_t n, BTab tb)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float a[2 * NW + 1];
        pst_passfilter<NW>(tb, pp[i], a);
#pragma unroll
        for (int w = 0; w <= 2 * NW; w++) fi[(size_t)w * n + i] = a[w];
        pst_passfilter<NW>(tb, qq[i], a);
#pragma unroll
        for (int w = 0; w <= 2 * NW; w++) fx[(size_t)w * n + i] = a[w];
    }
}

// forward: y[0:N] (+)= inline PWD of x, y[N:2N] (+)= xline PWD.  DOTS: also the five dots of
// ps_cgstep (:853-863) of the freshly computed gg with Ss and rr.
template <int NW, bool ADD, bool DOTS>
__global__ void __launch_bounds__(256)
pwd3_fwd_kernel(const float *__restrict__ x, const float *__restrict__ fi, const float *__restrict__ fx,
                float *__restrict__ y, const float *__restrict__ Ss, const float *__restrict__ rr,
                int n1, int n2, int n3, int nj1, int nj2, int z0, int n3g, const float *__restrict__ xnext,
                unsigned gpp, int tg, double *__restrict__ partial)
{
    // n3 = planes held here (global planes [z0, z0 + n3) of n3g); xnext = plane z0 + n3 of x (next rank's first)
    const size_t n = (size_t)n1 * n2 * n3;
    const long pl = (long)n1 * n2;
    double acc[5] = {0., 0., 0., 0., 0.};
    // one block walks a group of tg traces of one plane (the pieces of the canonical sums, pst_common.cuh), its threads
    // walk i1: the indices come without a division per voxel
    const int i3 = (int)(blockIdx.x / gpp), t0 = (int)(blockIdx.x - (unsigned)i3 * gpp) * tg;
    const int t1 = t0 + tg < n2 ? t0 + tg : n2;
    for (int i2 = t0; i2 < t1; i2++) {
      const long tr = (long)i3 * n2 + i2;
      for (int i1 = threadIdx.x; i1 < n1; i1 += blockDim.x) {
        const size_t i = (size_t)tr * n1 + i1;
        float yi = ADD ? y[i] : 0.f, yx = ADD ? y[i + n] : 0.f;
        // shifts (w - NW) * nj on rows [NW*nj, n1 - NW*nj)  (allpass3_lop soint3d_cfuns.c:640-662,684-706)
        if (i1 >= NW * nj1 && i1 < n1 - NW * nj1 && i2 < n2 - 1) {
#pragma unroll
            for (int w = 0; w <= 2 * NW; w++) { const int s = (w - NW) * nj1; yi += (x[i + n1 + s] - x[i - s]) * fi[(size_t)w * n + i]; }
        }
        if (i1 >= NW * nj2 && i1 < n1 - NW * nj2 && z0 + i3 < n3g - 1) {
            const float *xu = (i3 < n3 - 1) ? x + i + pl : xnext + (i - (size_t)(n3 - 1) * pl);   // sample (i1, i2) of plane i3 + 1
#pragma unroll
            for (int w = 0; w <= 2 * NW; w++) { const int s = (w - NW) * nj2; yx += (xu[s] - x[i - s]) * fx[(size_t)w * n + i]; }
        }
        y[i] = yi;
        y[i + n] = yx;
        if (DOTS) {
            const float s0 = Ss[i], s1 = Ss[i + n], r0 = rr[i], r1 = rr[i + n];
            acc[0] += (double)yi * yi + (double)yx * yx;          // gg.gg
            acc[1] += (double)s0 * s0 + (double)s1 * s1;          // Ss.Ss
            acc[2] += (double)yi * s0 + (double)yx * s1;          // gg.Ss
            acc[3] += (double)yi * r0 + (double)yx * r1;          // gg.rr
            acc[4] += (double)s0 * r0 + (double)s1 * r1;          // Ss.rr
        }
      }
    }
    if (DOTS) pst_block_reduce<5>(acc, partial);
}

// adjoint as a gather, then the known-sample mask (ps_solver :1062-1066); partial of g.g
template <int NW>
__global__ void __launch_bounds__(256)
pwd3_adj_kernel(const float *__restrict__ yy, const float *__restrict__ fi, const float *__restrict__ fx,
                const unsigned char *__restrict__ known, float *__restrict__ g, int n1, int n2, int n3,
                int nj1, int nj2, int z0, int n3g, const float *__restrict__ yprev, const float *__restrict__ fxprev,
                unsigned gpp, int tg, double *__restrict__ partial)
{
    // yprev / fxprev: xline residual and xline taps ([w][plane]) of global plane z0 - 1 (previous rank's last)
    const size_t n = (size_t)n1 * n2 * n3;
    const long pl = (long)n1 * n2;
    double acc[1] = {0.};
    const int j3 = (int)(blockIdx.x / gpp), t0 = (int)(blockIdx.x - (unsigned)j3 * gpp) * tg;
    const int t1 = t0 + tg < n2 ? t0 + tg : n2;
    for (int j2 = t0; j2 < t1; j2++) {
      const long tr = (long)j3 * n2 + j2;
      for (int j1 = threadIdx.x; j1 < n1; j1 += blockDim.x) {
        const size_t j = (size_t)tr * n1 + j1;
        float v = 0.f;
        // inline operator: "+" targets of sources in trace j2-1 (source index ascending = shift descending)
        if (j2 >= 1) {
#pragma unroll
            for (int s = NW; s >= -NW; s--) {
                const int ix = j1 - s * nj1;
                if (ix >= NW * nj1 && ix < n1 - NW * nj1) { const size_t i = j - n1 - s * nj1; v += yy[i] * fi[(size_t)(s + NW) * n + i]; }
            }
        }
        if (j2 <= n2 - 2) {
#pragma unroll
            for (int s = -NW; s <= NW; s++) {
                const int ix = j1 + s * nj1;
                if (ix >= NW * nj1 && ix < n1 - NW * nj1) { const size_t i = j + s * nj1; v -= yy[i] * fi[(size_t)(s + NW) * n + i]; }
            }
        }
        // xline operator
        if (j3 >= 1) {
#pragma unroll
            for (int s = NW; s >= -NW; s--) {
                const int ix = j1 - s * nj2;
                if (ix >= NW * nj2 && ix < n1 - NW * nj2) { const size_t i = j - pl - s * nj2; v += yy[n + i] * fx[(size_t)(s + NW) * n + i]; }
            }
        } else if (z0 >= 1) {                                  // sources in the previous rank's last plane
#pragma unroll
            for (int s = NW; s >= -NW; s--) {
                const int ix = j1 - s * nj2;
                if (ix >= NW * nj2 && ix < n1 - NW * nj2) { const size_t i = j - s * nj2; v += yprev[i] * fxprev[(size_t)(s + NW) * pl + i]; }
            }
        }
        if (z0 + j3 <= n3g - 2) {
#pragma unroll
            for (int s = -NW; s <= NW; s++) {
                const int ix = j1 + s * nj2;
                if (ix >= NW * nj2 && ix < n1 - NW * nj2) { const size_t i = j + s * nj2; v -= yy[n + i] * fx[(size_t)(s + NW) * n + i]; }
            }
        }
        if (known[j]) v = 0.0f;
        g[j] = v;
        acc[0] += (double)v * v;
      }
    }
    pst_block_reduce<1>(acc, partial);
}

// ps_cgstep tail (:865-875): S = beta S + alfa g; Ss = beta Ss + alfa gg; x += S; rr += Ss; partial rr.rr
__global__ void __launch_bounds__(256)
cgstep_update_kernel(float *__restrict__ x, float *__restrict__ S, const float *__restrict__ g,
                     float *__restrict__ rr, float *__restrict__ Ss, const float *__restrict__ gg,
                     float alfa, float beta, Span Sp, double *__restrict__ partial)
{
    double acc[1] = {0.};
    const size_t n = Sp.n;
    size_t i0, i1, step;
    pst_span(Sp, 1, i0, i1, step);
    for (size_t i = i0; i < i1; i += step) {
        float s = S[i];
        s *= beta;
        s += alfa * g[i];
        S[i] = s;
        x[i] += s;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const size_t k = i + h * n;
            float t = Ss[k];
            t *= beta;
            t += alfa * gg[k];
            Ss[k] = t;
            const float r = rr[k] + t;
            rr[k] = r;
            acc[0] += (double)r * r;
        }
    }
    pst_block_reduce<1>(acc, partial);
}

__global__ void __launch_bounds__(256)
sumsq2_kernel(const float *__restrict__ v, Span Sp, double *__restrict__ partial)      // both components of v[2][n]
{
    double acc[1] = {0.};
    size_t i0, i1, step;
    pst_span(Sp, 1, i0, i1, step);
    for (size_t i = i0; i < i1; i += step) {
        acc[0] += (double)v[i] * v[i];
        acc[0] += (double)v[i + Sp.n] * v[i + Sp.n];
    }
    pst_block_reduce<1>(acc, partial);
}

__global__ void known_kernel(const float *__restrict__ ref, unsigned char *__restrict__ known, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        known[i] = (ref[i] != 0.f);
}

// Right-hand side noise of csoint3d (var > 0): a * N(0,1) drawn on the HOST exactly like the reference --
// MT19937 seeded per call (init_genrand / genrand_real1, soint3d_cfuns.c:2304-2370) feeding a Box-Muller pair
// generator that hands out the SECOND value of a pair first (ps_randn_one_bm :2372-2402) -- so that the same
// seed gives the same interpolation.  The draws are sequential by construction; they are produced in chunks and
// copied into rr as -d (ps_solver :1018-1024).
namespace {
struct NoiseGen {
    uint32_t mt[624]; int mti; bool have; float kept;
    explicit NoiseGen(uint32_t s) : mti(624), have(false), kept(0.f)
    {
        mt[0] = s;
        for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    }
    uint32_t u32()
    {
        if (mti >= 624) {
            for (int k = 0; k < 624; k++) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            mti = 0;
        }
        uint32_t y = mt[mti++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    float normal()
    {
        if (have) { have = false; return kept; }
        double x1, x2;
        do { x1 = u32() * (1.0 / 4294967295.0); } while (x1 == 0.0);
        x2 = u32() * (1.0 / 4294967295.0);
        const double z1 = sqrt(-2.0 * log(x1)), z2 = 2.0 * 3.14159265358979323846264338328 * x2;
        const double y1 = z1 * cos(z2), y2 = z1 * sin(z2);
        have = true; kept = (float)y1;
        return (float)y2;
    }
};
}  // namespace

// The stream is sequential over the WHOLE cube (2 * nglob draws: inline residual, then xline residual).  A rank of a
// distributed context draws all of it and keeps the two runs that fall on its slab [off, off + n) -- every rank sees
// the reference's sequence, at the price of drawing it nranks times.
static int soint3d_noise_rhs(pst_ctx *c, float *d_rr, size_t nglob, size_t off, size_t n, int seed, float var)
{
    NoiseGen G((uint32_t)(unsigned long)seed);
    const float a = sqrtf(var);
    const size_t chunk = (size_t)1 << 24;
    std::vector<float> h(std::min(chunk, n));
    size_t pos = 0;                                               // draws consumed so far
    for (int half = 0; half < 2; half++) {
        const size_t start = (size_t)half * nglob + off;
        for (; pos < start; pos++) (void)G.normal();
        for (size_t o = 0; o < n; o += chunk) {
            const size_t m = std::min(chunk, n - o);
            for (size_t i = 0; i < m; i++) { const float d = a * G.normal(); h[i] = -d; }
            pos += m;
            PST_CUDA(cudaMemcpyAsync(d_rr + (size_t)half * n + o, h.data(), m * sizeof(float), cudaMemcpyHostToDevice, c->stream));
            PST_CUDA(cudaStreamSynchronize(c->stream));
        }
    }
    return PST_OK;
}

template <int NW>
static int soint3d_run(pst_ctx *c, const float *d_din, const float *d_mask, const float *d_pp, const float *d_qq,
                       int n1, int n2, int n3, int nj1, int nj2, int niter, int seed, float var, int verb, float *d_out)
{
    static const BTab tb = make_btab(NW);
    constexpr int NA = 2 * NW + 1;
    // distributed contexts: n3 is the GLOBAL plane count, every pointer is this rank's n3-slab.  What crosses
    // slabs: plane z1 of the operand of L (next rank's first plane), plane z0-1 of the xline residual
    // for L' (previous rank's last plane; its taps once), and the five dots (all-reduced in pst_finish_reduce).
    const bool dist = c->comm != nullptr && c->nranks > 1;
    const int n3g = n3;
    int z0 = 0;
    if (dist) {
        int za = 0, zb = n3g;
        PST_TRY(pst_ctx_slab(c, n3g, &za, &zb));
        z0 = za; n3 = zb - za;
        if (n3 < 1) { pst_set_error("soint3d: empty slab"); return PST_EUNSUP; }
    }
    const size_t n = (size_t)n1 * n2 * n3, pln = (size_t)n1 * n2;
    PST_TRY(pst_arena_reserve(c, (size_t)(2 * NA + 8) * n * sizeof(float) + n + (NA + 4) * pln * sizeof(float) + 64 * 256));
    pst_arena_reset(c);
    float *fi, *fx, *g, *S, *rr, *gg, *Ss;
    float *xnext = nullptr, *yprev = nullptr, *fxprev = nullptr, *dum_lo = nullptr, *dum_hi = nullptr;
    unsigned char *known;
    PST_TRY(pst_arena_get(c, NA * n, &fi));
    PST_TRY(pst_arena_get(c, NA * n, &fx));
    PST_TRY(pst_arena_get(c, n, &g));
    PST_TRY(pst_arena_get(c, n, &S));
    PST_TRY(pst_arena_get(c, 2 * n, &rr));
    PST_TRY(pst_arena_get(c, 2 * n, &gg));
    PST_TRY(pst_arena_get(c, 2 * n, &Ss));
    PST_TRY(pst_arena_get(c, n, &known));
    if (dist) {
        PST_TRY(pst_arena_get(c, pln, &xnext));
        PST_TRY(pst_arena_get(c, pln, &yprev));
        PST_TRY(pst_arena_get(c, NA * pln, &fxprev));
        PST_TRY(pst_arena_get(c, pln, &dum_lo));
        PST_TRY(pst_arena_get(c, pln, &dum_hi));
    }
    // halo of the operand of L: my first plane goes to the previous rank, the next rank's first plane comes here
    auto halo_next = [&](const float *v) -> int {
        if (!dist) return PST_OK;
        return pst_comm_halo_exchange(c, v, v + (n - pln), dum_lo, xnext, pln);
    };
    // halo of the xline residual for L': my last plane goes to the next rank, the previous rank's last comes here
    auto halo_prev = [&](const float *v, float *into) -> int {
        if (!dist) return PST_OK;
        return pst_comm_halo_exchange(c, v, v + (n - pln), into, dum_hi, pln);
    };
    float *x = d_out;
    const int threads = 256, grid = pst_grid_for(c, n, threads);
    // canonical sums (pst_common.cuh): pieces = groups of tg traces of one plane (stencil kernels) / pst_red_ch() elements
    // of one plane (vector kernels); the CG scalars do not depend on the slab decomposition
    const unsigned ch = pst_red_ch(pln, n3g);
    const int tg = n1 >= (int)ch ? 1 : (int)(ch / (unsigned)n1);
    const unsigned gpp = (unsigned)((n2 + tg - 1) / tg), gridt = gpp * (unsigned)n3;
    const Span Sp = pst_span_canon(pln, n3, n3g);
    const unsigned gridc = Sp.ppp * (unsigned)n3;
    PST_TRY(pst_reserve_partials(c, gridt > gridc ? gridt : gridc, n3g));
    auto finish_t = [&](int nv, int rec) { return pst_finish_reduce_canon(c, (int)gpp, n3, z0, n3g, nv, rec); };
    auto finish_c = [&](int nv, int rec) { return pst_finish_reduce_canon(c, (int)Sp.ppp, n3, z0, n3g, nv, rec); };
    double h[PST_RED_SLOTS];
    PST_LAUNCHB(c, PST_K_OTHER, 48.0 * (double)n, (pwd3_taps_kernel<NW><<<grid, threads, 0, c->stream>>>(d_pp, d_qq, fi, fx, n, tb)));
    for (int w = 0; dist && w < NA; w++) PST_TRY(halo_prev(fx + (size_t)w * n, fxprev + (size_t)w * pln));   // once: taps are fixed
    PST_LAUNCH(c, PST_K_OTHER, (known_kernel<<<grid, threads, 0, c->stream>>>(d_mask ? d_mask : d_din, known, n)));
    // ps_solver :1018-1040 with dat = 0 (var = 0): rr = -0; x = x0 = data; rr += L x
    PST_CUDA(cudaMemcpyAsync(x, d_din, n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    if (var != 0.f) PST_TRY(soint3d_noise_rhs(c, rr, pln * (size_t)n3g, pln * (size_t)z0, n, seed, var));
    else PST_LAUNCH(c, PST_K_OTHER, (fill_kernel<<<grid, threads, 0, c->stream>>>(rr, -0.0f, 2 * n)));
    PST_TRY(halo_next(x));
    PST_LAUNCHB(c, PST_K_ALLPASS, 60.0 * (double)n,
                (pwd3_fwd_kernel<NW, true, false><<<gridt, threads, 0, c->stream>>>(x, fi, fx, rr, nullptr, nullptr, n1, n2, n3, nj1, nj2, z0, n3g, xnext, gpp, tg, c->d_partial)));
    PST_CUDA(cudaMemsetAsync(S, 0, n * sizeof(float), c->stream));
    PST_CUDA(cudaMemsetAsync(Ss, 0, 2 * n * sizeof(float), c->stream));
    // dpr0 = rr.rr (:1042)
    PST_LAUNCH(c, PST_K_OTHER, (sumsq2_kernel<<<gridc, threads, 0, c->stream>>>(rr, Sp, c->d_partial)));
    PST_TRY(finish_c(1, 8));
    PST_TRY(pst_fetch_record(c, 8, 1, h));
    const double dpr0 = h[0];
    double dpg0 = 1., rr2 = dpr0;
    bool first = true;
    for (int iter = 0; iter < niter; iter++) {
        PST_TRY(halo_prev(rr + n, yprev));
        PST_LAUNCHB(c, PST_K_ALLPASS, (8.0 * NA + 13.0) * (double)n,
                    (pwd3_adj_kernel<NW><<<gridt, threads, 0, c->stream>>>(rr, fi, fx, known, g, n1, n2, n3, nj1, nj2, z0, n3g, yprev, fxprev, gpp, tg, c->d_partial)));
        PST_TRY(finish_t(1, 8));
        PST_TRY(halo_next(g));
        PST_LAUNCHB(c, PST_K_ALLPASS, (8.0 * NA + 28.0) * (double)n,
                    (pwd3_fwd_kernel<NW, false, true><<<gridt, threads, 0, c->stream>>>(g, fi, fx, gg, Ss, rr, n1, n2, n3, nj1, nj2, z0, n3g, xnext, gpp, tg, c->d_partial)));
        PST_TRY(finish_t(5, 9));
        PST_TRY(pst_fetch_record(c, 8, 1, h));
        const double g2 = h[0];
        double dpr, dpg;
        if (iter == 0) { dpg0 = g2; dpr = 1.; dpg = 1.; }
        else { dpr = rr2 / dpr0; dpg = g2 / dpg0; }
        if (verb) printf("[pst] soint3d iteration %d res %g grad %g\n", iter + 1, dpr, dpg);
        if (dpr < 1.e-12 || dpg < 1.e-12) break;             // TOLERANCE (:889)
        PST_TRY(pst_fetch_record(c, 9, 5, h));
        const double gdg = h[0], sds = h[1], gds = h[2], ggr = h[3], ssr = h[4];
        double alfa, beta;
        if (first) {                                          // forget (:838-846)
            first = false;
            beta = 0.0;
            if (gdg <= 0.) continue;
            alfa = -ggr / gdg;
        } else {
            if (gdg == 0. || sds == 0.) continue;
            double determ = 1.0 - (gds / gdg) * (gds / sds);
            const double EPS = (double)1.e-12f;
            if (determ > EPS) determ *= gdg * sds; else determ = gdg * sds * EPS;
            const double gdr = -ggr, sdr = -ssr;
            alfa = (sds * gdr - gds * sdr) / determ;
            beta = (-gds * gdr + gdg * sdr) / determ;
        }
        PST_LAUNCHB(c, PST_K_CGDIR, 60.0 * (double)n,
                    (cgstep_update_kernel<<<gridc, threads, 0, c->stream>>>(x, S, g, rr, Ss, gg, (float)alfa, (float)beta, Sp, c->d_partial)));
        PST_TRY(finish_c(1, 10));
        PST_TRY(pst_fetch_record(c, 10, 1, h));
        rr2 = h[0];
        c->stats.cg_iterations++;
    }
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

extern "C" int pst_soint3d_dev(pst_ctx *c, const float *d_din, const float *d_mask, const float *d_dipi,
                               const float *d_dipx, int n1, int n2, int n3, int nw, int nj1, int nj2, int niter,
                               int drift, int seed, int hasmask, float var, int verb, float *d_out)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    if (n1 < 1 || n2 < 1 || n3 < 1 || niter < 0) { pst_set_error("soint3d: bad dimensions"); return PST_EINVAL; }
    if (nw != 1 && nw != 2) { pst_set_error("soint3d: order=%d unsupported (1 or 2)", nw); return PST_EUNSUP; }
    if (nj1 < 1 || nj2 < 1) { pst_set_error("soint3d: njs must be >= 1"); return PST_EINVAL; }
    if (n1 < 2 * nw * std::max(nj1, nj2) + 1) { pst_set_error("soint3d: n1 too short"); return PST_EINVAL; }
    if (drift != 0) { pst_set_error("soint3d: drift is not implemented on the GPU path (data-dependent scatter)"); return PST_EUNSUP; }
    if (var < 0.f) { pst_set_error("soint3d: var < 0"); return PST_EINVAL; }
    if (hasmask && !d_mask) { pst_set_error("soint3d: hasmask=1 needs a mask"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    const float *m = hasmask ? d_mask : nullptr;
    if (nw == 1) return soint3d_run<1>(c, d_din, m, d_dipi, d_dipx, n1, n2, n3, nj1, nj2, niter, seed, var, verb, d_out);
    return soint3d_run<2>(c, d_din, m, d_dipi, d_dipx, n1, n2, n3, nj1, nj2, niter, seed, var, verb, d_out);
}
