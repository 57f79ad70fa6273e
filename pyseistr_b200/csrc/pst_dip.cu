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

#include <math.h>

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
// One block walks traces (grid-stride), threads walk i1: fully coalesced.
template <int NW, bool DER, bool LS>
__global__ void __launch_bounds__(256)
allpass_kernel(const float *__restrict__ u, const float *__restrict__ p_in,
               const float *__restrict__ dp, float lam, float *__restrict__ p_out,
               float *__restrict__ y, int n1, int n2, int n3, int xline, BTab tb,
               double *__restrict__ partial)
{
    const long ntr = (long)n2 * n3;
    const long ip = xline ? (long)n1 * n2 : (long)n1;
    double acc2[1] = {0.0};
    for (long tr = blockIdx.x; tr < ntr; tr += gridDim.x) {
        const int i2 = (int)(tr % n2), i3 = (int)(tr / n2);
        const bool live_tr = xline ? (i3 < n3 - 1) : (i2 < n2 - 1);
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
                            unsigned char *__restrict__ m_x, int n1, int n2, int n3)
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
                if (i3 < n3 - 1)
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
// G3/G4: streaming kernels of ps_divne and ps_conjgrad with fused double reductions.

// divne step 1 (:802-809): optional mask zeroing (dip3 :1656-1663), num,den *= 1/hypot(den,eps)
// in double; partial of sum(den^2) (:811).
__global__ void __launch_bounds__(256)
divne_prescale_kernel(float *__restrict__ num, float *__restrict__ den,
                      const unsigned char *__restrict__ mask, float eps, size_t n,
                      double *__restrict__ partial)
{
    double acc[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
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
                        size_t n, double *__restrict__ partial)
{
    double acc[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
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

// gp = eps*p + S(gx)  (:303,:318); tmp <- gp as the input of the second shaping; partial gp.gp
__global__ void __launch_bounds__(256)
cg_gp_kernel(const float *__restrict__ p, float *__restrict__ tmp, float *__restrict__ gp,
             float eps, size_t n, double *__restrict__ partial)
{
    double acc[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float g = eps * p[i];
        g += tmp[i];
        gp[i] = g;
        tmp[i] = g;
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

__global__ void fill_kernel(float *__restrict__ x, float v, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = v;
}

// =======================================================================================
// host side
// =======================================================================================

struct DipGeom { int n1, n2, n3, r1, r2, r3; size_t n; };

static size_t tri_scratch_floats(const DipGeom &g)
{
    // largest extended volume of the strided-axis sweeps (+ axis 1 when it falls back)
    size_t a1 = (size_t)(g.n1 + 2 * (size_t)g.r1) * g.n2 * g.n3;
    size_t a2 = (size_t)g.n1 * (g.n2 + 2 * (size_t)g.r2) * g.n3;
    size_t a3 = (size_t)g.n1 * g.n2 * (g.n3 + 2 * (size_t)g.r3);
    size_t m = a1 > a2 ? a1 : a2;
    return m > a3 ? m : a3;
}

static int tri_lines_launch(pst_ctx *c, int cls, float *x, float *scr, long nlines, long na, long sa, long sb,
                            long d, int nx, int nb)
{
    const float wt = (float)(1.0 / ((double)nb * nb));       // ps_triangle_init :421
    const float w2 = (float)(2. * wt);
    const int threads = 128;
    const long blocks = (nlines + threads - 1) / threads;
    PST_LAUNCH(c, cls,
        if (nb <= nx)
            tri_lines_kernel<<<(unsigned)blocks, threads, 0, c->stream>>>(x, scr, nlines, na, sa, sb, d, nx, nb, wt, w2);
        else
            tri_lines_literal_kernel<<<(unsigned)blocks, threads, 0, c->stream>>>(x, scr, nlines, na, sa, sb, d, nx, nb, wt, w2));
    c->stats.smooth_passes++;
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

// ps_trianglen_lop body (:692-700): axes 1, 2, 3 in turn, in place
int pst_smooth3_inplace(pst_ctx *c, float *x, float *scr, int n1, int n2, int n3, int r1, int r2, int r3)
{
    if (r1 > 1) {
        const long nlines = (long)n2 * n3;
        const int np = n1 + 2 * r1;
        const int pitch = (np % 2) ? np : np + 1;
        constexpr int LPC = 16;
        const size_t smem = (size_t)LPC * pitch * sizeof(float);
        if (r1 <= n1 && smem <= 200 * 1024) {
            const float wt = (float)(1.0 / ((double)r1 * r1));
            const float w2 = (float)(2. * wt);
            static bool attr_done = false;
            if (!attr_done) {
                PST_CUDA(cudaFuncSetAttribute(tri_axis1_kernel<LPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr_done = true;
            }
            long blocks = (nlines + LPC - 1) / LPC;
            const long cap = (long)c->sm_count * 16;
            if (blocks > cap) blocks = cap;
            PST_LAUNCH(c, PST_K_TRI1, (tri_axis1_kernel<LPC><<<(unsigned)blocks, 128, smem, c->stream>>>(x, nlines, n1, r1, wt, w2, pitch)));
            c->stats.smooth_passes++;
            PST_CUDA(cudaGetLastError());
        } else {
            PST_TRY(tri_lines_launch(c, PST_K_TRI1, x, scr, nlines, nlines, n1, 0, 1, n1, r1));
        }
    }
    if (r2 > 1) PST_TRY(tri_lines_launch(c, PST_K_TRI2, x, scr, (long)n1 * n3, n1, 1, (long)n1 * n2, n1, n2, r2));
    if (r3 > 1) PST_TRY(tri_lines_launch(c, PST_K_TRI3, x, scr, (long)n1 * n2, (long)n1 * n2, 1, 0, (long)n1 * n2, n3, r3));
    return PST_OK;
}

template <int NW>
static int allpass_launch_nw(pst_ctx *c, const float *u, const float *p_in, const float *dp, float lam,
                             float *p_out, float *y, int n1, int n2, int n3, int xline, bool der,
                             bool ls, int rec)
{
    static const BTab tb = make_btab(NW);
    const long ntr = (long)n2 * n3;
    long blocks = ntr;
    const long cap = (long)c->sm_count * 8;
    if (blocks > cap) blocks = cap;
    const int threads = n1 >= 256 ? 256 : (n1 >= 128 ? 128 : 64);
    PST_LAUNCH(c, PST_K_ALLPASS,
        if (ls)
            allpass_kernel<NW, false, true><<<(unsigned)blocks, threads, 0, c->stream>>>(u, p_in, dp, lam, p_out, y, n1, n2, n3, xline, tb, c->d_partial);
        else if (der)
            allpass_kernel<NW, true, false><<<(unsigned)blocks, threads, 0, c->stream>>>(u, p_in, nullptr, 0.f, nullptr, y, n1, n2, n3, xline, tb, c->d_partial);
        else
            allpass_kernel<NW, false, false><<<(unsigned)blocks, threads, 0, c->stream>>>(u, p_in, nullptr, 0.f, nullptr, y, n1, n2, n3, xline, tb, c->d_partial));
    PST_CUDA(cudaGetLastError());
    return pst_finish_reduce(c, (int)blocks, 1, rec);
}

int pst_allpass_launch(pst_ctx *c, const float *u, const float *p_in, const float *dp, float lam,
                       float *p_out, float *y, int n1, int n2, int n3, int nw, int xline, bool der,
                       bool ls, int rec)
{
    if (nw == 1) return allpass_launch_nw<1>(c, u, p_in, dp, lam, p_out, y, n1, n2, n3, xline, der, ls, rec);
    if (nw == 2) return allpass_launch_nw<2>(c, u, p_in, dp, lam, p_out, y, n1, n2, n3, xline, der, ls, rec);
    pst_set_error("order=%d unsupported (1 or 2)", nw);
    return PST_EUNSUP;
}

struct CgWork {
    float *p, *r, *sp, *sx, *sr, *gp, *tmp, *scr;
};

// ps_divne (:796-827) + ps_conjgrad (:257-383, prec=NULL, hasp0=false, eps=1*1, tol=1e-6).
// num/den are overwritten (den becomes the weight); rat receives the CG model x.
int pst_divne_run(pst_ctx *c, const DipGeom &g, float *num, float *den, float *rat,
                  const unsigned char *mask, const CgWork &w, int liter, float eps_div, int *iters_run)
{
    const size_t n = g.n;
    const float eps = 1.f * 1.f, tol = 1.e-6f;
    const int threads = 256;
    const int grid = pst_grid_for(c, n, threads);
    double h[PST_RED_SLOTS];
    if (iters_run) *iters_run = 0;

    PST_LAUNCH(c, PST_K_CGVEC, (divne_prescale_kernel<<<grid, threads, 0, c->stream>>>(num, den, mask, eps_div, n, c->d_partial)));
    PST_TRY(pst_finish_reduce(c, grid, 1, 0));
    PST_TRY(pst_fetch_record(c, 0, 1, h));
    if (h[0] == 0.0) {
        PST_LAUNCH(c, PST_K_OTHER, (fill_kernel<<<grid, threads, 0, c->stream>>>(rat, 0.f, n)));
        return PST_OK;
    }
    const double norm = sqrt((double)n / h[0]);
    PST_LAUNCH(c, PST_K_CGVEC, (divne_scale_init_kernel<<<grid, threads, 0, c->stream>>>(num, den, norm, w.r, w.p, rat, n, c->d_partial)));
    PST_TRY(pst_finish_reduce(c, grid, 1, 0));
    PST_TRY(pst_fetch_record(c, 0, 1, h));
    if (h[0] == 0.0) return PST_OK;               // zero residual: p = x = 0 (:299-303)

    double gn = 0., gnp = 0., g0 = 0., alpha, beta;
    float a_pending = 0.f;
    bool pending = false;
    int iter;
    for (iter = 0; iter < liter; iter++) {
        PST_LAUNCH(c, PST_K_CGVEC,
            if (pending)
                cg_head_kernel<true><<<grid, threads, 0, c->stream>>>(w.p, rat, w.r, w.sp, w.sx, w.sr, den, a_pending, eps, w.tmp, n);
            else
                cg_head_kernel<false><<<grid, threads, 0, c->stream>>>(w.p, rat, w.r, w.sp, w.sx, w.sr, den, 0.f, eps, w.tmp, n));
        pending = false;
        PST_TRY(pst_smooth3_inplace(c, w.tmp, w.scr, g.n1, g.n2, g.n3, g.r1, g.r2, g.r3));
        PST_LAUNCH(c, PST_K_CGVEC, (cg_gp_kernel<<<grid, threads, 0, c->stream>>>(w.p, w.tmp, w.gp, eps, n, c->d_partial)));
        PST_TRY(pst_finish_reduce(c, grid, 1, 1));
        PST_TRY(pst_smooth3_inplace(c, w.tmp, w.scr, g.n1, g.n2, g.n3, g.r1, g.r2, g.r3));
        PST_TRY(pst_fetch_record(c, 1, 1, h));
        gn = h[0];
        if (iter == 0) {
            g0 = gn;
            PST_LAUNCH(c, PST_K_CGVEC, (cg_dir_kernel<true><<<grid, threads, 0, c->stream>>>(w.gp, w.tmp, den, w.sp, w.sx, w.sr, 0.f, n, c->d_partial)));
        } else {
            alpha = gn / gnp;
            const double dg = gn / g0;
            if (alpha < tol || dg < tol) break;
            PST_LAUNCH(c, PST_K_CGVEC, (cg_dir_kernel<false><<<grid, threads, 0, c->stream>>>(w.gp, w.tmp, den, w.sp, w.sx, w.sr, (float)alpha, n, c->d_partial)));
        }
        PST_TRY(pst_finish_reduce(c, grid, 3, 2));
        PST_TRY(pst_fetch_record(c, 2, 3, h));
        beta = h[0] + (double)eps * (h[1] - h[2]);
        alpha = -gn / beta;
        a_pending = (float)alpha;
        pending = true;
        gnp = gn;
        c->stats.cg_iterations++;
    }
    if (pending) {      // only the model x (= rat) is consumed after the last iteration
        PST_LAUNCH(c, PST_K_CGVEC, (cg_tail_kernel<<<grid, threads, 0, c->stream>>>(rat, w.sx, a_pending, n)));
    }
    if (iters_run) *iters_run = iter;
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

// dip3 (:1619-1691) for one direction.  p holds the initial dip (zeros) and receives the result.
static int gauss_newton(pst_ctx *c, const DipGeom &g, const float *u, float *p, const unsigned char *mask,
                        int xline, int niter, int liter, int nw, float *u1, float *u2, float *dp,
                        float *ptrial, const CgWork &w, int verb)
{
    double h[PST_RED_SLOTS];
    float *pcur = p, *pnext = ptrial;
    PST_TRY(pst_allpass_launch(c, u, pcur, nullptr, 0.f, nullptr, u2, g.n1, g.n2, g.n3, nw, xline, false, false, 3));
    PST_TRY(pst_fetch_record(c, 3, 1, h));
    double usum = h[0];
    for (int iter = 0; iter < niter; iter++) {
        PST_TRY(pst_allpass_launch(c, u, pcur, nullptr, 0.f, nullptr, u1, g.n1, g.n2, g.n3, nw, xline, true, false, 4));
        int its = 0;
        PST_TRY(pst_divne_run(c, g, u2, u1, dp, mask, w, liter, 1.0f, &its));
        float lam = 1.f;
        double usum2 = 0.;
        int k;
        for (k = 0; k < 8; k++) {
            PST_TRY(pst_allpass_launch(c, u, pcur, dp, lam, pnext, u2, g.n1, g.n2, g.n3, nw, xline, false, true, 3));
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

extern "C" int pst_dip_dev(pst_ctx *c, const float *d_din, const float *d_mask, int n1, int n2, int n3,
                           int niter, int liter, int order, int r1, int r2, int r3, int verb,
                           float *d_dip_out)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_TRY(check_dip_args(n1, n2, n3, niter, liter, order, r1, r2, r3));
    PST_CUDA(cudaSetDevice(c->device));
    DipGeom g{n1, n2, n3, r1, r2, r3, (size_t)n1 * n2 * n3};
    const size_t n = g.n;
    const size_t scr = tri_scratch_floats(g);
    const size_t need = (11 * n + scr) * sizeof(float) + 2 * n + 16 * 256;
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
    if (d_mask) {
        PST_TRY(pst_arena_get(c, n, &m_in));
        PST_TRY(pst_arena_get(c, n, &m_x));
        const int grid = (int)min((long)n2 * n3, (long)c->sm_count * 8);
        PST_LAUNCH(c, PST_K_OTHER,
            if (order == 1) mask_kernel<1><<<grid, 128, 0, c->stream>>>(d_mask, m_in, m_x, n1, n2, n3);
            else            mask_kernel<2><<<grid, 128, 0, c->stream>>>(d_mask, m_in, m_x, n1, n2, n3));
    }
    const int ndip = (n3 == 1) ? 1 : 2;
    PST_CUDA(cudaMemsetAsync(d_dip_out, 0, ndip * n * sizeof(float), c->stream));
    PST_TRY(gauss_newton(c, g, d_din, d_dip_out, m_in, 0, niter, liter, order, u1, u2, dp, ptrial, w, verb));
    if (ndip == 2)
        PST_TRY(gauss_newton(c, g, d_din, d_dip_out + n, m_x, 1, niter, liter, order, u1, u2, dp, ptrial, w, verb));
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

extern "C" int pst_allpass_dev(pst_ctx *c, const float *d_u, const float *d_sigma, int n1, int n2, int n3,
                               int order, int xline, int der, float *d_y)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    if (n1 < 2 * order + 1) { pst_set_error("allpass: n1 too short"); return PST_EINVAL; }
    PST_TRY(pst_allpass_launch(c, d_u, d_sigma, nullptr, 0.f, nullptr, d_y, n1, n2, n3, order, xline, der != 0, false, 5));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    return PST_OK;
}

extern "C" int pst_smooth3_dev(pst_ctx *c, float *d_x, int n1, int n2, int n3, int r1, int r2, int r3)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    if (r1 < 1 || r2 < 1 || r3 < 1 || n1 < 1 || n2 < 1 || n3 < 1) { pst_set_error("smooth3: bad arguments"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    DipGeom g{n1, n2, n3, r1, r2, r3, (size_t)n1 * n2 * n3};
    const size_t scr = tri_scratch_floats(g);
    PST_TRY(pst_arena_reserve(c, scr * sizeof(float) + 4096));
    pst_arena_reset(c);
    float *s;
    PST_TRY(pst_arena_get(c, scr, &s));
    PST_TRY(pst_smooth3_inplace(c, d_x, s, n1, n2, n3, r1, r2, r3));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    return PST_OK;
}

extern "C" int pst_divne_dev(pst_ctx *c, float *d_num, float *d_den, float *d_rat, int n1, int n2, int n3,
                             int r1, int r2, int r3, int liter, int *iters_run)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    if (r1 < 1 || r2 < 1 || r3 < 1 || n1 < 1 || n2 < 1 || n3 < 1) { pst_set_error("divne: bad arguments"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    DipGeom g{n1, n2, n3, r1, r2, r3, (size_t)n1 * n2 * n3};
    const size_t n = g.n, scr = tri_scratch_floats(g);
    PST_TRY(pst_arena_reserve(c, (7 * n + scr) * sizeof(float) + 16 * 256));
    pst_arena_reset(c);
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
    PST_CUDA(cudaStreamSynchronize(c->stream));
    return PST_OK;
}
