// Plane-wave spray + structure-oriented mean / median on sm_100a.
//
// Replaces (reference pyseistr/src/sof3d_cfuns.c): sf_banded_define/solve :159-264,
// pwd_define/pwd_set :401-498, regularization/predict1_step/predict2_step :548-661,
// update_init/get_update :980-1054, the spray loops of csomean3d :1455-1521 and csomf3d
// :1639-1705, the mean :1523-1537 and mf/sf_quantile :1058-1252,:1707-1736; and
// (pyseistr/src/sof_cfuns.c) pwspray_lop :948-1031, pwsmooth_lop/set :1076-1132,
// csomean2d :1433-1532, csomf2d :1534-1672.
//
// Formulation.  The reference sprays source-by-source into a materialised stack
// u[n2*n3][np][n1].  Here the same values are produced target-by-target and level-by-level:
// slot (a,b) at location j holds the prediction of trace j from source j-(a,b); it is one
// implicit plane-wave prediction from the slot(s) one step closer to the source
// (SURVEY A.8): (a-sgn a, b) at j-sgn a with the inline dip, (a, b-sgn b) at j-sgn b*n2 with
// the xline dip.  All targets of one slot are independent: one thread per target trace
// runs the serial banded LDL' factor/solve along n1.  Volumes are held "trace-minor"
// ([i3][i1][i2], i2 fastest) so that the 32 lanes of a warp walk n1 in lock step with
// coalesced accesses.  Only the slots the output needs are computed.
// Arithmetic follows the reference operation by operation (no FMA), so every sprayed
// value is bit-identical and the median selection is exact.
#include "pst_common.cuh"
#include "pst_predict_core.h"
#include "pst_predict_warp.h"

#include <math.h>
#include <stdlib.h>
#include <algorithm>

#define PST_MAXSLOT 81

// Budget of the per-chunk slot + factor-scratch buffers.  One thread runs one target trace, so a chunk must
// hold enough planes to fill the GPU (148 SMs x 1536+ threads): at 1000x1024x1024 the first session's 6 GB gave
// 85 planes = 87 K threads (29 % of capacity, ncu: long-scoreboard bound); 24 GB gives 340 planes.
static double spray_chunk_bytes()
{
    static const double v = []() { const char *e = getenv("PST_SPRAY_CHUNK_GB"); const double g = e ? atof(e) : 24.0; return (g > 0.5 ? g : 0.5) * 1.0e9; }();
    return v;
}


// One thread = one target trace.  Forward sweep: taps -> W'W bands (+regularisation) ->
// LDL' column k -> rhs -> forward substitution; (d, o[0..NB), b) go to scratch.  Backward
// sweep: back substitution, result stored to the slot volume.
template <int NW, bool TWO>
__global__ void __launch_bounds__(128)
predict_kernel(const PredArgs A)
{
    constexpr int NA = 2 * NW + 1, NB = 2 * NW, NC = NB + 1;
    const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int zl = A.zla + blockIdx.y;
    if (i2 >= A.n2) return;
    const int n1 = A.n1, n2 = A.n2;
    const long base = (long)zl * n1 * n2 + i2;
    float *out = A.out + base;
    // slot stays zero when its source lies outside the cube (csomf3d :1656)
    // (a source outside the LOCAL trace range of a cut panel feeds only halo targets nobody reads: zero as well)
    const int s2 = i2 - A.a, s3 = (A.ze0 + zl) - A.b, s2g = s2 + A.t_off;
    if (s2 < 0 || s2 >= n2 || s2g < 0 || s2g >= A.ntg || s3 < 0 || s3 >= A.n3) {
        for (int k = 0; k < n1; k++) out[(long)k * n2] = 0.f;
        return;
    }
    const float *x1 = A.in1 + base + A.in1_off;
    const float *g1 = A.sg1 + base + A.sg1_off;
    const float *x2 = TWO ? A.in2 + base + A.in2_off : nullptr;
    const float *g2 = TWO ? A.sg2 + base + A.sg2_off : nullptr;
    float *scr = A.scr + ((long)blockIdx.y * n1 * NC) * n2 + i2;
    const bool f1 = A.forw1 != 0, f2 = A.forw2 != 0;
    const RegC rg = A.reg;

    // sliding windows, index c <-> sample i-NW+c
    float W1[NA][NA], T1[NA], X1[NA];
    float W2[TWO ? NA : 1][NA], T2[TWO ? NA : 1], X2[TWO ? NA : 1];
    float O[NB][NB], D[NB], Bh[NB];        // history: index h <-> sample k-1-h
#pragma unroll
    for (int c = 0; c < NA; c++) {
        T1[c] = 0.f; X1[c] = 0.f;
        if (TWO) { T2[c] = 0.f; X2[c] = 0.f; }
#pragma unroll
        for (int j = 0; j < NA; j++) { W1[c][j] = 0.f; if (TWO) W2[c][j] = 0.f; }
    }
#pragma unroll
    for (int h = 0; h < NB; h++) {
        D[h] = 0.f; Bh[h] = 0.f;
#pragma unroll
        for (int m = 0; m < NB; m++) O[h][m] = 0.f;
    }
    // Invariant at the top of step i (leading sample kk = i + NW):  X[c] = inp[kk - NW + c] = inp[i + c].
    // The loop starts at i = -NW so the windows fill before column 0; entries with a negative
    // sample index are never used (tmp is only formed for kk >= NW).
    {
        float t1[NA], t2[NA];
#pragma unroll
        for (int c = 0; c < NA; c++) {
            const int idx = c - NW;
            t1[c] = (idx >= 0 && idx < n1) ? x1[(long)idx * n2] : 0.f;
            if (TWO) t2[c] = (idx >= 0 && idx < n1) ? x2[(long)idx * n2] : 0.f;
        }
#pragma unroll
        for (int c = 0; c < NA; c++) { X1[c] = t1[c]; if (TWO) X2[c] = t2[c]; }
    }

    // Loads run PF steps ahead of their use (register queues, loop unrolled by PF so the queue index is static):
    // one thread owns one trace, so without this every step waits for its own slope / input sample
    // (ncu, first version: 10 long-scoreboard stall cycles per issued instruction).
    constexpr int PF = 4;
    float gq1[PF], gq2[PF], xq1[PF], xq2[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) {
        const int kk = u;                                   // step i = -NW + u uses slope sample kk = i + NW = u
        gq1[u] = (kk < n1) ? g1[(long)kk * n2] : 0.f;
        gq2[u] = (TWO && kk < n1) ? g2[(long)kk * n2] : 0.f;
        const int idx = kk + 1 + NW;                        // and feeds input sample (kk + 1) + NW to the window
        xq1[u] = (idx < n1) ? x1[(long)idx * n2] : 0.f;
        xq2[u] = (TWO && idx < n1) ? x2[(long)idx * n2] : 0.f;
    }
    for (int i0 = -NW; i0 < n1; i0 += PF) {
#pragma unroll
      for (int u = 0; u < PF; u++) {
        const int i = i0 + u;
        if (i >= n1) break;
        // ---- leading sample kk = i + NW enters the windows at c = NA-1
        const int kk = i + NW;
        const float gv1 = gq1[u], gv2 = gq2[u], xn1 = xq1[u], xn2 = xq2[u];
        {                                                   // refill the queue slot for step i + PF
            const int kn = kk + PF;
            gq1[u] = (kn < n1) ? g1[(long)kn * n2] : 0.f;
            if (TWO) gq2[u] = (kn < n1) ? g2[(long)kn * n2] : 0.f;
            const int idn = kn + 1 + NW;
            xq1[u] = (idn < n1) ? x1[(long)idn * n2] : 0.f;
            if (TWO) xq2[u] = (idn < n1) ? x2[(long)idn * n2] : 0.f;
        }
        {
            float a1[NA], a2[NA];
            float tm1 = 0.f, tm2 = 0.f;
            if (kk < n1) {
                spray_taps<NW>(A.tb, gv1, f1, a1);
                if (TWO) spray_taps<NW>(A.tb, gv2, f2, a2);
                if (kk >= NW && kk < n1 - NW) {               // pwd_set :481-486
#pragma unroll
                    for (int j = 0; j < NA; j++) {
                        tm1 += a1[j] * X1[j];
                        if (TWO) tm2 += a2[j] * X2[j];
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < NA; j++) { a1[j] = 0.f; a2[j] = 0.f; }
            }
#pragma unroll
            for (int j = 0; j < NA; j++) { W1[NA - 1][j] = a1[j]; if (TWO) W2[NA - 1][j] = a2[j]; }
            T1[NA - 1] = tm1;
            if (TWO) T2[NA - 1] = tm2;
        }
        if (i >= 0) {
            // ---- matrix column i: regularisation + W'W (pwd_define :426-444)
            float dg = rg.d_in;
            if (i == 0 || i == n1 - 1) dg = rg.d_e0;
            if (i == 1 || i == n1 - 2) dg = rg.d_e1;
            float of[NB];
            of[0] = (i == 0 || i == n1 - 2) ? rg.o0_e : rg.o0_in;
            of[1] = rg.o1;
#pragma unroll
            for (int m = 2; m < NB; m++) of[m] = 0.0f;
            float rhs1 = 0.f, rhs2 = 0.f;
#pragma unroll
            for (int j = 0; j < NA; j++) {
                const int k = i + j - NW;
                if (k >= NW && k < n1 - NW) { const float aj = W1[j][j]; dg += aj * aj; }
            }
#pragma unroll
            for (int m = 0; m < NB; m++) {
#pragma unroll
                for (int j = m + 1; j < NA; j++) {
                    const int k = i + j - NW;
                    if (k >= NW && k < n1 - NW) of[m] += W1[j][j - m - 1] * W1[j][j];
                }
            }
            if (TWO) {
#pragma unroll
                for (int j = 0; j < NA; j++) {
                    const int k = i + j - NW;
                    if (k >= NW && k < n1 - NW) { const float aj = W2[j][j]; dg += aj * aj; }
                }
#pragma unroll
                for (int m = 0; m < NB; m++) {
#pragma unroll
                    for (int j = m + 1; j < NA; j++) {
                        const int k = i + j - NW;
                        if (k >= NW && k < n1 - NW) of[m] += W2[j][j - m - 1] * W2[j][j];
                    }
                }
            }
            // ---- rhs (pwd_set :487-496), end terms (predict1/2_step :611-619,:640-658)
#pragma unroll
            for (int j = 0; j < NA; j++) {
                const int k = i + j - NW;
                if (k >= NW && k < n1 - NW) {
                    rhs1 += W1[j][j] * T1[j];
                    if (TWO) rhs2 += W2[j][j] * T2[j];
                }
            }
            float rhs = TWO ? (rhs1 + rhs2) : rhs1;
            if (i < 2 || i >= n1 - 2) {
                // X1[0] = inp[kk - NW] = inp[i]
                float te;
                if (TWO) te = (float)(0.5 * (double)(X1[0] + X2[0]));
                else te = X1[0];
                rhs += rg.eps2 * te;
            }
            // ---- LDL' column (sf_banded_define :169-184)
            float t = dg;
#pragma unroll
            for (int m = 0; m < NB; m++)
                if (m < i) t -= (O[m][m] * O[m][m]) * D[m];
            const float dk = t;
            float ok[NB];
#pragma unroll
            for (int q = 0; q < NB; q++) {
                float v = of[q];
#pragma unroll
                for (int m = 0; m < NB - q - 1; m++)
                    if (m < i) v -= (O[m][m] * O[m][q + m + 1]) * D[m];
                ok[q] = (q < n1 - i - 1) ? v / dk : 0.f;
            }
            // ---- forward substitution (sf_banded_solve :250-256)
            float bk = rhs;
#pragma unroll
            for (int m = 0; m < NB; m++)
                if (m < i) bk -= O[m][m] * Bh[m];
            // ---- spill column to scratch: the back substitution needs b_k / d_k (formed here: same operands, same
            // division) and o[.][k], not d_k and b_k separately
            float *sc = scr + (long)i * NC * n2;
            sc[0] = bk / dk;
#pragma unroll
            for (int q = 0; q < NB; q++) sc[(long)(1 + q) * n2] = ok[q];
            // ---- shift LDL history
#pragma unroll
            for (int h = NB - 1; h > 0; h--) {
                D[h] = D[h - 1]; Bh[h] = Bh[h - 1];
#pragma unroll
                for (int m = 0; m < NB; m++) O[h][m] = O[h - 1][m];
            }
            D[0] = dk; Bh[0] = bk;
#pragma unroll
            for (int m = 0; m < NB; m++) O[0][m] = ok[m];
        }
        // ---- shift sample windows; new trailing input for the next leading sample
#pragma unroll
        for (int c = 0; c < NA - 1; c++) {
            T1[c] = T1[c + 1]; X1[c] = X1[c + 1];
            if (TWO) { T2[c] = T2[c + 1]; X2[c] = X2[c + 1]; }
#pragma unroll
            for (int j = 0; j < NA; j++) { W1[c][j] = W1[c + 1][j]; if (TWO) W2[c][j] = W2[c + 1][j]; }
        }
        X1[NA - 1] = xn1;                      // X[NA-1] for the next step = inp[(kk+1) + NW] (0 beyond the trace)
        if (TWO) X2[NA - 1] = xn2;
      }
    }

    // ---- back substitution (sf_banded_solve :257-263)
    float Y[NB];
#pragma unroll
    for (int m = 0; m < NB; m++) Y[m] = 0.f;
    constexpr int PB = 4;                                   // factor columns in flight ahead of the recurrence
    float cq[PB][NC];
#pragma unroll
    for (int u = 0; u < PB; u++) {
        const int k = n1 - 1 - u;
#pragma unroll
        for (int q = 0; q < NC; q++) cq[u][q] = (k >= 0) ? scr[((long)k * NC + q) * n2] : 1.f;
    }
    for (int k0 = n1 - 1; k0 >= 0; k0 -= PB) {
#pragma unroll
      for (int u = 0; u < PB; u++) {
        const int k = k0 - u;
        if (k < 0) break;
        float col[NC];
#pragma unroll
        for (int q = 0; q < NC; q++) col[q] = cq[u][q];
        {
            const int kn = k - PB;
#pragma unroll
            for (int q = 0; q < NC; q++) cq[u][q] = (kn >= 0) ? scr[((long)kn * NC + q) * n2] : 1.f;
        }
        float t = col[0];
#pragma unroll
        for (int m = 0; m < NB; m++)
            if (m < n1 - k - 1) t -= col[1 + m] * Y[m];
        out[(long)k * n2] = t;
#pragma unroll
        for (int m = NB - 1; m > 0; m--) Y[m] = Y[m - 1];
        Y[0] = t;
      }
    }
}

// predict_fast_kernel: see pst_predict_core.h (the per-trace body is a host/device header so that
// tests/test_predict_core.py can run it on the host against the CPU restatement)
template <int NW, bool TWO, int MINB>
__global__ void __launch_bounds__(128, MINB)
predict_fast_kernel(const PredArgs A)
{
    const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i2 >= A.n2) return;
    predict_fast_trace<NW, TWO>(A, i2, A.zla + (int)blockIdx.y, (int)blockIdx.y);
}

// predict_warp_kernel: one WARP per target trace (pst_predict_warp.h) for launches with few traces -- 2-D panels (860 -
// 1280 traces per spray level), painting (one trace), small cubes -- where the thread-per-trace kernel leaves the GPU
// with ~1000 threads.  Same bits.  The factor scratch of a trace is [band+1][n1] here.
template <int NW, bool TWO>
__global__ void __launch_bounds__(128)
predict_warp_kernel(const PredArgs A)
{
    __shared__ PredWarpWS<NW, TWO> ws[4];
    constexpr int NC = 2 * NW + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i2 = blockIdx.x * 4 + warp;
    const int zl = A.zla + blockIdx.y;
    if (i2 >= A.n2) return;                                   // the whole warp
    const int n1 = A.n1;
    const long n2 = A.n2;
    const long base = (long)zl * n1 * n2 + i2;
    float *out = A.out + base;
    const int s2 = i2 - A.a, s3 = (A.ze0 + zl) - A.b, s2g = s2 + A.t_off;
    if (s2 < 0 || s2 >= A.n2 || s2g < 0 || s2g >= A.ntg || s3 < 0 || s3 >= A.n3) {
        for (int k = lane; k < n1; k += 32) out[(long)k * n2] = 0.f;
        return;
    }
    const float *x1 = A.in1 + base + A.in1_off;
    const float *g1 = A.sg1 + base + A.sg1_off;
    const float *x2 = TWO ? A.in2 + base + A.in2_off : nullptr;
    const float *g2 = TWO ? A.sg2 + base + A.sg2_off : nullptr;
    float *scr = A.scr + ((long)blockIdx.y * n2 + i2) * (long)NC * n1;
    predict_warp_trace<NW, TWO>(ws[warp], x1, g1, x2, g2, n2, A.forw1 != 0, A.forw2 != 0, A.reg, A.tb, n1, scr, out);
}

// ---------------------------------------------------------------------------------------
// layout changes: [i3][i2][i1] (reference order, i1 fastest) <-> [i3][i1][i2] (trace-minor)
__global__ void transpose_kernel(const float *__restrict__ in, float *__restrict__ out, int rows,
                                 int cols)
{
    // per plane: in is [rows][cols] with cols fastest; out is [cols][rows] with rows fastest
    __shared__ float tile[32][33];
    const long plane = (long)rows * cols * blockIdx.z;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int rr = r0 + r, cc = c0 + threadIdx.x;
        if (rr < rows && cc < cols) tile[r][threadIdx.x] = in[plane + (long)rr * cols + cc];
    }
    __syncthreads();
    for (int c = threadIdx.y; c < 32; c += blockDim.y) {
        const int cc = c0 + c, rr = r0 + threadIdx.x;
        if (rr < rows && cc < cols) out[plane + (long)cc * rows + rr] = tile[threadIdx.x][c];
    }
}

static int transpose_planes(pst_ctx *c, const float *in, float *out, int rows, int cols, int planes)
{
    // grid.z is limited to 65535: loop in batches
    const int zmax = 32768;
    for (int z0 = 0; z0 < planes; z0 += zmax) {
        const int nz = std::min(zmax, planes - z0);
        dim3 grid((cols + 31) / 32, (rows + 31) / 32, nz), block(32, 8);
        const long off = (long)rows * cols * z0;
        PST_LAUNCHB(c, PST_K_OTHER, 8.0 * (double)rows * cols * nz, (transpose_kernel<<<grid, block, 0, c->stream>>>(in + off, out + off, rows, cols)));
    }
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

// ---------------------------------------------------------------------------------------
struct SlotPtrs { const float *p[PST_MAXSLOT]; };

// mean over all np slots in slot order, / np (csomean3d :1531-1535)
__global__ void __launch_bounds__(256)
slot_mean_kernel(SlotPtrs S, int np, float *__restrict__ out, long zoff_out, long zoff_slot, long count)
{
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x) {
        float sum = 0;
        for (int s = 0; s < np; s++) sum = sum + S.p[s][zoff_slot + e];
        sum = sum / np;
        out[zoff_out + e] = sum;
    }
}

// exact m-th order statistic of nmf slot values (mf + sf_quantile :1058-1084,:1214-1252).
// Selection by rank counting: O(nmf^2) compares, no data-dependent memory traffic.
template <int NMF>
__global__ void __launch_bounds__(256)
slot_median_kernel(SlotPtrs S, int nmf_rt, float *__restrict__ out, long zoff_out, long zoff_slot, long count)
{
    const int nmf = NMF > 0 ? NMF : nmf_rt;
    const int m = (nmf - 1) / 2;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x) {
        float v[NMF > 0 ? NMF : PST_MAXSLOT];
#pragma unroll
        for (int s = 0; s < nmf; s++) v[s] = S.p[s][zoff_slot + e];
        float res = v[0];
#pragma unroll
        for (int s = 0; s < nmf; s++) {
            int lt = 0, le = 0;
#pragma unroll
            for (int q = 0; q < nmf; q++) { lt += (v[q] < v[s]); le += (v[q] <= v[s]); }
            if (lt <= m && m < le) res = v[s];
        }
        out[zoff_out + e] = res;
    }
}

// exact q-th order statistic of n values (sf_quantile :1058-1084 returns the same VALUE) by rank counting
template <int MAXN>
__device__ __forceinline__ float select_rank(const float (&v)[MAXN], int n, int q)
{
    float res = v[0];
    for (int s = 0; s < n; s++) {
        int lt = 0, le = 0;
        for (int t = 0; t < n; t++) { lt += (v[t] < v[s]); le += (v[t] <= v[s]); }
        if (lt <= q && q < le) res = v[s];
    }
    return res;
}

// Space-varying median filter over the slot axis (svmf, sof3d_cfuns.c:1110-1136,:1254-1352; sof_cfuns.c:1190-1216,
// :1334-1431; `option=2`), centre row kept.  One thread = one output trace: (1) the median of length nfw+2 of EVERY slot
// row and sample, its magnitudes summed sequentially in float in the reference's row-major order -> panel average;
// (2) per sample the window length nfw+2 / nfw / nfw-2 / nfw-4 from |x| against avg/2, avg, 2 avg, and the median of
// that length around the centre slot.  Slot rows outside [0, np) are edge replicas (boundary(), ifbound = 1).
// DEFINED-BEHAVIOUR VARIANT: the reference's first pass reads one row past its extended panel for the last slot row
// (:1283), so heap contents enter its panel average; here that row is the edge replica as well.
template <int MAXN>
__global__ void __launch_bounds__(128)
slot_svmf_kernel(SlotPtrs S, int np, int nfw, int n1, int n2, int planes, float *__restrict__ out, long zoff_out, long zoff_slot)
{
    const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int zl = blockIdx.y;
    if (i2 >= n2 || zl >= planes) return;
    const long base = (long)zl * n1 * n2 + i2;
    const int nfilter = nfw + 2, m = (nfilter - 1) / 2, m2 = (nfw - 1) / 2, keep = (np - 1) / 2;
    float win[MAXN];
    float sum = 0.f;
    for (int i = 0; i < np; i++)
        for (int j = 0; j < n1; j++) {
            const long e = zoff_slot + base + (long)j * n2;
            for (int k = 0; k < nfilter; k++) {
                int s = i + k - m2;                                  // (m + i + k - m2) - m
                s = s < 0 ? 0 : (s > np - 1 ? np - 1 : s);
                win[k] = S.p[s][e];
            }
            sum = sum + fabsf(select_rank<MAXN>(win, nfilter, m));
        }
    const float avg = sum / (float)(n1 * np);
    for (int j = 0; j < n1; j++) {
        const long e = zoff_slot + base + (long)j * n2;
        const float x = fabsf(S.p[keep][e]);
        int wl;
        if (x < avg) wl = (x < avg / 2) ? nfw + 2 : nfw;
        else         wl = (x > avg * 2) ? nfw - 4 : nfw - 2;
        const int h = (wl - 1) / 2;
        for (int k = 0; k < wl; k++) {
            int s = keep + k - h;
            s = s < 0 ? 0 : (s > np - 1 ? np - 1 : s);
            win[k] = S.p[s][e];
        }
        out[zoff_out + base + (long)j * n2] = select_rank<MAXN>(win, wl, h);
    }
}

// pwsmooth_lop forward (sof_cfuns.c:1099-1108): out = sum_is u_is * w_is * ws, is ascending;
// MODE 0: ws = 1 and out -> t (normalisation spray of ones); MODE 1: ws = (t != 0 ? 1/t : 0)
template <int MODE>
__global__ void __launch_bounds__(256)
slot_wsum_kernel(SlotPtrs S, int ns, const float *__restrict__ tnorm, float *__restrict__ out, long count)
{
    const int ns2 = 2 * ns + 1;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x) {
        float ws = 1.0f;
        if (MODE == 1) {
            const float t = tnorm[e];
            ws = (0.0f != t) ? (float)(1.0 / (double)t) : 0.0f;
        }
        float acc = 0.f;
        for (int is = 0; is < ns2; is++) {
            const float w = (float)(ns + 1 - abs(is - ns));
            acc += S.p[is][e] * w * ws;
        }
        out[e] = acc;
    }
}

__global__ void fill_kernel_s(float *__restrict__ x, float v, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = v;
}

// =======================================================================================
// host side
// =======================================================================================

struct SprayPlan {
    int n1, n2, n3, ns2, ns3, nw;
    int np2, np3, np;
    float eps_reg;                 // regularisation (already squared)
    bool live[PST_MAXSLOT];        // slots that must be produced
    // plane bookkeeping (global plane indices; n3 above is the GLOBAL extent): the trace-minor
    // input volumes hold planes [zs0, zs1); outputs are produced for planes [zt0, zt1).
    // Single GPU: zs = zt = [0, n3).  Distributed: zt = this rank's slab, zs = slab + ns3 halos.
    int zs0, zs1, zt0, zt1;
    int t_off = 0, ntg = 0;        // cut panels (see PredArgs); ntg = 0 means whole panels (ntg = n2)
};

static void plan_close_parents(SprayPlan &P)
{
    // a live slot needs its parent(s): walk from the farthest level inwards
    for (int lev = P.ns2 + P.ns3; lev >= 1; lev--)
        for (int s = 0; s < P.np; s++) {
            const int a = s % P.np2 - P.ns2, b = s / P.np2 - P.ns3;
            if (!P.live[s] || abs(a) + abs(b) != lev) continue;
            if (a != 0) P.live[(b + P.ns3) * P.np2 + (a - (a > 0 ? 1 : -1) + P.ns2)] = true;
            if (b != 0) P.live[(b - (b > 0 ? 1 : -1) + P.ns3) * P.np2 + (a + P.ns2)] = true;
        }
}

// Chunk height along n3 (target planes per chunk, *cz) and the planes a chunk's slot / scratch volumes hold
// (*nzl = chunk + ns3 planes of halo sources either side), from the PST_SPRAY_CHUNK_GB budget.  The ONE place that
// decides it: the arena reservation and spray_run must agree.
static void spray_chunk_planes(const SprayPlan &P, int nlive, int *cz, int *nzl)
{
    const int NC = 2 * P.nw + 1;
    const double bytes_per_plane = (double)P.n1 * P.n2 * 4.0 * (nlive + NC);
    int z = (int)(spray_chunk_bytes() / bytes_per_plane) - 2 * P.ns3;
    if (z < 1) z = 1;
    if (z > P.zt1 - P.zt0) z = P.zt1 - P.zt0;
    *cz = z;
    *nzl = std::min(P.zs1 - P.zs0, z + 2 * P.ns3);
}

template <int NW>
static void launch_predict(pst_ctx *c, const PredArgs &A, bool two)
{
    const int threads = A.n2 >= 128 ? 128 : (A.n2 >= 64 ? 64 : 32);
    dim3 grid((A.n2 + threads - 1) / threads, A.zlb - A.zla);
    // PST_PREDICT_FAST=0: the first formulation of the kernel; =2: two-parent kernel with 192 registers and 2 CTAs per
    // SM instead of 168 (a few spills) and 3 (A/B measurements; same results)
    static const int fast = []() { const char *e = getenv("PST_PREDICT_FAST"); return e ? atoi(e) : 1; }();
    // algorithmic flops per predicted sample: predict1 47 (nw=1) / 116 (nw=2), predict2 78 / 187
    const double fl = NW == 1 ? (two ? 78.0 : 47.0) : (two ? 187.0 : 116.0);
    // few traces: one warp per trace (break-even against one thread per trace measured / modelled at ~8000 traces).
    // PST_PREDICT_WARP=0: never, =1: always (tests), unset: by the trace count of the launch
    static const int warp_mode = []() { const char *e = getenv("PST_PREDICT_WARP"); return e ? atoi(e) : -1; }();
    const long traces = (long)A.n2 * (A.zlb - A.zla);
    if (warp_mode == 1 || (warp_mode < 0 && traces <= 6000)) {
        dim3 wgrid((A.n2 + 3) / 4, A.zlb - A.zla);
        PST_LAUNCHBF(c, PST_K_PREDICT, (two ? 20.0 : 12.0) * (double)A.n1 * traces, fl * (double)A.n1 * traces,
            if (two) predict_warp_kernel<NW, true><<<wgrid, 128, 0, c->stream>>>(A);
            else     predict_warp_kernel<NW, false><<<wgrid, 128, 0, c->stream>>>(A));
        return;
    }
    PST_LAUNCHBF(c, PST_K_PREDICT, (two ? 20.0 : 12.0) * (double)A.n1 * A.n2 * (A.zlb - A.zla), fl * (double)A.n1 * A.n2 * (A.zlb - A.zla),
        if (fast == 2 && two) predict_fast_kernel<NW, true, 2><<<grid, threads, 0, c->stream>>>(A);
        else if (fast && two) predict_fast_kernel<NW, true, 3><<<grid, threads, 0, c->stream>>>(A);
        else if (fast)        predict_fast_kernel<NW, false, 3><<<grid, threads, 0, c->stream>>>(A);
        else if (two)    predict_kernel<NW, true><<<grid, threads, 0, c->stream>>>(A);
        else             predict_kernel<NW, false><<<grid, threads, 0, c->stream>>>(A));
}

typedef int (*chunk_reduce_fn)(pst_ctx *c, void *user, const SprayPlan &P, float *const *slot, int ze0,
                               int z0, int z1);

// Chunk-wise staging between the caller's reference-order volumes and the trace-minor working volumes, so that the
// host-pointer entry points can run their transfers under the kernels (pst_pipe_*, pst_api.cu): the inputs of a chunk
// are transposed right before the chunk (after waiting for exactly those planes), its output planes are transposed
// back and handed to the download stream right after its reduction.
struct SprayIO {
    const float *src[3] = {nullptr, nullptr, nullptr};   // din, dipi, dipx of the slab, reference order [i3][i2][i1]
    float *dstT[3] = {nullptr, nullptr, nullptr};        // trace-minor, first SLAB plane (halo planes lie before it)
    int done = 0;                                        // slab planes [0, done) are staged
    const float *outT = nullptr;                         // trace-minor result, slab planes
    float *d_out = nullptr;                              // reference-order result
};

static int transpose_planes(pst_ctx *c, const float *in, float *out, int rows, int cols, int planes);

static int spray_stage_in(pst_ctx *c, const SprayPlan &P, SprayIO *io, int z_upto)
{
    const int nz = P.zt1 - P.zt0;
    const int hi = std::min(nz, z_upto - P.zt0);
    if (hi <= io->done) return PST_OK;
    PST_TRY(pst_pipe_wait_planes(c, hi));
    const size_t off = (size_t)P.n1 * P.n2 * io->done;
    for (int v = 0; v < 3; v++)
        if (io->src[v]) PST_TRY(transpose_planes(c, io->src[v] + off, io->dstT[v] + off, P.n2, P.n1, hi - io->done));
    io->done = hi;
    return PST_OK;
}

static int spray_stage_out(pst_ctx *c, const SprayPlan &P, SprayIO *io, int z0, int z1)
{
    const size_t off = (size_t)P.n1 * P.n2 * (z0 - P.zt0);
    PST_TRY(transpose_planes(c, io->outT + off, io->d_out + off, P.n1, P.n2, z1 - z0));
    return pst_pipe_emit(c, io->d_out, z0 - P.zt0, z1 - P.zt0);
}

// Spray every live slot for all planes, chunk by chunk along n3, and hand the slot volumes
// of each chunk to `reduce`.  dT/piT/pxT are full trace-minor volumes.
static int spray_run(pst_ctx *c, const SprayPlan &P, const float *dT, const float *piT, const float *pxT,
                     chunk_reduce_fn reduce, void *user, SprayIO *io = nullptr)
{
    const int n1 = P.n1, n2 = P.n2, n3 = P.n3, ns3 = P.ns3, nw = P.nw;   // n3: global extent
    const long plane = (long)n1 * n2;
    int nlive = 0;
    for (int s = 0; s < P.np; s++) nlive += P.live[s] ? 1 : 0;
    const int NC = 2 * nw + 1;
    // chunk height: bound slot + scratch memory to ~6 GB
    int cz, nzl_max;
    spray_chunk_planes(P, nlive, &cz, &nzl_max);
    float *slotbuf[PST_MAXSLOT];
    const int centre = ns3 * P.np2 + P.ns2;
    for (int s = 0; s < P.np; s++) {
        slotbuf[s] = nullptr;
        if (P.live[s] && s != centre) PST_TRY(pst_arena_get(c, (size_t)plane * nzl_max, &slotbuf[s]));
    }
    float *scr;
    PST_TRY(pst_arena_get(c, (size_t)plane * nzl_max * NC, &scr));
    const RegC reg = make_reg(P.eps_reg);
    const BTabS tb = make_btab_s(nw);

    for (int z0 = P.zt0; z0 < P.zt1; z0 += cz) {
        const int z1 = std::min(P.zt1, z0 + cz);
        const int ze0 = std::max(P.zs0, z0 - ns3), ze1 = std::min(P.zs1, z1 + ns3);
        if (io) PST_TRY(spray_stage_in(c, P, io, ze1));
        float *slot[PST_MAXSLOT];
        for (int s = 0; s < P.np; s++) slot[s] = slotbuf[s];
        slot[centre] = const_cast<float *>(dT) + (long)(ze0 - P.zs0) * plane;
        const float *pi_c = piT + (long)(ze0 - P.zs0) * plane, *px_c = pxT ? pxT + (long)(ze0 - P.zs0) * plane : nullptr;
        for (int lev = 1; lev <= P.ns2 + ns3; lev++) {
            for (int s = 0; s < P.np; s++) {
                const int a = s % P.np2 - P.ns2, b = s / P.np2 - ns3;
                if (!P.live[s] || abs(a) + abs(b) != lev) continue;
                int lo = z0 - (b >= 0 ? ns3 - b : 0), hi = z1 + (b <= 0 ? ns3 + b : 0);
                lo = std::max(lo, ze0); hi = std::min(hi, ze1);
                if (hi <= lo) continue;
                PredArgs A{};
                A.n1 = n1; A.n2 = n2; A.n3 = n3; A.ze0 = ze0; A.zla = lo - ze0; A.zlb = hi - ze0;
                A.a = a; A.b = b; A.reg = reg; A.tb = tb; A.out = slot[s]; A.scr = scr;
                A.t_off = P.t_off; A.ntg = P.ntg > 0 ? P.ntg : n2;
                int nin = 0;
                const float *inp[2]; const float *sg[2]; long ioff[2], soff[2]; int fw[2];
                if (a != 0) {                       // inline parent (get_update bit 1, :1661-1673)
                    const int sa = a > 0 ? 1 : -1;
                    inp[nin] = slot[(b + ns3) * P.np2 + (a - sa + P.ns2)];
                    sg[nin] = pi_c; ioff[nin] = -sa; fw[nin] = a > 0;
                    soff[nin] = a > 0 ? -1 : 0;     // up: parent-location dip; down: target-location dip
                    nin++;
                }
                if (b != 0) {                       // xline parent (bit 2, :1674-1686)
                    const int sb = b > 0 ? 1 : -1;
                    inp[nin] = slot[(b - sb + ns3) * P.np2 + (a + P.ns2)];
                    sg[nin] = px_c; ioff[nin] = -(long)sb * plane; fw[nin] = b > 0;
                    soff[nin] = b > 0 ? -plane : 0;
                    nin++;
                }
                A.in1 = inp[0]; A.sg1 = sg[0]; A.in1_off = ioff[0]; A.sg1_off = soff[0]; A.forw1 = fw[0];
                if (nin == 2) { A.in2 = inp[1]; A.sg2 = sg[1]; A.in2_off = ioff[1]; A.sg2_off = soff[1]; A.forw2 = fw[1]; }
                if (nw == 1) launch_predict<1>(c, A, nin == 2);
                else         launch_predict<2>(c, A, nin == 2);
                c->stats.predictions += (long long)(hi - lo) * n2;
            }
        }
        PST_CUDA(cudaGetLastError());
        PST_TRY(reduce(c, user, P, slot, ze0, z0, z1));
        if (io && io->d_out) PST_TRY(spray_stage_out(c, P, io, z0, z1));
    }
    return PST_OK;
}

struct ReduceOut { float *outT; int nmf; const float *tnorm; int mode; };

static int reduce_mean(pst_ctx *c, void *user, const SprayPlan &P, float *const *slot, int ze0, int z0, int z1)
{
    ReduceOut *R = (ReduceOut *)user;
    const long plane = (long)P.n1 * P.n2, count = plane * (z1 - z0);
    SlotPtrs S{};
    for (int s = 0; s < P.np; s++) S.p[s] = slot[s];
    const int grid = pst_grid_for(c, (size_t)count, 256, 2);
    PST_LAUNCH(c, PST_K_SLOTRED, (slot_mean_kernel<<<grid, 256, 0, c->stream>>>(S, P.np, R->outT, plane * (z0 - P.zt0), plane * (z0 - ze0), count)));
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

static int reduce_median(pst_ctx *c, void *user, const SprayPlan &P, float *const *slot, int ze0, int z0, int z1)
{
    ReduceOut *R = (ReduceOut *)user;
    const long plane = (long)P.n1 * P.n2, count = plane * (z1 - z0);
    const int nmf = R->nmf, m = (nmf - 1) / 2, cen = (P.np - 1) / 2;
    SlotPtrs S{};
    for (int q = 0; q < nmf; q++) {              // boundary(): edge replication along the slot axis
        int s = cen - m + q;
        s = std::max(0, std::min(P.np - 1, s));
        S.p[q] = slot[s];
    }
    const int grid = pst_grid_for(c, (size_t)count, 256, 2);
    const long zo = plane * (z0 - P.zt0), zs = plane * (z0 - ze0);
    KTimer kt(c, PST_K_SLOTRED);
    switch (nmf) {
        case 3: slot_median_kernel<3><<<grid, 256, 0, c->stream>>>(S, nmf, R->outT, zo, zs, count); break;
        case 5: slot_median_kernel<5><<<grid, 256, 0, c->stream>>>(S, nmf, R->outT, zo, zs, count); break;
        case 9: slot_median_kernel<9><<<grid, 256, 0, c->stream>>>(S, nmf, R->outT, zo, zs, count); break;
        case 13: slot_median_kernel<13><<<grid, 256, 0, c->stream>>>(S, nmf, R->outT, zo, zs, count); break;
        case 17: slot_median_kernel<17><<<grid, 256, 0, c->stream>>>(S, nmf, R->outT, zo, zs, count); break;
        case 19: slot_median_kernel<19><<<grid, 256, 0, c->stream>>>(S, nmf, R->outT, zo, zs, count); break;
        default: slot_median_kernel<0><<<grid, 256, 0, c->stream>>>(S, nmf, R->outT, zo, zs, count); break;
    }
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

static int reduce_svmf(pst_ctx *c, void *user, const SprayPlan &P, float *const *slot, int ze0, int z0, int z1)
{
    ReduceOut *R = (ReduceOut *)user;
    const long plane = (long)P.n1 * P.n2;
    SlotPtrs S{};
    for (int s = 0; s < P.np; s++) S.p[s] = slot[s];
    const int threads = P.n2 >= 128 ? 128 : (P.n2 >= 64 ? 64 : 32);
    dim3 grid((P.n2 + threads - 1) / threads, z1 - z0);
    const long zo = plane * (z0 - P.zt0), zs = plane * (z0 - ze0);
    PST_LAUNCH(c, PST_K_SLOTRED,
        if (R->nmf + 2 <= 13) slot_svmf_kernel<13><<<grid, threads, 0, c->stream>>>(S, P.np, R->nmf, P.n1, P.n2, z1 - z0, R->outT, zo, zs);
        else                  slot_svmf_kernel<PST_MAXSLOT + 2><<<grid, threads, 0, c->stream>>>(S, P.np, R->nmf, P.n1, P.n2, z1 - z0, R->outT, zo, zs));
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

static int reduce_wsum(pst_ctx *c, void *user, const SprayPlan &P, float *const *slot, int ze0, int z0, int z1)
{
    ReduceOut *R = (ReduceOut *)user;
    const long plane = (long)P.n1 * P.n2, count = plane * (z1 - z0);
    SlotPtrs S{};
    for (int s = 0; s < P.np; s++) S.p[s] = slot[s] + plane * (z0 - ze0);
    const int grid = pst_grid_for(c, (size_t)count, 256, 2);
    PST_LAUNCH(c, PST_K_SLOTRED,
        if (R->mode == 0) slot_wsum_kernel<0><<<grid, 256, 0, c->stream>>>(S, P.ns2, nullptr, R->outT + plane * (z0 - P.zt0), count);
        else              slot_wsum_kernel<1><<<grid, 256, 0, c->stream>>>(S, P.ns2, R->tnorm + plane * (z0 - P.zt0), R->outT + plane * (z0 - P.zt0), count));
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

static int check_spray_args(int n1, int n2, int n3, int ns2, int ns3, int order)
{
    if (n1 < 1 || n2 < 1 || n3 < 1 || ns2 < 0 || ns3 < 0) { pst_set_error("spray: bad dimensions/radii"); return PST_EINVAL; }
    if (order != 1 && order != 2) { pst_set_error("spray: order=%d unsupported (1 or 2)", order); return PST_EUNSUP; }
    if (n1 < 2 * order + 2) { pst_set_error("spray: n1=%d too short for order %d", n1, order); return PST_EINVAL; }
    if ((2 * ns2 + 1) * (2 * ns3 + 1) > PST_MAXSLOT) { pst_set_error("spray: (2ns2+1)(2ns3+1) > %d unsupported", PST_MAXSLOT); return PST_EUNSUP; }
    return PST_OK;
}

int pst_comm_halo_exchange(pst_ctx *c, const float *send_lo, const float *send_hi, float *recv_lo,
                           float *recv_hi, size_t count);                                   // pst_comm.cu

// kind: 0 = mean (3-D), 1 = median (3-D), 2 = weighted normalised smooth (2-D), 3 = median (2-D).
// Distributed contexts: n3 is the GLOBAL plane count, the pointers are this rank's slab; the
// ns3-plane halos of din / dipi / dipx are exchanged once and halo sources are sprayed redundantly.
static int spray_filter_dev(pst_ctx *c, int kind, const float *d_din, const float *d_dipi, const float *d_dipx,
                            int n1, int n2, int n3, int ns2, int ns3, int nmf, int order, float eps_reg,
                            float *d_out)
{
    PST_CUDA(cudaSetDevice(c->device));
    const bool dist = c->comm != nullptr && c->nranks > 1;
    int z0 = 0, z1 = n3;
    if (dist) {
        z0 = (int)(((long)n3 * c->rank) / c->nranks);
        z1 = (int)(((long)n3 * (c->rank + 1)) / c->nranks);
        if (n3 / c->nranks < std::max(ns3, 1)) { pst_set_error("spray: slabs thinner than the xline spray radius"); return PST_EUNSUP; }
    }
    const int nz = z1 - z0;
    SprayPlan P{};
    P.n1 = n1; P.n2 = n2; P.n3 = n3; P.ns2 = ns2; P.ns3 = ns3; P.nw = order;
    P.np2 = 2 * ns2 + 1; P.np3 = 2 * ns3 + 1; P.np = P.np2 * P.np3;
    P.eps_reg = eps_reg;
    P.zt0 = z0; P.zt1 = z1;
    P.zs0 = std::max(0, z0 - ns3); P.zs1 = std::min(n3, z1 + ns3);
    const int ne = P.zs1 - P.zs0;                       // stored planes (slab + halos)
    const int cen = (P.np - 1) / 2;
    for (int s = 0; s < P.np; s++) P.live[s] = (kind == 0 || kind == 2 || kind == 4 || kind == 5);   // SVMF: every slot row
    if (kind == 4 || kind == 5) {
        if (nmf < 5 || nmf + 2 > PST_MAXSLOT + 2 || (nmf % 2) == 0) { pst_set_error("SVMF: median length nmf=%d unsupported (odd, >= 5: the shortest window is nmf-4)", nmf); return PST_EUNSUP; }
    }
    if (kind == 1 || kind == 3) {
        if (nmf < 1 || nmf > PST_MAXSLOT || (nmf % 2) == 0) { pst_set_error("median length nmf=%d unsupported (odd, <= %d)", nmf, PST_MAXSLOT); return PST_EUNSUP; }
        const int m = (nmf - 1) / 2;
        for (int q = cen - m; q <= cen + m; q++) P.live[std::max(0, std::min(P.np - 1, q))] = true;
    }
    P.live[cen] = true;
    plan_close_parents(P);
    int nlive = 0;
    for (int s = 0; s < P.np; s++) nlive += P.live[s] ? 1 : 0;

    const long plane = (long)n1 * n2;
    const size_t n = (size_t)plane * nz, nex = (size_t)plane * ne;
    const int NC = 2 * order + 1;
    int cz_, nzl_;
    spray_chunk_planes(P, nlive, &cz_, &nzl_);
    const size_t nzl = (size_t)nzl_;
    const size_t need = (4 * nex + 2 * n + (size_t)plane * nzl * (nlive + NC)) * sizeof(float) + 64 * 256 + (size_t)nlive * 256;
    PST_TRY(pst_arena_reserve(c, need));
    pst_arena_reset(c);
    float *dT, *piT, *pxT = nullptr, *outT, *tnorm = nullptr;
    PST_TRY(pst_arena_get(c, nex, &dT));
    PST_TRY(pst_arena_get(c, nex, &piT));
    if (d_dipx) PST_TRY(pst_arena_get(c, nex, &pxT));
    PST_TRY(pst_arena_get(c, n, &outT));
    const size_t off = (size_t)plane * (z0 - P.zs0);    // slab position inside the stored range
    SprayIO io;
    io.src[0] = d_din; io.src[1] = d_dipi; io.src[2] = d_dipx;
    io.dstT[0] = dT + off; io.dstT[1] = piT + off; io.dstT[2] = pxT ? pxT + off : nullptr;
    io.outT = outT; io.d_out = d_out;
    // single GPU, one spray pass: inputs staged chunk by chunk inside spray_run.  Otherwise (halo exchange / the
    // normalisation pass of the 2-D smoother read everything first): staged here, outputs still leave chunk by chunk.
    const bool lazy_in = !dist && kind != 2;
    if (!lazy_in) PST_TRY(spray_stage_in(c, P, &io, z1));
    if (dist && ns3 > 0) {
        const size_t cnt = (size_t)plane * ns3;
        float *vols[3] = {dT, piT, pxT};
        for (float *v : vols) {
            if (!v) continue;
            PST_TRY(pst_comm_halo_exchange(c, v + off, v + off + n - cnt, v, v + off + n, cnt));
        }
    }
    ReduceOut R{outT, nmf, nullptr, 0};
    if (kind == 0) PST_TRY(spray_run(c, P, dT, piT, pxT, reduce_mean, &R, &io));
    else if (kind == 1 || kind == 3) PST_TRY(spray_run(c, P, dT, piT, pxT, reduce_median, &R, &io));
    else if (kind == 4 || kind == 5) PST_TRY(spray_run(c, P, dT, piT, pxT, reduce_svmf, &R, &io));
    else {
        // pwsmooth_set (sof_cfuns.c:1113-1132): normalisation = smooth of a volume of ones
        PST_TRY(pst_arena_get(c, n, &tnorm));
        float *ones;
        PST_TRY(pst_arena_get(c, nex, &ones));
        PST_LAUNCH(c, PST_K_OTHER, (fill_kernel_s<<<pst_grid_for(c, nex, 256), 256, 0, c->stream>>>(ones, 1.0f, nex)));
        const size_t mark = c->arena_used;
        ReduceOut R0{tnorm, 0, nullptr, 0};
        PST_TRY(spray_run(c, P, ones, piT, pxT, reduce_wsum, &R0));
        c->arena_used = mark;          // chunk buffers are reusable
        ReduceOut R1{outT, 0, tnorm, 1};
        PST_TRY(spray_run(c, P, dT, piT, pxT, reduce_wsum, &R1, &io));
    }
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

extern "C" int pst_somean3d_dev(pst_ctx *c, const float *d_din, const float *d_dipi, const float *d_dipx,
                                int n1, int n2, int n3, int ns2, int ns3, int order, float *d_out)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_TRY(check_spray_args(n1, n2, n3, ns2, ns3, order));
    const float eps = 0.01;                       // caller's eps overridden (sof3d_cfuns.c:1402)
    return spray_filter_dev(c, 0, d_din, d_dipi, d_dipx, n1, n2, n3, ns2, ns3, 0, order, eps * eps, d_out);
}

extern "C" int pst_somf3d_dev(pst_ctx *c, const float *d_din, const float *d_dipi, const float *d_dipx,
                              int n1, int n2, int n3, int ns2, int ns3, int nmf, int option, int order,
                              float *d_out)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_TRY(check_spray_args(n1, n2, n3, ns2, ns3, order));
    if (option != 1 && option != 2) { pst_set_error("somf3d: option=%d unknown (1 = MF, 2 = SVMF)", option); return PST_EINVAL; }
    const float eps = 0.01;                       // sof3d_cfuns.c:1586
    return spray_filter_dev(c, option == 1 ? 1 : 4, d_din, d_dipi, d_dipx, n1, n2, n3, ns2, ns3, nmf, order, eps * eps, d_out);
}

int pst_somean2d_dev(pst_ctx *c, const float *d_din, const float *d_dip, int n1, int n2, int n3, int ns,
                     int order, float eps, float *d_out)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_TRY(check_spray_args(n1, n2, n3, ns, 0, order));
    return spray_filter_dev(c, 2, d_din, d_dip, nullptr, n1, n2, n3, ns, 0, 0, order, eps * eps, d_out);
}

int pst_somf2d_dev(pst_ctx *c, const float *d_din, const float *d_dip, int n1, int n2, int n3, int ns,
                   int nmf, int option, int order, float eps, float *d_out)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_TRY(check_spray_args(n1, n2, n3, ns, 0, order));
    if (option != 1 && option != 2) { pst_set_error("somf2d: option=%d unknown (1 = MF, 2 = SVMF)", option); return PST_EINVAL; }
    return spray_filter_dev(c, option == 1 ? 3 : 5, d_din, d_dip, nullptr, n1, n2, n3, ns, 0, nmf, order, eps * eps, d_out);
}

// =======================================================================================
// csint3d: interpolation by shaping-regularised CG with the plane-wave smoother as shaping
// operator (reference soint3d_cfuns.c:2510-2640; ps_conjgrad with hasp0 = true :  p0 = data,
// L = known-data mask :1350-1372, S = pwsmooth3_lop :2231-2300 = inline 2-D pwsmooth, transpose,
// xline 2-D pwsmooth).  The forward smoother is the 2-D chain spray + normalised triangle
// weights already used by somean2d; its ADJOINT needs the adjoint spray (pwspray_lop(adj)
// :1963-2003) whose step is predict_step(adj = true) :1777-1804: solve the banded system FIRST,
// then apply (W'W)' and the end terms.
//
// Volumes live in two trace-minor layouts: A = [i3][i1][i2] (panels = xline planes, traces along
// i2) for the inline smoother and B = [i2][i1][i3] for the xline smoother; swap_ab_kernel moves
// between them (a strided 2-D transpose per i1).  The CG vectors stay in layout A.
// =======================================================================================

// in: [P][n1][Q] (Q fastest)  ->  out: [Q][n1][P] (P fastest); only q in [q_lo, q_lo + nq) is written, at q - q_lo
__global__ void swap_ab_kernel(const float *__restrict__ in, float *__restrict__ out, int P, int n1, int Q, int zero_plus,
                               int q_lo, int nq)
{
    __shared__ float tile[32][33];
    const int i1 = blockIdx.z;
    const int q0 = q_lo + blockIdx.x * 32, p0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int p = p0 + r, q = q0 + threadIdx.x;
        if (p < P && q < Q) tile[r][threadIdx.x] = in[((long)p * n1 + i1) * Q + q];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int q = q0 + r, p = p0 + threadIdx.x;
        if (p < P && q < q_lo + nq) {
            float v = tile[threadIdx.x][r];
            if (zero_plus) v = 0.f + v;                      // "smooth[] += xtmp" onto a zeroed volume (:2296)
            out[((long)(q - q_lo) * n1 + i1) * P + p] = v;
        }
    }
}

static int swap_ab(pst_ctx *c, const float *in, float *out, int P, int n1, int Q, bool zero_plus = false, int q_lo = 0, int nq = -1)
{
    const int zmax = 32768;
    if (n1 > zmax) { pst_set_error("sint3d: n1 > %d unsupported", zmax); return PST_EUNSUP; }
    if (nq < 0) nq = Q - q_lo;
    dim3 grid((nq + 31) / 32, (P + 31) / 32, n1), block(32, 8);
    PST_LAUNCHB(c, PST_K_OTHER, 8.0 * (double)P * n1 * nq, (swap_ab_kernel<<<grid, block, 0, c->stream>>>(in, out, P, n1, Q, zero_plus ? 1 : 0, q_lo, nq)));
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

struct AdjArgs {
    float *tr;                 // [panel][k][i2]: running adjoint trace of every model trace i (in/out)
    const float *data;         // data-side volume ("out" of pwsmooth_lop), same layout
    const float *tnorm;        // normalisation t; ws = (t != 0 ? 1/t : 0)
    const float *sg;           // slopes, same layout
    float *scr;                // factor scratch [panel in launch][k][NB+1][i2]
    float wslot;               // triangle weight of the slot feeding this level
    int shift;                 // data trace ip = i + shift
    int sg_shift;              // slope trace = i + sg_shift
    int forw;
    int n1, n2;
    int p0;                    // first panel of this launch
    int t_off, ntg;            // cut panels: global index of local trace 0, global trace count
    RegC reg;
    BTabS tb;
};

// One thread = one model trace i of one panel:  tr <- predict_step(adj)( tr + u_slot[ip] ),
// u_slot[ip] = (data[ip] * wslot) * ws[ip]  (pwsmooth_lop(adj) :2096-2104, pwspray_lop(adj) :1971-1999).
template <int NW>
__global__ void __launch_bounds__(128)
predict_adj_kernel(const AdjArgs A)
{
    constexpr int NA = 2 * NW + 1, NB = 2 * NW, NC = NB + 1;
    const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i2 >= A.n2) return;
    const int ip = i2 + A.shift, ipg = ip + A.t_off;
    // "continue": the running trace is left alone (data trace outside the panel; outside the local range of a cut panel
    // the model trace is a halo trace nobody reads)
    if (ip < 0 || ip >= A.n2 || ipg < 0 || ipg >= A.ntg) return;
    { const int ig = i2 + A.sg_shift; if (ig < 0 || ig >= A.n2) return; }
    const int n1 = A.n1, n2 = A.n2;
    const long pbase = (long)(A.p0 + blockIdx.y) * n1 * n2;
    float *tr = A.tr + pbase + i2;
    const float *dat = A.data + pbase + ip;
    const float *tn = A.tnorm + pbase + ip;
    const float *g1 = A.sg + pbase + (i2 + A.sg_shift);
    float *scr = A.scr + ((long)blockIdx.y * n1 * NC) * n2 + i2;
    const bool f1 = A.forw != 0;
    const RegC rg = A.reg;

    // ---- pass 1 (ascending): taps -> W'W bands + regularisation -> LDL' column -> forward substitution
    float W1[NA][NA];
    float O[NB][NB], D[NB], Bh[NB];
#pragma unroll
    for (int cc = 0; cc < NA; cc++)
#pragma unroll
        for (int j = 0; j < NA; j++) W1[cc][j] = 0.f;
#pragma unroll
    for (int h = 0; h < NB; h++) {
        D[h] = 0.f; Bh[h] = 0.f;
#pragma unroll
        for (int m = 0; m < NB; m++) O[h][m] = 0.f;
    }
    // loads run PF steps ahead of their use (register queues, see predict_kernel)
    constexpr int PF = 4;
    float gq[PF], tq[PF], dq[PF], rq[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) {
        const int kk = u, ii = u - NW;                      // step i = -NW + u: slope sample kk, right-hand side sample i
        gq[u] = (kk < n1) ? g1[(long)kk * n2] : 0.f;
        const bool in = ii >= 0 && ii < n1;
        tq[u] = in ? tn[(long)ii * n2] : 0.f;
        dq[u] = in ? dat[(long)ii * n2] : 0.f;
        rq[u] = in ? tr[(long)ii * n2] : 0.f;
    }
    for (int i0 = -NW; i0 < n1; i0 += PF) {
#pragma unroll
      for (int u = 0; u < PF; u++) {
        const int i = i0 + u;
        if (i >= n1) break;
        const int kk = i + NW;
        const float gv = gq[u], tv = tq[u], dv = dq[u], rv = rq[u];
        {
            const int kn = kk + PF, in2 = i + PF;
            gq[u] = (kn < n1) ? g1[(long)kn * n2] : 0.f;
            const bool in = in2 >= 0 && in2 < n1;
            tq[u] = in ? tn[(long)in2 * n2] : 0.f;
            dq[u] = in ? dat[(long)in2 * n2] : 0.f;
            rq[u] = in ? tr[(long)in2 * n2] : 0.f;
        }
        {
            float a1[NA];
            if (kk < n1) spray_taps<NW>(A.tb, gv, f1, a1);
            else {
#pragma unroll
                for (int j = 0; j < NA; j++) a1[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < NA; j++) W1[NA - 1][j] = a1[j];
        }
        if (i >= 0) {
            float dg = rg.d_in;
            if (i == 0 || i == n1 - 1) dg = rg.d_e0;
            if (i == 1 || i == n1 - 2) dg = rg.d_e1;
            float of[NB];
            of[0] = (i == 0 || i == n1 - 2) ? rg.o0_e : rg.o0_in;
            of[1] = rg.o1;
#pragma unroll
            for (int m = 2; m < NB; m++) of[m] = 0.0f;
#pragma unroll
            for (int j = 0; j < NA; j++) {
                const int k = i + j - NW;
                if (k >= NW && k < n1 - NW) { const float aj = W1[j][j]; dg += aj * aj; }
            }
#pragma unroll
            for (int m = 0; m < NB; m++) {
#pragma unroll
                for (int j = m + 1; j < NA; j++) {
                    const int k = i + j - NW;
                    if (k >= NW && k < n1 - NW) of[m] += W1[j][j - m - 1] * W1[j][j];
                }
            }
            // right-hand side: the running trace plus the slot contribution
            const float ws = (0.0f != tv) ? (float)(1.0 / (double)tv) : 0.0f;
            const float uu = dv * A.wslot * ws;
            const float rhs = rv + uu;
            float t = dg;
#pragma unroll
            for (int m = 0; m < NB; m++)
                if (m < i) t -= (O[m][m] * O[m][m]) * D[m];
            const float dk = t;
            float ok[NB];
#pragma unroll
            for (int q = 0; q < NB; q++) {
                float v = of[q];
#pragma unroll
                for (int m = 0; m < NB - q - 1; m++)
                    if (m < i) v -= (O[m][m] * O[m][q + m + 1]) * D[m];
                ok[q] = (q < n1 - i - 1) ? v / dk : 0.f;
            }
            float bk = rhs;
#pragma unroll
            for (int m = 0; m < NB; m++)
                if (m < i) bk -= O[m][m] * Bh[m];
            float *sc = scr + (long)i * NC * n2;
            sc[0] = bk / dk;
#pragma unroll
            for (int q = 0; q < NB; q++) sc[(long)(1 + q) * n2] = ok[q];
#pragma unroll
            for (int h = NB - 1; h > 0; h--) {
                D[h] = D[h - 1]; Bh[h] = Bh[h - 1];
#pragma unroll
                for (int m = 0; m < NB; m++) O[h][m] = O[h - 1][m];
            }
            D[0] = dk; Bh[0] = bk;
#pragma unroll
            for (int m = 0; m < NB; m++) O[0][m] = ok[m];
        }
#pragma unroll
        for (int cc = 0; cc < NA - 1; cc++)
#pragma unroll
            for (int j = 0; j < NA; j++) W1[cc][j] = W1[cc + 1][j];
      }
    }
    // ---- pass 2 (descending): back substitution, y stored in place
    float Y[NB];
#pragma unroll
    for (int m = 0; m < NB; m++) Y[m] = 0.f;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    constexpr int PB = 4;
    float cq[PB][NC];
#pragma unroll
    for (int u = 0; u < PB; u++) {
        const int k = n1 - 1 - u;
#pragma unroll
        for (int q = 0; q < NC; q++) cq[u][q] = (k >= 0) ? scr[((long)k * NC + q) * n2] : 1.f;
    }
    for (int k0 = n1 - 1; k0 >= 0; k0 -= PB) {
#pragma unroll
      for (int u = 0; u < PB; u++) {
        const int k = k0 - u;
        if (k < 0) break;
        float col[NC];
#pragma unroll
        for (int q = 0; q < NC; q++) col[q] = cq[u][q];
        {
            const int kn = k - PB;
#pragma unroll
            for (int q = 0; q < NC; q++) cq[u][q] = (kn >= 0) ? scr[((long)kn * NC + q) * n2] : 1.f;
        }
        float t = col[0];
#pragma unroll
        for (int m = 0; m < NB; m++)
            if (m < n1 - k - 1) t -= col[1 + m] * Y[m];
        tr[(long)k * n2] = t;
        if (k == 0) t0 = t;
        if (k == 1) t1 = t;
        if (k == n1 - 2) t2 = t;
        if (k == n1 - 1) t3 = t;
#pragma unroll
        for (int m = NB - 1; m > 0; m--) Y[m] = Y[m - 1];
        Y[0] = t;
      }
    }
    // ---- pass 3 (ascending): pwd_set(adj = true): tmp = W y (rows [nw, n-nw), scatter order), out = W' tmp
    // windows at leading row e: YW[c] = y[e - NW + c]; W1[c][.] / TM[c] <-> row e - 2NW + c
    float YW[NA], TM[NA];
#pragma unroll
    for (int cc = 0; cc < NA; cc++) {
        const int idx = cc - NW;
        YW[cc] = (idx >= 0 && idx < n1) ? tr[(long)idx * n2] : 0.f;
        TM[cc] = 0.f;
#pragma unroll
        for (int j = 0; j < NA; j++) W1[cc][j] = 0.f;
    }
    for (int e = 0; e < n1 + NW; e++) {
        float a1[NA];
        float tm = 0.f;
        if (e >= NW && e < n1 - NW) {
            spray_taps<NW>(A.tb, g1[(long)e * n2], f1, a1);
#pragma unroll
            for (int j = NA - 1; j >= 0; j--) tm += a1[j] * YW[NA - 1 - j];      // io index ascending
        } else {
#pragma unroll
            for (int j = 0; j < NA; j++) a1[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < NA; j++) W1[NA - 1][j] = a1[j];
        TM[NA - 1] = tm;
        const int m = e - NW;
        if (m >= 0) {
            float o = 0.f;
#pragma unroll
            for (int cc = 0; cc < NA; cc++) {
                const int i = m - NW + cc;
                if (i >= NW && i < n1 - NW) o += W1[cc][NA - 1 - cc] * TM[cc];
            }
            if (m == 0) o += rg.eps2 * t0;
            if (m == 1) o += rg.eps2 * t1;
            if (m == n1 - 2) o += rg.eps2 * t2;
            if (m == n1 - 1) o += rg.eps2 * t3;
            tr[(long)m * n2] = o;
        }
#pragma unroll
        for (int cc = 0; cc < NA - 1; cc++) {
            YW[cc] = YW[cc + 1]; TM[cc] = TM[cc + 1];
#pragma unroll
            for (int j = 0; j < NA; j++) W1[cc][j] = W1[cc + 1][j];
        }
        {
            const int idx = e + 1 + NW;
            YW[NA - 1] = (idx < n1) ? tr[(long)idx * n2] : 0.f;
        }
    }
}

// in = ((in + trP) + trM) + (data * w_centre) * ws      (pwspray_lop(adj) :1985-2001)
__global__ void __launch_bounds__(256)
adj_combine_kernel(float *__restrict__ in, const float *__restrict__ trP, const float *__restrict__ trM,
                   const float *__restrict__ data, const float *__restrict__ tnorm, float wc, int add, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float tv = tnorm[i];
        const float ws = (0.0f != tv) ? (float)(1.0 / (double)tv) : 0.0f;
        float v = add ? in[i] : 0.f;
        v += trP[i];
        v += trM[i];
        v += data[i] * wc * ws;
        in[i] = v;
    }
}

// 2-D plane-wave smoother over the traces of every panel of a trace-minor volume
struct Smoother2 {
    int n1, nt, npanel, ns, nw;        // nt traces per panel
    float eps_reg;
    const float *dip;                  // slopes (trace-minor)
    float *tnorm;                      // spray of ones through the weights (pwsmooth_set)
    SprayPlan P;
    int t_off = 0, ntg = 0;            // cut panels (distributed xline smoother): see PredArgs
};

static void smoother_plan(Smoother2 &S)
{
    SprayPlan &P = S.P;
    P = SprayPlan{};
    P.n1 = S.n1; P.n2 = S.nt; P.n3 = S.npanel; P.ns2 = S.ns; P.ns3 = 0; P.nw = S.nw;
    P.np2 = 2 * S.ns + 1; P.np3 = 1; P.np = P.np2;
    P.eps_reg = S.eps_reg;
    P.zs0 = 0; P.zs1 = S.npanel; P.zt0 = 0; P.zt1 = S.npanel;
    P.t_off = S.t_off; P.ntg = S.ntg;
    for (int s = 0; s < P.np; s++) P.live[s] = true;
}

static int smoother_fwd(pst_ctx *c, const Smoother2 &S, const float *in, float *out, bool norm)
{
    const size_t mark = c->arena_used;
    ReduceOut R{out, 0, norm ? S.tnorm : nullptr, norm ? 1 : 0};
    PST_TRY(spray_run(c, S.P, in, S.dip, nullptr, reduce_wsum, &R));
    c->arena_used = mark;
    return PST_OK;
}

static int smoother_set(pst_ctx *c, Smoother2 &S, float *ones_scratch)
{
    const size_t n = (size_t)S.n1 * S.nt * S.npanel;
    PST_LAUNCH(c, PST_K_OTHER, (fill_kernel_s<<<pst_grid_for(c, n, 256), 256, 0, c->stream>>>(ones_scratch, 1.0f, n)));
    return smoother_fwd(c, S, ones_scratch, S.tnorm, false);
}

// in (+)= S' data ; trP / trM: two work volumes
static int smoother_adj(pst_ctx *c, const Smoother2 &S, float *in, const float *data, bool add, float *trP, float *trM)
{
    const size_t plane = (size_t)S.n1 * S.nt, n = plane * S.npanel;
    const int NC = 2 * S.nw + 1;
    PST_CUDA(cudaMemsetAsync(trP, 0, n * sizeof(float), c->stream));
    PST_CUDA(cudaMemsetAsync(trM, 0, n * sizeof(float), c->stream));
    const size_t mark = c->arena_used;
    int cz = (int)std::max<double>(1.0, std::min<double>((double)S.npanel, 3.0e9 / ((double)plane * 4.0 * NC)));
    float *scr;
    PST_TRY(pst_arena_get(c, plane * (size_t)cz * NC, &scr));
    AdjArgs A{};
    A.data = data; A.tnorm = S.tnorm; A.sg = S.dip; A.scr = scr; A.n1 = S.n1; A.n2 = S.nt;
    A.t_off = S.t_off; A.ntg = S.ntg > 0 ? S.ntg : S.nt;
    A.reg = make_reg(S.eps_reg); A.tb = make_btab_s(S.nw);
    const int threads = S.nt >= 128 ? 128 : (S.nt >= 64 ? 64 : 32);
    for (int side = 0; side < 2; side++) {
        A.tr = side == 0 ? trP : trM;
        A.forw = side == 0 ? 1 : 0;
        for (int is = S.ns - 1; is >= 0; is--) {
            A.shift = side == 0 ? (is + 1) : -(is + 1);
            A.sg_shift = side == 0 ? is : -(is + 1);                     // dip[ip-1] (forw) / dip[ip]
            const int slot = side == 0 ? S.ns + is + 1 : S.ns - is - 1;
            A.wslot = (float)(S.ns + 1 - abs(slot - S.ns));
            for (int p0 = 0; p0 < S.npanel; p0 += cz) {
                A.p0 = p0;
                dim3 grid((S.nt + threads - 1) / threads, std::min(cz, S.npanel - p0));
                PST_LAUNCHBF(c, PST_K_PREDICT, 20.0 * (double)plane * grid.y, (S.nw == 1 ? 47.0 : 116.0) * (double)plane * grid.y,
                    if (S.nw == 1) predict_adj_kernel<1><<<grid, threads, 0, c->stream>>>(A);
                    else           predict_adj_kernel<2><<<grid, threads, 0, c->stream>>>(A));
                c->stats.predictions += (long long)grid.y * S.nt;
            }
        }
    }
    PST_LAUNCH(c, PST_K_SLOTRED, (adj_combine_kernel<<<pst_grid_for(c, n, 256), 256, 0, c->stream>>>(
        in, trP, trM, data, S.tnorm, (float)(S.ns + 1), add ? 1 : 0, n)));
    PST_CUDA(cudaGetLastError());
    c->arena_used = mark;
    return PST_OK;
}

// ---- vector kernels of ps_conjgrad as csint3d runs it (dip_cfuns.c-style shaping CG, soint3d_cfuns.c copy)
__global__ void __launch_bounds__(256)
sint_init_kernel(const float *__restrict__ d, float *__restrict__ p, float *__restrict__ r, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = d[i];
        p[i] = v; r[i] = -v;
    }
}
// r += L x (known samples) ; partial r.r
__global__ void __launch_bounds__(256)
sint_resid_kernel(float *__restrict__ r, const float *__restrict__ x, const unsigned char *__restrict__ known, Span Sp,
                  double *__restrict__ partial)
{
    double acc[1] = {0.};
    size_t i0, i1, step;
    pst_span(Sp, 1, i0, i1, step);
    for (size_t i = i0; i < i1; i += step) {
        float v = r[i];
        if (known[i]) v += x[i];
        r[i] = v;
        acc[0] += (double)v * v;
    }
    pst_block_reduce<1>(acc, partial);
}
// gp = eps p ; gx = -eps x (+ r where known)
__global__ void __launch_bounds__(256)
sint_grad_kernel(const float *__restrict__ p, const float *__restrict__ x, const float *__restrict__ r,
                 const unsigned char *__restrict__ known, float eps, float *__restrict__ gp, float *__restrict__ gx, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        gp[i] = eps * p[i];
        float g = -eps * x[i];
        if (known[i]) g += r[i];
        gx[i] = g;
    }
}
// gr = L gx ; s = g (+ alpha s) with the reference's swap ; partials: gn is formed before (separate kernel)
template <bool FIRST>
__global__ void __launch_bounds__(256)
sint_dir_kernel(float *__restrict__ gp, float *__restrict__ gx, const unsigned char *__restrict__ known,
                float *__restrict__ sp, float *__restrict__ sx, float *__restrict__ sr, float alpha, Span Sp,
                double *__restrict__ partial)
{
    double acc[3] = {0., 0., 0.};
    size_t i0, i1, step;
    pst_span(Sp, 1, i0, i1, step);
    for (size_t i = i0; i < i1; i += step) {
        const float gpi = gp[i], gxi = gx[i];
        float gri = 0.f;
        if (known[i]) gri += gxi;
        float a, b, cc;
        if (FIRST) { a = gpi; b = gxi; cc = gri; }
        else {
            a = gpi + alpha * sp[i];
            b = gxi + alpha * sx[i];
            cc = gri + alpha * sr[i];
        }
        sp[i] = a; sx[i] = b; sr[i] = cc;
        acc[0] += (double)cc * cc;
        acc[1] += (double)a * a;
        acc[2] += (double)b * b;
    }
    pst_block_reduce<3>(acc, partial);
}
__global__ void __launch_bounds__(256)
sint_update_kernel(float *__restrict__ p, float *__restrict__ x, float *__restrict__ r, const float *__restrict__ sp,
                   const float *__restrict__ sx, const float *__restrict__ sr, float alpha, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        p[i] += alpha * sp[i];
        x[i] += alpha * sx[i];
        r[i] += alpha * sr[i];
    }
}
__global__ void __launch_bounds__(256)
sint_sumsq_kernel(const float *__restrict__ v, Span Sp, double *__restrict__ partial)
{
    double acc[1] = {0.};
    size_t i0, i1, step;
    pst_span(Sp, 1, i0, i1, step);
    for (size_t i = i0; i < i1; i += step)
        acc[0] += (double)v[i] * v[i];
    pst_block_reduce<1>(acc, partial);
}
// known = (mask != 0), partial count
__global__ void __launch_bounds__(256)
sint_known_kernel(const float *__restrict__ mask, unsigned char *__restrict__ known, Span Sp, double *__restrict__ partial)
{
    double acc[1] = {0.};
    size_t i0, i1, step;
    pst_span(Sp, 1, i0, i1, step);
    for (size_t i = i0; i < i1; i += step) {
        const unsigned char k = mask[i] != 0.f;
        known[i] = k;
        acc[0] += k;
    }
    pst_block_reduce<1>(acc, partial);
}

// Distributed contexts (n3-slabs): n3 is the GLOBAL plane count, the pointers are this rank's slab.  The inline
// smoother works inside (n1 x n2) panels, i.e. inside the slab.  The xline smoother's panels are (n1 x n3): their
// traces are cut over the ranks, so each application first exchanges an ns2-plane halo of ITS INPUT with both
// neighbours (forward: the inline-smoothed model; adjoint: the data -- the adjoint is evaluated per model trace as a
// gather over the data traces within ns2, so it needs the same halo, not a halo-add) and then runs on the extended
// local trace range with global edge tests; halo traces are computed redundantly and dropped.  The normalisation of
// the halo traces comes from their owner (exchanged once).  The CG dots are all-reduced.  Same bits as one GPU.
extern "C" int pst_sint3d_dev(pst_ctx *c, const float *d_din, const float *d_dipi, const float *d_dipx,
                              const float *d_mask, int n1, int n2, int n3, int niter, int ns1, int ns2,
                              int order1, int order2, int verb, float eps, float *d_out)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    if (!d_din || !d_dipi || !d_dipx || !d_mask || !d_out) { pst_set_error("sint3d: null pointer"); return PST_EINVAL; }
    PST_TRY(check_spray_args(n1, n2, n3, ns1, 0, order1));
    PST_TRY(check_spray_args(n1, n3, n2, ns2, 0, order2));
    if (niter < 0) { pst_set_error("sint3d: niter < 0"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    const bool dist = c->comm != nullptr && c->nranks > 1;
    int z0 = 0, z1 = n3;
    if (dist) {
        z0 = (int)(((long)n3 * c->rank) / c->nranks);
        z1 = (int)(((long)n3 * (c->rank + 1)) / c->nranks);
        if (n3 / c->nranks < std::max(ns2, 1)) { pst_set_error("sint3d: slabs thinner than the xline smoothing radius"); return PST_EUNSUP; }
    }
    const int nz = z1 - z0;
    const int ze0 = std::max(0, z0 - ns2), ze1 = std::min(n3, z1 + ns2), nzl = ze1 - ze0, offs = z0 - ze0;
    const size_t plane = (size_t)n1 * n2;
    const size_t n = plane * nz, nex = plane * nzl;               // slab / slab + halos
    const double nglob = (double)plane * n3;
    const int NCmax = 2 * std::max(order1, order2) + 1, nsmax = std::max(ns1, ns2);
    // 9 CG vectors + 2 dips + 2 norms + 5 work volumes + mask + spray slots/scratch (bounded chunks)
    const double chunk = std::min<double>(spray_chunk_bytes() + 2.0 * 4.0 * nex, (double)nex * 4.0 * (2 * nsmax + 1 + NCmax));
    PST_TRY(pst_arena_reserve(c, (size_t)((11 * n + 9 * nex) * sizeof(float) + n + chunk + 3.0e9 + 64 * 4096)));
    pst_arena_reset(c);
    float *dipA, *dipB, *tnA, *tnB, *p, *x, *r, *sp, *sx, *sr, *gp, *gx, *wA1, *wA2, *wB1, *wB2, *dA, *eA;
    unsigned char *known;
    float **slabv[] = {&dipA, &tnA, &p, &x, &r, &sp, &sx, &sr, &gp, &gx, &dA};
    float **extv[] = {&dipB, &tnB, &wA1, &wA2, &wB1, &wB2, &eA};
    for (float **q : slabv) PST_TRY(pst_arena_get(c, n, q));
    for (float **q : extv) PST_TRY(pst_arena_get(c, nex, q));
    PST_TRY(pst_arena_get(c, n, &known));
    const int threads = 256, grid = pst_grid_for(c, n, threads);
    // canonical sums (pst_common.cuh): the CG scalars do not depend on the slab decomposition
    const Span Sp = pst_span_canon(plane, nz, n3);
    const unsigned gridc = Sp.ppp * (unsigned)nz;
    PST_TRY(pst_reserve_partials(c, gridc, n3));
    auto finish = [&](int nv, int rec) { return pst_finish_reduce_canon(c, (int)Sp.ppp, nz, z0, n3, nv, rec); };
    double h[PST_RED_SLOTS];
    float *const eS = eA + (size_t)offs * plane;                 // the slab inside the extended A volume

    // ns2 planes of the slab held in eA -> the neighbours' halos, theirs -> mine
    auto halo = [&]() -> int {
        if (!dist || ns2 == 0) return PST_OK;
        const size_t cnt = plane * (size_t)ns2;
        return pst_comm_halo_exchange(c, eS, eS + n - cnt, eA, eS + n, cnt);
    };

    // reference layout [i3][i2][i1] -> A = [i3][i1][i2];  B = [i2][i1][i3] = swap(A), i3 over slab + halos
    PST_TRY(transpose_planes(c, d_dipi, dipA, n2, n1, nz));
    PST_TRY(transpose_planes(c, d_dipx, eS, n2, n1, nz));
    PST_TRY(halo());
    PST_TRY(swap_ab(c, eA, dipB, nzl, n1, n2));
    PST_TRY(transpose_planes(c, d_din, dA, n2, n1, nz));
    PST_TRY(transpose_planes(c, d_mask, wA1, n2, n1, nz));
    PST_LAUNCH(c, PST_K_OTHER, (sint_known_kernel<<<gridc, threads, 0, c->stream>>>(wA1, known, Sp, c->d_partial)));
    PST_TRY(finish(1, 8));
    PST_TRY(pst_fetch_record(c, 8, 1, h));
    // "lam += 1." on a float saturates at 2^24 (soint3d_cfuns.c:2583-2592)
    float lam = (float)std::min(h[0], 16777216.0);
    lam = sqrtf(lam / (float)nglob);
    const float ceps = lam * lam;
    const double tol = 10 * 1.19209290e-07F;

    Smoother2 SA{n1, n2, nz, ns1, order1, eps * eps, dipA, tnA, {}}, SB{n1, nzl, n2, ns2, order2, eps * eps, dipB, tnB, {}};
    SB.t_off = ze0; SB.ntg = n3;
    smoother_plan(SA);
    smoother_plan(SB);
    PST_TRY(smoother_set(c, SA, wA1));
    PST_TRY(smoother_set(c, SB, wB1));
    if (dist && ns2 > 0) {
        // the normalisation of my halo traces lacks the sources beyond them: take it from the owning rank
        PST_TRY(swap_ab(c, tnB, eA, n2, n1, nzl));
        PST_TRY(halo());
        PST_TRY(swap_ab(c, eA, tnB, nzl, n1, n2));
    }
    // S (forward): out = swap(SB(swap(SA(in))))
    auto S_fwd = [&](const float *in, float *out) -> int {
        PST_TRY(smoother_fwd(c, SA, in, eS, true));
        PST_TRY(halo());
        PST_TRY(swap_ab(c, eA, wB1, nzl, n1, n2));
        PST_TRY(smoother_fwd(c, SB, wB1, wB2, true));
        PST_TRY(swap_ab(c, wB2, out, n2, n1, nzl, true, offs, nz));
        return PST_OK;
    };
    // S' (adjoint, accumulating): io += SA'(swap(SB'(swap(data))))
    auto S_adj_add = [&](float *io, const float *data) -> int {
        if (dist) {
            PST_CUDA(cudaMemcpyAsync(eS, data, n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
            PST_TRY(halo());
            PST_TRY(swap_ab(c, eA, wB1, nzl, n1, n2));
        } else {
            PST_TRY(swap_ab(c, data, wB1, n3, n1, n2));
        }
        // work volumes for the xline side: wB2 (result), wA1/wA2 reinterpreted as B-layout scratch (same size)
        PST_TRY(smoother_adj(c, SB, wB2, wB1, false, wA1, wA2));
        PST_TRY(swap_ab(c, wB2, wA1, n2, n1, nzl, false, offs, nz));
        PST_TRY(smoother_adj(c, SA, io, wA1, true, wA2, wB1));
        return PST_OK;
    };

    PST_LAUNCH(c, PST_K_OTHER, (sint_init_kernel<<<grid, threads, 0, c->stream>>>(dA, p, r, n)));
    PST_TRY(S_fwd(p, x));
    PST_LAUNCH(c, PST_K_OTHER, (sint_resid_kernel<<<gridc, threads, 0, c->stream>>>(r, x, known, Sp, c->d_partial)));
    PST_TRY(finish(1, 8));
    PST_TRY(pst_fetch_record(c, 8, 1, h));
    if (h[0] != 0.) {
        double gn, gnp = 0., alpha, beta, g0 = 0., dg;
        for (int iter = 0; iter < niter; iter++) {
            PST_LAUNCH(c, PST_K_CGHEAD, (sint_grad_kernel<<<grid, threads, 0, c->stream>>>(p, x, r, known, ceps, gp, gx, n)));
            PST_TRY(S_adj_add(gp, gx));
            PST_TRY(S_fwd(gp, gx));
            PST_LAUNCH(c, PST_K_OTHER, (sint_sumsq_kernel<<<gridc, threads, 0, c->stream>>>(gp, Sp, c->d_partial)));
            PST_TRY(finish(1, 8));
            PST_TRY(pst_fetch_record(c, 8, 1, h));
            gn = h[0];
            if (iter == 0) {
                g0 = gn;
                PST_LAUNCH(c, PST_K_CGDIR, (sint_dir_kernel<true><<<gridc, threads, 0, c->stream>>>(gp, gx, known, sp, sx, sr, 0.f, Sp, c->d_partial)));
            } else {
                alpha = gn / gnp;
                dg = gn / g0;
                if (alpha < tol || dg < tol) break;
                PST_LAUNCH(c, PST_K_CGDIR, (sint_dir_kernel<false><<<gridc, threads, 0, c->stream>>>(gp, gx, known, sp, sx, sr, (float)alpha, Sp, c->d_partial)));
            }
            PST_TRY(finish(3, 9));
            PST_TRY(pst_fetch_record(c, 9, 3, h));
            beta = h[0] + (double)ceps * (h[1] - h[2]);
            alpha = -gn / beta;
            if (verb) printf("[pst] sint3d iteration %d gn %g\n", iter + 1, gn);
            PST_LAUNCH(c, PST_K_CGHEAD, (sint_update_kernel<<<grid, threads, 0, c->stream>>>(p, x, r, sp, sx, sr, (float)alpha, n)));
            gnp = gn;
            c->stats.cg_iterations++;
        }
    }
    PST_TRY(transpose_planes(c, x, d_out, n1, n2, nz));
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

// cpaint2d / cpaint3d (reference pyseistr/src/paint_cfuns.c:1861-2024; the two entries are the same code): plane-wave
// painting.  The seed trace sits at trace i0 and is predicted outwards one trace at a time -- leftwards
// predict_step(false, false, trace, pp[i2]), rightwards predict_step(false, true, trace, pp[i2-1]) -- so the whole
// operator is ONE dependent chain of n2 - 1 trace predictions: a launch of the prediction kernel per trace, one thread
// each.  Latency-bound by construction; provided for completeness of the predict_step consumers.
extern "C" int pst_paint2d_dev(pst_ctx *c, const float *d_dip, const float *d_seed, int n1, int n2, int order, int i0,
                               float eps, float *d_out)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    if (!d_dip || !d_seed || !d_out) { pst_set_error("paint2d: null pointer"); return PST_EINVAL; }
    PST_TRY(check_spray_args(n1, n2, 1, 0, 0, order));
    if (i0 < 0 || i0 >= n2) { pst_set_error("paint2d: reference trace i0=%d outside [0, %d)", i0, n2); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    const int NC = 2 * order + 1;
    PST_TRY(pst_arena_reserve(c, (size_t)n1 * NC * sizeof(float) + 4096));
    pst_arena_reset(c);
    float *scr;
    PST_TRY(pst_arena_get(c, (size_t)n1 * NC, &scr));
    PST_CUDA(cudaMemcpyAsync(d_out + (size_t)i0 * n1, d_seed, (size_t)n1 * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    PredArgs A{};
    A.n1 = n1; A.n2 = 1; A.n3 = 1; A.ze0 = 0; A.zla = 0; A.zlb = 1; A.a = 0; A.b = 0; A.t_off = 0; A.ntg = 1;
    A.reg = make_reg(eps * eps); A.tb = make_btab_s(order); A.scr = scr;
    for (int side = 0; side < 2; side++) {
        const int step = side == 0 ? -1 : 1;
        for (int i2 = i0 + step; i2 >= 0 && i2 < n2; i2 += step) {
            A.in1 = d_out + (size_t)(i2 - step) * n1;
            A.sg1 = d_dip + (size_t)(side == 0 ? i2 : i2 - 1) * n1;
            A.forw1 = side;
            A.out = d_out + (size_t)i2 * n1;
            PST_LAUNCHBF(c, PST_K_PREDICT, 12.0 * n1, (order == 1 ? 47.0 : 116.0) * n1,
                if (order == 1) predict_warp_kernel<1, false><<<1, 128, 0, c->stream>>>(A);
                else            predict_warp_kernel<2, false><<<1, 128, 0, c->stream>>>(A));
            c->stats.predictions++;
        }
    }
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

// csomean2d with adj = 1 (sof_cfuns.c:1503-1508): out = S' din per slice
int pst_somean2d_adj_dev(pst_ctx *c, const float *d_din, const float *d_dip, int n1, int n2, int n3, int ns,
                         int order, float eps, float *d_out)
{
    PST_TRY(check_spray_args(n1, n2, n3, ns, 0, order));
    PST_CUDA(cudaSetDevice(c->device));
    // distributed contexts: n3 is the global plane count; the (n1 x n2) panels are independent, each rank smooths its own
    if (c->comm && c->nranks > 1) n3 = (int)(((long)n3 * (c->rank + 1)) / c->nranks) - (int)(((long)n3 * c->rank) / c->nranks);
    const size_t n = (size_t)n1 * n2 * n3;
    const int NC = 2 * order + 1;
    const double chunk = std::min<double>(spray_chunk_bytes() + 2.0 * 4.0 * n, (double)n * 4.0 * (2 * ns + 1 + NC));
    PST_TRY(pst_arena_reserve(c, (size_t)(7 * n * sizeof(float) + chunk + 3.0e9 + 64 * 4096)));
    pst_arena_reset(c);
    float *dipA, *tn, *dA, *oA, *w1, *w2;
    float **all[] = {&dipA, &tn, &dA, &oA, &w1, &w2};
    for (float **q : all) PST_TRY(pst_arena_get(c, n, q));
    PST_TRY(transpose_planes(c, d_dip, dipA, n2, n1, n3));
    PST_TRY(transpose_planes(c, d_din, dA, n2, n1, n3));
    Smoother2 S{n1, n2, n3, ns, order, eps * eps, dipA, tn, {}};
    smoother_plan(S);
    PST_TRY(smoother_set(c, S, w1));
    PST_TRY(smoother_adj(c, S, oA, dA, false, w1, w2));
    PST_TRY(transpose_planes(c, oA, d_out, n1, n2, n3));
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}
