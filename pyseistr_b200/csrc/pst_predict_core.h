// Per-trace core of the plane-wave prediction kernel (pst_spray.cu: predict_fast_kernel), as a host/device header:
// tests/test_predict_core.py compiles it with g++ (no FMA contraction, like the library) and checks it against the
// CPU restatement bit for bit.  Replaces (reference pyseistr/src/sof3d_cfuns.c) passfilter :311-329, pwd_define/pwd_set
// :401-498, regularization/predict1_step/predict2_step :548-661, sf_banded_define/solve :159-264.
#pragma once

#ifdef __CUDACC__
#define PST_PD __device__ __forceinline__
#else
#define PST_PD inline
#endif
#ifndef PST_MAXTAP
#define PST_MAXTAP 5
#endif

struct BTabS { double b[PST_MAXTAP]; };


// passfilter (sof3d_cfuns.c:311-329), taps reversed when forw (pwd_define :414-424)
template <int NW>
PST_PD void spray_taps(const BTabS &tb, float p, bool forw, float (&a)[2 * NW + 1])
{
    constexpr int NF = 2 * NW;
    float t[2 * NW + 1];
#pragma unroll
    for (int k = 0; k <= NF; k++) {
        double ak = tb.b[k];
#pragma unroll
        for (int j = 0; j < NF; j++) {
            const float f = (j < NF - k) ? ((float)(NF - j) - p) : ((p + (float)j) + 1.0f);
            ak *= (double)f;
        }
        t[k] = (float)ak;
    }
#pragma unroll
    for (int k = 0; k <= NF; k++) a[k] = forw ? t[NF - k] : t[k];
}

// regularisation constants in the reference's float/double placement (regularization :548-565)
struct RegC { float d_in, d_e0, d_e1, o0_in, o0_e, o1, eps2; };

inline BTabS make_btab_s(int nw)     // apfilt_init sof3d_cfuns.c:286-303
{
    BTabS t{};
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

inline RegC make_reg(float eps)
{
    RegC r;
    const float eps2 = eps;
    r.d_in = 6. * eps;
    r.d_e0 = eps2 + eps;
    r.d_e1 = eps2 + 5. * eps;
    r.o0_in = -4. * eps;
    r.o0_e = -2. * eps;
    r.o1 = eps;
    r.eps2 = eps2;
    return r;
}

struct PredArgs {
    // all volumes are chunk-local, trace-minor: elem(zl,k,i2) = (zl*n1 + k)*n2 + i2
    const float *in1, *in2;     // parent slot volumes
    const float *sg1, *sg2;     // slope volumes
    long in1_off, in2_off;      // parent location shift, in elements (+-1 or +-n1*n2)
    long sg1_off, sg2_off;      // slope location shift (0 = target, else = parent shift)
    int forw1, forw2;
    float *out;                 // slot volume being produced
    float *scr;                 // factor scratch [plane][k][NB+1][i2]: b/d, o[0..NB)
    int n1, n2, n3;
    int ze0;                    // global index of chunk-local plane 0
    int zla, zlb;               // chunk-local plane range to produce
    int a, b;                   // slot offsets: source = target - (a,b)
    int t_off, ntg;             // traces of a panel cut over ranks (distributed xline smoother of sint3d): global index
                                // of local trace 0 and global trace count; whole panels: 0, n2
    RegC reg;
    BTabS tb;
};


// ---------------------------------------------------------------------------------------
// predict_fast_kernel: the same arithmetic as predict_kernel, operation for operation, with the bookkeeping removed
// from the inner loop (ncu on predict_kernel<2,1>: 566 instructions per predicted sample, of which ~40 % are window
// shifts, 64-bit index arithmetic and boundary predicates):
//   * every sliding window (tap rows, W x products, inputs, LDL' history, load queues) is a ring indexed by
//     (sample mod NA): the kernel is unrolled by NA steps with the rotation as a template parameter, nothing is shifted;
//   * interior columns 2 NW <= i < n1 - 2 NW run a variant without any boundary test (EDGE = false); the first and
//     last few columns run the general variant;
//   * loads and stores walk running pointers.
template <int NW, bool TWO>
struct PredRing {
    static constexpr int NA = 2 * NW + 1, NB = 2 * NW;
    float W1[NA][NA], T1[NA], X1[NA];
    float W2[TWO ? NA : 1][NA], T2[TWO ? NA : 1], X2[TWO ? NA : 1];
    float O[NA][NB], D[NA], Bh[NA];          // LDL' history ring: column k at slot k mod NA (NB of them are live)
    float gq1[NA], gq2[TWO ? NA : 1], xq1[NA], xq2[TWO ? NA : 1];
    const float *g1, *g2, *x1, *x2;          // running pointers: slope sample kk + NA, input sample i + 2 NA of the step being run
    float *sc;                               // scratch column of the step being run
};

// slot of sample (base + d), base = R mod NA known at compile time
#define PST_SLOT(R, d) ((((R) + (d)) % NA + NA) % NA)

template <int NW, bool TWO, bool EDGE, int R>
PST_PD void predict_step(PredRing<NW, TWO> &S, const PredArgs &A, const RegC &rg, int i, int n1, long n2, bool f1, bool f2)
{
    constexpr int NA = 2 * NW + 1, NB = 2 * NW, NC = NB + 1;
    const int kk = i + NW;                                   // leading sample: its tap row enters the ring
    const float gv1 = S.gq1[R], xn1 = S.xq1[R];
    const float gv2 = TWO ? S.gq2[TWO ? R : 0] : 0.f, xn2 = TWO ? S.xq2[TWO ? R : 0] : 0.f;
    {                                                        // refill the queue slot for step i + NA
        // (guarded in the interior variant too: its last steps refill for steps beyond the trace)
        if (kk + NA < n1) { S.gq1[R] = *S.g1; if (TWO) S.gq2[TWO ? R : 0] = *S.g2; }
        else { S.gq1[R] = 0.f; if (TWO) S.gq2[TWO ? R : 0] = 0.f; }
        if (i + 2 * NA < n1) { S.xq1[R] = *S.x1; if (TWO) S.xq2[TWO ? R : 0] = *S.x2; }
        else { S.xq1[R] = 0.f; if (TWO) S.xq2[TWO ? R : 0] = 0.f; }
        S.g1 += n2; S.x1 += n2;
        if (TWO) { S.g2 += n2; S.x2 += n2; }
    }
    {
        constexpr int SK = PST_SLOT(R, NW);                  // slot of sample kk
        float a1[NA], a2[NA];
        float tm1 = 0.f, tm2 = 0.f;
        if (!EDGE || kk < n1) {
            spray_taps<NW>(A.tb, gv1, f1, a1);
            if (TWO) spray_taps<NW>(A.tb, gv2, f2, a2);
            if (!EDGE || (kk >= NW && kk < n1 - NW)) {       // pwd_set :481-486
#pragma unroll
                for (int j = 0; j < NA; j++) {
                    tm1 += a1[j] * S.X1[PST_SLOT(R, j)];     // inp[kk - NW + j] = inp[i + j]
                    if (TWO) tm2 += a2[j] * S.X2[TWO ? PST_SLOT(R, j) : 0];
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < NA; j++) { a1[j] = 0.f; a2[j] = 0.f; }
        }
#pragma unroll
        for (int j = 0; j < NA; j++) { S.W1[SK][j] = a1[j]; if (TWO) S.W2[TWO ? SK : 0][j] = a2[j]; }
        S.T1[SK] = tm1;
        if (TWO) S.T2[TWO ? SK : 0] = tm2;
    }
    if (!EDGE || i >= 0) {
        // ---- matrix column i: regularisation + W'W (pwd_define :426-444); row k = i + j - NW sits in slot (R + j - NW)
        float dg = rg.d_in;
        float of[NB];
        of[0] = rg.o0_in;
        if (EDGE) {
            if (i == 0 || i == n1 - 1) dg = rg.d_e0;
            if (i == 1 || i == n1 - 2) dg = rg.d_e1;
            if (i == 0 || i == n1 - 2) of[0] = rg.o0_e;
        }
        of[1] = rg.o1;
#pragma unroll
        for (int m = 2; m < NB; m++) of[m] = 0.0f;
        float rhs1 = 0.f, rhs2 = 0.f;
#pragma unroll
        for (int j = 0; j < NA; j++) {
            const int k = i + j - NW;
            if (!EDGE || (k >= NW && k < n1 - NW)) { const float aj = S.W1[PST_SLOT(R, j - NW)][j]; dg += aj * aj; }
        }
#pragma unroll
        for (int m = 0; m < NB; m++) {
#pragma unroll
            for (int j = m + 1; j < NA; j++) {
                const int k = i + j - NW;
                if (!EDGE || (k >= NW && k < n1 - NW)) of[m] += S.W1[PST_SLOT(R, j - NW)][j - m - 1] * S.W1[PST_SLOT(R, j - NW)][j];
            }
        }
        if (TWO) {
#pragma unroll
            for (int j = 0; j < NA; j++) {
                const int k = i + j - NW;
                if (!EDGE || (k >= NW && k < n1 - NW)) { const float aj = S.W2[TWO ? PST_SLOT(R, j - NW) : 0][j]; dg += aj * aj; }
            }
#pragma unroll
            for (int m = 0; m < NB; m++) {
#pragma unroll
                for (int j = m + 1; j < NA; j++) {
                    const int k = i + j - NW;
                    if (!EDGE || (k >= NW && k < n1 - NW))
                        of[m] += S.W2[TWO ? PST_SLOT(R, j - NW) : 0][j - m - 1] * S.W2[TWO ? PST_SLOT(R, j - NW) : 0][j];
                }
            }
        }
        // ---- rhs (pwd_set :487-496), end terms (predict1/2_step :611-619,:640-658)
#pragma unroll
        for (int j = 0; j < NA; j++) {
            const int k = i + j - NW;
            if (!EDGE || (k >= NW && k < n1 - NW)) {
                rhs1 += S.W1[PST_SLOT(R, j - NW)][j] * S.T1[PST_SLOT(R, j - NW)];
                if (TWO) rhs2 += S.W2[TWO ? PST_SLOT(R, j - NW) : 0][j] * S.T2[TWO ? PST_SLOT(R, j - NW) : 0];
            }
        }
        float rhs = TWO ? (rhs1 + rhs2) : rhs1;
        if (EDGE && (i < 2 || i >= n1 - 2)) {
            float te;                                         // inp[i] sits in slot R
            if (TWO) te = (float)(0.5 * (double)(S.X1[R] + S.X2[TWO ? R : 0]));
            else te = S.X1[R];
            rhs += rg.eps2 * te;
        }
        // ---- LDL' column (sf_banded_define :169-184); history h <-> column i - 1 - h in slot (R - 1 - h)
        float t = dg;
#pragma unroll
        for (int m = 0; m < NB; m++)
            if (!EDGE || m < i) t -= (S.O[PST_SLOT(R, -1 - m)][m] * S.O[PST_SLOT(R, -1 - m)][m]) * S.D[PST_SLOT(R, -1 - m)];
        const float dk = t;
        float ok[NB];
#pragma unroll
        for (int q = 0; q < NB; q++) {
            float v = of[q];
#pragma unroll
            for (int m = 0; m < NB - q - 1; m++)
                if (!EDGE || m < i) v -= (S.O[PST_SLOT(R, -1 - m)][m] * S.O[PST_SLOT(R, -1 - m)][q + m + 1]) * S.D[PST_SLOT(R, -1 - m)];
            ok[q] = (!EDGE || q < n1 - i - 1) ? v / dk : 0.f;
        }
        // ---- forward substitution (sf_banded_solve :250-256)
        float bk = rhs;
#pragma unroll
        for (int m = 0; m < NB; m++)
            if (!EDGE || m < i) bk -= S.O[PST_SLOT(R, -1 - m)][m] * S.Bh[PST_SLOT(R, -1 - m)];
        // ---- spill column to scratch: b_k / d_k and o[.][k]
        S.sc[0] = bk / dk;
#pragma unroll
        for (int q = 0; q < NB; q++) S.sc[(long)(1 + q) * n2] = ok[q];
        S.sc += (long)NC * n2;
        S.D[R] = dk; S.Bh[R] = bk;
#pragma unroll
        for (int m = 0; m < NB; m++) S.O[R][m] = ok[m];
    }
    // sample i leaves the input window, sample i + NA enters in its slot
    S.X1[R] = xn1;
    if (TWO) S.X2[TWO ? R : 0] = xn2;
}

template <int NW, bool TWO, bool EDGE, int R0, int COUNT>
struct PredUnroll {
    static PST_PD void run(PredRing<NW, TWO> &S, const PredArgs &A, const RegC &rg, int i, int n1, long n2, bool f1, bool f2)
    {
        constexpr int NA = 2 * NW + 1;
        if (!EDGE || i < n1) predict_step<NW, TWO, EDGE, ((R0 % NA) + NA) % NA>(S, A, rg, i, n1, n2, f1, f2);
        PredUnroll<NW, TWO, EDGE, R0 + 1, COUNT - 1>::run(S, A, rg, i + 1, n1, n2, f1, f2);
    }
};
template <int NW, bool TWO, bool EDGE, int R0>
struct PredUnroll<NW, TWO, EDGE, R0, 0> {
    static PST_PD void run(PredRing<NW, TWO> &, const PredArgs &, const RegC &, int, int, long, bool, bool) {}
};

template <int NW, bool TWO>
PST_PD void predict_fast_trace(const PredArgs &A, int i2, int zl, int by)
{
    constexpr int NA = 2 * NW + 1, NB = 2 * NW, NC = NB + 1;
    constexpr int P = ((2 * NW + NA - 1) / NA) * NA;          // first interior column that is a multiple of NA
    const int n1 = A.n1;
    const long n2 = A.n2;
    const long base = (long)zl * n1 * n2 + i2;
    float *out = A.out + base;
    // slot stays zero when its source lies outside the cube (csomf3d :1656)
    // (a source outside the LOCAL trace range of a cut panel feeds only halo targets nobody reads: zero as well)
    const int s2 = i2 - A.a, s3 = (A.ze0 + zl) - A.b, s2g = s2 + A.t_off;
    if (s2 < 0 || s2 >= A.n2 || s2g < 0 || s2g >= A.ntg || s3 < 0 || s3 >= A.n3) {
        for (int k = 0; k < n1; k++) out[(long)k * n2] = 0.f;
        return;
    }
    const float *x1 = A.in1 + base + A.in1_off;
    const float *g1 = A.sg1 + base + A.sg1_off;
    const float *x2 = TWO ? A.in2 + base + A.in2_off : nullptr;
    const float *g2 = TWO ? A.sg2 + base + A.sg2_off : nullptr;
    float *scr = A.scr + ((long)by * n1 * NC) * n2 + i2;
    const bool f1 = A.forw1 != 0, f2 = A.forw2 != 0;
    const RegC rg = A.reg;

    PredRing<NW, TWO> S;
#pragma unroll
    for (int c = 0; c < NA; c++) {
        S.T1[c] = 0.f; S.D[c] = 0.f; S.Bh[c] = 0.f;
        if (TWO) S.T2[TWO ? c : 0] = 0.f;
#pragma unroll
        for (int j = 0; j < NA; j++) { S.W1[c][j] = 0.f; if (TWO) S.W2[TWO ? c : 0][j] = 0.f; }
#pragma unroll
        for (int m = 0; m < NB; m++) S.O[c][m] = 0.f;
    }
    // the first step is i = -NW: the input window holds samples -NW .. NW (slot = sample mod NA), the queues the slope
    // samples kk = 0 .. NA-1 of steps -NW .. -NW+NA-1 and the input samples NW+1 .. NW+NA
#pragma unroll
    for (int c = 0; c < NA; c++) {
        const int smp = c - NW;
        constexpr int dummy = 0; (void)dummy;
        const float v1 = (smp >= 0 && smp < n1) ? x1[(long)smp * n2] : 0.f;
        const float v2 = (TWO && smp >= 0 && smp < n1) ? x2[(long)smp * n2] : 0.f;
        // slot of sample smp
        S.X1[((smp % NA) + NA) % NA] = v1;
        if (TWO) S.X2[TWO ? ((smp % NA) + NA) % NA : 0] = v2;
    }
#pragma unroll
    for (int u = 0; u < NA; u++) {
        const int i = -NW + u;                               // step
        const int kk = i + NW;                               // = u
        const int xin = i + NA;                              // input sample entering after step i
        const int r = ((i % NA) + NA) % NA;
        S.gq1[r] = (kk < n1) ? g1[(long)kk * n2] : 0.f;
        if (TWO) S.gq2[TWO ? r : 0] = (kk < n1) ? g2[(long)kk * n2] : 0.f;
        S.xq1[r] = (xin < n1) ? x1[(long)xin * n2] : 0.f;
        if (TWO) S.xq2[TWO ? r : 0] = (xin < n1) ? x2[(long)xin * n2] : 0.f;
    }
    // running pointers of the refill of step i = -NW: slope sample kk + NA = NA, input sample i + 2 NA = 2 NA - NW
    S.g1 = g1 + (long)NA * n2; S.x1 = x1 + (long)(2 * NA - NW) * n2;
    S.g2 = TWO ? g2 + (long)NA * n2 : nullptr; S.x2 = TWO ? x2 + (long)(2 * NA - NW) * n2 : nullptr;
    S.sc = scr;

    // ---- prologue: steps -NW .. P-1 (general variant)
    PredUnroll<NW, TWO, true, -NW, P + NW>::run(S, A, rg, -NW, n1, n2, f1, f2);
    // ---- interior: NA steps per trip, no boundary tests
    int i = P;
    for (; i + NA <= n1 - 2 * NW; i += NA) PredUnroll<NW, TWO, false, 0, NA>::run(S, A, rg, i, n1, n2, f1, f2);
    // ---- epilogue: fewer than NA + 2 NW columns remain
    PredUnroll<NW, TWO, true, 0, NA + 2 * NW>::run(S, A, rg, i, n1, n2, f1, f2);

    // ---- back substitution (sf_banded_solve :257-263)
    float Y[NB];
#pragma unroll
    for (int m = 0; m < NB; m++) Y[m] = 0.f;
    constexpr int PB = 4;                                   // factor columns in flight ahead of the recurrence
    float cq[PB][NC];
    const float *sp = scr + (long)(n1 - 1) * NC * n2;       // column n1 - 1
    float *op = out + (long)(n1 - 1) * n2;
#pragma unroll
    for (int u = 0; u < PB; u++) {
        const int k = n1 - 1 - u;
#pragma unroll
        for (int q = 0; q < NC; q++) cq[u][q] = (k >= 0) ? sp[(long)q * n2 - (long)u * NC * n2] : 1.f;
    }
    sp -= (long)PB * NC * n2;                               // column of the refill of the first step
    for (int k0 = n1 - 1; k0 >= 0; k0 -= PB) {
#pragma unroll
      for (int u = 0; u < PB; u++) {
        const int k = k0 - u;
        if (k < 0) break;
        float col[NC];
#pragma unroll
        for (int q = 0; q < NC; q++) col[q] = cq[u][q];
        if (k - PB >= 0) {
#pragma unroll
            for (int q = 0; q < NC; q++) cq[u][q] = sp[(long)q * n2];
        }
        sp -= (long)NC * n2;
        float t = col[0];
#pragma unroll
        for (int m = 0; m < NB; m++)
            if (m < n1 - k - 1) t -= col[1 + m] * Y[m];
        *op = t;
        op -= n2;
#pragma unroll
        for (int m = NB - 1; m > 0; m--) Y[m] = Y[m - 1];
        Y[0] = t;
      }
    }
}
#undef PST_SLOT

