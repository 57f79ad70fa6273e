// One plane-wave prediction of ONE trace by ONE WARP (the per-trace core of predict_warp_kernel, pst_spray.cu), for
// launches with few target traces (2-D panels: 860 - 1280 traces per spray level; painting: one trace; small cubes): the
// thread-per-trace kernel would leave a B200 with ~1000 threads, each alone with its 400 dependent instructions per sample.
// Here the 32 lanes share one trace:
//   * everything that is NOT the recurrence is done for 32 samples at a time, one sample per lane: tap polynomials (double
//     products), W x products, the bands of W'W + regularisation, the right-hand side;
//   * the banded LDL' recurrence and the two substitutions -- serial along the trace by nature -- run over the 32 columns of
//     a block with their inputs broadcast from shared memory, identically on every lane;
//   * loads and stores of a block are one coalesced row each when the trace is contiguous (ks = 1).
// Same arithmetic as predict_step in pst_predict_core.h (reference sof3d_cfuns.c:159-264, :311-329, :401-498, :548-661),
// operation for operation, so the results are bit-identical.
//
// The code is written phase by phase over a lane index so that it also runs on the host (PST_PW_LANES expands to a loop
// over the 32 lanes there, to the thread's own lane on the device; PST_PW_SYNC to nothing / __syncwarp()):
// tests/test_predict_core.py checks it against the CPU restatement bit for bit.
#pragma once
#include "pst_predict_core.h"

#ifdef __CUDA_ARCH__
#define PST_PW_LANES(lane) for (int lane = (int)(threadIdx.x & 31), once_ = 1; once_; once_ = 0)
#define PST_PW_SYNC() __syncwarp()
#define PST_PW_UNROLL _Pragma("unroll")
#else
#define PST_PW_LANES(lane) for (int lane = 0; lane < 32; lane++)
#define PST_PW_SYNC() ((void)0)
#define PST_PW_UNROLL
#endif

// per-warp workspace (shared memory on the device)
template <int NW, bool TWO>
struct PredWarpWS {
    static constexpr int NA = 2 * NW + 1, NB = 2 * NW;
    float x1[64], x2[TWO ? 64 : 1];                    // input samples, sample s at [s & 63]
    float W1[64][NA], W2[TWO ? 64 : 1][NA];            // tap rows, row k at [k & 63]
    float T1[64], T2[TWO ? 64 : 1];                    // W x of row k
    float cdg[32], cof[32][NB], crhs[32];              // matrix column i and right-hand side of the block's 32 columns
    float fq[32], fo[32][NB];                          // b/d and the band of the factor, block's 32 columns
};

// strided trace views: sample k of a trace at p[k * ks]
template <int NW, bool TWO>
PST_PD void predict_warp_trace(PredWarpWS<NW, TWO> &S, const float *x1, const float *g1, const float *x2, const float *g2, long ks,
                               bool f1, bool f2, const RegC &rg, const BTabS &tb, int n1, float *scr, float *out)
{
    constexpr int NA = 2 * NW + 1, NB = 2 * NW, NC = NB + 1;
    // ---- input samples 0 .. NW-1 (the first block of rows needs x up to 31 + NW; the loop loads [32 b + NW, 32 b + 32 + NW))
    PST_PW_LANES(lane) {
        if (lane < NW) {
            S.x1[lane] = lane < n1 ? x1[(long)lane * ks] : 0.f;
            if (TWO) S.x2[TWO ? lane : 0] = lane < n1 ? x2[(long)lane * ks] : 0.f;
        }
    }
    // LDL' history, identical on every lane: column i - 1 - h
    float O[NB][NB], D[NB], Bh[NB];
PST_PW_UNROLL
    for (int h = 0; h < NB; h++) {
        D[h] = 0.f; Bh[h] = 0.f;
PST_PW_UNROLL
        for (int m = 0; m < NB; m++) O[h][m] = 0.f;
    }
    const int nblk = (n1 + NW + 31) / 32;
    for (int b = 0; b < nblk; b++) {
        // ---- phase 1: input samples [32 b + NW, 32 b + 32 + NW)
        PST_PW_LANES(lane) {
            const int s = 32 * b + NW + lane;
            S.x1[s & 63] = s < n1 ? x1[(long)s * ks] : 0.f;
            if (TWO) S.x2[TWO ? (s & 63) : 0] = s < n1 ? x2[(long)s * ks] : 0.f;
        }
        PST_PW_SYNC();
        // ---- phase 2: tap row kk and its W x product (pwd_define :414-424, pwd_set :481-486)
        PST_PW_LANES(lane) {
            const int kk = 32 * b + lane;
            float a1[NA], a2[NA];
            float tm1 = 0.f, tm2 = 0.f;
            if (kk < n1) {
                spray_taps<NW>(tb, g1[(long)kk * ks], f1, a1);
                if (TWO) spray_taps<NW>(tb, g2[(long)kk * ks], f2, a2);
                if (kk >= NW && kk < n1 - NW) {
PST_PW_UNROLL
                    for (int j = 0; j < NA; j++) {
                        tm1 += a1[j] * S.x1[(kk - NW + j) & 63];
                        if (TWO) tm2 += a2[j] * S.x2[TWO ? ((kk - NW + j) & 63) : 0];
                    }
                }
            } else {
PST_PW_UNROLL
                for (int j = 0; j < NA; j++) { a1[j] = 0.f; a2[j] = 0.f; }
            }
PST_PW_UNROLL
            for (int j = 0; j < NA; j++) { S.W1[kk & 63][j] = a1[j]; if (TWO) S.W2[TWO ? (kk & 63) : 0][j] = a2[j]; }
            S.T1[kk & 63] = tm1;
            if (TWO) S.T2[TWO ? (kk & 63) : 0] = tm2;
        }
        PST_PW_SYNC();
        // ---- phase 3: matrix column i = 32 b - NW + lane: regularisation + W'W (pwd_define :426-444), right-hand side
        // (pwd_set :487-496) and end terms (predict1/2_step :611-619, :640-658)
        PST_PW_LANES(lane) {
            const int i = 32 * b - NW + lane;
            if (i >= 0 && i < n1) {
                float dg = rg.d_in;
                if (i == 0 || i == n1 - 1) dg = rg.d_e0;
                if (i == 1 || i == n1 - 2) dg = rg.d_e1;
                float of[NB];
                of[0] = (i == 0 || i == n1 - 2) ? rg.o0_e : rg.o0_in;
                of[1] = rg.o1;
PST_PW_UNROLL
                for (int m = 2; m < NB; m++) of[m] = 0.0f;
                float rhs1 = 0.f, rhs2 = 0.f;
PST_PW_UNROLL
                for (int j = 0; j < NA; j++) {
                    const int k = i + j - NW;
                    if (k >= NW && k < n1 - NW) { const float aj = S.W1[k & 63][j]; dg += aj * aj; }
                }
PST_PW_UNROLL
                for (int m = 0; m < NB; m++) {
PST_PW_UNROLL
                    for (int j = m + 1; j < NA; j++) {
                        const int k = i + j - NW;
                        if (k >= NW && k < n1 - NW) of[m] += S.W1[k & 63][j - m - 1] * S.W1[k & 63][j];
                    }
                }
                if (TWO) {
PST_PW_UNROLL
                    for (int j = 0; j < NA; j++) {
                        const int k = i + j - NW;
                        if (k >= NW && k < n1 - NW) { const float aj = S.W2[TWO ? (k & 63) : 0][j]; dg += aj * aj; }
                    }
PST_PW_UNROLL
                    for (int m = 0; m < NB; m++) {
PST_PW_UNROLL
                        for (int j = m + 1; j < NA; j++) {
                            const int k = i + j - NW;
                            if (k >= NW && k < n1 - NW) of[m] += S.W2[TWO ? (k & 63) : 0][j - m - 1] * S.W2[TWO ? (k & 63) : 0][j];
                        }
                    }
                }
PST_PW_UNROLL
                for (int j = 0; j < NA; j++) {
                    const int k = i + j - NW;
                    if (k >= NW && k < n1 - NW) {
                        rhs1 += S.W1[k & 63][j] * S.T1[k & 63];
                        if (TWO) rhs2 += S.W2[TWO ? (k & 63) : 0][j] * S.T2[TWO ? (k & 63) : 0];
                    }
                }
                float rhs = TWO ? (rhs1 + rhs2) : rhs1;
                if (i < 2 || i >= n1 - 2) {
                    float te;
                    if (TWO) te = (float)(0.5 * (double)(S.x1[i & 63] + S.x2[TWO ? (i & 63) : 0]));
                    else te = S.x1[i & 63];
                    rhs += rg.eps2 * te;
                }
                S.cdg[lane] = dg; S.crhs[lane] = rhs;
PST_PW_UNROLL
                for (int m = 0; m < NB; m++) S.cof[lane][m] = of[m];
            }
        }
        PST_PW_SYNC();
        // ---- phase 4: LDL' columns and forward substitution of the block, in order, identically on every lane
        // (sf_banded_define :169-184, sf_banded_solve :250-256)
        for (int j = 0; j < 32; j++) {
            const int i = 32 * b - NW + j;
            if (i < 0 || i >= n1) continue;
            float t = S.cdg[j];
PST_PW_UNROLL
            for (int m = 0; m < NB; m++)
                if (m < i) t -= (O[m][m] * O[m][m]) * D[m];
            const float dk = t;
            float ok[NB];
PST_PW_UNROLL
            for (int q = 0; q < NB; q++) {
                float v = S.cof[j][q];
PST_PW_UNROLL
                for (int m = 0; m < NB - q - 1; m++)
                    if (m < i) v -= (O[m][m] * O[m][q + m + 1]) * D[m];
                ok[q] = (q < n1 - i - 1) ? v / dk : 0.f;
            }
            float bk = S.crhs[j];
PST_PW_UNROLL
            for (int m = 0; m < NB; m++)
                if (m < i) bk -= O[m][m] * Bh[m];
            PST_PW_LANES(lane) {
                if (lane == j) {
                    S.fq[j] = bk / dk;
PST_PW_UNROLL
                    for (int q = 0; q < NB; q++) S.fo[j][q] = ok[q];
                }
            }
PST_PW_UNROLL
            for (int h = NB - 1; h > 0; h--) {
                D[h] = D[h - 1]; Bh[h] = Bh[h - 1];
PST_PW_UNROLL
                for (int m = 0; m < NB; m++) O[h][m] = O[h - 1][m];
            }
            D[0] = dk; Bh[0] = bk;
PST_PW_UNROLL
            for (int m = 0; m < NB; m++) O[0][m] = ok[m];
        }
        PST_PW_SYNC();
        // ---- phase 5: factor columns of the block to the scratch, [c][i] (one coalesced row per component)
        PST_PW_LANES(lane) {
            const int i = 32 * b - NW + lane;
            if (i >= 0 && i < n1) {
                scr[i] = S.fq[lane];
PST_PW_UNROLL
                for (int q = 0; q < NB; q++) scr[(long)(1 + q) * n1 + i] = S.fo[lane][q];
            }
        }
        PST_PW_SYNC();
    }
    // ---- back substitution (sf_banded_solve :257-263), blocks of 32 samples downwards
    float Y[NB];
PST_PW_UNROLL
    for (int m = 0; m < NB; m++) Y[m] = 0.f;
    for (int bb = (n1 - 1) / 32; bb >= 0; bb--) {
        PST_PW_LANES(lane) {
            const int k = 32 * bb + lane;
            if (k < n1) {
                S.fq[lane] = scr[k];
PST_PW_UNROLL
                for (int q = 0; q < NB; q++) S.fo[lane][q] = scr[(long)(1 + q) * n1 + k];
            }
        }
        PST_PW_SYNC();
        for (int j = 31; j >= 0; j--) {
            const int k = 32 * bb + j;
            if (k >= n1) continue;
            float t = S.fq[j];
PST_PW_UNROLL
            for (int m = 0; m < NB; m++)
                if (m < n1 - k - 1) t -= S.fo[j][m] * Y[m];
            PST_PW_LANES(lane) { if (lane == j) S.cdg[j] = t; }
PST_PW_UNROLL
            for (int m = NB - 1; m > 0; m--) Y[m] = Y[m - 1];
            Y[0] = t;
        }
        PST_PW_SYNC();
        PST_PW_LANES(lane) {
            const int k = 32 * bb + lane;
            if (k < n1) out[(long)k * ks] = S.cdg[lane];
        }
        PST_PW_SYNC();
    }
}
