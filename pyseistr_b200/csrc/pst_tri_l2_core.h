// Triangle smoothing of 32-sample blocks of one line with CHECKPOINT + RECOMPUTE, re-read served from L2
// (the per-line core of pst_tri_l2.cu).  Same arithmetic as ps_smooth2 (reference dip_cfuns.c:458-484,508-529,
// 564-580,616-625), bit for bit:
//   t_k = ((-wt x_k) + 2wt x_{k-nb}) + (-wt x_{k-2nb})          k in [0, np),  np = nx + 2nb
//   F_k = F_{k-1} + t_k        forward running sum  (float, sequential)
//   B_k = B_{k+1} + F_k        backward running sum (float, sequential)
//   y_i = (B_{i+nb} + B_{nb+nx+(nx-1-i)}[i >= nx-nb]) + B_{nb-1-i}[i < nb]
// Pass A walks the line upwards in blocks of 32 samples and keeps only F before every block (the checkpoints);
// pass B walks the blocks downwards, takes the block's x again, repeats its 32 forward additions from the checkpoint
// (same operations, same operands: same bits), runs the backward sum over them and applies the fold.  Against
// pst_tri_rc_core.h (same idea) this core
//   * never asks for a block that lies outside [0, nx) (the caller's block stream is exactly: blocks 0 .. nld-1
//     upwards, then min(nblk-1, nld) - 1 .. 0 downwards, nld = ceil(nx / 32), nblk = ceil(np / 32)),
//   * emits its outputs as ALIGNED windows of 32 samples [32 b, 32 b + 32) (the nb outputs a block owes to the window
//     above it are carried in registers), so that a window can leave as one box / 32 coalesced rows and an in-place
//     line never overwrites a sample that is still to be read: window b is exactly x block b, consumed before it is
//     stored,
//   * forms the two products -wt x and 2wt x once per sample and block (they are the operands of three different
//     steps; the sums keep the reference's association ((-wt x_k) + 2wt x_{k-nb}) + (-wt x_{k-2nb})), keeps only the
//     3 nb products a block needs from the block below it, and alternates two register sets in pass B so that no block
//     is copied.
// IO is the caller's block transport (warp-uniform on the GPU):
//   load(x)            the next block of the stream above, as 32 values of this line (zero beyond nx)
//   store(v, i0)       y[i0 + j] = v[j] for j in [0, 32), i0 + j < nx
// The header compiles for the host too (tests/test_tri_l2_core.py checks it against the CPU restatement).
#pragma once

#ifdef __CUDACC__
#define PST_L2_HD __host__ __device__ __forceinline__
#else
#define PST_L2_HD inline
#endif
// full unrolling matters on the device only (register arrays); host compilers do not know the pragma
#ifdef __CUDA_ARCH__
#define PST_L2_UNROLL _Pragma("unroll")
#else
#define PST_L2_UNROLL
#endif

namespace tri_l2 {

constexpr int RC = 32;

// number of blocks pass B takes again (pass A takes nld)
PST_L2_HD int reload_count(int nx, int nb)
{
    const int nblk = (nx + 2 * nb + RC - 1) / RC, nld = (nx + RC - 1) / RC;
    return nblk - 1 < nld ? nblk - 1 : nld;
}

// t and the forward sum over one block.  xhi = x of the block; mt / wtl = the products -wt x (last 2 nb samples) and
// 2wt x (last nb samples) of the block below.  KEEP: F of every step to Fb.  TAILS: leave this block's tails in mt / wtl.
template <int NB, bool KEEP, bool TAILS>
PST_L2_HD void fwd_block(const float *xhi, float *mt, float *wtl, float wm, float w2, float &F, float *Fb)
{
    static_assert(2 * NB <= RC, "block shorter than the stencil");
    float m[RC], w[RC];
PST_L2_UNROLL
    for (int j = 0; j < RC; j++) { m[j] = wm * xhi[j]; w[j] = w2 * xhi[j]; }
PST_L2_UNROLL
    for (int j = 0; j < RC; j++) {
        float v = m[j];             // 0 + wm x_k: a zero's sign cannot reach F (F starts at +0 and never becomes -0)
        v = v + ((j >= NB) ? w[j - NB] : wtl[j]);
        v = v + ((j >= 2 * NB) ? m[j - 2 * NB] : mt[j]);
        F = F + v;
        if (KEEP) Fb[j] = F;
    }
    if (TAILS) {
PST_L2_UNROLL
        for (int j = 0; j < NB; j++) wtl[j] = w[RC - NB + j];
PST_L2_UNROLL
        for (int j = 0; j < 2 * NB; j++) mt[j] = m[RC - 2 * NB + j];
    }
}

// one step of pass B: block b with x in xhi; the block below it is loaded into xlo (or zeroed)
template <int NB, class IO>
PST_L2_HD void back_step(IO &io, int b, int nx, int np, int nld, float wm, float w2, const float *ck, int cks,
                         float *xlo, const float *xhi, float &Bs, float *top, float *carry)
{
    if (b - 1 >= 0 && b - 1 < nld) io.load(xlo);
    else {
PST_L2_UNROLL
        for (int j = 0; j < RC; j++) xlo[j] = 0.f;
    }
    float mt[2 * NB], wtl[NB];
PST_L2_UNROLL
    for (int j = 0; j < NB; j++) wtl[j] = w2 * xlo[RC - NB + j];
PST_L2_UNROLL
    for (int j = 0; j < 2 * NB; j++) mt[j] = wm * xlo[RC - 2 * NB + j];
    float Fb[RC];
    float Fs = ck[(long)b * cks];
    fwd_block<NB, true, false>(xhi, mt, wtl, wm, w2, Fs, Fb);
    const int k0 = b * RC;
    if (k0 + RC <= np) {
PST_L2_UNROLL
        for (int j = RC - 1; j >= 0; j--) { Bs = Bs + Fb[j]; Fb[j] = Bs; }
    } else {
PST_L2_UNROLL
        for (int j = RC - 1; j >= 0; j--) { if (k0 + j < np) Bs = Bs + Fb[j]; Fb[j] = Bs; }
    }
    if (!(k0 >= 2 * NB && k0 + RC <= nx)) {
        // fold2 (:458-484): B of the top nb samples is added to the last nb outputs (right reflection, first),
        // B of the bottom nb samples to the first nb outputs (left reflection, second)
PST_L2_UNROLL
        for (int j = RC - 1; j >= 0; j--) {
            const int k = k0 + j;
            if (k < np) {
                if (k >= nx + NB) {
PST_L2_UNROLL
                    for (int q = 0; q < NB; q++) if (q == k - nx - NB) top[q] = Fb[j];
                } else if (k >= NB) {
                    const int i = k - NB;
                    if (i >= nx - NB) {
                        float tv = 0.f;
PST_L2_UNROLL
                        for (int q = 0; q < NB; q++) if (q == nx - 1 - i) tv = top[q];
                        Fb[j] = Fb[j] + tv;
                    }
                }
            }
        }
        if (k0 == 0) {
            // y_i = B_{i+nb} (+ right reflection, already in) + B_{nb-1-i} for i < nb: both sit in block 0
PST_L2_UNROLL
            for (int i = 0; i < NB; i++) Fb[NB + i] = Fb[NB + i] + Fb[NB - 1 - i];
        }
    }
    // window [k0, k0 + 32): y_i = B_{i+nb}; i = k0 + j' comes from this block's step j' + nb, or -- the top nb
    // samples of the window -- from the first nb steps of the block above (processed before, carried)
    float out[RC];
PST_L2_UNROLL
    for (int j = 0; j < RC - NB; j++) out[j] = Fb[j + NB];
PST_L2_UNROLL
    for (int j = 0; j < NB; j++) { out[RC - NB + j] = carry[j]; carry[j] = Fb[j]; }
    if (k0 < nx) io.store(out, k0);
}

template <int NB, class IO>
PST_L2_HD void smooth_line(IO &io, int nx, float wm, float w2, float *ck, int cks)
{
    const int np = nx + 2 * NB;
    const int nblk = (np + RC - 1) / RC, nld = (nx + RC - 1) / RC;
    float X0[RC], X1[RC];
    // ---- pass A: forward sum, keep F before every block
    float mt[2 * NB], wtl[NB];
PST_L2_UNROLL
    for (int j = 0; j < NB; j++) wtl[j] = 0.f;
PST_L2_UNROLL
    for (int j = 0; j < 2 * NB; j++) mt[j] = 0.f;
    float F = 0.f;
    for (int b = 0; b < nblk; b++) {
        if (b < nld) io.load(X0);
        else {
PST_L2_UNROLL
            for (int j = 0; j < RC; j++) X0[j] = 0.f;
        }
        ck[(long)b * cks] = F;
        fwd_block<NB, false, true>(X0, mt, wtl, wm, w2, F, nullptr);
    }
    // ---- pass B: blocks downwards; X0 holds x block nblk-1 now.  Two register sets alternate as "block" / "block below".
    float Bs = 0.f;
    float top[NB], carry[NB];
PST_L2_UNROLL
    for (int j = 0; j < NB; j++) { top[j] = 0.f; carry[j] = 0.f; }
    int b = nblk - 1;
    for (; b >= 1; b -= 2) {
        back_step<NB>(io, b, nx, np, nld, wm, w2, ck, cks, X1, X0, Bs, top, carry);
        back_step<NB>(io, b - 1, nx, np, nld, wm, w2, ck, cks, X0, X1, Bs, top, carry);
    }
    if (b == 0) back_step<NB>(io, 0, nx, np, nld, wm, w2, ck, cks, X1, X0, Bs, top, carry);
}

}  // namespace tri_l2
