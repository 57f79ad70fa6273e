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
//     above it are carried in registers), so that a window can leave as one TMA box and an in-place line never
//     overwrites a sample that is still to be read: window b is exactly x block b, consumed before it is stored.
// IO is the caller's block transport (warp-uniform on the GPU):
//   load(x)            the next block of the stream above, as 32 values of this line (zero beyond nx)
//   store(v, i0)       y[i0 + j] = v[j] for j in [0, 32), i0 + j < nx
// The header compiles for the host too (tests/test_tri_l2_core.py checks it against the oracle).
#pragma once
#include "pst_tri_rc_core.h"

namespace tri_l2 {

constexpr int RC = 32;

// number of blocks pass B takes again (pass A takes nld)
PST_RC_HD int reload_count(int nx, int nb)
{
    const int nblk = (nx + 2 * nb + RC - 1) / RC, nld = (nx + RC - 1) / RC;
    return nblk - 1 < nld ? nblk - 1 : nld;
}

template <int NB, class IO>
PST_RC_HD void smooth_line(IO &io, int nx, float wm, float w2, float *ck, int cks)
{
    static_assert(2 * NB <= RC, "block shorter than the stencil");
    const int np = nx + 2 * NB;
    const int nblk = (np + RC - 1) / RC, nld = (nx + RC - 1) / RC;
    float xlo[RC], xhi[RC];
    // ---- pass A: forward sum, keep F before every block
PST_RC_UNROLL
    for (int j = 0; j < RC; j++) xlo[j] = 0.f;
    float F = 0.f;
    for (int b = 0; b < nblk; b++) {
        if (b < nld) io.load(xhi);
        else {
PST_RC_UNROLL
            for (int j = 0; j < RC; j++) xhi[j] = 0.f;
        }
        ck[(long)b * cks] = F;
        tri_rc::fwd_block<NB, RC, false>(xlo, xhi, wm, w2, F, nullptr);
PST_RC_UNROLL
        for (int j = 0; j < RC; j++) xlo[j] = xhi[j];
    }
    // ---- pass B: blocks downwards; xlo holds x block nblk-1 now
PST_RC_UNROLL
    for (int j = 0; j < RC; j++) xhi[j] = xlo[j];
    float Bs = 0.f;
    float top[NB], carry[NB];
PST_RC_UNROLL
    for (int j = 0; j < NB; j++) { top[j] = 0.f; carry[j] = 0.f; }
    for (int b = nblk - 1; b >= 0; b--) {
        if (b - 1 >= 0 && b - 1 < nld) io.load(xlo);
        else {
PST_RC_UNROLL
            for (int j = 0; j < RC; j++) xlo[j] = 0.f;
        }
        float Fb[RC];
        float Fs = ck[(long)b * cks];
        tri_rc::fwd_block<NB, RC, true>(xlo, xhi, wm, w2, Fs, Fb);
        const int k0 = b * RC;
        if (k0 + RC <= np) {
PST_RC_UNROLL
            for (int j = RC - 1; j >= 0; j--) { Bs = Bs + Fb[j]; Fb[j] = Bs; }
        } else {
PST_RC_UNROLL
            for (int j = RC - 1; j >= 0; j--) { if (k0 + j < np) Bs = Bs + Fb[j]; Fb[j] = Bs; }
        }
        if (!(k0 >= 2 * NB && k0 + RC <= nx)) {
            // fold2 (:458-484): B of the top nb samples is added to the last nb outputs (right reflection, first),
            // B of the bottom nb samples to the first nb outputs (left reflection, second)
PST_RC_UNROLL
            for (int j = RC - 1; j >= 0; j--) {
                const int k = k0 + j;
                if (k < np) {
                    if (k >= nx + NB) top[k - nx - NB] = Fb[j];
                    else if (k >= NB) {
                        const int i = k - NB;
                        if (i >= nx - NB) Fb[j] = Fb[j] + top[nx - 1 - i];
                    }
                }
            }
            if (k0 == 0) {
                // y_i = B_{i+nb} (+ right reflection, already in) + B_{nb-1-i} for i < nb: both sit in block 0
PST_RC_UNROLL
                for (int i = 0; i < NB; i++) Fb[NB + i] = Fb[NB + i] + Fb[NB - 1 - i];
            }
        }
        // window [k0, k0 + 32): y_i = B_{i+nb}; i = k0 + j' comes from this block's step j' + nb, or -- the top nb
        // samples of the window -- from the first nb steps of the block above (processed before, carried)
        float out[RC];
PST_RC_UNROLL
        for (int j = 0; j < RC - NB; j++) out[j] = Fb[j + NB];
PST_RC_UNROLL
        for (int j = 0; j < NB; j++) { out[RC - NB + j] = carry[j]; carry[j] = Fb[j]; }
        if (k0 < nx) io.store(out, k0);
PST_RC_UNROLL
        for (int j = 0; j < RC; j++) xhi[j] = xlo[j];
    }
}

}  // namespace tri_l2
