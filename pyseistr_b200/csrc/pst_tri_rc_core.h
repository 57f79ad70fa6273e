// Triangle smoothing of one line with CHECKPOINT + RECOMPUTE (experimental second smoother, see
// pst_tri_rc.cu).  Same arithmetic as ps_smooth2 (reference dip_cfuns.c:458-484,508-529,564-580,
// 616-625), bit for bit:
//   t_k = ((-wt x_k) + 2wt x_{k-nb}) + (-wt x_{k-2nb})          k in [0, np),  np = nx + 2nb
//   F_k = F_{k-1} + t_k        forward running sum  (float, sequential)
//   B_k = B_{k+1} + F_k        backward running sum (float, sequential)
//   y_i = (B_{i+nb} + B_{nb+nx+(nx-1-i)}[i >= nx-nb]) + B_{nb-1-i}[i < nb]
// The backward sum needs F in descending k, i.e. the whole forward sum first.  Instead of keeping
// F of a line on chip (pst_tri_stream.cu: 4 B per sample, which limits an SM to 32 lines), pass A
// keeps only F at every RC-th sample (the checkpoints), and pass B walks the blocks downwards,
// re-reads the block's x, repeats the block's RC forward additions from its checkpoint (the same
// operations on the same operands: the same bits) and runs the backward sum over them.  A line then
// needs np/RC floats of state, thousands of lines are in flight per SM, and the kernel is a plain
// latency-hidden stream: 4 B read (pass A) + 4 B read + 4 B written (pass B) per sample.
//
// This header is the per-line core, shared by the CUDA kernels and by a host build
// (tests/test_tri_rc_core.py compiles it with g++ and checks it against the CPU restatement), so the
// arithmetic of the CUDA path is testable without a GPU.  IO is the line's load/store policy:
//   prefetch(m)                      start loading x block m (samples [m RC, (m+1) RC), zero outside [0, nx))
//   take(x)                          the prefetched block, as RC values of this line
//   store_block(v, i0, jlo, jhi)     y[i0 + j] = v[j] for j in [jlo, jhi)
//   store_one(i, v)                  y[i] = v
#pragma once

#ifdef __CUDACC__
#define PST_RC_HD __host__ __device__ __forceinline__
#else
#define PST_RC_HD inline
#endif

// full unrolling matters on the device only (register arrays); host compilers do not know the pragma
#ifdef __CUDA_ARCH__
#define PST_RC_UNROLL _Pragma("unroll")
#else
#define PST_RC_UNROLL
#endif

namespace tri_rc {

// RC steps of t_k and the forward sum.  xhi = x block b (x_k for the block's k), xlo = x block b-1;
// 2 NB <= RC so that the two delayed taps are in one of them.
template <int NB, int RC, bool KEEP>
PST_RC_HD void fwd_block(const float *xlo, const float *xhi, float wm, float w2, float &F, float *Fb)
{
    static_assert(2 * NB <= RC, "block shorter than the stencil");
PST_RC_UNROLL
    for (int j = 0; j < RC; j++) {
        const float xa = xhi[j];
        const float xb = (j >= NB) ? xhi[j - NB] : xlo[RC + j - NB];
        const float xc = (j >= 2 * NB) ? xhi[j - 2 * NB] : xlo[RC + j - 2 * NB];
        float v = wm * xa;          // 0 + wm x_k: a zero's sign cannot reach F (F starts at +0 and never becomes -0)
        v = v + w2 * xb;
        v = v + wm * xc;
        F = F + v;
        if (KEEP) Fb[j] = F;
    }
}

template <int NB, int RC, class IO>
PST_RC_HD void process_line(IO &io, int nx, float wm, float w2, float *ck, int cks)
{
    const int np = nx + 2 * NB;
    const int nblk = (np + RC - 1) / RC;
    float xlo[RC], xhi[RC];
    // ---- pass A: forward sum, keep F before every block
PST_RC_UNROLL
    for (int j = 0; j < RC; j++) xlo[j] = 0.f;
    io.prefetch(0);
    io.take(xhi);
    float F = 0.f;
    for (int b = 0; b < nblk; b++) {
        io.prefetch(b + 1);
        ck[(long)b * cks] = F;
        fwd_block<NB, RC, false>(xlo, xhi, wm, w2, F, nullptr);
PST_RC_UNROLL
        for (int j = 0; j < RC; j++) xlo[j] = xhi[j];
        io.take(xhi);
    }
    // ---- pass B: blocks downwards.  xlo holds x block nblk-1 now.
PST_RC_UNROLL
    for (int j = 0; j < RC; j++) xhi[j] = xlo[j];
    io.prefetch(nblk - 2);
    io.take(xlo);
    float Bs = 0.f;
    float top[NB], mid[NB];
PST_RC_UNROLL
    for (int j = 0; j < NB; j++) { top[j] = 0.f; mid[j] = 0.f; }
    for (int b = nblk - 1; b >= 0; b--) {
        // x block b-2 is requested before this block's outputs are stored and x block b-1 is already here:
        // an in-place line (y over x) never overwrites a sample that is still to be read (NB <= RC)
        io.prefetch(b - 2);
        float Fb[RC];
        float Fs = ck[(long)b * cks];
        fwd_block<NB, RC, true>(xlo, xhi, wm, w2, Fs, Fb);
        const int k0 = b * RC;
        if (k0 + RC <= np) {
PST_RC_UNROLL
            for (int j = RC - 1; j >= 0; j--) { Bs = Bs + Fb[j]; Fb[j] = Bs; }
        } else {
PST_RC_UNROLL
            for (int j = RC - 1; j >= 0; j--) { if (k0 + j < np) Bs = Bs + Fb[j]; Fb[j] = Bs; }
        }
        if (k0 >= 2 * NB && k0 + RC <= nx) {
            io.store_block(Fb, k0 - NB, 0, RC);                 // y_i = B_{i+nb}, no reflection in this block
        } else {
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
                        if (i < NB) mid[i] = Fb[j];
                    }
                }
            }
            const int jlo = 2 * NB - k0 > 0 ? 2 * NB - k0 : 0;
            const int jhi = nx + NB - k0 < RC ? nx + NB - k0 : RC;
            if (jhi > jlo) io.store_block(Fb, k0 - NB, jlo, jhi);
            if (k0 == 0) {
PST_RC_UNROLL
                for (int j = NB - 1; j >= 0; j--) io.store_one(NB - 1 - j, mid[NB - 1 - j] + Fb[j]);
            }
        }
PST_RC_UNROLL
        for (int j = 0; j < RC; j++) xhi[j] = xlo[j];
        io.take(xlo);
    }
}

// one line in memory with a constant stride between samples: lane = line on the strided axes
template <int RC>
struct StridedIO {
    const float *s; float *d; long st; int nx;
    float pre[RC];
    PST_RC_HD void prefetch(int m)
    {
        if (m >= 0 && (m + 1) * RC <= nx) {                     // whole block inside the line: a running pointer
            const float *p = s + (long)m * RC * st;
PST_RC_UNROLL
            for (int j = 0; j < RC; j++) { pre[j] = *p; p += st; }
        } else {
PST_RC_UNROLL
            for (int j = 0; j < RC; j++) {
                const int i = m * RC + j;
                pre[j] = (m >= 0 && i < nx) ? s[(long)i * st] : 0.f;
            }
        }
    }
    PST_RC_HD void take(float *x)
    {
PST_RC_UNROLL
        for (int j = 0; j < RC; j++) x[j] = pre[j];
    }
    PST_RC_HD void store_block(const float *v, int i0, int jlo, int jhi)
    {
        if (jlo == 0 && jhi == RC) {
            float *p = d + (long)i0 * st;
PST_RC_UNROLL
            for (int j = 0; j < RC; j++) { *p = v[j]; p += st; }
        } else {
PST_RC_UNROLL
            for (int j = 0; j < RC; j++)
                if (j >= jlo && j < jhi) d[(long)(i0 + j) * st] = v[j];
        }
    }
    PST_RC_HD void store_one(int i, float v) { d[(long)i * st] = v; }
};

}  // namespace tri_rc
