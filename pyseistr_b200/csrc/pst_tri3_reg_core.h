// Per-line core of the register kernels of the distributed axis-3 smoothing pass (pst_dip.cu: tri3_reg_fwd_kernel,
// tri3_reg_bwd_kernel).  Same arithmetic as ps_smooth2 along axis 3 (reference dip_cfuns.c:458-484,508-529,564-580,
// 616-625), bit for bit, with the two running sums crossing the ranks' n3-slabs in the reference's order:
//   t_k = ((-wt x_k) + 2wt x_{k-nb}) + (-wt x_{k-2nb})          k in [0, n3g + 2nb)
//   F_k = F_{k-1} + t_k        forward running sum  (float, sequential; crosses the slabs upwards as one carry per line)
//   B_k = B_{k+1} + F_k        backward running sum (float, sequential; crosses the slabs downwards)
//   y_i = (B_{i+nb} + B_{nb+n3g+(n3g-1-i)}[i >= n3g-nb]) + B_{nb-1-i}[i < nb]
// One thread owns one line and walks the line's local rows in CHUNKS of NZ planes held in registers; NZ and the role of a
// chunk are template parameters, so the role of every row (own plane, neighbouring plane, outside the cube) and every fold
// index is known at compile time.  The forward walk forms the stencil values t of its first chunk, receives its line's
// carry, adds the t's to it in order, chunk after chunk, and returns the running sum for the next rank; it stores nothing
// but the running sum before every chunk.  The backward walk goes down the chunks: it loads the same rows again, repeats
// the forward additions from the saved sum -- same operands, same order, same bits -- (top chunk: before it asks for the
// backward carry), then runs the backward sum with the fold.
// EDGE: 0 interior chunk (steps k = zc + nb + q, q < NZ), 1 first chunk of the cube (k = q < NZ + nb; left reflection),
// 2 last chunk of the cube (k = zc + nb + q, q < NZ + nb; right reflection).  Slot q holds plane k - 2nb of x, then t_k,
// then F_k.
// IO is the caller's transport: wait_halos() (block-wide on the GPU: every thread of the block calls it once, the neighbours'
// planes are valid afterwards), recv(l) (the incoming carry of line l).
// The header compiles for the host too (tests/test_tri3_reg_core.py runs the ranks one after the other and checks the
// result against the CPU restatement).
#pragma once

#ifdef __CUDACC__
#define PST_T3R_HD __device__ __forceinline__
#define PST_T3R_LD(p) __ldcg(p)
#else
#define PST_T3R_HD inline
#define PST_T3R_LD(p) (*(p))
#endif
#ifdef __CUDA_ARCH__
#define PST_T3R_UNROLL _Pragma("unroll")
#else
#define PST_T3R_UNROLL
#endif

namespace tri3_reg {

// ARGS: one rank's part of a pass (on the GPU the kernel's own parameter block, read from constant memory -- copying it into a
// struct costs the registers the line needs).  Fields used:
//   const float *x        the slab, [nz][L]
//   const float *hb       planes [z0 - nb, z0): the previous rank's last planes (unused on the first rank)
//   const float *ha       planes [z1, z1 + nb): the next rank's first planes
//   float *ha_keep        forward walk: where those planes are kept for the backward walk (null: ha stays valid)
//   float *csave          [nz / NZ][L]: the forward running sum before every chunk
//   float *dst            output slab, [nz][L] (may be x)
//   long L; int nz, n3g, K0, K1; float wt, w2      lines per plane; slab height; global height; first / one-past-last
//                                                 step of the rank (K0 = 0 on the first rank, K1 = n3g + 2 nb on the last)

template <int NB, int NZ, int EDGE>
struct Reg {
    static constexpr int N = (EDGE == 0) ? NZ : NZ + NB;       // steps
    static constexpr int S = N + 2 * NB;                       // x rows
    static constexpr int OWN0 = (EDGE == 1) ? 2 * NB : NB;     // slots [OWN0, OWN0 + NZ): planes of the chunk
    // own: the chunk's first plane; before: plane zc - nb (previous chunk or previous rank); after: plane zc + NZ (next chunk
    // or next rank / its kept copy; null: filled by the caller).  PART 0: the chunk's planes, 1: the planes around it, 2: both
    template <int PART>
    static PST_T3R_HD void load(const float *own, const float *before, const float *after, long L, float (&v)[S])
    {
PST_T3R_UNROLL
        for (int q = 0; q < S; q++) {
            if (q >= OWN0 && q < OWN0 + NZ) {
                if (PART != 1) v[q] = PST_T3R_LD(own + (long)(q - OWN0) * L);
            } else if (q < OWN0) {
                if (EDGE == 1) { if (PART != 1) v[q] = 0.f; }                             // planes < 0
                else if (PART != 0) v[q] = PST_T3R_LD(before + (long)q * L);
            } else {
                if (EDGE == 2) { if (PART != 1) v[q] = 0.f; }                             // planes >= n3g
                else if (PART != 0 && after) v[q] = PST_T3R_LD(after + (long)(q - OWN0 - NZ) * L);
            }
        }
    }
    static PST_T3R_HD void stencil(float wm, float w2, float (&v)[S])
    {
PST_T3R_UNROLL
        for (int q = 0; q < N; q++) {
            float t = wm * v[q + 2 * NB];
            t = t + w2 * v[q + NB];
            t = t + wm * v[q];
            v[q] = t;
        }
    }
};

// one chunk of the forward walk; returns the running sum after the chunk
template <int NB, int NZ, int EDGE, class ARGS, class IO>
PST_T3R_HD float fwd_chunk(const ARGS &A, IO &io, long l, bool live, int j, int nch, bool head, float s)
{
    using R = Reg<NB, NZ, EDGE>;
    float v[R::S];
    const float *own = A.x + (long)j * NZ * A.L + l;
    const float *before = j > 0 ? own - (long)NB * A.L : A.hb + l;
    const float *after = j < nch - 1 ? own + (long)NZ * A.L : A.ha + l;
    if (head) {
        // first chunk: its own planes are in flight while the neighbours' flags are awaited; then its carry
        if (live) R::template load<0>(own, before, after, A.L, v);
        io.wait_halos();
        if (!live) return 0.f;
        R::template load<1>(own, before, after, A.L, v);
    } else {
        R::template load<2>(own, before, after, A.L, v);
    }
    if (EDGE != 2 && j == nch - 1 && A.ha_keep) {
        // keep the planes read from the next rank for the backward walk: that rank overwrites them in its own backward
        // walk, which runs before mine
PST_T3R_UNROLL
        for (int a = 0; a < NB; a++) A.ha_keep[(long)a * A.L + l] = v[R::OWN0 + NZ + a];
    }
    R::stencil(-A.wt, A.w2, v);
    if (head && EDGE != 1) s = io.recv(l);
    A.csave[(long)j * A.L + l] = s;
PST_T3R_UNROLL
    for (int q = 0; q < R::N; q++) s += v[q];
    return s;
}

// the forward walk of line l; false: the thread is not live (it still took part in wait_halos).  *carry: the running sum
// at the top of the slab (for the next rank)
template <int NB, int NZ, class ARGS, class IO>
PST_T3R_HD bool fwd_line(const ARGS &A, IO &io, long l, bool live, float *carry)
{
    const int nch = A.nz / NZ;
    const bool first = A.K0 == 0, last = A.K1 == A.n3g + 2 * NB;
    float s = 0.f;
    // chunk 0 (block-uniform role)
    if (first) s = fwd_chunk<NB, NZ, 1>(A, io, l, live, 0, nch, true, s);
    else if (last && nch == 1) s = fwd_chunk<NB, NZ, 2>(A, io, l, live, 0, nch, true, s);
    else s = fwd_chunk<NB, NZ, 0>(A, io, l, live, 0, nch, true, s);
    if (!live) return false;
    for (int j = 1; j < nch; j++) {
        if (last && j == nch - 1) s = fwd_chunk<NB, NZ, 2>(A, io, l, true, j, nch, false, s);
        else s = fwd_chunk<NB, NZ, 0>(A, io, l, true, j, nch, false, s);
    }
    *carry = s;
    return true;
}

// one chunk of the backward walk.  keep[]: x of the first nb planes of the chunk above (this thread has already
// overwritten them with outputs); on return: those of this chunk.  Returns the backward running sum below the chunk.
template <int NB, int NZ, int EDGE, class ARGS, class IO>
PST_T3R_HD float bwd_chunk(const ARGS &A, IO &io, long l, int j, int nch, const float *ha_src, float (&keep)[NB], float s)
{
    using R = Reg<NB, NZ, EDGE>;
    constexpr int N = R::N;
    float v[R::S];
    const float *own = A.x + (long)j * NZ * A.L + l;
    const float *before = j > 0 ? own - (long)NB * A.L : A.hb + l;
    const bool top = j == nch - 1;
    // (no halo flags to wait for: the forward walk of this pass did, and the previous rank's planes stay untouched until
    // its backward thread of this line has received the carry this walk returns)
    R::template load<2>(own, before, top ? ha_src + l : nullptr, A.L, v);
    if (EDGE != 2 && !top) {
PST_T3R_UNROLL
        for (int a = 0; a < NB; a++) v[R::OWN0 + NZ + a] = keep[a];
    }
PST_T3R_UNROLL
    for (int a = 0; a < NB; a++) keep[a] = v[R::OWN0 + a];
    R::stencil(-A.wt, A.w2, v);
    float sf = A.csave[(long)j * A.L + l];
PST_T3R_UNROLL
    for (int q = 0; q < N; q++) { sf += v[q]; v[q] = sf; }
    // (the last rank receives a +0 too -- on the GPU the one its forward kernel left in its own mailbox: with no wait loop
    // at all ptxas gives this straight-line code 32 registers and spills the whole line)
    if (top) s = io.recv(l);
    float *dl = A.dst + (long)j * NZ * A.L + l;                // local row of sample gi = k - nb: q (EDGE 0, 2), q - nb (EDGE 1)
    float park[NB];                                            // EDGE 2: B of the right pad; EDGE 1: heads awaiting the left pad
PST_T3R_UNROLL
    for (int q = N - 1; q >= 0; q--) {
        s += v[q];
        if (EDGE == 2) {
            // right pad: k >= nb + n3g <=> q >= NZ: parked.  The last nb samples (q in [NZ - nb, NZ)) take
            // B_{nb + n3g + (n3g - 1 - gi)}, the value parked by step q' = 2 NZ - 1 - q
            if (q >= NZ) park[q >= NZ ? q - NZ : 0] = s;
            else {
                float y = s;
                if (q >= NZ - NB) y = y + park[q >= NZ - NB ? NZ - 1 - q : 0];
                dl[(long)q * A.L] = y;
            }
        } else if (EDGE == 1) {
            // k = q.  k >= 2nb: sample gi = k - nb; k in [nb, 2nb): heads (completed by the left pad); k < nb: left pad,
            // y_gi = head_gi + B_k with gi = nb - 1 - k
            if (q >= 2 * NB) dl[(long)(q - NB) * A.L] = s;
            else if (q >= NB) park[q >= NB && q < 2 * NB ? q - NB : 0] = s;
            else dl[(long)(NB - 1 - q) * A.L] = park[q < NB ? NB - 1 - q : 0] + s;
        } else {
            dl[(long)q * A.L] = s;
        }
    }
    return s;
}

// the backward walk of line l; returns the backward running sum at the bottom of the slab (for the previous rank)
template <int NB, int NZ, class ARGS, class IO>
PST_T3R_HD float bwd_line(const ARGS &A, IO &io, long l)
{
    const int nch = A.nz / NZ;
    const bool first = A.K0 == 0, last = A.K1 == A.n3g + 2 * NB;
    const float *ha_src = A.ha_keep ? A.ha_keep : A.ha;
    float keep[NB];
PST_T3R_UNROLL
    for (int a = 0; a < NB; a++) keep[a] = 0.f;
    float s = 0.f;
    for (int j = nch - 1; j >= 0; j--) {
        if (first && j == 0) s = bwd_chunk<NB, NZ, 1>(A, io, l, j, nch, ha_src, keep, s);
        else if (last && j == nch - 1) s = bwd_chunk<NB, NZ, 2>(A, io, l, j, nch, ha_src, keep, s);
        else s = bwd_chunk<NB, NZ, 0>(A, io, l, j, nch, ha_src, keep, s);
    }
    return s;
}

}  // namespace tri3_reg
