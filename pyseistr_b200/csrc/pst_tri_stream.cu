// G2, streaming path: triangle smoothing of every line of one axis (ps_smooth2,
// reference dip_cfuns.c:458-484,508-529,564-580,616-625) as ONE persistent, warp-specialised
// pipeline per SM.  Arithmetic is the reference's, bit for bit:
//   t_k = ((-wt x_k) + 2wt x_{k-nb}) + (-wt x_{k-2nb})          k in [0, L),  L = nx + 2nb
//   F_k = F_{k-1} + t_k        forward running sum  (float, sequential)
//   B_k = B_{k+1} + F_k        backward running sum (float, sequential)
//   y_i = (B_{i+nb} + B_{nb+nx+(nx-1-i)}[i >= nx-nb]) + B_{nb-1-i}[i < nb]
//
// The two running sums are serial per line; parallelism exists only across lines, and the whole
// F line must exist before its backward sum starts.  Design (per CTA = per SM):
//   * a tile is 32 whole lines.  ONE "chain" warp owns them, lane = line, and keeps F of the
//     tile in shared memory (k-minor-4 layout: one 128-bit access moves 4 consecutive k of a
//     lane's line, conflict-free).
//   * PING-PONG: the backward sum of tile p-1 reads its F values in DESCENDING k, the forward
//     sum of tile p produces F in ASCENDING k.  Tile p stores F_k at the slot tile p-1 has just
//     consumed (slot order alternates between tiles), so both chains run FUSED in one loop: per
//     step one F slot is read (backward chain of the previous tile) and rewritten (forward chain
//     of the next tile).  Two independent FADD chains per lane (ILP 2), no phases, and the F
//     buffer is exactly one tile: 128 B per sample of line length.
//   * everything that is parallel is taken off the chain warp: a loader thread streams x tiles
//     with TMA (cp.async.bulk.tensor, mbarrier complete_tx; the box overlaps the previous one
//     by 2nb rows so that no look-back state exists; out-of-range rows/columns are zero-filled
//     by the TMA unit, which is exactly "tap skipped"), four builder warps turn x into t_k and
//     write it in the chain's k-minor-4 layout, and four storer warps move the B values of the
//     backward chain to global memory (transposing for the contiguous axis, so every global
//     access is a coalesced 128-byte row).  A padded line's dummy steps LEAD the forward pass
//     (k = j - D): their t is +0 by zero fill, so the slots they leave behind add nothing to the
//     backward chain that reads them last.
//   * rings (x, t, B) are guarded by mbarrier full/empty pairs; a ring's slot count is a multiple
//     of the number of warps that consume it (see the PROTOCOL RULE at the constants); the chain
//     warp is warp 0 and warps 4 and 8 stay idle so that it does not share its scheduler's issue
//     slots with a busy warp; it probes the barriers of stage q+1 while stage q's FADD chains run.
// HBM traffic is the compulsory 8 B/voxel (ncu: profiles/r01b_ncu_summary.md).  Reflections (fold2)
// are applied by the chain warp, in place in the B ring slot, in the 2 of ~33 stages that touch a
// line end.  Optional epilogue in the storers: dst = eps*p + S(src) with the sum of dst^2 (the gp
// step of ps_conjgrad), operand prefetched into L2 by the otherwise idle lanes of the loader warp.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pst_tri_stream.cuh"

namespace {

constexpr int KB = 32;                 // samples (k) per pipeline stage
constexpr int GP = 132;                // floats per 4-sample group in the t / B rings (132 = 4 mod 32)
constexpr int TSTAGE = 8 * GP;         // floats per ring stage
constexpr int NTS = 4;                 // t ring slots
#ifndef TS_NOS
#define TS_NOS 8
#endif
#ifndef TS_NXS
#define TS_NXS 8
#endif
#ifndef TS_EARLY
#define TS_EARLY 1
#endif
// PROTOCOL RULE: a ring's slot count must be a multiple of the number of warps that consume it
// round-robin (each slot is then always waited on by the same warp).  mbarrier waits identify a
// phase by its PARITY only, so a waiter may never run two phases ahead of a barrier; with, say, 6
// slots and 4 consumer warps a warp can reach use n+1 of a slot before another warp's use n has
// completed, and its wait falls through (observed: corrupted tiles and launch failures).
constexpr int NOS = TS_NOS;            // B ring slots
constexpr int NBW = 4;                 // builder warps
constexpr int NSW = 4;                 // storer warps
constexpr int NWARPS = 12;
constexpr int NBMAX = 16;
constexpr int NSMAX = 48;               // stages per line the F hand-off barriers are laid out for (lines up to 48*32 steps)
#ifndef TS_SPLIT
#define TS_SPLIT 1                      // 1: forward and backward chains on two warps (0 and 4) of the same scheduler
#endif
constexpr int NXS_MAX = TS_NXS;
static_assert(NTS % NBW == 0 && NXS_MAX % NBW == 0 && NOS % NSW == 0, "ring slots must be a multiple of their consumer warps");

// row pitch of a contiguous-axis x stage: 32 + 2nb samples rounded up to a 16-byte multiple and to
// an ODD number of 16-byte units so that "lane = line" 128-bit reads are conflict-free.  The box
// starts at sample 32q - D - 2nb, a multiple of 4 because nx % 4 == 0 and (nx + 2nb + D) % 32 == 0
// (TMA needs a 16-byte aligned start in the contiguous dimension).
__host__ __device__ constexpr int xw_pad(int) { return 0; }
__host__ __device__ constexpr int xw_of(int nb)
{
    int w = (KB + 2 * nb + 3) / 4 * 4;
    return ((w / 4) % 2 == 0) ? w + 4 : w;
}

struct Args {
    float *dst;
    long d, sb;            // strided: element stride along the line, batch stride (floats)
    int na;                // strided: extent of the fast (lane) index; contiguous: number of lines
    int nx, nb;
    int tilesA;            // strided: tiles along na
    long ntiles;
    int NS, L, Lp;
    int nxs, xsf;          // x ring slots, floats per x stage
    int XW, pad;           // contiguous: row pitch of an x stage, left pad of the box
    float wm, w2;
    unsigned *err;
    // optional fused CG epilogue of the storers (ps_conjgrad dip_cfuns.c:303,318,320):
    //   gp = eps*p + S(gx) written to dst, partial sum of gp^2 (double) per CTA
    const float *epi_p;
    float epi_eps;
    double *partial;
    int partial_stride;
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_probe(uint64_t *b, unsigned parity)      // never suspends (test_wait, not try_wait)
{
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// spin with a wall-clock bound: a protocol bug must surface as an error, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity, unsigned *err)
{
    const unsigned a = smem_u32(b);
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    const long long t0 = clock64();
    for (;;) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000LL) { if (err) *err = 1u; __trap(); }
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *m, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *m, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ float tri_t3(float xa, float xb, float xc, float wm, float w2)
{
    float v = wm * xa;          // 0 + wm*x_k: the sign of a zero cannot reach a non-zero sum (F starts at +0)
    v = v + w2 * xb;
    v = v + wm * xc;
    return v;
}

// ring cursor: slot index and phase parity advance without divisions
struct Cursor {
    int slot; unsigned par; int n;
    __device__ __forceinline__ void init(long seq, int nslots) { n = nslots; slot = (int)(seq % nslots); par = (unsigned)((seq / nslots) & 1); }
    __device__ __forceinline__ void advance(int by) { slot += by; while (slot >= n) { slot -= n; par ^= 1u; } }
};

// One fast stage of the chain warp (8 groups of 4 steps, no line end in reach): straight-line code.
//   backward chain of the previous tile:  B += F_old[slot]   -> B ring
//   forward chain of the next tile:       Fs += t            -> F_new[slot]
// REV selects the slot order of this pass (ping-pong).
template <bool REV, bool FWD, bool BWD>
__device__ __forceinline__ void chain_stage_load(const float *Fl, const float *T, int G0, int NG, float4 (&tv)[8], float4 (&fv)[8])
{
#pragma unroll
    for (int g = 0; g < 8; g++) {
        const int PG = REV ? NG - 1 - (G0 + g) : G0 + g;
        if (BWD) fv[g] = *reinterpret_cast<const float4 *>(Fl + (size_t)PG * 128);
        if (FWD) tv[g] = *reinterpret_cast<const float4 *>(T + g * GP);
    }
}
template <bool REV, bool FWD, bool BWD>
__device__ __forceinline__ void chain_stage_compute(float *Fl, float *O, int G0, int NG, const float4 (&tv)[8], const float4 (&fv)[8],
                                                    float &Fs, float &B)
{
#pragma unroll
    for (int g = 0; g < 8; g++) {
        const int PG = REV ? NG - 1 - (G0 + g) : G0 + g;
        float4 o, fn;
        if (BWD) {
            if (!REV) { B += fv[g].x; o.x = B; B += fv[g].y; o.y = B; B += fv[g].z; o.z = B; B += fv[g].w; o.w = B; }
            else      { B += fv[g].w; o.x = B; B += fv[g].z; o.y = B; B += fv[g].y; o.z = B; B += fv[g].x; o.w = B; }
            *reinterpret_cast<float4 *>(O + g * GP) = o;
        }
        if (FWD) {
            if (!REV) { Fs += tv[g].x; fn.x = Fs; Fs += tv[g].y; fn.y = Fs; Fs += tv[g].z; fn.z = Fs; Fs += tv[g].w; fn.w = Fs; }
            else      { Fs += tv[g].x; fn.w = Fs; Fs += tv[g].y; fn.z = Fs; Fs += tv[g].z; fn.y = Fs; Fs += tv[g].w; fn.x = Fs; }
            *reinterpret_cast<float4 *>(Fl + (size_t)PG * 128) = fn;
        }
    }
}

// Reflections of fold2 for a stage that holds line-end samples, applied IN the B ring slot after the
// straight-line stage code (O = this lane's column of the slot).  Backward position of ring step jj
// of the stage is kb = kb_hi - jj.  Right-tail sums are parked in ER until their mirror sample comes
// by, heads (i < nb) are parked in EH until the left tail completes them; the storers map ring
// steps kb < nb to sample nb-1-kb and skip kb in [nb, 2nb) and the right tail.
__device__ __noinline__ void chain_fold_fix(float *__restrict__ O, float *__restrict__ ER, float *__restrict__ EH, int lane,
                                            int kb_hi, int nx, int nb, int L)
{
    const int kb_lo = kb_hi - (KB - 1);
    auto at = [&](int kb) -> float * { const int jj = kb_hi - kb; return O + (jj >> 2) * GP + (jj & 3); };
    if (kb_hi == L - 1) {
        // first stage of the pass: it holds the whole right end (2nb <= 32 steps), so every mirror pair
        // B_{nx+nb-1-r} + B_{nx+nb+r} is formed directly; all loads are independent
#pragma unroll 4
        for (int r = 0; r < nb; r++) {
            float *q = O + ((nb + r) >> 2) * GP + ((nb + r) & 3);
            const float *m = O + ((nb - 1 - r) >> 2) * GP + ((nb - 1 - r) & 3);
            *q = *q + *m;
        }
    } else {
        for (int kb = min(L - 1, kb_hi); kb >= max(nx + nb, kb_lo); kb--) ER[(kb - nx - nb) * 32 + lane] = *at(kb);
        for (int kb = min(nx + nb - 1, kb_hi); kb >= max(nx, kb_lo); kb--) { float *q = at(kb); *q = *q + ER[(nx - 1 - (kb - nb)) * 32 + lane]; }
    }
#pragma unroll 4
    for (int kb = min(2 * nb - 1, kb_hi); kb >= max(nb, kb_lo); kb--) EH[(kb - nb) * 32 + lane] = *at(kb);
#pragma unroll 4
    for (int kb = min(nb - 1, kb_hi); kb >= max(0, kb_lo); kb--) { float *q = at(kb); *q = EH[(nb - 1 - kb) * 32 + lane] + *q; }
}

// One pass of the chain warp over the NS stages of a line: forward chain of tile p (FWD) fused with
// the backward chain of tile p-1 (BWD).  The barrier state of stage q+1 is probed while the FADD
// chains of stage q run, so a ready stage costs no wait latency.
template <bool REV, bool FWD, bool BWD>
__device__ __forceinline__ void chain_pass(float *Fl, const float *Tr, float *Or, float *ER, float *EH, int lane,
                                           uint64_t *full_t, uint64_t *empty_t, uint64_t *full_o, uint64_t *empty_o,
                                           Cursor &ct, Cursor &co, int NS, int nx, int nb, int L, unsigned *err)
{
    const int NG = NS * (KB / 4);
    float Fs = 0.f, B = 0.f;
    int kb_hi = L - 1;
    bool rt = false, ro = false;
    for (int q = 0; q < NS; q++, kb_hi -= KB) {
        if (FWD && !rt) mbar_wait(full_t + ct.slot, ct.par, err);
        if (BWD && !ro) mbar_wait(empty_o + co.slot, co.par ^ 1u, err);
        const float *T = Tr + ct.slot * TSTAGE + 4 * lane;
        float *O = Or + co.slot * TSTAGE + 4 * lane;
        uint64_t *const rel_t = empty_t + ct.slot, *const sig_o = full_o + co.slot;
        float4 tv[8], fv[8];
        chain_stage_load<REV, FWD, BWD>(Fl, T, q * (KB / 4), NG, tv, fv);
        if (FWD) { ct.slot = ct.slot + 1 == NTS ? 0 : ct.slot + 1; ct.par ^= (ct.slot == 0); }
        if (BWD) { co.slot = co.slot + 1 == NOS ? 0 : co.slot + 1; co.par ^= (co.slot == 0); }
        if (q + 1 < NS) {
            if (FWD) rt = mbar_probe(full_t + ct.slot, ct.par);
            if (BWD) ro = mbar_probe(empty_o + co.slot, co.par ^ 1u);
        } else { rt = false; ro = false; }
        chain_stage_compute<REV, FWD, BWD>(Fl, O, q * (KB / 4), NG, tv, fv, Fs, B);
        if (BWD && (kb_hi >= nx || kb_hi - (KB - 1) < 2 * nb)) chain_fold_fix(O, ER, EH, lane, kb_hi, nx, nb, L);
        __syncwarp();
        if (lane == 0) {
            if (FWD) mbar_arrive(rel_t);
            if (BWD) mbar_arrive(sig_o);
        }
    }
}

// ---- the two chains on two warps (TS_SPLIT): warp 0 runs the backward chain of tile p-1, warp 4 the forward
// chain of tile p, one stage behind it.  They share the F buffer slot by slot: f_read[q] (backward warp -> forward
// warp: "stage q of this pass is in my registers, overwrite it") and f_written[q] (forward warp -> backward warp:
// "stage q of this pass is stored", consumed one pass later).  Both are indexed by the PHYSICAL stage of the F buffer
// (the slot order alternates from pass to pass).  Each barrier completes once per pass, so its parity is the pass
// parity and neither warp can run two phases ahead of the other.
template <bool REV>
__device__ __forceinline__ void bwd_pass(float *Fl, float *Or, float *ER, float *EH, int lane, uint64_t *full_o,
                                         uint64_t *empty_o, uint64_t *f_read, uint64_t *f_written, Cursor &co,
                                         unsigned wpar, int NS, int nx, int nb, int L, unsigned *err)
{
    const int NG = NS * (KB / 4);
    float B = 0.f;
    int kb_hi = L - 1;
    bool ro = false;
    for (int q = 0; q < NS; q++, kb_hi -= KB) {
        if (!ro) mbar_wait(empty_o + co.slot, co.par ^ 1u, err);
        const int ps = REV ? NS - 1 - q : q;                       // PHYSICAL stage of the F buffer (slot order alternates per pass)
        mbar_wait(f_written + ps, wpar, err);                      // stored by the forward warp in the previous pass
        float *O = Or + co.slot * TSTAGE + 4 * lane;
        uint64_t *const sig_o = full_o + co.slot;
        float4 fv[8];
#pragma unroll
        for (int g = 0; g < 8; g++) {
            const int G = q * (KB / 4) + g;
            fv[g] = *reinterpret_cast<const float4 *>(Fl + (size_t)(REV ? NG - 1 - G : G) * 128);
        }
        co.slot = co.slot + 1 == NOS ? 0 : co.slot + 1; co.par ^= (co.slot == 0);
        ro = (q + 1 < NS) ? mbar_probe(empty_o + co.slot, co.par ^ 1u) : false;
#pragma unroll
        for (int g = 0; g < 8; g++) {
            float4 o;
            if (!REV) { B += fv[g].x; o.x = B; B += fv[g].y; o.y = B; B += fv[g].z; o.z = B; B += fv[g].w; o.w = B; }
            else      { B += fv[g].w; o.x = B; B += fv[g].z; o.y = B; B += fv[g].y; o.z = B; B += fv[g].x; o.w = B; }
            *reinterpret_cast<float4 *>(O + g * GP) = o;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(f_read + ps);                   // all lanes' F values are in registers (consumed above)
        if (kb_hi >= nx || kb_hi - (KB - 1) < 2 * nb) chain_fold_fix(O, ER, EH, lane, kb_hi, nx, nb, L);
        __syncwarp();
        if (lane == 0) mbar_arrive(sig_o);
    }
}
template <bool REV>
__device__ __forceinline__ void fwd_pass(float *Fl, const float *Tr, int lane, uint64_t *full_t, uint64_t *empty_t,
                                         uint64_t *f_read, uint64_t *f_written, Cursor &ct, bool wait_read,
                                         unsigned rpar, int NS, unsigned *err)
{
    const int NG = NS * (KB / 4);
    float Fs = 0.f;
    bool rt = false;
    for (int q = 0; q < NS; q++) {
        if (!rt) mbar_wait(full_t + ct.slot, ct.par, err);
        const float *T = Tr + ct.slot * TSTAGE + 4 * lane;
        uint64_t *const rel_t = empty_t + ct.slot;
        float4 tv[8];
#pragma unroll
        for (int g = 0; g < 8; g++) tv[g] = *reinterpret_cast<const float4 *>(T + g * GP);
        ct.slot = ct.slot + 1 == NTS ? 0 : ct.slot + 1; ct.par ^= (ct.slot == 0);
        rt = (q + 1 < NS) ? mbar_probe(full_t + ct.slot, ct.par) : false;
        float4 fn[8];
#pragma unroll
        for (int g = 0; g < 8; g++) {
            if (!REV) { Fs += tv[g].x; fn[g].x = Fs; Fs += tv[g].y; fn[g].y = Fs; Fs += tv[g].z; fn[g].z = Fs; Fs += tv[g].w; fn[g].w = Fs; }
            else      { Fs += tv[g].x; fn[g].w = Fs; Fs += tv[g].y; fn[g].z = Fs; Fs += tv[g].z; fn[g].y = Fs; Fs += tv[g].w; fn[g].x = Fs; }
        }
        const int ps = REV ? NS - 1 - q : q;
        if (wait_read) mbar_wait(f_read + ps, rpar, err);           // the backward warp has taken the old values of this stage
#pragma unroll
        for (int g = 0; g < 8; g++) {
            const int G = q * (KB / 4) + g;
            *reinterpret_cast<float4 *>(Fl + (size_t)(REV ? NG - 1 - G : G) * 128) = fn[g];
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(rel_t); mbar_arrive(f_written + ps); }
    }
}

// x stage -> t stage, lane = line, radius known at compile time: every x sample is read from shared
// memory ONCE into a register window (the generic path reads it three times)
template <bool CONTIG, int NB>
__device__ __forceinline__ void build_stage_window(const float *X, float *T, int lane, float wm, float w2)
{
    constexpr int PAD = CONTIG ? xw_pad(NB) : 0;
    constexpr int W = CONTIG ? xw_of(NB) : KB + 2 * NB;
    float xw[W];
    if (CONTIG) {
        const float4 *row = reinterpret_cast<const float4 *>(X + lane * W);
#pragma unroll
        for (int r = 0; r < W / 4; r++) { const float4 v = row[r]; xw[4 * r] = v.x; xw[4 * r + 1] = v.y; xw[4 * r + 2] = v.z; xw[4 * r + 3] = v.w; }
    } else {
#pragma unroll
        for (int r = 0; r < W; r++) xw[r] = X[r * 32 + lane];
    }
#pragma unroll
    for (int g = 0; g < 8; g++) {
        float4 t;
        t.x = tri_t3(xw[PAD + 4 * g + 0 + 2 * NB], xw[PAD + 4 * g + 0 + NB], xw[PAD + 4 * g + 0], wm, w2);
        t.y = tri_t3(xw[PAD + 4 * g + 1 + 2 * NB], xw[PAD + 4 * g + 1 + NB], xw[PAD + 4 * g + 1], wm, w2);
        t.z = tri_t3(xw[PAD + 4 * g + 2 + 2 * NB], xw[PAD + 4 * g + 2 + NB], xw[PAD + 4 * g + 2], wm, w2);
        t.w = tri_t3(xw[PAD + 4 * g + 3 + 2 * NB], xw[PAD + 4 * g + 3 + NB], xw[PAD + 4 * g + 3], wm, w2);
        *reinterpret_cast<float4 *>(T + g * GP + 4 * lane) = t;
    }
}

// NB > 0: radius fixed at compile time (register-window builders); NB == 0: any radius <= NBMAX
template <bool CONTIG, int NB>
__global__ void __launch_bounds__(NWARPS * 32, 1)
tri_stream_kernel(const __grid_constant__ CUtensorMap tmap, const Args A)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *const Fbuf = reinterpret_cast<float *>(smem_raw);
    float *const Xr = Fbuf + (size_t)A.Lp * 32;
    float *const Tr = Xr + (size_t)A.nxs * A.xsf;
    float *const Or = Tr + NTS * TSTAGE;
    float *const ER = Or + NOS * TSTAGE;
    float *const EH = ER + A.nb * 32;
    uint64_t *const bars = reinterpret_cast<uint64_t *>(EH + A.nb * 32);
    uint64_t *const full_x = bars, *const empty_x = full_x + NXS_MAX;
    uint64_t *const full_t = empty_x + NXS_MAX, *const empty_t = full_t + NTS;
    uint64_t *const full_o = empty_t + NTS, *const empty_o = full_o + NOS;
    uint64_t *const f_read = empty_o + NOS, *const f_written = f_read + NSMAX;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < A.nxs; i++) { mbar_init(full_x + i, 1); mbar_init(empty_x + i, 1); }
        for (int i = 0; i < NTS; i++) { mbar_init(full_t + i, 1); mbar_init(empty_t + i, 1); }
        for (int i = 0; i < NOS; i++) { mbar_init(full_o + i, 1); mbar_init(empty_o + i, 1); }
        for (int i = 0; i < NSMAX; i++) { mbar_init(f_read + i, 1); mbar_init(f_written + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const long m = (A.ntiles - (long)blockIdx.x + (long)gridDim.x - 1) / (long)gridDim.x;   // tiles of this CTA
    if (m <= 0) return;
    const int NS = A.NS, nb = NB > 0 ? NB : A.nb, nx = A.nx, L = A.L, Lp = A.Lp;
    const long nstage = m * NS;
    double acc = 0.0;                                        // storers: fused sum of gp^2

    if (TS_SPLIT && warp == 0) {
        // ================================ backward chain warp ================================
        float *const Fl = Fbuf + 4 * lane;
        Cursor co;
        co.init(0, NOS);
        for (long p = 1; p <= m; p++) {
            // reads what the forward warp stored in pass p-1: phase p-1 of f_written
            if (p & 1) bwd_pass<true>(Fl, Or, ER, EH, lane, full_o, empty_o, f_read, f_written, co, (unsigned)((p - 1) & 1), NS, nx, nb, L, A.err);
            else bwd_pass<false>(Fl, Or, ER, EH, lane, full_o, empty_o, f_read, f_written, co, (unsigned)((p - 1) & 1), NS, nx, nb, L, A.err);
        }
    } else if (TS_SPLIT && warp == 4) {
        // ================================ forward chain warp ================================
        float *const Fl = Fbuf + 4 * lane;
        Cursor ct;
        ct.init(0, NTS);
        for (long p = 0; p < m; p++) {
            // overwrites what the backward warp reads in the same pass: phase p-1 of f_read (none in pass 0)
            if (p & 1) fwd_pass<true>(Fl, Tr, lane, full_t, empty_t, f_read, f_written, ct, p > 0, (unsigned)((p - 1) & 1), NS, A.err);
            else fwd_pass<false>(Fl, Tr, lane, full_t, empty_t, f_read, f_written, ct, p > 0, (unsigned)((p - 1) & 1), NS, A.err);
        }
    } else if (!TS_SPLIT && warp == 0) {
        // ================================ chain warp ================================
        float *const Fl = Fbuf + 4 * lane;
        Cursor ct, co;
        ct.init(0, NTS);
        co.init(0, NOS);
#define TS_PASS(R, F, Bk) chain_pass<R, F, Bk>(Fl, Tr, Or, ER, EH, lane, full_t, empty_t, full_o, empty_o, ct, co, NS, nx, nb, L, A.err)
        TS_PASS(false, true, false);                          // tile 0: forward only
        for (long p = 1; p < m; p++) {
            if (p & 1) TS_PASS(true, true, true);
            else TS_PASS(false, true, true);
        }
        if (m & 1) TS_PASS(true, false, true);                // last tile: backward only
        else TS_PASS(false, false, true);
#undef TS_PASS
    } else if (warp == 1) {
        // ================================ loader ================================
        // lane 0 streams the x tiles (TMA).  With a fused epilogue the whole warp also pulls the epilogue
        // operand of the tile whose BACKWARD sum runs in the same pass into L2 (one 128-byte row or line
        // segment per lane): the loader runs a ring ahead of the chain, which hides the DRAM latency the
        // storers would otherwise wait for.
        Cursor cx;
        cx.init(0, A.nxs);
        const unsigned xbytes = (unsigned)A.xsf * 4u;
        for (long p = 0; p <= m; p++) {
            const long tile = (long)blockIdx.x + p * (long)gridDim.x;
            int c0 = 0, ib = 0;
            if (!CONTIG) { const long b = tile / A.tilesA; ib = (int)b; c0 = (int)(tile - b * A.tilesA) * 32; }
            // epilogue operand rows of tile p-1
            const float *pf = nullptr;
            if (A.epi_p && p > 0) {
                const long tprev = tile - (long)gridDim.x;
                if (CONTIG) {
                    const long line = tprev * 32 + lane;
                    if (line < A.na) pf = A.epi_p + line * nx;
                } else {
                    const long b = tprev / A.tilesA;
                    pf = A.epi_p + b * A.sb + (long)(tprev - b * A.tilesA) * 32;
                }
            }
            if (p == m && !A.epi_p) break;
            for (int q = 0; q < NS; q++) {
                if (p < m && lane == 0) {
                    mbar_wait(empty_x + cx.slot, cx.par ^ 1u, A.err);
                    mbar_arrive_expect_tx(full_x + cx.slot, xbytes);
                    float *dstx = Xr + (size_t)cx.slot * A.xsf;
                    if (CONTIG) tma_load_2d(dstx, &tmap, q * KB - (Lp - L) - 2 * nb - A.pad, (int)(tile * 32), full_x + cx.slot);
                    else tma_load_3d(dstx, &tmap, c0, q * KB - (Lp - L) - 2 * nb, ib, full_x + cx.slot);
                }
                if (p < m) cx.advance(1);
                __syncwarp();
                if (pf) {
                    const int kb0 = L - 1 - q * KB;                 // stage q of the backward pass covers samples (kb0-31 .. kb0) - nb
                    if (CONTIG) {
                        const int i0 = min(max(kb0 - (KB - 1) - nb, 0), nx - 1), i1 = min(max(kb0 - nb, 0), nx - 1);
                        prefetch_l2(pf + i0);
                        prefetch_l2(pf + i1);
                    } else {
                        const int i = kb0 - lane - nb;
                        if (i >= 0 && i < nx) prefetch_l2(pf + (long)i * A.d);
                    }
                }
            }
        }
    } else if (warp == 2 || warp == 3 || warp == 5 || warp == 6) {
        // ================================ builders: x -> t ================================
        const int bw = warp < 4 ? warp - 2 : warp - 3;       // 0..3
        Cursor cx, ct;
        cx.init(bw, A.nxs);
        ct.init(bw, NTS);
        const float wm = A.wm, w2 = A.w2;
        for (long seq = bw; seq < nstage; seq += NBW) {
            mbar_wait(full_x + cx.slot, cx.par, A.err);
            mbar_wait(empty_t + ct.slot, ct.par ^ 1u, A.err);
            const float *X = Xr + (size_t)cx.slot * A.xsf;
            float *T = Tr + ct.slot * TSTAGE;
            if (NB > 0) {
                build_stage_window<CONTIG, (NB > 0 ? NB : 1)>(X, T, lane, wm, w2);
            } else if (!CONTIG) {
                // x stage rows r <-> k = 32q - D - 2nb + r (D = Lp - L dummy steps lead the forward pass), 32 lanes per row
                const float *xc = X + lane, *xb = xc + nb * 32, *xa = xb + nb * 32;
#pragma unroll 2
                for (int g = 0; g < 8; g++) {
                    float a[4], b[4], c[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) { a[e] = xa[(4 * g + e) * 32]; b[e] = xb[(4 * g + e) * 32]; c[e] = xc[(4 * g + e) * 32]; }
                    float4 t;
                    t.x = tri_t3(a[0], b[0], c[0], wm, w2);
                    t.y = tri_t3(a[1], b[1], c[1], wm, w2);
                    t.z = tri_t3(a[2], b[2], c[2], wm, w2);
                    t.w = tri_t3(a[3], b[3], c[3], wm, w2);
                    *reinterpret_cast<float4 *>(T + g * GP + 4 * lane) = t;
                }
            } else {
                // x stage rows = lines, XW samples each: sample m <-> k = 32q - D - 2nb - pad + m; lane = step in stage
                const float *xc = X + A.pad + lane, *xb = xc + nb, *xa = xb + nb;
                float *tl = T + (lane >> 2) * GP + (lane & 3);
                const int XW = A.XW;
#pragma unroll 4
                for (int c = 0; c < 32; c++) {
                    const float t = tri_t3(xa[c * XW], xb[c * XW], xc[c * XW], wm, w2);
                    tl[4 * c] = t;
                }
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(full_t + ct.slot); mbar_arrive(empty_x + cx.slot); }
            cx.advance(NBW);
            ct.advance(NBW);
        }
    } else if (warp == 7 || warp == 9 || warp == 10 || warp == 11) {
        // ================================ storers: B ring -> global ================================
        const int sw = warp == 7 ? 0 : warp - 8;             // 0..3
        Cursor co;
        co.init(sw, NOS);
        for (long seq = sw; seq < nstage; seq += NSW) {
            const long p = seq / NS;
            const int q = (int)(seq - p * NS);
            const long tile = (long)blockIdx.x + p * (long)gridDim.x;
            if (CONTIG) {
                // lane = step within the stage: consecutive lanes store consecutive samples of a line
                const long line0 = tile * 32;
                const int nl = (int)min((long)32, (long)A.na - line0);
                const int kb = L - 1 - (q * KB + lane);
                int i;
                bool valid;
                if (kb >= nx + nb || kb < 0) { i = 0; valid = false; }
                else if (kb >= nb) { i = kb - nb; valid = i >= nb; }
                else { i = nb - 1 - kb; valid = true; }
                float pv[32];
                if (A.epi_p && valid) {                          // operands of the epilogue: in flight before the stage is ready
                    const float *pl = A.epi_p + line0 * nx + i;
#pragma unroll
                    for (int c = 0; c < 32; c++) pv[c] = (c < nl) ? pl[(long)c * nx] : 0.f;
                }
                mbar_wait(full_o + co.slot, co.par, A.err);
                const float *ol = Or + co.slot * TSTAGE + (lane >> 2) * GP + (lane & 3);
                float *dl = A.dst + line0 * nx + i;
                if (valid) {
                    if (A.epi_p) {
#pragma unroll
                        for (int c = 0; c < 32; c++) {
                            if (c >= nl) break;
                            float gq = A.epi_eps * pv[c];
                            gq += ol[4 * c];
                            dl[(long)c * nx] = gq;
                            acc += (double)gq * (double)gq;
                        }
                    } else {
#pragma unroll 8
                        for (int c = 0; c < nl; c++) dl[(long)c * nx] = ol[4 * c];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_o + co.slot);
            } else {
                // lane = line: a warp row is 128 contiguous bytes
                const long ib = tile / A.tilesA;
                const int c0 = (int)(tile - ib * A.tilesA) * 32;
                const bool live = (c0 + lane) < A.na;
                float *dtile = A.dst + ib * A.sb + c0 + lane;
                const int kb0 = L - 1 - q * KB;                        // backward position of step 0 of the stage
                const bool interior = kb0 < nx + nb && kb0 - (KB - 1) >= 2 * nb;   // every step is a plain sample
                float pv[32];
                if (A.epi_p && live && interior) {                     // epilogue operands: in flight before the stage is ready
                    const float *pp = A.epi_p + (dtile - A.dst) + (long)(kb0 - nb) * A.d;
#pragma unroll
                    for (int j = 0; j < 32; j++) pv[j] = pp[-(long)j * A.d];
                }
                mbar_wait(full_o + co.slot, co.par, A.err);
                const float *ol = Or + co.slot * TSTAGE + 4 * lane;
                float4 o[8];
#pragma unroll
                for (int g = 0; g < 8; g++) o[g] = *reinterpret_cast<const float4 *>(ol + g * GP);
                __syncwarp();
                if (TS_EARLY && lane == 0) mbar_arrive(empty_o + co.slot);        // slot is in registers: release it before storing
                if (interior) {
                    float *op = dtile + (long)(kb0 - nb) * A.d;
                    const long d = A.d;
                    if (live && A.epi_p) {
#pragma unroll
                        for (int g = 0; g < 8; g++) {
                            const float ov[4] = {o[g].x, o[g].y, o[g].z, o[g].w};
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                float gq = A.epi_eps * pv[4 * g + e];
                                gq += ov[e];
                                op[-(long)(4 * g + e) * d] = gq;
                                acc += (double)gq * (double)gq;
                            }
                        }
                    } else if (live) {
#pragma unroll
                        for (int g = 0; g < 8; g++) {
                            op[0] = o[g].x; op[-d] = o[g].y; op[-2 * d] = o[g].z; op[-3 * d] = o[g].w;
                            op -= 4 * d;
                        }
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 8; g++) {
                        const float oa[4] = {o[g].x, o[g].y, o[g].z, o[g].w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int kb = kb0 - 4 * g - e;
                            int i;
                            bool valid;
                            if (kb >= nx + nb || kb < 0) { i = 0; valid = false; }
                            else if (kb >= nb) { i = kb - nb; valid = i >= nb; }
                            else { i = nb - 1 - kb; valid = true; }
                            if (valid && live) {
                                float gq = oa[e];
                                if (A.epi_p) {
                                    gq = A.epi_eps * A.epi_p[(dtile - A.dst) + (long)i * A.d];
                                    gq += oa[e];
                                    acc += (double)gq * (double)gq;
                                }
                                dtile[(long)i * A.d] = gq;
                            }
                        }
                    }
                }
            }
            if (!CONTIG && !TS_EARLY) { __syncwarp(); if (lane == 0) mbar_arrive(empty_o + co.slot); }
            co.advance(NSW);
        }
    }
    if (A.partial) {
        // deterministic CTA sum: fixed lane tree, then warps in index order
        __shared__ double red[NWARPS];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < NWARPS; w++) t += red[w];
            A.partial[(size_t)blockIdx.x * A.partial_stride] = t;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

size_t smem_bytes(bool contig, int Lp, int nxs, int xsf, int nb)
{
    (void)contig;
    size_t fl = (size_t)Lp * 32 + (size_t)nxs * xsf + (size_t)NTS * TSTAGE + (size_t)NOS * TSTAGE + (size_t)2 * nb * 32;
    return fl * 4 + (size_t)(2 * NXS_MAX + 2 * NTS + 2 * NOS + 2 * NSMAX) * 8;
}

template <int NB>
int launch_nb(bool contig, unsigned grid, size_t smem, cudaStream_t stream, const CUtensorMap &tm, const Args &A)
{
    // the opt-in shared-memory limit is a per-device function attribute: remember it per device
    static bool attr[2][64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -4;
    if (contig) {
        if (!attr[0][dev]) { if (cudaFuncSetAttribute(tri_stream_kernel<true, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess) return -4; attr[0][dev] = true; }
        tri_stream_kernel<true, NB><<<grid, NWARPS * 32, smem, stream>>>(tm, A);
    } else {
        if (!attr[1][dev]) { if (cudaFuncSetAttribute(tri_stream_kernel<false, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess) return -4; attr[1][dev] = true; }
        tri_stream_kernel<false, NB><<<grid, NWARPS * 32, smem, stream>>>(tm, A);
    }
    return 0;
}

}  // namespace

// Plan: can the streaming kernel run this axis?  (alignment for TMA, radius bounds, F tile fits)
bool pst_tri_stream_ok(int axis, int n1, int n2, int n3, int nb, const void *src, const void *dst)
{
    const int nn[3] = {n1, n2, n3};
    const int nx = nn[axis];
    if (nb < 2 || nb > NBMAX || nb > nx) return false;
    if (n1 % 4 != 0) return false;
    if ((((uintptr_t)src) & 15) || (((uintptr_t)dst) & 15)) return false;
    if ((long)n1 * n2 >= (1L << 31) || (long)n2 * n3 >= (1L << 31)) return false;
    if (!get_encode()) return false;
    const int L = nx + 2 * nb, Lp = (L + KB - 1) / KB * KB;
    const int XW = xw_of(nb);
    const int xsf = axis == 0 ? 32 * XW : (KB + 2 * nb) * 32;
    if (Lp / KB > NSMAX) return false;
    return smem_bytes(axis == 0, Lp, NBW, xsf, nb) <= 226 * 1024;
}

int pst_tri_stream_launch(cudaStream_t stream, int sm_count, int axis, const float *src, float *dst,
                          int n1, int n2, int n3, int nb, unsigned *d_err, const pst_tri_stream_epi *epi, int *grid_out)
{
    const int nn[3] = {n1, n2, n3};
    const int nx = nn[axis];
    const bool contig = axis == 0;
    Args A{};
    A.dst = dst; A.nx = nx; A.nb = nb; A.err = d_err;
    if (epi) { A.epi_p = epi->p; A.epi_eps = epi->eps; A.partial = epi->partial; A.partial_stride = epi->partial_stride; }
    A.L = nx + 2 * nb;
    A.Lp = (A.L + KB - 1) / KB * KB;
    A.NS = A.Lp / KB;
    const float wt = (float)(1.0 / ((double)nb * nb));          // ps_triangle_init dip_cfuns.c:421
    A.wm = -wt;
    A.w2 = (float)(2. * wt);
    A.pad = xw_pad(nb);
    A.XW = xw_of(nb);
    A.xsf = contig ? 32 * A.XW : (KB + 2 * nb) * 32;
    int nxs = NXS_MAX;
    while (nxs > NBW && smem_bytes(contig, A.Lp, nxs, A.xsf, nb) > 226 * 1024) nxs -= NBW;   // multiples of NBW only
    A.nxs = nxs;
    const size_t smem = smem_bytes(contig, A.Lp, nxs, A.xsf, nb);
    if (smem > 226 * 1024) return -1;

    EncodeTiledFn enc = get_encode();
    if (!enc) return -2;
    CUtensorMap tm;
    cuuint64_t gdim[3], gstr[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    int rank;
    if (contig) {
        rank = 2;
        A.na = n2 * n3;                                          // lines
        A.ntiles = ((long)A.na + 31) / 32;
        gdim[0] = (cuuint64_t)n1; gdim[1] = (cuuint64_t)A.na;
        gstr[0] = (cuuint64_t)n1 * 4;
        box[0] = (cuuint32_t)A.XW; box[1] = 32;
    } else {
        rank = 3;
        if (axis == 1) { A.na = n1; A.d = n1; A.sb = (long)n1 * n2; gdim[0] = n1; gdim[1] = n2; gdim[2] = n3; }
        else { A.na = n1 * n2; A.d = (long)n1 * n2; A.sb = (long)n1 * n2 * n3; gdim[0] = (cuuint64_t)n1 * n2; gdim[1] = n3; gdim[2] = 1; }
        A.tilesA = (A.na + 31) / 32;
        A.ntiles = (long)A.tilesA * (long)gdim[2];
        gstr[0] = (cuuint64_t)A.d * 4; gstr[1] = (cuuint64_t)A.sb * 4;
        box[0] = 32; box[1] = (cuuint32_t)(KB + 2 * nb); box[2] = 1;
    }
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void *)src, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -3;

    const int ctas_per_sm = (int)((227 * 1024) / (smem + 1024)) > 0 ? (int)((227 * 1024) / (smem + 1024)) : 1;
    long grid = (long)sm_count * ctas_per_sm;
    if (grid > A.ntiles) grid = A.ntiles;
    if (grid_out) *grid_out = (int)grid;
    int rc = 0;
    switch (nb) {
#define TS_CASE(N) case N: rc = launch_nb<N>(contig, (unsigned)grid, smem, stream, tm, A); break;
        TS_CASE(2) TS_CASE(3) TS_CASE(4) TS_CASE(5) TS_CASE(6) TS_CASE(7) TS_CASE(8) TS_CASE(10)
#undef TS_CASE
        default: rc = launch_nb<0>(contig, (unsigned)grid, smem, stream, tm, A); break;
    }
    if (rc) return rc;
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}
