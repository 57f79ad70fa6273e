// Device code of the SYSTOLIC triangle smoother (experimental, see pst_tri_sys.cu): ps_smooth2 (reference
// dip_cfuns.c:458-484,508-529,564-580,616-625) of the lines of a strided axis, bit for bit:
//   t_k = ((-wt x_k) + 2wt x_{k-nb}) + (-wt x_{k-2nb})          k in [0, L),  L = nx + 2nb
//   F_k = F_{k-1} + t_k        forward running sum  (float, sequential)
//   B_k = B_{k+1} + F_k        backward running sum (float, sequential)
//   y_i = (B_{i+nb} + B_{nb+nx+(nx-1-i)}[i >= nx-nb]) + B_{nb-1-i}[i < nb]
// A tile is 32 whole lines (lane = line, a warp row = 128 contiguous bytes).  The padded line (D dummy steps lead
// it, so that its last step is the line's last) is cut into nseg <= 8 SEGMENTS of SEG steps; each segment of a tile
// belongs to one chain warp, which keeps the segment's t -> F -> B values IN REGISTERS (R[SEG], statically indexed:
// every loop over the segment is fully unrolled).  The forward sum sweeps the segments upwards, the backward sum
// downwards; a warp receives the running sum of the neighbouring segment through a 32-float mailbox in shared memory
// (mbarrier, one phase per tile), runs SEG dependent FADDs, and passes it on.  The ORIENTATION alternates from tile
// to tile (even tiles: warp w owns segment w; odd tiles: segment nseg-1-w), so the forward sweep of tile p+1 follows
// the backward sweep of tile p warp by warp: two sweeps are always in flight on different warps, and everything that
// is not the dependent chain (building t from x, the fold, the stores) is done by warps that would otherwise wait.
// One loader thread streams x with TMA, one box per segment (SEG + 2nb rows: every segment re-reads its 2nb halo
// rows, from L2), zero fill outside the volume = "tap skipped".  HBM traffic is the compulsory 8 B per sample, F never
// touches shared memory, and no warp is special: against pst_tri_stream.cu (one forward and one backward chain warp
// per SM, F through shared memory) the dependent chain gets 8 warps instead of 2.
//
// This header holds only device code over a handful of primitives (mbarrier, TMA, warp / CTA barriers) so that
// tests/native/tri_sys_emul.cpp can run the kernel on the host, thread for thread, and check protocol and indexing
// without a GPU.
#pragma once

namespace tri_sys_k {

constexpr int NCW = 8;                      // chain warps = maximum number of segments
constexpr int NTHREADS = (NCW + 1) * 32;    // + the loader warp

struct Args {
    float *dst;
    long d, sb;            // element stride along the line, batch (slab) stride
    int na;                // extent of the lane index
    int nx;
    int tilesA;            // tiles along na
    long ntiles;
    int nseg, L, D;        // segments in use, nx + 2nb, leading dummy steps (nseg * SEG - L)
    float wm, w2;
    unsigned *err;
};

// shared-memory layout (floats, then barriers)
template <bool CONTIG, int NB, int SEG>
struct Layout {
    static constexpr int XROWS = SEG + 2 * NB;                  // samples of a segment's x box along the line
    // contiguous axis: the box is [32 lines][XW samples]; XW is a multiple of 4 (TMA) with XW/4 odd, so that
    // "lane = line" 128-bit reads are conflict-free
    static constexpr int XW4 = (XROWS + 3) / 4 * 4;
    static constexpr int XW = (XW4 / 4) % 2 == 0 ? XW4 + 4 : XW4;
    static constexpr int XBOX = CONTIG ? 32 * XW : XROWS * 32;  // floats per box
    static constexpr size_t xs = 0;                             // [NCW][XBOX]
    static constexpr size_t scratch = xs + (size_t)NCW * XBOX;  // [SEG][32]  bottom zone of segment 0
    static constexpr size_t mf = scratch + (size_t)SEG * 32;    // [NCW][32] forward carries INTO segment s
    static constexpr size_t mb = mf + NCW * 32;                 // [NCW][32] backward carries INTO segment s
    static constexpr size_t tiles = mb + NCW * 32;              // contiguous axis: [NCW][32][33] store transposition
    static constexpr size_t floats = tiles + (CONTIG ? (size_t)NCW * 32 * 33 : 0);
    static constexpr size_t fl_al = (floats + 3) / 4 * 4;       // barriers start 16-byte aligned
    static constexpr size_t bytes = fl_al * 4 + 6 * NCW * sizeof(mbar_t);
};

PST_SYS_DEV float tri_t3(float xa, float xb, float xc, float wm, float w2)
{
    float v = wm * xa;          // 0 + wm*x_k: the sign of a zero cannot reach a non-zero sum (F starts at +0)
    v = v + w2 * xb;
    v = v + wm * xc;
    return v;
}

// ILS (strided axes only): the last and the interior segments store their outputs from INSIDE the backward chain loop
// (the chain issues one FADD per 4 cycles: the stores ride in its shadow) instead of in a pass of their own.
// PRE: t of a warp's segment of the NEXT tile is built IN PLACE in that segment's x box (row j <- t_j) while the warp
// waits for a carry of the current tile, and only moved to registers (SEG shared-memory loads) when the next tile
// starts: the lag between the backward sweep of a tile and the forward sweep of the next drops from store + build to
// store + load.
template <bool CONTIG, int NB, int SEG>
PST_SYS_DEV void sys_prebuild(float *box, int lane, float wm, float w2)
{
    typedef Layout<CONTIG, NB, SEG> LY;
    if (CONTIG) {
        float4 *const X4 = reinterpret_cast<float4 *>(box + lane * LY::XW);
        constexpr int HW = (2 * NB + 3) / 4 + 1;                 // 16-byte words held ahead of the chunk being built
        float xw[LY::XW];
        PST_SYS_UNROLL
        for (int q = 0; q < HW; q++) { const float4 v = X4[q]; xw[4 * q] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w; }
        PST_SYS_UNROLL
        for (int c = 0; c < SEG / 4; c++) {
            if (c + HW < LY::XW / 4) {
                const int q = c + HW < LY::XW / 4 ? c + HW : 0;
                const float4 v = X4[q]; xw[4 * q] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w;
            }
            float4 t;
            t.x = tri_t3(xw[4 * c + 2 * NB], xw[4 * c + NB], xw[4 * c], wm, w2);
            t.y = tri_t3(xw[4 * c + 1 + 2 * NB], xw[4 * c + 1 + NB], xw[4 * c + 1], wm, w2);
            t.z = tri_t3(xw[4 * c + 2 + 2 * NB], xw[4 * c + 2 + NB], xw[4 * c + 2], wm, w2);
            t.w = tri_t3(xw[4 * c + 3 + 2 * NB], xw[4 * c + 3 + NB], xw[4 * c + 3], wm, w2);
            X4[c] = t;
        }
    } else {
        float *const X = box + lane;
        constexpr int PF = 4;                                    // rows loaded ahead of the step being built
        float xr[LY::XROWS];
        PST_SYS_UNROLL
        for (int r = 0; r < 2 * NB + PF; r++) xr[r] = X[r * 32];
        PST_SYS_UNROLL
        for (int j = 0; j < SEG; j++) {
            if (j + 2 * NB + PF < LY::XROWS) xr[j + 2 * NB + PF < LY::XROWS ? j + 2 * NB + PF : 0] = X[(j + 2 * NB + PF) * 32];
            X[j * 32] = tri_t3(xr[j + 2 * NB], xr[j + NB], xr[j], wm, w2);
        }
    }
}

template <bool CONTIG, int NB, int SEG>
PST_SYS_DEV void sys_rload(const float *box, int lane, float *R)
{
    typedef Layout<CONTIG, NB, SEG> LY;
    if (CONTIG) {
        const float4 *const X4 = reinterpret_cast<const float4 *>(box + lane * LY::XW);
        PST_SYS_UNROLL
        for (int c = 0; c < SEG / 4; c++) { const float4 v = X4[c]; R[4 * c] = v.x; R[4 * c + 1] = v.y; R[4 * c + 2] = v.z; R[4 * c + 3] = v.w; }
    } else {
        PST_SYS_UNROLL
        for (int j = 0; j < SEG; j++) R[j] = box[j * 32 + lane];
    }
}

template <bool CONTIG, int NB, int SEG, bool ILS, bool PRE>
PST_SYS_GLOBAL(NTHREADS, (SEG <= 68 ? 2 : 1)) void tri_sys_kernel(const PST_SYS_TMAP_PARAM tmap, const Args A)
{
    static_assert(2 * NB <= SEG && SEG % 4 == 0, "segment shorter than the fold zones");
    typedef Layout<CONTIG, NB, SEG> LY;
    PST_SYS_SMEM(smem_f);
    float *const Xs = smem_f + LY::xs;
    float *const Sc = smem_f + LY::scratch;
    float *const Mf = smem_f + LY::mf;
    float *const Mb = smem_f + LY::mb;
    mbar_t *const bars = reinterpret_cast<mbar_t *>(smem_f + LY::fl_al);
    // full_x / empty_x / cf complete one phase per tile and every phase of tile p is complete before any warp starts
    // tile p+1 (a warp leaves tile p only after its backward carry, i.e. after the whole forward sweep).  The backward
    // carries are different: the warp that finishes tile p first waits for its backward carry of tile p+1 while the
    // backward sweep of tile p is still running, so a single barrier per segment would be waited on two phases ahead
    // (a parity wait then falls through: found by the host emulation).  cb is therefore doubled by tile parity, which
    // makes the waiter of each barrier always the same warp (same orientation), hence sequential.
    // full_x is doubled by tile parity as well: with PRE a warp waits for the NEXT tile's box while it works on the
    // current one, and nothing orders that wait after the landing of the current tile's box of the same segment.
    mbar_t *const full_x2 = bars, *const empty_x = bars + 2 * NCW, *const cf = bars + 3 * NCW, *const cb2 = bars + 4 * NCW;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 6 * NCW; i++) mbar_init(bars + i, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const long m = (A.ntiles - (long)blockIdx.x + (long)gridDim.x - 1) / (long)gridDim.x;   // tiles of this CTA
    const int nseg = A.nseg, D = A.D;
    if (m <= 0) return;

    if (warp == NCW) {
        // ================================ loader ================================
        if (lane != 0) return;
        const unsigned xbytes = (unsigned)(LY::XBOX * 4);
        for (long p = 0; p < m; p++) {
            const long tile = (long)blockIdx.x + p * (long)gridDim.x;
            const long b = CONTIG ? 0 : tile / A.tilesA;
            const int c0 = CONTIG ? 0 : (int)(tile - b * A.tilesA) * 32;
            for (int s = 0; s < nseg; s++) {
                mbar_t *const fx = full_x2 + (p & 1) * NCW + s;
                mbar_wait(empty_x + s, (unsigned)((p & 1) ^ 1), A.err);
                mbar_arrive_expect_tx(fx, xbytes);
                // box sample r of segment s  <->  x sample s*SEG - D - 2nb + r  (a multiple of 4 on the contiguous
                // axis, where TMA needs a 16-byte aligned start: SEG % 4 == 0 and D + 2nb = nseg*SEG - nx, nx % 4 == 0)
                if (CONTIG) tma_load_2d(Xs + (size_t)s * LY::XBOX, &tmap, s * SEG - D - 2 * NB, (int)(tile * 32), fx);
                else tma_load_3d(Xs + (size_t)s * LY::XBOX, &tmap, c0, s * SEG - D - 2 * NB, (int)b, fx);
            }
        }
        return;
    }
    if (warp >= nseg) return;

    // ================================ chain warps ================================
    float R[SEG];
    for (long p = 0; p < m; p++) {
        const unsigned par = (unsigned)(p & 1);
        const int s = par ? nseg - 1 - warp : warp;              // orientation alternates
        mbar_t *const cb = cb2 + par * NCW;
        const unsigned par2 = (unsigned)((p >> 1) & 1);
        const long tile = (long)blockIdx.x + p * (long)gridDim.x;
        const long b = CONTIG ? 0 : tile / A.tilesA;
        const int c0 = CONTIG ? 0 : (int)(tile - b * A.tilesA) * 32;
        // contiguous axis: lane = line tile*32 + lane of A.na lines; strided: lane = fast index c0 + lane of A.na
        const int rows = CONTIG ? (int)((long)A.na - tile * 32 < 32 ? (long)A.na - tile * 32 : 32) : 32;
        const bool live = CONTIG ? lane < rows : c0 + lane < A.na;
        // ---- t of the segment from its x box (every x sample is read once: the window lives in registers)
        float *const box = Xs + (size_t)s * LY::XBOX;
        if (!PRE) {
            mbar_wait(full_x2 + par * NCW + s, par2, A.err);
            if (CONTIG) {
                const float4 *row = reinterpret_cast<const float4 *>(box + lane * LY::XW);
                float xr[LY::XW];
                PST_SYS_UNROLL
                for (int q = 0; q < LY::XW / 4; q++) { const float4 v = row[q]; xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w; }
                PST_SYS_UNROLL
                for (int j = 0; j < SEG; j++) R[j] = tri_t3(xr[j + 2 * NB], xr[j + NB], xr[j], A.wm, A.w2);
            } else {
                const float *X = box + lane;
                float xr[LY::XROWS];
                PST_SYS_UNROLL
                for (int r = 0; r < LY::XROWS; r++) xr[r] = X[r * 32];
                PST_SYS_UNROLL
                for (int j = 0; j < SEG; j++) R[j] = tri_t3(xr[j + 2 * NB], xr[j + NB], xr[j], A.wm, A.w2);
            }
        } else {
            // t of this segment was built in place during the previous tile (this warp owned nothing else of that box)
            if (p == 0) { mbar_wait(full_x2 + par * NCW + s, par2, A.err); sys_prebuild<CONTIG, NB, SEG>(box, lane, A.wm, A.w2); }
            sys_rload<CONTIG, NB, SEG>(box, lane, R);
            mbar_fence_proxy();                                  // my in-place writes before the TMA unit refills the box
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_x + s);
        // with PRE: build t of my segment of the next tile while I wait for a carry -- before the forward sum for the
        // segments the forward sweep reaches late, after it for the others
        const bool pre_more = PRE && p + 1 < m;
        const int sn = nseg - 1 - s;
        const bool pre_before = 2 * s >= nseg;
        if (pre_more && pre_before) {
            mbar_wait(full_x2 + (par ^ 1u) * NCW + sn, (unsigned)(((p + 1) >> 1) & 1), A.err);
            sys_prebuild<CONTIG, NB, SEG>(Xs + (size_t)sn * LY::XBOX, lane, A.wm, A.w2);
        }
        // ---- forward sum over the segment
        float F = 0.f;
        if (s > 0) { mbar_wait(cf + s, par, A.err); F = Mf[s * 32 + lane]; }
        PST_SYS_UNROLL
        for (int j = 0; j < SEG; j++) { F = F + R[j]; R[j] = F; }
        if (s < nseg - 1) {
            Mf[(s + 1) * 32 + lane] = F;
            __syncwarp();
            if (lane == 0) mbar_arrive(cf + s + 1);
        }
        if (pre_more && !pre_before) {
            mbar_wait(full_x2 + (par ^ 1u) * NCW + sn, (unsigned)(((p + 1) >> 1) & 1), A.err);
            sys_prebuild<CONTIG, NB, SEG>(Xs + (size_t)sn * LY::XBOX, lane, A.wm, A.w2);
        }
        // ---- backward sum over the segment
        float B = 0.f;
        if (s < nseg - 1) { mbar_wait(cb + s, par2, A.err); B = Mb[s * 32 + lane]; }
        const int i0 = s * SEG - D - NB;                         // output sample of step 0 (y_i = B_{i+nb}, k = s*SEG + j - D)
        const bool inloop = ILS && !CONTIG && s > 0;
        if (inloop) {
            // byte pointer stepped by a byte stride: two integer instructions per store
            char *q = reinterpret_cast<char *>(A.dst + b * A.sb + c0 + lane + (long)(i0 + SEG - 1) * A.d);
            const long d = A.d * 4;
            if (s == nseg - 1) {
                // top nb steps first (kept: they are the right reflections of the nb outputs that follow)
                PST_SYS_UNROLL
                for (int j = SEG - 1; j >= 0; j--) {
                    B = B + R[j];
                    if (j >= SEG - NB) R[j] = B;
                    else {
                        float v = B;
                        if (j >= SEG - 2 * NB) v = v + R[2 * (SEG - NB) - 1 - j];
                        if (live) *reinterpret_cast<float *>(q) = v;
                    }
                    q -= d;
                }
            } else {
                PST_SYS_UNROLL
                for (int j = SEG - 1; j >= 0; j--) { B = B + R[j]; if (live) *reinterpret_cast<float *>(q) = B; q -= d; }
            }
        } else {
            PST_SYS_UNROLL
            for (int j = SEG - 1; j >= 0; j--) { B = B + R[j]; R[j] = B; }
        }
        if (s > 0) {
            Mb[(s - 1) * 32 + lane] = B;
            __syncwarp();
            if (lane == 0) mbar_arrive(cb + s - 1);
        }
        if (inloop) continue;
        // ---- fold2 and the stores.  Step j of segment s is k = s*SEG + j - D; y_i = B_{i+nb} goes to sample i = k - nb.
        int jlo = 0, jhi = SEG;                                  // steps with a plain output
        if (s == nseg - 1) {
            // last segment (k up to L-1): the top nb steps are the right reflections of the nb before them;
            // both sit at fixed positions of this segment
            PST_SYS_UNROLL
            for (int r = 0; r < NB; r++) R[SEG - NB - 1 - r] = R[SEG - NB - 1 - r] + R[SEG - NB + r];
            jhi = SEG - NB;
        } else if (s == 0) {
            // first segment: D dummy steps, then k in [0, nb) (left reflections), k in [nb, 2nb) (the outputs they
            // complete), then plain outputs.  D is a run-time value: the two zones go through shared memory.
            jlo = D + 2 * NB;                                    // <= SEG (checked by the plan)
            float *const sc = Sc + lane;
            PST_SYS_UNROLL
            for (int j = 0; j < SEG; j++)
                if (j < jlo) sc[j * 32] = R[j];
        }
        if (CONTIG) {
            // 32 lines x 32 steps at a time through the warp's padded tile: every global store is a row of a line
            float *const Tt = smem_f + LY::tiles + (size_t)warp * (32 * 33);
            float *const dline = A.dst + tile * 32 * (long)A.nx;
            PST_SYS_UNROLL
            for (int c = 0; c < (SEG + 31) / 32; c++) {
                PST_SYS_UNROLL
                for (int jj = 0; jj < 32; jj++)
                    if (32 * c + jj < SEG) Tt[lane * 33 + jj] = R[32 * c + jj < SEG ? 32 * c + jj : 0];
                __syncwarp();
                const int j = 32 * c + lane;
                if (j >= jlo && j < jhi) {
                    float *q = dline + i0 + j;
                    PST_SYS_UNROLL
                    for (int r = 0; r < 32; r++)
                        if (r < rows) q[(long)r * A.nx] = Tt[r * 33 + lane];
                }
                __syncwarp();
            }
            if (s == 0 && live) {
                const float *const sc = Sc + lane;
                float *q = dline + (long)lane * A.nx;
                PST_SYS_UNROLL
                for (int i = 0; i < NB; i++) q[i] = sc[(D + NB + i) * 32] + sc[(D + NB - 1 - i) * 32];
            }
        } else {
            float *const dcol = A.dst + b * A.sb + c0 + lane;
            const long d = A.d;
            const long db = d * 4;                               // byte stride: two integer instructions per store
            if (s == nseg - 1) {
                if (live) {
                    char *q = reinterpret_cast<char *>(dcol + (long)i0 * d);
                    PST_SYS_UNROLL
                    for (int j = 0; j < SEG - NB; j++) { *reinterpret_cast<float *>(q) = R[j]; q += db; }
                }
            } else if (s > 0) {
                if (live) {
                    char *q = reinterpret_cast<char *>(dcol + (long)i0 * d);
                    PST_SYS_UNROLL
                    for (int j = 0; j < SEG; j++) { *reinterpret_cast<float *>(q) = R[j]; q += db; }
                }
            } else if (live) {
                const float *const sc = Sc + lane;
                PST_SYS_UNROLL
                for (int j = 2 * NB; j < SEG; j++)
                    if (j >= jlo) dcol[(long)(i0 + j) * d] = R[j];
                PST_SYS_UNROLL
                for (int i = 0; i < NB; i++) dcol[(long)i * d] = sc[(D + NB + i) * 32] + sc[(D + NB - 1 - i) * 32];
            }
        }
    }
}

// launch geometry, shared by pst_tri_sys_launch and the host emulation
struct Plan {
    bool ok;
    int SEG, nseg, L, D, nx, nb;
    long na, d, sb, nslab, ntiles;
    int tilesA;
    float wm, w2;
};

inline bool nb_built(int nb) { return (nb >= 2 && nb <= 8) || nb == 10; }

inline Plan make_plan(int axis, int n1, int n2, int n3, int nb)
{
    Plan P{};
    if (axis < 0 || axis > 2) return P;
    P.nx = axis == 0 ? n1 : (axis == 1 ? n2 : n3);
    P.nb = nb;
    if (!nb_built(nb) || P.nx < 2 * nb) return P;
    if (n1 % 4 != 0) return P;                                   // TMA: global strides are multiples of 16 bytes
    P.L = P.nx + 2 * nb;
    // shortest built segment length that cuts the line into 2..NCW segments with both fold zones inside the end
    // segments (the bottom zone sits behind the D leading dummy steps of segment 0)
    const int segs[2] = {68, 132};
    bool found = false;
    for (int c = 0; c < 2 && !found; c++) {
        P.SEG = segs[c];
        if (P.L > NCW * P.SEG || 2 * nb > P.SEG) continue;
        P.nseg = (P.L + P.SEG - 1) / P.SEG;
        P.D = P.nseg * P.SEG - P.L;
        found = P.nseg >= 2 && P.D + 2 * nb <= P.SEG;
    }
    if (!found) return P;
    if (axis == 0) {
        P.na = (long)n2 * n3;                                    // lines
        P.d = 1; P.sb = 0; P.nslab = 1;
    } else {
        P.na = axis == 1 ? n1 : (long)n1 * n2;
        P.d = P.na;
        P.sb = axis == 1 ? (long)n1 * n2 : 0;
        P.nslab = axis == 1 ? n3 : 1;
    }
    if (P.na >= (1L << 31)) return P;
    P.tilesA = (int)((P.na + 31) / 32);
    P.ntiles = (long)P.tilesA * P.nslab;
    const float wt = (float)(1.0 / ((double)nb * nb));          // ps_triangle_init dip_cfuns.c:421
    P.wm = -wt;
    P.w2 = (float)(2. * wt);
    P.ok = true;
    return P;
}

inline Args make_args(const Plan &P, float *dst, unsigned *err)
{
    Args A{};
    A.dst = dst; A.d = P.d; A.sb = P.sb; A.na = (int)P.na; A.nx = P.nx; A.tilesA = P.tilesA; A.ntiles = P.ntiles;
    A.nseg = P.nseg; A.L = P.L; A.D = P.D; A.wm = P.wm; A.w2 = P.w2; A.err = err;
    return A;
}

}  // namespace tri_sys_k
