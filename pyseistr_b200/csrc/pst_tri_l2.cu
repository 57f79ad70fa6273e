// G2, the triangle smoother of every axis: checkpoint + recompute with the second read served from L2.
//
// ps_smooth2 (reference dip_cfuns.c:458-484,508-529,564-580,616-625) needs, per line, a forward running sum over the
// whole line before the backward running sum can start; both are sequential float sums and are kept that way (bit
// parity).  The first smoothers kept F of 32 lines on chip (pst_tri_stream.cu: shared memory, pst_tri_sys.cu:
// registers) and were bound by the one or two warps per SM that own the dependent FADD chains (0.53 / 0.62 of the
// HBM roofline on B200); the plain checkpoint + recompute kernel (pst_tri_rc.cu) has every lane own a line, is
// HBM-bound, but moves 12 B per sample because ~57 K lines x 4 KB in flight do not fit the L2.  This kernel keeps the
// recompute formulation and sizes the set of lines in flight FOR the 126 MB L2:
//   * a warp owns a tile of 32 lines (lane = line) and is a pipeline of its own: no inter-warp synchronisation;
//   * x arrives as TMA boxes of 32 lines x 32 samples (4 KB) in a per-warp ring of shared-memory slots, several
//     boxes ahead of the arithmetic (cp.async.bulk.tensor + mbarrier complete_tx); strided axes: box = 32 rows of
//     128 contiguous bytes, lane = column; contiguous axis: box = 32 lines x 128 bytes with the 128-byte swizzle, so
//     that "lane = line" 128-bit shared-memory reads are conflict-free;
//   * pass A streams the tile upwards once (HBM), pass B streams it downwards again -- 3 to 6 warps per SM x 148 SMs
//     x up to 128 KB per tile (64 KB on average over a tile's life) in flight, so the second read hits L2 (pass-A loads carry an evict-last policy,
//     pass-B loads and the stores evict-first) and HBM sees the compulsory 8 B per sample;
//   * outputs leave as aligned 32 x 32 boxes through TMA stores (bulk groups), zero-clipped at the volume edges.
// The per-line arithmetic is pst_tri_l2_core.h (host-testable, tests/test_tri_l2_core.py).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "pst_tri_l2.cuh"
#include "pst_tri_l2_core.h"

namespace {

typedef uint64_t mbar_t;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(mbar_t *b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(mbar_t *b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// spin with a wall-clock bound: a protocol bug must surface as an error, never as a hung GPU
__device__ __forceinline__ void mbar_wait(mbar_t *b, unsigned parity)
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
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *m, int c0, int c1, mbar_t *bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *m, int c0, int c1, int c2, mbar_t *bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, const void *src, int c0, int c1, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *m, const void *src, int c0, int c1, int c2, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                 ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr int BOX = 32 * 32;                 // floats per box (4 KB)
constexpr int MAXBLK = 160;                  // checkpoints per lane: lines up to 160 * 32 - 2 nb samples

struct Args {
    float *dst;            // strided axes: outputs are stored directly (coalesced 128-byte rows)
    long d, sb;            // strided axes: element stride along the line, slab stride
    long ntiles;
    int tilesA;            // strided: tiles along the fast index (per slab); contiguous: unused
    int na;                // strided: extent of the fast index
    int nx;
    int nld, nrel;         // blocks per tile loaded in pass A / again in pass B
    int nblk;              // checkpoints per line
    int nslot;             // ring slots per warp
    int hints;             // bit 0: evict-last on pass-A loads, bit 1: evict-first on pass-B loads, bit 2: on stores
    float wm, w2;
};

// The block transport of one warp: ring of TMA boxes in; out: coalesced rows (strided axes) or TMA boxes through two
// staging boxes (contiguous axis).  Every member function is called by all 32 lanes (warp-uniform control flow); lane 0
// issues the TMA operations.  All ring / stream positions are advanced incrementally (no divisions in the loop).
template <bool CONTIG>
struct WarpIO {
    const CUtensorMap *tin, *tout;
    float *ring, *stage;
    mbar_t *full;
    const Args *A;
    long tile_step;
    long left_issue;                            // loads still to issue over all of this warp's tiles
    // issue side: next load = block stream position i_ql of tile i_tile
    long i_tile; int i_ql, i_slot, i_c0, i_slab;
    // take side
    int t_slot; unsigned t_phase;
    int per_tile;                               // blocks per tile in the stream
    int lane;
    int nstore;
    uint64_t polA, polB, polS;
    // tile being computed (for the stores)
    int cur_c0, cur_slab;

    __device__ __forceinline__ void coords(long tile, int &c0, int &slab) const
    {
        if (CONTIG) { c0 = (int)tile * 32; slab = 0; }
        else { slab = (int)tile / A->tilesA; c0 = ((int)tile - slab * A->tilesA) * 32; }
    }
    // put the next load of the stream in flight (TMA by lane 0; every lane advances the position)
    __device__ __forceinline__ void issue()
    {
        const int m = i_ql < A->nld ? i_ql : (A->nrel - 1) - (i_ql - A->nld);
        if (lane == 0) {
            const uint64_t pol = i_ql < A->nld ? polA : polB;
            mbar_expect_tx(full + i_slot, BOX * 4);
            if (CONTIG) tma_load_2d(ring + (size_t)i_slot * BOX, tin, m * 32, i_c0, full + i_slot, pol);
            else tma_load_3d(ring + (size_t)i_slot * BOX, tin, i_c0, m * 32, i_slab, full + i_slot, pol);
        }
        left_issue--;
        if (++i_slot == A->nslot) i_slot = 0;
        if (++i_ql == per_tile) {
            i_ql = 0;
            i_tile += tile_step;
            coords(i_tile, i_c0, i_slab);
        }
    }
    __device__ __forceinline__ void start(long tile_first, long ntiles_mine)
    {
        left_issue = ntiles_mine * per_tile;
        i_tile = tile_first; i_ql = 0; i_slot = 0;
        coords(i_tile, i_c0, i_slab);
        t_slot = 0; t_phase = 0;
        for (int s = 0; s < A->nslot && left_issue > 0; s++) issue();
    }
    __device__ __forceinline__ void load(float *x)
    {
        mbar_wait(full + t_slot, t_phase);
        const float *b = ring + (size_t)t_slot * BOX;
        if (CONTIG) {
            // box = [line][32 samples], 16-byte chunks XOR-swizzled by (line & 7)
            const float4 *row = reinterpret_cast<const float4 *>(b + lane * 32);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const float4 v = row[c ^ (lane & 7)];
                x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = b[j * 32 + lane];
        }
        __syncwarp();                                   // every lane has read the slot: it may be refilled
        if (++t_slot == A->nslot) { t_slot = 0; t_phase ^= 1u; }
        if (left_issue > 0) issue();                    // refills the slot just read (issue and take advance in step)
    }
    __device__ __forceinline__ void store(const float *v, int i0)
    {
        if (CONTIG) {
            float *st = stage + (size_t)(nstore & 1) * BOX;
            if (lane == 0) bulk_wait_read_1();          // the store that used this staging box two windows ago has read it
            __syncwarp();
            float4 *row = reinterpret_cast<float4 *>(st + lane * 32);
#pragma unroll
            for (int c = 0; c < 8; c++) row[c ^ (lane & 7)] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            fence_async();                              // my generic-proxy writes before the TMA unit reads the box
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(tout, st, i0, cur_c0, polS);
                bulk_commit();
            }
            nstore++;
        } else {
            // lane = fast index: every row of the window is one 128-byte store of the warp
            const int col = cur_c0 + lane;
            if (col < A->na) {
                float *q = A->dst + (long)cur_slab * A->sb + col + (long)i0 * A->d;
                const long db = A->d;
                if (i0 + 32 <= A->nx) {
#pragma unroll
                    for (int j = 0; j < 32; j++) { __stcs(q, v[j]); q += db; }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) { if (i0 + j < A->nx) __stcs(q, v[j]); q += db; }
                }
            }
        }
    }
};

// per-warp shared memory: ring [nslot][BOX] | stage [2][BOX] (contiguous axis) | checkpoints [nblk][32] | barriers
__host__ __device__ inline size_t per_warp_bytes(bool contig, int nslot, int nblk)
{
    size_t b = ((size_t)(nslot + (contig ? 2 : 0)) * BOX + (size_t)nblk * 32) * 4 + 16 * sizeof(mbar_t);
    return (b + 1023) & ~(size_t)1023;
}

template <bool CONTIG, int NB>
__global__ void __launch_bounds__(256, 1)
tri_l2_kernel(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, const Args A)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    unsigned char *base = smem_raw + (size_t)warp * per_warp_bytes(CONTIG, A.nslot, A.nblk);
    float *ring = reinterpret_cast<float *>(base);
    float *stage = ring + (size_t)A.nslot * BOX;
    float *ck = stage + (CONTIG ? 2 * BOX : 0);
    mbar_t *full = reinterpret_cast<mbar_t *>(ck + (size_t)A.nblk * 32);
    if (lane == 0) {
        for (int s = 0; s < A.nslot; s++) mbar_init(full + s, 1);
        fence_init();
    }
    __syncwarp();

    WarpIO<CONTIG> io;
    io.tin = &tin; io.tout = &tout; io.ring = ring; io.stage = stage; io.full = full; io.A = &A; io.lane = lane;
    const long tile_first = (long)blockIdx.x * nwarp + warp;
    io.tile_step = (long)gridDim.x * nwarp;
    const long ntiles_mine = tile_first < A.ntiles ? (A.ntiles - tile_first + io.tile_step - 1) / io.tile_step : 0;
    io.per_tile = A.nld + A.nrel;
    io.nstore = 0;
    io.polA = (A.hints & 1) ? policy_evict_last() : policy_evict_normal();
    io.polB = (A.hints & 2) ? policy_evict_first() : policy_evict_normal();
    io.polS = (A.hints & 4) ? policy_evict_first() : policy_evict_normal();
    if (ntiles_mine == 0) return;
    io.start(tile_first, ntiles_mine);
    long tile = tile_first;
    for (long p = 0; p < ntiles_mine; p++, tile += io.tile_step) {
        io.coords(tile, io.cur_c0, io.cur_slab);
        tri_l2::smooth_line<NB>(io, A.nx, A.wm, A.w2, ck + lane, 32);
    }
    if (CONTIG) {
        if (lane == 0) bulk_wait_all();                 // the staging boxes must outlive their stores
        __syncwarp();
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

struct Plan {
    bool ok;
    int nx, nb, nld, nrel, nblk, tilesA;
    long ntiles, na, d, sb;
    float wm, w2;
};

bool nb_built(int nb) { return (nb >= 2 && nb <= 8) || nb == 10 || nb == 16; }

Plan make_plan(int axis, int n1, int n2, int n3, int nb)
{
    Plan P{};
    if (axis < 0 || axis > 2) return P;
    const int nn[3] = {n1, n2, n3};
    P.nx = nn[axis]; P.nb = nb;
    if (!nb_built(nb) || nb > P.nx) return P;
    if (n1 % 4 != 0) return P;                                   // TMA: global strides are multiples of 16 bytes
    P.nblk = (P.nx + 2 * nb + 31) / 32;
    if (P.nblk > MAXBLK) return P;
    P.nld = (P.nx + 31) / 32;
    P.nrel = tri_l2::reload_count(P.nx, nb);
    if (axis == 0) {
        P.na = (long)n2 * n3;
        P.tilesA = 0;
        P.ntiles = (P.na + 31) / 32;
        P.d = 1; P.sb = 0;
    } else {
        P.na = axis == 1 ? n1 : (long)n1 * n2;
        P.tilesA = (int)((P.na + 31) / 32);
        P.ntiles = (long)P.tilesA * (axis == 1 ? n3 : 1);
        P.d = P.na;
        P.sb = axis == 1 ? (long)n1 * n2 : 0;
    }
    if (P.na >= (1L << 31) || P.ntiles * 32 >= (1L << 31)) return P;
    const float wt = (float)(1.0 / ((double)nb * nb));          // ps_triangle_init dip_cfuns.c:421
    P.wm = -wt;
    P.w2 = (float)(2. * wt);
    P.ok = true;
    return P;
}

int encode_map(EncodeTiledFn enc, CUtensorMap *tm, int axis, const float *ptr, int n1, int n2, int n3)
{
    cuuint64_t gdim[3], gstr[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    int rank;
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
    if (axis == 0) {
        rank = 2;
        gdim[0] = (cuuint64_t)n1; gdim[1] = (cuuint64_t)n2 * n3;
        gstr[0] = (cuuint64_t)n1 * 4;
        box[0] = 32; box[1] = 32;
        sw = CU_TENSOR_MAP_SWIZZLE_128B;
    } else {
        rank = 3;
        if (axis == 1) { gdim[0] = (cuuint64_t)n1; gdim[1] = (cuuint64_t)n2; gdim[2] = (cuuint64_t)n3; }
        else { gdim[0] = (cuuint64_t)n1 * n2; gdim[1] = (cuuint64_t)n3; gdim[2] = 1; }
        gstr[0] = gdim[0] * 4;
        gstr[1] = gdim[0] * gdim[1] * 4;
        box[0] = 32; box[1] = 32; box[2] = 1;
    }
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void *)ptr, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -3;
}

int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

template <bool CONTIG, int NB>
int launch(cudaStream_t stream, unsigned grid, int warps, size_t smem, const CUtensorMap &ti, const CUtensorMap &to, const Args &A)
{
    static bool attr[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -4;
    if (!attr[dev]) {
        if (cudaFuncSetAttribute(tri_l2_kernel<CONTIG, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return -4;
        attr[dev] = true;
    }
    tri_l2_kernel<CONTIG, NB><<<grid, warps * 32, smem, stream>>>(ti, to, A);
    return 0;
}

}  // namespace

bool pst_tri_l2_ok(int axis, int n1, int n2, int n3, int nb, const void *src, const void *dst)
{
    if ((((uintptr_t)src) & 15) || (((uintptr_t)dst) & 15)) return false;
    if (!make_plan(axis, n1, n2, n3, nb).ok) return false;
    return get_encode() != nullptr;
}

// Tunables (environment, read once): PST_TRI_L2_WARPS = warps per SM (2..6, default 4: 148 x 4 tiles of 128 KB = 76 MB
// of lines in flight), PST_TRI_L2_SLOTS = ring slots per warp (default: what fits), PST_TRI_L2_HINTS = L2 policy bits.
int pst_tri_l2_launch(cudaStream_t stream, int sm_count, int axis, const float *src, float *dst, int n1, int n2, int n3, int nb)
{
    const Plan P = make_plan(axis, n1, n2, n3, nb);
    if (!P.ok) return -1;
    EncodeTiledFn enc = get_encode();
    if (!enc) return -2;
    // measured on B200, 1000x1024x1024 (profiles/r02_tri_ncu_summary.md): contiguous axis 6 warps x 4 slots (2.07 ms per
    // pass), strided axes 8 warps x 2 slots (2.05 - 2.29 ms); more warps in flight miss the L2 on the second read
    static const int warps_env = env_int("PST_TRI_L2_WARPS", 0);
    static const int slots_env = env_int("PST_TRI_L2_SLOTS", 0);
    static const int hints_env = env_int("PST_TRI_L2_HINTS", 7);
    const bool contig = axis == 0;
    int warps = warps_env > 0 ? warps_env : (contig ? 6 : 8);
    if (warps_env <= 0 && contig) {
        // tiles are dealt to the warps round-robin: with few tiles per warp (a slab of a multi-GPU run: 4096 tiles over
        // 148 x 6 warps = 4.6 rounds) the last, partial round costs a whole one.  Pick the warp count with the cheapest
        // rounds x warps, weighted by the steady-state cost per tile measured at 5 / 6 / 7 / 8 warps (2.176 / 2.073 /
        // 2.135 / 2.158 ms per pass at 1000x1024x1024).
        const double f[4] = {1.050, 1.000, 1.030, 1.041};
        double best = 0.;
        for (int w = 5; w <= 8; w++) {
            const double rounds = (double)((P.ntiles + (long)sm_count * w - 1) / ((long)sm_count * w));
            const double cost = rounds * w * f[w - 5];
            if (best == 0. || cost < best - 1e-9) { best = cost; warps = w; }
        }
    }
    warps = warps < 1 ? 1 : (warps > 8 ? 8 : warps);
    const size_t budget = 227 * 1024;
    int nslot = slots_env > 0 ? slots_env : (contig ? 4 : 2);
    if (nslot > 16) nslot = 16;
    while (nslot > 2 && per_warp_bytes(contig, nslot, P.nblk) * warps > budget) nslot--;
    while (warps > 1 && per_warp_bytes(contig, nslot, P.nblk) * warps > budget) warps--;
    if (per_warp_bytes(contig, nslot, P.nblk) * warps > budget) return -1;
    const size_t smem = per_warp_bytes(contig, nslot, P.nblk) * warps;
    Args A{};
    A.dst = dst; A.d = P.d; A.sb = P.sb; A.na = (int)(P.na < (1L << 31) ? P.na : 0);
    A.ntiles = P.ntiles; A.tilesA = P.tilesA; A.nx = P.nx; A.nld = P.nld; A.nrel = P.nrel; A.nblk = P.nblk; A.nslot = nslot;
    A.hints = hints_env; A.wm = P.wm; A.w2 = P.w2;
    CUtensorMap ti, to;
    if (encode_map(enc, &ti, axis, src, n1, n2, n3) || encode_map(enc, &to, axis, dst, n1, n2, n3)) return -3;
    long grid = sm_count;
    const long need = (P.ntiles + warps - 1) / warps;
    if (grid > need) grid = need;
    int rc = 0;
    switch (nb) {
#define L2_CASE(N) case N: rc = axis == 0 ? launch<true, N>(stream, (unsigned)grid, warps, smem, ti, to, A) \
                                          : launch<false, N>(stream, (unsigned)grid, warps, smem, ti, to, A); break;
        L2_CASE(2) L2_CASE(3) L2_CASE(4) L2_CASE(5) L2_CASE(6) L2_CASE(7) L2_CASE(8) L2_CASE(10) L2_CASE(16)
#undef L2_CASE
        default: return -1;
    }
    if (rc) return rc;
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}
