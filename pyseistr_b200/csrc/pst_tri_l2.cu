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
//     x 128 KB per tile = 57 - 114 MB in flight, so the second read hits L2 (pass-A loads carry an evict-last policy,
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
constexpr int MAXBLK = 72;                   // checkpoints per lane: lines up to 72 * 32 - 2 nb samples

struct Args {
    long ntiles;
    int tilesA;            // strided: tiles along the fast index (per slab); contiguous: unused
    int nx;
    int nld, nrel;         // blocks per tile loaded in pass A / again in pass B
    int nslot;             // ring slots per warp
    int hints;             // bit 0: evict-last on pass-A loads, bit 1: evict-first on pass-B loads, bit 2: on stores
    float wm, w2;
};

// The block transport of one warp: ring of TMA boxes in, two staging boxes out.  Every member function is called by
// all 32 lanes (warp-uniform control flow); lane 0 issues the TMA operations.
template <bool CONTIG>
struct WarpIO {
    const CUtensorMap *tin, *tout;
    float *ring, *stage;
    mbar_t *full;
    const Args *A;
    long tile_first, tile_step, ntiles_mine;   // this warp's tiles: tile_first + p * tile_step
    long q_issue, q_take, q_total;             // positions in the warp's whole block stream (all of its tiles)
    int per_tile;                              // blocks per tile in the stream
    int lane;
    int nstore;
    uint64_t polA, polB, polS;
    // tile coordinates of the tile being computed (for the stores)
    int cur_c0, cur_slab;

    __device__ __forceinline__ void coords(long tile, int &c0, int &slab) const
    {
        if (CONTIG) { c0 = (int)(tile * 32); slab = 0; }
        else { slab = (int)(tile / A->tilesA); c0 = (int)(tile - (long)slab * A->tilesA) * 32; }
    }
    // lane 0: put the load of stream position q in flight
    __device__ __forceinline__ void issue(long q)
    {
        const long p = q / per_tile;
        const int ql = (int)(q - p * per_tile);
        const int m = ql < A->nld ? ql : (A->nrel - 1) - (ql - A->nld);
        const uint64_t pol = ql < A->nld ? polA : polB;
        int c0, slab;
        coords(tile_first + p * tile_step, c0, slab);
        const int s = (int)(q % A->nslot);
        mbar_expect_tx(full + s, BOX * 4);
        if (CONTIG) tma_load_2d(ring + (size_t)s * BOX, tin, m * 32, c0, full + s, pol);
        else tma_load_3d(ring + (size_t)s * BOX, tin, c0, m * 32, slab, full + s, pol);
    }
    __device__ __forceinline__ void start()
    {
        q_take = 0; q_issue = 0;
        if (lane == 0)
            for (; q_issue < q_total && q_issue < A->nslot; q_issue++) issue(q_issue);
        q_issue = q_total < A->nslot ? q_total : A->nslot;
    }
    __device__ __forceinline__ void load(float *x)
    {
        const int s = (int)(q_take % A->nslot);
        mbar_wait(full + s, (unsigned)((q_take / A->nslot) & 1));
        const float *b = ring + (size_t)s * BOX;
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
        q_take++;
        if (q_issue < q_total) {
            if (lane == 0) issue(q_issue);
            q_issue++;
        }
    }
    __device__ __forceinline__ void store(const float *v, int i0)
    {
        float *st = stage + (size_t)(nstore & 1) * BOX;
        if (lane == 0) bulk_wait_read_1();              // the store that used this staging box two windows ago has read it
        __syncwarp();
        if (CONTIG) {
            float4 *row = reinterpret_cast<float4 *>(st + lane * 32);
#pragma unroll
            for (int c = 0; c < 8; c++) row[c ^ (lane & 7)] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) st[j * 32 + lane] = v[j];
        }
        fence_async();                                  // my generic-proxy writes before the TMA unit reads the box
        __syncwarp();
        if (lane == 0) {
            if (CONTIG) tma_store_2d(tout, st, i0, cur_c0, polS);
            else tma_store_3d(tout, st, cur_c0, i0, cur_slab, polS);
            bulk_commit();
        }
        nstore++;
    }
};

template <bool CONTIG, int NB>
__global__ void __launch_bounds__(192, 1)
tri_l2_kernel(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, const Args A)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    // per warp: ring [nslot][BOX] | stage [2][BOX] | checkpoints [MAXBLK][32] | barriers [nslot]
    const size_t per_warp = ((size_t)(A.nslot + 2) * BOX + MAXBLK * 32) * 4 + 1024;
    unsigned char *base = smem_raw + (size_t)warp * per_warp;
    float *ring = reinterpret_cast<float *>(base);
    float *stage = ring + (size_t)A.nslot * BOX;
    float *ck = stage + 2 * BOX;
    mbar_t *full = reinterpret_cast<mbar_t *>(ck + MAXBLK * 32);
    if (lane == 0) {
        for (int s = 0; s < A.nslot; s++) mbar_init(full + s, 1);
        fence_init();
    }
    __syncwarp();

    WarpIO<CONTIG> io;
    io.tin = &tin; io.tout = &tout; io.ring = ring; io.stage = stage; io.full = full; io.A = &A; io.lane = lane;
    io.tile_first = (long)blockIdx.x * nwarp + warp;
    io.tile_step = (long)gridDim.x * nwarp;
    io.ntiles_mine = io.tile_first < A.ntiles ? (A.ntiles - io.tile_first + io.tile_step - 1) / io.tile_step : 0;
    io.per_tile = A.nld + A.nrel;
    io.q_total = io.ntiles_mine * io.per_tile;
    io.nstore = 0;
    io.polA = (A.hints & 1) ? policy_evict_last() : 0;
    io.polB = (A.hints & 2) ? policy_evict_first() : 0;
    io.polS = (A.hints & 4) ? policy_evict_first() : 0;
    if ((A.hints & 1) == 0) { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); io.polA = p; }
    if ((A.hints & 2) == 0) { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); io.polB = p; }
    if ((A.hints & 4) == 0) { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); io.polS = p; }
    if (io.ntiles_mine == 0) return;
    io.start();
    for (long p = 0; p < io.ntiles_mine; p++) {
        io.coords(io.tile_first + p * io.tile_step, io.cur_c0, io.cur_slab);
        tri_l2::smooth_line<NB>(io, A.nx, A.wm, A.w2, ck + lane, 32);
    }
    if (lane == 0) bulk_wait_all();                     // the staging boxes must outlive their stores
    __syncwarp();
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
    int nx, nb, nld, nrel, tilesA;
    long ntiles, na;
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
    const int nblk = (P.nx + 2 * nb + 31) / 32;
    if (nblk > MAXBLK) return P;
    P.nld = (P.nx + 31) / 32;
    P.nrel = tri_l2::reload_count(P.nx, nb);
    if (axis == 0) {
        P.na = (long)n2 * n3;
        P.tilesA = 0;
        P.ntiles = (P.na + 31) / 32;
    } else {
        P.na = axis == 1 ? n1 : (long)n1 * n2;
        P.tilesA = (int)((P.na + 31) / 32);
        P.ntiles = (long)P.tilesA * (axis == 1 ? n3 : 1);
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
    static const int warps_env = env_int("PST_TRI_L2_WARPS", 4);
    static const int slots_env = env_int("PST_TRI_L2_SLOTS", 0);
    static const int hints_env = env_int("PST_TRI_L2_HINTS", 7);
    int warps = warps_env < 1 ? 1 : (warps_env > 6 ? 6 : warps_env);
    const size_t budget = 227 * 1024;
    const size_t fixed = ((size_t)2 * BOX + MAXBLK * 32) * 4 + 1024;
    int nslot = (int)((budget / warps - fixed) / (BOX * 4));
    if (slots_env > 0 && slots_env < nslot) nslot = slots_env;
    if (nslot > 16) nslot = 16;
    if (nslot < 2) return -1;
    const size_t per_warp = ((size_t)(nslot + 2) * BOX + MAXBLK * 32) * 4 + 1024;
    const size_t smem = per_warp * warps;
    Args A{};
    A.ntiles = P.ntiles; A.tilesA = P.tilesA; A.nx = P.nx; A.nld = P.nld; A.nrel = P.nrel; A.nslot = nslot;
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
