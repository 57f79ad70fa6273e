// Device code of the checkpoint + recompute smoother (see pst_tri_rc.cu).  Kept in a header of its own so that
// tests/native/tri_rc_emul.cpp can run these very kernels on the host (one std::thread per CUDA thread, a
// barrier per warp for __syncwarp) and check their indexing and the warp transposition without a GPU.
#pragma once
#include "pst_tri_rc_core.h"

namespace tri_rc_k {

constexpr int TPB = 128;

template <int NB, int RC>
__global__ void __launch_bounds__(TPB, RC == 32 ? 3 : 5)
tri_rc_strided_kernel(const float *src, float *dst, long na, long d, long sb, int nx, float wm, float w2)
{
    extern __shared__ float ck[];
    const long a = (long)blockIdx.x * TPB + threadIdx.x;
    if (a >= na) return;
    const long base = a + (long)blockIdx.y * sb;
    tri_rc::StridedIO<RC> io;
    io.s = src + base; io.d = dst + base; io.st = d; io.nx = nx;
    tri_rc::process_line<NB, RC>(io, nx, wm, w2, ck + threadIdx.x, TPB);
}

// 32 lines x 32 samples per warp, transposed through a [32][33] tile
struct ContigIO {
    const float *s; float *d;      // first of the warp's lines
    int n1, nx, rows, lane;
    float *tile;
    float raw[32];
    bool live;
    __device__ __forceinline__ void prefetch(int m)
    {
        live = m >= 0 && m * 32 < nx;
        if (!live) return;
        const int i = m * 32 + lane;
#pragma unroll
        for (int r = 0; r < 32; r++) raw[r] = (r < rows && i < nx) ? s[(long)r * n1 + i] : 0.f;
    }
    __device__ __forceinline__ void take(float *x)
    {
        if (!live) {
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = 0.f;
            return;
        }
#pragma unroll
        for (int r = 0; r < 32; r++) tile[r * 33 + lane] = raw[r];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j++) x[j] = tile[lane * 33 + j];
        __syncwarp();
    }
    __device__ __forceinline__ void store_block(const float *v, int i0, int jlo, int jhi)
    {
#pragma unroll
        for (int j = 0; j < 32; j++) tile[lane * 33 + j] = v[j];
        __syncwarp();
        if (lane >= jlo && lane < jhi) {
#pragma unroll
            for (int r = 0; r < 32; r++)
                if (r < rows) d[(long)r * n1 + i0 + lane] = tile[r * 33 + lane];
        }
        __syncwarp();
    }
    __device__ __forceinline__ void store_one(int i, float v)
    {
        if (lane < rows) d[(long)lane * n1 + i] = v;        // the nb left-reflected outputs of a line: rare
    }
};

template <int NB>
__global__ void __launch_bounds__(TPB, 3)
tri_rc_contig_kernel(const float *src, float *dst, long nlines, int n1, float wm, float w2, int nblk)
{
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long line0 = ((long)blockIdx.x * (TPB / 32) + warp) * 32;
    if (line0 >= nlines) return;                                // whole warp
    ContigIO io;
    io.s = src + line0 * n1; io.d = dst + line0 * n1; io.n1 = n1; io.nx = n1;
    io.rows = (int)(nlines - line0 < 32 ? nlines - line0 : 32);
    io.lane = lane; io.tile = sm + (size_t)nblk * TPB + warp * (32 * 33); io.live = false;
    tri_rc::process_line<NB, 32>(io, n1, wm, w2, sm + threadIdx.x, TPB);
}

// launch geometry, shared by pst_tri_rc_launch and the host emulation
struct Plan {
    bool ok;
    int RC, nblk, nx;
    long na, d, sb, nlines;
    unsigned gx, gy;
    size_t smem;
    float wm, w2;
};

inline bool nb_built(int nb) { return (nb >= 2 && nb <= 8) || nb == 10 || nb == 16; }

inline Plan make_plan(int axis, int n1, int n2, int n3, int nb, int rc_pref)
{
    Plan P{};
    const int nn[3] = {n1, n2, n3};
    P.nx = nn[axis];
    if (axis < 0 || axis > 2 || !nb_built(nb) || nb > P.nx) return P;
    P.RC = (axis != 0 && rc_pref == 16 && 2 * nb <= 16) ? 16 : 32;
    P.nblk = (P.nx + 2 * nb + P.RC - 1) / P.RC;
    const float wt = (float)(1.0 / ((double)nb * nb));          // ps_triangle_init dip_cfuns.c:421
    P.wm = -wt;
    P.w2 = (float)(2. * wt);
    if (axis == 0) {
        P.nlines = (long)n2 * n3;
        const long groups = (P.nlines + 31) / 32;
        const long blocks = (groups + TPB / 32 - 1) / (TPB / 32);
        if (blocks >= (1L << 31)) return P;
        P.gx = (unsigned)blocks; P.gy = 1;
        P.smem = (size_t)P.nblk * TPB * 4 + (size_t)(TPB / 32) * 32 * 33 * 4;
    } else {
        P.na = axis == 1 ? n1 : (long)n1 * n2;
        P.d = P.na;
        P.sb = axis == 1 ? (long)n1 * n2 : 0;
        if (axis == 1 && n3 > 65535) return P;
        P.gx = (unsigned)((P.na + TPB - 1) / TPB); P.gy = axis == 1 ? (unsigned)n3 : 1u;
        P.smem = (size_t)P.nblk * TPB * 4;
    }
    if (P.smem > 100 * 1024) return P;
    P.ok = true;
    return P;
}

}  // namespace tri_rc_k
