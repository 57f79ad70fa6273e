// G2, second smoother (EXPERIMENTAL, off by default: PST_TRI_RC=1|2): triangle smoothing of every line of one
// axis by checkpoint + recompute (pst_tri_rc_core.h; bit-identical to ps_smooth2, reference dip_cfuns.c:564-625).
// Where the streaming kernel (pst_tri_stream.cu) keeps F of 32 lines per SM on chip and is bound by one chain
// warp per SM, this one keeps np/RC floats per line, so every lane of every resident warp owns a line and the SM
// hides latency with occupancy; the price is a second read of x (12 B per sample instead of 8).
//   strided axes (2, 3): lane = 32 consecutive i1 -> every load/store of a warp is one 128-byte row
//   contiguous axis (1): lane = line; a warp moves 32 lines x 32 samples as 32 coalesced rows and transposes them
//                        through a padded shared-memory tile (conflict-free both ways)
// Checkpoints live in shared memory ([block][thread]).  No inter-warp synchronisation anywhere.
#include <cuda_runtime.h>
#include <stdint.h>

#include "pst_tri_rc.cuh"
#include "pst_tri_rc_core.h"

namespace {

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

template <int NB, int RC>
int launch_strided(cudaStream_t st, dim3 grid, size_t smem, const float *src, float *dst, long na, long d, long sb,
                   int nx, float wm, float w2)
{
    static bool attr[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -4;
    if (smem > 48 * 1024 && !attr[dev]) {
        if (cudaFuncSetAttribute(tri_rc_strided_kernel<NB, RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) return -4;
        attr[dev] = true;
    }
    tri_rc_strided_kernel<NB, RC><<<grid, TPB, smem, st>>>(src, dst, na, d, sb, nx, wm, w2);
    return 0;
}

template <int NB>
int launch_contig(cudaStream_t st, unsigned grid, size_t smem, const float *src, float *dst, long nlines, int n1,
                  float wm, float w2, int nblk)
{
    static bool attr[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -4;
    if (smem > 48 * 1024 && !attr[dev]) {
        if (cudaFuncSetAttribute(tri_rc_contig_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) return -4;
        attr[dev] = true;
    }
    tri_rc_contig_kernel<NB><<<grid, TPB, smem, st>>>(src, dst, nlines, n1, wm, w2, nblk);
    return 0;
}

int block_len(int nb, int rc_pref) { return (rc_pref == 16 && 2 * nb <= 16) ? 16 : 32; }

bool nb_built(int nb) { return (nb >= 2 && nb <= 8) || nb == 10 || nb == 16; }

}  // namespace

bool pst_tri_rc_ok(int axis, int n1, int n2, int n3, int nb, int rc_pref)
{
    const int nn[3] = {n1, n2, n3};
    const int nx = nn[axis];
    if (!nb_built(nb) || nb > nx) return false;
    const int RC = axis == 0 ? 32 : block_len(nb, rc_pref);
    const int nblk = (nx + 2 * nb + RC - 1) / RC;
    const size_t smem = (size_t)nblk * TPB * 4 + (axis == 0 ? (size_t)(TPB / 32) * 32 * 33 * 4 : 0);
    if (smem > 100 * 1024) return false;
    if (axis == 1 && n3 > 65535) return false;
    if ((long)nx * 2 >= (1L << 30)) return false;
    return true;
}

int pst_tri_rc_launch(cudaStream_t stream, int axis, const float *src, float *dst, int n1, int n2, int n3, int nb,
                      int rc_pref)
{
    if (!pst_tri_rc_ok(axis, n1, n2, n3, nb, rc_pref)) return -1;
    const float wt = (float)(1.0 / ((double)nb * nb));          // ps_triangle_init dip_cfuns.c:421
    const float wm = -wt, w2 = (float)(2. * wt);
    int rc = 0;
    if (axis == 0) {
        const long nlines = (long)n2 * n3;
        const int nblk = (n1 + 2 * nb + 31) / 32;
        const size_t smem = (size_t)nblk * TPB * 4 + (size_t)(TPB / 32) * 32 * 33 * 4;
        const long groups = (nlines + 31) / 32;
        const unsigned grid = (unsigned)((groups + TPB / 32 - 1) / (TPB / 32));
        switch (nb) {
#define RC_CASE(N) case N: rc = launch_contig<N>(stream, grid, smem, src, dst, nlines, n1, wm, w2, nblk); break;
            RC_CASE(2) RC_CASE(3) RC_CASE(4) RC_CASE(5) RC_CASE(6) RC_CASE(7) RC_CASE(8) RC_CASE(10) RC_CASE(16)
#undef RC_CASE
            default: return -1;
        }
    } else {
        const int nx = axis == 1 ? n2 : n3;
        const long na = axis == 1 ? n1 : (long)n1 * n2;
        const long d = na;
        const long sb = axis == 1 ? (long)n1 * n2 : 0;
        const int RC = block_len(nb, rc_pref);
        const int nblk = (nx + 2 * nb + RC - 1) / RC;
        const size_t smem = (size_t)nblk * TPB * 4;
        dim3 grid((unsigned)((na + TPB - 1) / TPB), axis == 1 ? (unsigned)n3 : 1u);
        switch (nb) {
#define RC_CASE(N) \
    case N: \
        rc = (RC == 16) ? launch_strided<N, (2 * N <= 16 ? 16 : 32)>(stream, grid, smem, src, dst, na, d, sb, nx, wm, w2) \
                        : launch_strided<N, 32>(stream, grid, smem, src, dst, na, d, sb, nx, wm, w2); \
        break;
            RC_CASE(2) RC_CASE(3) RC_CASE(4) RC_CASE(5) RC_CASE(6) RC_CASE(7) RC_CASE(8) RC_CASE(10) RC_CASE(16)
#undef RC_CASE
            default: return -1;
        }
    }
    if (rc) return rc;
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}
