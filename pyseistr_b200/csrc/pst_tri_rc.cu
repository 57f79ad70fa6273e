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
#include "pst_tri_rc_kernels.cuh"

namespace {

using namespace tri_rc_k;

template <int NB, int RC>
int launch_strided(cudaStream_t st, const Plan &P, const float *src, float *dst)
{
    static bool attr[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -4;
    if (P.smem > 48 * 1024 && !attr[dev]) {
        if (cudaFuncSetAttribute(tri_rc_strided_kernel<NB, RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) return -4;
        attr[dev] = true;
    }
    tri_rc_strided_kernel<NB, RC><<<dim3(P.gx, P.gy), TPB, P.smem, st>>>(src, dst, P.na, P.d, P.sb, P.nx, P.wm, P.w2);
    return 0;
}

template <int NB>
int launch_contig(cudaStream_t st, const Plan &P, const float *src, float *dst)
{
    static bool attr[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -4;
    if (P.smem > 48 * 1024 && !attr[dev]) {
        if (cudaFuncSetAttribute(tri_rc_contig_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) return -4;
        attr[dev] = true;
    }
    tri_rc_contig_kernel<NB><<<P.gx, TPB, P.smem, st>>>(src, dst, P.nlines, P.nx, P.wm, P.w2, P.nblk);
    return 0;
}

}  // namespace

bool pst_tri_rc_ok(int axis, int n1, int n2, int n3, int nb, int rc_pref)
{
    return make_plan(axis, n1, n2, n3, nb, rc_pref).ok;
}

int pst_tri_rc_launch(cudaStream_t stream, int axis, const float *src, float *dst, int n1, int n2, int n3, int nb,
                      int rc_pref)
{
    const Plan P = make_plan(axis, n1, n2, n3, nb, rc_pref);
    if (!P.ok) return -1;
    int rc = 0;
    if (axis == 0) {
        switch (nb) {
#define RC_CASE(N) case N: rc = launch_contig<N>(stream, P, src, dst); break;
            RC_CASE(2) RC_CASE(3) RC_CASE(4) RC_CASE(5) RC_CASE(6) RC_CASE(7) RC_CASE(8) RC_CASE(10) RC_CASE(16)
#undef RC_CASE
            default: return -1;
        }
    } else {
        switch (nb) {
#define RC_CASE(N) \
    case N: \
        rc = (P.RC == 16) ? launch_strided<N, (2 * N <= 16 ? 16 : 32)>(stream, P, src, dst) : launch_strided<N, 32>(stream, P, src, dst); \
        break;
            RC_CASE(2) RC_CASE(3) RC_CASE(4) RC_CASE(5) RC_CASE(6) RC_CASE(7) RC_CASE(8) RC_CASE(10) RC_CASE(16)
#undef RC_CASE
            default: return -1;
        }
    }
    if (rc) return rc;
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}
