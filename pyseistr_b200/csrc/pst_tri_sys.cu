// G2, third smoother (EXPERIMENTAL, off by default: PST_TRI_SYS=1): systolic, register-resident triangle smoothing
// of the strided axes.  Design and device code: pst_tri_sys_kernels.cuh.  This file binds the primitives the device
// code is written over (mbarrier, TMA) to PTX and holds the launcher.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "pst_tri_sys.cuh"

typedef uint64_t mbar_t;
#define PST_SYS_DEV __device__ __forceinline__
#define PST_SYS_GLOBAL(T, MINB) __global__ __launch_bounds__(T, MINB)
#define PST_SYS_TMAP_PARAM __grid_constant__ CUtensorMap
#define PST_SYS_UNROLL _Pragma("unroll")
#define PST_SYS_SMEM(name) \
    extern __shared__ __align__(1024) unsigned char pst_sys_smem_raw[]; \
    float *const name = reinterpret_cast<float *>(pst_sys_smem_raw)

namespace {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(mbar_t *b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_fence_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(mbar_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(mbar_t *b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// spin with a wall-clock bound: a protocol bug must surface as an error, never as a hung GPU
__device__ __forceinline__ void mbar_wait(mbar_t *b, unsigned parity, unsigned *err)
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
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *m, int c0, int c1, mbar_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *m, int c0, int c1, int c2, mbar_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

}  // namespace

#include "pst_tri_sys_kernels.cuh"

namespace {

using namespace tri_sys_k;

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

template <bool CONTIG, int NB, int SEG, bool ILS, bool PRE>
int launch(cudaStream_t stream, unsigned grid, const CUtensorMap &tm, const Args &A)
{
    static bool attr[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -4;
    const size_t smem = Layout<CONTIG, NB, SEG>::bytes;
    if (!attr[dev]) {
        if (cudaFuncSetAttribute(tri_sys_kernel<CONTIG, NB, SEG, ILS, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -4;
        attr[dev] = true;
    }
    tri_sys_kernel<CONTIG, NB, SEG, ILS, PRE><<<grid, NTHREADS, smem, stream>>>(tm, A);
    return 0;
}

template <int NB>
int launch_seg(bool contig, int SEG, cudaStream_t stream, unsigned grid, const CUtensorMap &tm, const Args &A)
{
    // PST_TRI_SYS_ILS=1: outputs stored from inside the backward chain loop (strided axes)
    // PST_TRI_SYS_PRE=1: t of the next tile built in place in the x box while waiting for a carry
    static const bool ils = []() { const char *e = getenv("PST_TRI_SYS_ILS"); return e && e[0] == '1'; }();
    static const bool pre = []() { const char *e = getenv("PST_TRI_SYS_PRE"); return e && e[0] == '1'; }();
#define SYS_GO(C, I, P) (SEG == 68 ? launch<C, NB, 68, I, P>(stream, grid, tm, A) : launch<C, NB, 132, I, P>(stream, grid, tm, A))
    if (contig) return pre ? SYS_GO(true, false, true) : SYS_GO(true, false, false);
    if (ils) return pre ? SYS_GO(false, true, true) : SYS_GO(false, true, false);
    return pre ? SYS_GO(false, false, true) : SYS_GO(false, false, false);
#undef SYS_GO
}

template <int NB>
int box_width(int SEG) { return SEG == 68 ? Layout<true, NB, 68>::XW : Layout<true, NB, 132>::XW; }

}  // namespace

bool pst_tri_sys_ok(int axis, int n1, int n2, int n3, int nb, const void *src, const void *dst)
{
    if ((((uintptr_t)src) & 15) || (((uintptr_t)dst) & 15)) return false;
    if (!make_plan(axis, n1, n2, n3, nb).ok) return false;
    return get_encode() != nullptr;
}

int pst_tri_sys_launch(cudaStream_t stream, int sm_count, int axis, const float *src, float *dst, int n1, int n2,
                       int n3, int nb, unsigned *d_err)
{
    const Plan P = make_plan(axis, n1, n2, n3, nb);
    if (!P.ok) return -1;
    EncodeTiledFn enc = get_encode();
    if (!enc) return -2;
    const Args A = make_args(P, dst, d_err);
    CUtensorMap tm;
    cuuint64_t gdim[3], gstr[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    int rank = 3;
    if (axis == 0) {
        rank = 2;
        int xw = 0;
        switch (nb) {
#define SYS_CASE(N) case N: xw = box_width<N>(P.SEG); break;
            SYS_CASE(2) SYS_CASE(3) SYS_CASE(4) SYS_CASE(5) SYS_CASE(6) SYS_CASE(7) SYS_CASE(8) SYS_CASE(10)
#undef SYS_CASE
            default: return -1;
        }
        gdim[0] = (cuuint64_t)n1; gdim[1] = (cuuint64_t)P.na;
        gstr[0] = (cuuint64_t)n1 * 4;
        box[0] = (cuuint32_t)xw; box[1] = 32;
    } else {
        if (axis == 1) { gdim[0] = (cuuint64_t)n1; gdim[1] = (cuuint64_t)n2; gdim[2] = (cuuint64_t)n3; }
        else { gdim[0] = (cuuint64_t)n1 * n2; gdim[1] = (cuuint64_t)n3; gdim[2] = 1; }
        gstr[0] = (cuuint64_t)P.d * 4;
        gstr[1] = (cuuint64_t)(axis == 1 ? P.sb : P.d * (long)n3) * 4;
        box[0] = 32; box[1] = (cuuint32_t)(P.SEG + 2 * nb); box[2] = 1;
    }
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void *)src, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -3;
    const int per_sm = (P.SEG == 68 && axis != 0) ? 2 : 1;
    long grid = (long)sm_count * per_sm;
    if (grid > P.ntiles) grid = P.ntiles;
    int rc = 0;
    switch (nb) {
#define SYS_CASE(N) case N: rc = launch_seg<N>(axis == 0, P.SEG, stream, (unsigned)grid, tm, A); break;
        SYS_CASE(2) SYS_CASE(3) SYS_CASE(4) SYS_CASE(5) SYS_CASE(6) SYS_CASE(7) SYS_CASE(8) SYS_CASE(10)
#undef SYS_CASE
        default: return -1;
    }
    if (rc) return rc;
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}
