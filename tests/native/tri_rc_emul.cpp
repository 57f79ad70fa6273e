// Host emulation of the checkpoint + recompute smoother KERNELS (pyseistr_b200/csrc/pst_tri_rc_kernels.cuh): the
// kernels' own source is compiled with g++ under a minimal shim of the CUDA execution model -- one std::thread per
// CUDA thread, thread-local threadIdx / blockIdx, one barrier per warp for __syncwarp, blocks run one after the
// other with "shared memory" poisoned in between -- and launched with the geometry the real launcher uses
// (tri_rc_k::make_plan).  Test infrastructure (tests/test_tri_rc_core.py); small volumes only.
#include <pthread.h>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

struct dim3e { unsigned x, y, z; };
static thread_local dim3e threadIdx, blockIdx;
static thread_local pthread_barrier_t *warp_barrier;
static inline void __syncwarp() { pthread_barrier_wait(warp_barrier); }
// dynamic shared memory of the block that is running (each kernel declares one extern array; a block-scope
// extern declaration names a member of the enclosing namespace)
namespace tri_rc_k {
float ck[32768];
float sm[32768];
}

#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__

#include "pst_tri_rc_kernels.cuh"

using namespace tri_rc_k;

template <class F>
static void run_grid(const Plan &P, F body)
{
    for (unsigned by = 0; by < P.gy; by++)
        for (unsigned bx = 0; bx < P.gx; bx++) {
            for (size_t i = 0; i < 32768; i++) { ck[i] = NAN; sm[i] = NAN; }
            pthread_barrier_t bars[TPB / 32];
            for (auto &b : bars) pthread_barrier_init(&b, nullptr, 32);
            std::vector<std::thread> th;
            for (unsigned t = 0; t < (unsigned)TPB; t++)
                th.emplace_back([&, t]() {
                    threadIdx = {t, 0, 0};
                    blockIdx = {bx, by, 0};
                    warp_barrier = &bars[t / 32];
                    body();
                });
            for (auto &t : th) t.join();
            for (auto &b : bars) pthread_barrier_destroy(&b);
        }
}

template <int NB>
static void go(int axis, const Plan &P, const float *src, float *dst)
{
    if (axis == 0) run_grid(P, [&]() { tri_rc_contig_kernel<NB>(src, dst, P.nlines, P.nx, P.wm, P.w2, P.nblk); });
    else if (P.RC == 16) run_grid(P, [&]() { tri_rc_strided_kernel<NB, (2 * NB <= 16 ? 16 : 32)>(src, dst, P.na, P.d, P.sb, P.nx, P.wm, P.w2); });
    else run_grid(P, [&]() { tri_rc_strided_kernel<NB, 32>(src, dst, P.na, P.d, P.sb, P.nx, P.wm, P.w2); });
}

extern "C" int tri_rc_emul(const float *src, float *dst, int n1, int n2, int n3, int axis, int nb, int rc)
{
    const Plan P = make_plan(axis, n1, n2, n3, nb, rc);
    if (!P.ok) return -1;
    if (P.smem > sizeof(ck)) return -2;
    switch (nb) {
#define CASE(N) case N: go<N>(axis, P, src, dst); return 0;
        CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(10) CASE(16)
#undef CASE
    }
    return -3;
}
