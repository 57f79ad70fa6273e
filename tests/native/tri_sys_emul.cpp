// Host emulation of the systolic smoother KERNEL (pyseistr_b200/csrc/pst_tri_sys_kernels.cuh): the kernel's own source
// compiled with g++ over host versions of its primitives -- one std::thread per CUDA thread, thread-local
// threadIdx / blockIdx, mbarriers with the hardware's phase-parity semantics (a waiter two phases ahead falls
// through, exactly the bug class the protocol has to exclude), TMA boxes copied with zero fill outside the tensor,
// a barrier per warp for __syncwarp -- launched with the geometry the real launcher uses (tri_sys_k::make_plan).
// CTAs run one after the other.  Test infrastructure (tests/test_tri_sys_emul.py); small volumes only.
#include <pthread.h>
#include <sched.h>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

struct dim3e { unsigned x, y, z; };
static thread_local dim3e threadIdx, blockIdx;
static dim3e gridDim;
static thread_local pthread_barrier_t *warp_barrier;
static pthread_barrier_t cta_barrier;
static inline void __syncwarp() { pthread_barrier_wait(warp_barrier); }
static inline void __syncthreads() { pthread_barrier_wait(&cta_barrier); }

struct mbar_t {
    pthread_mutex_t mu;
    unsigned phase;        // index of the current (incomplete) phase
    int pending, count;
    long tx;
};
static std::atomic<int> g_timeout{0};
// schedule fuzzing (PST_EMUL_JITTER=n): random pauses around every barrier operation, so that repeated runs
// explore interleavings the OS scheduler alone would not produce
static int g_jitter = 0;
static thread_local unsigned g_rng = 0;
static inline void jitter()
{
    if (!g_jitter) return;
    if (!g_rng) g_rng = 2463534242u ^ (unsigned)(size_t)&g_rng ^ (unsigned)std::chrono::steady_clock::now().time_since_epoch().count();
    g_rng ^= g_rng << 13; g_rng ^= g_rng >> 17; g_rng ^= g_rng << 5;
    const unsigned r = g_rng % 16;
    if (r < 4) std::this_thread::sleep_for(std::chrono::microseconds((g_rng >> 8) % (unsigned)g_jitter));
    else if (r < 8) sched_yield();
}
static inline void mbar_complete_if_done(mbar_t *b) { if (b->pending == 0 && b->tx == 0) { b->phase++; b->pending = b->count; } }
static inline void mbar_init(mbar_t *b, unsigned count)
{
    pthread_mutex_init(&b->mu, nullptr);
    b->phase = 0; b->pending = (int)count; b->count = (int)count; b->tx = 0;
}
static inline void mbar_fence_init() {}
static inline void mbar_fence_proxy() {}
static inline void mbar_arrive(mbar_t *b)
{
    jitter();
    pthread_mutex_lock(&b->mu); b->pending--; mbar_complete_if_done(b); pthread_mutex_unlock(&b->mu);
}
static inline void mbar_arrive_expect_tx(mbar_t *b, unsigned bytes)
{
    jitter();
    pthread_mutex_lock(&b->mu); b->tx += bytes; b->pending--; mbar_complete_if_done(b); pthread_mutex_unlock(&b->mu);
}
static inline void mbar_complete_tx(mbar_t *b, unsigned bytes)
{
    pthread_mutex_lock(&b->mu); b->tx -= bytes; mbar_complete_if_done(b); pthread_mutex_unlock(&b->mu);
}
// try_wait.parity P: true when the most recently completed phase has parity P
static inline void mbar_wait(mbar_t *b, unsigned parity, unsigned *)
{
    jitter();
    const auto t0 = std::chrono::steady_clock::now();
    for (long spin = 0;; spin++) {
        pthread_mutex_lock(&b->mu);
        const bool ok = (b->phase & 1u) != parity;
        pthread_mutex_unlock(&b->mu);
        if (ok) { jitter(); return; }
        sched_yield();
        if ((spin & 1023) == 1023 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) {
            g_timeout = 1;
            fprintf(stderr, "tri_sys_emul: barrier wait timed out (deadlock)\n");
            abort();
        }
    }
}

// tiled tensor map, rank 3, float32, zero fill
struct TMapE {
    const float *base;
    long dim[3];
    long stride[3];        // elements; stride[0] = 1
    int box[3];
};
static inline void tma_load_3d(void *dst, const TMapE *m, int c0, int c1, int c2, mbar_t *bar)
{
    float *o = (float *)dst;
    for (int z = 0; z < m->box[2]; z++)
        for (int y = 0; y < m->box[1]; y++)
            for (int x = 0; x < m->box[0]; x++) {
                const long i0 = c0 + x, i1 = (long)c1 + y, i2 = (long)c2 + z;
                const bool in = i0 >= 0 && i0 < m->dim[0] && i1 >= 0 && i1 < m->dim[1] && i2 >= 0 && i2 < m->dim[2];
                *o++ = in ? m->base[i0 + i1 * m->stride[1] + i2 * m->stride[2]] : 0.f;
            }
    mbar_complete_tx(bar, (unsigned)((size_t)m->box[0] * m->box[1] * m->box[2] * 4));
}

static inline void tma_load_2d(void *dst, const TMapE *m, int c0, int c1, mbar_t *bar) { tma_load_3d(dst, m, c0, c1, 0, bar); }
struct float4 { float x, y, z, w; };

alignas(1024) static unsigned char g_smem[256 * 1024];

#define PST_SYS_DEV static inline
#define PST_SYS_GLOBAL(T, MINB) static
#define PST_SYS_TMAP_PARAM TMapE
#define PST_SYS_UNROLL
#define PST_SYS_SMEM(name) float *const name = reinterpret_cast<float *>(g_smem)

#include "pst_tri_sys_kernels.cuh"

using namespace tri_sys_k;

template <bool CONTIG, int NB, int SEG, bool ILS, bool PRE>
static void run(const Plan &P, const float *src, float *dst, int axis, int n1, int n2, int n3, int sm_count)
{
    TMapE tm{};
    tm.base = src;
    if (CONTIG) {
        tm.dim[0] = n1; tm.dim[1] = P.na; tm.dim[2] = 1;
        tm.stride[0] = 1; tm.stride[1] = n1; tm.stride[2] = 0;
        tm.box[0] = Layout<true, NB, SEG>::XW; tm.box[1] = 32; tm.box[2] = 1;
    } else {
        if (axis == 1) { tm.dim[0] = n1; tm.dim[1] = n2; tm.dim[2] = n3; }
        else { tm.dim[0] = (long)n1 * n2; tm.dim[1] = n3; tm.dim[2] = 1; }
        tm.stride[0] = 1; tm.stride[1] = P.d; tm.stride[2] = axis == 1 ? P.sb : P.d * (long)n3;
        tm.box[0] = 32; tm.box[1] = SEG + 2 * NB; tm.box[2] = 1;
    }
    const Args A = make_args(P, dst, nullptr);
    const int per_sm = (SEG == 68 && !CONTIG) ? 2 : 1;
    long grid = (long)sm_count * per_sm;
    if (grid > P.ntiles) grid = P.ntiles;
    gridDim = {(unsigned)grid, 1, 1};
    if (Layout<CONTIG, NB, SEG>::bytes > sizeof(g_smem)) abort();
    for (unsigned bx = 0; bx < (unsigned)grid; bx++) {
        memset(g_smem, 0xff, sizeof(g_smem));                   // NaN pattern: a read of unwritten shared memory shows
        pthread_barrier_init(&cta_barrier, nullptr, NTHREADS);
        pthread_barrier_t bars[NTHREADS / 32];
        for (auto &b : bars) pthread_barrier_init(&b, nullptr, 32);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < (unsigned)NTHREADS; t++)
            th.emplace_back([&, t]() {
                threadIdx = {t, 0, 0};
                blockIdx = {bx, 0, 0};
                warp_barrier = &bars[t / 32];
                tri_sys_kernel<CONTIG, NB, SEG, ILS, PRE>(tm, A);
            });
        for (auto &t : th) t.join();
        for (auto &b : bars) pthread_barrier_destroy(&b);
        pthread_barrier_destroy(&cta_barrier);
    }
}

template <int NB>
static void dispatch(const Plan &P, const float *src, float *dst, int axis, int n1, int n2, int n3, int sm_count, bool ils, bool pre)
{
#define GO(C, I, Q) do { if (P.SEG == 68) run<C, NB, 68, I, Q>(P, src, dst, axis, n1, n2, n3, sm_count); else run<C, NB, 132, I, Q>(P, src, dst, axis, n1, n2, n3, sm_count); } while (0)
    if (axis == 0) { if (pre) GO(true, false, true); else GO(true, false, false); }
    else if (ils) { if (pre) GO(false, true, true); else GO(false, true, false); }
    else { if (pre) GO(false, false, true); else GO(false, false, false); }
#undef GO
}

extern "C" int tri_sys_emul(const float *src, float *dst, int n1, int n2, int n3, int axis, int nb, int sm_count)
{
    const Plan P = make_plan(axis, n1, n2, n3, nb);
    if (!P.ok) return -1;
    const char *ej = getenv("PST_EMUL_JITTER");
    g_jitter = ej ? atoi(ej) : 0;
    const char *ei = getenv("PST_TRI_SYS_ILS"), *ep = getenv("PST_TRI_SYS_PRE");
    const bool ils = ei && ei[0] == '1', pre = ep && ep[0] == '1';
    switch (nb) {
#define CASE(N) case N: dispatch<N>(P, src, dst, axis, n1, n2, n3, sm_count, ils, pre); return 0;
        CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(10)
#undef CASE
    }
    return -3;
}

// the plan alone (for the tests: which shapes are eligible, and how they are cut)
extern "C" int tri_sys_plan(int n1, int n2, int n3, int axis, int nb, int *out)
{
    const Plan P = make_plan(axis, n1, n2, n3, nb);
    out[0] = P.SEG; out[1] = P.nseg; out[2] = P.D; out[3] = (int)P.ntiles;
    return P.ok ? 1 : 0;
}
