// Shared internals of libpst_b200: context, workspace arena, deterministic reductions.
// Compiled for sm_100a only, with -fmad=false: the reference's results depend on the
// order and (non-)contraction of float operations, so every
// kernel spells out reference-order arithmetic and the compiler may not fuse mul+add.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/pst_b200.h"

#define PST_MAXTAP 5          // order (nw) 1 or 2 -> 3 or 5 taps
#define PST_RED_SLOTS 8       // doubles per reduction record

void pst_set_error(const char *fmt, ...);

#define PST_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            pst_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,             \
                          cudaGetErrorString(e__));                                 \
            return PST_ECUDA;                                                       \
        }                                                                           \
    } while (0)

#define PST_TRY(call)                                                               \
    do {                                                                            \
        int r__ = (call);                                                           \
        if (r__ != PST_OK) return r__;                                              \
    } while (0)

struct pst_comm;   // multi-GPU communicator (pst_comm.cu)

// what the distributed axis-3 kernels need from the carry mailboxes (pst_comm.cu)
struct pst_mailbox_view {
    float *cf_in, *cb_in;            // my mailbox: forward carries from rank-1, backward carries from rank+1
    unsigned *ff_in, *fb_in;         // per-CTA flags raised by the producer (value = epoch)
    float *cf_out, *cb_out;          // neighbours' mailboxes (peer memory): rank+1's cf, rank-1's cb
    unsigned *ff_out, *fb_out;
    unsigned *err;                   // set when a wait times out
    unsigned epoch;
    // tile kernels: self-validating 8-byte {carry bits, epoch} pairs, one per line (no fence, no flag)
    uint2 *pf_in, *pb_in, *pf_out, *pb_out;
    // peer-memory halos of the axis-3 taps: "my current input is complete" flags (value = epoch).  hr_in_*: in my
    // mailbox, raised by the previous / next rank; hr_out_*: the slots I raise in theirs
    unsigned *hr_in_prev, *hr_in_next, *hr_out_prev, *hr_out_next;
};

struct pst_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // reduction plumbing: per-block partials -> one record of PST_RED_SLOTS doubles
    double *d_partial = nullptr;     // [max_blocks * PST_RED_SLOTS]
    double *d_planes = nullptr;      // [PST_RED_SLOTS][planes_cap]: per-plane sums of the canonical reductions
    int planes_cap = 0;
    double *d_red = nullptr;         // [64 * PST_RED_SLOTS] ring of records
    double *h_red = nullptr;         // pinned mirror
    int max_blocks = 0;
    // device-resident control block of the shaping CG (pst_dip.cu: CgCtl) + pinned mirror of its stop flag
    void *d_cgctl = nullptr;
    volatile int *h_cgstop = nullptr;
    // simple bump arena over one big device allocation, reset per call
    char *arena = nullptr;
    size_t arena_size = 0, arena_used = 0;
    pst_stats stats{};
    // per-launch profiling (pst_ctx_set_profile): pool of event pairs, resolved lazily
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;    // 2 events per timed launch
    std::vector<int> prof_cls;
    std::vector<double> prof_bytes, prof_flops;
    size_t prof_used = 0;
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;
    pst_comm *comm = nullptr;        // null for single-GPU contexts
    int rank = 0, nranks = 1;
    // host-pointer entry points: copy streams + plane-granular pipelining of the transfers under the kernels
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Pipe {
        bool on = false;
        std::vector<int> up_planes;            // planes [0, up_planes[k]) of every input volume are uploaded once ...
        std::vector<cudaEvent_t> up_events;    // ... up_events[k] has fired (recorded on s_in)
        float *h_out = nullptr;                // where finished output planes go (host), null = no emit
        size_t plane = 0;                      // floats per plane
        int waited = 0;                        // planes the compute stream already waits for
    } pipe;
};

// plane-granular transfer pipeline (pst_api.cu); all no-ops when c->pipe.on is false
int pst_pipe_wait_planes(pst_ctx *c, int zhi);                              // c->stream waits until planes < zhi are on the device
int pst_pipe_emit(pst_ctx *c, const float *d_src_plane0, int z0, int z1);  // planes [z0, z1) of the output are final: D2H them

// ---- per-launch profiling -----------------------------------------------------------------
void pst_prof_resolve(pst_ctx *c);                       // sync + accumulate pending pairs
struct KTimer {                                          // RAII around one kernel launch
    pst_ctx *c; size_t slot; bool on;
    KTimer(pst_ctx *ctx, int cls, double bytes = 0.0, double flops = 0.0) : c(ctx), slot(0), on(ctx->prof)
    {
        c->stats.kernel_launches++;
        if (!on) return;
        if (c->prof_used + 2 > c->prof_ev.size()) pst_prof_resolve(c);
        slot = c->prof_used;
        c->prof_used += 2;
        c->prof_cls[slot / 2] = cls;
        c->prof_bytes[slot / 2] = bytes;
        c->prof_flops[slot / 2] = flops;
        cudaEventRecord(c->prof_ev[slot], c->stream);
    }
    ~KTimer() { if (on) cudaEventRecord(c->prof_ev[slot + 1], c->stream); }
};

// wrap the launch statement(s) of ONE kernel: counts it and, when profiling, times it
#define PST_LAUNCH(c, cls, ...) do { KTimer kt__((c), (cls)); __VA_ARGS__; } while (0)
// same, with the algorithmic bytes the launch moves (for the live roofline)
#define PST_LAUNCHB(c, cls, bytes, ...) do { KTimer kt__((c), (cls), (double)(bytes)); __VA_ARGS__; } while (0)
// same, plus the algorithmic flops (ALU-bound kernels)
#define PST_LAUNCHBF(c, cls, bytes, flops, ...) do { KTimer kt__((c), (cls), (double)(bytes), (double)(flops)); __VA_ARGS__; } while (0)

// ---- arena ---------------------------------------------------------------------------
int pst_arena_reserve(pst_ctx *c, size_t bytes);            // (re)allocate if too small
int pst_arena_alloc(pst_ctx *c, size_t bytes, void **p);    // 256-byte aligned bump
inline void pst_arena_reset(pst_ctx *c) { c->arena_used = 0; }

template <typename T>
inline int pst_arena_get(pst_ctx *c, size_t count, T **p)
{
    void *v = nullptr;
    int rc = pst_arena_alloc(c, count * sizeof(T), &v);
    *p = (T *)v;
    return rc;
}

// ---- launch geometry -----------------------------------------------------------------
inline int pst_grid_for(const pst_ctx *c, size_t n, int threads, int per_thread = 4)
{
    size_t want = (n + (size_t)threads * per_thread - 1) / ((size_t)threads * per_thread);
    size_t cap = (size_t)c->sm_count * 8;          // a multiple of the SM count
    if (want < 1) want = 1;
    if (want > cap) want = cap;
    return (int)want;
}

// ---- deterministic block reduction of NV doubles --------------------------------------
// Lanes are combined in a fixed shuffle tree, warps in index order, blocks in index order by
// pst_finish_reduce: results are run-to-run reproducible (no atomics).
template <int NV>
__device__ __forceinline__ void pst_block_reduce(double (&v)[NV], double *partial_out)
{
    __shared__ double sh[NV][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; q++) {
        double x = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) sh[q][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
        const int nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int q = 0; q < NV; q++) {
            double x = lane < nwarp ? sh[q][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0) partial_out[(size_t)blockIdx.x * PST_RED_SLOTS + q] = x;
        }
    }
}

// ---- canonical (partition-independent) sums ---------------------------------------------------------------------
// A volume [n1 n2][nz] is summed as: elements of one PIECE (pst_red_ch() consecutive elements of ONE n3-plane; block =
// piece, fixed thread pattern + fixed trees) -> pieces of a plane in index order (pst_plane_sums_kernel) -> planes of
// the GLOBAL cube in a fixed order (pst_final_sums_kernel).  No step looks at how many planes THIS rank owns, so an
// n3-slab decomposition over any number of ranks forms every CG / divne / line-search scalar with exactly the
// additions of the single-GPU run: ranks fill their planes of the [nv][n3 global] table, the others stay +0, and the
// all-reduce of that table adds only zeros (exact in any order).
// piece length: never a function of the slab.  8192 elements (32 per thread): measured against 32768 at 1000x1024x1024,
// CG gp / direction / head 186 / 547 / 725 (grid-stride head) -> 181 / 531 / 657 ms per step, and a 64-plane slab of
// 500x512x512 still fills 148 SMs
inline unsigned pst_red_ch(size_t n12, int nzg)
{
    static const unsigned forced = []() { const char *e = getenv("PST_RED_CH"); const long v = e ? atol(e) : 0; return (v >= 1024 && v % 1024 == 0) ? (unsigned)v : 0u; }();
    if (forced) return forced;                 // A/B (set it on every rank alike)
    (void)n12; (void)nzg;
    return 8192u;
}
struct Span { size_t n; unsigned n12, ppp, ch; };  // ppp > 0: block = piece (blockIdx.x % ppp) of plane (blockIdx.x / ppp); 0: grid-stride over n
inline Span pst_span_canon(size_t n12, int nz, int nzg)
{
    Span S; S.n = n12 * (size_t)nz; S.n12 = (unsigned)n12; S.ch = pst_red_ch(n12, nzg); S.ppp = (unsigned)((n12 + S.ch - 1) / S.ch);
    return S;
}
#ifdef __CUDACC__
// this thread's elements: i0, i0 + step, ... < i1 (width = elements per thread and step: 1 or 4)
__device__ __forceinline__ void pst_span(const Span &S, int width, size_t &i0, size_t &i1, size_t &step)
{
    if (S.ppp) {
        const unsigned z = blockIdx.x / S.ppp, p = blockIdx.x - z * S.ppp;
        const unsigned o = p * S.ch;
        const unsigned len = (S.n12 - o < S.ch) ? S.n12 - o : S.ch;
        const size_t base = (size_t)z * S.n12 + o;
        i0 = base + (size_t)width * threadIdx.x; i1 = base + len; step = (size_t)width * blockDim.x;
    } else {
        i0 = (size_t)width * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); i1 = S.n; step = (size_t)width * gridDim.x * blockDim.x;
    }
}
#endif
// pieces [nz][ppp] of c->d_partial -> record `rec`; z0 / nzg place this rank's planes in the global cube
int pst_finish_reduce_canon(pst_ctx *c, int ppp, int nz, int z0, int nzg, int nv, int rec);
int pst_reserve_partials(pst_ctx *c, size_t nblocks, int nzg);

// sum the per-block partials into record `rec` (device), optionally fetch to host
int pst_finish_reduce(pst_ctx *c, int nblocks, int nv, int rec);
int pst_fetch_record(pst_ctx *c, int rec, int nv, double *host_out);   // syncs the stream
