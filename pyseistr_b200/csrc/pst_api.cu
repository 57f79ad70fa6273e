// C-ABI surface of libpst_b200 (include/pst_b200.h): context, arena, deterministic reduction
// plumbing and the host-pointer entry points that mirror the reference's CPython extension
// functions (copy in -> compute on the GPU -> copy out).  No CPU fallback anywhere: without a
// usable sm_100 device every call fails loudly.
#include "pst_common.cuh"

#include <stdarg.h>
#include <string.h>

static thread_local char g_err[1024] = "";

void pst_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *pst_last_error(void) { return g_err; }
extern "C" const char *pst_version(void) { return "pst_b200 0.1 (sm_100a)"; }

extern "C" int pst_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---- arena -----------------------------------------------------------------------------
int pst_arena_reserve(pst_ctx *c, size_t bytes)
{
    bytes += 1 << 20;
    if (c->arena && c->arena_size >= bytes) return PST_OK;
    if (c->arena) { cudaStreamSynchronize(c->stream); cudaFree(c->arena); c->arena = nullptr; c->arena_size = 0; }
    cudaError_t e = cudaMalloc((void **)&c->arena, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        pst_set_error("device workspace of %.2f GB unavailable: %s", bytes / 1e9, cudaGetErrorString(e));
        return PST_ENOMEM;
    }
    c->arena_size = bytes;
    c->arena_used = 0;
    return PST_OK;
}

int pst_arena_alloc(pst_ctx *c, size_t bytes, void **p)
{
    const size_t a = (c->arena_used + 255) & ~(size_t)255;
    if (a + bytes > c->arena_size) {
        pst_set_error("internal: workspace arena overflow (%zu + %zu > %zu)", a, bytes, c->arena_size);
        return PST_ENOMEM;
    }
    *p = c->arena + a;
    c->arena_used = a + bytes;
    return PST_OK;
}

// ---- reductions ------------------------------------------------------------------------
__global__ void finish_reduce_kernel(const double *__restrict__ partial, int nblocks, int nv,
                                     double *__restrict__ rec)
{
    // one warp per value; fixed strided order + fixed shuffle tree => reproducible
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q >= nv) return;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += partial[(size_t)b * PST_RED_SLOTS + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) rec[q] = s;
}

int pst_comm_allreduce_record(pst_ctx *c, double *d_rec, int nv);   // pst_comm.cu

int pst_finish_reduce(pst_ctx *c, int nblocks, int nv, int rec)
{
    if (nblocks > c->max_blocks) { pst_set_error("internal: reduction grid too large"); return PST_EINVAL; }
    {
        KTimer kt(c, PST_K_OTHER);
        finish_reduce_kernel<<<1, 32 * PST_RED_SLOTS, 0, c->stream>>>(c->d_partial, nblocks, nv,
                                                                       c->d_red + (size_t)rec * PST_RED_SLOTS);
    }
    PST_CUDA(cudaGetLastError());
    if (c->comm) PST_TRY(pst_comm_allreduce_record(c, c->d_red + (size_t)rec * PST_RED_SLOTS, nv));
    return PST_OK;
}

// ---- canonical sums (pst_common.cuh) ----
// one warp per (value, GLOBAL plane): the plane's pieces in lane-strided order + the fixed shuffle tree; +0 for planes
// of other ranks
__global__ void plane_sums_kernel(const double *__restrict__ partial, int ppp, int nz, int z0, int nzg, int nv,
                                  double *__restrict__ planes)
{
    const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (long)nzg * nv) return;
    const int q = (int)(w / nzg), zg = (int)(w - (long)q * nzg), z = zg - z0;
    double s = 0.0;
    if (z >= 0 && z < nz) {
        for (int p = lane; p < ppp; p += 32) s += partial[((size_t)z * ppp + p) * PST_RED_SLOTS + q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    }
    if (lane == 0) planes[(size_t)q * nzg + zg] = s;
}

// one block per value: planes in thread-strided order, lanes by the shuffle tree, warps in index order
__global__ void __launch_bounds__(256) final_sums_kernel(const double *__restrict__ planes, int nzg, double *__restrict__ rec)
{
    __shared__ double sh[8];
    const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double s = 0.0;
    for (int z = threadIdx.x; z < nzg; z += 256) s += planes[(size_t)q * nzg + z];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) sh[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = sh[0];
        for (int k = 1; k < 8; k++) t += sh[k];
        rec[q] = t;
    }
}

int pst_reserve_partials(pst_ctx *c, size_t nblocks, int nzg)
{
    if (nblocks > (size_t)c->max_blocks) {
        PST_CUDA(cudaStreamSynchronize(c->stream));
        if (c->d_partial) cudaFree(c->d_partial);
    if (c->d_planes) cudaFree(c->d_planes);
        c->d_partial = nullptr; c->max_blocks = 0;
        PST_CUDA(cudaMalloc((void **)&c->d_partial, nblocks * PST_RED_SLOTS * sizeof(double)));
        c->max_blocks = (int)nblocks;
    }
    if (nzg > c->planes_cap) {
        PST_CUDA(cudaStreamSynchronize(c->stream));
        if (c->d_planes) cudaFree(c->d_planes);
        c->d_planes = nullptr; c->planes_cap = 0;
        PST_CUDA(cudaMalloc((void **)&c->d_planes, (size_t)nzg * PST_RED_SLOTS * sizeof(double)));
        c->planes_cap = nzg;
    }
    return PST_OK;
}

int pst_finish_reduce_canon(pst_ctx *c, int ppp, int nz, int z0, int nzg, int nv, int rec)
{
    if ((size_t)ppp * nz > (size_t)c->max_blocks || nzg > c->planes_cap || nv > PST_RED_SLOTS) {
        pst_set_error("internal: canonical reduction without pst_reserve_partials");
        return PST_EINVAL;
    }
    {
        KTimer kt(c, PST_K_OTHER);
        const long warps = (long)nzg * nv;
        plane_sums_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, c->stream>>>(c->d_partial, ppp, nz, z0, nzg, nv, c->d_planes);
    }
    PST_CUDA(cudaGetLastError());
    if (c->comm) PST_TRY(pst_comm_allreduce_record(c, c->d_planes, nv * nzg));
    {
        KTimer kt(c, PST_K_OTHER);
        final_sums_kernel<<<nv, 256, 0, c->stream>>>(c->d_planes, nzg, c->d_red + (size_t)rec * PST_RED_SLOTS);
    }
    PST_CUDA(cudaGetLastError());
    return PST_OK;
}

int pst_fetch_record(pst_ctx *c, int rec, int nv, double *host_out)
{
    PST_CUDA(cudaMemcpyAsync(c->h_red + (size_t)rec * PST_RED_SLOTS, c->d_red + (size_t)rec * PST_RED_SLOTS,
                             nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    for (int q = 0; q < nv; q++) host_out[q] = c->h_red[(size_t)rec * PST_RED_SLOTS + q];
    return PST_OK;
}

// ---- context ---------------------------------------------------------------------------
static int ctx_init(pst_ctx *c, int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        pst_set_error("no CUDA device visible: libpst_b200 has no CPU fallback");
        return PST_ENODEV;
    }
    if (device < 0 || device >= ndev) { pst_set_error("device %d out of range (0..%d)", device, ndev - 1); return PST_ENODEV; }
    cudaDeviceProp prop;
    PST_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        pst_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return PST_ENODEV;
    }
    PST_CUDA(cudaSetDevice(device));
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    PST_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PST_CUDA(cudaEventCreate(&c->ev0));
    PST_CUDA(cudaEventCreate(&c->ev1));
    PST_CUDA(cudaEventCreate(&c->tm0));
    PST_CUDA(cudaEventCreate(&c->tm1));
    c->max_blocks = c->sm_count * 32;
    PST_CUDA(cudaMalloc((void **)&c->d_partial, (size_t)c->max_blocks * PST_RED_SLOTS * sizeof(double)));
    PST_CUDA(cudaMalloc((void **)&c->d_red, 64 * PST_RED_SLOTS * sizeof(double)));
    PST_CUDA(cudaMallocHost((void **)&c->h_red, 64 * PST_RED_SLOTS * sizeof(double)));
    PST_CUDA(cudaMalloc(&c->d_cgctl, 256));
    PST_CUDA(cudaMallocHost((void **)&c->h_cgstop, 64));
    return PST_OK;
}

void pst_comm_destroy(pst_ctx *c);   // pst_comm.cu

extern "C" void pst_ctx_destroy(pst_ctx *c);

extern "C" int pst_ctx_create(int device, pst_ctx **out)
{
    if (!out) { pst_set_error("null out pointer"); return PST_EINVAL; }
    *out = nullptr;
    pst_ctx *c = new pst_ctx();
    int rc = ctx_init(c, device);
    if (rc != PST_OK) { pst_ctx_destroy(c); return rc; }      // releases whatever ctx_init got as far as creating
    *out = c;
    return PST_OK;
}

extern "C" void pst_ctx_destroy(pst_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm) pst_comm_destroy(c);
    if (c->arena) cudaFree(c->arena);
    if (c->d_partial) cudaFree(c->d_partial);
    if (c->d_red) cudaFree(c->d_red);
    if (c->h_red) cudaFreeHost(c->h_red);
    if (c->d_cgctl) cudaFree(c->d_cgctl);
    if (c->h_cgstop) cudaFreeHost((void *)c->h_cgstop);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->tm0) cudaEventDestroy(c->tm0);
    if (c->tm1) cudaEventDestroy(c->tm1);
    for (auto &e : c->prof_ev) cudaEventDestroy(e);
    for (auto &e : c->ev_pool) cudaEventDestroy(e);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int pst_ctx_stats(pst_ctx *c, pst_stats *out)
{
    if (!c || !out) { pst_set_error("null argument"); return PST_EINVAL; }
    cudaSetDevice(c->device);
    pst_prof_resolve(c);
    *out = c->stats;
    return PST_OK;
}

extern "C" int pst_ctx_reset_stats(pst_ctx *c)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    cudaSetDevice(c->device);
    pst_prof_resolve(c);
    memset(&c->stats, 0, sizeof(c->stats));
    return PST_OK;
}

void pst_prof_resolve(pst_ctx *c)
{
    if (c->prof_used == 0) return;
    cudaStreamSynchronize(c->stream);
    for (size_t s = 0; s < c->prof_used; s += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->prof_ev[s], c->prof_ev[s + 1]) == cudaSuccess) {
            const int k = c->prof_cls[s / 2];
            c->stats.class_ms[k] += ms;
            c->stats.class_launches[k]++;
            c->stats.class_bytes[k] += c->prof_bytes[s / 2];
            c->stats.class_flops[k] += c->prof_flops[s / 2];
        }
    }
    c->prof_used = 0;
}

extern "C" int pst_ctx_set_profile(pst_ctx *c, int on)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    if (on && c->prof_ev.empty()) {
        const size_t pairs = 4096;
        c->prof_ev.resize(2 * pairs);
        c->prof_cls.resize(pairs);
        c->prof_bytes.resize(pairs);
        c->prof_flops.resize(pairs);
        for (auto &e : c->prof_ev) PST_CUDA(cudaEventCreate(&e));
    }
    if (!on) pst_prof_resolve(c);
    c->prof = on != 0;
    return PST_OK;
}

extern "C" int pst_timer_start(pst_ctx *c)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    PST_CUDA(cudaEventRecord(c->tm0, c->stream));
    return PST_OK;
}

extern "C" int pst_timer_stop(pst_ctx *c, double *elapsed_ms)
{
    if (!c || !elapsed_ms) { pst_set_error("null argument"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    PST_CUDA(cudaEventRecord(c->tm1, c->stream));
    PST_CUDA(cudaEventSynchronize(c->tm1));
    float ms = 0.f;
    PST_CUDA(cudaEventElapsedTime(&ms, c->tm0, c->tm1));
    *elapsed_ms = ms;
    return PST_OK;
}

// ---- raw memory helpers ------------------------------------------------------------------
extern "C" int pst_dev_alloc(pst_ctx *c, size_t bytes, void **d_ptr)
{
    if (!c || !d_ptr) { pst_set_error("null argument"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); pst_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return PST_ENOMEM; }
    return PST_OK;
}
extern "C" int pst_dev_free(pst_ctx *c, void *d_ptr)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    PST_CUDA(cudaFree(d_ptr));
    return PST_OK;
}
extern "C" int pst_h2d(pst_ctx *c, void *d_dst, const void *h_src, size_t bytes)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    PST_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c->stream));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    return PST_OK;
}
extern "C" int pst_d2h(pst_ctx *c, void *h_dst, const void *d_src, size_t bytes)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    PST_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    return PST_OK;
}
extern "C" int pst_host_alloc_pinned(size_t bytes, void **h_ptr)
{
    if (!h_ptr) { pst_set_error("null argument"); return PST_EINVAL; }
    cudaError_t e = cudaMallocHost(h_ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); pst_set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); return PST_ENOMEM; }
    return PST_OK;
}
extern "C" int pst_host_free_pinned(void *h_ptr)
{
    PST_CUDA(cudaFreeHost(h_ptr));
    return PST_OK;
}
extern "C" int pst_sync(pst_ctx *c)
{
    if (!c) { pst_set_error("null context"); return PST_EINVAL; }
    PST_CUDA(cudaSetDevice(c->device));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    return PST_OK;
}

// ---- host-pointer entry points -----------------------------------------------------------
// Scoped device buffers for one call (outside the arena, which the *_dev calls reset).
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes)
    {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; pst_set_error("cudaMalloc(%.2f GB) failed: %s", bytes / 1e9, cudaGetErrorString(e)); return PST_ENOMEM; }
        return PST_OK;
    }
    float *f() { return (float *)p; }
};

struct CallTimer {
    pst_ctx *c;
    explicit CallTimer(pst_ctx *ctx) : c(ctx)
    {
        pst_prof_resolve(c);
        memset(&c->stats, 0, sizeof(c->stats));
        cudaEventRecord(c->ev0, c->stream);
    }
    void stop()
    {
        cudaEventRecord(c->ev1, c->stream);
        cudaEventSynchronize(c->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        c->stats.device_ms = ms;
    }
};

// ---- transfer pipeline of the host-pointer entry points ------------------------------------------------------
// Uploads run on s_in in plane chunks (an event per chunk), downloads on s_out; the compute stream waits only for
// the planes a kernel is about to read (pst_pipe_wait_planes) and hands finished output planes to s_out
// (pst_pipe_emit), so PCIe runs under the kernels in both directions.
static int pipe_streams(pst_ctx *c)
{
    if (!c->s_in) PST_CUDA(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    if (!c->s_out) PST_CUDA(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    return PST_OK;
}
static int pipe_event(pst_ctx *c, cudaEvent_t *e)
{
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t ne;
        PST_CUDA(cudaEventCreateWithFlags(&ne, cudaEventDisableTiming));
        c->ev_pool.push_back(ne);
    }
    *e = c->ev_pool[c->ev_used++];
    return PST_OK;
}
// upload nvol volumes of nz planes each, interleaved in chunks of `cz` planes; events land in c->pipe
static int pipe_upload(pst_ctx *c, int nvol, float *const *dst, const float *const *src, size_t plane, int nz, int cz)
{
    PST_TRY(pipe_streams(c));
    c->ev_used = 0;
    c->pipe = pst_ctx::Pipe();
    c->pipe.plane = plane;
    // the staging buffers were just allocated: nothing of an earlier call may still be reading them
    for (int z = 0; z < nz; z += cz) {
        const int ze = std::min(nz, z + cz);
        for (int v = 0; v < nvol; v++)
            PST_CUDA(cudaMemcpyAsync(dst[v] + (size_t)z * plane, src[v] + (size_t)z * plane, (size_t)(ze - z) * plane * sizeof(float),
                                     cudaMemcpyHostToDevice, c->s_in));
        cudaEvent_t e;
        PST_TRY(pipe_event(c, &e));
        PST_CUDA(cudaEventRecord(e, c->s_in));
        c->pipe.up_planes.push_back(ze);
        c->pipe.up_events.push_back(e);
        c->stats.h2d_bytes += (double)nvol * (ze - z) * plane * sizeof(float);
    }
    c->pipe.on = true;
    return PST_OK;
}
int pst_pipe_wait_planes(pst_ctx *c, int zhi)
{
    pst_ctx::Pipe &P = c->pipe;
    if (!P.on || zhi <= P.waited) return PST_OK;
    for (size_t k = 0; k < P.up_planes.size(); k++) {
        if (P.up_planes[k] <= P.waited) continue;
        PST_CUDA(cudaStreamWaitEvent(c->stream, P.up_events[k], 0));
        P.waited = P.up_planes[k];
        if (P.waited >= zhi) break;
    }
    return PST_OK;
}
int pst_pipe_emit(pst_ctx *c, const float *d_src_plane0, int z0, int z1)
{
    pst_ctx::Pipe &P = c->pipe;
    if (!P.on || !P.h_out || z1 <= z0) return PST_OK;
    cudaEvent_t e;
    PST_TRY(pipe_event(c, &e));
    PST_CUDA(cudaEventRecord(e, c->stream));
    PST_CUDA(cudaStreamWaitEvent(c->s_out, e, 0));
    PST_CUDA(cudaMemcpyAsync(P.h_out + (size_t)z0 * P.plane, d_src_plane0 + (size_t)z0 * P.plane, (size_t)(z1 - z0) * P.plane * sizeof(float),
                             cudaMemcpyDeviceToHost, c->s_out));
    c->stats.d2h_bytes += (double)(z1 - z0) * P.plane * sizeof(float);
    return PST_OK;
}
// end of a pipelined call: everything emitted has landed; the pipe is switched off
static int pipe_finish(pst_ctx *c)
{
    c->pipe.on = false;
    PST_CUDA(cudaStreamSynchronize(c->stream));
    if (c->s_in) PST_CUDA(cudaStreamSynchronize(c->s_in));
    if (c->s_out) PST_CUDA(cudaStreamSynchronize(c->s_out));
    return PST_OK;
}
struct PipeGuard {           // an error return must not leave a stale pipe behind
    pst_ctx *c;
    ~PipeGuard() { if (c->pipe.on) { c->pipe.on = false; cudaStreamSynchronize(c->stream); if (c->s_in) cudaStreamSynchronize(c->s_in); if (c->s_out) cudaStreamSynchronize(c->s_out); } }
};

static int up(pst_ctx *c, DevBuf &b, const float *h, size_t n)
{
    PST_TRY(b.alloc(n * sizeof(float)));
    PST_CUDA(cudaMemcpyAsync(b.p, h, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    c->stats.h2d_bytes += (double)n * sizeof(float);
    return PST_OK;
}
static int down(pst_ctx *c, float *h, const DevBuf &b, size_t n)
{
    PST_CUDA(cudaMemcpyAsync(h, b.p, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    c->stats.d2h_bytes += (double)n * sizeof(float);
    return PST_OK;
}

// planes held by this rank: all of them on a single-GPU context
static int slab_planes(pst_ctx *c, int n3)
{
    if (c->nranks <= 1) return n3;
    return (int)(((long)n3 * (c->rank + 1)) / c->nranks) - (int)(((long)n3 * c->rank) / c->nranks);
}

#define PST_ENTRY(c)                                                        \
    if (!(c)) { pst_set_error("null context"); return PST_EINVAL; }         \
    PST_CUDA(cudaSetDevice((c)->device));

extern "C" int pst_dip(pst_ctx *c, const float *din, const float *mask, int n1, int n2, int n3,
                       int niter, int liter, int order, float eps_dv, float eps_cg, float tol_cg,
                       int r1, int r2, int r3, int verb, float *dip_out)
{
    (void)eps_dv; (void)eps_cg; (void)tol_cg;      // ignored by the reference's C
    PST_ENTRY(c);
    if (!din || !dip_out || n1 < 1 || n2 < 1 || n3 < 1) { pst_set_error("dip: null pointer or bad shape"); return PST_EINVAL; }
    const int nz = slab_planes(c, n3);
    const size_t plane = (size_t)n1 * n2, n = plane * nz;      /* n = this rank's slab */
    CallTimer t(c);
    DevBuf d, m, o;
    PST_TRY(d.alloc(n * sizeof(float)));
    if (mask) PST_TRY(m.alloc(n * sizeof(float)));
    PST_TRY(o.alloc((n3 == 1 ? n : 2 * n) * sizeof(float)));
    // transfers on the copy streams: the inputs go up while the workspace is set up; the inline dip goes down while the
    // xline dip is being estimated (pst_dip_dev hands it over through the pipe), the xline dip after the call
    PipeGuard guard{c};
    float *dst[2] = {d.f(), m.f()};
    const float *src[2] = {din, mask};
    PST_TRY(pipe_upload(c, mask ? 2 : 1, dst, src, plane, nz, nz));
    c->pipe.h_out = dip_out;
    PST_TRY(pst_dip_dev(c, d.f(), mask ? m.f() : nullptr, n1, n2, n3, niter, liter, order, r1, r2, r3, verb, o.f()));
    if (n3 == 1) PST_TRY(pst_pipe_emit(c, o.f(), 0, nz));
    else         PST_TRY(pst_pipe_emit(c, o.f(), nz, 2 * nz));
    PST_TRY(pipe_finish(c));
    t.stop();
    return PST_OK;
}

// csomean3d / csomf3d on host buffers: the three inputs go up in plane chunks on the copy stream while the spray works
// on the chunks that have arrived, finished output planes go down on a second copy stream under the next chunk
static int spray3d_host(pst_ctx *c, int median, const float *din, const float *dipi, const float *dipx, int n1, int n2, int n3,
                        int ns2, int ns3, int nmf, int option, int order, float *out)
{
    const int nz = slab_planes(c, n3);
    const size_t plane = (size_t)n1 * n2, n = plane * nz;
    CallTimer t(c);
    DevBuf own[3], o;
    float *dev[3], *dst[3];
    const float *src[3] = {din, dipi, dipx};
    const int nup = 3;
    for (int v = 0; v < 3; v++) {
        PST_TRY(own[v].alloc(n * sizeof(float)));
        dev[v] = dst[v] = own[v].f();
    }
    PST_TRY(o.alloc(n * sizeof(float)));
    PipeGuard guard{c};
    // ~0.4 GB per volume and chunk: fine enough to start early, coarse enough for full PCIe speed
    const int cz = (int)std::max<size_t>(1, std::min<size_t>((size_t)nz, (size_t)100e6 / plane + 1));
    PST_TRY(pipe_upload(c, nup, dst, src, plane, nz, cz));
    c->pipe.h_out = out;
    int rc = median ? pst_somf3d_dev(c, dev[0], dev[1], dev[2], n1, n2, n3, ns2, ns3, nmf, option, order, o.f())
                    : pst_somean3d_dev(c, dev[0], dev[1], dev[2], n1, n2, n3, ns2, ns3, order, o.f());
    if (rc != PST_OK) return rc;
    PST_TRY(pipe_finish(c));
    t.stop();
    return PST_OK;
}

extern "C" int pst_somean3d(pst_ctx *c, const float *din, const float *dipi, const float *dipx,
                            int n1, int n2, int n3, int ns2, int ns3, int order, float eps, int verb,
                            float *out)
{
    (void)eps; (void)verb;                          // eps overridden with 0.01 by the reference (Q2)
    PST_ENTRY(c);
    if (!din || !dipi || !dipx || !out || n1 < 1 || n2 < 1 || n3 < 1) { pst_set_error("somean3d: null pointer or bad shape"); return PST_EINVAL; }
    return spray3d_host(c, 0, din, dipi, dipx, n1, n2, n3, ns2, ns3, 0, 1, order, out);
}

extern "C" int pst_somf3d(pst_ctx *c, const float *din, const float *dipi, const float *dipx,
                          int n1, int n2, int n3, int ns2, int ns3, int nmf, int option, int order,
                          float eps, int verb, float *out)
{
    (void)eps; (void)verb;
    PST_ENTRY(c);
    if (!din || !dipi || !dipx || !out || n1 < 1 || n2 < 1 || n3 < 1) { pst_set_error("somf3d: null pointer or bad shape"); return PST_EINVAL; }
    return spray3d_host(c, 1, din, dipi, dipx, n1, n2, n3, ns2, ns3, nmf, option, order, out);
}

int pst_somean2d_dev(pst_ctx *c, const float *d_din, const float *d_dip, int n1, int n2, int n3, int ns,
                     int order, float eps, float *d_out);
int pst_somean2d_adj_dev(pst_ctx *c, const float *d_din, const float *d_dip, int n1, int n2, int n3, int ns,
                         int order, float eps, float *d_out);
int pst_somf2d_dev(pst_ctx *c, const float *d_din, const float *d_dip, int n1, int n2, int n3, int ns,
                   int nmf, int option, int order, float eps, float *d_out);

extern "C" int pst_somean2d(pst_ctx *c, const float *din, const float *dip, int n1, int n2, int n3,
                            int ns, int order, int adj, float eps, int verb, float *out)
{
    (void)verb;
    PST_ENTRY(c);
    if (!din || !dip || !out || n1 < 1 || n2 < 1 || n3 < 1) { pst_set_error("somean2d: null pointer or bad shape"); return PST_EINVAL; }

    const size_t n = (size_t)n1 * n2 * slab_planes(c, n3);
    CallTimer t(c);
    DevBuf d, a, o;
    PST_TRY(up(c, d, din, n)); PST_TRY(up(c, a, dip, n));
    PST_TRY(o.alloc(n * sizeof(float)));
    if (adj) PST_TRY(pst_somean2d_adj_dev(c, d.f(), a.f(), n1, n2, n3, ns, order, eps, o.f()));
    else PST_TRY(pst_somean2d_dev(c, d.f(), a.f(), n1, n2, n3, ns, order, eps, o.f()));
    PST_TRY(down(c, out, o, n));
    t.stop();
    return PST_OK;
}

extern "C" int pst_somf2d(pst_ctx *c, const float *din, const float *dip, int n1, int n2, int n3,
                          int ns, int nmf, int option, int order, float eps, int verb, float *out)
{
    (void)verb;
    PST_ENTRY(c);
    if (!din || !dip || !out || n1 < 1 || n2 < 1 || n3 < 1) { pst_set_error("somf2d: null pointer or bad shape"); return PST_EINVAL; }
    const size_t n = (size_t)n1 * n2 * slab_planes(c, n3);
    CallTimer t(c);
    DevBuf d, a, o;
    PST_TRY(up(c, d, din, n)); PST_TRY(up(c, a, dip, n));
    PST_TRY(o.alloc(n * sizeof(float)));
    PST_TRY(pst_somf2d_dev(c, d.f(), a.f(), n1, n2, n3, ns, nmf, option, order, eps, o.f()));
    PST_TRY(down(c, out, o, n));
    t.stop();
    return PST_OK;
}

extern "C" int pst_soint3d(pst_ctx *c, const float *din, const float *mask, const float *dipi, const float *dipx,
                           int n1, int n2, int n3, int nw, int nj1, int nj2, int niter, int drift, int seed,
                           int hasmask, float var, int verb, float *out)
{
    PST_ENTRY(c);
    if (!din || !dipi || !dipx || !out || n1 < 1 || n2 < 1 || n3 < 1) { pst_set_error("soint3d: null pointer or bad shape"); return PST_EINVAL; }
    if (hasmask && !mask) { pst_set_error("soint3d: hasmask=1 needs a mask"); return PST_EINVAL; }
    const size_t n = (size_t)n1 * n2 * slab_planes(c, n3);      /* this rank's slab in a distributed context */
    CallTimer t(c);
    DevBuf d, m, a, b, o;
    PST_TRY(up(c, d, din, n));
    if (hasmask) PST_TRY(up(c, m, mask, n));
    PST_TRY(up(c, a, dipi, n)); PST_TRY(up(c, b, dipx, n));
    PST_TRY(o.alloc(n * sizeof(float)));
    PST_TRY(pst_soint3d_dev(c, d.f(), hasmask ? m.f() : nullptr, a.f(), b.f(), n1, n2, n3, nw, nj1, nj2, niter, drift, seed,
                            hasmask, var, verb, o.f()));
    PST_TRY(down(c, out, o, n));
    t.stop();
    return PST_OK;
}

extern "C" int pst_smooth3(pst_ctx *c, const float *x, int n1, int n2, int n3, int r1, int r2, int r3,
                           int repeat, int adj, float *out)
{
    PST_ENTRY(c);
    if (!x || !out || n1 < 1 || n2 < 1 || n3 < 1) { pst_set_error("smooth3: null pointer or bad shape"); return PST_EINVAL; }
    const size_t n = (size_t)n1 * n2 * slab_planes(c, n3);
    CallTimer t(c);
    DevBuf d;
    PST_TRY(up(c, d, x, n));
    PST_TRY(pst_smooth3_dev(c, d.f(), n1, n2, n3, r1, r2, r3, repeat, adj));
    PST_TRY(down(c, out, d, n));
    t.stop();
    return PST_OK;
}

extern "C" int pst_smoothcf(pst_ctx *c, const float *x, int n1, int n2, int n3, int repeat, int adj, int r1, int r2, int r3,
                            int diff1, int diff2, int diff3, int box1, int box2, int box3, float *out)
{
    PST_ENTRY(c);
    if (!x || !out || n1 < 1 || n2 < 1 || n3 < 1) { pst_set_error("smoothcf: null pointer or bad shape"); return PST_EINVAL; }
    const size_t n = (size_t)n1 * n2 * slab_planes(c, n3);
    CallTimer t(c);
    DevBuf d;
    PST_TRY(up(c, d, x, n));
    PST_TRY(pst_smoothcf_dev(c, d.f(), n1, n2, n3, repeat, adj, r1, r2, r3, diff1, diff2, diff3, box1, box2, box3));
    PST_TRY(down(c, out, d, n));
    t.stop();
    return PST_OK;
}

extern "C" int pst_paint2d(pst_ctx *c, const float *dip, const float *seed, int n1, int n2, int order, int i0, float eps,
                           int verb, float *out)
{
    (void)verb;
    PST_ENTRY(c);
    if (!dip || !seed || !out || n1 < 1 || n2 < 1) { pst_set_error("paint2d: null pointer or bad shape"); return PST_EINVAL; }
    if (c->nranks > 1) { pst_set_error("paint2d: single-GPU contexts only (one dependent chain of trace predictions)"); return PST_EUNSUP; }
    const size_t n = (size_t)n1 * n2;
    CallTimer t(c);
    DevBuf d, s, o;
    PST_TRY(up(c, d, dip, n));
    PST_TRY(up(c, s, seed, (size_t)n1));
    PST_TRY(o.alloc(n * sizeof(float)));
    PST_TRY(pst_paint2d_dev(c, d.f(), s.f(), n1, n2, order, i0, eps, o.f()));
    PST_TRY(down(c, out, o, n));
    t.stop();
    return PST_OK;
}

extern "C" int pst_sint3d(pst_ctx *c, const float *din, const float *dipi, const float *dipx, const float *mask,
                          int n1, int n2, int n3, int niter, int ns1, int ns2, int order1, int order2, int verb,
                          float eps, float *out)
{
    PST_ENTRY(c);
    if (!din || !dipi || !dipx || !mask || !out || n1 < 1 || n2 < 1 || n3 < 1) { pst_set_error("sint3d: null pointer or bad shape"); return PST_EINVAL; }
    const size_t n = (size_t)n1 * n2 * slab_planes(c, n3);      /* this rank's slab in a distributed context */
    CallTimer t(c);
    DevBuf d, m, a, b, o;
    PST_TRY(up(c, d, din, n));
    PST_TRY(up(c, m, mask, n));
    PST_TRY(up(c, a, dipi, n)); PST_TRY(up(c, b, dipx, n));
    PST_TRY(o.alloc(n * sizeof(float)));
    PST_TRY(pst_sint3d_dev(c, d.f(), a.f(), b.f(), m.f(), n1, n2, n3, niter, ns1, ns2, order1, order2, verb, eps, o.f()));
    PST_TRY(down(c, out, o, n));
    t.stop();
    return PST_OK;
}
