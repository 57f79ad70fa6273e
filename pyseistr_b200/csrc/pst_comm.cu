// Multi-GPU plumbing: n3-slab decomposition, one process per GPU, NCCL over NVLink.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"; if the host process already loaded one —
// e.g. torch's bundled copy — that instance is reused), so libpst_b200.so has no link-time
// dependency and single-GPU use never touches it.  Only the few entry points used here are
// declared; their ABI is stable across NCCL 2.x.
//
// What crosses slabs:
//   * CG / divne / line-search scalars: all-reduce of <= 8 doubles (pst_comm_allreduce_record);
//   * plane halos (xline stencil, spray inputs, axis-3 smoothing taps): pst_comm_halo_*;
//   * axis-3 running sums: a carry plane handed rank to rank (pst_comm_send / pst_comm_recv).
#include "pst_common.cuh"

#include <dlfcn.h>
#include <string.h>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8 };   // ncclDataType_t values (nccl.h)
enum { ncclSum = 0 };                          // ncclRedOp_t

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.handle) return PST_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // reuse the host process's copy
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { pst_set_error("cannot load libnccl.so.2: %s", dlerror()); return PST_ECOMM; }
#define PST_SYM(field, name)                                                            \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                         \
    if (!g_nccl.field) { pst_set_error("libnccl: missing symbol %s", name); return PST_ECOMM; }
    PST_SYM(GetUniqueId, "ncclGetUniqueId");
    PST_SYM(CommInitRank, "ncclCommInitRank");
    PST_SYM(CommDestroy, "ncclCommDestroy");
    PST_SYM(AllReduce, "ncclAllReduce");
    PST_SYM(Send, "ncclSend");
    PST_SYM(Recv, "ncclRecv");
    PST_SYM(AllGather, "ncclAllGather");
    PST_SYM(GroupStart, "ncclGroupStart");
    PST_SYM(GroupEnd, "ncclGroupEnd");
    PST_SYM(GetErrorString, "ncclGetErrorString");
#undef PST_SYM
    g_nccl.handle = h;
    return PST_OK;
}

#define PST_NCCL(call)                                                                  \
    do {                                                                                \
        ncclResult_t r__ = (call);                                                      \
        if (r__ != 0) {                                                                 \
            pst_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
            return PST_ECOMM;                                                           \
        }                                                                               \
    } while (0)

// Peer-memory mailbox of the axis-3 carry pipeline: every rank owns one
// device buffer [carry_fwd[L] | carry_bwd[L] | flag_fwd[nblk] | flag_bwd[nblk] | err], exported with
// CUDA IPC; rank r writes its outgoing carries straight into the neighbour's buffer over NVLink and
// then raises the neighbour's per-CTA flag (release at system scope).
struct pst_comm {
    ncclComm_t comm = nullptr;
    char *mbox = nullptr;            // my mailbox (device)
    char *mbox_prev = nullptr;       // rank-1's mailbox, mapped (null on rank 0)
    char *mbox_next = nullptr;       // rank+1's mailbox, mapped (null on the last rank)
    size_t mbox_L = 0, mbox_bytes = 0;
    unsigned epoch = 0;
    // neighbours' workspace arenas mapped with CUDA IPC (peer-memory halos of the axis-3 smoothing taps)
    char *arena_prev = nullptr, *arena_next = nullptr;
    char *arena_mapped_for = nullptr;      // my arena pointer at the time of the exchange
    size_t arena_mapped_size = 0;
};

int pst_comm_allreduce_record(pst_ctx *c, double *d_rec, int nv)
{
    if (!c->comm) return PST_OK;
    PST_NCCL(g_nccl.AllReduce(d_rec, d_rec, (size_t)nv, ncclFloat64, ncclSum, c->comm->comm, c->stream));
    return PST_OK;
}

// point-to-point on the context's stream
int pst_comm_send(pst_ctx *c, const float *d_buf, size_t count, int peer)
{
    PST_NCCL(g_nccl.Send(d_buf, count, ncclFloat32, peer, c->comm->comm, c->stream));
    return PST_OK;
}
int pst_comm_recv(pst_ctx *c, float *d_buf, size_t count, int peer)
{
    PST_NCCL(g_nccl.Recv(d_buf, count, ncclFloat32, peer, c->comm->comm, c->stream));
    return PST_OK;
}

// one NCCL group with an optional send and an optional receive (null pointer = skip)
int pst_comm_sendrecv(pst_ctx *c, const float *send, size_t nsend, int peer_out, float *recv, size_t nrecv, int peer_in)
{
    if ((!send || nsend == 0) && (!recv || nrecv == 0)) return PST_OK;
    PST_NCCL(g_nccl.GroupStart());
    if (send && nsend) PST_NCCL(g_nccl.Send(send, nsend, ncclFloat32, peer_out, c->comm->comm, c->stream));
    if (recv && nrecv) PST_NCCL(g_nccl.Recv(recv, nrecv, ncclFloat32, peer_in, c->comm->comm, c->stream));
    PST_NCCL(g_nccl.GroupEnd());
    return PST_OK;
}

// Exchange plane halos with both neighbours in one NCCL group:
//   send_lo (count floats) -> rank-1,   recv_lo <- rank-1   (their send_hi)
//   send_hi               -> rank+1,   recv_hi <- rank+1   (their send_lo)
// Edge ranks skip the missing side (their recv buffer is left untouched).
int pst_comm_halo_exchange(pst_ctx *c, const float *send_lo, const float *send_hi, float *recv_lo,
                           float *recv_hi, size_t count)
{
    if (!c->comm || count == 0) return PST_OK;
    PST_NCCL(g_nccl.GroupStart());
    if (c->rank > 0) {
        PST_NCCL(g_nccl.Send(send_lo, count, ncclFloat32, c->rank - 1, c->comm->comm, c->stream));
        PST_NCCL(g_nccl.Recv(recv_lo, count, ncclFloat32, c->rank - 1, c->comm->comm, c->stream));
    }
    if (c->rank < c->nranks - 1) {
        PST_NCCL(g_nccl.Send(send_hi, count, ncclFloat32, c->rank + 1, c->comm->comm, c->stream));
        PST_NCCL(g_nccl.Recv(recv_hi, count, ncclFloat32, c->rank + 1, c->comm->comm, c->stream));
    }
    PST_NCCL(g_nccl.GroupEnd());
    return PST_OK;
}

void pst_comm_destroy(pst_ctx *c)
{
    if (c->comm) {
        if (c->comm->arena_prev) cudaIpcCloseMemHandle(c->comm->arena_prev);
        if (c->comm->arena_next) cudaIpcCloseMemHandle(c->comm->arena_next);
        if (c->comm->mbox_prev) cudaIpcCloseMemHandle(c->comm->mbox_prev);
        if (c->comm->mbox_next) cudaIpcCloseMemHandle(c->comm->mbox_next);
        if (c->comm->mbox) cudaFree(c->comm->mbox);
        if (c->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm->comm);
        delete c->comm;
    }
    c->comm = nullptr;
}

extern "C" int pst_comm_unique_id(void *id128)
{
    if (!id128) { pst_set_error("null argument"); return PST_EINVAL; }
    PST_TRY(nccl_load());
    ncclUniqueId id;
    PST_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return PST_OK;
}

extern "C" int pst_ctx_create_dist(int device, int rank, int nranks, const void *nccl_id128, pst_ctx **ctx)
{
    if (!ctx) { pst_set_error("null out pointer"); return PST_EINVAL; }
    if (nranks < 1 || rank < 0 || rank >= nranks) { pst_set_error("bad rank %d / nranks %d", rank, nranks); return PST_EINVAL; }
    PST_TRY(pst_ctx_create(device, ctx));
    if (nranks == 1) return PST_OK;
    pst_ctx *c = *ctx;
    if (!nccl_id128) { pst_ctx_destroy(c); *ctx = nullptr; pst_set_error("null NCCL id"); return PST_EINVAL; }
    int rc = nccl_load();
    if (rc != PST_OK) { pst_ctx_destroy(c); *ctx = nullptr; return rc; }
    ncclUniqueId id;
    memcpy(&id, nccl_id128, 128);
    c->comm = new pst_comm();
    c->rank = rank;
    c->nranks = nranks;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm->comm, nranks, id, rank);
    if (r != 0) {
        pst_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c->comm; c->comm = nullptr;
        pst_ctx_destroy(c); *ctx = nullptr;
        return PST_ECOMM;
    }
    return PST_OK;
}

// The library's fixed slab rule: rank r owns global planes [n3*r/G, n3*(r+1)/G).
extern "C" int pst_ctx_slab(pst_ctx *c, int n3, int *z0, int *z1)
{
    if (!c || !z0 || !z1) { pst_set_error("null argument"); return PST_EINVAL; }
    *z0 = (int)(((long)n3 * c->rank) / c->nranks);
    *z1 = (int)(((long)n3 * (c->rank + 1)) / c->nranks);
    return PST_OK;
}

// ---- carry mailboxes ---------------------------------------------------------------------
static size_t mbox_layout(size_t L, size_t *off_cb, size_t *off_ff, size_t *off_fb, size_t *off_err,
                          size_t *off_pf = nullptr, size_t *off_pb = nullptr)
{
    const size_t nblk = (L + 31) / 32;          // one flag per CTA; the narrowest tile is 32 lines
    size_t o = 0;
    o += L * sizeof(float);                 *off_cb = o;
    o += L * sizeof(float);                 *off_ff = o;
    o += nblk * sizeof(unsigned);           *off_fb = o;
    o += nblk * sizeof(unsigned);           *off_err = o;
    o += 64;
    o = (o + 255) & ~(size_t)255;
    // self-validating carries of the tile kernels: one 8-byte {carry, epoch} pair per line and direction
    if (off_pf) *off_pf = o;
    o += L * 8;
    if (off_pb) *off_pb = o;
    o += L * 8;
    return (o + 255) & ~(size_t)255;
}

// Collective: make sure every rank's mailbox can hold L carries per direction and that the
// neighbours' mailboxes are mapped.  Returns the pointers the kernels need.
int pst_comm_mailbox(pst_ctx *c, size_t L, pst_mailbox_view *v)
{
    pst_comm *m = c->comm;
    size_t ocb, off, ofb, oerr;
    if (m->mbox_L < L) {
        PST_CUDA(cudaStreamSynchronize(c->stream));
        if (m->mbox_prev) { cudaIpcCloseMemHandle(m->mbox_prev); m->mbox_prev = nullptr; }
        if (m->mbox_next) { cudaIpcCloseMemHandle(m->mbox_next); m->mbox_next = nullptr; }
        if (m->mbox) { cudaFree(m->mbox); m->mbox = nullptr; }
        m->mbox_bytes = mbox_layout(L, &ocb, &off, &ofb, &oerr);
        PST_CUDA(cudaMalloc((void **)&m->mbox, m->mbox_bytes));
        PST_CUDA(cudaMemset(m->mbox, 0, m->mbox_bytes));
        m->mbox_L = L;
        m->epoch = 0;
        // exchange IPC handles through NCCL (all-gather of 64-byte handles)
        cudaIpcMemHandle_t mine;
        PST_CUDA(cudaIpcGetMemHandle(&mine, m->mbox));
        char *d_h = nullptr;
        PST_CUDA(cudaMalloc((void **)&d_h, sizeof(mine) * (size_t)(c->nranks + 1)));
        PST_CUDA(cudaMemcpy(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice));
        // the memset and the handle upload ran on the legacy default stream, c->stream is non-blocking: order them
        // before the all-gather (whose completion then also proves every peer's mailbox is zeroed)
        PST_CUDA(cudaDeviceSynchronize());
        PST_NCCL(g_nccl.AllGather(d_h, d_h + sizeof(mine), sizeof(mine), 0 /* ncclInt8 */, m->comm, c->stream));
        PST_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<cudaIpcMemHandle_t> all((size_t)c->nranks);
        PST_CUDA(cudaMemcpy(all.data(), d_h + sizeof(mine), sizeof(mine) * (size_t)c->nranks, cudaMemcpyDeviceToHost));
        PST_CUDA(cudaFree(d_h));
        if (c->rank > 0)
            PST_CUDA(cudaIpcOpenMemHandle((void **)&m->mbox_prev, all[(size_t)c->rank - 1], cudaIpcMemLazyEnablePeerAccess));
        if (c->rank < c->nranks - 1)
            PST_CUDA(cudaIpcOpenMemHandle((void **)&m->mbox_next, all[(size_t)c->rank + 1], cudaIpcMemLazyEnablePeerAccess));
    }
    size_t opf, opb;
    mbox_layout(m->mbox_L, &ocb, &off, &ofb, &oerr, &opf, &opb);
    v->pf_in = (uint2 *)(m->mbox + opf);
    v->pb_in = (uint2 *)(m->mbox + opb);
    v->pf_out = m->mbox_next ? (uint2 *)(m->mbox_next + opf) : nullptr;
    v->pb_out = m->mbox_prev ? (uint2 *)(m->mbox_prev + opb) : nullptr;
    v->cf_in = (float *)m->mbox;                        // written by rank-1
    v->cb_in = (float *)(m->mbox + ocb);                // written by rank+1
    v->ff_in = (unsigned *)(m->mbox + off);
    v->fb_in = (unsigned *)(m->mbox + ofb);
    v->err = (unsigned *)(m->mbox + oerr);
    v->cf_out = m->mbox_next ? (float *)m->mbox_next : nullptr;
    v->ff_out = m->mbox_next ? (unsigned *)(m->mbox_next + off) : nullptr;
    v->cb_out = m->mbox_prev ? (float *)(m->mbox_prev + ocb) : nullptr;
    v->fb_out = m->mbox_prev ? (unsigned *)(m->mbox_prev + ofb) : nullptr;
    v->hr_in_prev = (unsigned *)(m->mbox + oerr + 16);
    v->hr_in_next = (unsigned *)(m->mbox + oerr + 32);
    v->hr_out_prev = m->mbox_prev ? (unsigned *)(m->mbox_prev + oerr + 32) : nullptr;   // I am the previous rank's "next"
    v->hr_out_next = m->mbox_next ? (unsigned *)(m->mbox_next + oerr + 16) : nullptr;   // and the next rank's "previous"
    v->epoch = ++m->epoch;
    return PST_OK;
}

// Collective: map the neighbours' workspace arenas (CUDA IPC) so that kernels can read halo planes straight from peer
// memory.  The arenas are re-exchanged whenever ANY rank's arena moved since the last exchange (decided with an
// all-reduce, so that every rank takes the same branch).  *prev / *next: mapped bases (null on the edge ranks).
int pst_comm_map_arenas(pst_ctx *c, char **prev, char **next)
{
    pst_comm *m = c->comm;
    *prev = *next = nullptr;
    if (!m || !c->arena) return PST_OK;
    int *d_flag = nullptr;
    PST_CUDA(cudaMalloc((void **)&d_flag, 64 + sizeof(cudaIpcMemHandle_t) * (size_t)(c->nranks + 1)));
    int changed = (m->arena_mapped_for != c->arena || m->arena_mapped_size != c->arena_size) ? 1 : 0;
    PST_CUDA(cudaMemcpy(d_flag, &changed, sizeof(int), cudaMemcpyHostToDevice));
    PST_CUDA(cudaDeviceSynchronize());
    PST_NCCL(g_nccl.AllReduce(d_flag, d_flag, 1, 2 /* ncclInt32 */, 2 /* ncclMax */, m->comm, c->stream));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    PST_CUDA(cudaMemcpy(&changed, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (changed) {
        if (m->arena_prev) { cudaIpcCloseMemHandle(m->arena_prev); m->arena_prev = nullptr; }
        if (m->arena_next) { cudaIpcCloseMemHandle(m->arena_next); m->arena_next = nullptr; }
        cudaIpcMemHandle_t mine;
        PST_CUDA(cudaIpcGetMemHandle(&mine, c->arena));
        char *d_h = (char *)d_flag + 64;                     // (the allocation is 256-byte aligned)
        PST_CUDA(cudaMemcpy(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice));
        PST_CUDA(cudaDeviceSynchronize());
        PST_NCCL(g_nccl.AllGather(d_h, d_h + sizeof(mine), sizeof(mine), 0 /* ncclInt8 */, m->comm, c->stream));
        PST_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<cudaIpcMemHandle_t> all((size_t)c->nranks);
        PST_CUDA(cudaMemcpy(all.data(), d_h + sizeof(mine), sizeof(mine) * (size_t)c->nranks, cudaMemcpyDeviceToHost));
        if (c->rank > 0)
            PST_CUDA(cudaIpcOpenMemHandle((void **)&m->arena_prev, all[(size_t)c->rank - 1], cudaIpcMemLazyEnablePeerAccess));
        if (c->rank < c->nranks - 1)
            PST_CUDA(cudaIpcOpenMemHandle((void **)&m->arena_next, all[(size_t)c->rank + 1], cudaIpcMemLazyEnablePeerAccess));
        m->arena_mapped_for = c->arena;
        m->arena_mapped_size = c->arena_size;
    }
    PST_CUDA(cudaFree(d_flag));
    *prev = m->arena_prev;
    *next = m->arena_next;
    return PST_OK;
}

// after a call: did any kernel give up waiting for a neighbour?
int pst_comm_check(pst_ctx *c)
{
    if (!c->comm || !c->comm->mbox) return PST_OK;
    size_t ocb, off, ofb, oerr;
    mbox_layout(c->comm->mbox_L, &ocb, &off, &ofb, &oerr);
    unsigned e = 0;
    PST_CUDA(cudaMemcpyAsync(&e, c->comm->mbox + oerr, sizeof(e), cudaMemcpyDeviceToHost, c->stream));
    PST_CUDA(cudaStreamSynchronize(c->stream));
    if (e) { pst_set_error("axis-3 carry pipeline: timed out waiting for a neighbour rank"); return PST_ECOMM; }
    return PST_OK;
}
