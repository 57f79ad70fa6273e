// Multi-GPU plumbing (slab decomposition along n3, one process per GPU).
// Round-1 state: single-GPU contexts only; the distributed entry points are declared in the
// C-ABI and report PST_EUNSUP until the NCCL halo / carry exchange lands.
#include "pst_common.cuh"

#include <string.h>

struct pst_comm { int dummy; };

int pst_comm_allreduce_record(pst_ctx *c, double *d_rec, int nv)
{
    (void)c; (void)d_rec; (void)nv;
    return PST_OK;
}

void pst_comm_destroy(pst_ctx *c)
{
    delete c->comm;
    c->comm = nullptr;
}

extern "C" int pst_comm_unique_id(void *id128)
{
    if (!id128) { pst_set_error("null argument"); return PST_EINVAL; }
    memset(id128, 0, 128);
    pst_set_error("multi-GPU communicator not built in this version");
    return PST_EUNSUP;
}

extern "C" int pst_ctx_create_dist(int device, int rank, int nranks, const void *nccl_id128, pst_ctx **ctx)
{
    (void)nccl_id128;
    if (nranks == 1 && rank == 0) return pst_ctx_create(device, ctx);
    pst_set_error("multi-GPU slab decomposition not built in this version (nranks=%d)", nranks);
    return PST_EUNSUP;
}
