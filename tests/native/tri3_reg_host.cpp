// Host build of the per-line core of the register kernels of the multi-GPU axis-3 smoothing pass
// (pyseistr_b200/csrc/pst_tri3_reg_core.h): the same code the CUDA kernels run, with the ranks taken one after the other --
// forward walks of rank 0 .. N-1 (each handing its carries to the next), then backward walks of rank N-1 .. 0 -- over the
// n3-slabs of one volume, in place, like the kernels.  Test infrastructure (tests/test_tri3_reg_core.py).
#include <vector>

#include "pst_tri3_reg_core.h"

struct HostArgs {
    const float *x, *hb, *ha;
    float *ha_keep, *csave, *dst;
    long L;
    int nz, n3g, K0, K1;
    float wt, w2;
};
struct HostIO {
    const float *carry;                       // incoming carries of this rank and direction (null: none is ever asked for)
    int waits = 0;
    void wait_halos() { waits++; }
    float recv(long l) { return carry[l]; }
};

template <int NB, int NZ>
static int run(float *x, int n1, int n2, int n3, int nranks)
{
    const long L = (long)n1 * n2;
    const int nz = n3 / nranks, nch = nz / NZ;
    std::vector<std::vector<float>> csave(nranks), keep(nranks), cf(nranks + 1), cb(nranks + 1);
    for (int r = 0; r < nranks; r++) { csave[r].assign((size_t)nch * L, -7.f); keep[r].assign((size_t)NB * L, -7.f); }
    for (int r = 0; r <= nranks; r++) { cf[r].assign(L, -7.f); cb[r].assign(L, 0.f); }
    auto args = [&](int r) {
        HostArgs A;
        const int z0 = r * nz, z1 = z0 + nz;
        const bool first = r == 0, last = r == nranks - 1;
        A.x = x + (size_t)z0 * L; A.dst = x + (size_t)z0 * L;
        A.hb = first ? nullptr : x + (size_t)(z0 - NB) * L;
        A.ha = last ? nullptr : x + (size_t)z1 * L;
        A.ha_keep = last ? nullptr : keep[r].data();
        A.csave = csave[r].data();
        A.L = L; A.nz = nz; A.n3g = n3;
        A.K0 = first ? 0 : z0 + NB; A.K1 = last ? n3 + 2 * NB : z1 + NB;
        A.wt = (float)(1.0 / ((double)NB * NB)); A.w2 = (float)(2. * A.wt);
        return A;
    };
    for (int r = 0; r < nranks; r++) {
        const HostArgs A = args(r);
        HostIO io{r > 0 ? cf[r].data() : nullptr};
        for (long l = 0; l < L; l++) {
            float s = 0.f;
            if (!tri3_reg::fwd_line<NB, NZ>(A, io, l, true, &s)) return -5;
            cf[r + 1][l] = s;
        }
        if (io.waits != L) return -6;          // every thread passes the halo flags exactly once per pass
    }
    for (int r = nranks - 1; r >= 0; r--) {
        const HostArgs A = args(r);
        HostIO io{cb[r + 1].data()};           // (the last rank receives +0)
        for (long l = 0; l < L; l++) cb[r][l] = tri3_reg::bwd_line<NB, NZ>(A, io, l);
    }
    return 0;
}

template <int NB>
static int run_nb(float *x, int n1, int n2, int n3, int nranks, int NZ)
{
    return NZ == 128 ? run<NB, 128>(x, n1, n2, n3, nranks) : run<NB, 32>(x, n1, n2, n3, nranks);
}

// mirrors tri3_reg_chunk() of pst_dip.cu
extern "C" int tri3_reg_host(float *x, int n1, int n2, int n3, int nb, int nranks)
{
    if (nranks < 2 || n3 % nranks != 0) return -1;
    const int nz = n3 / nranks;
    const int NZ = nz % 128 == 0 ? 128 : ((nz % 32 == 0 && nz <= 96) ? 32 : 0);
    if (!NZ || NZ < 2 * nb) return -1;
    switch (nb) {
        case 2: return run_nb<2>(x, n1, n2, n3, nranks, NZ);
        case 3: return run_nb<3>(x, n1, n2, n3, nranks, NZ);
        case 4: return run_nb<4>(x, n1, n2, n3, nranks, NZ);
        case 5: return run_nb<5>(x, n1, n2, n3, nranks, NZ);
        case 6: return run_nb<6>(x, n1, n2, n3, nranks, NZ);
        case 8: return run_nb<8>(x, n1, n2, n3, nranks, NZ);
    }
    return -3;
}
