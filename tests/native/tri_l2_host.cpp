// Host build of the L2-resident checkpoint + recompute smoother core (pyseistr_b200/csrc/pst_tri_l2_core.h): the same
// per-line code the CUDA kernel runs, one line at a time, with a block transport that checks the stream the kernel's
// TMA ring delivers (blocks 0 .. nld-1 upwards, then nrel-1 .. 0 downwards).  Test infrastructure
// (tests/test_tri_l2_core.py).
#include <vector>

#include "pst_tri_l2_core.h"

struct HostIO {
    const float *s; float *d; long st; int nx, nld, nrel, q;
    void load(float *x)
    {
        const int m = q < nld ? q : (nrel - 1) - (q - nld);
        q++;
        for (int j = 0; j < 32; j++) { const int i = m * 32 + j; x[j] = (m >= 0 && i < nx) ? s[(long)i * st] : 0.f; }
    }
    void store(const float *v, int i0)
    {
        for (int j = 0; j < 32; j++) if (i0 + j < nx) d[(long)(i0 + j) * st] = v[j];
    }
};

template <int NB>
static int run(const float *src, float *dst, long na, long nslab, long sb, long d, int nx)
{
    const float wt = (float)(1.0 / ((double)NB * NB));
    const float wm = -wt, w2 = (float)(2. * wt);
    std::vector<float> ck((nx + 2 * NB + 31) / 32 + 1);
    const int nld = (nx + 31) / 32, nrel = tri_l2::reload_count(nx, NB);
    for (long s = 0; s < nslab; s++)
        for (long a = 0; a < na; a++) {
            const long base = a + s * sb;
            HostIO io{src + base, dst + base, d, nx, nld, nrel, 0};
            tri_l2::smooth_line<NB>(io, nx, wm, w2, ck.data(), 1);
            if (io.q != nld + nrel) return -9;          // the core consumed exactly the stream the kernel issues
        }
    return 0;
}

extern "C" int tri_l2_host(const float *src, float *dst, int n1, int n2, int n3, int axis, int nb)
{
    long na, nslab, sb, d;
    int nx;
    if (axis == 0) { na = 1; nslab = (long)n2 * n3; sb = n1; d = 1; nx = n1; }
    else if (axis == 1) { na = n1; nslab = n3; sb = (long)n1 * n2; d = n1; nx = n2; }
    else { na = (long)n1 * n2; nslab = 1; sb = 0; d = (long)n1 * n2; nx = n3; }
    if (nb > nx) return -1;
#define CASE(N) case N: return run<N>(src, dst, na, nslab, sb, d, nx);
    switch (nb) {
        CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(10) CASE(16)
    }
#undef CASE
    return -3;
}
