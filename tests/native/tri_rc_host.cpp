// Host build of the checkpoint + recompute smoother core (pyseistr_b200/csrc/pst_tri_rc_core.h): the same
// per-line code the CUDA kernels run, one line at a time.  Test infrastructure (tests/test_tri_rc_core.py).
#include <vector>

#include "pst_tri_rc_core.h"

template <int NB, int RC>
static void run(const float *src, float *dst, long na, long nslab, long sb, long d, int nx)
{
    const float wt = (float)(1.0 / ((double)NB * NB));
    const float wm = -wt, w2 = (float)(2. * wt);
    std::vector<float> ck((nx + 2 * NB + RC - 1) / RC + 1);
    for (long s = 0; s < nslab; s++)
        for (long a = 0; a < na; a++) {
            const long base = a + s * sb;
            tri_rc::StridedIO<RC> io{src + base, dst + base, d, nx, {}};
            tri_rc::process_line<NB, RC>(io, nx, wm, w2, ck.data(), 1);
        }
}

extern "C" int tri_rc_host(const float *src, float *dst, int n1, int n2, int n3, int axis, int nb, int rc)
{
    long na, nslab, sb, d;
    int nx;
    if (axis == 0) { na = 1; nslab = (long)n2 * n3; sb = n1; d = 1; nx = n1; }
    else if (axis == 1) { na = n1; nslab = n3; sb = (long)n1 * n2; d = n1; nx = n2; }
    else { na = (long)n1 * n2; nslab = 1; sb = 0; d = (long)n1 * n2; nx = n3; }
    if (nb > nx) return -1;
#define CASE(N) \
    case N: \
        if (rc == 32) run<N, 32>(src, dst, na, nslab, sb, d, nx); \
        else if (rc == 16 && 2 * N <= 16) run<N, (2 * N <= 16 ? 16 : 32)>(src, dst, na, nslab, sb, d, nx); \
        else return -2; \
        return 0;
    switch (nb) {
        CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(10) CASE(16)
    }
#undef CASE
    return -3;
}
