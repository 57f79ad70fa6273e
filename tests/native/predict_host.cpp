// Host build of the plane-wave prediction core (pyseistr_b200/csrc/pst_predict_core.h): the per-trace code of
// predict_fast_kernel, run for ntr traces stored trace-minor ([k][i2], like the kernel's volumes).  Test infrastructure
// (tests/test_predict_core.py).
#include <cstring>
#include <vector>

#include "pst_predict_core.h"
#include "pst_predict_warp.h"

// x1/x2, g1/g2, out: [n1][ntr] (i2 fastest).  two = 0: one parent.
extern "C" int predict_host(const float *x1, const float *g1, const float *x2, const float *g2, int n1, int ntr, int nw,
                            int two, int forw1, int forw2, float eps, float *out)
{
    PredArgs A;
    memset(&A, 0, sizeof(A));
    A.in1 = x1; A.sg1 = g1; A.in2 = x2; A.sg2 = g2; A.forw1 = forw1; A.forw2 = forw2;
    A.out = out;
    std::vector<float> scr((size_t)n1 * ntr * (2 * nw + 1));
    A.scr = scr.data();
    A.n1 = n1; A.n2 = ntr; A.n3 = 1; A.ze0 = 0; A.zla = 0; A.zlb = 1; A.a = 0; A.b = 0; A.t_off = 0; A.ntg = ntr;
    A.reg = make_reg(eps);
    A.tb = make_btab_s(nw);
    for (int i2 = 0; i2 < ntr; i2++) {
        if (nw == 1 && !two) predict_fast_trace<1, false>(A, i2, 0, 0);
        else if (nw == 1) predict_fast_trace<1, true>(A, i2, 0, 0);
        else if (nw == 2 && !two) predict_fast_trace<2, false>(A, i2, 0, 0);
        else if (nw == 2) predict_fast_trace<2, true>(A, i2, 0, 0);
        else return -1;
    }
    return 0;
}

// the warp-per-trace formulation (pst_predict_warp.h), lanes run one after the other phase by phase
template <int NW, bool TWO>
static void warp_one(const float *x1, const float *g1, const float *x2, const float *g2, long ks, int forw1, int forw2,
                     float eps, int n1, float *scr, float *out)
{
    static PredWarpWS<NW, TWO> ws;
    predict_warp_trace<NW, TWO>(ws, x1, g1, x2, g2, ks, forw1 != 0, forw2 != 0, make_reg(eps), make_btab_s(NW), n1, scr, out);
}

extern "C" int predict_warp_host(const float *x1, const float *g1, const float *x2, const float *g2, int n1, int ntr, int nw,
                                 int two, int forw1, int forw2, float eps, float *out)
{
    std::vector<float> scr((size_t)n1 * (2 * nw + 1));
    for (int t = 0; t < ntr; t++) {
        if (nw == 1 && !two) warp_one<1, false>(x1 + t, g1 + t, x2 + t, g2 + t, ntr, forw1, forw2, eps, n1, scr.data(), out + t);
        else if (nw == 1) warp_one<1, true>(x1 + t, g1 + t, x2 + t, g2 + t, ntr, forw1, forw2, eps, n1, scr.data(), out + t);
        else if (nw == 2 && !two) warp_one<2, false>(x1 + t, g1 + t, x2 + t, g2 + t, ntr, forw1, forw2, eps, n1, scr.data(), out + t);
        else if (nw == 2) warp_one<2, true>(x1 + t, g1 + t, x2 + t, g2 + t, ntr, forw1, forw2, eps, n1, scr.data(), out + t);
        else return -1;
    }
    return 0;
}
