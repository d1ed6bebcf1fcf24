// Compile/link check of xpoly_b200/host/xp_six.hpp against the reference's headers: the
// reference's own example (example.cpp:54-93) through SIX<FloatMat,Float>::maxm and
// Lineq::has_solution-style MIP calls, routed to libxpoly_b200.so.  Running it needs a GPU.
#include "ltype.h"
#include "comf.h"
#include "strbuf.h"
#include "smempool.h"
#include "rational.h"
#include "flty.h"
#include "sstl.h"
#include "matt.h"
#include "bs.h"
#include "sbs.h"
#include "sgraph.h"
#include "xmat.h"
#include "linsys.h"
#include "lpsol.h"
#include "xp_six.hpp"

using namespace xcom;

int main()
{
    // max 2x1 - x2  s.t. 2x1 - x2 <= 2, x1 - 5x2 <= -4, x >= 0   (example.cpp:54-61)
    FloatMat leq(2, 3), tgtf(1, 3), vc(2, 3), eq, sol;
    double L[2][3] = {{2, -1, 2}, {1, -5, -4}};
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 3; j++) leq.set(i, j, Float(L[i][j]));
    tgtf.set(0, 0, Float(2.0));
    tgtf.set(0, 1, Float(-1.0));
    tgtf.set(0, 2, Float(0.0));
    vc.set(0, 0, Float(-1.0));
    vc.set(1, 1, Float(-1.0));
    SIX<FloatMat, Float> six;
    Float maxv;
    UINT st = six.maxm(maxv, sol, tgtf, vc, eq, leq);
    printf("status %u max %.17g x = (%.17g, %.17g)\n", st, maxv.f(), sol.get(0, 0).f(), sol.get(0, 1).f());

    {   // the only reference entry that exposes the solver state: TwoStageMethod (lpsol.h:291-301),
        // here with phase 1 (b has a negative entry); every IN OUT argument is printed
        FloatMat l2(leq), t2(tgtf), v2(vc), ssol;
        Vector<bool> nvs, bvs;
        Vector<INT> b2e, e2b;
        Float mv(0.0);
        INT rhs = 2;
        SIX<FloatMat, Float> s2;
        UINT stt = s2.TwoStageMethod(l2, v2, t2, ssol, mv, nvs, bvs, b2e, e2b, rhs);
        printf("two_stage status %u maxv %.17g rhs_idx %d eq2bv %d %d bv2eq %d %d %d %d nv %d%d%d%d\n", stt, mv.f(), (int)rhs,
               (int)e2b.get(0), (int)e2b.get(1), (int)b2e.get(0), (int)b2e.get(1), (int)b2e.get(2), (int)b2e.get(3),
               (int)nvs.get(0), (int)nvs.get(1), (int)nvs.get(2), (int)nvs.get(3));
        printf("two_stage tableau %u x %u:", l2.get_row_size(), l2.get_col_size());
        for (UINT i = 0; i < l2.get_row_size(); i++)
            for (UINT j = 0; j < l2.get_col_size(); j++) printf(" %.17g", l2.get(i, j).f());
        printf("\ntwo_stage tgtf:");
        for (UINT j = 0; j < t2.get_col_size(); j++) printf(" %.17g", t2.get(0, j).f());
        printf("\ntwo_stage slack_sol:");
        for (UINT j = 0; j < ssol.get_col_size(); j++) printf(" %.17g", ssol.get(0, j).f());
        printf("\ntwo_stage vc %u x %u diag:", v2.get_row_size(), v2.get_col_size());
        for (UINT j = 0; j < v2.get_row_size(); j++) printf(" %.17g", v2.get(j, j).f());
        printf("\n");
    }

    RMat rleq(2, 3), rtg(1, 3), rvc(2, 3), req, rsol;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 3; j++) rleq.set(i, j, Rational((int)L[i][j], 1));
    rtg.set(0, 0, Rational(2, 1));
    rtg.set(0, 1, Rational(-1, 1));
    rtg.set(0, 2, Rational(0, 1));
    rvc.set(0, 0, Rational(-1, 1));
    rvc.set(1, 1, Rational(-1, 1));
    SIX<RMat, Rational> rsix;
    Rational rv;
    UINT st2 = rsix.minm(rv, rsol, rtg, rvc, req, rleq);
    printf("minm_rat status %u v %d/%d\n", st2, rv.num(), rv.den());
    MIP<RMat, Rational> mip;
    UINT st3 = mip.maxm(rv, rsol, rtg, rvc, req, rleq);
    printf("mip_max_rat status %u v %d/%d x = (%d/%d, %d/%d)\n", st3, rv.num(), rv.den(), rsol.get(0, 0).num(),
           rsol.get(0, 0).den(), rsol.get(0, 1).num(), rsol.get(0, 1).den());
    // Lineq::has_solution's dependence queries, collected and answered at once
    static const int QA[6][3] = {{-1, 0, -1}, {1, 0, 10}, {0, -1, -1}, {0, 1, 10}, {1, -1, 1}, {-1, 1, -1}};
    static const int QB[5][3] = {{-1, 0, -1}, {1, 0, 10}, {0, -1, -1}, {0, 1, 10}, {1, -1, -20}};
    static const int QC[6][3] = {{-1, 0, -1}, {1, 0, 10}, {0, -1, -1}, {0, 1, 10}, {2, -2, 1}, {-2, 2, -1}};
    RMat qa(6, 3), qb(5, 3), qc(6, 3), qe;
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 3; j++) {
            qa.set(i, j, Rational(QA[i][j], 1));
            qc.set(i, j, Rational(QC[i][j], 1));
            if (i < 5) qb.set(i, j, Rational(QB[i][j], 1));
        }
    XpHasSolutionBatch hb;
    hb.add(qa, qe, 2);
    hb.add(qb, qe, 2);
    hb.add(qc, qe, 2);
    hb.run(true, true);
    printf("has_solution %d %d %d\n", (int)hb.get(0), (int)hb.get(1), (int)hb.get(2));
    return st == SIX_SUCC && maxv.f() == 2.0 ? 0 : 1;
}
