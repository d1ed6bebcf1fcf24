// TEST INFRASTRUCTURE ONLY.  The producer side of the small-LP batches, driven
// through the UNMODIFIED reference: DepPoly::is_empty (src/eng/poly.cpp:530-573) is what
// DepPolyMgr::buildDepPoly calls per (reference pair, loop depth) (poly.cpp:1166-1195); it runs
// the reference's pre-filter Lineq::reduce (src/com/linsys.cpp:359) and then asks
// Lineq::has_solution (linsys.cpp:830-906).  This program builds dependence polyhedra of
// two-deep loop nests with affine subscripts, calls the reference's own is_empty on each, and
// -- through GNU ld's --wrap on the mangled name of Lineq::has_solution, no reference source is
// touched -- records every system that reaches has_solution after the pre-filter together with
// the reference's answer.
//
//   ref_producer record <out.json>   write the recorded queries (golden fixture generator)
//   ref_producer replay              answer the same recorded queries in ONE batch with
//                                    XpHasSolutionBatch (xp_six.hpp -> libxpoly_b200.so, GPU)
//                                    and compare every answer with the reference's
#include "ltype.h"
#include "comf.h"
#include "strbuf.h"
#include "smempool.h"
#include "sstl.h"
#include "matt.h"
#include "bs.h"
#include "sbs.h"
#include "sgraph.h"
#include "rational.h"
#include "flty.h"
#include "xmat.h"
#include "linsys.h"
#include "lpsol.h"
using namespace xcom;
#include "depvecs.h"
#include "poly.h"
#ifdef XP_WITH_ADAPTOR
#include "xp_six.hpp"
#endif

#include <stdio.h>
#include <string.h>

#include <vector>

struct Recorded {
    std::vector<int> leq; // rows x (rhs_idx + 1) integers (dependence polyhedra are integral)
    int rows, rhs_idx;
    bool is_int, is_unique, answer;
};
static std::vector<Recorded> g_rec;

extern "C" bool __real__ZN4xcom5Lineq12has_solutionERKNS_4RMatES3_RS1_jbb(Lineq *self, RMat const &leq, RMat const &eq,
                                                                           RMat &vc, UINT rhs_idx, bool is_int,
                                                                           bool is_unique);
extern "C" bool __wrap__ZN4xcom5Lineq12has_solutionERKNS_4RMatES3_RS1_jbb(Lineq *self, RMat const &leq, RMat const &eq,
                                                                           RMat &vc, UINT rhs_idx, bool is_int,
                                                                           bool is_unique)
{
    const bool ans = __real__ZN4xcom5Lineq12has_solutionERKNS_4RMatES3_RS1_jbb(self, leq, eq, vc, rhs_idx, is_int, is_unique);
    Recorded r;
    r.rows = (int)leq.get_row_size();
    r.rhs_idx = (int)rhs_idx;
    r.is_int = is_int;
    r.is_unique = is_unique;
    r.answer = ans;
    bool integral = eq.size() == 0; // (is_empty passes no equalities)
    for (UINT i = 0; i < leq.get_row_size(); i++)
        for (UINT j = 0; j < leq.get_col_size(); j++) {
            Rational v = leq.get(i, j);
            if (v.den() != 1) integral = false;
            r.leq.push_back(v.num());
        }
    if (integral) g_rec.push_back(r);
    return ans;
}

static unsigned long long g_s = 88172645463325252ULL;
static int rnd(int lo, int hi)
{ // xorshift64
    g_s ^= g_s << 13;
    g_s ^= g_s >> 7;
    g_s ^= g_s << 17;
    return lo + (int)(g_s % (unsigned long long)(hi - lo + 1));
}

// Dependence polyhedron of  A[a1*i + b1*j + c1]  (iteration (i, j))  against
// A[a2*i' + b2*j' + c2]  (iteration (i', j')),  1 <= i, i' <= N, 1 <= j, j' <= M, variables
// (i, j, i', j'), rows `coeffs . x <= rhs`; the subscript equality as two inequalities; `depth`
// 1 / 2: carried by the outer / inner loop (i' >= i + 1, or i' == i and j' >= j + 1); 0: no
// ordering condition (loop independent test).
static void build(DepPoly &dp, int N, int M, int a1, int b1, int c1, int a2, int b2, int c2, int depth)
{
    std::vector<std::vector<int> > rows;
    for (int v = 0; v < 4; v++) {
        std::vector<int> lo(5, 0), hi(5, 0);
        lo[v] = -1, lo[4] = -1;                   // -x <= -1
        hi[v] = 1, hi[4] = (v % 2 == 0) ? N : M;  //  x <= N / M
        rows.push_back(lo);
        rows.push_back(hi);
    }
    std::vector<int> e(5, 0);
    e[0] = a1, e[1] = b1, e[2] = -a2, e[3] = -b2, e[4] = c2 - c1;
    rows.push_back(e);
    for (int k = 0; k < 5; k++) e[k] = -e[k];
    rows.push_back(e);
    if (depth == 1) {
        std::vector<int> r(5, 0);
        r[0] = 1, r[2] = -1, r[4] = -1; // i - i' <= -1
        rows.push_back(r);
    } else if (depth == 2) {
        std::vector<int> r(5, 0), s(5, 0), t(5, 0);
        r[0] = 1, r[2] = -1;            // i - i' <= 0
        s[0] = -1, s[2] = 1;            // i' - i <= 0
        t[1] = 1, t[3] = -1, t[4] = -1; // j - j' <= -1
        rows.push_back(r);
        rows.push_back(s);
        rows.push_back(t);
    }
    dp.reinit((UINT)rows.size(), 5);
    for (size_t i = 0; i < rows.size(); i++)
        for (int j = 0; j < 5; j++) dp.set((UINT)i, (UINT)j, Rational(rows[i][j], 1));
    dp.rhs_idx = 4;
}

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    std::vector<bool> empties;
    for (int k = 0; k < 400; k++) { // the loop of DepPolyMgr::buildDepPoly: one is_empty per pair and depth
        DepPoly dp;
        build(dp, rnd(3, 12), rnd(3, 12), rnd(-2, 3), rnd(-2, 3), rnd(-3, 6), rnd(-2, 3), rnd(-2, 3), rnd(-3, 6), k % 3);
        empties.push_back(dp.is_empty(true, NULL));
    }
    if (!strcmp(argv[1], "record")) {
        FILE *f = fopen(argc > 2 ? argv[2] : "deppoly_queries.json", "w");
        if (!f) return 3;
        size_t n_empty = 0;
        for (size_t k = 0; k < empties.size(); k++) n_empty += empties[k];
        fprintf(f, "{\"generator\": \"oracle/ref_producer.cpp record (unmodified reference: DepPoly::is_empty -> Lineq::reduce -> "
                   "Lineq::has_solution)\", \"polyhedra\": %zu, \"empty\": %zu, \"queries\": [", empties.size(), n_empty);
        for (size_t q = 0; q < g_rec.size(); q++) {
            const Recorded &r = g_rec[q];
            fprintf(f, "%s{\"rows\": %d, \"rhs_idx\": %d, \"is_int\": %d, \"is_unique\": %d, \"answer\": %d, \"leq\": [", q ? ", " : "",
                    r.rows, r.rhs_idx, (int)r.is_int, (int)r.is_unique, (int)r.answer);
            for (size_t e = 0; e < r.leq.size(); e++) fprintf(f, "%s%d", e ? "," : "", r.leq[e]);
            fprintf(f, "]}");
        }
        fprintf(f, "]}\n");
        fclose(f);
        printf("recorded %zu has_solution queries from %zu polyhedra (%zu empty)\n", g_rec.size(), empties.size(), n_empty);
        return 0;
    }
#ifdef XP_WITH_ADAPTOR
    if (!strcmp(argv[1], "replay")) {
        // the same queries, collected with has_solution's own arguments and answered in ONE batch
        XpHasSolutionBatch hb;
        std::vector<RMat> keep(g_rec.size());
        RMat noeq;
        for (size_t q = 0; q < g_rec.size(); q++) {
            const Recorded &r = g_rec[q];
            keep[q].reinit(r.rows, r.rhs_idx + 1);
            for (int i = 0; i < r.rows; i++)
                for (int j = 0; j <= r.rhs_idx; j++) keep[q].set(i, j, Rational(r.leq[(size_t)i * (r.rhs_idx + 1) + j], 1));
            hb.add(keep[q], noeq, (UINT)r.rhs_idx);
        }
        hb.run(true, true);
        size_t bad = 0;
        for (size_t q = 0; q < g_rec.size(); q++) bad += hb.get((UINT)q) != g_rec[q].answer;
        printf("replayed %zu queries in one batch: %zu mismatches, %u undecided\n", g_rec.size(), bad, hb.failed());
        return bad == 0 && hb.failed() == 0 ? 0 : 1;
    }
#endif
    return 2;
}
