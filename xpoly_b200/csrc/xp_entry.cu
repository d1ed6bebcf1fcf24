// Entry-level C ABI: SIX<Mat,T>::maxm / minm, MIP<Mat,T>::maxm / minm and
// Lineq::has_solution over raw row-major arrays.  The host half (verify,
// normalize, dual, calcFinalSolution, B&B bookkeeping) is xp_host_six.hpp; every
// TwoStageMethod (slack form, phase 1, solveSlackForm) runs on the GPU:
//   - LPs that fit shared memory go to the one-CTA-per-LP kernels, many per launch;
//   - larger FP64 LPs go to the HBM-resident path (xp_six_two_stage_f64_large, phase 1 on
//     the device too).
// There is no CPU solve path in here.
#include "xp_batch_core.cuh"

#include "../host/xp_host_six.hpp"

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

using namespace xph;

namespace {

// ---------------------------------------------------------------- FP64 side
typedef TwoStageResult<F64> ResF;

bool fits_smem(const xp_ctx *ctx, int m, int n, size_t key_bytes)
{
    return xpb_smem_bytes(m, n, 8, key_bytes) <= ctx->smem_optin;
}

// Host-side loops over independent items (trees of a B&B batch, LPs of a ragged launch) run
// on a small pool of worker threads (created on first use, parked on a condition variable in
// between: a B&B batch calls this four times per wave, far too often to spawn threads).
class HostPool {
    std::vector<std::thread> th_;
    std::mutex m_, busy_;
    std::condition_variable cv_, done_;
    std::function<void()> job_;
    unsigned long long gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;

public:
    explicit HostPool(unsigned n)
    {
        for (unsigned t = 0; t < n; t++)
            th_.emplace_back([this]() {
                unsigned long long seen = 0;
                for (;;) {
                    std::function<void()> j;
                    {
                        std::unique_lock<std::mutex> l(m_);
                        cv_.wait(l, [&]() { return stop_ || gen_ != seen; });
                        if (stop_) return;
                        seen = gen_;
                        j = job_;
                    }
                    j();
                    {
                        std::lock_guard<std::mutex> l(m_);
                        if (--pending_ == 0) done_.notify_one();
                    }
                }
            });
    }
    ~HostPool()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    unsigned workers() const { return (unsigned)th_.size(); }
    // Runs f on every worker and on the caller; false if the pool is in use by another host
    // thread (the caller then runs f alone).
    bool run(const std::function<void()> &f)
    {
        if (th_.empty() || !busy_.try_lock()) return false;
        {
            std::lock_guard<std::mutex> l(m_);
            job_ = f;
            pending_ = (int)th_.size();
            gen_++;
        }
        cv_.notify_all();
        f();
        {
            std::unique_lock<std::mutex> l(m_);
            done_.wait(l, [&]() { return pending_ == 0; });
        }
        busy_.unlock();
        return true;
    }
};
HostPool &host_pool()
{
    unsigned T = std::thread::hardware_concurrency();
    if (T > 16) T = 16;
    static HostPool p(T > 1 ? T - 1 : 0);
    return p;
}

// run f(i) for i in [0, n).  The exact policy's overflow flag is thread-local and sticky, so
// every item carries its own copy (ovf[i], in/out): an item never sees another item's
// overflow, whatever thread it ran on -- the result does not depend on the schedule.
template <class P, class F>
void for_items(size_t n, std::vector<char> &ovf, F f)
{
    auto one = [&](size_t i) {
        P::overflow() = ovf[i] != 0;
        f(i);
        ovf[i] = P::overflow();
    };
    const bool keep = P::overflow();
    bool done = false;
    if (n >= 64) {
        HostPool &pool = host_pool();
        std::atomic<size_t> next(0);
        size_t chunk = n / (8 * (pool.workers() + 1));
        if (chunk < 1) chunk = 1;
        if (chunk > 128) chunk = 128;
        done = pool.run([&]() {
            const bool mine = P::overflow();
            for (;;) {
                const size_t b = next.fetch_add(chunk);
                if (b >= n) break;
                const size_t e = b + chunk < n ? b + chunk : n;
                for (size_t i = b; i < e; i++) one(i);
            }
            P::overflow() = mine;
        });
    }
    if (!done)
        for (size_t i = 0; i < n; i++) one(i);
    P::overflow() = keep;
}

// TwoStageMethod (lpsol.h:1906-1930) for one FP64 LP on the HBM-resident path; phase 1
// (constructBasicFeasibleSolution, :838-988) runs on the device as well.
int two_stage_large_f64(xp_ctx *ctx, const Mat<F64> &leq, const Mat<F64> &tg, uint32_t max_iter,
                        ResF &R, const std::vector<double> *vc_diag = nullptr,
                        const std::vector<double> *vc_rhs = nullptr)
{
    const int m = leq.r, n = leq.c - 1, Cm = n + m + 1;
    R.slack_sol.assign(Cm, 0.0);
    R.tgtf.assign(Cm, 0.0);
    R.eq2bv.assign(m, 0);
    R.maxv = 0.0;
    int32_t st = 0;
    int rc = xp_six_two_stage_f64_large_vc(ctx, m, n, leq.a.data(), tg.a.data(), vc_diag ? vc_diag->data() : nullptr,
                                           vc_rhs ? vc_rhs->data() : nullptr, max_iter, XP_RULE_REFERENCE, &st,
                                           &R.maxv, R.slack_sol.data(), R.tgtf.data(), R.eq2bv.data(), nullptr,
                                           nullptr);
    if (rc) return rc;
    R.status = st;
    return 0;
}

// TwoStageMethod for a set of normalised FP64 LPs: one ragged batched launch for
// everything that fits shared memory, the HBM-resident path for the rest.
int two_stage_many_f64(xp_ctx *ctx, const std::vector<const Mat<F64> *> &leqs,
                       const std::vector<const Mat<F64> *> &tgs, uint32_t max_iter,
                       std::vector<ResF> &out)
{
    const int B = (int)leqs.size();
    out.assign(B, ResF());
    std::vector<int> small;
    for (int k = 0; k < B; k++) {
        const int m = leqs[k]->r, n = leqs[k]->c - 1;
        if (m < 1 || n < 1) {
            out[k].status = XP_ERR_BAD_ARG;
            continue;
        }
        if (fits_smem(ctx, m, n, 16)) small.push_back(k);
        else {
            int rc = two_stage_large_f64(ctx, *leqs[k], *tgs[k], max_iter, out[k]);
            if (rc) return rc;
        }
    }
    if (small.empty()) return 0;
    const int S = (int)small.size();
    std::vector<int32_t> ms(S), ns(S);
    std::vector<int64_t> lo(S), to(S);
    size_t ll = 0, tl = 0;
    int ldo = 0, ldm = 0;
    for (int s = 0; s < S; s++) {
        const Mat<F64> &L = *leqs[small[s]];
        ms[s] = L.r;
        ns[s] = L.c - 1;
        lo[s] = (int64_t)ll;
        to[s] = (int64_t)tl;
        ll += L.a.size();
        tl += (size_t)L.c;
        ldo = std::max(ldo, L.r + L.c);
        ldm = std::max(ldm, L.r);
    }
    std::vector<double> lp(ll), tp(tl), maxv(S), sol((size_t)S * ldo), tgo((size_t)S * ldo);
    std::vector<int32_t> status(S), e2b((size_t)S * ldm);
    for (int s = 0; s < S; s++) {
        const Mat<F64> &L = *leqs[small[s]], &T = *tgs[small[s]];
        std::copy(L.a.begin(), L.a.end(), lp.begin() + lo[s]);
        std::copy(T.a.begin(), T.a.end(), tp.begin() + to[s]);
    }
    int rc = xp_six_two_stage_f64_ragged(ctx, S, ms.data(), ns.data(), lo.data(), to.data(),
                                         lp.data(), ll, tp.data(), tl, max_iter, XP_RULE_REFERENCE,
                                         ldo, ldm, status.data(), maxv.data(), sol.data(),
                                         tgo.data(), e2b.data(), nullptr, nullptr);
    if (rc) return rc;
    for (int s = 0; s < S; s++) {
        ResF &R = out[small[s]];
        const int Cm = ms[s] + ns[s] + 1;
        R.status = status[s];
        R.maxv = maxv[s];
        R.slack_sol.assign(sol.begin() + (size_t)s * ldo, sol.begin() + (size_t)s * ldo + Cm);
        R.tgtf.assign(tgo.begin() + (size_t)s * ldo, tgo.begin() + (size_t)s * ldo + Cm);
        R.eq2bv.assign(e2b.begin() + (size_t)s * ldm, e2b.begin() + (size_t)s * ldm + ms[s]);
    }
    return 0;
}

// --------------------------------------------------------------- exact side
typedef TwoStageResult<Q> ResQ;

long long lcm_ll(long long a, long long b, bool &ovf)
{
    long long g = Q::gcdll(a, b);
    __int128 l = (__int128)(a / g) * b;
    if (l > (__int128)0x7fffffffffffffffLL) {
        ovf = true;
        return 1;
    }
    return (long long)l;
}

// TwoStageMethod for a set of normalised exact LPs.  Rows (and the objective)
// are scaled to integers by the lcm of their denominators -- a positive row
// scaling changes no pivoting decision -- and the results are scaled back:
// slack i of a row scaled by k_i reads s_i = s'_i / k_i, its objective
// coefficient c(s_i) = c'(s'_i) * k_i / k_0, everything else divides by k_0.
int two_stage_group_q(xp_ctx *ctx, const std::vector<const Mat<Q> *> &leqs,
                      const std::vector<const Mat<Q> *> &tgs, const std::vector<int> &live,
                      uint32_t max_iter, std::vector<ResQ> &out);

// LPs that fit shared memory go out as one ragged launch; the others (c5's
// 201 x 402 root tableau) as a second one whose CTAs keep their state in a
// global-memory slab (one maxm/maxn bound per launch decides which).
int two_stage_many_q(xp_ctx *ctx, const std::vector<const Mat<Q> *> &leqs,
                     const std::vector<const Mat<Q> *> &tgs, uint32_t max_iter,
                     std::vector<ResQ> &out)
{
    const int B = (int)leqs.size();
    out.assign(B, ResQ());
    std::vector<int> small, big;
    for (int k = 0; k < B; k++) {
        const Mat<Q> &L = *leqs[k];
        if (L.r < 1 || L.c < 2) {
            out[k].status = XP_ERR_BAD_ARG;
            continue;
        }
        (fits_smem(ctx, L.r, L.c - 1, 24) ? small : big).push_back(k);
    }
    int rc = two_stage_group_q(ctx, leqs, tgs, small, max_iter, out);
    if (rc) return rc;
    return two_stage_group_q(ctx, leqs, tgs, big, max_iter, out);
}

int two_stage_group_q(xp_ctx *ctx, const std::vector<const Mat<Q> *> &leqs,
                      const std::vector<const Mat<Q> *> &tgs, const std::vector<int> &live,
                      uint32_t max_iter, std::vector<ResQ> &out)
{
    const int B = (int)leqs.size();
    std::vector<int32_t> ms(B), ns(B);
    std::vector<int64_t> lo(B), to(B);
    std::vector<std::vector<long long>> scale(B); // k_1..k_m then k_0
    size_t ll = 0, tl = 0;
    int ldo = 0, ldm = 0;
    const int S = (int)live.size();
    if (S == 0) return 0;
    for (int s = 0; s < S; s++) {
        const Mat<Q> &L = *leqs[live[s]];
        ms[s] = L.r;
        ns[s] = L.c - 1;
        lo[s] = (int64_t)ll;
        to[s] = (int64_t)tl;
        ll += L.a.size();
        tl += (size_t)L.c;
        ldo = std::max(ldo, L.r + L.c);
        ldm = std::max(ldm, L.r);
    }
    std::vector<int64_t> lp(ll), tp(tl);
    std::vector<char> iovf(S, 0);
    for_items<Q>((size_t)S, iovf, [&](size_t s) {
        const Mat<Q> &L = *leqs[live[s]], &T = *tgs[live[s]];
        std::vector<long long> &sc = scale[live[s]];
        sc.assign(L.r + 1, 1);
        bool ovf = false;
        for (int i = 0; i < L.r; i++) {
            long long k = 1;
            for (int j = 0; j < L.c; j++) k = lcm_ll(k, L.at(i, j).den, ovf);
            sc[i] = k;
            for (int j = 0; j < L.c; j++) {
                __int128 v = (__int128)L.at(i, j).num * (k / L.at(i, j).den);
                if (v > (__int128)0x7fffffffffffffffLL || v < -(__int128)0x7fffffffffffffffLL) ovf = true;
                lp[lo[s] + (size_t)i * L.c + j] = (int64_t)v;
            }
        }
        long long k0 = 1;
        for (int j = 0; j < T.c; j++) k0 = lcm_ll(k0, T.at(0, j).den, ovf);
        sc[L.r] = k0;
        for (int j = 0; j < T.c; j++) {
            __int128 v = (__int128)T.at(0, j).num * (k0 / T.at(0, j).den);
            if (v > (__int128)0x7fffffffffffffffLL || v < -(__int128)0x7fffffffffffffffLL) ovf = true;
            tp[to[s] + j] = (int64_t)v;
        }
        if (ovf) out[live[s]].status = XP_ERR_OVERFLOW;
    });
    std::vector<int64_t> maxv((size_t)S * 2), sn((size_t)S * ldo), sd((size_t)S * ldo),
        tn((size_t)S * ldo), td((size_t)S * ldo);
    std::vector<int32_t> status(S), e2b((size_t)S * ldm);
    int rc = xp_six_two_stage_i64_ragged(ctx, S, ms.data(), ns.data(), lo.data(), to.data(),
                                         lp.data(), ll, tp.data(), tl, max_iter, XP_RULE_REFERENCE,
                                         ldo, ldm, status.data(), maxv.data(), sn.data(), sd.data(),
                                         tn.data(), td.data(), e2b.data(), nullptr, nullptr);
    if (rc) return rc;
    for_items<Q>((size_t)S, iovf, [&](size_t s) {
        ResQ &R = out[live[s]];
        if (R.status == XP_ERR_OVERFLOW) return; // input scaling already overflowed
        const int m = ms[s], n = ns[s], Cm = m + n + 1;
        const std::vector<long long> &sc = scale[live[s]];
        const Q::T k0 = Q::from_int(sc[m]);
        R.status = status[s];
        R.maxv = Q::div(Q::make(maxv[2 * (size_t)s], maxv[2 * (size_t)s + 1]), k0);
        R.slack_sol.resize(Cm);
        R.tgtf.resize(Cm);
        for (int j = 0; j < Cm; j++) {
            const size_t o = (size_t)s * ldo + j;
            Q::T sv = Q::make(sn[o], sd[o] ? sd[o] : 1), tv = Q::make(tn[o], td[o] ? td[o] : 1);
            if (j >= n && j < n + m) {
                const Q::T ki = Q::from_int(sc[j - n]);
                sv = Q::div(sv, ki);
                tv = Q::mul(tv, ki);
            }
            R.slack_sol[j] = sv;
            R.tgtf[j] = Q::div(tv, k0);
        }
        R.eq2bv.assign(e2b.begin() + (size_t)s * ldm, e2b.begin() + (size_t)s * ldm + m);
        if (Q::overflow()) R.status = XP_ERR_OVERFLOW; // a scaled-back value left int64
    });
    return 0;
}

// ---------------------------------------------------------------- plumbing
Mat<F64> mat_f64(int r, int c, const double *src)
{
    Mat<F64> M(r, c);
    if (src) std::copy(src, src + (size_t)r * c, M.a.begin());
    return M;
}
Mat<Q> mat_q(int r, int c, const xp_rat *src)
{
    Mat<Q> M(r, c);
    if (src)
        for (size_t e = 0; e < (size_t)r * c; e++)
            M.a[e] = Q::make(src[e].num, src[e].den ? src[e].den : 1);
    return M;
}
bool to_rat(Q::T v, xp_rat *o)
{
    if (v.num > 0x7fffffffLL || v.num < -0x7fffffffLL || v.den > 0x7fffffffLL) return false;
    o->num = (int32_t)v.num;
    o->den = (int32_t)v.den;
    return true;
}

template <class P>
struct Many;
template <>
struct Many<F64> {
    static int run(xp_ctx *c, const std::vector<const Mat<F64> *> &l,
                   const std::vector<const Mat<F64> *> &t, uint32_t it, std::vector<ResF> &o)
    {
        return two_stage_many_f64(c, l, t, it, o);
    }
};
template <>
struct Many<Q> {
    static int run(xp_ctx *c, const std::vector<const Mat<Q> *> &l,
                   const std::vector<const Mat<Q> *> &t, uint32_t it, std::vector<ResQ> &o)
    {
        return two_stage_many_q(c, l, t, it, o);
    }
};

// General variable constraints (lpsol.h:798-802 reads vc(i,i) and vc(i,rhs) in the feasibility
// check of the optimal exit, nothing else).  FP64: the HBM-resident path takes the two vectors
// and checks them on the device.  Exact: the row-sum half of is_feasible is an identity, so the
// verdict of the batched kernel (which assumes -x <= 0) is re-decided on the solution row.
int two_stage_vc(xp_ctx *ctx, const SixJob<F64> &job, uint32_t max_iter, std::vector<ResF> &R)
{
    R.assign(1, ResF());
    return two_stage_large_f64(ctx, job.lp_leq, job.lp_tgtf, max_iter, R[0], &job.N.vc_diag, &job.N.vc_rhs);
}
int two_stage_vc(xp_ctx *ctx, const SixJob<Q> &job, uint32_t max_iter, std::vector<ResQ> &R)
{
    std::vector<const Mat<Q> *> l{&job.lp_leq}, t{&job.lp_tgtf};
    int rc = two_stage_many_q(ctx, l, t, max_iter, R);
    if (rc) return rc;
    if (R[0].status == XP_SIX_SUCC || R[0].status == XP_SIX_OPTIMAL_IS_INFEASIBLE) {
        const int rhs = job.lp_leq.r + job.lp_leq.c - 1;
        const bool bad = violates_vc<Q>(R[0].slack_sol, rhs, job.N.vc_diag, job.N.vc_rhs);
        R[0].status = bad ? XP_SIX_OPTIMAL_IS_INFEASIBLE : XP_SIX_SUCC;
        R[0].maxv = bad ? Q::zero() : R[0].tgtf[rhs]; // maxv = tgtf[rhs] on success (:1119), 0 otherwise
    }
    return 0;
}

// SIX::minm of an LP beyond shared memory whose normalisation is the identity (no equalities,
// every variable constrained by -x <= 0): the explicit dual (calcDualMaxm, lpsol.h:1585-1655)
// is built ON THE DEVICE from the caller's primal -- no host copy, no host transposition of the
// constraint matrix.  Returns false if the LP does not qualify (the general host path runs).
bool minm_dual_on_device(xp_ctx *, const Mat<Q> &, const Mat<Q> &, const Mat<Q> &, const Mat<Q> &, uint32_t, int &,
                         Q::T &, std::vector<Q::T> &, std::vector<int32_t> *)
{
    return false;
}
bool minm_dual_on_device(xp_ctx *ctx, const Mat<F64> &tg, const Mat<F64> &vc, const Mat<F64> &eq,
                         const Mat<F64> &leq, uint32_t max_iter, int &st_out, double &v, std::vector<double> &sol,
                         std::vector<int32_t> *eq2bv)
{
    if (!eq.empty() || leq.empty() || getenv("XP_HOST_DUAL")) return false;
    const int m = leq.r, n = leq.c - 1;
    if (fits_smem(ctx, n, m, 16)) return false; // the dual has n rows and m variables
    if (vc.r != n || vc.c != n + 1) return false;
    for (int i = 0; i < n; i++) // -x_i <= 0 and nothing else in row / column i
        for (int j = 0; j <= n; j++)
            if (j == i ? !(vc.at(i, i) < 0.0) : vc.at(i, j) != 0.0) return false;
    SixJob<F64> job; // only what finish() reads for a min problem (:1713-1716)
    job.is_min = true;
    job.tgtf_orig = tg;
    job.m_rhs = n;
    job.dn = m;
    job.dm = n;
    job.N.rhs_idx = n;
    ResF R;
    R.slack_sol.assign((size_t)m + n + 1, 0.0);
    R.tgtf.assign((size_t)m + n + 1, 0.0);
    R.eq2bv.assign(n, 0);
    R.maxv = 0.0;
    int32_t st = 0;
    int rc = xp_six_two_stage_f64_large_dual(ctx, m, n, leq.a.data(), tg.a.data(), max_iter, XP_RULE_REFERENCE, &st,
                                             &R.maxv, R.slack_sol.data(), R.tgtf.data(), R.eq2bv.data(), nullptr,
                                             nullptr);
    if (rc) {
        st_out = rc;
        return true;
    }
    R.status = st;
    if (eq2bv) *eq2bv = R.eq2bv;
    st_out = job.finish(R, v, sol);
    return true;
}

// One SIX::maxm / minm.
template <class P>
int solve_one(xp_ctx *ctx, bool is_min, const Mat<P> &tg, const Mat<P> &vc, const Mat<P> &eq,
              const Mat<P> &leq, uint32_t max_iter, typename P::T &v,
              std::vector<typename P::T> &sol, std::vector<int32_t> *eq2bv)
{
    if (is_min) {
        int st_dev = 0;
        v = P::zero();
        if (minm_dual_on_device(ctx, tg, vc, eq, leq, max_iter, st_dev, v, sol, eq2bv)) return st_dev;
    }
    SixJob<P> job;
    int st = job.prepare(is_min, tg, vc, eq, leq);
    v = P::zero();
    if (st) return st;
    std::vector<TwoStageResult<P>> R;
    std::vector<const Mat<P> *> l{&job.lp_leq}, t{&job.lp_tgtf};
    // (minm solves the explicit dual, whose variable constraints are its own -I, :1630-1636)
    int rc = (!job.N.std_vc && !is_min) ? two_stage_vc(ctx, job, max_iter, R) : Many<P>::run(ctx, l, t, max_iter, R);
    if (rc) return rc;
    if (eq2bv) *eq2bv = R[0].eq2bv;
    return job.finish(R[0], v, sol);
}

// A set of B&B trees advanced in lockstep: one batched GPU call per wave of node
// relaxations, decisions replayed per tree in DFS order (MipTree).
template <class P>
int run_trees(xp_ctx *ctx, std::vector<MipTree<P>> &trees)
{
    const size_t NT = trees.size();
    std::vector<char> tovf(NT, 0); // per-tree overflow state of the exact policy
    std::vector<SixJob<P>> slot(NT);
    std::vector<char> has(NT);
    for (;;) {
        // host preparation of every live tree's next node LP (normalize, dual), in parallel
        for_items<P>(NT, tovf, [&](size_t t) {
            has[t] = 0;
            for (;;) { // node LPs whose host preparation already fails are fed back at once
                if (trees[t].done) return;
                const typename MipTree<P>::Frame &f = trees[t].top();
                SixJob<P> job;
                int st = job.prepare(!trees[t].is_max, trees[t].tgtf, trees[t].vc, f.eq, f.leq);
                if (st == 0) {
                    slot[t] = std::move(job);
                    has[t] = 1;
                    return;
                }
                trees[t].feed(st, P::zero(), std::vector<typename P::T>());
            }
        });
        std::vector<int> act;
        for (size_t t = 0; t < NT; t++)
            if (has[t]) act.push_back((int)t);
        if (act.empty()) break;
        std::vector<const Mat<P> *> l, tg;
        for (int t : act) {
            l.push_back(&slot[t].lp_leq);
            tg.push_back(&slot[t].lp_tgtf);
        }
        std::vector<TwoStageResult<P>> R;
        int rc = Many<P>::run(ctx, l, tg, 10000u, R); // six.set_param(m_indent, 10000), :2441
        if (rc) return rc;
        // final solution of every node LP and the tree's accept / prune decision, in parallel
        // (trees are independent; within a tree the order is the reference's DFS order)
        std::vector<char> aovf(act.size());
        for (size_t k = 0; k < act.size(); k++) aovf[k] = tovf[act[k]];
        for_items<P>(act.size(), aovf, [&](size_t k) {
            typename P::T v;
            std::vector<typename P::T> sol;
            int st = slot[act[k]].finish(R[k], v, sol);
            trees[act[k]].feed(st, v, sol);
        });
        for (size_t k = 0; k < act.size(); k++) tovf[act[k]] = aovf[k];
    }
    for (size_t t = 0; t < NT; t++)
        if (tovf[t]) P::overflow() = true;
    return 0;
}

int mip_status(int st) { return st; }

} // namespace

// ============================================================== C entry points
#define XP_ENTRY_GUARD(ctx) \
    if (!(ctx)) return XP_ERR_BAD_ARG; \
    XP_CUDA_OK(ctx, cudaSetDevice((ctx)->device)); \
    Q::overflow() = false;

// eq2bv_out holds the basis of the LP that was actually solved: the normalised primal
// (m + 2k rows) for maxm, its explicit dual (one row per normalised variable: n plus the
// free-variable splits, at most 2n) for minm.  The documented capacities (m + 2k / 2n,
// include/xpoly_b200.h) bound the copy; entries past the solved LP's row count are untouched.
static void copy_basis(const std::vector<int32_t> &e2b, bool is_min, int m, int n, int k, int32_t *out)
{
    if (!out) return;
    const size_t cap = is_min ? (size_t)2 * n : (size_t)m + 2 * (size_t)k;
    std::copy(e2b.begin(), e2b.begin() + (e2b.size() < cap ? e2b.size() : cap), out);
}

static int six_f64(xp_ctx *ctx, bool is_min, int m, int n, const double *tgtf, const double *vc,
                   int k, const double *eq, const double *leq, uint32_t max_iter, double *v,
                   double *sol, int32_t *eq2bv_out)
{
    XP_ENTRY_GUARD(ctx);
    if (n < 1 || (m < 1 && k < 1) || !tgtf || !v) return XP_ERR_BAD_ARG;
    Mat<F64> T = mat_f64(1, n + 1, tgtf), V = vc ? mat_f64(n, n + 1, vc) : default_vc<F64>(n);
    Mat<F64> E = k > 0 ? mat_f64(k, n + 1, eq) : Mat<F64>(), L = m > 0 ? mat_f64(m, n + 1, leq) : Mat<F64>();
    std::vector<double> s;
    std::vector<int32_t> e2b;
    int st = solve_one<F64>(ctx, is_min, T, V, E, L, max_iter, *v, s, &e2b);
    if (st == XP_SIX_SUCC && sol) std::copy(s.begin(), s.begin() + n + 1, sol);
    copy_basis(e2b, is_min, m, n, k, eq2bv_out);
    return st;
}

extern "C" int xp_six_maxm_f64(xp_ctx *ctx, int m, int n, const double *tgtf, const double *vc,
                               int k, const double *eq, const double *leq, uint32_t max_iter,
                               double *maxv, double *sol, int32_t *eq2bv_out)
{
    return six_f64(ctx, false, m, n, tgtf, vc, k, eq, leq, max_iter, maxv, sol, eq2bv_out);
}
extern "C" int xp_six_minm_f64(xp_ctx *ctx, int m, int n, const double *tgtf, const double *vc,
                               int k, const double *eq, const double *leq, uint32_t max_iter,
                               double *minv, double *sol, int32_t *eq2bv_out)
{
    return six_f64(ctx, true, m, n, tgtf, vc, k, eq, leq, max_iter, minv, sol, eq2bv_out);
}

static int six_rat(xp_ctx *ctx, bool is_min, int m, int n, const xp_rat *tgtf, const xp_rat *vc,
                   int k, const xp_rat *eq, const xp_rat *leq, uint32_t max_iter, xp_rat *v,
                   xp_rat *sol, int32_t *eq2bv_out)
{
    XP_ENTRY_GUARD(ctx);
    if (n < 1 || (m < 1 && k < 1) || !tgtf || !v) return XP_ERR_BAD_ARG;
    Mat<Q> T = mat_q(1, n + 1, tgtf), V = vc ? mat_q(n, n + 1, vc) : default_vc<Q>(n);
    Mat<Q> E = k > 0 ? mat_q(k, n + 1, eq) : Mat<Q>(), L = m > 0 ? mat_q(m, n + 1, leq) : Mat<Q>();
    Q::T val;
    std::vector<Q::T> s;
    std::vector<int32_t> e2b;
    int st = solve_one<Q>(ctx, is_min, T, V, E, L, max_iter, val, s, &e2b);
    v->num = 0;
    v->den = 1;
    if (st == XP_SIX_SUCC) {
        bool ok = to_rat(val, v);
        for (int j = 0; j <= n && sol; j++) ok &= to_rat(s[j], &sol[j]);
        if (!ok) return XP_ERR_OVERFLOW;
    }
    copy_basis(e2b, is_min, m, n, k, eq2bv_out);
    return st;
}

extern "C" int xp_six_maxm_rat(xp_ctx *ctx, int m, int n, const xp_rat *tgtf, const xp_rat *vc,
                               int k, const xp_rat *eq, const xp_rat *leq, uint32_t max_iter,
                               xp_rat *maxv, xp_rat *sol, int32_t *eq2bv_out)
{
    return six_rat(ctx, false, m, n, tgtf, vc, k, eq, leq, max_iter, maxv, sol, eq2bv_out);
}
extern "C" int xp_six_minm_rat(xp_ctx *ctx, int m, int n, const xp_rat *tgtf, const xp_rat *vc,
                               int k, const xp_rat *eq, const xp_rat *leq, uint32_t max_iter,
                               xp_rat *minv, xp_rat *sol, int32_t *eq2bv_out)
{
    return six_rat(ctx, true, m, n, tgtf, vc, k, eq, leq, max_iter, minv, sol, eq2bv_out);
}

// ---- batched entry level: uniform shape, vc = -I, no equalities ----
template <class P, class In, class Out, class MK, class WR>
static int solve_batch(xp_ctx *ctx, int is_min, int batch, int m, int n, const In *tgtf,
                       const In *leq, uint32_t max_iter, int32_t *status, Out *v, Out *sol, MK mk,
                       WR wr)
{
    XP_ENTRY_GUARD(ctx);
    if (batch < 0 || m < 1 || n < 1 || !tgtf || !leq) return XP_ERR_BAD_ARG;
    const Mat<P> V = default_vc<P>(n), E;
    std::vector<SixJob<P>> jobs(batch);
    std::vector<int> st0(batch);
    std::vector<const Mat<P> *> l, t;
    std::vector<int> idx;
    for (int b = 0; b < batch; b++) {
        Mat<P> T = mk(1, n + 1, tgtf + (size_t)b * (n + 1));
        Mat<P> L = mk(m, n + 1, leq + (size_t)b * m * (n + 1));
        st0[b] = jobs[b].prepare(is_min != 0, T, V, E, L);
        if (st0[b] == 0) {
            l.push_back(&jobs[b].lp_leq);
            t.push_back(&jobs[b].lp_tgtf);
            idx.push_back(b);
        }
    }
    std::vector<TwoStageResult<P>> R;
    int rc = Many<P>::run(ctx, l, t, max_iter, R);
    if (rc) return rc;
    for (int b = 0; b < batch; b++)
        if (st0[b] && status) status[b] = st0[b];
    for (size_t k = 0; k < idx.size(); k++) {
        const int b = idx[k];
        typename P::T val;
        std::vector<typename P::T> s;
        int st = jobs[b].finish(R[k], val, s);
        st = wr(st, val, s, v ? v + b : nullptr, sol ? sol + (size_t)b * (n + 1) : nullptr, n + 1);
        if (status) status[b] = st;
    }
    return 0;
}

extern "C" int xp_six_solve_f64_batch(xp_ctx *ctx, int is_min, int batch, int m, int n,
                                      const double *tgtf, const double *leq, uint32_t max_iter,
                                      int32_t *status, double *v, double *sol)
{
    return solve_batch<F64, double, double>(
        ctx, is_min, batch, m, n, tgtf, leq, max_iter, status, v, sol, mat_f64,
        [](int st, double val, const std::vector<double> &s, double *vo, double *so, int n1) {
            if (vo) *vo = val;
            if (st == XP_SIX_SUCC && so) std::copy(s.begin(), s.begin() + n1, so);
            return st;
        });
}

extern "C" int xp_six_solve_rat_batch(xp_ctx *ctx, int is_min, int batch, int m, int n,
                                      const xp_rat *tgtf, const xp_rat *leq, uint32_t max_iter,
                                      int32_t *status, xp_rat *v, xp_rat *sol)
{
    return solve_batch<Q, xp_rat, xp_rat>(
        ctx, is_min, batch, m, n, tgtf, leq, max_iter, status, v, sol, mat_q,
        [](int st, Q::T val, const std::vector<Q::T> &s, xp_rat *vo, xp_rat *so, int n1) {
            bool ok = true;
            if (vo) {
                vo->num = 0;
                vo->den = 1;
                if (st == XP_SIX_SUCC) ok &= to_rat(val, vo);
            }
            if (st == XP_SIX_SUCC && so)
                for (int j = 0; j < n1; j++) ok &= to_rat(s[j], so + j);
            return ok ? st : XP_ERR_OVERFLOW;
        });
}

// ---- MIP ----
extern "C" int xp_mip_solve_rat(xp_ctx *ctx, int is_min, int is_bin, int m, int n,
                                const xp_rat *tgtf, int k, const xp_rat *eq, const xp_rat *leq,
                                xp_rat *v, xp_rat *sol, int32_t *n_nodes)
{
    return xp_mip_solve_rat_ri(ctx, is_min, is_bin, m, n, tgtf, k, eq, leq, nullptr, v, sol, n_nodes);
}

extern "C" int xp_mip_solve_rat_ri(xp_ctx *ctx, int is_min, int is_bin, int m, int n,
                                   const xp_rat *tgtf, int k, const xp_rat *eq, const xp_rat *leq,
                                   const uint8_t *rational_indicator, xp_rat *v, xp_rat *sol,
                                   int32_t *n_nodes)
{
    XP_ENTRY_GUARD(ctx);
    if (n < 1 || (m < 1 && k < 1) || !tgtf || !v) return XP_ERR_BAD_ARG;
    std::vector<MipTree<Q>> trees(1);
    if (rational_indicator) trees[0].allow_rational.assign(rational_indicator, rational_indicator + n + 1);
    trees[0].start(mat_q(1, n + 1, tgtf), default_vc<Q>(n), k > 0 ? mat_q(k, n + 1, eq) : Mat<Q>(),
                   m > 0 ? mat_q(m, n + 1, leq) : Mat<Q>(), !is_min, is_bin != 0);
    int rc = run_trees<Q>(ctx, trees);
    if (rc) return rc;
    if (n_nodes) *n_nodes = trees[0].nodes;
    int st = trees[0].ret_status;
    // RecusivePart shares one `v` down the whole recursion (lpsol.h:2426-2447), so on a
    // failing status the caller still sees the value of the last node LP solved.
    v->num = 0;
    v->den = 1;
    bool ok = to_rat(trees[0].v, v);
    if (st == XP_IP_SUCC) {
        for (int j = 0; j <= n && sol; j++) ok &= to_rat(trees[0].sol[j], &sol[j]);
        if (!ok || !Q::ok()) return XP_ERR_OVERFLOW;
    }
    return mip_status(st);
}

extern "C" int xp_mip_solve_f64(xp_ctx *ctx, int is_min, int is_bin, int m, int n,
                                const double *tgtf, int k, const double *eq, const double *leq,
                                double *v, double *sol, int32_t *n_nodes)
{
    XP_ENTRY_GUARD(ctx);
    if (n < 1 || (m < 1 && k < 1) || !tgtf || !v) return XP_ERR_BAD_ARG;
    std::vector<MipTree<F64>> trees(1);
    trees[0].start(mat_f64(1, n + 1, tgtf), default_vc<F64>(n),
                   k > 0 ? mat_f64(k, n + 1, eq) : Mat<F64>(),
                   m > 0 ? mat_f64(m, n + 1, leq) : Mat<F64>(), !is_min, is_bin != 0);
    int rc = run_trees<F64>(ctx, trees);
    if (rc) return rc;
    if (n_nodes) *n_nodes = trees[0].nodes;
    int st = trees[0].ret_status;
    *v = trees[0].v;
    if (st == XP_IP_SUCC && sol) std::copy(trees[0].sol.begin(), trees[0].sol.begin() + n + 1, sol);
    return st;
}

extern "C" int xp_mip_solve_rat_batch(xp_ctx *ctx, int is_min, int is_bin, int batch, int m, int n,
                                      const xp_rat *tgtf, const xp_rat *leq, int32_t *status,
                                      xp_rat *v, xp_rat *sol, int32_t *n_nodes)
{
    XP_ENTRY_GUARD(ctx);
    if (batch < 0 || m < 1 || n < 1 || !tgtf || !leq) return XP_ERR_BAD_ARG;
    std::vector<MipTree<Q>> trees(batch);
    for (int b = 0; b < batch; b++)
        trees[b].start(mat_q(1, n + 1, tgtf + (size_t)b * (n + 1)), default_vc<Q>(n), Mat<Q>(),
                       mat_q(m, n + 1, leq + (size_t)b * m * (n + 1)), !is_min, is_bin != 0);
    int rc = run_trees<Q>(ctx, trees);
    if (rc) return rc;
    for (int b = 0; b < batch; b++) {
        int st = trees[b].ret_status;
        if (n_nodes) n_nodes[b] = trees[b].nodes;
        bool vok = true;
        if (v) {
            v[b].num = 0;
            v[b].den = 1;
            vok = to_rat(trees[b].v, &v[b]);
        }
        if (st == XP_IP_SUCC) {
            bool ok = vok;
            for (int j = 0; j <= n && sol; j++) ok &= to_rat(trees[b].sol[j], &sol[(size_t)b * (n + 1) + j]);
            if (!ok) st = XP_ERR_OVERFLOW;
        }
        if (status) status[b] = st;
    }
    return 0;
}

// ---- Lineq::has_solution (linsys.cpp:830-906), batched ----
int xp_has_solution_device(xp_ctx *ctx, const std::vector<int> &sel, const int32_t *ns, const int32_t *ms,
                           const int64_t *leq_off, const xp_rat *leq_pool, int is_int_sol, int is_unique_sol,
                           int32_t *result);
bool xp_has_solution_device_fits(int n, int m, int k, const xp_rat *leq);

namespace {

// Systems Ls[b] (leq, may be empty) / Es[b] (eq, may be empty) over ns[b] variables, vc = -I.
// result[b] = 1 / 0, or a negative code for a system on which the reference itself would run
// into undefined behaviour (the other systems are still answered).
int has_solution_core(xp_ctx *ctx, const std::vector<Mat<Q>> &Ls, const std::vector<Mat<Q>> &Es,
                      const std::vector<int> &ns, int is_int_sol, int is_unique_sol, int32_t *result)
{
    const int batch = (int)Ls.size();
    std::vector<Mat<Q>> Ts(batch), Vs(batch);
    std::vector<int> pending;
    for (int b = 0; b < batch; b++) {
        result[b] = 0;
        if (Ls[b].empty() && Es[b].empty()) continue; // linsys.cpp:845-847
        if (Ls[b].empty()) { // tgtf(1, leq.get_col_size()) is 1 x 0 and is then written, :851-854
            result[b] = XP_ERR_REFERENCE_UB;
            continue;
        }
        const int n = ns[b];
        Vs[b] = default_vc<Q>(n);
        Ts[b] = Mat<Q>(1, n + 1); // all-ones objective, reviseTargetFunc (lpsol.h:2052-2074)
        for (int j = 0; j < n; j++) {
            bool nz = !Ls[b].col_all_eq(j, Q::zero());
            if (!Es[b].empty() && !Es[b].col_all_eq(j, Q::zero())) nz = true;
            if (nz) Ts[b].at(0, j) = Q::from_int(1);
        }
        pending.push_back(b);
    }
    for (int pass = 0; pass < 2 && !pending.empty(); pass++) { // max first, then min
        const bool is_min = pass == 1;
        std::vector<int> st(pending.size(), 0);
        if (is_int_sol) {
            std::vector<MipTree<Q>> trees(pending.size());
            for (size_t k = 0; k < pending.size(); k++) {
                const int b = pending[k];
                trees[k].start(Ts[b], Vs[b], Es[b], Ls[b], !is_min, false);
            }
            int rc = run_trees<Q>(ctx, trees);
            if (rc) return rc;
            for (size_t k = 0; k < pending.size(); k++) st[k] = trees[k].ret_status;
        } else {
            std::vector<SixJob<Q>> jobs(pending.size());
            std::vector<const Mat<Q> *> l, t;
            std::vector<size_t> live;
            for (size_t k = 0; k < pending.size(); k++) {
                const int b = pending[k];
                int e = jobs[k].prepare(is_min, Ts[b], Vs[b], Es[b], Ls[b]);
                if (e) {
                    st[k] = e < 0 ? e : XP_ERR_BAD_ARG;
                    continue;
                }
                live.push_back(k);
                l.push_back(&jobs[k].lp_leq);
                t.push_back(&jobs[k].lp_tgtf);
            }
            std::vector<ResQ> R;
            int rc = two_stage_many_q(ctx, l, t, XP_NO_ITER_LIMIT, R);
            if (rc) return rc;
            for (size_t i = 0; i < live.size(); i++) {
                Q::T val;
                std::vector<Q::T> sv;
                st[live[i]] = jobs[live[i]].finish(R[i], val, sv);
            }
        }
        std::vector<int> next;
        for (size_t k = 0; k < pending.size(); k++) {
            const int b = pending[k];
            if (st[k] < 0) result[b] = st[k];
            // IP_SUCC == SIX_SUCC == 0 and IP_UNBOUND == SIX_UNBOUND == 1
            else if (st[k] == 0 || (!is_unique_sol && st[k] == 1)) result[b] = 1;
            else next.push_back(b);
        }
        pending.swap(next);
    }
    return 0;
}

} // namespace

extern "C" int xp_has_solution_rat_batch(xp_ctx *ctx, int batch, int m, int n, const xp_rat *leq,
                                         int is_int_sol, int is_unique_sol, int32_t *result)
{
    XP_ENTRY_GUARD(ctx);
    if (batch < 0 || m < 1 || n < 1 || !leq || !result) return XP_ERR_BAD_ARG;
    std::vector<int32_t> ns(batch, n), ms(batch, m);
    std::vector<int64_t> off(batch);
    for (int b = 0; b < batch; b++) off[b] = (int64_t)b * m * (n + 1);
    int rc = xp_has_solution_rat_ragged(ctx, batch, ns.data(), ms.data(), off.data(), leq, (size_t)batch * m * (n + 1),
                                        nullptr, nullptr, nullptr, 0, is_int_sol, is_unique_sol, result);
    if (rc) return rc;
    for (int b = 0; b < batch; b++)
        if (result[b] < 0) return result[b]; // uniform API: any failing system fails the call
    return 0;
}

// The real caller shape (DepPolyMgr::buildDepPoly, poly.cpp:1166-1195: one query per
// reference pair and loop depth): systems of different sizes, inequalities and equalities.
extern "C" int xp_has_solution_rat_ragged(xp_ctx *ctx, int batch, const int32_t *ns,
                                          const int32_t *ms, const int64_t *leq_off,
                                          const xp_rat *leq_pool, size_t leq_pool_len,
                                          const int32_t *ks, const int64_t *eq_off,
                                          const xp_rat *eq_pool, size_t eq_pool_len,
                                          int is_int_sol, int is_unique_sol, int32_t *result)
{
    XP_ENTRY_GUARD(ctx);
    if (batch < 0 || !ns || !ms || !result) return XP_ERR_BAD_ARG;
    for (int b = 0; b < batch; b++) {
        const int n = ns[b], m = ms[b], k = ks ? ks[b] : 0;
        if (n < 1 || m < 0 || k < 0) return XP_ERR_BAD_ARG;
        if ((m > 0 && (!leq_off || !leq_pool)) || (k > 0 && (!eq_off || !eq_pool))) return XP_ERR_BAD_ARG;
        // every system must lie inside its pool (offsets and lengths in xp_rat elements)
        if (m > 0 && (leq_off[b] < 0 || (size_t)leq_off[b] + (size_t)m * (n + 1) > leq_pool_len)) return XP_ERR_BAD_ARG;
        if (k > 0 && (eq_off[b] < 0 || (size_t)eq_off[b] + (size_t)k * (n + 1) > eq_pool_len)) return XP_ERR_BAD_ARG;
    }
    // Queries of the dependence-feasibility shape (integer rows, no equalities, small) are answered
    // whole on the device, one warp per query (xp_has_solution_dev.cu); the others -- equalities,
    // rational inputs, larger systems -- and any query that overflowed int64 there go through the
    // lock-step path below (host B&B state machine, batched node relaxations).
    std::vector<int> dev_sel, host_sel;
    const bool no_dev = getenv("XP_HS_HOST") != nullptr;
    for (int b = 0; b < batch; b++) {
        const int k = ks ? ks[b] : 0;
        if (!no_dev && ms[b] > 0 && xp_has_solution_device_fits(ns[b], ms[b], k, leq_pool + leq_off[b])) dev_sel.push_back(b);
        else host_sel.push_back(b);
    }
    if (!dev_sel.empty()) {
        int rc = xp_has_solution_device(ctx, dev_sel, ns, ms, leq_off, leq_pool, is_int_sol, is_unique_sol, result);
        if (rc) return rc;
        for (int b : dev_sel)
            if (result[b] == XP_ERR_OVERFLOW || result[b] == XP_ERR_TOO_LARGE) host_sel.push_back(b);
    }
    if (host_sel.empty()) return 0;
    const int hb = (int)host_sel.size();
    std::vector<Mat<Q>> Ls(hb), Es(hb);
    std::vector<int> nv(hb);
    for (int s = 0; s < hb; s++) {
        const int b = host_sel[s], n = ns[b], m = ms[b], k = ks ? ks[b] : 0;
        nv[s] = n;
        if (m > 0) Ls[s] = mat_q(m, n + 1, leq_pool + leq_off[b]);
        if (k > 0) Es[s] = mat_q(k, n + 1, eq_pool + eq_off[b]);
    }
    std::vector<int32_t> hres(hb, 0);
    int rc = has_solution_core(ctx, Ls, Es, nv, is_int_sol, is_unique_sol, hres.data());
    if (rc) return rc;
    for (int s = 0; s < hb; s++) result[host_sel[s]] = hres[s];
    return 0;
}
