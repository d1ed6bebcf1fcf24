// HBM-resident FP64 simplex: SIX<FloatMat,Float>::solveSlackForm on the device.
//
// Reference semantics (all in /root/reference/src/com/lpsol.h):
//   solveSlackForm :1007-1191, findPivotBV :552-663, findPivotNVandBVPair
//   :670-773, pivot :1455-1511, PivotPairTab :68-154, is_feasible :783-822.
//
// Two kernels per simplex iteration, no host round trip inside a batch:
//   k_select  (1 CTA)  pricing (+ the reference's zeroing of basic reduced
//             costs), ratio test on the already-extracted entering column, tabu
//             table upkeep, pivot-row scaling, objective-row update, basis swap
//             and a side-effect-free peek at the NEXT entering column.
//   k_sweep   (grid)   the rank-1 update a[i][j] += (-a[i][q]) * row_p[j] as a
//             128-bit row-major stream; while streaming it also extracts the
//             updated NEXT entering column and the constant column into
//             contiguous buffers, so the next ratio test never touches the
//             tableau with a strided read.
// Algorithmic HBM bytes per pivot: 2*(m+1)*C*8 (read+write of every entry).
//
// The pair-tabu table is a bit matrix (n x n bits) plus per-row / per-column
// population counters, which makes canBeNVCandidate / canBeBVCandidate O(1)
// and exactly equivalent to the reference's byte-matrix scans.
#include "xp_common.cuh"

#include <cstring>
#include <vector>

namespace {

struct LpState {
    int status;
    unsigned cnt;
    unsigned max_iter;
    int sweep_pending;
    int p;      // pivot row of the pending sweep
    int q;      // entering column of the pending sweep
    int q_next; // column the pending sweep extracts (-1: none)
    int cur;    // colbuf[cur] holds (after the pending sweep) column col_tag
    int col_tag;
    unsigned n_log;
    int infeasible;
    int pad;
    double maxv;
};

struct LpDev {
    int m, C, n; // n = rhs_idx = C-1
    int W;       // tabu words per row
    double *tab, *tgtf, *prow, *colbuf[2], *rhsbuf, *sol;
    const double *vc_diag, *vc_rhs; // may be null
    uint8_t *nvset;
    int32_t *bv2eq, *eq2bv;
    uint32_t *tabu;
    int32_t *row_cnt, *col_cnt;
    int32_t *log;
    unsigned log_cap;
    LpState *st;
};

constexpr int SEL_THREADS = 1024;
constexpr int INT_BIG = 0x7fffffff;

__device__ __forceinline__ bool tabu_get(const LpDev &d, int nv, int bv)
{
    return (d.tabu[(size_t)nv * d.W + (bv >> 5)] >> (bv & 31)) & 1u;
}

// Strided read of column j into dst (slow path only: start of a solve, after a
// disableNV retry, and inside the fallback pair search).
__device__ void gather_col(const LpDev &d, int j, double *dst)
{
    for (int i = threadIdx.x; i < d.m; i += blockDim.x) dst[i] = d.tab[(size_t)i * d.C + j];
    __syncthreads();
}

// findPivotBV (lpsol.h:552-663) on a contiguous copy of column q.
// Returns the pivot ROW or -1.
__device__ int ratio_test(const LpDev &d, int q, const double *col, XpMinIdx *shm)
{
    const int n = d.n;
    XpMinIdx best;
    best.v = 0.0;
    best.i = -1;
    for (int i = threadIdx.x; i < d.m; i += blockDim.x) { // pass 1, :571-612
        double a = col[i];
        if (xp_fle(a, 0.0)) continue;
        int bv = d.eq2bv[i];
        if (tabu_get(d, q, bv)) continue;
        if (d.col_cnt[bv] >= n - 1) continue; // !canBeBVCandidate
        XpMinIdx c;
        c.v = xp_div(d.rhsbuf[i], a);
        c.i = i;
        best = xp_better(best, c);
    }
    best = xp_block_argmin(best, shm);
    if (best.i >= 0) return best.i;
    best.v = 0.0;
    best.i = -1;
    for (int i = threadIdx.x; i < d.m; i += blockDim.x) { // pass 2, :623-658
        int bv = d.eq2bv[i];
        if (tabu_get(d, q, bv)) continue;
        if (d.col_cnt[bv] >= n - 1) continue;
        double a = col[i];
        if (xp_feq(a, 0.0)) continue;
        XpMinIdx c;
        c.v = xp_div(d.rhsbuf[i], a);
        c.i = i;
        best = xp_better(best, c);
    }
    best = xp_block_argmin(best, shm);
    return best.i;
}

// PivotPairTab::disableNV (lpsol.h:114-121) with counter upkeep.
__device__ void disable_nv(const LpDev &d, int q)
{
    const int n = d.n;
    for (int w = threadIdx.x; w < d.W; w += blockDim.x) {
        uint32_t want = 0xffffffffu;
        int base = w << 5;
        if (base + 32 > n) want = (n - base >= 32) ? 0xffffffffu : ((1u << (n - base)) - 1u);
        if ((q >> 5) == w) want &= ~(1u << (q & 31));
        uint32_t old = d.tabu[(size_t)q * d.W + w];
        uint32_t add = want & ~old;
        d.tabu[(size_t)q * d.W + w] = old | want;
        while (add) {
            int b = __ffs(add) - 1;
            add &= add - 1;
            d.col_cnt[base + b] += 1; // distinct columns per thread: no race
        }
    }
    if (threadIdx.x == 0) d.row_cnt[q] = n - 1;
    __syncthreads();
}

__global__ void __launch_bounds__(SEL_THREADS, 1) k_select(LpDev d)
{
    __shared__ XpMinIdx shm[33];
    __shared__ int shi[33];
    LpState *st = d.st;
    const int tid = threadIdx.x;
    const int n = d.n, C = d.C;

    const int status0 = st->status;
    const unsigned cnt0 = st->cnt, max_iter = st->max_iter;
    int cur = st->cur, col_tag = st->col_tag;
    __syncthreads();
    if (tid == 0) st->sweep_pending = 0;
    if (status0 != XPI_RUNNING) return;
    if (cnt0 >= max_iter) { // while (cnt < m_max_iter), :1039
        if (tid == 0) st->status = XP_SIX_TIME_OUT;
        return;
    }

    int q = -1, p = -1;
    for (;;) {
        // ---- pricing, :1054-1069 ----
        int best = INT_BIG;
        int anypos = 0;
        for (int j = tid; j < n; j += blockDim.x) {
            if (d.nvset[j] && d.tgtf[j] > 0.0) {
                anypos = 1;
                if (best == INT_BIG && d.row_cnt[j] < n - 1) best = j; // canBeNVCandidate
            }
        }
        best = xp_block_min_int(best, shi);
        anypos = __syncthreads_or(anypos);
        // basic columns scanned before the break have their reduced cost forced to 0 (:1059)
        const int zlim = best == INT_BIG ? n : best;
        for (int j = tid; j < zlim; j += blockDim.x)
            if (!d.nvset[j]) d.tgtf[j] = 0.0;
        __syncthreads();

        if (best == INT_BIG) {
            if (!anypos) { // optimal exit; feasibility is checked by k_feas_*
                if (tid == 0) st->status = XPI_OPT_PENDING;
                return;
            }
            // ---- findPivotNVandBVPair, :670-773 ----
            // Pass A: eligible c_j > 0; pass B additionally c_j == 0 (tolerant).
            // findPivotBV is pure, so the c_j > 0 columns that failed in pass A
            // are not retried in pass B (same outcome, less work).
            int found = 0;
            for (int pass = 0; pass < 2 && !found; pass++) {
                int last = -1;
                for (;;) {
                    int cand = INT_BIG;
                    for (int j = last + 1 + tid; j < n; j += blockDim.x) {
                        if (!d.nvset[j] || d.row_cnt[j] >= n - 1) continue;
                        double c = d.tgtf[j];
                        bool take = pass == 0 ? (c > 0.0) : (!(c > 0.0) && xp_feq(c, 0.0));
                        if (take) {
                            cand = j;
                            break;
                        }
                    }
                    cand = xp_block_min_int(cand, shi);
                    if (cand == INT_BIG) break;
                    gather_col(d, cand, d.colbuf[cur]);
                    col_tag = cand;
                    int r = ratio_test(d, cand, d.colbuf[cur], shm);
                    if (r >= 0) {
                        q = cand;
                        p = r;
                        found = 1;
                        break;
                    }
                    last = cand;
                }
            }
            if (!found) {
                if (tid == 0) {
                    st->status = XP_SIX_UNBOUND;
                    st->col_tag = col_tag;
                }
                return;
            }
            break;
        }
        q = best;
        if (col_tag != q) {
            gather_col(d, q, d.colbuf[cur]);
            col_tag = q;
        }
        p = ratio_test(d, q, d.colbuf[cur], shm);
        if (p >= 0) break;
        disable_nv(d, q); // :1146-1151, retry without counting an iteration
    }

    // ---- genPair (:1156) + pivot bookkeeping ----
    const double *col = d.colbuf[cur];
    const int bv = d.eq2bv[p];
    const double pv = col[p];
    const double cq = d.tgtf[q];
    __syncthreads(); // everyone has read eq2bv[p], tgtf[q] before they change
    if (tid == 0) {
        uint32_t *w = &d.tabu[(size_t)q * d.W + (bv >> 5)];
        uint32_t bit = 1u << (bv & 31);
        if (!(*w & bit)) {
            *w |= bit;
            d.row_cnt[q] += 1;
            d.col_cnt[bv] += 1;
        }
        unsigned k = st->n_log;
        if (k < d.log_cap) {
            d.log[3 * k] = q;
            d.log[3 * k + 1] = bv;
            d.log[3 * k + 2] = p;
        }
        st->n_log = k + 1;
    }
    // ---- pivot, steps on row p and the objective row (:1471-1501) ----
    const double r = xp_div(1.0, pv);
    const bool r_one = xp_feq(r, 1.0), r_zero = xp_feq(r, 0.0);
    const bool cq_zero = xp_feq(cq, 0.0), cq_one = xp_feq(cq, 1.0);
    double *rowp = d.tab + (size_t)p * C;
    for (int j = tid; j < C; j += blockDim.x) {
        double x = xp_scale(rowp[j], r, r_one, r_zero); // mulOfRow(eqnum, 1/pivot)
        rowp[j] = x;
        d.prow[j] = x;
        double t = xp_mul(x, -1.0);                     // nvexp.mul(-1)
        if (j >= n) t = -t;                             // constant column keeps its sign
        t = cq_zero ? 0.0 : (cq_one ? t : xp_mul(t, cq)); // nvexp.mul(tgtf[nv])
        d.tgtf[j] = xp_add(t, d.tgtf[j]);               // tgtf.addRowToRow
    }
    if (tid == 0) { // :1504-1510
        d.nvset[q] = 0;
        d.nvset[bv] = 1;
        d.eq2bv[p] = q;
        d.bv2eq[q] = p;
        d.bv2eq[bv] = -1;
    }
    __syncthreads();
    // ---- peek: the column the NEXT iteration will price in (no side effects) ----
    int nxt = INT_BIG;
    if (cnt0 + 1 < max_iter) {
        for (int j = tid; j < n; j += blockDim.x) {
            if (d.nvset[j] && d.tgtf[j] > 0.0 && d.row_cnt[j] < n - 1) {
                nxt = j;
                break;
            }
        }
    }
    nxt = xp_block_min_int(nxt, shi);
    if (tid == 0) {
        st->p = p;
        st->q = q;
        st->q_next = nxt == INT_BIG ? -1 : nxt;
        st->cur = cur ^ 1; // the sweep reads colbuf[cur], writes colbuf[cur^1]
        st->col_tag = nxt == INT_BIG ? -1 : nxt;
        st->cnt = cnt0 + 1;
        st->sweep_pending = 1;
    }
}

// Rank-1 update + extraction.  Each thread owns VEC adjacent columns and walks
// `rows_per_cta` rows; the pivot-row slice lives in registers, the multipliers
// -a[i][q] for the CTA's rows are staged in shared memory.
template <int VEC, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) k_sweep(LpDev d, int rows_per_cta)
{
    extern __shared__ double s_f[];
    const LpState *st = d.st;
    if (!st->sweep_pending) return;
    const int p = st->p, qn = st->q_next, C = d.C, n = d.n, m = d.m;
    const double *fcol = d.colbuf[st->cur ^ 1];
    double *ncol = d.colbuf[st->cur];
    const int r0 = blockIdx.y * rows_per_cta;
    const int r1 = min(m, r0 + rows_per_cta);
    for (int i = r0 + threadIdx.x; i < r1; i += THREADS) s_f[i - r0] = -fcol[i];
    __syncthreads();
    const int j0 = (blockIdx.x * THREADS + threadIdx.x) * VEC;
    if (j0 >= C) return;

    if (VEC == 2) {
        const double2 pr = *reinterpret_cast<const double2 *>(d.prow + j0);
        const int exq = (qn == j0) ? 0 : (qn == j0 + 1 ? 1 : -1);
        const int exr = (n == j0) ? 0 : (n == j0 + 1 ? 1 : -1);
        double *base = d.tab + j0;
        int i = r0;
        for (; i + UNROLL <= r1; i += UNROLL) {
            double2 a[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++)
                a[u] = *reinterpret_cast<const double2 *>(base + (size_t)(i + u) * C);
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const double f = s_f[i + u - r0];
                double2 v;
                v.x = xp_add(a[u].x, xp_mul(f, pr.x));
                v.y = xp_add(a[u].y, xp_mul(f, pr.y));
                if (i + u == p) v = a[u]; // row p was rewritten by k_select
                *reinterpret_cast<double2 *>(base + (size_t)(i + u) * C) = v;
                if (exq >= 0) ncol[i + u] = exq ? v.y : v.x;
                if (exr >= 0) d.rhsbuf[i + u] = exr ? v.y : v.x;
            }
        }
        for (; i < r1; i++) {
            double2 a = *reinterpret_cast<const double2 *>(base + (size_t)i * C);
            const double f = s_f[i - r0];
            double2 v;
            v.x = xp_add(a.x, xp_mul(f, pr.x));
            v.y = xp_add(a.y, xp_mul(f, pr.y));
            if (i == p) v = a;
            *reinterpret_cast<double2 *>(base + (size_t)i * C) = v;
            if (exq >= 0) ncol[i] = exq ? v.y : v.x;
            if (exr >= 0) d.rhsbuf[i] = exr ? v.y : v.x;
        }
    } else {
        const double pr = d.prow[j0];
        const bool exq = qn == j0, exr = n == j0;
        double *base = d.tab + j0;
        for (int i = r0; i < r1; i++) {
            double a = base[(size_t)i * C];
            double v = xp_add(a, xp_mul(s_f[i - r0], pr));
            if (i == p) v = a;
            base[(size_t)i * C] = v;
            if (exq) ncol[i] = v;
            if (exr) d.rhsbuf[i] = v;
        }
    }
}

// ---- optimal exit: sol + is_feasible (lpsol.h:1089-1127, :783-822) ----
__global__ void k_feas_sol(LpDev d)
{
    LpState *st = d.st;
    if (st->status != XPI_OPT_PENDING) return;
    int bad = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < d.C; j += gridDim.x * blockDim.x) {
        double s = 0.0;
        if (j < d.n && !d.nvset[j]) s = d.rhsbuf[d.bv2eq[j]];
        d.sol[j] = s;
        if (j < d.n) { // vc(i,i) * sol(i) > vc(i,rhs), :798-802
            double dg = d.vc_diag ? d.vc_diag[j] : -1.0;
            double rh = d.vc_rhs ? d.vc_rhs[j] : 0.0;
            if (xp_mul(dg, s) > rh) bad = 1;
        }
    }
    if (bad) atomicOr(&st->infeasible, 1);
}

// One thread per row: the reference's left-to-right sum (:805-809).  Terms of
// non-basic columns are exact +-0 products and cannot change the running sum,
// so only basic columns are visited (same value, bit for bit).
__global__ void k_feas_rows(LpDev d)
{
    LpState *st = d.st;
    if (st->status != XPI_OPT_PENDING) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.m) return;
    const double *row = d.tab + (size_t)i * d.C;
    double sum = 0.0;
    for (int j = 0; j < d.n; j++) {
        if (d.nvset[j]) continue;
        sum = xp_add(sum, xp_mul(row[j], d.sol[j]));
    }
    if (!xp_feq(sum, row[d.n])) atomicOr(&st->infeasible, 1);
}

__global__ void k_feas_done(LpDev d)
{
    LpState *st = d.st;
    if (st->status != XPI_OPT_PENDING) return;
    if (st->infeasible) {
        st->status = XP_SIX_OPTIMAL_IS_INFEASIBLE;
    } else {
        st->status = XP_SIX_SUCC;
        st->maxv = d.tgtf[d.n]; // :1119
    }
}

__global__ void k_init(LpDev d, unsigned max_iter, int fresh)
{
    LpState *st = d.st;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int stride = gridDim.x * blockDim.x;
    if (fresh) {
        for (size_t k = t; k < (size_t)d.n * d.W; k += stride) d.tabu[k] = 0u; // newPPT, :1021
        for (int k = t; k < d.n; k += stride) {
            d.row_cnt[k] = 0;
            d.col_cnt[k] = 0;
        }
        for (int k = t; k < d.m; k += stride) d.rhsbuf[k] = d.tab[(size_t)k * d.C + d.n];
        for (int k = t; k < d.C; k += stride) d.sol[k] = 0.0; // sol.reinit, :1028
    }
    if (t == 0) {
        if (fresh) {
            st->cnt = 0;
            st->n_log = 0;
            st->cur = 0;
            st->col_tag = -1;
            st->infeasible = 0;
            st->maxv = 0.0; // :1027
            st->status = XPI_RUNNING;
        } else if (st->status == XP_SIX_TIME_OUT && st->cnt < max_iter) {
            st->status = XPI_RUNNING; // resume after a bounded run
        }
        st->max_iter = max_iter;
        st->sweep_pending = 0;
    }
}

__global__ void k_slack_form(double *tab, double *tgtf, const double *leq, const double *tg, int m,
                             int n, uint8_t *nvset, int32_t *bv2eq, int32_t *eq2bv)
{
    // SIX::slack (lpsol.h:1405-1433) + identity basis (:1821-1841): [A | I | b]
    const int C = n + m + 1;
    size_t total = (size_t)m * C;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(e / C), j = (int)(e % C);
        double v;
        if (j < n) v = leq[(size_t)i * (n + 1) + j];
        else if (j < n + m) v = (j - n == i) ? 1.0 : 0.0;
        else v = leq[(size_t)i * (n + 1) + n];
        tab[e] = v;
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < C; j += gridDim.x * blockDim.x) {
        tgtf[j] = j < n ? tg[j] : (j < n + m ? 0.0 : tg[n]);
        if (j < n + m) {
            nvset[j] = j < n;
            bv2eq[j] = j < n ? -1 : j - n;
        }
        if (j < m) eq2bv[j] = n + j;
    }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{ // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t seed, uint64_t idx)
{
    return (double)(mix64(seed ^ mix64(idx)) >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void k_fill_synth(double *tab, double *tgtf, int m, int n, uint64_t seed, uint8_t *nvset,
                             int32_t *bv2eq, int32_t *eq2bv)
{
    // SURVEY 8(d) dense family in slack form: A_ij~U(0,1), b_i = 1+U*n, c_j~U(0,1)
    const int C = n + m + 1;
    size_t total = (size_t)m * C;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(e / C), j = (int)(e % C);
        double v;
        if (j < n) v = u01(seed, (uint64_t)i * (n + 1) + j);
        else if (j < n + m) v = (j - n == i) ? 1.0 : 0.0;
        else v = 1.0 + u01(seed, (uint64_t)i * (n + 1) + n) * n;
        tab[e] = v;
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < C; j += gridDim.x * blockDim.x) {
        tgtf[j] = j < n ? u01(seed, (uint64_t)m * (n + 1) + j) : 0.0;
        if (j < n + m) {
            nvset[j] = j < n;
            bv2eq[j] = j < n ? -1 : j - n;
        }
        if (j < m) eq2bv[j] = n + j;
    }
}

__global__ void k_checksum(const double *a, size_t nelem, unsigned long long *out)
{
    unsigned long long s = 0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nelem;
         e += (size_t)gridDim.x * blockDim.x) {
        unsigned long long b = (unsigned long long)__double_as_longlong(a[e]);
        s += mix64(b ^ mix64((uint64_t)e)); // position-keyed, order-independent sum
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

} // namespace

struct xp_lp_f64 {
    xp_ctx *ctx;
    LpDev d;
    LpState *h_st; // pinned
    double *vc_diag, *vc_rhs;
    // optional per-launch timing of the sweep kernel (CUDA events on the ctx stream)
    bool profile = false;
    std::vector<cudaEvent_t> evs;
    uint64_t prof_sweeps = 0;
    unsigned cnt_at_entry = 0;
    double prof_sweep_ms = 0.0, prof_gap_ms = 0.0;
};

constexpr int PROF_MAX_SWEEPS = 4096;

static int sweep_launch(xp_lp_f64 *lp)
{
    xp_ctx *ctx = lp->ctx;
    const LpDev &d = lp->d;
    const int m = d.m, C = d.C;
    if ((C & 1) == 0) {
        constexpr int TH = 256;
        int ctiles = (C / 2 + TH - 1) / TH;
        // aim for >= 8 CTAs per SM worth of row tiles, 8..64 rows per CTA
        int want = ctx->sm_count * 8;
        int rpc = (int)(((long long)m * ctiles + want - 1) / want);
        rpc = rpc < 8 ? 8 : (rpc > 64 ? 64 : rpc);
        rpc = (rpc + 7) & ~7;
        dim3 grid(ctiles, (m + rpc - 1) / rpc);
        k_sweep<2, TH, 8><<<grid, TH, rpc * sizeof(double), ctx->stream>>>(d, rpc);
    } else {
        constexpr int TH = 128;
        int ctiles = (C + TH - 1) / TH;
        int rpc = 16;
        dim3 grid(ctiles, (m + rpc - 1) / rpc);
        k_sweep<1, TH, 1><<<grid, TH, rpc * sizeof(double), ctx->stream>>>(d, rpc);
    }
    ctx->launches++;
    return 0;
}

extern "C" int xp_lp_f64_create(xp_ctx *ctx, int m, int C, xp_lp_f64 **out)
{
    if (!ctx || !out || m < 1 || C < 2) return XP_ERR_BAD_ARG;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    xp_lp_f64 *lp = new xp_lp_f64();
    lp->ctx = ctx;
    LpDev &d = lp->d;
    memset(&d, 0, sizeof d);
    d.m = m;
    d.C = C;
    d.n = C - 1;
    d.W = (d.n + 31) / 32;
    d.log_cap = 1u << 16;
    const size_t n = d.n;
#define ALLOC(ptr, bytes) XP_CUDA_OK(ctx, cudaMalloc((void **)&(ptr), (bytes)))
    ALLOC(d.tab, (size_t)m * C * sizeof(double));
    ALLOC(d.tgtf, C * sizeof(double));
    ALLOC(d.prow, C * sizeof(double));
    ALLOC(d.colbuf[0], m * sizeof(double));
    ALLOC(d.colbuf[1], m * sizeof(double));
    ALLOC(d.rhsbuf, m * sizeof(double));
    ALLOC(d.sol, C * sizeof(double));
    ALLOC(lp->vc_diag, n * sizeof(double));
    ALLOC(lp->vc_rhs, n * sizeof(double));
    ALLOC(d.nvset, n + 1);
    ALLOC(d.bv2eq, n * sizeof(int32_t));
    ALLOC(d.eq2bv, m * sizeof(int32_t));
    ALLOC(d.tabu, n * (size_t)d.W * sizeof(uint32_t));
    ALLOC(d.row_cnt, n * sizeof(int32_t));
    ALLOC(d.col_cnt, n * sizeof(int32_t));
    ALLOC(d.log, (size_t)d.log_cap * 3 * sizeof(int32_t));
    ALLOC(d.st, sizeof(LpState));
#undef ALLOC
    XP_CUDA_OK(ctx, cudaMemset(d.st, 0, sizeof(LpState)));
    XP_CUDA_OK(ctx, cudaMallocHost((void **)&lp->h_st, sizeof(LpState)));
    *out = lp;
    return 0;
}

extern "C" void xp_lp_f64_destroy(xp_lp_f64 *lp)
{
    if (!lp) return;
    LpDev &d = lp->d;
    cudaSetDevice(lp->ctx->device);
    cudaStreamSynchronize(lp->ctx->stream);
    void *ptrs[] = {d.tab,    d.tgtf,      d.prow,     d.colbuf[0], d.colbuf[1], d.rhsbuf,
                    d.sol,    lp->vc_diag, lp->vc_rhs, d.nvset,     d.bv2eq,     d.eq2bv,
                    d.tabu,   d.row_cnt,   d.col_cnt,  d.log,       d.st};
    for (void *p : ptrs) cudaFree(p);
    for (cudaEvent_t e : lp->evs) cudaEventDestroy(e);
    cudaFreeHost(lp->h_st);
    delete lp;
}

static int lp_reset(xp_lp_f64 *lp)
{
    xp_ctx *ctx = lp->ctx;
    k_init<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(lp->d, 0u, 1);
    ctx->launches++;
    XP_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

extern "C" int xp_lp_f64_upload(xp_lp_f64 *lp, const double *tableau, const double *tgtf,
                                const uint8_t *nvset, const uint8_t *bvset, const int32_t *bv2eq,
                                const int32_t *eq2bv, const double *vc_diag, const double *vc_rhs)
{
    if (!lp || !tableau || !tgtf || !nvset || !bv2eq || !eq2bv) return XP_ERR_BAD_ARG;
    (void)bvset; // the complement of nvset on [0, rhs_idx)
    xp_ctx *ctx = lp->ctx;
    LpDev &d = lp->d;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d.tab, tableau, (size_t)d.m * d.C * sizeof(double),
                                    cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d.tgtf, tgtf, d.C * sizeof(double), cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d.nvset, nvset, d.n, cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d.bv2eq, bv2eq, d.n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d.eq2bv, eq2bv, d.m * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    d.vc_diag = d.vc_rhs = nullptr;
    if (vc_diag) {
        XP_CUDA_OK(ctx, cudaMemcpyAsync(lp->vc_diag, vc_diag, d.n * sizeof(double),
                                        cudaMemcpyHostToDevice, s));
        d.vc_diag = lp->vc_diag;
    }
    if (vc_rhs) {
        XP_CUDA_OK(ctx, cudaMemcpyAsync(lp->vc_rhs, vc_rhs, d.n * sizeof(double),
                                        cudaMemcpyHostToDevice, s));
        d.vc_rhs = lp->vc_rhs;
    }
    return lp_reset(lp);
}

extern "C" int xp_lp_f64_upload_leq(xp_lp_f64 *lp, const double *leq, const double *tgtf, int n)
{
    if (!lp || !leq || !tgtf) return XP_ERR_BAD_ARG;
    xp_ctx *ctx = lp->ctx;
    LpDev &d = lp->d;
    const int m = d.m;
    if (d.C != n + m + 1) return XP_ERR_BAD_ARG;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    void *scr = nullptr;
    size_t bytes = ((size_t)m * (n + 1) + (n + 1)) * sizeof(double);
    int rc = xp_ctx_scratch(ctx, bytes, &scr);
    if (rc) return rc;
    double *d_leq = (double *)scr, *d_tg = d_leq + (size_t)m * (n + 1);
    cudaStream_t s = ctx->stream;
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d_leq, leq, (size_t)m * (n + 1) * sizeof(double),
                                    cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d_tg, tgtf, (n + 1) * sizeof(double), cudaMemcpyHostToDevice, s));
    k_slack_form<<<ctx->sm_count * 4, 256, 0, s>>>(d.tab, d.tgtf, d_leq, d_tg, m, n, d.nvset,
                                                   d.bv2eq, d.eq2bv);
    ctx->launches++;
    d.vc_diag = d.vc_rhs = nullptr;
    XP_CUDA_OK(ctx, cudaGetLastError());
    return lp_reset(lp);
}

extern "C" int xp_lp_f64_fill_synthetic(xp_lp_f64 *lp, uint64_t seed)
{
    if (!lp) return XP_ERR_BAD_ARG;
    xp_ctx *ctx = lp->ctx;
    LpDev &d = lp->d;
    const int m = d.m, n = d.C - 1 - m;
    if (n < 1) return XP_ERR_BAD_ARG;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    k_fill_synth<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d.tab, d.tgtf, m, n, seed, d.nvset,
                                                             d.bv2eq, d.eq2bv);
    ctx->launches++;
    d.vc_diag = d.vc_rhs = nullptr;
    XP_CUDA_OK(ctx, cudaGetLastError());
    return lp_reset(lp);
}

extern "C" int xp_lp_f64_solve(xp_lp_f64 *lp, uint32_t max_iter, int rule)
{
    if (!lp) return XP_ERR_BAD_ARG;
    if (rule != XP_RULE_REFERENCE) return XP_ERR_BAD_ARG;
    xp_ctx *ctx = lp->ctx;
    LpDev &d = lp->d;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    XP_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, s));
    k_init<<<1, 32, 0, s>>>(d, max_iter, 0);
    ctx->launches++;
    if (lp->profile) {
        XP_CUDA_OK(ctx, cudaMemcpyAsync(lp->h_st, d.st, sizeof(LpState), cudaMemcpyDeviceToHost, s));
        XP_CUDA_OK(ctx, cudaStreamSynchronize(s));
        lp->cnt_at_entry = lp->h_st->cnt;
    }
    // Each select+sweep pair is one simplex iteration; batches run without any
    // host round trip, the host only polls the status word between batches.
    int batch = 8;
    int n_prof = 0; // sweeps bracketed by events in this call
    for (;;) {
        for (int b = 0; b < batch; b++) {
            k_select<<<1, SEL_THREADS, 0, s>>>(d);
            ctx->launches++;
            const bool prof = lp->profile && n_prof < PROF_MAX_SWEEPS;
            if (prof) XP_CUDA_OK(ctx, cudaEventRecord(lp->evs[2 * n_prof], s));
            sweep_launch(lp);
            if (prof) {
                XP_CUDA_OK(ctx, cudaEventRecord(lp->evs[2 * n_prof + 1], s));
                n_prof++;
            }
        }
        XP_CUDA_OK(ctx, cudaGetLastError());
        XP_CUDA_OK(ctx, cudaMemcpyAsync(lp->h_st, d.st, sizeof(LpState), cudaMemcpyDeviceToHost, s));
        XP_CUDA_OK(ctx, cudaStreamSynchronize(s));
        if (lp->h_st->status != XPI_RUNNING) break;
        if (batch < 64) batch *= 2;
        unsigned long long left = (unsigned long long)max_iter - lp->h_st->cnt;
        if ((unsigned long long)batch > left + 1) batch = (int)(left + 1);
    }
    if (lp->profile) {
        // only the first (iterations done in this call) sweeps did real work
        long long real = (long long)lp->h_st->cnt - (long long)lp->cnt_at_entry;
        if (real > n_prof) real = n_prof;
        for (long long k = 0; k < real; k++) {
            float ms = 0.f;
            XP_CUDA_OK(ctx, cudaEventElapsedTime(&ms, lp->evs[2 * k], lp->evs[2 * k + 1]));
            lp->prof_sweep_ms += ms;
            lp->prof_sweeps++;
            if (k > 0) {
                XP_CUDA_OK(ctx, cudaEventElapsedTime(&ms, lp->evs[2 * k - 1], lp->evs[2 * k]));
                lp->prof_gap_ms += ms;
            }
        }
    }
    if (lp->h_st->status == XPI_OPT_PENDING) {
        k_feas_sol<<<ctx->sm_count, 256, 0, s>>>(d);
        k_feas_rows<<<(d.m + 127) / 128, 128, 0, s>>>(d);
        k_feas_done<<<1, 1, 0, s>>>(d);
        ctx->launches += 3;
        XP_CUDA_OK(ctx, cudaGetLastError());
        XP_CUDA_OK(ctx, cudaMemcpyAsync(lp->h_st, d.st, sizeof(LpState), cudaMemcpyDeviceToHost, s));
    }
    XP_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, s));
    XP_CUDA_OK(ctx, cudaStreamSynchronize(s));
    XP_CUDA_OK(ctx, cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev0, ctx->ev1));
    return lp->h_st->status;
}

extern "C" int xp_lp_f64_profile(xp_lp_f64 *lp, int enable)
{
    if (!lp) return XP_ERR_BAD_ARG;
    xp_ctx *ctx = lp->ctx;
    if (enable && lp->evs.empty()) {
        lp->evs.resize(2 * PROF_MAX_SWEEPS);
        for (auto &e : lp->evs) XP_CUDA_OK(ctx, cudaEventCreate(&e));
    }
    lp->profile = enable != 0;
    lp->prof_sweeps = 0;
    lp->prof_sweep_ms = lp->prof_gap_ms = 0.0;
    return 0;
}

extern "C" int xp_lp_f64_profile_read(xp_lp_f64 *lp, uint64_t *n_sweeps, double *sweep_ms,
                                      double *gap_ms)
{
    if (!lp) return XP_ERR_BAD_ARG;
    if (n_sweeps) *n_sweeps = lp->prof_sweeps;
    if (sweep_ms) *sweep_ms = lp->prof_sweep_ms;
    if (gap_ms) *gap_ms = lp->prof_gap_ms;
    return 0;
}

extern "C" int xp_lp_f64_download(xp_lp_f64 *lp, double *tableau, double *tgtf, uint8_t *nvset,
                                  uint8_t *bvset, int32_t *bv2eq, int32_t *eq2bv, double *maxv,
                                  double *sol, uint32_t *iters, int32_t *pivot_log, uint32_t log_cap)
{
    if (!lp) return XP_ERR_BAD_ARG;
    xp_ctx *ctx = lp->ctx;
    LpDev &d = lp->d;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
#define D2H(dst, src, bytes) \
    if (dst) XP_CUDA_OK(ctx, cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, s))
    D2H(tableau, d.tab, (size_t)d.m * d.C * sizeof(double));
    D2H(tgtf, d.tgtf, d.C * sizeof(double));
    D2H(nvset, d.nvset, (size_t)d.n);
    D2H(bv2eq, d.bv2eq, d.n * sizeof(int32_t));
    D2H(eq2bv, d.eq2bv, d.m * sizeof(int32_t));
    D2H(sol, d.sol, d.C * sizeof(double));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(lp->h_st, d.st, sizeof(LpState), cudaMemcpyDeviceToHost, s));
    std::vector<uint8_t> nv;
    if (bvset && !nvset) {
        nv.resize(d.n);
        XP_CUDA_OK(ctx, cudaMemcpyAsync(nv.data(), d.nvset, d.n, cudaMemcpyDeviceToHost, s));
    }
    XP_CUDA_OK(ctx, cudaStreamSynchronize(s));
    if (pivot_log && log_cap) {
        unsigned k = lp->h_st->n_log < log_cap ? lp->h_st->n_log : log_cap;
        if (k > d.log_cap) k = d.log_cap;
        XP_CUDA_OK(ctx, cudaMemcpy(pivot_log, d.log, (size_t)k * 3 * sizeof(int32_t),
                                   cudaMemcpyDeviceToHost));
    }
#undef D2H
    if (bvset) {
        const uint8_t *src = nvset ? nvset : nv.data();
        for (int j = 0; j < d.n; j++) bvset[j] = !src[j];
    }
    if (maxv) *maxv = lp->h_st->maxv;
    if (iters) *iters = lp->h_st->cnt;
    return 0;
}

extern "C" int xp_lp_f64_checksum(xp_lp_f64 *lp, uint64_t *sum_tableau, uint64_t *sum_tgtf)
{
    if (!lp) return XP_ERR_BAD_ARG;
    xp_ctx *ctx = lp->ctx;
    LpDev &d = lp->d;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    void *scr = nullptr;
    int rc = xp_ctx_scratch(ctx, 16, &scr);
    if (rc) return rc;
    unsigned long long *acc = (unsigned long long *)scr;
    XP_CUDA_OK(ctx, cudaMemsetAsync(acc, 0, 16, ctx->stream));
    k_checksum<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d.tab, (size_t)d.m * d.C, acc);
    k_checksum<<<8, 256, 0, ctx->stream>>>(d.tgtf, (size_t)d.C, acc + 1);
    ctx->launches += 2;
    unsigned long long h[2];
    XP_CUDA_OK(ctx, cudaMemcpyAsync(h, acc, 16, cudaMemcpyDeviceToHost, ctx->stream));
    XP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (sum_tableau) *sum_tableau = h[0];
    if (sum_tgtf) *sum_tgtf = h[1];
    return 0;
}

// Host-buffer convenience: the call a maintainer binds in place of
// SIX<FloatMat,Float>::solveSlackForm.
extern "C" int xp_six_slack_f64(xp_ctx *ctx, double *tableau, double *tgtf, int m, int C,
                                uint8_t *nvset, uint8_t *bvset, int32_t *bv2eq, int32_t *eq2bv,
                                const double *vc_diag, const double *vc_rhs, uint32_t max_iter,
                                int rule, double *maxv, double *sol, uint32_t *iters,
                                int32_t *pivot_log, uint32_t log_cap)
{
    if (!ctx || !tableau || !tgtf || !nvset || !bv2eq || !eq2bv) return XP_ERR_BAD_ARG;
    // Device buffers are kept on the ctx between calls of the same shape.
    xp_lp_f64 *lp = (xp_lp_f64 *)ctx->cached_lp;
    int rc = 0;
    if (!lp || lp->d.m != m || lp->d.C != C) {
        if (lp) xp_lp_f64_destroy(lp);
        ctx->cached_lp = nullptr;
        lp = nullptr;
        rc = xp_lp_f64_create(ctx, m, C, &lp);
        if (rc) return rc;
        ctx->cached_lp = lp;
    }
    rc = xp_lp_f64_upload(lp, tableau, tgtf, nvset, bvset, bv2eq, eq2bv, vc_diag, vc_rhs);
    if (rc) return rc;
    int st = xp_lp_f64_solve(lp, max_iter, rule);
    if (st >= 0) {
        rc = xp_lp_f64_download(lp, tableau, tgtf, nvset, bvset, bv2eq, eq2bv, maxv, sol, iters,
                                pivot_log, log_cap);
        if (rc) st = rc;
    }
    return st;
}

void xp_large_release_cached(xp_ctx *ctx)
{
    if (ctx->cached_lp) xp_lp_f64_destroy((xp_lp_f64 *)ctx->cached_lp);
    ctx->cached_lp = nullptr;
}
