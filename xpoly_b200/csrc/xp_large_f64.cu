// HBM-resident FP64 simplex: SIX<FloatMat,Float>::solveSlackForm on the device,
// on one GPU or column-sharded over up to 8 GPUs (one process per GPU).
//
// Reference semantics (all in /root/reference/src/com/lpsol.h):
//   solveSlackForm :1007-1191, findPivotBV :552-663, findPivotNVandBVPair
//   :670-773, pivot :1455-1511, PivotPairTab :68-154, is_feasible :783-822.
//
// Two kernels per simplex iteration, no host round trip inside a batch:
//   k_select  (1 CTA)  completes the pricing exchange, runs the ratio test on the
//             already-extracted entering column, keeps the tabu table, scales
//             the pivot row, updates the objective row and the replicated
//             constant column, swaps the basis, and prices the NEXT iteration
//             (side-effect free) so the sweep can extract that column.
//   k_sweep   (grid)   the rank-1 update a[i][j] += (-a[i][q]) * row_p[j] as a
//             128-bit row-major stream; while streaming it extracts the updated
//             NEXT entering column into a contiguous buffer (and, sharded, into
//             every peer's buffer over NVLink), so the next ratio test never
//             touches the tableau with a strided read.
// Algorithmic HBM bytes per pivot: 2*(m+1)*C*8 (read+write of every entry).
//
// Column sharding: rank g owns columns [lo_g, hi_g) of the tableau
// and of the objective row; the constant column, the basis maps and the tabu
// table are replicated and kept identical by construction (every rank takes the
// same decisions from the same bits).  The only data that cross GPUs per pivot
// are one 8-byte candidate word per rank (pricing arg-min, lowest index wins)
// and the entering column (m+1 doubles), both written straight into the peers'
// exchange blocks (CUDA IPC mappings) by the kernels themselves.
//
// The pair-tabu table is a bit matrix (n x n bits) plus per-row / per-column
// population counters, which makes canBeNVCandidate / canBeBVCandidate O(1)
// and exactly equivalent to the reference's byte-matrix scans.
#include "xp_common.cuh"

#include <cstring>
#include <vector>

namespace {

constexpr int MAXR = XP_MAX_RANKS;
constexpr int SEL_THREADS = 1024;
constexpr int PT = 8; // independent loads in flight per thread in k_select
constexpr int INT_BIG = 0x7fffffff;
constexpr unsigned long long SPIN_LIMIT = 6000000000ULL; // ~3 s of SM clocks

struct LpState {
    int status;
    unsigned cnt;
    unsigned max_iter;
    int sweep_pending;
    int p;        // pivot row of the pending sweep
    int q_next;   // column the pending sweep extracts (global index, -1: none)
    unsigned n_log;
    int infeasible;
    unsigned xseq; // candidate exchanges initiated so far
    unsigned xs;   // slow-path column fetches so far
    unsigned swp;  // sweeps issued so far (slot parity, `done` tag)
    unsigned fe;   // feasibility-chain epoch
    int fast;      // exchange #xseq is in flight and sweep #swp extracts its column
    int pad;
    double maxv;
    double tg_rhs; // replica of the objective row's constant term
};

// Exchange block of one rank (one cudaMalloc, exported over CUDA IPC).  Every
// word has a single writer; sequence numbers only grow.
struct XHdr {
    unsigned long long cand[2][MAXR]; // (seq << 32) | anypos << 31 | candidate, by rank
    unsigned long long done[MAXR];    // sweep number whose extracted column is in slot[rank]
    unsigned long long arrive[MAXR];  // slow fetch: rank reached fetch #xs
    unsigned long long xflag[MAXR];   // slow fetch: owner's column #xs is in xslot
    unsigned long long feas_in;       // feasibility chain: partial sums from rank-1 are in feas[]
    unsigned long long feas_res[MAXR]; // (epoch << 1) | infeasible, broadcast by the last rank
};
constexpr size_t XHDR_BYTES = 1024;
static_assert(sizeof(XHdr) <= XHDR_BYTES, "exchange header");

struct LpDev {
    int m, C, n;  // global shape; n = rhs_idx = C-1
    int W;        // tabu words per row
    int rank, G;  // column shard
    int col0, Cl; // first local column, local width (row stride of tab)
    int mpad;     // doubles per exchanged column (m+1 padded)
    double *tab, *tgtf, *prow, *fcol, *rhsbuf, *sol;
    const double *vc_diag, *vc_rhs; // may be null
    uint8_t *nvset;
    int32_t *bv2eq, *eq2bv;
    uint32_t *tabu;
    int32_t *row_cnt, *col_cnt;
    int32_t *log;
    unsigned log_cap;
    unsigned *feas_ctr, *sweep_ctr;
    LpState *st;
    unsigned char *xb[MAXR]; // exchange blocks: xb[rank] is local, the rest peer mappings
};

// ---- exchange-block addressing ----
__host__ __device__ __forceinline__ size_t xoff_slot(const LpDev &d, int r, int par)
{
    return XHDR_BYTES + ((size_t)(r * 2 + par) * d.mpad) * sizeof(double);
}
__host__ __device__ __forceinline__ size_t xoff_xslot(const LpDev &d)
{
    return XHDR_BYTES + ((size_t)(MAXR * 2) * d.mpad) * sizeof(double);
}
__host__ __device__ __forceinline__ size_t xoff_feas(const LpDev &d)
{
    return XHDR_BYTES + ((size_t)(MAXR * 2 + 1) * d.mpad) * sizeof(double);
}
__host__ __device__ __forceinline__ size_t xblock_bytes(const LpDev &d)
{
    return XHDR_BYTES + ((size_t)(MAXR * 2 + 2) * d.mpad) * sizeof(double);
}
// Column range of rank r: even split in units of two columns (128-bit accesses).
__host__ __device__ __forceinline__ int shard_lo(int C, int G, int r)
{
    long long pairs = (C + 1) / 2;
    long long lo = 2 * (pairs * r / G);
    return lo > C ? C : (int)lo;
}
__device__ __forceinline__ int owner_of(const LpDev &d, int j)
{
    int r = (int)(((long long)(j / 2) * d.G) / ((d.C + 1) / 2));
    while (r + 1 < d.G && shard_lo(d.C, d.G, r + 1) <= j) r++;
    while (r > 0 && shard_lo(d.C, d.G, r) > j) r--;
    return r;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }

// Threads 0..G-1 each write `w` to the word at byte offset `off` of rank t's
// block.  Callers __syncthreads() first when the word guards data.
__device__ __forceinline__ void publish(const LpDev &d, size_t off, unsigned long long w)
{
    if ((int)threadIdx.x < d.G) {
        __threadfence_system();
        st_release_sys((unsigned long long *)(d.xb[threadIdx.x] + off), w);
    }
}
// Thread t < cnt waits until pred(word t at local offset off + 8*t).  Returns
// false on timeout (block-uniform).
template <class Pred>
__device__ __forceinline__ bool wait_words(const LpDev &d, size_t off, int first, int cnt, Pred pred)
{
    int bad = 0;
    if ((int)threadIdx.x < cnt) {
        const unsigned long long *w =
            (const unsigned long long *)(d.xb[d.rank] + off) + first + threadIdx.x;
        const unsigned long long t0 = clock64();
        unsigned spins = 0;
        while (!pred(ld_acquire_sys(w))) {
            if ((++spins & 1023u) == 0 && clock64() - t0 > SPIN_LIMIT) {
                bad = 1;
                break;
            }
        }
    }
    return !__syncthreads_or(bad);
}

__device__ __forceinline__ bool tabu_get(const LpDev &d, int nv, int bv)
{
    return (d.tabu[(size_t)nv * d.W + (bv >> 5)] >> (bv & 31)) & 1u;
}

// Pricing over the local slice (lpsol.h:1054-1069): lowest eligible index with
// c_j > 0, and whether any non-basic c_j > 0 exists at all.  Only j > after.
__device__ __forceinline__ void price_local(const LpDev &d, int after, int mode, int &best, int &anypos)
{
    // mode 0: c_j > 0 (pricing / pair-search pass A); mode 1: c_j == 0 tolerant (pass B)
    const int tid = threadIdx.x;
    const int nl = min(d.Cl, d.n - d.col0); // local columns that are variables
    best = INT_BIG;
    anypos = 0;
    for (int base = 0; base < nl; base += SEL_THREADS * PT) {
        double c[PT];
        int nv[PT], rc[PT];
#pragma unroll
        for (int u = 0; u < PT; u++) {
            const int jl = base + u * SEL_THREADS + tid;
            const bool ok = jl < nl;
            const int g = d.col0 + (ok ? jl : 0);
            c[u] = ok ? d.tgtf[jl] : 0.0;
            nv[u] = ok ? d.nvset[g] : 0;
            rc[u] = ok ? d.row_cnt[g] : INT_BIG;
        }
#pragma unroll
        for (int u = 0; u < PT; u++) {
            const int g = d.col0 + base + u * SEL_THREADS + tid;
            if (!nv[u]) continue;
            const bool pos = c[u] > 0.0;
            if (pos) anypos = 1;
            const bool take = mode == 0 ? pos : (!pos && xp_feq(c[u], 0.0));
            if (take && g > after && best == INT_BIG && rc[u] < d.n - 1) best = g;
        }
    }
}

// findPivotBV (lpsol.h:552-663) on a contiguous copy of column q (col) and the
// replicated constant column.  Returns the pivot ROW or -1.
__device__ int ratio_test(const LpDev &d, int q, const double *col, XpMinIdx *shm)
{
    const int n = d.n, tid = threadIdx.x;
    for (int pass = 0; pass < 2; pass++) {
        XpMinIdx best;
        best.v = 0.0;
        best.i = -1;
        for (int base = 0; base < d.m; base += SEL_THREADS * PT) {
            double a[PT], rh[PT];
            int bv[PT];
#pragma unroll
            for (int u = 0; u < PT; u++) {
                const int i = base + u * SEL_THREADS + tid;
                const bool ok = i < d.m;
                a[u] = ok ? ld_cg(col + i) : 0.0;
                rh[u] = ok ? d.rhsbuf[i] : 0.0;
                bv[u] = ok ? d.eq2bv[i] : -1;
            }
            uint32_t tw[PT];
            int cc[PT];
#pragma unroll
            for (int u = 0; u < PT; u++) {
                // pass 1 (:571-612) takes a > 0 (tolerant), pass 2 (:623-658) any a != 0
                const bool cand = bv[u] >= 0 && (pass == 0 ? !xp_fle(a[u], 0.0) : !xp_feq(a[u], 0.0));
                tw[u] = cand ? d.tabu[(size_t)q * d.W + (bv[u] >> 5)] : 0xffffffffu;
                cc[u] = cand ? d.col_cnt[bv[u]] : INT_BIG;
                if (!cand) bv[u] = -1;
            }
#pragma unroll
            for (int u = 0; u < PT; u++) {
                if (bv[u] < 0) continue;
                if ((tw[u] >> (bv[u] & 31)) & 1u) continue; // is_handle(q, bv)
                if (cc[u] >= n - 1) continue;               // !canBeBVCandidate
                XpMinIdx c;
                c.v = xp_div(rh[u], a[u]);
                c.i = base + u * SEL_THREADS + tid;
                best = xp_better(best, c);
            }
        }
        best = xp_block_argmin(best, shm);
        if (best.i >= 0) return best.i;
    }
    return -1;
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

struct SelCtx {
    unsigned xseq, xs;
    bool ok; // false after a peer timeout
};

// All-ranks arg-min of the local pricing result (lowest index wins; `anypos`
// is OR-ed).  Split in two halves so the exchange of the next iteration can be
// in flight during the sweep.
__device__ __forceinline__ void cand_publish(const LpDev &d, SelCtx &x, int best, int anypos)
{
    x.xseq++;
    const unsigned long long w = ((unsigned long long)x.xseq << 32) |
                                 ((unsigned long long)(anypos ? 1u : 0u) << 31) |
                                 (unsigned long long)(unsigned)best;
    __syncthreads();
    publish(d, offsetof(XHdr, cand) + (size_t)(x.xseq & 1) * MAXR * 8 + (size_t)d.rank * 8, w);
}
__device__ __forceinline__ void cand_complete(const LpDev &d, SelCtx &x, int &q, int &anypos)
{
    const unsigned seq = x.xseq;
    const size_t off = offsetof(XHdr, cand) + (size_t)(seq & 1) * MAXR * 8;
    if (!wait_words(d, off, 0, d.G, [seq](unsigned long long w) { return (unsigned)(w >> 32) == seq; }))
        x.ok = false;
    const unsigned long long *w = (const unsigned long long *)(d.xb[d.rank] + off);
    q = INT_BIG;
    anypos = 0;
    for (int r = 0; r < d.G; r++) {
        const unsigned long long v = ld_acquire_sys(w + r);
        q = min(q, (int)(v & 0x7fffffffu));
        anypos |= (int)((v >> 31) & 1u);
    }
    __syncthreads(); // every thread has read the words before anyone publishes the next exchange
}

// Slow path: column j as of now, from its owner's tableau, to every rank's
// xslot ([m] carries c_j).  Collective; used at the start of a solve, after a
// disableNV retry and inside the pair search.
__device__ const double *fetch_col(const LpDev &d, SelCtx &x, int j)
{
    double *mine = (double *)(d.xb[d.rank] + xoff_xslot(d));
    x.xs++;
    const unsigned xs = x.xs;
    const int owner = d.G > 1 ? owner_of(d, j) : 0;
    if (d.G > 1) {
        __syncthreads();
        publish(d, offsetof(XHdr, arrive) + (size_t)d.rank * 8, xs);
    }
    if (owner == d.rank) {
        if (d.G > 1 &&
            !wait_words(d, offsetof(XHdr, arrive), 0, d.G, [xs](unsigned long long w) { return w >= xs; }))
            x.ok = false;
        const int jl = j - d.col0;
        for (int base = 0; base < d.m + 1; base += SEL_THREADS * PT) {
            double v[PT];
#pragma unroll
            for (int u = 0; u < PT; u++) {
                const int i = base + u * SEL_THREADS + threadIdx.x;
                v[u] = i < d.m ? d.tab[(size_t)i * d.Cl + jl] : (i == d.m ? d.tgtf[jl] : 0.0);
            }
#pragma unroll
            for (int u = 0; u < PT; u++) {
                const int i = base + u * SEL_THREADS + threadIdx.x;
                if (i > d.m) continue;
                for (int r = 0; r < d.G; r++) ((double *)(d.xb[r] + xoff_xslot(d)))[i] = v[u];
            }
        }
        if (d.G > 1) {
            __syncthreads();
            publish(d, offsetof(XHdr, xflag) + (size_t)d.rank * 8, xs);
        }
    }
    if (d.G > 1) {
        if (!wait_words(d, offsetof(XHdr, xflag), owner, 1, [xs](unsigned long long w) { return w >= xs; }))
            x.ok = false;
    } else {
        __syncthreads();
    }
    return mine;
}

__global__ void __launch_bounds__(SEL_THREADS, 1) k_select(LpDev d)
{
    __shared__ XpMinIdx shm[33];
    __shared__ int shi[33];
    LpState *st = d.st;
    const int tid = threadIdx.x;
    const int n = d.n, m = d.m, Cl = d.Cl;

    const int status0 = st->status;
    const unsigned cnt0 = st->cnt, max_iter = st->max_iter;
    const int fast0 = st->fast;
    const unsigned swp0 = st->swp;
    SelCtx x;
    x.xseq = st->xseq;
    x.xs = st->xs;
    x.ok = true;
    __syncthreads();
    if (tid == 0) st->sweep_pending = 0;
    if (status0 != XPI_RUNNING) return;
    if (cnt0 >= max_iter) { // while (cnt < m_max_iter), :1039
        if (tid == 0) st->status = XP_SIX_TIME_OUT;
        return;
    }
#define SEL_EXIT(code)                  \
    do {                                \
        if (tid == 0) {                 \
            st->status = (code);        \
            st->xseq = x.xseq;          \
            st->xs = x.xs;              \
            st->fast = 0;               \
        }                               \
        return;                         \
    } while (0)

    int q = -1, p = -1;
    const double *col = nullptr;
    bool first = true;
    for (;;) {
        // ---- pricing, :1054-1069 ----
        int best, anypos;
        if (!(first && fast0)) {
            int lb, la;
            price_local(d, -1, 0, lb, la);
            lb = xp_block_min_int(lb, shi);
            la = __syncthreads_or(la);
            if (d.G > 1) cand_publish(d, x, lb, la);
            best = lb;
            anypos = la;
        }
        if (d.G > 1) {
            cand_complete(d, x, best, anypos);
            if (!x.ok) SEL_EXIT(XP_ERR_PEER);
        } else if (first && fast0) {
            best = st->q_next < 0 ? INT_BIG : st->q_next;
            anypos = st->pad;
        }
        // basic columns scanned before the break have their reduced cost forced to 0 (:1059)
        {
            const int zlim = (best == INT_BIG ? n : best) - d.col0;
            const int zl = min(zlim, Cl);
            for (int base = 0; base < zl; base += SEL_THREADS * PT) {
                int nv[PT];
#pragma unroll
                for (int u = 0; u < PT; u++) {
                    const int jl = base + u * SEL_THREADS + tid;
                    nv[u] = jl < zl ? d.nvset[d.col0 + jl] : 1;
                }
#pragma unroll
                for (int u = 0; u < PT; u++) {
                    const int jl = base + u * SEL_THREADS + tid;
                    if (!nv[u]) d.tgtf[jl] = 0.0;
                }
            }
            __syncthreads();
        }
        if (best == INT_BIG) {
            if (!anypos) SEL_EXIT(XPI_OPT_PENDING); // optimal exit; feasibility is checked by k_feas_*
            // ---- findPivotNVandBVPair, :670-773 ----
            // Pass A: eligible c_j > 0; pass B additionally c_j == 0 (tolerant).
            // findPivotBV is pure, so the c_j > 0 columns that failed in pass A
            // are not retried in pass B (same outcome, less work).
            int found = 0;
            for (int pass = 0; pass < 2 && !found; pass++) {
                int last = -1;
                for (;;) {
                    int cand, dummy;
                    price_local(d, last, pass, cand, dummy);
                    cand = xp_block_min_int(cand, shi);
                    if (d.G > 1) {
                        cand_publish(d, x, cand, 0);
                        cand_complete(d, x, cand, dummy);
                        if (!x.ok) SEL_EXIT(XP_ERR_PEER);
                    }
                    if (cand == INT_BIG) break;
                    col = fetch_col(d, x, cand);
                    if (!x.ok) SEL_EXIT(XP_ERR_PEER);
                    int r = ratio_test(d, cand, col, shm);
                    if (r >= 0) {
                        q = cand;
                        p = r;
                        found = 1;
                        break;
                    }
                    last = cand;
                }
            }
            if (!found) SEL_EXIT(XP_SIX_UNBOUND);
            break;
        }
        q = best;
        if (first && fast0) {
            const int owner = d.G > 1 ? owner_of(d, q) : 0;
            if (d.G > 1 && owner != d.rank) {
                const unsigned w = swp0;
                if (!wait_words(d, offsetof(XHdr, done), owner, 1,
                                [w](unsigned long long v) { return v >= w; }))
                    SEL_EXIT(XP_ERR_PEER);
            }
            col = (const double *)(d.xb[d.rank] + xoff_slot(d, owner, swp0 & 1));
        } else {
            col = fetch_col(d, x, q);
            if (!x.ok) SEL_EXIT(XP_ERR_PEER);
        }
        first = false;
        p = ratio_test(d, q, col, shm);
        if (p >= 0) break;
        disable_nv(d, q); // :1146-1151, retry without counting an iteration
    }

    // ---- genPair (:1156) + pivot bookkeeping ----
    const int bv = d.eq2bv[p];
    const double pv = ld_cg(col + p);
    const double cq = ld_cg(col + m); // c_q travels with the column
    const double rhs_p = d.rhsbuf[p];
    __syncthreads(); // everyone has read eq2bv[p], rhsbuf[p] before they change
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
    const double prow_rhs = xp_scale(rhs_p, r, r_one, r_zero);
    // multipliers f_i = -a[i][q] for the sweep, and the replicated constant column
    for (int base = 0; base < m; base += SEL_THREADS * PT) {
        double a[PT], rh[PT];
#pragma unroll
        for (int u = 0; u < PT; u++) {
            const int i = base + u * SEL_THREADS + tid;
            a[u] = i < m ? ld_cg(col + i) : 0.0;
            rh[u] = i < m ? d.rhsbuf[i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < PT; u++) {
            const int i = base + u * SEL_THREADS + tid;
            if (i >= m) continue;
            const double f = -a[u];
            d.fcol[i] = f;
            d.rhsbuf[i] = i == p ? prow_rhs : xp_add(rh[u], xp_mul(f, prow_rhs));
        }
    }
    double *rowp = d.tab + (size_t)p * Cl;
    for (int base = 0; base < Cl; base += SEL_THREADS * PT) {
        double xr[PT], tg[PT];
#pragma unroll
        for (int u = 0; u < PT; u++) {
            const int jl = base + u * SEL_THREADS + tid;
            xr[u] = jl < Cl ? rowp[jl] : 0.0;
            tg[u] = jl < Cl ? d.tgtf[jl] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < PT; u++) {
            const int jl = base + u * SEL_THREADS + tid;
            if (jl >= Cl) continue;
            const double xv = xp_scale(xr[u], r, r_one, r_zero); // mulOfRow(eqnum, 1/pivot)
            rowp[jl] = xv;
            d.prow[jl] = xv;
            double t = xp_mul(xv, -1.0);                      // nvexp.mul(-1)
            if (d.col0 + jl >= n) t = -t;                     // constant column keeps its sign
            t = cq_zero ? 0.0 : (cq_one ? t : xp_mul(t, cq)); // nvexp.mul(tgtf[nv])
            d.tgtf[jl] = xp_add(t, tg[u]);                    // tgtf.addRowToRow
        }
    }
    if (tid == 0) { // :1504-1510, and the replica of tgtf[rhs]
        d.nvset[q] = 0;
        d.nvset[bv] = 1;
        d.eq2bv[p] = q;
        d.bv2eq[q] = p;
        d.bv2eq[bv] = -1;
        double t = -xp_mul(prow_rhs, -1.0);
        t = cq_zero ? 0.0 : (cq_one ? t : xp_mul(t, cq));
        st->tg_rhs = xp_add(t, st->tg_rhs);
    }
    __syncthreads();
    // ---- price the NEXT iteration (no side effects) and start its exchange ----
    int nxt = INT_BIG, anyn = 0, fast1 = 0;
    if (cnt0 + 1 < max_iter) {
        price_local(d, -1, 0, nxt, anyn);
        nxt = xp_block_min_int(nxt, shi);
        anyn = __syncthreads_or(anyn);
        fast1 = 1;
        if (nxt != INT_BIG && tid < d.G) // c_q of my candidate rides in slot[rank][par][m]
            ((double *)(d.xb[tid] + xoff_slot(d, d.rank, (swp0 + 1) & 1)))[m] = d.tgtf[nxt - d.col0];
        if (d.G > 1) cand_publish(d, x, nxt, anyn);
    }
    if (tid == 0) {
        st->p = p;
        st->q_next = nxt == INT_BIG ? -1 : nxt;
        st->pad = anyn;
        st->cnt = cnt0 + 1;
        st->swp = swp0 + 1;
        st->fast = fast1;
        st->xseq = x.xseq;
        st->xs = x.xs;
        st->sweep_pending = 1;
    }
#undef SEL_EXIT
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
    const int p = st->p, Cl = d.Cl, m = d.m;
    const int qn = st->q_next < 0 ? -1 : st->q_next - d.col0; // local index of my candidate
    const size_t slot = xoff_slot(d, d.rank, st->swp & 1);
    const int r0 = blockIdx.y * rows_per_cta;
    const int r1 = min(m, r0 + rows_per_cta);
    for (int i = r0 + threadIdx.x; i < r1; i += THREADS) s_f[i - r0] = d.fcol[i];
    __syncthreads();
    const int j0 = (blockIdx.x * THREADS + threadIdx.x) * VEC;

    if (j0 >= Cl) {
        // nothing to update in this thread
    } else if (VEC == 2) {
        const double2 pr = *reinterpret_cast<const double2 *>(d.prow + j0);
        const int exq = (qn == j0) ? 0 : (qn == j0 + 1 ? 1 : -1);
        double *base = d.tab + j0;
        int i = r0;
        for (; i + UNROLL <= r1; i += UNROLL) {
            double2 a[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++)
                a[u] = *reinterpret_cast<const double2 *>(base + (size_t)(i + u) * Cl);
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const double f = s_f[i + u - r0];
                double2 v;
                v.x = xp_add(a[u].x, xp_mul(f, pr.x));
                v.y = xp_add(a[u].y, xp_mul(f, pr.y));
                if (i + u == p) v = a[u]; // row p was rewritten by k_select
                *reinterpret_cast<double2 *>(base + (size_t)(i + u) * Cl) = v;
                if (exq >= 0) {
                    const double e = exq ? v.y : v.x;
                    for (int r = 0; r < d.G; r++) ((double *)(d.xb[r] + slot))[i + u] = e;
                }
            }
        }
        for (; i < r1; i++) {
            double2 a = *reinterpret_cast<const double2 *>(base + (size_t)i * Cl);
            const double f = s_f[i - r0];
            double2 v;
            v.x = xp_add(a.x, xp_mul(f, pr.x));
            v.y = xp_add(a.y, xp_mul(f, pr.y));
            if (i == p) v = a;
            *reinterpret_cast<double2 *>(base + (size_t)i * Cl) = v;
            if (exq >= 0) {
                const double e = exq ? v.y : v.x;
                for (int r = 0; r < d.G; r++) ((double *)(d.xb[r] + slot))[i] = e;
            }
        }
    } else {
        const double pr = d.prow[j0];
        const bool exq = qn == j0;
        double *base = d.tab + j0;
        for (int i = r0; i < r1; i++) {
            double a = base[(size_t)i * Cl];
            double v = xp_add(a, xp_mul(s_f[i - r0], pr));
            if (i == p) v = a;
            base[(size_t)i * Cl] = v;
            if (exq)
                for (int r = 0; r < d.G; r++) ((double *)(d.xb[r] + slot))[i] = v;
        }
    }
    if (d.G > 1) {
        // The last CTA to finish tells every rank that sweep #swp is complete on this
        // rank, i.e. that slot[rank][swp & 1] holds my candidate's column everywhere.
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (atomicAdd(d.sweep_ctr, 1u) == gridDim.x * gridDim.y - 1) {
                *d.sweep_ctr = 0;
                __threadfence_system();
                for (int r = 0; r < d.G; r++)
                    st_release_sys((unsigned long long *)(d.xb[r] + offsetof(XHdr, done)) + d.rank,
                                   (unsigned long long)st->swp);
            }
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
// so only basic columns are visited (same value, bit for bit).  Sharded: the
// running sums travel rank to rank in column order (feas[] of the next rank).
__global__ void k_feas_rows(LpDev d)
{
    LpState *st = d.st;
    if (st->status != XPI_OPT_PENDING) return;
    __shared__ int s_bad;
    const unsigned long long ep = (unsigned long long)st->fe + 1;
    if (threadIdx.x == 0) s_bad = 0;
    if (d.G > 1 && d.rank > 0) {
        if (threadIdx.x == 0) {
            const unsigned long long *w =
                (const unsigned long long *)(d.xb[d.rank] + offsetof(XHdr, feas_in));
            const unsigned long long t0 = clock64();
            while (ld_acquire_sys(w) < ep)
                if (clock64() - t0 > SPIN_LIMIT) {
                    s_bad = 1;
                    break;
                }
        }
    }
    __syncthreads();
    if (s_bad) {
        if (threadIdx.x == 0) st->infeasible = 2; // peer timeout
        return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d.m) {
        const double *row = d.tab + (size_t)i * d.Cl;
        double sum = 0.0;
        if (d.G > 1 && d.rank > 0) sum = ld_cg((const double *)(d.xb[d.rank] + xoff_feas(d)) + i);
        const int nl = min(d.Cl, d.n - d.col0);
        for (int jl = 0; jl < nl; jl++) {
            const int g = d.col0 + jl;
            if (d.nvset[g]) continue;
            sum = xp_add(sum, xp_mul(row[jl], d.sol[g]));
        }
        if (d.rank + 1 < d.G) ((double *)(d.xb[d.rank + 1] + xoff_feas(d)))[i] = sum;
        else if (!xp_feq(sum, d.rhsbuf[i])) atomicOr(&st->infeasible, 1);
    }
    if (d.G > 1 && d.rank + 1 < d.G) { // last CTA hands the chain to the next rank
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (atomicAdd(d.feas_ctr, 1u) == gridDim.x - 1) {
                *d.feas_ctr = 0;
                __threadfence_system();
                st_release_sys((unsigned long long *)(d.xb[d.rank + 1] + offsetof(XHdr, feas_in)), ep);
            }
        }
    }
}

__global__ void k_feas_done(LpDev d)
{
    LpState *st = d.st;
    if (st->status != XPI_OPT_PENDING) return;
    const unsigned long long ep = (unsigned long long)st->fe + 1;
    int inf = st->infeasible;
    if (d.G > 1) {
        // every rank checked the variable bounds (replicated); the row sums end on the last rank
        if (d.rank == d.G - 1) {
            for (int r = 0; r < d.G; r++) {
                __threadfence_system();
                st_release_sys((unsigned long long *)(d.xb[r] + offsetof(XHdr, feas_res)), (ep << 2) | (unsigned)inf);
            }
        }
        const unsigned long long *w = (const unsigned long long *)(d.xb[d.rank] + offsetof(XHdr, feas_res));
        const unsigned long long t0 = clock64();
        unsigned long long v;
        while (((v = ld_acquire_sys(w)) >> 2) < ep)
            if (clock64() - t0 > SPIN_LIMIT) {
                v = (ep << 2) | 2u;
                break;
            }
        inf |= (int)(v & 3u);
    }
    st->fe = (unsigned)ep;
    if (inf & 2) {
        st->status = XP_ERR_PEER;
    } else if (inf) {
        st->status = XP_SIX_OPTIMAL_IS_INFEASIBLE;
    } else {
        st->status = XP_SIX_SUCC;
        st->maxv = st->tg_rhs; // :1119
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
        for (int k = t; k < d.C; k += stride) d.sol[k] = 0.0; // sol.reinit, :1028
    }
    if (t == 0) {
        if (fresh) {
            st->cnt = 0;
            st->n_log = 0;
            st->infeasible = 0;
            st->fast = 0;
            st->q_next = -1;
            st->maxv = 0.0; // :1027
            st->status = XPI_RUNNING;
        } else if (st->status == XP_SIX_TIME_OUT && st->cnt < max_iter) {
            st->status = XPI_RUNNING; // resume after a bounded run
        }
        st->max_iter = max_iter;
        st->sweep_pending = 0;
    }
}

// SIX::slack (lpsol.h:1405-1433) + identity basis (:1821-1841): [A | I | b],
// local slice [col0, col0+Cl) plus the replicated constant column.
template <class Gen>
__device__ __forceinline__ void fill_slack_form(const LpDev &d, int nvars, Gen gen)
{
    const int m = d.m, n = nvars, C = d.C;
    const size_t total = (size_t)m * d.Cl;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d.Cl), j = d.col0 + (int)(e % d.Cl);
        double v;
        if (j < n) v = gen(i, j);
        else if (j < n + m) v = (j - n == i) ? 1.0 : 0.0;
        else v = gen(i, n);
        d.tab[e] = v;
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < C; j += gridDim.x * blockDim.x) {
        if (j >= d.col0 && j < d.col0 + d.Cl) d.tgtf[j - d.col0] = j < n ? gen(m, j) : (j < n + m ? 0.0 : gen(m, n));
        if (j < n + m) {
            d.nvset[j] = j < n;
            d.bv2eq[j] = j < n ? -1 : j - n;
        }
        if (j < m) {
            d.eq2bv[j] = n + j;
            d.rhsbuf[j] = gen(j, n);
        }
        if (j == 0) d.st->tg_rhs = gen(m, n);
    }
}

struct GenLeq { // entries of the caller's leq (rows 0..m-1) and objective (row m)
    const double *leq, *tg;
    int m, n;
    __device__ __forceinline__ double operator()(int i, int j) const
    {
        return i < m ? leq[(size_t)i * (n + 1) + j] : tg[j];
    }
};

__global__ void k_slack_form(LpDev d, const double *leq, const double *tg, int n)
{
    GenLeq g;
    g.leq = leq;
    g.tg = tg;
    g.m = d.m;
    g.n = n;
    fill_slack_form(d, n, g);
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

struct GenSynth { // SURVEY 8(d) dense family: A_ij~U(0,1), b_i = 1+U*n, c_j~U(0,1), c_rhs = 0
    uint64_t seed;
    int m, n;
    __device__ __forceinline__ double operator()(int i, int j) const
    {
        if (i == m) return j < n ? u01(seed, (uint64_t)m * (n + 1) + j) : 0.0;
        const double u = u01(seed, (uint64_t)i * (n + 1) + j);
        return j < n ? u : 1.0 + u * n;
    }
};

__global__ void k_fill_synth(LpDev d, int n, uint64_t seed)
{
    GenSynth g;
    g.seed = seed;
    g.m = d.m;
    g.n = n;
    fill_slack_form(d, n, g);
}

// replicas after a raw upload (the caller's full arrays)
__global__ void k_set_tg_rhs(LpDev d, double v) { d.st->tg_rhs = v; }
__global__ void k_rhs_from_tab(LpDev d)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.m; i += gridDim.x * blockDim.x)
        d.rhsbuf[i] = d.tab[(size_t)i * d.Cl + (d.n - d.col0)];
}

__global__ void k_checksum(const double *a, int rows, int Cl, int col0, int C, unsigned long long *out)
{
    unsigned long long s = 0;
    const size_t nelem = (size_t)rows * Cl;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nelem;
         e += (size_t)gridDim.x * blockDim.x) {
        const size_t i = e / Cl, j = col0 + e % Cl;
        unsigned long long b = (unsigned long long)__double_as_longlong(a[e]);
        s += mix64(b ^ mix64((uint64_t)(i * C + j))); // keyed by the GLOBAL position: shard sums add up
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
    unsigned char *xblock = nullptr;       // this rank's exchange block
    void *peer_map[MAXR] = {nullptr};      // IPC mappings to close
    bool attached = false;
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
    const int m = d.m, Cl = d.Cl;
    if ((Cl & 1) == 0) {
        constexpr int TH = 256;
        int ctiles = (Cl / 2 + TH - 1) / TH;
        // aim for >= 8 CTAs per SM worth of row tiles, 8..64 rows per CTA
        int want = ctx->sm_count * 8;
        int rpc = (int)(((long long)m * ctiles + want - 1) / want);
        rpc = rpc < 8 ? 8 : (rpc > 64 ? 64 : rpc);
        rpc = (rpc + 7) & ~7;
        dim3 grid(ctiles, (m + rpc - 1) / rpc);
        k_sweep<2, TH, 8><<<grid, TH, rpc * sizeof(double), ctx->stream>>>(d, rpc);
    } else {
        constexpr int TH = 128;
        int ctiles = (Cl + TH - 1) / TH;
        int rpc = 16;
        dim3 grid(ctiles, (m + rpc - 1) / rpc);
        k_sweep<1, TH, 1><<<grid, TH, rpc * sizeof(double), ctx->stream>>>(d, rpc);
    }
    ctx->launches++;
    return 0;
}

static int lp_create(xp_ctx *ctx, int m, int C, int rank, int G, xp_lp_f64 **out)
{
    if (!ctx || !out || m < 1 || C < 2 || G < 1 || G > MAXR || rank < 0 || rank >= G)
        return XP_ERR_BAD_ARG;
    if (G > 1 && (C + 1) / 2 < G) return XP_ERR_BAD_ARG;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    xp_lp_f64 *lp = new xp_lp_f64();
    lp->ctx = ctx;
    LpDev &d = lp->d;
    memset(&d, 0, sizeof d);
    d.m = m;
    d.C = C;
    d.n = C - 1;
    d.W = (d.n + 31) / 32;
    d.rank = rank;
    d.G = G;
    d.col0 = shard_lo(C, G, rank);
    d.Cl = (rank + 1 < G ? shard_lo(C, G, rank + 1) : C) - d.col0;
    d.mpad = (m + 1 + 15) & ~15;
    d.log_cap = 1u << 16;
    const size_t n = d.n, Cl = d.Cl;
#define ALLOC(ptr, bytes) XP_CUDA_OK(ctx, cudaMalloc((void **)&(ptr), (bytes)))
    ALLOC(d.tab, (size_t)m * Cl * sizeof(double));
    ALLOC(d.tgtf, Cl * sizeof(double));
    ALLOC(d.prow, Cl * sizeof(double));
    ALLOC(d.fcol, m * sizeof(double));
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
    ALLOC(d.feas_ctr, 16);
    ALLOC(d.st, sizeof(LpState));
    ALLOC(lp->xblock, xblock_bytes(d));
#undef ALLOC
    XP_CUDA_OK(ctx, cudaMemset(d.st, 0, sizeof(LpState)));
    XP_CUDA_OK(ctx, cudaMemset(d.feas_ctr, 0, 16));
    d.sweep_ctr = d.feas_ctr + 1;
    XP_CUDA_OK(ctx, cudaMemset(lp->xblock, 0, xblock_bytes(d)));
    d.xb[rank] = lp->xblock;
    lp->attached = G == 1;
    XP_CUDA_OK(ctx, cudaMallocHost((void **)&lp->h_st, sizeof(LpState)));
    *out = lp;
    return 0;
}

extern "C" int xp_lp_f64_create(xp_ctx *ctx, int m, int C, xp_lp_f64 **out)
{
    return lp_create(ctx, m, C, 0, 1, out);
}

extern "C" int xp_lp_f64_create_sharded(xp_ctx *ctx, int m, int C, int rank, int nranks,
                                        xp_lp_f64 **out)
{
    return lp_create(ctx, m, C, rank, nranks, out);
}

extern "C" int xp_lp_f64_local_cols(const xp_lp_f64 *lp, int *col0, int *ncols)
{
    if (!lp) return XP_ERR_BAD_ARG;
    if (col0) *col0 = lp->d.col0;
    if (ncols) *ncols = lp->d.Cl;
    return 0;
}

extern "C" int xp_lp_f64_peer_handle(xp_lp_f64 *lp, void *handle)
{
    if (!lp || !handle) return XP_ERR_BAD_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) <= XP_PEER_HANDLE_BYTES, "handle size");
    xp_ctx *ctx = lp->ctx;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    XP_CUDA_OK(ctx, cudaIpcGetMemHandle(&h, lp->xblock));
    memset(handle, 0, XP_PEER_HANDLE_BYTES);
    memcpy(handle, &h, sizeof h);
    return 0;
}

extern "C" int xp_lp_f64_peer_attach(xp_lp_f64 *lp, const void *handles)
{
    if (!lp || !handles) return XP_ERR_BAD_ARG;
    xp_ctx *ctx = lp->ctx;
    LpDev &d = lp->d;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    for (int r = 0; r < d.G; r++) {
        if (r == d.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char *)handles + (size_t)r * XP_PEER_HANDLE_BYTES, sizeof h);
        void *p = nullptr;
        XP_CUDA_OK(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        lp->peer_map[r] = p;
        d.xb[r] = (unsigned char *)p;
    }
    lp->attached = true;
    return 0;
}

// Same-process variant (tests, single-process multi-device hosts): `all` holds
// the nranks handles in rank order; devices get peer access enabled if they differ.
extern "C" int xp_lp_f64_peer_attach_local(xp_lp_f64 *lp, xp_lp_f64 *const *all)
{
    if (!lp || !all) return XP_ERR_BAD_ARG;
    xp_ctx *ctx = lp->ctx;
    LpDev &d = lp->d;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    for (int r = 0; r < d.G; r++) {
        if (!all[r] || all[r]->d.G != d.G || all[r]->d.rank != r || all[r]->d.m != d.m ||
            all[r]->d.C != d.C)
            return XP_ERR_BAD_ARG;
        if (r == d.rank) continue;
        const int pd = all[r]->ctx->device;
        if (pd != ctx->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(pd, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else XP_CUDA_OK(ctx, e);
        }
        d.xb[r] = all[r]->xblock;
    }
    lp->attached = true;
    return 0;
}

extern "C" void xp_lp_f64_destroy(xp_lp_f64 *lp)
{
    if (!lp) return;
    LpDev &d = lp->d;
    cudaSetDevice(lp->ctx->device);
    cudaStreamSynchronize(lp->ctx->stream);
    for (int r = 0; r < MAXR; r++)
        if (lp->peer_map[r]) cudaIpcCloseMemHandle(lp->peer_map[r]);
    void *ptrs[] = {d.tab,     d.tgtf,    d.prow,  d.fcol,     d.rhsbuf,  d.sol, lp->vc_diag,
                    lp->vc_rhs, d.nvset,  d.bv2eq, d.eq2bv,    d.tabu,    d.row_cnt,
                    d.col_cnt, d.log,     d.st,    d.feas_ctr, lp->xblock};
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

// Full (global) host arrays in; every rank keeps its column slice and the replicas.
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
    const size_t pitch = (size_t)d.C * sizeof(double);
    XP_CUDA_OK(ctx, cudaMemcpy2DAsync(d.tab, (size_t)d.Cl * sizeof(double), tableau + d.col0, pitch,
                                      (size_t)d.Cl * sizeof(double), d.m, cudaMemcpyHostToDevice, s));
    if (d.col0 + d.Cl == d.C) { // the constant column is in my slice
        k_rhs_from_tab<<<(d.m + 255) / 256, 256, 0, s>>>(d);
        ctx->launches++;
    } else {
        XP_CUDA_OK(ctx, cudaMemcpy2DAsync(d.rhsbuf, sizeof(double), tableau + d.n, pitch,
                                          sizeof(double), d.m, cudaMemcpyHostToDevice, s));
    }
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d.tgtf, tgtf + d.col0, d.Cl * sizeof(double), cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d.nvset, nvset, d.n, cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d.bv2eq, bv2eq, d.n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d.eq2bv, eq2bv, d.m * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    k_set_tg_rhs<<<1, 1, 0, s>>>(d, tgtf[d.n]);
    ctx->launches++;
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
    k_slack_form<<<ctx->sm_count * 4, 256, 0, s>>>(d, d_leq, d_tg, n);
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
    k_fill_synth<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d, n, seed);
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
    if (!lp->attached) {
        ctx->err = "sharded LP: peers not attached (xp_lp_f64_peer_attach)";
        return XP_ERR_BAD_ARG;
    }
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
    // Every rank of a sharded LP sees the same status words, hence issues the
    // same launches.
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
    if (lp->h_st->status == XP_ERR_PEER) ctx->err = "sharded LP: timed out waiting for a peer GPU";
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

// Full-size host arrays out: a sharded rank writes only its own columns of
// `tableau` / `tgtf` (the caller merges ranks); the replicated state is complete.
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
    if (tableau)
        XP_CUDA_OK(ctx, cudaMemcpy2DAsync(tableau + d.col0, (size_t)d.C * sizeof(double), d.tab,
                                          (size_t)d.Cl * sizeof(double), (size_t)d.Cl * sizeof(double),
                                          d.m, cudaMemcpyDeviceToHost, s));
    D2H(tgtf ? tgtf + d.col0 : nullptr, d.tgtf, d.Cl * sizeof(double));
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
    k_checksum<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d.tab, d.m, d.Cl, d.col0, d.C, acc);
    k_checksum<<<8, 256, 0, ctx->stream>>>(d.tgtf, 1, d.Cl, d.col0, d.C, acc + 1);
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
