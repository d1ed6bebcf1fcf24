// HBM-resident FP64 simplex: SIX<FloatMat,Float>::solveSlackForm on the device,
// on one GPU or column-sharded over up to 8 GPUs (one process per GPU).
//
// Reference semantics (all in /root/reference/src/com/lpsol.h):
//   solveSlackForm :1007-1191, findPivotBV :552-663, findPivotNVandBVPair
//   :670-773, pivot :1455-1511, PivotPairTab :68-154, is_feasible :783-822.
//
// Blocked (delayed-update) formulation.  The reference touches every tableau
// entry once per pivot: a[i][j] = a[i][j] + (-a[i][q]) * row_p[j], product
// rounded, then sum rounded (lpsol.h:1485-1489).  Here the tableau in HBM is
// only brought up to date every k pivots ("flush"); in between, the k pivot rows
// P[s][.] and multiplier columns F[s][.] = -a[.][q_s] are kept aside and the few
// entries the next decision needs -- the entering column (m values) and the
// leaving row (C values) -- are evaluated on demand by replaying the pending
// updates on them in order:
//     v = A[i][j];  for s = 0..t-1:  v = (i == p_s) ? P[s][j] : add(v, mul(F[s][i], P[s][j]))
// which is, operation for operation and rounding for rounding, what the
// reference would have stored.  The flush applies the same recurrence to every
// entry in one pass.  Pivot sequence and every bit of the state are therefore
// unchanged, while the tableau streams through HBM once per k pivots instead of
// once per pivot.  k = 1 is the reference's own schedule.
//
// Kernels (no host round trip inside a block):
//   k_pcol   (grid over rows)     entering column now (strided read + replay),
//            its multipliers F[t], both passes of the ratio test as block
//            arg-mins; the last CTA reduces, keeps the tabu table and swaps the basis.
//   k_prow   (grid over columns)  leaving row now (read + replay), scaled into
//            P[t]; objective row and replicated constant column updated; pricing
//            of the next iteration; the last CTA reduces (and, sharded, runs
//            the all-ranks arg-min of the candidates).
//   k_flush  (grid over tiles)    the rank-t update of the whole tableau slice as a
//            128-bit row-major stream, P in registers, F in shared memory.
// Rare paths (ratio test fails -> disableNV and re-price, findPivotNVandBVPair,
// start of a solve) run on one CTA ("slow path") with the same device functions.
//
// Algorithmic HBM bytes per pivot: 2*(m+1)*C*8; bytes actually
// moved per pivot: that / k plus O((m + C) * k).
//
// Column sharding: rank g owns columns [lo_g, hi_g) of the tableau,
// of P and of the objective row; the constant column, F, the basis maps and the
// tabu table are replicated and kept identical by construction (every rank
// takes the same decisions from the same bits).  Per pivot the owner of the
// entering column writes its m+1 values straight into every peer's F buffer and
// each rank publishes one 8-byte pricing candidate (lowest index wins) -- stores
// into CUDA-IPC-mapped peer memory over NVLink issued by the kernels themselves.
//
// The pair-tabu table is a bit matrix (n x n bits) plus per-row / per-column
// population counters, which makes canBeNVCandidate / canBeBVCandidate O(1)
// and exactly equivalent to the reference's byte-matrix scans.
#include "xp_common.cuh"

#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>
#include <vector>

namespace cg = cooperative_groups;

namespace {

constexpr int MAXR = XP_MAX_RANKS;
constexpr int KMAX = XP_MAX_BLOCK; // pivots per flush, upper bound
constexpr int NH = 8;              // blocks whose factors (F, records, pivot-row marks) are kept: ring slot = blk % NH
constexpr int TH = 256;            // threads per CTA of the panel kernels / the slow path
constexpr int INT_BIG = 0x7fffffff;
constexpr unsigned long long SPIN_LIMIT = 6000000000ULL; // ~3 s of SM clocks

struct LpState {
    int status;
    unsigned cnt;
    unsigned max_iter;
    int t;     // pivots pending in the open block (rows of P / F in use)
    int kblk;  // block size
    int kadapt; // > 0: bounded run, block sizes chosen per block with this upper bound
    int blk;   // blocks flushed so far (parity selects the F buffer)
    int q;     // entering column of the next step (INT_BIG: none)
    int anypos;
    int slow;  // next k_pcol must price / search on one CTA
    int pivot_pending;
    int p, bv, s0p; // pivot row, leaving variable, last_piv[p] before this step
    int zero_upto;  // basic columns j < zero_upto still owe the reference's tgtf[j] = 0 (:1059)
    unsigned n_log;
    int infeasible;
    unsigned xseq; // candidate exchanges so far
    unsigned xs;   // slow-path column fetches so far
    unsigned cseq; // entering columns published so far
    unsigned fe;   // feasibility-chain epoch
    unsigned wseq; // windowed-panel launches that did work so far (never reset: flag words only grow)
    int wb_t0;      // first pivot of the open block whose deferred columns k_prow_bulk still owes
    int wb_pending; // k_wpanel made pivots [wb_t0, t): their P rows / objective entries beyond the window are due
    int n_touched;
    int touched[KMAX]; // rows with last_piv >= 0
    int hist_t[NH];    // pivots of the closed blocks still in the ring (k_block_snapshot)
    unsigned wcnt;     // pivots made by k_wpanel so far (cnt - wcnt: pivots that needed the full-width kernels)
    int wfail;         // the last k_wpanel run ended on a failing ratio test (k_pcol takes that column)
    int qmax;          // highest entering column k_wpanel has used since the host last looked (k_qmax_reset)
    int rest_pending;  // lookahead: the block in ring slot rest_slot is closed on the window tiles only,
    int rest_slot;     // the tiles beyond the window still owe it (k_flush_w, slot == SLOT_LAG)
    double r, cq, prow_rhs;
    double maxv;
    double tg_rhs; // replica of the objective row's constant term
};

// Exchange block of one rank (one cudaMalloc, exported over CUDA IPC).  Every
// word has a single writer; sequence numbers only grow.
struct XHdr {
    unsigned long long cand[2][MAXR]; // (seq << 32) | anypos << 31 | candidate, by rank
    unsigned long long colflag[MAXR]; // column event number whose data the rank has pushed
    unsigned long long arrive[MAXR];  // slow fetch: rank reached fetch #xs
    unsigned long long feas_in;       // feasibility chain: partial sums from rank-1 are in feas[]
    unsigned long long feas_res;      // (epoch << 2) | flags, broadcast by the last rank
    // k_panel, sharded: the pivot the owner of the entering column decided (ratio test on
    // its own copy of the column), by parity of the column event number
    unsigned long long piv[2][9]; // a, rh, cq (2 words each), p, bv, s0: payload | tag << 32
    // windowed panel (k_wpanel), sharded: rank 0 decides a whole run of pivots alone and hands
    // the peers its exit state + one record per pivot (xoff_rec) + the multiplier columns
    unsigned long long wflag;       // leader -> peer: number of the windowed launch whose results are in place
    unsigned long long wack[MAXR];  // peer -> leader: last windowed launch the peer has consumed
    int wexit[8];                   // leader -> peer: t, q, anypos, zero_upto, slow, status, ratio test failed, highest q
};
constexpr size_t XHDR_BYTES = 1024;
static_assert(sizeof(XHdr) <= XHDR_BYTES, "exchange header");

struct PartA {
    double v1, v2;
    int i1, i2;
    int pad[2];
};

struct LpDev {
    int m, C, n;  // global shape; n = rhs_idx = C-1
    int W;        // tabu words per row
    int rank, G;  // column shard
    int col0, Cl; // first local column, local width (row stride of tab and P)
    int mpad;     // doubles per F row (m+1 padded)
    int gridA, gridB;
    int w, wrpc, wwpc; // windowed panel: window = global columns [0, w) (0: off), rows / window columns per CTA
    double *tab, *tgtf, *P, *rhsbuf, *sol;
    const double *vc_diag, *vc_rhs; // may be null
    uint8_t *nvset;
    int32_t *bv2eq, *eq2bv, *last_piv;
    int32_t *hist_lp; // [NH][m] pivot-row marks of closed blocks (null unless the handle keeps a history)
    uint32_t *tabu;
    int32_t *row_cnt, *col_cnt;
    int32_t *log;
    unsigned log_cap;
    PartA *partA;
    int2 *partB;
    unsigned *ctr; // [0] pcol ticket, [1] prow ticket, [2] flush ticket, [3] feas ticket
    LpState *st;
    unsigned char *xb[MAXR]; // exchange blocks: xb[rank] is local, the rest peer mappings
};

// ---- exchange-block addressing ----
__host__ __device__ __forceinline__ size_t xoff_F(const LpDev &d, int par, int s)
{
    return XHDR_BYTES + ((size_t)(par * KMAX + s) * d.mpad) * sizeof(double); // par: ring slot of the block
}
__host__ __device__ __forceinline__ size_t xoff_feas(const LpDev &d)
{
    return XHDR_BYTES + ((size_t)(NH * KMAX) * d.mpad) * sizeof(double);
}
// k_panel, sharded: landing zone of the entering column, two tagged 8-byte words per row
__host__ __device__ __forceinline__ size_t xoff_land(const LpDev &d, int par)
{
    return XHDR_BYTES + ((size_t)(NH * KMAX + 1 + 2 * par) * d.mpad) * sizeof(double); // par: parity of the column event
}
// k_wpanel: one record per pivot of a block, by ring slot of the block
struct WRec {
    double r, cq, prow_rhs; // 1 / pivot element, c_q, scaled constant term of the pivot row
    int p, q, bv, s0p;      // pivot row, entering / leaving variable, last_piv[p] before this pivot
};
__host__ __device__ __forceinline__ size_t xoff_rec(const LpDev &d, int par)
{
    return XHDR_BYTES + ((size_t)(NH * KMAX + 5) * d.mpad) * sizeof(double) + (size_t)par * KMAX * sizeof(WRec);
}
__host__ __device__ __forceinline__ size_t xblock_bytes(const LpDev &d)
{
    return XHDR_BYTES + ((size_t)(NH * KMAX + 5) * d.mpad) * sizeof(double) + (size_t)NH * KMAX * sizeof(WRec);
}
__device__ __forceinline__ double *Fptr(const LpDev &d, int r, int par, int s)
{
    return (double *)(d.xb[r] + xoff_F(d, par, s));
}
// Column range of rank r, in units of two columns (128-bit accesses).  An even split, except that
// rank 0 -- the leader of windowed runs, whose slice must hold the pricing window -- keeps at least
// LEADER_MIN columns when the even share would be smaller (8 GPUs at c3: 3072 + 7 x 1902 instead of
// 8 x 2048: the entering column of the dense family climbs past 2048 after ~3 600 pivots, and every
// pivot outside the window costs two NVLink exchanges).  Mirrored by sharded.shard_bounds().
constexpr int LEADER_MIN = 3072;
__host__ __device__ __forceinline__ int shard_lo(int C, int G, int r)
{
    const long long pairs = (C + 1) / 2;
    if (G > 2 && pairs / G < LEADER_MIN / 2 && pairs >= LEADER_MIN) { // (needs room for the peers as well)
        if (r == 0) return 0;
        const long long rest = pairs - LEADER_MIN / 2;
        const long long lo = LEADER_MIN + 2 * (rest * (r - 1) / (G - 1));
        return lo > C ? C : (int)lo;
    }
    long long lo = 2 * (pairs * r / G);
    return lo > C ? C : (int)lo;
}
__device__ __forceinline__ int owner_of(const LpDev &d, int j)
{
    if (d.G == 1) return 0;
    int r = (int)(((long long)(j / 2) * d.G) / ((d.C + 1) / 2));
    if (r >= d.G) r = d.G - 1;
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
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Size of the next block of a bounded run (max_iter known): as few passes over the tableau as
// the upper bound allows, sizes in multiples of 8 (the flush kernel's unrolled step count) and
// none of them nearly empty, e.g. 200 pivots -> 32 32 32 32 24 24 24.  A pure function of
// replicated state, so every rank of a sharded LP picks the same size.
__device__ __forceinline__ void next_block(LpState *st)
{
    const int kmax = st->kadapt;
    if (kmax <= 0 || st->max_iter == XP_NO_ITER_LIMIT || st->cnt >= st->max_iter) return;
    const unsigned left = st->max_iter - st->cnt;
    const unsigned nb = (left + kmax - 1) / kmax;
    unsigned k = (left + nb - 1) / nb;
    if (k >= 16) k = (k + 7) & ~7u;
    if (k > (unsigned)kmax) k = kmax;
    if (k > left) k = left;
    st->kblk = k < 1 ? 1 : (int)k;
}

// Threads 0..G-1 each write `w` to the word at byte offset `off` of rank t's
// block.  Callers __syncthreads() first when the word guards data.
__device__ __forceinline__ void publish(const LpDev &d, size_t off, unsigned long long w)
{
    if ((int)threadIdx.x < d.G) {
        __threadfence_system();
        st_release_sys((unsigned long long *)(d.xb[threadIdx.x] + off), w);
    }
}
// Thread t < cnt waits until pred(word first+t at local offset off).  Returns
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

// ---------------------------------------------------------------------------
// Replay of the pending updates (see the header comment).
// ---------------------------------------------------------------------------
// Entry (i, ql) now, at step t; pq[s] = P[s][ql].  All loads are issued before the
// dependent add chain starts (one memory round trip, not t).
__device__ __forceinline__ double cur_in_col(const LpDev &d, int par, int t, int i, int ql,
                                             const double *pq)
{
    const int s0 = d.last_piv[i];
    const double a0 = d.tab[(size_t)i * d.Cl + ql];
    const double *f0 = Fptr(d, d.rank, par, 0) + i;
    double f[KMAX];
#pragma unroll
    for (int s = 0; s < KMAX; s++) f[s] = s < t ? ld_cg(f0 + (size_t)s * d.mpad) : 0.0;
    double v = a0;
#pragma unroll
    for (int s = 0; s < KMAX; s++) {
        if (s == s0) v = pq[s];
        else if (s > s0 && s < t) v = xp_add(v, xp_mul(f[s], pq[s]));
    }
    return v;
}
// Entry (p, jl) now, at step t; fp[s] = F[s][p], s0 = last_piv[p].
__device__ __forceinline__ double cur_in_row(const LpDev &d, int t, int p, int s0, int jl,
                                             const double *fp)
{
    const double a0 = d.tab[(size_t)p * d.Cl + jl];
    const double *p0 = d.P + jl;
    double pr[KMAX];
#pragma unroll
    for (int s = 0; s < KMAX; s++) pr[s] = s < t ? ld_cg(p0 + (size_t)s * d.Cl) : 0.0;
    double v = a0;
#pragma unroll
    for (int s = 0; s < KMAX; s++) {
        if (s == s0) v = pr[s];
        else if (s > s0 && s < t) v = xp_add(v, xp_mul(fp[s], pr[s]));
    }
    return v;
}

// Ratio-test keys of one row (lpsol.h:571-612 pass 1, :623-658 pass 2).
__device__ __forceinline__ void ratio_keys(const LpDev &d, int q, int i, double a, XpMinIdx &b1,
                                           XpMinIdx &b2)
{
    if (xp_feq(a, 0.0)) return; // neither pass takes a == 0 (tolerant)
    const int bv = d.eq2bv[i];
    if ((d.tabu[(size_t)q * d.W + (bv >> 5)] >> (bv & 31)) & 1u) return; // is_handle(q, bv)
    if (d.col_cnt[bv] >= d.n - 1) return;                                // !canBeBVCandidate
    XpMinIdx c;
    c.v = xp_div(d.rhsbuf[i], a);
    c.i = i;
    b2 = xp_better(b2, c);
    if (a > 0.0) b1 = xp_better(b1, c); // !(a <= 0) with the tolerant ==
}

// ---------------------------------------------------------------------------
// One-CTA ("slow path") building blocks.  Sharded: every rank runs them at the
// same logical point, so the exchanges inside are collective.
// ---------------------------------------------------------------------------
struct Seq {
    unsigned xseq, xs, cseq;
    bool ok; // false after a peer timeout
};

// All-ranks arg-min of a pricing result (lowest index wins; anypos is OR-ed).
__device__ void cand_exchange(const LpDev &d, Seq &x, int &best, int &anypos)
{
    if (d.G == 1) return;
    x.xseq++;
    const unsigned seq = x.xseq;
    const unsigned long long w = ((unsigned long long)seq << 32) |
                                 ((unsigned long long)(anypos ? 1u : 0u) << 31) |
                                 (unsigned long long)(unsigned)best;
    const size_t off = offsetof(XHdr, cand) + (size_t)(seq & 1) * MAXR * 8;
    __syncthreads();
    publish(d, off + (size_t)d.rank * 8, w);
    if (!wait_words(d, off, 0, d.G, [seq](unsigned long long v) { return (unsigned)(v >> 32) == seq; }))
        x.ok = false;
    const unsigned long long *wp = (const unsigned long long *)(d.xb[d.rank] + off);
    best = INT_BIG;
    anypos = 0;
    for (int r = 0; r < d.G; r++) {
        const unsigned long long v = ld_acquire_sys(wp + r);
        best = min(best, (int)(v & 0x7fffffffu));
        anypos |= (int)((v >> 31) & 1u);
    }
    __syncthreads(); // everyone has read the words before the next exchange may start
}

// Pricing over the local slice (lpsol.h:1054-1069 / the candidate scans of
// findPivotNVandBVPair): lowest eligible index > after.
// mode 0: c_j > 0; mode 1: c_j == 0 (tolerant) and not > 0.
__device__ void sp_price(const LpDev &d, int after, int mode, int *shi, int &best, int &anypos)
{
    const int nl = min(d.Cl, d.n - d.col0);
    best = INT_BIG;
    anypos = 0;
    for (int jl = threadIdx.x; jl < nl; jl += blockDim.x) {
        const int g = d.col0 + jl;
        if (!d.nvset[g]) continue;
        const double c = d.tgtf[jl];
        const bool pos = c > 0.0;
        if (pos) anypos = 1;
        const bool take = mode == 0 ? pos : (!pos && xp_feq(c, 0.0));
        if (take && g > after && best == INT_BIG && d.row_cnt[g] < d.n - 1) best = g;
    }
    best = xp_block_min_int(best, shi);
    anypos = __syncthreads_or(anypos);
}

// tgtf[j] = 0 for basic j < limit (lpsol.h:1059), local slice.
__device__ void sp_zero(const LpDev &d, int limit)
{
    const int zl = min(min(limit, d.n) - d.col0, d.Cl);
    for (int jl = threadIdx.x; jl < zl; jl += blockDim.x)
        if (!d.nvset[d.col0 + jl]) d.tgtf[jl] = 0.0;
    __syncthreads();
}

// Column q as of now into F[par][t] of every rank ([m] carries c_q).
__device__ void sp_column(const LpDev &d, Seq &x, int par, int t, int q, double *s_pq)
{
    const int owner = owner_of(d, q);
    x.cseq++;
    if (d.G > 1) { // nobody may still be reading F[par][t] from an earlier fetch of this step
        x.xs++;
        const unsigned xs = x.xs;
        __syncthreads();
        publish(d, offsetof(XHdr, arrive) + (size_t)d.rank * 8, xs);
        if (owner == d.rank &&
            !wait_words(d, offsetof(XHdr, arrive), 0, d.G, [xs](unsigned long long w) { return w >= xs; }))
            x.ok = false;
    }
    if (owner == d.rank) {
        const int ql = q - d.col0;
        __syncthreads();
        if ((int)threadIdx.x < t) s_pq[threadIdx.x] = d.P[(size_t)threadIdx.x * d.Cl + ql];
        __syncthreads();
        for (int i = threadIdx.x; i <= d.m; i += blockDim.x) {
            const double v = i < d.m ? -cur_in_col(d, par, t, i, ql, s_pq) : d.tgtf[ql];
            for (int r = 0; r < d.G; r++) Fptr(d, r, par, t)[i] = v;
        }
        __syncthreads();
        if (d.G > 1) publish(d, offsetof(XHdr, colflag) + (size_t)d.rank * 8, x.cseq);
    }
    if (d.G > 1) {
        const unsigned cs = x.cseq;
        if (!wait_words(d, offsetof(XHdr, colflag), owner, 1, [cs](unsigned long long w) { return w >= cs; }))
            x.ok = false;
    }
    __syncthreads();
}

// findPivotBV (lpsol.h:552-663) on F[par][t] (= -column).  Returns the pivot ROW or -1.
__device__ int sp_ratio(const LpDev &d, int par, int t, int q, XpMinIdx *shm)
{
    XpMinIdx b1, b2;
    b1.v = b2.v = 0.0;
    b1.i = b2.i = -1;
    const double *f = Fptr(d, d.rank, par, t);
    for (int i = threadIdx.x; i < d.m; i += blockDim.x) ratio_keys(d, q, i, -ld_cg(f + i), b1, b2);
    b1 = xp_block_argmin(b1, shm);
    if (b1.i >= 0) return b1.i;
    b2 = xp_block_argmin(b2, shm);
    return b2.i;
}

// PivotPairTab::disableNV (lpsol.h:114-121) with counter upkeep.
__device__ void sp_disable_nv(const LpDev &d, int q)
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

// genPair (:1156), the pivot log, and the basis swap of SIX::pivot (:1504-1510);
// leaves the scalars k_prow needs in the state.  One CTA.
__device__ void sp_commit_pivot(const LpDev &d, LpState *st, const Seq &x, int par, int t, int q, int p)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        const double *f = Fptr(d, d.rank, par, t);
        const int bv = d.eq2bv[p];
        const double pv = -ld_cg(f + p);
        const double cq = ld_cg(f + d.m);
        uint32_t *w = &d.tabu[(size_t)q * d.W + (bv >> 5)];
        const uint32_t bit = 1u << (bv & 31);
        if (!(*w & bit)) {
            *w |= bit;
            d.row_cnt[q] += 1;
            d.col_cnt[bv] += 1;
        }
        const unsigned k = st->n_log;
        if (k < d.log_cap) {
            d.log[3 * k] = q;
            d.log[3 * k + 1] = bv;
            d.log[3 * k + 2] = p;
        }
        st->n_log = k + 1;
        const double r = xp_div(1.0, pv); // mulOfRow(eqnum, 1 / pivot), :1471
        st->r = r;
        st->cq = cq;
        st->prow_rhs = xp_scale(d.rhsbuf[p], r, xp_feq(r, 1.0), xp_feq(r, 0.0));
        st->q = q;
        st->p = p;
        st->bv = bv;
        st->s0p = d.last_piv[p];
        d.nvset[q] = 0;
        d.nvset[bv] = 1;
        d.eq2bv[p] = q;
        d.bv2eq[q] = p;
        d.bv2eq[bv] = -1;
        st->pivot_pending = 1;
        st->xseq = x.xseq;
        st->xs = x.xs;
        st->cseq = x.cseq;
    }
}

__device__ void sp_exit(LpState *st, const Seq &x, int code)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        st->status = code;
        st->xseq = x.xseq;
        st->xs = x.xs;
        st->cseq = x.cseq;
    }
}

// The part of solveSlackForm's loop body (:1054-1156) that picks (q, p), on
// one CTA.  failed_q >= 0: the ratio test on that column just failed.
__device__ void sp_select(const LpDev &d, LpState *st, Seq &x, int par, int t, int failed_q,
                          XpMinIdx *shm, int *shi, double *s_pq)
{
    const int n = d.n;
    // zeroing owed by an earlier fast-path scan (same basis since): make it physical
    const int owed = st->zero_upto;
    __syncthreads();
    if (owed > 0) sp_zero(d, owed);
    if (threadIdx.x == 0) st->zero_upto = 0;
    if (failed_q >= 0) sp_disable_nv(d, failed_q); // :1146-1151, retry without counting an iteration
    int q = -1, p = -1;
    for (;;) {
        int best, anypos;
        sp_price(d, -1, 0, shi, best, anypos);
        cand_exchange(d, x, best, anypos);
        if (!x.ok) return sp_exit(st, x, XP_ERR_PEER);
        sp_zero(d, best == INT_BIG ? n : best);
        if (best == INT_BIG) {
            if (!anypos) return sp_exit(st, x, XPI_OPT_PENDING); // feasibility: k_feas_*
            // ---- findPivotNVandBVPair, :670-773 ----
            // Pass A: eligible c_j > 0; pass B additionally c_j == 0 (tolerant).
            // findPivotBV is pure, so the c_j > 0 columns that failed in pass A
            // are not retried in pass B (same outcome, less work).
            bool found = false;
            for (int pass = 0; pass < 2 && !found; pass++) {
                int last = -1;
                for (;;) {
                    int cand, dummy;
                    sp_price(d, last, pass, shi, cand, dummy);
                    cand_exchange(d, x, cand, dummy);
                    if (!x.ok) return sp_exit(st, x, XP_ERR_PEER);
                    if (cand == INT_BIG) break;
                    sp_column(d, x, par, t, cand, s_pq);
                    if (!x.ok) return sp_exit(st, x, XP_ERR_PEER);
                    const int r = sp_ratio(d, par, t, cand, shm);
                    if (r >= 0) {
                        q = cand;
                        p = r;
                        found = true;
                        break;
                    }
                    last = cand;
                }
            }
            if (!found) return sp_exit(st, x, XP_SIX_UNBOUND); // :1138-1141
            break;
        }
        sp_column(d, x, par, t, best, s_pq);
        if (!x.ok) return sp_exit(st, x, XP_ERR_PEER);
        p = sp_ratio(d, par, t, best, shm);
        if (p >= 0) {
            q = best;
            break;
        }
        sp_disable_nv(d, best);
    }
    sp_commit_pivot(d, st, x, par, t, q, p);
}

// ---------------------------------------------------------------------------
// k_pcol: entering column, multipliers, ratio test.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TH) k_pcol(LpDev d)
{
    __shared__ XpMinIdx shm[33];
    __shared__ int shi[33];
    __shared__ double s_pq[KMAX];
    __shared__ int s_flag;
    LpState *st = d.st;
    const int tid = threadIdx.x;
    if (st->status != XPI_RUNNING) return;
    if (st->cnt >= st->max_iter) { // while (cnt < m_max_iter), :1039
        if (blockIdx.x == 0 && tid == 0) st->status = XP_SIX_TIME_OUT;
        return;
    }
    const int t = st->t, par = st->blk & (NH - 1);
    if (t >= st->kblk) return; // block full: k_flush comes first
    Seq x;
    x.xseq = st->xseq;
    x.xs = st->xs;
    x.cseq = st->cseq;
    x.ok = true;
    if (st->slow) { // decided by the previous kernel; only CTA 0 works
        if (blockIdx.x == 0) sp_select(d, st, x, par, t, -1, shm, shi, s_pq);
        return;
    }
    const int q = st->q;
    const int owner = owner_of(d, q);
    const bool mine = owner == d.rank;
    const int ql = q - d.col0;
    const unsigned cs = x.cseq + 1;
    if (mine) {
        if (tid < t) s_pq[tid] = d.P[(size_t)tid * d.Cl + ql];
        __syncthreads();
    } else {
        if (!wait_words(d, offsetof(XHdr, colflag), owner, 1, [cs](unsigned long long w) { return w >= cs; })) {
            if (tid == 0) st->status = XP_ERR_PEER;
            return;
        }
    }
    XpMinIdx b1, b2;
    b1.v = b2.v = 0.0;
    b1.i = b2.i = -1;
    double *fmine = Fptr(d, d.rank, par, t);
    for (int i = blockIdx.x * TH + tid; i < d.m; i += gridDim.x * TH) {
        double a;
        if (mine) {
            a = cur_in_col(d, par, t, i, ql, s_pq);
            for (int r = 0; r < d.G; r++) Fptr(d, r, par, t)[i] = -a;
        } else {
            a = -ld_cg(fmine + i);
        }
        ratio_keys(d, q, i, a, b1, b2);
    }
    if (mine && blockIdx.x == 0 && tid == 0) {
        const double cq = d.tgtf[ql];
        for (int r = 0; r < d.G; r++) Fptr(d, r, par, t)[d.m] = cq;
    }
    b1 = xp_block_argmin(b1, shm);
    b2 = xp_block_argmin(b2, shm);
    if (tid == 0) {
        PartA pa;
        pa.v1 = b1.v;
        pa.i1 = b1.i;
        pa.v2 = b2.v;
        pa.i2 = b2.i;
        pa.pad[0] = pa.pad[1] = 0;
        d.partA[blockIdx.x] = pa;
        if (d.G > 1 && mine) __threadfence_system();
        else __threadfence();
        s_flag = atomicAdd(&d.ctr[0], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_flag) return;
    // ---- last CTA: everything above is complete on this rank ----
    __threadfence();
    if (tid == 0) d.ctr[0] = 0;
    x.cseq = cs;
    if (d.G > 1 && mine) publish(d, offsetof(XHdr, colflag) + (size_t)d.rank * 8, cs);
    b1.i = b2.i = -1;
    b1.v = b2.v = 0.0;
    for (int k = tid; k < (int)gridDim.x; k += TH) {
        const PartA *pa = &d.partA[k];
        XpMinIdx c;
        c.v = __ldcg(&pa->v1);
        c.i = __ldcg(&pa->i1);
        b1 = xp_better(b1, c);
        c.v = __ldcg(&pa->v2);
        c.i = __ldcg(&pa->i2);
        b2 = xp_better(b2, c);
    }
    b1 = xp_block_argmin(b1, shm);
    b2 = xp_block_argmin(b2, shm);
    const int p = b1.i >= 0 ? b1.i : b2.i;
    if (p < 0) sp_select(d, st, x, par, t, q, shm, shi, s_pq);
    else sp_commit_pivot(d, st, x, par, t, q, p);
}

// ---------------------------------------------------------------------------
// k_prow: leaving row -> P[t], objective row, constant column, next pricing.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TH) k_prow(LpDev d)
{
    __shared__ int shi[33];
    __shared__ double s_fp[KMAX];
    __shared__ int s_flag;
    LpState *st = d.st;
    const int tid = threadIdx.x;
    if (!st->pivot_pending || st->status != XPI_RUNNING) return;
    const int t = st->t, par = st->blk & (NH - 1), n = d.n, Cl = d.Cl;
    const int p = st->p, q = st->q, bv = st->bv, s0p = st->s0p, zero_upto = st->zero_upto;
    const double r = st->r, cq = st->cq, prow_rhs = st->prow_rhs;
    const bool r_one = xp_feq(r, 1.0), r_zero = xp_feq(r, 0.0);
    const bool cq_zero = xp_feq(cq, 0.0), cq_one = xp_feq(cq, 1.0);
    if (tid < t) s_fp[tid] = ld_cg(Fptr(d, d.rank, par, tid) + p);
    __syncthreads();
    int cand = INT_BIG, anypos = 0;
    double *Pt = d.P + (size_t)t * Cl;
    for (int jl = blockIdx.x * TH + tid; jl < Cl; jl += gridDim.x * TH) {
        const int g = d.col0 + jl;
        const double v = cur_in_row(d, t, p, s0p, jl, s_fp);
        const double xv = xp_scale(v, r, r_one, r_zero); // mulOfRow(eqnum, 1/pivot), :1471
        Pt[jl] = xv;
        double tg = d.tgtf[jl];
        const int nvg = g < n ? d.nvset[g] : 0;
        // the zeroing the pricing scan of this iteration owed (:1059), judged on the
        // basis as it was at that scan: bv was basic, q was not
        if (g < zero_upto && (g == bv || (!nvg && g != q))) tg = 0.0;
        double tt = xp_mul(xv, -1.0);                        // nvexp.mul(-1), :1496
        if (g >= n) tt = -tt;                                // constant column keeps its sign
        tt = cq_zero ? 0.0 : (cq_one ? tt : xp_mul(tt, cq)); // nvexp.mul(tgtf[nv])
        const double tn = xp_add(tt, tg);                    // tgtf.addRowToRow, :1501
        d.tgtf[jl] = tn;
        if (nvg && tn > 0.0) { // pricing of the next iteration, :1054-1069
            anypos = 1;
            if (cand == INT_BIG && d.row_cnt[g] < n - 1) cand = g;
        }
    }
    const double *ft = Fptr(d, d.rank, par, t);
    for (int i = blockIdx.x * TH + tid; i < d.m; i += gridDim.x * TH)
        d.rhsbuf[i] = i == p ? prow_rhs : xp_add(d.rhsbuf[i], xp_mul(ld_cg(ft + i), prow_rhs));
    cand = xp_block_min_int(cand, shi);
    anypos = __syncthreads_or(anypos);
    if (tid == 0) {
        d.partB[blockIdx.x] = make_int2(cand, anypos);
        __threadfence();
        s_flag = atomicAdd(&d.ctr[1], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_flag) return;
    // ---- last CTA ----
    __threadfence();
    if (tid == 0) d.ctr[1] = 0;
    cand = INT_BIG;
    anypos = 0;
    for (int k = tid; k < (int)gridDim.x; k += TH) {
        const int2 pb = __ldcg(&d.partB[k]);
        cand = min(cand, pb.x);
        anypos |= pb.y;
    }
    cand = xp_block_min_int(cand, shi);
    anypos = __syncthreads_or(anypos);
    Seq x;
    x.xseq = st->xseq;
    x.xs = st->xs;
    x.cseq = st->cseq;
    x.ok = true;
    cand_exchange(d, x, cand, anypos);
    const bool optimal = cand == INT_BIG && !anypos;
    if (optimal) { // the scan found nothing: every basic reduced cost is forced to 0 (:1059)
        __syncthreads();
        sp_zero(d, n);
    }
    if (tid == 0) {
        double tt = -xp_mul(prow_rhs, -1.0); // replica of the constant term of the objective row
        tt = cq_zero ? 0.0 : (cq_one ? tt : xp_mul(tt, cq));
        st->tg_rhs = xp_add(tt, st->tg_rhs);
        d.last_piv[p] = t;
        if (s0p < 0) st->touched[st->n_touched++] = p;
        st->t = t + 1;
        st->cnt += 1;
        if (q > st->qmax) st->qmax = q;
        st->pivot_pending = 0;
        st->q = cand;
        st->anypos = anypos;
        st->xseq = x.xseq;
        if (!x.ok) {
            st->status = XP_ERR_PEER;
        } else if (optimal) {
            st->zero_upto = 0;
            st->status = XPI_OPT_PENDING; // feasibility is checked by k_feas_*
        } else {
            st->zero_upto = cand == INT_BIG ? n : cand;
            st->slow = cand == INT_BIG; // some c_j > 0 but none eligible: pair search
        }
    }
}

// ---------------------------------------------------------------------------
// k_panel: the fast path of k_pcol + k_prow for up to (kblk - t) consecutive
// pivots in ONE persistent cooperative kernel (single GPU).  CTA c owns a fixed
// range of rows (entering column, ratio test, constant column) and a fixed
// range of columns (leaving row, objective row, pricing); two grid barriers per
// pivot replace two kernel boundaries, every CTA reduces the per-CTA partial
// results redundantly (deterministic, so all agree), and CTA 0 keeps the tabu
// table / basis maps while the others already work on the next column.  Any
// rare event (ratio test fails, no eligible candidate, optimum) ends the kernel
// with the state in global memory; the k_pcol / k_prow pair that follows in the
// stream handles that one pivot on the slow path, and the next k_panel resumes.
// ---------------------------------------------------------------------------
struct PanA { // per-CTA ratio-test result, both passes, with everything the pivot needs
    double v1, rh1, a1, v2, rh2, a2;
    int i1, bv1, s01, i2, bv2, s02;
};
struct PanB { // per-CTA pricing result
    double c;
    int cand, anypos;
};

// block arg-min of two (value, row) keys at once; result valid in every thread
__device__ __forceinline__ void argmin2_block(XpMinIdx &k1, XpMinIdx &k2, XpMinIdx *sh /* 2 x 9 */)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    k1 = xp_warp_argmin(k1);
    k2 = xp_warp_argmin(k2);
    __syncthreads();
    if (lane == 0) {
        sh[w] = k1;
        sh[9 + w] = k2;
    }
    __syncthreads();
    if (w == 0) {
        XpMinIdx y1, y2;
        y1.i = y2.i = -1;
        y1.v = y2.v = 0.0;
        if (lane < nw) {
            y1 = sh[lane];
            y2 = sh[9 + lane];
        }
        y1 = xp_warp_argmin(y1);
        y2 = xp_warp_argmin(y2);
        if (lane == 0) {
            sh[8] = y1;
            sh[17] = y2;
        }
    }
    __syncthreads();
    k1 = sh[8];
    k2 = sh[17];
}

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Returns false if the other CTAs did not arrive within SPIN_LIMIT (never on a
// healthy device: the launch is cooperative, all CTAs are resident).
__device__ __forceinline__ bool grid_barrier(unsigned long long *bar, unsigned long long target,
                                             bool sys_fence = false)
{
    __shared__ int s_ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sys_fence) __threadfence_system(); // this CTA wrote into peer GPUs
        else __threadfence();
        atomicAdd(bar, 1ULL);
        int ok = 1;
        const unsigned long long t0 = clock64();
        unsigned spins = 0;
        while (ld_acquire_gpu(bar) < target) {
            if ((++spins & 4095u) == 0 && clock64() - t0 > SPIN_LIMIT) {
                ok = 0;
                break;
            }
        }
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

// Tagged words between GPUs: 4 bytes of payload + the 4-byte number of the column event, so a
// word is valid the moment its tag matches (8-byte stores are single-copy atomic) and neither
// side needs a system-scope fence on the pivot path.
__device__ __forceinline__ void ll_put(unsigned long long *w, unsigned tag, unsigned payload)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(w),
                 "l"(((unsigned long long)tag << 32) | (unsigned long long)payload)
                 : "memory");
}
__device__ __forceinline__ void ll_put_f64(unsigned long long *w, unsigned tag, double x)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    ll_put(w, tag, (unsigned)b);
    ll_put(w + 1, tag, (unsigned)(b >> 32));
}
// Polls two words until both carry `tag`; false on timeout.
__device__ __forceinline__ bool ll_get_f64(const unsigned long long *w, unsigned tag, double &x)
{
    const unsigned long long t0 = clock64();
    unsigned spins = 0;
    for (;;) {
        unsigned long long a, b;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(w) : "memory");
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(b) : "l"(w + 1) : "memory");
        if ((unsigned)(a >> 32) == tag && (unsigned)(b >> 32) == tag) {
            x = __longlong_as_double((long long)(((b & 0xffffffffULL) << 32) | (a & 0xffffffffULL)));
            return true;
        }
        if ((++spins & 1023u) == 0 && clock64() - t0 > SPIN_LIMIT) return false;
    }
}
__device__ __forceinline__ bool ll_get_u32(const unsigned long long *w, unsigned tag, unsigned &x)
{
    const unsigned long long t0 = clock64();
    unsigned spins = 0;
    for (;;) {
        unsigned long long a;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(w) : "memory");
        if ((unsigned)(a >> 32) == tag) {
            x = (unsigned)a;
            return true;
        }
        if ((++spins & 1023u) == 0 && clock64() - t0 > SPIN_LIMIT) return false;
    }
}

constexpr int PANEL_NBAR = 2 * KMAX + 2; // barrier arrivals per CTA and launch (padded on exit)

__host__ __device__ inline size_t panel_smem_bytes(int rpc, int cpc)
{
    return (size_t)KMAX * ((size_t)rpc + cpc) * sizeof(double) + (size_t)2 * rpc * sizeof(int);
}

// CTA c owns rows [c*rpc, ..) and local columns [c*cpc, ..).  For its rows it keeps
// eq2bv, last_piv and the multipliers F[0..t) of the open block in shared memory, for
// its columns the pivot rows P[0..t): the replay of the pending updates runs out of
// shared memory, and the only global reads on the critical path of a pivot are the
// tableau column / row themselves and the per-CTA partial results.
__global__ void __launch_bounds__(TH, 1)
k_panel(LpDev d, PanA *partA, PanB *partB, unsigned long long *bar, unsigned long long bar_base,
        int rpc, int cpc, unsigned long long *dbg)
{
    extern __shared__ double s_dyn[]; // sF[KMAX][rpc] | sP[KMAX][cpc] | s_e2b[rpc] | s_lp[rpc]
    __shared__ XpMinIdx s_mi[18];
    __shared__ int shi[33];
    __shared__ double s_c[33];
    __shared__ double s_pq[KMAX], s_fp[KMAX];
    __shared__ double s_wd[2]; // winner: rhs, pivot element
    __shared__ int s_wi[2];    // winner: leaving variable, last_piv
    LpState *st = d.st;
    const int tid = threadIdx.x, c = blockIdx.x, NB = gridDim.x;
    const int n = d.n, m = d.m, Cl = d.Cl, G = d.G, col0 = d.col0;
    int nbar = 0;
    double *sF = s_dyn, *sP = s_dyn + (size_t)KMAX * rpc;
    int *s_e2b = (int *)(sP + (size_t)KMAX * cpc), *s_lp = s_e2b + rpc;
    const int r_lo = min(m, c * rpc), r_hi = min(m, r_lo + rpc);
    const int c_lo = min(Cl, c * cpc), c_hi = min(Cl, c_lo + cpc);

    int t = st->t;
    const int kblk = st->kblk, par = st->blk & (NH - 1);
    unsigned cnt = st->cnt;
    const unsigned max_iter = st->max_iter;
    int q = st->q, zero_upto = st->zero_upto, anypos = st->anypos;
    // (a launch that finds the block already full, or the iteration budget spent, must not pay
    // for bringing the open block's factors into shared memory)
    const bool go = st->status == XPI_RUNNING && !st->slow && !st->pivot_pending && q != INT_BIG &&
                    t < kblk && cnt < max_iter;
    double tg_rhs = st->tg_rhs;
    unsigned n_log = st->n_log;
    int n_touched = st->n_touched;
    unsigned xseq = st->xseq, cseq = st->cseq;
    int slow_out = 0;
    int q_prev = -1, bv_prev = -1; // the pivot whose basis swap CTA 0 may still be writing
    bool dirty = false;
    int qmx = -1; // highest entering column of this launch (the host sizes the pricing window by it)
    if (go) {
        for (int i = r_lo + tid; i < r_hi; i += TH) {
            s_e2b[i - r_lo] = d.eq2bv[i];
            s_lp[i - r_lo] = d.last_piv[i];
        }
        // resuming inside an open block: bring its factors into shared memory
        for (int e = tid; e < t * (r_hi - r_lo); e += TH) {
            const int s = e / (r_hi - r_lo), li = e - s * (r_hi - r_lo);
            sF[(size_t)s * rpc + li] = ld_cg(Fptr(d, d.rank, par, s) + r_lo + li);
        }
        for (int e = tid; e < t * (c_hi - c_lo); e += TH) {
            const int s = e / (c_hi - c_lo), lj = e - s * (c_hi - c_lo);
            sP[(size_t)s * cpc + lj] = ld_cg(d.P + (size_t)s * Cl + c_lo + lj);
        }
        __syncthreads();
    }
    // c_q: one GPU carries it from the pricing partials; sharded, it travels with the column
    double cq = (go && G == 1 && q < n) ? ld_cg(d.tgtf + q) : 0.0;
#define PANEL_T(k)                                                  \
    if (dbg && c == 0 && tid == 0) {                                \
        unsigned long long now__;                                   \
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now__));    \
        dbg[k] += now__ - tprev;                                    \
        tprev = now__;                                              \
    }
    unsigned long long tprev = 0;
    if (dbg && c == 0 && tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tprev));
    while (go && t < kblk && cnt < max_iter) {
        // ================= phase A: entering column, multipliers, ratio test =================
        // Sharded: only the owner of column q runs it (its CTAs write the multipliers into
        // every rank's F[t] on the way) and then hands the decision (p, pivot element, ...) to
        // the peers together with the "column landed" flag; peers go straight to phase B.
        const int owner = owner_of(d, q);
        const bool mine = owner == d.rank;
        const int ql = q - col0;
        cseq++;
        double *Ft = Fptr(d, d.rank, par, t);
        int p, bv, s0p;
        double piv_a, piv_rh;
        if (mine) {
            if (tid < t) s_pq[tid] = ld_cg(d.P + (size_t)tid * Cl + ql);
            __syncthreads();
            XpMinIdx b1, b2;
            b1.i = b2.i = -1;
            b1.v = b2.v = 0.0;
            double x1_rh = 0.0, x1_a = 0.0, x2_rh = 0.0, x2_a = 0.0; // extras of this thread's best rows
            int x1_bv = 0, x1_s0 = 0, x2_bv = 0, x2_s0 = 0;
            for (int i = r_lo + tid; i < r_hi; i += TH) {
                const int li = i - r_lo;
                const int rbv = s_e2b[li], s0 = s_lp[li];
                const double a0 = d.tab[(size_t)i * Cl + ql];
                const double rh = d.rhsbuf[i];
                const uint32_t tw = __ldcg(d.tabu + (size_t)q * d.W + (rbv >> 5));
                const int cc = __ldcg(d.col_cnt + rbv);
                double a = a0;
                if (s0 >= 0) a = s_pq[s0];
#pragma unroll 4
                for (int s = s0 + 1; s < t; s++) a = xp_add(a, xp_mul(sF[(size_t)s * rpc + li], s_pq[s]));
                __stcg(Ft + i, -a);
                for (int r = 0; r < G; r++)
                    if (r != d.rank)
                        ll_put_f64((unsigned long long *)(d.xb[r] + xoff_land(d, cseq & 1)) + 2 * i, cseq, -a);
                sF[(size_t)t * rpc + li] = -a;
                if (xp_feq(a, 0.0)) continue;           // neither pass takes a == 0 (tolerant)
                if ((tw >> (rbv & 31)) & 1u) continue;   // is_handle(q, bv), :589
                if (cc >= n - 1) continue;               // !canBeBVCandidate, :596
                XpMinIdx k;
                k.v = xp_div(rh, a);
                k.i = i;
                const XpMinIdx n2 = xp_better(b2, k); // pass 2, :623-658
                if (n2.i == i) x2_rh = rh, x2_a = a, x2_bv = rbv, x2_s0 = s0;
                b2 = n2;
                if (a > 0.0) { // pass 1, :571-612
                    const XpMinIdx n1 = xp_better(b1, k);
                    if (n1.i == i) x1_rh = rh, x1_a = a, x1_bv = rbv, x1_s0 = s0;
                    b1 = n1;
                }
            }
            PANEL_T(0) // phase A loads + replay
            {
                const int my1 = b1.i, my2 = b2.i;
                argmin2_block(b1, b2, s_mi);
                PanA *dst = partA + c;
                if (b1.i >= 0 && my1 == b1.i) { // the thread that owns the CTA's winner writes it out
                    __stcg(&dst->v1, b1.v), __stcg(&dst->rh1, x1_rh), __stcg(&dst->a1, x1_a);
                    __stcg(&dst->i1, b1.i), __stcg(&dst->bv1, x1_bv), __stcg(&dst->s01, x1_s0);
                }
                if (b2.i >= 0 && my2 == b2.i) {
                    __stcg(&dst->v2, b2.v), __stcg(&dst->rh2, x2_rh), __stcg(&dst->a2, x2_a);
                    __stcg(&dst->i2, b2.i), __stcg(&dst->bv2, x2_bv), __stcg(&dst->s02, x2_s0);
                }
                if (tid == 0) {
                    if (b1.i < 0) __stcg(&dst->i1, -1);
                    if (b2.i < 0) __stcg(&dst->i2, -1);
                }
            }
            PANEL_T(1) // block arg-min + partial store
            nbar++;
            if (!grid_barrier(bar, bar_base + (unsigned long long)nbar * NB)) {
                if (tid == 0) st->status = XP_ERR_PEER;
                return;
            }
            if (G > 1) cq = ld_cg(d.tgtf + ql);
            PANEL_T(2) // barrier 1
            double w_rh1 = 0.0, w_a1 = 0.0, w_rh2 = 0.0, w_a2 = 0.0;
            int w_bv1 = 0, w_s01 = 0, w_bv2 = 0, w_s02 = 0;
            b1.i = b2.i = -1;
            b1.v = b2.v = 0.0;
            if (tid < NB) {
                const PanA *src = partA + tid;
                b1.i = __ldcg(&src->i1), b2.i = __ldcg(&src->i2);
                b1.v = __ldcg(&src->v1), w_rh1 = __ldcg(&src->rh1), w_a1 = __ldcg(&src->a1);
                w_bv1 = __ldcg(&src->bv1), w_s01 = __ldcg(&src->s01);
                b2.v = __ldcg(&src->v2), w_rh2 = __ldcg(&src->rh2), w_a2 = __ldcg(&src->a2);
                w_bv2 = __ldcg(&src->bv2), w_s02 = __ldcg(&src->s02);
            }
            {
                const int my1 = b1.i, my2 = b2.i;
                argmin2_block(b1, b2, s_mi);
                if (b1.i >= 0) {
                    if (tid < NB && my1 == b1.i) s_wd[0] = w_rh1, s_wd[1] = w_a1, s_wi[0] = w_bv1, s_wi[1] = w_s01;
                } else if (b2.i >= 0) {
                    if (tid < NB && my2 == b2.i) s_wd[0] = w_rh2, s_wd[1] = w_a2, s_wi[0] = w_bv2, s_wi[1] = w_s02;
                }
                __syncthreads();
            }
            PANEL_T(3) // partial reduce
            p = b1.i >= 0 ? b1.i : b2.i;
            bv = s_wi[0], s0p = s_wi[1];
            piv_rh = s_wd[0], piv_a = s_wd[1];
            if (G > 1 && c == 0 && tid < G && tid != d.rank) { // the decision, as tagged words, to peer `tid`
                unsigned long long *dst = ((XHdr *)d.xb[tid])->piv[cseq & 1];
                ll_put_f64(dst + 0, cseq, piv_a), ll_put_f64(dst + 2, cseq, piv_rh), ll_put_f64(dst + 4, cseq, cq);
                ll_put(dst + 6, cseq, (unsigned)p), ll_put(dst + 7, cseq, (unsigned)bv);
                ll_put(dst + 8, cseq, (unsigned)s0p);
            }
        } else {
            // wait for the owner's decision (nine tagged words in my own exchange block)
            __shared__ double s_rd[3];
            __shared__ int s_ri[3];
            int fail = 0;
            if (tid < 6) {
                const unsigned long long *src = ((const XHdr *)d.xb[d.rank])->piv[cseq & 1];
                if (tid < 3) {
                    double x = 0.0;
                    fail = !ll_get_f64(src + 2 * tid, cseq, x);
                    s_rd[tid] = x;
                } else {
                    unsigned x = 0;
                    fail = !ll_get_u32(src + 3 + tid, cseq, x);
                    s_ri[tid - 3] = (int)x;
                }
            }
            if (__syncthreads_or(fail)) {
                if (tid == 0) st->status = XP_ERR_PEER;
                return;
            }
            piv_a = s_rd[0], piv_rh = s_rd[1], cq = s_rd[2];
            p = s_ri[0], bv = s_ri[1], s0p = s_ri[2];
            PANEL_T(3)
        }
        if (p < 0) break; // ratio test failed: k_pcol redoes this column and takes the slow path
        const double r = xp_div(1.0, piv_a); // mulOfRow(eqnum, 1 / pivot), :1471
        const bool r_one = xp_feq(r, 1.0), r_zero = xp_feq(r, 0.0);
        const bool cq_zero = xp_feq(cq, 0.0), cq_one = xp_feq(cq, 1.0);
        const double prow_rhs = xp_scale(piv_rh, r, r_one, r_zero);
        // ================= phase B: leaving row, objective row, constant column, pricing ====
        if (tid < t) s_fp[tid] = ld_cg(Fptr(d, d.rank, par, tid) + p);
        __syncthreads();
        int cand = INT_BIG, anyp = 0;
        double ccand = 0.0;
        double *Pt = d.P + (size_t)t * Cl;
        for (int jl = c_lo + tid; jl < c_hi; jl += TH) {
            const int lj = jl - c_lo;
            const int g = col0 + jl; // global column index
            const double a0 = d.tab[(size_t)p * Cl + jl];
            const double tg0 = d.tgtf[jl];
            int nvraw = g < n ? (int)__ldcg(d.nvset + g) : 0;
            // CTA 0 may still be writing the previous pivot's swap: that one is applied by hand
            nvraw = g == bv_prev ? 1 : (g == q_prev ? 0 : nvraw);
            const int rc = g < n ? __ldcg(d.row_cnt + g) : INT_BIG;
            double v = a0;
            if (s0p >= 0) v = sP[(size_t)s0p * cpc + lj];
#pragma unroll 4
            for (int s = s0p + 1; s < t; s++) v = xp_add(v, xp_mul(s_fp[s], sP[(size_t)s * cpc + lj]));
            const double xv = xp_scale(v, r, r_one, r_zero);
            sP[(size_t)t * cpc + lj] = xv;
            __stcg(Pt + jl, xv);
            double tg = tg0;
            if (g < zero_upto && g < n && !nvraw) tg = 0.0;      // zeroing owed by the scan (:1059)
            double tt = xp_mul(xv, -1.0);                        // nvexp.mul(-1), :1496
            if (g >= n) tt = -tt;                                // constant column keeps its sign
            tt = cq_zero ? 0.0 : (cq_one ? tt : xp_mul(tt, cq)); // nvexp.mul(tgtf[nv])
            const double tn = xp_add(tt, tg);                    // tgtf.addRowToRow, :1501
            d.tgtf[jl] = tn;
            const int nvnew = g == bv ? 1 : (g == q ? 0 : nvraw); // basis after the swap
            if (nvnew && tn > 0.0) { // pricing of the next iteration, :1054-1069
                anyp = 1;
                if (cand == INT_BIG && rc < n - 1) {
                    cand = g;
                    ccand = tn;
                }
            }
        }
        int lost = 0;
        for (int i = r_lo + tid; i < r_hi; i += TH) {
            double f = 0.0;
            if (mine) {
                f = sF[(size_t)t * rpc + (i - r_lo)];
            } else { // the owner wrote this step's multipliers into my landing zone as tagged words
                if (!ll_get_f64((const unsigned long long *)(d.xb[d.rank] + xoff_land(d, cseq & 1)) + 2 * i, cseq, f))
                    lost = 1;
                __stcg(Ft + i, f); // plain copy for the flush and for later replays
                sF[(size_t)t * rpc + (i - r_lo)] = f;
            }
            d.rhsbuf[i] = i == p ? prow_rhs : xp_add(d.rhsbuf[i], xp_mul(f, prow_rhs));
        }
        if (G > 1 && __syncthreads_or(lost)) {
            if (tid == 0) st->status = XP_ERR_PEER;
            return;
        }
        PANEL_T(4) // phase B loads + replay + stores
        if (p >= r_lo && p < r_hi && tid == 0) { // my row caches follow the swap
            s_e2b[p - r_lo] = q;
            s_lp[p - r_lo] = t;
        }
        {
            const int mycand = cand;
            cand = xp_block_min_int(cand, shi);
            anyp = __syncthreads_or(anyp);
            if (mycand == cand && cand != INT_BIG) s_c[32] = ccand; // unique owner of the minimum
            __syncthreads();
            if (tid == 0) {
                PanB *dst = partB + c;
                __stcg(&dst->c, cand != INT_BIG ? s_c[32] : 0.0);
                __stcg(&dst->cand, cand);
                __stcg(&dst->anypos, anyp);
            }
        }
        PANEL_T(5) // block min + partial store
        nbar++;
        if (!grid_barrier(bar, bar_base + (unsigned long long)nbar * NB)) {
            if (tid == 0) st->status = XP_ERR_PEER;
            return;
        }
        PANEL_T(6) // barrier 2
        {
            int cd = INT_BIG, ap = 0;
            double cv = 0.0;
            if (tid < NB) {
                const PanB *src = partB + tid;
                cd = __ldcg(&src->cand);
                ap = __ldcg(&src->anypos);
                cv = __ldcg(&src->c);
            }
            const int mycd = cd;
            cd = xp_block_min_int(cd, shi);
            ap = __syncthreads_or(ap);
            if (mycd == cd && cd != INT_BIG) s_c[32] = cv;
            __syncthreads();
            if (G > 1) { // all-ranks arg-min: one 8-byte word per rank, lowest index wins
                xseq++;
                const size_t off = offsetof(XHdr, cand) + (size_t)(xseq & 1) * MAXR * 8;
                if (c == 0 && tid < G) // self-validating word: no fence
                    ll_put((unsigned long long *)(d.xb[tid] + off) + d.rank, xseq,
                           (ap ? 0x80000000u : 0u) | (unsigned)cd);
                const unsigned seq = xseq;
                if (!wait_words(d, off, 0, G, [seq](unsigned long long v) { return (unsigned)(v >> 32) == seq; })) {
                    if (tid == 0) st->status = XP_ERR_PEER;
                    return;
                }
                const unsigned long long *wp = (const unsigned long long *)(d.xb[d.rank] + off);
                cd = INT_BIG;
                ap = 0;
                for (int r = 0; r < G; r++) {
                    const unsigned long long v = ld_acquire_sys(wp + r);
                    cd = min(cd, (int)(v & 0x7fffffffu));
                    ap |= (int)((v >> 31) & 1u);
                }
            }
            // ---- CTA 0 keeps the books of this pivot while the others move on ----
            if (c == 0 && tid == 0) {
                uint32_t *w = &d.tabu[(size_t)q * d.W + (bv >> 5)]; // genPair, :1156
                const uint32_t bit = 1u << (bv & 31);
                if (!(*w & bit)) {
                    *w |= bit;
                    d.row_cnt[q] += 1;
                    d.col_cnt[bv] += 1;
                }
                if (n_log < d.log_cap) {
                    d.log[3 * n_log] = q;
                    d.log[3 * n_log + 1] = bv;
                    d.log[3 * n_log + 2] = p;
                }
                n_log++;
                d.nvset[q] = 0; // :1504-1510
                d.nvset[bv] = 1;
                d.eq2bv[p] = q;
                d.bv2eq[q] = p;
                d.bv2eq[bv] = -1;
                d.last_piv[p] = t;
                if (s0p < 0) st->touched[n_touched++] = p;
                double tt = -xp_mul(prow_rhs, -1.0); // replica of the objective row's constant term
                tt = cq_zero ? 0.0 : (cq_one ? tt : xp_mul(tt, cq));
                tg_rhs = xp_add(tt, tg_rhs);
            }
            PANEL_T(7) // partial reduce + bookkeeping
            if (dbg && c == 0 && tid == 0) dbg[15] += 1;
            t++;
            cnt++;
            dirty = true;
            if (q > qmx) qmx = q;
            q_prev = q;
            bv_prev = bv;
            q = cd;
            anypos = ap;
            if (G == 1) cq = cd != INT_BIG ? s_c[32] : 0.0;
            zero_upto = cd == INT_BIG ? n : cd;
            __syncthreads();
            if (cd == INT_BIG) { // no eligible candidate: the slow path re-prices (optimum / pair search)
                slow_out = 1;
                break;
            }
        }
    }
    if (c == 0 && tid == 0 && dirty) {
        st->t = t;
        st->cnt = cnt;
        st->q = q;
        st->anypos = anypos;
        st->zero_upto = zero_upto;
        st->slow = slow_out;
        st->tg_rhs = tg_rhs;
        st->n_log = n_log;
        st->n_touched = n_touched;
        if (qmx > st->qmax) st->qmax = qmx;
    }
    if (c == 0 && tid == 0 && go) { // exchange counters advance even if no pivot completed
        st->xseq = xseq;
        st->cseq = cseq;
    }
    if (tid == 0 && nbar < PANEL_NBAR) { // keep the barrier counter in step with the host's base
        __threadfence();
        atomicAdd(bar, (unsigned long long)(PANEL_NBAR - nbar));
    }
}

// Which column tiles (of FT_TC = 256 columns) and which block a tableau pass works on.  The
// default is "every tile, the open block, close it".  A handle whose columns arrive in pieces
// (xp_six_two_stage_f64_large uploading behind the solve) runs the live pass on the tiles it
// has, and later replays closed blocks out of the ring (slot >= 0: F, records and pivot-row
// marks of that block, no closing) on the tiles that arrived late.
struct ColSet {
    int ct0a, ct1a, ct0b, ct1b; // tile ranges [ct0a, ct1a) u [ct0b, ct1b)
    int slot;                   // -1: the open block (live); >= 0: ring slot of a closed block; SLOT_LAG: see below
    int close;                  // the last CTA closes the block
    int chunk = 0;              // SLOT_LAG: row blocks per work chunk (0: default)
};
// Lookahead (windowed panel, on the rank that runs k_wpanel): a finished block is closed by k_block_close before any
// tile has taken it (rest_pending / rest_slot), and the pass over the tableau is launched with
// slot == SLOT_LAG *beside* the next block's k_wpanel -- programmatic dependent launch, the
// cluster holds 16 SMs, the pass the others.  It replays the owed block out of the ring exactly
// as the streamed upload's late tiles do, range a (the window tiles) first; k_wpanel spins until
// ctr[4] says they are all in place.  Apart from those window tiles k_wpanel touches the ring
// slot of the block it is deciding, the P rows of the window columns and the replicas (rhsbuf,
// objective row of the window); the pass touches the owed block's slot and the P rows of the
// columns beyond the window: no word in common.
constexpr int SLOT_LAG = -2;

#include "xp_large_wpanel.cuh"

// First pricing of a fresh solve restricted to the window (lpsol.h:1054-1069): the lowest
// eligible non-basic column with c_j > 0, handed to k_wpanel exactly as k_prow hands over its
// pricing result (the zeroing of basic columns passed by the scan is owed, zero_upto).  The
// ordinary start -- sp_select on the slow path -- needs every column; a handle whose columns are
// still arriving starts here.  Finds nothing inside the window: the state stays "slow".
__global__ void __launch_bounds__(1024) k_first_price_window(LpDev d)
{
    __shared__ int shi[33];
    LpState *st = d.st;
    if (st->status != XPI_RUNNING || !st->slow || st->pivot_pending || st->cnt != 0 || st->t != 0) return;
    const int lim = min(d.w, d.n);
    int best = INT_BIG;
    for (int j = threadIdx.x; j < lim; j += blockDim.x)
        if (best == INT_BIG && d.nvset[j] && d.tgtf[j] > 0.0 && d.row_cnt[j] < d.n - 1) best = j;
    best = xp_block_min_int(best, shi);
    if (threadIdx.x == 0 && best != INT_BIG) {
        st->q = best;
        st->anypos = 1;
        st->zero_upto = best;
        st->slow = 0;
    }
}

// Before a live pass on a handle that keeps a history: pivot count and pivot-row marks of the
// block about to close go into its ring slot.
__global__ void k_block_snapshot(LpDev d)
{
    LpState *st = d.st;
    const int t = st->t;
    if (t == 0 || (t < st->kblk && st->status == XPI_RUNNING)) return; // nothing to close
    const int slot = st->blk & (NH - 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.m; i += gridDim.x * blockDim.x)
        d.hist_lp[(size_t)slot * d.m + i] = d.last_piv[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) st->hist_t[slot] = t;
}

// while (cnt < m_max_iter), lpsol.h:1039 -- what the next k_pcol would find (nothing between here and
// there changes the state once the count is reached).  Said before the pass over the tableau, a
// bounded run needs no extra round of launches to report it, and the pass closes the last,
// partial block right away.
__global__ void k_timeout(LpDev d)
{
    LpState *st = d.st;
    if (st->status == XPI_RUNNING && st->cnt >= st->max_iter) st->status = XP_SIX_TIME_OUT;
}

__global__ void k_qmax_reset(LpDev d) { d.st->qmax = 0; }

// Lookahead: the same snapshot, and the block is closed right here -- before any tile has taken
// it.  The whole pass is then owed (rest_pending) and runs out of the ring slot beside the next
// block's k_wpanel, window tiles first (ctr[4] counts them for the cluster).  One CTA.
__global__ void __launch_bounds__(1024) k_block_close(LpDev d)
{
    LpState *st = d.st;
    const int t = st->t;
    // while (cnt < m_max_iter), lpsol.h:1039 -- what the next k_pcol would find (nothing between
    // here and there changes the state once the count is reached); said here, a bounded run needs no
    // extra round of launches to report it
    const bool timeout = st->status == XPI_RUNNING && st->cnt >= st->max_iter;
    const bool running = st->status == XPI_RUNNING && !timeout;
    __syncthreads();
    if (timeout && threadIdx.x == 0) st->status = XP_SIX_TIME_OUT;
    if (t == 0 || (t < st->kblk && running)) return; // nothing to close
    const int slot = st->blk & (NH - 1), nt = st->n_touched;
    for (int i = threadIdx.x; i < d.m; i += blockDim.x) d.hist_lp[(size_t)slot * d.m + i] = d.last_piv[i];
    __syncthreads();
    for (int k = threadIdx.x; k < nt; k += blockDim.x) d.last_piv[st->touched[k]] = -1;
    if (threadIdx.x == 0) {
        st->hist_t[slot] = t;
        st->rest_pending = 1;
        st->rest_slot = slot;
        d.ctr[4] = 0; // window units of the owed pass in place
        d.ctr[5] = 0; // chunks of the owed pass handed out
        st->n_touched = 0;
        st->t = 0;
        st->blk += 1;
        next_block(st);
    }
}

// ---------------------------------------------------------------------------
// k_flush: apply the t pending pivots to the tableau slice.
// Each thread owns VEC adjacent columns and keeps P[0..KB)[cols] in registers;
// the CTA walks `rows_per_cta` rows whose multipliers sit in shared memory.
// ---------------------------------------------------------------------------
template <int KB, int VEC, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) k_flush(LpDev d, int rows_per_cta)
{
    extern __shared__ double s_f[]; // [rows_per_cta][KB], then last_piv as ints
    __shared__ int s_flag;
    LpState *st = d.st;
    const int t = st->t;
    if (t == 0) return;
    if (t < st->kblk && st->status == XPI_RUNNING) return; // block still open
    const int par = st->blk & (NH - 1), Cl = d.Cl, m = d.m;
    int *s_lp = (int *)(s_f + (size_t)rows_per_cta * KB);
    const int r0 = blockIdx.y * rows_per_cta;
    const int r1 = min(m, r0 + rows_per_cta);
    for (int e = threadIdx.x; e < (r1 - r0) * KB; e += THREADS) {
        const int i = e / KB, s = e - i * KB;
        s_f[e] = s < t ? ld_cg(Fptr(d, d.rank, par, s) + r0 + i) : 0.0;
    }
    for (int i = r0 + threadIdx.x; i < r1; i += THREADS) s_lp[i - r0] = d.last_piv[i];
    __syncthreads();
    const int j0 = (blockIdx.x * THREADS + threadIdx.x) * VEC;
    if (j0 < Cl) {
        double pr[KB][VEC];
#pragma unroll
        for (int s = 0; s < KB; s++) {
            if (s < t) {
                if (VEC == 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(d.P + (size_t)s * Cl + j0);
                    pr[s][0] = v.x;
                    pr[s][VEC - 1] = v.y;
                } else {
                    pr[s][0] = d.P[(size_t)s * Cl + j0];
                }
            } else {
#pragma unroll
                for (int c = 0; c < VEC; c++) pr[s][c] = 0.0;
            }
        }
        double *base = d.tab + j0;
        for (int i = r0; i < r1; i += UNROLL) {
            double a[UNROLL][VEC];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                if (i + u < r1) {
                    if (VEC == 2) {
                        const double2 v = *reinterpret_cast<const double2 *>(base + (size_t)(i + u) * Cl);
                        a[u][0] = v.x;
                        a[u][VEC - 1] = v.y;
                    } else {
                        a[u][0] = base[(size_t)(i + u) * Cl];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                if (i + u >= r1) continue;
                const double *f = s_f + (size_t)(i + u - r0) * KB;
                const int s0 = s_lp[i + u - r0];
                if (s0 < 0) {
#pragma unroll
                    for (int s = 0; s < KB; s++) {
                        if (s < t) {
                            const double fs = f[s];
#pragma unroll
                            for (int c = 0; c < VEC; c++) a[u][c] = xp_add(a[u][c], xp_mul(fs, pr[s][c]));
                        }
                    }
                } else { // this row was a pivot row at step s0: restart from P[s0]
#pragma unroll
                    for (int c = 0; c < VEC; c++) {
                        double v = d.P[(size_t)s0 * Cl + j0 + c];
                        for (int s = s0 + 1; s < t; s++)
                            v = xp_add(v, xp_mul(f[s], d.P[(size_t)s * Cl + j0 + c]));
                        a[u][c] = v;
                    }
                }
                if (VEC == 2) {
                    double2 v;
                    v.x = a[u][0];
                    v.y = a[u][VEC - 1];
                    *reinterpret_cast<double2 *>(base + (size_t)(i + u) * Cl) = v;
                } else {
                    base[(size_t)(i + u) * Cl] = a[u][0];
                }
            }
        }
    }
    // the last CTA closes the block
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_flag = atomicAdd(&d.ctr[2], 1u) == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (!s_flag) return;
    if (threadIdx.x == 0) {
        d.ctr[2] = 0;
        for (int k = 0; k < st->n_touched; k++) d.last_piv[st->touched[k]] = -1;
        st->n_touched = 0;
        st->t = 0;
        st->blk += 1;
        next_block(st);
    }
}

// ---------------------------------------------------------------------------
// k_flush_t: the same update with the operands of the rank-t recurrence in
// shared memory (any t <= KMAX, no per-t register blocking): a CTA of THREADS
// threads owns 2*THREADS adjacent columns, keeps P[0..t)[its columns] in shared
// memory for its whole life, and walks its row group 64 rows at a time with the
// rows' multipliers staged beside it; each thread carries a TR x 2 register
// tile of tableau entries through the t steps (one 128-bit shared load of P and
// TR/2 broadcast loads of F per step for 4*TR non-fused FP64 operations).
// ---------------------------------------------------------------------------
constexpr int FT_ROWS = 64; // rows staged per __syncthreads pair

// THREADS = HALVES * LANES: LANES threads span a tile's 2*LANES columns, and the
// HALVES thread groups take alternate TR-row tiles of the staged rows, sharing
// the P tile (more warps per byte of shared memory).  Work units are
// (column tile, block of FT_ROWS rows); every CTA takes one contiguous,
// equally sized run of units, so the grid is exactly one resident wave, all
// SMs finish together, and a CTA reloads its P tile at most once.  The
// multipliers of unit u+1 and the tableau entries of the next tile are in
// flight while the current tile runs its t steps.
__device__ __forceinline__ void cp_async_cg16(void *smem, const void *g, int src_bytes)
{ // 16-byte asynchronous copy global -> shared (L2 only); bytes beyond src_bytes are zero-filled
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_ca4(void *smem, const void *g)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(g) : "memory");
}
// mbarrier in shared memory: every thread's cp.async copies arrive on it when they land
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count)
{
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(unsigned long long *b)
{ // arrives (without adding to the expected count) once all cp.async of this thread so far have completed
    asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(b);
    for (unsigned long long spin = 0;; spin++) {
        unsigned ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(a), "r"(parity)
                     : "memory");
        if (ok) return;
        if (spin > (1ull << 28)) __trap(); // cannot happen: every thread arrives once per phase
    }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int TR, int LANES, int HALVES>
__global__ void __launch_bounds__(LANES *HALVES) k_flush_t(LpDev d, int groups, int nbuf)
{
    extern __shared__ double sm[]; // sP[t][2*LANES] | sF[nbuf][t][FT_ROWS] | s_lp[nbuf][FT_ROWS], nbuf = 3 (or 2)
    __shared__ int s_flag;
    __shared__ unsigned long long s_mbar[3]; // "multipliers of unit lu have landed", lu % 3
    constexpr int THREADS = LANES * HALVES, TC = 2 * LANES;
    constexpr int TPB = FT_ROWS / (HALVES * TR); // tiles per unit and half
    LpState *st = d.st;
    const int t = st->t;
    if (t == 0) return;
    if (t < st->kblk && st->status == XPI_RUNNING) return; // block still open
    const int par = st->blk & (NH - 1), Cl = d.Cl, m = d.m, tid = threadIdx.x;
    const int lane = tid % LANES, half = tid / LANES;
    double *sP = sm, *sF = sm + (size_t)t * TC;
    int *s_lp = (int *)(sF + (size_t)nbuf * t * FT_ROWS);
    const double *sPl = sP + 2 * lane;
    const int nrb = (m + FT_ROWS - 1) / FT_ROWS;
    const int ctiles = (Cl + TC - 1) / TC;
    const long long units = (long long)ctiles * nrb;
    int u0, u1;
    if (groups > 0) { // HBM-bound regime: CTA = (column tile, row group); groups advance in step
        const int ct = blockIdx.x % ctiles, g = blockIdx.x / ctiles;
        const int bpg = (nrb + groups - 1) / groups;
        u0 = ct * nrb + min(nrb, g * bpg);
        u1 = ct * nrb + min(nrb, (g + 1) * bpg);
    } else { // FP64-bound regime: equal contiguous runs, one resident wave, no idle SM
        u0 = (int)(units * blockIdx.x / gridDim.x);
        u1 = (int)(units * (blockIdx.x + 1) / gridDim.x);
    }

    auto load_P = [&](int ct) {
        for (int e = tid; e < t * LANES; e += THREADS) {
            const int s = e / LANES, l = e - s * LANES;
            const int j = ct * TC + 2 * l;
            double2 v = make_double2(0.0, 0.0);
            if (j < Cl) v = *reinterpret_cast<const double2 *>(d.P + (size_t)s * Cl + j);
            *reinterpret_cast<double2 *>(sP + (size_t)s * TC + 2 * l) = v;
        }
    };
    // multipliers + pivot-row marks of unit u: global -> shared buffer b (of three) with cp.async,
    // no registers in between; every thread's copies arrive on the buffer's mbarrier.  Three
    // buffers and no block barrier: a warp waits for "unit landed", never for its siblings, and
    // nobody can be more than one unit ahead of the slowest warp (the copies of unit lu+1 are
    // issued at the start of unit lu, and unit lu+1 only opens once all 256 threads have done so),
    // so the buffer being refilled -- last read two units ago -- is free.
    auto copy_F = [&](int rb, int b) { // rb: first row of the unit
        double *dst = sF + (size_t)b * t * FT_ROWS;
        for (int e = tid; e < t * (FT_ROWS / 2); e += THREADS) {
            const int s = e / (FT_ROWS / 2), r = 2 * (e - s * (FT_ROWS / 2));
            const double *row = Fptr(d, d.rank, par, s);
            const int left = m - (rb + r); // rows of this pair that exist: >= 2, 1 or <= 0 (zero fill)
            cp_async_cg16(dst + (size_t)s * FT_ROWS + r, row + (left > 0 ? rb + r : 0), left >= 2 ? 16 : (left == 1 ? 8 : 0));
        }
        if (tid < FT_ROWS) // rows past m get some row's mark: never looked at (row < m is tested first)
            cp_async_ca4(s_lp + b * FT_ROWS + tid, d.last_piv + min(rb + tid, m - 1));
        cp_async_arrive(&s_mbar[b]);
    };
    // tile k (0..TPB) of unit u: first row and first column of this thread
    double2 nx[TR];
    auto load_tile = [&](int ct, int rb, int k) { // column tile, first row of the unit, tile in the unit
        const int r = rb + (k * HALVES + half) * TR;
        const int j = ct * TC + 2 * lane;
#pragma unroll
        for (int w = 0; w < TR; w++)
            nx[w] = (j < Cl && r + w < m) ? *reinterpret_cast<const double2 *>(d.tab + (size_t)(r + w) * Cl + j)
                                          : make_double2(0.0, 0.0);
    };
    if (tid == 0) {
        mbar_init(&s_mbar[0], THREADS);
        mbar_init(&s_mbar[1], THREADS);
        mbar_init(&s_mbar[2], THREADS);
    }
    __syncthreads();
    // Memory-bound passes (few steps per unit) keep the plain scheme -- wait for the copies,
    // block barrier per unit -- which measures 5 % faster there; compute-bound ones drop it.
    const bool mb = t >= 12 && nbuf == 3; // (two buffers: the host found no room for three at two CTAs per SM)
    int ct = u0 < u1 ? u0 / nrb : 0, rb = u0 < u1 ? (u0 % nrb) * FT_ROWS : 0; // advanced without divisions
    if (u0 < u1) {
        load_P(ct);
        copy_F(rb, 0);
        load_tile(ct, rb, 0);
    }
    if (!mb) cp_async_wait_all();
    __syncthreads(); // the P tile (plain stores)
    for (int u = u0; u < u1; u++) {
        int ct1 = ct, rb1 = rb + FT_ROWS; // the next unit
        if (rb1 >= nrb * FT_ROWS) {
            ct1 = ct + 1;
            rb1 = 0;
        }
        const int j0 = ct * TC + 2 * lane;
        const bool active = j0 < Cl; // Cl is even on this path
        const int lu = u - u0, buf = lu % nbuf;
        const double *sFu = sF + (size_t)buf * t * FT_ROWS;
        const int *lpu = s_lp + buf * FT_ROWS;
        const bool more = u + 1 < u1;
        if (mb) mbar_wait(&s_mbar[buf], (unsigned)(lu / 3) & 1u); // this unit's multipliers have landed
        if (more) copy_F(rb1, (lu + 1) % nbuf);
        for (int k = 0; k < TPB; k++) {
            double2 a[TR];
#pragma unroll
            for (int w = 0; w < TR; w++) a[w] = nx[w];
            if (k + 1 < TPB) load_tile(ct, rb, k + 1);
            else if (more) load_tile(ct1, rb1, 0);
            const int rc = (k * HALVES + half) * TR, row = rb + rc;
            if (!active || row >= m) continue;
            // marks are -1 (not a pivot row of this block) or the step: any >= 0 <=> AND >= 0
            bool special = false;
            if (TR == 8) {
                const int4 l0 = *reinterpret_cast<const int4 *>(lpu + rc), l1 = *reinterpret_cast<const int4 *>(lpu + rc + 4);
                special = (l0.x & l0.y & l0.z & l0.w & l1.x & l1.y & l1.z & l1.w) >= 0;
            } else {
#pragma unroll
                for (int w = 0; w < TR; w++) special |= (row + w < m) && lpu[rc + w] >= 0;
            }
            if (!special) {
#pragma unroll 8
                for (int s = 0; s < t; s++) {
                    const double2 p2 = *reinterpret_cast<const double2 *>(sPl + (size_t)s * TC);
                    const double *f = sFu + (size_t)s * FT_ROWS + rc;
#pragma unroll
                    for (int w = 0; w < TR; w += 2) {
                        const double2 f2 = *reinterpret_cast<const double2 *>(f + w);
                        a[w].x = xp_add(a[w].x, xp_mul(f2.x, p2.x));
                        a[w].y = xp_add(a[w].y, xp_mul(f2.x, p2.y));
                        a[w + 1].x = xp_add(a[w + 1].x, xp_mul(f2.y, p2.x));
                        a[w + 1].y = xp_add(a[w + 1].y, xp_mul(f2.y, p2.y));
                    }
                }
            } else { // a row of this tile was a pivot row at step s0: restart it from P[s0]
#pragma unroll
                for (int w = 0; w < TR; w++) {
                    if (row + w >= m) continue;
                    const int s0 = lpu[rc + w];
                    double2 v = a[w];
                    if (s0 >= 0) v = *reinterpret_cast<const double2 *>(sPl + (size_t)s0 * TC);
                    for (int s = s0 + 1; s < t; s++) {
                        const double2 p2 = *reinterpret_cast<const double2 *>(sPl + (size_t)s * TC);
                        const double fs = sFu[(size_t)s * FT_ROWS + rc + w];
                        v.x = xp_add(v.x, xp_mul(fs, p2.x));
                        v.y = xp_add(v.y, xp_mul(fs, p2.y));
                    }
                    a[w] = v;
                }
            }
#pragma unroll
            for (int w = 0; w < TR; w++)
                if (row + w < m) *reinterpret_cast<double2 *>(d.tab + (size_t)(row + w) * Cl + j0) = a[w];
        }
        if (more && ct1 != ct) { // next unit starts a new column tile: swap the P tile
            __syncthreads();
            load_P(ct1);
            __syncthreads();
        }
        if (!mb) {
            cp_async_wait_all();
            __syncthreads();
        }
        ct = ct1;
        rb = rb1;
    }
    // the last CTA closes the block
    if (tid == 0) {
        __threadfence();
        s_flag = atomicAdd(&d.ctr[2], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_flag) return;
    if (tid == 0) {
        d.ctr[2] = 0;
        for (int k = 0; k < st->n_touched; k++) d.last_piv[st->touched[k]] = -1;
        st->n_touched = 0;
        st->t = 0;
        st->blk += 1;
        next_block(st);
    }
}

// ---------------------------------------------------------------------------
// k_flush_w: k_flush_t with a wider register tile -- 8 rows x 4 columns per thread (two
// column pairs 2*LANES apart, so every 128-bit shared load of P stays conflict free) instead
// of 8 x 2.  Per step a thread then issues 6 shared loads (2 of P, 4 broadcast loads of F) for
// 64 non-fused FP64 operations instead of 5 for 32: the load/store unit, which k_flush_t keeps
// 66 % busy beside a 72 % busy FP64 pipe, gets 20 % fewer wavefronts per operation and the
// loop overhead is spread over twice the arithmetic.  The 64 accumulators leave no registers
// for the next tile at two CTAs per SM, so the next tile is prefetched into L2 instead.
// Same work units, multiplier staging and closing protocol as k_flush_t; same arithmetic per
// entry in the same order.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <int TR, int LANES, int GROUPS>
__global__ void __launch_bounds__(LANES *GROUPS, 2) k_flush_w(LpDev d, ColSet cs, int nbuf)
{
    extern __shared__ double sm[]; // sP[t][4*LANES] | sF[nbuf][t][FT_ROWS] | s_lp[nbuf][FT_ROWS]
    __shared__ int s_flag;
    __shared__ unsigned long long s_mbar[3];
    constexpr int THREADS = LANES * GROUPS, TC = 4 * LANES;
    constexpr int TPB = FT_ROWS / (GROUPS * TR); // tiles per unit and row group
    LpState *st = d.st;
    const bool live = cs.slot == -1, lag = cs.slot == SLOT_LAG;
    // The lagging pass is launched beside k_wpanel (programmatic stream serialization) and never
    // needs its results -- but the kernels behind it in the stream do, and they only wait for
    // THIS grid: one CTA of the lagging pass (the last to finish; CTA 0 if there is nothing to
    // do) therefore leaves through griddepcontrol.wait, so the grid cannot complete before the
    // cluster has (a no-op when launched the plain way).  The others exit at once: the cluster
    // waits for the window tiles of every CTA, so no CTA may wait for the cluster while it
    // keeps another one from starting.
    auto leave = [&]() {
        if (lag) asm volatile("griddepcontrol.wait;" ::: "memory");
    };
    if (lag && !st->rest_pending) {
        if (blockIdx.x == 0) leave();
        return;
    }
    const int slot = lag ? st->rest_slot : cs.slot;
    const int t = live ? st->t : st->hist_t[slot];
    if (t == 0 || (live && t < st->kblk && st->status == XPI_RUNNING)) { // nothing owed / block still open
        if (blockIdx.x == 0) leave();
        return;
    }
    const int par = live ? (st->blk & (NH - 1)) : slot, Cl = d.Cl, m = d.m, tid = threadIdx.x;
    const int32_t *marks = live ? d.last_piv : d.hist_lp + (size_t)slot * m;
    const int lane = tid % LANES, grp = tid / LANES;
    double *sP = sm, *sF = sm + (size_t)t * TC;
    int *s_lp = (int *)(sF + (size_t)nbuf * t * FT_ROWS);
    const double *sPl0 = sP + 2 * lane, *sPl1 = sP + 2 * LANES + 2 * lane;
    const int nrb = (m + FT_ROWS - 1) / FT_ROWS;
    const int nta = cs.ct1a - cs.ct0a, ntb = cs.ct1b - cs.ct0b;
    // Work units = (column tile, block of FT_ROWS rows), tile-major over the tiles of the set
    // (range a, then range b).  Ordinarily every CTA takes one contiguous share.  The lagging
    // pass hands them out as it goes instead -- ctr[5] counts chunks of cs.chunk row blocks of one
    // tile, single units over the last two tiles -- because its CTAs do not run alike: the
    // ones on the 16 SMs of the k_wpanel cluster only start when the cluster is through.  The
    // window tiles (range a) come first; their completion is counted in ctr[4] for the cluster.
    const int nt = nta + ntb, CH = cs.chunk > 0 ? cs.chunk : 4;
    const int cpt = (nrb + CH - 1) / CH, tail_t = nt < 2 ? nt : 2;
    const int nbig = (nt - tail_t) * cpt, nchunks = nbig + tail_t * nrb;
    __shared__ int s_chunk;
    auto load_P = [&](int ct) {
        for (int e = tid; e < t * (TC / 2); e += THREADS) {
            const int s = e / (TC / 2), l = e - s * (TC / 2);
            const int j = ct * TC + 2 * l;
            double2 v = make_double2(0.0, 0.0);
            if (j < Cl) v = *reinterpret_cast<const double2 *>(d.P + (size_t)s * Cl + j);
            *reinterpret_cast<double2 *>(sP + (size_t)s * TC + 2 * l) = v;
        }
    };
    auto copy_F = [&](int rb, int b) {
        double *dst = sF + (size_t)b * t * FT_ROWS;
        for (int e = tid; e < t * (FT_ROWS / 2); e += THREADS) {
            const int s = e / (FT_ROWS / 2), r = 2 * (e - s * (FT_ROWS / 2));
            const double *row = Fptr(d, d.rank, par, s);
            const int left = m - (rb + r);
            cp_async_cg16(dst + (size_t)s * FT_ROWS + r, row + (left > 0 ? rb + r : 0), left >= 2 ? 16 : (left == 1 ? 8 : 0));
        }
        if (tid < FT_ROWS) cp_async_ca4(s_lp + b * FT_ROWS + tid, marks + min(rb + tid, m - 1));
        cp_async_arrive(&s_mbar[b]);
    };
    auto prefetch_tile = [&](int ct, int rb, int k) { // this thread's lines of a tile to come
        const int r = rb + (k * GROUPS + grp) * TR;
        const int j = ct * TC + 2 * lane;
        if ((lane & 7) == 0) { // one request per 128-byte line
#pragma unroll
            for (int w = 0; w < TR; w++) {
                if (r + w >= m) break;
                if (j < Cl) prefetch_l2(d.tab + (size_t)(r + w) * Cl + j);
                if (j + 2 * LANES < Cl) prefetch_l2(d.tab + (size_t)(r + w) * Cl + j + 2 * LANES);
            }
        }
    };
    if (tid == 0) {
        mbar_init(&s_mbar[0], THREADS);
        mbar_init(&s_mbar[1], THREADS);
        mbar_init(&s_mbar[2], THREADS);
    }
    __syncthreads();
    const bool mb = t >= 12 && nbuf == 3;
    int lu = 0; // units this CTA has taken so far (selects the multiplier buffer and the mbarrier phase)
    int cur_tile = -1; // tile whose pivot rows sP holds
    auto tile_of = [&](int k) { return k < nta ? cs.ct0a + k : cs.ct0b + (k - nta); };
    for (int it = 0;; it++) {
    int u0, u1, sig = 0; // this share: units [u0, u1) in set order; units to report in ctr[4]
    if (!lag) {
        if (it > 0) break;
        const long long units = (long long)nt * nrb;
        u0 = (int)(units * blockIdx.x / gridDim.x), u1 = (int)(units * (blockIdx.x + 1) / gridDim.x);
    } else {
        if (tid == 0) s_chunk = (int)atomicAdd(&d.ctr[5], 1u);
        __syncthreads();
        const int c = s_chunk;
        if (c >= nchunks) break;
        int k, r0, r1;
        if (c < nbig) k = c / cpt, r0 = (c - k * cpt) * CH, r1 = min(r0 + CH, nrb);
        else k = (nt - tail_t) + (c - nbig) / nrb, r0 = (c - nbig) % nrb, r1 = r0 + 1;
        u0 = k * nrb + r0, u1 = k * nrb + r1;
        if (k < nta) sig = u1 - u0;
    }
    int cti = u0 < u1 ? u0 / nrb : 0, rb = u0 < u1 ? (u0 % nrb) * FT_ROWS : 0; // tile index in the set
    int ct = tile_of(cti);
    if (u0 < u1) {
        if (ct != cur_tile) load_P(ct), cur_tile = ct;
        copy_F(rb, lu % nbuf);
        prefetch_tile(ct, rb, 0);
    }
    if (!mb) cp_async_wait_all();
    __syncthreads();
    for (int u = u0; u < u1; u++) {
        int ct1 = ct, cti1 = cti, rb1 = rb + FT_ROWS;
        if (rb1 >= nrb * FT_ROWS) {
            cti1 = cti + 1;
            ct1 = tile_of(cti1);
            rb1 = 0;
        }
        const int j0 = ct * TC + 2 * lane, j1 = j0 + 2 * LANES;
        const bool act0 = j0 < Cl, act1 = j1 < Cl; // Cl is even on this path
        const int buf = lu % nbuf;
        const double *sFu = sF + (size_t)buf * t * FT_ROWS;
        const int *lpu = s_lp + buf * FT_ROWS;
        const bool more = u + 1 < u1;
        if (mb) mbar_wait(&s_mbar[buf], (unsigned)(lu / 3) & 1u);
        if (more) copy_F(rb1, (lu + 1) % nbuf);
#pragma unroll 1
        for (int k = 0; k < TPB; k++) {
            const int rc = (k * GROUPS + grp) * TR, row = rb + rc;
            double2 a[TR], b[TR];
            const bool full = act1 && row + TR <= m; // interior tile: no per-entry predicates or selects
            double *t0p = d.tab + (size_t)row * Cl + j0;
            if (full) {
#pragma unroll
                for (int w = 0; w < TR; w++) {
                    a[w] = *reinterpret_cast<const double2 *>(t0p + (size_t)w * Cl);
                    b[w] = *reinterpret_cast<const double2 *>(t0p + (size_t)w * Cl + 2 * LANES);
                }
            } else {
#pragma unroll
                for (int w = 0; w < TR; w++) {
                    a[w] = (act0 && row + w < m) ? *reinterpret_cast<const double2 *>(t0p + (size_t)w * Cl) : make_double2(0.0, 0.0);
                    b[w] = (act1 && row + w < m) ? *reinterpret_cast<const double2 *>(t0p + (size_t)w * Cl + 2 * LANES)
                                                 : make_double2(0.0, 0.0);
                }
            }
            if (k + 1 < TPB) prefetch_tile(ct, rb, k + 1);
            else if (more) prefetch_tile(ct1, rb1, 0);
            if (!act0 || row >= m) continue;
            bool special = false;
            if (TR == 8) {
                const int4 l0 = *reinterpret_cast<const int4 *>(lpu + rc), l1 = *reinterpret_cast<const int4 *>(lpu + rc + 4);
                special = (l0.x & l0.y & l0.z & l0.w & l1.x & l1.y & l1.z & l1.w) >= 0;
            } else {
#pragma unroll
                for (int w = 0; w < TR; w++) special |= (row + w < m) && lpu[rc + w] >= 0;
            }
            if (!special) {
#pragma unroll 4
                for (int s = 0; s < t; s++) {
                    const double2 p2 = *reinterpret_cast<const double2 *>(sPl0 + (size_t)s * TC);
                    const double2 q2 = *reinterpret_cast<const double2 *>(sPl1 + (size_t)s * TC);
                    const double *f = sFu + (size_t)s * FT_ROWS + rc;
#pragma unroll
                    for (int w = 0; w < TR; w += 2) {
                        const double2 f2 = *reinterpret_cast<const double2 *>(f + w);
                        a[w].x = xp_add(a[w].x, xp_mul(f2.x, p2.x));
                        a[w].y = xp_add(a[w].y, xp_mul(f2.x, p2.y));
                        b[w].x = xp_add(b[w].x, xp_mul(f2.x, q2.x));
                        b[w].y = xp_add(b[w].y, xp_mul(f2.x, q2.y));
                        a[w + 1].x = xp_add(a[w + 1].x, xp_mul(f2.y, p2.x));
                        a[w + 1].y = xp_add(a[w + 1].y, xp_mul(f2.y, p2.y));
                        b[w + 1].x = xp_add(b[w + 1].x, xp_mul(f2.y, q2.x));
                        b[w + 1].y = xp_add(b[w + 1].y, xp_mul(f2.y, q2.y));
                    }
                }
            } else { // a row of this tile was a pivot row at step s0: restart it from P[s0]
#pragma unroll
                for (int w = 0; w < TR; w++) {
                    if (row + w >= m) continue;
                    const int s0 = lpu[rc + w];
                    double2 v = a[w], x = b[w];
                    if (s0 >= 0) {
                        v = *reinterpret_cast<const double2 *>(sPl0 + (size_t)s0 * TC);
                        x = *reinterpret_cast<const double2 *>(sPl1 + (size_t)s0 * TC);
                    }
                    for (int s = s0 + 1; s < t; s++) {
                        const double2 p2 = *reinterpret_cast<const double2 *>(sPl0 + (size_t)s * TC);
                        const double2 q2 = *reinterpret_cast<const double2 *>(sPl1 + (size_t)s * TC);
                        const double fs = sFu[(size_t)s * FT_ROWS + rc + w];
                        v.x = xp_add(v.x, xp_mul(fs, p2.x));
                        v.y = xp_add(v.y, xp_mul(fs, p2.y));
                        x.x = xp_add(x.x, xp_mul(fs, q2.x));
                        x.y = xp_add(x.y, xp_mul(fs, q2.y));
                    }
                    a[w] = v;
                    b[w] = x;
                }
            }
            if (full) {
#pragma unroll
                for (int w = 0; w < TR; w++) {
                    *reinterpret_cast<double2 *>(t0p + (size_t)w * Cl) = a[w];
                    *reinterpret_cast<double2 *>(t0p + (size_t)w * Cl + 2 * LANES) = b[w];
                }
            } else {
#pragma unroll
                for (int w = 0; w < TR; w++)
                    if (row + w < m) {
                        *reinterpret_cast<double2 *>(t0p + (size_t)w * Cl) = a[w];
                        if (act1) *reinterpret_cast<double2 *>(t0p + (size_t)w * Cl + 2 * LANES) = b[w];
                    }
            }
        }
        if (more && ct1 != ct) {
            __syncthreads();
            load_P(ct1);
            cur_tile = ct1;
            __syncthreads();
        }
        if (!mb) {
            cp_async_wait_all();
            __syncthreads();
        }
        ct = ct1;
        cti = cti1;
        rb = rb1;
        lu++;
    }
    if (lag) { // everybody is through with this share (sP, s_chunk); window units are reported
        __syncthreads();
        if (tid == 0 && sig) {
            __threadfence();
            atomicAdd(&d.ctr[4], (unsigned)sig);
        }
    }
    } // shares
    if (!cs.close && !lag) return;
    if (tid == 0) {
        __threadfence();
        s_flag = atomicAdd(&d.ctr[2], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_flag) return;
    if (tid == 0) {
        d.ctr[2] = 0;
        if (lag) { // (runs beside k_wpanel: touches nothing but these two words)
            st->rest_pending = 0;
            return leave();
        }
        for (int k = 0; k < st->n_touched; k++) d.last_piv[st->touched[k]] = -1;
        st->n_touched = 0;
        st->t = 0;
        st->blk += 1;
        next_block(st);
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
        if (threadIdx.x == 0) atomicOr(&st->infeasible, 2); // peer timeout
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
            if (atomicAdd(&d.ctr[3], 1u) == gridDim.x - 1) {
                d.ctr[3] = 0;
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
                st_release_sys((unsigned long long *)(d.xb[r] + offsetof(XHdr, feas_res)),
                               (ep << 2) | (unsigned)inf);
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

__global__ void k_init(LpDev d, unsigned max_iter, int kblk, int fresh)
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
        for (int k = t; k < d.m; k += stride) d.last_piv[k] = -1;
    }
    if (t == 0) {
        if (fresh) {
            st->cnt = 0;
            st->n_log = 0;
            st->infeasible = 0;
            st->t = 0;
            st->n_touched = 0;
            st->q = INT_BIG;
            st->slow = 1; // first pricing of a solve runs on the slow path
            st->pivot_pending = 0;
            st->zero_upto = 0;
            st->maxv = 0.0; // :1027
            st->status = XPI_RUNNING;
            st->kblk = kblk > 0 ? kblk : 1;
            st->kadapt = 0;
            st->blk = 0; // ring slots restart (the previous solve is over on every rank)
            st->wb_pending = 0;
            for (int k = 0; k < NH; k++) st->hist_t[k] = 0;
            st->rest_pending = 0;
            st->wcnt = 0;
            st->wfail = 0;
            st->qmax = 0;
        } else {
            if (st->status == XP_SIX_TIME_OUT && st->cnt < max_iter) st->status = XPI_RUNNING; // resume
            if (kblk != 0 && st->t == 0) { // kblk < 0: adaptive with upper bound -kblk
                st->kblk = kblk > 0 ? kblk : -kblk;
                st->kadapt = kblk > 0 ? 0 : -kblk;
            }
        }
        st->max_iter = max_iter;
        if (!fresh && st->t == 0) next_block(st);
    }
}

// SIX::slack (lpsol.h:1405-1433) + identity basis (:1821-1841): [A | I | b],
// local slice [col0, col0+Cl) plus the replicated constant column.
// Rows [r0, r1) of the tableau; the launch that holds row 0 also writes the objective row and
// the basis maps (a row-chunked upload builds the slack form chunk by chunk as the rows arrive).
template <class Gen>
__device__ __forceinline__ void fill_slack_form(const LpDev &d, int nvars, Gen gen, int r0 = 0, int r1 = 0x7fffffff)
{
    const int m = d.m, n = nvars, C = d.C;
    if (r1 > m) r1 = m;
    const size_t total = (size_t)r1 * d.Cl;
    for (size_t e = (size_t)r0 * d.Cl + (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d.Cl), j = d.col0 + (int)(e % d.Cl);
        double v;
        if (j < n) v = gen(i, j);
        else if (j < n + m) v = (j - n == i) ? 1.0 : 0.0;
        else v = gen(i, n);
        d.tab[e] = v;
    }
    for (int j = r0 + blockIdx.x * blockDim.x + threadIdx.x; j < r1; j += gridDim.x * blockDim.x) {
        d.eq2bv[j] = n + j;
        d.rhsbuf[j] = gen(j, n);
    }
    if (r0 > 0) return;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < C; j += gridDim.x * blockDim.x) {
        if (j >= d.col0 && j < d.col0 + d.Cl)
            d.tgtf[j - d.col0] = j < n ? gen(m, j) : (j < n + m ? 0.0 : gen(m, n));
        if (j < n + m) {
            d.nvset[j] = j < n;
            d.bv2eq[j] = j < n ? -1 : j - n;
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

__global__ void k_slack_form(LpDev d, const double *leq, const double *tg, int n, int r0, int r1)
{
    GenLeq g;
    g.leq = leq;
    g.tg = tg;
    g.m = d.m;
    g.n = n;
    fill_slack_form(d, n, g, r0, r1);
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

// The same for global columns [g0, g1) only (single GPU), the constant column coming from a
// separate vector b: a column-chunked upload builds the slack form as the chunks arrive.
__global__ void k_slack_form_cols(LpDev d, const double *leq, const double *b, const double *tg, int n, int g0,
                                  int g1, int first)
{
    const int m = d.m, C = d.C, wd = g1 - g0;
    const size_t total = (size_t)m * wd;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / wd), j = g0 + (int)(e % wd);
        double v;
        if (j < n) v = leq[(size_t)i * (n + 1) + j];
        else if (j < n + m) v = (j - n == i) ? 1.0 : 0.0;
        else v = b[i];
        d.tab[(size_t)i * C + j] = v;
    }
    for (int j = g0 + blockIdx.x * blockDim.x + threadIdx.x; j < g1; j += gridDim.x * blockDim.x)
        d.tgtf[j] = j < n ? tg[j] : (j < n + m ? 0.0 : tg[n]);
    if (!first) return;
    // the basis maps belong to the solve, not to a piece of columns: all of them with the first
    // piece (a later piece must not undo the swaps made while it was on its way)
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n + m; j += gridDim.x * blockDim.x) {
        d.nvset[j] = j < n;
        d.bv2eq[j] = j < n ? -1 : j - n;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        d.eq2bv[i] = n + i;
        d.rhsbuf[i] = b[i];
        if (i == 0) d.st->tg_rhs = tg[n];
    }
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
    unsigned char *xblock = nullptr;  // this rank's exchange block
    void *peer_map[MAXR] = {nullptr}; // IPC mappings to close
    bool attached = false;
    int kblk = 0; // 0: automatic
    unsigned cnt_host = 0; // iteration count after the last solve (0 after an upload)
    // persistent panel kernel (single GPU)
    PanA *panA = nullptr;
    PanB *panB = nullptr;
    unsigned long long *bar = nullptr, bar_base = 0;
    int panel_nb = 0, panel_rpc = 0, panel_cpc = 0;
    bool use_panel = true;
    int ft_min = 2, ft_balanced_min = 10, ft_smem_set = 0, ft_occ = 1, ft_occ_k = -1, ft_nbuf = 3; // k_flush_t: smallest k that uses it, launch cache
    int ft_wide = 1, fw_smem_set = 0, fw_occ = 1, fw_occ_k = -1, fw_nbuf = 3; // k_flush_w (8 x 4 register tile) for the FP64-bound passes
    unsigned long long *panel_dbg = nullptr; // XP_PANEL_DBG=1: per-phase ns accumulators (16 words)
    unsigned long long *wpanel_dbg = nullptr; // the same for the windowed panel
    int window = 0;  // requested window: 0 automatic, < 0 off, > 0 forced width (xp_lp_f64_set_window)
    bool wpanel_ready = false; // kernel attributes set
    // optional per-launch timing of the flush kernel (CUDA events on the ctx stream)
    bool profile = false;
    int shared_sms = 0; // SMs the last solve's tableau passes left to k_wpanel (lookahead), else 0
    // lookahead: xp_lp_f64_solve returned SIX_TIME_OUT with the last closed block still owed to the
    // tableau -- the next solve applies it beside its first k_wpanel, anything else drains first (lp_drain)
    bool owed = false;
    ColSet owed_pass;
    int owed_kblk = 0;
    // automatic window: it follows the entering column (see lp_solve); w_max is the configured width
    int w_max = 0, wwpc_max = 0, q_prev = 0;
    bool w_auto = false;
    bool pess = true;      // the next batch of blocks carries the general-path kernels (see lp_solve)
    unsigned gen_seen = 0; // pivots made outside k_wpanel as of the last poll
    std::vector<cudaEvent_t> evs;
    uint64_t prof_sweeps = 0;
    double prof_sweep_ms = 0.0, prof_gap_ms = 0.0;
};

constexpr int PROF_MAX_SWEEPS = 4096;

extern "C" void xp_lp_f64_destroy(xp_lp_f64 *lp);

static int auto_block(const LpDev &d)
{
    // measured on B200 (c3, 8192 x 16384): pivots/s keeps growing up to k = 32, where the
    // flush (FP64-pipe bound, ~19 us per pivot) and the panel (~15 us per pivot) are comparable
    const long long cells = (long long)d.m * d.Cl;
    if (cells >= (4LL << 20)) return 32;
    if (cells >= (1LL << 18)) return 16;
    return 4;
}

template <int KB>
static void flush_launch_kb(xp_ctx *ctx, const LpDev &d)
{
    const int m = d.m, Cl = d.Cl;
    constexpr int THR = 256;
    const bool vec2 = (Cl & 1) == 0 && KB <= 16;
    int ctiles = vec2 ? (Cl / 2 + THR - 1) / THR : (Cl + THR - 1) / THR;
    int want = ctx->sm_count * 8; // >= 8 CTAs per SM worth of row tiles, 8..64 rows per CTA
    int rpc = (int)(((long long)m * ctiles + want - 1) / want);
    rpc = rpc < 8 ? 8 : (rpc > 64 ? 64 : rpc);
    rpc = (rpc + 7) & ~7;
    dim3 grid(ctiles, (m + rpc - 1) / rpc);
    size_t smem = (size_t)rpc * KB * sizeof(double) + (size_t)rpc * sizeof(int);
    if (vec2) k_flush<KB, (KB <= 16 ? 2 : 1), THR, (KB <= 4 ? 8 : 4)><<<grid, THR, smem, ctx->stream>>>(d, rpc);
    else k_flush<KB, 1, THR, 4><<<grid, THR, smem, ctx->stream>>>(d, rpc);
    ctx->launches++;
}

constexpr int FT_TR = 8, FT_LANES = 128, FT_HALVES = 2, FT_THREADS = FT_LANES * FT_HALVES;
constexpr int FW_LANES = 64, FW_GROUPS = 4; // k_flush_w: same 256-column tile, 8 x 4 entries per thread

static size_t flush_t_smem(int kblk, int nbuf = 3)
{
    return ((size_t)kblk * 2 * FT_LANES + (size_t)nbuf * kblk * FT_ROWS) * sizeof(double) + nbuf * FT_ROWS * sizeof(int);
}

static ColSet all_tiles(const LpDev &d)
{
    ColSet cs;
    cs.ct0a = 0;
    cs.ct1a = (d.Cl + 4 * FW_LANES - 1) / (4 * FW_LANES);
    cs.ct0b = cs.ct1b = 0;
    cs.slot = -1;
    cs.close = 1;
    return cs;
}

static int flush_w_launch(xp_lp_f64 *lp, int kblk, const ColSet *set = nullptr, int sm_reserve = 0, bool beside = false)
{
    xp_ctx *ctx = lp->ctx;
    const LpDev &d = lp->d;
    const ColSet cs = set ? *set : all_tiles(d);
    auto kern = k_flush_w<FT_TR, FW_LANES, FW_GROUPS>;
    if (lp->fw_smem_set == 0) {
        XP_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flush_t_smem(KMAX)));
        lp->fw_smem_set = 1;
    }
    if (lp->fw_occ_k != kblk) {
        int occ = 1;
        XP_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FT_THREADS, flush_t_smem(kblk)));
        lp->fw_nbuf = 3;
        if (occ < 2) {
            int occ2 = 1;
            XP_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kern, FT_THREADS, flush_t_smem(kblk, 2)));
            if (occ2 > occ) occ = occ2, lp->fw_nbuf = 2;
        }
        lp->fw_occ = occ < 1 ? 1 : occ;
        lp->fw_occ_k = kblk;
    }
    const size_t smem = flush_t_smem(kblk, lp->fw_nbuf);
    const int ctiles = (cs.ct1a - cs.ct0a) + (cs.ct1b - cs.ct0b);
    const long long units = (long long)ctiles * ((d.m + FT_ROWS - 1) / FT_ROWS);
    long long grid = (long long)lp->fw_occ * (ctx->sm_count - sm_reserve); // one resident wave on the SMs it gets
    if (grid > units) grid = units;
    if (grid < 1) return 0;
    if (beside) { // may start as soon as the kernel in front (k_wpanel) has, and never waits for it
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(FT_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        XP_CUDA_OK(ctx, cudaLaunchKernelEx(&cfg, kern, d, cs, lp->fw_nbuf));
    } else {
        kern<<<(unsigned)grid, FT_THREADS, smem, ctx->stream>>>(d, cs, lp->fw_nbuf);
    }
    ctx->launches++;
    return 0;
}

static int flush_t_launch(xp_lp_f64 *lp, int kblk)
{
    xp_ctx *ctx = lp->ctx;
    const LpDev &d = lp->d;
    if (lp->ft_wide && kblk >= lp->ft_balanced_min) return flush_w_launch(lp, kblk); // FP64-bound passes
    size_t smem = flush_t_smem(kblk);
    if (lp->ft_smem_set < (int)smem) {
        XP_CUDA_OK(ctx, cudaFuncSetAttribute(k_flush_t<FT_TR, FT_LANES, FT_HALVES>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flush_t_smem(KMAX)));
        lp->ft_smem_set = (int)flush_t_smem(KMAX);
    }
    if (lp->ft_occ_k != kblk) {
        int occ = 1;
        XP_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_flush_t<FT_TR, FT_LANES, FT_HALVES>,
                                                                      FT_THREADS, smem));
        lp->ft_nbuf = 3;
        const char *fb = getenv("XP_FLUSH_NBUF"); // "2" forces the two-buffer fallback (tests)
        if (occ < 2 || (fb && fb[0] == '2')) { // three multiplier buffers cost the second CTA per SM here: fall back to two
            int occ2 = 1;
            XP_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_flush_t<FT_TR, FT_LANES, FT_HALVES>,
                                                                          FT_THREADS, flush_t_smem(kblk, 2)));
            if (occ2 > occ || (fb && fb[0] == '2')) {
                occ = occ2;
                lp->ft_nbuf = 2;
            }
        }
        lp->ft_occ = occ < 1 ? 1 : occ;
        lp->ft_occ_k = kblk;
    }
    smem = flush_t_smem(kblk, lp->ft_nbuf);
    const int ctiles = (d.Cl + 2 * FT_LANES - 1) / (2 * FT_LANES);
    const long long units = (long long)ctiles * ((d.m + FT_ROWS - 1) / FT_ROWS);
    long long grid = (long long)lp->ft_occ * ctx->sm_count; // exactly one resident wave
    if (grid > units) grid = units;
    int groups = 0;
    if (kblk < lp->ft_balanced_min) { // memory-bound: row groups that move through the rows together
        groups = (int)(grid / ctiles);
        const int nrb = (d.m + FT_ROWS - 1) / FT_ROWS;
        if (groups < 1) groups = 1;
        if (groups > nrb) groups = nrb;
        grid = (long long)groups * ctiles;
    }
    k_flush_t<FT_TR, FT_LANES, FT_HALVES><<<(unsigned)grid, FT_THREADS, smem, ctx->stream>>>(d, groups, lp->ft_nbuf);
    ctx->launches++;
    return 0;
}

static int flush_launch(xp_lp_f64 *lp, int kblk)
{
    xp_ctx *ctx = lp->ctx;
    const LpDev &d = lp->d;
    if ((d.Cl & 1) == 0 && kblk >= lp->ft_min) return flush_t_launch(lp, kblk);
    if (kblk <= 1) flush_launch_kb<1>(ctx, d);
    else if (kblk <= 2) flush_launch_kb<2>(ctx, d);
    else if (kblk <= 4) flush_launch_kb<4>(ctx, d);
    else if (kblk <= 8) flush_launch_kb<8>(ctx, d);
    else if (kblk <= 12) flush_launch_kb<12>(ctx, d);
    else if (kblk <= 16) flush_launch_kb<16>(ctx, d);
    else if (kblk <= 24) flush_launch_kb<24>(ctx, d);
    else flush_launch_kb<32>(ctx, d);
    return 0;
}

template <int KB>
static cudaError_t preload_flush()
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, k_flush<KB, 1, 256, 4>);
    if (e != cudaSuccess) return e;
    return cudaFuncGetAttributes(&fa, k_flush<KB, (KB <= 16 ? 2 : 1), 256, (KB <= 4 ? 8 : 4)>);
}

static int lp_create_fill(xp_ctx *ctx, int m, int C, int rank, int G, xp_lp_f64 *lp);
static int window_config(xp_lp_f64 *lp);
static int lp_drain(xp_lp_f64 *lp);

static int lp_create(xp_ctx *ctx, int m, int C, int rank, int G, xp_lp_f64 **out)
{
    if (!ctx || !out || m < 1 || C < 2 || G < 1 || G > MAXR || rank < 0 || rank >= G)
        return XP_ERR_BAD_ARG;
    if (G > 1 && (C + 1) / 2 < G) return XP_ERR_BAD_ARG;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    xp_lp_f64 *lp = new xp_lp_f64();
    lp->ctx = ctx;
    memset(&lp->d, 0, sizeof lp->d);
    const int rc = lp_create_fill(ctx, m, C, rank, G, lp);
    if (rc) { // e.g. out of memory half way: give back what was allocated (destroy takes nulls)
        const std::string why = ctx->err;
        xp_lp_f64_destroy(lp);
        cudaGetLastError();
        ctx->err = why;
        return rc;
    }
    *out = lp;
    return 0;
}

static int lp_create_fill(xp_ctx *ctx, int m, int C, int rank, int G, xp_lp_f64 *lp)
{
    LpDev &d = lp->d;
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
    d.gridA = (m + TH - 1) / TH;
    if (d.gridA > 64) d.gridA = 64;
    int span = d.Cl > m ? d.Cl : m;
    d.gridB = (span + TH - 1) / TH;
    if (d.gridB > 128) d.gridB = 128;
    const size_t n = d.n, Cl = d.Cl;
#define ALLOC(ptr, bytes) XP_CUDA_OK(ctx, cudaMalloc((void **)&(ptr), (bytes)))
    ALLOC(d.tab, (size_t)m * Cl * sizeof(double));
    ALLOC(d.tgtf, Cl * sizeof(double));
    ALLOC(d.P, (size_t)KMAX * Cl * sizeof(double));
    ALLOC(d.rhsbuf, m * sizeof(double));
    ALLOC(d.sol, C * sizeof(double));
    ALLOC(lp->vc_diag, n * sizeof(double));
    ALLOC(lp->vc_rhs, n * sizeof(double));
    ALLOC(d.nvset, n + 1);
    ALLOC(d.bv2eq, n * sizeof(int32_t));
    ALLOC(d.eq2bv, m * sizeof(int32_t));
    ALLOC(d.last_piv, m * sizeof(int32_t));
    ALLOC(d.tabu, n * (size_t)d.W * sizeof(uint32_t));
    ALLOC(d.row_cnt, n * sizeof(int32_t));
    ALLOC(d.col_cnt, n * sizeof(int32_t));
    ALLOC(d.log, (size_t)d.log_cap * 3 * sizeof(int32_t));
    ALLOC(d.partA, 64 * sizeof(PartA));
    ALLOC(d.partB, 128 * sizeof(int2));
    ALLOC(d.ctr, 64);
    ALLOC(d.st, sizeof(LpState));
    ALLOC(lp->xblock, xblock_bytes(d));
    ALLOC(d.hist_lp, (size_t)NH * m * sizeof(int32_t));
#undef ALLOC
    {
        int span2 = d.Cl > m ? d.Cl : m;
        int nb = (span2 + TH - 1) / TH;
        const char *e = getenv("XP_PANEL_CTAS");
        int cap = e ? atoi(e) : 64;
        if (cap < 1) cap = 1;
        if (cap > 128) cap = 128;
        lp->panel_nb = nb < 1 ? 1 : (nb > cap ? cap : nb);
        lp->panel_rpc = (m + lp->panel_nb - 1) / lp->panel_nb;
        lp->panel_cpc = (d.Cl + lp->panel_nb - 1) / lp->panel_nb;
        const char *fm = getenv("XP_FLUSH_T_MIN");
        if (fm) lp->ft_min = atoi(fm);
        const char *fb = getenv("XP_FLUSH_BALANCED_MIN");
        if (fb) lp->ft_balanced_min = atoi(fb);
        const char *fw = getenv("XP_FLUSH_WIDE");
        if (fw) lp->ft_wide = atoi(fw);
        const char *u = getenv("XP_NO_PANEL");
        lp->use_panel = !(u && atoi(u));
        // the panel kernel keeps the open block's factors of its rows / columns in shared
        // memory; tableaux too large for that (or for one resident wave) use k_pcol / k_prow
        while (lp->use_panel && panel_smem_bytes(lp->panel_rpc, lp->panel_cpc) > ctx->smem_optin - 4096 &&
               lp->panel_nb < ctx->sm_count) {
            lp->panel_nb = lp->panel_nb * 2 > ctx->sm_count ? ctx->sm_count : lp->panel_nb * 2;
            lp->panel_rpc = (m + lp->panel_nb - 1) / lp->panel_nb;
            lp->panel_cpc = (d.Cl + lp->panel_nb - 1) / lp->panel_nb;
        }
        if (panel_smem_bytes(lp->panel_rpc, lp->panel_cpc) > ctx->smem_optin - 4096) lp->use_panel = false;
        if (lp->use_panel)
            XP_CUDA_OK(ctx, cudaFuncSetAttribute(k_panel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)ctx->smem_optin - 4096)); // one setting for every LP shape
        XP_CUDA_OK(ctx, cudaMalloc((void **)&lp->panA, 256 * sizeof(PanA)));
        XP_CUDA_OK(ctx, cudaMalloc((void **)&lp->panB, 256 * sizeof(PanB)));
        XP_CUDA_OK(ctx, cudaMalloc((void **)&lp->bar, 64));
        XP_CUDA_OK(ctx, cudaMemset(lp->bar, 0, 64));
        const char *dbg = getenv("XP_PANEL_DBG");
        if (dbg && atoi(dbg)) {
            XP_CUDA_OK(ctx, cudaMalloc((void **)&lp->panel_dbg, 128));
            XP_CUDA_OK(ctx, cudaMemset(lp->panel_dbg, 0, 128));
            XP_CUDA_OK(ctx, cudaMalloc((void **)&lp->wpanel_dbg, 128));
            XP_CUDA_OK(ctx, cudaMemset(lp->wpanel_dbg, 0, 128));
        }
        const int wrc = window_config(lp);
        if (wrc) return wrc;
    }
    XP_CUDA_OK(ctx, cudaMemset(d.st, 0, sizeof(LpState)));
    XP_CUDA_OK(ctx, cudaMemset(d.ctr, 0, 64));
    XP_CUDA_OK(ctx, cudaMemset(d.last_piv, 0xff, m * sizeof(int32_t)));
    XP_CUDA_OK(ctx, cudaMemset(lp->xblock, 0, xblock_bytes(d)));
    d.xb[rank] = lp->xblock;
    lp->attached = G == 1;
    XP_CUDA_OK(ctx, cudaMallocHost((void **)&lp->h_st, sizeof(LpState)));
    // Load every kernel of this path now: a lazy module load while another shard's
    // kernel is spinning on a peer flag would wait for that kernel, i.e. forever.
    cudaFuncAttributes fa;
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_pcol));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_prow));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_panel));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_wpanel));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_wpanel_peer));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_prow_bulk));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_flush_t<FT_TR, FT_LANES, FT_HALVES>));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_flush_w<FT_TR, 64, 4>));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_block_close));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_timeout));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_qmax_reset));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_block_snapshot));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_init));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_feas_sol));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_feas_rows));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_feas_done));
    XP_CUDA_OK(ctx, cudaFuncGetAttributes(&fa, k_checksum));
    XP_CUDA_OK(ctx, preload_flush<1>());
    XP_CUDA_OK(ctx, preload_flush<2>());
    XP_CUDA_OK(ctx, preload_flush<4>());
    XP_CUDA_OK(ctx, preload_flush<8>());
    XP_CUDA_OK(ctx, preload_flush<12>());
    XP_CUDA_OK(ctx, preload_flush<16>());
    XP_CUDA_OK(ctx, preload_flush<24>());
    XP_CUDA_OK(ctx, preload_flush<32>());
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

extern "C" int xp_lp_f64_set_block(xp_lp_f64 *lp, int pivots_per_flush)
{
    if (!lp || pivots_per_flush < 0 || pivots_per_flush > KMAX) return XP_ERR_BAD_ARG;
    if (int rc = lp_drain(lp)) return rc;
    lp->kblk = pivots_per_flush;
    return 0;
}

extern "C" int xp_lp_f64_set_window(xp_lp_f64 *lp, int width)
{
    if (!lp) return XP_ERR_BAD_ARG;
    if (int rc = lp_drain(lp)) return rc;
    lp->window = width;
    return window_config(lp);
}

extern "C" int xp_lp_f64_window(const xp_lp_f64 *lp) { return lp ? lp->d.w : 0; }

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
    void *ptrs[] = {d.tab,   d.tgtf,     d.P,     d.rhsbuf, d.sol,     lp->vc_diag, lp->vc_rhs,
                    d.nvset, d.bv2eq,    d.eq2bv, d.tabu,   d.row_cnt, d.col_cnt,   d.log,
                    d.st,    d.last_piv, d.partA, d.partB,  d.ctr,     lp->xblock, d.hist_lp};
    for (void *p : ptrs) cudaFree(p);
    cudaFree(lp->panA);
    cudaFree(lp->panB);
    cudaFree(lp->bar);
    if (lp->panel_dbg) {
        unsigned long long h[16];
        cudaMemcpy(h, lp->panel_dbg, 128, cudaMemcpyDeviceToHost);
        if (h[15])
            fprintf(stderr,
                    "[xp panel] steps %llu: A %.2f  argminA %.2f  bar1 %.2f  redA %.2f  B %.2f  minB %.2f  "
                    "bar2 %.2f  redB+book %.2f us/step\n",
                    h[15], h[0] / 1e3 / h[15], h[1] / 1e3 / h[15], h[2] / 1e3 / h[15], h[3] / 1e3 / h[15],
                    h[4] / 1e3 / h[15], h[5] / 1e3 / h[15], h[6] / 1e3 / h[15], h[7] / 1e3 / h[15]);
        cudaFree(lp->panel_dbg);
    }
    if (lp->wpanel_dbg) {
        unsigned long long h[16];
        cudaMemcpy(h, lp->wpanel_dbg, 128, cudaMemcpyDeviceToHost);
        if (h[15])
            fprintf(stderr, "[xp wpanel] steps %llu (window %d): A %.2f  exchA %.2f  B %.2f  exchB %.2f us/step\n",
                    h[15], lp->d.w, h[0] / 1e3 / h[15], h[1] / 1e3 / h[15], h[2] / 1e3 / h[15], h[3] / 1e3 / h[15]);
        cudaFree(lp->wpanel_dbg);
    }
    for (cudaEvent_t e : lp->evs) cudaEventDestroy(e);
    cudaFreeHost(lp->h_st);
    delete lp;
}

// The pass a bounded solve left owed (lookahead), on the handle's stream.  Every entry point that
// reads or reshapes the tableau calls this first.
static int lp_drain(xp_lp_f64 *lp)
{
    if (!lp->owed) return 0;
    lp->owed = false;
    const int rc = flush_w_launch(lp, lp->owed_kblk, &lp->owed_pass);
    if (rc) return rc;
    XP_CUDA_OK(lp->ctx, cudaGetLastError());
    return 0;
}

static int lp_reset(xp_lp_f64 *lp)
{
    xp_ctx *ctx = lp->ctx;
    lp->cnt_host = 0;
    lp->owed = false; // (a fresh LP: k_init clears rest_pending)
    lp->pess = true;  // (the first pricing of a solve is the slow path's)
    lp->gen_seen = 0;
    if (lp->w_max > 0) { // the window starts at its configured width
        lp->d.w = lp->w_max;
        lp->d.wwpc = lp->wwpc_max;
        lp->q_prev = lp->w_max;
    }
    const int k = lp->kblk > 0 ? lp->kblk : auto_block(lp->d);
    k_init<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(lp->d, 0u, k, 1);
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

namespace {
// Slack form of this rank's slice from a COMPACT copy of the A-columns it owns: `A` holds columns
// [a_lo, a_lo + wa) of leq (m x wa), b the constant column, tg the objective.  Replicated state
// (basis maps, constant-column replica) is written by every rank.
__global__ void k_slack_form_slice(LpDev d, const double *A, int wa, int a_lo, const double *b, const double *tg, int n)
{
    const int m = d.m, C = d.C, Cl = d.Cl;
    const size_t total = (size_t)m * Cl;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / Cl), j = d.col0 + (int)(e % Cl);
        double v;
        if (j < n) v = A[(size_t)i * wa + (j - a_lo)];
        else if (j < n + m) v = (j - n == i) ? 1.0 : 0.0;
        else v = b[i];
        d.tab[e] = v;
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < C; j += gridDim.x * blockDim.x) {
        if (j >= d.col0 && j < d.col0 + Cl) d.tgtf[j - d.col0] = j < n ? tg[j] : (j < n + m ? 0.0 : tg[n]);
        if (j < n + m) {
            d.nvset[j] = j < n;
            d.bv2eq[j] = j < n ? -1 : j - n;
        }
        if (j < m) {
            d.eq2bv[j] = n + j;
            d.rhsbuf[j] = b[j];
        }
        if (j == 0) d.st->tg_rhs = tg[n];
    }
}
} // namespace

// Works on a sharded handle too: every rank uploads only the columns of leq that fall into its
// slice (a 2-D copy; pinned host memory gives the full PCIe rate) plus the constant column and
// the objective; slack columns are generated on the device.
extern "C" int xp_lp_f64_upload_leq(xp_lp_f64 *lp, const double *leq, const double *tgtf, int n)
{
    if (!lp || !leq || !tgtf) return XP_ERR_BAD_ARG;
    xp_ctx *ctx = lp->ctx;
    LpDev &d = lp->d;
    const int m = d.m;
    if (d.C != n + m + 1) return XP_ERR_BAD_ARG;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const int a_lo = d.col0 < n ? d.col0 : n, a_hi = d.col0 + d.Cl < n ? d.col0 + d.Cl : n;
    const int wa = a_hi > a_lo ? a_hi - a_lo : 0;
    if (ctx->stage_bytes < (size_t)m * sizeof(double)) {
        if (ctx->stage) cudaFreeHost(ctx->stage);
        ctx->stage = nullptr;
        ctx->stage_bytes = 0;
        XP_CUDA_OK(ctx, cudaMallocHost(&ctx->stage, (size_t)m * sizeof(double)));
        ctx->stage_bytes = (size_t)m * sizeof(double);
    }
    double *h_b = (double *)ctx->stage;
    for (int i = 0; i < m; i++) h_b[i] = leq[(size_t)i * (n + 1) + n];
    void *scr = nullptr;
    const size_t bytes = ((size_t)m * (wa > 0 ? wa : 1) + m + (n + 1)) * sizeof(double);
    int rc = xp_ctx_scratch(ctx, bytes, &scr);
    if (rc) return rc;
    double *d_A = (double *)scr, *d_b = d_A + (size_t)m * (wa > 0 ? wa : 1), *d_tg = d_b + m;
    cudaStream_t s = ctx->stream;
    if (wa > 0)
        XP_CUDA_OK(ctx, cudaMemcpy2DAsync(d_A, (size_t)wa * sizeof(double), leq + a_lo, (size_t)(n + 1) * sizeof(double),
                                          (size_t)wa * sizeof(double), m, cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d_b, h_b, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d_tg, tgtf, (n + 1) * sizeof(double), cudaMemcpyHostToDevice, s));
    k_slack_form_slice<<<ctx->sm_count * 4, 256, 0, s>>>(d, d_A, wa, a_lo, d_b, d_tg, n);
    ctx->launches++;
    d.vc_diag = d.vc_rhs = nullptr;
    XP_CUDA_OK(ctx, cudaGetLastError());
    XP_CUDA_OK(ctx, cudaStreamSynchronize(s)); // the staging buffer is reused by the next call
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

// Windowed panel: window width and the cluster's row / column split (see xp_large_wpanel.cuh).
// A pure function of (m, C, G, requested width), so every rank of a sharded LP agrees.
static int window_config(xp_lp_f64 *lp)
{
    xp_ctx *ctx = lp->ctx;
    LpDev &d = lp->d;
    d.w = d.wrpc = d.wwpc = 0;
    lp->w_max = lp->wwpc_max = 0;
    lp->w_auto = false;
    int want = lp->window;
    if (const char *e = getenv("XP_WINDOW"))
        if (want == 0) want = atoi(e);
    if (want < 0 || !lp->use_panel) return 0;
    const int Cl0 = d.G > 1 ? shard_lo(d.C, d.G, 1) : d.C; // rank 0's slice holds the window
    const int rpc = (d.m + WNC - 1) / WNC;
    if (rpc > WTH * WRPT) return 0;
    int w;
    if (want > 0) {
        w = want < Cl0 ? want : Cl0;
    } else { // automatic: large LPs only (small ones are launch-bound either way)
        if (d.m < 2048 || Cl0 < 2048) return 0;
        w = Cl0 < 4096 ? (Cl0 & ~15) : 4096;
    }
    int wpc = (w + WNC - 1) / WNC;
    while (wpc > 1 && (wpc > WTH || wpanel_smem_bytes(rpc, wpc) + 8192 > ctx->smem_optin)) wpc--;
    if (wpc > WTH || wpanel_smem_bytes(rpc, wpc) + 8192 > ctx->smem_optin) return 0;
    if (wpc * WNC < w) w = wpc * WNC;
    if (!lp->wpanel_ready) {
        XP_CUDA_OK(ctx, cudaFuncSetAttribute(k_wpanel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        XP_CUDA_OK(ctx, cudaFuncSetAttribute(k_wpanel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)ctx->smem_optin - 8192));
        XP_CUDA_OK(ctx, cudaFuncSetAttribute(k_prow_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(2 * KMAX * WB_TH * sizeof(double))));
        lp->wpanel_ready = true;
    }
    if (d.rank == 0) { // can the device hold one such cluster at all?
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(WNC);
        cfg.blockDim = dim3(WTHB);
        cfg.dynamicSmemBytes = wpanel_smem_bytes(rpc, wpc);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = WNC;
        at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int nc = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, k_wpanel, &cfg);
        if (e != cudaSuccess || nc < 1) {
            cudaGetLastError();
            if (d.G > 1) { // the peers cannot know: a sharded LP needs the same answer everywhere
                ctx->err = "windowed panel: a 16-CTA cluster does not fit this device (set XP_WINDOW=-1 on every rank)";
                return XP_ERR_CUDA;
            }
            return 0;
        }
    }
    d.w = w;
    d.wrpc = rpc;
    d.wwpc = wpc;
    lp->w_max = w;
    lp->wwpc_max = wpc;
    lp->q_prev = w;
    // (following the entering column pays where the slice of rank 0 is much wider than the window,
    // so that the pass over the rest hides the cluster: measured at c3, 2 GPUs +11 %, 4 GPUs -4 %)
    lp->w_auto = want == 0 && Cl0 >= 2 * w && !getenv("XP_WINDOW_FIXED");
    return 0;
}

// `beside` (lookahead): a pass of k_flush_w launched right behind the cluster that starts as
// soon as the cluster has and runs on the other SMs; `after` is recorded behind that pass.
static cudaError_t wpanel_launch(xp_lp_f64 *lp, int bulk_lo = -1, int bulk_hi = -1, const ColSet *beside = nullptr,
                                 int kblk = 0, cudaEvent_t after = nullptr)
{
    const LpDev &d = lp->d;
    cudaStream_t s = lp->ctx->stream;
    lp->ctx->launches += 2;
    cudaError_t e;
    if (d.rank == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(WNC);
        cfg.blockDim = dim3(WTHB);
        cfg.dynamicSmemBytes = wpanel_smem_bytes(d.wrpc, d.wwpc);
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = WNC;
        at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const unsigned wait_units = beside ? (unsigned)((beside->ct1a - beside->ct0a) * ((d.m + FT_ROWS - 1) / FT_ROWS)) : 0u;
        e = cudaLaunchKernelEx(&cfg, k_wpanel, d, lp->wpanel_dbg, wait_units);
    } else {
        k_wpanel_peer<<<1, 32, 0, s>>>(d);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) return e;
    if (beside) {
        // (work is handed out as it goes: the CTAs that wait for the cluster's SMs join late).
        // XP_LOOKAHEAD_SERIAL=1 launches the pass the plain way -- what a profiler that serialises
        // kernels makes of it: the cluster gives up waiting and the next k_wpanel does the work (tests)
        static const bool serial = getenv("XP_LOOKAHEAD_SERIAL") != nullptr;
        if (flush_w_launch(lp, kblk, beside, 0, !serial)) return cudaErrorLaunchFailure;
        if (after && (e = cudaEventRecord(after, s)) != cudaSuccess) return e;
    }
    // the columns the window left out (and, on peers, the replicated bookkeeping)
    const int jl0 = bulk_lo >= 0 ? bulk_lo : (d.w - d.col0 > 0 ? d.w - d.col0 : 0);
    const int jl1 = bulk_hi >= 0 ? bulk_hi : d.Cl;
    int grid = (jl1 - jl0 + WB_TH - 1) / WB_TH;
    if (grid < 1) grid = 1;
    if (grid > 2 * lp->ctx->sm_count) grid = 2 * lp->ctx->sm_count;
    k_prow_bulk<<<grid, WB_TH, 2 * KMAX * WB_TH * sizeof(double), s>>>(d, 0, 0, jl0, jl1);
    return cudaGetLastError();
}

static cudaError_t panel_launch(xp_lp_f64 *lp)
{
    LpDev d = lp->d;
    PanA *pa = lp->panA;
    PanB *pb = lp->panB;
    unsigned long long *bar = lp->bar;
    unsigned long long base = lp->bar_base;
    int rpc = lp->panel_rpc, cpc = lp->panel_cpc;
    unsigned long long *dbg = lp->panel_dbg;
    void *args[] = {&d, &pa, &pb, &bar, &base, &rpc, &cpc, &dbg};
    lp->bar_base += (unsigned long long)PANEL_NBAR * lp->panel_nb;
    return cudaLaunchCooperativeKernel((void *)k_panel, dim3(lp->panel_nb), dim3(TH), args,
                                       panel_smem_bytes(rpc, cpc), lp->ctx->stream);
}

// defer: a run that stops at max_iter may leave its last closed block owed to the tableau
// (lookahead; the next solve applies it beside its first k_wpanel, lp_drain() otherwise).
static int lp_solve(xp_lp_f64 *lp, uint32_t max_iter, int rule, bool defer)
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
    const int kblk = lp->kblk > 0 ? lp->kblk : auto_block(d); // upper bound of every block
    // bounded run with automatic blocking: the device picks each block's size (next_block)
    const bool adaptive = lp->kblk == 0 && max_iter != XP_NO_ITER_LIMIT && max_iter > lp->cnt_host;
    XP_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, s));
    k_init<<<1, 32, 0, s>>>(d, max_iter, adaptive ? -kblk : kblk, 0);
    ctx->launches++;
    // One block = kblk x (k_pcol, k_prow) + k_flush, no host round trip inside;
    // the host polls the status word between batches of blocks.  Every rank of a
    // sharded LP sees the same status words, hence issues the same launches.
    int blocks = 1;
    if (max_iter != XP_NO_ITER_LIMIT && max_iter > lp->cnt_host) {
        // bounded run: the number of blocks is known, issue them without polling in between
        const unsigned long long need = ((unsigned long long)max_iter - lp->cnt_host + kblk - 1) / kblk + 1;
        blocks = need > 8 ? 8 : (int)need;
    }
    int n_prof = 0; // flushes bracketed by events in this call
    // Lookahead (windowed panel, k_flush_w; on the rank that runs k_wpanel): a finished block is
    // closed at once (k_block_close) and the tableau slice takes it one step later, beside the
    // next block's k_wpanel.  The peers of a sharded LP keep the plain order -- the launches that
    // talk to each other (k_wpanel / k_wpanel_peer, k_pcol, k_prow, k_panel) stay paired -- and so
    // does a leader whose slice is (nearly) all window: with less than a quarter of its tiles
    // to overlap, the pass is better off with the whole device and no cluster beside it.
    constexpr int TCW = 4 * FW_LANES;
    const int tiles = (d.Cl + TCW - 1) / TCW;
    const bool look = lp->use_panel && d.w > 0 && d.rank == 0 && (d.Cl & 1) == 0 && lp->ft_wide && kblk >= lp->ft_min &&
                      kblk >= lp->ft_balanced_min && d.w % TCW == 0 && 4 * (d.w / TCW) <= 3 * tiles && ctx->sm_count > 2 * WNC &&
                      !getenv("XP_NO_LOOKAHEAD");
    lp->shared_sms = look ? WNC : 0;
    ColSet owed_pass; // window tiles first, then the others
    owed_pass.ct0a = 0, owed_pass.ct1a = look ? d.w / TCW : 0, owed_pass.ct0b = owed_pass.ct1a, owed_pass.ct1b = tiles;
    owed_pass.slot = SLOT_LAG, owed_pass.close = 0;
    // Automatic window: it follows the entering column.  The reference enters the lowest-index
    // column with c_j > 0, which on the dense family climbs slowly (0.6 columns per pivot at c3 after
    // a burst in the first 200); everything the window holds has to be in place before the next
    // block can be decided, so a window of 1.5 x the highest column seen lately (+ 256, in tiles,
    // never above the configured width) keeps that serial part short.  Decided where the host polls
    // the state anyway, from replicated words (every rank of a sharded LP arrives at the same width);
    // a fresh LP starts at the configured width.  The width is a schedule, not an approximation: a
    // pivot outside the window takes the full-width kernels, and the window grows.
    auto adapt_window = [&](const LpState &h) {
        if (!lp->w_auto || lp->w_max <= 0 || d.w <= 0) return cudaSuccess;
        int q_hi = h.qmax;
        if (h.q != INT_BIG && h.q > q_hi) q_hi = h.q;
        const int q_ref = q_hi > lp->q_prev ? q_hi : lp->q_prev;
        lp->q_prev = q_hi;
        long long w_new = ((long long)q_ref + q_ref / 2 + 257 + TCW - 1) / TCW * TCW;
        if (w_new < 2 * TCW) w_new = 2 * TCW;
        if (w_new >= lp->w_max) w_new = lp->w_max;
        if ((int)w_new != d.w) {
            d.w = (int)w_new;
            d.wwpc = d.w == lp->w_max ? lp->wwpc_max : (d.w + WNC - 1) / WNC;
            if (look) owed_pass.ct1a = owed_pass.ct0b = (d.w + TCW - 1) / TCW < tiles ? (d.w + TCW - 1) / TCW : tiles;
        }
        k_qmax_reset<<<1, 1, 0, s>>>(d);
        ctx->launches++;
        return cudaGetLastError();
    };
    if (const char *e = getenv("XP_LAG_CHUNK")) owed_pass.chunk = atoi(e); // (tuning)
    if (lp->owed && (!look || lp->owed_kblk != kblk || lp->owed_pass.ct1a != owed_pass.ct1a || lp->owed_pass.chunk != owed_pass.chunk))
        if (int rc = lp_drain(lp)) return rc; // (the schedule changed since the call that left it)
    bool owed = lp->owed; // a closed block may be waiting for the tableau pass
    lp->owed = false;
    const bool windowed = lp->use_panel && d.w > 0; // (the same on every rank of a sharded LP)
    const bool self_timeout = windowed; // k_block_close / k_timeout report SIX_TIME_OUT: no extra block for it
    if (self_timeout && max_iter != XP_NO_ITER_LIMIT && max_iter > lp->cnt_host) {
        const unsigned long long need = ((unsigned long long)max_iter - lp->cnt_host + kblk - 1) / kblk;
        blocks = need > 8 ? 8 : (int)need;
    }
    // Optimistic batches: as long as k_wpanel decides every pivot alone, the kernels of
    // the general path (k_pcol, k_prow, the second k_wpanel, k_panel) find nothing to do -- five
    // empty launches per block.  The host leaves them out while the state it polls says the last
    // batch needed none (every pivot counted by k_wpanel, the next one decidable inside the window)
    // and puts them back for the next batch otherwise.  A batch that meets an exception without
    // them simply stops deciding there -- every remaining launch finds the block open and
    // returns -- until the host has looked.  Sharded: every rank polls the same replicated words
    // after the same batch, hence takes the same decision (the kernels left out are the ones
    // that talk to each other).
    const bool may_skip = windowed && !getenv("XP_NO_OPTIMISTIC");
    bool pess = !may_skip || lp->pess;
    unsigned gen_seen = lp->gen_seen; // pivots not made by k_wpanel, as of the last poll
    unsigned cnt_seen = lp->cnt_host;
    auto general_needed = [&](const LpState &h) {
        const bool go = !h.slow && !h.wfail && !h.pivot_pending && h.q != INT_BIG && h.q < d.w;
        const unsigned gen = h.cnt - h.wcnt;
        const bool moved = gen != gen_seen;
        const bool stalled = h.cnt == cnt_seen; // (whatever the reason: a batch without a single pivot)
        gen_seen = gen;
        cnt_seen = h.cnt;
        return !go || moved || stalled;
    };
    const bool dbg_tl = look && getenv("XP_BLOCK_DBG") != nullptr; // stderr: where one block's time goes
    bool dbg_done = false;
    cudaEvent_t dbg_ev[8] = {};
    if (dbg_tl)
        for (auto &e : dbg_ev) cudaEventCreate(&e);
    for (;;) {
        for (int b = 0; b < blocks; b++) {
            if (look) {
                // k_wpanel decides this block || the tableau takes the previous one; then the
                // deferred columns' pivot rows, the slow-path pair, the full-width panel
                const bool prof = owed && lp->profile && n_prof < PROF_MAX_SWEEPS;
                const bool tl = dbg_tl && b == 3 && !dbg_done;
                if (prof) XP_CUDA_OK(ctx, cudaEventRecord(lp->evs[2 * n_prof], s));
                if (tl) cudaEventRecord(dbg_ev[0], s);
                if (!owed) XP_CUDA_OK(ctx, wpanel_launch(lp)); // first block of the call: nothing is owed (every call ends drained)
                else XP_CUDA_OK(ctx, wpanel_launch(lp, -1, -1, &owed_pass, kblk, prof ? lp->evs[2 * n_prof + 1] : nullptr));
                if (prof) n_prof++;
                owed = true;
                if (tl) cudaEventRecord(dbg_ev[1], s);
                if (pess) { // (optimistic batches leave these out: see below)
                    k_pcol<<<d.gridA, TH, 0, s>>>(d);
                    if (tl) cudaEventRecord(dbg_ev[2], s);
                    k_prow<<<d.gridB, TH, 0, s>>>(d);
                    if (tl) cudaEventRecord(dbg_ev[3], s);
                    XP_CUDA_OK(ctx, wpanel_launch(lp));
                    if (tl) cudaEventRecord(dbg_ev[4], s);
                    XP_CUDA_OK(ctx, panel_launch(lp));
                    ctx->launches += 3;
                } else if (tl) {
                    cudaEventRecord(dbg_ev[2], s), cudaEventRecord(dbg_ev[3], s), cudaEventRecord(dbg_ev[4], s);
                }
                if (tl) cudaEventRecord(dbg_ev[5], s);
                k_block_close<<<1, 1024, 0, s>>>(d);
                if (tl) cudaEventRecord(dbg_ev[6], s), dbg_done = true;
                ctx->launches++;
                continue;
            }
            if (lp->use_panel && d.w > 0) {
                // windowed fast path (one cluster decides the whole block inside the window, the
                // deferred columns follow in bulk); the pair in the middle takes whatever single
                // pivot needs the slow path or lies outside the window; the full-width panel
                // finishes a block the window cannot
                XP_CUDA_OK(ctx, wpanel_launch(lp));
                if (pess) {
                    k_pcol<<<d.gridA, TH, 0, s>>>(d);
                    k_prow<<<d.gridB, TH, 0, s>>>(d);
                    XP_CUDA_OK(ctx, wpanel_launch(lp));
                    XP_CUDA_OK(ctx, panel_launch(lp));
                    ctx->launches += 3;
                }
                k_timeout<<<1, 1, 0, s>>>(d);
                ctx->launches++;
            } else if (lp->use_panel) {
                // fast path: all pivots of the block in one persistent kernel; the pair in
                // the middle takes whatever single pivot needs the slow path
                XP_CUDA_OK(ctx, panel_launch(lp));
                k_pcol<<<d.gridA, TH, 0, s>>>(d);
                k_prow<<<d.gridB, TH, 0, s>>>(d);
                XP_CUDA_OK(ctx, panel_launch(lp));
                ctx->launches += 4;
            } else {
                for (int k = 0; k < kblk; k++) {
                    k_pcol<<<d.gridA, TH, 0, s>>>(d);
                    k_prow<<<d.gridB, TH, 0, s>>>(d);
                }
                ctx->launches += 2 * kblk;
            }
            const bool prof = lp->profile && n_prof < PROF_MAX_SWEEPS;
            if (prof) XP_CUDA_OK(ctx, cudaEventRecord(lp->evs[2 * n_prof], s));
            {
                int frc = flush_launch(lp, kblk);
                if (frc) return frc;
            }
            if (prof) {
                XP_CUDA_OK(ctx, cudaEventRecord(lp->evs[2 * n_prof + 1], s));
                n_prof++;
            }
        }
        XP_CUDA_OK(ctx, cudaGetLastError());
        XP_CUDA_OK(ctx, cudaMemcpyAsync(lp->h_st, d.st, sizeof(LpState), cudaMemcpyDeviceToHost, s));
        XP_CUDA_OK(ctx, cudaStreamSynchronize(s));
        XP_CUDA_OK(ctx, adapt_window(*lp->h_st));
        pess = general_needed(*lp->h_st) || !may_skip;
        if (lp->h_st->status != XPI_RUNNING) break;
        unsigned long long left = (unsigned long long)max_iter - lp->h_st->cnt;
        int want = blocks < 8 ? blocks * 2 : 8;
        unsigned long long need = self_timeout ? (left + kblk - 1) / kblk : left / kblk + 1; // the extra block reports TIME_OUT
        blocks = (unsigned long long)want > need ? (int)need : want;
        if (blocks < 1) blocks = 1;
    }
    if (dbg_tl) {
        if (dbg_done) {
            cudaStreamSynchronize(s);
            static const char *nm[6] = {"k_wpanel || k_flush_w, k_prow_bulk", "k_pcol", "k_prow", "k_wpanel, k_prow_bulk (2nd)",
                                        "k_panel", "k_block_close"};
            for (int k = 0; k < 6; k++) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, dbg_ev[k], dbg_ev[k + 1]);
                fprintf(stderr, "[xp block] %-40s %8.1f us\n", nm[k], ms * 1e3);
            }
        }
        for (auto &e : dbg_ev) cudaEventDestroy(e);
    }
    lp->pess = pess;
    lp->gen_seen = gen_seen;
    if (!lp->h_st->rest_pending) {
        // nothing owed (the polled state is final: no launch since)
    } else if (look && defer && lp->h_st->status == XP_SIX_TIME_OUT) { // resumable: leave it to the next call
        lp->owed = true;
        lp->owed_pass = owed_pass;
        lp->owed_kblk = kblk;
    } else if (look) { // the last closed block
        const bool prof = lp->profile && n_prof < PROF_MAX_SWEEPS;
        if (prof) XP_CUDA_OK(ctx, cudaEventRecord(lp->evs[2 * n_prof], s));
        int frc = flush_w_launch(lp, kblk, &owed_pass);
        if (frc) return frc;
        if (prof) {
            XP_CUDA_OK(ctx, cudaEventRecord(lp->evs[2 * n_prof + 1], s));
            n_prof++;
        }
    }
    if (lp->profile) {
        for (int k = 0; k < n_prof; k++) {
            float ms = 0.f;
            XP_CUDA_OK(ctx, cudaEventElapsedTime(&ms, lp->evs[2 * k], lp->evs[2 * k + 1]));
            if (ms < 0.020f) continue; // an empty launch (block still open / already terminal)
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
    lp->cnt_host = lp->h_st->cnt;
    return lp->h_st->status;
}

extern "C" int xp_lp_f64_solve(xp_lp_f64 *lp, uint32_t max_iter, int rule) { return lp_solve(lp, max_iter, rule, true); }

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

extern "C" int xp_lp_f64_pass_shared_sms(const xp_lp_f64 *lp) { return lp ? lp->shared_sms : 0; }

// (debugging aid, not in the header) status, cnt, t, kblk, blk, q, slow, pivot_pending, wseq,
// wb_pending, rest_pending, rest_slot, n_touched, wcnt, qmax, wfail of the device state
extern "C" int xp_lp_f64_debug_state(xp_lp_f64 *lp, long long *out16)
{
    if (!lp || !out16) return XP_ERR_BAD_ARG;
    LpState h;
    if (cudaMemcpy(&h, lp->d.st, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) return XP_ERR_CUDA;
    const long long v[16] = {h.status, h.cnt, h.t, h.kblk, h.blk, h.q, h.slow, h.pivot_pending, h.wseq, h.wb_pending,
                             h.rest_pending, h.rest_slot, h.n_touched, h.wcnt, h.qmax, h.wfail};
    for (int k = 0; k < 16; k++) out16[k] = v[k];
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
    if (int rc = lp_drain(lp)) return rc;
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
    if (int rc = lp_drain(lp)) return rc;
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
    lp->kblk = ctx->slack_block;
    if (lp->window != ctx->slack_window) {
        rc = xp_lp_f64_set_window(lp, ctx->slack_window);
        if (rc) return rc;
    }
    rc = xp_lp_f64_upload(lp, tableau, tgtf, nvset, bvset, bv2eq, eq2bv, vc_diag, vc_rhs);
    if (rc) return rc;
    int st = lp_solve(lp, max_iter, rule, false);
    if (st >= 0) {
        rc = xp_lp_f64_download(lp, tableau, tgtf, nvset, bvset, bv2eq, eq2bv, maxv, sol, iters,
                                pivot_log, log_cap);
        if (rc) st = rc;
    }
    return st;
}

// Block size used by xp_six_slack_f64 on this ctx (0 = automatic).
extern "C" int xp_ctx_set_block(xp_ctx *ctx, int pivots_per_flush)
{
    if (!ctx || pivots_per_flush < 0 || pivots_per_flush > KMAX) return XP_ERR_BAD_ARG;
    ctx->slack_block = pivots_per_flush;
    return 0;
}

// Pricing window used by xp_six_slack_f64 / xp_six_two_stage_f64_large on this ctx
// (see xp_lp_f64_set_window).
extern "C" int xp_ctx_set_window(xp_ctx *ctx, int width)
{
    if (!ctx) return XP_ERR_BAD_ARG;
    ctx->slack_window = width;
    return 0;
}

void xp_large_release_cached(xp_ctx *ctx)
{
    if (ctx->cached_lp) xp_lp_f64_destroy((xp_lp_f64 *)ctx->cached_lp);
    ctx->cached_lp = nullptr;
    if (ctx->cached_aux) xp_lp_f64_destroy((xp_lp_f64 *)ctx->cached_aux);
    ctx->cached_aux = nullptr;
}

// Device buffers of the host-pointer entry points are kept on the ctx between calls of the
// same shape (a 1 GiB cudaMalloc / cudaFree per call would cost more than the upload).
static int cached_handle(xp_ctx *ctx, void **slot, int m, int C, xp_lp_f64 **out)
{
    xp_lp_f64 *lp = (xp_lp_f64 *)*slot;
    if (!lp || lp->d.m != m || lp->d.C != C || lp->d.G != 1) {
        if (lp) xp_lp_f64_destroy(lp);
        *slot = nullptr;
        lp = nullptr;
        int rc = xp_lp_f64_create(ctx, m, C, &lp);
        if (rc) return rc;
        *slot = lp;
    }
    *out = lp;
    return 0;
}

// ---------------------------------------------------------------------------
// TwoStageMethod on the HBM-resident path with phase 1 ON THE DEVICE
// (SURVEY 8f3): constructBasicFeasibleSolution (lpsol.h:838-988) -- auxiliary
// column, forced first pivot, aux solve, pivoting xa out, objective restoration
// by substitution, column deletion -- without the tableau ever visiting the
// host: the only upload is the caller's leq (half the slack form), the only
// downloads are O(C) vectors.  Single GPU (the sharded path takes slack forms).
// ---------------------------------------------------------------------------
namespace {

// [A | -1 | I | b], objective -xa, identity basis (lpsol.h:860-875, :1405-1433); d.C = n+m+2.
__global__ void k_aux_form(LpDev d, const double *leq, int n)
{
    const int m = d.m, C = d.C, s0 = n + 1;
    const size_t total = (size_t)m * C;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / C), j = (int)(e % C);
        double v;
        if (j < n) v = leq[(size_t)i * (n + 1) + j];
        else if (j == n) v = -1.0;
        else if (j < C - 1) v = (j - s0 == i) ? 1.0 : 0.0;
        else v = leq[(size_t)i * (n + 1) + n];
        d.tab[e] = v;
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < C; j += gridDim.x * blockDim.x) {
        d.tgtf[j] = j == n ? -1.0 : 0.0;
        if (j < C - 1) {
            d.nvset[j] = j < s0;
            d.bv2eq[j] = j < s0 ? -1 : j - s0;
        }
        if (j < m) {
            d.eq2bv[j] = s0 + j;
            d.rhsbuf[j] = leq[(size_t)j * (n + 1) + n];
        }
        if (j == 0) d.st->tg_rhs = 0.0;
    }
}

// SIX::pivot (lpsol.h:1455-1511) at a GIVEN (p, q), first half: the scaled pivot row into
// P[0][.], the multipliers -a[i][q] into F[0][0][.], the objective row and the basis maps.
// One CTA: the scalars it reads (a[p][q], c_q, eq2bv[p]) are overwritten by it.
__global__ void __launch_bounds__(1024) k_xpiv_prepare(LpDev d, int p, int q)
{
    __shared__ double s_pv, s_cq;
    __shared__ int s_bv;
    const int tid = threadIdx.x, C = d.C, n = d.n, m = d.m;
    if (tid == 0) {
        s_pv = d.tab[(size_t)p * C + q];
        s_cq = d.tgtf[q];
        s_bv = d.eq2bv[p];
    }
    __syncthreads();
    const double cq = s_cq;
    const double r = xp_div(1.0, s_pv);
    const bool r_one = xp_feq(r, 1.0), r_zero = xp_feq(r, 0.0);
    const bool cq_zero = xp_feq(cq, 0.0), cq_one = xp_feq(cq, 1.0);
    double *F = Fptr(d, 0, 0, 0);
    for (int i = tid; i < m; i += blockDim.x) F[i] = i == p ? 0.0 : -d.tab[(size_t)i * C + q]; // :1485
    const double *rowp = d.tab + (size_t)p * C;
    for (int j = tid; j < C; j += blockDim.x) {
        const double x = xp_scale(rowp[j], r, r_one, r_zero); // :1471
        d.P[j] = x;
        double t = xp_mul(x, -1.0); // objective row, :1496-1501
        if (j >= n) t = -t;
        t = cq_zero ? 0.0 : (cq_one ? t : xp_mul(t, cq));
        const double c = xp_add(t, d.tgtf[j]);
        d.tgtf[j] = c;
        if (j == n) d.st->tg_rhs = c;
    }
    if (tid == 0) { // :1504-1510
        const int bv = s_bv;
        d.nvset[q] = 0;
        d.nvset[bv] = 1;
        d.eq2bv[p] = q;
        d.bv2eq[q] = p;
        d.bv2eq[bv] = -1;
    }
}

// ... second half: the rank-1 elimination over the whole tableau (:1481-1490), row p := P[0].
__global__ void k_xpiv_apply(LpDev d, int p)
{
    const int C = d.C;
    const double *F = Fptr(d, 0, 0, 0);
    const size_t total = (size_t)d.m * C;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / C), j = (int)(e % C);
        const double x = d.P[j];
        d.tab[e] = i == p ? x : xp_add(d.tab[e], xp_mul(F[i], x));
    }
}

// lpsol.h:944-953 with FloatMat::substit (xmat.cpp:1491-1520), is_eq = false: the original
// objective with every basic variable substituted by its row, one variable after the other
// (each step reads coefficients the previous ones produced).  One CTA.
__global__ void __launch_bounds__(1024) k_restore_objective(LpDev d, const double *tg, int n_orig)
{
    __shared__ int s_list[1024];
    __shared__ int s_cnt;
    const int tid = threadIdx.x, C = d.C, rhs = d.n;
    for (int j = tid; j < C; j += blockDim.x)
        d.tgtf[j] = j < n_orig ? tg[j] : (j == rhs ? tg[n_orig] : 0.0);
    __syncthreads();
    for (int i0 = 0; i0 < rhs; i0 += 1024) {
        // basic variables of this chunk, ascending
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        const int i = i0 + tid;
        const bool basic = i < rhs && !d.nvset[i];
        const unsigned bal = __ballot_sync(0xffffffffu, basic);
        __shared__ int s_woff[33];
        if ((tid & 31) == 0) s_woff[tid >> 5] = __popc(bal);
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int w = 0; w < 32; w++) {
                const int c = s_woff[w];
                s_woff[w] = acc;
                acc += c;
            }
            s_cnt = acc;
        }
        __syncthreads();
        if (basic) s_list[s_woff[tid >> 5] + __popc(bal & ((1u << (tid & 31)) - 1u))] = i;
        __syncthreads();
        const int cnt = s_cnt;
        for (int k = 0; k < cnt; k++) {
            const int v = s_list[k];
            const double ci = d.tgtf[v];
            if (xp_feq(ci, 0.0)) continue; // uniform: every thread reads the same word
            const double *ex = d.tab + (size_t)d.bv2eq[v] * C;
            const double ev = ex[v];
            const bool skip = xp_feq(ev, 0.0);
            double s = -1.0;
            if (!xp_feq(ci, ev)) s = xp_div(-ci, ev);
            const bool s_zero = xp_feq(s, 0.0), s_one = xp_feq(s, 1.0);
            __syncthreads(); // everybody has read c_v
            for (int j = tid; j < C; j += blockDim.x) {
                double tj = d.tgtf[j];
                if (j >= rhs) tj = xp_mul(tj, -1.0);
                if (!skip) {
                    const double x = s_zero ? 0.0 : (s_one ? ex[j] : xp_mul(ex[j], s));
                    tj = xp_add(x, tj);
                }
                if (j >= rhs) tj = xp_mul(tj, -1.0);
                d.tgtf[j] = tj;
            }
            __syncthreads();
        }
        __syncthreads();
    }
    if (tid == 0) d.st->tg_rhs = d.tgtf[rhs];
}

// Drop column xa of `a` into `b` (b.C = a.C - 1) and re-index the maps (lpsol.h:956-986).
__global__ void k_drop_column(LpDev a, LpDev b, int xa)
{
    const int m = a.m, Ca = a.C, Cb = b.C;
    const size_t total = (size_t)m * Cb;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / Cb), k = (int)(e % Cb);
        b.tab[e] = a.tab[(size_t)i * Ca + (k < xa ? k : k + 1)];
    }
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Cb; k += gridDim.x * blockDim.x) {
        const int j = k < xa ? k : k + 1;
        b.tgtf[k] = a.tgtf[j];
        if (k < Cb - 1) {
            b.nvset[k] = a.nvset[j];
            b.bv2eq[k] = a.bv2eq[j];
        }
        if (k < m) {
            const int bv = a.eq2bv[k];
            b.eq2bv[k] = bv > xa ? bv - 1 : bv;
            b.rhsbuf[k] = a.rhsbuf[k];
        }
        if (k == 0) b.st->tg_rhs = a.st->tg_rhs;
    }
}

int xpiv_at(xp_lp_f64 *lp, int p, int q)
{
    xp_ctx *ctx = lp->ctx;
    cudaStream_t s = ctx->stream;
    k_xpiv_prepare<<<1, 1024, 0, s>>>(lp->d, p, q);
    k_xpiv_apply<<<ctx->sm_count * 8, 256, 0, s>>>(lp->d, p);
    k_rhs_from_tab<<<(lp->d.m + 255) / 256, 256, 0, s>>>(lp->d);
    ctx->launches += 3;
    XP_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

struct LpRef { // handles live on the ctx (cached_handle): nothing to release on the exit paths
    xp_lp_f64 *lp = nullptr;
};

} // namespace

// Upload behind the solve (no phase 1, bounded run of at most NH blocks).  The pricing window
// holds the columns that can enter, so the caller's LP goes up in column pieces: a window
// [0, ws) first, the rest of A -- [ws, n) -- in up to three more while the device is already
// deciding.  The "early" tiles (window + slack identity + constant column, the latter two
// generated on the device) run the whole bounded solve: k_wpanel decides, k_prow_bulk /
// k_flush_w keep the early tiles up to date and every closed block stays in the ring (F, records,
// pivot-row marks).  Each late piece, once landed, replays the closed blocks in order on its
// tiles -- the same operations per entry in the same order, only later.  Anything the window
// cannot decide (a pricing scan that leaves it, a failing ratio test) simply leaves the block
// open: by the time the host looks, the late tiles have caught up and the ordinary full-width
// solve continues from that state.
// Returns 1 if it took the job (state ready for xp_lp_f64_solve), 0 if the shape does not
// qualify (caller uploads the plain way), < 0 on error.
static int two_stage_streamed(xp_ctx *ctx, xp_lp_f64 *lp, int m, int n, const double *leq, double *d_leq,
                              const double *d_tg, const double *d_b, uint32_t max_iter)
{
    LpDev &d = lp->d;
    const int TCW = 4 * FW_LANES, C = d.C, w = lp->w_max > 0 ? lp->w_max : d.w; // (the configured width: a fresh LP starts there)
    const int kblk = lp->kblk > 0 ? lp->kblk : auto_block(d);
    if (getenv("XP_NO_STREAM_UPLOAD")) return 0;
    if (!lp->use_panel || w <= 0 || w % TCW || n <= w + TCW || !lp->ft_wide || (d.Cl & 1)) return 0;
    if (max_iter == XP_NO_ITER_LIMIT || max_iter < 64 || kblk < 16) return 0;
    const unsigned nb = (max_iter + kblk - 1) / kblk;
    if (nb > (unsigned)NH) return 0;
    size_t min_mb = 64; // small uploads: nothing to hide
    if (const char *e = getenv("XP_STREAM_MIN_MB")) min_mb = (size_t)atoi(e); // (tests)
    if ((size_t)m * (n + 1) * sizeof(double) < (min_mb << 20)) return 0;
    cudaStream_t s = ctx->stream;
    int rc = xp_ctx_pipe(ctx);
    if (rc) return rc;
    const int late_end = ((n + TCW - 1) / TCW) * TCW < C ? ((n + TCW - 1) / TCW) * TCW : C;
    const int tiles = (C + TCW - 1) / TCW;
    const size_t pitch = (size_t)(n + 1) * sizeof(double);
    // Column pieces, in upload order.  The first is the window the bounded run decides in: the
    // sooner it lands the sooner the device starts, so it is narrower than the resident solve's
    // window (a run of at most NH blocks enters low columns; leaving the window is handled, only
    // slower).  The last is small -- its replay is all that remains once the upload ends -- and
    // the columns between go up in two halves so the first half replays while the second lands.
    int ws = w > 3 * 1024 ? 3 * 1024 : w;
    if (const char *e = getenv("XP_STREAM_FIRST")) { // (tests)
        const int v = atoi(e) / TCW * TCW;
        if (v >= TCW && v <= w) ws = v;
    }
    int cut[5], np = 0;
    cut[np++] = 0;
    cut[np++] = ws;
    {
        const int rt = (late_end - ws) / TCW; // late tiles
        if (rt >= 12) {
            const int last = 4, mid = rt - last;
            cut[np++] = ws + (mid / 2) * TCW;
            cut[np++] = ws + mid * TCW;
        } else if (rt >= 6) {
            cut[np++] = ws + (rt - 2) * TCW;
        }
        cut[np++] = late_end;
    }
    const int npieces = np - 1; // <= 4 <= XP_PIPE_MAX
    const bool dbg = getenv("XP_STREAM_DBG") != nullptr;
    cudaEvent_t te[12] = {};
    if (dbg)
        for (auto &e : te) cudaEventCreate(&e);
    XP_CUDA_OK(ctx, cudaEventRecord(ctx->pipe_begin, s)); // d_tg, d_b and the previous call's kernels
    XP_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->pipe_copy, ctx->pipe_begin, 0));
    if (dbg) cudaEventRecord(te[0], ctx->pipe_copy);
    for (int p = 0; p < npieces; p++) {
        const int c0 = cut[p], c1 = cut[p + 1] < n ? cut[p + 1] : n;
        if (c1 > c0)
            XP_CUDA_OK(ctx, cudaMemcpy2DAsync(d_leq + c0, pitch, leq + c0, pitch, (size_t)(c1 - c0) * sizeof(double), m,
                                              cudaMemcpyHostToDevice, ctx->pipe_copy));
        XP_CUDA_OK(ctx, cudaEventRecord(ctx->pipe_up[p], ctx->pipe_copy));
        if (dbg) cudaEventRecord(te[1 + p], ctx->pipe_copy);
    }
    rc = lp_reset(lp);
    if (rc) return rc;
    const int w_full = d.w, wwpc_full = d.wwpc;
    d.w = ws;
    d.wwpc = (ws + WNC - 1) / WNC;
    struct Restore {
        LpDev &d;
        int w, wwpc;
        ~Restore() { d.w = w, d.wwpc = wwpc; }
    } restore{d, w_full, wwpc_full};
    // ---- early tiles: slack form, then the whole bounded solve ----
    XP_CUDA_OK(ctx, cudaStreamWaitEvent(s, ctx->pipe_up[0], 0));
    const int g = ctx->sm_count * 4;
    k_slack_form_cols<<<g, 256, 0, s>>>(d, d_leq, d_b, d_tg, n, 0, ws, 1);
    if (late_end < C) k_slack_form_cols<<<g, 256, 0, s>>>(d, d_leq, d_b, d_tg, n, late_end, C, 0);
    k_init<<<1, 32, 0, s>>>(d, max_iter, lp->kblk == 0 ? -kblk : kblk, 0);
    k_first_price_window<<<1, 1024, 0, s>>>(d);
    ctx->launches += 4;
    // (lookahead, as in xp_lp_f64_solve: a finished block is closed at once and the early tiles
    // take it beside the next block's k_wpanel, window tiles first)
    ColSet early;
    early.ct0a = 0, early.ct1a = ws / TCW, early.ct0b = late_end / TCW, early.ct1b = late_end < C ? tiles : late_end / TCW;
    if (!getenv("XP_NO_LOOKAHEAD")) {
        early.slot = SLOT_LAG, early.close = 0;
        for (unsigned b = 0; b < nb; b++) {
            if (b == 0) XP_CUDA_OK(ctx, wpanel_launch(lp, late_end, C));
            else XP_CUDA_OK(ctx, wpanel_launch(lp, late_end, C, &early, kblk));
            k_block_close<<<1, 1024, 0, s>>>(d);
            ctx->launches++;
        }
        rc = flush_w_launch(lp, kblk, &early); // the last closed block
        if (rc) return rc;
    } else {
        early.slot = -1, early.close = 1;
        for (unsigned b = 0; b < nb; b++) {
            XP_CUDA_OK(ctx, wpanel_launch(lp, late_end, C));
            k_block_snapshot<<<32, 256, 0, s>>>(d);
            ctx->launches++;
            rc = flush_w_launch(lp, kblk, &early);
            if (rc) return rc;
        }
    }
    // ---- late pieces: slack form when each has landed, then the closed blocks in order ----
    if (dbg) cudaEventRecord(te[6], s);
    auto bulk_grid = [&](int c0, int c1) {
        int bg = (c1 - c0 + WB_TH - 1) / WB_TH;
        return bg > 2 * ctx->sm_count ? 2 * ctx->sm_count : bg;
    };
    for (int p = 1; p < npieces; p++) {
        const int c0 = cut[p], c1 = cut[p + 1];
        XP_CUDA_OK(ctx, cudaStreamWaitEvent(s, ctx->pipe_up[p], 0));
        k_slack_form_cols<<<g, 256, 0, s>>>(d, d_leq, d_b, d_tg, n, c0, c1, 0);
        ctx->launches++;
        ColSet late;
        late.ct0a = c0 / TCW, late.ct1a = c1 / TCW, late.ct0b = late.ct1b = 0;
        late.close = 0;
        for (unsigned b = 0; b < nb; b++) {
            k_prow_bulk<<<bulk_grid(c0, c1), WB_TH, 2 * KMAX * WB_TH * sizeof(double), s>>>(d, 1, (int)b, c0, c1);
            ctx->launches++;
            late.slot = (int)b;
            rc = flush_w_launch(lp, kblk, &late);
            if (rc) return rc;
        }
        if (dbg) cudaEventRecord(te[6 + p], s);
    }
    XP_CUDA_OK(ctx, cudaGetLastError());
    XP_CUDA_OK(ctx, cudaMemcpyAsync(lp->h_st, d.st, sizeof(LpState), cudaMemcpyDeviceToHost, s));
    XP_CUDA_OK(ctx, cudaStreamSynchronize(s));
    if (dbg) {
        float t = 0;
        fprintf(stderr, "[xp stream] pieces");
        for (int p = 0; p < npieces; p++) {
            cudaEventElapsedTime(&t, te[0], te[1 + p]);
            fprintf(stderr, " [%d,%d) up at %.2f ms;", cut[p], cut[p + 1], t);
        }
        cudaEventElapsedTime(&t, te[0], te[6]);
        fprintf(stderr, " early solve done at %.2f;", t);
        for (int p = 1; p < npieces; p++) {
            cudaEventElapsedTime(&t, te[0], te[6 + p]);
            fprintf(stderr, " replay %d done at %.2f;", p, t);
        }
        fprintf(stderr, " (cnt %u, t %d, status %d)\n", lp->h_st->cnt, lp->h_st->t, lp->h_st->status);
        for (auto &ev : te) cudaEventDestroy(ev);
    }
    if (lp->h_st->t > 0) { // a block was left open (exception inside the window run): its pivot rows for the late tiles
        k_prow_bulk<<<bulk_grid(ws, late_end), WB_TH, 2 * KMAX * WB_TH * sizeof(double), s>>>(d, 2, 0, ws, late_end);
        ctx->launches++;
        XP_CUDA_OK(ctx, cudaGetLastError());
    }
    lp->cnt_host = lp->h_st->cnt;
    return 1;
}

// Variable constraints of a handle: `vd` / `vr` (n_struct entries: diagonal and constant column of
// the caller's `vc`, lpsol.h:798-802) for the structural variables, -1 / 0 for everything the
// solver adds itself (slacks; the auxiliary variable at column `xa` if xa >= 0).
static int set_vc(xp_lp_f64 *lp, const double *vd, const double *vr, int n_struct, int xa)
{
    xp_ctx *ctx = lp->ctx;
    LpDev &d = lp->d;
    d.vc_diag = d.vc_rhs = nullptr;
    if (!vd && !vr) return 0;
    std::vector<double> hd(d.n, -1.0), hr(d.n, 0.0);
    for (int j = 0; j < n_struct; j++) {
        const int k = (xa >= 0 && j >= xa) ? j + 1 : j;
        if (vd) hd[k] = vd[j];
        if (vr) hr[k] = vr[j];
    }
    XP_CUDA_OK(ctx, cudaMemcpyAsync(lp->vc_diag, hd.data(), d.n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(lp->vc_rhs, hr.data(), d.n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    XP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream)); // hd / hr go out of scope
    d.vc_diag = lp->vc_diag;
    d.vc_rhs = lp->vc_rhs;
    return 0;
}

extern "C" int xp_six_two_stage_f64_large_vc(xp_ctx *ctx, int m, int n, const double *leq, const double *tgtf,
                                             const double *vc_diag, const double *vc_rhs, uint32_t max_iter,
                                             int rule, int32_t *status, double *maxv, double *slack_sol,
                                             double *tgtf_out, int32_t *eq2bv, uint32_t *iters, uint32_t *pivots);

extern "C" int xp_six_two_stage_f64_large(xp_ctx *ctx, int m, int n, const double *leq,
                                          const double *tgtf, uint32_t max_iter, int rule,
                                          int32_t *status, double *maxv, double *slack_sol,
                                          double *tgtf_out, int32_t *eq2bv, uint32_t *iters,
                                          uint32_t *pivots)
{
    return xp_six_two_stage_f64_large_vc(ctx, m, n, leq, tgtf, nullptr, nullptr, max_iter, rule, status, maxv,
                                         slack_sol, tgtf_out, eq2bv, iters, pivots);
}

// State the last xp_six_two_stage_f64_large[_vc] / xp_six_slack_f64 call left on the device, as
// SIX::TwoStageMethod hands it back through its IN OUT arguments (lpsol.h:291-301): the final
// tableau m x C, objective row, basis maps.  Any pointer may be NULL.
extern "C" int xp_ctx_last_lp_download(xp_ctx *ctx, double *tableau, double *tgtf, uint8_t *nvset, uint8_t *bvset,
                                       int32_t *bv2eq, int32_t *eq2bv)
{
    if (!ctx || !ctx->cached_lp) return XP_ERR_BAD_ARG;
    return xp_lp_f64_download((xp_lp_f64 *)ctx->cached_lp, tableau, tgtf, nvset, bvset, bv2eq, eq2bv, nullptr, nullptr,
                              nullptr, nullptr, 0);
}

namespace {

struct Stage1 { // the stage1 decision (lpsol.h:1794-1803) and the forced first pivot row (:894-904)
    bool pos = false, bneg = false;
    int prow = 0;
};

// (-A^T | c) and -b^T of calcDualMaxm (lpsol.h:1602-1623) from the primal on the device:
// dual[i][j] = (-1) * A[j][i], dual[i][mp] = c_i, dual objective = (-1) * b_j, constant (-1) * 0.
// 32 x 32 tiles through shared memory: both sides coalesced.
__global__ void k_build_dual(const double *leq, const double *tg, int mp, int np, double *dleq, double *dtg, double *db)
{
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
    const int tiles_i = (np + 31) / 32, tiles_j = (mp + 31) / 32;
    for (int tix = blockIdx.x; tix < tiles_i * tiles_j; tix += gridDim.x) {
        const int bi = (tix % tiles_i) * 32, bj = (tix / tiles_i) * 32; // dual rows bi.., dual columns bj..
        for (int r = ty; r < 32; r += 8) { // primal row bj + r, primal columns bi + tx
            const int pj = bj + r, pi = bi + tx;
            tile[r][tx] = (pj < mp && pi < np) ? leq[(size_t)pj * (np + 1) + pi] : 0.0;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int di = bi + r, dj = bj + tx;
            if (di < np && dj < mp) dleq[(size_t)di * (mp + 1) + dj] = xp_mul(tile[tx][r], -1.0);
        }
        __syncthreads();
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
        dleq[(size_t)i * (mp + 1) + mp] = tg[i];
        db[i] = tg[i];
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= mp; j += gridDim.x * blockDim.x)
        dtg[j] = xp_mul(j < mp ? leq[(size_t)j * (np + 1) + np] : 0.0, -1.0);
}

} // namespace

// The normalised LP (m x (n+1), objective, constant column) is either still on the host (`leq`
// non-null: uploaded here, behind the solve where that pays) or already in d_leq / d_tg / d_b.
static int two_stage_impl(xp_ctx *ctx, int m, int n, const double *leq, Stage1 s1, double *d_leq, double *d_tg,
                          double *d_b, const double *vc_diag, const double *vc_rhs, uint32_t max_iter, int rule,
                          int32_t *status, double *maxv, double *slack_sol, double *tgtf_out, int32_t *eq2bv,
                          uint32_t *iters, uint32_t *pivots)
{
    cudaStream_t s = ctx->stream;
    const int Cm = n + m + 1, prow = s1.prow;
    const bool aux = !s1.pos || s1.bneg;
    int rc = 0;
    if (maxv) *maxv = 0.0;
    if (iters) *iters = 0;
    if (pivots) *pivots = 0;
    unsigned n_piv = 0;
    LpRef A, M;
    rc = cached_handle(ctx, &ctx->cached_lp, m, Cm, &M.lp);
    if (rc) return rc;
    M.lp->kblk = ctx->slack_block;
    if (M.lp->window != ctx->slack_window) {
        rc = xp_lp_f64_set_window(M.lp, ctx->slack_window);
        if (rc) return rc;
    }
    int streamed = 0;
    if (!aux) {
        rc = set_vc(M.lp, vc_diag, vc_rhs, n, -1); // (only the optimal exit reads it)
        if (rc) return rc;
        if (leq) streamed = two_stage_streamed(ctx, M.lp, m, n, leq, d_leq, d_tg, d_b, max_iter);
        if (streamed < 0) return streamed;
    }
    if (!streamed) {
    // No phase 1: the rows go up in chunks on the copy stream and k_slack_form builds [A | I | b]
    // chunk by chunk behind them, so the slack form is complete one chunk's kernel after the last
    // byte has arrived.  Phase 1 (rare) needs the whole LP first: one copy.
    const size_t row_bytes = (size_t)(n + 1) * sizeof(double);
    int n_chunks = 1;
    if (leq && !aux && (size_t)m * row_bytes >= ((size_t)64 << 20)) n_chunks = XP_PIPE_MAX < m ? XP_PIPE_MAX : m;
    if (n_chunks > 1) {
        rc = xp_ctx_pipe(ctx);
        if (rc) return rc;
        XP_CUDA_OK(ctx, cudaEventRecord(ctx->pipe_begin, s)); // d_tg and the previous call's kernels
        XP_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->pipe_copy, ctx->pipe_begin, 0));
        for (int c = 0; c < n_chunks; c++) {
            const int r0 = (int)((long long)m * c / n_chunks), r1 = (int)((long long)m * (c + 1) / n_chunks);
            XP_CUDA_OK(ctx, cudaMemcpyAsync(d_leq + (size_t)r0 * (n + 1), leq + (size_t)r0 * (n + 1),
                                            (size_t)(r1 - r0) * row_bytes, cudaMemcpyHostToDevice, ctx->pipe_copy));
            XP_CUDA_OK(ctx, cudaEventRecord(ctx->pipe_up[c], ctx->pipe_copy));
        }
    } else if (leq) {
        XP_CUDA_OK(ctx, cudaMemcpyAsync(d_leq, leq, (size_t)m * row_bytes, cudaMemcpyHostToDevice, s));
    }
    if (!aux) {
        for (int c = 0; c < n_chunks; c++) {
            const int r0 = (int)((long long)m * c / n_chunks), r1 = (int)((long long)m * (c + 1) / n_chunks);
            if (n_chunks > 1) XP_CUDA_OK(ctx, cudaStreamWaitEvent(s, ctx->pipe_up[c], 0));
            k_slack_form<<<ctx->sm_count * 4, 256, 0, s>>>(M.lp->d, d_leq, d_tg, n, r0, r1);
            ctx->launches++;
        }
        XP_CUDA_OK(ctx, cudaGetLastError());
    } else {
        const int xa = n, Ca = Cm + 1;
        rc = cached_handle(ctx, &ctx->cached_aux, m, Ca, &A.lp);
        if (rc) return rc;
        A.lp->kblk = ctx->slack_block;
        if (A.lp->window != ctx->slack_window) {
            rc = xp_lp_f64_set_window(A.lp, ctx->slack_window);
            if (rc) return rc;
        }
        k_aux_form<<<ctx->sm_count * 4, 256, 0, s>>>(A.lp->d, d_leq, n);
        ctx->launches++;
        rc = set_vc(A.lp, vc_diag, vc_rhs, n, xa); // the auxiliary solve's own optimal exit checks them too
        if (rc) return rc;
        rc = xpiv_at(A.lp, prow, xa); // forced first pivot, :892-908
        if (rc) return rc;
        n_piv++;
        rc = lp_reset(A.lp);
        if (rc) return rc;
        int st = lp_solve(A.lp, max_iter, rule, false);
        if (st < 0) return st;
        n_piv += A.lp->h_st->cnt;
        if (pivots) *pivots = n_piv;
        if (st != XP_SIX_SUCC || !xp_feq(A.lp->h_st->maxv, 0.0)) { // :912-922
            *status = XP_SIX_NO_PRI_FEASIBLE_SOL;
            return 0;
        }
        // xa still basic: pivot it out on the first non-basic column with a non-zero entry in
        // its row, :924-941 (one row and the basis flags come to the host: O(C) bytes)
        std::vector<uint8_t> nv(Ca - 1);
        std::vector<int32_t> b2e(Ca - 1);
        XP_CUDA_OK(ctx, cudaMemcpyAsync(nv.data(), A.lp->d.nvset, Ca - 1, cudaMemcpyDeviceToHost, s));
        XP_CUDA_OK(ctx, cudaMemcpyAsync(b2e.data(), A.lp->d.bv2eq, (Ca - 1) * sizeof(int32_t),
                                        cudaMemcpyDeviceToHost, s));
        XP_CUDA_OK(ctx, cudaStreamSynchronize(s));
        if (!nv[xa]) {
            const int eqnum = b2e[xa];
            std::vector<double> row(Ca);
            XP_CUDA_OK(ctx, cudaMemcpyAsync(row.data(), A.lp->d.tab + (size_t)eqnum * Ca, Ca * sizeof(double),
                                            cudaMemcpyDeviceToHost, s));
            XP_CUDA_OK(ctx, cudaStreamSynchronize(s));
            int cand = -1;
            for (int j = 0; j < Ca - 1; j++)
                if (nv[j] && !xp_feq(row[j], 0.0)) {
                    cand = j;
                    break;
                }
            if (cand < 0) { // reference ASSERTs (:937)
                *status = XP_ERR_REFERENCE_UB;
                return 0;
            }
            rc = xpiv_at(A.lp, eqnum, cand);
            if (rc) return rc;
            n_piv++;
        }
        k_restore_objective<<<1, 1024, 0, s>>>(A.lp->d, d_tg, n); // :944-953
        k_drop_column<<<ctx->sm_count * 4, 256, 0, s>>>(A.lp->d, M.lp->d, xa); // :956-986
        ctx->launches += 2;
        XP_CUDA_OK(ctx, cudaGetLastError());
        rc = set_vc(M.lp, vc_diag, vc_rhs, n, -1);
        if (rc) return rc;
    }
    rc = lp_reset(M.lp);
    if (rc) return rc;
    } // !streamed
    int st = lp_solve(M.lp, max_iter, rule, false);
    if (st < 0) return st;
    uint32_t it = 0;
    rc = xp_lp_f64_download(M.lp, nullptr, tgtf_out, nullptr, nullptr, nullptr, eq2bv, maxv, slack_sol, &it,
                            nullptr, 0);
    if (rc) return rc;
    if (iters) *iters = it;
    if (pivots) *pivots = n_piv + it;
    *status = st;
    return 0;
}


extern "C" int xp_six_two_stage_f64_large_vc(xp_ctx *ctx, int m, int n, const double *leq, const double *tgtf,
                                             const double *vc_diag, const double *vc_rhs, uint32_t max_iter,
                                             int rule, int32_t *status, double *maxv, double *slack_sol,
                                             double *tgtf_out, int32_t *eq2bv, uint32_t *iters, uint32_t *pivots)
{
    if (!ctx || m < 1 || n < 1 || !leq || !tgtf || !status) return XP_ERR_BAD_ARG;
    if (rule != XP_RULE_REFERENCE) return XP_ERR_BAD_ARG;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    // stage1 decision on the caller's arrays, :1794-1803 (the constant column is gathered into a
    // small pinned buffer on the way: the piecewise upload sends it ahead of the matrix)
    if (ctx->stage_bytes < (size_t)m * sizeof(double)) {
        if (ctx->stage) cudaFreeHost(ctx->stage);
        ctx->stage = nullptr;
        ctx->stage_bytes = 0;
        XP_CUDA_OK(ctx, cudaMallocHost(&ctx->stage, (size_t)m * sizeof(double)));
        ctx->stage_bytes = (size_t)m * sizeof(double);
    }
    double *h_b = (double *)ctx->stage;
    Stage1 s1;
    for (int j = 0; j < n; j++) s1.pos |= tgtf[j] > 0.0;
    for (int i = 0; i < m; i++) {
        const double b = leq[(size_t)i * (n + 1) + n];
        h_b[i] = b;
        s1.bneg |= b < 0.0;
        if (h_b[s1.prow] > b) s1.prow = i;
    }
    // the caller's LP in device scratch (the only bulk upload of the call)
    void *scr = nullptr;
    const size_t in_elems = (size_t)m * (n + 1) + (n + 1) + m;
    int rc = xp_ctx_scratch(ctx, in_elems * sizeof(double), &scr);
    if (rc) return rc;
    double *d_leq = (double *)scr, *d_tg = d_leq + (size_t)m * (n + 1), *d_b = d_tg + (n + 1);
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d_tg, tgtf, (n + 1) * sizeof(double), cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d_b, h_b, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, s));
    return two_stage_impl(ctx, m, n, leq, s1, d_leq, d_tg, d_b, vc_diag, vc_rhs, max_iter, rule, status, maxv,
                          slack_sol, tgtf_out, eq2bv, iters, pivots);
}

// TwoStageMethod on the explicit DUAL of a normalised primal LP (SIX::minm, lpsol.h:1661-1732 with
// calcDualMaxm :1585-1655), the dual built on the device: the caller hands over the PRIMAL
// (leq mp x (np+1), tgtf np+1; x >= 0, no equalities) exactly as it would to the max entry, the
// transposition -A^T, the dual objective -b and its slack form never exist on the host.  Outputs
// are those of the dual LP (np rows, mp variables): slack_sol / tgtf_out mp+np+1 entries, eq2bv np.
extern "C" int xp_six_two_stage_f64_large_dual(xp_ctx *ctx, int mp, int np, const double *leq, const double *tgtf,
                                               uint32_t max_iter, int rule, int32_t *status, double *maxv,
                                               double *slack_sol, double *tgtf_out, int32_t *eq2bv,
                                               uint32_t *iters, uint32_t *pivots)
{
    if (!ctx || mp < 1 || np < 1 || !leq || !tgtf || !status) return XP_ERR_BAD_ARG;
    if (rule != XP_RULE_REFERENCE) return XP_ERR_BAD_ARG;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int m = np, n = mp; // the dual's shape
    Stage1 s1;                // dual objective -b_j, dual constant column c_i
    for (int j = 0; j < mp; j++) s1.pos |= -leq[(size_t)j * (np + 1) + np] > 0.0;
    for (int i = 0; i < np; i++) {
        s1.bneg |= tgtf[i] < 0.0;
        if (tgtf[s1.prow] > tgtf[i]) s1.prow = i;
    }
    void *scr = nullptr;
    const size_t prim = (size_t)mp * (np + 1) + (np + 1), dual = (size_t)m * (n + 1) + (n + 1) + m;
    int rc = xp_ctx_scratch(ctx, (prim + dual) * sizeof(double), &scr);
    if (rc) return rc;
    double *d_leq = (double *)scr, *d_tg = d_leq + (size_t)m * (n + 1), *d_b = d_tg + (n + 1);
    double *p_leq = d_b + m, *p_tg = p_leq + (size_t)mp * (np + 1);
    XP_CUDA_OK(ctx, cudaMemcpyAsync(p_leq, leq, (size_t)mp * (np + 1) * sizeof(double), cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(p_tg, tgtf, (size_t)(np + 1) * sizeof(double), cudaMemcpyHostToDevice, s));
    k_build_dual<<<ctx->sm_count * 4, 256, 0, s>>>(p_leq, p_tg, mp, np, d_leq, d_tg, d_b);
    ctx->launches++;
    XP_CUDA_OK(ctx, cudaGetLastError());
    return two_stage_impl(ctx, m, n, nullptr, s1, d_leq, d_tg, d_b, nullptr, nullptr, max_iter, rule, status, maxv,
                          slack_sol, tgtf_out, eq2bv, iters, pivots);
}

// Position-keyed checksums (xp_lp_f64_checksum) of the tableau and objective row the last
// xp_six_two_stage_f64_large / xp_six_slack_f64 call on this ctx left on the device: the final
// tableau never travels to the host on that path, this is how a test still sees all of it.
extern "C" int xp_ctx_last_lp_checksum(xp_ctx *ctx, uint64_t *sum_tableau, uint64_t *sum_tgtf)
{
    if (!ctx || !ctx->cached_lp) return XP_ERR_BAD_ARG;
    return xp_lp_f64_checksum((xp_lp_f64 *)ctx->cached_lp, sum_tableau, sum_tgtf);
}
