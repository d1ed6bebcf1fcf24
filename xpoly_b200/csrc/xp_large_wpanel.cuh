// Windowed panel (included by xp_large_f64.cu inside its anonymous namespace).
//
// The reference prices the LOWEST-index non-basic column with c_j > 0 (lpsol.h:1054-1069), so
// the column that enters next is decided by the objective coefficients to its left alone, and
// on the dense LPs of SURVEY 8(d) it stays among the first few thousand columns for thousands
// of pivots.  k_wpanel exploits that: it carries the objective row, the pivot rows P[s][.] and
// the pricing state only for the WINDOW = global columns [0, w); the decision chain of a pivot
//     entering column (all m rows: strided read + replay)  -> ratio test -> pivot row p
//     -> row p over the window (read + replay) -> objective row -> pricing -> next column
// then fits one thread-block CLUSTER of 16 CTAs: CTA c keeps the multipliers F[0..t) of its
// m/16 rows and the pivot rows P[0..t) of its w/16 window columns in shared memory, partial
// arg-mins travel through distributed shared memory and the two reductions per pivot are
// hardware cluster barriers (no global-memory barrier, no L2 round trip for the partials).
// What the window leaves out -- P[s][j] and c_j for j >= w -- depends on the decisions only
// (p_s, 1/pivot, c_q and the multipliers F[s][p_u]); k_prow_bulk computes it afterwards for
// all deferred columns at once from one record per pivot (WRec).  Together the two kernels
// leave exactly the state the full-width k_panel would have left (same operations in the same
// order per entry, hence the same bits).  If pricing finds nothing inside the window the run
// ends there and the full-width slow path re-prices (sp_select), so the window is a schedule,
// never an approximation.
//
// Column-sharded LPs: the window lies in rank 0's slice, so rank 0 ("leader") decides the whole
// run alone, writing each multiplier column and record into the peers' exchange blocks as it
// goes; the peers wait for ONE flag per run (k_wpanel_peer) instead of exchanging two words
// per pivot, then every rank runs k_prow_bulk on its own columns (peers also replay the
// replicated bookkeeping -- basis maps, tabu table, constant column -- from the records).
//
// Reference semantics implemented here: pricing lpsol.h:1054-1069, findPivotBV :552-663 (pass 1
// and, when it finds no row, pass 2), pivot :1455-1511, genPair :1156.

constexpr int WNC = 16;        // CTAs of the cluster (non-portable size on sm_100a)
constexpr int WTH = 512;       // worker threads per CTA
constexpr int WTHB = WTH + 32; // + one warp that keeps the books (CTA 0) and never delays the workers
constexpr int WRPT = 2;        // rows per worker thread, at most
constexpr unsigned long long KEY_NONE = ~0ULL;

__host__ __device__ inline size_t wpanel_smem_bytes(int rpc, int wpc)
{
    // sF[KMAX][rpc] | sP[KMAX][wpc] | s_rh[2][rpc] | s_tg[wpc] | s_e2b[rpc] s_lp[rpc] s_nv[wpc] s_rc[wpc] | s_el[rpc]
    return (size_t)8 * ((size_t)KMAX * rpc + (size_t)KMAX * wpc + 2 * (size_t)rpc + wpc) +
           (size_t)4 * (2 * (size_t)rpc + 2 * (size_t)wpc) + (((size_t)rpc + 15) & ~(size_t)15);
}

// Order-preserving map of a double onto unsigned integers (-0 and +0 compare equal in the
// reference's `minbval > v`, so -0 is folded onto +0 first).
__device__ __forceinline__ unsigned long long f64_key(double v)
{
    const long long b = __double_as_longlong(xp_add(v, 0.0));
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ULL));
}

// Warp arg-min of (key, index): lowest key, ties -> lowest index (the reference's first strict
// minimum in row order).  Empty lanes pass (KEY_NONE, INT_BIG).  Result in every lane.
__device__ __forceinline__ void warp_min_key(unsigned long long &k, int &i)
{
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    const int mi = __reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? i : INT_BIG);
    k = ((unsigned long long)mh << 32) | ml;
    i = mi;
}

// Partial results of one CTA as the other CTAs of the cluster see them (two slots each: an
// exchange only reuses a slot after another cluster barrier has passed in between).
struct WXchg {
    unsigned long long k[2][WNC]; // ratio test: key of the CTA's best row
    double rh[2][WNC], a[2][WNC]; //   its constant term and entry in the entering column
    int i[2][WNC], bv[2][WNC], s0[2][WNC];
    double cc[2][WNC]; // pricing: c_j of the CTA's candidate
    unsigned c[2][WNC]; //   candidate | anypos << 31
};

__global__ void __launch_bounds__(WTHB, 1) k_wpanel(LpDev d, unsigned long long *dbg, unsigned wait_units)
{
    cg::cluster_group cl = cg::this_cluster();
    extern __shared__ double s_dyn[];
    __shared__ WXchg X;
    __shared__ unsigned long long s_wk[WTH / 32];
    __shared__ double s_wrh[WTH / 32], s_wa[WTH / 32], s_wcc[WTH / 32];
    __shared__ int s_wi[WTH / 32], s_wbv[WTH / 32], s_ws0[WTH / 32], s_wc[WTH / 32], s_wany[WTH / 32];
    __shared__ double s_pq[KMAX], s_fp[KMAX];
    __shared__ int s_bad;
    LpState *st = d.st;
    // Lookahead: the previous block may still be owed to the tableau (k_block_close).  The pass
    // that applies it -- k_flush_w with SLOT_LAG, launched behind this kernel with programmatic
    // stream serialization -- may start now, on the SMs this cluster leaves; it takes the window
    // tiles first and counts them in ctr[4].  (The flag is read before the trigger: the pass
    // clears it when it is through.)
    const bool owed = wait_units != 0 && st->rest_pending != 0;
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = (int)cl.block_rank();
    const bool worker = tid < WTH;
    const int n = d.n, m = d.m, Cl = d.Cl, G = d.G, rpc = d.wrpc, wpc = d.wwpc;

    int t = st->t;
    const int kblk = st->kblk, par = st->blk & (NH - 1);
    unsigned cnt = st->cnt;
    const unsigned max_iter = st->max_iter;
    int q = st->q, zero_upto = st->zero_upto, anypos = st->anypos;
    // every CTA (and, sharded, every rank) evaluates the same predicate on the same state
    const bool go = st->status == XPI_RUNNING && !st->slow && !st->pivot_pending && q != INT_BIG &&
                    q < d.w && t < kblk && cnt < max_iter;
    if (c == 0 && tid == 0) st->wb_pending = 0; // (k_prow_bulk of the previous run is over; nobody else reads it here)
    if (!go) return;
    if (owed) { // the window columns (and nothing else of the tableau) are read below
        // CTA 0 watches the counter for the whole cluster.  If the pass does not show up -- the
        // driver may serialise the two launches (profilers do) -- the cluster leaves without
        // having touched anything: the pass then runs behind it, and the second k_wpanel of
        // the block's launch sequence finds the tableau complete and does the work.
        if (c == 0 && tid == 0) {
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
            unsigned spins = 0;
            int ok = 1;
            while (ld_acquire_gpu_u32(&d.ctr[4]) < wait_units) {
                if ((++spins & 255u) == 0) {
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                    if (t1 - t0 > 20000000ULL) { // 20 ms
                        ok = 0;
                        break;
                    }
                }
            }
            s_bad = !ok;
        }
        cl.sync();
        const int bad0 = *cl.map_shared_rank(&s_bad, 0);
        cl.sync(); // (s_bad is reused below)
        if (bad0) return;
        __threadfence(); // order the tableau reads below behind CTA 0's acquire
    }
    const int t_in = t;
    const unsigned wseq = st->wseq + 1; // number of this windowed launch
    double tg_rhs = st->tg_rhs;
    unsigned n_log = st->n_log;
    int n_touched = st->n_touched;

    double *sF = s_dyn, *sP = sF + (size_t)KMAX * rpc;
    double *s_rh = sP + (size_t)KMAX * wpc, *s_tg = s_rh + 2 * (size_t)rpc;
    int *s_e2b = (int *)(s_tg + wpc), *s_lp = s_e2b + rpc, *s_nv = s_lp + rpc, *s_rc = s_nv + wpc;
    unsigned char *s_el = (unsigned char *)(s_rc + wpc);
    const int r_lo = min(m, c * rpc), nrows = min(m, r_lo + rpc) - r_lo;
    const int c_lo = min(d.w, c * wpc), ncols = min(d.w, c_lo + wpc) - c_lo;

    if (tid == 0) s_bad = 0;
    __syncthreads();
    if (G > 1 && tid > 0 && tid < G) {
        // the peers must have consumed the previous windowed launch (its exit words are single
        // buffered) -- which also means their flush of the block before last is over, so the
        // F / P / record buffers of this block's parity are free
        const unsigned long long *w = &((const XHdr *)d.xb[0])->wack[tid];
        const unsigned long long t0 = clock64();
        unsigned spins = 0;
        while (ld_acquire_sys(w) + 1 < wseq)
            if ((++spins & 1023u) == 0 && clock64() - t0 > SPIN_LIMIT) {
                s_bad = 1; // keep going (cluster-uniform control flow); reported at the end
                break;
            }
    }
    for (int li = tid; li < nrows; li += WTHB) {
        s_e2b[li] = d.eq2bv[r_lo + li];
        s_lp[li] = d.last_piv[r_lo + li];
        s_rh[li] = d.rhsbuf[r_lo + li];
    }
    for (int lj = tid; lj < ncols; lj += WTHB) {
        const int g = c_lo + lj;
        s_tg[lj] = d.tgtf[g];
        s_nv[lj] = g < n ? (int)d.nvset[g] : 0;
        s_rc[lj] = g < n ? d.row_cnt[g] : INT_BIG;
    }
    // resuming inside an open block: bring its factors into shared memory
    for (int e = tid; e < t * nrows; e += WTHB) {
        const int s = e / nrows, li = e - s * nrows;
        sF[(size_t)s * rpc + li] = ld_cg(Fptr(d, d.rank, par, s) + r_lo + li);
    }
    for (int e = tid; e < t * ncols; e += WTHB) {
        const int s = e / ncols, lj = e - s * ncols;
        sP[(size_t)s * wpc + lj] = ld_cg(d.P + (size_t)s * Cl + c_lo + lj);
    }
    double cq = ld_cg(d.tgtf + q);
    cl.sync();

    unsigned long long tprev = 0;
    if (dbg && c == 0 && tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tprev));
#define WPANEL_T(k)                                              \
    if (dbg && c == 0 && tid == 0) {                             \
        unsigned long long now__;                                \
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now__)); \
        dbg[k] += now__ - tprev;                                 \
        tprev = now__;                                           \
    }
    int cur = 0;      // s_rh buffer holding the constant column as of now
    unsigned xn = 0;  // exchanges so far (slot = xn & 1)
    int slow_out = 0, fail_out = 0; // how the run ended: re-price on the slow path / the ratio test on q failed
    int qmx = q; // highest entering column of this run (the host sizes the window by it)
    WRec *recs[MAXR];
    for (int r = 0; r < G; r++) recs[r] = (WRec *)(d.xb[r] + xoff_rec(d, par));

    while (t < kblk && cnt < max_iter) {
        // ============ phase A: entering column, multipliers, ratio test (lpsol.h:552-663) ============
        const int ql = q; // the leader's slice starts at column 0
        double a0[WRPT], rh[WRPT];
        int rbv[WRPT], s0[WRPT], cc[WRPT];
        unsigned tw[WRPT];
        bool ok[WRPT];
#pragma unroll
        for (int k = 0; k < WRPT; k++) {
            const int li = tid + k * WTH;
            ok[k] = worker && li < nrows;
            a0[k] = rh[k] = 0.0;
            rbv[k] = s0[k] = cc[k] = 0;
            tw[k] = 0;
            if (ok[k]) {
                a0[k] = d.tab[(size_t)(r_lo + li) * Cl + ql];
                rbv[k] = s_e2b[li];
                s0[k] = s_lp[li];
                rh[k] = s_rh[(size_t)cur * rpc + li];
                tw[k] = __ldcg(d.tabu + (size_t)q * d.W + (rbv[k] >> 5));
                cc[k] = __ldcg(d.col_cnt + rbv[k]);
            }
        }
        if (tid < t) { // P[s][q], s < t, from the CTA that owns window column q
            const int oc = ql / wpc;
            s_pq[tid] = cl.map_shared_rank(sP, oc)[(size_t)tid * wpc + (ql - oc * wpc)];
        }
        __syncthreads();
        double *Ft = Fptr(d, 0, par, t);
        unsigned long long key = KEY_NONE;
        int bi = INT_BIG, e_bv = 0, e_s0 = 0;
        double e_rh = 0.0, e_a = 0.0;
#pragma unroll
        for (int k = 0; k < WRPT; k++) {
            if (!ok[k]) continue;
            const int li = tid + k * WTH, i = r_lo + li;
            double a = a0[k];
            if (s0[k] >= 0) a = s_pq[s0[k]];
#pragma unroll 4
            for (int s = s0[k] + 1; s < t; s++) a = xp_add(a, xp_mul(sF[(size_t)s * rpc + li], s_pq[s]));
            const double f = -a;
            sF[(size_t)t * rpc + li] = f;
            __stcg(Ft + i, f);
            for (int r = 1; r < G; r++) Fptr(d, r, par, t)[i] = f; // peers: plain stores, fenced once per launch
            // neither pass takes a == 0 (tolerant), a handled pair (:589) or an exhausted leaving variable (:596)
            const bool el = !xp_feq(a, 0.0) && !((tw[k] >> (rbv[k] & 31)) & 1u) && cc[k] < n - 1;
            s_el[li] = el;
            if (el && a > 0.0) { // pass 1, :571-612
                const unsigned long long kk = f64_key(xp_div(rh[k], a));
                if (kk < key || (kk == key && i < bi)) key = kk, bi = i, e_rh = rh[k], e_a = a, e_bv = rbv[k], e_s0 = s0[k];
            }
        }
        WPANEL_T(0)
        int p = INT_BIG, bv = 0, s0p = 0;
        double piv_a = 0.0, piv_rh = 0.0;
        for (int pass = 1; pass <= 2; pass++) {
            if (pass == 2) { // pass 1 found no row anywhere: any a != 0 qualifies (:623-658)
                key = KEY_NONE, bi = INT_BIG;
#pragma unroll
                for (int k = 0; k < WRPT; k++) {
                    const int li = tid + k * WTH;
                    if (!ok[k] || !s_el[li]) continue;
                    const double a = -sF[(size_t)t * rpc + li];
                    const unsigned long long kk = f64_key(xp_div(rh[k], a));
                    if (kk < key || (kk == key && r_lo + li < bi))
                        key = kk, bi = r_lo + li, e_rh = rh[k], e_a = a, e_bv = rbv[k], e_s0 = s0[k];
                }
            }
            const int slot = (int)(xn++ & 1u);
            { // CTA arg-min, sent to every CTA of the cluster
                unsigned long long wk = key;
                int wi = bi;
                warp_min_key(wk, wi);
                if (worker && wi != INT_BIG && bi == wi) s_wrh[warp] = e_rh, s_wa[warp] = e_a, s_wbv[warp] = e_bv, s_ws0[warp] = e_s0;
                if (worker && lane == 0) s_wk[warp] = wk, s_wi[warp] = wi;
                __syncthreads();
                if (warp == 0) {
                    unsigned long long k2 = lane < WTH / 32 ? s_wk[lane] : KEY_NONE;
                    int i2 = lane < WTH / 32 ? s_wi[lane] : INT_BIG;
                    const int mine2 = i2;
                    warp_min_key(k2, i2);
                    const unsigned who = __ballot_sync(0xffffffffu, lane < WTH / 32 && mine2 == i2 && i2 != INT_BIG);
                    const int ww = who ? __ffs(who) - 1 : 0;
                    if (lane < WNC) {
                        WXchg *R = cl.map_shared_rank(&X, lane);
                        R->k[slot][c] = k2;
                        R->i[slot][c] = i2;
                        R->rh[slot][c] = s_wrh[ww];
                        R->a[slot][c] = s_wa[ww];
                        R->bv[slot][c] = s_wbv[ww];
                        R->s0[slot][c] = s_ws0[ww];
                    }
                }
            }
            cl.sync();
            { // every warp reduces the 16 partials (same inputs, same result everywhere)
                unsigned long long k2 = lane < WNC ? X.k[slot][lane] : KEY_NONE;
                int i2 = lane < WNC ? X.i[slot][lane] : INT_BIG;
                const int mine2 = i2;
                warp_min_key(k2, i2);
                const unsigned who = __ballot_sync(0xffffffffu, lane < WNC && mine2 == i2 && i2 != INT_BIG);
                const int cw = who ? __ffs(who) - 1 : 0;
                p = i2;
                piv_rh = X.rh[slot][cw], piv_a = X.a[slot][cw], bv = X.bv[slot][cw], s0p = X.s0[slot][cw];
            }
            if (p != INT_BIG) break;
        }
        WPANEL_T(1)
        if (p == INT_BIG) { // ratio test failed: k_pcol redoes this column and takes the slow path
            fail_out = 1;
            break;
        }
        const double r = xp_div(1.0, piv_a); // mulOfRow(eqnum, 1 / pivot), :1471
        const bool r_one = xp_feq(r, 1.0), r_zero = xp_feq(r, 0.0);
        const bool cq_zero = xp_feq(cq, 0.0), cq_one = xp_feq(cq, 1.0);
        const double prow_rhs = xp_scale(piv_rh, r, r_one, r_zero);
        // ============ phase B: row p over the window, objective row, constant column, pricing ============
        const int lj = tid;
        const bool okc = worker && lj < ncols;
        double b0 = 0.0;
        if (okc) b0 = d.tab[(size_t)p * Cl + c_lo + lj];
        if (tid < t) { // F[s][p], s < t, from the CTA that owns row p
            const int orow = p / rpc;
            s_fp[tid] = cl.map_shared_rank(sF, orow)[(size_t)tid * rpc + (p - orow * rpc)];
        }
        __syncthreads();
        int cand = INT_BIG, anyp = 0;
        double ccand = 0.0;
        if (okc) {
            const int g = c_lo + lj;
            double v = b0;
            if (s0p >= 0) v = sP[(size_t)s0p * wpc + lj];
#pragma unroll 4
            for (int s = s0p + 1; s < t; s++) v = xp_add(v, xp_mul(s_fp[s], sP[(size_t)s * wpc + lj]));
            const double xv = xp_scale(v, r, r_one, r_zero);
            sP[(size_t)t * wpc + lj] = xv;
            __stcg(d.P + (size_t)t * Cl + g, xv);
            const int nvraw = s_nv[lj];
            double tg = s_tg[lj];
            if (g < zero_upto && g < n && !nvraw) tg = 0.0;      // zeroing owed by the scan (:1059)
            double tt = xp_mul(xv, -1.0);                        // nvexp.mul(-1), :1496
            if (g >= n) tt = -tt;                                // constant column keeps its sign
            tt = cq_zero ? 0.0 : (cq_one ? tt : xp_mul(tt, cq)); // nvexp.mul(tgtf[nv])
            const double tn = xp_add(tt, tg);                    // tgtf.addRowToRow, :1501
            s_tg[lj] = tn;
            const int nvnew = g == bv ? 1 : (g == q ? 0 : nvraw); // basis after the swap, :1504-1510
            s_nv[lj] = nvnew;
            int rc = s_rc[lj];
            if (g == q) s_rc[lj] = ++rc; // genPair(q, bv): the pair is new (the ratio test skips handled pairs)
            if (nvnew && tn > 0.0) { // pricing of the next iteration, :1054-1069
                anyp = 1;
                if (rc < n - 1) cand = g, ccand = tn;
            }
        }
#pragma unroll
        for (int k = 0; k < WRPT; k++) { // constant column of my rows
            if (!ok[k]) continue;
            const int li = tid + k * WTH;
            s_rh[(size_t)(cur ^ 1) * rpc + li] =
                r_lo + li == p ? prow_rhs : xp_add(rh[k], xp_mul(sF[(size_t)t * rpc + li], prow_rhs));
        }
        WPANEL_T(2)
        const int slot = (int)(xn++ & 1u);
        {
            const int wc = __reduce_min_sync(0xffffffffu, cand);
            const int wany = __any_sync(0xffffffffu, anyp);
            if (worker && wc != INT_BIG && cand == wc) s_wcc[warp] = ccand;
            if (worker && lane == 0) s_wc[warp] = wc, s_wany[warp] = wany;
            __syncthreads();
            if (warp == 0) {
                const int c2 = lane < WTH / 32 ? s_wc[lane] : INT_BIG;
                const int a2 = lane < WTH / 32 ? s_wany[lane] : 0;
                const int mc = __reduce_min_sync(0xffffffffu, c2);
                const int ma = __any_sync(0xffffffffu, a2);
                const unsigned who = __ballot_sync(0xffffffffu, lane < WTH / 32 && c2 == mc && mc != INT_BIG);
                const int ww = who ? __ffs(who) - 1 : 0;
                if (lane < WNC) {
                    WXchg *R = cl.map_shared_rank(&X, lane);
                    R->c[slot][c] = (unsigned)mc | (ma ? 0x80000000u : 0u);
                    R->cc[slot][c] = s_wcc[ww];
                }
            }
        }
        cl.sync();
        int cd, ap;
        double cq_next;
        {
            const unsigned v = lane < WNC ? X.c[slot][lane] : 0x7fffffffu;
            const int c2 = (int)(v & 0x7fffffffu);
            cd = __reduce_min_sync(0xffffffffu, c2);
            ap = __any_sync(0xffffffffu, (v >> 31) & 1u);
            const unsigned who = __ballot_sync(0xffffffffu, lane < WNC && c2 == cd && cd != INT_BIG);
            cq_next = X.cc[slot][who ? __ffs(who) - 1 : 0];
        }
        WPANEL_T(3)
        // the thread that owns row p follows the swap (it is the only reader of these two entries)
        if (worker && p >= r_lo && p < r_lo + nrows && (p - r_lo) % WTH == tid) {
            s_e2b[p - r_lo] = q;
            s_lp[p - r_lo] = t;
        }
        if (c == 0 && warp == WTH / 32) { // ---- the book-keeping warp of CTA 0 ----
            if (lane < G) { // one record per pivot, to every rank (k_prow_bulk reads it)
                WRec rec;
                rec.r = r, rec.cq = cq, rec.prow_rhs = prow_rhs;
                rec.p = p, rec.q = q, rec.bv = bv, rec.s0p = s0p;
                recs[lane][t] = rec;
            }
            if (lane == 0) {
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
        }
        if (dbg && c == 0 && tid == 0) dbg[15] += 1;
        t++;
        cnt++;
        cur ^= 1;
        q = cd;
        if (cd != INT_BIG && cd > qmx) qmx = cd;
        anypos = ap;
        cq = cq_next;
        zero_upto = cd == INT_BIG ? 0 : cd;
        if (cd == INT_BIG) { // nothing eligible inside the window: the full-width slow path re-prices
            slow_out = 1;    // (its scan starts from column 0 and zeroes physically: nothing is owed)
            break;
        }
    }
    __syncthreads();
    // ---- write the window state back ----
    for (int li = tid; li < nrows; li += WTHB) d.rhsbuf[r_lo + li] = s_rh[(size_t)cur * rpc + li];
    for (int lj = tid; lj < ncols; lj += WTHB) d.tgtf[c_lo + lj] = s_tg[lj];
    const int bad = s_bad;
    if (c == 0 && warp == WTH / 32) {
        if (lane == 0) {
            st->t = t;
            st->cnt = cnt;
            st->q = q;
            st->anypos = anypos;
            st->zero_upto = zero_upto;
            st->slow = slow_out;
            st->wfail = fail_out;
            if (qmx > st->qmax) st->qmax = qmx;
            st->tg_rhs = tg_rhs;
            st->n_log = n_log;
            st->n_touched = n_touched;
            st->wseq = wseq;
            st->wcnt += (unsigned)(t - t_in);
            st->wb_t0 = t_in;
            st->wb_pending = t > t_in;
            if (bad) st->status = XP_ERR_PEER;
        }
        if (G > 1 && lane > 0 && lane < G) { // exit state for peer `lane`, then the flag
            XHdr *H = (XHdr *)d.xb[lane];
            H->wexit[0] = t, H->wexit[1] = q, H->wexit[2] = anypos, H->wexit[3] = zero_upto;
            H->wexit[4] = slow_out, H->wexit[5] = bad ? XP_ERR_PEER : XPI_RUNNING, H->wexit[6] = fail_out, H->wexit[7] = qmx;
        }
    }
    if (G > 1) {
        // every store of this launch into peer memory (multiplier columns by all CTAs, records and
        // exit words by CTA 0) must be visible before the flag: fence by each writer, cluster
        // barrier, then one releasing store per peer
        __threadfence_system();
        cl.sync();
        if (c == 0 && warp == WTH / 32 && lane > 0 && lane < G)
            st_release_sys(&((XHdr *)d.xb[lane])->wflag, (unsigned long long)wseq);
    }
#undef WPANEL_T
}

// Peers of a windowed launch: wait for the leader's flag, take over its exit state.
__global__ void k_wpanel_peer(LpDev d)
{
    LpState *st = d.st;
    if (threadIdx.x != 0) return;
    const int q = st->q, t = st->t;
    const bool go = st->status == XPI_RUNNING && !st->slow && !st->pivot_pending && q != INT_BIG &&
                    q < d.w && t < st->kblk && st->cnt < st->max_iter;
    st->wb_pending = 0;
    if (!go) return;
    const unsigned wseq = st->wseq + 1;
    XHdr *H = (XHdr *)d.xb[d.rank];
    const unsigned long long t0 = clock64();
    unsigned spins = 0;
    bool ok = true;
    while (ld_acquire_sys(&H->wflag) < wseq)
        if ((++spins & 1023u) == 0 && clock64() - t0 > SPIN_LIMIT) {
            ok = false;
            break;
        }
    st->wseq = wseq;
    if (!ok) {
        st->status = XP_ERR_PEER;
        return;
    }
    volatile int *we = H->wexit;
    const int t1 = we[0];
    st->q = we[1];
    st->anypos = we[2];
    st->zero_upto = we[3];
    st->slow = we[4];
    st->wfail = we[6];
    if (we[7] > st->qmax) st->qmax = we[7];
    if (we[5] != XPI_RUNNING) st->status = we[5];
    st->wb_t0 = t;
    st->wb_pending = t1 > t;
    st->wcnt += (unsigned)(t1 - t);
    st->t = t1; // cnt, n_log, tg_rhs, basis maps, tabu table: replayed from the records by k_prow_bulk
    __threadfence_system();
    st_release_sys(&((XHdr *)d.xb[0])->wack[d.rank], (unsigned long long)wseq);
}

// ---------------------------------------------------------------------------
// k_prow_bulk: what k_wpanel deferred.  For every local column outside the window and every
// pivot s of [wb_t0, t): the pivot row entry P[s][j] (row p_s as of step s: tableau entry +
// replay of the steps before s, scaled by 1/pivot, lpsol.h:1471) and the objective entry
// (:1496-1501) -- the same operations in the same order as phase B of the panel kernels, one
// column per thread, sequential in s.  Peers of a sharded LP (rank > 0) also replay the
// replicated bookkeeping from the records: constant column, basis maps, tabu table, pivot log.
// ---------------------------------------------------------------------------
constexpr int WB_TH = 128;

// mode 0 (live): pivots [wb_t0, t) of the open block, local columns [jlo, jhi) -- what the run
//   of k_wpanel just before left out;
// mode 1 (replay): every pivot of the closed block in ring slot `slot`, for columns that were
//   not there when the block was decided (a piecewise upload);
// mode 2: every pivot so far of the OPEN block (all made by k_wpanel), for such columns.
__global__ void __launch_bounds__(WB_TH) k_prow_bulk(LpDev d, int mode, int slot, int jlo, int jhi)
{
    extern __shared__ double s_bulk[]; // sA[KMAX][WB_TH] | sPr[KMAX][WB_TH]
    __shared__ double s_L[KMAX][KMAX]; // s_L[s][u] = F[u][p_s], u < s
    __shared__ WRec s_rec[KMAX];
    LpState *st = d.st;
    if (mode == 0 && !st->wb_pending) return;
    const int t0 = mode == 0 ? st->wb_t0 : 0, t1 = mode == 1 ? st->hist_t[slot] : st->t;
    const int par = mode == 1 ? slot : (st->blk & (NH - 1)), tid = threadIdx.x;
    const int n = d.n, Cl = jhi;
    if (t1 <= t0) return;
    const WRec *recs = (const WRec *)(d.xb[d.rank] + xoff_rec(d, par));
    for (int s = t0 + tid; s < t1; s += WB_TH) s_rec[s] = recs[s];
    __syncthreads();
    for (int e = tid; e < (t1 - t0) * KMAX; e += WB_TH) {
        const int s = t0 + e / KMAX, u = e % KMAX;
        s_L[s][u] = u < s ? ld_cg(Fptr(d, d.rank, par, u) + s_rec[s].p) : 0.0;
    }
    __syncthreads();
    double *sA = s_bulk, *sPr = s_bulk + (size_t)KMAX * WB_TH;
    const int ld = d.Cl;
    for (int jb = jlo + blockIdx.x * WB_TH; jb < Cl; jb += gridDim.x * WB_TH) {
        const int jl = jb + tid;
        if (jl < Cl) {
            const int g = d.col0 + jl;
            for (int s = t0; s < t1; s++) sA[s * WB_TH + tid] = d.tab[(size_t)s_rec[s].p * ld + jl];
            for (int u = 0; u < t0; u++) sPr[u * WB_TH + tid] = ld_cg(d.P + (size_t)u * ld + jl);
            double tg = d.tgtf[jl];
            for (int s = t0; s < t1; s++) {
                const WRec rc = s_rec[s];
                double v = sA[s * WB_TH + tid];
                if (rc.s0p >= 0) v = sPr[rc.s0p * WB_TH + tid];
#pragma unroll 8
                for (int u = rc.s0p + 1; u < s; u++) v = xp_add(v, xp_mul(s_L[s][u], sPr[u * WB_TH + tid]));
                const double xv = xp_scale(v, rc.r, xp_feq(rc.r, 1.0), xp_feq(rc.r, 0.0));
                sPr[s * WB_TH + tid] = xv;
                d.P[(size_t)s * ld + jl] = xv;
                double tt = xp_mul(xv, -1.0);
                if (g >= n) tt = -tt;
                tt = xp_feq(rc.cq, 0.0) ? 0.0 : (xp_feq(rc.cq, 1.0) ? tt : xp_mul(tt, rc.cq));
                tg = xp_add(tt, tg); // no zeroing out here: the scan stopped inside the window (zero_upto <= q < w)
            }
            d.tgtf[jl] = tg;
        }
    }
    if (d.rank > 0 && mode == 0) { // replicated state the leader kept while deciding
        for (int i = blockIdx.x * WB_TH + tid; i < d.m; i += gridDim.x * WB_TH) {
            double rh = d.rhsbuf[i];
            for (int s = t0; s < t1; s++) {
                const double pr = s_rec[s].prow_rhs;
                rh = i == s_rec[s].p ? pr : xp_add(rh, xp_mul(ld_cg(Fptr(d, d.rank, par, s) + i), pr));
            }
            d.rhsbuf[i] = rh;
        }
        if (blockIdx.x == 0 && tid == 0) {
            unsigned n_log = st->n_log;
            int n_touched = st->n_touched;
            double tg_rhs = st->tg_rhs;
            for (int s = t0; s < t1; s++) {
                const WRec rc = s_rec[s];
                uint32_t *w = &d.tabu[(size_t)rc.q * d.W + (rc.bv >> 5)];
                const uint32_t bit = 1u << (rc.bv & 31);
                if (!(*w & bit)) {
                    *w |= bit;
                    d.row_cnt[rc.q] += 1;
                    d.col_cnt[rc.bv] += 1;
                }
                if (n_log < d.log_cap) {
                    d.log[3 * n_log] = rc.q;
                    d.log[3 * n_log + 1] = rc.bv;
                    d.log[3 * n_log + 2] = rc.p;
                }
                n_log++;
                d.nvset[rc.q] = 0;
                d.nvset[rc.bv] = 1;
                d.eq2bv[rc.p] = rc.q;
                d.bv2eq[rc.q] = rc.p;
                d.bv2eq[rc.bv] = -1;
                d.last_piv[rc.p] = s;
                if (rc.s0p < 0) st->touched[n_touched++] = rc.p;
                double tt = -xp_mul(rc.prow_rhs, -1.0);
                tt = xp_feq(rc.cq, 0.0) ? 0.0 : (xp_feq(rc.cq, 1.0) ? tt : xp_mul(tt, rc.cq));
                tg_rhs = xp_add(tt, tg_rhs);
            }
            st->n_log = n_log;
            st->n_touched = n_touched;
            st->tg_rhs = tg_rhs;
            st->cnt += (unsigned)(t1 - t0);
        }
    }
}

