// dd_lsap.cuh -- scipy-exact rectangular linear sum assignment + CPython-exact set-difference order.
//
// dd_lsap_solve restates scipy 1.18.1 optimize/rectangular_lsap/rectangular_lsap.cpp (the solver
// behind deep_sort/linear_assignment.py:58) INCLUDING the rules that decide ties, because the
// order of `unmatched_detections` (hence new track ids) depends on them:
//   * rows are inserted one by one, serially;
//   * the not-yet-scanned column list starts in REVERSE order and a chosen column is removed by
//     swap-with-last;
//   * scanning the list in order, the running minimum is replaced when a column's reduced path cost
//     is strictly lower, or equal and the column is unassigned.  Equivalent order-free form used by
//     the parallel scan: lowest value wins; among equal values an unassigned column beats an
//     assigned one, the LAST unassigned one wins, otherwise the FIRST assigned one.
// Only the column scan, the dual update and the list compactions are spread over the lanes of the
// group; all arithmetic is the same f64 add/sub sequence as scipy (no multiplications -> no FMA).
#pragma once
#include "dd_common.cuh"

struct DDLsapScratch {
    double* u;                 // [n]   row duals
    double* v;                 // [n]   column duals
    double* spc;               // [n]   shortest path costs
    short* path;               // [n]
    short* col4row;            // [n]
    short* row4col;            // [n]
    short* remaining;          // [n]
    unsigned char* SR;         // [n]
    unsigned char* SC;         // [n]
};

DD_HD size_t dd_lsap_scratch_bytes(int n) {
    size_t b = (size_t)n * (3 * 8 + 4 * 2 + 2);
    return (b + 15) & ~(size_t)15;
}

DD_HD void dd_lsap_carve(char* mem, int n, DDLsapScratch& s) {
    s.u = (double*)mem;
    s.v = s.u + n;
    s.spc = s.v + n;
    s.path = (short*)(s.spc + n);
    s.col4row = s.path + n;
    s.row4col = s.col4row + n;
    s.remaining = s.row4col + n;
    s.SR = (unsigned char*)(s.remaining + n);
    s.SC = s.SR + n;
}

#if defined(__CUDA_ARCH__)
#define DD_INF __longlong_as_double(0x7ff0000000000000LL)
#else
#define DD_INF ((double)INFINITY)
#endif

// Solve min-cost assignment for an nr x nc problem with nr <= nc (the caller transposes otherwise,
// exactly like scipy).  cost(i, j) returns the f64 cost.  On return col4row[i] (all rows assigned)
// and row4col[j] (-1 = free column).  Returns 0, or -1 when the matrix is infeasible.
template <class G, class CostFn>
DD_HD int dd_lsap_solve(const G& g, int nr, int nc, const CostFn& cost, DDLsapScratch& s) {
    for (int i = g.lane; i < nr; i += G::NL) { s.u[i] = 0.0; s.col4row[i] = -1; }
    for (int j = g.lane; j < nc; j += G::NL) { s.v[j] = 0.0; s.row4col[j] = -1; s.path[j] = -1; }
    g.sync();
    for (int cur = 0; cur < nr; ++cur) {
        // ---- fast path: the first scan of scipy's search, without materialising its work arrays.
        // With spc = +inf everywhere and minVal = 0 the scan gives spc[j] = (0 + c(cur,j)) - u[cur] - v[j]
        // for every column in list order it = nc-1-j.  If the winner is an unassigned column it is the sink,
        // the path is (cur -> sink), the dual update is u[cur] += minVal and v[sink] -= (minVal - minVal),
        // and nothing else of the search state survives the row.  Otherwise fall through to the full search.
        {
            const double ui = s.u[cur];
            DDKey best;
            best.val = DD_INF;
            best.pref = -0x7fffffff;
            bool have = false;
            for (int it = g.lane; it < nc; it += G::NL) {
                const int j = nc - 1 - it;
                DDKey k;
                k.val = dd_sub(dd_sub(dd_add(0.0, cost(cur, j)), ui), s.v[j]);
                k.pref = (s.row4col[j] == -1) ? (it + 1) : -it;
                if (!(k.val < DD_INF)) { k.val = DD_INF; }          // r < inf fails in scipy: spc stays inf
                if (!have || dd_key_better(k, best)) { best = k; have = true; }
            }
            best = g.best(best);
            if (!(best.val < DD_INF)) return -1;
            if (best.pref > 0) {
                const int sink = nc - 1 - (best.pref - 1);
                g.sync();
                if (g.lane == 0) {
                    s.u[cur] = dd_add(ui, best.val);
                    s.v[sink] = dd_sub(s.v[sink], dd_sub(best.val, best.val));
                    s.row4col[sink] = (short)cur;
                    s.col4row[cur] = (short)sink;
                }
                g.sync();
                continue;
            }
        }
        for (int j = g.lane; j < nc; j += G::NL) {
            s.remaining[j] = (short)(nc - 1 - j);
            s.SC[j] = 0;
            s.spc[j] = DD_INF;
        }
        for (int i = g.lane; i < nr; i += G::NL) s.SR[i] = 0;
        g.sync();
        int nrem = nc;
        double minVal = 0.0;
        int i = cur;
        int sink = -1;
        while (sink == -1) {
            if (g.lane == 0) s.SR[i] = 1;
            const double ui = s.u[i];
            DDKey best;
            best.val = DD_INF;
            best.pref = -0x7fffffff;
            int best_it = -1;
            for (int it = g.lane; it < nrem; it += G::NL) {
                const int j = s.remaining[it];
                const double r = dd_sub(dd_sub(dd_add(minVal, cost(i, j)), ui), s.v[j]);
                double sp = s.spc[j];
                if (r < sp) {
                    s.path[j] = (short)i;
                    s.spc[j] = r;
                    sp = r;
                }
                DDKey k;
                k.val = sp;
                k.pref = (s.row4col[j] == -1) ? (it + 1) : -it;
                if (best_it < 0 || dd_key_better(k, best)) { best = k; best_it = it; }
            }
            if (best_it < 0) { best.val = DD_INF; best.pref = -0x7fffffff; }
            best = g.best(best);
            minVal = best.val;
            if (!(minVal < DD_INF)) return -1;                 // infeasible (inf or NaN)
            const int index = best.pref > 0 ? best.pref - 1 : -best.pref;
            const int j = s.remaining[index];
            const int r4c = s.row4col[j];
            g.sync();                                          // all lanes have read remaining[]
            if (r4c == -1) sink = j; else i = r4c;
            if (g.lane == 0) {
                s.SC[j] = 1;
                s.remaining[index] = s.remaining[nrem - 1];
            }
            --nrem;
            g.sync();
        }
        // dual update (rectangular_lsap.cpp: u[cur] += minVal; visited rows / columns)
        for (int i2 = g.lane; i2 < nr; i2 += G::NL) {
            if (i2 == cur) s.u[i2] = dd_add(s.u[i2], minVal);
            else if (s.SR[i2]) s.u[i2] = dd_add(s.u[i2], dd_sub(minVal, s.spc[s.col4row[i2]]));
        }
        for (int j2 = g.lane; j2 < nc; j2 += G::NL)
            if (s.SC[j2]) s.v[j2] = dd_sub(s.v[j2], dd_sub(minVal, s.spc[j2]));
        g.sync();
        if (g.lane == 0) {                                     // augment along the path
            int j = sink;
            while (true) {
                const int i2 = s.path[j];
                s.row4col[j] = (short)i2;
                const int t = s.col4row[i2];
                s.col4row[i2] = (short)j;
                j = t;
                if (i2 == cur) break;
            }
        }
        g.sync();
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// CPython 3.12 Objects/setobject.c emulation for small non-negative ints (hash(k) == k):
// iteration order of list(set(a) - set(m)) -- deep_sort/linear_assignment.py:140.
// Serial (one lane).  Tables hold key+1 (0 = empty slot); no dummies ever appear in this flow.
// ------------------------------------------------------------------------------------------------
DD_HD int dd_set_table_slots(int n_keys) {      // table size CPython ends with after n inserts
    int mask = 7, fill = 0;
    for (int k = 0; k < n_keys; ++k) {
        ++fill;
        if (fill * 5 >= mask * 3) {
            int ns = 8;
            while (ns <= fill * 4) ns <<= 1;
            mask = ns - 1;
        }
    }
    // set_merge into an empty set may pre-size to the smallest power of two > 2 * n_keys
    int ns2 = 8;
    while (ns2 <= n_keys * 2) ns2 <<= 1;
    return (mask + 1) > ns2 ? (mask + 1) : ns2;
}

DD_HD void dd_set_insert_clean(short* tbl, int mask, int key) {   // set_insert_clean / set_add_entry
    unsigned perturb = (unsigned)key;
    int i = key & mask;
    while (true) {
        if (tbl[i] == 0) { tbl[i] = (short)(key + 1); return; }
        if (i + 9 <= mask) {
            for (int j = 1; j <= 9; ++j)
                if (tbl[i + j] == 0) { tbl[i + j] = (short)(key + 1); return; }
        }
        perturb >>= 5;
        i = (int)(((unsigned)i * 5u + 1u + perturb) & (unsigned)mask);
    }
}

// Rebuild `src` (mask ms) into `dst` with `newsize` slots, re-inserting in slot order
// (set_table_resize).  Returns the new mask.
DD_HD int dd_set_rebuild(const short* src, int ms, short* dst, int newsize) {
    for (int i = 0; i < newsize; ++i) dst[i] = 0;
    for (int i = 0; i <= ms; ++i)
        if (src[i]) dd_set_insert_clean(dst, newsize - 1, src[i] - 1);
    return newsize - 1;
}

// Insert with CPython's growth rule (set_add_entry: fill*5 >= mask*3 -> resize to > used*4).
// tbl / tmp are two buffers of `cap` slots that ping-pong on resize; returns via references.
DD_HD void dd_set_add(short*& tbl, short*& tmp, int& mask, int& fill, int key) {
    dd_set_insert_clean(tbl, mask, key);
    ++fill;
    if (fill * 5 >= mask * 3) {
        int ns = 8;
        while (ns <= fill * 4) ns <<= 1;
        mask = dd_set_rebuild(tbl, mask, tmp, ns);
        short* t = tbl; tbl = tmp; tmp = t;
    }
}

// a[0..na): distinct keys in list order (the `track_indices` list); is_matched[k] != 0 when key k is
// in the subtracted set; nm = size of the subtracted set.  Writes the iteration order of the
// difference to out[], returns its length.  bufA/bufB/bufC: three tables of `cap` slots each.
DD_HD int dd_set_difference_order_serial(const short* a, int na, const unsigned char* is_matched,
                                         int nm, short* out, short* bufA, short* bufB, short* bufC,
                                         int cap) {
    (void)cap;
    // 1. so = set(a)   (set_update_internal over a list -> set_add_key per element)
    short *A = bufA, *At = bufB;
    int maskA = 7, fillA = 0;
    for (int i = 0; i < 8; ++i) A[i] = 0;
    for (int k = 0; k < na; ++k) dd_set_add(A, At, maskA, fillA, a[k]);
    short* R = bufC;
    short* Rt = (A == bufA) ? bufB : bufA;
    int n_out = 0;
    if ((na >> 2) > nm) {
        // 2a. set_copy_and_difference: copy `so` (set_merge into an empty set), then discard.
        int maskR = 7;
        if (na * 5 >= maskR * 3) {
            int ns = 8;
            while (ns <= na * 2) ns <<= 1;
            maskR = ns - 1;
        }
        for (int i = 0; i <= maskR; ++i) R[i] = 0;
        if (maskR == maskA) {
            for (int i = 0; i <= maskA; ++i) R[i] = A[i];
        } else {
            for (int i = 0; i <= maskA; ++i)
                if (A[i]) dd_set_insert_clean(R, maskR, A[i] - 1);
        }
        for (int i = 0; i <= maskR; ++i)
            if (R[i] && !is_matched[R[i] - 1]) out[n_out++] = (short)(R[i] - 1);
        return n_out;
    }
    // 2b. fresh result set; walk `so` in slot order, add survivors.
    int maskR = 7, fillR = 0;
    for (int i = 0; i < 8; ++i) R[i] = 0;
    for (int i = 0; i <= maskA; ++i) {
        if (!A[i]) continue;
        const int key = A[i] - 1;
        if (is_matched[key]) continue;
        dd_set_add(R, Rt, maskR, fillR, key);
    }
    for (int i = 0; i <= maskR; ++i)
        if (R[i]) out[n_out++] = (short)(R[i] - 1);
    return n_out;
}

// Fast path for the case the tracker always produces: a = [0, 1, ..., n-1].
// set(range(n)) places key k in slot k at every table size (k < fill <= 3/5 mask), so its slot order is
// ascending and only the survivors matter:
//   (n >> 2) > nm  : set_copy_and_difference -> ascending survivors (caller already has them);
//   otherwise      : the survivors, in ascending order, are inserted into a fresh set with CPython's
//                    growth rule and read back in slot order (dd_set_order_from_survivors).
// Only the insertions are inherently serial; dd_set_build_from_survivors does them (one lane) and returns
// (table << 16) | mask -- table 0: the set lives in bufA, 1: in bufB -- so that a group can read the slots back with an
// ordered compaction (dd_match_stream).
DD_HD int dd_set_build_from_survivors(const short* surv_asc, int k, short* bufA, short* bufB) {
    short *R = bufA, *Rt = bufB;
    int maskR = 7, fillR = 0;
    for (int i = 0; i < 8; ++i) R[i] = 0;
    for (int i = 0; i < k; ++i) dd_set_add(R, Rt, maskR, fillR, surv_asc[i]);
    return ((R == bufA ? 0 : 1) << 16) | maskR;
}
DD_HD int dd_set_order_from_survivors(const short* surv_asc, int k, short* out, short* bufA, short* bufB) {
    const int tm = dd_set_build_from_survivors(surv_asc, k, bufA, bufB);
    const short* R = (tm >> 16) ? bufB : bufA;
    const int maskR = tm & 0xffff;
    int n_out = 0;
    for (int i = 0; i <= maskR; ++i)
        if (R[i]) out[n_out++] = (short)(R[i] - 1);
    return n_out;
}
