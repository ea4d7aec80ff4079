// cpk_host_sweep.hpp -- item list of the LDL' sweeps for the global-memory walks
// Host-side part of libcpk_b200 (included by cpk_host.cu only; uses its fail() and the
// constants of cpk_device.cuh).
#pragma once

// ---------------------------------------------------------------------------
// LDL' sweep builder.  Produces the unified item list of DevSweep: forward items
// (level order), then backward items (level order), with the rows that need no
// work compiled away (see cpk_device.cuh) and the segment table that tells the
// kernel how to deal items to warps (blocks of consecutive items for bulk
// levels, one item per warp for chains).
// ---------------------------------------------------------------------------
struct HSweep {
    int nitems = 0, nfwd = 0;
    std::vector<int> seg, levptr, sptr, col, rid, pidx, flags, partner;
    std::vector<double> val, d, e, dp;
    // "lone" rows: no work in the forward sweep, no entry in the backward one, gathered by nobody:
    // y_i = z_i / d_i.  Kept out of the item list: a plain streamed pass (DevSweep::lone_*)
    std::vector<int> lone_pidx;
    std::vector<double> lone_d;
    // level walk: first item of every CTA's share of every level ([levels][nsplit+1], see DevSweep)
    int nsplit = 0;
    std::vector<int> ctasplit;
    int64_t n_trivial = 0, n_fused = 0, tail_f = 0, tail_b = 0, n_warprow = 0, max_len = 0;
    int lev_f_eff = 0, lev_b_eff = 0;
};

struct SweepRow { int row; int level; int len; };
constexpr int kLongRow = 8;         // rows with more entries are walked by a whole warp
typedef std::vector<std::pair<int, double>> EncRow;     // (column code, value), in accumulation order

// Appends the rows (already filtered) of one direction.  `entries(r)` returns the
// encoded dependency list of LDL row r.
// The items of a level are then DEALT to the CTAs that will walk it: every CTA gets one contiguous
// share, balanced by the estimated cost of the items (longest-processing-time-first), and inside a share
// the items are ordered by kind, so that the batches the warps draw from it are of one kind.  Cost in
// dependent round trips: a lane item of <= 2 entries per row or a warp-row of <= 64 entries is walked
// in batches of two (3 trips per batch), anything else item by item (3 trips + 2 per chunk of
// 4 entries per lane).
static int item_kind(const HSweep &W, int it)
{
    const int width = (W.sptr[it + 1] - W.sptr[it]) / 32;
    const int f0 = W.flags[(size_t)it * 32];
    bool partner = false;
    for (int l = 0; l < 32; ++l) partner = partner || (W.flags[(size_t)it * 32 + l] & F_PARTNER);
    if (partner) return 2;
    if (f0 & F_WARPROW) return width <= 2 ? 1 : 2;
    return width <= 2 ? 0 : 2;
}
static int item_cost(const HSweep &W, int it)
{
    const int width = (W.sptr[it + 1] - W.sptr[it]) / 32;
    return item_kind(W, it) == 2 ? 6 + 4 * ((std::max(width, 1) + 3) / 4) : 3;     // in half round trips
}
static void deal_level(HSweep &W, int first, int n, int nctas)
{
    std::vector<int> order;             // new position -> old item (relative)
    order.reserve(n);
    std::vector<int> split((size_t)nctas + 1, first);
    if (n > 0) {
        std::vector<int> kind(n), cost(n), idx(n);
        for (int i = 0; i < n; ++i) { kind[i] = item_kind(W, first + i); cost[i] = item_cost(W, first + i); idx[i] = i; }
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return cost[x] > cost[y]; });
        std::vector<std::vector<int>> share(nctas);
        typedef std::pair<long long, int> Load;     // (cost so far, CTA)
        std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
        for (int c = 0; c < nctas; ++c) heap.push(Load(0, c));
        for (int i : idx) {
            Load l = heap.top(); heap.pop();
            share[l.second].push_back(i);
            heap.push(Load(l.first + cost[i], l.second));
        }
        for (int c = 0; c < nctas; ++c) {
            // generic items first (the long poles), then warp-rows, then lane items; original order inside a kind
            // (rows of equal length are ordered by their index in the user vector: coalesced gathers)
            std::stable_sort(share[c].begin(), share[c].end(), [&](int x, int y) {
                const int kx = kind[x] == 2 ? 0 : kind[x] == 1 ? 1 : 2, ky = kind[y] == 2 ? 0 : kind[y] == 1 ? 1 : 2;
                if (kx != ky) return kx < ky;
                return x < y;
            });
            split[c] = first + (int)order.size();
            for (int i : share[c]) order.push_back(i);
        }
        split[nctas] = first + n;
    }
    W.ctasplit.insert(W.ctasplit.end(), split.begin(), split.end());
    if (n <= 1) return;
    // rebuild the item arrays of this level in the new order
    const int s0 = W.sptr[first];
    std::vector<int> col, sptr(1, s0), rid, pidx, flags, partner;
    std::vector<double> val, d, e, dp;
    {
        const size_t ne = (size_t)(W.sptr[first + n] - s0), ns = (size_t)n * 32;
        col.reserve(ne); val.reserve(ne); sptr.reserve((size_t)n + 1);
        rid.reserve(ns); pidx.reserve(ns); flags.reserve(ns); partner.reserve(ns); d.reserve(ns); e.reserve(ns); dp.reserve(ns);
    }
    for (int k = 0; k < n; ++k) {
        const int it = first + order[k];
        const int b = W.sptr[it], en = W.sptr[it + 1];
        col.insert(col.end(), W.col.begin() + b, W.col.begin() + en);
        val.insert(val.end(), W.val.begin() + b, W.val.begin() + en);
        sptr.push_back(s0 + (int)col.size());
        for (int l = 0; l < 32; ++l) {
            const size_t sl = (size_t)it * 32 + l;
            rid.push_back(W.rid[sl]); pidx.push_back(W.pidx[sl]); flags.push_back(W.flags[sl]); partner.push_back(W.partner[sl]);
            d.push_back(W.d[sl]); e.push_back(W.e[sl]); dp.push_back(W.dp[sl]);
        }
    }
    std::copy(col.begin(), col.end(), W.col.begin() + s0);
    std::copy(val.begin(), val.end(), W.val.begin() + s0);
    for (int k = 0; k <= n; ++k) W.sptr[first + k] = sptr[k];
    const size_t base = (size_t)first * 32;
    std::copy(rid.begin(), rid.end(), W.rid.begin() + base);
    std::copy(pidx.begin(), pidx.end(), W.pidx.begin() + base);
    std::copy(flags.begin(), flags.end(), W.flags.begin() + base);
    std::copy(partner.begin(), partner.end(), W.partner.begin() + base);
    std::copy(d.begin(), d.end(), W.d.begin() + base);
    std::copy(e.begin(), e.end(), W.e.begin() + base);
    std::copy(dp.begin(), dp.end(), W.dp.begin() + base);
}

// Small systems (one CTA, or a grid that holds a warp per item): every warp owns a contiguous range of equal
// COUNT; the items, sorted by cost, are dealt round-robin to the warps, so every range holds the same mix.
static void deal_level_rr(HSweep &W, int first, int n, int nw)
{
    if (n <= 1 || nw <= 1) return;
    std::vector<int> order;             // new position -> old item (relative)
    order.reserve(n);
    // warp w gets [n*w/nw, n*(w+1)/nw): fill these ranges by dealing
    std::vector<std::vector<int>> hand(std::min(nw, n));
    const int nh = (int)hand.size();
    // ranges have sizes that differ by at most one; deal so that range sizes match
    std::vector<int> cap(nh);
    for (int w = 0; w < nh; ++w) cap[w] = (int)((long long)n * (w + 1) / nh) - (int)((long long)n * w / nh);
    int w = 0;
    for (int i = 0; i < n; ++i) {
        int tries = 0;
        while ((int)hand[w].size() >= cap[w] && tries < nh) { w = (w + 1) % nh; ++tries; }
        hand[w].push_back(i);
        w = (w + 1) % nh;
    }
    for (auto &h : hand) for (int i : h) order.push_back(i);
    // rebuild the item arrays of this level in the new order
    const int s0 = W.sptr[first];
    std::vector<int> col, sptr(1, s0), rid, pidx, flags, partner;
    std::vector<double> val, d, e, dp;
    {
        const size_t ne = (size_t)(W.sptr[first + n] - s0), ns = (size_t)n * 32;
        col.reserve(ne); val.reserve(ne); sptr.reserve((size_t)n + 1);
        rid.reserve(ns); pidx.reserve(ns); flags.reserve(ns); partner.reserve(ns); d.reserve(ns); e.reserve(ns); dp.reserve(ns);
    }
    for (int k = 0; k < n; ++k) {
        const int it = first + order[k];
        const int b = W.sptr[it], en = W.sptr[it + 1];
        col.insert(col.end(), W.col.begin() + b, W.col.begin() + en);
        val.insert(val.end(), W.val.begin() + b, W.val.begin() + en);
        sptr.push_back(s0 + (int)col.size());
        for (int l = 0; l < 32; ++l) {
            const size_t sl = (size_t)it * 32 + l;
            rid.push_back(W.rid[sl]); pidx.push_back(W.pidx[sl]); flags.push_back(W.flags[sl]); partner.push_back(W.partner[sl]);
            d.push_back(W.d[sl]); e.push_back(W.e[sl]); dp.push_back(W.dp[sl]);
        }
    }
    std::copy(col.begin(), col.end(), W.col.begin() + s0);
    std::copy(val.begin(), val.end(), W.val.begin() + s0);
    for (int k = 0; k <= n; ++k) W.sptr[first + k] = sptr[k];
    const size_t base = (size_t)first * 32;
    std::copy(rid.begin(), rid.end(), W.rid.begin() + base);
    std::copy(pidx.begin(), pidx.end(), W.pidx.begin() + base);
    std::copy(flags.begin(), flags.end(), W.flags.begin() + base);
    std::copy(partner.begin(), partner.end(), W.partner.begin() + base);
    std::copy(d.begin(), d.end(), W.d.begin() + base);
    std::copy(e.begin(), e.end(), W.e.begin() + base);
    std::copy(dp.begin(), dp.end(), W.dp.begin() + base);
}

template <class Entries, class Flags>
static void sweep_append(HSweep &W, std::vector<SweepRow> rows, const std::vector<int64_t> &perm,
                         const std::vector<double> &dd, const std::vector<double> &ee, const std::vector<int> &partner,
                         int grid_warps, Entries entries, Flags flags_of)
{
    // the team shape that will walk this system: whole grid for large N, one CTA otherwise
    // shares per CTA + work queue: large systems with SHALLOW sweeps (cfg 3 / cfg 4: 1+1 levels of tens of
    // thousands of items; cfg 3 solve 3.31 -> 3.16 ms, batches of N = 80 000 systems on sub-teams 3.84 -> 3.40 ms).
    // Sweeps of many small levels keep the round-robin deal: measured 8-10 % slower with shares (stress g=40,
    // 16+2 levels), and the sync-free walk of deep sweeps hands consecutive items to a warp anyway.
    const bool large = (int)perm.size() > 24576;
    const bool big = large && W.lev_f_eff + W.lev_b_eff <= 4;
    const int deal_ctas = big ? std::max(grid_warps / kWarpsPerCta, 1) : 0;      // 0: equal counts per warp (deal_level_rr)
    W.nsplit = deal_ctas;
    // inside a level any order is valid: group rows of equal (short) length and
    // order them by their index in the user vector, so that the P' gather and the
    // P scatter of consecutive lanes touch consecutive addresses
    for (auto &rw : rows) W.max_len = std::max<int64_t>(W.max_len, rw.len);
    // Rows longer than the level's threshold become warp-rows (one warp per row: the shortest chain
    // for a level of few rows, where latency is all that counts).  A WIDE level of a large system
    // (several rows per warp of the grid) is bound by item throughput instead -- a warp-row of 10..40
    // entries keeps a whole warp busy for the same dependent round trips as a lane item of 32 rows
    // -- so a level with MANY such rows (several per warp of the grid: merged levels of a filled
    // factor hold hundreds of thousands) keeps rows of up to 32 entries as lane items.
    static const int wide_thr = [] { const char *e = getenv("CPK_LDL_WIDE_ROW"); return e ? atoi(e) : 32; }();
    std::vector<long long> lvl_mid;        // rows of kLongRow+1 .. wide_thr entries per level
    for (auto &rw : rows) {
        if ((size_t)rw.level >= lvl_mid.size()) lvl_mid.resize((size_t)rw.level + 1, 0);
        if (rw.len > kLongRow && rw.len <= wide_thr) lvl_mid[rw.level]++;
    }
    auto long_thr = [&](int level) {
        if ((int)perm.size() <= 24576) return kLongRow;
        return lvl_mid[level] >= 4LL * grid_warps ? std::max(wide_thr, kLongRow) : kLongRow;
    };
    std::stable_sort(rows.begin(), rows.end(), [&](const SweepRow &a, const SweepRow &b) {
        if (a.level != b.level) return a.level < b.level;
        const int thr = long_thr(a.level);
        const int la = std::min(a.len, thr + 1), lb = std::min(b.len, thr + 1);
        if (la != lb) return la > lb;
        return perm[a.row] < perm[b.row];
    });
    if (W.sptr.empty()) W.sptr.push_back(0);
    // items per level
    std::vector<std::pair<int, int>> level_items;       // (first item, count) per level, in order
    size_t i0 = 0;
    while (i0 < rows.size()) {
        const int lev = rows[i0].level;
        const int first_item = W.nitems;
        const int thr = long_thr(lev);
        while (i0 < rows.size() && rows[i0].level == lev) {
            if (rows[i0].len > thr) {
                // one long row = one item, entries spread over the 32 lanes
                const int r = rows[i0].row;
                const EncRow &er = entries(r);
                const size_t width = (er.size() + 31) / 32;
                const size_t base = W.col.size();
                W.col.resize(base + width * 32, -1);
                W.val.resize(base + width * 32, 0.0);
                for (size_t jx = 0; jx < er.size(); ++jx) { W.col[base + jx] = er[jx].first; W.val[base + jx] = er[jx].second; }
                for (int lane = 0; lane < 32; ++lane) {
                    if (lane == 0) {
                        W.rid.push_back(r); W.pidx.push_back((int)perm[r]); W.flags.push_back(flags_of(r) | F_WARPROW); W.d.push_back(dd[r]);
                        if (partner[r] >= 0) { W.partner.push_back(partner[r]); W.e.push_back(ee[std::min(r, partner[r])]); W.dp.push_back(dd[partner[r]]); }
                        else { W.partner.push_back(-1); W.e.push_back(0.0); W.dp.push_back(1.0); }
                    } else {
                        W.rid.push_back(-1); W.pidx.push_back(0); W.flags.push_back(0); W.d.push_back(1.0);
                        W.partner.push_back(-1); W.e.push_back(0.0); W.dp.push_back(1.0);
                    }
                }
                W.sptr.push_back((int)W.col.size());
                W.nitems++; W.n_warprow++;
                ++i0;
                continue;
            }
            size_t i1 = i0;
            while (i1 < rows.size() && i1 - i0 < 32 && rows[i1].level == lev && rows[i1].len <= thr) ++i1;
            int width = 0;
            for (size_t t = i0; t < i1; ++t) width = std::max(width, rows[t].len);
            const size_t base = W.col.size();
            W.col.resize(base + (size_t)width * 32, -1);
            W.val.resize(base + (size_t)width * 32, 0.0);
            for (int lane = 0; lane < 32; ++lane) {
                const size_t t = i0 + lane;
                if (t < i1) {
                    const int r = rows[t].row;
                    W.rid.push_back(r);
                    W.pidx.push_back((int)perm[r]);
                    W.flags.push_back(flags_of(r));
                    W.d.push_back(dd[r]);
                    if (partner[r] >= 0) {
                        W.partner.push_back(partner[r]);
                        W.e.push_back(ee[std::min(r, partner[r])]);
                        W.dp.push_back(dd[partner[r]]);
                    } else { W.partner.push_back(-1); W.e.push_back(0.0); W.dp.push_back(1.0); }
                    const EncRow &er = entries(r);
                    for (size_t jx = 0; jx < er.size(); ++jx) {
                        W.col[base + jx * 32 + lane] = er[jx].first;
                        W.val[base + jx * 32 + lane] = er[jx].second;
                    }
                } else {
                    W.rid.push_back(-1); W.pidx.push_back(0); W.flags.push_back(0); W.d.push_back(1.0);
                    W.partner.push_back(-1); W.e.push_back(0.0); W.dp.push_back(1.0);
                }
            }
            W.sptr.push_back((int)W.col.size());
            W.nitems++;
            i0 = i1;
        }
        level_items.emplace_back(first_item, W.nitems - first_item);
        W.levptr.push_back(first_item);
        if (big) deal_level(W, first_item, W.nitems - first_item, deal_ctas);
        else deal_level_rr(W, first_item, W.nitems - first_item, large ? grid_warps : kWarpsPerCta);
    }
    // segments: a level with at least 2 items per warp is "bulk" (blocks of
    // consecutive items per warp); runs of smaller levels
    // are merged into one chain segment dealt round-robin.
    int chain_first = -1, chain_end = -1;
    auto flush_chain = [&]() {
        if (chain_first >= 0) { W.seg.push_back(chain_first); W.seg.push_back(chain_end); W.seg.push_back(1); }
        chain_first = -1;
    };
    static const double bulk_min = [] { const char *e = getenv("CPK_LDL_BULK_MIN"); return e ? atof(e) : 2.0; }();
    static const int blk_max = [] { const char *e = getenv("CPK_LDL_BLK"); return e ? atoi(e) : 8; }();
    for (auto &li : level_items) {
        if (li.second >= bulk_min * grid_warps) {
            flush_chain();
            const int blk = std::max(1, std::min(blk_max, li.second / grid_warps));
            W.seg.push_back(li.first); W.seg.push_back(li.first + li.second); W.seg.push_back(blk);
        } else {
            if (chain_first < 0) chain_first = li.first;
            chain_end = li.first + li.second;
        }
    }
    flush_chain();
}

