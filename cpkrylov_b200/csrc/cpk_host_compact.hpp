// cpk_host_compact.hpp -- stream builder of the compact (one-CTA) LDL' walk
// Host-side part of libcpk_b200 (included by cpk_host.cu only; uses its fail() and the
// constants of cpk_device.cuh).
#pragma once

// ---------------------------------------------------------------------------
// Stream builder of the compact (one-CTA) walk: see DevCompact in cpk_device.cuh.
// A "step" is a set of items that may run concurrently and ends with a CTA
// barrier; a dependency level is one step (plus one step per extra part of rows
// too long for a single item).  Runs of one-item steps become CW_SEQ groups.
// ---------------------------------------------------------------------------
struct CwItemH {
    int kind;                               // CW_ROWS / CW_WARPROW / CW_DCHUNK
    int width, stride;                      // D chunk: count in `width`, has2x2 in `stride`
    int z;                                  // warp-row: target; D chunk: first row
    std::vector<unsigned char> data;        // multiple of 16 bytes
};

struct CwStream {
    std::vector<unsigned char> bytes;       // finished blocks
    int nblk = 0;
    // block under construction
    std::vector<int> table;                 // 4 ints per slot, kCwWarps slots per step
    std::vector<unsigned char> blob;
    long long n_levels = 0, n_steps = 0, n_items = 0;
    int rot = 0;                            // first warp of the next step (rotates, so that the warp
                                            // busy in one step is rarely the one busy in the next)
    size_t used(size_t more_steps, size_t more_bytes) const {
        return 16 + 4 * table.size() + 16 * kCwWarps * more_steps + blob.size() + more_bytes;
    }
    void flush() {
        if (table.empty()) return;
        const size_t nsteps = table.size() / (4 * kCwWarps);
        const size_t area = 16 + 4 * table.size();
        std::vector<unsigned char> blk(kCwBlock, 0);
        int *h = reinterpret_cast<int *>(blk.data());
        h[0] = (int)nsteps;
        int *dst = h + 4;
        for (size_t i = 0; i < table.size(); i += 4) {
            dst[i] = table[i] + ((table[i + 1] & 15) ? (int)area : 0);
            dst[i + 1] = table[i + 1]; dst[i + 2] = table[i + 2]; dst[i + 3] = table[i + 3];
        }
        if (!blob.empty()) memcpy(blk.data() + area, blob.data(), blob.size());
        bytes.insert(bytes.end(), blk.begin(), blk.end());
        ++nblk;
        table.clear(); blob.clear();
    }
    // appends one step (a row of kCwWarps empty slots) and returns its first slot
    size_t new_step() { table.resize(table.size() + 4 * kCwWarps, 0); ++n_steps; return table.size() - 4 * kCwWarps; }
    void set_slot(size_t step0, int w, int off, const CwItemH &it) {
        int *sl = &table[step0 + 4 * w];
        sl[0] = off; sl[1] = it.kind | (it.width << 8) | (it.stride << 24); sl[2] = it.z;
    }
    int put_data(const CwItemH &it) {
        const int off = (int)blob.size();
        blob.insert(blob.end(), it.data.begin(), it.data.end());
        ++n_items;
        return off;
    }
    void barrier_after(size_t step0) { for (int w = 0; w < kCwWarps; ++w) table[step0 + 4 * w + 1] |= CW_BARRIER; }
    // One dependency level (or the D pass): its items run concurrently, a CTA barrier
    // closes it.  More than kCwWarps items take several steps, only the last has the barrier.
    void level(const std::vector<CwItemH> &items) {
        if (items.empty()) return;
        ++n_levels;
        if (items[0].kind == CW_DCHUNK) {
            size_t st = 0;
            for (auto &it : items) {        // a chunk is shared by all warps: warp w takes rows 32w .. 32w+31
                if (used(1, it.data.size()) > (size_t)kCwBlock) flush();
                st = new_step();
                const int off = put_data(it);
                for (int w = 0; w < kCwWarps; ++w) if (32 * w < it.width) set_slot(st, w, off, it);
            }
            barrier_after(st);
            return;
        }
        size_t i = 0, st = 0;
        while (i < items.size()) {
            if (used(1, items[i].data.size()) > (size_t)kCwBlock) flush();      // the level continues in the next block
            st = new_step();
            for (int q = 0; q < kCwWarps && i < items.size(); ++q) {
                if (used(0, items[i].data.size()) > (size_t)kCwBlock) break;
                set_slot(st, (rot + q) % kCwWarps, put_data(items[i]), items[i]);
                ++i;
            }
            rot = (rot + 5) % kCwWarps;
        }
        barrier_after(st);
    }
};

static void cw_put(std::vector<unsigned char> &b, const void *src, size_t n)
{
    const unsigned char *s = static_cast<const unsigned char *>(src);
    if (n) b.insert(b.end(), s, s + n);
}

// One row of a level, addresses absolute in the shared vector: sv[tgt] -= sum val_k * sv[col_k]
struct CwRowH { int tgt; std::vector<int> col; std::vector<double> val; int len() const { return (int)col.size(); } };

// Emits the rows of ONE dependency level (independent of one another): short rows 32 per item,
// long rows one warp each (in parts of <= 512 entries; the extra parts follow as levels of their own).
static void cw_emit_level(CwStream &S, std::vector<CwRowH> &rows)
{
    constexpr int kMaxWidth = 16;           // entries per lane in one item
    if (rows.empty()) return;
    typedef std::vector<CwItemH> Items;
    std::stable_sort(rows.begin(), rows.end(), [&](const CwRowH &a, const CwRowH &b) { return a.len() < b.len(); });
    Items first;
    std::vector<Items> later;           // later[j-1] = parts j of the long rows
    size_t k = 0;
    // short rows: up to 32 per item, one lane each
    while (k < rows.size() && rows[k].len() <= kLongRow) {
        size_t k1 = k;
        int width = 0;
        while (k1 < rows.size() && k1 - k < 32 && rows[k1].len() <= kLongRow) { width = std::max(width, rows[k1].len()); ++k1; }
        const int stride = ((int)(k1 - k) + 3) & ~3;        // lanes stored (idle lanes of a thin item are not)
        CwItemH it{width <= 2 ? CW_ROWS2 : CW_ROWS, width, stride, 0, {}};
        if (width <= 2) {
            // one 32-byte record per lane: {target, col0, col1, 0, val0, val1}
            for (int q = 0; q < stride; ++q) {
                int rec_i[4] = {-1, -1, -1, 0};
                double rec_v[2] = {0.0, 0.0};
                if (k + q < k1) {
                    const CwRowH &r = rows[k + q];
                    rec_i[0] = r.tgt;
                    for (int j = 0; j < r.len(); ++j) { rec_i[1 + j] = r.col[j]; rec_v[j] = r.val[j]; }
                }
                cw_put(it.data, rec_i, 16);
                cw_put(it.data, rec_v, 16);
            }
        } else {
            int tgt[32];
            for (int q = 0; q < 32; ++q) tgt[q] = (k + q < k1) ? rows[k + q].tgt : -1;
            cw_put(it.data, tgt, (size_t)4 * stride);
            std::vector<double> val((size_t)width * stride, 0.0);
            std::vector<int> col((size_t)width * stride, -1);
            for (size_t q = k; q < k1; ++q) {
                const CwRowH &r = rows[q];
                for (int j = 0; j < r.len(); ++j) {
                    val[(size_t)j * stride + (q - k)] = r.val[j];
                    col[(size_t)j * stride + (q - k)] = r.col[j];
                }
            }
            cw_put(it.data, val.data(), val.size() * 8);
            cw_put(it.data, col.data(), col.size() * 4);
        }
        first.push_back(std::move(it));
        k = k1;
    }
    // long rows: the 32 lanes share the row, <= 32*kMaxWidth entries per part
    for (; k < rows.size(); ++k) {
        const CwRowH &r = rows[k];
        const int len = r.len();
        int part = 0;
        for (int e0 = 0; e0 < len; e0 += 32 * kMaxWidth, ++part) {
            const int cnt = std::min(len - e0, 32 * kMaxWidth);
            const int width = (cnt + 31) / 32;
            CwItemH it{CW_WARPROW, width, 32, r.tgt, {}};
            std::vector<double> val((size_t)width * 32, 0.0);
            std::vector<int> col((size_t)width * 32, -1);
            for (int j = 0; j < cnt; ++j) { val[j] = r.val[e0 + j]; col[j] = r.col[e0 + j]; }
            cw_put(it.data, val.data(), val.size() * 8);
            cw_put(it.data, col.data(), col.size() * 4);
            if (part == 0) first.push_back(std::move(it));
            else {
                if ((int)later.size() < part) later.resize(part);
                later[part - 1].push_back(std::move(it));
            }
        }
    }
    S.level(first);
    for (auto &st : later) S.level(st);
}

// Slots of the staging area behind the sweep values (see cw_sweep): DevCompact / cw_smem_bytes reserve them.
struct CwStage { int base = -1, used = 0; };        // base < 0: chain merging off

// items of one sweep direction.  R.row(i) = dependency list of LDL row i as (row id, value);
// `toff`/`coff`: offsets of the target and of the dependencies in the shared vector sv[2N].
//
// CHAIN MERGING (stage.base >= 0).  A level costs the walkers a barrier and a slot fetch (~400 cycles)
// on top of its arithmetic, and the tail of a filled factor is a chain of levels of ONE or two rows
// (cvxqp1: 47 consecutive one-row levels).  A run of such levels is rewritten as two levels: the
// rows of the run substitute their in-run dependencies, so that each becomes a linear form in the
// values the vector holds when the run starts -- sv[i] = sv[i] + delta_i, delta_i = -sum_c F_i(c) sv[c],
// F_i = L_i,: - sum_{j in run} L_ij F_j -- and, the sweep being in place, the deltas are first
// collected in a staging area behind the vector (level 1: nobody's target is anybody's input) and
// then added to their rows (level 2).  The staging slots are zeroed when a solve starts, every slot
// is used by one row only.  Guards: entries <= 8 x the original ones, |F| <= 1e3 x max|L|; a run the guards cut short continues as a new group.
static void cw_sweep(CwStream &S, int N, const HCsr &R, const std::vector<int> &lev, int nlev, int toff, int coff, CwStage *stage = nullptr)
{
    std::vector<std::vector<int>> byLevel(nlev);
    for (int i = 0; i < N; ++i) if (R.len(i) > 0) byLevel[lev[i]].push_back(i);
    // ---- runs of tiny levels
    static const int kChainItems = [] { const char *e = getenv("CPK_CW_CHAIN_ITEMS"); return e ? atoi(e) : 2; }();
    static const int kChainFill = [] { const char *e = getenv("CPK_CW_CHAIN_FILL"); return e ? atoi(e) : 8; }();
    constexpr int kChainMin = 4;
    std::vector<int> run_end(nlev, -1);         // run_end[l0] = last level of the run that starts at l0
    if (stage && stage->base >= 0) {
        auto tiny = [&](int l) {
            if (byLevel[l].empty()) return false;
            int nshort = 0, nlong = 0;
            for (int r : byLevel[l]) { if (R.len(r) <= kLongRow) ++nshort; else ++nlong; }
            return (nshort + 31) / 32 + nlong <= kChainItems;
        };
        for (int l = 0; l < nlev;) {
            if (!tiny(l)) { ++l; continue; }
            int e = l;
            while (e + 1 < nlev && tiny(e + 1)) ++e;
            if (e - l + 1 >= kChainMin) run_end[l] = e;
            l = e + 1;
        }
    }
    double maxL = 1.0;
    for (double v : R.val) maxL = std::max(maxL, std::fabs(v));
    std::vector<double> acc;
    std::vector<int> mark, touched;
    for (int l = 0; l < nlev; ++l) {
        if (run_end[l] >= 0) {
            // ---- try to merge the levels l .. run_end[l] (or a prefix of the run)
            if (acc.empty()) { acc.assign((size_t)N, 0.0); mark.assign((size_t)N, -1); }
            std::vector<int> rows_g;                          // rows of the merged levels, level order
            std::unordered_map<int, size_t> pos;              // row -> index in F
            std::vector<std::vector<std::pair<int, double>>> F;
            long long orig = 0, fill = 0;
            int last = l - 1;
            for (int ll = l; ll <= run_end[l]; ++ll) {
                bool ok = stage->used + (int)rows_g.size() + (int)byLevel[ll].size() <= kCwStage;
                std::vector<std::vector<std::pair<int, double>>> Fl;
                long long lorig = 0, lfill = 0;
                for (size_t q = 0; q < byLevel[ll].size() && ok; ++q) {
                    const int i = byLevel[ll][q];
                    touched.clear();
                    auto add = [&](int c, double v) {
                        if (mark[c] != i) { mark[c] = i; acc[c] = 0.0; touched.push_back(c); }
                        acc[c] += v;
                    };
                    for (int64_t k = R.ptr[i]; k < R.ptr[i + 1]; ++k) {
                        const int j = R.col[k]; const double a = R.val[k];
                        ++lorig;
                        add(j, a);
                        auto it = pos.find(j);
                        if (it != pos.end()) for (auto &x : F[it->second]) add(x.first, -a * x.second);
                    }
                    std::sort(touched.begin(), touched.end());
                    std::vector<std::pair<int, double>> fi;
                    for (int c : touched) { fi.emplace_back(c, acc[c]); if (!(std::fabs(acc[c]) <= 1e3 * maxL)) ok = false; }
                    lfill += (long long)fi.size();
                    Fl.push_back(std::move(fi));
                }
                if (ok && fill + lfill > (long long)kChainFill * (orig + lorig) + 256) ok = false;
                if (!ok) break;
                for (size_t q = 0; q < byLevel[ll].size(); ++q) { pos[byLevel[ll][q]] = F.size(); rows_g.push_back(byLevel[ll][q]); F.push_back(std::move(Fl[q])); }
                orig += lorig; fill += lfill; last = ll;
            }
            if (getenv("CPK_VERBOSE"))
                fprintf(stderr, "[cpk] compact walk: run of tiny levels %d..%d, merged %d..%d (%zu rows, %lld -> %lld entries, staging used %d)\n",
                        l, run_end[l], l, last, rows_g.size(), orig, fill, stage->used);
            if (last - l + 1 >= kChainMin) {
                std::vector<CwRowH> comp, commit;
                for (size_t q = 0; q < rows_g.size(); ++q) {
                    const int slot = stage->base + stage->used++;
                    CwRowH r{slot, {}, {}};
                    for (auto &x : F[q]) { r.col.push_back(coff + x.first); r.val.push_back(x.second); }
                    comp.push_back(std::move(r));
                    commit.push_back(CwRowH{toff + rows_g[q], {slot}, {-1.0}});
                }
                cw_emit_level(S, comp);
                cw_emit_level(S, commit);
                if (last < run_end[l]) run_end[last + 1] = run_end[l];      // the rest of the run may form a group of its own
                l = last;               // the loop's ++l moves past the merged levels
                continue;
            }
            if (l < run_end[l]) run_end[l + 1] = run_end[l];
        }
        std::vector<int> &rows = byLevel[l];
        if (rows.empty()) continue;
        std::vector<CwRowH> lr;
        lr.reserve(rows.size());
        for (int r : rows) {
            CwRowH h{toff + r, {}, {}};
            for (int64_t k = R.ptr[r]; k < R.ptr[r + 1]; ++k) { h.col.push_back(coff + R.col[k]); h.val.push_back(R.val[k]); }
            lr.push_back(std::move(h));
        }
        cw_emit_level(S, lr);
    }
}

static void cw_dpass(CwStream &S, int N, const std::vector<double> &d, const std::vector<double> &e, const std::vector<int> &partner)
{
    std::vector<CwItemH> items;
    for (int i0 = 0; i0 < N;) {
        bool has2 = false;
        int cnt = std::min(32 * kCwWarps, N - i0);      // one row per walker thread
        for (int r = 0; r < cnt; ++r) has2 = has2 || partner[i0 + r] >= 0;
        if (has2) cnt = std::min(cnt, 256);
        CwItemH it{CW_DCHUNK, cnt, has2 ? 1 : 0, i0, {}};
        cw_put(it.data, &d[i0], (size_t)cnt * 8);
        if (has2) {
            std::vector<double> ee(cnt, 0.0), dp(cnt, 1.0);
            std::vector<int> pr(cnt, -1);
            for (int r = 0; r < cnt; ++r) {
                const int i = i0 + r, q = partner[i];
                if (q < 0) continue;
                pr[r] = q; dp[r] = d[q]; ee[r] = e[std::min(i, q)];
            }
            cw_put(it.data, ee.data(), (size_t)cnt * 8);
            cw_put(it.data, dp.data(), (size_t)cnt * 8);
            cw_put(it.data, pr.data(), (size_t)cnt * 4);
        }
        while (it.data.size() % 16) it.data.push_back(0);
        items.push_back(std::move(it));
        i0 += cnt;
    }
    S.level(items);
}

