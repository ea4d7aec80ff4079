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

// items of one sweep direction.  R.row(i) = dependency list of LDL row i as (row id, value);
// `toff`/`coff`: offsets of the target and of the dependencies in the shared vector sv[2N].
static void cw_sweep(CwStream &S, int N, const HCsr &R, const std::vector<int> &lev, int nlev, int toff, int coff)
{
    constexpr int kMaxWidth = 16;           // entries per lane in one item
    std::vector<std::vector<int>> byLevel(nlev);
    for (int i = 0; i < N; ++i) if (R.len(i) > 0) byLevel[lev[i]].push_back(i);
    typedef std::vector<CwItemH> Items;
    auto emit = [&](Items &&items) { S.level(items); };
    for (int l = 0; l < nlev; ++l) {
        std::vector<int> &rows = byLevel[l];
        if (rows.empty()) continue;
        std::stable_sort(rows.begin(), rows.end(), [&](int a, int b) { return R.len(a) < R.len(b); });
        Items first;
        std::vector<Items> later;           // later[j-1] = parts j of the long rows
        size_t k = 0;
        // short rows: up to 32 per item, one lane each
        while (k < rows.size() && R.len(rows[k]) <= kLongRow) {
            size_t k1 = k;
            int width = 0;
            while (k1 < rows.size() && k1 - k < 32 && R.len(rows[k1]) <= kLongRow) { width = std::max(width, R.len(rows[k1])); ++k1; }
            const int stride = ((int)(k1 - k) + 3) & ~3;        // lanes stored (idle lanes of a thin item are not)
            CwItemH it{width <= 2 ? CW_ROWS2 : CW_ROWS, width, stride, 0, {}};
            if (width <= 2) {
                // one 32-byte record per lane: {target, col0, col1, 0, val0, val1}
                for (int q = 0; q < stride; ++q) {
                    int rec_i[4] = {-1, -1, -1, 0};
                    double rec_v[2] = {0.0, 0.0};
                    if (k + q < k1) {
                        const int r = rows[k + q];
                        rec_i[0] = toff + r;
                        for (int j = 0; j < R.len(r); ++j) { rec_i[1 + j] = coff + R.col[R.ptr[r] + j]; rec_v[j] = R.val[R.ptr[r] + j]; }
                    }
                    cw_put(it.data, rec_i, 16);
                    cw_put(it.data, rec_v, 16);
                }
            } else {
                int tgt[32];
                for (int q = 0; q < 32; ++q) tgt[q] = (k + q < k1) ? toff + rows[k + q] : -1;
                cw_put(it.data, tgt, (size_t)4 * stride);
                std::vector<double> val((size_t)width * stride, 0.0);
                std::vector<int> col((size_t)width * stride, -1);
                for (size_t q = k; q < k1; ++q) {
                    const int r = rows[q];
                    for (int j = 0; j < R.len(r); ++j) {
                        val[(size_t)j * stride + (q - k)] = R.val[R.ptr[r] + j];
                        col[(size_t)j * stride + (q - k)] = coff + R.col[R.ptr[r] + j];
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
            const int r = rows[k], len = R.len(r);
            int part = 0;
            for (int e0 = 0; e0 < len; e0 += 32 * kMaxWidth, ++part) {
                const int cnt = std::min(len - e0, 32 * kMaxWidth);
                const int width = (cnt + 31) / 32;
                CwItemH it{CW_WARPROW, width, 32, toff + r, {}};
                std::vector<double> val((size_t)width * 32, 0.0);
                std::vector<int> col((size_t)width * 32, -1);
                for (int j = 0; j < cnt; ++j) { val[j] = R.val[R.ptr[r] + e0 + j]; col[j] = coff + R.col[R.ptr[r] + e0 + j]; }
                cw_put(it.data, val.data(), val.size() * 8);
                cw_put(it.data, col.data(), col.size() * 4);
                if (part == 0) first.push_back(std::move(it));
                else {
                    if ((int)later.size() < part) later.resize(part);
                    later[part - 1].push_back(std::move(it));
                }
            }
        }
        emit(std::move(first));
        for (auto &st : later) emit(std::move(st));
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

