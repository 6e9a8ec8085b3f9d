// analysis.cpp -- host-side integer analysis (see layout.hpp for the reference counterparts).
#include "layout.hpp"

#include "../../include/opmb200.h"

#include <algorithm>
#include <numeric>

namespace opmb200 {

namespace {
inline int64_t find_in_row(const int32_t* rowptr, const int32_t* col, int64_t i, int32_t j)
{
    const int32_t* b = col + rowptr[i];
    const int32_t* e = col + rowptr[i + 1];
    const int32_t* it = std::lower_bound(b, e, j);
    return (it != e && *it == j) ? (it - col) : -1;
}
} // namespace

// Opm::getMatrixRowColoring: level[i] = 1 + max level over the rows i depends on; rows of a
// level keep their natural order (stable sort == counting sort).
int row_coloring(int64_t n, const int32_t* rowptr, const int32_t* col, int type, int32_t* color,
                 int32_t* level_rows, int32_t* level_ptr)
{
    if (type < 0 || type > 2)
        return -OPMB200_INVALID_ARGUMENT;
    std::fill(color, color + n, 0);
    int32_t nlev = 0;
    if (type == 2) { // UPPER: dependencies on the entries right of the diagonal, rows descending
        for (int64_t i = n - 1; i >= 0; --i) {
            const int64_t kd = find_in_row(rowptr, col, i, (int32_t)i);
            if (kd < 0)
                return -OPMB200_DIAGONAL_MISSING;
            int32_t c = 0;
            for (int64_t k = kd + 1; k < rowptr[i + 1]; ++k)
                c = std::max(c, color[col[k]] + 1);
            color[i] = c;
            nlev = std::max(nlev, c + 1);
        }
    } else { // LOWER / SYMMETRIC: entries left of the diagonal, rows ascending
        for (int64_t i = 0; i < n; ++i) {
            int32_t c = 0;
            int64_t k = rowptr[i];
            for (; k < rowptr[i + 1] && col[k] != i; ++k) {
                const int32_t j = col[k];
                if (type == 0 && find_in_row(rowptr, col, j, (int32_t)i) < 0)
                    continue; // SYMMETRIC: only when A_ji is stored as well
                c = std::max(c, color[j] + 1);
            }
            if (k == rowptr[i + 1])
                return -OPMB200_DIAGONAL_MISSING; // the reference walks until it meets the diagonal
            color[i] = c;
            nlev = std::max(nlev, c + 1);
        }
    }
    std::vector<int32_t> cnt(nlev + 1, 0);
    for (int64_t i = 0; i < n; ++i)
        ++cnt[color[i] + 1];
    for (int32_t l = 0; l < nlev; ++l)
        cnt[l + 1] += cnt[l];
    for (int32_t l = 0; l <= nlev; ++l)
        level_ptr[l] = cnt[l];
    for (int64_t i = 0; i < n; ++i)
        level_rows[cnt[color[i]]++] = (int32_t)i;
    return nlev;
}

void partition_simple(int32_t num_cells, int32_t num_domains, int32_t* part)
{
    // Opm::partitionCellsSimple: num_cells / num_domains each, the first (num_cells mod
    // num_domains) domains get one more; contiguous in the natural ordering.
    const int32_t q = num_cells / num_domains, r = num_cells % num_domains;
    int32_t c = 0;
    for (int32_t d = 0; d < num_domains; ++d)
        for (int32_t e = c + q + (d < r ? 1 : 0); c < e; ++c)
            part[c] = d;
}

namespace {
// Group ids of the "chunks" schedule for a given chunk size, plus a dataflow estimate of the sweep
// length in units of one local step (an external dependency costs `t_ext` steps).
struct ChunkPlan {
    std::vector<int32_t> grp;      // group id per row
    std::vector<int32_t> grp_chunk; // chunk of every group
    int32_t n_groups = 0;
    int32_t n_chunks = 0;
    double est_steps = 0;
    double total_steps = 0;
};

// chunk ids of contiguous runs of R rows of the natural ordering.  A chunk ends after R rows, at the
// owner/ghost border, and wherever the wavefront restarts: a row whose global level lies below the
// level of the chunk's first row could have started before this chunk did, so queueing it behind
// the chunk would serialise independent work (on a box grid: the first row of the next plane
// behind the last strip of the current plane).
int32_t contiguous_chunks(int64_t n, int64_t n_interior, const std::vector<int32_t>& glev, int64_t R,
                          std::vector<int32_t>& chunk_id)
{
    chunk_id.resize(n);
    int32_t c = -1;
    int64_t begin = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (i == 0 || i - begin >= R || i == n_interior || glev[i] < glev[begin]) {
            ++c;
            begin = i;
        }
        chunk_id[i] = c;
    }
    return c + 1;
}

// chunk ids of TJ x TK tiles of grid lines on a box grid in natural order (nx cells per line, ny
// lines per plane): one line per row of a step, so that the j-1 AND the k-1 neighbour of most rows
// are served by the chunk's own ring.  Tiles are numbered by the time the wavefront reaches them
// (TJ*tj + TK*tk: the step at which tile (tj, tk) can start), so that the tiles the in-order ticket
// keeps resident are the ones the front is working on.  Ghost rows (>= n_interior) get contiguous
// chunks behind the tiles.  The caller validates the result (chunk_order_valid): the pattern need
// not be a clean 7-point stencil.
int32_t tile_chunks(int64_t n, int64_t n_interior, int64_t nx, int64_t ny, int TJ, int TK, std::vector<int32_t>& chunk_id)
{
    chunk_id.resize(n);
    const int64_t nxy = nx * ny;
    const int64_t ntj = (ny + TJ - 1) / TJ;
    const int64_t nz = n_interior > 0 ? (n_interior - 1) / nxy + 1 : 0;
    const int64_t ntk = (nz + TK - 1) / TK;
    std::vector<int64_t> order(ntj * ntk);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        return TJ * (a % ntj) + TK * (a / ntj) < TJ * (b % ntj) + TK * (b / ntj);
    });
    std::vector<int32_t> rank(ntj * ntk, 0);
    for (size_t r = 0; r < order.size(); ++r)
        rank[order[r]] = (int32_t)r;
    int32_t nc = 0;
    for (int64_t i = 0; i < n_interior; ++i) {
        const int64_t k = i / nxy, j = (i % nxy) / nx;
        chunk_id[i] = rank[(k / TK) * ntj + j / TJ];
        nc = std::max(nc, chunk_id[i] + 1);
    }
    for (int64_t i = n_interior; i < n; ++i)
        chunk_id[i] = nc + (int32_t)((i - n_interior) / 1024);
    if (n > n_interior)
        nc = chunk_id[n - 1] + 1;
    return nc;
}

// every dependency must point to the same or an earlier chunk (lower sweep; the upper sweep walks
// the chunks backwards), otherwise in-order chunk start could deadlock
template <class RowBegin, class RowEnd>
bool chunk_order_valid(int64_t n, const int32_t* col, const std::vector<int32_t>& chunk_id, RowBegin row_begin, RowEnd row_end)
{
    for (int64_t r = 0; r < n; ++r)
        for (int64_t k = row_begin(r); k < row_end(r); ++k) {
            const int32_t c = col[k];
            if ((c < r && chunk_id[c] > chunk_id[r]) || (c > r && chunk_id[c] < chunk_id[r]))
                return false;
        }
    return true;
}

template <class RowBegin, class RowEnd>
void plan_chunks(int64_t n, const int32_t* col, const std::vector<int32_t>& chunk_id, int32_t n_chunks,
                 RowBegin row_begin, RowEnd row_end, double t_ext, int R, ChunkPlan& P)
{
    P.n_chunks = n_chunks;
    auto chunk_of = [&](int64_t i) { return (int64_t)chunk_id[i]; };
    std::vector<int32_t> llev(n, 0);
    for (int64_t r = 0; r < n; ++r) {
        const int64_t cr = chunk_of(r);
        for (int64_t k = row_begin(r); k < row_end(r); ++k) {
            const int32_t c = col[k];
            if (c < r && chunk_of(c) == cr)
                llev[r] = std::max(llev[r], llev[c] + 1);
        }
        for (int64_t k = row_begin(r); k < row_end(r); ++k) {
            const int32_t c = col[k];
            if (c > r && chunk_of(c) == cr)
                llev[c] = std::max(llev[c], llev[r] + 1);
        }
    }
    std::vector<int32_t> nl(P.n_chunks + 1, 0);
    for (int64_t i = 0; i < n; ++i)
        nl[chunk_of(i) + 1] = std::max(nl[chunk_of(i) + 1], llev[i] + 1);
    for (int32_t c = 0; c < P.n_chunks; ++c)
        nl[c + 1] += nl[c];
    P.n_groups = nl[P.n_chunks];
    P.grp.resize(n);
    P.grp_chunk.assign(P.n_groups, 0);
    std::vector<int32_t> width(P.n_groups, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int32_t g = nl[chunk_of(i)] + llev[i];
        P.grp[i] = g;
        P.grp_chunk[g] = (int32_t)chunk_of(i);
        ++width[g];
    }
    // Dataflow estimate of the lower sweep: a group starts when the previous group of its chunk is
    // done and its external dependencies (rows of earlier chunks) have arrived.  Chunks are
    // evaluated in order, which is a valid evaluation order because external producers always
    // belong to earlier chunks.
    std::vector<double> done(P.n_groups, 0.0);
    double total = 0, tmax = 0;
    // bucket rows by group
    std::vector<int32_t> gptr(P.n_groups + 1, 0), grows(n);
    for (int64_t i = 0; i < n; ++i)
        ++gptr[P.grp[i] + 1];
    for (int32_t g = 0; g < P.n_groups; ++g)
        gptr[g + 1] += gptr[g];
    {
        std::vector<int32_t> cur(gptr.begin(), gptr.end() - 1);
        for (int64_t i = 0; i < n; ++i)
            grows[cur[P.grp[i]]++] = (int32_t)i;
    }
    for (int32_t g = 0; g < P.n_groups; ++g) {
        double start = (g > 0 && P.grp_chunk[g - 1] == P.grp_chunk[g]) ? done[g - 1] : 0.0;
        for (int32_t t = gptr[g]; t < gptr[g + 1]; ++t) {
            const int64_t r = grows[t];
            for (int64_t k = row_begin(r); k < row_end(r); ++k) {
                const int32_t c = col[k];
                if (c < r && P.grp_chunk[P.grp[c]] != P.grp_chunk[g])
                    start = std::max(start, done[P.grp[c]] + t_ext);
            }
        }
        const double cost = (width[g] + R - 1) / R;
        done[g] = start + cost;
        total += cost;
        tmax = std::max(tmax, done[g]);
    }
    P.est_steps = tmax;
    P.total_steps = total;
}
} // namespace

int build_layout(int b, int64_t n, int64_t nnzb, const int32_t* rowptr, const int32_t* col, int64_t n_interior,
                 bool want_ilu0, int schedule_mode, int chunk_rows, Layout& L, std::string& err)
{
    if (b < 1 || b > 4) {
        err = "block size must be 1..4";
        return OPMB200_INVALID_ARGUMENT;
    }
    if (n < 0 || n_interior < 0 || n_interior > n || !rowptr || (!col && nnzb > 0) || rowptr[0] != 0
        || rowptr[n] != nnzb || nnzb >= (int64_t(1) << 31) - 64 * kSlice) {
        err = "inconsistent BCSR arguments";
        return OPMB200_INVALID_ARGUMENT;
    }
    L.b = b;
    L.n = n;
    L.nnzb = nnzb;
    L.n_interior = n_interior;

    // ---- validation + diagonal positions ----------------------------------------------------
    std::vector<int32_t> diag(n);
    for (int64_t i = 0; i < n; ++i) {
        if (rowptr[i + 1] < rowptr[i]) {
            err = "rowptr not monotone";
            return OPMB200_INVALID_ARGUMENT;
        }
        int64_t kd = -1;
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            if (col[k] < 0 || col[k] >= n || (k > rowptr[i] && col[k] <= col[k - 1])) {
                err = "column indices must be ascending and in range (row " + std::to_string(i) + ")";
                return OPMB200_INVALID_ARGUMENT;
            }
            if (col[k] == i)
                kd = k;
        }
        if (kd < 0) {
            err = "diagonal entry missing in row " + std::to_string(i);
            return OPMB200_DIAGONAL_MISSING;
        }
        diag[i] = (int32_t)kd;
    }

    // ---- reference-exact LOWER colouring of the pattern as given ----------------------------
    {
        std::vector<int32_t> color(n);
        L.ref_level_rows.resize(n);
        L.ref_level_ptr.resize(n + 1);
        const int nl = row_coloring(n, rowptr, col, 1, color.data(), L.ref_level_rows.data(), L.ref_level_ptr.data());
        if (nl < 0) {
            err = "diagonal entry missing";
            return -nl;
        }
        L.ref_level_ptr.resize(nl + 1);
    }

    // ---- schedule levels on pattern(A) U pattern(A^T), ghost rows reduced to their diagonal ---
    // entry (r,c) of an owner row: c < r  => r waits for c ; c > r => c waits for r.
    auto row_begin = [&](int64_t i) { return i < n_interior ? (int64_t)rowptr[i] : (int64_t)diag[i]; };
    auto row_end = [&](int64_t i) { return i < n_interior ? (int64_t)rowptr[i + 1] : (int64_t)diag[i] + 1; };
    std::vector<int32_t> lev(n, 0);
    bool symmetric = true;
    for (int64_t r = 0; r < n; ++r) {
        for (int64_t k = row_begin(r); k < row_end(r); ++k) {
            const int32_t c = col[k];
            if (c < r)
                lev[r] = std::max(lev[r], lev[c] + 1);
        }
        for (int64_t k = row_begin(r); k < row_end(r); ++k) {
            const int32_t c = col[k];
            if (c > r)
                lev[c] = std::max(lev[c], lev[r] + 1);
            if (c != r && c < n_interior && r < n_interior && symmetric && find_in_row(rowptr, col, c, (int32_t)r) < 0)
                symmetric = false;
        }
    }
    L.symmetric = symmetric;
    // the tile walkers stage at most 4 dependency blocks per row and direction; wider rows (NNC- or
    // well-like) keep the level schedule
    int max_wl = 0, max_wu = 0;
    for (int64_t i = 0; i < n_interior; ++i) {
        max_wl = std::max<int>(max_wl, diag[i] - rowptr[i]);
        max_wu = std::max<int>(max_wu, rowptr[i + 1] - 1 - diag[i]);
    }
    const bool auto_mode = schedule_mode == 2;
    const int max_slots = b >= 3 ? 4 : 3; // instantiated tile walkers (solver.cu, DISPATCH_BS)
    if (schedule_mode != 0 && (max_wl > max_slots || max_wu > max_slots || n >= (int64_t)kTwExt))
        schedule_mode = 0;
    // rows of one CTA step of the tile walkers == one slice of the SELL layout the SpMV and the factorisation
    // kernels work on
    const int R = kTwRows;
    static_assert(kTwRows == kSlice, "a step is a slice");
    int32_t nlev = 0;
    for (int64_t i = 0; i < n; ++i)
        nlev = std::max(nlev, lev[i] + 1);
    if (n == 0)
        nlev = 0;
    std::vector<int32_t> grp_chunk;
    if (schedule_mode != 0 && n > 0) {
        // chunk size: given, or the candidate with the shortest estimated sweep (critical path of
        // the chunk dataflow, or total work over the ~296 tile walkers a B200 keeps resident)
        ChunkPlan best;
        const double t_ext = 6.0, resident = 296.0;
        std::vector<int64_t> cands;
        int64_t line = 0, plane = 0; // nx and nx*ny of a box grid in natural order (0: not recognised)
        if (chunk_rows > 0) {
            cands.push_back(chunk_rows);
        }
        if (chunk_rows <= 0) {
            // "line length" of the ordering = the most frequent lower offset larger than 1 (nx on a
            // box grid)
            std::vector<int64_t> offs;
            const int64_t stride = std::max<int64_t>(1, n_interior / 200000);
            for (int64_t r = 0; r < n_interior; r += stride)
                for (int64_t k = row_begin(r); k < row_end(r); ++k)
                    if (col[k] < r - 1)
                        offs.push_back(r - col[k]);
            std::sort(offs.begin(), offs.end());
            int64_t bestc = 0, bestp = 0;
            for (size_t i = 0; i < offs.size();) {
                size_t j = i;
                while (j < offs.size() && offs[j] == offs[i])
                    ++j;
                // prefer the smallest frequent offset: a plane offset is as frequent as a line offset
                if ((int64_t)(j - i) > bestc + bestc / 4) {
                    bestc = (int64_t)(j - i);
                    line = offs[i];
                }
                i = j;
            }
            for (size_t i = 0; i < offs.size();) { // plane offset: the most frequent offset beyond the line offset
                size_t j = i;
                while (j < offs.size() && offs[j] == offs[i])
                    ++j;
                if (offs[i] > line && (int64_t)(j - i) > bestp + bestp / 4) {
                    bestp = (int64_t)(j - i);
                    plane = offs[i];
                }
                i = j;
            }
            if (bestp * 4 < bestc)
                plane = 0;
            for (int64_t C : {(int64_t)1024, (int64_t)2048, (int64_t)4096, (int64_t)8192, (R / 2) * line, R * line, 2 * R * line})
                if (chunk_rows == 0 && !auto_mode && C >= 256 && C <= 65536 && std::find(cands.begin(), cands.end(), C) == cands.end())
                    cands.push_back(C);
        }
        double best_cost = -1;
        auto consider = [&](const std::vector<int32_t>& chunk_id, int32_t nc, int64_t tag) {
            ChunkPlan P;
            plan_chunks(n, col, chunk_id, nc, row_begin, row_end, t_ext, R, P);
            const double padding = P.total_steps * R / (double)std::max<int64_t>(n, 1);
            const double cost = std::max(P.est_steps, P.total_steps / resident) + 200.0 * std::max(0.0, padding - 1.6);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best = std::move(P);
                L.chunk_rows = (int)tag;
            }
        };
        std::vector<int32_t> chunk_id;
        for (int64_t C : cands) {
            const int32_t nc = contiguous_chunks(n, n_interior, lev, C, chunk_id);
            consider(chunk_id, nc, C);
        }
        // box grids: tiles of R grid lines, one line per row of a step (chunk_rows reports
        // -(TJ*100 + TK); a negative request forces that tile shape)
        bool tiled = false;
        if (chunk_rows <= 0 && line > 1 && plane > line && plane % line == 0) {
            for (int TK = 1; TK <= R; ++TK) {
                if (R % TK)
                    continue;
                const int TJ = R / TK;
                if (chunk_rows < 0 && chunk_rows != -(TJ * 100 + TK))
                    continue;
                if (chunk_rows == 0 && (TJ < 2 || TK < 2))
                    continue;
                const int32_t nc = tile_chunks(n, n_interior, line, plane / line, TJ, TK, chunk_id);
                if (chunk_order_valid(n, col, chunk_id, row_begin, row_end)) {
                    consider(chunk_id, nc, -(TJ * 100 + TK));
                    tiled = true;
                }
            }
        }
        if (auto_mode && !tiled) {
            schedule_mode = 0; // "auto": only box grids leave the level schedule
        } else if (best_cost < 0) { // a forced tile shape that the pattern does not admit: contiguous chunks
            const int32_t nc = contiguous_chunks(n, n_interior, lev, 2048, chunk_id);
            consider(chunk_id, nc, 2048);
        }
        if (schedule_mode != 0) {
            schedule_mode = 1;
            lev = best.grp; // group id replaces the level from here on
            nlev = best.n_groups;
            grp_chunk = best.grp_chunk;
            L.n_chunks = best.n_chunks;
            L.est_steps = best.est_steps;
        }
    }
    L.schedule_mode = schedule_mode;
    L.n_levels = nlev;
    L.level_q0.assign(nlev + 1, 0);
    for (int64_t i = 0; i < n; ++i)
        ++L.level_q0[lev[i] + 1];
    for (int32_t l = 0; l < nlev; ++l)
        L.level_q0[l + 1] += L.level_q0[l];
    L.r2n.resize(n);
    L.n2r.resize(n);
    if (schedule_mode == 0) {
        // Level schedule: inside a level the rows are sorted by (lower width, upper width, natural index) before
        // they are cut into 32-row slices (sigma-sorted SELL): a slice is as wide as its widest row, so rows of
        // equal width share slices.  Irregular patterns (inactive cells, NNCs: the Norne-sized C2) lose 1.50 ->
        // 1.35 of padding; on box grids nothing changes.  Counting sort on (level, min(wl,7), min(wu,7)).
        std::vector<int64_t> cnt((size_t)nlev * 64 + 1, 0);
        auto key = [&](int64_t i) {
            const int wl = i < n_interior ? std::min<int>(diag[i] - rowptr[i], 7) : 0;
            const int wu = i < n_interior ? std::min<int>(rowptr[i + 1] - 1 - diag[i], 7) : 0;
            return (size_t)lev[i] * 64 + (size_t)wl * 8 + wu;
        };
        for (int64_t i = 0; i < n; ++i)
            ++cnt[key(i) + 1];
        for (size_t k = 0; k + 1 < cnt.size(); ++k)
            cnt[k + 1] += cnt[k];
        for (int64_t i = 0; i < n; ++i) {
            const int32_t q = (int32_t)cnt[key(i)]++;
            L.r2n[q] = (int32_t)i;
            L.n2r[i] = q;
        }
    } else {
        std::vector<int32_t> cur(L.level_q0.begin(), L.level_q0.end());
        for (int64_t i = 0; i < n; ++i) {
            const int32_t q = cur[lev[i]]++;
            L.r2n[q] = (int32_t)i;
            L.n2r[i] = q;
        }
    }

    // ---- mode 1: cut the groups into the steps of the tile walkers ------------------------------
    // A step holds <= R rows (one CTA step) and <= kTwMaxExt dependencies per direction that the
    // chunk's ring does not serve (one per lane of a poll warp).  A dependency is served by the ring
    // iff it lies in the same chunk at most `window` positions away: the ring holds tw_ring
    // positions and a step overwrites at most R of them, so everything up to tw_ring - R positions
    // before the first row of a step (behind its last row, upper sweep) is intact.
    int tw_window = 0;
    if (schedule_mode == 1) {
        L.tw_rows = R;
        L.tw_ring = 256;
        while (L.tw_ring < 4 * L.tw_rows)
            L.tw_ring *= 2;
        L.tw_slots[0] = std::max(3, max_wl);
        L.tw_slots[1] = std::max(3, max_wu);
        tw_window = L.tw_ring - 2 * L.tw_rows;
        auto n_ext = [&](int64_t i, bool upper) {
            if (i >= n_interior)
                return 0;
            int cnt = 0;
            const int32_t q = L.n2r[i];
            for (int64_t k = upper ? diag[i] + 1 : rowptr[i]; k < (upper ? rowptr[i + 1] : diag[i]); ++k) {
                const int32_t c = col[k], pp = L.n2r[c];
                if (!(grp_chunk[lev[c]] == grp_chunk[lev[i]] && (upper ? pp - q : q - pp) <= tw_window))
                    ++cnt;
            }
            return cnt;
        };
        std::vector<int32_t> step_chunk;
        L.step_q0.clear();
        for (int32_t g = 0; g < nlev; ++g) {
            int32_t q = L.level_q0[g];
            while (q < L.level_q0[g + 1]) {
                L.step_q0.push_back(q);
                step_chunk.push_back(grp_chunk[g]);
                int count = 0, elo = 0, eup = 0;
                while (q < L.level_q0[g + 1] && count < R) {
                    const int64_t i = L.r2n[q];
                    const int a = n_ext(i, false), u = n_ext(i, true);
                    if (count > 0 && (elo + a > kTwMaxExt || eup + u > kTwMaxExt))
                        break;
                    ++count;
                    elo += a;
                    eup += u;
                    ++q;
                }
            }
        }
        L.n_steps = (int)step_chunk.size();
        L.step_q0.push_back((int32_t)n);
        L.chunk_step0.assign(L.n_chunks + 1, 0);
        for (int32_t st = 0; st < L.n_steps; ++st) // steps are ordered by chunk
            L.chunk_step0[step_chunk[st] + 1] = st + 1;
        for (int32_t c = 0; c < L.n_chunks; ++c)
            L.chunk_step0[c + 1] = std::max(L.chunk_step0[c + 1], L.chunk_step0[c]);
        L.step_flags.assign(L.n_steps, 0);
        for (int32_t st = 0; st < L.n_steps; ++st)
            if (L.r2n[L.step_q0[st]] >= n_interior)
                L.step_flags[st] = 1;
        // the chunk of a ROW is needed again for the dependency codes below
        std::vector<int32_t> row_chunk(n);
        for (int64_t i = 0; i < n; ++i)
            row_chunk[i] = grp_chunk[lev[i]];
        lev.swap(row_chunk); // lev[i] = chunk of row i from here on
        // steps replace the groups: no slice crosses a step
        grp_chunk = step_chunk;
        nlev = L.n_steps;
        L.n_levels = nlev;
        L.level_q0 = L.step_q0;
    }

    // ---- slices --------------------------------------------------------------------------------
    L.slice_q0.clear();
    L.slice_level.clear();
    L.level_slice0.assign(nlev + 1, 0);
    for (int32_t l = 0; l < nlev; ++l) {
        L.level_slice0[l] = (int32_t)L.slice_level.size();
        for (int32_t q = L.level_q0[l]; q < L.level_q0[l + 1]; q += kSlice) {
            L.slice_q0.push_back(q);
            L.slice_level.push_back(l);
        }
    }
    L.level_slice0[nlev] = (int32_t)L.slice_level.size();
    L.n_slices = (int)L.slice_level.size();
    if (schedule_mode == 1) {
        L.chunk_slice0.assign(L.n_chunks + 1, 0);
        for (int32_t g = 0; g < nlev; ++g) // groups are ordered by chunk
            L.chunk_slice0[grp_chunk[g] + 1] = L.level_slice0[g + 1];
        for (int32_t c = 0; c < L.n_chunks; ++c)
            L.chunk_slice0[c + 1] = std::max(L.chunk_slice0[c + 1], L.chunk_slice0[c]);
    }
    L.slice_q0.push_back((int32_t)n);
    L.slice_wl.assign(L.n_slices, 0);
    L.slice_wu.assign(L.n_slices, 0);
    L.slice_base.assign(L.n_slices + 1, 0);
    L.slice_lrank.assign(L.n_slices + 1, 0);
    for (int s = 0; s < L.n_slices; ++s) {
        int wl = 0, wu = 0;
        for (int32_t q = L.slice_q0[s]; q < L.slice_q0[s + 1]; ++q) {
            const int64_t i = L.r2n[q];
            if (i < n_interior) {
                wl = std::max<int>(wl, diag[i] - rowptr[i]);
                wu = std::max<int>(wu, rowptr[i + 1] - 1 - diag[i]);
            }
        }
        L.slice_wl[s] = wl;
        L.slice_wu[s] = wu;
        L.slice_base[s + 1] = L.slice_base[s] + wl + 1 + wu;
        L.slice_lrank[s + 1] = L.slice_lrank[s] + wl;
        if ((int64_t)L.slice_base[s] + wl + 1 + wu >= (int64_t(1) << 31) / kSlice) {
            err = "matrix too large for 32-bit slot ids";
            return OPMB200_INVALID_ARGUMENT;
        }
    }
    L.n_slot_rows = L.slice_base[L.n_slices];

    // ---- slots -----------------------------------------------------------------------------------
    const int64_t nslots = L.n_slot_rows * kSlice;
    L.slot_col.assign(nslots, -1);
    L.slot_src.assign(nslots, -1);
    std::vector<int32_t> slot_of_native(nnzb, -1);
    for (int s = 0; s < L.n_slices; ++s) {
        const int64_t base = L.slice_base[s];
        const int wl = L.slice_wl[s];
        for (int32_t q = L.slice_q0[s]; q < L.slice_q0[s + 1]; ++q) {
            const int lane = q - L.slice_q0[s];
            const int64_t i = L.r2n[q];
            const int64_t gd = (base + wl) * kSlice + lane;
            L.slot_col[gd] = q;
            if (i >= n_interior) {
                L.slot_src[gd] = -2; // ghost row: identity (ISTLSolver.cpp:56-75)
                continue;
            }
            L.slot_src[gd] = diag[i];
            slot_of_native[diag[i]] = (int32_t)gd;
            int sl = 0;
            for (int64_t k = rowptr[i]; k < diag[i]; ++k, ++sl) { // L: ascending column
                const int64_t g = (base + sl) * kSlice + lane;
                L.slot_col[g] = L.n2r[col[k]];
                L.slot_src[g] = (int32_t)k;
                slot_of_native[k] = (int32_t)g;
            }
            int su = 0;
            for (int64_t k = rowptr[i + 1] - 1; k > diag[i]; --k, ++su) { // U: descending column
                const int64_t g = (base + wl + 1 + su) * kSlice + lane;
                L.slot_col[g] = L.n2r[col[k]];
                L.slot_src[g] = (int32_t)k;
                slot_of_native[k] = (int32_t)g;
            }
        }
    }

    // ---- mode 1: the order in which the factorisation kernels take the slices ---------------------
    // They run one warp per slice with an in-order ticket.  In (chunk, step) order consecutive slices depend
    // on each other and only a few chunks would be in flight; ordered by the level of a slice in the slice
    // dependency graph they sweep the whole front at once, like the level schedule does.
    L.factor_order.clear();
    if (schedule_mode == 1) {
        std::vector<int32_t> slice_of(n), slev(L.n_slices, 0);
        for (int sl = 0; sl < L.n_slices; ++sl)
            for (int32_t q = L.slice_q0[sl]; q < L.slice_q0[sl + 1]; ++q)
                slice_of[q] = sl;
        for (int sl = 0; sl < L.n_slices; ++sl) // dependencies point to earlier slices
            for (int32_t q = L.slice_q0[sl]; q < L.slice_q0[sl + 1]; ++q) {
                const int64_t i = L.r2n[q];
                if (i >= n_interior)
                    continue;
                for (int64_t k = rowptr[i]; k < diag[i]; ++k)
                    slev[sl] = std::max(slev[sl], slev[slice_of[L.n2r[col[k]]]] + 1);
            }
        L.factor_order.resize(L.n_slices);
        std::iota(L.factor_order.begin(), L.factor_order.end(), 0);
        std::stable_sort(L.factor_order.begin(), L.factor_order.end(), [&](int32_t x, int32_t y) { return slev[x] < slev[y]; });
    }

    // ---- mode 1: dependency codes, external lists and block slots of every step -------------------
    for (int d = 0; d < 2; ++d) {
        L.tw_code[d].clear();
        L.tw_ext[d].clear();
        L.tw_next[d].clear();
        L.tw_slot[d].clear();
        L.tw_pub[d].clear();
    }
    if (schedule_mode == 1) {
        const int RP = L.tw_rp();
        for (int d = 0; d < 2; ++d) {
            const int S = L.tw_slots[d];
            L.tw_code[d].assign((size_t)L.n_steps * S * RP, kTwRing | L.tw_ring); // "no dependency": the zero record behind the ring
            L.tw_ext[d].assign((size_t)L.n_steps * kTwMaxExt, -1);
            L.tw_next[d].assign(L.n_steps, 0);
            L.tw_slot[d].assign((size_t)S * n, -1);
            L.tw_pub[d].assign((size_t)L.n_steps * 4, 0);
        }
        for (int32_t st = 0; st < L.n_steps; ++st) {
            for (int32_t q = L.step_q0[st]; q < L.step_q0[st + 1]; ++q) {
                const int rho = q - L.step_q0[st];
                const int64_t i = L.r2n[q];
                if (i >= n_interior)
                    continue; // ghost rows: identity
                for (int d = 0; d < 2; ++d) {
                    const int S = L.tw_slots[d];
                    const int w = d ? rowptr[i + 1] - 1 - diag[i] : diag[i] - rowptr[i];
                    for (int k = 0; k < w; ++k) { // lower: ascending column; upper: descending (slot order of the SELL layout)
                        const int64_t e = d ? (int64_t)rowptr[i + 1] - 1 - k : (int64_t)rowptr[i] + k;
                        const int32_t c = col[e], pp = L.n2r[c];
                        int32_t code;
                        if (lev[c] == lev[i] && (d ? pp - q : q - pp) <= tw_window) {
                            code = kTwRing | (pp & (L.tw_ring - 1));
                        } else {
                            const int l = L.tw_next[d][st]++;
                            L.tw_ext[d][(size_t)st * kTwMaxExt + l] = pp;
                            code = kTwExt | l;
                        }
                        L.tw_code[d][((size_t)st * S + k) * RP + rho] = code;
                        L.tw_slot[d][(size_t)k * n + q] = slot_of_native[e];
                    }
                }
            }
        }
        // rows somebody polls
        std::vector<int32_t> step_of(n);
        for (int32_t st = 0; st < L.n_steps; ++st)
            for (int32_t q = L.step_q0[st]; q < L.step_q0[st + 1]; ++q)
                step_of[q] = st;
        for (int d = 0; d < 2; ++d)
            for (int32_t st = 0; st < L.n_steps; ++st)
                for (int l = 0; l < L.tw_next[d][st]; ++l) {
                    const int32_t pp = L.tw_ext[d][(size_t)st * kTwMaxExt + l];
                    const int rho = pp - L.step_q0[step_of[pp]];
                    L.tw_pub[d][(size_t)step_of[pp] * 4 + (rho >> 5)] |= (int32_t)(1u << (rho & 31));
                }
    }

    // ---- DILU transposed-entry map and ILU0 update pairs -------------------------------------------
    const int64_t nl = L.n_l_slot_rows() * kSlice;
    L.l_transpose.assign(nl, -1);
    if (want_ilu0)
        L.trip_ptr.assign(nl + 1, 0);
    L.trip_src.clear();
    L.trip_dst.clear();
    for (int pass = 0; pass < (want_ilu0 ? 2 : 1); ++pass) {
        // pass 0: transposes + pair counts ; pass 1: fill pairs
        for (int s = 0; s < L.n_slices; ++s) {
            for (int32_t q = L.slice_q0[s]; q < L.slice_q0[s + 1]; ++q) {
                const int lane = q - L.slice_q0[s];
                const int64_t i = L.r2n[q];
                if (i >= n_interior)
                    continue;
                int sl = 0;
                for (int64_t kij = rowptr[i]; kij < diag[i]; ++kij, ++sl) {
                    const int64_t cl = ((int64_t)L.slice_lrank[s] + sl) * kSlice + lane; // compact L slot
                    const int32_t j = col[kij];
                    if (pass == 0) {
                        const int64_t kji = find_in_row(rowptr, col, j, (int32_t)i);
                        L.l_transpose[cl] = kji >= 0 ? slot_of_native[kji] : -1;
                    }
                    if (!want_ilu0)
                        continue;
                    // ParallelOverlappingILU0_impl.hpp:66-86: merge row j (right of its diagonal)
                    // with row i (right of ij)
                    int64_t jk = diag[j] + 1, ik = kij + 1;
                    int32_t w = pass ? L.trip_ptr[cl] : 0;
                    while (ik < rowptr[i + 1] && jk < rowptr[j + 1]) {
                        if (col[ik] == col[jk]) {
                            if (pass) {
                                L.trip_src[w] = slot_of_native[jk];
                                L.trip_dst[w] = slot_of_native[ik];
                            }
                            ++w;
                            ++ik;
                            ++jk;
                        } else if (col[ik] < col[jk]) {
                            ++ik;
                        } else {
                            ++jk;
                        }
                    }
                    if (!pass)
                        L.trip_ptr[cl + 1] = w; // count, prefix-summed below
                }
            }
        }
        if (want_ilu0 && pass == 0) {
            int64_t tot = 0;
            for (int64_t c = 0; c < nl; ++c) {
                const int32_t cnt = L.trip_ptr[c + 1];
                L.trip_ptr[c] = (int32_t)tot;
                tot += cnt;
                if (tot >= (int64_t(1) << 31)) {
                    err = "too many ILU0 update pairs";
                    return OPMB200_INVALID_ARGUMENT;
                }
            }
            // shift: trip_ptr[c] now holds starts; rebuild the n+1 form
            std::vector<int32_t> starts(L.trip_ptr.begin(), L.trip_ptr.begin() + nl);
            for (int64_t c = 0; c < nl; ++c)
                L.trip_ptr[c] = starts[c];
            L.trip_ptr[nl] = (int32_t)tot;
            L.trip_src.assign(tot, -1);
            L.trip_dst.assign(tot, -1);
        }
    }
    if (want_ilu0) {
        L.row_static.assign(n, 1);
        for (int s = 0; s < L.n_slices; ++s) {
            const int64_t gd0 = ((int64_t)L.slice_base[s] + L.slice_wl[s]) * kSlice;
            for (int32_t q = L.slice_q0[s]; q < L.slice_q0[s + 1]; ++q) {
                const int lane = q - L.slice_q0[s];
                const int64_t i = L.r2n[q];
                if (i >= n_interior)
                    continue;
                for (int sl = 0; sl < diag[i] - rowptr[i]; ++sl) {
                    const int64_t cl = ((int64_t)L.slice_lrank[s] + sl) * kSlice + lane;
                    for (int32_t t = L.trip_ptr[cl]; t < L.trip_ptr[cl + 1]; ++t)
                        if (L.trip_dst[t] != gd0 + lane)
                            L.row_static[q] = 0;
                }
            }
        }
    }
    return OPMB200_SUCCESS;
}

int localize(int64_t n_global, const int32_t* rowptr, const int32_t* col, const int32_t* part, int32_t rank,
             int64_t* n_local, int64_t* n_interior, int64_t* nnzb_local, int32_t* out_l2g, int32_t* out_rowptr,
             int32_t* out_col, int64_t* out_src)
{
    std::vector<int32_t> owners, ghosts;
    std::vector<char> is_ghost(n_global, 0);
    int64_t nnz = 0;
    for (int64_t g = 0; g < n_global; ++g) {
        if (part[g] != rank)
            continue;
        owners.push_back((int32_t)g);
        nnz += rowptr[g + 1] - rowptr[g];
        for (int64_t k = rowptr[g]; k < rowptr[g + 1]; ++k)
            if (part[col[k]] != rank)
                is_ghost[col[k]] = 1;
    }
    for (int64_t g = 0; g < n_global; ++g)
        if (is_ghost[g])
            ghosts.push_back((int32_t)g);
    *n_interior = (int64_t)owners.size();
    *n_local = (int64_t)(owners.size() + ghosts.size());
    *nnzb_local = nnz + (int64_t)ghosts.size();
    if (!out_l2g)
        return OPMB200_SUCCESS;
    std::vector<int32_t> g2l(n_global, -1);
    int32_t l = 0;
    for (int32_t g : owners) {
        out_l2g[l] = g;
        g2l[g] = l++;
    }
    for (int32_t g : ghosts) {
        out_l2g[l] = g;
        g2l[g] = l++;
    }
    int64_t w = 0;
    out_rowptr[0] = 0;
    std::vector<std::pair<int32_t, int64_t>> tmp;
    for (size_t li = 0; li < owners.size(); ++li) {
        const int64_t g = owners[li];
        tmp.clear();
        for (int64_t k = rowptr[g]; k < rowptr[g + 1]; ++k)
            tmp.emplace_back(g2l[col[k]], k);
        std::sort(tmp.begin(), tmp.end());
        for (auto& pr : tmp) {
            out_col[w] = pr.first;
            out_src[w] = pr.second;
            ++w;
        }
        out_rowptr[li + 1] = (int32_t)w;
    }
    for (size_t gi = 0; gi < ghosts.size(); ++gi) {
        const int32_t li = (int32_t)(owners.size() + gi);
        out_col[w] = li;
        out_src[w] = -1;
        ++w;
        out_rowptr[li + 1] = (int32_t)w;
    }
    return OPMB200_SUCCESS;
}

} // namespace opmb200
