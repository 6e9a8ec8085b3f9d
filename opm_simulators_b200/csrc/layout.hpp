// layout.hpp -- host-side analysis of a BCSR sparsity pattern into the device layout.
//
// Reference counterparts: Opm::getMatrixRowColoring (GraphColoring.hpp:246-307), the reorder maps
// of MultithreadDILU (DILU.hpp:83-91), createReorderedMatrix / extractLowerAndUpperMatrices
// (gpuistl/detail/coloringAndReorderingUtils.hpp:37-66) and the diagonal/transposed-entry index
// searches the reference does per element on the device (DILUKernels.cu:207-268).  Here all of
// it is integer work done once per sparsity pattern on the host.
//
// Device layout ("level-ordered SELL-32"):
//   * rows are renumbered by (schedule level, natural index): position q <-> natural row r2n[q];
//   * each level is cut into slices of <= 32 consecutive positions, one warp lane per row;
//   * a slice owns wL + 1 + wU consecutive "slot rows" of 32 block slots each: its strictly-lower
//     blocks (ascending natural column), its diagonal block, its strictly-upper blocks
//     (descending natural column -- the order DILU.hpp:293 and convertToCRS traverse them);
//     slot id g = slot_row*32 + lane;  value element e of slot g lives at
//     (g & ~31)*b*b + e*32 + (g & 31)  => every warp-wide load is one contiguous 256-byte line;
//   * slot_col[g] = POSITION of the column (-1: empty slot), slot_src[g] = index of the block in
//     the caller's BCSR values (-1: empty, -2: identity block of a ghost row).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace opmb200 {

constexpr int kSlice = 32;
constexpr int32_t kTwRing = 1 << 30, kTwExt = 1 << 29; // dependency codes of the tile walkers
constexpr int kTwMaxExt = 64;                        // external dependencies per step (two per poll lane)
constexpr int kTwWarps = 4;                          // compute warps of a tile walker
constexpr int kTwRows = 32;                          // rows of a step (one SELL slice): kTwWarps x 8 rows, b lanes each

struct Layout {
    int b = 0;
    int64_t n = 0, n_interior = 0, nnzb = 0;
    bool symmetric = true;

    // reference-exact artefacts (what the parity tests compare bit for bit)
    std::vector<int32_t> ref_level_rows, ref_level_ptr; // getMatrixRowColoring(A, LOWER)

    // schedule groups: rows of one group are mutually independent and depend only on rows of
    // earlier groups (lower sweep) / later groups (upper sweep).
    //   mode 0 "levels": group = level of pattern(A) U pattern(A^T) (== the reference level sets
    //                    when the pattern is structurally symmetric)
    //   mode 1 "tiles":  rows are cut into chunks (tiles of grid lines on a box grid, else contiguous
    //                    runs of the natural ordering), group = (chunk, level inside the chunk), a
    //                    group is cut into STEPS of <= tw_rows rows; one CTA walks one chunk step by
    //                    step (tile_kernels.cuh), so only dependencies that cross a chunk boundary
    //                    travel through the L2
    //   mode 2 "auto":   request only: tiles when the pattern admits them, else levels
    int schedule_mode = 0;
    int chunk_rows = 0;
    int n_levels = 0;              // number of groups
    std::vector<int32_t> level_q0; // [n_levels+1] first position of each group
    std::vector<int32_t> r2n, n2r; // position <-> natural row
    int n_chunks = 0;
    std::vector<int32_t> chunk_slice0; // [n_chunks+1] (mode 1)
    double est_steps = 0;              // schedule estimate of the chosen chunking (mode 1)

    // slices
    int n_slices = 0;
    std::vector<int32_t> slice_q0;    // [n_slices+1]
    std::vector<int32_t> slice_base;  // [n_slices+1] first slot row
    std::vector<int32_t> slice_wl;    // [n_slices]
    std::vector<int32_t> slice_wu;    // [n_slices]
    std::vector<int32_t> slice_level; // [n_slices]
    std::vector<int32_t> slice_lrank; // [n_slices+1] number of L slot rows before the slice
    std::vector<int32_t> level_slice0; // [n_levels+1]
    int64_t n_slot_rows = 0;

    std::vector<int32_t> slot_col; // [n_slot_rows*32]
    std::vector<int32_t> slot_src; // [n_slot_rows*32]
    // ---- mode 1: step tables of the tile walkers (tile_kernels.cuh) ------------------------------
    int tw_rows = 0;                  // rows per CTA step: kTwRows
    int tw_ring = 0;                  // positions of a chunk's shared-memory ring (power of two)
    int tw_slots[2] = {3, 3};         // dependency slots per row: lower, upper (3 or 4)
    int n_steps = 0;
    std::vector<int32_t> step_q0;     // [n_steps+1]
    std::vector<int32_t> chunk_step0; // [n_chunks+1]
    std::vector<int32_t> factor_order; // [n_slices] the factorisation kernels' ticket k works on slice factor_order[k]
    std::vector<int32_t> step_flags;  // [n_steps] bit 0: the step's rows are ghost rows
    // per direction d (0 lower, 1 upper), by step s (NOT in walking order):
    //   tw_code[d][(s*S + k)*RP + rho]: where row rho of the step finds dependency k:
    //        kTwRing + tw_ring none (the zero record behind the ring) | kTwRing + ring index | kTwExt + slot of the
    //        step's external list
    //   tw_ext[d][s*32 + l]: position of external dependency l (-1 beyond tw_next[d][s])
    //   tw_slot[d][k*n + q]: SELL slot of dependency block k of the row at position q (-1 none)
    //   tw_pub[d][s*4 + w]: bit rho of word w set = some other chunk (or a far step of this one) polls the
    //        result of row rho of the step: it is published with a strong store
    std::vector<int32_t> tw_code[2], tw_ext[2], tw_next[2], tw_slot[2], tw_pub[2];
    int tw_rp() const { return (tw_rows + 3) & ~3; }
    // DILU: for every L slot (compact L numbering) the slot of the transposed block, or -1
    std::vector<int32_t> l_transpose; // [n_l_slot_rows*32]
    // ILU0: per compact L slot, the (source slot in row j, destination slot in row i) update pairs
    std::vector<int32_t> trip_ptr; // [n_l_slot_rows*32 + 1]
    std::vector<int32_t> trip_src, trip_dst;
    // ILU0: 1 where the elimination of a row (by position) only ever updates its own diagonal block,
    // i.e. its U blocks stay the caller's values and may be read before the row is finished
    std::vector<int32_t> row_static; // [n]

    int64_t n_l_slot_rows() const { return slice_lrank.empty() ? 0 : slice_lrank.back(); }
};

// GraphColoring.hpp:246-307; type 0 SYMMETRIC, 1 LOWER, 2 UPPER.  Returns number of levels or a
// negative opmb200_status.
int row_coloring(int64_t n, const int32_t* rowptr, const int32_t* col, int type, int32_t* color,
                 int32_t* level_rows, int32_t* level_ptr);

// Builds everything above.  Returns 0 or an opmb200_status (diagonal missing, bad arguments).
// schedule_mode: 0 levels, 1 tiles, 2 auto; chunk_rows: rows per chunk (> 0: contiguous chunks of that many
// rows, 0: chosen automatically, -(TJ*100+TK): that tile shape)
int build_layout(int b, int64_t n, int64_t nnzb, const int32_t* rowptr, const int32_t* col, int64_t n_interior,
                 bool want_ilu0, int schedule_mode, int chunk_rows, Layout& L, std::string& err);

void partition_simple(int32_t num_cells, int32_t num_domains, int32_t* part);

int localize(int64_t n_global, const int32_t* rowptr, const int32_t* col, const int32_t* part, int32_t rank,
             int64_t* n_local, int64_t* n_interior, int64_t* nnzb_local, int32_t* out_l2g, int32_t* out_rowptr,
             int32_t* out_col, int64_t* out_src);

} // namespace opmb200
