/*
 * opmb200.h -- C ABI of libopmb200.so: a B200-native (sm_100a) drop-in for OPM Flow's
 * per-Newton-step linear solve, BiCGSTAB preconditioned by ILU0 or DILU on the b x b block-CSR
 * Jacobian (b = 1..4).  No CPU fallback: every compute entry point runs hand-written CUDA
 * kernels and fails with OPMB200_CUDA_ERROR when no device is usable.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * opm-simulators tree).  INTEGRATION.md shows the Dune-side adapter classes that bind them.
 *
 * Conventions
 *   - Matrices are passed exactly as Dune::BCRSMatrix<Opm::MatrixBlock<double,b,b>> stores them
 *     (gpuistl/GpuSparseMatrix.cpp:164-167, gpubridge/GpuBridge.cpp:211-228): rowptr[n+1],
 *     colidx[nnzb] ascending per row with the diagonal present, values[nnzb*b*b] row-major per
 *     block, contiguous from &A[0][0][0][0].  Vectors are n*b doubles (Dune::BlockVector).
 *   - Every `double*` argument may be a HOST pointer (pageable or pinned) or a DEVICE pointer;
 *     the library detects which (cudaPointerGetAttributes) and stages host data itself.
 *   - The caller owns all the memory it passes; the handle owns all device memory, streams,
 *     graphs and its reference to the communicator (SURVEY.md section 8b "Ownership").
 *   - One handle = one device = one caller thread; calls on a handle are blocking at return
 *     (AbstractISTLSolver.hpp:43-212: prepare then solve, strictly sequential).
 *   - Return value: 0 on success, else an opmb200_status; opmb200_last_error() holds the text the
 *     adapter puts into the exception it throws (error conventions below).
 */
#ifndef OPMB200_H
#define OPMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPMB200_VERSION 100

typedef enum opmb200_status {
    OPMB200_SUCCESS = 0,
    OPMB200_INVALID_ARGUMENT = 1,   /* null pointer, bad sizes, unsupported block size            */
    OPMB200_BAD_OPTIONS = 2,        /* malformed JSON, unknown "solver" / "preconditioner.type":
                                       adapter throws std::invalid_argument
                                       (PreconditionerFactory_impl.hpp:98-106,
                                        FlexibleSolver_impl.hpp:326-329)                          */
    OPMB200_MATRIX_BLOCK_ERROR = 3, /* singular diagonal block: adapter throws
                                       Dune::MatrixBlockError (matrixblock.hh:209-217,
                                       ParallelOverlappingILU0_impl.hpp:588-606); all ranks agree  */
    OPMB200_SOLVER_ABORT = 4,       /* BiCGSTAB breakdown or NaN/Inf defect: Dune::SolverAbort     */
    OPMB200_CUDA_ERROR = 5,         /* std::runtime_error, like OPM_GPU_SAFE_CALL                  */
    OPMB200_NCCL_ERROR = 6,
    OPMB200_DIAGONAL_MISSING = 7,   /* Dune::ISTLError "diagonal entry missing"                    */
    OPMB200_NOT_PREPARED = 8        /* solve/apply before the first update_values                  */
} opmb200_status;

/* == Dune::InverseOperatorResult (filled by IterativeSolver::Iteration, dune-istl solver.hh) */
typedef struct opmb200_result {
    int iterations;   /* (int) of the half-step counter 0.5, 1, 1.5 ...                           */
    double reduction; /* |r| / |r0|                                                               */
    int converged;    /* |r| < reduction*|r0| or |r| < 1e-30; non-convergence is NOT an error here:
                         AbstractISTLSolver::checkConvergence (AbstractISTLSolver.hpp:192-211)
                         judges it                                                                */
    double conv_rate; /* reduction^(1/it)                                                         */
    double elapsed;   /* seconds, wall clock of the call                                          */
} opmb200_result;

typedef struct opmb200_solver opmb200_solver; /* opaque */
typedef struct opmb200_comm opmb200_comm;     /* opaque: one NCCL rank                             */

/* Owner/copy index lists, the content of Dune::OwnerOverlapCopyCommunication as the reference's
 * GPU sender flattens it (gpuistl/GpuAwareMPISender.hpp:164-222): for neighbour k,
 * send_rows[send_ptr[k]..send_ptr[k+1]) are local OWNER rows whose values the peer holds as
 * copies, recv_rows[recv_ptr[k]..recv_ptr[k+1]) are local COPY (ghost) rows owned by the peer,
 * both in the order agreed with the peer. */
typedef struct opmb200_halo {
    int n_neighbors;
    const int* neighbor_rank;
    const int* send_ptr;
    const int* send_rows;
    const int* recv_ptr;
    const int* recv_rows;
} opmb200_halo;

typedef struct opmb200_info {
    int block_size;
    int64_t n_rows, n_interior, nnzb;
    int n_levels;          /* level sets of getMatrixRowColoring(A, LOWER)                         */
    int n_slices;          /* 32-row scheduling slices of the level-ordered layout                 */
    int64_t padded_blocks; /* block slots of the device layout (>= nnzb)                           */
    int structurally_symmetric;
    int preconditioner;    /* 0 none, 1 dilu, 2 ilu0                                               */
    double relaxation;
    double tol;
    int maxiter;
    int n_ranks;
    double t_analysis_s;   /* host analysis in opmb200_create                                      */
    double t_update_ms;    /* device time of the last update_values (H2D + relayout + factorise)   */
    double t_solve_ms;     /* device time of the last solve                                        */
    int64_t kernel_launches; /* kernels launched by this handle so far                             */
    int schedule;          /* what was built: 0 level-scheduled sweeps, 1 tile walkers              */
    int n_chunks;          /* chunks of the chunked schedule                                       */
    int chunk_rows;        /* rows per chunk (chosen by the analysis when not given)               */
    double est_steps;      /* analysis estimate of the sweep length in local steps                 */
} opmb200_info;

/* ---- library ------------------------------------------------------------------------------ */
int opmb200_version(void);
const char* opmb200_last_error(void);
int opmb200_device_count(int* count);
/* gpuistl/set_device.cpp: bind rank -> device */
int opmb200_set_device(int device);

/* ---- host-side integer analysis (bit-exact against the reference) -------------------------- */
/* Opm::getMatrixRowColoring (GraphColoring.hpp:246-307); type 0 SYMMETRIC, 1 LOWER, 2 UPPER.
 * color[n], level_rows[n], level_ptr[n+1] (first *n_levels+1 entries valid). */
int opmb200_row_coloring(int64_t n, const int32_t* rowptr, const int32_t* colidx, int type,
                         int32_t* color, int32_t* level_rows, int32_t* level_ptr, int32_t* n_levels);
/* The sweep schedule of a sparsity pattern, host only (what opmb200_create builds; the threaded
 * reference orders rows by level set, DILU.hpp:82-91, 306-363).  schedule 0 = level sets, 1 = tiles
 * (chunk_rows > 0 forces contiguous chunks of that many rows, 0 chooses, -(TJ*100+TK) forces that tile
 * shape; box grids get tiles of TJ x TK grid lines, reported as chunk_rows = -(TJ*100+TK)), 2 = auto
 * (tiles on box grids whose rows have at most 4 lower / upper blocks, else level sets; *schedule_out
 * tells).  position_to_row[n]: rows in schedule order; slice_first[n+1] (first *n_slices+1 valid):
 * first position of every <= 32-row slice; chunk_first_slice[n+2] (first *n_chunks+1 valid; tiles only).
 * Invariant the sweeps rely on: a row depends only on rows of earlier slices, and only on rows of
 * the same or an earlier chunk.  Output pointers may be NULL. */
int opmb200_plan_schedule(int block_size, int64_t n_rows, int64_t nnzb, const int32_t* rowptr, const int32_t* colidx,
                          int64_t n_interior, int schedule, int chunk_rows, int32_t* n_slices, int32_t* n_chunks,
                          int32_t* chunk_rows_out, double* est_steps, int32_t* position_to_row, int32_t* slice_first,
                          int32_t* chunk_first_slice);
/* The step tables of the tile walkers for the same arguments, host only (tests replay them on the CPU).
 * info[8] = {schedule built, rows per step R, ring positions, dependency slots lower, upper, n_steps,
 * n_chunks, chunk_rows}.  With info[0] == 1 and non-NULL pointers: step_first[n_steps+1] positions,
 * chunk_first_step[n_chunks+1], step_flags[n_steps] (bit 0: ghost rows) and, for `direction` (0 lower,
 * 1 upper) with S slots and RP = (R+3)&~3: codes[n_steps*S*RP] ((1<<30) + ring positions: none | (1<<30) + ring index |
 * (1<<29) + external slot), ext[n_steps*64] positions, n_ext[n_steps].  Call once with NULL arrays to size. */
int opmb200_plan_tiles(int block_size, int64_t n_rows, int64_t nnzb, const int32_t* rowptr, const int32_t* colidx,
                       int64_t n_interior, int schedule, int chunk_rows, int32_t* info, int32_t* position_to_row,
                       int32_t* step_first, int32_t* chunk_first_step, int32_t* step_flags, int direction,
                       int32_t* codes, int32_t* ext, int32_t* n_ext);
/* Opm::partitionCellsSimple (opm/simulators/flow/partitionCells.cpp:734-751) */
int opmb200_partition_simple(int32_t num_cells, int32_t num_domains, int32_t* part);
/* Ghost-last local system of one rank with one overlap layer (FlowGenericVanguard.hpp:79,
 * ISTLSolver.hpp:299-306): owners first (ascending global index), then the cells adjacent to an
 * owner (ascending); owner rows complete, ghost rows hold only an identity diagonal
 * (== after makeOverlapRowsInvalid, ISTLSolver.cpp:56-75).  Call with out_* == NULL to size
 * (returns n_local, n_interior, nnzb_local), then again with buffers:
 * l2g[n_local], rowptr[n_local+1], colidx[nnzb_local], src[nnzb_local] (index of the global
 * block each local block copies, -1 for ghost diagonals). */
int opmb200_localize(int64_t n_global, const int32_t* rowptr, const int32_t* colidx, const int32_t* part,
                     int32_t rank, int64_t* n_local, int64_t* n_interior, int64_t* nnzb_local,
                     int32_t* out_l2g, int32_t* out_rowptr, int32_t* out_colidx, int64_t* out_src);

/* ---- communicator (block-Jacobi across subdomains) ------------------- */
/* 128-byte NCCL unique id made on rank 0 and broadcast by the host (MPI_Bcast in Flow,
 * torch.distributed in bench.py). */
int opmb200_comm_unique_id(void* id128);
int opmb200_comm_create(int rank, int n_ranks, const void* id128, opmb200_comm** out);
int opmb200_comm_destroy(opmb200_comm* comm);

/* ---- solver handle ------------------------------------------------------------------------- */
/* == Dune::FlexibleSolver<Op>(op, [comm,] prm, weights, pressureIndex) -> init -> initOpPrecSp +
 *    initSolver (FlexibleSolver_impl.hpp:142-330) and the GpuDILU / OpmGpuILU0 constructors'
 *    analysis (gpuistl/GpuDILU.cpp:48, OpmGpuILU0.cpp:42).
 * json_options: the solver property tree as JSON text (what --linear-solver=file.json loads,
 *    setupPropertyTree.cpp:190-203), or NULL for the defaults.  Keys read: solver ("bicgstab"),
 *    tol (1e-2), maxiter (200), verbosity (0), preconditioner.type ("paroverilu0" | "ilu0" |
 *    "dilu" | the reference's GPU aliases "opmilu0", "opmgpuilu0", "gpudilu"), preconditioner.
 *    relaxation (1.0), preconditioner.ilulevel (0; > 0 is rejected).  The reference's GPU tuning keys
 *    preconditioner.split_matrix / tune_gpu_kernels / reorder (StandardPreconditioners_gpu_serial.hpp:77-80, 92-96)
 *    are accepted, type-checked and have no effect (one implementation here); preconditioner.
 *    mixed_precision_scheme != 0 is rejected (fp64 storage only).  Values may be JSON strings or numbers
 *    (boost::property_tree stores strings).
 * Stream ordering: every call works on the handle's own stream and returns after that stream has drained.  A
 *    DEVICE pointer argument (values, x, b, ...) must be complete when the call is made -- the library does not
 *    order itself behind the stream that produced it (synchronise that stream, or its event, first).
 * n_interior: number of owner rows (== n_rows when serial); rows >= n_interior are ghosts.
 * comm, halo: NULL when serial. */
int opmb200_create(const char* json_options, int block_size, int64_t n_rows, int64_t nnzb,
                   const int32_t* rowptr, const int32_t* colidx, int64_t n_interior, opmb200_comm* comm,
                   const opmb200_halo* halo, opmb200_solver** out);
int opmb200_destroy(opmb200_solver* s);

/* == GpuSparseMatrixWrapper::updateNonzeroValues + PreconditionerWithUpdate::update()
 *    (gpuistl/ISTLSolverGPUISTL.hpp:425-440; DILU.hpp:110-118;
 *    ParallelOverlappingILU0_impl.hpp:431-610).  values: nnzb*b*b doubles, host or device. */
int opmb200_update_values(opmb200_solver* s, const double* values);

/* == Dune::Preconditioner::apply(v, d) incl. the BlockPreconditioner halo copy
 *    (DILU.hpp:135-143, ParallelOverlappingILU0_impl.hpp:361-419, OwningBlockPreconditioner.hpp) */
int opmb200_precond_apply(opmb200_solver* s, double* v, const double* d);

/* == Dune::LinearOperator::apply(x, y): y = A x            (WellOperators.hpp:432-443) */
int opmb200_op_apply(opmb200_solver* s, const double* x, double* y);
/* == Dune::LinearOperator::applyscaleadd(alpha, x, y): y += alpha A x   (:446-456) */
int opmb200_op_applyscaleadd(opmb200_solver* s, double alpha, const double* x, double* y);
/* == Dune::ScalarProduct::dot / norm (owner rows, summed over ranks; gpuistl/GpuSender.hpp:89-110) */
int opmb200_dot(opmb200_solver* s, const double* x, const double* y, double* result);

/* == Dune::InverseOperator::apply(x, b, [reduction,] res) of Dune::BiCGSTABSolver
 *    (FlexibleSolver_impl.hpp:94-107, 214-220).  x: in = initial guess, out = solution;
 *    b: OVERWRITTEN with the final residual, as Dune does.  reduction < 0: use "tol". */
int opmb200_solve(opmb200_solver* s, double* x, double* b, double reduction, opmb200_result* res);

/* ---- standard wells kept outside the matrix (matrix-add-well-contributions=false, Flow's default) ----
 * == WellModelMatrixAdapter / WellModelGhostLastMatrixAdapter (WellOperators.hpp:224-287, 300-360) around
 *    WellModelAsLinearOperator::apply / applyscaleadd (:84-109, 144-164) -> StandardWellEquations::apply
 *    (wells/StandardWellEquations.cpp:132-148); the reference's GPU twin is WellContributionsCuda::apply
 *    (gpubridge/cuda/cuWellContributions.cu:37-130, 166-190).
 * After this call every operator application of the handle (opmb200_solve's SpMVs, opmb200_op_apply,
 * opmb200_op_applyscaleadd) is  (A - sum_w C_w^T D_w^-1 B_w) x ; the preconditioner stays that of A, as in Flow.
 *   well_ptr[n_wells+1]  perforation range of every well;  cells[n_perf]  local row of each perforation
 *   B, C   [n_perf][dim_wells][block_size] row-major (duneB_ / duneC_: one dim_wells x block_size block per
 *          perforation; C is applied transposed),  Dinv [n_wells][dim_wells][dim_wells] (invDuneD_)
 * All host arrays (the well equations live on the host); copied at every call, the index tables are rebuilt only
 * when the structure changed.  n_wells == 0 removes the wells.  Call again after every well assembly. */
int opmb200_set_wells(opmb200_solver* s, int n_wells, int dim_wells, const int32_t* well_ptr, const int32_t* cells,
                      const double* B, const double* C, const double* Dinv);

/* ---- CPR pieces either side of the ILU0/DILU smoother -------------------------------------------------
 * All vectors in the caller's natural order, host or device pointers; `transpose` as in the reference
 * (false: weights multiply the equations (rows), the default "quasiimpes"/"trueimpes" CPR; true: CPRT).
 * == Amg::getQuasiImpesWeights (getQuasiImpesWeights.hpp:64-111; gpuistl/detail/cpr_amg_operations.cu:35-76):
 *    weights[n*b] from the diagonal blocks of the values of the last opmb200_update_values */
int opmb200_cpr_quasi_impes_weights(opmb200_solver* s, int pressure_index, int transpose, double* weights);
/* == PressureTransferPolicy::calculateCoarseEntries (PressureTransferPolicy.hpp; cpr_amg_operations.cu:79-123):
 *    coarse_values[k] for every block k of the caller's BCSR (blocks of ghost rows: 0) */
int opmb200_cpr_coarse_entries(opmb200_solver* s, const double* weights, int pressure_index, int transpose,
                               double* coarse_values);
/* == PressureTransferPolicy::moveToCoarseLevel (cpr_amg_operations.cu:126-151): coarse[n] from fine[n*b] */
int opmb200_cpr_restrict(opmb200_solver* s, const double* weights, int pressure_index, int transpose,
                         const double* fine, double* coarse);
/* == PressureTransferPolicy::moveToFineLevel (PressureTransferPolicy.hpp:148-162; cpr_amg_operations.cu:154-178):
 *    transpose == false writes only the pressure component of fine[n*b] */
int opmb200_cpr_prolongate(opmb200_solver* s, const double* weights, int pressure_index, int transpose,
                           const double* coarse, double* fine);

/* ---- introspection (parity tests) ---------------------------------------------------------- */
int opmb200_get_info(opmb200_solver* s, opmb200_info* info);
/* level sets exactly as getMatrixRowColoring(A, LOWER) returns them: level_ptr[n_levels+1],
 * level_rows[n] (pass NULL to skip an output) */
int opmb200_get_levels(opmb200_solver* s, int32_t* level_ptr, int32_t* level_rows);
/* DILU.hpp:83-91 reordered_to_natural / natural_to_reordered */
int opmb200_get_reorder(opmb200_solver* s, int32_t* reordered_to_natural, int32_t* natural_to_reordered);
/* MultithreadDILU::getDiagonal(): n*b*b doubles, natural row order (host pointer) */
int opmb200_get_dinv(opmb200_solver* s, double* dinv);
/* the in-place block-ILU0 factor in the caller's BCSR order (L strictly lower, inverted
 * diagonal, U strictly upper): nnzb*b*b doubles (host pointer) */
int opmb200_get_ilu0(opmb200_solver* s, double* lu);
/* residual norm after every half step of the last solve (hist[0] = |r0|); returns count */
int opmb200_get_history(opmb200_solver* s, double* hist, int capacity, int* count);

/* ---- measurement ---------------------------------------------------------------------------
 * Times `reps` back-to-back launches of one kernel group on the handle's own stream with CUDA
 * events (after `warmup` untimed launches) and returns the average milliseconds per launch and
 * the algorithmic bytes one launch moves (SURVEY.md section 8d formulas).
 * what: 0 SpMV (y = A x), 1 preconditioner apply (lower + upper sweep), 2 preconditioner update
 *       (relayout + factorisation), 3 the fused BiCGSTAB vector kernels of one iteration,
 *       4 lower sweep only, 5 upper sweep only, 6 (experiment) upper sweep with an SpMV beside it,
 *       7 CPR quasi-IMPES weights, 8 CPR coarse entries, 9 CPR restriction + prolongation.
 *       With wells set (opmb200_set_wells), 0 times the well-corrected operator (well kernel + SpMV).
 * A measurement aid, not part of the solve path: it overwrites the handle's work vectors and device scalars, and
 * what = 2 re-runs the last update from its device-resident values (a caller's DEVICE value buffer must still be
 * alive).  Solves after it are unaffected (every solve re-initialises its state). */
int opmb200_time_kernel(opmb200_solver* s, int what, int warmup, int reps, double* ms_per_launch,
                        double* algorithmic_bytes);

/* ---- peer-to-peer collectives (optional, multi-GPU) -------------------------------------------
 * Replaces the NCCL calls of the Krylov loop by collectives that run INSIDE the library's own
 * kernels over NVLink peer memory (one process per GPU, peers mapped with CUDA IPC): the last CTA
 * of every fused reduction kernel pushes its partial sums into the peers' mailboxes and adds the
 * partials of all ranks in rank order (= OwnerOverlapCopyCommunication::sum, gpuistl/GpuSender.hpp:
 * 89-95), and the halo copy writes owner rows straight into the neighbour's receive buffer
 * (= copyOwnerToAll, gpuistl/GpuAwareMPISender.hpp:55-136).
 * Every rank exports one blob, the host all-gathers them in rank order (MPI_Allgather in Flow,
 * torch.distributed in bench.py) and hands the concatenation to opmb200_p2p_import. */
#define OPMB200_P2P_BLOB_BYTES 1024
int opmb200_p2p_export(opmb200_solver* s, void* blob);
int opmb200_p2p_import(opmb200_solver* s, const void* blobs_in_rank_order);

/* Device-time stopwatch on the handle's own stream (CUDA events): start, run any sequence of
 * calls on the handle, stop -> elapsed milliseconds between the two events. */
int opmb200_timer_start(opmb200_solver* s);
int opmb200_timer_stop(opmb200_solver* s, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* OPMB200_H */
