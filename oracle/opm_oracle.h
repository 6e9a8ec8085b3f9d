/*
 * opm_oracle.h -- CPU restatement ("oracle") of OPM Flow's per-Newton-step linear solve.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker / the timed CPU baseline.  The product (libopmb200.so) never
 * links, loads or calls it.
 *
 * Parity status: PINNED for colouring, partition, DILU, the whole-solve golden vectors and
 * the un-preconditioned Krylov loop (see tests/test_oracle_golden.py for the list of
 * reference golden vectors it reproduces); the BiCGSTAB loop and the block-ILU0 kernel of
 * dune-istl (>= 2.9, not vendored in /root/reference, see dune.module:12) are restated from
 * the published algorithm and anchored on the reference's call sites -- iteration COUNTS on
 * multi-iteration solves and parallel halo results have no stored numbers in the reference
 * and are therefore "parity unpinned" (DESIGN.md section 2).  The same holds for the well operator
 * (orc_well_apply, orc_par_set_wells) and for the numpy restatements of the CPR transfer pieces in oracle.py: the
 * reference compares two live implementations (tests/gpuistl/test_GpuPressureTransferPolicy.cpp) and stores no
 * numbers; they are anchored on dense linear algebra (tests/test_oracle_wells_cpr.py).
 *
 * All matrices are block-CSR exactly as Dune::BCRSMatrix<Opm::MatrixBlock<double,b,b>> lays
 * them out (opm/simulators/linalg/gpuistl/GpuSparseMatrix.cpp:164-167): rowptr[n+1],
 * col[nnzb] ascending per row with the diagonal present, val[nnzb*b*b] row-major per block.
 */
#ifndef OPM_ORACLE_H
#define OPM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_OK = 0,
    ORC_ERR_DIAG_MISSING = 1, /* "diagonal entry missing" ISTLError */
    ORC_ERR_SINGULAR = 2,     /* Dune::MatrixBlockError */
    ORC_ERR_BREAKDOWN = 3,    /* Dune::SolverAbort: rho/omega/h breakdown */
    ORC_ERR_NAN = 4,          /* Dune::SolverAbort: defect is NaN/Inf */
    ORC_ERR_ARG = 5
};

enum { ORC_COLOR_SYMMETRIC = 0, ORC_COLOR_LOWER = 1, ORC_COLOR_UPPER = 2 };
enum { ORC_PREC_NONE = 0, ORC_PREC_DILU = 1, ORC_PREC_ILU0 = 2 };

typedef struct orc_result {
    int iterations;    /* (int) of Dune's half-step counter */
    double reduction;  /* norm / norm0 */
    int converged;
    double conv_rate;  /* reduction^(1/it) */
    double it;         /* the half-step counter itself (0, 0.5, 1, ...) */
    double norm0;
    double norm;
} orc_result;

/* GraphColoring.hpp:246-307.  color[n], level_rows[n], level_ptr[n+1]; returns #levels (<0 on error) */
int orc_row_coloring(int n, const int* rowptr, const int* col, int type,
                     int* color, int* level_rows, int* level_ptr);
/* DILU.hpp:83-91 */
void orc_reorder_maps(int n, const int* level_rows, int* reordered_to_natural, int* natural_to_reordered);
/* partitionCells.cpp:734-751 */
void orc_partition_simple(int num_cells, int num_domains, int* part);

/* matrixblock.hh:48-68,70-233 ; gpuistl/detail/deviceBlockOperations.hpp:37-114 */
int orc_invert_block(int b, double* blk);

/* WellOperators.hpp:432-468 (interior == n gives Dune::MatrixAdapter) */
void orc_spmv(int n, int b, const int* rowptr, const int* col, const double* val, int interior,
              const double* x, double* y);
void orc_spmv_scaleadd(int n, int b, const int* rowptr, const int* col, const double* val, int interior,
                       double alpha, const double* x, double* y);

/* WellOperators.hpp:84-91,144-164 + StandardWellEquations.cpp:132-148: y -= C^T (Dinv (B x)) well after well.
 * wptr[nw+1] perforation ranges, cells[nperf] perforated cell of each perforation, B/C [nperf][dw][b],
 * Dinv [nw][dw][dw], all row-major */
void orc_well_apply(int nw, int dw, int b, const int* wptr, const int* cells, const double* B, const double* C,
                    const double* Dinv, const double* x, double* y);

/* ISTLSolver.cpp:56-75 */
void orc_make_overlap_rows_invalid(int n, int b, const int* rowptr, const int* col, double* val, int interior);

/* DILU.hpp:186-206, 253-304 */
int orc_dilu_update(int n, int b, const int* rowptr, const int* col, const double* val, double* dinv);
void orc_dilu_apply(int n, int b, const int* rowptr, const int* col, const double* val, const double* dinv,
                    const double* d, double* v);

/* ParallelOverlappingILU0_impl.hpp:42-99 (== Dune::ILU::blockILU0Decomposition when interior == n) */
int orc_ilu0_decompose(int n, int b, const int* rowptr, const int* col, double* lu, int interior);
/* ParallelOverlappingILU0_impl.hpp:361-419 without the copyOwnerToAll / relaxation tail */
void orc_ilu0_apply(int n, int b, const int* rowptr, const int* col, const double* lu, int interior,
                    const double* d, double* v);

/* ---- the (possibly multi-subdomain) solver: P subdomains emulate P MPI ranks in one process ---- */
typedef struct orc_par orc_par;
orc_par* orc_par_create(int nsub, int b, long nglobal);
void orc_par_destroy(orc_par* h);
/* l2g may be NULL for a serial system (nsub == 1).  Arrays are borrowed, not copied. */
int orc_par_set_sub(orc_par* h, int p, int n, int interior, const int* rowptr, const int* col,
                    const double* val, const int* l2g);
/* wells of subdomain p stay outside the matrix (WellModelMatrixAdapter, WellOperators.hpp:224-287): every
 * op.apply / applyscaleadd of the solver adds the well operator.  Arrays are borrowed.  nw = 0 removes them. */
int orc_par_set_wells(orc_par* h, int p, int nw, int dw, const int* wptr, const int* cells, const double* B,
                      const double* C, const double* Dinv);
/* builds the preconditioner on every subdomain; kind = ORC_PREC_*; w = ILU0 relaxation */
int orc_par_prec_update(orc_par* h, int kind, double w);
/* Dune::BlockPreconditioner::apply: local apply, copyOwnerToAll, (ILU0: relaxation) */
int orc_par_prec_apply(orc_par* h, double** v, double** d);
void orc_par_copy_owner_to_all(orc_par* h, double** v);
double orc_par_dot(orc_par* h, double** x, double** y);
/* Dune::BiCGSTABSolver::apply; x in/out, b overwritten with the residual; op_repeats >= 1
 * (tests/test_preconditionerfactory.cpp:231-276 RepeatingOperator); hist may be NULL, else
 * receives norm after every half step (hist[0] = norm0), at most 2*maxiter+1 entries. */
int orc_par_bicgstab(orc_par* h, double** x, double** b, double reduction, int maxiter, int op_repeats,
                     orc_result* res, double* hist, int* nhist);
/* access to factor storage for tests */
const double* orc_par_dinv(orc_par* h, int p);
const double* orc_par_lu(orc_par* h, int p);

#ifdef __cplusplus
}
#endif
#endif
