/*
 * opm_oracle.c -- CPU restatement of the reference algorithms on the hot path.
 * TEST INFRASTRUCTURE ONLY (see opm_oracle.h).  Plain C99, serial, fp64, compiled with
 * -ffp-contract=off so that every operation is a single IEEE rounding in the order the
 * reference performs it.  Each function cites the reference file:line it follows
 * (paths relative to /root/reference).
 */
#include "opm_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXB 8

/* ------------------------------------------------------------------------------------------
 * dense block helpers == Dune::FieldMatrix / DenseMatrix members used by the reference
 * ---------------------------------------------------------------------------------------- */
/* y = A x          (DenseMatrix::mv: y[i] = 0; y[i] += A[i][j]*x[j]) */
static void blk_mv(int b, const double* A, const double* x, double* y)
{
    for (int i = 0; i < b; ++i) {
        double s = 0.0;
        for (int j = 0; j < b; ++j)
            s += A[i * b + j] * x[j];
        y[i] = s;
    }
}
/* y += A x         (DenseMatrix::umv) */
static void blk_umv(int b, const double* A, const double* x, double* y)
{
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < b; ++j)
            y[i] += A[i * b + j] * x[j];
}
/* y -= A x         (DenseMatrix::mmv) */
static void blk_mmv(int b, const double* A, const double* x, double* y)
{
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < b; ++j)
            y[i] -= A[i * b + j] * x[j];
}
/* y += alpha A x   (DenseMatrix::usmv: y[i] += alpha * A[i][j] * x[j]) */
static void blk_usmv(int b, double alpha, const double* A, const double* x, double* y)
{
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < b; ++j)
            y[i] += alpha * A[i * b + j] * x[j];
}
/* C = A B          (FieldMatrix operator*) ; C must not alias A or B */
static void blk_mm(int b, const double* A, const double* B, double* C)
{
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < b; ++j) {
            double s = 0.0;
            for (int k = 0; k < b; ++k)
                s += A[i * b + k] * B[k * b + j];
            C[i * b + j] = s;
        }
}

/* Dune DenseMatrix::invert(): LU with partial pivoting (fallback path of matrixblock.hh:205-224
 * and the path of every block size >= 5).  Singular iff a pivot is exactly zero. */
static int blk_invert_lu(int n, double* M)
{
    double A[MAXB * MAXB];
    int piv[MAXB];
    memcpy(A, M, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) {
        double pivmax = fabs(A[i * n + i]);
        int imax = i;
        for (int k = i + 1; k < n; ++k) {
            const double a = fabs(A[k * n + i]);
            if (a > pivmax) {
                pivmax = a;
                imax = k;
            }
        }
        if (imax != i)
            for (int j = 0; j < n; ++j) {
                const double t = A[i * n + j];
                A[i * n + j] = A[imax * n + j];
                A[imax * n + j] = t;
            }
        piv[i] = imax;
        if (!(pivmax != 0.0))
            return ORC_ERR_SINGULAR;
        for (int k = i + 1; k < n; ++k) {
            const double f = A[k * n + i] / A[i * n + i];
            A[k * n + i] = f;
            for (int j = i + 1; j < n; ++j)
                A[k * n + j] -= f * A[i * n + j];
        }
    }
    for (int i = 0; i < n * n; ++i)
        M[i] = 0.0;
    for (int i = 0; i < n; ++i)
        M[i * n + i] = 1.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j)
            for (int k = 0; k < n; ++k)
                M[i * n + k] -= A[i * n + j] * M[j * n + k];
    for (int i = n; i > 0;) {
        --i;
        for (int k = 0; k < n; ++k) {
            for (int j = i + 1; j < n; ++j)
                M[i * n + k] -= A[i * n + j] * M[j * n + k];
            M[i * n + k] /= A[i * n + i];
        }
    }
    for (int i = n; i > 0;) {
        --i;
        if (i != piv[i])
            for (int j = 0; j < n; ++j) {
                const double t = M[j * n + i];
                M[j * n + i] = M[j * n + piv[i]];
                M[j * n + piv[i]] = t;
            }
    }
    return ORC_OK;
}

/* signed 3x3 minor of a 4x4 matrix: rows r[0..2], cols c[0..2], expanded along its first column
 * into six triple products added left to right -- the term order of matrixblock.hh:72-190. */
static double minor3(const double* m, const int* r, const int* c, double sign)
{
#define E(i, j) m[r[i] * 4 + c[j]]
    double s = sign * E(0, 0) * E(1, 1) * E(2, 2);
    s -= sign * E(0, 0) * E(1, 2) * E(2, 1);
    s -= sign * E(1, 0) * E(0, 1) * E(2, 2);
    s += sign * E(1, 0) * E(0, 2) * E(2, 1);
    s += sign * E(2, 0) * E(0, 1) * E(1, 2);
    s -= sign * E(2, 0) * E(0, 2) * E(1, 1);
#undef E
    return s;
}

/* Opm::MatrixBlock::invert (matrixblock.hh:255-283):
 *   1..3 -> Dune::FMatrixHelp::invertMatrix closed forms (same expression tree as
 *           gpuistl/detail/deviceBlockOperations.hpp:37-114, "based on Dune cpu code"),
 *   4    -> adjugate / determinant with LU fallback when |det| < 1e-40 (matrixblock.hh:192-224),
 *   >=5  -> Dune's pivoted LU. */
int orc_invert_block(int b, double* a)
{
    if (b == 1) {
        a[0] = 1.0 / a[0];
        return ORC_OK;
    }
    if (b == 2) {
        const double det_1 = 1.0 / (a[0] * a[3] - a[1] * a[2]);
        const double a00 = a[0];
        a[0] = a[3] * det_1;
        a[1] = -a[1] * det_1;
        a[2] = -a[2] * det_1;
        a[3] = a00 * det_1;
        return ORC_OK;
    }
    if (b == 3) {
        const double m00 = a[0], m01 = a[1], m02 = a[2];
        const double m10 = a[3], m11 = a[4], m12 = a[5];
        const double m20 = a[6], m21 = a[7], m22 = a[8];
        const double p0011 = m00 * m11, p0012 = m00 * m12;
        const double p0110 = m01 * m10, p0210 = m02 * m10;
        const double p0120 = m01 * m20, p0220 = m02 * m20;
        const double rdet = 1.0
            / (p0011 * m22 - p0012 * m21 - p0110 * m22 + p0210 * m21 + p0120 * m12 - p0220 * m11);
        a[0] = (m11 * m22 - m12 * m21) * rdet;
        a[1] = -(m01 * m22 - m02 * m21) * rdet;
        a[2] = (m01 * m12 - m02 * m11) * rdet;
        a[3] = -(m10 * m22 - m12 * m20) * rdet;
        a[4] = (m00 * m22 - p0220) * rdet;
        a[5] = -(p0012 - p0210) * rdet;
        a[6] = (m10 * m21 - m11 * m20) * rdet;
        a[7] = -(m00 * m21 - p0120) * rdet;
        a[8] = (p0011 - p0110) * rdet;
        return ORC_OK;
    }
    if (b == 4) {
        double m[16], inv[16];
        memcpy(m, a, sizeof m);
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                /* inverse[i][j] = cofactor(j,i): delete row j and column i of m */
                int r[3], c[3], nr = 0, nc = 0;
                for (int k = 0; k < 4; ++k) {
                    if (k != j)
                        r[nr++] = k;
                    if (k != i)
                        c[nc++] = k;
                }
                inv[i * 4 + j] = minor3(m, r, c, ((i + j) & 1) ? -1.0 : 1.0);
            }
        const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
        if (fabs(det) < 1e-40) {
            const int rc = blk_invert_lu(4, a);
            if (rc != ORC_OK)
                for (int k = 0; k < 16; ++k)
                    a[k] = NAN;
            return rc;
        }
        const double rdet = 1.0 / det;
        for (int k = 0; k < 16; ++k)
            a[k] = inv[k] * rdet;
        return ORC_OK;
    }
    if (b > MAXB)
        return ORC_ERR_ARG;
    return blk_invert_lu(b, a);
}

/* index of block (i,j) in row i, or -1 (BCRSMatrix row find) */
static int find_col(const int* rowptr, const int* col, int i, int j)
{
    int lo = rowptr[i], hi = rowptr[i + 1] - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (col[mid] == j)
            return mid;
        if (col[mid] < j)
            lo = mid + 1;
        else
            hi = mid - 1;
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * Opm::getMatrixRowColoring, GraphColoring.hpp:246-307
 * ---------------------------------------------------------------------------------------- */
int orc_row_coloring(int n, const int* rowptr, const int* col, int type, int* color, int* level_rows,
                     int* level_ptr)
{
    int ncolors = 0;
    int* cnt = (int*)calloc((size_t)n + 1, sizeof(int));
    for (int i = 0; i < n; ++i)
        color[i] = 0;
    if (type == ORC_COLOR_SYMMETRIC || type == ORC_COLOR_LOWER) {
        for (int i = 0; i < n; ++i) {
            int k = rowptr[i];
            for (; k < rowptr[i + 1] && col[k] != i; ++k) {
                const int j = col[k];
                if (type == ORC_COLOR_SYMMETRIC && find_col(rowptr, col, j, i) < 0)
                    continue;
                if (color[j] + 1 > color[i])
                    color[i] = color[j] + 1;
            }
            if (k == rowptr[i + 1]) { /* the reference iterates until it meets the diagonal */
                free(cnt);
                return -ORC_ERR_DIAG_MISSING;
            }
            if (color[i] >= ncolors)
                ncolors = color[i] + 1;
            ++cnt[color[i]];
        }
    } else if (type == ORC_COLOR_UPPER) {
        for (int i = n - 1; i >= 0; --i) {
            const int kd = find_col(rowptr, col, i, i);
            if (kd < 0) {
                free(cnt);
                return -ORC_ERR_DIAG_MISSING;
            }
            for (int k = kd + 1; k < rowptr[i + 1]; ++k) {
                const int j = col[k];
                if (color[j] + 1 > color[i])
                    color[i] = color[j] + 1;
            }
            if (color[i] >= ncolors)
                ncolors = color[i] + 1;
            ++cnt[color[i]];
        }
    } else {
        free(cnt);
        return -ORC_ERR_ARG;
    }
    /* std::stable_sort of 0..n-1 by colour == counting sort in natural order (:300-303) */
    level_ptr[0] = 0;
    for (int c = 0; c < ncolors; ++c)
        level_ptr[c + 1] = level_ptr[c] + cnt[c];
    for (int c = 0; c < ncolors; ++c)
        cnt[c] = level_ptr[c];
    for (int i = 0; i < n; ++i)
        level_rows[cnt[color[i]]++] = i;
    free(cnt);
    return ncolors;
}

/* DILU.hpp:83-91 */
void orc_reorder_maps(int n, const int* level_rows, int* reordered_to_natural, int* natural_to_reordered)
{
    for (int k = 0; k < n; ++k) {
        reordered_to_natural[k] = level_rows[k];
        natural_to_reordered[level_rows[k]] = k;
    }
}

/* Opm::partitionCellsSimple, opm/simulators/flow/partitionCells.cpp:734-751 */
void orc_partition_simple(int num_cells, int num_domains, int* part)
{
    const int dom_sz = num_cells / num_domains;
    const int rem = num_cells % num_domains;
    int begin = 0;
    for (int d = 0; d < num_domains; ++d) {
        const int end = begin + dom_sz + (d < rem ? 1 : 0);
        for (int c = begin; c < end; ++c)
            part[c] = d;
        begin = end;
    }
}

/* ------------------------------------------------------------------------------------------
 * SpMV: GhostLastMatrixAdapter::apply / applyscaleadd, WellOperators.hpp:432-468
 * ---------------------------------------------------------------------------------------- */
void orc_spmv(int n, int b, const int* rowptr, const int* col, const double* val, int interior,
              const double* x, double* y)
{
    const int bb = b * b;
    for (int i = 0; i < interior; ++i) {
        double* yi = y + (size_t)i * b;
        for (int r = 0; r < b; ++r)
            yi[r] = 0.0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
            blk_umv(b, val + (size_t)k * bb, x + (size_t)col[k] * b, yi);
    }
    for (size_t s = (size_t)interior * b; s < (size_t)n * b; ++s)
        y[s] = 0.0;
}

void orc_spmv_scaleadd(int n, int b, const int* rowptr, const int* col, const double* val, int interior,
                       double alpha, const double* x, double* y)
{
    const int bb = b * b;
    for (int i = 0; i < interior; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
            blk_usmv(b, alpha, val + (size_t)k * bb, x + (size_t)col[k] * b, y + (size_t)i * b);
    for (size_t s = (size_t)interior * b; s < (size_t)n * b; ++s)
        y[s] = 0.0;
}

/* detail::makeOverlapRowsInvalid, ISTLSolver.cpp:56-75 (overlap rows == rows >= interior) */
void orc_make_overlap_rows_invalid(int n, int b, const int* rowptr, const int* col, double* val, int interior)
{
    const int bb = b * b;
    for (int i = interior; i < n; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            double* blk = val + (size_t)k * bb;
            for (int e = 0; e < bb; ++e)
                blk[e] = 0.0;
            if (col[k] == i)
                for (int r = 0; r < b; ++r)
                    blk[r * b + r] = 1.0;
        }
}

/* ------------------------------------------------------------------------------------------
 * DILU: MultithreadDILU::serialUpdate / serialApply, DILU.hpp:186-206, 253-304
 * ---------------------------------------------------------------------------------------- */
int orc_dilu_update(int n, int b, const int* rowptr, const int* col, const double* val, double* dinv)
{
    const int bb = b * b;
    double t1[MAXB * MAXB], t2[MAXB * MAXB], tmp[MAXB * MAXB];
    int rc_all = ORC_OK;
    for (int i = 0; i < n; ++i) {
        const int kd = find_col(rowptr, col, i, i);
        if (kd < 0)
            return ORC_ERR_DIAG_MISSING;
        memcpy(dinv + (size_t)i * bb, val + (size_t)kd * bb, sizeof(double) * bb);
    }
    for (int i = 0; i < n; ++i) {
        memcpy(tmp, dinv + (size_t)i * bb, sizeof(double) * bb);
        for (int k = rowptr[i]; k < rowptr[i + 1] && col[k] < i; ++k) {
            const int j = col[k];
            const int kji = find_col(rowptr, col, j, i);
            if (kji >= 0) {
                /* Dinv_temp -= (A_ij * Dinv_j) * A_ji       (:200) */
                blk_mm(b, val + (size_t)k * bb, dinv + (size_t)j * bb, t1);
                blk_mm(b, t1, val + (size_t)kji * bb, t2);
                for (int e = 0; e < bb; ++e)
                    tmp[e] -= t2[e];
            }
        }
        const int rc = orc_invert_block(b, tmp);
        if (rc != ORC_OK && rc_all == ORC_OK)
            rc_all = rc;
        memcpy(dinv + (size_t)i * bb, tmp, sizeof(double) * bb);
    }
    return rc_all;
}

void orc_dilu_apply(int n, int b, const int* rowptr, const int* col, const double* val, const double* dinv,
                    const double* d, double* v)
{
    const int bb = b * b;
    double rhs[MAXB];
    /* lower solve (D + L_A) y = d, y stored in v      (:267-281) */
    for (int i = 0; i < n; ++i) {
        for (int r = 0; r < b; ++r)
            rhs[r] = d[(size_t)i * b + r];
        for (int k = rowptr[i]; k < rowptr[i + 1] && col[k] < i; ++k)
            blk_mmv(b, val + (size_t)k * bb, v + (size_t)col[k] * b, rhs);
        blk_mv(b, dinv + (size_t)i * bb, rhs, v + (size_t)i * b);
    }
    /* upper solve (D + U_A) v = D y, rows and columns descending      (:286-302) */
    for (int i = n - 1; i >= 0; --i) {
        for (int r = 0; r < b; ++r)
            rhs[r] = 0.0;
        for (int k = rowptr[i + 1] - 1; k >= rowptr[i] && col[k] > i; --k)
            blk_umv(b, val + (size_t)k * bb, v + (size_t)col[k] * b, rhs);
        blk_mmv(b, dinv + (size_t)i * bb, rhs, v + (size_t)i * b);
    }
}

/* ------------------------------------------------------------------------------------------
 * block ILU0: detail::ghost_last_bilu0_decomposition, ParallelOverlappingILU0_impl.hpp:42-99
 * ---------------------------------------------------------------------------------------- */
int orc_ilu0_decompose(int n, int b, const int* rowptr, const int* col, double* lu, int interior)
{
    const int bb = b * b;
    double t[MAXB * MAXB];
    (void)n;
    for (int i = 0; i < interior; ++i) {
        const int endi = rowptr[i + 1];
        int ij = rowptr[i];
        for (; ij < endi && col[ij] < i; ++ij) {
            const int j = col[ij];
            const int jj = find_col(rowptr, col, j, j);
            if (jj < 0)
                return ORC_ERR_DIAG_MISSING;
            /* L_ij = A_ij * A_jj^{-1}  (rightmultiply; A_jj already holds its inverse) */
            blk_mm(b, lu + (size_t)ij * bb, lu + (size_t)jj * bb, t);
            memcpy(lu + (size_t)ij * bb, t, sizeof(double) * bb);
            /* A_ik -= L_ij * A_jk for matching k > j, two-pointer merge   (:66-86) */
            int jk = jj + 1, ik = ij + 1;
            const int endj = rowptr[j + 1];
            while (ik < endi && jk < endj) {
                if (col[ik] == col[jk]) {
                    blk_mm(b, lu + (size_t)ij * bb, lu + (size_t)jk * bb, t); /* B = L_ij * A_jk (leftmultiply) */
                    for (int e = 0; e < bb; ++e)
                        lu[(size_t)ik * bb + e] -= t[e];
                    ++ik;
                    ++jk;
                } else if (col[ik] < col[jk]) {
                    ++ik;
                } else {
                    ++jk;
                }
            }
        }
        if (ij == endi || col[ij] != i)
            return ORC_ERR_DIAG_MISSING;
        const int rc = orc_invert_block(b, lu + (size_t)ij * bb);
        if (rc != ORC_OK)
            return rc;
    }
    return ORC_OK;
}

/* ParallelOverlappingILU0::apply, ParallelOverlappingILU0_impl.hpp:383-411.  The reference first
 * splits the factor into lower_/upper_/inv_ CRS arrays (convertToCRS :102-191, upper stored in
 * reverse row and reverse column order); the traversal below visits the blocks in that same order
 * directly on the in-place factor. */
void orc_ilu0_apply(int n, int b, const int* rowptr, const int* col, const double* lu, int interior,
                    const double* d, double* v)
{
    const int bb = b * b;
    double rhs[MAXB];
    (void)n;
    for (int i = 0; i < interior; ++i) {
        for (int r = 0; r < b; ++r)
            rhs[r] = d[(size_t)i * b + r];
        for (int k = rowptr[i]; k < rowptr[i + 1] && col[k] < i; ++k)
            blk_mmv(b, lu + (size_t)k * bb, v + (size_t)col[k] * b, rhs);
        for (int r = 0; r < b; ++r)
            v[(size_t)i * b + r] = rhs[r]; /* L_ii = I */
    }
    for (int i = interior - 1; i >= 0; --i) {
        for (int r = 0; r < b; ++r)
            rhs[r] = v[(size_t)i * b + r];
        int k = rowptr[i + 1] - 1;
        for (; k >= rowptr[i] && col[k] > i; --k)
            blk_mmv(b, lu + (size_t)k * bb, v + (size_t)col[k] * b, rhs);
        blk_mv(b, lu + (size_t)k * bb, rhs, v + (size_t)i * b); /* k now at the (inverted) diagonal */
    }
}

/* ------------------------------------------------------------------------------------------
 * P subdomains in one process == P MPI ranks of Flow (block-Jacobi, owner-masked dots)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int n, interior;
    const int *rowptr, *col, *l2g;
    const double* val;
    double* dinv;
    double* lu;
    /* standard wells kept outside the matrix (matrix-add-well-contributions=false); borrowed arrays */
    int nw, dw;
    const int *wptr, *wcells;
    const double *wB, *wC, *wDinv;
    double* wscratch; /* scaleAddRes_ of WellModelAsLinearOperator */
} orc_sub;

struct orc_par {
    int nsub, b;
    long nglobal;
    orc_sub* sub;
    int kind;
    double w;
    double* gscratch;
};

orc_par* orc_par_create(int nsub, int b, long nglobal)
{
    orc_par* h = (orc_par*)calloc(1, sizeof(orc_par));
    h->nsub = nsub;
    h->b = b;
    h->nglobal = nglobal;
    h->sub = (orc_sub*)calloc((size_t)nsub, sizeof(orc_sub));
    h->kind = ORC_PREC_NONE;
    h->w = 1.0;
    h->gscratch = (double*)calloc((size_t)nglobal * b + 1, sizeof(double));
    return h;
}

void orc_par_destroy(orc_par* h)
{
    if (!h)
        return;
    for (int p = 0; p < h->nsub; ++p) {
        free(h->sub[p].dinv);
        free(h->sub[p].lu);
        free(h->sub[p].wscratch);
    }
    free(h->sub);
    free(h->gscratch);
    free(h);
}

int orc_par_set_sub(orc_par* h, int p, int n, int interior, const int* rowptr, const int* col,
                    const double* val, const int* l2g)
{
    if (p < 0 || p >= h->nsub || interior > n)
        return ORC_ERR_ARG;
    orc_sub* s = &h->sub[p];
    s->n = n;
    s->interior = interior;
    s->rowptr = rowptr;
    s->col = col;
    s->val = val;
    s->l2g = l2g;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * Standard wells as a linear operator: y -= C^T (D^-1 (B x)), well after well
 * (WellOperators.hpp:84-91 apply, :144-164 applySingleWell; StandardWellEquations.cpp:132-148:
 * Bx = B.mv(x), invDBx = invD.mv(Bx), Ax -= C.mmtv(invDBx)).  Perforation p of well w touches
 * cell cells[p]; B and C hold one dw x b block (row-major) per perforation, Dinv one dw x dw
 * block per well.  The running sums follow Dune's mv / mmtv loop order.
 * ---------------------------------------------------------------------------------------- */
void orc_well_apply(int nw, int dw, int b, const int* wptr, const int* cells, const double* B, const double* C,
                    const double* Dinv, const double* x, double* y)
{
    double z1[16], z2[16];
    for (int w = 0; w < nw; ++w) {
        for (int r = 0; r < dw; ++r)
            z1[r] = 0.0;
        for (int p = wptr[w]; p < wptr[w + 1]; ++p) { /* BCRSMatrix::mv: block.umv per entry */
            const double* Bp = B + (size_t)p * dw * b;
            const double* xc = x + (size_t)cells[p] * b;
            for (int r = 0; r < dw; ++r)
                for (int c = 0; c < b; ++c)
                    z1[r] += Bp[r * b + c] * xc[c];
        }
        const double* Dw = Dinv + (size_t)w * dw * dw;
        for (int r = 0; r < dw; ++r) {
            z2[r] = 0.0;
            for (int c = 0; c < dw; ++c)
                z2[r] += Dw[r * dw + c] * z1[c];
        }
        for (int p = wptr[w]; p < wptr[w + 1]; ++p) { /* BCRSMatrix::mmtv: y[j] -= A[i][j]^T x[i] */
            const double* Cp = C + (size_t)p * dw * b;
            double* yc = y + (size_t)cells[p] * b;
            for (int r = 0; r < dw; ++r)
                for (int c = 0; c < b; ++c)
                    yc[c] -= Cp[r * b + c] * z2[r];
        }
    }
}

int orc_par_set_wells(orc_par* h, int p, int nw, int dw, const int* wptr, const int* cells, const double* B,
                      const double* C, const double* Dinv)
{
    if (p < 0 || p >= h->nsub || nw < 0 || dw < 0 || dw > 16)
        return ORC_ERR_ARG;
    orc_sub* s = &h->sub[p];
    s->nw = nw;
    s->dw = dw;
    s->wptr = wptr;
    s->wcells = cells;
    s->wB = B;
    s->wC = C;
    s->wDinv = Dinv;
    return ORC_OK;
}

/* WellModelMatrixAdapter / WellModelGhostLastMatrixAdapter::apply tail (WellOperators.hpp:244-250, 336-339):
 * wellOper.apply(x, y) then ghostLastProject(y) */
static void sub_well_apply(const orc_sub* s, int b, const double* x, double* y)
{
    if (s->nw <= 0)
        return;
    orc_well_apply(s->nw, s->dw, b, s->wptr, s->wcells, s->wB, s->wC, s->wDinv, x, y);
    for (size_t k = (size_t)s->interior * b; k < (size_t)s->n * b; ++k)
        y[k] = 0.0;
}
/* WellModelAsLinearOperator::applyscaleadd (WellOperators.hpp:93-109): scaleAddRes = 0; apply(x, scaleAddRes);
 * y.axpy(alpha, scaleAddRes); then ghostLastProject(y) (:353-355) */
static void sub_well_applyscaleadd(orc_sub* s, int b, double alpha, const double* x, double* y)
{
    if (s->nw <= 0)
        return;
    const size_t len = (size_t)s->n * b;
    if (!s->wscratch)
        s->wscratch = (double*)malloc(len * sizeof(double));
    for (size_t k = 0; k < len; ++k)
        s->wscratch[k] = 0.0;
    orc_well_apply(s->nw, s->dw, b, s->wptr, s->wcells, s->wB, s->wC, s->wDinv, x, s->wscratch);
    for (size_t k = 0; k < len; ++k)
        y[k] += alpha * s->wscratch[k];
    for (size_t k = (size_t)s->interior * b; k < len; ++k)
        y[k] = 0.0;
}

const double* orc_par_dinv(orc_par* h, int p) { return h->sub[p].dinv; }
const double* orc_par_lu(orc_par* h, int p) { return h->sub[p].lu; }

/* PreconditionerWithUpdate::update(): MultithreadDILU::update (DILU.hpp:110-118) or
 * ParallelOverlappingILU0::update (ParallelOverlappingILU0_impl.hpp:431-610, ILU(0), no reordering);
 * a failure on any rank fails all ranks (comm.min vote, :598-606). */
int orc_par_prec_update(orc_par* h, int kind, double w)
{
    const int bb = h->b * h->b;
    int rc_all = ORC_OK;
    h->kind = kind;
    h->w = w;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < h->nsub; ++p) {
        orc_sub* s = &h->sub[p];
        const size_t nnzb = (size_t)s->rowptr[s->n];
        int rc = ORC_OK;
        if (kind == ORC_PREC_DILU) {
            if (!s->dinv)
                s->dinv = (double*)malloc(sizeof(double) * (size_t)s->n * bb + 8);
            rc = orc_dilu_update(s->n, h->b, s->rowptr, s->col, s->val, s->dinv);
        } else if (kind == ORC_PREC_ILU0) {
            if (!s->lu)
                s->lu = (double*)malloc(sizeof(double) * nnzb * bb + 8);
            memcpy(s->lu, s->val, sizeof(double) * nnzb * bb);
            rc = orc_ilu0_decompose(s->n, h->b, s->rowptr, s->col, s->lu, s->interior);
        } else if (kind != ORC_PREC_NONE) {
            rc = ORC_ERR_ARG;
        }
        if (rc != ORC_OK) {
#pragma omp critical
            if (rc_all == ORC_OK)
                rc_all = rc;
        }
    }
    return rc_all;
}

/* OwnerOverlapCopyCommunication::copyOwnerToAll(v,v): every copy (ghost) entry receives the value
 * held by its owner.  gpuistl/GpuOwnerOverlapCopy.hpp, gpuistl/GpuAwareMPISender.hpp:164-222 */
void orc_par_copy_owner_to_all(orc_par* h, double** v)
{
    const int b = h->b;
    if (h->nsub == 1 && !h->sub[0].l2g)
        return;
    for (int p = 0; p < h->nsub; ++p) {
        const orc_sub* s = &h->sub[p];
        for (int i = 0; i < s->interior; ++i)
            for (int r = 0; r < b; ++r)
                h->gscratch[(size_t)s->l2g[i] * b + r] = v[p][(size_t)i * b + r];
    }
    for (int p = 0; p < h->nsub; ++p) {
        const orc_sub* s = &h->sub[p];
        for (int i = s->interior; i < s->n; ++i)
            for (int r = 0; r < b; ++r)
                v[p][(size_t)i * b + r] = h->gscratch[(size_t)s->l2g[i] * b + r];
    }
}

/* Dune::BlockPreconditioner::apply (OwningBlockPreconditioner.hpp:31-92;
 * gpuistl/GpuBlockPreconditioner.hpp:65-81): local apply, then copyOwnerToAll.
 * ParallelOverlappingILU0 does its own copyOwnerToAll and then the relaxation (:413-417). */
int orc_par_prec_apply(orc_par* h, double** v, double** d)
{
    const int b = h->b;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < h->nsub; ++p) {
        const orc_sub* s = &h->sub[p];
        if (h->kind == ORC_PREC_DILU)
            orc_dilu_apply(s->n, b, s->rowptr, s->col, s->val, s->dinv, d[p], v[p]);
        else if (h->kind == ORC_PREC_ILU0)
            orc_ilu0_apply(s->n, b, s->rowptr, s->col, s->lu, s->interior, d[p], v[p]);
        else /* "nothing": tests/test_preconditionerfactory.cpp:183-198 NothingPreconditioner, v = d */
            memcpy(v[p], d[p], sizeof(double) * (size_t)s->n * b);
    }
    if (h->kind != ORC_PREC_NONE)
        orc_par_copy_owner_to_all(h, v);
    if (h->kind == ORC_PREC_ILU0 && fabs(h->w - 1.0) > 1e-15)
#pragma omp parallel for schedule(static)
        for (int p = 0; p < h->nsub; ++p)
            for (size_t s = 0; s < (size_t)h->sub[p].n * b; ++s)
                v[p][s] *= h->w;
    return ORC_OK;
}

/* SeqScalarProduct / OverlappingSchwarzScalarProduct::dot: owner entries only, summed over ranks
 * (gpuistl/GpuSender.hpp:89-95).  BlockVector::dot adds one block dot at a time. */
double orc_par_dot(orc_par* h, double** x, double** y)
{
    const int b = h->b;
    double total = 0.0;
    double* part = (double*)malloc(sizeof(double) * (size_t)h->nsub);
#pragma omp parallel for schedule(static)
    for (int p = 0; p < h->nsub; ++p) {
        double sum = 0.0;
        for (int i = 0; i < h->sub[p].interior; ++i) {
            double blk = 0.0;
            for (int r = 0; r < b; ++r)
                blk += x[p][(size_t)i * b + r] * y[p][(size_t)i * b + r];
            sum += blk;
        }
        part[p] = sum;
    }
    for (int p = 0; p < h->nsub; ++p) /* MPI_Allreduce: rank order */
        total += part[p];
    free(part);
    return total;
}

static double par_norm(orc_par* h, double** x) { return sqrt(orc_par_dot(h, x, x)); }

static double** vec_alloc(orc_par* h)
{
    double** v = (double**)malloc(sizeof(double*) * (size_t)h->nsub);
    for (int p = 0; p < h->nsub; ++p)
        v[p] = (double*)calloc((size_t)h->sub[p].n * h->b + 1, sizeof(double));
    return v;
}
static void vec_free(orc_par* h, double** v)
{
    for (int p = 0; p < h->nsub; ++p)
        free(v[p]);
    free(v);
}
static void vec_copy(orc_par* h, double** dst, double** src)
{
#pragma omp parallel for schedule(static)
    for (int p = 0; p < h->nsub; ++p)
        memcpy(dst[p], src[p], sizeof(double) * (size_t)h->sub[p].n * h->b);
}
static void vec_zero(orc_par* h, double** v)
{
    for (int p = 0; p < h->nsub; ++p)
        memset(v[p], 0, sizeof(double) * (size_t)h->sub[p].n * h->b);
}
/* y += a x */
static void vec_axpy(orc_par* h, double a, double** x, double** y)
{
#pragma omp parallel for schedule(static)
    for (int p = 0; p < h->nsub; ++p)
        for (size_t s = 0; s < (size_t)h->sub[p].n * h->b; ++s)
            y[p][s] += a * x[p][s];
}

/* op.apply(x,y) (y = A x) with the RepeatingOperator generalisation */
static void op_apply(orc_par* h, int repeats, double** x, double** y, double** t1)
{
    const int b = h->b;
    if (repeats <= 1) {
#pragma omp parallel for schedule(static)
        for (int p = 0; p < h->nsub; ++p) {
            const orc_sub* s = &h->sub[p];
            orc_spmv(s->n, b, s->rowptr, s->col, s->val, s->interior, x[p], y[p]);
            sub_well_apply(s, b, x[p], y[p]);
        }
        return;
    }
    vec_copy(h, t1, x);
    for (int rr = 0; rr < repeats; ++rr) {
        for (int p = 0; p < h->nsub; ++p) {
            const orc_sub* s = &h->sub[p];
            orc_spmv(s->n, b, s->rowptr, s->col, s->val, s->interior, t1[p], y[p]);
            sub_well_apply(s, b, t1[p], y[p]);
        }
        vec_copy(h, t1, y);
    }
}
/* op.applyscaleadd(alpha,x,y) (y += alpha A x) */
static void op_applyscaleadd(orc_par* h, int repeats, double alpha, double** x, double** y, double** t1,
                             double** t2)
{
    const int b = h->b;
    if (repeats <= 1) {
#pragma omp parallel for schedule(static)
        for (int p = 0; p < h->nsub; ++p) {
            orc_sub* s = &h->sub[p];
            orc_spmv_scaleadd(s->n, b, s->rowptr, s->col, s->val, s->interior, alpha, x[p], y[p]);
            sub_well_applyscaleadd(s, b, alpha, x[p], y[p]);
        }
        return;
    }
    op_apply(h, repeats, x, t2, t1);
    for (int p = 0; p < h->nsub; ++p)
        for (size_t s = 0; s < (size_t)h->sub[p].n * b; ++s)
            y[p][s] += t2[p][s] * alpha;
}

/* IterativeSolver::Iteration::step (dune-istl solver.hh): returns 1 when converged, <0 on NaN */
static int iteration_step(double it, double def, double reduction, orc_result* res, double* hist, int* nhist)
{
    if (!isfinite(def))
        return -1;
    if (it == 0.0)
        res->norm0 = def;
    res->norm = def;
    res->it = it;
    if (hist)
        hist[(*nhist)++] = def;
    res->converged = (def < res->norm0 * reduction || def < 1e-30);
    return res->converged;
}

/* Dune::BiCGSTABSolver<X>::apply(x, b, res) -- dune-istl (>= 2.9) solvers.hh, not vendored in
 * /root/reference; constructed at FlexibleSolver_impl.hpp:214-220.  In-tree restatements that agree
 * with the step list below: gpubridge/cuda/cusparseSolverBackend.cu:97-308 (same half-iteration
 * counter, same norm < tol*norm0 test).  SURVEY.md section 8c lists the conventions that cannot be
 * verified against the Dune source here. */
int orc_par_bicgstab(orc_par* h, double** x, double** b, double reduction, int maxiter, int op_repeats,
                     orc_result* res, double* hist, int* nhist)
{
    const double EPSILON = 1e-80;
    int rc = ORC_OK, nh = 0, st;
    double it = 0.0, rho = 1, rho_new, alpha = 1, beta, hh, omega = 1, norm;
    double** r = b;
    double** p = vec_alloc(h);
    double** v = vec_alloc(h);
    double** t = vec_alloc(h);
    double** y = vec_alloc(h);
    double** rt = vec_alloc(h);
    double** w1 = vec_alloc(h);
    double** w2 = vec_alloc(h);
    memset(res, 0, sizeof *res);
    if (!nhist)
        nhist = &nh;
    *nhist = 0;

    /* _prec->pre(x,r): BlockPreconditioner makes x consistent; serial: no-op */
    orc_par_copy_owner_to_all(h, x);
    /* r = b - A x */
    op_applyscaleadd(h, op_repeats, -1.0, x, r, w1, w2);
    vec_copy(h, rt, r);
    norm = par_norm(h, r);
    st = iteration_step(0.0, norm, reduction, res, hist, nhist);
    if (st < 0) {
        rc = ORC_ERR_NAN;
        goto done;
    }
    if (st)
        goto done;
    vec_zero(h, p);
    vec_zero(h, v);

    for (it = 0.5; it < maxiter; it += .5) {
        rho_new = orc_par_dot(h, rt, r);
        if (fabs(rho) <= EPSILON || fabs(omega) <= EPSILON) {
            rc = ORC_ERR_BREAKDOWN;
            break;
        }
        if (it < 1) {
            vec_copy(h, p, r);
        } else {
            beta = (norm == 0.0) ? 0.0 : (rho_new / rho) * (alpha / omega);
            vec_axpy(h, -omega, v, p); /* p = r + beta (p - omega v) */
#pragma omp parallel for schedule(static)
            for (int q = 0; q < h->nsub; ++q)
                for (size_t s = 0; s < (size_t)h->sub[q].n * h->b; ++s) {
                    p[q][s] *= beta;
                    p[q][s] += r[q][s];
                }
        }
        /* y = W^-1 p ; v = A y */
        vec_zero(h, y);
        orc_par_prec_apply(h, y, p);
        op_apply(h, op_repeats, y, v, w1);
        hh = orc_par_dot(h, rt, v);
        if (fabs(hh) < EPSILON) {
            rc = ORC_ERR_BREAKDOWN;
            break;
        }
        alpha = (norm == 0.0) ? 0.0 : rho_new / hh;
        vec_axpy(h, alpha, y, x);
        vec_axpy(h, -alpha, v, r);
        norm = par_norm(h, r);
        st = iteration_step(it, norm, reduction, res, hist, nhist);
        if (st < 0) {
            rc = ORC_ERR_NAN;
            break;
        }
        if (st)
            break;

        it += .5;
        /* y = W^-1 r ; t = A y */
        vec_zero(h, y);
        orc_par_prec_apply(h, y, r);
        op_apply(h, op_repeats, y, t, w1);
        hh = orc_par_dot(h, t, t);
        omega = (norm == 0.0) ? 0.0 : orc_par_dot(h, t, r) / hh;
        vec_axpy(h, omega, y, x);
        vec_axpy(h, -omega, t, r);
        rho = rho_new;
        norm = par_norm(h, r);
        st = iteration_step(it, norm, reduction, res, hist, nhist);
        if (st < 0) {
            rc = ORC_ERR_NAN;
            break;
        }
        if (st)
            break;
    }
done:
    /* Iteration::~Iteration -> _finalize */
    res->iterations = (int)res->it;
    res->reduction = res->norm0 > 0 ? res->norm / res->norm0 : 0.0;
    res->conv_rate = res->it > 0 ? pow(res->reduction, 1.0 / res->it) : 0.0;
    vec_free(h, p);
    vec_free(h, v);
    vec_free(h, t);
    vec_free(h, y);
    vec_free(h, rt);
    vec_free(h, w1);
    vec_free(h, w2);
    return rc;
}
