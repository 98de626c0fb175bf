/* oracle/mkl_shim/mkl_spblas.h -- TEST INFRASTRUCTURE ONLY.
 * Scalar stand-in for the three inspector-executor Sparse BLAS entry points the
 * reference uses (src/Utils.cpp:397,408,413).  4-array CSR, zero based, NO COPY:
 * the handle aliases the caller's arrays (the reference relies on this,
 * src/QuasiNewtonLearner.cpp:26).  y = alpha*op(A)*x + beta*y; beta==0 overwrites. */
#pragma once
#include "mkl_types.h"
#include <cstdlib>

typedef enum { SPARSE_STATUS_SUCCESS = 0, SPARSE_STATUS_NOT_INITIALIZED = 1,
               SPARSE_STATUS_INVALID_VALUE = 3 } sparse_status_t;
typedef enum { SPARSE_INDEX_BASE_ZERO = 0, SPARSE_INDEX_BASE_ONE = 1 } sparse_index_base_t;
typedef enum { SPARSE_OPERATION_NON_TRANSPOSE = 10, SPARSE_OPERATION_TRANSPOSE = 11 } sparse_operation_t;
typedef enum { SPARSE_MATRIX_TYPE_GENERAL = 20 } sparse_matrix_type_t;
typedef enum { SPARSE_FILL_MODE_LOWER = 40, SPARSE_FILL_MODE_UPPER = 41 } sparse_fill_mode_t;
typedef enum { SPARSE_DIAG_NON_UNIT = 50, SPARSE_DIAG_UNIT = 51 } sparse_diag_type_t;

struct matrix_descr {
    sparse_matrix_type_t type;
    sparse_fill_mode_t mode;
    sparse_diag_type_t diag;
};

struct shim_sparse_matrix {
    MKL_INT rows, cols;
    MKL_INT *rs, *re, *ci;
    double *val;
};
typedef shim_sparse_matrix* sparse_matrix_t;

static inline sparse_status_t mkl_sparse_d_create_csr(sparse_matrix_t* A, sparse_index_base_t,
        MKL_INT rows, MKL_INT cols, MKL_INT* rows_start, MKL_INT* rows_end,
        MKL_INT* col_indx, double* values)
{
    shim_sparse_matrix* m = (shim_sparse_matrix*)std::malloc(sizeof(shim_sparse_matrix));
    if (!m) return SPARSE_STATUS_NOT_INITIALIZED;
    m->rows = rows; m->cols = cols; m->rs = rows_start; m->re = rows_end;
    m->ci = col_indx; m->val = values;
    *A = m;
    return SPARSE_STATUS_SUCCESS;
}

static inline sparse_status_t mkl_sparse_destroy(sparse_matrix_t A)
{
    std::free(A);
    return SPARSE_STATUS_SUCCESS;
}

static inline sparse_status_t mkl_sparse_d_mv(sparse_operation_t op, double alpha,
        const sparse_matrix_t A, matrix_descr, const double* x, double beta, double* y)
{
    if (!A) return SPARSE_STATUS_NOT_INITIALIZED;
    if (op == SPARSE_OPERATION_NON_TRANSPOSE) {
        for (MKL_INT i = 0; i < A->rows; ++i) {
            double s = 0.0;
            for (MKL_INT j = A->rs[i]; j < A->re[i]; ++j) s += A->val[j] * x[A->ci[j]];
            y[i] = (beta == 0.0) ? alpha * s : alpha * s + beta * y[i];
        }
    } else {
        for (MKL_INT c = 0; c < A->cols; ++c) y[c] = (beta == 0.0) ? 0.0 : beta * y[c];
        for (MKL_INT i = 0; i < A->rows; ++i) {
            const double xi = alpha * x[i];
            for (MKL_INT j = A->rs[i]; j < A->re[i]; ++j) y[A->ci[j]] += A->val[j] * xi;
        }
    }
    return SPARSE_STATUS_SUCCESS;
}
