/* oracle/mkl_shim/mkl.h -- TEST INFRASTRUCTURE ONLY.
 * Scalar stand-in for the Intel MKL symbols the reference uses (VML, CBLAS L1,
 * service functions; full call-site list in SURVEY.md section 8c).  Every loop is
 * the textbook definition in double precision, so results differ from real MKL
 * only in summation order / <1 ulp transcendental rounding. */
#pragma once
#include "mkl_types.h"
#include "mkl_spblas.h"
#include "mkl_dss.h"
#include <cmath>
#include <cstddef>

struct MKLVersion {
    int MajorVersion, MinorVersion, UpdateVersion;
    const char* ProductStatus;
    const char* Build;
    const char* Processor;
    const char* Platform;
};

static inline void mkl_get_version(MKLVersion* v)
{
    v->MajorVersion = 0; v->MinorVersion = 0; v->UpdateVersion = 0;
    v->ProductStatus = "stand-in"; v->Build = "oracle/mkl_shim";
    v->Processor = "scalar C++"; v->Platform = "any";
}
static inline int mkl_get_max_threads() { return 1; }
static inline void mkl_set_num_threads(int) {}

static inline void vdExp(MKL_INT n, const double* a, double* y) { for (MKL_INT i = 0; i < n; ++i) y[i] = std::exp(a[i]); }
static inline void vdLn(MKL_INT n, const double* a, double* y) { for (MKL_INT i = 0; i < n; ++i) y[i] = std::log(a[i]); }
static inline void vdDiv(MKL_INT n, const double* a, const double* b, double* y) { for (MKL_INT i = 0; i < n; ++i) y[i] = a[i] / b[i]; }
static inline void vdMul(MKL_INT n, const double* a, const double* b, double* y) { for (MKL_INT i = 0; i < n; ++i) y[i] = a[i] * b[i]; }
static inline void vdAdd(MKL_INT n, const double* a, const double* b, double* y) { for (MKL_INT i = 0; i < n; ++i) y[i] = a[i] + b[i]; }

static inline double cblas_ddot(MKL_INT n, const double* x, MKL_INT incx, const double* y, MKL_INT incy)
{
    double s = 0.0;
    for (MKL_INT i = 0; i < n; ++i) s += x[(size_t)i * incx] * y[(size_t)i * incy];
    return s;
}
static inline void cblas_daxpy(MKL_INT n, double a, const double* x, MKL_INT incx, double* y, MKL_INT incy)
{
    for (MKL_INT i = 0; i < n; ++i) y[(size_t)i * incy] += a * x[(size_t)i * incx];
}
static inline void cblas_daxpby(MKL_INT n, double a, const double* x, MKL_INT incx, double b, double* y, MKL_INT incy)
{
    for (MKL_INT i = 0; i < n; ++i) y[(size_t)i * incy] = a * x[(size_t)i * incx] + b * y[(size_t)i * incy];
}
static inline void cblas_dscal(MKL_INT n, double a, double* x, MKL_INT incx)
{
    for (MKL_INT i = 0; i < n; ++i) x[(size_t)i * incx] *= a;
}
static inline void cblas_dcopy(MKL_INT n, const double* x, MKL_INT incx, double* y, MKL_INT incy)
{
    for (MKL_INT i = 0; i < n; ++i) y[(size_t)i * incy] = x[(size_t)i * incx];
}
static inline size_t cblas_idamax(MKL_INT n, const double* x, MKL_INT incx)
{
    size_t best = 0;
    for (MKL_INT i = 1; i < n; ++i)
        if (std::fabs(x[(size_t)i * incx]) > std::fabs(x[best * incx])) best = (size_t)i;
    return best;
}
static inline size_t cblas_idamin(MKL_INT n, const double* x, MKL_INT incx)
{
    size_t best = 0;
    for (MKL_INT i = 1; i < n; ++i)
        if (std::fabs(x[(size_t)i * incx]) < std::fabs(x[best * incx])) best = (size_t)i;
    return best;
}
