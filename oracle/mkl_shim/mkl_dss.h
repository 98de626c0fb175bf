/* oracle/mkl_shim/mkl_dss.h -- TEST INFRASTRUCTURE ONLY.
 * Dense stand-in for the MKL DSS calls the reference makes
 * (src/HessianLearner.cpp:37,52,104,110,303; src/Utils.cpp:322-372).
 * Structure = upper triangle of a symmetric matrix in CSR (every diagonal present).
 * factor: dense copy; solve: Gaussian elimination with partial pivoting;
 * "Inertia": eigenvalue signs by cyclic Jacobi; "Determinant": {pow10, mantissa}.
 * Adequate at fixture sizes (n+k <= a few hundred); O(n^3). */
#pragma once
#include "mkl_types.h"
#include <vector>
#include <cmath>
#include <cstring>

typedef void* _MKL_DSS_HANDLE_t;

#define MKL_DSS_DEFAULTS 0
#define MKL_DSS_SUCCESS 0
#define MKL_DSS_ZERO_BASED_INDEXING 131072
#define MKL_DSS_MSG_LVL_SUCCESS -2147483647
#define MKL_DSS_MSG_LVL_DEBUG -2147483646
#define MKL_DSS_MSG_LVL_INFO -2147483645
#define MKL_DSS_MSG_LVL_WARNING -2147483644
#define MKL_DSS_MSG_LVL_ERROR -2147483643
#define MKL_DSS_MSG_LVL_FATAL -2147483642
#define MKL_DSS_TERM_LVL_SUCCESS 1073741832
#define MKL_DSS_TERM_LVL_ERROR 1073741864
#define MKL_DSS_SYMMETRIC 536870976
#define MKL_DSS_AUTO_ORDER 268435520
#define MKL_DSS_MY_ORDER 268435584
#define MKL_DSS_GET_ORDER 268435712
#define MKL_DSS_METIS_ORDER 268435840
#define MKL_DSS_METIS_OPENMP_ORDER 268435968
#define MKL_DSS_POSITIVE_DEFINITE 134217792
#define MKL_DSS_INDEFINITE 134217856
#define MKL_DSS_REFINEMENT_OFF 4096
#define MKL_DSS_REFINEMENT_ON 8192
#define MKL_DSS_FAILURE -5

struct shim_dss {
    MKL_INT n;
    std::vector<MKL_INT> row, col;
    std::vector<double> A;   /* dense n*n, row major, symmetric */
};

#define dss_create(handle, opt) shim_dss_create(&(handle), &(opt))

static inline MKL_INT shim_dss_create(_MKL_DSS_HANDLE_t* h, const MKL_INT*)
{
    *h = new shim_dss();
    return MKL_DSS_SUCCESS;
}

static inline MKL_INT dss_delete(_MKL_DSS_HANDLE_t& h, const MKL_INT&)
{
    delete (shim_dss*)h;
    h = nullptr;
    return MKL_DSS_SUCCESS;
}

static inline MKL_INT dss_define_structure(_MKL_DSS_HANDLE_t& h, const MKL_INT&,
        const MKL_INT* rowIndex, const MKL_INT& nRows, const MKL_INT&,
        const MKL_INT* columns, const MKL_INT& nNonZeros)
{
    shim_dss* s = (shim_dss*)h;
    s->n = nRows;
    s->row.assign(rowIndex, rowIndex + nRows + 1);
    s->col.assign(columns, columns + nNonZeros);
    return MKL_DSS_SUCCESS;
}

static inline MKL_INT dss_reorder(_MKL_DSS_HANDLE_t&, const MKL_INT&, const MKL_INT*)
{
    return MKL_DSS_SUCCESS;   /* ordering is irrelevant to a dense solve */
}

static inline MKL_INT dss_factor_real(_MKL_DSS_HANDLE_t& h, const MKL_INT&, const void* values)
{
    shim_dss* s = (shim_dss*)h;
    const double* v = (const double*)values;
    const MKL_INT n = s->n;
    s->A.assign((size_t)n * n, 0.0);
    for (MKL_INT i = 0; i < n; ++i)
        for (MKL_INT j = s->row[i]; j < s->row[i + 1]; ++j) {
            s->A[(size_t)i * n + s->col[j]] = v[j];
            s->A[(size_t)s->col[j] * n + i] = v[j];
        }
    return MKL_DSS_SUCCESS;
}

static inline MKL_INT dss_solve_real(_MKL_DSS_HANDLE_t& h, const MKL_INT&,
        const void* rhsValues, const MKL_INT& nRhs, void* solValues)
{
    shim_dss* s = (shim_dss*)h;
    const MKL_INT n = s->n;
    const double* b = (const double*)rhsValues;
    double* x = (double*)solValues;
    for (MKL_INT r = 0; r < nRhs; ++r) {
        std::vector<double> M(s->A);
        std::vector<double> y(b + (size_t)r * n, b + (size_t)(r + 1) * n);
        for (MKL_INT c = 0; c < n; ++c) {
            MKL_INT piv = c;
            for (MKL_INT i = c + 1; i < n; ++i)
                if (std::fabs(M[(size_t)i * n + c]) > std::fabs(M[(size_t)piv * n + c])) piv = i;
            if (piv != c) {
                for (MKL_INT j = 0; j < n; ++j) std::swap(M[(size_t)c * n + j], M[(size_t)piv * n + j]);
                std::swap(y[c], y[piv]);
            }
            const double d = M[(size_t)c * n + c];
            for (MKL_INT i = c + 1; i < n; ++i) {
                const double f = M[(size_t)i * n + c] / d;
                if (f != 0.0) {
                    for (MKL_INT j = c; j < n; ++j) M[(size_t)i * n + j] -= f * M[(size_t)c * n + j];
                    y[i] -= f * y[c];
                }
            }
        }
        for (MKL_INT i = n - 1; i >= 0; --i) {
            double t = y[i];
            for (MKL_INT j = i + 1; j < n; ++j) t -= M[(size_t)i * n + j] * x[(size_t)r * n + j];
            x[(size_t)r * n + i] = t / M[(size_t)i * n + i];
        }
    }
    return MKL_DSS_SUCCESS;
}

static inline void shim_dss_eigenvalues(const shim_dss* s, std::vector<double>& ev)
{
    const MKL_INT n = s->n;
    std::vector<double> M(s->A);
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0;
        for (MKL_INT i = 0; i < n; ++i)
            for (MKL_INT j = i + 1; j < n; ++j) off += M[(size_t)i * n + j] * M[(size_t)i * n + j];
        if (off < 1e-300) break;
        for (MKL_INT p = 0; p < n; ++p)
            for (MKL_INT q = p + 1; q < n; ++q) {
                const double apq = M[(size_t)p * n + q];
                if (apq == 0.0) continue;
                const double app = M[(size_t)p * n + p], aqq = M[(size_t)q * n + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (MKL_INT k = 0; k < n; ++k) {
                    const double akp = M[(size_t)k * n + p], akq = M[(size_t)k * n + q];
                    M[(size_t)k * n + p] = c * akp - sn * akq;
                    M[(size_t)k * n + q] = sn * akp + c * akq;
                }
                for (MKL_INT k = 0; k < n; ++k) {
                    const double apk = M[(size_t)p * n + k], aqk = M[(size_t)q * n + k];
                    M[(size_t)p * n + k] = c * apk - sn * aqk;
                    M[(size_t)q * n + k] = sn * apk + c * aqk;
                }
            }
    }
    ev.resize(n);
    for (MKL_INT i = 0; i < n; ++i) ev[i] = M[(size_t)i * n + i];
}

static inline MKL_INT dss_statistics(_MKL_DSS_HANDLE_t& h, const MKL_INT&, const char* what, double* ret)
{
    shim_dss* s = (shim_dss*)h;
    std::vector<double> ev;
    shim_dss_eigenvalues(s, ev);
    double scale = 0.0;
    for (double e : ev) scale = std::fmax(scale, std::fabs(e));
    const double tiny = scale * 1e-13;
    if (std::strcmp(what, "Inertia") == 0) {
        ret[0] = ret[1] = ret[2] = 0.0;
        for (double e : ev) {
            if (e > tiny) ret[0] += 1; else if (e < -tiny) ret[1] += 1; else ret[2] += 1;
        }
        return MKL_DSS_SUCCESS;
    }
    if (std::strcmp(what, "Determinant") == 0) {
        double mant = 1.0, p10 = 0.0;
        for (double e : ev) {
            mant *= e;
            if (mant == 0.0) break;
            const double l = std::floor(std::log10(std::fabs(mant)));
            mant /= std::pow(10.0, l);
            p10 += l;
        }
        ret[0] = p10; ret[1] = mant;
        return MKL_DSS_SUCCESS;
    }
    return MKL_DSS_FAILURE;
}
