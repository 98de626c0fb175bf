// w-fsa_b200/host/ldlt.hpp -- dense symmetric-indefinite factorisation (Bunch-Kaufman diagonal
// pivoting, P A P^T = L D L^T with 1x1 and 2x2 blocks in D).  Stands where the reference calls
// MKL DSS with MKL_DSS_SYMMETRIC + MKL_DSS_INDEFINITE (/root/reference/src/HessianLearner.cpp:
// 37,104,110,303 and src/Utils.cpp:296-355): solve, inertia and determinant of the augmented
// KKT matrix.  O(n^3/3); fine for n+k up to a few thousand.
#pragma once
#include <vector>

namespace wfsa {

class SymIndefinite {
public:
    // a: full symmetric n x n, row major (only the lower triangle is read)
    void Factor(int n, const std::vector<double>& a);
    void Solve(const double* rhs, double* sol) const;
    void Inertia(int& positive, int& negative, int& zero) const;
    // determinant = sign * exp(logabs); sign == 0 if singular
    void LogDet(double& logabs, int& sign) const;
private:
    int n_ = 0;
    std::vector<double> A_;   // column-major lower triangle holds L and D
    std::vector<int> piv_;    // >= 0: 1x1 pivot row; < 0: 2x2 block, partner row = -piv-1
    double& at(int i, int j) { return A_[(size_t)j * n_ + i]; }
    double at(int i, int j) const { return A_[(size_t)j * n_ + i]; }
};

}  // namespace wfsa
