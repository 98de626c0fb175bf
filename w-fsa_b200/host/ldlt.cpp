// w-fsa_b200/host/ldlt.cpp -- see ldlt.hpp.  Unblocked Bunch-Kaufman on the lower triangle.
#include "ldlt.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <functional>
#include <thread>
#include <utility>

namespace wfsa {

namespace {
// The trailing update of a pivot step touches every column right of the pivot and nothing else of the matrix: columns are
// dealt out cyclically to a team of host threads that lives for one factorisation and meets at a spinning barrier twice per
// step (the pivot search and the row/column interchanges stay with the calling thread).  Every element sees the same
// operations in the same order as in the serial loop: the factor is bit-identical for any team size.
class ColumnTeam {
public:
    explicit ColumnTeam(int threads) : n_(threads)
    {
        for (int t = 1; t < n_; ++t) pool_.emplace_back([this, t] { worker(t); });
    }
    ~ColumnTeam()
    {
        stop_ = true;
        gen_.fetch_add(1, std::memory_order_release);
        for (auto& th : pool_) th.join();
    }
    int size() const { return n_; }
    // runs job(t, size) on every member (t = 0 on the caller) and returns when all are done
    void run(const std::function<void(int, int)>& job)
    {
        if (n_ == 1) { job(0, 1); return; }
        job_ = &job;
        done_.store(0, std::memory_order_relaxed);
        gen_.fetch_add(1, std::memory_order_release);
        job(0, n_);
        while (done_.load(std::memory_order_acquire) != n_ - 1) { }
    }
private:
    void worker(int t)
    {
        unsigned seen = 0;
        for (;;) {
            unsigned g;
            while ((g = gen_.load(std::memory_order_acquire)) == seen) { }
            seen = g;
            if (stop_) return;
            (*job_)(t, n_);
            done_.fetch_add(1, std::memory_order_release);
        }
    }
    int n_;
    std::vector<std::thread> pool_;
    std::atomic<unsigned> gen_{0};
    std::atomic<int> done_{0};
    const std::function<void(int, int)>* job_ = nullptr;
    bool stop_ = false;
};
}  // namespace

void SymIndefinite::Factor(int n, const std::vector<double>& a)
{
    n_ = n;
    A_.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) at(i, j) = a[(size_t)i * n + j];
    piv_.assign(n, 0);
    const double alpha = (1.0 + std::sqrt(17.0)) / 8.0;
    // (an unblocked factorisation streams the trailing matrix once per step: n^3/6 doubles of memory traffic -- 76 GB for the
    //  3 849 x 3 849 KKT matrix of config 4 -- so it is worth a team from a few hundred rows on)
    unsigned hw = std::thread::hardware_concurrency();
    ColumnTeam team(n >= 512 ? (int)std::max(1u, std::min(hw ? hw : 1u, 32u)) : 1);
    const int serial_below = 256;                            // columns left: not worth waking the team
    int k = 0;
    while (k < n) {
        int kstep = 1, kp = k;
        const double absakk = std::fabs(at(k, k));
        int imax = k; double colmax = 0.0;
        for (int i = k + 1; i < n; ++i) if (std::fabs(at(i, k)) > colmax) { colmax = std::fabs(at(i, k)); imax = i; }
        if (std::fmax(absakk, colmax) == 0.0 || std::isnan(absakk) || std::isnan(colmax)) {
            kp = k;   // singular column: leave it (D gets a zero)
        } else {
            if (absakk >= alpha * colmax) kp = k;
            else {
                double rowmax = 0.0;
                for (int j = k; j < imax; ++j) rowmax = std::fmax(rowmax, std::fabs(at(imax, j)));
                for (int i = imax + 1; i < n; ++i) rowmax = std::fmax(rowmax, std::fabs(at(i, imax)));
                if (absakk >= alpha * colmax * (colmax / rowmax)) kp = k;
                else if (std::fabs(at(imax, imax)) >= alpha * rowmax) kp = imax;
                else { kp = imax; kstep = 2; }
            }
            const int kk = k + kstep - 1;
            if (kp != kk) {
                for (int i = kp + 1; i < n; ++i) std::swap(at(i, kk), at(i, kp));
                for (int j = kk + 1; j < kp; ++j) std::swap(at(j, kk), at(kp, j));
                std::swap(at(kk, kk), at(kp, kp));
                if (kstep == 2) std::swap(at(k + 1, k), at(kp, k));
            }
            if (kstep == 1) {
                if (k < n - 1) {
                    const double r1 = 1.0 / at(k, k);
                    auto update = [&](int t, int T) {
                        for (int j = k + 1 + t; j < n; j += T) {
                            const double f = r1 * at(j, k);
                            if (f != 0.0) for (int i = j; i < n; ++i) at(i, j) -= at(i, k) * f;
                        }
                    };
                    if (n - k > serial_below && team.size() > 1) team.run(update); else update(0, 1);
                    for (int i = k + 1; i < n; ++i) at(i, k) *= r1;
                }
            } else if (k < n - 2) {
                double d21 = at(k + 1, k);
                const double d11 = at(k + 1, k + 1) / d21, d22 = at(k, k) / d21;
                const double t = 1.0 / (d11 * d22 - 1.0);
                d21 = t / d21;
                // (the serial loop overwrites at(j, k) and at(j, k+1) with the multipliers as it goes and later columns read the
                //  ORIGINAL entries of their own row j only, so the team first updates the columns, then stores the multipliers)
                std::vector<double> wks((size_t)n, 0.0), wkp1s((size_t)n, 0.0);
                for (int j = k + 2; j < n; ++j) {
                    wks[j] = d21 * (d11 * at(j, k) - at(j, k + 1));
                    wkp1s[j] = d21 * (d22 * at(j, k + 1) - at(j, k));
                }
                auto update = [&](int t, int T) {
                    for (int j = k + 2 + t; j < n; j += T) {
                        const double wk = wks[j], wkp1 = wkp1s[j];
                        for (int i = j; i < n; ++i) at(i, j) -= at(i, k) * wk + at(i, k + 1) * wkp1;
                    }
                };
                if (n - k > serial_below && team.size() > 1) team.run(update); else update(0, 1);
                for (int j = k + 2; j < n; ++j) { at(j, k) = wks[j]; at(j, k + 1) = wkp1s[j]; }
            }
        }
        if (kstep == 1) piv_[k] = kp;
        else { piv_[k] = -kp - 1; piv_[k + 1] = -kp - 1; }
        k += kstep;
    }
}

void SymIndefinite::Solve(const double* rhs, double* b) const
{
    const int n = n_;
    for (int i = 0; i < n; ++i) b[i] = rhs[i];
    int k = 0;
    while (k < n) {
        if (piv_[k] >= 0) {
            const int kp = piv_[k];
            if (kp != k) std::swap(b[k], b[kp]);
            for (int i = k + 1; i < n; ++i) b[i] -= b[k] * at(i, k);
            b[k] /= at(k, k);
            k += 1;
        } else {
            const int kp = -piv_[k] - 1;
            if (kp != k + 1) std::swap(b[k + 1], b[kp]);
            for (int i = k + 2; i < n; ++i) b[i] -= b[k] * at(i, k) + b[k + 1] * at(i, k + 1);
            const double akm1k = at(k + 1, k);
            const double akm1 = at(k, k) / akm1k, ak = at(k + 1, k + 1) / akm1k;
            const double denom = akm1 * ak - 1.0;
            const double bkm1 = b[k] / akm1k, bk = b[k + 1] / akm1k;
            b[k] = (ak * bkm1 - bk) / denom;
            b[k + 1] = (akm1 * bk - bkm1) / denom;
            k += 2;
        }
    }
    k = n - 1;
    while (k >= 0) {
        if (piv_[k] >= 0) {
            double s = 0.0;
            for (int i = k + 1; i < n; ++i) s += at(i, k) * b[i];
            b[k] -= s;
            const int kp = piv_[k];
            if (kp != k) std::swap(b[k], b[kp]);
            k -= 1;
        } else {
            double s0 = 0.0, s1 = 0.0;
            for (int i = k + 1; i < n; ++i) { s0 += at(i, k) * b[i]; s1 += at(i, k - 1) * b[i]; }
            b[k] -= s0;
            b[k - 1] -= s1;
            const int kp = -piv_[k] - 1;
            if (kp != k) std::swap(b[k], b[kp]);
            k -= 2;
        }
    }
}

void SymIndefinite::Inertia(int& pos, int& neg, int& zero) const
{
    pos = neg = zero = 0;
    int k = 0;
    while (k < n_) {
        if (piv_[k] >= 0) {
            const double d = at(k, k);
            if (d > 0) ++pos; else if (d < 0) ++neg; else ++zero;
            k += 1;
        } else {
            const double a = at(k, k), b = at(k + 1, k), c = at(k + 1, k + 1);
            const double det = a * c - b * b;
            if (det < 0) { ++pos; ++neg; }
            else if (det > 0) { if (a + c > 0) pos += 2; else neg += 2; }
            else { ++zero; if (a + c > 0) ++pos; else if (a + c < 0) ++neg; else ++zero; }
            k += 2;
        }
    }
}

void SymIndefinite::LogDet(double& logabs, int& sign) const
{
    logabs = 0.0; sign = 1;
    int k = 0;
    while (k < n_) {
        double d;
        if (piv_[k] >= 0) { d = at(k, k); k += 1; }
        else { d = at(k, k) * at(k + 1, k + 1) - at(k + 1, k) * at(k + 1, k); k += 2; }
        if (d == 0.0 || std::isnan(d)) { sign = 0; logabs = -INFINITY; return; }
        if (d < 0) sign = -sign;
        logabs += std::log(std::fabs(d));
    }
}

}  // namespace wfsa
