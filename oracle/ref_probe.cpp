// oracle/ref_probe.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference classes (compiled from /root/reference/src by
// oracle/Makefile against oracle/mkl_shim) through their own C++ interface and dumps,
// at full double precision, everything the golden fixtures need:
//   * the frozen path matrices P, M, the constraint map C and p   (inc/Learner.h:143-166)
//   * the trimmed-index of every edge, keyed by (state, kind, label) -- recovered with
//     Learner::Init(0, sentinel) + Learner::RewriteWeights           (src/Learner.cpp:45-58,574-579)
//   * logq / KL at caller-supplied x via Learner::ComputeModeledProbs + ComputeObjective
//                                                                     (src/Learner.cpp:515-553)
//   * relative path probabilities r at the same x
//   * a K-epoch optimisation trajectory (info rows + final per-edge weights)
// Output is one JSON object on stdout.  Usage:
//   ref_probe <fsa> <corpus> <QuasiNewton|Hessian> <initflags> <epochs> <eta> <tol> [xfile]
// xfile: text, one evaluation point per line, n doubles each (n = trimmed parameters).
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include <memory>

#include "mkl.h"
#include "Fsa.h"
#include "Corpus.h"
#include "Learner.h"
#include "QuasiNewtonLearner.h"
#include "HessianLearner.h"
#include "Recognize.h"

template <class Base>
struct Probe : public Base
{
    using Base::_x; using Base::p; using Base::q; using Base::logq;
    using Base::relative_path_probs;
    using Base::Crow; using Base::Ccol;
    using Base::Prow; using Base::Pcol; using Base::Pdata;
    using Base::Mrow; using Base::Mcol;
    void SetX(const double* x) { for (MKL_INT i = 0; i < this->GetNumberOfParameters(); ++i) _x[i] = x[i]; }
};

static void jnum(double v)
{
    if (std::isnan(v)) printf("\"nan\"");
    else if (std::isinf(v)) printf(v > 0 ? "\"inf\"" : "\"-inf\"");
    else printf("%.17g", v);
}
static void jstr(const char* s)
{
    putchar('"');
    for (; *s; ++s) {
        if (*s == '"' || *s == '\\') { putchar('\\'); putchar(*s); }
        else if ((unsigned char)*s < 0x20) printf("\\u%04x", (unsigned)*s);
        else putchar(*s);
    }
    putchar('"');
}
template <class V> static void jvec(const V& v)
{
    putchar('[');
    for (size_t i = 0; i < v.size(); ++i) { if (i) putchar(','); jnum((double)v[i]); }
    putchar(']');
}

// every edge with its current logprob, in the reference's own iteration order
static void dump_edges(const Fsa& fsa)
{
    putchar('[');
    bool first = true;
    for (const auto& st : fsa.GetTransitionMtx()) {
        for (const auto& e : st.second.emissions) {
            if (!first) putchar(','); first = false;
            printf("{\"state\":"); jstr(st.first); printf(",\"kind\":\"E\",\"label\":"); jstr(e.str);
            printf(",\"raw\":%d,\"logprob\":", (int)e.index); jnum(e.logprob); putchar('}');
        }
        for (const auto& t : st.second.transitions) {
            if (!first) putchar(','); first = false;
            printf("{\"state\":"); jstr(st.first); printf(",\"kind\":\"T\",\"label\":"); jstr(t.next->first);
            printf(",\"raw\":%d,\"logprob\":", (int)t.index); jnum(t.logprob); putchar('}');
        }
    }
    putchar(']');
}

template <class L>
static int run(const char* fsa_fn, const char* corpus_fn, int flags, int epochs, double eta, double tol, const char* xfile)
{
    Corpus corpus;
    FILE* f = fopen(corpus_fn, "r");
    if (!f) { fprintf(stderr, "cannot open %s\n", corpus_fn); return 2; }
    corpus.Read(f); fclose(f);
    const double corpus_sum = corpus.Sum();
    corpus.Renormalize();                       // src/main.cpp:154

    Fsa fsa;
    f = fopen(fsa_fn, "r");
    if (!f) { fprintf(stderr, "cannot open %s\n", fsa_fn); return 2; }
    fsa.Read(f); fclose(f);

    Probe<L> L_;
    L_.BuildFrom(fsa, corpus, true);            // src/main.cpp:206

    printf("{\"corpus_size\":%zu,\"corpus_sum\":", corpus.size()); jnum(corpus_sum);
    printf(",\"states\":%zu,\"transitions\":%zu,\"emissions\":%zu,\"raw_parameters\":%zu,\"raw_constraints\":%zu",
        fsa.GetNumberOfStates(), fsa.GetNumberOfTransitions(), fsa.GetNumberOfEmissions(),
        fsa.GetNumberOfParameters(), fsa.GetNumberOfConstraints());
    printf(",\"strings\":%d,\"paths\":%d,\"common_support\":", (int)L_.GetNumberOfStrings(), (int)L_.GetNumberOfPaths());
    jnum(L_.GetCommonSupport());
    printf(",\"unique_paths\":%s,\"n\":%d,\"k\":%d", L_.HasUniquePaths() ? "true" : "false",
        (int)L_.GetNumberOfParameters(), (int)L_.GetNumberOfConstraints());
    printf(",\"raw_edges\":"); dump_edges(fsa);   // file weights + raw indices

    // accepting paths of every corpus string, counted by the reference's own recogniser
    // (inc/Recognize.h:62-96); 0 = not recognised
    {
        size_t count = 0;
        Recognizer<int> rec(fsa.GetEndState(),
            [](int& h, const Fsa::Transitions::value_type&, const Fsa::Emissions::value_type&) -> int& { return h; },
            [&](const int&) { ++count; });
        const auto& start = fsa.GetTransitionMtx().at(fsa.GetStartState());
        printf(",\"corpus\":[");
        bool first = true;
        for (const auto& w : corpus) {
            count = 0;
            rec.RecognizeBFS(w.first.c_str(), start, 0);
            if (!first) putchar(','); first = false;
            printf("{\"word\":"); jstr(w.first.c_str()); printf(",\"p\":"); jnum(w.second);
            printf(",\"paths\":%zu}", count);
        }
        printf("]");
    }

    const int n = L_.GetNumberOfParameters();
    if (n == 0 || L_.GetNumberOfStrings() == 0) {   // src/main.cpp:217-228
        printf(",\"degenerate\":true}\n");
        return 0;
    }
    L_.Finalize();                               // src/main.cpp:229

    // x as read from the file (after Trim compaction), before any Init flag
    printf(",\"x_file\":"); jvec(std::vector<double>(L_._x.begin(), L_._x.begin() + n));
    printf(",\"Ccol\":"); jvec(L_.Ccol);
    printf(",\"Prow\":"); jvec(L_.Prow); printf(",\"Pcol\":"); jvec(L_.Pcol); printf(",\"Pdata\":"); jvec(L_.Pdata);
    printf(",\"Mrow\":"); jvec(L_.Mrow); printf(",\"Mcol\":"); jvec(L_.Mcol);
    printf(",\"p\":"); jvec(L_.p);

    // index map: sentinel x_i = i + 1, then RewriteWeights -> each edge shows its trimmed index
    {
        std::vector<double> sentinel(n);
        for (int i = 0; i < n; ++i) sentinel[i] = i + 1;
        std::vector<double> keep(L_._x.begin(), L_._x.begin() + n);
        L_.SetX(sentinel.data());
        Fsa tmp(fsa);
        L_.RewriteWeights(tmp);
        printf(",\"sentinel_edges\":"); dump_edges(tmp);
        L_.SetX(keep.data());
    }

    // evaluations at given points
    printf(",\"evals\":[");
    if (xfile) {
        std::ifstream xs(xfile);
        std::string line; bool first = true;
        while (std::getline(xs, line)) {
            std::istringstream iss(line);
            std::vector<double> x; double v;
            while (iss >> v) x.push_back(v);
            if ((int)x.size() != n) continue;
            std::vector<double> keep(L_._x.begin(), L_._x.begin() + n);
            L_.SetX(x.data());
            L_.ComputeModeledProbs();
            L_.ComputeObjective();
            if (!first) putchar(','); first = false;
            printf("{\"x\":"); jvec(x);
            printf(",\"logq\":"); jvec(L_.logq);
            printf(",\"kl\":"); jnum(L_.GetKLDistance());
            printf(",\"r\":"); jvec(L_.relative_path_probs);
            putchar('}');
            L_.SetX(keep.data());
        }
    }
    printf("]");

    // optimisation trajectory                                  src/main.cpp:251-304
    L_.Init(flags);
    printf(",\"x_init\":"); jvec(std::vector<double>(L_._x.begin(), L_._x.begin() + n));
    printf(",\"trajectory\":[");
    bool halted = false; int last_epoch = 0;
    std::string err;
    try {
        for (int e = 1; e <= epochs; ++e) {
            L_.OptimizationStep(eta, false);
            const auto info = L_.GetOptimizationInfo();
            if (e > 1) putchar(',');
            printf("{\"epoch\":%d,\"info\":", e); jvec(info);
            printf(",\"x\":"); jvec(std::vector<double>(L_._x.begin(), L_._x.begin() + n));
            putchar('}');
            last_epoch = e;
            bool bad = false;
            for (double v : info) if (!std::isfinite(v)) bad = true;
            if (bad) { err = "non-finite info"; break; }
            if (L_.HaltCondition(tol)) { halted = true; break; }
        }
    } catch (const std::exception& ex) { err = ex.what(); }
    printf("],\"halted\":%s,\"last_epoch\":%d,\"error\":", halted ? "true" : "false", last_epoch); jstr(err.c_str());
    {
        Fsa tmp(fsa);
        L_.RewriteWeights(tmp);
        printf(",\"final_edges\":"); dump_edges(tmp);
    }
    printf("}\n");
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 8) {
        fprintf(stderr, "usage: %s fsa corpus QuasiNewton|Hessian initflags epochs eta tol [xfile]\n", argv[0]);
        return 2;
    }
    const int flags = atoi(argv[4]), epochs = atoi(argv[5]);
    const double eta = atof(argv[6]), tol = atof(argv[7]);
    const char* xfile = argc > 8 ? argv[8] : nullptr;
    try {
        if (strcmp(argv[3], "Hessian") == 0)
            return run<HessianLearner>(argv[1], argv[2], flags, epochs, eta, tol, xfile);
        return run<QuasiNewtonLearner>(argv[1], argv[2], flags, epochs, eta, tol, xfile);
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
}
