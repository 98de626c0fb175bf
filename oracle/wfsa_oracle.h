/* oracle/wfsa_oracle.h -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's
 * algorithm for the objective / gradient / H_f path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (w-fsa_b200/) never does.
 *
 * Pinned against the reference itself: tests/test_oracle.py checks every function below
 * against tests/golden/*.json, which oracle/make_golden.py produced by running the unmodified
 * reference (oracle/_ref/ref_probe). */
#ifndef WFSA_ORACLE_H
#define WFSA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* same field layout as wfsa_fsa_desc / wfsa_corpus_desc of include/wfsa_dev.h */
typedef struct {
    int32_t n_states, start_state, end_state, n_symbols, n_raw_params;
    const int32_t *emis_row, *emis_tok_off, *emis_tok, *emis_param, *trans_row, *trans_dst, *trans_param;
} oracle_fsa;
typedef struct {
    int64_t n_strings;
    const int64_t* offsets;
    const int32_t* tokens;
    const double* p;
} oracle_corpus;

/* Reference algorithm: enumerate every accepting path of every string
 * (/root/reference/inc/Recognize.h:35-96), weigh it with exp(sum of log-weights of its edges)
 * (src/Learner.cpp:530-533), q = sum over paths (:536), r = path posterior (:542-545),
 * edge_exp[e] = sum_s p_s sum_paths r * count_e (src/QuasiNewtonLearner.cpp:117-123, per edge).
 * ltw / lew: log-weight of every transition / emission edge (-inf = trimmed away).
 * Edge ids: transitions 0..n_trans-1, emissions n_trans..n_trans+n_emis-1.
 * Returns 0, or -1 if a string has more than max_paths paths. */
int oracle_enum_eval(const oracle_fsa* f, const oracle_corpus* c, const double* ltw, const double* lew,
                     double* path_count, double* logq, double* edge_exp, int64_t max_paths);

/* H_f of src/HessianLearner.cpp:381-547: edge_param[e] = trimmed parameter of edge e or <0;
 * H is n x n row major, full symmetric. */
int oracle_enum_hessian(const oracle_fsa* f, const oracle_corpus* c, const double* ltw, const double* lew,
                        const int32_t* edge_param, int32_t n, double* H, int64_t max_paths);

/* The same quantities by a forward-backward dynamic program, scaled
 * linear domain for single-token automata (sparse active sets), dense log domain otherwise.
 * OpenMP over strings; nthreads <= 0 = all cores.  Used as the CPU baseline at sizes where
 * enumeration is infeasible, and as a second opinion in the parity tests. */
int oracle_dp_eval(const oracle_fsa* f, const oracle_corpus* c, const double* ltw, const double* lew,
                   double* path_count_or_null, double* logq, double* edge_exp, int nthreads);
int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
