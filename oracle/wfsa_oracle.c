/* oracle/wfsa_oracle.c -- TEST INFRASTRUCTURE ONLY (see wfsa_oracle.h).
 *
 * Part 1 restates the reference: explicit enumeration of accepting paths
 * (/root/reference/inc/Recognize.h:35-96; which parameters a step counts:
 * src/Learner.cpp:285-293), then the per-string algebra of src/Learner.cpp:515-553,
 * src/QuasiNewtonLearner.cpp:93-125 and src/HessianLearner.cpp:381-547.
 * Part 2 is an independent CPU forward-backward for sizes where
 * enumeration is infeasible.  Plain C11 + OpenMP. */
#include "wfsa_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t* edges;      /* concatenated edge ids of all paths */
    int64_t n_edges, cap_edges;
    int64_t* off;        /* path offsets */
    int64_t n_paths, cap_paths;
} PathBuf;

static void pb_init(PathBuf* b) { memset(b, 0, sizeof(*b)); b->cap_paths = 16; b->off = malloc(sizeof(int64_t) * 17); b->off[0] = 0; }
static void pb_free(PathBuf* b) { free(b->edges); free(b->off); }
static void pb_reset(PathBuf* b) { b->n_edges = 0; b->n_paths = 0; b->off[0] = 0; }
static void pb_push(PathBuf* b, const int32_t* e, int n)
{
    if (b->n_edges + n > b->cap_edges) { b->cap_edges = (b->n_edges + n) * 2 + 64; b->edges = realloc(b->edges, sizeof(int32_t) * b->cap_edges); }
    memcpy(b->edges + b->n_edges, e, sizeof(int32_t) * n);
    b->n_edges += n;
    if (b->n_paths + 1 > b->cap_paths) { b->cap_paths *= 2; b->off = realloc(b->off, sizeof(int64_t) * (b->cap_paths + 1)); }
    b->off[++b->n_paths] = b->n_edges;
}

typedef struct {
    const oracle_fsa* f;
    const int32_t* tok;
    int len, n_trans;
    int32_t* stack;
    int depth;
    PathBuf* out;
    int64_t max_paths;
    int overflow;
} EnumCtx;

/* inc/Recognize.h:35-60 (DFS order; BFS enumerates the same set of paths) */
static void enum_dfs(EnumCtx* c, int pos, int state)
{
    const oracle_fsa* f = c->f;
    if (c->overflow) return;
    for (int t = f->trans_row[state]; t < f->trans_row[state + 1]; ++t) {
        const int v = f->trans_dst[t];
        if (v == f->end_state) {
            if (pos == c->len) {      /* the string has been consumed; only the transition counts */
                if (c->out->n_paths >= c->max_paths) { c->overflow = 1; return; }
                c->stack[c->depth] = t;
                pb_push(c->out, c->stack, c->depth + 1);
            }
            continue;
        }
        for (int e = f->emis_row[v]; e < f->emis_row[v + 1]; ++e) {
            const int e0 = f->emis_tok_off[e], el = f->emis_tok_off[e + 1] - e0;
            if (pos + el > c->len) continue;
            int m = 1;
            for (int k = 0; k < el; ++k) if (c->tok[pos + k] != f->emis_tok[e0 + k]) { m = 0; break; }
            if (!m) continue;
            c->stack[c->depth] = t;
            c->stack[c->depth + 1] = c->n_trans + e;
            c->depth += 2;
            enum_dfs(c, pos + el, v);
            c->depth -= 2;
        }
    }
}

static int enumerate(const oracle_fsa* f, const int32_t* tok, int len, PathBuf* out, int64_t max_paths)
{
    EnumCtx c;
    c.f = f; c.tok = tok; c.len = len; c.n_trans = f->trans_row[f->n_states];
    c.stack = malloc(sizeof(int32_t) * (size_t)(2 * ((size_t)len + 1) * (f->n_states + 1) + 4));
    c.depth = 0; c.out = out; c.max_paths = max_paths; c.overflow = 0;
    pb_reset(out);
    enum_dfs(&c, 0, f->start_state);
    free(c.stack);
    return c.overflow ? -1 : 0;
}

static double edge_lw(int e, int n_trans, const double* ltw, const double* lew) { return e < n_trans ? ltw[e] : lew[e - n_trans]; }

/* path posteriors r of one string; returns log q (or -inf) */
static double path_posteriors(const PathBuf* pb, int n_trans, const double* ltw, const double* lew, double* r)
{
    double m = -INFINITY;
    for (int64_t l = 0; l < pb->n_paths; ++l) {
        double s = 0.0;
        for (int64_t i = pb->off[l]; i < pb->off[l + 1]; ++i) s += edge_lw(pb->edges[i], n_trans, ltw, lew);   /* P.x */
        r[l] = s;
        if (s > m) m = s;
    }
    if (!(m > -INFINITY)) return -INFINITY;
    double q = 0.0;
    for (int64_t l = 0; l < pb->n_paths; ++l) { r[l] = exp(r[l] - m); q += r[l]; }               /* q = M.exp(P.x) */
    for (int64_t l = 0; l < pb->n_paths; ++l) r[l] /= q;                                        /* r /= M^t.q   */
    return m + log(q);
}

int oracle_enum_eval(const oracle_fsa* f, const oracle_corpus* c, const double* ltw, const double* lew,
                     double* path_count, double* logq, double* edge_exp, int64_t max_paths)
{
    const int n_trans = f->trans_row[f->n_states], n_emis = f->emis_row[f->n_states];
    if (edge_exp) memset(edge_exp, 0, sizeof(double) * (size_t)(n_trans + n_emis));
    PathBuf pb; pb_init(&pb);
    double* r = NULL; int64_t rcap = 0;
    int rc = 0;
    for (int64_t s = 0; s < c->n_strings; ++s) {
        const int len = (int)(c->offsets[s + 1] - c->offsets[s]);
        if (enumerate(f, c->tokens + c->offsets[s], len, &pb, max_paths) != 0) { rc = -1; break; }
        if (path_count) path_count[s] = (double)pb.n_paths;
        if (pb.n_paths > rcap) { rcap = pb.n_paths * 2; r = realloc(r, sizeof(double) * rcap); }
        const double lq = pb.n_paths ? path_posteriors(&pb, n_trans, ltw, lew, r) : -INFINITY;
        if (logq) logq[s] = lq;
        if (edge_exp && lq > -INFINITY)
            for (int64_t l = 0; l < pb.n_paths; ++l)
                for (int64_t i = pb.off[l]; i < pb.off[l + 1]; ++i) edge_exp[pb.edges[i]] += c->p[s] * r[l];
    }
    free(r); pb_free(&pb);
    return rc;
}

int oracle_enum_hessian(const oracle_fsa* f, const oracle_corpus* c, const double* ltw, const double* lew,
                        const int32_t* edge_param, int32_t n, double* H, int64_t max_paths)
{
    const int n_trans = f->trans_row[f->n_states];
    memset(H, 0, sizeof(double) * (size_t)n * n);
    PathBuf pb; pb_init(&pb);
    double* r = NULL; int64_t rcap = 0;
    double* P = NULL; size_t pcap = 0;
    int32_t* colof = malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t* cols = malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    double* g = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int rc = 0;
    for (int64_t s = 0; s < c->n_strings; ++s) {
        const int len = (int)(c->offsets[s + 1] - c->offsets[s]);
        if (enumerate(f, c->tokens + c->offsets[s], len, &pb, max_paths) != 0) { rc = -1; break; }
        if (pb.n_paths < 2) continue;                              /* only equivocal strings, :411 */
        if (pb.n_paths > rcap) { rcap = pb.n_paths * 2; r = realloc(r, sizeof(double) * rcap); }
        if (!(path_posteriors(&pb, n_trans, ltw, lew, r) > -INFINITY)) continue;
        /* union of the parameters on any path */
        int D = 0;
        for (int i = 0; i < n; ++i) colof[i] = -1;
        for (int64_t i = 0; i < pb.n_edges; ++i) {
            const int j = edge_param[pb.edges[i]];
            if (j >= 0 && colof[j] < 0) { colof[j] = D; cols[D++] = j; }
        }
        const size_t need = (size_t)pb.n_paths * (size_t)(D > 0 ? D : 1);
        if (need > pcap) { pcap = need * 2; P = realloc(P, sizeof(double) * pcap); }
        memset(P, 0, sizeof(double) * need);
        for (int64_t l = 0; l < pb.n_paths; ++l)
            for (int64_t i = pb.off[l]; i < pb.off[l + 1]; ++i) {
                const int j = edge_param[pb.edges[i]];
                if (j >= 0) P[(size_t)l * D + colof[j]] += 1.0;
            }
        /* drop (column,count) pairs identical on every path, :432-435 */
        int K = 0;
        for (int d = 0; d < D; ++d) {
            int same = 1;
            for (int64_t l = 1; l < pb.n_paths; ++l) if (P[(size_t)l * D + d] != P[d]) { same = 0; break; }
            if (!same) { cols[K] = cols[d]; for (int64_t l = 0; l < pb.n_paths; ++l) P[(size_t)l * D + K] = P[(size_t)l * D + d]; ++K; }
        }
        for (int a = 0; a < K; ++a) {
            g[a] = 0.0;
            for (int64_t l = 0; l < pb.n_paths; ++l) g[a] += P[(size_t)l * D + a] * r[l];        /* :523 */
        }
        for (int a = 0; a < K; ++a)
            for (int b = 0; b < K; ++b) {
                double hjk = 0.0;
                for (int64_t l = 0; l < pb.n_paths; ++l) hjk -= P[(size_t)l * D + a] * P[(size_t)l * D + b] * r[l];   /* :538-540 */
                hjk += g[a] * g[b];
                H[(size_t)cols[a] * n + cols[b]] += c->p[s] * hjk;                               /* :543 */
            }
    }
    free(r); free(P); free(colof); free(cols); free(g); pb_free(&pb);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* Part 2: forward-backward                                                                    */
/* ------------------------------------------------------------------------------------------ */
int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static double logaddexp(double a, double b)
{
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    const double m = a > b ? a : b;
    return m + log1p(exp(-fabs(a - b)));
}

typedef struct { int32_t dst, tid, eid; } Arc;
typedef struct {
    int single;               /* every consumed emission is exactly one token */
    int n_trans, n_emis, S, A;
    int64_t* row;             /* [S*A+1] arcs by (src, symbol) */
    Arc* arcs;
    int32_t* final_tid;       /* [S] */
    int32_t* eps_order;       /* [S] */
} DpModel;

static void dp_build(const oracle_fsa* f, DpModel* m)
{
    memset(m, 0, sizeof(*m));
    m->S = f->n_states; m->A = f->n_symbols;
    m->n_trans = f->trans_row[m->S]; m->n_emis = f->emis_row[m->S];
    m->single = 1;
    for (int s = 0; s < m->S; ++s) {
        if (s == f->start_state || s == f->end_state) continue;
        for (int e = f->emis_row[s]; e < f->emis_row[s + 1]; ++e)
            if (f->emis_tok_off[e + 1] - f->emis_tok_off[e] != 1) m->single = 0;
    }
    m->final_tid = malloc(sizeof(int32_t) * m->S);
    for (int s = 0; s < m->S; ++s) {
        m->final_tid[s] = -1;
        for (int t = f->trans_row[s]; t < f->trans_row[s + 1]; ++t) if (f->trans_dst[t] == f->end_state) m->final_tid[s] = t;
    }
    if (m->single) {
        const size_t R = (size_t)m->S * (m->A > 0 ? m->A : 1);
        m->row = calloc(R + 1, sizeof(int64_t));
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                int64_t acc = 0;
                for (size_t i = 0; i <= R; ++i) { const int64_t c = m->row[i]; m->row[i] = acc; acc += c; }
                m->arcs = malloc(sizeof(Arc) * (size_t)(acc > 0 ? acc : 1));
            }
            int64_t* fill = pass ? calloc(R + 1, sizeof(int64_t)) : NULL;
            for (int u = 0; u < m->S; ++u) {
                if (u == f->end_state) continue;
                for (int t = f->trans_row[u]; t < f->trans_row[u + 1]; ++t) {
                    const int v = f->trans_dst[t];
                    if (v == f->end_state || v == f->start_state) continue;
                    for (int e = f->emis_row[v]; e < f->emis_row[v + 1]; ++e) {
                        const size_t r = (size_t)u * m->A + f->emis_tok[f->emis_tok_off[e]];
                        if (!pass) m->row[r]++;
                        else { Arc a = {v, t, e}; m->arcs[m->row[r] + fill[r]++] = a; }
                    }
                }
            }
            free(fill);
        }
    } else {
        /* order in which empty-emission arcs only go forward */
        char* eps = calloc(m->S, 1);
        int* indeg = calloc(m->S, sizeof(int));
        for (int s = 0; s < m->S; ++s) {
            if (s == f->start_state || s == f->end_state) continue;
            for (int e = f->emis_row[s]; e < f->emis_row[s + 1]; ++e) if (f->emis_tok_off[e + 1] == f->emis_tok_off[e]) eps[s] = 1;
        }
        for (int u = 0; u < m->S; ++u) for (int t = f->trans_row[u]; t < f->trans_row[u + 1]; ++t) if (eps[f->trans_dst[t]]) indeg[f->trans_dst[t]]++;
        m->eps_order = malloc(sizeof(int32_t) * m->S);
        int head = 0, tail = 0;
        for (int s = 0; s < m->S; ++s) if (!indeg[s]) m->eps_order[tail++] = s;
        while (head < tail) {
            const int u = m->eps_order[head++];
            for (int t = f->trans_row[u]; t < f->trans_row[u + 1]; ++t) { const int v = f->trans_dst[t]; if (eps[v] && --indeg[v] == 0) m->eps_order[tail++] = v; }
        }
        for (int s = 0; tail < m->S && s < m->S; ++s) if (indeg[s] > 0) m->eps_order[tail++] = s;   /* cycle: undefined, caller's problem */
        free(eps); free(indeg);
    }
}

static void dp_free(DpModel* m) { free(m->row); free(m->arcs); free(m->final_tid); free(m->eps_order); }

/* scaled linear-domain forward-backward over sparse active sets (single-token automata) */
typedef struct {
    int32_t* st; double* al; int64_t cap;      /* lattice entries */
    int64_t* pos_off; double* cscale; int pcap;
    double* dense; double* beta; int32_t* touched;
} DpScratch;

static double dp_single(const oracle_fsa* f, const DpModel* m, const double* tw, const double* ew,
                        const int32_t* tok, int len, double ps, double* edge_exp, DpScratch* w)
{
    const int A = m->A;
    if (len == 0) {
        const int ft = m->final_tid[f->start_state];
        if (ft < 0 || !(tw[ft] > 0)) return -INFINITY;
        if (edge_exp) edge_exp[ft] += ps;
        return log(tw[ft]);
    }
    if (len + 1 > w->pcap) { w->pcap = len + 64; w->pos_off = realloc(w->pos_off, sizeof(int64_t) * (w->pcap + 1)); w->cscale = realloc(w->cscale, sizeof(double) * w->pcap); }
    int64_t n = 0;
    double logq = 0.0;
    w->pos_off[0] = 0;
    for (int t = 0; t < len; ++t) {
        const int c = tok[t];
        if (c < 0 || c >= A) return -INFINITY;
        int nt = 0;
        const int64_t b0 = t ? w->pos_off[t - 1] : 0, b1 = t ? w->pos_off[t] : 1;
        for (int64_t i = b0; i < b1; ++i) {
            const int u = t ? w->st[i] : f->start_state;
            const double au = t ? w->al[i] : 1.0;
            const size_t r = (size_t)u * A + c;
            for (int64_t k = m->row[r]; k < m->row[r + 1]; ++k) {
                const Arc a = m->arcs[k];
                const double x = au * tw[a.tid] * ew[a.eid];
                if (x == 0.0) continue;
                if (w->dense[a.dst] == 0.0) w->touched[nt++] = a.dst;
                w->dense[a.dst] += x;
            }
        }
        if (nt == 0) return -INFINITY;
        double sum = 0.0;
        for (int i = 0; i < nt; ++i) sum += w->dense[w->touched[i]];
        if (n + nt > w->cap) { w->cap = (n + nt) * 2 + 256; w->st = realloc(w->st, sizeof(int32_t) * w->cap); w->al = realloc(w->al, sizeof(double) * w->cap); }
        for (int i = 0; i < nt; ++i) { const int v = w->touched[i]; w->st[n] = v; w->al[n] = w->dense[v] / sum; w->dense[v] = 0.0; ++n; }
        w->cscale[t] = sum;
        logq += log(sum);
        w->pos_off[t + 1] = n;     /* entries of position t: [pos_off[t], pos_off[t+1]) */
    }
    /* careful: pos_off[t] is the START of position t */
    double qfin = 0.0;
    for (int64_t i = w->pos_off[len - 1]; i < w->pos_off[len]; ++i) { const int ft = m->final_tid[w->st[i]]; if (ft >= 0) qfin += w->al[i] * tw[ft]; }
    if (!(qfin > 0.0)) return -INFINITY;
    logq += log(qfin);
    if (!edge_exp) return logq;
    /* backward */
    for (int64_t i = w->pos_off[len - 1]; i < w->pos_off[len]; ++i) {
        const int u = w->st[i], ft = m->final_tid[u];
        const double b = ft >= 0 ? tw[ft] / qfin : 0.0;
        w->beta[u] = b;
        if (ft >= 0) edge_exp[ft] += ps * w->al[i] * b;
    }
    for (int t = len - 2; t >= -1; --t) {
        const int cn = tok[t + 1];
        const double inv = 1.0 / w->cscale[t + 1];
        const int64_t b0 = t >= 0 ? w->pos_off[t] : 0, b1 = t >= 0 ? w->pos_off[t + 1] : 1;
        /* new betas are written after all reads of position t+1's betas: stage them in al's sibling */
        for (int64_t i = b0; i < b1; ++i) {
            const int u = t >= 0 ? w->st[i] : f->start_state;
            const double au = t >= 0 ? w->al[i] : 1.0;
            const size_t r = (size_t)u * A + cn;
            double b = 0.0;
            for (int64_t k = m->row[r]; k < m->row[r + 1]; ++k) {
                const Arc a = m->arcs[k];
                const double term = tw[a.tid] * ew[a.eid] * w->beta[a.dst] * inv;
                if (term == 0.0) continue;
                b += term;
                const double gma = ps * au * term;
                edge_exp[a.tid] += gma;
                edge_exp[m->n_trans + a.eid] += gma;
            }
            w->dense[i - b0] = b;      /* dense[] is free here (all zero); reuse as staging, cleared below */
        }
        for (int64_t i = w->pos_off[t + 1]; i < w->pos_off[t + 2]; ++i) w->beta[w->st[i]] = 0.0;
        for (int64_t i = b0; i < b1; ++i) { if (t >= 0) w->beta[w->st[i]] = w->dense[i - b0]; w->dense[i - b0] = 0.0; }
    }
    for (int64_t i = w->pos_off[0]; i < w->pos_off[1]; ++i) w->beta[w->st[i]] = 0.0;
    return logq;
}

/* dense log-domain forward-backward for emissions of any length (incl. empty) */
static double dp_generic(const oracle_fsa* f, const DpModel* m, const double* ltw, const double* lew,
                         const int32_t* tok, int len, double ps, double* edge_exp)
{
    const int S = m->S;
    const size_t sz = (size_t)(len + 1) * S;
    double* la = malloc(sizeof(double) * sz * 2);
    double* lb = la + sz;
    for (size_t i = 0; i < 2 * sz; ++i) la[i] = -INFINITY;
    la[f->start_state] = 0.0;
    double lq = -INFINITY;
    for (int pos = 0; pos <= len; ++pos)
        for (int oi = 0; oi < S; ++oi) {
            const int u = m->eps_order[oi];
            const double au = la[(size_t)pos * S + u];
            if (au == -INFINITY || u == f->end_state) continue;
            for (int t = f->trans_row[u]; t < f->trans_row[u + 1]; ++t) {
                const int v = f->trans_dst[t];
                const double wv = au + ltw[t];
                if (wv == -INFINITY) continue;
                if (v == f->end_state) { if (pos == len) lq = logaddexp(lq, wv); continue; }
                for (int e = f->emis_row[v]; e < f->emis_row[v + 1]; ++e) {
                    const int e0 = f->emis_tok_off[e], el = f->emis_tok_off[e + 1] - e0;
                    if (pos + el > len) continue;
                    int ok = 1;
                    for (int k = 0; k < el; ++k) if (tok[pos + k] != f->emis_tok[e0 + k]) { ok = 0; break; }
                    if (!ok) continue;
                    double* d = la + (size_t)(pos + el) * S + v;
                    *d = logaddexp(*d, wv + lew[e]);
                }
            }
        }
    if (!(lq > -INFINITY) || !edge_exp) { free(la); return lq; }
    for (int pos = len; pos >= 0; --pos)
        for (int oi = S - 1; oi >= 0; --oi) {
            const int u = m->eps_order[oi];
            const double au = la[(size_t)pos * S + u];
            if (au == -INFINITY || u == f->end_state) continue;
            double bu = -INFINITY;
            for (int t = f->trans_row[u]; t < f->trans_row[u + 1]; ++t) {
                const int v = f->trans_dst[t];
                if (ltw[t] == -INFINITY) continue;
                if (v == f->end_state) {
                    if (pos == len) { bu = logaddexp(bu, ltw[t]); edge_exp[t] += ps * exp(au + ltw[t] - lq); }
                    continue;
                }
                for (int e = f->emis_row[v]; e < f->emis_row[v + 1]; ++e) {
                    const int e0 = f->emis_tok_off[e], el = f->emis_tok_off[e + 1] - e0;
                    if (pos + el > len) continue;
                    int ok = 1;
                    for (int k = 0; k < el; ++k) if (tok[pos + k] != f->emis_tok[e0 + k]) { ok = 0; break; }
                    if (!ok) continue;
                    const double bv = lb[(size_t)(pos + el) * S + v];
                    const double term = ltw[t] + lew[e] + bv;
                    if (term == -INFINITY) continue;
                    bu = logaddexp(bu, term);
                    const double gma = ps * exp(au + term - lq);
                    edge_exp[t] += gma;
                    edge_exp[m->n_trans + e] += gma;
                }
            }
            lb[(size_t)pos * S + u] = bu;
        }
    free(la);
    return lq;
}

int oracle_dp_eval(const oracle_fsa* f, const oracle_corpus* c, const double* ltw, const double* lew,
                   double* path_count, double* logq, double* edge_exp, int nthreads)
{
    DpModel m;
    dp_build(f, &m);
    const int ne = m.n_trans + m.n_emis;
    double* tw = malloc(sizeof(double) * (size_t)(m.n_trans > 0 ? m.n_trans : 1));
    double* ew = malloc(sizeof(double) * (size_t)(m.n_emis > 0 ? m.n_emis : 1));
    for (int i = 0; i < m.n_trans; ++i) tw[i] = exp(ltw[i]);
    for (int i = 0; i < m.n_emis; ++i) ew[i] = exp(lew[i]);
    int nth = nthreads > 0 ? nthreads : oracle_max_threads();
    if (nth < 1) nth = 1;
    double* acc = edge_exp ? calloc((size_t)nth * (size_t)(ne > 0 ? ne : 1), sizeof(double)) : NULL;
#ifdef _OPENMP
#pragma omp parallel num_threads(nth)
#endif
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        DpScratch w; memset(&w, 0, sizeof(w));
        w.dense = calloc((size_t)m.S + 1, sizeof(double));
        w.beta = calloc((size_t)m.S + 1, sizeof(double));
        w.touched = malloc(sizeof(int32_t) * ((size_t)m.S + 1));
        double* my = acc ? acc + (size_t)tid * ne : NULL;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
        for (int64_t s = 0; s < c->n_strings; ++s) {
            const int len = (int)(c->offsets[s + 1] - c->offsets[s]);
            const int32_t* tok = c->tokens + c->offsets[s];
            const double lq = m.single ? dp_single(f, &m, tw, ew, tok, len, c->p[s], my, &w)
                                       : dp_generic(f, &m, ltw, lew, tok, len, c->p[s], my);
            if (logq) logq[s] = lq;
            if (path_count) path_count[s] = lq > -INFINITY ? exp(lq) : 0.0;
        }
        free(w.st); free(w.al); free(w.pos_off); free(w.cscale); free(w.dense); free(w.beta); free(w.touched);
    }
    if (edge_exp) {
        memset(edge_exp, 0, sizeof(double) * (size_t)ne);
        for (int t = 0; t < nth; ++t) for (int e = 0; e < ne; ++e) edge_exp[e] += acc[(size_t)t * ne + e];
        free(acc);
    }
    free(tw); free(ew);
    dp_free(&m);
    return 0;
}
