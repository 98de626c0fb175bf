// w-fsa_b200/csrc/kernels_seg.cuh -- sm_100a kernels over the SEGMENTED compiled lattices (lattice.hpp).
//
// Every string's trimmed lattice is cut at the nodes all of its accepting paths share.  Between two
// cuts lies either one edge (a bridge) or a small DAG (a region); path weights factor over segments:
//     log q_s = sum_{bridges of s} log w[arc]  +  sum_{regions of s} log q_region
//     E_s[count(arc)] = [arc is a bridge of s]  +  posterior of the arc inside its region
// so one evaluation (Learner::ComputeModeledProbs + ComputeObjective + ComputeGrad,
// /root/reference/src/Learner.cpp:515-553, src/QuasiNewtonLearner.cpp:93-125) becomes
//   KR  kr_regions : forward-backward over every distinct region TYPE once (thread per type):
//                    lq[type] = log q_type,  acc[arc] += W_type * posterior   (64-bit fixed-point REDs),
//                    loglik += W_type * lq[type]
// The bridges need no per-string work: with c[arc] = sum_s p_s * (times arc is a bridge of s), computed when the
// lattices are compiled, their part of the gradient is the constant c and their part of the log-likelihood is
// sum_arc c[arc] * log w[arc] (k_prep6).  Only a caller that wants log q_s of every string runs
//   KS  ks_strings : thread per string: log q_s = sum of log w over its bridge arcs (16-bit ids, table in
//                    shared memory) + sum of lq[type] over its regions.
#pragma once
#include "kernels.cuh"

namespace wfsa {

struct KRParams {
    const double* __restrict__ aw;        // [n_arcs] a(u,v) * b(v,e)
    const uint32_t* __restrict__ words;   // word j of lane l of group g at goff[g] + j*32 + l
    const int64_t* __restrict__ goff;     // [n_groups+1]
    const int32_t* __restrict__ grows;    // [n_groups] 0x10000|paths<<8|length = path form; 4/8/12/16 = small DAG; else big DAG (multiple of 16)
    const double* __restrict__ typeW;     // [n_groups*32]
    double* lq;                           // [n_groups*32 + 1] log q per type (last = dummy, stays 0)
    long long n_groups;
    double* xs;                           // big regions: per warp [xs_rows][32]
    size_t xs_rows;
    unsigned int* counter;                // dynamic group scheduler
    unsigned long long* acc;              // [replicas][n_arcs]
    double fx_scale;
    int n_arcs, replicas;
    int no_long_reds;                     // timing experiment of k_eval6 (debug bit 3): long path-form types skip their REDs
    // log-likelihood: sum_s p_s log q_s = sum_types W_type * lq_type + sum_arcs (bridge count of the arc) * log w[arc].
    // The first sum is accumulated here (fixed point, one RED per CTA), the second one arrives as per-CTA partials of
    // the weight kernel in llpart[n_llpart][2] = (value, non-finite terms) and is added by CTA 0.
    double ll_scale;
    unsigned long long* red;              // red[0] fixed-point loglik, red[1] non-finite terms
    const long long* __restrict__ llpart;
    int n_llpart;
};

// 64-bit add without a return value, spelled as `red` so that ptxas emits REDG in every kernel: in a kernel that also
// holds a grid barrier (k_eval6) it keeps atomicAdd() with an unused result as ATOMG, whose lanes travel the return path
// (measured: 0 RED sectors in ncu, region phase 95 us instead of 55 us).
__device__ __forceinline__ void red_add64(unsigned long long* p, unsigned long long v)
{
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// W * log q of one region type into the thread's fixed-point sum (W = 0: a padding lane)
__device__ __forceinline__ void kr_loglik(const KRParams& P, double W, bool ok, double lq, long long& ll)
{
    if (W > 0.0) {
        if (ok) ll += __double2ll_rn(W * lq * P.ll_scale);
        else atomicAdd(P.red + 1, 1ull);
    }
}

// Adds one fixed-point value per lane into acc[key].  Region types are sorted by their arcs, so at the first
// edge positions a whole warp usually holds ONE arc: then the 32 values are summed with shuffles (integers:
// exact, order independent) and a single RED is issued; otherwise every lane issues its own.
// Must be called by all 32 lanes; key < 0 = nothing to add.
__device__ __forceinline__ void red_uniform(unsigned long long* acc_g, int key, long long v, int lane)
{
    const int k0 = __shfl_sync(FULL, key, 0);
    if (__all_sync(FULL, key == k0)) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0 && key >= 0 && v) red_add64(acc_g + key, (unsigned long long)v);
    } else if (key >= 0 && v) red_add64(acc_g + key, (unsigned long long)v);
}

// One region in PATH FORM per thread: PP paths (padded with zero-weight paths) of L edges each, arc of edge l of
// path p at row l*PP + p.  q = sum_p prod_l w[arc]; every edge of path p gets the posterior r_p / q.  Registers
// only: PP independent multiply chains, no pool, no flags (the reference's P.x / exp / M algebra,
// src/Learner.cpp:530-545, for one small segment).
template <int PP, int ACC>
__device__ __forceinline__ void kr_paths(const KRParams& P, const double* aw, int L, long long g, long long off, int lane,
                                         unsigned long long* acc_g, long long& ll)
{
    const uint32_t* wp = P.words + off + lane;
    const double W = P.typeW[g * 32 + lane];
    double r[PP];
#pragma unroll
    for (int p = 0; p < PP; ++p) r[p] = 1.0;
    for (int l = 0; l < L; ++l) {
        uint32_t a[PP];
#pragma unroll
        for (int p = 0; p < PP; ++p) a[p] = __ldg(wp + (size_t)(l * PP + p) * 32);
#pragma unroll
        for (int p = 0; p < PP; ++p) r[p] *= aw[a[p]];
    }
    double q = 0.0;
#pragma unroll
    for (int p = 0; p < PP; ++p) q += r[p];
    const bool ok = W > 0.0 && q > 0.0 && isfinite(q);
    if (W > 0.0) {
        const double lq = ok ? log(q) : -INFINITY;
        P.lq[g * 32 + lane] = lq;
        kr_loglik(P, W, ok, lq, ll);
    }
    if (ACC == ACC_NONE) return;
    const double sc = ok ? W * P.fx_scale / q : 0.0;
    long long v[PP];
#pragma unroll
    for (int p = 0; p < PP; ++p) v[p] = __double2ll_rn(r[p] * sc);
    for (int l = 0; l < L; ++l) {
        uint32_t a[PP];
#pragma unroll
        for (int p = 0; p < PP; ++p) a[p] = __ldg(wp + (size_t)(l * PP + p) * 32);
#pragma unroll
        for (int p = 0; p < PP; ++p) {
            if (ACC == ACC_GLOBAL && l == 0 && p < 2) red_uniform(acc_g, v[p] ? (int)a[p] : -1, v[p], lane);
            else if (v[p]) red_add64(acc_g + a[p], (unsigned long long)v[p]);
        }
    }
}

// One big region per thread: the KL stream loop (CHECK words rescale, x values on a stack).  STAGED = false: words
// streamed from HBM (wp = P.words + goff[g] + lane), stack in HBM (one slab per warp).  STAGED = true (k_eval6): the
// caller copied the nw word rows of the group to shared memory (wp) and the stack lives there as well, so that the serial
// chain of a region (up to ~100 dependent steps, twice) runs at shared-memory latency.
template <int ACC, bool STAGED>
__device__ __forceinline__ void kr_big_t(const KRParams& P, const double* aw, double* pool, int NT, long long g, int lane,
                                         const uint32_t* wp, int nw, double* xs, unsigned long long* acc_g, long long& ll, double* trash)
{
    // The chain of a region is serial (every step reads the node value the previous step wrote), and ONE warp walks it:
    // what a step costs is the latency of its dependent instructions.  Both chain loops are therefore free of branches
    // inside a batch of eight words: every lane does the same loads, arithmetic and stores, and a lane whose word is
    // padding, FIN or belongs to a shorter region works on slot 0 / arc 0 and stores into `trash` (one double per
    // thread).  The decode and the weight loads of step k + 1 then overlap the chain of step k.  The posteriors are
    // rounded inside the backward chain and parked in the x stack; a third loop without dependencies issues the REDs.
    // Measured before (45 instructions with four branches per step): ~300 cycles per step, the 96-row group of a
    // 125 k-string shard alone took 31 us of a 37 us region phase.
    // STAGED: words and x stack are in shared memory; address them as such (a generic store to the shared window sits in the
    // same in-order LSU queue as the chain's LDS/STS and takes longer to resolve)
    const unsigned int wp_sa = STAGED ? (unsigned int)__cvta_generic_to_shared(wp) : 0u;
    const unsigned int xs_sa = STAGED ? (unsigned int)__cvta_generic_to_shared(xs) : 0u;
    auto ldw = [&](int i) -> uint32_t {
        if (STAGED) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(wp_sa + (unsigned int)i * 128u)); return v; }
        return __ldcs(wp + (size_t)i * 32);
    };
    auto xs_ld = [&](int i) -> long long {
        if (STAGED) { long long v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(xs_sa + (unsigned int)i * 256u)); return v; }
        return reinterpret_cast<const long long*>(xs)[(size_t)i * 32];
    };
    auto xs_st = [&](int i, long long v) {
        // (no memory clobber: the x stack is touched by these volatile statements only, which keep their order among themselves;
        //  a clobber would pin the chain's pool accesses around every store)
        if (STAGED) asm volatile("st.shared.b64 [%0], %1;" :: "r"(xs_sa + (unsigned int)i * 256u), "l"(v));
        else reinterpret_cast<long long*>(xs)[(size_t)i * 32] = v;
    };
    const double W = P.typeW[g * 32 + lane];
    pool[0] = 1.0;
    int E = 0, EQ = 0;
    double qh = 0.0;
    bool any = false;
    for (int i0 = 0; i0 < nw; i0 += 8) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = ldw(i0 + j);
        const bool has_chk = (i0 & 8) != 0;                   // word 7 of every second batch is a CHECK word (same index in all lanes)
        if (has_chk) w[7] = 0u;
        // decode first (addresses and weights of the eight steps are independent of the chain), then the chain itself:
        // per step  LDS node -> DMUL -> DADD -> select -> STS  and nothing else
        const double* ps[8]; double* pd[8]; double wv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t wj = w[j];
            const bool edge = (wj & kLEdge) != 0;
            const int src = edge ? (wj >> 19) & 15 : (wj & 15), arc = edge ? (int)(wj & 0xffff) : 0;
            ps[j] = pool + src * NT;                          // (FIN: the exit node)
            pd[j] = edge ? pool + ((wj >> 23) & 15) * NT : trash;
            wv[j] = aw[arc];
            any |= edge;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t wj = w[j];
            const bool fin = (wj & (kLEdge | kLFin)) == kLFin;
            const double a = *ps[j];
            const double xv = a * wv[j];
            const double sum = *pd[j] + xv;
            *pd[j] = (wj & kLFirstIn) ? xv : sum;
            xs_st(i0 + j, __double_as_longlong(xv));
            qh = fin ? a : qh;
            EQ = fin ? E : EQ;
        }
        if (has_chk) {
            const uint32_t m0 = ldw(i0 + 7) & 0xffffu;
            xs_st(i0 + 7, (long long)E);
            if (m0) {
                int emax = 0;
                for (uint32_t m = m0; m; m &= m - 1) emax = max(emax, pool_exp(pool + (__ffs(m) - 1) * NT));
                if (emax != 0 && (emax < 1023 - kLBand || emax > 1023 + kLBand)) {
                    const int shift = 1023 - emax;
                    for (uint32_t m = m0; m; m &= m - 1) { double* q = pool + (__ffs(m) - 1) * NT; *q = scalbn(*q, shift); }
                    E -= shift;
                }
            }
        }
    }
    const bool ok = any && qh > 0.0 && isfinite(qh);
    if (any) {
        const double lq = ok ? log(qh) + (double)EQ * 0.69314718055994530942 : -INFINITY;
        P.lq[g * 32 + lane] = lq;
        kr_loglik(P, W, ok, lq, ll);
    }
    const double sc0 = ok ? W * P.fx_scale / qh : 0.0;
    double sc = sc0;
    int F = 0;
    for (int i0 = nw - 8; i0 >= 0; i0 -= 8) {
        uint32_t w[8];
        double xv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { w[j] = ldw(i0 + j); xv[j] = __longlong_as_double(xs_ld(i0 + j)); }
        const bool has_chk = (i0 & 8) != 0;
        if (has_chk) {
            const uint32_t m0 = w[7] & 0xffffu;
            if (m0) {
                const int Et = (int)__double_as_longlong(xv[7]);
                int emax = 0;
                for (uint32_t m = m0; m; m &= m - 1) emax = max(emax, pool_exp(pool + (__ffs(m) - 1) * NT));
                if (emax != 0 && (emax < 1023 - kLBand || emax > 1023 + kLBand)) {
                    const int shift = 1023 - emax;
                    for (uint32_t m = m0; m; m &= m - 1) { double* q = pool + (__ffs(m) - 1) * NT; *q = scalbn(*q, shift); }
                    F -= shift;
                }
                sc = scalbn(sc0, Et + F - EQ);
            }
            w[7] = 0u;
        }
        const double* pdst[8]; double* psrc[8]; double wv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t wj = w[j];
            const bool edge = (wj & kLEdge) != 0;
            const int slot = edge ? (wj >> 19) & 15 : (wj & 15), arc = edge ? (int)(wj & 0xffff) : 0;   // (FIN: the exit node, beta = 1)
            pdst[j] = pool + ((wj >> 23) & 15) * NT;
            psrc[j] = (wj & (kLEdge | kLFin)) ? pool + slot * NT : trash;
            wv[j] = aw[arc];
        }
#pragma unroll
        for (int j = 7; j >= 0; --j) {
            const uint32_t wj = w[j];
            const bool edge = (wj & kLEdge) != 0;
            const bool fin = (wj & (kLEdge | kLFin)) == kLFin;
            const double bd = *pdst[j];
            const double c = wv[j] * bd;
            const double sum = *psrc[j] + c;
            *psrc[j] = fin ? 1.0 : ((wj & kLLastOut) ? c : sum);
            const long long v = __double2ll_rn(xv[j] * bd * sc);
            xs_st(i0 + j, (edge && ok) ? v : 0ll);
        }
    }
    if (ACC == ACC_NONE) return;
    for (int i0 = 0; i0 < nw; i0 += 8) {                      // the REDs: no dependencies between the steps
        uint32_t w[8];
        long long v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { w[j] = ldw(i0 + j); v[j] = xs_ld(i0 + j); }
#pragma unroll
        for (int j = 0; j < 8; ++j) if (v[j]) red_add64(acc_g + (w[j] & 0xffff), (unsigned long long)v[j]);
    }
}

// ---- one big DAG group on TWO warps (k_eval6, a staged group on a dedicated CTA) ---------------------------------------
// The chain of a region costs a lone warp ~4 cycles per instruction whatever is done about its dependencies (interleaving
// the forward with the backward chain in one instruction stream was measured twice: slower).  But the two chains are
// independent until the posteriors -- alpha needs the weights only, beta as well -- so they can run on two warps, i.e. two
// schedulers: warp 0 walks the words forward (x stack: alpha_src * w per edge, E at the CHECK words), warp 1 backward on its own
// pool (y stack: beta_dst per edge, F at the CHECK words), and warp 0 then forms (x * y) * sc in the order and with the scale
// factors of kr_big_t -- bitwise the same posteriors -- and issues the REDs.  All addresses are shared-memory byte addresses of
// the calling lane's column: wp_sa words [rows][32] u32, xs_sa / ys_sa stacks [rows][32] f64, pool at stride NT doubles.
__device__ __forceinline__ double bg_lds(unsigned int a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void bg_sts(unsigned int a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v)); }
__device__ __forceinline__ uint32_t bg_ldw(unsigned int wp_sa, int i) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(wp_sa + (unsigned int)i * 128u)); return v; }
__device__ __forceinline__ void bg_st64(unsigned int base, int i, long long v) { asm volatile("st.shared.b64 [%0], %1;" :: "r"(base + (unsigned int)i * 256u), "l"(v)); }
__device__ __forceinline__ long long bg_ld64(unsigned int base, int i) { long long v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(base + (unsigned int)i * 256u)); return v; }
__device__ __forceinline__ int bg_exp(unsigned int a) { int hi; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"(a + 4u)); return (hi >> 20) & 0x7ff; }

// forward chain; returns q (its exponent in EQ) and whether the lane holds a region at all
__device__ __forceinline__ void kr_big_fwd2(const double* aw, unsigned int p_sa, unsigned int ps, unsigned int tr_sa, unsigned int wp_sa, int nw,
                                            unsigned int xs_sa, double& qh, int& EQ, bool& any)
{
    bg_sts(p_sa, 1.0);
    int E = 0;
    EQ = 0; qh = 0.0; any = false;
    for (int i0 = 0; i0 < nw; i0 += 8) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = bg_ldw(wp_sa, i0 + j);
        const bool chk = (i0 & 8) != 0;
        const uint32_t m0 = chk ? w[7] & 0xffffu : 0u;
        if (chk) w[7] = 0u;
        unsigned int fs[8], fd[8];
        double fw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t wj = w[j];
            const bool edge = (wj & kLEdge) != 0;
            fs[j] = p_sa + (edge ? (wj >> 19) & 15u : (wj & 15u)) * ps;
            fd[j] = edge ? p_sa + ((wj >> 23) & 15u) * ps : tr_sa;
            fw[j] = aw[edge ? (wj & 0xffffu) : 0u];
            any |= edge;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t wj = w[j];
            const bool fin = (wj & (kLEdge | kLFin)) == kLFin;
            const double a = bg_lds(fs[j]);
            const double xv = a * fw[j];
            const double old = bg_lds(fd[j]);
            bg_sts(fd[j], (wj & kLFirstIn) ? xv : old + xv);
            bg_st64(xs_sa, i0 + j, __double_as_longlong(xv));
            qh = fin ? a : qh;
            EQ = fin ? E : EQ;
        }
        if (chk) {
            bg_st64(xs_sa, i0 + 7, (long long)E);
            if (m0) {
                int emax = 0;
                for (uint32_t m = m0; m; m &= m - 1) emax = max(emax, bg_exp(p_sa + (unsigned int)(__ffs(m) - 1) * ps));
                if (emax != 0 && (emax < 1023 - kLBand || emax > 1023 + kLBand)) {
                    const int shift = 1023 - emax;
                    for (uint32_t m = m0; m; m &= m - 1) { const unsigned int a = p_sa + (unsigned int)(__ffs(m) - 1) * ps; bg_sts(a, scalbn(bg_lds(a), shift)); }
                    E -= shift;
                }
            }
        }
    }
}

// backward chain on the caller's own pool: beta of the target node of every edge to the y stack, F to the CHECK slots
__device__ __forceinline__ void kr_big_bwd2(const double* aw, unsigned int p_sa, unsigned int ps, unsigned int tr_sa, unsigned int wp_sa, int nw,
                                            unsigned int ys_sa)
{
    int F = 0;
    for (int i0 = nw - 8; i0 >= 0; i0 -= 8) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = bg_ldw(wp_sa, i0 + j);
        const bool chk = (i0 & 8) != 0;
        if (chk) {                                             // the CHECK word comes first in backward order
            const uint32_t m0 = w[7] & 0xffffu;
            if (m0) {
                int emax = 0;
                for (uint32_t m = m0; m; m &= m - 1) emax = max(emax, bg_exp(p_sa + (unsigned int)(__ffs(m) - 1) * ps));
                if (emax != 0 && (emax < 1023 - kLBand || emax > 1023 + kLBand)) {
                    const int shift = 1023 - emax;
                    for (uint32_t m = m0; m; m &= m - 1) { const unsigned int a = p_sa + (unsigned int)(__ffs(m) - 1) * ps; bg_sts(a, scalbn(bg_lds(a), shift)); }
                    F -= shift;
                }
            }
            bg_st64(ys_sa, i0 + 7, (long long)F);
            w[7] = 0u;
        }
        unsigned int bd_[8], bs[8];
        double bw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t wj = w[j];
            const bool edge = (wj & kLEdge) != 0;
            bd_[j] = p_sa + ((wj >> 23) & 15u) * ps;
            bs[j] = (wj & (kLEdge | kLFin)) ? p_sa + (edge ? (wj >> 19) & 15u : (wj & 15u)) * ps : tr_sa;
            bw[j] = aw[edge ? (wj & 0xffffu) : 0u];
        }
#pragma unroll
        for (int j = 7; j >= 0; --j) {
            const uint32_t wj = w[j];
            const bool fin = (wj & (kLEdge | kLFin)) == kLFin;
            const double bd = bg_lds(bd_[j]);
            const double c = bw[j] * bd;
            const double old = bg_lds(bs[j]);
            bg_sts(bs[j], fin ? 1.0 : ((wj & kLLastOut) ? c : old + c));
            if (!(j == 7 && chk)) bg_st64(ys_sa, i0 + j, __double_as_longlong(bd));
        }
    }
}

// posteriors (x * y) * sc and their REDs; sc follows the CHECK words exactly as in kr_big_t
__device__ __forceinline__ void kr_big_red2(unsigned int wp_sa, int nw, unsigned int xs_sa, unsigned int ys_sa, bool ok, double sc0, int EQ,
                                            unsigned long long* acc_g)
{
    double sc = sc0;
    for (int i0 = nw - 8; i0 >= 0; i0 -= 8) {
        uint32_t w[8];
        long long x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { w[j] = bg_ldw(wp_sa, i0 + j); x[j] = bg_ld64(xs_sa, i0 + j); y[j] = bg_ld64(ys_sa, i0 + j); }
        if (i0 & 8) {
            if (w[7] & 0xffffu) sc = scalbn(sc0, (int)x[7] + (int)y[7] - EQ);
            w[7] = 0u;
        }
#pragma unroll
        for (int j = 7; j >= 0; --j) {
            if (!(w[j] & kLEdge) || !ok) continue;
            const long long v = __double2ll_rn(__longlong_as_double(x[j]) * __longlong_as_double(y[j]) * sc);
            if (v) red_add64(acc_g + (w[j] & 0xffff), (unsigned long long)v);
        }
    }
}

template <int ACC>
__device__ __forceinline__ void kr_big(const KRParams& P, const double* aw, double* pool, int NT, long long g, int lane,
                                       double* xs, unsigned long long* acc_g, long long& ll, double* trash)
{
    const long long o = P.goff[g];
    kr_big_t<ACC, false>(P, aw, pool, NT, g, lane, P.words + o + lane, (int)((P.goff[g + 1] - o) >> 5), xs, acc_g, ll, trash);
}

template <int ACC, int MAXNT, bool AWG = false>
__global__ void __launch_bounds__(MAXNT, 1) kr_regions(const KRParams P)
{
    extern __shared__ unsigned long long smem[];
    __shared__ double s_trash[MAXNT];                         // where the lanes without an edge store (kr_big_t)
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
    // [n_arcs + 1] arc weights, the last entry the zero weight of padding: a copy in shared memory, or (AWG: the table does not
    // fit) P.aw itself, which then has the zero entry behind it
    const double* aw = AWG ? P.aw : reinterpret_cast<const double*>(smem);
    double* pool = (AWG ? reinterpret_cast<double*>(smem) : reinterpret_cast<double*>(smem) + P.n_arcs + 1) + tid;   // slot s of this thread at pool[s*NT]
    if (!AWG) { double* t = reinterpret_cast<double*>(smem); for (int i = tid; i <= P.n_arcs; i += NT) t[i] = i < P.n_arcs ? P.aw[i] : 0.0; }
    __syncthreads();
    const long long gwarp = ((long long)blockIdx.x * NT + tid) >> 5;
    double* const xs = P.xs + (size_t)gwarp * P.xs_rows * 32 + lane;
    unsigned long long* const acc_g = P.acc + (size_t)(blockIdx.x % P.replicas) * (size_t)P.n_arcs;
    // groups are sorted by cost (DAG regions first, then path form by paths and length) and handed out dynamically, one
    // at a time; the id of the next group is requested while the current one is processed
    long long g = 0, ll = 0;
    if (lane == 0) g = (long long)atomicAdd(P.counter, 1u);
    g = __shfl_sync(FULL, g, 0);
    while (g < P.n_groups) {
        long long gn = 0;
        if (lane == 0) gn = (long long)atomicAdd(P.counter, 1u);
        const int rows = P.grows[g];
        const long long off = P.goff[g];
        if (rows & 0x10000) {                                  // path form: paths << 8 | length
            const int L = rows & 0xff;
            switch ((rows >> 8) & 0xff) {
                case 2: kr_paths<2, ACC>(P, aw, L, g, off, lane, acc_g, ll); break;
                case 3: kr_paths<3, ACC>(P, aw, L, g, off, lane, acc_g, ll); break;
                case 4: kr_paths<4, ACC>(P, aw, L, g, off, lane, acc_g, ll); break;
                case 6: kr_paths<6, ACC>(P, aw, L, g, off, lane, acc_g, ll); break;
                default: kr_paths<8, ACC>(P, aw, L, g, off, lane, acc_g, ll); break;
            }
        } else kr_big<ACC>(P, aw, pool, NT, g, lane, xs, acc_g, ll, &s_trash[tid]);
        g = __shfl_sync(FULL, gn, 0);
    }
    // the CTA's share of the log-likelihood: integer sums (exact, order independent), one RED per CTA
#pragma unroll
    for (int o = 16; o; o >>= 1) ll += __shfl_xor_sync(FULL, ll, o);
    __syncthreads();                                           // the pool is free now
    long long* part = reinterpret_cast<long long*>(pool - tid);
    if (lane == 0) part[tid >> 5] = ll;
    __syncthreads();
    if (tid == 0) {
        long long s = 0;
        for (int w = 0; w < (NT >> 5); ++w) s += part[w];
        unsigned long long nf = 0;
        if (blockIdx.x == 0)
            for (int k = 0; k < P.n_llpart; ++k) { s += P.llpart[2 * k]; nf += (unsigned long long)P.llpart[2 * k + 1]; }
        if (s) atomicAdd(P.red, (unsigned long long)s);
        if (nf) atomicAdd(P.red + 1, nf);
    }
}

// ------------------------------------------------------------------------------------------
struct KSParams {
    const double* __restrict__ logaw;     // [n_arcs] log weight of every combined arc
    const uint32_t* __restrict__ words;   // chunk-interleaved super-groups (lattice.hpp, SegmentedCorpus)
    const int64_t* __restrict__ sgoff;    // [n_sgroups+1]
    const int32_t* __restrict__ gref;     // [n_sgroups*16] leading rows of a group that hold region type ids
    const double* __restrict__ lq;        // log q per region type (written by kr_regions)
    const double* __restrict__ p;         // [n_sgroups*16*32] p_s in group order (0 = padding lane)
    double* logq;                         // [n_sgroups*16*32] log q_s in group order
    long long n_sgroups;
    unsigned int* counter;
    int n_arcs;
};

// KS: one CTA of 16 warps per super-group of 16 groups, one warp per group, one thread per string.
//   log q_s = sum over the string's bridge arcs of log w[arc]  (two 16-bit ids per word; 8-byte table in shared
//             memory; the host scheduled the ids so that a half-warp reads 16 different bank pairs)
//           + sum over its regions of lq[type]                  (gathers from L2)
// The words of a super-group are interleaved chunk by chunk, so the 16 warps stream ONE contiguous region
// (DRAM sees long bursts: 5.6 TB/s against 4.3 TB/s with one private block per warp, profiles/microbench_stream.cu).
// Measured dead ends, kept in the history: a cp.async (LDGSTS) ring and a TMA bulk-copy ring per warp were both
// slower than this register double buffer (the limit was DRAM burst locality, not bytes in flight).
constexpr int kKsRows = 8, kKsWarps = 16;      // = kKsChunkRows, kKsSuper of lattice.hpp

// AWG: the table does not fit shared memory and is read from HBM/L2 (P.logaw has n_arcs + 16 entries, the last 16 zero)
template <bool AWG = false>
__global__ void __launch_bounds__(kKsWarps * 32, 2) ks_strings(const KSParams P)
{
    extern __shared__ __align__(128) unsigned long long smem[];
    __shared__ long long s_next[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* tab = AWG ? P.logaw : reinterpret_cast<const double*>(smem);   // [n_arcs + 16]; ids n_arcs.. (padding, one per bank pair) -> 0
    if (!AWG) { double* t = reinterpret_cast<double*>(smem); for (int i = tid; i < P.n_arcs + 16; i += kKsWarps * 32) t[i] = i < P.n_arcs ? P.logaw[i] : 0.0; }
    if (tid == 0) s_next[0] = (long long)atomicAdd(P.counter, 1u);
    __syncthreads();
    int par = 0;
    for (long long sg = s_next[0]; sg < P.n_sgroups;) {
        if (tid == 0) s_next[par ^ 1] = (long long)atomicAdd(P.counter, 1u);   // the next super-group, fetched early
        const long long o = P.sgoff[sg];
        const int chunks = (int)((P.sgoff[sg + 1] - o) / (kKsWarps * kKsRows * 32));
        const long long g = sg * kKsWarps + warp;
        const int nref = P.gref[g];
        const double ps = P.p[g * 32 + lane];
        const uint32_t* wp = P.words + o + (size_t)warp * (kKsRows * 32) + lane;   // chunk c at wp + c * (16*8*32)
        constexpr size_t CS = (size_t)kKsWarps * kKsRows * 32;
        double s0 = 0.0, s1 = 0.0, r = 0.0, v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        uint32_t a[kKsRows];
#pragma unroll
        for (int j = 0; j < kKsRows; ++j) a[j] = __ldcs(wp + j * 32);
        for (int c = 0; c < chunks; ++c) {
            uint32_t b[kKsRows];
            if (c + 1 < chunks) {
#pragma unroll
                for (int j = 0; j < kKsRows; ++j) b[j] = __ldcs(wp + (size_t)(c + 1) * CS + j * 32);
            }
            if (c * kKsRows < nref) {                           // chunk with region references (the first one, rarely two)
#pragma unroll
                for (int j = 0; j < kKsRows; ++j) {
                    const int i = c * kKsRows + j;
                    if (i < nref) {
                        const double v = P.lq[a[j]];            // the first four are added at the end (latency overlaps)
                        if (i == 0) v0 = v; else if (i == 1) v1 = v; else if (i == 2) v2 = v; else if (i == 3) v3 = v; else r += v;
                    } else { s0 += tab[a[j] & 0xffffu]; s1 += tab[a[j] >> 16]; }
                }
            } else {
#pragma unroll
                for (int j = 0; j < kKsRows; ++j) { s0 += tab[a[j] & 0xffffu]; s1 += tab[a[j] >> 16]; }
            }
            if (c + 1 < chunks) {
#pragma unroll
                for (int j = 0; j < kKsRows; ++j) a[j] = b[j];
            }
        }
        const double lqs = (s0 + s1) + (r + ((v0 + v1) + (v2 + v3)));
        if (ps != 0.0) P.logq[g * 32 + lane] = lqs;
        __syncthreads();
        par ^= 1;
        sg = s_next[par];
    }
}

// no region types at all: the bridge partials are the whole log-likelihood
__global__ void k_add_llpart(const long long* __restrict__ llpart, int n, unsigned long long* red)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        long long s = 0, nf = 0;
        for (int k = 0; k < n; ++k) { s += llpart[2 * k]; nf += llpart[2 * k + 1]; }
        if (s) atomicAdd(red, (unsigned long long)s);
        if (nf) atomicAdd(red + 1, (unsigned long long)nf);
    }
}

// log q of the strings of the segmented path, group order -> string id order
__global__ void k_scatter_logq(long long n, const int32_t* __restrict__ ksid, const double* __restrict__ logq_k, double* logq)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && ksid[i] >= 0) logq[ksid[i]] = logq_k[i];
}

// The bridge part of the log-likelihood: sum over arcs of (sum_s p_s * times the arc is a bridge of s) * log w[arc].
// const_acc holds the first factor in the fixed point of the gradient accumulators (it is the constant part of the
// gradient as well).  Called by all threads of a CTA with the arc of the thread (c = 0 beyond the last
// arc); CTAs that hold arcs store their partial [value, non-finite terms] to llpart[blockIdx.x] (no reset needed).
__device__ __forceinline__ void bridge_loglik_partial(unsigned long long c, double l, double inv_fx, double ll_scale, long long* llpart)
{
    __shared__ long long s_part[32][2];
    long long v = 0, nf = 0;
    if (c) {
        if (isfinite(l)) v = __double2ll_rn((double)(long long)c * inv_fx * l * ll_scale);
        else nf = 1;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { v += __shfl_xor_sync(FULL, v, o); nf += __shfl_xor_sync(FULL, nf, o); }
    if ((threadIdx.x & 31) == 0) { s_part[threadIdx.x >> 5][0] = v; s_part[threadIdx.x >> 5][1] = nf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { v += s_part[w][0]; nf += s_part[w][1]; }
        llpart[2 * blockIdx.x] = v;
        llpart[2 * blockIdx.x + 1] = nf;
    }
}

// per combined arc: aw = a(u,v) * b(v,e) and its logarithm, straight from x
__global__ void __launch_bounds__(256) k_arc_weights_log(int n_arcs, const int32_t* __restrict__ arc_tid, const int32_t* __restrict__ arc_eid,
                                  const int32_t* __restrict__ trans_tp, const int32_t* __restrict__ emis_tp,
                                  const double* __restrict__ x, double* aw, double* logaw,
                                  const unsigned long long* __restrict__ const_acc, double inv_fx, double ll_scale, long long* llpart)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double l = 0.0;
    if (i < n_arcs) {
        l = logweight_of(trans_tp[arc_tid[i]], x, 0) + (arc_eid[i] < 0 ? 0.0 : logweight_of(emis_tp[arc_eid[i]], x, 0));
        logaw[i] = l;
        aw[i] = exp(l);
    }
    bridge_loglik_partial(i < n_arcs ? const_acc[i] : 0ull, l, inv_fx, ll_scale, llpart);
}


// Device-side barrier over the ranks of the communicator through peer memory (benchmark helper: it lines the ranks
// up between two timed evaluations).  One warp: lane r tells rank r "rank `rank` reached epoch e", then waits for
// rank r's word.  Words of epoch e live at bar[(e & 1) * nranks + sender].  Bounded wait (~2 s).
struct PeerBarrierParams {
    unsigned long long* peers[8];
    size_t bar_off;
    int nranks, rank;
    unsigned long long epoch;
};
__global__ void k_peer_barrier(const PeerBarrierParams P)
{
    const int lane = threadIdx.x;
    if (lane < P.nranks) {
        const size_t slot = P.bar_off + (size_t)(P.epoch & 1ull) * P.nranks;
        *reinterpret_cast<volatile unsigned long long*>(P.peers[lane] + slot + P.rank) = P.epoch;
        const volatile unsigned long long* src = P.peers[P.rank] + slot + lane;
        const long long t0 = clock64();
        while (*src != P.epoch)
            if (clock64() - t0 > 4000000000ll) break;
    }
}

}  // namespace wfsa
