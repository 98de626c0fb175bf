// w-fsa_b200/csrc/kernels_k7.cuh -- K7: position-synchronous, pair-batched forward-backward for DENSE automata
// (more than 32 states emit one symbol: BASELINE.json config 5, 4096 states / 256 symbols).
//
// K3 walks one string per CTA: every position costs a chain of dependent, uncoalesced look-ups
// (state -> row -> entries -> weights) that nothing amortises, 2 350 L2 sectors per position and sweep.  Here ALL strings of
// a batch advance one position per launch, and the strings of a step are grouped by the symbol pair (c_{t-1}, c_t) they
// sit on: one CTA takes up to kK7Chunk strings of ONE pair, loads the pair's sub-matrix of the transition table once
// (thread j = candidate j of c_t: its in-edges and their weights go to registers), and applies it to the alpha vectors
// of its strings, which stream through shared memory with coalesced loads and stores.  What is left per string and
// position is the alpha / beta vector itself (8 * V bytes in, 8 * V bytes out): the kernels run at HBM bandwidth, and the
// lattice of a whole batch (180 GB of HBM3e: ~0.5 M strings per batch for config 5) is kept for the backward sweep.
// The posteriors of a pair's arcs are summed over the strings of the chunk as fixed-point integers in registers: one RED
// per (arc, chunk) instead of one per (arc, string).
//
// Same recursion, same order of the sums and same power-of-two rescaling rule as K3 (kernels.cuh): alpha, beta and the
// rounded arc posteriors equal K3's; what the reference computes by path enumeration + sparse algebra
// (/root/reference/src/Learner.cpp:515-553, src/QuasiNewtonLearner.cpp:93-125).
#pragma once
#include "kernels.cuh"

namespace wfsa {

constexpr int kK7Chunk = 16;                 // strings of one pair per CTA (alpha tile in shared memory)
constexpr int kK7Reg = 8;                    // table entries of a thread kept in registers (more: re-read)

constexpr int kK7Planes = 8;                 // in-/out-edges of a candidate kept in the pair planes (more: read through the row word)
constexpr uint32_t kK7Sent = 0xfffffc00u;    // padding entry of a plane: transition field all ones, slot 0

// strings perm[start .. start+cnt) of this step sit on the pair (cp, c); cp = n_sym at position 0.  c0c / c0p: first slot of
// the candidates of c / cp; kf / kb: planes of the pair that hold entries (forward / backward tables)
struct K7Desc { int cp, c, start, cnt, c0c, c0p, kf, kb; };

struct K7Params {
    FastTablesD T;
    EvalWeightsD W;
    EvalOutD O;
    const K7Desc* __restrict__ desc;         // descriptors of the step
    const int32_t* __restrict__ perm;        // batch-local string indices of the step, grouped by pair
    const double* src;                       // fwd: alpha_{t-1} rows; bwd: beta~_{t+1} rows      [r][V]
    double* dst;                             // fwd: alpha_t rows;     bwd: beta~_t rows
    const double* __restrict__ lat;          // bwd: alpha_t rows
    const int* exp_src;                      // fwd: cumulative exponent at t-1 [r]
    int* exp_dst;                            // fwd: cumulative exponent at t   [r]
    const int* __restrict__ exp_t;           // bwd: cumulative exponent of alpha_t [r]
    int* F;                                  // bwd: cumulative exponent of beta~ [r] (updated in place)
    const int* __restrict__ EQ;              // bwd: exponent of q [r]
    const double* __restrict__ scale;        // bwd: p_s / q_s [r]
    int V, rescale;
    // Pair planes (k7_build_planes): everything thread j needs for the pair (cp, c), laid out so that the threads of a CTA
    // read consecutive words -- entry k of candidate j at pe[(key * kK7Planes + k) * V + j], its row word (start << 8 | count
    // in the fent / bent table) at pr[key * V + j], key = cp * n_sym + c.  One round trip instead of the chain
    // cand_off -> slot_state -> row -> entries of the per-state tables.
    const uint32_t* __restrict__ pe;
    const uint32_t* __restrict__ pr;
};

// Pair planes of one direction.  fwd: thread j = candidate j of c, row (state, cp) of the in-edge table; bwd: thread j =
// candidate j of cp, row (state, c) of the out-edge table.  One CTA of V threads per key; kmax[key] = planes in use.
__global__ void k7_build_planes(const FastTablesD T, int V, int fwd, uint32_t* pe, uint32_t* pr, int* kmax)
{
    __shared__ int s_max;
    const int A = T.n_sym, j = threadIdx.x;
    const size_t key = blockIdx.x;
    const int cp = (int)(key / A), c = (int)(key % A);
    if (j == 0) s_max = 0;
    __syncthreads();
    uint32_t row = 0;
    const uint32_t* ent = fwd ? T.fent : T.bent;
    if (fwd || cp < A) {
        const int sym = fwd ? c : cp;
        const uint32_t c0 = T.cand_off[sym], n = T.cand_off[sym + 1] - c0;
        if ((uint32_t)j < n) {
            const uint32_t st = T.slot_state[c0 + j];
            row = fwd ? T.frow[(size_t)st * (A + 1) + cp] : T.brow[(size_t)st * A + c];
        }
    }
    const int cnt = row & ((1u << kRowCntBitsD) - 1);
    const uint32_t st = row >> kRowCntBitsD;
    pr[key * V + j] = row;
    for (int k = 0; k < kK7Planes; ++k) pe[(key * kK7Planes + k) * V + j] = k < cnt ? ent[st + k] : kK7Sent;
    if (cnt) atomicMax(&s_max, min(cnt, kK7Planes));
    __syncthreads();
    if (j == 0) kmax[key] = s_max;
}

// the largest biased exponent of one value per thread over the CTA (-1: all zero)
__device__ __forceinline__ int k7_block_emax(double a, int* s_red, int nwarps, int warp, int lane)
{
    const int wmax = __reduce_max_sync(FULL, a != 0.0 ? biased_exp(a) : -1);
    __syncthreads();                                         // s_red of the previous string has been read
    if (lane == 0) s_red[warp] = wmax;
    __syncthreads();
    int m = -1;
    for (int w = 0; w < nwarps; ++w) m = max(m, s_red[w]);
    return m;
}

// forward step t: alpha_t[v] = b(v, c_t) * sum_u a(u,v) alpha_{t-1}[u] for every string of the chunk
__global__ void __launch_bounds__(512, 2) k7_fwd(const K7Params P)
{
    extern __shared__ double k7_tile[];                     // [kK7Chunk][V]
    __shared__ int s_red[32];
    const FastTablesD& T = P.T;
    const int V = P.V, j = threadIdx.x, lane = j & 31, warp = j >> 5, nwarps = V >> 5, A = T.n_sym;
    const K7Desc d = P.desc[blockIdx.x];
    const uint32_t c0 = T.cand_off[d.c], ncand = T.cand_off[d.c + 1] - c0;
    const bool valid = (uint32_t)j < ncand;
    double w[kK7Reg], swj = 0.0;
    int sidx[kK7Reg], cnt = 0;
    uint32_t st = 0;
    if (valid) {
        const uint32_t slot = c0 + j;
        const uint32_t row = T.frow[(size_t)T.slot_state[slot] * (A + 1) + d.cp];
        cnt = row & ((1u << kRowCntBitsD) - 1);
        st = row >> kRowCntBitsD;
#pragma unroll
        for (int k = 0; k < kK7Reg; ++k) {
            w[k] = 0.0; sidx[k] = 0;
            if (k < cnt) { const uint32_t ent = T.fent[st + k]; w[k] = P.W.tw[ent >> kSlotBitsD]; sidx[k] = ent & ((1u << kSlotBitsD) - 1); }
        }
        swj = P.W.sw[slot];
    }
    // per-string scalars once (a load of the string index and, behind it, of its exponent inside the loop over the strings
    // cost two dependent round trips PER STRING: 1.6 us per string in the first version)
    __shared__ int s_r[kK7Chunk], s_E[kK7Chunk];
    if (j < d.cnt) { const int r = P.perm[d.start + j]; s_r[j] = r; s_E[j] = d.cp == A ? 0 : P.exp_src[r]; }
    __syncthreads();
    // the alpha vectors of the chunk's strings: coalesced rows into shared memory, all loads in flight together
#pragma unroll 4
    for (int i = 0; i < d.cnt; ++i) k7_tile[i * V + j] = d.cp == A ? (j == 0 ? 1.0 : 0.0) : P.src[(size_t)s_r[i] * V + j];
    __syncthreads();
    for (int i = 0; i < d.cnt; ++i) {
        const int r = s_r[i];
        double v = 0.0;
        if (valid) {
            const double* sv = k7_tile + i * V;
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < kK7Reg; ++k) if (k < cnt) s = fma(w[k], sv[sidx[k]], s);
            for (int k = kK7Reg; k < cnt; ++k) { const uint32_t ent = T.fent[st + k]; s = fma(P.W.tw[ent >> kSlotBitsD], sv[ent & ((1u << kSlotBitsD) - 1)], s); }
            v = s * swj;
        }
        int E = s_E[i];
        if (P.rescale) {
            const int emax = k7_block_emax(v, s_red, nwarps, warp, lane);
            if (emax >= 0 && (emax < 1023 - kRescaleBand || emax > 1023 + kRescaleBand)) {
                const int shift = 1023 - emax;
                v = scalbn(v, shift);
                E -= shift;
            }
        }
        P.dst[(size_t)r * V + j] = v;
        if (j == 0) P.exp_dst[r] = E;
    }
}

// strings whose last position is t: q, log q, log-likelihood; scale[r] = p_s / q_s, EQ[r]; one warp per string.
// The sum over the candidates runs in K3's order (per group of 32 candidates a shuffle tree, then the groups in sequence).
struct K7FinParams {
    FastTablesD T; EvalWeightsD W; EvalOutD O;
    const int32_t* __restrict__ sid;         // [NB] string id of batch-local index r
    const int32_t* __restrict__ last_tok;    // [NB] token of the last position
    const double* __restrict__ p;            // [n_strings] p_s by string id
    const double* __restrict__ lat;          // alpha_t rows
    const int* __restrict__ exp_t;
    double* scale; int* EQ; int* F;
    double* bt;                              // part B: beta~_t rows to initialise
    int r0, r1, V;
};

__global__ void __launch_bounds__(256) k7_finish_q(const K7FinParams P)
{
    const int lane = threadIdx.x & 31;
    const int r = P.r0 + (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (r >= P.r1) return;
    const int c = P.last_tok[r], sid = P.sid[r];
    const uint32_t c0 = P.T.cand_off[c], ncand = P.T.cand_off[c + 1] - c0;
    double qh = 0.0;
    for (int b = 0; b < P.V; b += 32) {
        const int jj = b + lane;
        double term = 0.0;
        if ((uint32_t)jj < ncand) {
            const double al = P.lat[(size_t)r * P.V + jj];
            term = al * (al != 0.0 ? P.W.fw[c0 + jj] : 0.0);
        }
        qh += warp_sum(term);
    }
    if (lane == 0) {
        const double ps = P.p[sid];
        const int E = P.exp_t[r];
        if (!(qh > 0.0) || !isfinite(qh)) {
            if (P.O.logq) P.O.logq[sid] = -INFINITY;
            atomicAdd(P.O.red + 1, 1ull);
            P.scale[r] = 0.0; P.EQ[r] = 0;
        } else {
            const double lq = log(qh) + (double)E * 0.69314718055994530942;
            if (P.O.logq) P.O.logq[sid] = lq;
            atomicAdd(P.O.red, (unsigned long long)__double2ll_rn(ps * lq * P.O.ll_scale));
            P.scale[r] = (1.0 / qh) * ps;
            P.EQ[r] = E;
        }
    }
}

// the same strings at the start of their backward sweep: beta~_t = fin * b(v, c_t), posterior of the final transitions
__global__ void __launch_bounds__(256) k7_finish_beta(const K7FinParams P)
{
    const int lane = threadIdx.x & 31;
    const int r = P.r0 + (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (r >= P.r1) return;
    const int c = P.last_tok[r];
    const uint32_t c0 = P.T.cand_off[c], ncand = P.T.cand_off[c + 1] - c0;
    const double sc = P.scale[r] * P.O.fx_scale;            // (1/q) p_s 2^k; alpha_t and q share the exponent
    for (int jj = lane; jj < P.V; jj += 32) {
        double bt = 0.0;
        if ((uint32_t)jj < ncand && sc != 0.0) {
            const double al = P.lat[(size_t)r * P.V + jj];
            const double fin = al != 0.0 ? P.W.fw[c0 + jj] : 0.0;
            if (al != 0.0 && fin != 0.0) {
                const long long v = __double2ll_rn(al * fin * sc);
                if (v) atomicAdd(P.O.acc_global + P.T.n_arcs + P.T.slot_state[c0 + jj], (unsigned long long)v);
            }
            bt = al != 0.0 ? fin * P.W.sw[c0 + jj] : 0.0;
        }
        P.bt[(size_t)r * P.V + jj] = bt;
    }
    if (lane == 0) P.F[r] = 0;
}

// backward step t over the descriptors of forward step t+1 (pair (c_t, c_{t+1})): thread j = candidate j of c_t,
//   beta~_t[u] = b(u, c_t) * sum_v a(u,v) beta~_{t+1}[v],   acc[arc] += alpha_t[u] a(u,v) beta~_{t+1}[v] p_s / q_s
__global__ void __launch_bounds__(512, 2) k7_bwd(const K7Params P)
{
    extern __shared__ double k7_tile[];
    __shared__ int s_red[32];
    const FastTablesD& T = P.T;
    const int V = P.V, j = threadIdx.x, lane = j & 31, warp = j >> 5, nwarps = V >> 5, A = T.n_sym;
    const K7Desc d = P.desc[blockIdx.x];
    const bool from_start = d.cp == A;                      // t = -1: the arcs out of the start state, thread j = entry j of its row
    const uint32_t c0 = T.cand_off[d.cp], ncand = T.cand_off[d.cp + 1] - c0;
    bool valid = (uint32_t)j < ncand;
    double w[kK7Reg], swj = 0.0;
    int didx[kK7Reg], cnt = 0;
    uint32_t st = 0;
#pragma unroll
    for (int k = 0; k < kK7Reg; ++k) { w[k] = 0.0; didx[k] = 0; }
    if (from_start) {
        const uint32_t row = T.brow[(size_t)T.start_state * A + d.c];
        valid = j < (int)(row & ((1u << kRowCntBitsD) - 1));
        if (valid) {
            st = (row >> kRowCntBitsD) + j; cnt = 1;
            const uint32_t ent = T.bent[st];
            w[0] = P.W.tw[ent >> kSlotBitsD]; didx[0] = ent & ((1u << kSlotBitsD) - 1);
        }
    } else if (valid) {
        const uint32_t slot = c0 + j;
        const uint32_t row = T.brow[(size_t)T.slot_state[slot] * A + d.c];
        cnt = row & ((1u << kRowCntBitsD) - 1);
        st = row >> kRowCntBitsD;
#pragma unroll
        for (int k = 0; k < kK7Reg; ++k)
            if (k < cnt) { const uint32_t ent = T.bent[st + k]; w[k] = P.W.tw[ent >> kSlotBitsD]; didx[k] = ent & ((1u << kSlotBitsD) - 1); }
        swj = P.W.sw[slot];
    }
    __shared__ int s_r[kK7Chunk];
    __shared__ double s_sc[kK7Chunk];                        // p_s / q_s * 2^k * 2^(E_t + F_{t+1} - EQ) per string
    if (j < d.cnt) {
        const int r = P.perm[d.start + j];
        s_r[j] = r;
        const int de = (from_start ? 0 : P.exp_t[r]) + P.F[r] - P.EQ[r];
        double sc = P.scale[r] * P.O.fx_scale;
        if (de != 0) sc = scalbn(sc, de);
        s_sc[j] = sc;
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < d.cnt; ++i) k7_tile[i * V + j] = P.src[(size_t)s_r[i] * V + j];
    long long lacc[kK7Reg];
#pragma unroll
    for (int k = 0; k < kK7Reg; ++k) lacc[k] = 0;
    double al_next = (from_start || d.cnt == 0) ? 1.0 : P.lat[(size_t)s_r[0] * V + j];      // alpha_t of the next string, one ahead
    __syncthreads();
    for (int i = 0; i < d.cnt; ++i) {
        const int r = s_r[i];
        const double al = al_next;
        if (!from_start && i + 1 < d.cnt) al_next = P.lat[(size_t)s_r[i + 1] * V + j];
        double bt = 0.0;
        if (valid && al != 0.0) {
            const double sc = s_sc[i];
            const double* sv = k7_tile + i * V;
            double b = 0.0;
#pragma unroll
            for (int k = 0; k < kK7Reg; ++k)
                if (k < cnt) {
                    const double term = w[k] * sv[didx[k]];
                    b += term;
                    if (term != 0.0) lacc[k] += __double2ll_rn(al * term * sc);
                }
            for (int k = kK7Reg; k < cnt; ++k) {
                const uint32_t ent = T.bent[st + k];
                const double term = P.W.tw[ent >> kSlotBitsD] * sv[ent & ((1u << kSlotBitsD) - 1)];
                b += term;
                if (term != 0.0) { const long long v = __double2ll_rn(al * term * sc); if (v) atomicAdd(P.O.acc_global + st + k, (unsigned long long)v); }
            }
            bt = b * swj;
        }
        if (from_start) continue;                            // no beta~ before the first position
        if (P.rescale) {
            const int emax = k7_block_emax(bt, s_red, nwarps, warp, lane);      // (F[r] was read in the prologue)
            if (emax >= 0 && (emax < 1023 - kRescaleBand || emax > 1023 + kRescaleBand)) {
                const int shift = 1023 - emax;
                bt = scalbn(bt, shift);
                if (j == 0) P.F[r] -= shift;
            }
        }
        P.dst[(size_t)r * V + j] = bt;
    }
    if (valid) {
#pragma unroll
        for (int k = 0; k < kK7Reg; ++k) if (lacc[k]) atomicAdd(P.O.acc_global + st + k, (unsigned long long)lacc[k]);
    }
}


// ---- the same two steps on the pair planes ------------------------------------------------------------------------------
// What changed against k7_fwd / k7_bwd (same arithmetic, same order of every sum): the table entries of thread j come from
// the pair planes in ONE coalesced round trip; the alpha / beta rows of the chunk go to shared memory with cp.async while
// the weights are gathered; shared-memory addresses are byte offsets computed once; the exponent check of a rescale step
// takes one pass over the chunk and a fix-up pass for the (rare) strings that leave the band, instead of two block barriers
// per string; the alpha_t values of the backward step are fetched four strings ahead.

__device__ __forceinline__ void k7_cp_async8(void* smem_dst, const void* gsrc)
{
    const unsigned int sa = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ double k7_lds(const char* base, unsigned int off)
{
    return *reinterpret_cast<const double*>(base + off);
}

template <int MAXV, int MINB>
__global__ void __launch_bounds__(MAXV, MINB) k7_fwd2(const K7Params P)
{
    extern __shared__ double k7_tile[];                     // [kK7Chunk][V]
    __shared__ int s_r[kK7Chunk], s_E[kK7Chunk], s_emax[kK7Chunk];
    const FastTablesD& T = P.T;
    const int V = P.V, j = threadIdx.x, lane = j & 31, A = T.n_sym;
    const K7Desc d = P.desc[blockIdx.x];
    const size_t key = (size_t)d.cp * A + d.c;
    if (j < d.cnt) { s_r[j] = P.perm[d.start + j]; s_emax[j] = -1; }
    uint32_t ent[kK7Planes];
#pragma unroll
    for (int k = 0; k < kK7Planes; ++k) ent[k] = k < d.kf ? P.pe[(key * kK7Planes + k) * V + j] : kK7Sent;
    const uint32_t row = P.pr[key * V + j];
    const double swj = P.W.sw[min(d.c0c + j, T.n_slots - 1)];
    __syncthreads();
    if (d.cp == A) {
        for (int i = 0; i < d.cnt; ++i) k7_tile[i * V + j] = j == 0 ? 1.0 : 0.0;
    } else {
        for (int i = 0; i < d.cnt; ++i) k7_cp_async8(k7_tile + i * V + j, P.src + (size_t)s_r[i] * V + j);
    }
    if (j < d.cnt) s_E[j] = d.cp == A ? 0 : P.exp_src[s_r[j]];
    double w[kK7Planes];
    unsigned int off[kK7Planes];
#pragma unroll
    for (int k = 0; k < kK7Planes; ++k) {
        w[k] = (ent[k] >> kSlotBitsD) != (kK7Sent >> kSlotBitsD) ? P.W.tw[ent[k] >> kSlotBitsD] : 0.0;
        off[k] = (ent[k] & ((1u << kSlotBitsD) - 1)) * 8u;
    }
    const int cnt = row & ((1u << kRowCntBitsD) - 1);
    const uint32_t st = row >> kRowCntBitsD;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    const char* tile = reinterpret_cast<const char*>(k7_tile);
    const unsigned int rowb = (unsigned int)V * 8u;
    for (int i = 0; i < d.cnt; ++i) {
        const char* sv = tile + (unsigned int)i * rowb;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kK7Planes; ++k) s = fma(w[k], k7_lds(sv, off[k]), s);
        for (int k = kK7Planes; k < cnt; ++k) {
            const uint32_t e = T.fent[st + k];
            s = fma(P.W.tw[e >> kSlotBitsD], k7_lds(sv, (e & ((1u << kSlotBitsD) - 1)) * 8u), s);
        }
        const double v = s * swj;
        if (P.rescale) {
            const int wmax = __reduce_max_sync(FULL, v != 0.0 ? biased_exp(v) : -1);
            if (lane == 0 && wmax >= 0) atomicMax(&s_emax[i], wmax);
        }
        P.dst[(size_t)s_r[i] * V + j] = v;
    }
    if (P.rescale) {
        __syncthreads();
        for (int i = 0; i < d.cnt; ++i) {
            const int emax = s_emax[i];
            if (emax >= 0 && (emax < 1023 - kRescaleBand || emax > 1023 + kRescaleBand)) {
                double* q = P.dst + (size_t)s_r[i] * V + j;
                *q = scalbn(*q, 1023 - emax);                 // this thread's own store of the loop above
                if (j == i) s_E[i] -= 1023 - emax;
            }
        }
    }
    if (j < d.cnt) P.exp_dst[s_r[j]] = s_E[j];
}

template <int MAXV, int MINB>
__global__ void __launch_bounds__(MAXV, MINB) k7_bwd2(const K7Params P)
{
    extern __shared__ double k7_tile[];
    __shared__ int s_r[kK7Chunk], s_emax[kK7Chunk];
    __shared__ double s_sc[kK7Chunk];                        // p_s / q_s * 2^k * 2^(E_t + F_{t+1} - EQ) per string
    const FastTablesD& T = P.T;
    const int V = P.V, j = threadIdx.x, lane = j & 31, A = T.n_sym;
    const K7Desc d = P.desc[blockIdx.x];                     // (cp < n_sym: the step out of the start state runs k7_bwd)
    const size_t key = (size_t)d.cp * A + d.c;
    if (j < d.cnt) { s_r[j] = P.perm[d.start + j]; s_emax[j] = -1; }
    uint32_t ent[kK7Planes];
#pragma unroll
    for (int k = 0; k < kK7Planes; ++k) ent[k] = k < d.kb ? P.pe[(key * kK7Planes + k) * V + j] : kK7Sent;
    const uint32_t row = P.pr[key * V + j];
    const double swj = P.W.sw[min(d.c0p + j, T.n_slots - 1)];
    __syncthreads();
    for (int i = 0; i < d.cnt; ++i) k7_cp_async8(k7_tile + i * V + j, P.src + (size_t)s_r[i] * V + j);
    constexpr int PF = 2;                                    // strings whose alpha_t value is fetched ahead
    double aln[PF];
#pragma unroll
    for (int ii = 0; ii < PF; ++ii) aln[ii] = ii < d.cnt ? P.lat[(size_t)s_r[ii] * V + j] : 0.0;
    if (j < d.cnt) {
        const int r = s_r[j];
        const int de = P.exp_t[r] + P.F[r] - P.EQ[r];
        double sc = P.scale[r] * P.O.fx_scale;
        if (de != 0) sc = scalbn(sc, de);
        s_sc[j] = sc;
    }
    double w[kK7Planes];
    unsigned int off2[kK7Planes / 2];                        // two 16-bit byte offsets per register
#pragma unroll
    for (int k = 0; k < kK7Planes; ++k) w[k] = (ent[k] >> kSlotBitsD) != (kK7Sent >> kSlotBitsD) ? P.W.tw[ent[k] >> kSlotBitsD] : 0.0;
#pragma unroll
    for (int k = 0; k < kK7Planes; k += 2)
        off2[k / 2] = ((ent[k] & ((1u << kSlotBitsD) - 1)) * 8u) | (((ent[k + 1] & ((1u << kSlotBitsD) - 1)) * 8u) << 16);
    const int cnt = row & ((1u << kRowCntBitsD) - 1);
    const uint32_t st = row >> kRowCntBitsD;
    long long lacc[kK7Planes];
#pragma unroll
    for (int k = 0; k < kK7Planes; ++k) lacc[k] = 0;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    const char* tile = reinterpret_cast<const char*>(k7_tile);
    const unsigned int rowb = (unsigned int)V * 8u;
    for (int i0 = 0; i0 < d.cnt; i0 += PF) {
        double al[PF];
#pragma unroll
        for (int ii = 0; ii < PF; ++ii) al[ii] = aln[ii];
#pragma unroll
        for (int ii = 0; ii < PF; ++ii) if (i0 + PF + ii < d.cnt) aln[ii] = P.lat[(size_t)s_r[i0 + PF + ii] * V + j];
#pragma unroll
        for (int ii = 0; ii < PF; ++ii) {
            const int i = i0 + ii;
            if (i >= d.cnt) break;
            double bt = 0.0;
            if (al[ii] != 0.0) {
                const double sc = s_sc[i];
                const char* sv = tile + (unsigned int)i * rowb;
                double b = 0.0;
#pragma unroll
                for (int k = 0; k < kK7Planes; ++k) {
                    const double term = w[k] * k7_lds(sv, (k & 1) ? off2[k / 2] >> 16 : off2[k / 2] & 0xffffu);
                    b += term;
                    lacc[k] += __double2ll_rn(al[ii] * term * sc);
                }
                for (int k = kK7Planes; k < cnt; ++k) {
                    const uint32_t e = T.bent[st + k];
                    const double term = P.W.tw[e >> kSlotBitsD] * k7_lds(sv, (e & ((1u << kSlotBitsD) - 1)) * 8u);
                    b += term;
                    if (term != 0.0) { const long long v = __double2ll_rn(al[ii] * term * sc); if (v) atomicAdd(P.O.acc_global + st + k, (unsigned long long)v); }
                }
                bt = b * swj;
            }
            if (P.rescale) {
                const int wmax = __reduce_max_sync(FULL, bt != 0.0 ? biased_exp(bt) : -1);
                if (lane == 0 && wmax >= 0) atomicMax(&s_emax[i], wmax);
            }
            P.dst[(size_t)s_r[i] * V + j] = bt;
        }
    }
    if (P.rescale) {
        __syncthreads();
        for (int i = 0; i < d.cnt; ++i) {
            const int emax = s_emax[i];
            if (emax >= 0 && (emax < 1023 - kRescaleBand || emax > 1023 + kRescaleBand)) {
                double* q = P.dst + (size_t)s_r[i] * V + j;
                *q = scalbn(*q, 1023 - emax);
                if (j == i) P.F[s_r[i]] -= 1023 - emax;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kK7Planes; ++k) if (lacc[k]) atomicAdd(P.O.acc_global + st + k, (unsigned long long)lacc[k]);
}

}  // namespace wfsa
