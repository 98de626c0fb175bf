// w-fsa_b200/csrc/lattice.cpp -- see lattice.hpp.
#include "lattice.hpp"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <thread>

namespace wfsa {

void build_lattice_arcs(const HostFsa& f, const GenericLayout& g, LatticeArcs& A)
{
    A = LatticeArcs();
    const int S = f.n_states, NS = f.n_sym;
    A.n_sym = NS; A.n_states = S;
    A.final_arc.assign(S, -1);
    A.eps_rank.assign(S, 0);
    for (int i = 0; i < (int)g.eps_order.size(); ++i) A.eps_rank[g.eps_order[i]] = i;
    A.layered = true;
    for (int s = 0; s < S; ++s) {
        if (s == f.start || s == f.end) continue;
        for (int e = f.emis_row[s]; e < f.emis_row[s + 1]; ++e) {
            if (f.emis_len(e) == 0) A.has_eps = true;
            if (f.emis_len(e) != 1) A.layered = false;
        }
    }
    // arcs in (source, transition, emission) order; the compile index is a CSR over (source, first token)
    std::vector<int32_t> key;          // row of every non-final arc
    std::vector<int32_t> dst;
    for (int u = 0; u < S; ++u) {
        if (u == f.end) continue;
        for (int t = f.trans_row[u]; t < f.trans_row[u + 1]; ++t) {
            const int v = f.trans_dst[t];
            if (v == f.end) {
                A.final_arc[u] = (int32_t)A.arc_tid.size();
                A.arc_tid.push_back(t); A.arc_eid.push_back(-1);
                key.push_back(-1); dst.push_back(-1);
                continue;
            }
            for (int e = f.emis_row[v]; e < f.emis_row[v + 1]; ++e) {
                const int c = f.emis_len(e) == 0 ? NS : f.emis_tok[f.emis_tok_off[e]];
                A.arc_tid.push_back(t); A.arc_eid.push_back(e);
                key.push_back(u * (NS + 1) + c); dst.push_back(v);
            }
        }
    }
    A.n_arcs = (int)A.arc_tid.size();
    const size_t n_rows = (size_t)S * (NS + 1);
    A.row.assign(n_rows + 1, 0);
    for (int a = 0; a < A.n_arcs; ++a) if (key[a] >= 0) A.row[key[a] + 1]++;
    for (size_t r = 0; r < n_rows; ++r) A.row[r + 1] += A.row[r];
    const int n_ent = A.row[n_rows];
    A.ent_arc.resize(n_ent); A.ent_dst.resize(n_ent); A.ent_eid.resize(n_ent);
    std::vector<int32_t> fill(A.row.begin(), A.row.end() - 1);
    for (int a = 0; a < A.n_arcs; ++a) {
        if (key[a] < 0) continue;
        const int i = fill[key[a]]++;
        A.ent_arc[i] = a; A.ent_dst[i] = dst[a]; A.ent_eid[i] = A.arc_eid[a];
    }
}

int compile_lattice(const HostFsa& f, const LatticeArcs& A, const uint8_t* alive, const int32_t* tok, int len,
                    int n_slots, LatticeScratch& S, std::vector<uint32_t>& words, std::vector<int32_t>& bridge_arcs)
{
    const int NSt = f.n_states, NS = f.n_sym;
    const size_t need = (size_t)(len + 1) * NSt;
    if (S.node_of.size() < need) S.node_of.resize(need, -1);
    if ((int)S.bucket.size() < len + 1) S.bucket.resize(len + 1);
    S.npos.clear(); S.nstate.clear(); S.esrc.clear(); S.edst.clear(); S.earc.clear();
    int end_node = -1;
    auto get_node = [&](int pos, int st) -> int {
        int32_t& slot = S.node_of[(size_t)pos * NSt + st];
        if (slot < 0) {
            slot = (int32_t)S.npos.size();
            S.npos.push_back(pos); S.nstate.push_back(st);
            S.bucket[pos].push_back(slot);
        }
        return slot;
    };
    auto process = [&](int n) {
        const int pos = S.npos[n], u = S.nstate[n];
        if (pos < len) {
            const int c = tok[pos];
            if (c >= 0 && c < NS) {
                const size_t r = (size_t)u * (NS + 1) + c;
                for (int i = A.row[r]; i < A.row[r + 1]; ++i) {
                    const int a = A.ent_arc[i];
                    if (alive && !alive[a]) continue;
                    const int e = A.ent_eid[i];
                    const int e0 = f.emis_tok_off[e], el = f.emis_tok_off[e + 1] - e0;
                    if (el > 1) {
                        if (pos + el > len) continue;
                        bool m = true;
                        for (int k = 1; k < el; ++k) if (tok[pos + k] != f.emis_tok[e0 + k]) { m = false; break; }
                        if (!m) continue;
                    }
                    const int d = get_node(pos + el, A.ent_dst[i]);
                    S.esrc.push_back(n); S.edst.push_back(d); S.earc.push_back(a);
                }
            }
        }
        if (A.has_eps) {
            const size_t r = (size_t)u * (NS + 1) + NS;
            for (int i = A.row[r]; i < A.row[r + 1]; ++i) {
                const int a = A.ent_arc[i];
                if (alive && !alive[a]) continue;
                const int d = get_node(pos, A.ent_dst[i]);       // later in eps order than u
                S.esrc.push_back(n); S.edst.push_back(d); S.earc.push_back(a);
            }
        }
        if (pos == len && A.final_arc[u] >= 0 && (!alive || alive[A.final_arc[u]])) {
            if (end_node < 0) {
                end_node = (int)S.npos.size();
                S.npos.push_back(len + 1); S.nstate.push_back(f.end);
            }
            S.esrc.push_back(n); S.edst.push_back(end_node); S.earc.push_back(A.final_arc[u]);
        }
    };
    // ---- forward reachability, nodes processed in a topological order
    get_node(0, f.start);
    for (int pos = 0; pos <= len; ++pos) {
        std::vector<int32_t>& B = S.bucket[pos];
        if (!A.has_eps) {
            for (size_t bi = 0; bi < B.size(); ++bi) process(B[bi]);
        } else {
            S.done.assign(B.size(), 0);
            for (;;) {
                int best = -1;
                if (S.done.size() < B.size()) S.done.resize(B.size(), 0);
                for (size_t bi = 0; bi < B.size(); ++bi)
                    if (!S.done[bi] && (best < 0 || A.eps_rank[S.nstate[B[bi]]] < A.eps_rank[S.nstate[B[best]]])) best = (int)bi;
                if (best < 0) break;
                S.done[best] = 1;
                process(B[best]);
            }
        }
    }
    const int n_nodes = (int)S.npos.size(), n_e = (int)S.esrc.size();
    // reset the scratch index for the next string
    for (int n = 0; n < n_nodes; ++n) if (S.npos[n] <= len) S.node_of[(size_t)S.npos[n] * NSt + S.nstate[n]] = -1;
    for (int pos = 0; pos <= len; ++pos) S.bucket[pos].clear();
    if (end_node < 0) return 0;
    // ---- co-reachability (edges are ordered by source in topological order => reverse sweep)
    S.coreach.assign(n_nodes, 0);
    S.coreach[end_node] = 1;
    for (int e = n_e - 1; e >= 0; --e) if (S.coreach[S.edst[e]]) S.coreach[S.esrc[e]] = 1;
    if (!S.coreach[0]) return 0;
    S.nout.assign(n_nodes, 0);
    S.per_pos.assign(len + 2, 0);
    for (int n = 0; n < n_nodes; ++n) if (S.coreach[n]) S.per_pos[S.npos[n]]++;
    for (int e = 0; e < n_e; ++e) if (S.coreach[S.edst[e]]) S.nout[S.esrc[e]]++;
    // ---- emit the stream
    const size_t base = words.size(), bridge_base = bridge_arcs.size();
    S.nslot.assign(n_nodes, -1);
    uint32_t free_mask = n_slots >= 32 ? 0xffffffffu : ((1u << n_slots) - 1u), live = 0;
    auto alloc = [&]() -> int {
        if (!free_mask) return -1;
        const int s = __builtin_ctz(free_mask);
        free_mask &= free_mask - 1; live |= 1u << s;
        return s;
    };
    auto release = [&](int s) { free_mask |= 1u << s; live &= ~(1u << s); };
    S.nslot[0] = alloc();
    int cur_src = -1;
    size_t last_edge = 0;
    auto check_word = [&]() { if (((words.size() - base) & (kCheckEvery - 1)) == kCheckEvery - 1) words.push_back(live & 0xffffu); };
    for (int e = 0; e < n_e; ++e) {
        const int d = S.edst[e];
        if (!S.coreach[d]) continue;
        const int s = S.esrc[e];
        if (s != cur_src) {
            if (cur_src >= 0) { words[last_edge] |= kLatLastOut; release(S.nslot[cur_src]); }
            cur_src = s;
        }
        check_word();
        uint32_t w = kLatEdge | (uint32_t)S.earc[e];
        if (S.nslot[d] < 0) {
            S.nslot[d] = alloc();
            if (S.nslot[d] < 0) { words.resize(base); bridge_arcs.resize(bridge_base); return -1; }
            w |= kLatFirstIn;
        }
        if (A.layered && S.per_pos[S.npos[s]] == 1 && S.nout[s] == 1) { w |= kLatBridge; bridge_arcs.push_back(S.earc[e]); }
        w |= (uint32_t)S.nslot[d] << kLatDstShift | (uint32_t)S.nslot[s] << kLatSrcShift;
        last_edge = words.size();
        words.push_back(w);
    }
    words[last_edge] |= kLatLastOut;
    release(S.nslot[cur_src]);
    check_word();
    words.push_back(kLatFin | (uint32_t)S.nslot[end_node]);
    return 1;
}

void compile_corpus(const HostFsa& f, const LatticeArcs& A, const uint8_t* alive, const int32_t* tokens,
                    const int64_t* offs, const double* p, const std::vector<int32_t>& ids, int n_slots, double fx_scale,
                    bool use_bridges, int max_stream_words, CompiledCorpus& out)
{
    out = CompiledCorpus();
    const size_t n = ids.size();
    out.const_acc.assign(A.n_arcs, 0);
    unsigned hw = std::thread::hardware_concurrency();
    const int T = (int)std::max<size_t>(1, std::min<size_t>({(size_t)(hw ? hw : 4), (size_t)64, n / 256 + 1}));
    struct Local {
        std::vector<uint32_t> words;
        std::vector<int64_t> off;            // start of string (i*T + t) in words
        std::vector<int32_t> cnt;            // words, 0 rejected, -1 overflow
        std::vector<long long> cacc;
        int64_t n_edges = 0, n_bridge = 0;
    };
    std::vector<Local> loc(T);
    LatticeArcs Anb;                          // copy with bridges disabled when requested
    const LatticeArcs* AP = &A;
    if (!use_bridges && A.layered) { Anb = A; Anb.layered = false; AP = &Anb; }
    auto work = [&](int t) {
        Local& L = loc[t];
        LatticeScratch S;
        std::vector<int32_t> bridges;
        L.cacc.assign(A.n_arcs, 0);
        for (size_t i = t; i < n; i += T) {
            const int32_t sid = ids[i];
            const int len = (int)(offs[sid + 1] - offs[sid]);
            bridges.clear();
            L.off.push_back((int64_t)L.words.size());
            const size_t before = L.words.size();
            int rc = compile_lattice(f, *AP, alive, tokens + offs[sid], len, n_slots, S, L.words, bridges);
            if (rc == 1 && L.words.size() - before > (size_t)max_stream_words) { rc = -1; L.words.resize(before); }
            if (rc == 1) {
                L.cnt.push_back((int32_t)(L.words.size() - before));
                const long long c = llrint(p[sid] * fx_scale);
                for (int32_t a : bridges) L.cacc[a] += c;
                L.n_bridge += (int64_t)bridges.size();
                for (size_t k = before; k < L.words.size(); ++k) L.n_edges += (L.words[k] >> 31);
            } else L.cnt.push_back(rc);
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < T; ++t) {
        for (int a = 0; a < A.n_arcs; ++a) out.const_acc[a] += loc[t].cacc[a];
        out.n_edges += loc[t].n_edges; out.n_bridge += loc[t].n_bridge;
    }
    // strings with a stream, longest stream first (ties in input order => deterministic)
    std::vector<int64_t> ok;                 // indices into ids
    for (size_t i = 0; i < n; ++i) {
        const int c = loc[i % T].cnt[i / T];
        if (c > 0) ok.push_back((int64_t)i);
        else if (c < 0) out.overflow.push_back(ids[i]);
        else out.rejected.push_back(ids[i]);
    }
    std::stable_sort(ok.begin(), ok.end(), [&](int64_t a, int64_t b) { return loc[a % T].cnt[a / T] > loc[b % T].cnt[b / T]; });
    const int64_t groups = ((int64_t)ok.size() + 31) / 32;
    out.goff.assign(groups + 1, 0);
    out.gsid.assign((size_t)groups * 32, -1);
    for (int64_t g = 0; g < groups; ++g) {
        const int64_t a = ok[g * 32];
        int64_t mx = loc[a % T].cnt[a / T];                       // sorted: the first of the group is the longest
        mx = (mx + kCheckEvery - 1) / kCheckEvery * kCheckEvery;
        out.goff[g + 1] = out.goff[g] + mx * 32;
        out.max_words = std::max(out.max_words, mx);
    }
    out.n_words = out.goff[groups];
    out.words.assign((size_t)out.n_words + 32, 0u);
    auto fill = [&](int t) {
        for (int64_t g = t; g < groups; g += T) {
            uint32_t* dst = out.words.data() + out.goff[g];
            for (int l = 0; l < 32 && g * 32 + l < (int64_t)ok.size(); ++l) {
                const int64_t i = ok[g * 32 + l];
                const Local& L = loc[i % T];
                const uint32_t* src = L.words.data() + L.off[i / T];
                const int c = L.cnt[i / T];
                for (int k = 0; k < c; ++k) dst[(size_t)k * 32 + l] = src[k];
                out.gsid[(size_t)g * 32 + l] = ids[i];
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(fill, t);
        fill(0);
        for (auto& x : th) x.join();
    }
}

}  // namespace wfsa
