// w-fsa_b200/csrc/layout.cpp -- see layout.hpp.
#include "layout.hpp"

#include <algorithm>
#include <map>
#include <queue>
#include <set>

namespace wfsa {

static std::string fail(int& status, int code, const std::string& msg)
{
    status = code;
    return msg;
}

std::string copy_and_validate(const wfsa_fsa_desc* d, HostFsa& f, int& status)
{
    status = WFSA_OK;
    if (!d) return fail(status, WFSA_ERR_INVALID, "fsa descriptor is NULL");
    if (d->n_states < 2) return fail(status, WFSA_ERR_INVALID, "an automaton needs a start and an end state");
    if (d->start_state < 0 || d->start_state >= d->n_states || d->end_state < 0 || d->end_state >= d->n_states ||
        d->start_state == d->end_state)
        return fail(status, WFSA_ERR_INVALID, "start/end state ids out of range or equal");
    if (d->n_symbols < 0 || d->n_raw_params < 0) return fail(status, WFSA_ERR_INVALID, "negative sizes");
    if (!d->emis_row || !d->trans_row || !d->emis_tok_off) return fail(status, WFSA_ERR_INVALID, "NULL CSR arrays");
    f.n_states = d->n_states; f.start = d->start_state; f.end = d->end_state;
    f.n_sym = d->n_symbols; f.n_raw = d->n_raw_params;
    f.emis_row.assign(d->emis_row, d->emis_row + d->n_states + 1);
    f.trans_row.assign(d->trans_row, d->trans_row + d->n_states + 1);
    const int ne = f.emis_row.back(), nt = f.trans_row.back();
    if (f.emis_row[0] != 0 || f.trans_row[0] != 0 || ne < 0 || nt < 0)
        return fail(status, WFSA_ERR_INVALID, "CSR row arrays must start at 0");
    for (int s = 0; s < f.n_states; ++s)
        if (f.emis_row[s] > f.emis_row[s + 1] || f.trans_row[s] > f.trans_row[s + 1])
            return fail(status, WFSA_ERR_INVALID, "CSR row arrays must be non-decreasing");
    if ((ne && (!d->emis_param)) || (nt && (!d->trans_dst || !d->trans_param)))
        return fail(status, WFSA_ERR_INVALID, "NULL edge arrays");
    f.emis_tok_off.assign(d->emis_tok_off, d->emis_tok_off + ne + 1);
    if (f.emis_tok_off[0] != 0) return fail(status, WFSA_ERR_INVALID, "emis_tok_off must start at 0");
    for (int e = 0; e < ne; ++e)
        if (f.emis_tok_off[e] > f.emis_tok_off[e + 1]) return fail(status, WFSA_ERR_INVALID, "emis_tok_off must be non-decreasing");
    const int ntok = f.emis_tok_off.back();
    if (ntok && !d->emis_tok) return fail(status, WFSA_ERR_INVALID, "NULL emis_tok");
    f.emis_tok.assign(d->emis_tok, d->emis_tok + ntok);
    f.emis_param.assign(d->emis_param, d->emis_param + ne);
    f.trans_dst.assign(d->trans_dst, d->trans_dst + nt);
    f.trans_param.assign(d->trans_param, d->trans_param + nt);
    for (int t : f.emis_tok)
        if (t < 0 || t >= f.n_sym) return fail(status, WFSA_ERR_INVALID, "emission token out of [0, n_symbols)");
    std::vector<char> seen(f.n_raw, 0);
    auto check_param = [&](int p) -> bool {
        if (p == -1) return true;
        if (p < 0 || p >= f.n_raw || seen[p]) return false;
        seen[p] = 1;
        return true;
    };
    for (int p : f.emis_param) if (!check_param(p)) return fail(status, WFSA_ERR_INVALID, "emission parameter id invalid or used twice");
    for (int p : f.trans_param) if (!check_param(p)) return fail(status, WFSA_ERR_INVALID, "transition parameter id invalid or used twice");
    for (int s = 0; s < f.n_states; ++s) {
        std::set<int> dsts;
        for (int t = f.trans_row[s]; t < f.trans_row[s + 1]; ++t) {
            const int v = f.trans_dst[t];
            if (v < 0 || v >= f.n_states) return fail(status, WFSA_ERR_INVALID, "transition target out of range");
            // the reference rejects arcs into the start state at parse time (src/Fsa.cpp:175-178)
            if (v == f.start) return fail(status, WFSA_ERR_INVALID, "transition into the start state");
            if (!dsts.insert(v).second) return fail(status, WFSA_ERR_INVALID, "duplicate transition (src/Fsa.cpp:171-174)");
        }
        std::set<std::vector<int32_t>> es;
        for (int e = f.emis_row[s]; e < f.emis_row[s + 1]; ++e) {
            std::vector<int32_t> w(f.emis_tok.begin() + f.emis_tok_off[e], f.emis_tok.begin() + f.emis_tok_off[e + 1]);
            if (!es.insert(w).second) return fail(status, WFSA_ERR_INVALID, "duplicate emission (src/Fsa.cpp:145-147)");
        }
    }
    return "";
}

std::string build_fast_layout(const HostFsa& f, FastLayout& L, int& status)
{
    status = WFSA_OK;
    L = FastLayout();
    L.n_sym = f.n_sym; L.n_states = f.n_states;
    for (int s = 0; s < f.n_states; ++s) {
        if (s == f.start || s == f.end) continue;   // their emissions are never consumed
        for (int e = f.emis_row[s]; e < f.emis_row[s + 1]; ++e)
            if (f.emis_len(e) != 1) return "";        // not a fast-path automaton (ok = false)
    }
    const int S = f.n_states, A = f.n_sym;
    // slots grouped by symbol
    std::vector<std::vector<std::pair<int, int>>> by_sym(A);   // (state, emission edge)
    for (int s = 0; s < S; ++s) {
        if (s == f.start || s == f.end) continue;
        for (int e = f.emis_row[s]; e < f.emis_row[s + 1]; ++e) by_sym[f.emis_tok[f.emis_tok_off[e]]].push_back({s, e});
    }
    L.cand_off.assign(A + 2, 0);
    std::vector<std::vector<std::pair<int, int>>> slots_of_state(S);   // (symbol, index in E[symbol])
    for (int c = 0; c < A; ++c) {
        L.cand_off[c] = (uint32_t)L.slot_state.size();
        int j = 0;
        for (auto& se : by_sym[c]) {
            L.slot_state.push_back((uint32_t)se.first);
            L.slot_emis.push_back(se.second);
            slots_of_state[se.first].push_back({c, j++});
        }
        L.max_cand = std::max(L.max_cand, j);
    }
    L.cand_off[A] = (uint32_t)L.slot_state.size();
    L.slot_state.push_back((uint32_t)f.start);     // the START pseudo symbol
    L.slot_emis.push_back(-1);
    slots_of_state[f.start].push_back({A, 0});
    L.cand_off[A + 1] = (uint32_t)L.slot_state.size();
    L.n_slots = (int)L.slot_state.size();
    L.max_cand = std::max(L.max_cand, 1);
    if (L.max_cand > kMaxCand)
        return fail(status, WFSA_ERR_LIMIT, "more than 1024 states emit one symbol");
    if (f.n_trans() >= kMaxTid) return fail(status, WFSA_ERR_LIMIT, "too many transitions for the packed tables");

    // final transitions
    std::vector<int> final_tid(S, -1);
    for (int s = 0; s < S; ++s)
        for (int t = f.trans_row[s]; t < f.trans_row[s + 1]; ++t)
            if (f.trans_dst[t] == f.end) final_tid[s] = t;
    L.start_final_tid = final_tid[f.start];
    L.slot_final.resize(L.n_slots);
    for (int i = 0; i < L.n_slots; ++i) L.slot_final[i] = final_tid[L.slot_state[i]];
    L.slot_final[L.cand_off[A]] = -1;    // the START slot never ends a non-empty string

    // rows
    std::vector<std::vector<uint32_t>> frows((size_t)S * (A + 1)), brows((size_t)S * A);
    std::vector<std::vector<int32_t>> brow_eid((size_t)S * A);
    for (int u = 0; u < S; ++u) {
        if (u == f.end) continue;
        for (int t = f.trans_row[u]; t < f.trans_row[u + 1]; ++t) {
            const int v = f.trans_dst[t];
            if (v == f.end) continue;
            for (auto& sl : slots_of_state[u])         // predecessor slots (symbol c_prev, index j)
                frows[(size_t)v * (A + 1) + sl.first].push_back(((uint32_t)t << kSlotBits) | (uint32_t)sl.second);
            for (auto& sl : slots_of_state[v]) {       // successor slots (symbol c_next, index j)
                brows[(size_t)u * A + sl.first].push_back(((uint32_t)t << kSlotBits) | (uint32_t)sl.second);
                brow_eid[(size_t)u * A + sl.first].push_back(L.slot_emis[L.cand_off[sl.first] + sl.second]);
            }
        }
    }
    auto flatten = [&](std::vector<std::vector<uint32_t>>& rows, std::vector<uint32_t>& row, std::vector<uint32_t>& ent) -> bool {
        row.resize(rows.size());
        for (size_t r = 0; r < rows.size(); ++r) {
            if (rows[r].size() >= (1u << kRowCntBits) || ent.size() >= kMaxRowStart) return false;
            row[r] = ((uint32_t)ent.size() << kRowCntBits) | (uint32_t)rows[r].size();
            L.max_row = std::max(L.max_row, (int)rows[r].size());
            ent.insert(ent.end(), rows[r].begin(), rows[r].end());
        }
        return true;
    };
    if (!flatten(frows, L.frow, L.fent) || !flatten(brows, L.brow, L.bent))
        return fail(status, WFSA_ERR_LIMIT, "a (state,symbol) row has 256+ arcs or the arc table exceeds 16M entries");
    L.n_arcs = (int)L.bent.size();
    L.arc_tid.resize(L.n_arcs);
    L.arc_eid.reserve(L.n_arcs);
    for (auto& r : brow_eid) L.arc_eid.insert(L.arc_eid.end(), r.begin(), r.end());
    for (int a = 0; a < L.n_arcs; ++a) L.arc_tid[a] = (int32_t)(L.bent[a] >> kSlotBits);
    L.ok = true;
    L.state_final.assign(final_tid.begin(), final_tid.end());
    L.state_final[f.start] = -1;        // start->end only accepts the empty string (handled separately)
    // compact 16-bit tables for the thread-per-string and warp-per-string kernels
    if (L.n_arcs < 65536 && S < 32767) {
        L.brow16.resize(L.brow.size() + 1);
        for (size_t r = 0; r < L.brow.size(); ++r) L.brow16[r] = (uint16_t)(L.brow[r] >> kRowCntBits);
        L.brow16[L.brow.size()] = (uint16_t)L.n_arcs;
        L.arc_dst16.resize(L.n_arcs);
        {
            size_t a = 0;
            for (int u = 0; u < S; ++u)
                for (int c = 0; c < A; ++c) {
                    const uint32_t row = L.brow[(size_t)u * A + c];
                    const int cnt = row & ((1u << kRowCntBits) - 1);
                    for (int k = 0; k < cnt; ++k, ++a) {
                        const uint32_t slot = L.bent[a] & ((1u << kSlotBits) - 1);
                        L.arc_dst16[a] = (uint16_t)L.slot_state[L.cand_off[c] + slot];
                    }
                }
        }
        L.compact_ok = true;
        if (L.max_cand <= 32) {
            L.bent_dst.resize(L.n_arcs);
            for (int a = 0; a < L.n_arcs; ++a) L.bent_dst[a] = (uint8_t)(L.bent[a] & ((1u << kSlotBits) - 1));
            L.slot_state16.resize(L.n_slots);
            for (int i = 0; i < L.n_slots; ++i) L.slot_state16[i] = (uint16_t)L.slot_state[i];
            L.warp_ok = true;
        }
    }
    return "";
}

std::string build_generic_layout(const HostFsa& f, GenericLayout& G, int& status)
{
    status = WFSA_OK;
    G = GenericLayout();
    const int S = f.n_states;
    std::vector<char> has_eps(S, 0);
    for (int s = 0; s < S; ++s) {
        if (s == f.start || s == f.end) continue;
        for (int e = f.emis_row[s]; e < f.emis_row[s + 1]; ++e) {
            G.max_emis_len = std::max(G.max_emis_len, f.emis_len(e));
            if (f.emis_len(e) == 0) has_eps[s] = 1;
        }
    }
    for (int s = 0; s < S; ++s) G.n_eps_states += has_eps[s];
    // Kahn's algorithm on the sub-graph of arcs u -> v where v has an empty emission
    std::vector<int> indeg(S, 0);
    for (int u = 0; u < S; ++u)
        for (int t = f.trans_row[u]; t < f.trans_row[u + 1]; ++t)
            if (has_eps[f.trans_dst[t]]) ++indeg[f.trans_dst[t]];
    std::queue<int> q;
    for (int s = 0; s < S; ++s) if (!indeg[s]) q.push(s);
    while (!q.empty()) {
        const int u = q.front(); q.pop();
        G.eps_order.push_back(u);
        for (int t = f.trans_row[u]; t < f.trans_row[u + 1]; ++t) {
            const int v = f.trans_dst[t];
            if (has_eps[v] && --indeg[v] == 0) q.push(v);
        }
    }
    if ((int)G.eps_order.size() != S)
        return fail(status, WFSA_ERR_EPS_CYCLE,
                    "the automaton has a cycle of empty emissions: the sum over paths diverges "
                    "(the reference's DFS does not terminate on it, inc/Recognize.h:35-60)");
    return "";
}

}  // namespace wfsa
