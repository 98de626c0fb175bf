// w-fsa_b200/csrc/layout.hpp -- host-side lowering of a wfsa_fsa_desc to the device layout.
//
// The reference walks a graph of C-string-keyed nodes (/root/reference/inc/Fsa.h:26-66) and
// matches emission strings by prefix (inc/Recognize.h:49-57).  The device layout replaces that
// with dense edge ids and, for automata whose emissions are all exactly one token long, with
// "combined arcs" in CSR grouped by (state, symbol):
//
//   slot      = one (state v, symbol c) emission edge; the slots of symbol c are the
//               *candidates* a string position carrying c can be in (one lane / thread each)
//   fwd row   (v, c_prev)  -> { (slot index of predecessor u inside E[c_prev], transition id) }
//   bwd row   (u, c_next)  -> { (slot index of successor  v inside E[c_next], transition id) }
//               the position of an entry in the bwd table is the id of the combined arc
//               (u --a(u,v)--> v emits c_next); gradient accumulators are per combined arc.
//   START is a pseudo symbol (id n_symbols) whose only slot is the start state.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/wfsa_dev.h"

namespace wfsa {

struct HostFsa {
    int n_states = 0, start = 0, end = 0, n_sym = 0, n_raw = 0;
    std::vector<int32_t> emis_row, emis_tok_off, emis_tok, emis_param;
    std::vector<int32_t> trans_row, trans_dst, trans_param;
    int n_emis() const { return (int)emis_param.size(); }
    int n_trans() const { return (int)trans_dst.size(); }
    int emis_len(int e) const { return emis_tok_off[e + 1] - emis_tok_off[e]; }
};

// Limits of the packed 32-bit table entries.
constexpr int kSlotBits = 10;              // candidate index inside a symbol's list (< 1024)
constexpr int kMaxCand = 1 << kSlotBits;
constexpr int kMaxTid = 1 << (32 - kSlotBits);
constexpr int kRowCntBits = 8;             // entries per row (< 256)
constexpr uint32_t kMaxRowStart = 1u << (32 - kRowCntBits);

struct FastLayout {
    bool ok = false;                 // automaton is single-token ("token = symbol" fast path)
    int n_sym = 0, n_states = 0, n_slots = 0, n_arcs = 0, max_cand = 0, max_row = 0;
    std::vector<uint32_t> cand_off;  // [n_sym + 2]  (index n_sym = START)
    std::vector<uint32_t> slot_state;
    std::vector<int32_t> slot_emis;  // emission edge id, -1 for the START slot
    std::vector<int32_t> slot_final; // transition edge id state->end or -1
    std::vector<uint32_t> frow, fent;    // rows: n_states * (n_sym + 1)
    std::vector<uint32_t> brow, bent;    // rows: n_states * n_sym
    std::vector<int32_t> arc_tid;    // [n_arcs] transition edge of combined arc
    std::vector<int32_t> arc_eid;    // [n_arcs] emission edge of the arc's target slot
    int start_final_tid = -1;        // transition start->end (accepts the empty string)
    // Compact tables of the warp-per-string kernel (staged in shared memory): the bwd CSR above with
    // 16-bit row starts, 8-bit target slots and 16-bit slot->state; usable when at most 32 states
    // emit one symbol and the automaton has fewer than 65536 combined arcs.
    bool warp_ok = false;                // compact tables + at most 32 states emit one symbol
    bool compact_ok = false;             // fewer than 65536 arcs and 32767 states: 16-bit tables below exist
    std::vector<uint16_t> brow16;        // [n_states*n_sym + 1] row starts (arc ids are bwd-CSR positions)
    std::vector<uint8_t> bent_dst;       // [n_arcs] target slot inside E[c_next]
    std::vector<uint16_t> slot_state16;  // [n_slots]
    std::vector<uint16_t> arc_dst16;     // [n_arcs] target STATE (thread-per-string kernel)
    std::vector<int32_t> state_final;    // [n_states] transition edge state->end or -1
};

struct GenericLayout {
    // plain edge lists + an order of states in which empty-emission arcs only go forward
    std::vector<int32_t> eps_order;  // [n_states]
    int n_eps_states = 0;            // states that have an empty emission
    int max_emis_len = 0;
};

// returns "" on success, otherwise a message; `status` receives the wfsa_status
std::string copy_and_validate(const wfsa_fsa_desc* d, HostFsa& out, int& status);
std::string build_fast_layout(const HostFsa& f, FastLayout& out, int& status);
std::string build_generic_layout(const HostFsa& f, GenericLayout& out, int& status);

}  // namespace wfsa
