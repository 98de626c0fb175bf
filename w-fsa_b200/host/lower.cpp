// w-fsa_b200/host/lower.cpp -- see lower.hpp.
#include "lower.hpp"

namespace wfsa {

wfsa_fsa_desc LoweredFsa::desc() const
{
    wfsa_fsa_desc d{};
    d.n_states = n_states; d.start_state = start; d.end_state = end; d.n_symbols = n_symbols; d.n_raw_params = n_raw;
    d.emis_row = emis_row.data(); d.emis_tok_off = emis_tok_off.data(); d.emis_tok = emis_tok.data();
    d.emis_param = emis_param.data(); d.trans_row = trans_row.data(); d.trans_dst = trans_dst.data();
    d.trans_param = trans_param.data();
    return d;
}

wfsa_corpus_desc LoweredCorpus::desc() const
{
    wfsa_corpus_desc d{};
    d.n_strings = (int64_t)p.size(); d.offsets = offsets.data(); d.tokens = tokens.data(); d.p = p.data();
    return d;
}

void lower_fsa(const Fsa& fsa, LoweredFsa& L)
{
    L = LoweredFsa();
    const auto& S = fsa.States();
    L.n_states = (int)S.size(); L.start = fsa.StartIndex(); L.end = fsa.EndIndex();
    L.n_raw = (int)fsa.GetNumberOfParameters();
    for (int& s : L.sym_of_byte) s = -1;
    for (const auto& st : S)
        for (const auto& e : st.emissions)
            for (unsigned char c : e.str)
                if (L.sym_of_byte[c] < 0) L.sym_of_byte[c] = L.n_symbols++;
    L.emis_row.push_back(0); L.trans_row.push_back(0); L.emis_tok_off.push_back(0);
    for (size_t si = 0; si < S.size(); ++si) {
        const auto& st = S[si];
        for (size_t ei = 0; ei < st.emissions.size(); ++ei) {
            const auto& e = st.emissions[ei];
            for (unsigned char c : e.str) L.emis_tok.push_back(L.sym_of_byte[c]);
            L.emis_tok_off.push_back((int32_t)L.emis_tok.size());
            L.emis_param.push_back(e.index);
            L.emis_edge.emplace_back((int)si, (int)ei);
        }
        L.emis_row.push_back((int32_t)L.emis_param.size());
        for (size_t ti = 0; ti < st.transitions.size(); ++ti) {
            L.trans_dst.push_back(st.transitions[ti].next);
            L.trans_param.push_back(st.transitions[ti].index);
            L.trans_edge.emplace_back((int)si, (int)ti);
        }
        L.trans_row.push_back((int32_t)L.trans_dst.size());
    }
}

void lower_corpus(const Corpus& corpus, const LoweredFsa& fsa, size_t first, size_t count, LoweredCorpus& out)
{
    out = LoweredCorpus();
    out.offsets.push_back(0);
    for (size_t i = first; i < first + count && i < corpus.size(); ++i) {
        for (unsigned char c : corpus[i].first) out.tokens.push_back(fsa.sym_of_byte[c]);
        out.offsets.push_back((int64_t)out.tokens.size());
        out.p.push_back(corpus[i].second);
    }
}

std::vector<size_t> balanced_ranges(const Corpus& corpus, int parts)
{
    std::vector<size_t> cut(parts + 1, corpus.size());
    size_t total = 0;
    for (const auto& w : corpus) total += w.first.size() + 1;
    cut[0] = 0;
    size_t acc = 0; int k = 1;
    for (size_t i = 0; i < corpus.size() && k < parts; ++i) {
        acc += corpus[i].first.size() + 1;
        while (k < parts && acc * parts >= total * (size_t)k) cut[k++] = i + 1;
    }
    return cut;
}

}  // namespace wfsa
