// w-fsa_b200/host/lower.hpp -- "model compiler": Fsa + Corpus -> the index-based descriptors
// of include/wfsa_dev.h.  Characters become dense symbol ids (the alphabet is the set of bytes
// that occur in emission strings); a corpus byte outside the alphabet becomes token -1, which
// no emission matches, so the string is unrecognised exactly as in the reference's prefix
// matching (/root/reference/inc/Recognize.h:52,87).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/wfsa_dev.h"
#include "fsa.hpp"

namespace wfsa {

struct LoweredFsa {
    std::vector<int32_t> emis_row, emis_tok_off, emis_tok, emis_param, trans_row, trans_dst, trans_param;
    int n_states = 0, start = 0, end = 0, n_symbols = 0, n_raw = 0;
    int sym_of_byte[256];
    // edge id -> (state, position inside the state's list) for reporting
    std::vector<std::pair<int, int>> emis_edge, trans_edge;
    wfsa_fsa_desc desc() const;
};

struct LoweredCorpus {
    std::vector<int64_t> offsets;
    std::vector<int32_t> tokens;
    std::vector<double> p;
    wfsa_corpus_desc desc() const;
};

void lower_fsa(const Fsa& fsa, LoweredFsa& out);
// strings [first, first+count) of the corpus
void lower_corpus(const Corpus& corpus, const LoweredFsa& fsa, size_t first, size_t count, LoweredCorpus& out);
// length-balanced contiguous ranges: cut the corpus into `parts` ranges of ~equal token count
std::vector<size_t> balanced_ranges(const Corpus& corpus, int parts);

}  // namespace wfsa
