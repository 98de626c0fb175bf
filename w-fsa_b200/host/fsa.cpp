// w-fsa_b200/host/fsa.cpp -- see fsa.hpp.
#include "fsa.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <unordered_map>
#include <unordered_set>

namespace wfsa {

bool read_file(FILE* f, std::string& out)
{
    if (!f) return false;
    out.clear();
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) out.append(buf, got);
    return !ferror(f);
}

std::pair<std::string, char> get_word(const std::string& text, size_t& pos, const std::string& sep)
{
    const size_t begin = pos, n = text.size();
    for (size_t i = pos; i < n; ++i) {
        const char c = text[i];
        if (c == '\0') { pos = n; return {text.substr(begin, i - begin), '\0'}; }
        if (!sep.empty() && c == sep[0] && text.compare(i, sep.size(), sep) == 0) {
            const size_t after = i + sep.size();
            if (sep == "\n") { pos = after; return {text.substr(begin, i - begin), '\n'}; }
            if (after < n && text[after] == '\n') {       // separator then end of line: newline stays unread
                pos = after;
                return {text.substr(begin, i - begin), '\n'};
            }
            pos = after;
            if (after >= n) return {text.substr(begin, i - begin), '\0'};
            return {text.substr(begin, i - begin), sep.back()};
        }
        if (c == '\n') { pos = i + 1; return {text.substr(begin, i - begin), '\n'}; }
    }
    pos = n;
    return {text.substr(begin), '\0'};
}

static bool has_prefix(const std::string& word, const std::string& prefix)
{
    return word.compare(0, prefix.size(), prefix) == 0;
}

// ---------------------------------------------------------------------------------------------
int Fsa::state_id(const std::string& name)
{
    auto it = name_index.find(name);
    if (it != name_index.end()) return it->second;
    states.emplace_back();
    states.back().name = name;
    name_index.emplace(name, (int)states.size() - 1);
    return (int)states.size() - 1;
}

void Fsa::Read(FILE* input)
{
    std::string content;
    if (!read_file(input, content)) throw FsaError("Unable to read file!");
    ReadText(content);
}

void Fsa::ReadText(const std::string& content)
{
    states.clear(); name_index.clear(); m1 = m2 = n = 0; start_idx = end_idx = -1;
    const size_t lines = (size_t)std::count(content.begin(), content.end(), '\n');
    const size_t expected = (std::max<size_t>(lines, 3) - 3) / 2 + 1;        // src/Fsa.cpp:115-123
    size_t pos = 0;
    separator = get_word(content, pos, "\n").first;
    start_state = get_word(content, pos, "\n").first;
    end_state = get_word(content, pos, "\n").first;
    if (separator.empty()) separator = " ";
    for (const std::string* x : {&start_state, &end_state})
        if (has_prefix(*x, separator))
            throw FsaError("Invalid FSA format! Start or end state contains the separator! \"" + separator + "\" is in \"" + *x + "\"");
    if (start_state == end_state)
        throw FsaError("Invalid FSA format! Start and end states should be different! \"" + start_state + "\"==\"" + end_state + "\"");
    while (pos < content.size() && content[pos] != '\0') read_one_state(content, pos);
    if (states.size() > expected) {
        std::ostringstream o;
        o << "Invalid FSA format! There are more states than rows in the automaton file! " << states.size() << " > " << expected;
        throw FsaError(o.str());
    }
    for (size_t i = 0; i < states.size(); ++i) {
        if (states[i].name == start_state) start_idx = (int)i;
        if (states[i].name == end_state) end_idx = (int)i;
    }
    if (start_idx < 0) throw FsaError("Invalid FSA format! The start state \"" + start_state + "\" is not defined!");
    end_artificial = end_idx < 0;
    if (end_idx < 0) end_idx = state_id(end_state);
    assign_indices();
}

void Fsa::read_one_state(const std::string& text, size_t& pos)
{
    auto result = get_word(text, pos, separator);
    const std::string this_state = result.first;
    if (this_state.empty() || has_prefix(this_state, end_state)) {   // comment / blank line, src/Fsa.cpp:130-134
        get_word(text, pos, "\n");
        return;
    }
    std::vector<Emission> emissions;
    do {
        result = get_word(text, pos, separator);
        const std::string word = result.first;
        if (this_state == start_state && !word.empty())
            throw FsaError("Invalid FSA format! Start state should emit empty string instead of \"" + word + "\"!");
        for (const auto& e : emissions)
            if (e.str == word)
                throw FsaError("Invalid FSA format! Emission \"" + word + "\" of state \"" + this_state + "\" appears more than once!");
        result = get_word(text, pos, separator);
        Emission e; e.str = word; e.logprob = std::atof(result.first.c_str());
        emissions.push_back(e);
    } while (result.second != '\n' && result.second != '\0');

    if (get_word(text, pos, separator).first != this_state)
        throw FsaError("Invalid FSA format! You should enlist transitions of \"" + this_state + "\" after emissions of the same state!");
    std::vector<std::pair<std::string, double>> transitions;
    do {
        result = get_word(text, pos, separator);
        const std::string word = result.first;
        for (const auto& t : transitions)
            if (t.first == word)
                throw FsaError("Invalid FSA format! Transition \"" + this_state + "\" -> \"" + word + "\" appears more than once!");
        if (word == start_state)
            throw FsaError("Invalid FSA format! \"" + this_state + "\" connects to start state \"" + start_state + "\"!");
        result = get_word(text, pos, separator);
        transitions.emplace_back(word, std::atof(result.first.c_str()));
    } while (result.second != '\n' && result.second != '\0');

    const int self = state_id(this_state);
    std::vector<Transition> tr;
    for (const auto& t : transitions) {
        Transition x; x.next = state_id(t.first); x.logprob = t.second;
        tr.push_back(x);
    }
    State& s = states[self];
    s.emissions = emissions;       // a second definition of a state replaces the first (src/Fsa.cpp:204)
    s.transitions = tr;
    s.defined = true;
}

void Fsa::assign_indices()
{
    m1 = m2 = n = 0;
    for (auto& s : states) {
        if (s.emissions.size() == 1) s.emissions[0].index = -1;
        else for (auto& e : s.emissions) e.index = (int)(n++);
        if (s.transitions.size() == 1) s.transitions[0].index = -1;
        else for (auto& t : s.transitions) t.index = (int)(n++);
        m1 += s.transitions.size();
        m2 += s.emissions.size();
    }
}

std::string Fsa::DumpString(bool full) const
{
    std::string out = separator + "\n" + start_state + "\n" + end_state + "\n";
    char buf[64];
    auto num = [&](double v) { snprintf(buf, sizeof(buf), full ? "%.17g" : "%g", v); return std::string(buf); };
    for (const auto& s : states) {
        if (s.name == end_state) continue;
        out += s.name;
        for (const auto& e : s.emissions) out += separator + e.str + separator + num(e.logprob);
        out += "\n" + s.name;
        for (const auto& t : s.transitions) out += separator + states[t.next].name + separator + num(t.logprob);
        out += "\n";
    }
    return out;
}

void Fsa::Dump(FILE* out) const
{
    const std::string s = DumpString(false);
    fwrite(s.data(), 1, s.size(), out);
}

// ---------------------------------------------------------------------------------------------
void Corpus::Read(FILE* input)
{
    std::string content;
    if (!read_file(input, content)) throw CorpusError("Cannot read file!");
    ReadText(content);
}

void Corpus::ReadText(const std::string& content)
{
    clear();
    std::unordered_set<std::string> words;
    size_t pos = 0;
    auto result = get_word(content, pos, "\n");
    separator = result.first;
    if (separator.empty()) separator = " ";
    std::string word;
    while (result.second != '\0') {
        word.clear();
        bool empty = true;
        do {
            result = get_word(content, pos, separator);
            if (result.second == '\n' || result.second == '\0') {
                if (!empty) {
                    if (words.insert(word).second) emplace_back(word, std::atof(result.first.c_str()));
                    else throw CorpusError("\"" + word + "\" is duplicate!");
                }
                break;
            }
            empty = false;
            word += result.first;
        } while (result.second);
    }
    for (const auto& w : *this) {
        if (!std::isnormal(w.second) || w.second < 0) {
            std::ostringstream o;
            o << "\"" << w.first << "\" has probability " << w.second << "!";
            throw CorpusError(o.str());
        }
    }
}

void Corpus::Renormalize()
{
    const double s = Sum();
    for (auto& w : *this) w.second /= s;
}

double Corpus::Sum() const
{
    double s = 0.0;
    for (const auto& w : *this) s += w.second;
    return s;
}

}  // namespace wfsa
