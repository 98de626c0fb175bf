// w-fsa_b200/host/fsa.hpp -- the Fsa / Corpus file formats of w-fsa, parsed on the host.
//
// Same formats, same error conditions and the same parameter-numbering rule as the reference
// (/root/reference/src/Fsa.cpp:73-238, src/Corpus.cpp:9-80, src/Utils.cpp:20-80); the
// in-memory model is index based (vectors in file order) instead of a hash map of C strings,
// because its only consumer is the lowering to the device layout (lower.hpp).
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace wfsa {

struct MyError : public std::runtime_error { using std::runtime_error::runtime_error; };
struct FsaError : public MyError { using MyError::MyError; };
struct CorpusError : public MyError { using MyError::MyError; };
struct LearnerError : public MyError { using MyError::MyError; };

// Splits `text` at `pos` into the next word; mirrors GetWord's contract (src/Utils.cpp:20-80):
// a word ends at the separator (multi-character allowed), at '\n' or at the end of the text;
// the returned terminator is the last separator character, '\n' or '\0'.  A separator that is
// directly followed by '\n' ends the word with terminator '\n' but leaves the newline unread.
std::pair<std::string, char> get_word(const std::string& text, size_t& pos, const std::string& sep);

class Fsa {
public:
    struct Emission { std::string str; double logprob = 0.0; int index = -1; };
    struct Transition { int next = -1; double logprob = 0.0; int index = -1; };
    struct State { std::string name; std::vector<Emission> emissions; std::vector<Transition> transitions; bool defined = false; };

    void Read(FILE* input);                       // src/Fsa.cpp:73-113
    void ReadText(const std::string& content);
    void Dump(FILE* out) const;                   // src/Fsa.cpp:47-71 (%g)
    std::string DumpString(bool full_precision = false) const;

    // the reference counts the end state only if some transition names it
    size_t GetNumberOfStates() const { return states.size() - (end_artificial ? 1 : 0); }
    size_t GetNumberOfTransitions() const { return m1; }
    size_t GetNumberOfEmissions() const { return m2; }
    size_t GetNumberOfParameters() const { return n; }
    size_t GetNumberOfFreeParameters() const { return m1 + m2 - 2 * (GetNumberOfStates() - 1); }   // src/Fsa.cpp:255-258
    size_t GetNumberOfConstraints() const { return n - GetNumberOfFreeParameters(); }
    const std::string& GetStartState() const { return start_state; }
    const std::string& GetEndState() const { return end_state; }
    int StartIndex() const { return start_idx; }
    int EndIndex() const { return end_idx; }
    const std::vector<State>& States() const { return states; }
    std::vector<State>& States() { return states; }
    const std::string& Separator() const { return separator; }

private:
    int state_id(const std::string& name);        // creates an (undefined) state on first sight
    void read_one_state(const std::string& text, size_t& pos);
    void assign_indices();                        // src/Fsa.cpp:207-238
    std::vector<State> states;
    std::unordered_map<std::string, int> name_index;
    std::string separator, start_state, end_state;
    int start_idx = -1, end_idx = -1;
    bool end_artificial = false;
    size_t m1 = 0, m2 = 0, n = 0;
};

class Corpus : public std::vector<std::pair<std::string, double>> {
public:
    void Read(FILE* input);                       // src/Corpus.cpp:9-61
    void ReadText(const std::string& content);
    void Renormalize();                           // src/Corpus.cpp:67-72
    double Sum() const;                           // src/Corpus.cpp:74-80
private:
    std::string separator;
};

bool read_file(FILE* f, std::string& out);

}  // namespace wfsa
