"""Synthetic automata and corpora of the BASELINE.json shapes.

HMM-shaped like the reference's parameterisation: every state emits `n_emis` distinct
one-token symbols and has `n_succ` distinct successors plus a transition to the end state;
the start state has `n_succ` successors.  Strings are random walks from the start state
(so every string is recognised), length uniform in [lmin, lmax], integer weights in [1, 5].
Config 4: make_model(256, 64, 8, 4); config 5: make_model(4096, 256, 64, 16).
"""
import numpy as np

from . import Lowered


class Model:
    def __init__(self, n_states, n_sym, n_succ, n_emis, seed):
        rng = np.random.RandomState(seed)
        self.S, self.A, self.D, self.E, self.seed = n_states, n_sym, n_succ, n_emis, seed
        self.emis = np.stack([rng.choice(n_sym, n_emis, replace=False) for _ in range(n_states)]).astype(np.int32)
        self.succ = np.stack([rng.choice(n_states, n_succ, replace=False) for _ in range(n_states)]).astype(np.int32)
        self.start_succ = rng.choice(n_states, n_succ, replace=False).astype(np.int32)

    # ---- descriptors without going through text (symbols are abstract ids) ----
    def lowered(self):
        S, E, D = self.S, self.E, self.D
        start, end = S, S + 1                      # states 0..S-1 emit; S = start; S+1 = end
        low = Lowered.__new__(Lowered)
        low.desc = None
        low.n_states, low.start, low.end, low.n_symbols = S + 2, start, end, self.A
        emis_row = np.concatenate([np.arange(S + 1) * E, [S * E + 1, S * E + 1]]).astype(np.int32)
        # state S (start) has one empty emission (never consumed), the end state none
        emis_tok = self.emis.reshape(-1).astype(np.int32)
        tok_off = np.concatenate([np.arange(S * E + 1), [S * E]]).astype(np.int32)
        raw = 0
        eparam = np.empty(S * E + 1, dtype=np.int32)
        tparam = np.empty(S * (D + 1) + D, dtype=np.int32)
        tdst = np.empty(S * (D + 1) + D, dtype=np.int32)
        trow = np.empty(S + 3, dtype=np.int32)
        for s in range(S):
            if E > 1:
                eparam[s * E:(s + 1) * E] = np.arange(raw, raw + E); raw += E
            else:
                eparam[s * E] = -1
            b = s * (D + 1)
            trow[s] = b
            tdst[b:b + D] = self.succ[s]; tdst[b + D] = end
            tparam[b:b + D + 1] = np.arange(raw, raw + D + 1); raw += D + 1
        eparam[S * E] = -1
        b = S * (D + 1)
        trow[S] = b
        tdst[b:b + D] = self.start_succ
        if D > 1:
            tparam[b:b + D] = np.arange(raw, raw + D); raw += D
        else:
            tparam[b] = -1
        trow[S + 1] = b + D
        trow[S + 2] = b + D
        low.n_raw = raw
        low.emis_row, low.emis_tok_off, low.emis_tok, low.emis_param = emis_row, tok_off, emis_tok, eparam
        low.trans_row, low.trans_dst, low.trans_param = trow, tdst, tparam
        low.n_trans, low.n_emis = len(tdst), len(eparam)
        low.sym, low.edges, low.trans_edges, low.emis_edges, low.words = None, None, None, None, None
        return low

    # ---- the reference's text format (one-character symbols: needs n_sym <= 90) ----
    def text(self, sep="\t"):
        assert self.A <= 90, "one-byte symbols only"
        ch = lambda c: chr(33 + int(c))
        lines = [sep, "^", "$", sep.join(["^", "", "0"]), sep.join(["^"] + [x for t in self.start_succ for x in ("s%d" % t, "0")])]
        for s in range(self.S):
            lines.append(sep.join(["s%d" % s] + [x for c in self.emis[s] for x in (ch(c), "0")]))
            lines.append(sep.join(["s%d" % s] + [x for t in self.succ[s] for x in ("s%d" % t, "0")] + ["$", "0"]))
        return "\n".join(lines) + "\n"

    def corpus(self, n_strings, lmin, lmax, seed):
        """Random walks; returns (offsets int64, tokens int32, weights float64)."""
        rng = np.random.RandomState(seed)
        lens = rng.randint(lmin, lmax + 1, size=n_strings)
        offsets = np.zeros(n_strings + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        tokens = np.empty(int(offsets[-1]), dtype=np.int32)
        state = self.start_succ[rng.randint(0, self.D, size=n_strings)]
        alive = np.arange(n_strings)
        for t in range(int(lens.max())):
            alive = alive[lens[alive] > t]
            if t > 0:
                state[alive] = self.succ[state[alive], rng.randint(0, self.D, size=len(alive))]
            tokens[offsets[alive] + t] = self.emis[state[alive], rng.randint(0, self.E, size=len(alive))]
        weights = rng.randint(1, 6, size=n_strings).astype(np.float64)
        return offsets, tokens, weights

    def corpus_text(self, offsets, tokens, weights, sep="\t"):
        out = [sep]
        seen = set()
        for i in range(len(weights)):
            w = "".join(chr(33 + int(c)) for c in tokens[offsets[i]:offsets[i + 1]])
            if w in seen:
                continue                     # the format forbids duplicates (src/Corpus.cpp:40-45)
            seen.add(w)
            out.append(sep.join([w, str(int(weights[i]))]))
        return "\n".join(out) + "\n"


def make_model(n_states=256, n_sym=64, n_succ=8, n_emis=4, seed=1234):
    return Model(n_states, n_sym, n_succ, n_emis, seed)


def balanced_ranges(offsets, parts):
    """Cut a corpus into `parts` contiguous ranges of ~equal token count (+1 per string)."""
    n = len(offsets) - 1
    cost = (offsets[1:] - offsets[:-1]) + 1
    cum = np.concatenate([[0], np.cumsum(cost)])
    total = cum[-1]
    cuts = [0]
    for k in range(1, parts):
        cuts.append(int(np.searchsorted(cum, total * k / parts, side="left")))
    cuts.append(n)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts
