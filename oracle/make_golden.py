#!/usr/bin/env python3
"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.json by running the UNMODIFIED reference (oracle/_ref/ref_probe,
built by `make -C oracle ref` from /root/reference/src against oracle/mkl_shim) on

  * every automaton/corpus pair the reference's CTest list exercises plus the pairs it
    leaves untested, and
  * seeded random automata (multi-character and empty emissions, ambiguous segmentations,
    dead ends, unrecognised strings) and a down-scaled config-4-shaped automaton.

For every case it stores the inputs (file texts), the reference's structural results
(path counts per string, trimmed index of every edge, P/M/C, p), the reference's own
logq / KL / path posteriors at several x, the gradient and H_f evaluated in numpy float64
from the reference's matrices with the reference's formulas
(src/QuasiNewtonLearner.cpp:93-125, src/HessianLearner.cpp:381-547), and full optimisation
trajectories of both optimisers.

Can only run where /root/reference exists (the build container).  The JSON it writes is
committed; tests never need the reference.
    python oracle/make_golden.py
"""
import json
import os
import random
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("WFSA_REFERENCE", "/root/reference")
PROBE = os.path.join(HERE, "_ref", "ref_probe")
OUT = os.path.join(ROOT, "tests", "golden")


def run_probe(fsa_path, corpus_path, optimizer, flags, epochs, eta=1.0, tol=1e-6, xfile=None):
    cmd = [PROBE, fsa_path, corpus_path, optimizer, str(flags), str(epochs), repr(eta), repr(tol)]
    if xfile:
        cmd.append(xfile)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if res.returncode != 0:
        return {"probe_error": res.stderr.strip().splitlines()[-1] if res.stderr.strip() else "exit %d" % res.returncode,
                "exit": res.returncode}
    return json.loads(res.stdout)


def fnum(v):
    if isinstance(v, str):
        return float(v)
    return float(v)


def csr_dense(row, col, data, ncols):
    nrows = len(row) - 1
    A = np.zeros((nrows, ncols))
    for i in range(nrows):
        for j in range(int(row[i]), int(row[i + 1])):
            A[i, int(col[j])] += fnum(data[j]) if data is not None else 1.0
    return A


def grad_and_hf(d, r):
    """grad = P^T (r * (-M^T p)); H_f per src/HessianLearner.cpp:498-547 with the index sets of
    AssembleH (:409-443).  Unique paths: grad = -P^T p, H_f = 0."""
    n = d["n"]
    P = csr_dense(d["Prow"], d["Pcol"], d["Pdata"], n)
    L = P.shape[0]
    M = csr_dense(d["Mrow"], d["Mcol"], None, L)
    p = np.array([fnum(v) for v in d["p"]])
    H = np.zeros((n, n))
    if d["unique_paths"]:
        return (-(P.T @ p)), H
    r = np.array([fnum(v) for v in r])
    grad = P.T @ (r * (-(M.T @ p)))
    Mrow = [int(v) for v in d["Mrow"]]
    for s in range(len(p)):
        lo, hi = Mrow[s], Mrow[s + 1]
        if hi - lo < 2:
            continue
        rows = P[lo:hi]
        union = np.where((rows != 0).any(axis=0))[0]
        # drop (col,count) pairs identical in every path
        idx = [j for j in union if not np.all(rows[:, j] == rows[0, j])]
        g = {j: float((rows[:, j] * r[lo:hi]).sum()) for j in idx}
        for a_, j in enumerate(idx):
            for k in idx[a_:]:
                hjk = g[j] * g[k] - float((rows[:, j] * rows[:, k] * r[lo:hi]).sum())
                H[j, k] += p[s] * hjk
    return grad, H


def edge_key(e):
    return (e["state"], e["kind"], e["label"])


def trimmed_of(sentinel_logprob):
    v = fnum(sentinel_logprob)
    if v == float("-inf"):
        return -2
    if v == 0.0:
        return -1
    return int(round(v)) - 1


def build_case(name, fsa_text, corpus_text, runs, n_random_x=2, seed=0, note="", slim=False):
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "a.wfsa")
        fc = os.path.join(tmp, "c.corpus")
        with open(fa, "w", newline="") as f:
            f.write(fsa_text)
        with open(fc, "w", newline="") as f:
            f.write(corpus_text)
        base = run_probe(fa, fc, "QuasiNewton", 0, 0)
        case = {"name": name, "note": note, "fsa_text": fsa_text, "corpus_text": corpus_text}
        if "probe_error" in base:
            case["reference_error"] = base["probe_error"]
            return case
        for k in ("corpus_size", "corpus_sum", "states", "transitions", "emissions", "raw_parameters",
                  "raw_constraints", "strings", "paths", "common_support", "unique_paths", "n", "k", "corpus"):
            case[k] = base[k]
        case["degenerate"] = bool(base.get("degenerate", False))
        raw = {edge_key(e): e for e in base["raw_edges"]}
        if case["degenerate"]:
            case["edges"] = [{"state": e["state"], "kind": e["kind"], "label": e["label"], "raw": e["raw"],
                              "file_logprob": e["logprob"]} for e in base["raw_edges"]]
            return case
        sent = {edge_key(e): e for e in base["sentinel_edges"]}
        edges = []
        for key, e in raw.items():
            t = trimmed_of(sent[key]["logprob"]) if e["raw"] >= 0 else -1
            edges.append({"state": e["state"], "kind": e["kind"], "label": e["label"], "raw": e["raw"],
                          "trimmed": t, "file_logprob": e["logprob"]})
        case["edges"] = edges
        # slim cases (large n) keep only what the parity tests compare: no P/M, no dense H_f
        for k in (("Ccol", "p", "x_file") if slim else
                  ("Ccol", "Prow", "Pcol", "Pdata", "Mrow", "Mcol", "p", "x_file")):
            case[k] = base[k]
        n = base["n"]
        rng = np.random.RandomState(seed)
        xs = [np.zeros(n), np.array([fnum(v) for v in base["x_file"]])]
        for _ in range(n_random_x):
            xs.append(rng.normal(-1.0, 0.7, size=n))
        xfile = os.path.join(tmp, "x.txt")
        with open(xfile, "w") as f:
            for x in xs:
                f.write(" ".join(repr(float(v)) for v in x) + "\n")
        ev = run_probe(fa, fc, "QuasiNewton", 0, 0, xfile=xfile)
        case["evals"] = []
        for e in ev["evals"]:
            grad, H = grad_and_hf(base, e["r"])
            rec = {"x": e["x"], "logq": e["logq"], "kl": e["kl"], "grad": [float(v) for v in grad]}
            if not slim:
                rec["r"] = e["r"]
                rec["Hf"] = [[float(v) for v in row] for row in H]
            case["evals"].append(rec)
        case["runs"] = []
        for (opt, flags, epochs, eta, tol) in runs:
            t = run_probe(fa, fc, opt, flags, epochs, eta, tol)
            if "probe_error" in t:
                case["runs"].append({"optimizer": opt, "flags": flags, "epochs": epochs, "eta": eta, "tol": tol,
                                     "reference_error": t["probe_error"]})
                continue
            case["runs"].append({
                "optimizer": opt, "flags": flags, "epochs": epochs, "eta": eta, "tol": tol,
                "x_init": t["x_init"],
                "trajectory": ([{"epoch": r_["epoch"], "info": r_["info"]} for r_ in t["trajectory"][:-1]]
                               + t["trajectory"][-1:]) if slim else t["trajectory"],
                "halted": t["halted"],
                "last_epoch": t["last_epoch"], "error": t["error"],
                "final_edges": [{"state": e["state"], "kind": e["kind"], "label": e["label"], "logprob": e["logprob"]}
                                for e in t["final_edges"]]})
        return case


# ----------------------------------------------------------------------------------------------
# random automata in the reference's text format (src/Fsa.cpp:73-205, src/Corpus.cpp:9-61)
# ----------------------------------------------------------------------------------------------
def random_automaton(rng, n_states, alphabet, max_emis_len, allow_eps, sep=" ", max_emis=3, max_trans=3):
    names = ["q%d" % i for i in range(n_states)]
    emis = {}
    for i, s in enumerate(names):
        k = rng.randint(1, max_emis)
        choices = set()
        for _ in range(50):         # bounded: fewer than k distinct strings may exist
            if len(choices) >= k:
                break
            ln = rng.randint(0 if (allow_eps and i > 0) else 1, max_emis_len)
            choices.add("".join(rng.choice(alphabet) for _ in range(ln)))
        emis[s] = sorted(choices)
    trans = {}
    # the start state may go anywhere; an epsilon-emitting state may only be entered from a
    # lower-numbered state, so the epsilon sub-graph is acyclic (the reference itself does not
    # terminate on epsilon cycles, inc/Recognize.h:35-60)
    start_targets = rng.sample(names, min(len(names), rng.randint(1, max_trans)))
    trans["^"] = start_targets
    for i, s in enumerate(names):
        k = rng.randint(1, max_trans)
        cands = []
        for j, t in enumerate(names):
            if "" in emis[t] and j <= i:
                continue
            cands.append(t)
        cands.append("$")
        trans[s] = rng.sample(cands, min(len(cands), k))
        if "$" not in trans[s] and rng.random() < 0.4:
            trans[s].append("$")
    lines = [sep if sep != " " else "", "^", "$"]
    def w():
        return "%.4f" % rng.uniform(-2.0, 0.5)
    lines.append(sep.join(["^", "", "0"]))
    lines.append(sep.join(["^"] + [x for t in trans["^"] for x in (t, w())]))
    for s in names:
        lines.append(sep.join([s] + [x for e in emis[s] for x in (e, w())]))
        lines.append(sep.join([s] + [x for t in trans[s] for x in (t, w())]))
    return "\n".join(lines) + "\n", names, emis, trans


def random_walk_string(rng, emis, trans, max_steps):
    s = "^"
    out = ""
    for _ in range(max_steps):
        t = rng.choice(trans[s])
        if t == "$":
            return out
        out += rng.choice(emis[t])
        s = t
    return None


def random_corpus(rng, emis, trans, alphabet, n_strings, max_steps, max_len, sep=" "):
    words = {}
    tries = 0
    while len(words) < n_strings and tries < 2000:
        tries += 1
        if rng.random() < 0.8:
            wv = random_walk_string(rng, emis, trans, max_steps)
        else:
            wv = "".join(rng.choice(alphabet) for _ in range(rng.randint(1, 4)))
        if wv is None or wv == "" or len(wv) > max_len or wv in words:
            continue
        words[wv] = rng.choice([1, 1, 2, 3, 0.5, 2.5])
    lines = [sep if sep != " " else ""]
    for wv, c in words.items():
        lines.append(sep.join([wv, repr(c) if isinstance(c, float) else str(c)]))
    return "\n".join(lines) + "\n"


def c4_shaped(rng, n_states, n_sym, n_succ, n_emis, n_strings, lmin, lmax):
    """Down-scaled BASELINE config 4 generator: every state emits n_emis
    distinct one-character symbols, has n_succ successors plus a transition to the end state."""
    alphabet = [chr(33 + i) for i in range(n_sym)]
    names = ["s%d" % i for i in range(n_states)]
    emis = {s: rng.sample(alphabet, n_emis) for s in names}
    trans = {s: rng.sample(names, n_succ) + ["$"] for s in names}
    trans["^"] = rng.sample(names, n_succ)
    sep = "\t"
    lines = [sep, "^", "$", sep.join(["^", "", "0"]), sep.join(["^"] + [x for t in trans["^"] for x in (t, "0")])]
    for s in names:
        lines.append(sep.join([s] + [x for e in emis[s] for x in (e, "0")]))
        lines.append(sep.join([s] + [x for t in trans[s] for x in (t, "0")]))
    fsa_text = "\n".join(lines) + "\n"
    words = {}
    while len(words) < n_strings:
        ln = rng.randint(lmin, lmax)
        s = "^"
        out = ""
        for _ in range(ln):
            s = rng.choice([t for t in trans[s] if t != "$"])
            out += rng.choice(emis[s])
        if out not in words:
            words[out] = rng.randint(1, 5)
    ctext = "\n".join([sep] + [sep.join([wv, str(c)]) for wv, c in words.items()]) + "\n"
    return fsa_text, ctext


def main():
    if not os.path.exists(PROBE):
        sys.exit("build the reference first: make -C oracle ref")
    os.makedirs(OUT, exist_ok=True)
    data = os.path.join(REF, "data")

    def rd(fn):
        with open(os.path.join(data, fn), newline="") as f:
            return f.read()

    std_runs = [("QuasiNewton", 7, 30, 1.0, 1e-6), ("Hessian", 31, 20, 1.0, 1e-6),
                ("Hessian", 7, 20, 1.0, 1e-6), ("QuasiNewton", 6, 10, 0.5, 0.0),
                ("QuasiNewton", 39, 10, 1.0, 0.0), ("Hessian", 15, 6, 0.7, 0.0)]
    pairs = [("test.wfsa", "test.corpus"), ("test.list.wfsa", "test.corpus"), ("test2.wfsa", "test.corpus"),
             ("test3.wfsa", "test.corpus"), ("test4.wfsa", "test.corpus"), ("test.loop.wfsa", "test.corpus"),
             ("talk.wfsa", "talk.corpus"), ("talk.wfsa", "test.corpus"), ("test5.wfsa", "test5.corpus"),
             ("test5_2.wfsa", "test5.corpus"), ("test.wfsa.win", "test.corpus")]
    fixtures = []
    for a, c in pairs:
        name = "%s+%s" % (a, c)
        print("fixture", name, file=sys.stderr)
        fixtures.append(build_case(name, rd(a), rd(c), std_runs, seed=len(fixtures),
                                   note="reference data fixture (inputs embedded verbatim)"))
    with open(os.path.join(OUT, "fixtures.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "cases": fixtures}, f, separators=(",", ":"))

    rnd = []
    rng = random.Random(20261018)
    specs = []
    for i in range(36):
        specs.append(dict(n_states=rng.randint(2, 6), alphabet="ab" if i % 3 else "abc",
                          max_emis_len=rng.choice([1, 2, 3]), allow_eps=(i % 2 == 0),
                          sep=[" ", "\t", "::", ";"][i % 4]))
    for i, sp in enumerate(specs):
        fsa_text, names, emis, trans = random_automaton(rng, sp["n_states"], sp["alphabet"], sp["max_emis_len"],
                                                        sp["allow_eps"], sp["sep"])
        ctext = random_corpus(rng, emis, trans, sp["alphabet"], rng.randint(3, 12), 7, 9, sp["sep"])
        print("random", i, file=sys.stderr)
        case = build_case("random%02d" % i, fsa_text, ctext,
                          [("QuasiNewton", 7, 8, 1.0, 0.0), ("Hessian", 15, 5, 1.0, 0.0)], seed=100 + i,
                          note="seeded random automaton: %r" % sp)
        rnd.append(case)
    with open(os.path.join(OUT, "random.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "cases": rnd}, f, separators=(",", ":"))

    c4 = []
    for i, (S, A, D, E, N, lo, hi) in enumerate([(24, 8, 3, 2, 60, 4, 14), (64, 16, 4, 3, 100, 6, 18),
                                                  (256, 64, 8, 4, 120, 8, 24)]):
        rng2 = random.Random(4000 + i)
        fsa_text, ctext = c4_shaped(rng2, S, A, D, E, N, lo, hi)
        print("c4-shaped", i, file=sys.stderr)
        c4.append(build_case("c4shape_%dx%d" % (S, A), fsa_text, ctext,
                             [("QuasiNewton", 7, 5, 1.0, 0.0)], n_random_x=1, seed=200 + i, slim=True,
                             note="down-scaled config-4 generator S=%d syms=%d D=%d E=%d" % (S, A, D, E)))
    with open(os.path.join(OUT, "c4shape.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "cases": c4}, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
