"""Pins the CPU oracle (oracle/wfsa_oracle.c) against golden vectors produced by the unmodified
reference (tests/golden/*.json, generator oracle/make_golden.py): path counts, used-parameter
flags / Trim, log q, KL, gradient and H_f, for both the enumeration restatement and the CPU
forward-backward.  Only after this does the GPU suite trust the oracle."""
import numpy as np
import pytest

import wfsa_b200 as W
from helpers import fnum, good_cases, key, vec_tol_ok
from oracle import oracle as O


def setup_case(case):
    d = W.parse(case["fsa_text"], case["corpus_text"])
    low = W.Lowered(d)
    zeros_t, zeros_e = np.zeros(low.n_trans), np.zeros(low.n_emis)
    pc, _, ee = O.enum_eval(low, zeros_t, zeros_e)
    trimmed, n, Ccol = O.trim(low, ee > 0)
    return d, low, pc, trimmed, n, Ccol


def perm_golden_to_mine(case, low, trimmed):
    params = np.concatenate([low.trans_param, low.emis_param])
    mine = {key(e): (trimmed[r] if r >= 0 else -1) for e, r in zip(low.edges, params)}
    perm = np.full(case["n"], -1, dtype=np.int64)
    for e in case["edges"]:
        if e["trimmed"] >= 0:
            perm[e["trimmed"]] = mine[key(e)]
    return perm, mine


@pytest.mark.parametrize("case", good_cases(), ids=lambda c: c["name"])
def test_oracle_matches_reference(case):
    d, low, pc, trimmed, n, Ccol = setup_case(case)
    # structure: path counts, Trim
    assert [int(c) for c in pc] == [w["paths"] for w in case["corpus"]]
    assert n == case["n"] and (int(Ccol.max()) + 1 if n else 0) == case["k"]
    perm, mine = perm_golden_to_mine(case, low, trimmed)
    assert (perm >= 0).all() and len(set(perm.tolist())) == n
    for e in case["edges"]:
        assert (mine[key(e)] == -2) == (e["trimmed"] == -2) and (mine[key(e)] == -1) == (e["trimmed"] == -1)
    # same partition into constraints
    gC = np.array(case["Ccol"], dtype=int)
    for a in range(n):
        for b in range(a + 1, n):
            assert (gC[a] == gC[b]) == (Ccol[perm[a]] == Ccol[perm[b]])
    rec = pc > 0
    p = low.p
    plogp = float(np.sum(p[rec] * np.log(p[rec])))
    params = np.concatenate([low.trans_param, low.emis_param])
    edge_tp = np.array([trimmed[r] if r >= 0 else -1 for r in params], dtype=np.int32)
    for ev in case["evals"]:
        xg = np.array([fnum(v) for v in ev["x"]])
        x = np.zeros(n)
        x[perm] = xg
        ltw, lew = low.edge_logweights(x, trimmed)
        for name, fn in (("enum", O.enum_eval), ("dp", lambda *a: O.dp_eval(*a, nthreads=2))):
            _, lq, ee = fn(low, ltw, lew)
            ref_lq = np.array([fnum(v) for v in ev["logq"]])
            assert np.allclose(lq[rec], ref_lq, rtol=1e-12, atol=1e-12), name
            kl = plogp - float(np.sum(p[rec] * lq[rec]))
            assert abs(kl - fnum(ev["kl"])) <= 1e-12 * max(1.0, abs(fnum(ev["kl"]))), name
            grad = np.zeros(n)
            for e, t in enumerate(edge_tp):
                if t >= 0:
                    grad[t] = -ee[e]
            ok, err = vec_tol_ok(grad[perm], [fnum(v) for v in ev["grad"]], 1e-11)
            assert ok, (name, err)
        if "Hf" in ev:
            H = O.enum_hessian(low, ltw, lew, edge_tp, n)
            Hg = np.array(ev["Hf"], dtype=float)
            Hg = np.triu(Hg) + np.triu(Hg, 1).T               # golden stores the upper triangle (j <= k)
            assert np.allclose(H[np.ix_(perm, perm)], Hg, rtol=1e-10, atol=1e-13)


def test_dp_oracle_threads_agree():
    case = [c for c in good_cases(("c4shape",)) if c["name"] == "c4shape_64x16"][0]
    d, low, pc, trimmed, n, Ccol = setup_case(case)
    rng = np.random.RandomState(3)
    x = rng.normal(-1, 0.5, size=n)
    ltw, lew = low.edge_logweights(x, trimmed)
    _, lq1, ee1 = O.dp_eval(low, ltw, lew, nthreads=1)
    _, lq4, ee4 = O.dp_eval(low, ltw, lew, nthreads=4)
    assert np.array_equal(lq1, lq4) and np.allclose(ee1, ee4, rtol=1e-13, atol=1e-16)
