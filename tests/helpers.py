"""Shared helpers of the test-suite: golden cases and index maps keyed by (state, kind, label).

Parameter numbering differs between the reference (std::unordered_map iteration order,
/root/reference/src/Fsa.cpp:207-238) and this build (file order), so every comparison maps
through edge keys, never through indices ("Parameter ordering")."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def fnum(v):
    return float(v) if isinstance(v, str) else v


def load_cases(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)["cases"]


def all_cases():
    out = []
    for name in ("fixtures", "random", "c4shape"):
        out.extend(load_cases(name))
    return out


def good_cases(names=("fixtures", "random", "c4shape")):
    out = []
    for name in names:
        out.extend(c for c in load_cases(name) if "reference_error" not in c and not c.get("degenerate"))
    return out


def key(e):
    return (e["state"], e["kind"], e["label"])


def golden_to_mine(case, my_edges):
    """perm[g] = index in MY trimmed numbering of the golden trimmed parameter g."""
    mine = {key(e): e for e in my_edges}
    n = case["n"]
    perm = np.full(n, -1, dtype=np.int64)
    for e in case["edges"]:
        if e.get("trimmed", -1) >= 0:
            perm[e["trimmed"]] = mine[key(e)]["trimmed"]
    assert (perm >= 0).all()
    return perm


def recognised_words(case):
    return [w for w in case["corpus"] if w["paths"] > 0]


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = np.maximum(np.abs(b), floor)
    scale[scale == 0] = 1.0
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def vec_tol_ok(mine, ref, rtol=1e-9):
    """|mine - ref| <= rtol * max(|ref_i|, 1e-6 * max|ref|) for every component (fixed-point accumulators
    have an absolute quantum of 2^-fx)."""
    mine, ref = np.asarray(mine, dtype=float), np.asarray(ref, dtype=float)
    if ref.size == 0:
        return True, 0.0
    floor = max(1e-6 * float(np.max(np.abs(ref))), 1e-300)
    err = np.abs(mine - ref) / np.maximum(np.abs(ref), floor)
    return bool(np.all(err <= rtol)), float(err.max())
