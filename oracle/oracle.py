"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_build/libwfsa_oracle.so.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libwfsa_oracle.so")
I32P, I64P, F64P = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.oracle_enum_eval.argtypes = [C.c_void_p, C.c_void_p, F64P, F64P, F64P, F64P, F64P, C.c_int64]
        L.oracle_enum_hessian.argtypes = [C.c_void_p, C.c_void_p, F64P, F64P, I32P, C.c_int32, F64P, C.c_int64]
        L.oracle_dp_eval.argtypes = [C.c_void_p, C.c_void_p, F64P, F64P, F64P, F64P, F64P, C.c_int]
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def max_threads():
    return lib().oracle_max_threads()


def enum_eval(low, ltw, lew, first=0, count=None, max_paths=2000000):
    """(path_count, logq, edge_exp) by path enumeration -- the reference's algorithm."""
    fd, cd = low.fsa_desc(), low.corpus_desc(first, count)
    n = int(cd.n_strings)
    pc, lq = np.zeros(max(n, 1)), np.zeros(max(n, 1))
    ee = np.zeros(max(low.n_trans + low.n_emis, 1))
    ltw = np.ascontiguousarray(ltw, dtype=np.float64); lew = np.ascontiguousarray(lew, dtype=np.float64)
    rc = lib().oracle_enum_eval(C.byref(fd), C.byref(cd), _p(ltw, F64P), _p(lew, F64P), _p(pc, F64P), _p(lq, F64P), _p(ee, F64P), max_paths)
    if rc != 0:
        raise RuntimeError("oracle_enum_eval: too many paths")
    return pc[:n], lq[:n], ee[:low.n_trans + low.n_emis]


def enum_hessian(low, ltw, lew, edge_param, n, first=0, count=None, max_paths=2000000):
    fd, cd = low.fsa_desc(), low.corpus_desc(first, count)
    H = np.zeros((max(n, 1), max(n, 1)))
    ep = np.ascontiguousarray(edge_param, dtype=np.int32)
    ltw = np.ascontiguousarray(ltw, dtype=np.float64); lew = np.ascontiguousarray(lew, dtype=np.float64)
    rc = lib().oracle_enum_hessian(C.byref(fd), C.byref(cd), _p(ltw, F64P), _p(lew, F64P), _p(ep, I32P), n, _p(H, F64P), max_paths)
    if rc != 0:
        raise RuntimeError("oracle_enum_hessian: too many paths")
    return H[:n, :n]


def dp_eval(low, ltw, lew, first=0, count=None, nthreads=0, want_grad=True, want_counts=False):
    """(path_count|None, logq, edge_exp|None) by CPU forward-backward."""
    fd, cd = low.fsa_desc(), low.corpus_desc(first, count)
    n = int(cd.n_strings)
    lq = np.zeros(max(n, 1))
    pc = np.zeros(max(n, 1)) if want_counts else None
    ee = np.zeros(max(low.n_trans + low.n_emis, 1)) if want_grad else None
    ltw = np.ascontiguousarray(ltw, dtype=np.float64); lew = np.ascontiguousarray(lew, dtype=np.float64)
    lib().oracle_dp_eval(C.byref(fd), C.byref(cd), _p(ltw, F64P), _p(lew, F64P), _p(pc, F64P), _p(lq, F64P), _p(ee, F64P), nthreads)
    return (pc[:n] if want_counts else None), lq[:n], (ee[:low.n_trans + low.n_emis] if want_grad else None)


def trim(low, edge_used):
    """Learner::Trim (src/Learner.cpp:350-425) restated on raw parameter ids: returns
    (trimmed[n_raw], n, Ccol[n]).  Constraint groups are runs of consecutive raw ids that belong to
    one state's emissions or one state's transitions."""
    params = np.concatenate([low.trans_param, low.emis_param])
    used = np.zeros(low.n_raw, dtype=bool)
    for e, r in enumerate(params):
        if r >= 0 and edge_used[e]:
            used[r] = True
    # constraint id of every raw parameter: group = (state, kind)
    group = np.full(low.n_raw, -1, dtype=np.int64)
    for s in range(low.n_states):
        for kind, row, par in ((0, low.emis_row, low.emis_param), (1, low.trans_row, low.trans_param)):
            for e in range(row[s], row[s + 1]):
                if par[e] >= 0:
                    group[par[e]] = s * 2 + kind
    trimmed = np.where(used, 0, -2).astype(np.int32)
    for gid in np.unique(group):
        members = np.where(group == gid)[0]
        alive = [m for m in members if trimmed[m] == 0]
        if len(alive) == 1:
            trimmed[alive[0]] = -1
    n, Ccol, last, k = 0, [], None, -1
    for i in range(low.n_raw):
        if trimmed[i] == 0:
            trimmed[i] = n
            n += 1
            if group[i] != last:
                k += 1
                last = group[i]
            Ccol.append(k)
    return trimmed, n, np.array(Ccol, dtype=np.int32)
