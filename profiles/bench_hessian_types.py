#!/usr/bin/env python3
"""profiles/bench_hessian_types.py -- wfsa_dev_hessian on a config-4 corpus, blocks derived from the compiled region types
(no path enumeration per string).  usage: python profiles/bench_hessian_types.py [n_strings] [reps]
Prints one JSON line: blocks (= region types with more than one path), paths, sum of D*D over the blocks (the cells the
kernel scatters), ms per call."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "w-fsa_b200", "python"))
import wfsa_b200 as W  # noqa: E402
from wfsa_b200 import synth  # noqa: E402

n_strings = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model = synth.make_model(256, 64, 8, 4, seed=1234)
low = model.lowered()
offs, toks, w = model.corpus(n_strings, 32, 128, seed=1235)
low.set_tokens(offs, toks, w / w.sum())
dev = W.Device(low)
rec, pc, used = dev.structure()
trimmed = np.where(used > 0, 0, -2).astype(np.int32)
n = 0
for i in range(len(trimmed)):
    if trimmed[i] == 0:
        trimmed[i] = n
        n += 1
dev.set_param_map(trimmed, n, rec)
x = np.random.RandomState(0).normal(-1.0, 0.3, size=n)
H = np.zeros((n, n))
rmin = C.c_double()
t0 = time.perf_counter()
dev._ck(dev.L.wfsa_dev_hessian(dev.h, W._p(x, W.F64P), W._p(H, W.F64P), C.byref(rmin)))      # builds the blocks on first use
first_s = time.perf_counter() - t0
t0 = time.perf_counter()
for _ in range(reps):
    dev._ck(dev.L.wfsa_dev_hessian(dev.h, W._p(x, W.F64P), W._p(H, W.F64P), C.byref(rmin)))
ms = (time.perf_counter() - t0) / reps * 1e3
t0 = time.perf_counter()
for _ in range(reps):
    dev._ck(dev.L.wfsa_dev_hessian(dev.h, W._p(x, W.F64P), None, C.byref(rmin)))
ms_rmin = (time.perf_counter() - t0) / reps * 1e3
S = W.segmented_compile(low, trimmed)
po, co = S["hb_path_off"], S["hb_col_off"]
D = np.diff(co)
L = np.diff(po)
print(json.dumps({"strings": n_strings, "paths_of_all_strings": float(pc.sum()), "n": n, "blocks": int(len(D)), "block_paths": int(L.sum()),
                  "sum_DD": int((D.astype(np.int64) ** 2).sum()), "mean_D": float(D.mean()), "max_D": int(D.max()), "max_paths": int(L.max()),
                  "first_call_s": first_s, "ms_per_call": ms, "ms_rmin_only": ms_rmin, "rmin": rmin.value,
                  "symmetric": bool(np.abs(H - H.T).max() <= 1e-15), "trace": float(np.trace(H))}))
dev.close()
