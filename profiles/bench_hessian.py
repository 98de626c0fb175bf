#!/usr/bin/env python3
"""profiles/bench_hessian.py -- times wfsa_dev_hessian (k5_hessian: H_f = sum_s p_s (g g^T - P^T diag(r) P) with the
contraction on the FP64 tensor cores) on synthetic path blocks, for the ncu DMMA-pipe evidence.

    python profiles/bench_hessian.py [--blocks 200000] [--paths 32] [--cols 24] [--reps 5]

Prints one JSON line: blocks/s, contraction FLOP/s (2*L*D*D per block, the symmetric half is what the kernel issues),
and checks the result of a sample of blocks against numpy."""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=200000)
    ap.add_argument("--paths", type=int, default=32)
    ap.add_argument("--cols", type=int, default=24)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    model = synth.make_model(64, 16, 4, 3, seed=5)
    low = model.lowered()
    offs, toks, w = model.corpus(200, 8, 30, seed=6)
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
    rng = np.random.RandomState(1)
    nb, L, D = a.blocks, a.paths, min(a.cols, n)
    cols = np.stack([rng.choice(n, D, replace=False) for _ in range(min(nb, 4096))]).astype(np.int32)
    cols = np.ascontiguousarray(np.tile(cols, ((nb + len(cols) - 1) // len(cols), 1))[:nb])
    counts = rng.randint(0, 3, size=(nb, L, D)).astype(np.float64)
    ps = rng.uniform(0.5, 1.0, size=nb) / nb
    po = np.arange(nb + 1, dtype=np.int64) * L
    co = np.arange(nb + 1, dtype=np.int64) * D
    vo = np.arange(nb + 1, dtype=np.int64) * L * D
    pb = W.PathBlocks(nb, W._p(po, W.I64P), W._p(co, W.I64P), W._p(cols.reshape(-1), W.I32P), W._p(vo, W.I64P),
                      W._p(counts.reshape(-1), W.F64P), W._p(ps, W.F64P))
    dev._ck(dev.L.wfsa_dev_set_path_blocks(dev.h, C.byref(pb)))
    x = rng.normal(-1, 0.5, size=n)
    H = np.zeros((n, n))
    rmin = C.c_double()
    dev._ck(dev.L.wfsa_dev_hessian(dev.h, W._p(x, W.F64P), W._p(H, W.F64P), C.byref(rmin)))   # warm-up
    t = time.perf_counter()
    for _ in range(a.reps):
        dev._ck(dev.L.wfsa_dev_hessian(dev.h, W._p(x, W.F64P), W._p(H, W.F64P), C.byref(rmin)))
    dt = (time.perf_counter() - t) / a.reps
    # check a sample against numpy (float64; the device accumulates 64-bit fixed point)
    Href = np.zeros((n, n))
    for b in range(nb):
        if b >= 2000:
            break
    sample = min(nb, 2000)
    dev2_H = np.zeros((n, n))
    pb2 = W.PathBlocks(sample, W._p(po[:sample + 1], W.I64P), W._p(co[:sample + 1], W.I64P), W._p(cols.reshape(-1), W.I32P),
                       W._p(vo[:sample + 1], W.I64P), W._p(counts.reshape(-1), W.F64P), W._p(ps, W.F64P))
    dev._ck(dev.L.wfsa_dev_set_path_blocks(dev.h, C.byref(pb2)))
    dev._ck(dev.L.wfsa_dev_hessian(dev.h, W._p(x, W.F64P), W._p(dev2_H, W.F64P), C.byref(rmin)))
    for b in range(sample):
        c = cols[b]
        M = counts[b]
        s = M @ x[c]
        r = np.exp(s - s.max())
        r /= r.sum()
        g = M.T @ r
        Href[np.ix_(c, c)] += ps[b] * (np.outer(g, g) - M.T @ (r[:, None] * M))
    ok = bool(np.allclose(dev2_H, Href, rtol=1e-9, atol=1e-13))
    flop = 2.0 * L * D * D * nb
    print(json.dumps({"blocks": nb, "paths": L, "cols": D, "n": n, "ms_per_call": dt * 1e3, "blocks_per_s": nb / dt,
                      "contraction_gflops": flop / dt / 1e9, "sample_matches_numpy": ok}))
    dev.close()


if __name__ == "__main__":
    main()
