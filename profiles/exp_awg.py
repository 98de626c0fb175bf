"""Evaluation time around the shared-memory limit of the arc weights: usage  python profiles/exp_awg.py [n_strings]
512 states (weights in shared memory), 1024 states (weights in HBM/L2: the AWG instances), and 1024 states on the CTA-per-string
kernel the library fell back to before (forced, on a 20 k-string sample)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "w-fsa_b200", "python"))
import numpy as np
import wfsa_b200 as W
from wfsa_b200 import synth

n_strings = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
for n_states, kernel, ns in ((512, 0, n_strings), (1024, 0, n_strings), (1024, 2, 20000)):
    model = synth.make_model(n_states, 64, 8, 4, seed=77)
    low = model.lowered()
    offs, toks, w = model.corpus(ns, 32, 128, seed=78)
    low.set_tokens(offs, toks, w / w.sum())
    dev = W.Device(low, force_kernel=kernel)
    rec, pc, used = dev.structure()
    trimmed = np.where(used > 0, 0, -2).astype(np.int32)
    n = 0
    for i in range(len(trimmed)):
        if trimmed[i] == 0:
            trimmed[i] = n; n += 1
    dev.set_param_map(trimmed, n, rec)
    x = np.random.RandomState(0).normal(-1.0, 0.3, size=n)
    dev.upload_x(x)
    for _ in range(3):
        dev.eval_launch()
    dev.sync()
    steps = 10
    dev.timer_begin_steps()
    for _ in range(steps):
        dev.l2_flush()
        dev.eval_launch()
    dev.timer_end()
    ms, k = dev.timer_step_ms()
    info = dev.info()
    print(json.dumps({"states": n_states, "combined_arcs": info["n_arcs"], "kernel": info["kernel"], "strings": ns, "symbols": int(offs[-1]),
                      "ms_per_step": ms / k, "symbols_per_s": float(offs[-1]) / (ms / k * 1e-3), "smem_bytes": info["smem_bytes"]}), flush=True)
    dev.close()
