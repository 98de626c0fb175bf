"""Experiment driver for the single-launch evaluation kernel: one corpus, several library configurations.
usage: python profiles/exp_eval6.py [n_strings] -- prints ms/step and the in-kernel phase split per configuration."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "w-fsa_b200", "python"))
import numpy as np
import wfsa_b200 as W
from wfsa_b200 import synth

n_strings = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
configs = json.loads(sys.argv[2]) if len(sys.argv) > 2 else [{}]
model = synth.make_model(256, 64, 8, 4, seed=1234)
low = model.lowered()
offs, toks, w = model.corpus(n_strings, 32, 128, seed=1235)
low.set_tokens(offs, toks, w / w.sum())
for cfg in configs:
    env = cfg.get("env", {})
    for k, v in env.items():
        os.environ[k] = str(v)
    variant = (cfg.get("replicas", 0) << 8) | ((cfg.get("threads", 0) // 32) << 24)
    dev = W.Device(low, accum_variant=variant)
    rec, pc, used = dev.structure()
    trimmed = np.where(used > 0, 0, -2).astype(np.int32)
    n = 0
    for i in range(len(trimmed)):
        if trimmed[i] == 0:
            trimmed[i] = n; n += 1
    dev.set_param_map(trimmed, n, rec)
    x = np.random.RandomState(0).normal(-1.0, 0.3, size=n)
    dev.upload_x(x)
    for _ in range(5):
        dev.eval_launch()
    dev.sync()
    dev.eval6_phases(True)
    steps = 20
    dev.timer_begin_steps()
    for _ in range(steps):
        if not cfg.get("no_flush"):
            dev.l2_flush()
        dev.eval_launch()
    dev.timer_end()
    ms, k = dev.timer_step_ms()
    ph = dev.eval6_phases(True) / steps / 1e3
    ll, g = dev.eval_fetch()
    # e2e
    t = []
    for _ in range(steps):
        dev.l2_flush(); dev.sync()
        t0 = time.perf_counter(); dev.eval(x, want_logq=False); t.append(time.perf_counter() - t0)
    info = dev.info()
    print(json.dumps({"cfg": cfg, "ms_per_step": ms / k, "phases_us": [round(v, 2) for v in ph], "e2e_ms": float(np.mean(t)) * 1e3,
                      "e2e_min_ms": float(np.min(t)) * 1e3, "block": info["block"], "loglik": ll, "gsum": float(np.abs(g).sum())}), flush=True)
    dev.close()
    for k in env:
        os.environ.pop(k, None)
