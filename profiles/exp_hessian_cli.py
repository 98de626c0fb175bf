"""The Hessian optimiser at config-4 shape through the `wfsa` executable: usage  python profiles/exp_hessian_cli.py [n_strings] [epochs]
Writes the synthetic automaton and corpus in the reference's text formats, runs `wfsa -opt Hessian -i 15` (uniform start, normalised,
multipliers initialised, H_f included) and prints the epoch table with wall-clock seconds per stage."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "w-fsa_b200", "python"))
from wfsa_b200 import synth

n_strings = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model = synth.make_model(256, 64, 8, 4, seed=1234)
offs, toks, w = model.corpus(n_strings, 32, 128, seed=1235)
with tempfile.TemporaryDirectory() as tmp:
    fa, fc = os.path.join(tmp, "c4.wfsa"), os.path.join(tmp, "c4.corpus")
    open(fa, "w").write(model.text())
    open(fc, "w").write(model.corpus_text(offs, toks, w))
    exe = os.path.join(ROOT, "w-fsa_b200", "_build", "wfsa")
    for opt, init in (("Hessian", "15"), ("Hessian", "7"), ("QuasiNewton", "7")):
        t0 = time.perf_counter()
        r = subprocess.run([exe, "-a", fa, "-c", fc, "-opt", opt, "-i", init, "-e", str(epochs), "-tol", "0", "-s"], capture_output=True, text=True)
        dt = time.perf_counter() - t0
        rows = [ln for ln in r.stderr.splitlines() if ln[:1].isdigit() and "\t" in ln]
        t1 = time.perf_counter()
        r0 = subprocess.run([exe, "-a", fa, "-c", fc, "-opt", opt, "-i", init, "-e", "0", "-s"], capture_output=True, text=True)
        d0 = time.perf_counter() - t1
        print("%s -i %s: rc %d, %d epochs, %.2f s in total, %.2f s without epochs -> %.3f s per epoch" % (opt, init, r.returncode, len(rows), dt, d0, (dt - d0) / max(len(rows), 1)))
        for ln in r.stderr.splitlines():
            if ln.startswith("epoch") or (ln[:1].isdigit() and "\t" in ln) or "trimming" in ln or "parameters" in ln or "paths:" in ln:
                print("   ", ln)
        if r.returncode != 0:
            print(r.stderr[-1500:])
