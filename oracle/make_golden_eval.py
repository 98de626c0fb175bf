#!/usr/bin/env python3
"""oracle/make_golden_eval.py -- the `-eval` Result line of the UNMODIFIED reference (oracle/_ref/wfsa_ref,
/root/reference/src/main.cpp:306-323, HessianLearner::GetOptimizationResult src/HessianLearner.cpp:349-372 and
QuasiNewtonLearner's) for every good case of tests/golden/{fixtures,random}.json -> tests/golden/eval.json.

    make -C oracle ref && python oracle/make_golden_eval.py

Runs in this container only (needs the reference build); the GPU tests read the committed JSON."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "wfsa_ref")
RUNS = [("Hessian", 31, 20, 1e-6), ("Hessian", 7, 20, 1e-6), ("QuasiNewton", 7, 30, 1e-6)]


def main():
    out = {"generator": "oracle/make_golden_eval.py: oracle/_ref/wfsa_ref -opt O -i F -e E -tol T -n -eval -s", "cases": []}
    for name in ("fixtures", "random"):
        for c in json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))["cases"]:
            if "reference_error" in c or c.get("degenerate"):
                continue
            with tempfile.TemporaryDirectory() as tmp:
                fa, fc = os.path.join(tmp, "a.wfsa"), os.path.join(tmp, "a.corpus")
                open(fa, "w", newline="").write(c["fsa_text"])
                open(fc, "w", newline="").write(c["corpus_text"])
                for opt, flags, epochs, tol in RUNS:
                    r = subprocess.run([REF, "-a", fa, "-c", fc, "-opt", opt, "-i", str(flags), "-e", str(epochs), "-tol", repr(tol), "-n", "-eval", "-s"],
                                       capture_output=True, text=True, timeout=600)
                    line = [ln for ln in r.stderr.splitlines() if ln.startswith("Result:")]
                    rows = [ln for ln in r.stderr.splitlines() if ln[:1].isdigit() and "\t" in ln]
                    out["cases"].append({"name": c["name"], "set": name, "optimizer": opt, "flags": flags, "epochs": epochs, "tol": tol,
                                         "returncode": r.returncode, "epochs_run": len(rows),
                                         "result": line[0].split()[1:] if line else None})
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "eval.json"), "w"), indent=0)
    print(len(out["cases"]), "runs,", sum(1 for x in out["cases"] if x["result"]), "with a Result line")


if __name__ == "__main__":
    sys.exit(main())
