"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C ABI
(libwfsa_b200.so): the Session tests drive the C++ host mirror of the reference's Learner
classes, the Device tests call wfsa_dev_* directly.  Checked against
  * golden vectors produced by the unmodified reference (tests/golden, oracle/make_golden.py),
  * the CPU oracle (oracle/wfsa_oracle.c, itself pinned by tests/test_cpu_oracle.py),
  * size-independent properties at BASELINE.json's full config-4 size.
Tolerances: log q, KL: 1e-9 relative (observed ~1e-15); gradient / H_f: 1e-9 relative per
component with a floor of 1e-6 * max|.| (the accumulators are 64-bit fixed point);
learned weights after a fixed epoch count: 1e-6 absolute (north_star)."""
import math

import numpy as np
import pytest

import wfsa_b200 as W
from helpers import fnum, golden_to_mine, good_cases, key, load_cases, vec_tol_ok
from oracle import oracle as O
from wfsa_b200 import synth

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------------------------
# Sessions against the reference's golden vectors
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", good_cases(), ids=lambda c: c["name"])
def test_structure_objective_gradient_hessian(case):
    s = W.Session(case["fsa_text"], case["corpus_text"], "Hessian")
    d = s.describe()
    assert d["strings"] == case["strings"] and fnum(d["paths"]) == case["paths"]
    assert abs(fnum(d["common_support"]) - fnum(case["common_support"])) < 1e-14
    assert d["unique_paths"] == case["unique_paths"] and d["n"] == case["n"] and d["k"] == case["k"]
    assert [int(fnum(pc)) for _, pc in d["shard"]] == [w["paths"] for w in case["corpus"]]
    mine = {key(e): e for e in d["edges"]}
    for e in case["edges"]:
        m = mine[key(e)]["trimmed"]
        assert (m == -2) == (e["trimmed"] == -2) and (m == -1) == (e["trimmed"] == -1), key(e)
    perm = golden_to_mine(case, d["edges"])
    # x read from the file (after Trim compaction)
    xf = np.array([fnum(v) for v in d["x"]])
    assert np.array_equal(xf[perm], np.array([fnum(v) for v in case["x_file"]]))
    for ev in case["evals"]:
        x = np.zeros(case["n"])
        x[perm] = [fnum(v) for v in ev["x"]]
        r = s.eval(x)
        ref_lq = np.array([fnum(v) for v in ev["logq"]])
        assert np.allclose(r["logq"], ref_lq, rtol=1e-9, atol=1e-12)
        assert abs(r["kl"] - fnum(ev["kl"])) <= 1e-9 * max(1.0, abs(fnum(ev["kl"])))
        ok, err = vec_tol_ok(r["grad"][perm], [fnum(v) for v in ev["grad"]], 1e-9)
        assert ok, err
        if "Hf" in ev:
            Hg = np.array(ev["Hf"], dtype=float)
            Hg = np.triu(Hg) + np.triu(Hg, 1).T
            H = s.hessian(x)[np.ix_(perm, perm)]
            assert np.allclose(H, Hg, rtol=1e-9, atol=1e-11 * max(1.0, np.abs(Hg).max()))
    s.close()


# Hessian-optimiser runs whose KKT matrix is singular along a non-identifiable direction (the optimum is a set, not a point):
# the Newton step along the null space is rounding noise shaped by the linear solver's pivoting, so the WEIGHTS of these
# runs are compared to 1e-5 / 1e-3 only; every printed column is still compared tightly.  Every other run: 1e-6 (north_star).
SINGULAR_KKT = {
    "talk.wfsa+talk.corpus",   # the reference's own optimum differs between its two optimisers: weights are not identifiable
    "random00",                # measured: the only other runs that leave 1e-6 (3 of 117 runs); both random automata have a state
    "random34",                # whose emissions never disambiguate two transitions, i.e. a flat direction of the objective
}


def _runs():
    out = []
    for case in good_cases():
        for i, run in enumerate(case.get("runs", [])):
            if "reference_error" in run:
                continue
            out.append(pytest.param(case, run, id="%s-%s-i%d-e%d" % (case["name"], run["optimizer"], run["flags"], run["epochs"])))
    return out


@pytest.mark.parametrize("case,run", _runs())
def test_optimisation_trajectory(case, run):
    """Same optimiser, same -i flags, same epochs: KL / graderr / g / lambda per epoch and the learned
    weights must follow the reference's run (src/main.cpp:271-304)."""
    s = W.Session(case["fsa_text"], case["corpus_text"], run["optimizer"])
    d = s.describe()
    perm = golden_to_mine(case, d["edges"])
    s.init(run["flags"])
    x0 = s.x()
    assert np.allclose(x0[perm], [fnum(v) for v in run["x_init"]][:case["n"]], rtol=1e-12, atol=1e-12)
    hess = run["optimizer"] == "Hessian"
    # QuasiNewton is a diagonal iteration: weights must agree to 1e-6 (north_star).  The Hessian optimiser
    # solves an indefinite KKT system that is singular along non-identifiable directions (talk, random34:
    # the optimum is not unique); there the step along the null space is whatever the linear
    # solver's pivoting makes of rounding noise (MKL DSS, the oracle's stand-in and our Bunch-Kaufman all
    # differ), so weights are compared to 1e-3 / 1e-5 relative while every printed column (KL, graderr,
    # g_min, g_max, lambda_min, inertia, rmin) is still compared tightly at every epoch.
    singular = hess and case["name"] in SINGULAR_KKT
    xtol = dict(rtol=1e-5, atol=1e-3) if singular else dict(rtol=1e-6, atol=1e-6)
    ref_rows = run["trajectory"]
    n_cmp = len(ref_rows) - (1 if run["error"] else 0)       # the row that made the reference stop is not compared
    halted = False
    for row in ref_rows[:n_cmp]:
        info = s.step(run["eta"])
        ref = [fnum(v) for v in row["info"]]
        cols = [0, 1, 2, 3, 6] if hess else [0, 1, 2, 3, 4]
        for c in cols:
            assert math.isclose(info[c], ref[c], rel_tol=2e-6, abs_tol=2e-8), (row["epoch"], c, info[c], ref[c])
        if hess and not run["error"]:
            assert (info[4], info[5]) == (ref[4], ref[5]), ("inertia", row["epoch"])
            if (run["flags"] & 8) and not case["unique_paths"]:      # smallest path posterior
                assert math.isclose(info[7], ref[7], rel_tol=2e-6, abs_tol=1e-12), ("rmin", row["epoch"])
        if not hess and not case["unique_paths"]:                    # QuasiNewton prints the smallest path posterior as well
            assert math.isclose(info[5], ref[5], rel_tol=2e-6, abs_tol=1e-12), ("rmin", row["epoch"], info[5], ref[5])
        if "x" in row:
            xr = np.array([fnum(v) for v in row["x"]][:case["n"]])
            assert np.allclose(s.x()[perm], xr, **xtol), row["epoch"]
        halted = s.halt(run["tol"])
        if halted and run["tol"] > 0:
            break
    if not run["error"]:
        # with -tol 0 halting means graderr == g == 0 EXACTLY: the reference's path algebra leaves 1e-17 of
        # rounding noise there, a lattice whose posteriors are exactly 1 does not -- not a comparable quantity
        assert run["tol"] == 0 or halted == run["halted"]
        dump = W.parse(s.dump(True), case["corpus_text"])
        mine = {key(e): fnum(e["file_logprob"]) for e in dump["edges"]}
        for e in run["final_edges"]:
            ref = fnum(e["logprob"])
            if math.isinf(ref):
                assert mine[key(e)] == ref
            else:
                assert abs(mine[key(e)] - ref) <= xtol["atol"] + xtol["rtol"] * abs(ref), (key(e), mine[key(e)], ref)
    s.close()


def test_evaluation_result_line():
    """-eval output of HessianLearner (src/HessianLearner.cpp:349-372) on talk: golden values in SURVEY.md 8c."""
    case = [c for c in load_cases("fixtures") if c["name"] == "talk.wfsa+talk.corpus"][0]
    s = W.Session(case["fsa_text"], case["corpus_text"], "Hessian")
    s.init(31)
    for _ in range(20):
        s.step(1.0)
        if s.halt(1e-6):
            break
    s.renormalize()
    r = s.result()
    ref = [-0.27031007207211, 0.27031007207211, 0.549306144334055, 0.346573590279973, -38.3012866549895, 3.58351893845611, 4, 1]
    assert np.allclose(r[[0, 1, 2, 3, 5, 6, 7]], np.array(ref)[[0, 1, 2, 3, 5, 6, 7]], rtol=1e-9, atol=1e-12)
    # r[4] is the log-determinant of a singular Hessian (the optimum is not unique): pure rounding noise,
    # solver dependent: the reference build prints -38.3 (|det| ~ 1e-17) here and inf -- its
    # answer for a non-positive determinant, src/Utils.cpp:349-351 -- on test5; both are the same statement
    assert r[4] == np.inf or -60 < r[4] < -30
    s.close()


def _eval_runs():
    out = []
    cases = {c["name"]: c for c in load_cases("fixtures") + load_cases("random")}
    for r in load_cases("eval"):
        if r["returncode"] == 0 and r["name"] in cases and r["optimizer"] == "Hessian":      # (QuasiNewton prints an empty Result line)
            out.append(pytest.param(cases[r["name"]], r, id="%s-%s-i%d" % (r["name"], r["optimizer"], r["flags"])))
    return out


@pytest.mark.parametrize("case,run", _eval_runs())
def test_evaluation_result_line_all_cases(case, run):
    """The `-eval` Result line (src/main.cpp:306-323; HessianLearner::GetOptimizationResult, src/HessianLearner.cpp:349-372:
    KL, -log common support, log model volume, log auxiliary volume, log det of the Hessian, log det of the auxiliary Hessian,
    free parameters, auxiliary parameters; QuasiNewtonLearner prints an empty line) after the reference's own run
    `-opt O -i F -e E -tol T -n -eval`, for every fixture and every random automaton the reference accepts
    (tests/golden/eval.json, written by oracle/make_golden_eval.py from the unmodified reference build)."""
    s = W.Session(case["fsa_text"], case["corpus_text"], run["optimizer"])
    s.init(run["flags"])
    for _ in range(run["epochs"]):
        s.step(1.0)
        if s.halt(run["tol"]):
            break
    s.renormalize()
    r = s.result()
    ref = [fnum(v) for v in run["result"]]
    assert len(r) == 8 and len(ref) == 8
    for i in (0, 1, 2, 3, 5, 6, 7):
        assert math.isclose(r[i], ref[i], rel_tol=1e-6, abs_tol=1e-8), (i, r[i], ref[i])
    # r[4], the log-determinant of the n x n Hessian in the weights (src/HessianLearner.cpp:219-260), is +inf for a non-positive
    # determinant (src/Utils.cpp:349-351).  It is compared when both sides call it positive: at the stopping tolerance of the
    # run (1e-6 on the gradient) it agrees to ~1e-3.  Where the Hessian has an eigenvalue at rounding level (non-identifiable
    # directions) the SIGN of the determinant is noise of the factorisation, and one side may print inf: 2 of the 77 runs
    # (random11 -i 7, random24 -i 31), both with |log det| of the other side far from zero.
    if math.isfinite(ref[4]) and math.isfinite(r[4]) and case["name"] not in SINGULAR_KKT:
        assert math.isclose(r[4], ref[4], rel_tol=2e-3, abs_tol=2e-3), (r[4], ref[4])
    s.close()


@pytest.mark.parametrize("case", [c for c in load_cases("fixtures") + load_cases("random") if c.get("degenerate")],
                         ids=lambda c: c["name"])
def test_degenerate_inputs(case):
    """"Empty automaton!" / no recognised string: the reference exits 1 (src/main.cpp:217-228)."""
    s = W.Session(case["fsa_text"], case["corpus_text"], "QuasiNewton")
    d = s.describe()
    assert d["degenerate"] and d["strings"] == case["strings"] and fnum(d["paths"]) == case["paths"]
    assert d["degenerate_message"] == ("Empty automaton!" if d["n"] == 0 else "Automaton cannot generate any of the strings!")
    with pytest.raises(W.WfsaError) as ei:
        s.init(7)
    assert ei.value.code == 102
    s.close()


def test_epsilon_cycle_is_rejected():
    fsa = "\n^\n$\n^  0\n^ a 0\na  0 x 0\na a 0 $ 0\n"          # state a emits "" and loops on itself
    with pytest.raises(W.WfsaError) as ei:
        W.Session(fsa, "\nx 1\n", "QuasiNewton")
    assert "cycle of empty emissions" in ei.value.message


# ----------------------------------------------------------------------------------------------
# The raw device ABI against the CPU oracle
# ----------------------------------------------------------------------------------------------
def build_device(low, **kw):
    dev = W.Device(low, **kw)
    rec, pc, used = dev.structure()
    params = np.concatenate([low.trans_param, low.emis_param])
    edge_used = np.array([r < 0 or used[r] for r in params])
    trimmed, n, Ccol = O.trim(low, edge_used)
    dev.set_param_map(trimmed, n, rec)
    return dev, rec, pc, trimmed, n


def oracle_grad(low, trimmed, n, ee):
    params = np.concatenate([low.trans_param, low.emis_param])
    g = np.zeros(n)
    for e, r in enumerate(params):
        if r >= 0 and trimmed[r] >= 0:
            g[trimmed[r]] = -ee[e]
    return g


@pytest.fixture(scope="module")
def medium():
    model = synth.make_model(256, 64, 8, 4, seed=11)
    low = model.lowered()
    offs, toks, w = model.corpus(3000, 32, 128, seed=12)
    # a few strings that are not recognised: unknown symbol, random junk
    toks = toks.copy()
    toks[offs[5]] = -1
    rng = np.random.RandomState(5)
    toks[offs[9]:offs[10]] = rng.randint(0, 64, size=offs[10] - offs[9])
    low.set_tokens(offs, toks, w / w.sum())
    return model, low


@pytest.mark.parametrize("kernel,accum,variant", [(6, 0, 0), (6, 0, 4 << 16), (5, 0, 0), (5, 0, 4 << 16), (5, 0, 4), (4, 0, 0), (4, 0, 4 << 16), (1, 1, 0), (1, 1, 1), (1, 2, 0), (2, 0, 0), (7, 0, 0), (3, 0, 0)])
def test_kernels_match_cpu_oracle(medium, kernel, accum, variant):
    model, low = medium
    count = 3000 if kernel != 3 else 300          # the generic kernel is the slow, dense one
    dev, rec, pc, trimmed, n = build_device(low, force_kernel=kernel, accum_mode=accum, accum_variant=variant, count=count)
    assert dev.info()["kernel"] == kernel
    zt, ze = np.zeros(low.n_trans), np.zeros(low.n_emis)
    opc, _, _ = O.dp_eval(low, zt, ze, count=count, want_grad=False, want_counts=True)
    assert np.array_equal(rec.astype(bool), opc > 0) and not rec[5] and not rec[9]
    assert np.allclose(pc, opc, rtol=1e-12)
    rng = np.random.RandomState(1)
    for x in (np.zeros(n), rng.normal(-1.5, 0.8, size=n)):
        ll, logq, grad = dev.eval(x)
        ltw, lew = low.edge_logweights(x, trimmed)
        _, olq, oee = O.dp_eval(low, ltw, lew, count=count)
        r = rec.astype(bool)
        assert np.allclose(logq[r], olq[r], rtol=1e-12, atol=1e-10)
        assert np.all(np.isneginf(logq[~r]))
        p = low.p[:count]
        assert abs(ll - float(np.sum(p[r] * olq[r]))) <= 1e-10 * abs(ll)
        # oracle gradient restricted to recognised strings == all strings (unrecognised contribute 0)
        ok, err = vec_tol_ok(grad, oracle_grad(low, trimmed, n, oee), 1e-9)
        assert ok, err
    if kernel == 5:
        # 16 pool slots hold every string of this corpus; with 4 slots the ambiguous stretches overflow onto the
        # warp-per-string kernel; variant 4 disables the constant folding of bridge edges
        i = dev.info()
        assert i["n_active_strings"] == int(rec.sum()) and i["lattice_edges"] > 0
        assert (i["n_overflow_strings"] > 0) == (variant == 4 << 16)
        assert (i["lattice_bridge_edges"] > 0) == (variant != 4)
    if kernel == 6:
        # segmented form: bridges + merged region types; with 4 pool slots some regions do not fit and their
        # strings go to the warp-per-string kernel
        i = dev.info()
        assert i["n_active_strings"] == int(rec.sum()) and i["lattice_bridge_edges"] > 0
        assert 0 < i["seg_types"] <= i["seg_region_instances"] and i["seg_type_edges"] <= i["seg_region_edges"]
        assert (i["n_overflow_strings"] > 0) == (variant == 4 << 16)
    if kernel == 4:
        # K = 12 (default) handles every string of this corpus on the thread-per-string kernel;
        # K = 4 pushes the strings whose active set exceeds 4 states onto the warp-per-string kernel
        import ctypes as C
        i = W.DevInfo()
        dev._ck(dev.L.wfsa_dev_get_info(dev.h, C.byref(i)))
        assert i.n_active_strings == int(rec.sum())
    dev.close()


def test_segmented_objective_without_per_string_pass(medium):
    """Kernel 6 finishes [loglik, grad] from the region types and the folded bridge constants alone; log q of every
    string is computed only when it is fetched (ks_strings on demand) and must agree with that log-likelihood."""
    model, low = medium
    dev, rec, pc, trimmed, n = build_device(low, force_kernel=6)
    r = rec.astype(bool)
    p = low.p[:len(rec)]
    rng = np.random.RandomState(3)
    with pytest.raises(W.WfsaError):                    # nothing has been evaluated for this parameter map yet
        dev.eval_fetch()
    for x in (rng.normal(-1.0, 0.5, size=n), rng.normal(-12.0, 3.0, size=n)):
        dev.upload_x(x)
        dev.eval_launch()
        ll0, g0 = dev.eval_fetch()                      # objective and gradient only
        ll1, logq, g1 = dev.eval_fetch(want_logq=True)  # now the per-string pass runs, on the same evaluation
        ll2, logq2, g2 = dev.eval_fetch(want_logq=True) # and is not repeated
        assert ll0 == ll1 == ll2 and np.array_equal(g0, g1) and np.array_equal(g0, g2) and np.array_equal(logq, logq2)
        assert abs(float(np.dot(p[r], logq[r])) - ll0) <= 1e-12 * abs(ll0)
        ll3, logq3, g3 = dev.eval(x)                    # the one-call form
        assert ll3 == ll0 and np.array_equal(g3, g0) and np.array_equal(logq3, logq)
        ltw, lew = low.edge_logweights(x, trimmed)
        _, olq, oee = O.dp_eval(low, ltw, lew, count=len(rec))
        assert np.allclose(logq[r], olq[r], rtol=1e-12, atol=1e-10)
        assert abs(ll0 - float(np.sum(p[r] * olq[r]))) <= 1e-10 * abs(ll0)
    dev.close()


def test_accumulation_variants_are_bitwise_identical(medium):
    """64-bit fixed-point accumulation: shared-memory (split / CAS) and global REDs, and repeated runs,
    give the same bits; two half-shards add up to the whole."""
    model, low = medium
    outs = []
    rng = np.random.RandomState(2)
    x = None
    for accum, variant in ((1, 0), (1, 1), (2, 0), (1, 0)):
        dev, rec, pc, trimmed, n = build_device(low, force_kernel=1, accum_mode=accum, accum_variant=variant)
        if x is None:
            x = rng.normal(-1.0, 0.5, size=n)
        outs.append(dev.eval(x))
        dev.close()
    for o in outs[1:]:
        assert o[0] == outs[0][0] and np.array_equal(o[2], outs[0][2]) and np.array_equal(o[1], outs[0][1])
    # additivity over shards (what the multi-GPU all-reduce relies on)
    tot_ll, tot_g = 0.0, np.zeros_like(outs[0][2])
    for first, count in ((0, 1400), (1400, 1600)):
        dev = W.Device(low, force_kernel=1, first=first, count=count)
        rec, pc, used = dev.structure()
        dev.set_param_map(trimmed, n, rec)      # the map of the whole corpus (used flags are all-reduced in real runs)
        ll, _, g = dev.eval(x)
        tot_ll += ll
        tot_g += g
        dev.close()
    assert abs(tot_ll - outs[0][0]) <= 1e-13 * abs(outs[0][0])
    assert np.allclose(tot_g, outs[0][2], rtol=1e-13, atol=1e-18)


def test_edge_cases():
    # emissions all one token; "" is in the language (start -> end); b is a dead end
    fsa = "\n^\n$\n^  0\n^ a -0.5 b -1 $ -2\na x -1 y -0.3\na a -0.7 $ -0.9\nb x 0\nb b 0\n"
    words = [("", 2.0), ("x", 1.0), ("xy", 1.0), ("zz", 1.0), ("xq", 3.0), ("yyyy", 1.0)]
    d = W.parse(fsa, "\nx 1\n")
    low = W.Lowered(d, corpus=words)
    for kernel in (6, 5, 4, 1, 2, 3):
        dev, rec, pc, trimmed, n = build_device(low, force_kernel=kernel)
        assert rec.tolist() == [1, 1, 1, 0, 0, 1] and pc.tolist() == [1, 1, 1, 0, 0, 1]
        x = np.array([-0.4, -1.1, -0.2, -0.6, -0.8, -1.3])[:n]
        ll, logq, grad = dev.eval(x)
        ltw, lew = low.edge_logweights(x, trimmed)
        _, olq, oee = O.enum_eval(low, ltw, lew)
        assert np.allclose(logq[rec > 0], olq[rec > 0], rtol=1e-13) and np.all(np.isneginf(logq[rec == 0]))
        ok, err = vec_tol_ok(grad, oracle_grad(low, trimmed, n, oee), 1e-9)
        assert ok, (kernel, err)
        dev.close()
    # an empty shard is legal (a rank may own no strings)
    low0 = W.Lowered(d, corpus=[])
    dev = W.Device(low0)
    rec, pc, used = dev.structure()
    assert len(rec) == 0 and not used.any()
    dev.set_param_map(np.full(low0.n_raw, -2, dtype=np.int32), 0, rec)
    ll, _, g = dev.eval(np.zeros(0))
    assert ll == 0.0 and len(g) == 0
    dev.close()


def test_long_strings_need_rescaling():
    """4000-token strings with small weights: q ~ e^-14000, far below DBL_MIN; the lazy power-of-two
    rescaling must keep log q and the posteriors exact (the reference underflows here)."""
    model = synth.make_model(64, 16, 4, 3, seed=21)
    low = model.lowered()
    offs, toks, w = model.corpus(40, 3500, 4000, seed=22)
    low.set_tokens(offs, toks, w / w.sum())
    for kernel in (6, 5, 4, 1, 2):
        dev, rec, pc, trimmed, n = build_device(low, force_kernel=kernel)
        assert rec.all()
        if kernel == 5:          # streams of up to 8192 words stay on the compiled-lattice kernel
            assert dev.info()["n_overflow_strings"] < len(rec)
        rng = np.random.RandomState(4)
        for x in (rng.normal(-3.0, 1.0, size=n), rng.normal(+2.5, 1.0, size=n)):
            ll, logq, grad = dev.eval(x)
            ltw, lew = low.edge_logweights(x, trimmed)
            _, olq, oee = O.dp_eval(low, ltw, lew)
            assert np.all(np.isfinite(logq)) and np.allclose(logq, olq, rtol=1e-11)
            ok, err = vec_tol_ok(grad, oracle_grad(low, trimmed, n, oee), 1e-9)
            assert ok, (kernel, err)
        dev.close()


@pytest.mark.parametrize("kernel", [2, 7])
def test_config5_shape_dense_kernels(kernel):
    """Down-scaled config 5 (dense: more than 32 states emit a symbol): CTA-per-string kernel K3 and the
    position-synchronous pair-batched kernel K7, each against the CPU forward-backward; K7 runs K3's sums in K3's order, so the two agree to the last bits."""
    model = synth.make_model(512, 64, 16, 8, seed=31)       # 64 states per symbol
    low = model.lowered()
    offs, toks, w = model.corpus(400, 16, 48, seed=32)
    toks = toks.copy()
    toks[offs[7] + 3] = -1                                  # one string with an unknown symbol
    low.set_tokens(offs, toks, w / w.sum())
    dev, rec, pc, trimmed, n = build_device(low, force_kernel=kernel)
    assert dev.info()["kernel"] == kernel and rec.sum() == 399
    x = np.random.RandomState(6).normal(-1.0, 0.4, size=n)
    ll, logq, grad = dev.eval(x)
    ltw, lew = low.edge_logweights(x, trimmed)
    _, olq, oee = O.dp_eval(low, ltw, lew)
    r = rec.astype(bool)
    assert np.allclose(logq[r], olq[r], rtol=1e-12) and np.isneginf(logq[~r]).all()
    ok, err = vec_tol_ok(grad, oracle_grad(low, trimmed, n, oee), 1e-9)
    assert ok, err
    assert abs(ll - float(np.dot(low.p[r], olq[r]))) <= 1e-11 * abs(ll)
    if kernel == 7:
        d3 = W.Device(low, force_kernel=2)
        d3.structure()
        d3.set_param_map(trimmed, n, rec)
        ll3, logq3, grad3 = d3.eval(x)
        d3.close()
        # (identical sums in identical order, except the association of one product in the posterior of the final
        #  transitions: equal to the last bits, not bitwise)
        assert np.allclose(grad3, grad, rtol=1e-13, atol=0) and np.allclose(logq3[r], logq[r], rtol=1e-15)
    dev.close()


def test_config5_true_shape_parity():
    """BASELINE config 5 at its named shape (4096 states / 256 symbols, 64 successors, 16 emissions: ~256 states per symbol,
    4.2 M combined arcs): 300 strings through the batched kernel K7 against the CPU forward-backward, with a small
    lattice budget so that the corpus is cut into several batches, and long strings so that the power-of-two
    rescaling is exercised."""
    import os
    model = synth.make_model(4096, 256, 64, 16, seed=4321)
    low = model.lowered()
    offs, toks, w = model.corpus(300, 32, 128, seed=41)
    low.set_tokens(offs, toks, w / w.sum())
    os.environ["WFSA_K7_ROWS"] = "9000"
    try:
        dev, rec, pc, trimmed, n = build_device(low, force_kernel=7)     # (the default for this shape from 200 000 strings on)
    finally:
        os.environ.pop("WFSA_K7_ROWS", None)
    info = dev.info()
    assert info["kernel"] == 7 and rec.all() and info["pool_slots"] >= 2     # (pool_slots reports the number of batches)
    x = np.random.RandomState(8).normal(-3.0, 1.0, size=n)                  # small weights: alpha underflows without rescaling
    ll, logq, grad = dev.eval(x)
    ltw, lew = low.edge_logweights(x, trimmed)
    _, olq, oee = O.dp_eval(low, ltw, lew)
    assert np.allclose(logq, olq, rtol=1e-12)
    ok, err = vec_tol_ok(grad, oracle_grad(low, trimmed, n, oee), 1e-9)
    assert ok, err
    ll2, _, grad2 = dev.eval(x, want_logq=False)
    assert ll2 == ll and np.array_equal(grad2, grad), "run-to-run bitwise determinism"
    dev.close()


def test_full_size_config4_properties():
    """BASELINE config 4 at full size (1M strings, ~80M tokens): size-independent identities (every position emits exactly
    one symbol and takes exactly one transition, +1 into the end state), run-to-run bitwise determinism, a 20k-string
    sample of log q, and -- at a random x -- log q of every string, the log-likelihood and the whole gradient against the
    CPU forward-backward over the full corpus."""
    model = synth.make_model(256, 64, 8, 4, seed=1234)
    low = model.lowered()
    offs, toks, w = model.corpus(1000000, 32, 128, seed=1235)
    p = w / w.sum()
    low.set_tokens(offs, toks, p)
    dev, rec, pc, trimmed, n = build_device(low)
    assert rec.all() and dev.info()["kernel"] == 6 and dev.info()["n_overflow_strings"] == 0
    assert n == low.n_raw            # nothing is trimmed at this size
    x = np.zeros(n)
    ll, logq, grad = dev.eval(x)
    lens = np.diff(offs)
    params = np.concatenate([low.trans_param, low.emis_param])
    is_emis = np.concatenate([np.zeros(low.n_trans, bool), np.ones(low.n_emis, bool)])
    ge = -sum(grad[trimmed[r]] for r, em in zip(params, is_emis) if r >= 0 and em)
    gt = -sum(grad[trimmed[r]] for r, em in zip(params, is_emis) if r >= 0 and not em)
    assert abs(ge - float(np.sum(p * lens))) <= 1e-10 * ge
    assert abs(gt - float(np.sum(p * (lens + 1)))) <= 1e-10 * gt
    assert abs(ll - float(np.sum(p * logq))) <= 1e-10 * abs(ll)
    ll2, logq2, grad2 = dev.eval(x)
    assert ll2 == ll and np.array_equal(grad, grad2) and np.array_equal(logq, logq2)
    ltw, lew = low.edge_logweights(x, trimmed)
    _, olq, _ = O.dp_eval(low, ltw, lew, first=500000, count=20000, want_grad=False)
    assert np.allclose(logq[500000:520000], olq, rtol=1e-12)
    # ... and the whole evaluation at a random x against the CPU forward-backward over ALL 1M strings (the oracle takes a
    # few seconds on the host cores): log q of every string, the log-likelihood and every gradient component, 1e-9
    x = np.random.RandomState(3).normal(-1.0, 0.3, size=n)
    ll, logq, grad = dev.eval(x)
    ltw, lew = low.edge_logweights(x, trimmed)
    _, olq, oee = O.dp_eval(low, ltw, lew, nthreads=O.max_threads())
    assert np.allclose(logq, olq, rtol=1e-11)
    assert abs(ll - float(np.dot(p, olq))) <= 1e-11 * abs(ll)
    ok, err = vec_tol_ok(grad, oracle_grad(low, trimmed, n, oee), 1e-9)
    assert ok, err
    dev.close()


@pytest.mark.parametrize("n_states,expect_segmented,variant", [(512, True, 0), (1024, True, 0), (1024, True, 4 << 16), (2048, False, 0)])
def test_automata_around_the_arc_limits(n_states, expect_segmented, variant):
    """The limits of DESIGN.md section 5.  512 states (16.9 k combined arcs, 135 KB of weights): the segmented path with the
    weights in shared memory.  1024 states (32.8 k arcs, 262 KB of weights): the same kernels reading the weights from HBM/L2
    (their AWG instances).  2048 states (65.6 k arcs): beyond the 16-bit arc ids of the compiled form, whatever the library picks
    instead.  Whatever runs, the results are the oracle's.  (Variant 4 << 16: four pool slots, so that some strings overflow and
    the evaluation takes the general pipeline -- kr_regions with the weights in HBM -- instead of the single-launch kernel.)"""
    model = synth.make_model(n_states, 64, 8, 4, seed=77)
    low = model.lowered()
    offs, toks, w = model.corpus(20000, 32, 128, seed=78)
    low.set_tokens(offs, toks, w / w.sum())
    dev, rec, pc, trimmed, n = build_device(low, accum_variant=variant)
    info = dev.info()
    assert rec.all() and (info["kernel"] == 6) == expect_segmented, info["kernel"]
    assert (info["n_overflow_strings"] > 0) == (variant != 0)
    x = np.random.RandomState(4).normal(-1.0, 0.3, size=n)
    ll, logq, grad = dev.eval(x)
    ltw, lew = low.edge_logweights(x, trimmed)
    _, olq, oee = O.dp_eval(low, ltw, lew, nthreads=O.max_threads())
    assert np.allclose(logq, olq, rtol=1e-11)
    assert abs(ll - float(np.dot(low.p, olq))) <= 1e-11 * abs(ll)
    ok, err = vec_tol_ok(grad, oracle_grad(low, trimmed, n, oee), 1e-9)
    assert ok, (info["kernel"], err)
    print("\n%d states, %d combined arcs: kernel %d" % (n_states, info["n_arcs"], info["kernel"]))
    dev.close()


def test_pool_overflow_at_scale():
    """100 000 config-4 strings with FOUR pool slots per region instead of sixteen: every string with a region that
    needs more goes to the warp-per-string kernel (the cliff of DESIGN.md section 5).  Results must not change; the
    time of the mixed evaluation is printed next to the all-segmented one so that the cost of the cliff is on record."""
    import time
    model = synth.make_model(256, 64, 8, 4, seed=1234)
    low = model.lowered()
    offs, toks, w = model.corpus(100000, 32, 128, seed=55)
    low.set_tokens(offs, toks, w / w.sum())
    out = {}
    for name, variant in (("16 slots", 0), ("4 slots", 4 << 16)):
        dev, rec, pc, trimmed, n = build_device(low, accum_variant=variant)
        info = dev.info()
        assert info["kernel"] == 6 and rec.all()
        x = np.random.RandomState(9).normal(-1.0, 0.3, size=n)
        dev.eval(x, want_logq=False)
        t0 = time.perf_counter()
        for _ in range(5):
            ll, _, grad = dev.eval(x, want_logq=False)
        ms = (time.perf_counter() - t0) / 5 * 1e3
        _, logq, _ = dev.eval(x)
        out[name] = (ll, logq, grad, info["n_overflow_strings"], ms)
        dev.close()
    assert out["16 slots"][3] == 0 and out["4 slots"][3] > 100
    print("\npool overflow at scale: %d of 100000 strings on the warp-per-string kernel, %.3f ms per evaluation against %.3f ms"
          % (out["4 slots"][3], out["4 slots"][4], out["16 slots"][4]))
    ltw, lew = low.edge_logweights(x, trimmed)
    _, olq, oee = O.dp_eval(low, ltw, lew, nthreads=O.max_threads())
    og = oracle_grad(low, trimmed, n, oee)
    for name in out:
        ll, logq, grad = out[name][:3]
        assert np.allclose(logq, olq, rtol=1e-11), name
        assert abs(ll - float(np.dot(low.p, olq))) <= 1e-11 * abs(ll), name
        ok, err = vec_tol_ok(grad, og, 1e-9)
        assert ok, (name, err)


@pytest.mark.parametrize("kernel,variant", [(0, 0), (1, 0), (0, 4 << 16)])
def test_two_ranks_nccl(kernel, variant):
    """Two processes, two GPUs, the library's own ncclAllReduce of exact integers (tests/nccl_worker.py):
    identical bits on every rank and equal to the single-device evaluation.  Skipped on a one-GPU box
    (tests/test_cpu_multirank.py covers the sharding and the integer all-reduce on gloo)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + kernel + (7 if variant else 0)), os.path.join(here, "nccl_worker.py"), str(kernel), str(variant)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "nccl_worker ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("optimizer,init", [("QuasiNewton", 7), ("Hessian", 15)])
def test_cli_two_gpus_same_table(optimizer, init, tmp_path):
    """`wfsa --gpus 2` (one forked process per GPU, the training loop of /root/reference/src/main.cpp:271-304 on every
    rank, sums over the ranks inside the backend) prints the same optimisation table and writes the same weights as one
    GPU.  Skipped on a one-GPU box."""
    import os
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    exe = os.path.join(root, "w-fsa_b200", "_build", "wfsa")
    fsa = tmp_path / "talk.wfsa"; corpus = tmp_path / "talk.corpus"
    fsa.write_text("\n^\n$\n^  0\n^ TALK_N 0 TALK_V 0\nTALK_V talk 0\nTALK_V TENSE 0\nTALK_N talk 0\nTALK_N PLUR 0\nPLUR  0 s 0\nPLUR $ 0\nTENSE  0 s 0 ed 0\nTENSE $ 0\n")
    corpus.write_text("\ntalk 1\ntalks 2\ntalked 1\ntalking 1\ntalkings 1\n")
    outs = []
    for gpus in (1, 2):
        r = subprocess.run([exe, "-a", str(fsa), "-c", str(corpus), "-opt", optimizer, "-i", str(init), "-e", "6", "--full", "--gpus", str(gpus)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        table = [ln.split() for ln in r.stderr.splitlines() if ln[:2].isdigit() or ln[:1].isdigit() and "\t" in ln]
        assert len(table) >= 3, r.stderr[-3000:]
        outs.append((table, r.stdout))
    (t1, w1), (t2, w2) = outs
    assert len(t1) == len(t2)
    for a, b in zip(t1, t2):
        assert np.allclose([float(v) for v in a], [float(v) for v in b], rtol=1e-6, atol=1e-9), (a, b)
    n1 = [float(v) for ln in w1.splitlines() for v in ln.split() if v.replace(".", "").replace("-", "").replace("e", "").replace("+", "").isdigit()]
    n2 = [float(v) for ln in w2.splitlines() for v in ln.split() if v.replace(".", "").replace("-", "").replace("e", "").replace("+", "").isdigit()]
    assert len(n1) == len(n2) and len(n1) > 0 and np.allclose(n1, n2, rtol=1e-6, atol=1e-9)


def _device_hessian(dev, x, n):
    import ctypes as C
    H = np.zeros((n, n)); rmin = C.c_double()
    xx = np.ascontiguousarray(x, dtype=np.float64)
    dev._ck(dev.L.wfsa_dev_hessian(dev.h, W._p(xx, W.F64P), W._p(H, W.F64P), C.byref(rmin)))
    return H, rmin.value


def test_hessian_from_region_types_matches_enumeration(medium):
    """H_f with NO path enumeration per string: the segmented backend takes the blocks of the contraction from its
    compiled region types (Cov adds over the independent regions of a string, wfsa_dev.cu build_type_blocks).
    Checked against the CPU enumeration of every path of every string (HessianLearner::ComputeHf,
    /root/reference/src/HessianLearner.cpp:498-547, restated in oracle/wfsa_oracle.c), 1e-9 per entry."""
    model, low = medium
    dev, rec, pc, trimmed, n = build_device(low, force_kernel=6)
    assert dev.info()["kernel"] == 6
    x = np.random.RandomState(21).normal(-1.0, 0.5, size=n)
    H, rmin = _device_hessian(dev, x, n)
    dev.close()
    ltw, lew = low.edge_logweights(x, trimmed)
    params = np.concatenate([low.trans_param, low.emis_param])
    edge_param = np.array([trimmed[r] if r >= 0 else -1 for r in params], dtype=np.int32)
    edge_param[edge_param < 0] = -1
    keep = np.where(rec > 0)[0]
    sub = type(low).__new__(type(low)); sub.__dict__.update(low.__dict__)
    offs = np.concatenate([[0], np.cumsum((low.offsets[1:] - low.offsets[:-1])[keep])])
    toks = np.concatenate([low.tokens[low.offsets[i]:low.offsets[i + 1]] for i in keep])
    sub.set_tokens(offs, toks, low.p[keep])
    Href = O.enum_hessian(sub, ltw, lew, edge_param, n, max_paths=20000000)
    assert np.allclose(H, H.T, rtol=0, atol=1e-15)
    ok, err = vec_tol_ok(H.reshape(-1), Href.reshape(-1), rtol=1e-9)
    assert ok, err
    assert 0.0 < rmin <= 0.5


def test_hessian_at_scale_without_path_enumeration():
    """100 000 config-4-shaped strings (about 1.5 M paths): the host enumeration of round 1 refused corpora beyond
    4e6 paths; the type-derived blocks have no such limit.  Properties: symmetric, negative semi-definite
    (H_f = -sum_s p_s Cov_s), rows of a constraint group sum to zero against the group's indicator ... and additive:
    the sum of H_f over two halves of the corpus equals H_f of the whole."""
    model = synth.make_model(256, 64, 8, 4, seed=1234)
    low = model.lowered()
    offs, toks, w = model.corpus(100000, 32, 128, seed=77)
    low.set_tokens(offs, toks, w / w.sum())
    dev, rec, pc, trimmed, n = build_device(low)
    assert dev.info()["kernel"] == 6 and pc.sum() > 1.0e6
    x = np.random.RandomState(5).normal(-1.0, 0.3, size=n)
    H, rmin = _device_hessian(dev, x, n)
    dev.close()
    assert np.isfinite(H).all() and np.abs(H - H.T).max() <= 1e-15 and 0.0 < rmin < 1e-3
    rng = np.random.RandomState(6)
    for _ in range(5):
        v = rng.normal(size=n)
        assert v @ H @ v <= 1e-12
    half = 50000
    Hs = np.zeros_like(H)
    for first, count in ((0, half), (half, 100000 - half)):
        d2 = W.Device(low, first=first, count=count)
        d2.structure()
        d2.set_param_map(trimmed, n, None)
        Hh, _ = _device_hessian(d2, x, n)
        d2.close()
        Hs += Hh
    # every block adds its entries rounded to the fixed-point quantum of H (2^-52 here): absolute error ~ blocks x quantum
    assert np.allclose(Hs, H, rtol=1e-9, atol=1e-13), np.abs(Hs - H).max()


def test_hessian_contraction_random_blocks():
    """K5 (FP64 tensor-core contraction) on random dense path blocks against numpy."""
    rng = np.random.RandomState(9)
    fsa = "\n^\n$\n^  0\n^ a 0 b 0\na x 0 y 0\na a 0 $ 0 b 0\nb x 0 y 0\nb b 0 $ 0 a 0\n"
    low = W.Lowered(W.parse(fsa, "\nx 1\nxy 1\n"))
    dev, rec, pc, trimmed, n = build_device(low)
    nb = 37
    path_off, col_off, val_off, cols, counts, ps = [0], [0], [0], [], [], []
    for b in range(nb):
        L, D = rng.randint(2, 40), rng.randint(1, n + 1)
        c = rng.choice(n, D, replace=False)
        M = rng.randint(0, 4, size=(L, D)).astype(float)
        cols.extend(c.tolist()); counts.extend(M.reshape(-1).tolist()); ps.append(rng.uniform(0.01, 0.2))
        path_off.append(path_off[-1] + L); col_off.append(len(cols)); val_off.append(len(counts))
    arr = lambda v, t: np.ascontiguousarray(np.array(v, dtype=t))
    po, co, vo = arr(path_off, np.int64), arr(col_off, np.int64), arr(val_off, np.int64)
    cl, ct, pp = arr(cols, np.int32), arr(counts, np.float64), arr(ps, np.float64)
    import ctypes as C
    pb = W.PathBlocks(nb, W._p(po, W.I64P), W._p(co, W.I64P), W._p(cl, W.I32P), W._p(vo, W.I64P), W._p(ct, W.F64P), W._p(pp, W.F64P))
    dev._ck(dev.L.wfsa_dev_set_path_blocks(dev.h, C.byref(pb)))
    x = rng.normal(-1, 0.7, size=n)
    H = np.zeros((n, n)); rmin = C.c_double()
    dev._ck(dev.L.wfsa_dev_hessian(dev.h, W._p(x, W.F64P), W._p(H, W.F64P), C.byref(rmin)))
    Href = np.zeros((n, n)); rm = np.inf
    for b in range(nb):
        c = cl[co[b]:co[b + 1]]
        M = ct[vo[b]:vo[b + 1]].reshape(path_off[b + 1] - path_off[b], len(c))
        s = M @ x[c]
        r = np.exp(s - s.max()); r /= r.sum()
        rm = min(rm, r.min())
        g = M.T @ r
        Href[np.ix_(c, c)] += pp[b] * (np.outer(g, g) - M.T @ (r[:, None] * M))
    assert np.allclose(H, Href, rtol=1e-9, atol=1e-11) and abs(rmin.value - rm) <= 1e-12
    dev.close()
