"""
GPU tests of the reference-facing boundary: JSON string in -> JSON string out
through the C ABI (include/arbplf.h), the Python mirror of the reference's
module, and the arbplf-* executables.

1. Every golden input/output pair shipped by the reference (tests/golden,
   SURVEY.md appendix B) must be reproduced to 1e-11 relative.
2. The reduction semantics (selection order, duplicates, sum / avg / only /
   weighted) are checked against the 320-bit oracle.
3. The reference's invariance tests (site weights vs duplicated sites, rate
   matrix scaling, relabelling) are replayed.
"""
import copy
import json
import os
import subprocess

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = H.manifest()
SUPPORTED = {"ll", "deriv", "marginal", "dwell", "trans", "em_update", "hess"}

# Absolute floors, only where the reference's exact arithmetic cancels to a
# value far below the magnitude of the terms (see tests/test_engine_gpu.py):
# value = tolerance on |got - want| in units of the largest |entry| of the table.
CANCEL_FLOOR = 4e-15


def _run(program, doc):
    import phyly_b200.arbplf as A
    return json.loads(getattr(A, "arbplf_" + program)(json.dumps(doc)))


def _check(got, want, what, rtol=1e-11):
    assert got["columns"] == want["columns"], what
    assert len(got["data"]) == len(want["data"]), what
    scale = max([abs(r[-1]) for r in want["data"]] + [1.0])
    for r1, r2 in zip(got["data"], want["data"]):
        assert r1[:-1] == r2[:-1], (what, r1, r2)
        a, b = r1[-1], r2[-1]
        assert isinstance(a, float), (what, r1)
        assert abs(a - b) <= rtol * abs(b) + CANCEL_FLOOR * scale, (what, r1, r2)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden(case):
    assert case["program"] in SUPPORTED
    got = _run(case["program"], H.golden_in(case["name"]))
    _check(got, H.golden_out(case["name"]), case["name"])


def test_golden_long_branch_derivatives_relative():
    """examples/JC.long.branch: naive fp64 loses these (4e-5 relative); the
    double-double matrix kernel keeps them to 1e-11 relative with NO absolute floor."""
    for name in ("jc_long_deriv", "jc29_same_deriv", "jc29_diff_deriv", "jc30_same_deriv", "jc30_diff_deriv"):
        got = _run("deriv", H.golden_in(name))
        want = H.golden_out(name)
        for r1, r2 in zip(got["data"], want["data"]):
            if r2[-1] != 0.0:
                assert abs(r1[-1] - r2[-1]) <= 1e-11 * abs(r2[-1]), (name, r1, r2)


def test_golden_exact_zeros():
    """Derivatives at data-free sites / edges are literal 0.0 (constant-column shortcut)."""
    got = _run("deriv", H.golden_in("fels_deriv"))
    want = H.golden_out("fels_deriv")
    for r1, r2 in zip(got["data"], want["data"]):
        if r2[0] == 0:
            assert r1[-1] == 0.0 and r2[-1] == 0.0
    got = _run("ll", H.golden_in("fels_ll2"))
    assert got["data"][0][-1] == 0.0


RCASES = H.reduction_cases()


@pytest.mark.parametrize("name,program,doc", RCASES, ids=[c[0] for c in RCASES])
def test_reductions_against_oracle(name, program, doc):
    want = H.expected_json(name, program, doc)
    got = _run(program, doc)
    _check(got, want, name)


def test_cli_executables():
    exe = os.path.join(ROOT, "phyly_b200", "bin")
    for name, prog in (("fels_ll", "ll"), ("bpp_deriv", "deriv"), ("beast_anc_marginal", "marginal"),
                       ("mj_rewards", "dwell"), ("mj_jumps", "trans"), ("fels_hess_leaf", "hess"), ("fels_em_leaf", "em-update")):
        text = json.dumps(H.golden_in(name))
        p = subprocess.run([os.path.join(exe, "arbplf-" + prog)], input=text, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        assert p.stdout.endswith("\n")
        _check(json.loads(p.stdout), H.golden_out(name), name)
    # errors: non-zero exit status, message on stderr, nothing on stdout (arbplf-ll.c:13-14)
    p = subprocess.run([os.path.join(exe, "arbplf-ll")], input='{"model_and_data": {}}', capture_output=True, text=True)
    assert p.returncode != 0 and p.stdout == "" and "error" in p.stderr
    # arbplf-hess without an aggregating site reduction (arbplfhess.c:1196-1200)
    p = subprocess.run([os.path.join(exe, "arbplf-hess")], input=json.dumps(H.golden_in("fels_ll")), capture_output=True, text=True)
    assert p.returncode != 0 and p.stdout == ""
    # the other second-order programs and the certified ll run through their executables too
    doc = json.dumps(H.golden_in("fels_hess_leaf"))
    for prog in ("inv-hess", "newton-delta", "newton-update", "newton-refine"):
        p = subprocess.run([os.path.join(exe, "arbplf-" + prog)], input=doc, capture_output=True, text=True)
        assert p.returncode == 0 and json.loads(p.stdout)["columns"][-1] == "value", (prog, p.stderr)
    p = subprocess.run([os.path.join(exe, "arbplf-ll-certified")], input=json.dumps(H.golden_in("fels_ll")), capture_output=True, text=True)
    out = json.loads(p.stdout)
    assert out["lower"]["data"][0][-1] <= -11.297288182875496 <= out["upper"]["data"][0][-1]


def test_output_formatting_matches_jansson():
    import phyly_b200.arbplf as A
    s = A.arbplf_ll(json.dumps(H.golden_in("fels_ll")))
    assert s == '{"columns": ["site", "value"], "data": [[0, -11.297288182875496]]}'
    s = A.arbplf_ll(json.dumps(H.golden_in("fels_ll2")))
    assert '[0, 0.0]' in s
    s = A.arbplf_dwell(json.dumps(H.golden_in("fels_dwell_adenine")))
    assert "7.4565883396650014e-6" in s or "e-6" in s


BAD_REDUCTIONS = [
    {"site_reduction": "hello"}, {"site_reduction": {"aggregation": {"a": 1}}}, {"site_reduction": {"aggregation": "hello"}},
    {"site_reduction": {"aggregation": [1, "x"]}}, {"site_reduction": {"aggregation": [1]}},
    {"site_reduction": {"aggregation": [1, 2, 3]}}, {"site_reduction": {"selection": {"a": 1}}},
    {"site_reduction": {"selection": "hello"}}, {"site_reduction": {"selection": [-1]}},
    {"site_reduction": {"selection": [2]}}, {"site_reduction": {"selection": [0.5]}},
    {"site_reduction": {"selection": ["a"]}}, {"site_reduction": {"selection": [{"a": 1}]}},
    {"site_reduction": {"selection": [0, 1, 1], "aggregation": [1, 2]}},
    {"site_reduction": {"selection": [0], "aggregation": [1, 2]}},
    {"site_reduction": {"selection": [0, 1], "aggregation": "only"}},
    {"site_reduction": {"selection": None}}, {"site_reduction": {"bogus": 1}}, {"bogus_reduction": {}},
]


@pytest.mark.parametrize("i", range(len(BAD_REDUCTIONS)))
def test_bad_reductions_are_rejected(i):
    # test_scripts/test_bad_reduction_args.py:31-112
    import phyly_b200.arbplf as A
    from tests.test_host_cpu import GOOD
    doc = copy.deepcopy(GOOD)
    doc.pop("site_reduction")
    A.arbplf_ll(json.dumps(doc))            # the unmodified document is fine
    doc.update(BAD_REDUCTIONS[i])
    with pytest.raises(RuntimeError):
        A.arbplf_ll(json.dumps(doc))


def test_site_weights_equal_duplicated_sites():
    # test_scripts/test_site_weights.py:60-97
    base = H.random_problem(41, ntips=6, n=4, S=3, ncat=2)
    md = base["model_and_data"]
    dup = copy.deepcopy(md)
    dup["character_data"] = [md["character_data"][i] for i in (0, 0, 1, 2, 2, 2)]
    for prog, extra in (("ll", {}), ("deriv", {}), ("marginal", {}), ("dwell", {"state_reduction": {"aggregation": "sum"}}),
                        ("trans", {"trans_reduction": {"aggregation": "sum"}})):
        a = _run(prog, dict({"model_and_data": md, "site_reduction": {"aggregation": [2, 1, 3]}}, **extra))
        b = _run(prog, dict({"model_and_data": dup, "site_reduction": {"aggregation": "sum"}}, **extra))
        _check(a, b, "weights vs duplicates " + prog, rtol=1e-13)


def test_rate_matrix_scaling_invariance():
    # test_scripts/test_rate_divisor.py:96-262
    base = H.random_problem(42, ntips=5, n=4, S=4, ncat=1, divisor=3.0)
    md = base["model_and_data"]
    scaled = copy.deepcopy(md)
    scaled["rate_matrix"] = [[8 * x for x in row] for row in md["rate_matrix"]]
    scaled["rate_divisor"] = 24.0
    diag = copy.deepcopy(md)
    for i in range(4):
        diag["rate_matrix"][i][i] = 123.0          # the diagonal is ignored
    for prog in ("ll", "deriv", "marginal"):
        a = _run(prog, {"model_and_data": md})
        _check(_run(prog, {"model_and_data": scaled}), a, "scaled " + prog, rtol=1e-13)
        _check(_run(prog, {"model_and_data": diag}), a, "diag " + prog, rtol=1e-13)


def test_edge_permutation_equivariance():
    # test_scripts/test_ll_deriv.py / test_path.py:133-195
    base = H.random_problem(43, ntips=6, n=4, S=3, ncat=2)
    md = base["model_and_data"]
    E = len(md["edges"])
    perm = list(np.random.default_rng(5).permutation(E))
    pm = copy.deepcopy(md)
    pm["edges"] = [md["edges"][i] for i in perm]
    pm["edge_rate_coefficients"] = [md["edge_rate_coefficients"][i] for i in perm]
    a = _run("deriv", {"model_and_data": md, "site_reduction": {"aggregation": "sum"}})
    b = _run("deriv", {"model_and_data": pm, "site_reduction": {"aggregation": "sum"}})
    va = {r[0]: r[1] for r in a["data"]}
    vb = {perm[r[0]]: r[1] for r in b["data"]}
    for e in range(E):
        assert abs(va[e] - vb[e]) <= 1e-12 * abs(va[e]) + 1e-14


def test_em_update_against_oracle_with_mixture_and_weights():
    """arbplf-em-update beyond the reference's three goldens: a gamma + invariant mixture (one category of rate
    0, skipped per site when its likelihood is exactly 0), weighted site aggregation, a selection with
    duplicates."""
    from oracle import arbplf_oracle as O
    doc = copy.deepcopy(H.golden_in("beast_gtrgi"))
    S = len(doc["model_and_data"]["character_data"]) if "character_data" in doc["model_and_data"] else \
        len(doc["model_and_data"]["probability_array"])
    rng = np.random.default_rng(8)
    sel = [int(x) for x in rng.integers(0, S, 12)]
    for red in ({"aggregation": "sum"}, {"aggregation": "avg", "selection": sel},
                {"selection": sel, "aggregation": [float(x) for x in rng.random(len(sel)) + 0.1]}):
        doc["site_reduction"] = red
        want = O.run("em_update", doc, mode="mp")
        got = _run("em_update", doc)
        _check(got, want, "em_update %s" % json.dumps(red)[:40])


# ---------------------------------------------------------------------------
# second order programs (arbplfhess.c): hess / inv-hess / newton-delta / newton-update against the 320-bit oracle
# ---------------------------------------------------------------------------

SECOND_ORDER = [
    dict(seed=21, ntips=5, n=4, S=6, ncat=1),
    dict(seed=22, ntips=7, n=4, S=9, ncat=3, mixture="gamma", missing=0.2),
    dict(seed=23, ntips=6, n=4, S=7, ncat=2, root="equilibrium_distribution", root_degree=3, internal_data=True),
    dict(seed=24, ntips=5, n=3, S=8, ncat=2, root="uniform_distribution"),
    dict(seed=25, ntips=9, n=4, S=12, ncat=2, mixture="median_inv", max_degree=3),
    dict(seed=26, ntips=4, n=20, S=5, ncat=1),
]


def _second_order_expected(name, program, doc):
    return H.expected_json("so_%s_%s" % (program, name), program, doc)


@pytest.mark.parametrize("kw", SECOND_ORDER, ids=["random%d" % k["seed"] for k in SECOND_ORDER])
def test_second_order_programs_against_oracle(kw):
    prob = H.random_problem(**kw)
    S = len(prob["model_and_data"].get("probability_array", prob["model_and_data"].get("character_data")))
    w = [0.5 + (7 * i % 5) for i in range(S)]
    doc = {"model_and_data": prob["model_and_data"],
           "site_reduction": {"selection": list(range(S)), "aggregation": w}}
    name = "random%d" % kw["seed"]
    for program, rtol in (("hess", 1e-11), ("newton_delta", 1e-9), ("newton_update", 1e-9), ("inv_hess", 1e-9)):
        got = _run(program, doc)
        want = _second_order_expected(name, program, doc)
        # inverse / solve amplify the 1e-13 of the Hessian entries by its condition number: looser, stated bound
        _check(got, want, "%s %s" % (program, name), rtol=rtol)


def test_hess_closed_form_truncated_path():
    """test_scripts/test_path_exponential_absorbing.py:124-135: two-state path with an absorbing state, last node
    observed in state 1: every entry of the Hessian is the second derivative of log(1 - exp(-T)), T = total rate."""
    import math
    rates = [1.0, 2.0, 3.0]
    N = len(rates) + 1
    pa = [[[1, 0]] + [[1, 1]] * (N - 2) + [[0, 1]]]
    doc = {"model_and_data": {"edges": [[i, i + 1] for i in range(N - 1)], "edge_rate_coefficients": rates,
                              "rate_matrix": [[0, 1], [0, 0]], "probability_array": pa},
           "site_reduction": {"aggregation": "sum"}}
    got = _run("hess", doc)
    T = sum(rates)
    want = -math.exp(T) / math.expm1(T) ** 2
    assert len(got["data"]) == 9
    for r in got["data"]:
        assert abs(r[-1] - want) <= 1e-11 * abs(want), (r, want)


def test_newton_refine_reaches_a_stationary_point():
    prob = H.random_problem(seed=22, ntips=7, n=4, S=9, ncat=3, mixture="gamma", missing=0.2)
    S = len(prob["model_and_data"].get("probability_array", prob["model_and_data"].get("character_data")))
    doc = {"model_and_data": prob["model_and_data"], "site_reduction": {"aggregation": "sum"}}
    out = _run("newton_refine", doc)
    assert out["columns"] == ["edge", "value"]
    md = json.loads(json.dumps(prob["model_and_data"]))
    md["edge_rate_coefficients"] = [max(r[-1], 1e-300) for r in out["data"]]
    g = _run("deriv", {"model_and_data": md, "site_reduction": {"aggregation": "sum"}})
    ll0 = _run("ll", {"model_and_data": prob["model_and_data"], "site_reduction": {"aggregation": "sum"}})["data"][0][-1]
    ll1 = _run("ll", {"model_and_data": md, "site_reduction": {"aggregation": "sum"}})["data"][0][-1]
    assert ll1 >= ll0 - 1e-9
    # interior coordinates are stationary; a rate pushed to the boundary keeps a non-positive derivative
    for (e, x), (_, gv) in zip(out["data"], [(r[0], r[-1]) for r in g["data"]]):
        assert abs(gv) <= 1e-6 * S or (x < 1e-6 and gv <= 1e-6 * S), (e, x, gv)


# ---------------------------------------------------------------------------
# certified mode: the enclosure of arbplf-ll contains the reference's certified value
# ---------------------------------------------------------------------------

LL_GOLDENS = [c for c in CASES if c["program"] == "ll"]


@pytest.mark.parametrize("case", LL_GOLDENS, ids=[c["name"] for c in LL_GOLDENS])
def test_certified_enclosure_contains_the_golden_value(case):
    """north_star: 'In certified mode, the GPU enclosure must contain the Arb midpoint.'  Every arbplf-ll golden of
    the reference (correctly rounded doubles of its certified balls) lies inside [lower, upper], and the enclosure is
    tight: its width stays below 1e-10 of the value."""
    got = _run("ll_certified", H.golden_in(case["name"]))
    want = H.golden_out(case["name"])
    lo, hi = got["lower"], got["upper"]
    assert lo["columns"] == want["columns"] and hi["columns"] == want["columns"]
    assert len(lo["data"]) == len(want["data"]) == len(hi["data"])
    for a, b, w in zip(lo["data"], hi["data"], want["data"]):
        assert a[:-1] == w[:-1] == b[:-1]
        assert a[-1] <= w[-1] <= b[-1], (case["name"], a, w, b)
        assert b[-1] - a[-1] <= 1e-10 * max(abs(w[-1]), 1e-3), (case["name"], a, w, b)


def test_certified_enclosure_contains_the_oracle_on_random_problems():
    from oracle import arbplf_oracle as O
    for kw in (dict(seed=2, ntips=12, n=4, S=33, ncat=4, mixture="gamma", missing=0.2),
               dict(seed=4, ntips=10, n=4, S=8, ncat=2, root="uniform_distribution", internal_data=True),
               dict(seed=7, ntips=6, n=5, S=7, ncat=1, soft=True, internal_data=True),
               dict(seed=10, ntips=40, n=4, S=24, ncat=4, mixture="gamma", edge_scale=0.05),
               dict(seed=14, ntips=6, n=20, S=19, ncat=2, missing=0.2, root="equilibrium_distribution")):
        prob = H.random_problem(**kw)
        doc = {"model_and_data": prob["model_and_data"]}
        got = _run("ll_certified", doc)
        want = H.expected_json("ll_random%d" % kw["seed"], "ll", doc)
        for a, b, w in zip(got["lower"]["data"], got["upper"]["data"], want["data"]):
            assert a[-1] <= w[-1] <= b[-1], (kw["seed"], a, w, b)
            assert b[-1] - a[-1] <= 1e-10 * max(abs(w[-1]), 1e-3), (kw["seed"], a, w, b)
        # aggregated with weights of both signs
        S = len(want["data"])
        wts = [((-1) ** i) * (0.5 + i % 3) for i in range(S)]
        docw = dict(doc, site_reduction={"selection": list(range(S)), "aggregation": wts})
        g2 = _run("ll_certified", docw)
        tot = sum(wt * r[-1] for wt, r in zip(wts, want["data"]))
        assert g2["lower"]["data"][0][-1] <= tot <= g2["upper"]["data"][0][-1]


def test_large_tables_are_written_by_several_threads_in_the_same_bytes(monkeypatch):
    """A table with the site axis kept and >= 65536 rows is written by one run of rows per host thread
    (host/reduce.c:table_to_json): the text is the one a single thread writes, for every shape of table and for a
    selection on the site axis; a large document is also read by the threaded reader (host/json.c)."""
    import phyly_b200.arbplf as A
    import bench
    doc, N = bench.model_document(12)
    S = 70001
    codes = np.random.default_rng(8).integers(0, 5, (S, N)).astype(np.uint8)
    leaves_only = np.ones(N, dtype=bool)
    for a, b in doc["model_and_data"]["edges"]:
        leaves_only[a] = False
    codes[:, ~leaves_only] = 4
    text = bench.json_document_bytes(doc, codes).decode()
    assert len(text) > (1 << 20)
    tail = ', "site_reduction": {"aggregation": "sum"}}'
    assert text.endswith(tail)
    body = text[:-len(tail)]
    sel = list(range(S - 1, -1, -1))          # every site, backwards: the order of the selection is the order of the rows
    cases = [("ll", body + "}"),
             ("ll", body + ', "site_reduction": {"selection": %s}}' % json.dumps(sel)),
             ("deriv", body + "}"),
             ("deriv", body + ', "edge_reduction": {"aggregation": "sum"}}'),
             ("marginal", body + ', "site_reduction": {"selection": %s}}' % json.dumps(list(range(0, S, 3)))),
             ("dwell", body + ', "edge_reduction": {"aggregation": "avg"}}'),
             ("trans", body + ', "trans_reduction": {"aggregation": "sum"}}')]
    for prog, t in cases:
        fn = getattr(A, "arbplf_" + prog)
        monkeypatch.setenv("ARBPLF_HOST_THREADS", "1")
        one = fn(t)
        monkeypatch.setenv("ARBPLF_HOST_THREADS", "7")
        many = fn(t)
        monkeypatch.delenv("ARBPLF_HOST_THREADS")
        auto = fn(t)
        assert one == many == auto, prog
    got = json.loads(A.arbplf_ll(cases[1][1]))
    assert [r[0] for r in got["data"]] == sel
    fwd = json.loads(A.arbplf_ll(cases[0][1]))
    assert [r[1] for r in got["data"]] == [r[1] for r in fwd["data"]][::-1]
    # against the oracle on a few of the sites
    from oracle import arbplf_oracle as O
    idx = [0, 1, 35000, S - 1]
    md = dict(doc["model_and_data"])
    md["character_data"] = codes[idx].tolist()
    want = O.run_ll({"model_and_data": md}, mode="fp64")
    for k, i in enumerate(idx):
        assert abs(fwd["data"][i][1] - want["data"][k][1]) <= 1e-11 * abs(want["data"][k][1])


def test_repeated_calls_keep_the_alignment_on_the_device(monkeypatch, capfd):
    """An optimiser's loop: the same large document again and again with other edge rates.  From the second call on the
    reader recognises the character_data text (host/json.h) and the driver neither re-reads nor re-uploads it
    (drivers.c:ctx_load); every answer is the text a cold process gives (ARBPLF_NO_DATA_CACHE=1)."""
    import phyly_b200.arbplf as A
    import bench
    doc, N = bench.model_document(16)
    S = 40000
    rng = np.random.default_rng(12)
    codes = rng.integers(0, 5, (S, N)).astype(np.uint8)
    for a, b in doc["model_and_data"]["edges"]:
        codes[:, a] = 4
    text = bench.json_document_bytes(doc, codes).decode()
    assert len(text) > (1 << 20)
    rates = doc["model_and_data"]["edge_rate_coefficients"]

    def variant(scale):
        return text.replace(json.dumps(rates), json.dumps([r * scale for r in rates]), 1)

    docs = [text, variant(1.25), variant(0.5), text]
    monkeypatch.setenv("ARBPLF_NO_DATA_CACHE", "1")
    cold = [A.arbplf_deriv(t) for t in docs]
    cold_ll = A.arbplf_ll(docs[1].replace(', "site_reduction": {"aggregation": "sum"}', ""))
    monkeypatch.delenv("ARBPLF_NO_DATA_CACHE")
    monkeypatch.setenv("ARBPLF_JSON_TRACE", "1")
    A.arbplf_deriv(docs[0])
    capfd.readouterr()
    for t, want in zip(docs, cold):
        got = A.arbplf_deriv(t)
        err = capfd.readouterr().err
        assert "taken from the cache" in err and "no upload" in err
        assert got == want
    # another program on the same alignment, per-site output
    got_ll = A.arbplf_ll(docs[1].replace(', "site_reduction": {"aggregation": "sum"}', ""))
    assert "no upload" in capfd.readouterr().err
    assert got_ll == cold_ll
    # a different alignment in between, then the first one again: both are uploaded
    codes2 = codes.copy()
    codes2[:, -1] = (codes2[:, -1] + 1) % 4
    text2 = bench.json_document_bytes(doc, codes2).decode()
    monkeypatch.setenv("ARBPLF_NO_DATA_CACHE", "1")
    want2 = A.arbplf_deriv(text2)
    monkeypatch.delenv("ARBPLF_NO_DATA_CACHE")
    capfd.readouterr()
    assert A.arbplf_deriv(text2) == want2
    assert "no upload" not in capfd.readouterr().err
    assert A.arbplf_deriv(text) == cold[0]
    assert "no upload" not in capfd.readouterr().err
    assert A.arbplf_deriv(text) == cold[0]
    assert "no upload" in capfd.readouterr().err
    # a failed call in between forgets what the engine holds
    with pytest.raises(RuntimeError):
        A.arbplf_deriv(text.replace('"site_reduction": {"aggregation": "sum"}', '"site_reduction": {"aggregation": "nonsense"}'))
    capfd.readouterr()
    assert A.arbplf_deriv(text) == cold[0]
