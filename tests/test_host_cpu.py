"""
CPU-side tests of the C host layer and of the library boundary (no GPU work):
  * the shared library loads and exports every symbol declared in include/*.h;
  * the integer front end (CSR tree, BFS order, edge map) is bit exact against
    the oracle;
  * site-independent model quantities (mixture rates, equilibrium, scaled rate
    matrix in double-double) match the 320-bit oracle;
  * malformed models are rejected (test_scripts/test_bad_model_args.py of the
    reference) and nothing silently falls back to a CPU computation.
"""
import copy
import ctypes
import json
import os
import re

import numpy as np
import pytest

from oracle import arbplf_oracle as O
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from phyly_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol():
    lib = _lib()
    names = []
    for hdr in ("plf.h", "arbplf.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b((?:plf|arbplf)_[a-z0-9_]+)\s*\(", text)
    names = sorted(set(names))
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in include/ but not exported" % n


def _summary(doc):
    import phyly_b200.arbplf as A
    return json.loads(A.arbplf_model_summary(json.dumps({"model_and_data": doc["model_and_data"]})))


CASES = [c["name"] for c in H.manifest()]


@pytest.mark.parametrize("name", CASES)
def test_integer_front_end_is_bit_exact(name):
    doc = H.golden_in(name)
    s = _summary(doc)
    t = O.build_tree(doc["model_and_data"]["edges"])
    assert s["indptr"] == t.indptr
    assert s["indices"] == t.indices
    assert s["preorder"] == t.preorder
    assert s["order"] == t.order


@pytest.mark.parametrize("seed", range(8))
def test_integer_front_end_random_trees(seed):
    doc = H.random_problem(100 + seed, ntips=5 + 7 * seed, n=4, S=1, max_degree=2 + seed % 3)
    s = _summary(doc)
    t = O.build_tree(doc["model_and_data"]["edges"])
    assert (s["indptr"], s["indices"], s["preorder"], s["order"]) == (t.indptr, t.indices, t.preorder, t.order)


@pytest.mark.parametrize("name", ["beast_gtrg", "beast_hky85g", "beast_hky85i", "beast_gtrgi", "beast_gtri", "bpp_ll",
                                  "gell_test", "gell_driver", "beast_k80", "fels_ll", "mj_jumps"])
def test_model_quantities_match_oracle(name):
    doc = H.golden_in(name)
    s = _summary(doc)
    m = O.parse_model(doc["model_and_data"])
    params, cs = H.model_params(m)
    np.testing.assert_allclose(s["cat_rates"], params["cat_rates"], rtol=4e-16, atol=0)
    np.testing.assert_allclose(s["cat_prior"], params["cat_prior"], rtol=4e-16, atol=0)
    n = m.n
    q = np.array(s["q_hi"]).reshape(n, n)
    np.testing.assert_allclose(q, params["q_hi"], rtol=4e-16, atol=0)
    # hi + lo together carry ~31 digits
    from mpmath import mpf
    for i in range(n):
        for j in range(n):
            got = mpf(s["q_hi"][i * n + j]) + mpf(s["q_lo"][i * n + j])
            want = cs.Q[i, j]
            assert abs(got - want) <= abs(want) * mpf(10) ** -18, (i, j)
    np.testing.assert_allclose(s["edge_rates_csr"], params["edge_rates"], rtol=0, atol=0)
    if params["root_vec"] is not None:
        np.testing.assert_allclose(s["root_vec"], params["root_vec"], rtol=4e-16)


def test_gamma_rates_known_answers():
    # test_scripts/test_gamma_discretization.py:101-107 (Yang 1994, shape 0.5, 4 categories)
    doc = H.golden_in("fels_ll")
    doc["model_and_data"]["gamma_rate_mixture"] = {"gamma_shape": 0.5, "gamma_categories": 4}
    s = _summary(doc)
    want = [0.0333877533835995, 0.251915917593438, 0.820268481973649, 2.89442784704931]
    np.testing.assert_allclose(s["cat_rates"], want, rtol=2e-15)
    # tiny shape: effectively one category with rate 4 (test_gamma_discretization.py:146-184)
    doc["model_and_data"]["gamma_rate_mixture"] = {"gamma_shape": 1e-6, "gamma_categories": 4}
    s = _summary(doc)
    np.testing.assert_allclose(s["cat_rates"], [0, 0, 0, 4], atol=1e-12)
    # invariant sites: rates divided by (1-p), extra category of rate 0
    doc["model_and_data"]["gamma_rate_mixture"] = {"gamma_shape": 0.5, "gamma_categories": 4, "invariable_prior": 0.3}
    s = _summary(doc)
    np.testing.assert_allclose(s["cat_rates"], [w / 0.7 for w in want] + [0.0], rtol=2e-15)
    np.testing.assert_allclose(s["cat_prior"], [0.7 / 4] * 4 + [0.3], rtol=2e-16)


@pytest.mark.parametrize("shape", [0.05, 0.137064, 0.19249, 0.5, 1.0, 3.7, 25.0])
@pytest.mark.parametrize("ncat", [2, 4, 8])
def test_gamma_rates_vs_oracle(shape, ncat):
    from mpmath import mp
    mp.prec = O.MP_PREC_BITS
    doc = H.golden_in("fels_ll")
    for key, fn in (("gamma_rate_mixture", O.gamma_rates_mean), ("normalized_median_gamma_rate_mixture", O.gamma_rates_median)):
        d = copy.deepcopy(doc)
        d["model_and_data"][key] = {"gamma_shape": shape, "gamma_categories": ncat}
        got = _summary(d)["cat_rates"]
        want = [float(x) for x in fn(ncat, shape)]
        np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-300)


GOOD = {
    "model_and_data": {
        "edges": [[0, 1], [1, 2], [1, 3]],
        "edge_rate_coefficients": [2.0, 4.2, 0.5],
        "rate_matrix": [[0, 4.2, 3.0], [1.0, 0, 5.0], [6.0, 0.5, 0]],
        "probability_array": [
            [[0.6, 0.2, 0.2], [1, 1, 1], [1, 0, 0], [0, 0, 1]],
            [[0.6, 0.2, 0.2], [1, 1, 1], [0, 0, 1], [0, 0, 1]]]},
    "site_reduction": {"aggregation": "sum"},
}


def _bad_models():
    # test_scripts/test_bad_model_args.py:31-260
    def mod(f):
        x = copy.deepcopy(GOOD)
        f(x["model_and_data"])
        return x
    out = []
    for bad in ({"hello": "world"}, "hello", 0, [-1, 3], [0, 1, 2], [0], [1.0, 3]):
        out.append(mod(lambda m, b=bad: m["edges"].__setitem__(2, b)))
    for edges in ([[0, 1], [1, 5], [1, 6]], [[0, 1], [1, 2], [2, 2]], [[0, 1], [1, 2], [3, 3]],
                  [[0, 1], [1, 2], [1, 3], [2, 3]], [[0, 1], [1, 2], [2, 0]], [[0, 1], [2, 3], [3, 2]],
                  [[0, 1], [2, 1], [3, 1]], [[0, 1], [2, 3], [2, 4], [4, 3]]):
        out.append(mod(lambda m, e=edges: m.__setitem__("edges", e)))
    for erc in ("hello", [2.0, 4.2, 0.5, 1.0], [2.0, 4.2], [2.0, "x", 0.5], [2.0, "4.2", 0.5], [2.0, -4.2, 0.5]):
        out.append(mod(lambda m, e=erc: m.__setitem__("edge_rate_coefficients", e)))
    out.append(mod(lambda m: m.__setitem__("rate_matrix", "hello")))
    out.append(mod(lambda m: m["rate_matrix"].__setitem__(1, "hello")))
    out.append(mod(lambda m: m["rate_matrix"].__setitem__(1, [1.0, 0, 5.0, 1.0])))
    out.append(mod(lambda m: m["rate_matrix"].__setitem__(1, [1.0, 0])))
    out.append(mod(lambda m: m["rate_matrix"][1].__setitem__(2, "hello")))
    out.append(mod(lambda m: m["rate_matrix"][1].__setitem__(2, "5.0")))
    out.append(mod(lambda m: m["rate_matrix"][1].__setitem__(2, -5.0)))
    out.append(mod(lambda m: m.__setitem__("probability_array", "hello")))
    out.append(mod(lambda m: m["probability_array"].__setitem__(0, "hello")))
    out.append(mod(lambda m: m["probability_array"].__setitem__(0, {"a": 1})))
    out.append(mod(lambda m: m["probability_array"][0].__setitem__(1, "hello")))
    out.append(mod(lambda m: m["probability_array"][0].__setitem__(1, {"a": 1})))
    out.append(mod(lambda m: m["probability_array"][0][1].__setitem__(1, "hello")))
    out.append(mod(lambda m: m["probability_array"][0][1].__setitem__(1, -0.5)))
    out.append(mod(lambda m: m["probability_array"][0].append([1, 1, 1])))
    out.append(mod(lambda m: m["probability_array"][0].pop()))
    out.append(mod(lambda m: m["probability_array"][0][1].append(1)))
    out.append(mod(lambda m: m["probability_array"][0][1].pop()))
    # schema level
    out.append(mod(lambda m: m.__setitem__("unknown_key", 1)))
    out.append(mod(lambda m: m.pop("edges")))
    out.append(mod(lambda m: m.__setitem__("rate_divisor", -1)))
    out.append(mod(lambda m: m.__setitem__("rate_divisor", "bogus")))
    out.append(mod(lambda m: m.__setitem__("root_prior", "bogus")))
    out.append(mod(lambda m: m.__setitem__("root_prior", [0.5, 0.5])))
    out.append(mod(lambda m: (m.__setitem__("rate_mixture", {"rates": [1, 2], "prior": [0.5, 0.5]}),
                              m.__setitem__("gamma_rate_mixture", {"gamma_shape": 1, "gamma_categories": 2}))))
    out.append(mod(lambda m: m.__setitem__("gamma_rate_mixture", {"gamma_shape": 1, "gamma_categories": 2.0})))
    out.append(mod(lambda m: (m.__setitem__("character_data", [[0, 0, 0, 0]]), m.__setitem__("character_definitions", [[1, 1, 1]]))))
    return out


@pytest.mark.parametrize("i", range(len(_bad_models())))
def test_bad_models_are_rejected(i):
    import phyly_b200.arbplf as A
    doc = _bad_models()[i]
    with pytest.raises(O.OracleError):
        O.run("ll", doc, mode="fp64")
    with pytest.raises(RuntimeError):
        A.arbplf_model_summary(json.dumps({"model_and_data": doc["model_and_data"]}))
    with pytest.raises(RuntimeError):
        A.arbplf_ll(json.dumps(doc))


def test_malformed_json_is_rejected():
    import phyly_b200.arbplf as A
    for s in ("", "{", "[1, 2", '{"model_and_data": }', "3", '{"a": 1} trailing'):
        with pytest.raises(RuntimeError):
            A.arbplf_ll(s)


def test_second_order_programs_require_site_aggregation():
    # arbplfhess.c:1162-1207: "site_reduction" is required and must aggregate; rejected before any device work
    import phyly_b200.arbplf as A
    for f in (A.arbplf_hess, A.arbplf_inv_hess, A.arbplf_newton_delta, A.arbplf_newton_update,
              A.arbplf_newton_refine):
        with pytest.raises(RuntimeError):
            f(json.dumps({"model_and_data": GOOD["model_and_data"]}))
        with pytest.raises(RuntimeError):
            f(json.dumps(dict(GOOD, site_reduction={"selection": [0]})))
        with pytest.raises(RuntimeError):
            f(json.dumps(dict(GOOD, site_reduction={"aggregation": "sum"}, edge_reduction={"aggregation": "sum"})))


def test_em_update_requires_site_aggregation():
    # arbplfem.c:566-571: rejected while parsing, before any device work
    import phyly_b200.arbplf as A
    no_agg = {k: v for k, v in GOOD.items() if k != "site_reduction"}
    with pytest.raises(RuntimeError):
        A.arbplf_em_update(json.dumps(no_agg))
    with pytest.raises(RuntimeError):       # edge reduction is not part of the schema (strict unpack, arbplfem.c:517-523)
        A.arbplf_em_update(json.dumps(dict(GOOD, edge_reduction={"aggregation": "sum"})))


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import phyly_b200.arbplf as A
    from phyly_b200.engine import Engine, EngineError
    with pytest.raises(RuntimeError):
        A.arbplf_ll(json.dumps(GOOD))
    with pytest.raises(EngineError):
        Engine(0)


def test_integer_fast_path_of_the_json_reader():
    """Arrays of short non-negative integers (character codes) take a tight loop in host/json.c; everything it does not
    recognise -- leading zeros, signs, fractions, long integers, blanks -- must still get the general route's treatment."""
    import phyly_b200.arbplf as A
    md = {"edges": [[0, 1], [0, 2]], "edge_rate_coefficients": [1, 2], "rate_matrix": [[0, 1], [1, 0]],
          "character_definitions": [[1, 0], [0, 1], [1, 1]], "character_data": [[2, 0, 1], [2, 1, 1]]}
    good = json.dumps({"model_and_data": md})
    s = json.loads(A.arbplf_model_summary(good))
    assert s["edge_rates_csr"] == [1.0, 2.0]
    spaced = good.replace("[2, 0, 1]", "[ 2 ,0\n, 1 ]")
    assert json.loads(A.arbplf_model_summary(spaced)) == s
    for bad in (good.replace("[2, 0, 1]", "[2, 00, 1]"), good.replace("[2, 0, 1]", "[2, 01, 1]"),
                good.replace("[2, 0, 1]", "[2, -0, 1]").replace("-0", "-1"), good.replace("[2, 0, 1]", "[2, 0.5, 1]"),
                good.replace("[2, 0, 1]", "[2, 3, 1]"), good.replace("[2, 0, 1]", "[2, 12345678901234567890, 1]"),
                good.replace("[2, 0, 1]", "[2, 0, 1,]"), good.replace("[2, 0, 1]", "[2 0 1]")):
        with pytest.raises(RuntimeError):
            A.arbplf_model_summary(bad)
    # integers where reals are expected stay valid, long integers keep their value
    ok = good.replace('"edge_rate_coefficients": [1, 2]', '"edge_rate_coefficients": [1234567890123, 2]')
    assert json.loads(A.arbplf_model_summary(ok))["edge_rates_csr"][0] == 1234567890123.0


def _big_document(S, taxa=64, seed=0):
    """A cfg2-shaped document whose character_data is large enough for the threaded reader of host/json.c
    (>= 1 MB of text, >= 256 rows); returns (text, codes)."""
    import bench
    doc, N = bench.model_document(taxa)
    codes = np.random.default_rng(seed).integers(0, 5, (S, N)).astype(np.uint8)
    return _summary_document(bench, doc, codes), codes


def _summary_document(bench, doc, codes):
    text = bench.json_document_bytes(doc, codes).decode()
    extra = ', "site_reduction": {"aggregation": "sum"}'
    assert text.endswith(extra + "}")
    return text[:-len(extra) - 1] + "}"


def _code_sums(codes):
    flat = codes.astype(np.uint64).ravel()
    pos = np.arange(1, flat.size + 1, dtype=np.uint64)
    return str(int(flat.sum())), str(int((flat * pos).sum(dtype=np.uint64)))


def test_threaded_matrix_reader_matches_the_general_route(monkeypatch, capfd):
    """host/json.c reads a large matrix of short integers with several threads into shared blocks, host/model.c turns it
    into code bytes with several threads: the codes are those of the document, for any thread count, with blanks
    anywhere JSON allows them."""
    import phyly_b200.arbplf as A
    text, codes = _big_document(6000)
    assert len(text) > (1 << 20)
    want = _code_sums(codes)
    monkeypatch.setenv("ARBPLF_JSON_TRACE", "1")
    monkeypatch.setenv("ARBPLF_HOST_THREADS", "4")
    monkeypatch.setenv("ARBPLF_NO_DATA_CACHE", "1")      # (the cache of the last character_data has its own test)
    capfd.readouterr()
    A.arbplf_model_summary(text)
    assert "json: matrix of 6000 rows read by 4 threads" in capfd.readouterr().err       # the route under test is taken
    monkeypatch.delenv("ARBPLF_JSON_TRACE")
    for threads in ("1", "3", "16", None):
        if threads is None:
            monkeypatch.delenv("ARBPLF_HOST_THREADS", raising=False)
        else:
            monkeypatch.setenv("ARBPLF_HOST_THREADS", threads)
        s = json.loads(A.arbplf_model_summary(text))
        assert s["site_count"] == 6000 and s["definition_count"] == 5
        assert (s["codes_sum"], s["codes_weighted_sum"]) == want, threads
    monkeypatch.delenv("ARBPLF_HOST_THREADS", raising=False)
    spaced = text.replace("],[", " ]\n,\t[ ").replace(",", " , ", 3000)
    s = json.loads(A.arbplf_model_summary(spaced))
    assert (s["codes_sum"], s["codes_weighted_sum"]) == want
    # the matrix need not come last: text behind it (with brackets of its own) is handed to some threads, which drop it
    key = '"character_data": '
    i = text.index(key)
    j = text.index("]]", i) + 2
    moved = '{"model_and_data": {' + text[i:j] + ", " + text[len('{"model_and_data": {'):i].rstrip(", ") + text[j:]
    assert json.loads(moved)["model_and_data"].keys() == json.loads(text)["model_and_data"].keys()
    monkeypatch.setenv("ARBPLF_JSON_TRACE", "1")
    capfd.readouterr()
    s = json.loads(A.arbplf_model_summary(moved))
    assert "json: matrix of 6000 rows read by" in capfd.readouterr().err
    monkeypatch.delenv("ARBPLF_JSON_TRACE")
    assert (s["codes_sum"], s["codes_weighted_sum"]) == want
    # a short document (general route) with the same rows gives the same bytes
    small, codes_small = _big_document(40)
    s = json.loads(A.arbplf_model_summary(small))
    assert (s["codes_sum"], s["codes_weighted_sum"]) == _code_sums(codes_small)
    # more than 256 definitions: 4-byte codes
    import bench
    doc, N = bench.model_document(64)
    K = 300
    md = dict(doc["model_and_data"])
    md["character_definitions"] = [[1, 1, 1, 1] if k % 7 == 0 else [float(k % 4 == j) for j in range(4)] for k in range(K)]
    wide = np.random.default_rng(2).integers(0, K, (3000, N))
    md["character_data"] = wide.tolist()
    t = json.dumps({"model_and_data": md})
    assert len(t) > (1 << 20)
    s = json.loads(A.arbplf_model_summary(t))
    assert s["definition_count"] == K and (s["codes_sum"], s["codes_weighted_sum"]) == _code_sums(wide)


def test_threaded_matrix_reader_steps_aside_for_anything_else():
    """A row that is not a flat array of short non-negative integers sends the whole matrix to the general route, which
    rejects it (or accepts it) exactly as it would a small document; the tree's 'edges' -- also a matrix of
    integers -- may be read by either route."""
    import phyly_b200.arbplf as A
    text, codes = _big_document(6000, seed=1)
    N = codes.shape[1]
    row = "[" + ",".join(str(int(x)) for x in codes[3000]) + "]"
    assert text.count(row) >= 1
    cells = row[1:-1].split(",")

    def with_row(new):
        head, tail = text.split(row, 1)
        return head + new + tail

    bad_rows = [
        "[" + ",".join(["01"] + cells[1:]) + "]",                 # leading zero
        "[" + ",".join(["-1"] + cells[1:]) + "]",                 # negative
        "[" + ",".join(["0.0"] + cells[1:]) + "]",                # a real
        "[" + ",".join(["7"] + cells[1:]) + "]",                  # not a row of the definitions
        "[" + ",".join(["12345678901234567890"] + cells[1:]) + "]",
        row[:-1] + ",]",                                          # stray comma
        "[" + ",".join(cells[:-1]) + "]",                         # one node short
        "[" + ",".join(["[0]"] + cells[1:]) + "]",                # nested array
        "[" + ",".join(['"0"'] + cells[1:]) + "]",                # a string
        "[" + " ".join(cells) + "]",                              # no commas
        "0",                                                      # not an array
    ]
    for new in bad_rows:
        with pytest.raises(RuntimeError):
            A.arbplf_model_summary(with_row(new))
    # the same faults at the very end and the very start of the matrix
    last = "[" + ",".join(str(int(x)) for x in codes[-1]) + "]"
    assert text.count(last + "]") >= 1
    with pytest.raises(RuntimeError):
        A.arbplf_model_summary(text.replace(last + "]", last[:-1] + ",9]]", 1))
    with pytest.raises(RuntimeError):
        A.arbplf_model_summary(text.replace('"character_data": [[', '"character_data": [[-0,', 1))
    # an unterminated matrix is a syntax error, not a crash
    cut = text[:text.index(row) + 40]
    with pytest.raises(RuntimeError):
        A.arbplf_model_summary(cut)
    # a long integer that is valid JSON but no short code: general route, then the range check
    ok_edges = json.loads(A.arbplf_model_summary(text))
    # large tree: 'edges' has more than 256 rows and sits in front of > 1 MB of text
    import bench
    doc, N2 = bench.model_document(300)
    c2 = np.random.default_rng(5).integers(0, 5, (1200, N2)).astype(np.uint8)
    big = _summary_document(bench, doc, c2)
    assert len(big) > (1 << 20) and len(doc["model_and_data"]["edges"]) > 256
    sb = json.loads(A.arbplf_model_summary(big))
    small = _summary_document(bench, doc, c2[:1])
    ss = json.loads(A.arbplf_model_summary(small))
    for k in ("indptr", "indices", "preorder", "order", "edge_rates_csr"):
        assert sb[k] == ss[k], k
    assert (sb["codes_sum"], sb["codes_weighted_sum"]) == _code_sums(c2)
    assert ok_edges["site_count"] == 6000


def test_character_data_cache(monkeypatch, capfd):
    """The reader keeps the codes of the last large character_data matrix under a hash of its text (host/json.h): the
    same bytes in the next document are not read again, any other bytes are; what the model parser makes of a cached
    matrix (node count, definition count) is checked like a freshly read one."""
    import phyly_b200.arbplf as A
    monkeypatch.delenv("ARBPLF_NO_DATA_CACHE", raising=False)
    monkeypatch.setenv("ARBPLF_JSON_TRACE", "1")
    text, codes = _big_document(5000, seed=11)
    want = _code_sums(codes)

    def summary(t):
        capfd.readouterr()
        out = json.loads(A.arbplf_model_summary(t))
        return out, capfd.readouterr().err

    s1, err1 = summary(text)
    assert "read by" in err1 and "taken from the cache" not in err1
    s2, err2 = summary(text)
    assert "character_data of 5000 rows taken from the cache" in err2 and "read by" not in err2
    assert s1 == s2 and (s2["codes_sum"], s2["codes_weighted_sum"]) == want
    # other edge rates, same alignment: still the cached codes
    doc = json.loads(text[:text.index('"character_data"')].rstrip(", ") + "}}")
    rates = doc["model_and_data"]["edge_rate_coefficients"]
    changed = text.replace(json.dumps(rates), json.dumps([r * 1.5 for r in rates]), 1)
    assert changed != text
    s3, err3 = summary(changed)
    assert "taken from the cache" in err3
    assert (s3["codes_sum"], s3["codes_weighted_sum"]) == want
    assert s3["edge_rates_csr"] != s1["edge_rates_csr"]
    # one code differs (same length): read afresh, and the earlier text is no longer the cached one
    i = text.index("]]", text.index('"character_data"')) - 1
    assert text[i] in "01234"
    other = text[:i] + ("0" if text[i] != "0" else "1") + text[i + 1:]
    codes2 = codes.copy()
    codes2[-1, -1] = int(other[i])
    s4, err4 = summary(other)
    assert "read by" in err4 and "taken from the cache" not in err4
    assert (s4["codes_sum"], s4["codes_weighted_sum"]) == _code_sums(codes2)
    s5, err5 = summary(text)
    assert "read by" in err5
    assert (s5["codes_sum"], s5["codes_weighted_sum"]) == want
    # a cached matrix against a tree with another node count / fewer definitions: the parser's own errors
    s6, err6 = summary(text)
    assert "taken from the cache" in err6
    fewer_defs = text.replace(json.dumps(doc["model_and_data"]["character_definitions"]),
                              json.dumps(doc["model_and_data"]["character_definitions"][:4]), 1)
    assert fewer_defs != text
    with pytest.raises(RuntimeError):
        A.arbplf_model_summary(fewer_defs)
    assert "less than the character count" in capfd.readouterr().err
    summary(text)
    import bench
    doc_small, N_small = bench.model_document(20)
    md = dict(doc_small["model_and_data"])
    head = json.dumps({"model_and_data": md})
    j = text.index('"character_data"')
    k = text.index("]]", j) + 2
    grafted = head[:head.index('"character_data"')] + text[j:k] + "}}"
    with pytest.raises(RuntimeError):
        A.arbplf_model_summary(grafted)
    assert "failed to match the number of nodes" in capfd.readouterr().err
    # switched off
    monkeypatch.setenv("ARBPLF_NO_DATA_CACHE", "1")
    s7, err7 = summary(text)
    assert "taken from the cache" not in err7 and (s7["codes_sum"], s7["codes_weighted_sum"]) == want


def test_pack4_layout():
    """phyly_b200.engine.pack4 builds PLF_CODES_PACKED4 rows (include/plf.h): node 2j in the low, node 2j + 1 in the high
    nibble of byte j, rows of (N + 1) // 2 bytes, for odd and even node counts; codes above 15 are refused."""
    from phyly_b200.engine import pack4, CODES_PACKED4
    assert CODES_PACKED4 == 0
    rng = np.random.default_rng(4)
    for N in (1, 2, 7, 8, 127):
        codes = rng.integers(0, 16, (33, N)).astype(np.uint8)
        p = pack4(codes)
        assert p.shape == (33, (N + 1) // 2) and p.dtype == np.uint8
        for nd in range(N):
            assert np.array_equal((p[:, nd >> 1] >> ((nd & 1) * 4)) & 15, codes[:, nd])
        if N % 2:
            assert np.all(p[:, -1] >> 4 == 0)
        out = np.empty_like(p)
        assert pack4(codes, out=out) is out and np.array_equal(out, p)
    with pytest.raises(ValueError):
        pack4(np.array([[16, 0]], dtype=np.uint8))
