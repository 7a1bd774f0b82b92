"""
GPU parity tests of the device seam (include/plf.h) against the oracle.

Every comparison is against the oracle in 320-bit mode (which reproduces the
reference's golden outputs bit for bit, tests/test_oracle_golden.py), at the
tolerance the north star states: 1e-11 relative.  Exact zeros / ones that the
reference produces through its constant-column shortcut must stay exact.
The oracle values are cached in tests/golden/seam/*.npz (see
tests/helpers.py:seam_reference; regenerate by deleting the cache).
"""
import numpy as np
import pytest

from oracle import arbplf_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

RTOL = 1e-11
# A derivative is a sum of signed terms (rows of Q sum to zero).  Where the
# reference's exact arithmetic cancels to (nearly) zero, fp64 can only be
# accurate relative to the magnitude of the terms: |err| <= RTOL*|x| + CANCEL*sum|terms|.
CANCEL = 2e-14
# ll = log(L) with L an fp64 number: near L = 1 (sites with almost no data) the
# result cannot be better than a few ulp of L in absolute terms.
LL_ATOL = 2e-15


def _engine():
    from phyly_b200.engine import Engine
    return Engine(0)


def _assert_close(got, want, what, rtol=RTOL, atol=0.0):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = np.abs(got - want)
    tol = rtol * np.abs(want) + atol
    bad = ~(err <= tol)
    if bad.any():
        i = tuple(np.argwhere(bad)[0])
        raise AssertionError("%s: %d mismatches, first at %s: got %r want %r (rel %.3e)" % (
            what, bad.sum(), i, got[i], want[i], err[i] / max(abs(want[i]), 1e-300)))


def _paths(m, C, K):
    from phyly_b200 import engine as E
    t = m.tree
    maxdeg = max(t.indptr[i + 1] - t.indptr[i] for i in range(t.node_count))
    p = [E.PATH_GENERIC]
    if m.n == 4 and C <= 16 and maxdeg <= 3 and K <= 256:
        p.append(E.PATH_FUSED4)          # more than 4 categories: windows of <= 4 (run_fused_windows)
    if 16 < m.n <= 64 and K <= 256:
        p.append(E.PATH_AUTO)           # ll and edge forms on the tensor-pipe kernels (generic = tile / scalar kernels)
    return p


GOLDEN_MODELS = ["fels_deriv", "beast_gtrg", "beast_hky85i", "beast_gtrgi", "jc_long_deriv", "bpp_deriv",
                 "mj_jumps", "beast_anc_marginal", "jc29_same_deriv", "jc30_diff_deriv", "fels_ll2"]

RANDOM = [
    dict(seed=1, ntips=8, n=4, S=9, ncat=1),
    dict(seed=2, ntips=12, n=4, S=33, ncat=4, mixture="gamma", missing=0.2),
    dict(seed=3, ntips=9, n=4, S=17, ncat=3, root="equilibrium_distribution", root_degree=3),
    dict(seed=4, ntips=10, n=4, S=8, ncat=2, root="uniform_distribution", internal_data=True),
    dict(seed=5, ntips=7, n=4, S=6, ncat=5, mixture="median_inv"),
    dict(seed=6, ntips=6, n=2, S=7, ncat=2, root=None, divisor=None),
    dict(seed=7, ntips=6, n=5, S=7, ncat=1, soft=True, internal_data=True),
    dict(seed=8, ntips=4, n=20, S=5, ncat=1),
    dict(seed=9, ntips=9, n=4, S=12, ncat=1, max_degree=4),
    dict(seed=10, ntips=40, n=4, S=24, ncat=4, mixture="gamma", edge_scale=0.05),
    dict(seed=11, ntips=6, n=3, S=300, ncat=2, missing=0.5),
    dict(seed=12, ntips=24, n=4, S=70, ncat=2, root="equilibrium_distribution", missing=0.3),
    # codon- and amino-acid-sized state spaces: the FP64 tensor-pipe kernels (dmma.cu) against the 320-bit oracle
    dict(seed=13, ntips=4, n=61, S=9, ncat=1, missing=0.15),
    dict(seed=14, ntips=6, n=20, S=19, ncat=2, missing=0.2, root="equilibrium_distribution"),
    # more than 4 rate categories on the fused path: Gamma8, and Gamma6 + invariant (7 categories, one of rate 0)
    dict(seed=15, ntips=11, n=4, S=40, ncat=8, mixture="gamma", missing=0.1),
    dict(seed=16, ntips=9, n=4, S=25, ncat=7, mixture="median_inv", root="equilibrium_distribution"),
]


def _problems():
    out = []
    for name in GOLDEN_MODELS:
        out.append((name, H.golden_in(name)))
    for kw in RANDOM:
        out.append(("random%d" % kw["seed"], H.random_problem(**kw)))
    return out


PROBLEMS = _problems()
IDS = [p[0] for p in PROBLEMS]


def _setup(name, prob):
    m = O.parse_model(prob["model_and_data"])
    ref = H.seam_reference(name, prob)
    defs, codes = H.dedupe_rows(m.dense_pmat())
    return m, ref, defs.shape[0]


@pytest.mark.parametrize("name,prob", PROBLEMS, ids=IDS)
def test_matrices(name, prob):
    m, ref, K = _setup(name, prob)
    eng = _engine()
    cs = H.fill_engine(eng, m)
    P = eng.transition_matrices()
    D = eng.derivative_matrices()
    Ld, Lt, Lt_hi, Lt_lo, Lg = H.frechet_directions(m, cs)
    F = eng.frechet_matrices(Lg)
    _assert_close(P, ref["P"], name + " P", rtol=1e-14, atol=1e-300)
    for c in range(P.shape[0]):
        for e in range(P.shape[1]):
            # entries of Q.P cancel: accurate relative to the scale of |Q||P|
            _assert_close(D[c, e], ref["Dm"][c, e], "%s D[%d,%d]" % (name, c, e), rtol=1e-13,
                          atol=1e-28 * ref["Dscale"][c, e])
            _assert_close(F[c, e], ref["F"][c, e], "%s F[%d,%d]" % (name, c, e), rtol=1e-13,
                          atol=1e-28 * np.abs(ref["F"][c, e]).max())
    eng.close()


@pytest.mark.parametrize("name,prob", PROBLEMS, ids=IDS)
def test_ll_and_deriv(name, prob):
    m, ref, K = _setup(name, prob)
    S = m.site_count
    ll_w, D_w = ref["ll"], ref["D"]
    E = D_w.shape[1]
    rng = np.random.default_rng(0)
    w = rng.integers(0, 4, S).astype(np.float64) + rng.random(S)
    for path in _paths(m, int(ref["C"]), K):
        eng = _engine()
        H.fill_engine(eng, m)
        eng.set_path(path)
        site_ll, tot = eng.ll()
        _assert_close(site_ll, ll_w, "%s ll path%d" % (name, path), atol=LL_ATOL)
        assert H.close(tot, float(np.sum(ll_w)), 1e-11, 1e-13), (name, path, tot, np.sum(ll_w))
        r = eng.deriv(per_site=True, per_site_ll=True)
        _assert_close(r["site_ll"], ll_w, "%s deriv-ll path%d" % (name, path), atol=LL_ATOL)
        _assert_close(r["site_deriv"], D_w, "%s deriv path%d" % (name, path), atol=CANCEL * ref["Dabs"] + 1e-300)
        assert np.all(r["site_deriv"][D_w == 0.0] == 0.0), (name, path)     # exact zeros stay exact
        # weighted sums through the reduction kernels
        eng.set_site_weights(w)
        r2 = eng.deriv(per_site=False)
        want = (w[:, None] * D_w).sum(axis=0)
        mag = (np.abs(w[:, None] * D_w)).sum(axis=0).max()
        _assert_close(r2["sum_deriv"], want, "%s sum deriv path%d" % (name, path), atol=1e-13 * mag)
        assert H.close(r2["sum_ll"], float((w * ll_w).sum()), 1e-11, 1e-13), (name, path)
        mask = np.zeros(E, dtype=np.uint8)
        mask[::2] = 1
        r3 = eng.deriv(edge_mask=mask, per_site=False)
        _assert_close(r3["sum_deriv"][mask == 1], want[mask == 1], "%s masked deriv path%d" % (name, path),
                      atol=1e-13 * mag)
        assert np.all(r3["sum_deriv"][mask == 0] == 0.0)
        eng.close()


@pytest.mark.parametrize("name,prob", PROBLEMS, ids=IDS)
def test_marginal(name, prob):
    m, ref, K = _setup(name, prob)
    want = ref["marg"]
    S = m.site_count
    rng = np.random.default_rng(3)
    w = rng.integers(0, 4, S).astype(np.float64) + rng.random(S)
    for path in _paths(m, int(ref["C"]), K):       # generic kernels, and the fused 4-state kernel in marginal mode
        eng = _engine()
        H.fill_engine(eng, m)
        eng.set_path(path)
        sm, tot = eng.marginal()
        _assert_close(sm, want, "%s marginal path%d" % (name, path), atol=1e-300)
        assert np.all(sm[want == 0.0] == 0.0), (name, path)      # structural zeros stay exact
        _assert_close(tot, want.sum(axis=0), "%s marginal sums path%d" % (name, path), atol=1e-13 * S)
        # site-summed only (the warp-reduction route of the fused kernel), with weights
        eng.set_site_weights(w)
        _, tot_w = eng.marginal(per_site=False)
        _assert_close(tot_w, (w[:, None, None] * want).sum(axis=0), "%s weighted marginal sums path%d" % (name, path),
                      atol=1e-13 * w.sum())
        eng.close()


@pytest.mark.parametrize("name,prob", PROBLEMS, ids=IDS)
def test_dwell_and_trans(name, prob):
    m, ref, K = _setup(name, prob)
    S = m.site_count
    Xd, Xt = ref["Xd"], ref["Xt"]
    from phyly_b200 import engine as E
    for path in _paths(m, int(ref["C"]), K):
        eng = _engine()
        cs = H.fill_engine(eng, m)
        Ld, Lt, Lt_hi, Lt_lo, Lg = H.frechet_directions(m, cs)
        eng.set_path(path)
        so, tot = eng.edge_expect(E.KIND_DWELL, Ld)
        _assert_close(so, Xd, "%s dwell path%d" % (name, path), atol=1e-300)
        _assert_close(tot, Xd.sum(axis=0), "%s dwell sums path%d" % (name, path), atol=1e-13 * S)
        so, tot = eng.edge_expect(E.KIND_TRANS, Lt_hi, Lt_lo)
        _assert_close(so, Xt, "%s trans path%d" % (name, path), atol=1e-300)
        so2, tot2 = eng.edge_expect(E.KIND_TRANS, Lt_hi, Lt_lo, per_site=False)
        _assert_close(tot2, Xt.sum(axis=0), "%s trans sums path%d" % (name, path), atol=1e-13 * S)
        eng.close()


def test_deep_tree_scaling():
    """A 600-taxon tree underflows fp64 without per-site rescaling."""
    prob = H.random_problem(21, ntips=600, n=4, S=40, ncat=2, missing=0.02, edge_scale=0.3)
    m = O.parse_model(prob["model_and_data"])
    ref = H.seam_reference("deep600", prob)
    sites = ref["sites"]
    assert ref["ll"][0] < -745.0          # below log(min double): plain fp64 would underflow
    from phyly_b200 import engine as E
    for path in (E.PATH_GENERIC, E.PATH_FUSED4):
        eng = _engine()
        H.fill_engine(eng, m)
        eng.set_path(path)
        r = eng.deriv(per_site=True, per_site_ll=True)
        _assert_close(r["site_ll"][sites], ref["ll"], "deep ll path%d" % path)
        _assert_close(r["site_deriv"][sites], ref["D"], "deep deriv path%d" % path, atol=CANCEL * ref["Dabs"] + 1e-300)
        eng.close()


def test_zero_likelihood_is_an_error():
    prob = H.golden_in("fels_ll")
    md = prob["model_and_data"]
    md["edge_rate_coefficients"] = [0.0] * 7     # P = I everywhere, data disagree -> likelihood 0
    m = O.parse_model(md)
    from phyly_b200.engine import EngineError
    eng = _engine()
    H.fill_engine(eng, m)
    with pytest.raises(EngineError):
        eng.ll()
    eng.close()


def _cm_applicable(m, C, K):
    """The constant-memory kernels take binary internal nodes, <= 128 of them, and C * internal edges <= 254."""
    t = m.tree
    deg = [t.indptr[i + 1] - t.indptr[i] for i in range(t.node_count)]
    internal = [i for i in range(t.node_count) if deg[i] > 0]
    n_int_edges = sum(1 for i in range(t.node_count) for j in range(t.indptr[i], t.indptr[i + 1]) if deg[t.indices[j]] > 0)
    return (m.n == 4 and C <= 4 and K <= 16 and all(deg[i] == 2 for i in internal) and len(internal) <= 128
            and C * n_int_edges <= 254)


@pytest.mark.parametrize("config", [0, 1, 2])
@pytest.mark.parametrize("name,prob", PROBLEMS, ids=IDS)
def test_constant_memory_kernels(name, prob, config, monkeypatch):
    """The production choice at bench size (matrices in __constant__ memory, program as kernel parameter, 512 or
    384 threads) is forced onto the small problems so that it meets the oracle too: per-site and summed
    derivatives, edge masks, and a Frechet (trans) query."""
    m, ref, K = _setup(name, prob)
    C = int(ref["C"])
    if not _cm_applicable(m, C, K):
        pytest.skip("not a binary 4-state problem")
    from phyly_b200 import engine as E
    monkeypatch.setenv("PLF_F4_CONFIG", str(config))
    monkeypatch.setenv("PLF_F4_CONFIG_LL", "0" if config == 0 else "2")     # the ll-only constant-memory kernels
    ll_w, D_w = ref["ll"], ref["D"]
    S = m.site_count
    eng = _engine()
    cs = H.fill_engine(eng, m)
    eng.set_path(E.PATH_FUSED4)
    site_ll, tot = eng.ll()
    _assert_close(site_ll, ll_w, "%s cm%d ll-only" % (name, config), atol=LL_ATOL)
    assert H.close(tot, float(np.sum(ll_w)), 1e-11, 1e-13)
    r = eng.deriv(per_site=True, per_site_ll=True)
    _assert_close(r["site_ll"], ll_w, "%s cm%d ll" % (name, config), atol=LL_ATOL)
    _assert_close(r["site_deriv"], D_w, "%s cm%d deriv" % (name, config), atol=CANCEL * ref["Dabs"] + 1e-300)
    assert np.all(r["site_deriv"][D_w == 0.0] == 0.0)
    rng = np.random.default_rng(5)
    w = rng.integers(0, 4, S).astype(np.float64) + rng.random(S)
    eng.set_site_weights(w)
    mask = np.zeros(D_w.shape[1], dtype=np.uint8)
    mask[1::2] = 1
    r2 = eng.deriv(edge_mask=mask, per_site=False)
    want = (w[:, None] * D_w).sum(axis=0)
    mag = (np.abs(w[:, None] * D_w)).sum(axis=0).max()
    _assert_close(r2["sum_deriv"][mask == 1], want[mask == 1], "%s cm%d masked sums" % (name, config), atol=1e-13 * mag)
    assert np.all(r2["sum_deriv"][mask == 0] == 0.0)
    assert H.close(r2["sum_ll"], float((w * ll_w).sum()), 1e-11, 1e-13)
    eng.set_site_weights(np.ones(S))
    Ld, Lt, Lt_hi, Lt_lo, Lg = H.frechet_directions(m, cs)
    xt, _ = eng.edge_expect(E.KIND_TRANS, Lt_hi, Lt_lo)
    _assert_close(xt, ref["Xt"], "%s cm%d trans" % (name, config), atol=1e-300)
    eng.close()


def test_out_of_range_character_code_is_an_error():
    """A code that is not a row of the definition table is caught on the device while the data is transposed
    (the JSON front end validates codes itself, parsemodel.c:600-613; the binary seam must not trust them)."""
    from phyly_b200.engine import EngineError
    prob = H.random_problem(31, ntips=6, n=4, S=50, ncat=2)
    m = O.parse_model(prob["model_and_data"])
    defs, codes = H.dedupe_rows(m.dense_pmat())
    for use_async in (False, True):
        eng = _engine()
        H.fill_engine(eng, m)
        bad = np.ascontiguousarray(codes.astype(np.uint8))
        bad[17, 3] = defs.shape[0]              # first invalid code
        if use_async:
            eng.set_data_async(defs, bad)
            with pytest.raises(EngineError):
                eng.ll()
        else:
            with pytest.raises(EngineError):
                eng.set_data(defs, bad)
        with pytest.raises(EngineError):        # no data is installed afterwards
            eng.ll()
        eng.set_data(defs, np.ascontiguousarray(codes.astype(np.uint8)))
        eng.ll()
        eng.close()


def test_site_pattern_compression_is_bit_exact():
    """plf_compress_patterns against the oracle's restatement of the reference-side generator
    (examples/BEAST.GTRG/mknuc.py:57-66): patterns in order of first occurrence, integer counts, site -> pattern map."""
    rng = np.random.default_rng(3)
    eng = _engine()
    for S, N, K, dtype in ((1, 5, 3, np.uint8), (1000, 7, 2, np.uint8), (20000, 11, 5, np.uint8), (5000, 9, 300, np.int32),
                           (4097, 127, 5, np.uint8)):
        codes = rng.integers(0, K, (S, N)).astype(dtype)
        if S > 2000:
            codes[rng.integers(0, S, S // 2)] = codes[rng.integers(0, S, S // 2)]      # plenty of duplicates
        pat, cnt, smap = eng.compress_patterns(codes)
        opat, ocnt, osmap = O.compress_patterns(codes)
        assert pat.shape == opat.shape and np.array_equal(pat, opat), (S, N)
        assert np.array_equal(cnt, ocnt) and np.array_equal(smap, osmap), (S, N)
        assert cnt.sum() == S and np.array_equal(pat[smap], codes)
    eng.close()


def test_compressed_alignment_gives_the_same_sums():
    """test_scripts/test_site_weights.py:60-97: an alignment with repeated columns and its compressed form with the
    multiplicities as site weights give the same aggregated outputs; the map scatters per-pattern values back."""
    prob = H.random_problem(seed=2, ntips=12, n=4, S=33, ncat=4, mixture="gamma", missing=0.2)
    m = O.parse_model(prob["model_and_data"])
    defs, codes = H.dedupe_rows(m.dense_pmat())
    rng = np.random.default_rng(5)
    rep = rng.integers(0, codes.shape[0], 500)
    big = np.ascontiguousarray(codes[rep]).astype(np.uint8)
    eng = _engine()
    H.fill_engine(eng, m)
    eng.set_data(defs, big)
    full = eng.deriv(per_site=True, per_site_ll=True)
    pat, cnt, smap = eng.compress_patterns(big)
    assert pat.shape[0] <= codes.shape[0]
    eng.set_data(defs, pat)
    eng.set_site_weights(cnt.astype(np.float64))
    comp = eng.deriv(per_site=True, per_site_ll=True)
    assert abs(comp["sum_ll"] - full["sum_ll"]) <= 1e-12 * abs(full["sum_ll"])
    np.testing.assert_allclose(comp["sum_deriv"], full["sum_deriv"], rtol=1e-11, atol=1e-12 * np.abs(full["sum_deriv"]).max())
    assert np.array_equal(comp["site_ll"][smap], full["site_ll"])
    assert np.array_equal(comp["site_deriv"][smap], full["site_deriv"])
    eng.close()


def test_packed_nibble_codes_are_the_same_upload():
    """PLF_CODES_PACKED4 (two codes per byte, include/plf.h) against uint8 codes: identical per-site results, for an odd
    and an even node count, through plf_set_data and plf_set_data_async; more than 16 definitions are refused."""
    from phyly_b200 import engine as E
    import bench

    class A:
        pass
    args = A(); args.taxa = 24; args.sites = 5003
    pb = bench.build_problem(args, 0, 0)
    eng = pb["eng"]
    defs = np.array(bench.DEFS, dtype=np.float64)
    codes = pb["codes"]
    assert codes.shape[1] % 2 == 1
    w = 1.0 + np.random.default_rng(1).poisson(2.0, pb["S"]).astype(np.float64)
    eng.set_data(defs, codes); eng.set_site_weights(w)
    want = eng.deriv(per_site=True, per_site_ll=True)
    packed = E.pack4(codes)
    assert packed.shape == (pb["S"], (codes.shape[1] + 1) // 2)
    for asynchronous in (False, True):
        eng.set_data_packed(defs, packed, weights=w, asynchronous=asynchronous)
        got = eng.deriv(per_site=True, per_site_ll=True)
        assert np.array_equal(got["site_ll"], want["site_ll"])
        assert np.array_equal(got["site_deriv"], want["site_deriv"])
        assert got["sum_ll"] == want["sum_ll"] and np.array_equal(got["sum_deriv"], want["sum_deriv"])
    eng.close()
    # even node count: a star with three leaves
    eng = _engine()
    eng.set_tree([0, 3, 3, 3, 3], [1, 2, 3], [0, 1, 2, 3])
    Q = np.array([[0, 1, 2, 1], [1, 0, 1, 2], [2, 1, 0, 1], [1, 2, 1, 0]], dtype=np.float64)
    Q = Q - np.diag(Q.sum(axis=1))
    eng.set_model(Q, np.zeros((4, 4)), [0.1, 0.2, 0.3], [1.0], [1.0], 3, [0.25] * 4)
    c4 = np.random.default_rng(2).integers(0, 5, (301, 4)).astype(np.uint8)
    c4[:, 0] = 4
    eng.set_data(defs, c4)
    want_ll, _ = eng.ll()
    eng.set_data_packed(defs, E.pack4(c4))
    got_ll, _ = eng.ll()
    assert np.array_equal(got_ll, want_ll)
    with pytest.raises(E.EngineError):
        eng.set_data_packed(np.vstack([np.eye(4)] * 5), E.pack4(c4))
    eng.close()


def test_four_state_expm_kernel_gives_the_bits_of_the_generic_one(monkeypatch):
    """expm4_dd_kernel (half a warp per 4 x 4 matrix, shuffles) follows the arithmetic of expm_dd_kernel step by step:
    P and D are identical to the last bit, for short, long and zero-length branches, an invariant category and a rate
    matrix with structural zeros."""
    import bench

    class A:
        pass
    args = A(); args.taxa = 40; args.sites = 64
    pb = bench.build_problem(args, 0, 0)
    eng = pb["eng"]
    rng = np.random.default_rng(9)
    E = pb["E"]
    cases = [pb["edge_rates"], rng.exponential(1.0, E) * 10.0 ** rng.integers(-12, 3, E), np.zeros(E),
             np.full(E, 750.0), rng.exponential(0.01, E)]
    for rates in cases:
        monkeypatch.delenv("PLF_NO_EXPM4", raising=False)
        eng.set_edge_rates(rates)
        P4, D4 = eng.transition_matrices(), eng.derivative_matrices()
        monkeypatch.setenv("PLF_NO_EXPM4", "1")
        eng.set_edge_rates(rates)
        Pg, Dg = eng.transition_matrices(), eng.derivative_matrices()
        assert np.array_equal(P4, Pg) and np.array_equal(D4, Dg)
        assert np.all(np.isfinite(P4)) and np.allclose(P4.sum(axis=3), 1.0, rtol=0, atol=1e-14)
    eng.close()
    # a sparse rate matrix (structural zeros are skipped in both kernels) with an invariant category
    monkeypatch.delenv("PLF_NO_EXPM4", raising=False)
    prob = H.random_problem(77, ntips=9, n=4, S=5, ncat=3)
    md = prob["model_and_data"]
    md["rate_matrix"] = [[0, 1, 0, 0], [0, 0, 2, 0], [0.5, 0, 0, 3], [0, 0, 0, 0]]
    md["rate_divisor"] = 2.0
    md["root_prior"] = [0.1, 0.2, 0.3, 0.4]
    md["rate_mixture"] = {"rates": [0.0, 0.5, 3.0], "prior": [0.2, 0.3, 0.5]}
    m = O.parse_model(md)
    out = []
    for env in (None, "1"):
        if env:
            monkeypatch.setenv("PLF_NO_EXPM4", env)
        e2 = _engine()
        H.fill_engine(e2, m)
        out.append((e2.transition_matrices(), e2.derivative_matrices()))
        e2.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
