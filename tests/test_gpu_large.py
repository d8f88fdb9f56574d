"""GPU tier at BASELINE.json's full sizes AGAINST THE ORACLE: configs[2] (100 taxa, h = 3, 3,921,225 quartets) and
configs[4] (200 taxa, h = 5, 64,684,950 quartets, also through two quartet-sharded contexts).

The oracle accepts any quartet list, so the full device table is compared with it on >= 5,000 randomly drawn quartets:
which / typeHyb / formula bit-exact, expCF <= 1e-13, per-quartet logPseudoLik <= 1e-10 relative, at the network's own
parameters, at a random vector and at a vector on the 0 / 10 bounds (the setLength! caps).  The batched (run-length),
the small-batch and the per-quartet kernels must agree with each other on the full table, and the gradient pass
(96-bit fixed-point bins) is checked at 64.7 M quartets from a bad start (uniform observed CFs)."""
import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate

pytestmark = pytest.mark.gpu
NSAMPLE = 5000


@pytest.fixture(params=[None, False], ids=["run_table", "per_quartet"])
def table_mode(request, monkeypatch):
    """the default run table (distinct programs, summed weights) and the per-quartet kernels (SNAQ_OPT_PER_QUARTET)"""
    monkeypatch.setattr(S.Engine, "DEFAULT_DEDUP", request.param)
    return request.param


def loglik_rows(expcf, obs):
    """logPseudoLik(quartet), src/pseudolik.jl:1388-1410, from expected and observed CFs (numpy restatement)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        term = np.where(obs == 0.0, 0.0, 100.0 * obs * np.log(expcf / np.where(obs == 0.0, 1.0, obs)))
    term = np.where(expcf < 0.0, -1.0e15, term)
    return term[:, 0] + term[:, 1] + term[:, 2]


def bound_vector(x0, upper, seed):
    """every parameter at one of its bounds or at its own value (thirds): the 10.0 caps and gamma in {0, 1} are hit"""
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, 3, len(x0))
    return np.where(pick == 0, 0.0, np.where(pick == 1, upper, x0))


def check_against_oracle(oracle, flat, qt, obs, sample, vectors, read_rows, descr):
    """read_rows(x) -> (expcf, logpl) rows of the sampled quartets from the device"""
    sub = np.ascontiguousarray(qt[sample])
    for k, x in enumerate(vectors):
        want = oracle.extract(flat, sub, x)
        if k == 0:
            assert np.array_equal(descr["which"], want["which"])
            for i in range(len(sub)):                             # type 1 (a cycle on a pendant path) changes no CF and is not listed
                assert list(descr["types"][i][: descr["ntype"][i]]) == [t for t in want["types"][i][: want["ntype"][i]] if t != 1]
            tree = want["which"] == 1
            assert np.array_equal(descr["formula"][tree], want["formula"][tree])
            assert np.array_equal(descr["formula"][~tree], want["formula"][~tree])
        e, lp = read_rows(x)
        assert np.max(np.abs(e - want["expcf"])) <= 1e-13, (k, np.max(np.abs(e - want["expcf"])))
        ref = loglik_rows(want["expcf"], obs[sample])
        ok = np.isfinite(ref)
        assert np.array_equal(ok, np.isfinite(lp))
        assert np.max(np.abs(lp[ok] - ref[ok]) / np.maximum(1.0, np.abs(ref[ok]))) <= 1e-10


def test_config3_100_taxa_against_oracle(oracle, table_mode):
    n = 100
    net = simulate.random_level1_network(n, 3, 20261021)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    assert len(qt) == 3921225
    flat = net.flatten(taxa)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    eng.set_network(flat)
    x0, numht, upper = eng.get_params()
    p = oracle.parameters(flat)
    assert np.array_equal(numht, p["numht"]) and np.array_equal(x0, p["x"])
    e0, _, _ = eng.eval_quartets(x0)
    obs = simulate.obs_from_expcf(e0, 300, 11)
    eng.set_obscf(obs)
    rng = np.random.default_rng(5)
    sample = np.sort(rng.choice(len(qt), NSAMPLE, replace=False))
    xr = simulate.parameter_batch(x0, upper, 8, 21)[5]
    xb = bound_vector(x0, upper, 3)
    d = eng.descriptors()
    descr = {k: v[sample] for k, v in d.items()}

    def read_rows(x):
        e, lp, _ = eng.eval_quartets(x)
        return e[sample], lp[sample]

    check_against_oracle(oracle, flat, qt, obs, sample, [x0, xr, xb], read_rows, descr)
    # the kernels agree with each other on the FULL table: per-quartet read-back, small batch, batched (run-length)
    X = np.vstack([x0, xr, simulate.parameter_batch(x0, upper, 38, 4)])
    vb = eng.eval(X)                                               # 40 vectors: k_eval_runs
    vs = eng.eval(X[:3])                                           # k_eval_direct
    np.testing.assert_allclose(vs, vb[:3], rtol=1e-12)
    for i in (0, 1):
        _, lp, _ = eng.eval_quartets(X[i])
        assert abs(-np.sum(lp) - vb[i]) <= 1e-11 * abs(vb[i])
    # and with the oracle's total on the sample (every other quartet masked out)
    mask = np.zeros(len(qt), np.uint8)
    mask[sample] = 1
    eng.set_sampled(mask)
    got = eng.eval(X)
    want = oracle.evaluate(flat, qt[sample], obs[sample], X[:6])
    assert np.max(np.abs(got[:6] - want) / np.abs(want)) <= 1e-10
    f, g = eng.eval_grad(X[1])
    assert abs(f - want[1]) <= 1e-10 * abs(want[1]) and np.all(np.isfinite(g))


@pytest.fixture(scope="module")
def table200():
    n = 200
    net = simulate.random_level1_network(n, 5, 20261022)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    assert len(qt) == 64684950
    return net, taxa, qt


def test_config5_200_taxa_sharded_against_oracle(oracle, table200, table_mode):
    """two quartet-sharded contexts (rank 0 and 1 of 2) hold the 64.7 M quartets between them"""
    net, taxa, qt = table200
    flat = net.flatten(taxa)
    nq = len(qt)
    rng = np.random.default_rng(8)
    obs = rng.dirichlet([6.0, 2.0, 2.0], nq)
    d = S.DataCF(taxa, qt, obs)
    engs = [S.Engine(0, rank=r, world_size=2) for r in range(2)]
    begins = [nq * r // 2 for r in range(3)]
    for e in engs:
        e.set_data(d)
        e.set_network(flat)
    x0, numht, upper = engs[0].get_params()
    p = oracle.parameters(flat)
    assert np.array_equal(numht, p["numht"]) and np.array_equal(x0, p["x"])
    sample = np.sort(rng.choice(nq, NSAMPLE, replace=False))
    xr = simulate.parameter_batch(x0, upper, 8, 22)[6]
    xb = bound_vector(x0, upper, 4)

    def local_rows(arr_by_rank):
        out = []
        for q in sample:
            r = 0 if q < begins[1] else 1
            out.append(arr_by_rank[r][q - begins[r]])
        return np.array(out)

    descs = [e.descriptors() for e in engs]                        # the read-backs of a sharded context cover its own block
    assert [e.nq_local for e in engs] == [begins[1] - begins[0], begins[2] - begins[1]]
    descr = {k: local_rows([dd[k] for dd in descs]) for k in descs[0]}
    del descs

    def read_rows(x):
        es, ls = [], []
        for e in engs:
            ee, lp, _ = e.eval_quartets(x, want=("expcf", "logpl"))
            es.append(ee); ls.append(lp)
        return local_rows(es), local_rows(ls)

    check_against_oracle(oracle, flat, qt, obs, sample, [x0, xr, xb], read_rows, descr)
    # partial sums of the two ranks add up to the total of the per-quartet values, for every kernel path
    X = np.vstack([x0, xr, simulate.parameter_batch(x0, upper, 30, 5)])
    tot_b = engs[0].eval(X) + engs[1].eval(X)                      # 32 vectors: run-length kernel
    tot_s = engs[0].eval(X[:2]) + engs[1].eval(X[:2])              # single-call kernel
    np.testing.assert_allclose(tot_s, tot_b[:2], rtol=1e-12)
    lp_sum = 0.0
    for e in engs:
        _, lp, _ = e.eval_quartets(X[1], want=("logpl",))
        lp_sum += float(np.sum(lp))
    assert abs(-lp_sum - tot_b[1]) <= 1e-11 * abs(tot_b[1])
    # without a communicator a sharded context refuses the calls whose result would silently be a partial one
    with pytest.raises(S.SnaqError):
        engs[0].eval_grad(x0)
    with pytest.raises(S.SnaqError):
        engs[0].optbl(x0)


def test_config5_gradient_64M_quartets_bad_start(table200):
    """uniform observed CFs (objective ~ 2e10 at a bad start): the gradient of the full table against the deduplicated
    engine (whose sums are formed differently: one rounded contribution per RUN) and against central differences"""
    net, taxa, qt = table200
    flat = net.flatten(taxa)
    nq = len(qt)
    obs = np.full((nq, 3), 1 / 3)
    d = S.DataCF(taxa, qt, obs)
    eng = S.Engine(0, dedup=False)
    eng.set_data(d)
    eng.set_network(flat)
    x0, _, upper = eng.get_params()
    rng = np.random.default_rng(1)
    x = np.clip(x0 * rng.uniform(0.2, 3.0, len(x0)), 0.0, upper)   # a bad start
    f, g, h = eng.eval_grad(x, want_hdiag=True)
    f2, g2 = eng.eval_grad(x)
    assert f == f2 and np.array_equal(g, g2)                       # bit-reproducible, with and without the curvature bins
    assert f > 1e9
    v = eng.eval(x)[0]
    assert abs(f - v) <= 1e-11 * abs(v)
    dd = S.Engine(0, dedup=True)
    dd.set_data(d)
    dd.set_network(flat)
    fd, gd, hd = dd.eval_grad(x, want_hdiag=True)
    assert abs(fd - f) <= 1e-11 * abs(f)
    scale = 1.0 + np.abs(gd)
    assert np.max(np.abs(g - gd) / scale) <= 1e-6                  # 2^-32 rounding of 64.7 M contributions: ~1e-10 * sqrt(nq)
    assert np.max(np.abs(h - hd) / (1.0 + np.abs(hd))) <= 1e-6
    # central differences of the objective on a few coordinates (a wrapped bin would be off by multiples of 4.3e9)
    for i in rng.choice(len(x), 6, replace=False):
        hstep = 1e-4 * max(1.0, x[i])
        lo = max(0.0, x[i] - hstep); hi = min(upper[i], x[i] + hstep)
        xa, xb = x.copy(), x.copy()
        xa[i] = lo; xb[i] = hi
        va, vb = dd.eval(np.vstack([xa, xb]))
        fdiff = (vb - va) / (hi - lo)
        assert abs(fdiff - g[i]) <= 1e-3 * abs(g[i]) + 2e4, (i, fdiff, g[i])     # (the objective is ~2e10: differences carry ~1e3 of noise)


def test_gradient_bins_at_the_corners():
    """gamma -> 1 and lengths at the 10.0 cap make single derivatives ~1e7 (1 / expCF with expCF ~ exp(-10) / 3): a block's
    sum leaves the +-2^31 range that 64-bit 2^-32 bins had, and on the deduplicated table (one entry per run, weights
    summed) a SINGLE contribution does.  With the 96-bit bins the full and the deduplicated table give the same gradient,
    and it matches one-sided differences of the objective"""
    n = 50
    net = simulate.random_level1_network(n, 3, 77)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    obs = np.full((len(qt), 3), 1 / 3)
    d = S.DataCF(taxa, qt, obs)
    flat = net.flatten(taxa)
    eng = S.Engine(0, dedup=False); eng.set_data(d); eng.set_network(flat)
    dd = S.Engine(0, dedup=True); dd.set_data(d); dd.set_network(flat)
    x0, _, upper = eng.get_params()
    x = upper.copy()                                               # every length 10
    x[: eng.nh] = 1.0 - 1e-7                                       # gamma next to 1
    f, g = eng.eval_grad(x)
    fd, gd = dd.eval_grad(x)
    assert np.isfinite(f) and abs(f - fd) <= 1e-10 * abs(f)
    assert np.max(np.abs(g)) > 2.0 ** 31                           # beyond what one 64-bit 2^-32 bin could hold
    assert np.max(np.abs(g - gd) / (1.0 + np.abs(gd))) <= 2e-6    # 2^-32 rounding of 230,300 contributions per slot, twice
    for i in range(eng.nh):                                        # backward differences in the gammas (the steep coordinates)
        xa = x.copy(); xa[i] -= 1e-9
        va, vb = eng.eval(np.vstack([xa, x]))
        assert abs((vb - va) / (x[i] - xa[i]) - g[i]) <= 2e-2 * abs(g[i]) + 1e3, (i, (vb - va) / (x[i] - xa[i]), g[i])
