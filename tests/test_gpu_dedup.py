"""GPU tier: program deduplication (SNAQ_OPT_DEDUP, include/snaq_b200.h).  Quartets whose evaluation programs are identical
have the same expected CFs, and the pseudo-deviance is linear in the weights 100*obsCF, so one evaluation per distinct
program with the summed weights gives the same value.  Checked against the plain path and the oracle."""
import os

import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate

pytestmark = pytest.mark.gpu


def pair(n, h, seed, ngenes=80):
    net = None
    for attempt in range(30):
        try:
            net = simulate.random_level1_network(n, h, seed + 1000 * attempt, min_cycle=3 if seed % 3 == 0 else 4)
            break
        except S.SnaqError:
            continue
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    flat = net.flatten(taxa)
    plain, dd = S.Engine(0, dedup=False), S.Engine(0, dedup=True)
    for e in (plain, dd):
        e.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
        e.set_network(flat)
    x0, _, upper = plain.get_params()
    ex, _, _ = plain.eval_quartets(x0)
    obs = simulate.obs_from_expcf(ex, ngenes, seed)
    for e in (plain, dd):
        e.set_obscf(obs)
    return plain, dd, flat, qt, obs, x0, upper


@pytest.mark.parametrize("seed", range(6))
def test_dedup_matches_the_plain_path_and_the_oracle(oracle, seed):
    n, h = 10 + 3 * seed, seed % 5
    plain, dd, flat, qt, obs, x0, upper = pair(n, h, 40 + seed)
    nu = dd.counters()["distinct_programs"]
    assert 0 < nu < len(qt)
    assert plain.counters()["distinct_programs"] > 0           # the run table is always built (single-vector paths, k_eval_runs)
    X = simulate.parameter_batch(x0, upper, 70, seed)
    X[3] = np.where(upper > 1.5, 10.0, 0.5)                      # every length at its cap
    want = oracle.evaluate(flat, qt, obs, X)
    for P in (1, 3, 4, 9, 33, 70):                               # thread-per-program, batched and chunked paths
        a, b = plain.eval(X[:P]), dd.eval(X[:P])
        np.testing.assert_allclose(b, a, rtol=1e-12)
        np.testing.assert_allclose(b, want[:P], rtol=1e-10)
    fa, ga = plain.eval_grad(X[1])
    fb, gb = dd.eval_grad(X[1])
    assert abs(fa - fb) <= 1e-12 * abs(fa) and np.max(np.abs(ga - gb)) <= 1e-8 * (1 + np.max(np.abs(ga)))
    # per-quartet read-backs come from the full table: identical
    for u, v in zip(plain.eval_quartets(X[2]), dd.eval_quartets(X[2])):
        assert np.array_equal(u, v)
    # data refresh (bootstrap) and sampling masks rebuild the summed weights
    obs2 = simulate.obs_from_expcf(oracle.extract(flat, qt)["expcf"], 30, 99)
    mask = (np.arange(len(qt)) % 3 != 1).astype(np.uint8)
    for e in (plain, dd):
        e.set_obscf(obs2)
        e.set_sampled(mask)
    np.testing.assert_allclose(dd.eval(X[:20]), oracle.evaluate(flat, qt, obs2, X[:20], sampled=mask), rtol=1e-10)
    np.testing.assert_allclose(dd.eval(X[:20]), plain.eval(X[:20]), rtol=1e-12)


def test_dedup_optbl_same_optimum_and_perfect_data():
    plain, dd, flat, qt, obs, x0, upper = pair(18, 2, 61, ngenes=150)
    start = np.clip(x0 * 1.25, 0, upper)
    xa, ra = plain.optbl(start, 1e-12, 1e-10, 1e-10, 1e-10)
    xb, rb = dd.optbl(start, 1e-12, 1e-10, 1e-10, 1e-10)
    assert abs(ra["fmin"] - rb["fmin"]) <= 1e-9 * abs(ra["fmin"])
    assert abs(plain.eval(xb)[0] - rb["fmin"]) <= 1e-11 * abs(rb["fmin"])
    e, _, _ = dd.eval_quartets(x0)
    dd.set_obscf(e)                                              # perfect data: deviance 0 from summed weights too
    assert abs(dd.eval(x0)[0]) < 1e-9


def test_dedup_at_100_taxa_collapses_the_table():
    n = 100
    net = simulate.random_level1_network(n, 3, 20261021)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    plain, dd = S.Engine(0, dedup=False), S.Engine(0, dedup=True)
    rng = np.random.default_rng(3)
    obs = rng.dirichlet([4, 2, 2], size=len(qt))
    for e in (plain, dd):
        e.set_data(S.DataCF(taxa, qt, obs))
        e.set_network(net.flatten(taxa))
    assert dd.counters()["distinct_programs"] < 20000            # 3,921,225 quartets
    x0, _, upper = plain.get_params()
    X = simulate.parameter_batch(x0, upper, 40, 5)
    np.testing.assert_allclose(dd.eval(X), plain.eval(X), rtol=1e-12)
    np.testing.assert_allclose(dd.eval(X[:1]), plain.eval(X[:1]), rtol=1e-12)


@pytest.mark.parametrize("dedup", [False, True])
def test_negative_expcf_penalty_is_per_sampled_quartet(oracle, dedup):
    """The -1e15 of a negative expCF (src/pseudolik.jl:1394-1396) is per sampled quartet: not linear in the weights, so the
    deduplicated table carries the number of sampled quartets and the obs*log(obs) shares of every run."""
    from tests import kat
    net = S.readnewicklevel1(kat.NET_NEG)
    taxa = [str(i) for i in range(1, 7)]
    flat = net.flatten(taxa)
    qt = np.concatenate([simulate.all_quartets(6)] * 3)            # repeated rows: runs of identical programs
    rng = np.random.default_rng(2)
    obs = rng.dirichlet([3, 2, 2], size=len(qt))
    eng = S.Engine(0, dedup=dedup)
    eng.set_data(S.DataCF(taxa, qt, obs))
    eng.set_network(flat)
    x0, _, upper = eng.get_params()
    x = x0.copy()
    x[-2:] = [0.9, 0.8]                                            # gammaz sum > 1: negative expCFs
    X = np.stack([x0, x, x0 * 0.9, x])
    for mask in (None, np.zeros(len(qt), np.uint8), (np.arange(len(qt)) % 2).astype(np.uint8), (np.arange(len(qt)) % 3 != 0).astype(np.uint8)):
        eng.set_sampled(mask)
        want = oracle.evaluate(flat, qt, obs, X, sampled=mask)
        for P in (1, 2, 4):
            np.testing.assert_allclose(eng.eval(X[:P]), want[:P], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(eng.eval(np.tile(X, (10, 1))), np.tile(want, 10), rtol=1e-12, atol=1e-9)    # batched kernels
        f, g = eng.eval_grad(x)
        assert abs(f - want[1]) <= 1e-12 * max(1.0, abs(want[1]))
    assert want.max() > 1e15
