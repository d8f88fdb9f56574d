"""GPU tier: the CUDA path through the C ABI against the oracle, the reference's golden
vectors, and size-independent properties at BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star): descriptors bit-exact; logPseudoLik within 1e-10
relative; expCF compared at 1e-13 absolute (they are O(1))."""
import itertools
import os

import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate
from tests import kat

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
RTOL_LIK = 1e-10



@pytest.fixture(params=[None, False], ids=["run_table", "per_quartet"], autouse=True)
def table_mode(request, monkeypatch):
    """every test of this module runs on the default run table (distinct programs, summed weights) and on the per-quartet
    kernels (SNAQ_OPT_PER_QUARTET)"""
    monkeypatch.setattr(S.Engine, "DEFAULT_DEDUP", request.param)


def make_engine(net, taxa, qt, obs, **kw):
    d = S.DataCF(taxa, qt, obs)
    eng = S.Engine(0, **kw)
    eng.set_data(d)
    flat = net.flatten(taxa)
    eng.set_network(flat)
    return eng, d, flat


def nets5():
    return {"G": S.readnewicklevel1(kat.CASE_G), "F": kat.case_f_network(S), "I": S.readnewicklevel1(kat.CASE_I)}


# ---------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("case", ["G", "F", "I"])
def test_golden_5taxon_cases(case):
    net = nets5()[case]
    obs = np.tile([0.5, 0.4, 0.1], (5, 1))
    eng, d, flat = make_engine(net, kat.TAXA5, kat.q5_ids(), obs)
    x0, numht, upper = eng.get_params()
    assert list(numht) == kat.NUMHT[case]                          # test/test_hasEdge.jl
    np.testing.assert_allclose(x0, kat.HT[case], rtol=0, atol=1e-15)
    e, _, _ = eng.eval_quartets(x0)
    desc = eng.descriptors()
    for i, (which, types, t1, formula, expcf) in enumerate(kat.EXPCF[case]):   # test/test_calculateExpCF.jl
        assert desc["which"][i] == which
        assert list(desc["types"][i][: desc["ntype"][i]]) == types
        if which == 1:
            assert list(desc["formula"][i]) == formula
        np.testing.assert_allclose(e[i], expcf, rtol=0, atol=1e-15)


def test_golden_closed_forms_in_x():
    n = nets5()                                                    # test/test_optBLparts.jl:139-147,254-259,375-383
    for case, x, fn in (("G", [0.3, 0.2, 0.1, 2.0], kat.optbl_G), ("F", [0.3, 0.2, 0.1], kat.optbl_F),
                        ("I", [0.2, 0.2, 0.1, 0.1, 0.1], kat.optbl_I)):
        eng, _, _ = make_engine(n[case], kat.TAXA5, kat.q5_ids(), np.tile([0.5, 0.4, 0.1], (5, 1)))
        e, _, _ = eng.eval_quartets(np.array(x))
        np.testing.assert_allclose(e, fn(x), rtol=1e-13, atol=0)


def test_golden_likelihoods():
    # test/test_correctLik.jl:45,64
    eng, d, _ = make_engine(S.readnewicklevel1(kat.TREE5), kat.TAXA5, kat.q5_ids(), kat.OBS_TREE5)
    v = eng.eval(eng.get_params()[0])[0]
    assert abs(v - kat.LIK_TREE5) <= RTOL_LIK * kat.LIK_TREE5
    eng, d, _ = make_engine(S.readnewicklevel1(kat.CASE_G), kat.TAXA5, kat.q5_ids(), kat.OBS_CASEH)
    v = eng.eval(eng.get_params()[0])[0]
    assert abs(v - kat.LIK_CASEG_ON_H) <= RTOL_LIK * kat.LIK_CASEG_ON_H


@pytest.mark.parametrize("newick,lik", [(kat.NET0, kat.NET0_LIK), (kat.NET1, kat.NET1_LIK)])
def test_golden_example_outputs(newick, lik):
    # examples/net0.out:1 and net1.out:1 on examples/tableCF.txt, through topologyQpseudolik!
    d = S.readtableCF(os.path.join(HERE, "golden", "tableCF.txt"))
    net = S.readnewicklevel1(newick)
    val = S.topologyQpseudolik_(net, d)
    assert abs(val - lik) <= RTOL_LIK * lik
    assert net.loglik == val
    assert d.expCF.shape == (15, 3) and np.allclose(d.expCF.sum(axis=1), 1.0)
    # logPseudoLik(d) == -(sum of the per-quartet values), src/pseudolik.jl:1417-1423
    assert abs(-d.logPseudoLik.sum() - val) <= 1e-12 * val


@pytest.mark.parametrize("i", range(8))
def test_golden_more_example_outputs(i):
    # examples/net1.networks (5 scored topologies), net2.out:1, net3.out:1 on tableCF.txt; net1_snaq.out:1 on tableCFCI.csv
    cases = [(nw, lik, "tableCF.txt") for nw, lik in kat.NET1_NETWORKS]
    cases += [(kat.NET2_OUT[0], kat.NET2_OUT[1], "tableCF.txt"), (kat.NET3_OUT[0], kat.NET3_OUT[1], "tableCF.txt"),
              (kat.NET1_SNAQ_OUT[0], kat.NET1_SNAQ_OUT[1], "tableCFCI.csv")]
    newick, lik, table = cases[i]
    d = S.readtableCF(os.path.join(HERE, "golden", table))
    net = S.readnewicklevel1(newick)
    assert abs(S.topologyQpseudolik_(net, d) - lik) <= 1e-12 * lik
    f, g = d._engine.eval_grad(np.array(net.ht))
    assert abs(f - lik) <= 1e-12 * lik


def test_golden_scored_topologies_in_one_call():
    """examples/net1.networks: the five candidate topologies of the h=1 run scored by ONE snaq_eval_topologies call."""
    d = S.readtableCF(os.path.join(HERE, "golden", "tableCF.txt"))
    eng = S.Engine(0)
    eng.set_data(d)
    flats = [S.readnewicklevel1(nw).flatten(d.taxa) for nw, _ in kat.NET1_NETWORKS]
    got = eng.eval_topologies(flats)
    np.testing.assert_allclose(got, [lik for _, lik in kat.NET1_NETWORKS], rtol=1e-12)
    x_last, _, _ = eng.get_params()                                   # the context holds the last network
    assert abs(eng.eval(x_last)[0] - kat.NET1_NETWORKS[-1][1]) <= 1e-12 * kat.NET1_NETWORKS[-1][1]
    assert eng.eval_topologies([]).shape == (0,)


def test_golden_four_cycle_and_negative_expcf():
    net = S.readnewicklevel1(kat.NET_4CYCLE)                       # test/test_multipleAlleles.jl:66-72
    perms = np.array(list(itertools.permutations(range(4))), dtype=np.uint16)
    eng, _, _ = make_engine(net, ["a", "b", "c", "d"], perms, np.tile([0.6, 0.39, 0.01], (24, 1)))
    e, _, _ = eng.eval_quartets(eng.get_params()[0])
    np.testing.assert_allclose(e[0], kat.EXPCF_4CYCLE, rtol=0, atol=2e-16)
    net = S.readnewicklevel1(kat.NET_NEG)                          # test/test_calculateExpCF2.jl:7-35
    taxa = ["1", "2", "3", "4", "5", "6"]
    q = np.array([[taxa.index(t) for t in ["6", "5", "4", "1"]]], dtype=np.uint16)
    eng, _, _ = make_engine(net, taxa, q, [[0.5, 0.4, 0.1]])
    x = eng.get_params()[0]
    x[-2:] = [1.0, 0.067]
    e, _, _ = eng.eval_quartets(x)
    x2 = x.copy(); x2[-2:] = [0.067, 1.0]
    e2, _, _ = eng.eval_quartets(x2)
    assert min(e[0].min(), e2[0].min()) < 0 and (e[0][1] < 0 or e2[0][1] < 0)
    assert eng.eval(x)[0] > 9e14                                   # -1e15 penalty, src/pseudolik.jl:1394-1396
    assert eng.descriptors()["which"][0] == 2


def test_golden_perfect_data(oracle):
    net = S.readnewicklevel1(kat.NET_PERFECT)                      # test/test_perfectData.jl:1372-1382
    taxa = [str(i) for i in range(1, 16)]
    qt = simulate.all_quartets(15)
    eng, d, flat = make_engine(net, taxa, qt, np.full((1365, 3), 1 / 3))
    x0 = eng.get_params()[0]
    e, _, _ = eng.eval_quartets(x0)
    eng.set_obscf(e)
    assert abs(eng.eval(x0)[0]) <= 1e-10                           # reference: atol 1e-12 on 1365 quartets
    o = oracle.extract(flat, qt)
    assert np.max(np.abs(e - o["expcf"])) < 1e-14
    desc = eng.descriptors()
    assert np.array_equal(desc["which"], o["which"]) and np.array_equal(desc["formula"], o["formula"])


# ---------------------------------------------------------------- fuzz against the oracle
@pytest.mark.parametrize("seed", range(24))
def test_random_networks_against_oracle(oracle, seed):
    n = 8 + (seed * 5) % 19
    h = seed % 6
    net = None
    for attempt in range(20):                                      # small trees cannot always take h reticulations
        try:
            net = simulate.random_level1_network(n, h, 5000 + seed + 1000 * attempt, min_cycle=3 if seed % 3 == 0 else 4)
            break
        except S.SnaqError:
            if attempt % 4 == 3 and h > 0:
                h -= 1
    assert net is not None
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    flat = net.flatten(taxa)
    o = oracle.extract(flat, qt)
    obs = simulate.obs_from_expcf(o["expcf"], 30, seed)           # few genes: many exact zeros
    eng, d, _ = make_engine(net, taxa, qt, obs)
    x0, numht, upper = eng.get_params()
    p = oracle.parameters(flat)
    assert np.array_equal(numht, p["numht"]) and np.array_equal(x0, p["x"]) and np.array_equal(upper, p["upper"])
    # descriptors: bit exact
    desc = eng.descriptors()
    assert np.array_equal(desc["which"], o["which"])
    assert np.array_equal(desc["formula"], o["formula"])
    for i in range(len(qt)):
        assert list(desc["types"][i][: desc["ntype"][i]]) == [t for t in o["types"][i][: o["ntype"][i]] if t != 1]
    # likelihoods: lanes kernel (P=70) and thread-per-quartet kernel (P=3), incl. bounds and the 10.0 cap
    X = simulate.parameter_batch(x0, upper, 70, seed)
    rng = np.random.default_rng(seed)
    X[5] = np.where(upper > 1.5, rng.choice([0.0, 10.0, 5.0, 0.3], size=len(x0)), rng.choice([0.01, 0.5, 0.3], size=len(x0)))
    X[6] = np.where(upper > 1.5, 10.0, 0.5)
    want = oracle.evaluate(flat, qt, obs, X)
    got = eng.eval(X)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.max(np.abs(got[fin] - want[fin]) / np.abs(want[fin])) <= RTOL_LIK
    got3 = eng.eval(X[4:7])
    f3 = fin[4:7]
    assert np.max(np.abs(got3[f3] - want[4:7][f3]) / np.abs(want[4:7][f3])) <= RTOL_LIK
    # per-quartet read-back
    e, lp, dl = eng.eval_quartets(X[1])
    _, eo = oracle.evaluate(flat, qt, obs, X[1], want_expcf=True)
    assert np.max(np.abs(e - eo)) < 1e-13
    np.testing.assert_allclose(dl, np.abs(obs - eo).sum(axis=1), atol=1e-13)


def test_sampled_mask_obscf_refresh_and_errors(oracle):
    net = simulate.random_level1_network(14, 2, 77)
    taxa = [f"t{i + 1}" for i in range(14)]
    qt = simulate.all_quartets(14)
    flat = net.flatten(taxa)
    o = oracle.extract(flat, qt)
    obs = simulate.obs_from_expcf(o["expcf"], 100, 1)
    eng, d, _ = make_engine(net, taxa, qt, obs)
    x0, _, upper = eng.get_params()
    X = simulate.parameter_batch(x0, upper, 20, 3)
    full = eng.eval(X)
    mask = (np.arange(len(qt)) % 3 != 0).astype(np.uint8)
    eng.set_sampled(mask)
    a = eng.eval(X)
    np.testing.assert_allclose(a, oracle.evaluate(flat, qt, obs, X, sampled=mask), rtol=RTOL_LIK)
    eng.set_sampled(1 - mask)
    b = eng.eval(X)
    np.testing.assert_allclose(a + b, full, rtol=1e-12)            # the objective is a plain sum over quartets
    eng.set_sampled(None)
    obs2 = simulate.obs_from_expcf(o["expcf"], 100, 2)             # bootstrap-style refresh of obsCF
    eng.set_obscf(obs2)
    np.testing.assert_allclose(eng.eval(X), oracle.evaluate(flat, qt, obs2, X), rtol=RTOL_LIK)
    # error behaviour (src/snaq_optimization.jl:186-190,213,221,228; src/auxiliary.jl:119)
    bad = x0.copy(); bad[0] = 1.5 if upper[0] < 1.5 else -0.1
    with pytest.raises(S.SnaqError, match="between 0,1|nonnegative"):
        eng.eval(bad)
    with pytest.raises(S.SnaqError, match="same length|wrong length"):
        eng.eval(np.zeros(len(x0) + 1))
    eng2 = S.Engine(0)
    with pytest.raises(S.SnaqError, match="set_data|set_network"):
        eng2.eval(x0)
    eng2.set_data(d)
    with pytest.raises(S.SnaqError, match="set_network"):
        eng2.eval(x0)
    # a taxon of the data missing from the network
    with pytest.raises(S.SnaqError):
        e3 = S.Engine(0)
        e3.set_data(S.DataCF(taxa + ["extra"], qt, obs))
        qt_bad = qt.copy(); qt_bad[0, 3] = 14
        e3.set_data(S.DataCF(taxa + ["extra"], qt_bad, obs))
        e3.set_network(net.flatten(taxa + ["extra"]))


def test_quartet_sharding_partials_add_up(oracle):
    """world_size=2 contexts (run here on one GPU): the partial sums of the two shards add up to
    the unsharded objective — the value an NCCL all-reduce delivers."""
    net = simulate.random_level1_network(20, 3, 11)
    taxa = [f"t{i + 1}" for i in range(20)]
    qt = simulate.all_quartets(20)
    flat = net.flatten(taxa)
    obs = simulate.obs_from_expcf(oracle.extract(flat, qt)["expcf"], 100, 1)
    eng, d, _ = make_engine(net, taxa, qt, obs)
    x0, _, upper = eng.get_params()
    X = simulate.parameter_batch(x0, upper, 40, 5)
    full = eng.eval(X)
    parts = []
    for r in range(3):
        e, _, _ = make_engine(net, taxa, qt, obs, rank=r, world_size=3)
        parts.append(e.eval(X))
    np.testing.assert_allclose(sum(parts), full, rtol=1e-12)


# ---------------------------------------------------------------- BASELINE.json sizes
def test_config2_30_taxa_4096_vectors(oracle):
    """configs[1]: 30 taxa, h=2, 27,405 quartets x 4096 parameter vectors."""
    net = simulate.random_level1_network(30, 2, 20261020)
    taxa = [f"t{i + 1}" for i in range(30)]
    qt = simulate.all_quartets(30)
    assert len(qt) == 27405
    flat = net.flatten(taxa)
    obs = simulate.obs_from_expcf(oracle.extract(flat, qt)["expcf"], 300, 1)
    eng, d, _ = make_engine(net, taxa, qt, obs)
    x0, _, upper = eng.get_params()
    X = simulate.parameter_batch(x0, upper, 4096, 7)
    got = eng.eval(X)
    idx = np.r_[0:40, 4000:4096:8]
    want = oracle.evaluate(flat, qt, obs, X[idx])
    assert np.max(np.abs(got[idx] - want) / np.abs(want)) <= RTOL_LIK
    # batch composition must not matter: any row evaluated alone gives the same value
    for i in (0, 1, 2047, 4095):
        one = eng.eval(X[i])[0]
        assert abs(one - got[i]) <= 1e-12 * abs(got[i])
    # idempotence / determinism: same launch twice is bit-identical
    assert np.array_equal(got, eng.eval(X))


def test_config3_100_taxa_properties():
    """configs[2] size: 100 taxa, h=3, 3,921,225 quartets (the oracle is too slow here, so
    size-independent properties)."""
    n = 100
    net = simulate.random_level1_network(n, 3, 20261021)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    assert len(qt) == 3921225
    obs = np.full((len(qt), 3), 1 / 3)
    eng, d, flat = make_engine(net, taxa, qt, obs)
    x0, _, upper = eng.get_params()
    e, lp, dl = eng.eval_quartets(x0)
    assert np.allclose(e.sum(axis=1), 1.0, atol=1e-13) and (e > 0).all()
    # perfect data => deviance 0; and the total is the (negated) sum of the per-quartet values
    eng.set_obscf(e)
    assert abs(eng.eval(x0)[0]) < 1e-8
    obs2 = simulate.obs_from_expcf(e, 300, 3)
    eng.set_obscf(obs2)
    X = simulate.parameter_batch(x0, upper, 33, 9)
    v = eng.eval(X)                                                # lanes kernel
    _, lp0, _ = eng.eval_quartets(X[0])
    assert abs(-lp0.sum() - v[0]) <= 1e-11 * abs(v[0])
    v1 = eng.eval(X[:2])                                           # thread-per-quartet kernel
    np.testing.assert_allclose(v1, v[:2], rtol=1e-12)
    # halves of the quartet set add up (what the quartet-sharded all-reduce relies on)
    mask = (np.arange(len(qt)) % 2).astype(np.uint8)
    eng.set_sampled(mask); a = eng.eval(X[:4])
    eng.set_sampled(1 - mask); b = eng.eval(X[:4])
    np.testing.assert_allclose(a + b, v[:4], rtol=1e-12)
    # the structure of the descriptor table: one program per quartet, which in {1,2}
    desc = eng.descriptors()
    assert set(np.unique(desc["which"])) <= {1, 2}
    assert eng.counters()["descriptor_words"] >= len(qt) * 2
