"""Several contexts on one GPU driven from several host threads, and snaq_optbl_topologies (the
candidate topologies of the search loop, src/snaq_optimization.jl:1362-1432, optimised concurrently)."""
import threading

import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate

pytestmark = pytest.mark.gpu


def data_and_candidates(n, h, ncand, seed):
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    truth = simulate.random_level1_network(n, h, seed)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    eng.set_network(truth.flatten(taxa))
    x0, _, _ = eng.get_params()
    obs = simulate.obs_from_expcf(eng.eval_quartets(x0)[0], 300, seed)
    cands, s = [], 0
    while len(cands) < ncand:
        s += 1
        try:   # candidates of different sizes: 0..h hybrids, so the kernels' shared-memory tiles differ between threads
            cands.append(simulate.random_level1_network(n, (s + len(cands)) % (h + 1), 1000 * seed + s).flatten(taxa))
        except S.SnaqError:
            pass
    return taxa, qt, obs, cands


def sequential(taxa, qt, obs, cands, **kw):
    eng = S.Engine(0, **kw)
    eng.set_data(S.DataCF(taxa, qt, obs))
    out = []
    for c in cands:
        eng.set_network(c)
        x, _, _ = eng.get_params()
        out.append(eng.optbl(x))
    return out


def test_contexts_on_one_gpu_from_many_threads():
    """the dynamic shared-memory limit of a kernel is process-wide: concurrent contexts with different tile sizes must
    not invalidate each other's launches"""
    taxa, qt, obs, cands = data_and_candidates(16, 2, 12, 3)
    want = sequential(taxa, qt, obs, cands)
    T = 6
    engines = []
    for _ in range(T):
        e = S.Engine(0)
        e.set_data(S.DataCF(taxa, qt, obs))
        engines.append(e)
    got = {}

    def work(t):
        try:
            res = []
            for r in range(3 * len(cands)):
                c = (t + r) % len(cands)
                engines[t].set_network(cands[c])
                x, _, _ = engines[t].get_params()
                xm, info = engines[t].optbl(x)
                res.append((c, xm, info))
                f = engines[t].eval(xm)[0]
                assert f == info["fmin"] or abs(f - info["fmin"]) <= 1e-12 * abs(f)
            got[t] = res
        except Exception as ex:  # noqa: BLE001 - reported by the main thread
            got[t] = ex

    th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    for t in range(T):
        assert not isinstance(got[t], Exception), got[t]
        for c, xm, info in got[t]:
            assert np.array_equal(xm, want[c][0]) and info == want[c][1]


@pytest.mark.parametrize("workers", [1, 3, 8])
def test_optbl_topologies_equals_one_optbl_per_candidate(workers):
    taxa, qt, obs, cands = data_and_candidates(14, 1, 10, 5)
    want = sequential(taxa, qt, obs, cands)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, obs))
    eng.set_network(cands[0])
    before = eng.eval(eng.get_params()[0])[0]
    got = eng.optbl_topologies(cands, workers=workers)
    assert len(got) == len(cands)
    for (xm, info), (xw, iw) in zip(got, want):
        assert np.array_equal(xm, xw) and info == iw
    assert eng.eval(eng.get_params()[0])[0] == before          # the context's own network is untouched
    assert eng.optbl_topologies([]) == []


def test_optbl_topologies_follows_the_data_of_the_context():
    """bootstrap refresh (snaq_set_obscf) and the sampling mask reach the worker contexts"""
    taxa, qt, obs, cands = data_and_candidates(12, 1, 4, 7)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, obs))
    first = eng.optbl_topologies(cands, workers=2)
    rng = np.random.default_rng(1)
    obs2 = rng.dirichlet(np.ones(3), len(qt))
    mask = (rng.uniform(size=len(qt)) < 0.5).astype(np.uint8)
    eng.set_obscf(obs2)
    eng.set_sampled(mask)
    got = eng.optbl_topologies(cands, workers=2)
    ref = S.Engine(0)
    ref.set_data(S.DataCF(taxa, qt, obs2))
    ref.set_sampled(mask)
    for c, (xm, info) in zip(cands, got):
        ref.set_network(c)
        x, _, _ = ref.get_params()
        xw, iw = ref.optbl(x)
        assert np.array_equal(xm, xw) and info == iw
    assert any(a[1]["fmin"] != b[1]["fmin"] for a, b in zip(first, got))
    eng.set_sampled(None)
    eng.set_obscf(obs)
    again = eng.optbl_topologies(cands, workers=2)
    for a, b in zip(first, again):
        assert np.array_equal(a[0], b[0]) and a[1] == b[1]


def test_optbl_topologies_after_the_context_gets_another_table():
    """snaq_set_data with a table of another size: the worker contexts are re-created from the new data"""
    taxa, qt, obs, cands = data_and_candidates(12, 1, 4, 13)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, obs))
    eng.optbl_topologies(cands, workers=3)
    keep = np.arange(len(qt)) % 3 != 0
    qt2, obs2 = qt[keep], obs[keep]
    eng.set_data(S.DataCF(taxa, qt2, obs2))
    got = eng.optbl_topologies(cands, workers=3)
    want = sequential(taxa, qt2, obs2, cands)
    for (xm, info), (xw, iw) in zip(got, want):
        assert np.array_equal(xm, xw) and info == iw


def test_optbl_topologies_reports_the_failing_candidate():
    taxa, qt, obs, cands = data_and_candidates(10, 1, 3, 9)
    bad = simulate.random_level1_network(9, 0, 4).flatten(taxa)     # a network that misses taxon t10 of the data
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, obs))
    with pytest.raises(S.SnaqError, match="topology 2"):
        eng.optbl_topologies([cands[0], bad, cands[1]], workers=2)


def test_optbl_topologies_with_deduplicated_tables():
    taxa, qt, obs, cands = data_and_candidates(13, 2, 6, 11)
    want = sequential(taxa, qt, obs, cands, dedup=True)
    eng = S.Engine(0, dedup=True)
    eng.set_data(S.DataCF(taxa, qt, obs))
    for (xm, info), (xw, iw) in zip(eng.optbl_topologies(cands, workers=4), want):
        assert np.array_equal(xm, xw) and info == iw


def test_optBL_candidates_on_the_reference_example_networks():
    """examples/net1.networks: five topologies the reference optimised and scored on examples/tableCF.txt.  Optimising
    them again, all at once, can only keep or lower each score, and must agree with one optBL_ per network."""
    import os
    from . import kat
    here = os.path.dirname(os.path.abspath(__file__))
    d = S.readtableCF(os.path.join(here, "golden", "tableCF.txt"))
    nets = [S.readnewicklevel1(nw) for nw, _ in kat.NET1_NETWORKS]
    res = S.optBL_candidates_(nets, d, workers=5)
    for net, r, (nw, lik) in zip(nets, res, kat.NET1_NETWORKS):
        assert r["status"] in (3, 4, 5)
        assert net.loglik == r["fmin"] <= lik * (1 + 1e-9)
        assert r["fmin"] >= 0.9 * lik - 1e-6                 # the reference's optimum is a good one already
        one = S.readnewicklevel1(nw)
        r1 = S.optBL_(one, d)
        assert r1 == r and one.ht == net.ht
