"""hasEdge / indexht (updateHasEdge!, src/pseudolik.jl:392-464; parameters!(qnet,net), src/snaq_optimization.jl:106-180):
the structural rule of snaq.jl_b200/csrc/snaq_hasedge.h against the reference's golden vectors and the oracle,
for networks with at most one hybrid node (the case snaq_get_hasedge accepts)."""
import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate
from tests import hostsim_py as hs
from tests import kat


def nets():
    return {"G": S.readnewicklevel1(kat.CASE_G), "F": kat.case_f_network(S), "I": S.readnewicklevel1(kat.CASE_I)}


def indexht_from(he, bd1, nh, nt):
    return [i + 1 for i in range(len(he)) if he[i] and (not bd1 or i >= nh + nt)]


@pytest.mark.parametrize("case", ["G", "F", "I"])
def test_golden_hasedge_through_the_device_functions(oracle, case):
    # test/test_hasEdge.jl:21-41 (G), :65-85 (F), :106-126 (I)
    flat = nets()[case].flatten(kat.TAXA5)
    he, bd1 = hs.hasedge(flat, 5, kat.q5_ids())
    p = oracle.parameters(flat)
    for i in range(5):
        assert list(he[i]) == kat.HASEDGE[case][i], (case, i)
        assert indexht_from(he[i], bd1[i], p["nh"], p["nt"]) == kat.INDEXHT[case][i], (case, i)


@pytest.mark.parametrize("block", range(4))
def test_hasedge_matches_the_oracle_with_at_most_one_hybrid(oracle, block):
    nq = 0
    for seed in range(block * 30, block * 30 + 30):
        n = 8 + seed % 18
        h = seed % 2
        net = simulate.random_level1_network(n, h, seed, min_cycle=3 if seed % 3 == 0 else 4)
        taxa = [f"t{i + 1}" for i in range(n)]
        rng = np.random.default_rng(seed)
        rng.shuffle(net.leaf)                                   # the pruning order is net.leaf order
        flat = net.flatten(taxa)
        qt = simulate.all_quartets(n)
        o = oracle.extract(flat, qt)
        he, bd1 = hs.hasedge(flat, n, qt)
        assert np.array_equal(he, o["hasedge"]), seed
        p = oracle.parameters(flat)
        for i in range(0, len(qt), 7):
            assert indexht_from(he[i], bd1[i], p["nh"], p["nt"]) == list(o["indexht"][i][: o["nindexht"][i]]), (seed, i)
        nq += len(qt)
    assert nq > 50000


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["G", "F", "I"])
def test_golden_hasedge_through_the_abi(case):
    net = nets()[case]
    eng = S.Engine(0)
    eng.set_data(S.DataCF(kat.TAXA5, kat.q5_ids(), np.full((5, 3), 1 / 3)))
    eng.set_network(net.flatten(kat.TAXA5))
    he, ix = eng.hasedge()
    for i in range(5):
        assert list(he[i]) == kat.HASEDGE[case][i] and ix[i] == kat.INDEXHT[case][i], (case, i)


@pytest.mark.gpu
def test_hasedge_through_the_abi_matches_the_oracle_and_rejects_nested_cycles(oracle):
    for seed in (3, 8, 13, 21):
        n, h = 10 + seed % 9, seed % 2
        net = simulate.random_level1_network(n, h, 100 + seed, min_cycle=3 if seed % 3 == 0 else 4)
        taxa = [f"t{i + 1}" for i in range(n)]
        flat = net.flatten(taxa)
        qt = simulate.all_quartets(n)
        eng = S.Engine(0)
        eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
        eng.set_network(flat)
        he, ix = eng.hasedge()
        o = oracle.extract(flat, qt)
        assert np.array_equal(he, o["hasedge"])
        assert all(ix[i] == list(o["indexht"][i][: o["nindexht"][i]]) for i in range(len(qt)))
    net = simulate.random_level1_network(14, 2, 5)
    taxa = [f"t{i + 1}" for i in range(14)]
    qt = simulate.all_quartets(14)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    eng.set_network(net.flatten(taxa))
    with pytest.raises(S.SnaqError):
        eng.hasedge()
