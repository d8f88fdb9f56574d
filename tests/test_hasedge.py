"""hasEdge / indexht (updateHasEdge!, src/pseudolik.jl:392-464; parameters!(qnet,net), src/snaq_optimization.jl:106-180)
against the reference's golden vectors and the oracle: the structural rule of snaq.jl_b200/csrc/snaq_hasedge.h (networks
with at most one hybrid node) and the simulation of extractQuartet! of csrc/snaq_hesim.h (any level-1 network: nested
cycles, bad diamonds and triangles, any pruning order)."""
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


@pytest.mark.parametrize("case", ["G", "F", "I"])
def test_golden_hasedge_through_the_simulation(oracle, case):
    flat = nets()[case].flatten(kat.TAXA5)
    he, bd1, zp, (nh, nt) = hs.hasedge_sim(flat, 5, kat.q5_ids())
    for i in range(5):
        assert list(he[i]) == kat.HASEDGE[case][i], (case, i)
        assert hs.indexht_from(he[i], bd1[i], zp[i], nh, nt) == kat.INDEXHT[case][i], (case, i)


@pytest.mark.parametrize("block", range(4))
def test_simulated_hasedge_matches_the_oracle_with_nested_cycles(oracle, block):
    """h = 0..5 hybrid nodes, 8-21 taxa, bad diamonds / triangles allowed, net.leaf (the pruning order) shuffled:
    hasEdge and indexht bit-exact on every quartet"""
    nq = 0
    for seed in range(block * 40, block * 40 + 40):
        n = 8 + seed % 14
        h = seed % 6
        net = None
        for attempt in range(20):
            try:
                net = simulate.random_level1_network(n, h, 7000 + seed + 1000 * attempt, min_cycle=3 if seed % 3 == 0 else 4)
                break
            except S.SnaqError:
                if attempt % 4 == 3 and h > 0:
                    h -= 1
        taxa = [f"t{i + 1}" for i in range(n)]
        rng = np.random.default_rng(seed)
        rng.shuffle(net.leaf)
        flat = net.flatten(taxa)
        qt = simulate.all_quartets(n)
        o = oracle.extract(flat, qt)
        he, bd1, zp, (nh, nt) = hs.hasedge_sim(flat, n, qt)
        assert np.array_equal(he, o["hasedge"]), seed
        for i in range(0, len(qt), 3):
            assert hs.indexht_from(he[i], bd1[i], zp[i], nh, nt) == list(o["indexht"][i][: o["nindexht"][i]]), (seed, i)
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
def test_hasedge_through_the_abi_matches_the_oracle(oracle, monkeypatch):
    """any level-1 network through snaq_get_hasedge: h <= 1 by the structural rule AND by the simulation (SNAQ_HASEDGE_SIM=1),
    h = 2..5 (nested cycles) by the simulation; shuffled pruning order"""
    total = 0
    for seed in (3, 8, 13, 21, 30, 31, 32, 33, 34, 35, 36, 37):
        n, h = 10 + seed % 9, (seed % 2 if seed < 30 else 2 + seed % 4)
        net = None
        for attempt in range(20):
            try:
                net = simulate.random_level1_network(n, h, 100 + seed + 1000 * attempt, min_cycle=3 if seed % 3 == 0 else 4)
                break
            except S.SnaqError:
                if attempt % 4 == 3 and h > 0:
                    h -= 1
        taxa = [f"t{i + 1}" for i in range(n)]
        rng = np.random.default_rng(seed)
        rng.shuffle(net.leaf)
        flat = net.flatten(taxa)
        qt = simulate.all_quartets(n)
        o = oracle.extract(flat, qt)
        for force_sim in ("0", "1"):
            monkeypatch.setenv("SNAQ_HASEDGE_SIM", force_sim)
            eng = S.Engine(0)
            eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
            eng.set_network(flat)
            he, ix = eng.hasedge()
            assert np.array_equal(he, o["hasedge"]), (seed, force_sim)
            assert all(ix[i] == list(o["indexht"][i][: o["nindexht"][i]]) for i in range(len(qt))), (seed, force_sim)
        total += len(qt)
    assert total > 10000


@pytest.mark.gpu
def test_hasedge_at_100_taxa_on_sampled_quartets(oracle):
    """configs[2] size (h = 3, nested blocks): the simulation on the device against the oracle on 3,000 sampled quartets"""
    n = 100
    net = simulate.random_level1_network(n, 3, 20261021)
    taxa = [f"t{i + 1}" for i in range(n)]
    rng = np.random.default_rng(2)
    qt = np.sort(rng.choice(n, (3000, 4)), axis=1)
    qt = qt[(np.diff(qt, axis=1) > 0).all(axis=1)].astype(np.uint16)
    flat = net.flatten(taxa)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    eng.set_network(flat)
    he, ix = eng.hasedge()
    o = oracle.extract(flat, qt)
    assert np.array_equal(he, o["hasedge"])
    assert all(ix[i] == list(o["indexht"][i][: o["nindexht"][i]]) for i in range(len(qt)))


@pytest.mark.gpu
def test_quartet_edges_for_the_weighted_edge_choice(oracle):
    """snaq_get_quartet_edges = [e.number for e in d.quartet[idx].qnet.edge] (src/moves_snaq.jl:1451-1453)"""
    for seed, h in ((5, 0), (6, 1), (7, 3), (9, 4)):
        n = 14
        net = None
        for attempt in range(20):
            try:
                net = simulate.random_level1_network(n, h, 300 + seed + 1000 * attempt, min_cycle=3 if seed % 2 else 4)
                break
            except S.SnaqError:
                pass
        taxa = [f"t{i + 1}" for i in range(n)]
        flat = net.flatten(taxa)
        qt = simulate.all_quartets(n)
        eng = S.Engine(0)
        eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
        eng.set_network(flat)
        for q in range(0, len(qt), 37):
            assert eng.quartet_edges(q) == oracle.quartet_edges(flat, qt[q]), (seed, q)
