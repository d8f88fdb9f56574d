"""updateBL! (src/readquartetdata.jl:838-861): the oracle restatement on the CPU tier; the device reduction
(snaq_update_bl) against it on the GPU tier."""
import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate
from oracle import update_bl as ubl


def tree_and_data(n, seed, drop=0.0):
    net = simulate.random_level1_network(n, 0, seed)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    rng = np.random.default_rng(seed)
    if drop > 0:                                          # a data table that misses some quartets
        qt = qt[rng.uniform(size=len(qt)) >= drop]
    # shuffle the taxon order inside each quartet: the resolution column depends on it
    qt = np.array([rng.permutation(q) for q in qt], dtype=np.uint16)
    return net, taxa, qt


def test_oracle_update_bl_inverts_the_tree_formula(oracle):
    """On exact CFs of a tree the mean CF of an edge's quartets is 1 - 2/3 exp(-t): updateBL! returns t."""
    net, taxa, qt = tree_and_data(9, 3)
    flat = net.flatten(taxa)
    e = oracle.extract(flat, qt)["expcf"]
    for ed in net.edge:                                    # readnewicklevel1 marks missing lengths with 1.0 / -1.0
        pass
    tab = ubl.update_bl_table(flat, qt, e)
    internal = [ed for ed in net.edge if not ed.node[0].leaf and not ed.node[1].leaf]
    assert sorted(tab) == sorted(ed.number for ed in internal)
    for ed in internal:
        assert abs(tab[ed.number][1] - min(ed.length, 10.0)) < 1e-10
    # Nquartets = product of the sizes of the four parts
    assert sum(v[0] for v in tab.values()) <= len(qt) * len(internal)
    lengths, _ = ubl.update_bl(flat, qt, e)
    assert np.allclose(lengths, flat.edge_length)          # nothing is < 0 or == 1.0 among the internal edges here


def test_oracle_update_bl_all_concordant_gives_an_infinite_length():
    """every quartet of an edge fully concordant: -log(3/2 * 0) is Inf in the reference (:846), not an error"""
    net, taxa, qt = tree_and_data(5, 1)
    flat = net.flatten(taxa)
    parts = dict(ubl._parts(flat))
    enum, ps = next(iter(parts.items()))
    obs = np.full((len(qt), 3), 1 / 3)
    for i, q in enumerate(qt):                             # the split of each quartet that edge `enum` induces gets CF 1
        d = [int(t) for t in q]
        side = [next(k for k, p in enumerate(ps) if t in p) // 2 for t in d]
        if sorted(side) == [0, 0, 1, 1] and len({next(k for k, p in enumerate(ps) if t in p) for t in d}) == 4:
            mate = next(j for j in (1, 2, 3) if side[j] == side[0])
            obs[i] = 0.0
            obs[i, mate - 1] = 1.0
    tab = ubl.update_bl_table(flat, qt, obs)
    assert tab[enum][1] == np.inf
    lengths, _ = ubl.update_bl(flat, qt, obs)
    assert np.all(lengths >= 0)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(5))
def test_device_update_bl_matches_the_oracle(oracle, seed):
    n = 8 + 3 * seed
    net, taxa, qt = tree_and_data(n, 20 + seed, drop=0.3 if seed % 2 else 0.0)
    flat = net.flatten(taxa)
    e = oracle.extract(flat, qt)["expcf"]
    obs = simulate.obs_from_expcf(e, 40, seed)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, obs))
    eng.set_network(flat)
    nq, el = eng.update_bl()
    _, numht, _ = eng.get_params()
    tab = ubl.update_bl_table(flat, qt, obs)
    got = {int(numht[eng.nh + i]): (int(nq[i]), float(el[i])) for i in range(eng.nt) if nq[i] > 0}
    assert sorted(got) == sorted(tab)
    for k in tab:
        assert got[k][0] == tab[k][0]
        assert got[k][1] == pytest.approx(tab[k][1], rel=1e-9, abs=1e-9) or (np.isinf(tab[k][1]) and np.isinf(got[k][1]))


@pytest.mark.gpu
def test_updateBL_host_mirror_applies_the_reference_rule(oracle):
    net, taxa, qt = tree_and_data(12, 77)
    flat = net.flatten(taxa)
    obs = simulate.obs_from_expcf(oracle.extract(flat, qt)["expcf"], 100, 1)
    internal = [ed for ed in net.edge if not ed.node[0].leaf and not ed.node[1].leaf]
    internal[0].length = 1.0                               # "missing": replaced
    internal[1].length = -1.0                              # negative: replaced
    keep = internal[2].length
    flat = net.flatten(taxa)
    want, tab = ubl.update_bl(flat, qt, obs)
    d = S.DataCF(taxa, qt, obs)
    got_tab = S.updateBL_(net, d)
    assert {k: v[0] for k, v in got_tab.items()} == {k: v[0] for k, v in tab.items()}
    assert np.allclose([ed.length for ed in net.edge], want, rtol=1e-9, atol=1e-9)
    assert internal[2].length == keep
    hyb = simulate.random_level1_network(10, 1, 5)
    with pytest.raises(S.SnaqError):
        S.updateBL_(hyb, S.DataCF([f"t{i + 1}" for i in range(10)], simulate.all_quartets(10), np.full((210, 3), 1 / 3)))
