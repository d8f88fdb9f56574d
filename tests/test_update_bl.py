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


# ---- a known-answer vector worked out by hand on the reference's own example data -------------------------------------
# examples/tableCF.txt (15 quartets of A,B,C,D,E,O) and the ASTRAL tree of examples/astral.tre (docs/src/man/snaq_est.md:197-206),
# (E,(O,((C,D),(A,B)))).  edgesParts (src/readquartetdata.jl:866-889) gives an internal edge the four subtrees around it, and
# makeTable (:935-960) the quartets with ONE taxon in each, with the observed CF of the resolution the tree displays:
#   edge {A,B} | rest : parts {A},{B},{C,D},{E,O} -> ABCE, ABCO, ABDE, ABDO, CF12_34 = 0.8333.., 0.8666.., 0.8333.., 0.8666..
#                       mean 0.85      -> edgeL = -log(3/2 (1 - 0.85)) = -log 0.225
#   edge {C,D} | rest : parts {C},{D},{A,B},{E,O} -> ACDE, ACDO, BCDE, BCDO, the displayed split CD|xy is CF14_23 = 1, 1, 1, 1
#                       mean 1.0       -> edgeL = -log 0 = Inf (the reference does not guard it, :846)
#   edge {A,B,C,D} | {E,O} : parts {A,B},{C,D},{E},{O} -> ACEO, ADEO, BCEO, BDEO, CF12_34 = 0.5333.., 0.5333.., 0.6666.., 0.6666..
#                       mean 0.6       -> edgeL = -log(3/2 * 0.4) = -log 0.6
HAND = {frozenset("AB"): (4, -np.log(0.225)), frozenset("CD"): (4, np.inf), frozenset("EO"): (4, -np.log(0.6))}


def _hand_case():
    import os
    d = S.readtableCF(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tableCF.txt"))
    net = S.readnewicklevel1("(E,(O,((C,D),(A,B))));")
    names = {n.number: n.name for n in net.node if n.leaf}

    def side(edge):                                       # the smaller side of the bipartition an internal edge induces
        seen, stack = {edge.node[0].number}, [edge.node[0]]
        while stack:
            u = stack.pop()
            for e in u.edge:
                if e is edge:
                    continue
                v = e.node[1] if e.node[0] is u else e.node[0]
                if v.number not in seen:
                    seen.add(v.number)
                    stack.append(v)
        s = frozenset(names[k] for k in seen if k in names)
        return s if len(s) <= 3 and s in HAND else frozenset(names.values()) - s

    internal = {e.number: side(e) for e in net.edge if not e.node[0].leaf and not e.node[1].leaf}
    assert sorted(map(sorted, internal.values())) == sorted(map(sorted, HAND))
    return d, net, internal


def test_update_bl_restatement_reproduces_the_hand_computed_vector():
    d, net, internal = _hand_case()
    tab = ubl.update_bl_table(net.flatten(d.taxa), d.quartet_taxa, d.obsCF)
    assert sorted(tab) == sorted(internal)
    for num, s in internal.items():
        n, el = HAND[s]
        assert tab[num][0] == n
        assert (np.isinf(el) and np.isinf(tab[num][1])) or abs(tab[num][1] - el) < 1e-12


@pytest.mark.gpu
def test_device_update_bl_reproduces_the_hand_computed_vector():
    d, net, internal = _hand_case()
    eng = S.Engine(0)
    eng.set_data(d)
    eng.set_network(net.flatten(d.taxa))
    nq, el = eng.update_bl()
    _, numht, _ = eng.get_params()
    got = {int(numht[eng.nh + i]): (int(nq[i]), float(el[i])) for i in range(eng.nt)}
    for num, s in internal.items():
        n, want = HAND[s]
        assert got[num][0] == n
        assert (np.isinf(want) and np.isinf(got[num][1])) or abs(got[num][1] - want) < 1e-9
    # updateBL!'s rule (:850-858): lengths that readnewicklevel1 left at 1.0 are replaced, Inf is capped by setLength! at 10
    S.updateBL_(net, d)
    by = {s: next(e for e in net.edge if e.number == num).length for num, s in internal.items()}
    assert abs(by[frozenset("AB")] + np.log(0.225)) < 1e-9 and by[frozenset("CD")] == 10.0 and abs(by[frozenset("EO")] + np.log(0.6)) < 1e-9
