"""Pins the CPU oracle (oracle/snaq_oracle.cpp) against every golden vector the
reference's own tests and examples hold for the hot path."""
import itertools
import os

import numpy as np
import pytest

import snaq_jl_b200 as S
from tests import kat

HERE = os.path.dirname(os.path.abspath(__file__))


def nets():
    return {"G": S.readnewicklevel1(kat.CASE_G), "F": kat.case_f_network(S), "I": S.readnewicklevel1(kat.CASE_I)}


@pytest.mark.parametrize("case", ["G", "F", "I"])
def test_parameters_match_reference(oracle, case):
    # test/test_hasEdge.jl:28-29,72-73,113-114
    net = nets()[case]
    p = oracle.parameters(net.flatten(kat.TAXA5))
    assert list(p["numht"]) == kat.NUMHT[case]
    np.testing.assert_allclose(p["x"], kat.HT[case], rtol=0, atol=1e-15)
    assert list(net.numht) == kat.NUMHT[case]


def test_parameters_5taxon_newick_cases(oracle):
    # test/test_parameters.jl:40-65 (numht and ht for the newick-read 5-taxon networks)
    from math import exp
    cases = {
        "F": ("(((6:0.1,(4)11#H1)1:0.2,(11#H1,7))5:0.1,8:0.1,10:0.1);", [8, 31, 32], [0.1, 0.9 * (1 - exp(-.2)), 0.1 * (1 - exp(-1.))]),
        "G": (kat.CASE_G, [7, 3, 6, 9], [0.1, 0.2, 0.1, 1.0]),
        "H": ("((((6,4),#H1),7),(8)#H1,10);", [4, 3, 5, 7], [0.1, 1.0, 1.0, 1.0]),
        "J": ("((((6)#H1,4),7),8,(#H1,10));", [8, 4, 6, 10], [0.1, 1.0, 1.0, 1.0]),
        "I": (kat.CASE_I, [9, 4, 6, 9, 10], [0.1, 2.0, 1.0, 1.0, 1.0]),
    }
    for name, (nw, numht, ht) in cases.items():
        net = S.readnewicklevel1(nw)
        p = oracle.parameters(net.flatten(kat.TAXA5))
        assert list(p["numht"]) == numht, name
        np.testing.assert_allclose(p["x"], ht, rtol=1e-15, atol=0, err_msg=name)
        assert [i + 1 for i in net.index][: len(numht)] == ({"F": [8, 4, 6]}.get(name, numht)), name


@pytest.mark.parametrize("case", ["G", "F", "I"])
def test_hasedge_indexht(oracle, case):
    # test/test_hasEdge.jl:21-41, 65-85, 106-126
    r = oracle.extract(nets()[case].flatten(kat.TAXA5), kat.q5_ids())
    for i in range(5):
        assert list(r["hasedge"][i]) == kat.HASEDGE[case][i], (case, i)
        assert list(r["indexht"][i][: r["nindexht"][i]]) == kat.INDEXHT[case][i], (case, i)


@pytest.mark.parametrize("case", ["G", "F", "I"])
def test_calculate_expcf_stages(oracle, case):
    # test/test_calculateExpCF.jl: which, typeHyb, t1, formula, expCF (exact comparisons there)
    r = oracle.extract(nets()[case].flatten(kat.TAXA5), kat.q5_ids())
    for i, (which, types, t1, formula, expcf) in enumerate(kat.EXPCF[case]):
        assert r["which"][i] == which, (case, i)
        assert list(r["types"][i][: r["ntype"][i]]) == types, (case, i)
        if which == 1:
            assert list(r["formula"][i]) == formula, (case, i)
            # exact (!=) in the reference for cases G and F, approxEq for case I
            assert abs(r["t1"][i] - t1) <= (0 if case != "I" else 4e-16) * abs(t1) + (1e-16 if t1 == 0.3 else 0), (case, i, r["t1"][i], t1)
            # the reference asserts expCF == f(qnet.t1) exactly
            assert list(r["expcf"][i]) == kat._tree(r["t1"][i], formula), (case, i)
        else:
            assert r["t1"][i] == -1.0
            assert list(r["expcf"][i]) == expcf, (case, i)


def test_optbl_parts_closed_forms(oracle):
    # test/test_optBLparts.jl:139-147, 254-259, 375-383: expCF as functions of x
    n = nets()
    for case, x, fn in (("G", [0.3, 0.2, 0.1, 2.0], kat.optbl_G), ("F", [0.3, 0.2, 0.1], kat.optbl_F),
                        ("I", [0.2, 0.2, 0.1, 0.1, 0.1], kat.optbl_I)):
        flat = n[case].flatten(kat.TAXA5)
        r = oracle.extract(flat, kat.q5_ids(), x=x)
        np.testing.assert_allclose(r["expcf"], fn(x), rtol=1e-13, atol=0, err_msg=case)
        # same through the optBL! objective path (extract once, update!(qnet,x,net), recompute)
        obs = np.tile([0.5, 0.4, 0.1], (5, 1))
        _, ex = oracle.evaluate(flat, kat.q5_ids(), obs, x, want_expcf=True)
        np.testing.assert_allclose(ex, fn(x), rtol=1e-13, atol=0, err_msg=case)


def test_lik_on_tree(oracle):
    # test/test_correctLik.jl:33-45
    net = S.readnewicklevel1(kat.TREE5)
    flat = net.flatten(kat.TAXA5)
    r = oracle.extract(flat, kat.q5_ids())
    for i in range(1, 5):  # rows 2:5 in the reference test (row 1 spans two internal edges)
        srt = np.sort(r["expcf"][i])
        np.testing.assert_allclose(srt, [kat.EXPCF_TREE5_MINOR, kat.EXPCF_TREE5_MINOR, kat.EXPCF_TREE5_MAJOR], rtol=1e-14)
    lik = oracle.evaluate(flat, kat.q5_ids(), kat.OBS_TREE5, oracle.parameters(flat)["x"])
    assert abs(lik[0] - kat.LIK_TREE5) <= 1e-12 * kat.LIK_TREE5


def test_lik_of_network_case_g(oracle):
    # test/test_correctLik.jl:51-64
    net = S.readnewicklevel1(kat.CASE_G)
    flat = net.flatten(kat.TAXA5)
    lik = oracle.evaluate(flat, kat.q5_ids(), kat.OBS_CASEH, oracle.parameters(flat)["x"])
    assert abs(lik[0] - kat.LIK_CASEG_ON_H) <= 1e-13 * kat.LIK_CASEG_ON_H


@pytest.mark.parametrize("newick,lik", [(kat.NET0, kat.NET0_LIK), (kat.NET1, kat.NET1_LIK)])
def test_example_outputs(oracle, newick, lik):
    # examples/net0.out:1, examples/net1.out:1 on examples/tableCF.txt
    d = S.readtableCF(os.path.join(HERE, "golden", "tableCF.txt"))
    assert d.numQuartets == 15 and d.taxa == ["A", "B", "C", "D", "E", "O"]
    net = S.readnewicklevel1(newick)
    flat = net.flatten(d.taxa)
    val = oracle.evaluate(flat, d.quartet_taxa, d.obsCF, oracle.parameters(flat)["x"])[0]
    assert abs(val - lik) <= 2e-14 * lik, (val, lik)


def _more_examples():
    cases = [(nw, lik, "tableCF.txt") for nw, lik in kat.NET1_NETWORKS]
    cases += [(kat.NET2_OUT[0], kat.NET2_OUT[1], "tableCF.txt"), (kat.NET3_OUT[0], kat.NET3_OUT[1], "tableCF.txt"),
              (kat.NET1_SNAQ_OUT[0], kat.NET1_SNAQ_OUT[1], "tableCFCI.csv")]
    return cases


@pytest.mark.parametrize("i", range(8))
def test_more_example_outputs(oracle, i):
    # examples/net1.networks (5 scored topologies), net2.out:1, net3.out:1, net1_snaq.out:1 (a second data set, with zeros)
    newick, lik, table = _more_examples()[i]
    d = S.readtableCF(os.path.join(HERE, "golden", table))
    net = S.readnewicklevel1(newick)
    flat = net.flatten(d.taxa)
    val = oracle.evaluate(flat, d.quartet_taxa, d.obsCF, oracle.parameters(flat)["x"])[0]
    assert abs(val - lik) <= 2e-14 * lik, (val, lik)
    from tests import hostsim_py as hs                          # the engine's device functions, compiled for the host
    r = hs.run(flat, len(d.taxa), d.quartet_taxa, obs=d.obsCF)
    assert abs(r["lik"][0] - lik) <= 1e-13 * lik


def test_four_cycle_expcf(oracle):
    # test/test_multipleAlleles.jl:66-72: every ordering of the 4 taxa of a 4-taxon network
    net = S.readnewicklevel1(kat.NET_4CYCLE)
    taxa = ["a", "b", "c", "d"]
    flat = net.flatten(taxa)
    r = oracle.extract(flat, np.array([[0, 1, 2, 3]], dtype=np.uint16))
    np.testing.assert_allclose(r["expcf"][0], kat.EXPCF_4CYCLE, rtol=3e-16)
    # all 24 orderings permute the same three values consistently
    perms = np.array(list(itertools.permutations(range(4))), dtype=np.uint16)
    rp = oracle.extract(flat, perms)
    base = {frozenset({frozenset({0, 1}), frozenset({2, 3})}): kat.EXPCF_4CYCLE[0],
            frozenset({frozenset({0, 2}), frozenset({1, 3})}): kat.EXPCF_4CYCLE[1],
            frozenset({frozenset({0, 3}), frozenset({1, 2})}): kat.EXPCF_4CYCLE[2]}
    for p, e in zip(perms, rp["expcf"]):
        for j, (a, b) in enumerate([((0, 1), (2, 3)), ((0, 2), (1, 3)), ((0, 3), (1, 2))]):
            key = frozenset({frozenset({int(p[a[0]]), int(p[a[1]])}), frozenset({int(p[b[0]]), int(p[b[1]])})})
            assert abs(e[j] - base[key]) < 1e-15


def test_negative_expcf_bad_diamond(oracle):
    # test/test_calculateExpCF2.jl:7-35: gammaz 1.0 and 0.067 => second expCF negative, which stays 2
    net = S.readnewicklevel1(kat.NET_NEG)
    assert net.hybrid and any(h.isBadDiamondI for h in net.hybrid)
    bd = [h for h in net.hybrid if h.isBadDiamondI][0]
    maj, mino, _ = S.hybridEdges(bd)
    taxa = ["1", "2", "3", "4", "5", "6"]
    flat = net.flatten(taxa)
    p = oracle.parameters(flat)
    x = p["x"].copy()
    assert p["nhz"] == 2
    # node -5 / -7 of the reference are the minor / major side nodes: one of the two assignments
    neg = []
    for z in ([1.0, 0.067], [0.067, 1.0]):
        x[-2:] = z
        r = oracle.extract(flat, np.array([[taxa.index(t) for t in ["6", "5", "4", "1"]]], dtype=np.uint16), x=x)
        assert r["which"][0] == 2 and list(r["types"][0][: r["ntype"][0]]) == [5]
        neg.append(r["expcf"][0])
        assert min(r["expcf"][0]) < 0
    assert any(e[1] < 0 for e in neg)
    # logPseudoLik gets the -1e15 penalty per negative expCF (src/pseudolik.jl:1394-1396)
    val = oracle.evaluate(flat, np.array([[taxa.index(t) for t in ["6", "5", "4", "1"]]], dtype=np.uint16), [[0.5, 0.4, 0.1]], x)
    assert val[0] > 9e14


def test_perfect_data_15_taxa(oracle):
    # test/test_perfectData.jl:1372-1382: expCF fed back as data => pseudo-deviance 0
    net = S.readnewicklevel1(kat.NET_PERFECT)
    assert net.numhybrids == 3
    taxa = [str(i) for i in range(1, 16)]
    flat = net.flatten(taxa)
    qt = np.array(list(itertools.combinations(range(15), 4)), dtype=np.uint16)
    assert len(qt) == 1365
    r = oracle.extract(flat, qt)
    np.testing.assert_allclose(r["expcf"].sum(axis=1), 1.0, atol=1e-14)
    x = oracle.parameters(flat)["x"]
    val = oracle.evaluate(flat, qt, r["expcf"], x)[0]
    assert abs(val) <= 1e-12
    # and the stateless path (apply x to the net, then extract) agrees with the optBL! path
    _, ex = oracle.evaluate(flat, qt, r["expcf"], x, want_expcf=True)
    np.testing.assert_allclose(ex, r["expcf"], rtol=1e-14)
