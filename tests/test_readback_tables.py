"""fittedquartetCF / writeExpCF (src/descriptive.jl:24-59, src/readquartetdata.jl:4-14): the tables either side of the path.
The expected CFs come from the oracle here (CPU tier) and from the engine in the GPU test; a table written by writeExpCF is
read back by readtableCF unchanged."""
import os

import numpy as np
import pytest

import snaq_jl_b200 as S
from tests import kat

HERE = os.path.dirname(os.path.abspath(__file__))


def fitted(oracle):
    d = S.readtableCF(os.path.join(HERE, "golden", "tableCF.txt"))
    net = S.readnewicklevel1(kat.NET1)
    flat = net.flatten(d.taxa)
    d.expCF = oracle.extract(flat, d.quartet_taxa)["expcf"]
    return d


def test_fittedquartetcf_wide_and_long(oracle):
    d = fitted(oracle)
    w = S.fittedquartetCF(d)
    assert list(w) == ["tx1", "tx2", "tx3", "tx4", "obsCF12", "obsCF13", "obsCF14", "expCF12", "expCF13", "expCF14"]
    assert len(w["tx1"]) == d.numQuartets == 15
    assert [w[f"tx{k + 1}"][3] for k in range(4)] == d.quartet_names(3)
    np.testing.assert_array_equal(w["obsCF13"], d.obsCF[:, 1])
    np.testing.assert_array_equal(w["expCF14"], d.expCF[:, 2])
    lg = S.fittedquartetCF(d, "long")
    assert list(lg) == ["tx1", "tx2", "tx3", "tx4", "quartet", "obsCF", "expCF"]
    assert len(lg["quartet"]) == 45 and lg["quartet"][:3] == ["12_34", "13_24", "14_23"]
    assert lg["tx2"][3:6] == [d.quartet_names(1)[1]] * 3
    np.testing.assert_array_equal(lg["expCF"][6:9], d.expCF[2])
    with pytest.raises(S.SnaqError):
        S.fittedquartetCF(d, "tall")
    d.expCF = None
    with pytest.raises(S.SnaqError):
        S.fittedquartetCF(d)


def test_writeexpcf_round_trip(oracle, tmp_path):
    d = fitted(oracle)
    path = str(tmp_path / "expCF.csv")
    cols = S.writeExpCF(d, path)
    assert list(cols) == ["t1", "t2", "t3", "t4", "CF12_34", "CF13_24", "CF14_23"]
    back = S.readtableCF(path)
    assert back.taxa == d.taxa
    np.testing.assert_array_equal(back.quartet_taxa, d.quartet_taxa)
    np.testing.assert_array_equal(back.obsCF, d.expCF)          # repr() round-trips doubles exactly


@pytest.mark.gpu
def test_fitted_table_from_the_engine(oracle):
    d = S.readtableCF(os.path.join(HERE, "golden", "tableCF.txt"))
    net = S.readnewicklevel1(kat.NET1)
    assert abs(S.topologyQpseudolik_(net, d) - kat.NET1_LIK) <= 1e-10 * kat.NET1_LIK
    w = S.fittedquartetCF(d)
    want = oracle.extract(net.flatten(d.taxa), d.quartet_taxa)["expcf"]
    got = np.stack([w["expCF12"], w["expCF13"], w["expCF14"]], axis=1)
    assert np.max(np.abs(got - want)) <= 1e-13
