"""GPU tier: edge cases of the C ABI — empty and ragged quartet tables, batch sizes that cross every kernel
boundary, exact zeros in obsCF, all quartets unsampled, context reuse across topologies and data sets."""
import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate

pytestmark = pytest.mark.gpu



@pytest.fixture(params=[None, False], ids=["run_table", "per_quartet"], autouse=True)
def table_mode(request, monkeypatch):
    """every test of this module runs on the default run table (distinct programs, summed weights) and on the per-quartet
    kernels (SNAQ_OPT_PER_QUARTET)"""
    monkeypatch.setattr(S.Engine, "DEFAULT_DEDUP", request.param)


def build(n, h, seed, qt=None, ngenes=30):
    net = None
    for attempt in range(30):
        try:
            net = simulate.random_level1_network(n, h, seed + 1000 * attempt)
            break
        except S.SnaqError:
            continue
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n) if qt is None else qt
    return net, taxa, qt


def test_empty_quartet_table():
    net, taxa, _ = build(8, 1, 1)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, np.zeros((0, 4), np.uint16), np.zeros((0, 3))))
    eng.set_network(net.flatten(taxa))
    x0, _, upper = eng.get_params()
    assert np.array_equal(eng.eval(np.tile(x0, (3, 1))), np.zeros(3))        # sum over no quartets (src/pseudolik.jl:1418)
    e, lp, dl = eng.eval_quartets(x0)
    assert e.shape == (0, 3) and lp.shape == (0,)
    f, g = eng.eval_grad(x0)
    assert f == 0.0 and not g.any()
    xmin, res = eng.optbl(x0)
    assert res["fmin"] == 0.0


def test_ragged_table_and_batch_sizes_across_kernel_boundaries(oracle):
    """A table that holds an arbitrary subset of the quartets, in arbitrary taxon order; P = 1..70 crosses the
    thread-per-quartet (1, 2, 4 candidates per launch), lane-per-candidate (32-candidate groups) and chunked paths."""
    n = 13
    rng = np.random.default_rng(5)
    qt = simulate.all_quartets(n)
    qt = qt[rng.uniform(size=len(qt)) < 0.37]
    qt = np.array([rng.permutation(q) for q in qt], dtype=np.uint16)
    net, taxa, _ = build(n, 2, 9)
    flat = net.flatten(taxa)
    obs = simulate.obs_from_expcf(oracle.extract(flat, qt)["expcf"], 7, 3)     # 7 genes: many exact zeros
    assert (obs == 0).any()
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, obs))
    eng.set_network(flat)
    x0, _, upper = eng.get_params()
    X = simulate.parameter_batch(x0, upper, 70, 1)
    want = oracle.evaluate(flat, qt, obs, X)
    for P in (1, 2, 3, 4, 5, 8, 15, 16, 17, 31, 32, 33, 64, 70):
        np.testing.assert_allclose(eng.eval(X[:P]), want[:P], rtol=1e-10)
    e, lp, _ = eng.eval_quartets(X[5])
    np.testing.assert_allclose(-lp.sum(), want[5], rtol=1e-10)
    f, g = eng.eval_grad(X[5])
    assert abs(f - want[5]) <= 1e-10 * abs(want[5])


def test_large_batch_pipeline_equals_small_batches(oracle):
    net, taxa, qt = build(12, 2, 21)
    flat = net.flatten(taxa)
    obs = simulate.obs_from_expcf(oracle.extract(flat, qt)["expcf"], 50, 3)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, obs))
    eng.set_network(flat)
    x0, _, upper = eng.get_params()
    X = simulate.parameter_batch(x0, upper, 2500, 2)              # >= 2048: copy/compute pipeline in snaq_eval
    big = eng.eval(X)
    small = np.concatenate([eng.eval(X[i:i + 500]) for i in range(0, 2500, 500)])
    assert np.array_equal(big, small)                            # batch composition does not change a value
    bad = X.copy()
    bad[2100, 0] = 1.5                                           # a violation deep inside a chunk is still reported
    with pytest.raises(S.SnaqError, match="between 0,1|nonnegative"):
        eng.eval(bad)
    assert np.array_equal(eng.eval(X[:40]), big[:40])            # and the context stays usable


def test_all_unsampled_and_context_reuse(oracle):
    net, taxa, qt = build(11, 1, 33)
    flat = net.flatten(taxa)
    obs = simulate.obs_from_expcf(oracle.extract(flat, qt)["expcf"], 50, 3)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, obs))
    eng.set_network(flat)
    x0, _, upper = eng.get_params()
    eng.set_sampled(np.zeros(len(qt), np.uint8))
    assert eng.eval(x0)[0] == 0.0                                # unsampled quartets contribute 0 (src/pseudolik.jl:1390)
    _, _, dl = eng.eval_quartets(x0)
    assert not dl.any()                                          # deltaCF = 0 for unsampled quartets (:514)
    eng.set_sampled(None)
    v1 = eng.eval(x0)[0]
    # another topology (different number of parameters) on the same context and data
    net2, _, _ = build(11, 3, 34)
    flat2 = net2.flatten(taxa)
    eng.set_network(flat2)
    x2, _, _ = eng.get_params()
    assert len(x2) != len(x0)
    np.testing.assert_allclose(eng.eval(x2)[0], oracle.evaluate(flat2, qt, obs, x2)[0], rtol=1e-10)
    with pytest.raises(S.SnaqError, match="same length|wrong length"):
        eng.eval(x0)
    eng.set_network(flat)                                        # and back: same value as before, bit for bit
    assert eng.eval(x0)[0] == v1
    # new data on the same context: the network must be set again
    n3 = 9
    net3, taxa3, qt3 = build(n3, 1, 35)
    eng.set_data(S.DataCF(taxa3, qt3, np.full((len(qt3), 3), 1 / 3)))
    with pytest.raises(S.SnaqError, match="set_network"):
        eng.eval(x0)
    eng.set_network(net3.flatten(taxa3))
    x3, _, _ = eng.get_params()
    assert np.isfinite(eng.eval(x3)[0])


def test_persistent_batched_kernel_matches_the_default_and_the_oracle(oracle, monkeypatch):
    """k_eval_lanes_p is selected for big xf tiles (>= ~150 taxa); SNAQ_LANES_PERSISTENT=2 forces it on a small network."""
    net, taxa, qt = build(16, 3, 51)
    flat = net.flatten(taxa)
    obs = simulate.obs_from_expcf(oracle.extract(flat, qt)["expcf"], 60, 3)
    x0 = None
    got = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("SNAQ_LANES_PERSISTENT", mode)
        eng = S.Engine(0)
        eng.set_data(S.DataCF(taxa, qt, obs))
        eng.set_network(flat)
        x0, _, upper = eng.get_params()
        X = simulate.parameter_batch(x0, upper, 200, 4)
        got[mode] = eng.eval(X)
        assert np.array_equal(eng.eval(X[:64]), got[mode][:64])          # batch composition does not change a value
        assert np.array_equal(eng.eval(X), got[mode])                    # nor does repeating the launch
    np.testing.assert_allclose(got["2"], got["0"], rtol=1e-13)
    np.testing.assert_allclose(got["2"], oracle.evaluate(flat, qt, obs, X), rtol=1e-10)
