"""snaq_patch_network: after an in-place edit of the network (nearest-neighbour interchanges, as the search proposes them) only
the quartets with a taxon whose root path changed are classified again, and the device table is bit-identical to a
build from scratch with the same root; against snaq_set_network on a fresh context (its own root) the descriptors are
identical and the values agree to rounding."""
import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate

pytestmark = pytest.mark.gpu


def engines(n, h, seed):
    net = simulate.random_level1_network(n, h, seed)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    rng = np.random.default_rng(seed)
    obs = rng.dirichlet([5.0, 2.0, 2.0], len(qt))
    d = S.DataCF(taxa, qt, obs)
    out = []
    for _ in range(2):
        e = S.Engine(0)
        e.set_data(d)
        e.set_network(net.flatten(taxa))
        out.append(e)
    return net, taxa, qt, d, out


@pytest.mark.parametrize("n,h,seed", [(30, 2, 11), (60, 3, 12), (100, 3, 20261021)])
def test_patch_equals_rebuild(n, h, seed):
    net, taxa, qt, d, (ea, eb) = engines(n, h, seed)
    rng = np.random.default_rng(seed + 1)
    partial = 0
    for step in range(6):
        assert simulate.nni_move(net, rng) is not None
        flat = net.flatten(taxa)
        touched = ea.patch_network(flat)                           # incremental
        eb.patch_network(flat, force_all=True)                     # everything re-described, same root
        partial += touched >= 0
        assert ea.table_checksum() == eb.table_checksum(), (step, touched)
        x0, numht, upper = ea.get_params()
        xb, numhtb, _ = eb.get_params()
        assert np.array_equal(x0, xb) and np.array_equal(numht, numhtb)
        X = simulate.parameter_batch(x0, upper, 40, step)
        va, vb = ea.eval(X), eb.eval(X)
        assert np.array_equal(va, vb)                              # batched kernel, bit for bit
        assert np.array_equal(ea.eval(X[:3]), eb.eval(X[:3]))      # single-call kernel
        fa, ga = ea.eval_grad(X[1]); fb, gb = eb.eval_grad(X[1])
        assert fa == fb and np.array_equal(ga, gb)
        # a fresh context chooses its own root: same descriptors, same values to rounding
        ec = S.Engine(0)
        ec.set_data(d)
        ec.set_network(flat)
        vc = ec.eval(X)
        np.testing.assert_allclose(va, vc, rtol=1e-12)
        da, dc = ea.descriptors(), ec.descriptors()
        for k in da:
            assert np.array_equal(da[k], dc[k]), k
        ea_q = ea.eval_quartets(x0)
        ec_q = ec.eval_quartets(x0)
        assert np.max(np.abs(ea_q[0] - ec_q[0])) <= 1e-13
    assert partial >= 3                                            # most interchanges touch a minority of the taxa


def test_patch_falls_back_when_the_layout_changes():
    """adding a hybrid changes the parameter layout: everything is re-described, the result equals snaq_set_network"""
    n = 24
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    d = S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3))
    ea, eb = S.Engine(0), S.Engine(0)
    ea.set_data(d); eb.set_data(d)
    ea.set_network(simulate.random_level1_network(n, 1, 5).flatten(taxa))
    f2 = simulate.random_level1_network(n, 2, 6).flatten(taxa)
    assert ea.patch_network(f2) == -1
    eb.set_network(f2)
    x0, _, upper = ea.get_params()
    X = simulate.parameter_batch(x0, upper, 20, 1)
    np.testing.assert_allclose(ea.eval(X), eb.eval(X), rtol=1e-12)
    # and a patch is refused before any data / accepted as a first build
    ec = S.Engine(0)
    ec.set_data(d)
    assert ec.patch_network(f2) == -1
    np.testing.assert_allclose(ec.eval(X), eb.eval(X), rtol=1e-12)
