"""snaq_sample_quartets: step 1 of sampleEdgeQuartetWeighted (src/moves_snaq.jl:1443-1449) on the device, against the
serial inverse-CDF walk of StatsBase.sample(Weights(w)) restated in numpy over the deltaCF read-back."""
import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate

pytestmark = pytest.mark.gpu


def serial_walk(w, u):
    """t = rand()*sum(w); i = 1; cw = w[1]; while cw < t && i < n: i += 1; cw += w[i]  (0-based result)"""
    cum = np.cumsum(w)                                     # numpy's cumsum is the serial running sum
    t = u * w.sum()
    return np.minimum(np.searchsorted(cum, t, side="left"), len(w) - 1), cum, t


def setup(n, h, seed, ngenes=30):
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    net = simulate.random_level1_network(n, h, seed)
    flat = net.flatten(taxa)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    eng.set_network(flat)
    x, _, up = eng.get_params()
    eng.set_obscf(simulate.obs_from_expcf(eng.eval_quartets(x)[0], ngenes, seed))
    return eng, x, up, len(qt)


@pytest.mark.parametrize("n,h,seed", [(8, 1, 1), (14, 1, 2), (25, 2, 3), (40, 2, 4)])
def test_sample_equals_the_serial_walk(n, h, seed):
    eng, x, up, nq = setup(n, h, seed)
    rng = np.random.default_rng(seed)
    u = np.concatenate([[0.0, 0.5, 1.0 - 2.0 ** -53], rng.uniform(size=500)])
    w = eng.eval_quartets(x)[2]
    got = eng.sample_quartets(x, u)
    want, cum, t = serial_walk(w, u)
    assert got.min() >= 0 and got.max() < nq
    bad = got != want
    # a draw may differ from the serial walk only when u*sum is within rounding of a boundary of the running sum
    for k in np.nonzero(bad)[0]:
        lo, hi = sorted((got[k], want[k]))
        assert abs(cum[lo] - t[k]) <= 1e-12 * cum[-1] and not w[lo + 1:hi].any(), (k, got[k], want[k])
    assert bad.sum() <= 2
    assert np.all(w[got[3:]] > 0)                           # zero weights are never drawn
    assert np.array_equal(got, eng.sample_quartets(x, u))   # reproducible


def test_sample_follows_the_weights_and_the_mask():
    eng, x, up, nq = setup(20, 1, 7)
    rng = np.random.default_rng(0)
    mask = (rng.uniform(size=nq) < 0.3).astype(np.uint8)
    eng.set_sampled(mask)
    w = eng.eval_quartets(x)[2]
    assert not w[mask == 0].any()                          # unsampled quartets have weight 0 (src/pseudolik.jl:513-515)
    got = eng.sample_quartets(x, rng.uniform(size=20000))
    assert mask[got].all()
    # the ten heaviest quartets are drawn in proportion to their weights
    top = np.argsort(w)[-10:]
    freq = np.array([(got == i).mean() for i in top])
    assert np.allclose(freq, w[top] / w.sum(), atol=4 * np.sqrt(w[top].max() / w.sum() / 20000))
    # the weights follow x: at other parameters the same draws give other quartets
    x2 = np.clip(x * 0.3, 0, up)
    assert not np.array_equal(got[:200], eng.sample_quartets(x2, rng.uniform(size=200)))


def test_sample_arguments():
    eng, x, up, nq = setup(8, 0, 9)
    assert eng.sample_quartets(x, []).shape == (0,)
    with pytest.raises(S.SnaqError):
        eng.sample_quartets(x, [1.0])
    with pytest.raises(S.SnaqError):
        eng.sample_quartets(x[:-1], [0.5])
    # perfect fit: every weight is 0, the walk returns the first quartet like the reference's (t = 0)
    eng.set_obscf(eng.eval_quartets(x)[0])
    assert list(eng.sample_quartets(x, [0.0, 0.7])) == [0, 0]
