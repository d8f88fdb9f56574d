"""Bootstrap tables on the device: snaq_set_ci / snaq_resample_obscf restate sampleCFfromCI! (src/bootstrap.jl:52-70)
with a counter-based generator documented in include/snaq_b200.h (Julia's RNG stream cannot be reproduced); the same
generator in numpy gives the same table bit for bit, and the objective follows the new table."""
import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate

pytestmark = pytest.mark.gpu

M64 = (1 << 64) - 1


def splitmix64(z):
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def resample_host(lo, hi, seed):
    """c_i = (hi_i - lo_i) * U_i + lo_i; obs_i = c_i / (c_1 + c_2 + c_3)   (src/bootstrap.jl:60-67)"""
    nq = len(lo)
    u = np.empty((nq, 3))
    for q in range(nq):
        for i in range(3):
            ctr = 3 * q + i
            u[q, i] = (splitmix64((seed + ctr * 0x9E3779B97F4A7C15) & M64) >> 11) * 2.0 ** -53
    c = (hi - lo) * u + lo
    return c / ((c[:, 0] + c[:, 1]) + c[:, 2])[:, None]


def setup(n=9, h=1, seed=3):
    net = simulate.random_level1_network(n, h, seed)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    flat = net.flatten(taxa)
    eng.set_network(flat)
    x, _, _ = eng.get_params()
    obs = simulate.obs_from_expcf(eng.eval_quartets(x)[0], 50, seed)
    eng.set_obscf(obs)
    return eng, flat, taxa, qt, obs, x


def test_resampled_table_equals_the_host_restatement(oracle):
    eng, flat, taxa, qt, obs, x = setup()
    lo, hi = np.clip(obs - 0.05, 0, 1), np.clip(obs + 0.05, 0, 1)
    assert np.array_equal(eng.get_obscf(), obs)
    eng.set_ci(lo, hi)
    for seed in (0, 1, 20261019, 2 ** 63 + 12345):
        eng.resample_obscf(seed)
        got = eng.get_obscf()
        want = resample_host(lo, hi, seed)
        assert np.array_equal(got, want)
        assert np.all(got >= 0) and np.allclose(got.sum(1), 1.0, atol=1e-15)
        f = eng.eval(x)[0]                                  # the weights follow the new table
        fo = oracle.evaluate(flat, qt, want, x[None])[0]
        assert abs(f - fo) <= 1e-10 * abs(fo)
    eng.resample_obscf(7)
    a = eng.get_obscf()
    eng.resample_obscf(8)
    assert not np.array_equal(a, eng.get_obscf())
    eng.resample_obscf(7)
    assert np.array_equal(a, eng.get_obscf())               # a replicate is a function of its seed


def test_resampling_stays_inside_the_intervals_and_is_uniform():
    eng, flat, taxa, qt, obs, x = setup(n=14, h=0, seed=5)
    lo = np.tile([0.2, 0.1, 0.3], (len(qt), 1))
    hi = np.tile([0.4, 0.1, 0.5], (len(qt), 1))            # second CF: degenerate interval
    eng.set_ci(lo, hi)
    eng.resample_obscf(42)
    got = eng.get_obscf()
    c = got / got[:, 1:2] * 0.1                             # undo the normalisation through the fixed component
    assert np.all(c[:, 0] >= 0.2 - 1e-12) and np.all(c[:, 0] <= 0.4 + 1e-12)
    assert np.all(c[:, 2] >= 0.3 - 1e-12) and np.all(c[:, 2] <= 0.5 + 1e-12)
    u = (c[:, 0] - 0.2) / 0.2
    assert abs(u.mean() - 0.5) < 4 / np.sqrt(12 * len(u)) and abs(u.var() - 1 / 12) < 0.01
    assert abs(np.corrcoef(c[:, 0], c[:, 2])[0, 1]) < 0.1


def test_ci_arguments_and_worker_contexts():
    eng, flat, taxa, qt, obs, x = setup()
    with pytest.raises(S.SnaqError):
        eng.resample_obscf(1)                               # no intervals yet
    with pytest.raises(S.SnaqError):
        eng.set_ci(obs + 0.1, obs)                          # lo > hi
    with pytest.raises(S.SnaqError):
        eng.set_ci(np.zeros_like(obs), np.zeros_like(obs))  # nothing to normalise
    eng.set_ci(np.clip(obs - 0.05, 0, 1), np.clip(obs + 0.05, 0, 1))
    eng.resample_obscf(3)
    # the candidate pool sees the replicate's table
    got = eng.optbl_topologies([flat, flat], workers=2)
    ref = S.Engine(0)
    ref.set_data(S.DataCF(taxa, qt, eng.get_obscf()))
    ref.set_network(flat)
    xw, iw = ref.optbl(ref.get_params()[0])
    for xm, info in got:
        assert np.array_equal(xm, xw) and info == iw
