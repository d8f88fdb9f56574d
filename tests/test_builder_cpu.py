"""CPU tier: the engine's device functions (snaq.jl_b200/csrc/snaq_core.h), compiled for the host by the
test harness tests/hostsim, against the oracle: descriptors bit-exact, expCF and logPseudoLik within
tolerance, the specialised exp/log within 2.5e-16.  The same header is what nvcc compiles for the GPU."""
import ctypes as C

import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate
from tests import hostsim_py as hs
from tests import kat


def test_specialised_exp_log_accuracy():
    lib = hs.lib()

    def run(u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        e = np.zeros_like(u); l = np.zeros_like(u)
        lib.hostsim_math(len(u), u.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p))
        return e, l

    rng = np.random.default_rng(1)
    u = np.concatenate([rng.uniform(-12, 1, 400000), np.linspace(-10, 0.5, 20001), -np.logspace(-12, 1, 500)])
    e, _ = run(u)
    ref = np.exp(u.astype(np.longdouble))
    assert float(np.max(np.abs((e - ref) / ref))) < 2.5e-16
    z = np.concatenate([rng.uniform(1e-6, 2, 400000), np.exp(rng.uniform(-30, 5, 100000)), 1 - np.logspace(-15, -1, 500),
                        1 + np.logspace(-15, -1, 500), np.linspace(1 / 3, 1, 20001)])
    _, l = run(z)
    refl = np.log(z.astype(np.longdouble))
    err = np.abs(l - refl)
    rel = err / np.maximum(np.abs(refl), 1e-300)
    assert float(np.max(np.minimum(err, rel))) < 2.5e-16          # absolute or relative, whichever is smaller
    _, l = run(np.array([0.0, -1.0, np.inf]))
    assert l[0] == -np.inf and np.isnan(l[1]) and l[2] == np.inf  # rare arguments fall back to the library log


def test_division_by_three_without_a_division():
    """snaq_div3 (two FMAs after a multiplication) is g / 3.0 bit for bit: the reference's g*1/3 (src/pseudolik.jl:985-991)"""
    rng = np.random.default_rng(0)
    g = np.concatenate([rng.random(200000), rng.random(20000) * 2.0 ** -rng.integers(0, 60, 20000), rng.random(20000) * 3 - 1,
                        [0.0, 1.0, 0.5, 1 / 3, 2 / 3, 0.1, 3.0, 1e-300, -0.25]])
    q = np.zeros_like(g)
    hs.lib().hostsim_div3(len(g), hs._p(g), hs._p(q))
    assert np.array_equal(q, g / 3.0)


@pytest.mark.parametrize("case", ["G", "F", "I"])
def test_golden_cases_through_the_device_functions(case):
    net = {"G": S.readnewicklevel1(kat.CASE_G), "F": kat.case_f_network(S), "I": S.readnewicklevel1(kat.CASE_I)}[case]
    flat = net.flatten(kat.TAXA5)
    x0, numht, _ = hs.params(flat, 5)
    assert list(numht) == kat.NUMHT[case]
    r = hs.run(flat, 5, kat.q5_ids())
    for i, (which, types, t1, formula, expcf) in enumerate(kat.EXPCF[case]):
        assert r["which"][i] == which
        assert [t for t in r["types"][i] if t > 0] == types
        if which == 1:
            assert list(r["formula"][i]) == formula
        np.testing.assert_allclose(r["expcf"][i], expcf, rtol=0, atol=3e-16)


@pytest.mark.parametrize("block", range(6))
def test_random_networks_descriptors_and_values(oracle, block):
    nbad = 0
    for seed in range(block * 25, block * 25 + 25):
        n = 8 + seed % 18
        h = seed % 7
        try:
            net = simulate.random_level1_network(n, h, seed, min_cycle=3 if seed % 3 == 0 else 4)
        except S.SnaqError:
            continue
        taxa = [f"t{i + 1}" for i in range(n)]
        flat = net.flatten(taxa)
        qt = simulate.all_quartets(n)
        p = oracle.parameters(flat)
        x0, numht, _ = hs.params(flat, n)
        assert np.array_equal(x0, p["x"]) and np.array_equal(numht, p["numht"])
        rng = np.random.default_rng(seed)
        X = simulate.parameter_batch(p["x"], p["upper"], 6, seed)
        # the 10.0 cap together with negative type-4 terms: the reference's merge order matters there
        X[3] = np.where(p["upper"] > 1.5, rng.choice([0.0, 10.0, 5.0, 0.3], size=len(x0)), rng.choice([0.01, 0.5, 0.3], size=len(x0)))
        X[4] = np.where(p["upper"] > 1.5, 10.0, 0.5)
        o = oracle.extract(flat, qt)
        obs = simulate.obs_from_expcf(o["expcf"], 50, seed)
        r = hs.run(flat, n, qt, obs=obs, X=X)
        assert r["status"] == 0
        assert np.array_equal(r["which"], o["which"]) and np.array_equal(r["formula"], o["formula"])
        for qi in range(len(qt)):
            assert [t for t in o["types"][qi][: o["ntype"][qi]] if t != 1] == [t for t in r["types"][qi] if t > 0]
        want, ex = oracle.evaluate(flat, qt, obs, X, want_expcf=True)
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(r["lik"]), fin)
        assert np.max(np.abs(r["lik"][fin] - want[fin]) / np.abs(want[fin])) <= 1e-12
        assert np.max(np.abs(r["expcf"] - ex)) <= 1e-13
        nbad += 0
    assert nbad == 0


@pytest.mark.parametrize("seed", [61, 95, 103, 115, 7, 200, 333])
def test_cap_corner_with_a_pruned_bad_diamond_behind_a_type4_chain(oracle, seed):
    """Regression (found by a randomized sweep): when the pruned quartet holds more than one hybrid, cleanUpNode!
    (src/pseudolik.jl:852-857) merges the chain below a type-4 hybrid up to the next degree-3 node BEFORE the type-4
    length is computed; if the other junction is a pruned bad diamond I, its edge -log(1-gammaz) already exists and is
    part of that chain.  Only visible when lengths sit at the 10.0 cap and the type-4 term is negative."""
    n, h = 8 + seed % 22, seed % 7
    try:
        net = simulate.random_level1_network(n, h, 9000 + seed, min_cycle=3 if seed % 3 == 0 else 4)
    except S.SnaqError:
        pytest.skip("no network for this seed")
    taxa = [f"t{i + 1}" for i in range(n)]
    flat = net.flatten(taxa)
    qt = simulate.all_quartets(n)
    p = oracle.parameters(flat)
    rng = np.random.default_rng(seed)
    X = np.zeros((4, len(p["x"])))
    for k in range(4):
        X[k] = np.where(p["upper"] > 1.5, rng.choice([0.0, 10.0, 5.0, 0.3], size=len(p["x"])), rng.choice([0.01, 0.5, 0.3, 0.99], size=len(p["x"])))
    obs = np.full((len(qt), 3), 1 / 3)
    for k in range(4):
        try:
            want, ex = oracle.evaluate(flat, qt, obs, X[k:k + 1], want_expcf=True)
        except Exception:
            continue                                               # a vector the reference rejects (negative length too negative)
        r = hs.run(flat, n, qt, obs=obs, X=X[k:k + 1])
        assert np.max(np.abs(r["expcf"] - ex)) <= 1e-12


def test_negative_expcf_penalty_skips_unsampled_quartets(oracle):
    """-1e15 per negative expCF (src/pseudolik.jl:1394-1396) applies to SAMPLED quartets only (:1390)."""
    net = S.readnewicklevel1(kat.NET_NEG)
    taxa = [str(i) for i in range(1, 7)]
    flat = net.flatten(taxa)
    qt = simulate.all_quartets(6)
    obs = np.tile([0.5, 0.3, 0.2], (len(qt), 1))
    x = oracle.parameters(flat)["x"].copy()
    x[-2:] = [0.9, 0.8]                                           # gammaz sum > 1: negative expCFs
    for mask in (None, np.zeros(len(qt), np.uint8), (np.arange(len(qt)) % 2).astype(np.uint8)):
        want = oracle.evaluate(flat, qt, obs, x, sampled=mask)[0]
        got = hs.run(flat, 6, qt, obs=obs, X=x, sampled=mask)["lik"][0]
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (mask, got, want)
    assert oracle.evaluate(flat, qt, obs, x)[0] > 1e15
