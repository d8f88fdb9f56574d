"""GPU tier: snaq_eval_grad (objective + analytic gradient in one pass) and snaq_optbl (optBL!,
src/snaq_optimization.jl:341-390) through the C ABI, against the oracle.

The reference's optimiser is NLopt's BOBYQA: its iterates cannot be reproduced, so the
parity statement is on the optimum: same minimiser as a tightly converged bound-constrained optimisation
of the ORACLE objective (scipy L-BFGS-B), branch lengths and gammas within 1e-6 where the data identify
them, and the known optimum (deviance 0 at the simulating parameters) on perfect data."""
import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[None, False], ids=["run_table", "per_quartet"])
def table_mode(request, monkeypatch):
    """the default run table (distinct programs, summed weights) and the per-quartet kernels (SNAQ_OPT_PER_QUARTET)"""
    monkeypatch.setattr(S.Engine, "DEFAULT_DEDUP", request.param)
    return request.param


def make(n, h, seed, ngenes=None):
    net = None
    for attempt in range(30):                                   # small trees cannot always take h reticulations
        try:
            net = simulate.random_level1_network(n, h, seed + 1000 * attempt)
            break
        except S.SnaqError:
            continue
    assert net is not None
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    flat = net.flatten(taxa)
    eng.set_network(flat)
    x0, _, upper = eng.get_params()
    e, _, _ = eng.eval_quartets(x0)
    obs = e if ngenes is None else simulate.obs_from_expcf(e, ngenes, seed)
    eng.set_obscf(obs)
    return eng, flat, qt, obs, x0, upper


@pytest.mark.parametrize("seed", range(8))
def test_gradient_matches_central_differences_of_the_oracle(oracle, seed, table_mode):
    n, h = 9 + 2 * seed, seed % 5
    eng, flat, qt, obs, x0, upper = make(n, h, 300 + seed, ngenes=60)
    x = np.clip(simulate.parameter_batch(x0, upper, 4, seed)[2], 1e-3, upper - 1e-3)
    f, g = eng.eval_grad(x)
    assert abs(f - eng.eval(x)[0]) <= 1e-11 * abs(f)
    hstep = 1e-6
    X = np.tile(x, (2 * len(x), 1))
    for i in range(len(x)):
        X[2 * i, i] += hstep
        X[2 * i + 1, i] -= hstep
    ff = oracle.evaluate(flat, qt, obs, X)
    gfd = (ff[0::2] - ff[1::2]) / (2 * hstep)
    assert np.max(np.abs(g - gfd) / (1 + np.abs(gfd))) < 2e-5
    f2, g2 = eng.eval_grad(x)                                   # integer bins: bit-reproducible
    assert f2 == f and np.array_equal(g, g2)


@pytest.mark.parametrize("n,h,seed", [(45, 3, 11), (60, 4, 12)])
def test_streaming_gradient_kernel_on_large_tables(oracle, n, h, seed, monkeypatch):
    """tables with more than 65,536 additive quartets take k_eval_grad (TMA ring, program memo, grid-wide integer bins);
    the deduplicated engine takes the small kernel and the generic path for every entry: two independent routes to the
    same gradient and curvature, plus central differences of the oracle on a few coordinates"""
    monkeypatch.setattr(S.Engine, "DEFAULT_DEDUP", False)       # per-quartet table for `eng`
    eng, flat, qt, obs, x0, upper = make(n, h, seed, ngenes=80)
    rng = np.random.default_rng(seed)
    mask = (rng.uniform(size=len(qt)) < 0.8).astype(np.uint8)
    eng.set_sampled(mask)
    x = np.clip(simulate.parameter_batch(x0, upper, 4, seed)[3], 1e-3, upper - 1e-3)
    f, g, hd = eng.eval_grad(x, want_hdiag=True)
    f1, g1 = eng.eval_grad(x)
    assert f1 == f and np.array_equal(g1, g)
    assert abs(f - eng.eval(x)[0]) <= 1e-11 * abs(f)
    want = oracle.evaluate(flat, qt, obs, x[None], sampled=mask)[0]
    assert abs(f - want) <= 1e-10 * abs(want)
    ded = S.Engine(0, dedup=True)
    ded.set_data(S.DataCF([f"t{i + 1}" for i in range(n)], qt, obs))
    ded.set_sampled(mask)
    ded.set_network(flat)
    fd_, gd, hdd = ded.eval_grad(x, want_hdiag=True)
    assert abs(fd_ - f) <= 1e-10 * abs(f)
    assert np.max(np.abs(g - gd) / (1 + np.abs(gd))) < 1e-7
    assert np.max(np.abs(hd - hdd) / (1 + np.abs(hdd))) < 1e-7
    coords = rng.choice(len(x), size=6, replace=False)
    hstep = 1e-6
    X = np.tile(x, (2 * len(coords), 1))
    for k, i in enumerate(coords):
        X[2 * k, i] += hstep
        X[2 * k + 1, i] -= hstep
    ff = oracle.evaluate(flat, qt, obs, X, sampled=mask)
    gfd = (ff[0::2] - ff[1::2]) / (2 * hstep)
    noise = 4e-15 * abs(want) / (2 * hstep)
    assert np.max(np.abs(g[coords] - gfd) / (1 + np.abs(gfd))) < 2e-5 + noise
    # out-of-range candidates are reported by the streaming kernel too
    bad = x.copy()
    bad[0] = -0.1
    with pytest.raises(S.SnaqError):
        eng.eval_grad(bad)
    f3, g3 = eng.eval_grad(x)                                   # and the bins are clean for the next call
    assert f3 == f and np.array_equal(g3, g)


def test_curvature_diagonal_is_the_hessian_diagonal_on_a_tree(oracle, table_mode):
    eng, flat, qt, obs, x0, upper = make(12, 0, 41, ngenes=80)
    x = np.clip(x0 * 1.1, 1e-2, upper - 1e-2)
    f, g, hd = eng.eval_grad(x, want_hdiag=True)
    f2, g2 = eng.eval_grad(x)
    assert f2 == f and np.array_equal(g, g2)
    hstep = 1e-4
    X = np.tile(x, (2 * len(x), 1))
    for i in range(len(x)):
        X[2 * i, i] += hstep
        X[2 * i + 1, i] -= hstep
    ff = oracle.evaluate(flat, qt, obs, X)
    h_fd = (ff[0::2] - 2 * f + ff[1::2]) / hstep ** 2
    assert np.max(np.abs(hd - h_fd) / (1 + np.abs(h_fd))) < 1e-3
    # with hybrids: non-negative for lengths, zero for gammas
    eng2, _, _, _, x2, up2 = make(16, 3, 43, ngenes=80)
    _, _, hd2 = eng2.eval_grad(x2, want_hdiag=True)
    assert np.all(hd2 >= 0) and np.all(hd2[: eng2.nh] == 0)


@pytest.mark.parametrize("seed", range(4))
def test_optbl_recovers_the_simulating_parameters_on_perfect_data(seed):
    n, h = 10 + 3 * seed, 1 + seed % 3
    eng, flat, qt, obs, x0, upper = make(n, h, 500 + seed)
    rng = np.random.default_rng(seed)
    start = np.clip(x0 * rng.uniform(0.6, 1.5, len(x0)) + rng.uniform(0, 0.05, len(x0)), 0, upper)
    xmin, res = eng.optbl(start, ftol_rel=1e-12, ftol_abs=1e-10, xtol_rel=1e-10, xtol_abs=1e-10)   # fRelBL..: :44-47
    assert res["fmin"] < 1e-7 and res["status"] in (3, 4, 5)
    # perfect data: the fitted expected CFs are the observed ones (some parameter directions are nearly flat)
    e, _, _ = eng.eval_quartets(xmin)
    assert np.max(np.abs(e - obs)) < 2e-5
    if seed in (0, 1, 3):
        assert np.max(np.abs(xmin - x0)) < 1e-4


@pytest.mark.parametrize("seed", range(4))
def test_optbl_matches_a_tight_optimisation_of_the_oracle_objective(oracle, seed, table_mode):
    from scipy.optimize import minimize
    n, h = 8 + seed, seed % 3
    eng, flat, qt, obs, x0, upper = make(n, h, 700 + seed, ngenes=100)
    rng = np.random.default_rng(seed)
    start = np.clip(x0 * rng.uniform(0.7, 1.3, len(x0)), 1e-3, upper - 1e-3)

    def fo(x):
        return float(oracle.evaluate(flat, qt, obs, x.reshape(1, -1))[0])

    ref = minimize(fo, start, method="L-BFGS-B", bounds=list(zip(np.zeros_like(upper), upper)),
                   options=dict(ftol=1e-15, gtol=1e-9, maxiter=2000, maxfun=200000))
    xmin, res = eng.optbl(start, ftol_rel=1e-12, ftol_abs=1e-10, xtol_rel=1e-10, xtol_abs=1e-10)
    assert res["fmin"] <= ref.fun + 1e-7 * max(1.0, abs(ref.fun))       # at least as good as the CPU optimum
    assert abs(fo(xmin) - res["fmin"]) <= 1e-10 * max(1.0, abs(res["fmin"]))
    # where the optimum is interior and well determined the minimisers agree
    if abs(res["fmin"] - ref.fun) <= 1e-9 * max(1.0, abs(ref.fun)):
        interior = (ref.x > 1e-4) & (ref.x < upper - 1e-4)
        assert np.max(np.abs(xmin - ref.x)[interior]) < 5e-4


def test_optbl_reaches_the_optima_the_reference_published():
    """examples/net0.out, net1.out, net1_snaq.out: same -Ploglik and same branch lengths / gammas (1e-6) as the reference's
    own final optimisation, from a different start; the search-time-only networks of net1.networks: not worse."""
    import os
    from tests import kat
    from tests.test_optbl_cpu import _published_optima
    here, best, others = _published_optima()
    for nw, lik, table in best + others:
        d = S.readtableCF(os.path.join(here, "golden", table))
        net = S.readnewicklevel1(nw)
        x_ref = np.array(net.parameters(), dtype=np.float64)
        upper = np.array(net.upper(), dtype=np.float64)
        rng = np.random.default_rng(1)
        start = np.clip(np.where(upper > 1.5, 1.0, 0.3) * rng.uniform(0.5, 1.5, len(x_ref)), 0, upper)
        net.update_parameters(list(start))
        res = S.optBL_(net, d, 1e-12, 1e-10, 1e-10, 1e-10)                      # fRelBL, fAbsBL, xRelBL, xAbsBL
        if (nw, lik, table) in best:
            assert abs(net.loglik - lik) <= 1e-10 * lik
            assert np.max(np.abs(np.array(net.ht) - x_ref)) < 2e-6
        else:
            assert net.loglik <= lik + 1e-9


def test_optbl_default_tolerances_and_fd_mode_agree():
    eng, flat, qt, obs, x0, upper = make(14, 2, 901, ngenes=200)
    start = np.clip(x0 * 1.2, 0, upper)
    xa, ra = eng.optbl(start, ftol_rel=1e-12, ftol_abs=1e-10, xtol_rel=1e-10, xtol_abs=1e-10)
    xb, rb = eng.optbl(start, ftol_rel=1e-12, ftol_abs=1e-10, xtol_rel=1e-10, xtol_abs=1e-10, fd=True)
    assert abs(ra["fmin"] - rb["fmin"]) <= 1e-8 * abs(ra["fmin"])
    xc, rc = eng.optbl(start)                                   # search-time tolerances: 1e-6, 1e-6, 1e-2, 1e-3
    assert rc["fmin"] <= eng.eval(start)[0]
    assert rc["nlaunches"] <= ra["nlaunches"]
    with pytest.raises(S.SnaqError):
        eng.optbl(start, ftol_rel=0.0)
