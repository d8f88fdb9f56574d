"""How far from the converged optimum does optBL! stop under the reference's SEARCH tolerances (fRel = fAbs = 1e-6,
xRel = 1e-2, xAbs = 1e-3, src/snaq_optimization.jl:36-39)?  snaq_optbl (projected L-BFGS on the analytic gradient, run here
over the CPU build of the same optimiser source, tests/hostsim) against a bound-constrained derivative-free method of
BOBYQA's class on the ORACLE objective with the same tolerances (scipy's Powell: NLopt and its BOBYQA are not in the image,
SURVEY 8c).  The reference accepts a proposal iff the new value is lower by more than liktolAbs = 1e-6 (:1389), so what
matters is the gap to the value a converged optimisation would give.  CPU only.

  python profiles/run_optimizer_fidelity.py [nnets]  -> one JSON line (committed as profiles/r02_optimizer_fidelity.json)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.optimize import minimize

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate
from tests import hostsim_py as hs
from oracle import oracle

nnets = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rows = []
for k in range(nnets):
    n, h = 9 + k % 8, k % 4
    net = other = None
    for attempt in range(30):
        try:
            net = simulate.random_level1_network(n, h, 4000 + k + 1000 * attempt)
            other = simulate.random_level1_network(n, h, 9000 + k + 1000 * attempt)
            break
        except S.SnaqError:
            continue
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    flat = net.flatten(taxa)
    src = flat if k % 2 == 0 else other.flatten(taxa)          # odd cases: data simulated on ANOTHER network (a hard start)
    obs = simulate.obs_from_expcf(oracle.extract(src, qt)["expcf"], 100, k)
    p = oracle.parameters(flat)
    x0, upper = p["x"], p["upper"]
    rng = np.random.default_rng(k)
    start = np.clip(x0 * rng.uniform(0.5, 1.5, len(x0)), 1e-3, upper - 1e-3)
    fo = lambda x: float(oracle.evaluate(flat, qt, obs, np.clip(x, 0, upper).reshape(1, -1))[0])
    bounds = list(zip(np.zeros_like(upper), upper))
    xt, rt = hs.optbl(flat, n, qt, obs, start, 1e-12, 1e-10, 1e-10, 1e-10)
    ref = minimize(fo, xt, method="L-BFGS-B", bounds=bounds, options=dict(ftol=1e-15, gtol=1e-10, maxiter=500))
    fstar = min(rt["fmin"], ref.fun)
    xs, rs = hs.optbl(flat, n, qt, obs, start)                   # search tolerances
    t0 = time.time()
    pw = minimize(fo, start, method="Powell", bounds=bounds, options=dict(xtol=1e-2, ftol=1e-6, maxfev=1000))
    rows.append(dict(n=n, h=h, nparams=len(x0), other_data=bool(k % 2), fstar=fstar, ours_gap=rs["fmin"] - fstar, ours_evals=rs["nevals"],
                     ours_launches=rs["nlaunches"], ours_status=rs["status"], ours_proj_grad=rs["proj_grad"],
                     powell_gap=float(pw.fun) - fstar, powell_evals=int(pw.nfev), f_start=fo(start)))
og = np.array([r["ours_gap"] for r in rows]); pg = np.array([r["powell_gap"] for r in rows])
out = dict(networks=nnets, tolerances="fRel=fAbs=1e-6, xRel=1e-2, xAbs=1e-3", ours_gap_median=float(np.median(og)), ours_gap_max=float(og.max()),
           powell_gap_median=float(np.median(pg)), powell_gap_max=float(pg.max()), ours_not_worse_than_powell=int(np.sum(og <= pg + 1e-9)),
           ours_gap_below_liktolAbs=int(np.sum(og <= 1e-6)), powell_gap_below_liktolAbs=int(np.sum(pg <= 1e-6)),
           ours_evals_median=float(np.median([r["ours_evals"] for r in rows])), ours_launches_median=float(np.median([r["ours_launches"] for r in rows])),
           powell_evals_median=float(np.median([r["powell_evals"] for r in rows])), rows=rows)
print(json.dumps(out))
