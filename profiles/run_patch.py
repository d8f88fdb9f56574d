"""Per-candidate cost of the descriptor table after a nearest-neighbour interchange on the 100-taxon table: full rebuild
(snaq_set_network) against the incremental patch (snaq_patch_network), and optBL! at the search tolerances beside them."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snaq_jl_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
net = S.simulate.random_level1_network(n, 3, 20261021)
taxa = [f"t{i + 1}" for i in range(n)]
qt = S.simulate.all_quartets(n)
eng = S.Engine(0)
eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
eng.set_network(net.flatten(taxa))
x0, _, _ = eng.get_params()
eng.set_obscf(S.simulate.obs_from_expcf(eng.eval_quartets(x0)[0], 300, 5))
rng = np.random.default_rng(0)
rows = []
for k in range(40):
    S.simulate.nni_move(net, rng)
    flat = net.flatten(taxa)
    t0 = time.perf_counter(); touched = eng.patch_network(flat); t1 = time.perf_counter()
    eng.set_network(flat); t2 = time.perf_counter()
    eng.patch_network(flat, force_all=True); t3 = time.perf_counter()
    x, _, _ = eng.get_params()
    _, r = eng.optbl(x); t4 = time.perf_counter()
    rows.append((touched, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
a = np.array(rows[5:], dtype=float)
part = a[a[:, 0] >= 0]
print(json.dumps({"ntaxa": n, "quartets": len(qt), "moves": len(a), "partial_patches": int(len(part)),
                  "touched_taxa_median": float(np.median(part[:, 0])) if len(part) else None,
                  "patch_ms_median": float(np.median(a[:, 1])), "patch_ms_partial_median": float(np.median(part[:, 1])) if len(part) else None,
                  "set_network_ms_median": float(np.median(a[:, 2])), "patch_all_ms_median": float(np.median(a[:, 3])),
                  "optbl_search_ms_median": float(np.median(a[:, 4])),
                  "by_touched": sorted([(int(t), round(p, 3)) for t, p in a[:, :2]])}))
