"""Randomized parity sweep on the GPU: updateBL! tables, hasEdge/indexht (0-4 hybrids, nested cycles) and the per-quartet
read-backs under a sampling mask, each against the oracle.  Usage: python profiles/run_sweep.py [n_seeds [first_seed]]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snaq_jl_b200 as S
from snaq_jl_b200 import simulate
from oracle import oracle, update_bl as ubl

nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 80
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
nb = nh = nq_ = 0
for seed in range(first, first + nseeds):
    rng = np.random.default_rng(seed)
    # updateBL! on random trees, ragged permuted tables
    n = 6 + seed % 16
    net = simulate.random_level1_network(n, 0, 100 + seed)
    taxa = [f"t{i+1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    qt = qt[rng.uniform(size=len(qt)) < 0.7]
    qt = np.array([rng.permutation(q) for q in qt], dtype=np.uint16)
    if len(qt) == 0:
        continue
    flat = net.flatten(taxa)
    obs = simulate.obs_from_expcf(oracle.extract(flat, qt)["expcf"], 25, seed)
    eng = S.Engine(0); eng.set_data(S.DataCF(taxa, qt, obs)); eng.set_network(flat)
    cnt, el = eng.update_bl(); _, numht, _ = eng.get_params()
    tab = ubl.update_bl_table(flat, qt, obs)
    got = {int(numht[eng.nh + i]): (int(cnt[i]), float(el[i])) for i in range(eng.nt) if cnt[i] > 0}
    assert sorted(got) == sorted(tab), (seed, 'ubl keys')
    for k in tab:
        same = abs(got[k][1] - tab[k][1]) <= 1e-9 * max(1, abs(tab[k][1])) or (np.isinf(tab[k][1]) and np.isinf(got[k][1]))
        assert got[k][0] == tab[k][0] and same, (seed, k, got[k], tab[k])
    nb += 1
    # hasEdge with <= 1 hybrid
    h = seed % 5
    try:
        net = simulate.random_level1_network(n + 2, h, 300 + seed, min_cycle=3 if seed % 3 == 0 else 4)
    except S.SnaqError:
        continue
    taxa = [f"t{i+1}" for i in range(n + 2)]
    flat = net.flatten(taxa); qt = simulate.all_quartets(n + 2)
    o = oracle.extract(flat, qt)
    eng = S.Engine(0); eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3))); eng.set_network(flat)
    he, ix = eng.hasedge()
    assert np.array_equal(he, o["hasedge"]), (seed, 'hasedge')
    assert all(ix[i] == list(o["indexht"][i][:o["nindexht"][i]]) for i in range(len(qt))), (seed, 'indexht')
    nh += 1
    # per-quartet read-backs under a mask
    p = oracle.parameters(flat); X = simulate.parameter_batch(p["x"], p["upper"], 3, seed)
    obs = simulate.obs_from_expcf(o["expcf"], 20, seed)
    mask = (rng.uniform(size=len(qt)) < 0.6).astype(np.uint8)
    eng.set_obscf(obs); eng.set_sampled(mask)
    want, ex = oracle.evaluate(flat, qt, obs, X[2:3], sampled=mask, want_expcf=True)
    e, lp, dl = eng.eval_quartets(X[2])
    sm = mask.astype(bool)
    assert np.max(np.abs(e[sm] - ex[sm])) < 1e-13 and abs(-lp.sum() - want[0]) <= 1e-10 * abs(want[0]) and not dl[~sm].any()
    assert np.allclose(dl[sm], np.abs(obs[sm] - ex[sm]).sum(1), atol=1e-13)
    nq_ += 1
print('updateBL', nb, 'hasEdge', nh, 'readbacks', nq_, 'all equal to the oracle')
