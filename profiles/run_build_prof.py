"""Host wall time of snaq_set_network / snaq_patch_network at 100 taxa (3,921,225 quartets), and with SNAQ_BUILD_TIMING=1 the
stages of one build (each stage then ends with a wait, so the stage times add up to more than an untimed build).
  python profiles/run_build_prof.py [per_quartet]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, snaq_jl_b200 as S
n = 100
net = S.simulate.random_level1_network(n, 3, 20261021)
taxa = [f"t{i+1}" for i in range(n)]
qt = S.simulate.all_quartets(n)
eng = S.Engine(0, dedup=False if "per_quartet" in sys.argv else None)
eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
fl = net.flatten(taxa)
for _ in range(3):
    eng.set_network(fl)
t0 = time.perf_counter()
for _ in range(20):
    eng.set_network(fl)
print("set_network ms", (time.perf_counter() - t0) / 20 * 1e3)
rng = np.random.default_rng(0)
tp = []
for _ in range(20):
    tok = S.simulate.nni_move(net, rng)
    f2 = net.flatten(taxa)
    t0 = time.perf_counter()
    touched = eng.patch_network(f2)
    tp.append(((time.perf_counter() - t0) * 1e3, touched))
print("patch_network ms (touched taxa):", " ".join(f"{a:.2f}({b})" for a, b in tp))
os.environ["SNAQ_BUILD_TIMING"] = "1"
eng.set_network(fl)
eng.set_network(fl)
for _ in range(4):
    tok = S.simulate.nni_move(net, rng)
    t0 = time.perf_counter()
    touched = eng.patch_network(net.flatten(taxa))
    print("patch with stage timing: touched", touched, "ms", (time.perf_counter() - t0) * 1e3)
