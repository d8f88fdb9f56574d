"""Phase timeline of one k_eval_grad launch (SNAQ_TRACE=1): where the time of an objective+gradient pass goes."""
import ctypes as C
import os
import sys

os.environ["SNAQ_TRACE"] = "1"
sys.path.insert(0, '/root/repo')
import numpy as np

import snaq_jl_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
h = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dedup = len(sys.argv) > 3 and sys.argv[3] == "dedup"
net = S.simulate.random_level1_network(n, h, 20261021)
taxa = [f"t{i + 1}" for i in range(n)]
qt = S.simulate.all_quartets(n)
eng = S.Engine(0, dedup=dedup)
eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
eng.set_network(net.flatten(taxa))
x0, _, upper = eng.get_params()
for _ in range(5):
    eng.eval_grad(x0)
buf = (C.c_uint64 * (8 * 1024))()
nb = C.c_int32(0)
eng.lib.snaq_get_trace(eng.ctx, buf, 1024, C.byref(nb))
t = np.array(buf[: nb.value * 8], dtype=np.uint64).reshape(-1, 8)
t0 = t[:, 0].min()
t = (t - t0).astype(np.float64) / 1e3
print("blocks", nb.value, "n_full-ish nparams", len(x0), "counters", eng.counters())
for k, nm in ((0, "entry"), (1, "prologue_done"), (2, "A1_done"), (3, "A2_done"), (4, "G_done"), (5, "block_barrier"),
              (7, "partials_written"), (6, "exit")):
    print(f"{nm:17s} min {t[:, k].min():7.2f}  median {np.median(t[:, k]):7.2f}  max {t[:, k].max():7.2f} us")
