"""Phase timeline of one k_eval_direct launch (SNAQ_TRACE=1): where the time of a single objective call goes."""
import ctypes as C
import os
import sys

os.environ["SNAQ_TRACE"] = "1"
sys.path.insert(0, '/root/repo')
import numpy as np
import torch

import snaq_jl_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
P = int(sys.argv[2]) if len(sys.argv) > 2 else 1
net = S.simulate.random_level1_network(n, 3, 20261021)
taxa = [f"t{i + 1}" for i in range(n)]
qt = S.simulate.all_quartets(n)
eng = S.Engine(0)
eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
eng.set_network(net.flatten(taxa))
x0, _, upper = eng.get_params()
X = S.simulate.parameter_batch(x0, upper, P, 3)
for _ in range(5):
    eng.eval(X)
buf = (C.c_uint64 * (8 * 1024))()
nb = C.c_int32(0)
eng.lib.snaq_get_trace(eng.ctx, buf, 1024, C.byref(nb))
t = np.array(buf[: nb.value * 8], dtype=np.int64).reshape(-1, 8)[:, :5]
t0 = t[:, 0].min()
t = (t - t0) / 1e3
names = ["entry", "prologue_done", "generic_done", "additive_done", "exit"]
for k, nm in enumerate(names):
    print(f"{nm:15s} min {t[:, k].min():7.2f}  median {np.median(t[:, k]):7.2f}  max {t[:, k].max():7.2f} us")
