"""Latency of one objective call and one objective+gradient call on SMALL tables (the sizes SNaQ is mostly run at),
timed around the bare C-ABI calls (pointers prepared once), with the in-kernel phase timeline of k_eval_direct."""
import ctypes as C
import json
import os
import sys
import time

TRACE = os.environ.get("SNAQ_TRACE", "0") != "0"       # SNAQ_TRACE=1 adds the in-kernel timeline (and a few microseconds per call)
sys.path.insert(0, '/root/repo')
import numpy as np

import snaq_jl_b200 as S

res = {}
for n, h in ((8, 1), (12, 1), (20, 2), (30, 2), (50, 2), (100, 3)):
    net = S.simulate.random_level1_network(n, h, 20261021)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = S.simulate.all_quartets(n)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    eng.set_network(net.flatten(taxa))
    x0, _, upper = eng.get_params()
    e, _, _ = eng.eval_quartets(x0)
    eng.set_obscf(S.simulate.obs_from_expcf(e, 300, 5))
    npar = len(x0)
    x = np.ascontiguousarray(x0); out = np.zeros(1); g = np.zeros(npar); f = C.c_double(0)
    px, pout, pg = x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p)
    lib, ctx = eng.lib, eng.ctx
    r = {"nq": len(qt), "nparams": npar}
    for name, call in (("eval_us", lambda: lib.snaq_eval(ctx, 1, npar, px, pout)),
                       ("eval_grad_us", lambda: lib.snaq_eval_grad(ctx, npar, px, C.byref(f), pg, None))):
        for _ in range(50):
            call()
        N = 2000
        t = time.perf_counter()
        for _ in range(N):
            call()
        r[name] = (time.perf_counter() - t) / N * 1e6
    X = np.ascontiguousarray(S.simulate.parameter_batch(x0, upper, 20, 3)); o20 = np.zeros(20)
    pX, po = X.ctypes.data_as(C.c_void_p), o20.ctypes.data_as(C.c_void_p)
    for _ in range(50):
        lib.snaq_eval(ctx, 20, npar, pX, po)
    t = time.perf_counter()
    for _ in range(1000):
        lib.snaq_eval(ctx, 20, npar, pX, po)
    r["eval_P20_us"] = (time.perf_counter() - t) / 1000 * 1e6
    # the weighted quartet draw of a move proposal: on the device, against reading all the weights back (what the
    # reference does per proposal: weights = [q.deltaCF for q in d.quartet])
    u = np.array([0.37]); idx = np.zeros(1, np.int64)
    pu, pi = u.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p)
    for _ in range(20):
        lib.snaq_sample_quartets(ctx, npar, px, 1, pu, pi)
    t = time.perf_counter()
    for _ in range(300):
        lib.snaq_sample_quartets(ctx, npar, px, 1, pu, pi)
    r["sample_quartet_us"] = (time.perf_counter() - t) / 300 * 1e6
    dl = np.zeros(len(qt)); pdl = dl.ctypes.data_as(C.c_void_p)
    for _ in range(5):
        lib.snaq_eval_quartets(ctx, npar, px, None, None, pdl)
    t = time.perf_counter()
    for _ in range(50):
        lib.snaq_eval_quartets(ctx, npar, px, None, None, pdl)
    r["deltacf_readback_us"] = (time.perf_counter() - t) / 50 * 1e6
    if not TRACE:
        res[n] = r
        continue
    lib.snaq_eval(ctx, 1, npar, px, pout)
    buf = (C.c_uint64 * (8 * 1024))(); nb = C.c_int32(0)
    lib.snaq_get_trace(ctx, buf, 1024, C.byref(nb))
    t = np.array(buf[: nb.value * 8], dtype=np.uint64).reshape(-1, 8)[:, :5]
    t = t[(t[:, 0] > 0) & (t[:, 4] >= t[:, 0])]             # blocks that ran
    t = (t - t[:, 0].min()).astype(np.float64) / 1e3
    r["direct_kernel_phases_us"] = {"blocks": int(len(t)), "prologue_done_median": float(np.median(t[:, 1])),
                                    "loops_done_max": float(t[:, 3].max()), "exit_max": float(t[:, 4].max())}
    res[n] = r
print(json.dumps(res))
