"""ctypes wrapper of tests/hostsim/libhostsim.so: the engine's device functions
compiled for the CPU (test harness only; the product has no CPU path)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-C", _HERE, "libhostsim.so"], stdout=subprocess.DEVNULL)
        _LIB = C.CDLL(os.path.join(_HERE, "libhostsim.so"))
        _LIB.hostsim_last_error.restype = C.c_char_p
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def params(flat, ntaxa):
    n = C.c_int32(); nxe = C.c_int32()
    rc = lib().hostsim_params(flat.ptr(), ntaxa, C.byref(n), C.byref(nxe), None, None)
    if rc:
        raise RuntimeError(lib().hostsim_last_error().decode())
    x0 = np.zeros(n.value); numht = np.zeros(n.value, np.int32)
    lib().hostsim_params(flat.ptr(), ntaxa, C.byref(n), C.byref(nxe), _p(x0), _p(numht))
    return x0, numht, nxe.value


def run(flat, ntaxa, qt, obs=None, X=None, sampled=None, maxh=8):
    qt = np.ascontiguousarray(qt, dtype=np.uint16).reshape(-1, 4)
    nq = qt.shape[0]
    obs = np.tile([0.5, 0.4, 0.1], (nq, 1)) if obs is None else np.ascontiguousarray(obs, dtype=np.float64)
    if X is None:
        X = params(flat, ntaxa)[0]
    X = np.ascontiguousarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(1, -1)
    P = X.shape[0]
    out = dict(lik=np.zeros(P), expcf=np.zeros((nq, 3)), which=np.zeros(nq, np.int8), types=np.zeros((nq, maxh), np.int8),
               formula=np.zeros((nq, 3), np.int8), proglen=np.zeros(nq, np.int32))
    st = C.c_uint32(0)
    sm = None if sampled is None else np.ascontiguousarray(sampled, dtype=np.uint8)
    rc = lib().hostsim_run(flat.ptr(), ntaxa, C.c_int64(nq), _p(qt), _p(obs), _p(sm), P, _p(X), _p(out["lik"]),
                           _p(out["expcf"]), _p(out["which"]), _p(out["types"]), maxh, _p(out["formula"]),
                           _p(out["proglen"]), C.byref(st))
    if rc:
        raise RuntimeError(lib().hostsim_last_error().decode())
    out["status"] = st.value
    return out


def hasedge(flat, ntaxa, qt):
    qt = np.ascontiguousarray(qt, dtype=np.uint16).reshape(-1, 4)
    nq = qt.shape[0]
    npar = len(params(flat, ntaxa)[0])
    he = np.zeros((nq, max(npar, 1)), np.uint8)
    bd1 = np.zeros(nq, np.uint8)
    rc = lib().hostsim_hasedge(flat.ptr(), ntaxa, C.c_int64(nq), _p(qt), _p(he), _p(bd1))
    if rc:
        raise RuntimeError(lib().hostsim_last_error().decode())
    return he[:, :npar], bd1
