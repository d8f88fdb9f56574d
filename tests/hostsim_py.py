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


def hasedge_sim(flat, ntaxa, qt):
    """hasEdge / bad-diamond-I flag through the simulation of extractQuartet! (snaq_hesim.h): any level-1 network"""
    qt = np.ascontiguousarray(qt, dtype=np.uint16).reshape(-1, 4)
    nq = qt.shape[0]
    npar = len(params(flat, ntaxa)[0])
    he = np.zeros((nq, max(npar, 1)), np.uint8)
    bd1 = np.zeros(nq, np.uint8)
    cnh = C.c_int32(); cnt = C.c_int32(); cnz = C.c_int32()
    if lib().hostsim_counts(flat.ptr(), ntaxa, C.byref(cnh), C.byref(cnt), C.byref(cnz)):
        raise RuntimeError(lib().hostsim_last_error().decode())
    nh, nt, nz = cnh.value, cnt.value, cnz.value
    zp = np.full((nq, max(nz, 1)), -1, np.int16)
    rc = lib().hostsim_hasedge_sim(flat.ptr(), ntaxa, C.c_int64(nq), _p(qt), _p(he), _p(bd1), _p(zp))
    if rc:
        raise RuntimeError(lib().hostsim_last_error().decode())
    return he[:, :npar], bd1, zp[:, :nz], (nh, nt)


def indexht_from(he, bd1, zpos, nh, nt):
    """parameters!(qnet,net): gamma slots, t slots (slot order = qnet.edge order), then the gammaz slots in qnet.edge order"""
    if bd1:
        return [i + 1 for i in range(nh + nt, len(he)) if he[i]]
    out = [i + 1 for i in range(nh + nt) if he[i]]
    z = [(int(zpos[i - nh - nt]), i + 1) for i in range(nh + nt, len(he)) if he[i]]
    return out + [s for _, s in sorted(z)]


def grad(flat, ntaxa, qt, obs, x, sampled=None):
    """objective and analytic gradient at x (the device functions behind snaq_eval_grad, compiled for the CPU)"""
    qt = np.ascontiguousarray(qt, dtype=np.uint16).reshape(-1, 4)
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    g = np.zeros_like(x)
    f = C.c_double(0.0)
    st = C.c_uint32(0)
    sm = None if sampled is None else np.ascontiguousarray(sampled, dtype=np.uint8)
    rc = lib().hostsim_grad(flat.ptr(), ntaxa, C.c_int64(qt.shape[0]), _p(qt), _p(obs), _p(sm), _p(x), C.byref(f), _p(g), C.byref(st))
    if rc:
        raise RuntimeError(lib().hostsim_last_error().decode())
    return f.value, g, st.value


class _OptblOpts(C.Structure):
    _fields_ = [("ftol_rel", C.c_double), ("ftol_abs", C.c_double), ("xtol_rel", C.c_double), ("xtol_abs", C.c_double),
                ("maxeval", C.c_int32), ("ladder", C.c_int32), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class _OptblResult(C.Structure):
    _fields_ = [("fmin", C.c_double), ("proj_grad", C.c_double), ("nevals", C.c_int32), ("nlaunches", C.c_int32),
                ("niter", C.c_int32), ("status", C.c_int32)]


def optbl(flat, ntaxa, qt, obs, x0, ftol_rel=1e-6, ftol_abs=1e-6, xtol_rel=1e-2, xtol_abs=1e-3, maxeval=1000, ladder=20, fd=False):
    """the product's optimiser (snaq_optbl.cpp) over the CPU stand-in objective"""
    qt = np.ascontiguousarray(qt, dtype=np.uint16).reshape(-1, 4)
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    x = np.array(x0, dtype=np.float64, copy=True)
    o = _OptblOpts(ftol_rel, ftol_abs, xtol_rel, xtol_abs, maxeval, ladder, 1 if fd else 0, 0)
    r = _OptblResult()
    rc = lib().hostsim_optbl(flat.ptr(), ntaxa, C.c_int64(qt.shape[0]), _p(qt), _p(obs), C.byref(o), _p(x), C.byref(r))
    if rc:
        raise RuntimeError(f"optbl failed with code {rc}: " + lib().hostsim_last_error().decode())
    return x, dict(fmin=r.fmin, proj_grad=r.proj_grad, nevals=r.nevals, nlaunches=r.nlaunches, niter=r.niter, status=r.status)
