"""ctypes wrapper around oracle/libsnaq_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (the product never does).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libsnaq_oracle.so")
    src = os.path.join(_HERE, "snaq_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "snaq_b200.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsnaq_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libsnaq_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.oracle_last_error.restype = C.c_char_p
    return _LIB


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def _chk(rc):
    if rc != 0:
        raise OracleError(rc, lib().oracle_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def parameters(flat):
    n = C.c_int32(); nh = C.c_int32(); nt = C.c_int32(); nhz = C.c_int32()
    _chk(lib().oracle_parameters(flat.ptr(), C.byref(n), C.byref(nh), C.byref(nt), C.byref(nhz), None, None, None))
    x = np.zeros(n.value); numht = np.zeros(n.value, dtype=np.int32); up = np.zeros(n.value)
    _chk(lib().oracle_parameters(flat.ptr(), C.byref(n), C.byref(nh), C.byref(nt), C.byref(nhz), _p(x), _p(numht), _p(up)))
    return dict(x=x, numht=numht, upper=up, nh=nh.value, nt=nt.value, nhz=nhz.value)


def extract(flat, quartet_taxa, x=None, max_types=8):
    """extractQuartet! + calculateExpCFAll! per quartet; returns a dict of arrays."""
    qt = np.ascontiguousarray(quartet_taxa, dtype=np.uint16).reshape(-1, 4)
    nq = qt.shape[0]
    np_ = parameters(flat)["x"].size
    out = dict(which=np.zeros(nq, np.int8), ntype=np.zeros(nq, np.int8), types=np.zeros((nq, max_types), np.int8),
               formula=np.zeros((nq, 3), np.int8), hasedge=np.zeros((nq, max(np_, 1)), np.uint8),
               nindexht=np.zeros(nq, np.int32), indexht=np.full((nq, max(np_, 1)), -1, np.int32),
               expcf=np.zeros((nq, 3)), t1=np.zeros(nq))
    xx = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    _chk(lib().oracle_extract(flat.ptr(), C.c_int64(nq), _p(qt), _p(xx), _p(out["which"]), _p(out["ntype"]),
                              _p(out["types"]), C.c_int32(max_types), _p(out["formula"]), _p(out["hasedge"]),
                              _p(out["nindexht"]), _p(out["indexht"]), _p(out["expcf"]), _p(out["t1"])))
    out["hasedge"] = out["hasedge"][:, :np_]
    out["indexht"] = out["indexht"][:, :np_]
    return out


def quartet_edges(flat, taxa4):
    """edge numbers of q.qnet.edge after extractQuartet!(net, quartet), in order"""
    t = np.ascontiguousarray(taxa4, dtype=np.uint16)
    out = np.zeros(4096, np.int32)
    n = C.c_int32()
    _chk(lib().oracle_quartet_edges(flat.ptr(), _p(t), _p(out), C.c_int32(out.size), C.byref(n)))
    return [int(v) for v in out[: n.value]]


def evaluate(flat, quartet_taxa, obscf, X, sampled=None, threads=0, want_expcf=False):
    """logPseudoLik for each row of X (optBL! objective, stateless)."""
    qt = np.ascontiguousarray(quartet_taxa, dtype=np.uint16).reshape(-1, 4)
    ob = np.ascontiguousarray(obscf, dtype=np.float64).reshape(-1, 3)
    X = np.ascontiguousarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(1, -1)
    P, npar = X.shape
    out = np.zeros(P)
    ex = np.zeros((qt.shape[0], 3)) if want_expcf else None
    sm = None if sampled is None else np.ascontiguousarray(sampled, dtype=np.uint8)
    _chk(lib().oracle_eval(flat.ptr(), C.c_int64(qt.shape[0]), _p(qt), _p(ob), _p(sm), C.c_int32(P), C.c_int32(npar),
                           _p(X), _p(out), _p(ex), C.c_int32(threads)))
    return (out, ex) if want_expcf else out
