"""ctypes binding of the C ABI (include/snaq_b200.h) and the host-side mirror of the
reference's hot-path functions.

Reference functions mirrored here (a trailing ``_`` stands for Julia's ``!``):

===========================  ===============================================
``extractQuartet_(net, d)``   ``extractQuartet!``      src/pseudolik.jl:503-540
``calculateExpCFAll_(d,x,net)`` ``calculateExpCFAll!`` src/pseudolik.jl:1310-1324
``logPseudoLik(d)``           ``logPseudoLik``         src/pseudolik.jl:1388-1425
``topologyQpseudolik_(net,d)`` ``topologyQpseudolik!`` src/pseudolik.jl:1342-1378
``parameters_(net)``          ``parameters!``          src/snaq_optimization.jl:94-101
===========================  ===============================================

There is no CPU path: if the CUDA library is missing or no GPU is visible every
call raises ``SnaqError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from .datacf import DataCF
from .network import FlatNet, HybridNetwork, SnaqError, set_length

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SNAQ_B200_LIB") or os.path.join(_HERE, "csrc", "libsnaqb200.so")   # (override: the bounds-checked build)
_LIB = None

EXPORTS = ["snaq_create", "snaq_destroy", "snaq_last_error", "snaq_version", "snaq_set_data", "snaq_set_obscf",
           "snaq_set_sampled", "snaq_set_network", "snaq_patch_network", "snaq_num_params", "snaq_get_params",
           "snaq_eval", "snaq_eval_device", "snaq_eval_quartets", "snaq_get_descriptors", "snaq_get_counters",
           "snaq_check_status", "snaq_measure_fp64_peak", "snaq_get_trace", "snaq_eval_grad", "snaq_optbl", "snaq_update_bl", "snaq_get_hasedge",
           "snaq_eval_topologies", "snaq_optbl_topologies", "snaq_sample_quartets", "snaq_set_ci", "snaq_resample_obscf", "snaq_get_obscf",
           "snaq_comm_unique_id", "snaq_comm_init", "snaq_comm_info", "snaq_patch_info", "snaq_table_checksum", "snaq_get_quartet_edges", "snaq_debug_bounds"]


class _OptblOpts(C.Structure):
    _fields_ = [("ftol_rel", C.c_double), ("ftol_abs", C.c_double), ("xtol_rel", C.c_double), ("xtol_abs", C.c_double),
                ("maxeval", C.c_int32), ("ladder", C.c_int32), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class _OptblResult(C.Structure):
    _fields_ = [("fmin", C.c_double), ("proj_grad", C.c_double), ("nevals", C.c_int32), ("nlaunches", C.c_int32),
                ("niter", C.c_int32), ("status", C.c_int32)]


class _Opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("world_size", C.c_int32), ("flags", C.c_uint32)]


def load_library():
    """Load libsnaqb200.so (built in-tree by ``__graft_entry__.build()``); fails loudly."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SnaqError(f"CUDA engine not built: {LIB_PATH} is missing (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                            "there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        lib.snaq_last_error.restype = C.c_char_p
        lib.snaq_last_error.argtypes = [C.c_void_p]
        lib.snaq_version.restype = C.c_char_p
        _LIB = lib
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Engine:
    """One engine context = one GPU + one stream (``snaq_ctx``)."""

    DEFAULT_DEDUP: Optional[bool] = None     # what dedup=None means (the test suite runs both table modes through it)

    def __init__(self, device: int = 0, rank: int = 0, world_size: int = 1, dedup: Optional[bool] = None):
        """dedup: None / True = the default run table (each distinct quartet program evaluated once per vector with the
        summed weights of its quartets); False = the per-quartet kernels (SNAQ_OPT_PER_QUARTET, include/snaq_b200.h)"""
        self.lib = load_library()
        self.ctx = C.c_void_p()
        if dedup is None:
            dedup = Engine.DEFAULT_DEDUP
        self.per_quartet = dedup is False
        opts = _Opts(device, rank, world_size, 2 if dedup is False else (1 if dedup else 0))
        rc = self.lib.snaq_create(C.byref(opts), C.byref(self.ctx))
        if rc:
            raise SnaqError(f"snaq_create failed ({rc}): {self.lib.snaq_last_error(None).decode()}")
        self.device = device
        self.rank, self.world_size = rank, max(world_size, 1)
        self.nparams = None
        self.nq = 0
        self.q_begin, self.nq_local = 0, 0
        self._data_id = None

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.lib.snaq_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise SnaqError(self.lib.snaq_last_error(self.ctx).decode())

    # -- quartet-sharded mode --------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """the 128-byte id rank 0 creates and hands to the other ranks (snaq_comm_unique_id)"""
        lib = load_library()
        buf = C.create_string_buffer(128)
        if lib.snaq_comm_unique_id(buf):
            raise SnaqError(lib.snaq_last_error(None).decode())
        return buf.raw

    def comm_init(self, comm_id: bytes) -> None:
        """collective over the ranks of a quartet-sharded context: afterwards eval / eval_device / eval_grad / optbl
        return totals over all ranks (snaq_comm_init)"""
        if len(comm_id) != 128:
            raise SnaqError("the communicator id has 128 bytes")
        self._chk(self.lib.snaq_comm_init(self.ctx, C.c_char_p(comm_id)))

    def comm_info(self):
        r = C.c_int32(); w = C.c_int32(); h = C.c_int32(); f = C.c_int32()
        self._chk(self.lib.snaq_comm_info(self.ctx, C.byref(r), C.byref(w), C.byref(h), C.byref(f)))
        return dict(rank=r.value, world_size=w.value, has_comm=bool(h.value), fused_epilogue=bool(f.value))

    # -- data ----------------------------------------------------------------------
    def set_data(self, d: DataCF) -> None:
        self._chk(self.lib.snaq_set_data(self.ctx, C.c_int32(len(d.taxa)), C.c_int64(d.numQuartets),
                                         _p(d.quartet_taxa), _p(d.obsCF)))
        self.nq = d.numQuartets
        # a quartet-sharded context keeps the contiguous block [q_begin, q_begin + nq_local) (snaq_set_data); the per-quartet
        # read-backs (eval_quartets, descriptors, hasedge) of such a context cover its own block
        self.q_begin = self.nq * self.rank // self.world_size
        self.nq_local = self.nq * (self.rank + 1) // self.world_size - self.q_begin
        self.nparams = None
        if not bool(np.all(d.sampled)):
            self.set_sampled(d.sampled)
        self._data_id = id(d)

    def set_obscf(self, obscf: np.ndarray) -> None:
        ob = np.ascontiguousarray(obscf, dtype=np.float64)
        if ob.size != 3 * self.nq:
            raise SnaqError("obsCF has the wrong size")
        self._chk(self.lib.snaq_set_obscf(self.ctx, _p(ob)))

    def set_ci(self, lo: np.ndarray, hi: np.ndarray) -> None:
        """credibility intervals of the three CFs of every quartet, [nq, 3] each (the CI columns of the bootstrap table)"""
        lo = np.ascontiguousarray(lo, dtype=np.float64).reshape(-1, 3)
        hi = np.ascontiguousarray(hi, dtype=np.float64).reshape(-1, 3)
        if lo.shape != (self.nq, 3) or hi.shape != lo.shape:
            raise SnaqError("lo and hi must have one row of three values per quartet")
        self._chk(self.lib.snaq_set_ci(self.ctx, _p(lo), _p(hi)))

    def resample_obscf(self, seed: int) -> None:
        """``sampleCFfromCI!`` + ``readtableCF!`` on the device (src/bootstrap.jl:52-70): a new bootstrap table in HBM"""
        self._chk(self.lib.snaq_resample_obscf(self.ctx, C.c_uint64(seed)))

    def get_obscf(self) -> np.ndarray:
        out = np.zeros((self.nq, 3))
        self._chk(self.lib.snaq_get_obscf(self.ctx, _p(out)))
        return out

    def set_sampled(self, mask: Optional[np.ndarray]) -> None:
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._chk(self.lib.snaq_set_sampled(self.ctx, _p(m)))

    # -- topology --------------------------------------------------------------------
    def set_network(self, flat: FlatNet) -> None:
        self._chk(self.lib.snaq_set_network(self.ctx, flat.ptr()))
        self.set_network_meta()

    def patch_network(self, flat: FlatNet, touched_taxa=None, force_all: bool = False) -> int:
        """``extractQuartet!`` after an in-place edit of the network (snaq_patch_network): only quartets with a taxon whose
        root path changed are classified again.  Returns the number of taxa re-described (-1: all)."""
        tt = None if touched_taxa is None else np.ascontiguousarray(touched_taxa, dtype=np.int32)
        n = -1 if force_all else (0 if tt is None else int(tt.size))
        self._chk(self.lib.snaq_patch_network(self.ctx, flat.ptr(), _p(tt), C.c_int32(n)))
        self.set_network_meta()
        out = C.c_int64(0)
        self._chk(self.lib.snaq_patch_info(self.ctx, C.byref(out)))
        return out.value

    def debug_bounds(self):
        """(checks, violations) of the bounds-checked build (libsnaqb200_dbg.so); (-1, -1) for the normal library"""
        c = C.c_int64(); v = C.c_int64()
        self._chk(self.lib.snaq_debug_bounds(self.ctx, C.byref(c), C.byref(v)))
        return c.value, v.value

    def table_checksum(self):
        out = (C.c_uint64 * 4)()
        self._chk(self.lib.snaq_table_checksum(self.ctx, out))
        return tuple(int(v) for v in out)

    def set_network_meta(self) -> None:
        n = C.c_int32(); nh = C.c_int32(); nt = C.c_int32(); nz = C.c_int32()
        self._chk(self.lib.snaq_num_params(self.ctx, C.byref(n), C.byref(nh), C.byref(nt), C.byref(nz)))
        self.nparams, self.nh, self.nt, self.nhz = n.value, nh.value, nt.value, nz.value

    def eval_topologies(self, flats) -> np.ndarray:
        """pseudo-deviance of each flat network at its own parameters (topologyQpseudolik! over a list of candidates)"""
        from .network import _CFlat
        arr = (_CFlat * len(flats))(*[f.c for f in flats])
        out = np.zeros(len(flats))
        self._chk(self.lib.snaq_eval_topologies(self.ctx, C.c_int32(len(flats)), arr, _p(out)))
        if flats:
            self.set_network_meta()
        return out

    def optbl_topologies(self, flats, workers=0, ftol_rel=1e-6, ftol_abs=1e-6, xtol_rel=1e-2, xtol_abs=1e-3, maxeval=1000,
                         ladder=20, fd=False):
        """``optBL!`` of every candidate flat network, run concurrently on this GPU (snaq_optbl_topologies): returns a list
        of (xmin, info) in the order of ``flats``.  Each entry equals ``set_network`` + ``optbl`` on one context, bit for
        bit.  The first candidate that fails raises SnaqError ("topology <c>: ...") after all candidates have run."""
        from .network import _CFlat
        n = len(flats)
        arr = (_CFlat * max(n, 1))(*[f.c for f in flats])
        ldx = max([8] + [4 * len(f.edge_number) for f in flats])          # nparams <= #edges + 2 * #hybrids
        xmin = np.zeros((max(n, 1), ldx))
        npar = np.zeros(max(n, 1), np.int32)
        res = (_OptblResult * max(n, 1))()
        o = _OptblOpts(ftol_rel, ftol_abs, xtol_rel, xtol_abs, maxeval, ladder, 1 if fd else 0, 0)
        self._chk(self.lib.snaq_optbl_topologies(self.ctx, C.c_int32(n), arr, C.byref(o), C.c_int32(workers), _p(xmin),
                                                 C.c_int32(ldx), _p(npar), res))
        return [(xmin[c, : npar[c]].copy(), dict(fmin=res[c].fmin, proj_grad=res[c].proj_grad, nevals=res[c].nevals,
                                                 nlaunches=res[c].nlaunches, niter=res[c].niter, status=res[c].status))
                for c in range(n)]

    def get_params(self):
        x = np.zeros(self.nparams); numht = np.zeros(self.nparams, np.int32); up = np.zeros(self.nparams)
        self._chk(self.lib.snaq_get_params(self.ctx, _p(x), _p(numht), _p(up)))
        return x, numht, up

    # -- evaluation ------------------------------------------------------------------
    def eval(self, X) -> np.ndarray:
        """-loglik for each row of X (host buffers; H2D + kernel + D2H)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X.reshape(1, -1)
        out = np.zeros(X.shape[0])
        self._chk(self.lib.snaq_eval(self.ctx, C.c_int32(X.shape[0]), C.c_int32(X.shape[1]), _p(X), _p(out)))
        return out

    def eval_device(self, dX, dout, stream=None) -> None:
        """Same with torch CUDA tensors (float64, contiguous); no synchronisation."""
        P, n = int(dX.shape[0]), int(dX.shape[1])
        self._chk(self.lib.snaq_eval_device(self.ctx, C.c_int32(P), C.c_int32(n), C.c_void_p(dX.data_ptr()),
                                            C.c_void_p(dout.data_ptr()),
                                            C.c_void_p(stream.cuda_stream if stream is not None else 0)))

    def check_status(self) -> None:
        self._chk(self.lib.snaq_check_status(self.ctx))

    def eval_quartets(self, x, want=("expcf", "logpl", "deltacf")):
        x = np.ascontiguousarray(x, dtype=np.float64)
        e = np.zeros((self.nq_local, 3)) if "expcf" in want else None
        l = np.zeros(self.nq_local) if "logpl" in want else None
        dl = np.zeros(self.nq_local) if "deltacf" in want else None
        self._chk(self.lib.snaq_eval_quartets(self.ctx, C.c_int32(x.size), _p(x), _p(e), _p(l), _p(dl)))
        return e, l, dl

    def sample_quartets(self, x, u) -> np.ndarray:
        """indices (0-based, data order) of quartets drawn with weights deltaCF at x, one per uniform draw in ``u``:
        step 1 of sampleEdgeQuartetWeighted (src/moves_snaq.jl:1443-1449) without copying the weights to the host"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.ascontiguousarray(np.atleast_1d(u), dtype=np.float64)
        idx = np.zeros(u.size, np.int64)
        self._chk(self.lib.snaq_sample_quartets(self.ctx, C.c_int32(x.size), _p(x), C.c_int32(u.size), _p(u), _p(idx)))
        return idx

    def eval_grad(self, x, want_hdiag=False):
        """objective and its gradient at x in one pass over the quartets (snaq_eval_grad)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.zeros_like(x)
        h = np.zeros_like(x) if want_hdiag else None
        f = C.c_double(0.0)
        self._chk(self.lib.snaq_eval_grad(self.ctx, C.c_int32(x.size), _p(x), C.byref(f), _p(g), _p(h)))
        return (f.value, g, h) if want_hdiag else (f.value, g)

    def optbl(self, x0, ftol_rel=1e-6, ftol_abs=1e-6, xtol_rel=1e-2, xtol_abs=1e-3, maxeval=1000, ladder=20, fd=False):
        """``optBL!`` on the current network (snaq_optbl): returns (xmin, dict(fmin, nevals, nlaunches, niter, status, proj_grad))"""
        x = np.array(x0, dtype=np.float64, copy=True)
        if x.size != self.nparams:
            raise SnaqError(f"ht(net) (length {self.nparams}) and vector x (length {x.size}) need to have same length")
        o = _OptblOpts(ftol_rel, ftol_abs, xtol_rel, xtol_abs, maxeval, ladder, 1 if fd else 0, 0)
        r = _OptblResult()
        rc = self.lib.snaq_optbl(self.ctx, C.byref(o), _p(x), C.byref(r))
        if rc:
            msg = self.lib.snaq_last_error(self.ctx).decode()
            raise SnaqError(msg or "tolerances have to be positive, ftol (rel,abs), xtol (rel,abs)")
        return x, dict(fmin=r.fmin, proj_grad=r.proj_grad, nevals=r.nevals, nlaunches=r.nlaunches, niter=r.niter, status=r.status)

    def hasedge(self):
        """(hasEdge [nq, nparams] 0/1, indexht lists) of every quartet: updateHasEdge! / parameters!(qnet,net)"""
        npar = max(self.nparams, 1)
        he = np.zeros((self.nq_local, npar), np.uint8)
        ix = np.full((self.nq_local, npar), -1, np.int32)
        nix = np.zeros(self.nq_local, np.int32)
        self._chk(self.lib.snaq_get_hasedge(self.ctx, _p(he), _p(ix), _p(nix)))
        return he[:, : self.nparams], [list(ix[i, : nix[i]]) for i in range(self.nq_local)]

    def quartet_edges(self, q: int):
        """edge numbers of ``d.quartet[q].qnet.edge`` (0-based q in the data's order): step 2 of sampleEdgeQuartetWeighted"""
        out = np.zeros(4096, np.int32)
        n = C.c_int32()
        self._chk(self.lib.snaq_get_quartet_edges(self.ctx, C.c_int64(q), _p(out), C.c_int32(out.size), C.byref(n)))
        return [int(v) for v in out[: n.value]]

    def update_bl(self):
        """per edge-length parameter: (Nquartets, edgeL) of ``updateBL!`` (src/readquartetdata.jl:838-861)"""
        nq = np.zeros(max(self.nt, 1), np.int64)
        el = np.zeros(max(self.nt, 1), np.float64)
        self._chk(self.lib.snaq_update_bl(self.ctx, _p(nq), _p(el)))
        return nq[: self.nt], el[: self.nt]

    def descriptors(self, max_types: int = 8):
        which = np.zeros(self.nq_local, np.int8); ntype = np.zeros(self.nq_local, np.int8)
        types = np.zeros((self.nq_local, max_types), np.int8); formula = np.zeros((self.nq_local, 3), np.int8)
        self._chk(self.lib.snaq_get_descriptors(self.ctx, _p(which), _p(ntype), _p(types), C.c_int32(max_types),
                                                _p(formula), None))
        return dict(which=which, ntype=ntype, types=types, formula=formula)

    def measure_fp64_peak(self) -> float:
        v = C.c_double(0.0)
        self._chk(self.lib.snaq_measure_fp64_peak(self.ctx, C.byref(v)))
        return v.value

    def counters(self):
        out = (C.c_int64 * 8)()
        self._chk(self.lib.snaq_get_counters(self.ctx, out))
        v = list(out)
        return dict(eval_launches=v[0], quartet_evals=v[1], h2d_bytes=v[2], d2h_bytes=v[3], descriptor_words=v[4],
                    build_launches=v[5], single_call_bytes=v[6], distinct_programs=v[7])


# ------------------------------------------------------------------------------------
# reference-facing functions
# ------------------------------------------------------------------------------------
def _engine_for(d: DataCF, device: int = 0) -> Engine:
    """One engine per (DataCF, device): obsCF and the quartet list stay resident in HBM.  The engine lives on the
    DataCF object itself (no global cache keyed by id(): ids are reused once an object is collected)."""
    engines = d.__dict__.setdefault("_engines", {})
    eng = engines.get(device)
    if eng is None:
        eng = Engine(device)
        eng.set_data(d)
        engines[device] = eng
    d._engine = eng
    return eng


def parameters_(net: HybridNetwork):
    """``parameters!(net)``: src/snaq_optimization.jl:94-101."""
    return net.parameters()


def extractQuartet_(net: HybridNetwork, d: DataCF, device: int = 0) -> None:
    """``extractQuartet!(net,d)`` (src/pseudolik.jl:503-517): (re)builds the device descriptor
    table for ``net`` and fills expCF / deltaCF of every quartet at the network's own parameters."""
    eng = _engine_for(d, device)
    net.parameters()
    eng.set_network(net.flatten(d.taxa))
    if eng.nparams != len(net.ht):
        raise SnaqError("engine and host disagree on the number of parameters")
    d._net_id = id(net)
    d._x = np.array(net.ht, dtype=np.float64)
    d.expCF, d.logPseudoLik, d.deltaCF = eng.eval_quartets(d._x)


def calculateExpCFAll_(d: DataCF, x: Sequence[float], net: HybridNetwork) -> None:
    """``calculateExpCFAll!(d,x,net)`` (src/pseudolik.jl:1310-1324): expCF of every sampled
    quartet under x (stateless: equivalent to every qnet being `changed`)."""
    eng = getattr(d, "_engine", None)
    if eng is None or getattr(d, "_net_id", None) != id(net):
        raise SnaqError("qnet in quartets on data are not correctly updated with extractQuartet")
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.size != eng.nparams:
        raise SnaqError(f"ht(net) (length {eng.nparams}) and vector x (length {x.size}) need to have same length")
    d._x = x.copy()
    d.expCF, d.logPseudoLik, d.deltaCF = eng.eval_quartets(x)


def logPseudoLik(d: DataCF) -> float:
    """``logPseudoLik(d)`` (src/pseudolik.jl:1417-1425): -sum of the per-quartet values."""
    eng = getattr(d, "_engine", None)
    if eng is None or getattr(d, "_x", None) is None:
        raise SnaqError("expCF not updated: run extractQuartet_ / calculateExpCFAll_ first")
    return float(eng.eval(d._x)[0])


def optBL_(net: HybridNetwork, d: DataCF, ftolRel: float = 1e-6, ftolAbs: float = 1e-6, xtolRel: float = 1e-2,
           xtolAbs: float = 1e-3, device: int = 0):
    """``optBL!`` (src/snaq_optimization.jl:341-390): optimises every gamma / branch length / gammaz of ``net`` on the
    device (snaq_optbl), then ``updateParameters!(net, xmin)`` and ``loglik!(net, fmin)``.  Defaults are the
    search-time tolerances fRel, fAbs, xRel, xAbs (:36-39)."""
    if not (ftolRel > 0 and ftolAbs > 0 and xtolAbs > 0 and xtolRel > 0):
        raise SnaqError(f"tolerances have to be positive, ftol (rel,abs), xtol (rel,abs): {[ftolRel, ftolAbs, xtolRel, xtolAbs]}")
    extractQuartet_(net, d, device)
    eng = d._engine
    xmin, res = eng.optbl(np.array(net.ht, dtype=np.float64), ftolRel, ftolAbs, xtolRel, xtolAbs)
    net.update_parameters(list(xmin))
    net.loglik = res["fmin"]
    d._x = xmin.copy()
    return res


def optBL_candidates_(nets, d: DataCF, workers: int = 0, ftolRel: float = 1e-6, ftolAbs: float = 1e-6, xtolRel: float = 1e-2,
                      xtolAbs: float = 1e-3, device: int = 0):
    """``optBL!`` of several proposed networks over the same data, concurrently on one GPU (snaq_optbl_topologies; the
    loop it batches: src/snaq_optimization.jl:1362-1432).  Every net ends as ``optBL_`` would leave it (ht, loglik)."""
    if not (ftolRel > 0 and ftolAbs > 0 and xtolAbs > 0 and xtolRel > 0):
        raise SnaqError(f"tolerances have to be positive, ftol (rel,abs), xtol (rel,abs): {[ftolRel, ftolAbs, xtolRel, xtolAbs]}")
    eng = _engine_for(d, device)
    for net in nets:
        net.parameters()
    out = eng.optbl_topologies([net.flatten(d.taxa) for net in nets], workers, ftolRel, ftolAbs, xtolRel, xtolAbs)
    for net, (xmin, res) in zip(nets, out):
        net.update_parameters(list(xmin))
        net.loglik = res["fmin"]
    return [res for _, res in out]


def updateBL_(net: HybridNetwork, d: DataCF, device: int = 0):
    """``updateBL!`` (src/readquartetdata.jl:838-861): rough starting branch lengths of a TREE from the mean observed CF
    of the quartets that correspond exactly to each internal edge; returns {edge number: (Nquartets, edgeL)}."""
    if net.numhybrids > 0:
        raise SnaqError("updateBL! was created for a tree, and net here is not a tree, so no branch lengths updated")
    eng = _engine_for(d, device)
    net.parameters()
    eng.set_network(net.flatten(d.taxa))
    d._net_id = None                                   # the lengths change below: extractQuartet_ must run again
    _, numht, _ = eng.get_params()
    nq, el = eng.update_bl()
    tab = {int(numht[eng.nh + i]): (int(nq[i]), float(el[i])) for i in range(eng.nt) if nq[i] > 0}
    for e in net.edge:
        if e.number in tab and (e.length < 0.0 or e.length == 1.0):      # :850-853
            set_length(e, tab[e.number][1] if tab[e.number][1] > 0 else 0.0)
    for e in net.edge:
        if e.length < 0.0:                                                # :855-858
            set_length(e, 1.0)
    return tab


def topologyQpseudolik_(net: HybridNetwork, d: DataCF, device: int = 0) -> float:
    """``topologyQpseudolik!`` (src/pseudolik.jl:1342-1378): pseudo-deviance of a fixed network;
    updates ``net.loglik`` and the expected CFs stored in ``d``."""
    for e in net.edge:
        if e.hybrid and not e.gamma >= 0.0:
            raise SnaqError("hybrid edge has missing γ value. Cannot compute quartet pseudo-likelihood.")
    extractQuartet_(net, d, device)
    val = logPseudoLik(d)
    net.loglik = val
    return val
