"""``Quartet`` / ``DataCF`` (src/types.jl:217-273) as flat arrays, and ``readtableCF``
(src/readquartetdata.jl:97-217).

The reference keeps one heap object per quartet; here a DataCF is four arrays
(taxon ids nq x 4, obsCF nq x 3, ngenes, sampled) that go to HBM unchanged.
"""
from __future__ import annotations

import csv
import itertools
from typing import Iterable, List, Optional, Sequence

import numpy as np

from .network import SnaqError


class DataCF:
    def __init__(self, taxa: Sequence[str], quartet_taxa: np.ndarray, obscf: np.ndarray,
                 ngenes: Optional[np.ndarray] = None):
        self.taxa: List[str] = list(taxa)                      # tiplabels(d)
        self.quartet_taxa = np.ascontiguousarray(quartet_taxa, dtype=np.uint16).reshape(-1, 4)
        self.obsCF = np.ascontiguousarray(obscf, dtype=np.float64).reshape(-1, 3)
        nq = self.quartet_taxa.shape[0]
        if self.obsCF.shape[0] != nq:
            raise SnaqError("observed CF table and quartet list have different lengths")
        if np.any((self.obsCF < 0.0) | (self.obsCF > 1.0)):
            raise SnaqError("obsCF must be between (0,1)")      # src/types.jl:237-239
        self.ngenes = np.full(nq, -1.0) if ngenes is None else np.asarray(ngenes, dtype=np.float64)
        self.sampled = np.ones(nq, dtype=np.uint8)             # q.sampled
        # filled by the engine (q.qnet.expCF, q.logPseudoLik, q.deltaCF)
        self.expCF: Optional[np.ndarray] = None
        self.logPseudoLik: Optional[np.ndarray] = None
        self.deltaCF: Optional[np.ndarray] = None

    @property
    def numQuartets(self) -> int:
        return self.quartet_taxa.shape[0]

    def quartet_names(self, i: int) -> List[str]:
        return [self.taxa[t] for t in self.quartet_taxa[i]]

    @classmethod
    def from_rows(cls, rows: Iterable[Sequence], taxa: Optional[Sequence[str]] = None) -> "DataCF":
        """rows: (t1,t2,t3,t4,cf12_34,cf13_24,cf14_23[,ngenes])."""
        rows = [tuple(r) for r in rows]
        if taxa is None:
            seen = {}
            for r in rows:
                for t in r[:4]:
                    seen.setdefault(str(t), None)
            taxa = _sort_taxa(list(seen))
        tid = {t: i for i, t in enumerate(taxa)}
        qt = np.array([[tid[str(t)] for t in r[:4]] for r in rows], dtype=np.uint16).reshape(-1, 4)
        cf = np.array([[float(v) for v in r[4:7]] for r in rows], dtype=np.float64).reshape(-1, 3)
        ng = np.array([float(r[7]) if len(r) > 7 and r[7] not in ("", None) else -1.0 for r in rows])
        return cls(taxa, qt, cf, ng)

    @classmethod
    def all_quartets(cls, taxa: Sequence[str], obscf: np.ndarray) -> "DataCF":
        """All C(n,4) four-taxon sets in lexicographic order of taxon index
        (``allQuartets``, src/readquartetdata.jl:272-303)."""
        n = len(taxa)
        qt = np.array(list(itertools.combinations(range(n), 4)), dtype=np.uint16)
        return cls(taxa, qt, obscf)


def _sort_taxa(names: List[str]) -> List[str]:
    # sorttaxa!: numbers first in numeric order, then strings (src/auxiliary.jl)
    def key(s):
        try:
            return (0, int(s), "")
        except ValueError:
            return (1, 0, s)
    return sorted(names, key=key)


def readtableCF(path: str, delim: str = ",") -> DataCF:
    """``readtableCF(file)`` (src/readquartetdata.jl:97-140): taxon columns are found by name (t1 | tx1 | tax1 | taxon1,
    ...), CF columns by name (CF12_34 | CF12.34 | obsCF12 | CF1234, ...); an optional ``ngenes`` column is kept.
    Like the reference, a table whose columns cannot be identified is an error, not a guess."""
    with open(path, newline="") as fh:
        rd = csv.reader(fh, delimiter=delim)
        header = next(rd)
        body = [r for r in rd if r]
    low = [h.strip() for h in header]

    def find(*names):
        for i, h in enumerate(low):
            if h in names:
                return i
        return None

    tcol = [find(f"t{i}", f"tx{i}", f"tax{i}", f"taxon{i}") for i in (1, 2, 3, 4)]
    if any(c is None for c in tcol):
        raise SnaqError("columns for taxon names were not found")
    c1 = find("CF12_34", "CF12.34", "obsCF12", "CF1234")
    c2 = find("CF13_24", "CF13.24", "obsCF13", "CF1324")
    c3 = find("CF14_23", "CF14.23", "obsCF14", "CF1423")
    if c1 is None or c2 is None or c3 is None:
        raise SnaqError("Could not identify columns with quartet concordance factors (qCFs). Was expecting CF12_34, CF13_24 "
                        "and CF14_23 for the columns with CF values, or CF12.34 or obsCF12, etc.")
    cn = find("ngenes")
    rows = []
    for r in body:
        row = [r[tcol[0]].strip(), r[tcol[1]].strip(), r[tcol[2]].strip(), r[tcol[3]].strip(), r[c1], r[c2], r[c3]]
        if cn is not None:
            row.append(r[cn])
        rows.append(row)
    return DataCF.from_rows(rows)


# ---- read-back tables (src/descriptive.jl:24-59, src/readquartetdata.jl:4-14) -------------------------------------------
def fittedquartetCF(d: DataCF, format: str = "wide") -> dict:
    """Observed and expected CFs after a fit, as columns (a dict of equal-length lists / arrays in the reference's column
    order).  ``d.expCF`` is what ``Engine.eval_quartets`` / ``topologyQpseudolik_`` left there (q.qnet.expCF).
    wide: tx1..tx4, obsCF12, obsCF13, obsCF14, expCF12, expCF13, expCF14 (one row per four-taxon set);
    long: tx1..tx4, quartet ("12_34", "13_24", "14_23"), obsCF, expCF (three rows per four-taxon set)."""
    if d.expCF is None:
        raise SnaqError("expected CFs have not been calculated: evaluate the network on this DataCF first")
    names = [[d.taxa[t] for t in col] for col in d.quartet_taxa.T]
    if format == "wide":
        out = {f"tx{k + 1}": names[k] for k in range(4)}
        for j, s in enumerate(("12", "13", "14")):
            out["obsCF" + s] = d.obsCF[:, j].copy()
        for j, s in enumerate(("12", "13", "14")):
            out["expCF" + s] = d.expCF[:, j].copy()
        return out
    if format == "long":
        nq = d.numQuartets
        out = {f"tx{k + 1}": [v for v in names[k] for _ in range(3)] for k in range(4)}
        out["quartet"] = ["12_34", "13_24", "14_23"] * nq
        out["obsCF"] = d.obsCF.reshape(-1).copy()
        out["expCF"] = d.expCF.reshape(-1).copy()
        return out
    raise SnaqError(f"format {format} was not recognized. Should be :wide or :long.")


def writeExpCF(d: DataCF, path: Optional[str] = None, delim: str = ",") -> dict:
    """The expected CFs as the table readtableCF reads back: t1..t4, CF12_34, CF13_24, CF14_23 (src/readquartetdata.jl:4-14);
    written as CSV when ``path`` is given."""
    if d.expCF is None:
        raise SnaqError("expected CFs have not been calculated: evaluate the network on this DataCF first")
    cols = {f"t{k + 1}": [d.taxa[t] for t in d.quartet_taxa[:, k]] for k in range(4)}
    for j, s in enumerate(("CF12_34", "CF13_24", "CF14_23")):
        cols[s] = d.expCF[:, j].copy()
    if path is not None:
        with open(path, "w", newline="") as fh:
            w = csv.writer(fh, delimiter=delim)
            w.writerow(list(cols))
            for i in range(d.numQuartets):
                w.writerow([cols[f"t{k + 1}"][i] for k in range(4)] + [repr(float(cols[s][i])) for s in ("CF12_34", "CF13_24", "CF14_23")])
    return cols
