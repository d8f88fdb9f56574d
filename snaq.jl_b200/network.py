"""Host-side mirror of the reference's level-1 network model.

Only what the hot path needs on the host: the Node/Edge/HybridNetwork fields the
reference reads through the accessors of ``src/types.jl:5-63``, an extended-newick
reader that reproduces ``readnewicklevel1`` / ``readTopologyUpdate``
(``src/readwrite.jl:172-194``: ``cleanBL!`` :448-459, ``cleanAfterRead!`` :23-101,
``updateAllReadTopology!`` :116-142) including ``updateInCycle!``
(``src/update.jl:23-81``) and ``updateGammaz!`` (``src/update.jl:177-276``), the
parameter layout ``parameters(net)`` (``src/snaq_optimization.jl:54-92``), and the
flattening into the ``snaq_flatnet`` wire format of ``include/snaq_b200.h``.

The topology search, moves and rooting logic stay in the caller (SURVEY.md §2:
out of scope); this module exists so that tests, the benchmark and a Python user
can hand networks to the engine the way the Julia shim does.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

NODE_LEAF, NODE_HYBRID, NODE_HASHYBEDGE = 0x01, 0x02, 0x04
NODE_BD1, NODE_BD2, NODE_BADTRI = 0x08, 0x10, 0x20
EDGE_HYBRID, EDGE_MAJOR, EDGE_ISCHILD1, EDGE_IDENT, EDGE_FROMBD1 = 0x01, 0x02, 0x04, 0x08, 0x10


class SnaqError(RuntimeError):
    """Raised where the reference raises ``ErrorException``."""


def _length_clean(v: float) -> float:
    # cleanBL!: src/readwrite.jl:448-459
    if v < 0:
        return 1.0
    return min(v, 10.0)


class Edge:
    __slots__ = ("number", "length", "hybrid", "ismajor", "gamma", "node", "ischild1",
                 "istIdentifiable", "fromBadDiamondI", "inCycle")

    def __init__(self, number: int, length: float = 1.0, hybrid: bool = False, gamma: float = 1.0,
                 ismajor: Optional[bool] = None):
        self.number = number
        self.length = float(length)
        self.hybrid = hybrid
        self.gamma = float(gamma)
        self.ismajor = (gamma >= 0.5) if ismajor is None else ismajor
        self.node: List["Node"] = []
        self.ischild1 = True
        self.istIdentifiable = True
        self.fromBadDiamondI = False
        self.inCycle = -1

    @property
    def y(self) -> float:
        return math.exp(-self.length)

    @property
    def z(self) -> float:
        return 1.0 - math.exp(-self.length)

    def __repr__(self):
        return f"Edge({self.number}, {[n.number for n in self.node]}, len={self.length}, hyb={self.hybrid}, g={self.gamma})"


class Node:
    __slots__ = ("number", "leaf", "hybrid", "name", "edge", "hasHybEdge", "isBadDiamondI",
                 "isBadDiamondII", "isExtBadTriangle", "isVeryBadTriangle", "isBadTriangle",
                 "inCycle", "k", "gammaz", "prev")

    def __init__(self, number: int, leaf: bool, hybrid: bool = False, name: str = ""):
        self.number = number
        self.leaf = leaf
        self.hybrid = hybrid
        self.name = name
        self.edge: List[Edge] = []
        self.hasHybEdge = False
        self.isBadDiamondI = self.isBadDiamondII = False
        self.isExtBadTriangle = self.isVeryBadTriangle = self.isBadTriangle = False
        self.inCycle = -1
        self.k = -1
        self.gammaz = -1.0
        self.prev = None

    def __repr__(self):
        return f"Node({self.number}, leaf={self.leaf}, hyb={self.hybrid}, name={self.name!r})"


def getOtherNode(e: Edge, n: Node) -> Node:
    return e.node[1] if e.node[0] is n else e.node[0]


def set_length(e: Edge, length: float, negative: bool = False) -> None:
    """``setLength!``: src/auxiliary.jl:118-129."""
    if not (negative or length >= 0):
        raise SnaqError(f"length has to be nonnegative: {length}, cannot set to edge {e.number}")
    if not length >= -0.4054651081081644:
        raise SnaqError("length can be negative, but not too negative (greater than -log(1.5))")
    e.length = min(float(length), 10.0)


def hybridEdges(node: Node) -> Tuple[Optional[Edge], Optional[Edge], Optional[Edge]]:
    """PhyloNetworks ``hybridEdges(node)``."""
    if len(node.edge) != 3:
        raise SnaqError(f"node {node.number} has {len(node.edge)} edges instead of 3")
    if node.hybrid:
        maj = mino = tree = None
        for e in node.edge:
            if e.hybrid and e.ismajor:
                maj = e
            if e.hybrid and not e.ismajor:
                mino = e
            if not e.hybrid:
                tree = e
        return maj, mino, tree
    if node.hasHybEdge:
        hyb = tc = tree = None
        for e in node.edge:
            if e.hybrid:
                hyb = e
            if not e.hybrid and e.inCycle != -1:
                tc = e
            if not e.hybrid and e.inCycle == -1:
                tree = e
        return hyb, tc, tree
    for i, e in enumerate(node.edge):
        if getOtherNode(e, node).leaf:
            rest = [x for j, x in enumerate(node.edge) if j != i]
            return rest[0], rest[1], e
    return node.edge[0], node.edge[1], node.edge[2]


def isEdgeIdentifiable(edge: Edge) -> bool:
    """``isEdgeIdentifiable``: src/update.jl:282-310."""
    if edge.hybrid:
        node = edge.node[0] if edge.node[0].hybrid else edge.node[1]
        if not node.hybrid:
            raise SnaqError(f"hybrid edge {edge.number} pointing at tree node")
        _, _, tree = hybridEdges(node)
        return not getOtherNode(tree, node).leaf
    a, b = edge.node
    if not a.leaf and not b.leaf:
        if not a.hybrid and not b.hybrid and not edge.fromBadDiamondI:
            return True
        if a.hybrid or b.hybrid:
            h = a if a.hybrid else b
            return not h.isBadDiamondII and not h.isBadTriangle
    return False


class HybridNetwork:
    """Level-1 semi-directed network with the attributes snaq! maintains."""

    def __init__(self):
        self.node: List[Node] = []
        self.edge: List[Edge] = []
        self.hybrid: List[Node] = []
        self.leaf: List[Node] = []
        self.names: List[str] = []
        self.numBad = 0
        self.hasVeryBadTriangle = False
        self.loglik = -1.0
        # parameters!(net)
        self.ht: List[float] = []
        self.numht: List[int] = []
        self.index: List[int] = []

    # -- basic bookkeeping ---------------------------------------------------
    @property
    def numhybrids(self) -> int:
        return len(self.hybrid)

    @property
    def numtaxa(self) -> int:
        return len(self.leaf)

    def tiplabels(self) -> List[str]:
        return [n.name for n in self.leaf]

    def push_node(self, n: Node) -> None:
        self.node.append(n)
        if n.leaf:
            self.leaf.append(n)
        if n.hybrid:
            self.hybrid.append(n)

    def connect(self, e: Edge, a: Node, b: Node) -> None:
        """edge.node = [a, b] (child first as PN's parser does), push e to both node.edge."""
        e.node = [a, b]
        a.edge.append(e)
        b.edge.append(e)

    # -- level-1 attribute update -------------------------------------------
    def update_all(self) -> None:
        """``updateAllReadTopology!``: src/readwrite.jl:116-142 (inCycle, gammaz, flags)."""
        for n in self.node:
            n.inCycle = -1
            n.k = -1
            n.gammaz = -1.0
            n.isBadDiamondI = n.isBadDiamondII = n.isBadTriangle = False
            n.isExtBadTriangle = n.isVeryBadTriangle = False
        for e in self.edge:
            e.inCycle = -1
            e.fromBadDiamondI = False
        for n in self.node:
            n.hasHybEdge = any(e.hybrid for e in n.edge)
        for e in self.edge:
            if not e.hybrid:
                e.istIdentifiable = (not e.node[0].leaf) and (not e.node[1].leaf)
        self.numBad = 0
        self.hasVeryBadTriangle = False
        for h in self.hybrid:
            self._update_in_cycle(h)
        for e in self.edge:
            if e.hybrid:
                e.istIdentifiable = isEdgeIdentifiable(e)
        for h in self.hybrid:
            self._update_gammaz(h)
        self.parameters()

    def _update_in_cycle(self, h: Node) -> None:
        """``updateInCycle!``: src/update.jl:23-81.  The cycle is the unique path
        between the hybrid node and the parent of its minor edge in the tree made
        of the major edges, closed by the minor edge."""
        minor = [e for e in h.edge if e.hybrid and not e.ismajor and (e.node[0] is h or e.node[1] is h)
                 and self._child(e) is h]
        if len(minor) != 1:
            raise SnaqError(f"hybrid node {h.number} does not have exactly one minor parent edge")
        hybedge = minor[0]
        last = getOtherNode(hybedge, h)
        # breadth-first search from h along major edges, never through leaves
        prev: Dict[int, Tuple[Node, Edge]] = {}
        seen = {id(h)}
        queue = [h]
        found = False
        while queue and not found:
            nxt = []
            for cur in queue:
                for e in cur.edge:
                    if not e.ismajor:
                        continue
                    o = getOtherNode(e, cur)
                    if id(o) in seen or o.leaf:
                        continue
                    seen.add(id(o))
                    prev[id(o)] = (cur, e)
                    if o is last:
                        found = True
                        break
                    nxt.append(o)
                if found:
                    break
            queue = nxt
        if not found:
            raise SnaqError(f"hybrid node {h.number} does not create a cycle")
        h.inCycle = h.number
        hybedge.inCycle = h.number
        k = 1
        cur = last
        while cur is not h:
            if cur.inCycle != -1:
                raise SnaqError(f"cycle of hybrid {h.number} intersects the cycle of hybrid {cur.inCycle}: not level-1")
            cur.inCycle = h.number
            k += 1
            p, e = prev[id(cur)]
            e.inCycle = h.number
            cur = p
        h.k = k

    @staticmethod
    def _child(e: Edge) -> Node:
        return e.node[0] if e.ischild1 else e.node[1]

    def _update_gammaz(self, node: Node) -> None:
        """``updateGammaz!(net,node,allow=true)``: src/update.jl:177-276."""
        edge_maj, edge_min, tree_edge2 = hybridEdges(node)
        other_maj = getOtherNode(edge_maj, node)
        other_min = getOtherNode(edge_min, node)
        if node.k <= 2:
            raise SnaqError(f"cycle with only {node.k} nodes: parallel edges")
        if node.k == 4:
            _, edge_min2, tree_edge3 = hybridEdges(other_min)
            _, edge_maj2, tree_edge1 = hybridEdges(other_maj)
            other_min2 = getOtherNode(edge_min2, other_min)
            isLeaf1 = getOtherNode(tree_edge1, other_maj)
            isLeaf2 = getOtherNode(tree_edge2, node)
            isLeaf3 = getOtherNode(tree_edge3, other_min)
            tree_edge4 = None
            for e in other_min2.edge:
                if tree_edge4 is None and e.inCycle == -1 and not e.hybrid:
                    tree_edge4 = e
            same = other_min2 is getOtherNode(edge_maj2, other_maj)
            if same and isLeaf1.leaf and isLeaf2.leaf and isLeaf3.leaf:  # bad diamond I
                self.numBad += 1
                node.isBadDiamondI = True
                other_min.gammaz = edge_min.gamma * edge_min2.z
                other_maj.gammaz = edge_maj.gamma * edge_maj2.z
                for e in (edge_min2, edge_maj2, edge_maj, edge_min):
                    e.istIdentifiable = False
            elif (same and isLeaf1.leaf and not isLeaf2.leaf and isLeaf3.leaf
                  and getOtherNode(tree_edge4, other_min2).leaf):  # bad diamond II
                node.isBadDiamondII = True
                set_length(edge_maj, edge_maj.length + tree_edge2.length)
                set_length(tree_edge2, 0.0)
                tree_edge2.istIdentifiable = False
                edge_maj.istIdentifiable = True
                edge_min.istIdentifiable = True
        elif node.k == 3:
            if self.numtaxa <= 5:
                node.isVeryBadTriangle = True
                self.hasVeryBadTriangle = True
            else:
                _, _, tree_edge1 = hybridEdges(other_min)
                _, _, tree_edge3 = hybridEdges(other_maj)
                leaves = [getOtherNode(tree_edge1, other_min).leaf, getOtherNode(tree_edge2, node).leaf,
                          getOtherNode(tree_edge3, other_maj).leaf]
                nl = sum(leaves)
                if nl >= 2:
                    node.isExtBadTriangle = True
                    self.hasVeryBadTriangle = True
                elif nl == 1:
                    node.isVeryBadTriangle = True
                    self.hasVeryBadTriangle = True
                else:
                    node.isBadTriangle = True
                    set_length(edge_maj, edge_maj.length + tree_edge2.length)
                    set_length(tree_edge2, 0.0)
                    tree_edge2.istIdentifiable = False
        if node.k > 3 and not node.isBadDiamondI and not node.isBadDiamondII:
            _, tree_edge_incycle, _ = hybridEdges(other_min)
            tree_edge_incycle.istIdentifiable = True
            edge_maj.istIdentifiable = isEdgeIdentifiable(edge_maj)
            edge_min.istIdentifiable = isEdgeIdentifiable(edge_min)

    # -- parameters ------------------------------------------------------------
    def parameters(self) -> List[float]:
        """``parameters!(net)``: src/snaq_optimization.jl:54-101.  x = [gamma | t | gammaz]."""
        t, h, hz, n, n2, hzn, it, ih, ihz = ([] for _ in range(9))
        for i, e in enumerate(self.edge):
            if e.istIdentifiable:
                t.append(e.length); n.append(e.number); it.append(i)
            if e.hybrid and not e.ismajor:
                node = self._child(e)
                if not node.hybrid:
                    raise SnaqError(f"strange thing, hybrid edge {e.number} pointing at tree node {node.number}")
                if not node.isBadDiamondI:
                    h.append(e.gamma); n2.append(e.number); ih.append(i)
                else:
                    edges = hybridEdges(node)
                    oa, ob = getOtherNode(edges[0], node), getOtherNode(edges[1], node)
                    hz += [oa.gammaz, ob.gammaz]
                    hzn += [int(f"{node.number}1"), int(f"{node.number}2")]
                    ihz += [self.node.index(oa), self.node.index(ob)]
        self.ht = h + t + hz
        self.numht = n2 + n + hzn
        self.index = ih + it + ihz
        self._nh, self._nt, self._nhz = len(h), len(t), len(hz)
        return self.ht

    def upper(self) -> List[float]:
        """``upper(net)``: src/snaq_optimization.jl:275-279."""
        return [1.0] * self._nh + [10.0] * self._nt + [1.0] * self._nhz

    def update_parameters(self, x: Sequence[float]) -> None:
        """``updateParameters!(net,xmin)``: src/snaq_optimization.jl:255-271."""
        if len(x) != len(self.ht):
            raise SnaqError(f"xmin vector should have same length as ht(net) {len(self.ht)}, not {len(x)}")
        self.ht = [float(v) for v in x]
        for i, v in enumerate(self.ht):
            if i < self._nh:
                if not 0 <= v <= 1:
                    raise SnaqError(f"new gamma value should be between 0,1: {v}.")
                e = self.edge[self.index[i]]
                node = self._child(e)
                partner = [p for p in node.edge if p.hybrid and p is not e][0]
                # setgamma!(e, v, true): the partner gets 1-v and ismajor follows gamma
                if abs(v - 0.5) < 1e-12:
                    v = 0.5
                e.gamma, partner.gamma = v, 1.0 - v
                e.ismajor = v >= 0.5
                partner.ismajor = not e.ismajor
            elif i < self._nh + self._nt:
                set_length(self.edge[self.index[i]], v)
            else:
                if not 0 <= v <= 1:
                    raise SnaqError(f"new gammaz value should be between 0,1: {v}.")
                self.node[self.index[i]].gammaz = v

    # -- flattening -------------------------------------------------------------
    def flatten(self, taxa: Sequence[str]) -> "FlatNet":
        """Serialise into the ``snaq_flatnet`` wire format (include/snaq_b200.h).
        ``taxa`` is the data's taxon list; leaves get their 0-based index in it."""
        tid = {t: i for i, t in enumerate(taxa)}
        nn, ne = len(self.node), len(self.edge)
        nidx = {id(n): i for i, n in enumerate(self.node)}
        eidx = {id(e): i for i, e in enumerate(self.edge)}
        f = FlatNet()
        f.node_number = np.array([n.number for n in self.node], dtype=np.int32)
        fl = []
        for n in self.node:
            v = (NODE_LEAF if n.leaf else 0) | (NODE_HYBRID if n.hybrid else 0) | (NODE_HASHYBEDGE if n.hasHybEdge else 0)
            v |= (NODE_BD1 if n.isBadDiamondI else 0) | (NODE_BD2 if n.isBadDiamondII else 0) | (NODE_BADTRI if n.isBadTriangle else 0)
            fl.append(v)
        f.node_flags = np.array(fl, dtype=np.uint32)
        f.node_incycle = np.array([n.inCycle for n in self.node], dtype=np.int32)
        f.node_gammaz = np.array([n.gammaz for n in self.node], dtype=np.float64)
        tx = []
        for n in self.node:
            if n.leaf:
                if n.name not in tid:
                    raise SnaqError(f"taxon {n.name} of the network is not in the data")
                tx.append(tid[n.name])
            else:
                tx.append(-1)
        f.node_taxon = np.array(tx, dtype=np.int32)
        ned = np.full((nn, 3), -1, dtype=np.int32)
        for i, n in enumerate(self.node):
            if len(n.edge) > 3:
                raise SnaqError(f"node {n.number} has {len(n.edge)} edges (polytomy)")
            for j, e in enumerate(n.edge):
                ned[i, j] = eidx[id(e)]
        f.node_edge = ned.reshape(-1)
        f.edge_number = np.array([e.number for e in self.edge], dtype=np.int32)
        fl = []
        for e in self.edge:
            v = (EDGE_HYBRID if e.hybrid else 0) | (EDGE_MAJOR if e.ismajor else 0) | (EDGE_ISCHILD1 if e.ischild1 else 0)
            v |= (EDGE_IDENT if e.istIdentifiable else 0) | (EDGE_FROMBD1 if e.fromBadDiamondI else 0)
            fl.append(v)
        f.edge_flags = np.array(fl, dtype=np.uint32)
        f.edge_incycle = np.array([e.inCycle for e in self.edge], dtype=np.int32)
        f.edge_length = np.array([e.length for e in self.edge], dtype=np.float64)
        f.edge_gamma = np.array([e.gamma for e in self.edge], dtype=np.float64)
        f.edge_node = np.array([[nidx[id(e.node[0])], nidx[id(e.node[1])]] for e in self.edge], dtype=np.int32).reshape(-1)
        f.hybrid = np.array([nidx[id(n)] for n in self.hybrid], dtype=np.int32)
        f.leaf = np.array([nidx[id(n)] for n in self.leaf], dtype=np.int32)
        f.finish()
        return f

    # -- newick out (enough for round trips and logs) ----------------------------
    def to_newick(self, digits: Optional[int] = None) -> str:
        """Extended newick rooted at a node compatible with the hybrid directions
        (cf. ``writenewick_level1``, src/readwrite.jl:248-343)."""
        def fmt(v):
            return repr(round(v, digits)) if digits is not None else repr(float(v))
        # direct edges away from a root that is above every hybrid: pick an internal
        # node from which every hybrid node is reached through its parents
        root = None
        for cand in self.node:
            if cand.leaf or cand.hybrid:
                continue
            if self._orient(cand) is not None:
                root = cand
                break
        if root is None:
            raise SnaqError("could not find a root compatible with the hybrid edges")
        parent_of = self._orient(root)
        seen_h = set()

        def rec(n: Node, pe: Optional[Edge]) -> str:
            kids = []
            for e in n.edge:
                if e is pe:
                    continue
                c = getOtherNode(e, n)
                if parent_of.get(id(e)) is not n:
                    continue
                kids.append((e, c))
            label = n.name if (n.leaf or n.hybrid) else ""
            if n.hybrid and n.number in seen_h:
                s = label
            else:
                if n.hybrid:
                    seen_h.add(n.number)
                s = ("(" + ",".join(rec(c, e) for e, c in kids) + ")" if kids else "") + label
            if pe is not None:
                s += ":" + fmt(pe.length)
                if pe.hybrid:
                    s += "::" + fmt(pe.gamma)
            return s

        return rec(root, None) + ";"

    def _orient(self, root: Node):
        """parent_of[edge] = parent node when rooted at ``root``; None if a hybrid
        edge would point away from its hybrid node."""
        parent_of = {}
        stack = [(root, None)]
        visited = {id(root)}
        while stack:
            n, pe = stack.pop()
            for e in n.edge:
                if e is pe or id(e) in parent_of:
                    continue
                c = getOtherNode(e, n)
                if e.hybrid:
                    if c is not self._child(e):
                        return None
                    parent_of[id(e)] = n
                    if id(c) not in visited:
                        visited.add(id(c))
                        stack.append((c, e))
                    continue
                if n.hybrid and pe is not None and not pe.hybrid:
                    return None  # entered a hybrid node through its tree edge
                if id(c) in visited:
                    continue
                parent_of[id(e)] = n
                visited.add(id(c))
                stack.append((c, e))
        return parent_of


class _CFlat(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_edges", C.c_int32), ("n_hybrids", C.c_int32), ("n_leaves", C.c_int32),
                ("node_number", C.c_void_p), ("node_flags", C.c_void_p), ("node_incycle", C.c_void_p),
                ("node_gammaz", C.c_void_p), ("node_taxon", C.c_void_p), ("node_edge", C.c_void_p),
                ("edge_number", C.c_void_p), ("edge_flags", C.c_void_p), ("edge_incycle", C.c_void_p),
                ("edge_length", C.c_void_p), ("edge_gamma", C.c_void_p), ("edge_node", C.c_void_p),
                ("hybrid", C.c_void_p), ("leaf", C.c_void_p)]


class FlatNet:
    """numpy arrays + the ctypes struct pointing at them (keeps them alive)."""
    _arrays = ("node_number", "node_flags", "node_incycle", "node_gammaz", "node_taxon", "node_edge",
               "edge_number", "edge_flags", "edge_incycle", "edge_length", "edge_gamma", "edge_node",
               "hybrid", "leaf")

    def finish(self) -> None:
        s = _CFlat()
        s.n_nodes = len(self.node_number)
        s.n_edges = len(self.edge_number)
        s.n_hybrids = len(self.hybrid)
        s.n_leaves = len(self.leaf)
        for a in self._arrays:
            arr = np.ascontiguousarray(getattr(self, a))
            setattr(self, a, arr)
            setattr(s, a, arr.ctypes.data if arr.size else None)
        self.c = s

    def ptr(self):
        return C.byref(self.c)


# ------------------------------------------------------------------------------
# extended newick reader
# ------------------------------------------------------------------------------
class _Parser:
    def __init__(self, s: str):
        self.s = s.strip()
        self.i = 0
        self.net = HybridNetwork()
        self.num_internal = 1          # PN: internal nodes are -2, -3, ... in '(' order
        self.hyb_by_name: Dict[str, Node] = {}
        self.hyb_pending: Dict[str, List[Tuple[Edge, bool]]] = {}

    def peek(self) -> str:
        while self.i < len(self.s) and self.s[self.i].isspace():
            self.i += 1
        return self.s[self.i] if self.i < len(self.s) else ""

    def read_name(self) -> str:
        j = self.i
        while j < len(self.s) and self.s[j] not in "(),:;":
            j += 1
        name = self.s[self.i:j].strip()
        self.i = j
        return name

    def read_floats(self) -> List[Optional[float]]:
        vals: List[Optional[float]] = []
        while self.peek() == ":":
            self.i += 1
            j = self.i
            while j < len(self.s) and self.s[j] not in "(),:;":
                j += 1
            tok = self.s[self.i:j].strip()
            self.i = j
            vals.append(float(tok) if tok else None)
        return vals

    def new_edge(self, length, hybrid=False, gamma=None) -> Edge:
        e = Edge(len(self.net.edge) + 1, -1.0 if length is None else length, hybrid,
                 1.0 if not hybrid else (-1.0 if gamma is None else gamma), ismajor=True)
        self.net.edge.append(e)
        return e

    def subtree(self, parent: Optional[Node]) -> Node:
        """Reads one subtree and attaches it to ``parent`` (None for the root)."""
        net = self.net
        children_done = False
        n: Optional[Node] = None
        if self.peek() == "(":
            self.i += 1
            self.num_internal += 1
            n = Node(-self.num_internal, False)
            while True:
                self.subtree(n)
                c = self.peek()
                if c == ",":
                    self.i += 1
                    continue
                if c == ")":
                    self.i += 1
                    break
                raise SnaqError(f"newick: expected , or ) at position {self.i}")
            children_done = True
        name = self.read_name() if self.peek() not in ":,);" else ""
        vals = self.read_floats()
        length = vals[0] if len(vals) > 0 else None
        gamma = vals[2] if len(vals) > 2 else None
        is_hyb = "#" in name
        if is_hyb:
            hname = name[name.index("#"):]
            if hname in self.hyb_by_name:
                h = self.hyb_by_name[hname]
                if children_done:
                    # the occurrence that carries the subtree: move the children over
                    for e in list(n.edge):
                        e.node[e.node.index(n)] = h
                        h.edge.append(e)
                    n.edge = []
            else:
                if children_done:
                    h = n
                    h.hybrid = True
                    h.name = hname
                    net.names.append(hname)
                    h.number = len(net.names)
                else:
                    net.names.append(hname)
                    h = Node(len(net.names), False, True, hname)
                self.hyb_by_name[hname] = h
                net.push_node(h)
            if parent is None:
                raise SnaqError("newick: the root cannot be a hybrid node")
            e = self.new_edge(length, True, gamma)
            e.node = [h, parent]
            e.ischild1 = True
            h.edge.append(e)
            parent.edge.append(e)
            self.hyb_pending.setdefault(hname, []).append((e, children_done))
            return h
        if not children_done:
            if not name:
                raise SnaqError(f"newick: leaf without a name at position {self.i}")
            net.names.append(name)
            n = Node(len(net.names), True, False, name)
        else:
            n.name = name
        net.push_node(n)
        if parent is not None:
            e = self.new_edge(length)
            e.node = [n, parent]
            n.edge.append(e)
            parent.edge.append(e)
        return n

    def parse(self) -> HybridNetwork:
        root = self.subtree(None)
        if self.peek() != ";":
            raise SnaqError("newick: missing final ;")
        net = self.net
        # major / minor: by gamma when given; else the occurrence that carries the
        # subtree is major (Case H vs Case G of test/test_parameters.jl)
        for hname, lst in self.hyb_pending.items():
            if len(lst) != 2:
                raise SnaqError(f"hybrid node {hname} has {len(lst)} parent edges instead of 2")
            (e1, sub1), (e2, sub2) = lst
            g1, g2 = e1.gamma, e2.gamma
            if g1 == -1.0 and g2 != -1.0:
                g1 = 1.0 - g2
            if g2 == -1.0 and g1 != -1.0:
                g2 = 1.0 - g1
            if g1 == -1.0:  # both missing: defaults 0.9 / 0.1 (src/readwrite.jl:79-85)
                if sub1 or not sub2:
                    g1, g2 = 0.9, 0.1
                else:
                    g1, g2 = 0.1, 0.9
            if abs(g1 + g2 - 1.0) > 1e-6:
                raise SnaqError(f"hybrid edges of {hname} have gammas that do not sum to 1: {g1}, {g2}")
            e1.gamma, e2.gamma = g1, g2
            if g1 == g2:
                e1.ismajor, e2.ismajor = (sub1 or not sub2), not (sub1 or not sub2)
            else:
                e1.ismajor, e2.ismajor = g1 > g2, g2 > g1
        # cleanBL!
        for e in net.edge:
            e.length = _length_clean(e.length)
        # cleanAfterRead!: fuse degree-2 tree nodes (the root in particular)
        for n in list(net.node):
            if not n.leaf and not n.hybrid and len(n.edge) == 2:
                self._fuse(n)
            elif not n.leaf and not n.hybrid and len(n.edge) > 3:
                raise SnaqError(f"polytomy at node {n.number}: resolve it before calling the engine")
        for h in net.hybrid:
            ntree = sum(1 for e in h.edge if not e.hybrid)
            if ntree != 1:
                raise SnaqError(f"hybrid node {h.name} must have exactly one child edge, it has {ntree}")
        net.update_all()
        return net

    def _fuse(self, n: Node) -> None:
        net = self.net
        keep, drop = n.edge[0], n.edge[1]
        a = getOtherNode(keep, n)
        b = getOtherNode(drop, n)
        if drop.hybrid and not keep.hybrid:
            keep, drop, a, b = drop, keep, b, a
        keep.length = min(keep.length + drop.length, 10.0)
        # keep: a -- n  becomes  a -- b
        keep.node[keep.node.index(n)] = b
        b.edge[b.edge.index(drop)] = keep
        net.edge.remove(drop)
        net.node.remove(n)


def readnewicklevel1(s: str) -> HybridNetwork:
    """``readnewicklevel1`` / ``readTopologyUpdate``: src/readwrite.jl:172-194.
    Accepts a newick string or a path to a file holding one."""
    if "(" not in s:
        with open(s) as fh:
            s = fh.read()
    return _Parser(s).parse()


def network_from_lists(nodes: Sequence[Tuple[int, bool, bool, str, Sequence[int]]],
                       edges: Sequence[Tuple[int, float, bool, float, int, int]]) -> HybridNetwork:
    """Build a network by hand, like ``examples/case_f_example.jl``.
    nodes: (number, leaf, hybrid, name, [edge numbers in node.edge order]);
    edges: (number, length, hybrid, gamma, node_a, node_b) in edge.node order, with
    the hybrid (child) node first for hybrid edges.  Call ``update_all()`` after."""
    net = HybridNetwork()
    nby, eby = {}, {}
    for number, length, hybrid, gamma, a, b in edges:
        e = Edge(number, length, hybrid, gamma)
        eby[number] = e
        net.edge.append(e)
    for number, leaf, hybrid, name, elist in nodes:
        n = Node(number, leaf, hybrid, name)
        n.edge = [eby[k] for k in elist]
        nby[number] = n
        net.push_node(n)
    for number, length, hybrid, gamma, a, b in edges:
        e = eby[number]
        e.node = [nby[a], nby[b]]
        e.ischild1 = True
        if hybrid and not nby[a].hybrid:
            if not nby[b].hybrid:
                raise SnaqError(f"hybrid edge {number} does not touch a hybrid node")
            e.ischild1 = False
    net.names = [str(n.number) for n in net.node]
    return net
