"""Synthetic inputs for tests and the benchmark: random level-1
networks and CF tables simulated from them.  Everything is seeded.

The generator grows a random rooted binary tree by sequential random edge
insertion, then adds reticulations one at a time between two tree edges whose
connecting path touches no existing cycle (level-1), pointing the new hybrid edge
at an edge that is not an ancestor of the donor (so the network stays rootable).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .network import Edge, HybridNetwork, Node, SnaqError, getOtherNode


def random_level1_network(ntaxa: int, nhybrids: int, seed: int, min_cycle: int = 4,
                          mean_length: float = 1.0, taxa: Optional[Sequence[str]] = None,
                          max_tries: int = 2000) -> HybridNetwork:
    """Random level-1 network with ``ntaxa`` leaves and ``nhybrids`` reticulations.

    min_cycle=4 avoids triangles (bad triangles are only legal with >= 6 taxa and
    very bad ones never are); min_cycle=3 lets them appear so tests can cover them.
    Internal lengths ~ Exp(mean) truncated to [0.01, 5]; external lengths 1.0;
    gamma ~ U(0.05, 0.5) on the minor edge.
    """
    rng = np.random.default_rng(seed)
    if ntaxa < 4:
        raise SnaqError("need at least 4 taxa")
    names = list(taxa) if taxa is not None else [f"t{i + 1}" for i in range(ntaxa)]

    # --- rooted tree as parent/children arrays -------------------------------
    # node ids: 0..ntaxa-1 leaves; internal nodes appended.  root has 2 children.
    parent: List[int] = []
    children: List[List[int]] = []

    def new_node() -> int:
        parent.append(-1)
        children.append([])
        return len(parent) - 1

    for _ in range(ntaxa):
        new_node()
    root = new_node()
    for leaf in (0, 1):
        parent[leaf] = root
        children[root].append(leaf)
    for leaf in range(2, ntaxa):
        # pick a random existing edge (child id identifies the edge above it)
        cands = [v for v in range(len(parent)) if parent[v] >= 0]
        v = int(cands[rng.integers(len(cands))])
        p = parent[v]
        m = new_node()
        children[p][children[p].index(v)] = m
        parent[m] = p
        children[m] = [v, leaf]
        parent[v] = m
        parent[leaf] = m

    def draw_len() -> float:
        return float(min(5.0, max(0.01, rng.exponential(mean_length))))

    length = [1.0 if v < ntaxa else draw_len() for v in range(len(parent))]  # edge above v

    # --- reticulations ----------------------------------------------------------
    # hybrid records: (hyb node, donor node, recipient-parent (major parent))
    in_cycle = [False] * len(parent)
    retic: List[Tuple[int, int]] = []     # (minor parent node p, hybrid node h)
    gammas: List[float] = []
    minor_len: List[float] = []

    def ancestors(v: int) -> List[int]:
        out = []
        while v >= 0:
            out.append(v)
            v = parent[v]
        return out

    for _ in range(nhybrids):
        ok = False
        for _try in range(max_tries):
            cands = [v for v in range(len(parent)) if parent[v] >= 0]
            b = int(cands[rng.integers(len(cands))])       # recipient edge: parent[b] -> b
            d = int(cands[rng.integers(len(cands))])       # donor edge: parent[d] -> d
            if b == d:
                continue
            a = parent[b]
            if a == root:
                continue                                     # keep the (suppressed) root away from hybrid nodes
            # donor must not be inside the subtree below b (no directed cycle)
            if b in ancestors(d):
                continue
            # undirected tree path between a and the new donor node p (on the edge above d)
            anc_d = ancestors(parent[d])
            anc_a = ancestors(a)
            if d in anc_a:                                   # donor edge is above a: a ... d, p
                cyc_nodes = anc_a[: anc_a.index(d) + 1]
            elif a in anc_d:                                 # a is above the donor: p, parent[d] ... a
                cyc_nodes = anc_d[: anc_d.index(a) + 1]
            else:
                sa = set(anc_a)
                lca = next(x for x in anc_d if x in sa)
                cyc_nodes = anc_d[: anc_d.index(lca) + 1] + anc_a[: anc_a.index(lca)]
            k = len(cyc_nodes) + 2                             # + donor node p + hybrid node h
            if root in cyc_nodes:
                k -= 1                                         # the degree-2 root is suppressed when unrooting
            if k < min_cycle:
                continue
            if any(in_cycle[v] for v in cyc_nodes):
                continue
            # subdivide the donor edge with p and the recipient edge with h
            p = new_node()
            in_cycle.append(False)
            length.append(0.0)
            pd = parent[d]
            children[pd][children[pd].index(d)] = p
            parent[p] = pd
            children[p] = [d]
            parent[d] = p
            if d < ntaxa:
                length[p], length[d] = draw_len(), 1.0
            else:
                f = float(rng.uniform(0.2, 0.8))
                length[p], length[d] = max(0.01, length[d] * f), max(0.01, length[d] * (1 - f))
            h = new_node()
            in_cycle.append(False)
            length.append(0.0)
            a = parent[b]
            children[a][children[a].index(b)] = h
            parent[h] = a
            children[h] = [b]
            parent[b] = h
            if b < ntaxa:
                length[h], length[b] = draw_len(), 1.0
            else:
                f = float(rng.uniform(0.2, 0.8))
                length[h], length[b] = max(0.01, length[b] * f), max(0.01, length[b] * (1 - f))
            for v in cyc_nodes + [p, h]:
                in_cycle[v] = True
            retic.append((p, h))
            gammas.append(float(rng.uniform(0.05, 0.5)))
            minor_len.append(draw_len())
            ok = True
            break
        if not ok:
            raise SnaqError(f"could not place reticulation {len(retic) + 1} on a {ntaxa}-taxon tree (seed {seed})")

    # --- build the HybridNetwork (unrooted: suppress the degree-2 root) -------------
    net = HybridNetwork()
    nodes = {}
    hyb_ids = {h for _, h in retic}
    next_internal = 2
    next_hyb = 1000
    for v in range(len(parent)):
        if v == root:
            continue
        if v < ntaxa:
            n = Node(v + 1, True, False, names[v])
        elif v in hyb_ids:
            n = Node(next_hyb, False, True, f"H{next_hyb}")
            next_hyb += 1
        else:
            n = Node(-next_internal, False, False, "")
            next_internal += 1
        nodes[v] = n
    order = list(range(len(parent)))
    rng.shuffle(order)            # net.node order is arbitrary in the reference after moves
    for v in order:
        if v != root:
            net.push_node(nodes[v])
    net.names = [n.name for n in net.node]
    eno = 0

    def add_edge(child: int, par: int, ln: float, hybrid=False, gamma=1.0, major=True) -> Edge:
        nonlocal eno
        eno += 1
        e = Edge(eno, ln, hybrid, gamma, ismajor=major)
        e.node = [nodes[child], nodes[par]]
        e.ischild1 = True
        nodes[child].edge.append(e)
        nodes[par].edge.append(e)
        net.edge.append(e)
        return e

    r1, r2 = children[root]
    for v in range(len(parent)):
        if v == root or parent[v] == root:
            continue
        if v in hyb_ids:
            g = gammas[[h for _, h in retic].index(v)]
            add_edge(v, parent[v], length[v], True, 1.0 - g, True)
        else:
            add_edge(v, parent[v], length[v])
    # the two root edges fuse into one edge r1 -- r2
    if r1 in hyb_ids or r2 in hyb_ids:
        hchild, other = (r1, r2) if r1 in hyb_ids else (r2, r1)
        g = gammas[[h for _, h in retic].index(hchild)]
        add_edge(hchild, other, min(10.0, length[r1] + length[r2]), True, 1.0 - g, True)
    else:
        add_edge(r1, r2, min(10.0, length[r1] + length[r2]) if (r1 >= ntaxa and r2 >= ntaxa) else 1.0)
    for (p, h), g, ml in zip(retic, gammas, minor_len):
        add_edge(h, p, ml, True, g, False)
    rng.shuffle(net.edge)          # net.edge order defines the x layout: exercise arbitrary orders
    for n in net.node:
        rng.shuffle(n.edge)
    if len(net.hybrid) > 1:
        hy = list(net.hybrid)
        rng.shuffle(hy)
        net.hybrid = hy
    lf = list(net.leaf)
    rng.shuffle(lf)
    net.leaf = lf
    net.update_all()
    return net


def all_quartets(ntaxa: int) -> np.ndarray:
    """All C(n,4) four-taxon sets, lexicographic, as uint16 [nq,4] (vectorised)."""
    from math import comb
    nq = comb(ntaxa, 4)
    out = np.empty((nq, 4), dtype=np.uint16)
    pos = 0
    for a in range(ntaxa - 3):
        for b in range(a + 1, ntaxa - 2):
            m = ntaxa - b - 1            # candidates for c<d among b+1..n-1
            cnt = m * (m - 1) // 2
            if cnt <= 0:
                continue
            c, d = np.triu_indices(m, k=1)
            out[pos:pos + cnt, 0] = a
            out[pos:pos + cnt, 1] = b
            out[pos:pos + cnt, 2] = c + b + 1
            out[pos:pos + cnt, 3] = d + b + 1
            pos += cnt
    assert pos == nq
    return out


def obs_from_expcf(expcf: np.ndarray, ngenes: int, seed: int) -> np.ndarray:
    """obsCF = Multinomial(ngenes, expCF)/ngenes (gives exact zeros, which the
    likelihood must skip: src/pseudolik.jl:1398)."""
    rng = np.random.default_rng(seed)
    p = np.clip(np.asarray(expcf, dtype=np.float64), 0.0, None)
    p = p / p.sum(axis=1, keepdims=True)
    counts = rng.multinomial(ngenes, p)
    return counts.astype(np.float64) / float(ngenes)


def parameter_batch(x0: np.ndarray, upper: np.ndarray, P: int, seed: int) -> np.ndarray:
    """Rows 0: x0; 1..2n: x0 +- 0.1*range on one coordinate (BOBYQA's initial set,
    clipped to the bounds); rest: gamma/gammaz ~ U(0,0.5), t log-uniform in [1e-3,10]."""
    rng = np.random.default_rng(seed)
    n = len(x0)
    X = np.empty((P, n), dtype=np.float64)
    X[0] = x0
    r = 1
    for i in range(n):
        for sgn in (+1.0, -1.0):
            if r >= P:
                break
            X[r] = x0
            X[r, i] = min(upper[i], max(0.0, x0[i] + sgn * 0.1 * upper[i]))
            r += 1
    if r < P:
        m = P - r
        isT = upper > 1.5
        U = rng.uniform(0.0, 0.5, size=(m, n))
        T = np.exp(rng.uniform(np.log(1e-3), np.log(10.0), size=(m, n)))
        X[r:] = np.where(isT[None, :], T, U)
    return X


def nni_candidates(net: HybridNetwork):
    """internal tree edges on which :func:`nni_move` can act: both ends tree nodes outside every cycle"""
    out = []
    for e in net.edge:
        if e.hybrid:
            continue
        u, v = e.node
        if u.leaf or v.leaf or u.hybrid or v.hybrid or u.inCycle != -1 or v.inCycle != -1:
            continue
        out.append(e)
    return out


def nni_move(net: HybridNetwork, rng: np.random.Generator) -> Optional[Tuple]:
    """A nearest-neighbour interchange around a random internal tree edge away from the cycles, in place (the tree part
    of the reference's NNI move, src/moves_snaq.jl:1209-1330): one subtree hanging off each end of the edge is swapped.
    Returns a token for :func:`nni_undo` (the search's `undoMove`), or None when the network has no such edge."""
    cands = nni_candidates(net)
    if not cands:
        return None
    e = cands[int(rng.integers(len(cands)))]
    u, v = e.node
    a = [x for x in u.edge if x is not e][int(rng.integers(2))]
    c = [x for x in v.edge if x is not e][int(rng.integers(2))]
    _swap_ends(a, u, c, v)
    net.update_all()
    return (a, u, c, v)


def _swap_ends(a: Edge, u: Node, c: Edge, v: Node) -> None:
    """edge a leaves node u for v, edge c leaves v for u (positions in node.edge and edge.node are kept)"""
    a.node[a.node.index(u)] = v
    c.node[c.node.index(v)] = u
    iu, iv = u.edge.index(a), v.edge.index(c)
    u.edge[iu], v.edge[iv] = c, a


def nni_undo(net: HybridNetwork, token: Tuple) -> None:
    a, u, c, v = token
    _swap_ends(a, v, c, u)
    net.update_all()
