"""CPU restatement of ``updateBL!`` — TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).

Follows the reference literally on the flat network arrays (include/snaq_b200.h):
  edgesParts      src/readquartetdata.jl:866-889   the four leaf sets around every internal edge
  getDescendants! src/readquartetdata.jl:893-908   (major edges only; a tree has nothing else)
  makeTable       src/readquartetdata.jl:935-960   one row per data quartet with one taxon in each part
  resolution      src/readquartetdata.jl:966-990   which of 12|34, 13|24, 14|23 the row's CF is
  updateBL!       src/readquartetdata.jl:838-861   edgeL = -log(3/2 (1 - mean CF)); the replacement rule

Parity unpinned by reference golden vectors (the reference has no test of updateBL!); the restatement is the
definition the device reduction is checked against.  Pure Python loops: small cases only.
"""
import itertools
import math

import numpy as np

NODE_LEAF = 0x01
EDGE_MAJOR = 0x02


def _parts(flat):
    nn = len(flat.node_number)
    ne = len(flat.edge_number)
    node_edge = flat.node_edge.reshape(nn, 3)
    edge_node = flat.edge_node.reshape(ne, 2)
    is_leaf = (flat.node_flags & NODE_LEAF) != 0

    def other(e, n):
        a, b = edge_node[e]
        return b if a == n else a

    def descendants(node, edge, out):                     # :893-908
        if is_leaf[node]:
            out.append(node)
        else:
            for e in node_edge[node]:
                if e >= 0 and e != edge and (flat.edge_flags[e] & EDGE_MAJOR):
                    descendants(other(e, node), e, out)

    parts = []
    for e in range(ne):
        n1, n2 = edge_node[e]
        if is_leaf[n1] or is_leaf[n2]:                    # isInternalEdge
            continue
        e1 = [x for x in node_edge[n1] if x >= 0 and x != e]
        e2 = [x for x in node_edge[n2] if x >= 0 and x != e]
        ps = []
        for n, es in ((n1, e1), (n2, e2)):
            for x in es:
                out = []
                descendants(other(x, n), x, out)
                ps.append([int(flat.node_taxon[v]) for v in out])
        parts.append((int(flat.edge_number[e]), ps))
    return parts


def update_bl_table(flat, quartet_taxa, obscf):
    """the table `x` updateBL! returns: {edge number: (Nquartets, edgeL)} (:843-846)"""
    qt = np.asarray(quartet_taxa).reshape(-1, 4)
    obs = np.asarray(obscf, dtype=np.float64).reshape(-1, 3)
    rows = {}
    for i, q in enumerate(qt):
        rows.setdefault(tuple(sorted(int(t) for t in q)), []).append(i)
    out = {}
    for enum, (p1, p2, p3, p4) in _parts(flat):
        cfs = []
        for t1, t2, t3, t4 in itertools.product(p1, p2, p3, p4):
            for r in rows.get(tuple(sorted((t1, t2, t3, t4))), []):
                d = [int(t) for t in qt[r]]
                # resolution(nam, taxon) :966-990: the split t1 t2 | t3 t4 in the data's order of the four taxa
                mate = d.index(t2) if d[0] == t1 else (d.index(t1) if d[0] == t2 else (d.index(t4) if d[0] == t3 else d.index(t3)))
                cfs.append(obs[r, mate - 1])
        if cfs:
            arg = 1.5 * (1.0 - float(np.mean(cfs)))
            out[enum] = (len(cfs), -math.log(arg) if arg > 0.0 else math.inf)   # Julia: -log(0.0) == Inf
    return out


def update_bl(flat, quartet_taxa, obscf):
    """new edge lengths (edge order of the flat network) after updateBL! (:847-858)"""
    tab = update_bl_table(flat, quartet_taxa, obscf)
    length = np.array(flat.edge_length, dtype=np.float64)
    for i, enum in enumerate(flat.edge_number):
        if int(enum) in tab and (length[i] < 0.0 or length[i] == 1.0):
            el = tab[int(enum)][1]
            length[i] = el if el > 0 else 0.0
    length[length < 0.0] = 1.0
    return length, tab
