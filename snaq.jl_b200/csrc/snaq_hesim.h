/*
 * snaq_hesim.h — hasEdge / indexht for ANY level-1 network (nested cycles included).
 *
 * q.qnet.hasEdge (updateHasEdge!, src/pseudolik.jl:392-464) is defined operationally: "the edge number still exists in
 * the pruned quartet network and is still identifiable there".  With one cycle that is a structural property of the
 * quartet (snaq_hasedge.h); with a cycle inside a block of another cycle it depends on the ORDER in which
 * extractQuartet! deletes the other leaves (net.leaf order): whether the degree-2 remnants of an inner cycle have been
 * merged when the last leaf around them goes decides which of deleteLeaf!'s branches fire (src/pseudolik.jl:158-196,
 * :241-251, :305-355).  So for networks with two or more hybrid nodes this file does what the reference does, as an
 * integer-only simulation on a per-thread copy of the network: delete the n-4 other leaves one by one
 * (deleteLeaf!, :152-357, with deleteIntLeafWhile! :63-81, removeWeirdNodes! :695-735 and the bad-diamond-I rewrite
 * :206-239), then redundantCycle! (:559-612) and cleanExtEdges! (:636-643), and read the surviving edges off.
 * Branch lengths play no role (they never decide a branch of the pruning), so the state is adjacency + flags:
 * about 16 bytes per node + edge.  It is a parity read-back (nq x nparams bytes), off the hot path.
 *
 * SNAQ_HD like snaq_core.h: tests/hostsim runs the same code on the CPU against the oracle.
 */
#ifndef SNAQ_HESIM_H
#define SNAQ_HESIM_H

#include "snaq_core.h"

#define SNAQ_SIM_MAX 512            /* nodes / edges of the network (about 250 taxa) */
/* node flags */
#define SIM_N_LEAF 1u
#define SIM_N_HYB 2u
#define SIM_N_HASHYB 4u
#define SIM_N_BD1 8u
#define SIM_N_BD2 16u
#define SIM_N_BADTRI 32u
#define SIM_N_ALIVE 64u
/* edge flags */
#define SIM_E_HYB 1u
#define SIM_E_MAJOR 2u
#define SIM_E_IDENT 8u
#define SIM_E_FROMBD1 16u
#define SIM_E_INCYC 32u
#define SIM_E_ALIVE 64u

/* the network as the simulation starts from it (device memory; node / edge indices are positions in net.node / net.edge) */
typedef struct SnaqSimNet {
    int32_t n_nodes, n_edges, n_leaves, n_hybrids, nparams, nh, nt;
    const int16_t *node_edge;     /* [n_nodes][3] node.edge order, -1 padded */
    const uint8_t *node_flags;    /* [n_nodes]                                */
    const int16_t *node_incycle;  /* [n_nodes] INDEX of the cycle's hybrid node, or -1 */
    const int16_t *node_taxon;    /* [n_nodes] taxon id of a leaf, else -1   */
    const int16_t *edge_node;     /* [n_edges][2] edge.node order             */
    const uint8_t *edge_flags;    /* [n_edges]                                */
    const int32_t *edge_number;   /* [n_edges] PN edge.number                 */
    const int32_t *node_number;   /* [n_nodes] PN node.number                 */
    const int16_t *leaf_order;    /* [n_leaves] net.leaf                      */
    const int16_t *hyb_order;     /* [n_hybrids] net.hybrid                   */
    /* parameters: slot s of x is ... */
    const int16_t *slot_ref;      /* gamma: hybrid node; t: edge; gammaz: hybrid node (two consecutive slots: ind 1, 2)      */
    const int16_t *slot_alias;    /* t: 2*(hybrid node)+(ind-1) of a gammaz pseudo number equal to the edge's number, or -1;
                                     gammaz: an original edge that carries that pseudo number, or -1       */
} SnaqSimNet;

typedef struct SnaqSim {
    int16_t ne3[SNAQ_SIM_MAX][3];   /* node.edge */
    uint8_t nne[SNAQ_SIM_MAX];
    uint8_t nfl[SNAQ_SIM_MAX];
    int16_t ncy[SNAQ_SIM_MAX];
    int16_t en2[SNAQ_SIM_MAX][2];   /* edge.node */
    uint8_t enn[SNAQ_SIM_MAX];
    uint8_t efl[SNAQ_SIM_MAX];
    int16_t ren[SNAQ_SIM_MAX];      /* edge renamed by the bad-diamond-I rewrite: 2*(hybrid node)+(ind-1), else -1 */
    int nN, nE, err;
} SnaqSim;

#define SIM_FAIL(s) do { (s).err = __LINE__; return; } while (0)
#define SIM_FAILV(s, v) do { (s).err = __LINE__; return (v); } while (0)

SNAQ_HD int sim_other(const SnaqSim &g, int e, int n) { return g.en2[e][0] == n ? g.en2[e][1] : g.en2[e][0]; }
SNAQ_HD bool sim_leaf(const SnaqSim &g, int n) { return (g.nfl[n] & SIM_N_LEAF) != 0; }
SNAQ_HD bool sim_hyb(const SnaqSim &g, int n) { return (g.nfl[n] & SIM_N_HYB) != 0; }
SNAQ_HD bool sim_alive(const SnaqSim &g, int n) { return (g.nfl[n] & SIM_N_ALIVE) != 0; }
SNAQ_HD bool sim_ehyb(const SnaqSim &g, int e) { return (g.efl[e] & SIM_E_HYB) != 0; }

SNAQ_HD void sim_refresh(SnaqSim &g, int n) {                       /* hasHybEdge, as PN's setEdge! / removeEdge! maintain it */
    bool any = false;
    for (int i = 0; i < g.nne[n]; i++) any |= sim_ehyb(g, g.ne3[n][i]);
    g.nfl[n] = (uint8_t)((g.nfl[n] & ~SIM_N_HASHYB) | (any ? SIM_N_HASHYB : 0u));
}
SNAQ_HD void sim_removeEdge(SnaqSim &g, int n, int e) {             /* PN removeEdge!(node, edge) */
    int idx = -1;
    for (int i = 0; i < g.nne[n]; i++) if (g.ne3[n][i] == e) { idx = i; break; }
    if (idx < 0) SIM_FAIL(g);
    for (int i = idx; i + 1 < g.nne[n]; i++) g.ne3[n][i] = g.ne3[n][i + 1];
    g.nne[n]--;
    sim_refresh(g, n);
}
SNAQ_HD void sim_setEdge(SnaqSim &g, int n, int e) {                /* PN setEdge!(node, edge) */
    if (g.nne[n] >= 3) SIM_FAIL(g);
    g.ne3[n][g.nne[n]++] = (int16_t)e;
    sim_refresh(g, n);
}
SNAQ_HD void sim_removeNode(SnaqSim &g, int n, int e) {             /* PN removeNode!(node, edge) */
    int idx = -1;
    for (int i = 0; i < g.enn[e]; i++) if (g.en2[e][i] == n) { idx = i; break; }
    if (idx < 0) SIM_FAIL(g);
    for (int i = idx; i + 1 < g.enn[e]; i++) g.en2[e][i] = g.en2[e][i + 1];
    g.enn[e]--;
}
SNAQ_HD void sim_setNode(SnaqSim &g, int e, int n) {                /* PN setNode!(edge, node) */
    if (g.enn[e] >= 2) SIM_FAIL(g);
    g.en2[e][g.enn[e]++] = (int16_t)n;
    if (sim_leaf(g, n)) g.efl[e] &= (uint8_t)~SIM_E_IDENT;          /* an edge attached to a leaf is never identifiable */
}
SNAQ_HD void sim_deleteNode(SnaqSim &g, int n) {                    /* deleteNode!: src/auxiliary.jl:43-58 */
    if (!sim_alive(g, n)) SIM_FAIL(g);
    g.nfl[n] &= (uint8_t)~SIM_N_ALIVE;
}
SNAQ_HD void sim_deleteEdge(SnaqSim &g, int e) {                    /* deleteEdge!: src/auxiliary.jl:70-76 */
    if (!(g.efl[e] & SIM_E_ALIVE)) SIM_FAIL(g);
    g.efl[e] &= (uint8_t)~SIM_E_ALIVE;
}
/* PN hybridEdges(node): (major, minor, tree) at a hybrid node; (hybrid, tree edge in cycle, tree edge not in cycle) at a
 * tree node with hasHybEdge; else node.edge in order with a leaf-attached edge moved last */
SNAQ_HD void sim_hybridEdges(SnaqSim &g, int n, int &a, int &b, int &c) {
    a = b = c = -1;
    if (g.nne[n] != 3) SIM_FAIL(g);
    const int16_t *ed = g.ne3[n];
    if (sim_hyb(g, n)) {
        for (int i = 0; i < 3; i++) {
            const unsigned f = g.efl[ed[i]];
            if ((f & SIM_E_HYB) && (f & SIM_E_MAJOR)) a = ed[i];
            if ((f & SIM_E_HYB) && !(f & SIM_E_MAJOR)) b = ed[i];
            if (!(f & SIM_E_HYB)) c = ed[i];
        }
    } else if (g.nfl[n] & SIM_N_HASHYB) {
        for (int i = 0; i < 3; i++) {
            const unsigned f = g.efl[ed[i]];
            if (f & SIM_E_HYB) a = ed[i];
            if (!(f & SIM_E_HYB) && (f & SIM_E_INCYC)) b = ed[i];
            if (!(f & SIM_E_HYB) && !(f & SIM_E_INCYC)) c = ed[i];
        }
    } else {
        for (int i = 0; i < 3; i++) {
            if (sim_leaf(g, sim_other(g, ed[i], n))) {
                if (i == 0) { a = ed[1]; b = ed[2]; c = ed[0]; }
                else if (i == 1) { a = ed[0]; b = ed[2]; c = ed[1]; }
                else { a = ed[0]; b = ed[1]; c = ed[2]; }
                return;
            }
        }
        a = ed[0]; b = ed[1]; c = ed[2];
    }
}
SNAQ_HD void sim_hybridEdges2(SnaqSim &g, int n, int e, int &e1, int &e2) {   /* PN hybridEdges(node, edge) */
    e1 = e2 = -1;
    if (g.nne[n] != 3) SIM_FAIL(g);
    for (int i = 0; i < 3; i++)
        if (g.ne3[n][i] != e) { if (e1 < 0) e1 = g.ne3[n][i]; else e2 = g.ne3[n][i]; }
}
SNAQ_HD bool sim_isEdgeIdentifiable(SnaqSim &g, int e) {            /* src/update.jl:282-310 */
    if (sim_ehyb(g, e)) {
        const int node = sim_hyb(g, g.en2[e][0]) ? g.en2[e][0] : g.en2[e][1];
        if (!sim_hyb(g, node)) SIM_FAILV(g, false);
        int a, b, c;
        sim_hybridEdges(g, node, a, b, c);
        if (g.err || c < 0) return false;
        return !sim_leaf(g, sim_other(g, c, node));
    }
    const int na = g.en2[e][0], nb = g.en2[e][1];
    if (!sim_leaf(g, na) && !sim_leaf(g, nb)) {
        if (!sim_hyb(g, na) && !sim_hyb(g, nb) && !(g.efl[e] & SIM_E_FROMBD1)) return true;
        if (sim_hyb(g, na) || sim_hyb(g, nb)) {
            const int h = sim_hyb(g, na) ? na : nb;
            return !(g.nfl[h] & SIM_N_BD2) && !(g.nfl[h] & SIM_N_BADTRI);
        }
        return false;
    }
    return false;
}
SNAQ_HD void sim_makeNodeTree(SnaqSim &g, int h) {                  /* src/pseudolik.jl:12-24 */
    if (!sim_hyb(g, h)) SIM_FAIL(g);
    g.nfl[h] &= (uint8_t)~(SIM_N_BD1 | SIM_N_BD2 | SIM_N_BADTRI | SIM_N_HYB | SIM_N_HASHYB);
}
SNAQ_HD void sim_makeEdgeTree(SnaqSim &g, int e, int node) {        /* src/pseudolik.jl:37-56 */
    if (!sim_ehyb(g, e) || !sim_hyb(g, node)) SIM_FAIL(g);
    g.efl[e] = (uint8_t)((g.efl[e] & ~(SIM_E_HYB | SIM_E_INCYC)) | SIM_E_MAJOR);
    sim_refresh(g, sim_other(g, e, node));
    const bool id = sim_isEdgeIdentifiable(g, e);
    g.efl[e] = (uint8_t)((g.efl[e] & ~SIM_E_IDENT) | (id ? SIM_E_IDENT : 0u));
}
SNAQ_HD int sim_deleteIntLeaf(SnaqSim &g, int middle, int leaf) {   /* src/pseudolik.jl:90-114 */
    if (g.nne[middle] != 2) SIM_FAILV(g, middle);
    int leafedge, otheredge;
    if (sim_other(g, g.ne3[middle][0], middle) == leaf) { leafedge = g.ne3[middle][0]; otheredge = g.ne3[middle][1]; }
    else if (sim_other(g, g.ne3[middle][1], middle) == leaf) { leafedge = g.ne3[middle][1]; otheredge = g.ne3[middle][0]; }
    else SIM_FAILV(g, middle);
    const int othernode = sim_other(g, otheredge, middle);
    sim_removeNode(g, middle, leafedge);
    sim_removeEdge(g, othernode, otheredge);
    sim_setNode(g, leafedge, othernode);
    sim_setEdge(g, othernode, leafedge);
    sim_deleteNode(g, middle);
    sim_deleteEdge(g, otheredge);
    return othernode;
}
SNAQ_HD int sim_deleteIntLeafWhile(SnaqSim &g, int middle, int leaf) {   /* src/pseudolik.jl:63-68 */
    while (!g.err && g.nne[middle] == 2) middle = sim_deleteIntLeaf(g, middle, leaf);
    return middle;
}
SNAQ_HD int sim_removeNoLeafWhile(SnaqSim &g, int n) {              /* src/pseudolik.jl:615-633 */
    while (!g.err && !sim_leaf(g, n) && g.nne[n] == 1) {
        const int e = g.ne3[n][0];
        const int node = sim_other(g, e, n);
        sim_removeEdge(g, node, e);
        sim_deleteNode(g, n);
        sim_deleteEdge(g, e);
        n = node;
    }
    return n;
}
/* removeLoneHybrid!: src/pseudolik.jl:649-693; returns up to two nodes left with one edge */
SNAQ_HD int sim_removeLoneHybrid(SnaqSim &g, int n, int out[2]) {
    if (!sim_hyb(g, n) || g.nne[n] != 2) SIM_FAILV(g, 0);
    const int ea = g.ne3[n][0], eb = g.ne3[n][1];
    if (!(sim_ehyb(g, ea) && sim_ehyb(g, eb))) SIM_FAILV(g, 0);
    const int o[2] = {sim_other(g, ea, n), sim_other(g, eb, n)};
    sim_removeEdge(g, o[0], ea);
    sim_removeEdge(g, o[1], eb);
    sim_deleteEdge(g, ea);
    sim_deleteEdge(g, eb);
    sim_deleteNode(g, n);
    for (int k = 0; k < 2; k++) {
        if (g.nne[o[k]] == 2 && sim_alive(g, o[k])) {
            const int l1 = sim_other(g, g.ne3[o[k]][0], o[k]);
            const int lf = sim_leaf(g, l1) ? l1 : sim_other(g, g.ne3[o[k]][1], o[k]);
            if (sim_leaf(g, lf)) sim_deleteIntLeafWhile(g, o[k], lf);
        }
    }
    int cnt = 0;
    for (int k = 0; k < 2; k++) if (g.nne[o[k]] == 1) { if (sim_leaf(g, o[k])) SIM_FAILV(g, 0); out[cnt++] = o[k]; }
    return cnt;
}
SNAQ_HD void sim_removeWeirdNodes(SnaqSim &g, int n0) {             /* src/pseudolik.jl:695-735 */
    int list[16], nl = 0;
    list[nl++] = n0;
    while (nl > 0 && !g.err) {
        int n = list[--nl];
        int res[2], nres = 0;
        const bool lone1 = g.nne[n] == 1 && !sim_leaf(g, n);
        const bool lone2 = sim_hyb(g, n) && g.nne[n] == 2;
        if (lone1 || lone2) {                                       /* removeWeirdNode! */
            if (g.nne[n] == 1 && !sim_leaf(g, n)) n = sim_removeNoLeafWhile(g, n);
            if (sim_hyb(g, n) && g.nne[n] == 2) {
                if (!(sim_ehyb(g, g.ne3[n][0]) && sim_ehyb(g, g.ne3[n][1]))) SIM_FAIL(g);
                nres = sim_removeLoneHybrid(g, n, res);
            }
        }
        for (int k = 0; k < nres && nl < 16; k++) list[nl++] = res[k];
    }
}
/* the bad-diamond-I rewrite of deleteLeaf! (src/pseudolik.jl:206-239 and :256-286): other = tree node whose leaf goes,
 * ehyb = hybrid edge other -- hyb, etree = cycle tree edge at other, otherT = its far end */
SNAQ_HD void sim_deleteLeafBD1(SnaqSim &g, int other, int hyb, int ehyb, int etree, int otherT, bool firstBranch) {
    int emaj, emin, etr;
    sim_hybridEdges(g, hyb, emaj, emin, etr);
    if (g.err) return;
    const int edge4 = (ehyb == emaj) ? emin : emaj;
    const int other3 = sim_other(g, edge4, hyb);
    const int ind = (sim_other(g, emaj, hyb) == other3) ? 1 : 2;
    int x, edge3, edge5;
    sim_hybridEdges(g, other3, x, edge3, edge5);
    if (g.err || edge3 < 0 || edge5 < 0) { if (!g.err) g.err = __LINE__; return; }
    const int leaf5 = sim_other(g, edge5, other3);
    g.efl[etree] |= SIM_E_FROMBD1;
    sim_removeNode(g, other, etree);
    sim_removeEdge(g, hyb, ehyb);
    sim_setNode(g, etree, hyb);
    sim_setEdge(g, hyb, etree);
    g.ren[etree] = (int16_t)(2 * hyb + (ind - 1));                  /* edge.number = "<hybrid number><ind>" */
    sim_makeEdgeTree(g, edge4, hyb);
    sim_makeNodeTree(g, hyb);
    sim_removeEdge(g, otherT, edge3);
    sim_removeEdge(g, other3, edge3);
    sim_deleteNode(g, other);
    sim_deleteEdge(g, ehyb);
    sim_deleteEdge(g, edge3);
    sim_removeNode(g, other3, edge4);
    sim_removeEdge(g, leaf5, edge5);
    sim_setNode(g, edge4, leaf5);
    sim_setEdge(g, leaf5, edge4);
    sim_deleteNode(g, other3);
    sim_deleteEdge(g, edge5);
    if (firstBranch) g.nfl[hyb] &= (uint8_t)~SIM_N_HASHYB;
}
SNAQ_HD void sim_deleteLeaf(SnaqSim &g, int leaf) {                 /* src/pseudolik.jl:152-357 */
    if (!sim_leaf(g, leaf) || !sim_alive(g, leaf) || g.nne[leaf] != 1) SIM_FAIL(g);
    const int le = g.ne3[leaf][0];
    int other = sim_other(g, le, leaf);
    if (sim_hyb(g, other)) {                                        /* :158-196: the hybridisation vanishes */
        int edge1, edge2;
        sim_hybridEdges2(g, other, le, edge1, edge2);
        if (g.err || !(sim_ehyb(g, edge1) && sim_ehyb(g, edge2))) { if (!g.err) g.err = __LINE__; return; }
        const int o[2] = {sim_other(g, edge1, other), sim_other(g, edge2, other)};
        sim_removeEdge(g, o[0], edge1);
        sim_removeEdge(g, o[1], edge2);
        sim_deleteEdge(g, edge1);
        sim_deleteEdge(g, edge2);
        sim_deleteEdge(g, le);
        sim_deleteNode(g, other);
        sim_deleteNode(g, leaf);
        for (int k = 0; k < 2; k++) {
            if (g.nne[o[k]] == 2 && sim_alive(g, o[k])) {
                const int node = sim_other(g, g.ne3[o[k]][0], o[k]);
                const int leaf1 = sim_leaf(g, node) ? node : sim_other(g, g.ne3[o[k]][1], o[k]);
                if (sim_leaf(g, leaf1)) sim_deleteIntLeafWhile(g, o[k], leaf1);
            }
        }
        for (int k = 0; k < 2; k++) {
            if (g.nne[o[k]] == 1 && sim_alive(g, o[k])) {
                if (sim_leaf(g, o[k])) SIM_FAIL(g);
                sim_removeWeirdNodes(g, o[k]);
            }
        }
    } else if (g.nfl[other] & SIM_N_HASHYB) {                       /* :198-304: a parent of a hybrid node */
        int edge1, edge2;
        sim_hybridEdges2(g, other, le, edge1, edge2);
        if (g.err || !(sim_ehyb(g, edge1) || sim_ehyb(g, edge2))) { if (!g.err) g.err = __LINE__; return; }
        const int other1 = sim_other(g, edge1, other), other2 = sim_other(g, edge2, other);
        if (sim_hyb(g, other1)) {
            if (g.nfl[other1] & SIM_N_BD1) sim_deleteLeafBD1(g, other, other1, edge1, edge2, other2, true);
            else {
                sim_removeEdge(g, other, le);
                int a, b, c;
                sim_hybridEdges(g, other1, a, b, c);
                if (g.err || c < 0) { if (!g.err) g.err = __LINE__; return; }
                if (sim_leaf(g, sim_other(g, c, other1))) {         /* :241-251 */
                    sim_removeEdge(g, other2, edge2);
                    sim_removeNode(g, other, edge1);
                    sim_setNode(g, edge1, other2);
                    sim_setEdge(g, other2, edge1);
                    sim_deleteEdge(g, edge2);
                    sim_deleteNode(g, other);
                }
            }
            sim_deleteNode(g, leaf);
            sim_deleteEdge(g, le);
        } else if (sim_hyb(g, other2)) {
            if (g.nfl[other2] & SIM_N_BD1) sim_deleteLeafBD1(g, other, other2, edge2, edge1, other1, false);
            else {
                sim_removeEdge(g, other, le);
                int a, b, c;
                sim_hybridEdges(g, other2, a, b, c);
                if (g.err || c < 0) { if (!g.err) g.err = __LINE__; return; }
                if (sim_leaf(g, sim_other(g, c, other2))) {         /* :288-298 */
                    sim_removeEdge(g, other1, edge1);
                    sim_removeNode(g, other, edge2);
                    sim_setNode(g, edge2, other1);
                    sim_setEdge(g, other1, edge2);
                    sim_deleteEdge(g, edge1);
                    sim_deleteNode(g, other);
                }
            }
            sim_deleteNode(g, leaf);
            sim_deleteEdge(g, le);
        } else SIM_FAIL(g);
    } else {                                                        /* :305-355: plain tree node */
        if (g.nne[other] == 2) {
            const int edge1 = (g.ne3[other][0] == le) ? g.ne3[other][1] : g.ne3[other][0];
            int other1 = sim_other(g, edge1, other);
            sim_removeEdge(g, other1, edge1);
            sim_removeNode(g, other1, edge1);
            sim_deleteNode(g, other);
            sim_deleteNode(g, leaf);
            sim_deleteEdge(g, le);
            sim_deleteEdge(g, edge1);
            if (g.nne[other1] == 2 && sim_alive(g, other1)) {
                const int node = sim_other(g, g.ne3[other1][0], other1);
                const int leaf1 = sim_leaf(g, node) ? node : sim_other(g, g.ne3[other1][1], other1);
                if (sim_leaf(g, leaf1)) other1 = sim_deleteIntLeafWhile(g, other1, leaf1);
            }
            if (sim_hyb(g, other1) && g.nne[other1] == 2) sim_removeWeirdNodes(g, other1);
            if (g.nne[other1] == 1 && sim_alive(g, other1)) {
                if (sim_leaf(g, other1)) SIM_FAIL(g);
                sim_removeWeirdNodes(g, other1);
            }
        } else {
            int edge1, edge2;
            sim_hybridEdges2(g, other, le, edge1, edge2);
            if (g.err) return;
            const int other1 = sim_other(g, edge1, other), other2 = sim_other(g, edge2, other);
            sim_removeEdge(g, other, le);
            sim_deleteNode(g, leaf);
            sim_deleteEdge(g, le);
            if (sim_leaf(g, other1) || sim_leaf(g, other2)) {
                if (sim_leaf(g, other1) && sim_leaf(g, other2)) SIM_FAIL(g);
                const int newleaf = sim_leaf(g, other1) ? other1 : other2;
                const int middle = sim_deleteIntLeafWhile(g, other, newleaf);
                if (!g.err && sim_hyb(g, middle)) {
                    int a, b, c;
                    sim_hybridEdges(g, middle, a, b, c);
                    if (g.err || a < 0 || b < 0) { if (!g.err) g.err = __LINE__; return; }
                    g.efl[a] &= (uint8_t)~SIM_E_IDENT;
                    g.efl[b] &= (uint8_t)~SIM_E_IDENT;
                }
            }
        }
    }
}
/* redundantCycle!(net, node): src/pseudolik.jl:559-597 */
SNAQ_HD void sim_redundantCycleNode(SnaqSim &g, int n) {
    if (!sim_hyb(g, n)) SIM_FAIL(g);
    int ea, eb, ec;
    sim_hybridEdges(g, n, ea, eb, ec);
    if (g.err || ea < 0 || eb < 0 || !sim_ehyb(g, ea) || !sim_ehyb(g, eb)) { if (!g.err) g.err = __LINE__; return; }
    int other = sim_other(g, ea, n);
    if (g.nne[other] == 2) {
        int e = ea, guard = 0;
        while (g.nne[other] == 2) {
            const int ind = (e == g.ne3[other][0]) ? 1 : 0;
            e = g.ne3[other][ind];
            other = sim_other(g, e, other);
            if (++guard > SNAQ_SIM_MAX) SIM_FAIL(g);
        }
        if (other == n) {
            const int n1 = sim_other(g, ea, n);
            sim_deleteIntLeafWhile(g, n1, n);
            if (g.err) return;
            const int edge = sim_ehyb(g, g.ne3[n][0]) ? g.ne3[n][0] : g.ne3[n][1];
            if (g.en2[edge][0] == g.en2[edge][1]) {
                const int n3 = sim_other(g, ec, n);
                sim_removeEdge(g, n3, ec);
                sim_deleteNode(g, n);
                sim_deleteEdge(g, edge);
                sim_deleteEdge(g, ec);
                if (!sim_leaf(g, n3) && g.nne[n3] == 1) sim_removeNoLeafWhile(g, n3);
            }
        }
    }
}

/*
 * hasEdge bits of one quartet (cleared here) and the "keeps a bad diamond I whole" flag, by simulating extractQuartet!.
 * S: scratch state (thread-private).
 */
/* zpos (may be NULL): for every gammaz slot (nparams - nh - nt entries) the position in net.edge of the edge that carries
 * it in qnet, or -1: parameters!(qnet,net) lists these slots in qnet.edge order (src/snaq_optimization.jl:150-160),
 * which differs from slot order when two pruned bad diamonds I are left in one quartet network */
/* extractQuartet!(net, quartet) on the integer state: the un-merged q.qnet is what is alive in g afterwards */
SNAQ_HD int snaq_sim_prune(const SnaqSimNet &net, const uint16_t tx[4], SnaqSim &g);

SNAQ_HD int snaq_hasedge_sim(const SnaqSimNet &net, const uint16_t tx[4], SnaqSim &g, uint32_t *bits, int *bd1_type5, int16_t *zpos = nullptr) {
    const int N = net.n_nodes, E = net.n_edges;
    const int nw = (net.nparams + 31) / 32;
    for (int i = 0; i < nw; i++) bits[i] = 0;
    if (zpos) for (int i = net.nh + net.nt; i < net.nparams; i++) zpos[i - net.nh - net.nt] = -1;
    *bd1_type5 = 0;
    const int rc = snaq_sim_prune(net, tx, g);
    if (rc) return rc;
    (void)N;
    /* updateHasEdge!: src/pseudolik.jl:392-464.  An edge number is looked up among the surviving edges: the edge itself if
     * it was not renamed, or an edge that the bad-diamond-I rewrite gave the same number (first in net.edge order) */
    auto alive_e = [&](int e) { return e >= 0 && (g.efl[e] & SIM_E_ALIVE); };
    auto renamed_to = [&](int code) {                               /* surviving edge renamed to pseudo number `code`, or -1 */
        for (int e = 0; e < E; e++) if ((g.efl[e] & SIM_E_ALIVE) && g.ren[e] == code) return e;
        return -1;
    };
    auto internalNonId = [&](int e) { return !(g.efl[e] & SIM_E_IDENT) && !sim_leaf(g, g.en2[e][0]) && !sim_leaf(g, g.en2[e][1]); };
    auto setbit = [&](int s) { bits[s >> 5] |= 1u << (s & 31); };
    for (int s = 0; s < net.nparams; s++) {
        const int ref = net.slot_ref[s];
        if (s < net.nh) {                                           /* gamma: the hybrid node is still a hybrid of qnet */
            if (sim_alive(g, ref) && sim_hyb(g, ref)) setbit(s);
        } else if (s < net.nh + net.nt) {
            int found = (alive_e(ref) && g.ren[ref] < 0) ? ref : -1;
            if (net.slot_alias[s] >= 0) {
                const int r = renamed_to(net.slot_alias[s]);
                if (r >= 0 && (found < 0 || r < found)) found = r;
            }
            if (found >= 0 && (g.efl[found] & SIM_E_IDENT)) setbit(s);
        } else {                                                    /* gammaz pair of hybrid node ref: slots s (ind 1), s+1 (ind 2) */
            if (sim_alive(g, ref) && sim_hyb(g, ref)) { setbit(s); setbit(s + 1); }
            else {
                int idx[2];
                for (int k = 0; k < 2; k++) {
                    int f = renamed_to(2 * ref + k);
                    const int al = net.slot_alias[s + k];
                    if (alive_e(al) && g.ren[al] < 0 && (f < 0 || al < f)) f = al;
                    idx[k] = f;
                }
                if (idx[0] >= 0 && idx[1] < 0) { if (internalNonId(idx[0])) { setbit(s); if (zpos) zpos[s - net.nh - net.nt] = (int16_t)idx[0]; } }
                else if (idx[1] >= 0 && idx[0] < 0) { if (internalNonId(idx[1])) { setbit(s + 1); if (zpos) zpos[s + 1 - net.nh - net.nt] = (int16_t)idx[1]; } }
                else if (idx[0] >= 0 && idx[1] >= 0 && internalNonId(idx[0]) && internalNonId(idx[1])) return SNAQ_B_BLOCKS;
            }
            s++;
        }
    }
    for (int i = 0; i < net.n_hybrids; i++) {
        const int h = net.hyb_order[i];
        if (sim_alive(g, h) && sim_hyb(g, h) && (g.nfl[h] & SIM_N_BD1)) { *bd1_type5 = 1; break; }
    }
    return SNAQ_B_OK;
}

/* the edge numbers of q.qnet.edge, in order (sampleEdgeQuartetWeighted step 2, src/moves_snaq.jl:1451-1453); an edge
 * renamed by the bad-diamond-I rewrite carries "<hybrid node number><ind>".  Returns the count (<= cap written) or -1 */
SNAQ_HD int snaq_quartet_edges_sim(const SnaqSimNet &net, const uint16_t tx[4], SnaqSim &g, int32_t *out, int cap) {
    if (snaq_sim_prune(net, tx, g)) return -1;
    int n = 0;
    for (int e = 0; e < net.n_edges; e++) {
        if (!(g.efl[e] & SIM_E_ALIVE)) continue;
        int32_t num = net.edge_number[e];
        if (g.ren[e] >= 0) {
            const int32_t base = net.node_number[g.ren[e] >> 1], ind = (g.ren[e] & 1) + 1;
            num = base >= 0 ? base * 10 + ind : base * 10 - ind;    /* parse(Int, string(number) * string(ind)) */
        }
        if (n < cap) out[n] = num;
        n++;
    }
    return n;
}

SNAQ_HD int snaq_sim_prune(const SnaqSimNet &net, const uint16_t tx[4], SnaqSim &g) {
    const int N = net.n_nodes, E = net.n_edges;
    if (N > SNAQ_SIM_MAX || E > SNAQ_SIM_MAX) return SNAQ_B_OVERFLOW;
    g.nN = N; g.nE = E; g.err = 0;
    for (int n = 0; n < N; n++) {
        int ne = 0;
        for (int j = 0; j < 3; j++) { g.ne3[n][j] = net.node_edge[3 * n + j]; ne += net.node_edge[3 * n + j] >= 0; }
        g.nne[n] = (uint8_t)ne;
        g.nfl[n] = (uint8_t)(net.node_flags[n] | SIM_N_ALIVE);
        g.ncy[n] = net.node_incycle[n];
    }
    for (int e = 0; e < E; e++) {
        g.en2[e][0] = net.edge_node[2 * e]; g.en2[e][1] = net.edge_node[2 * e + 1];
        g.enn[e] = 2;
        g.efl[e] = (uint8_t)(net.edge_flags[e] | SIM_E_ALIVE);
        g.ren[e] = -1;
    }
    /* extractQuartet: src/pseudolik.jl:472-486 */
    for (int i = 0; i < net.n_leaves && !g.err; i++) {
        const int lf = net.leaf_order[i];
        const int t = net.node_taxon[lf];
        if (t == (int)tx[0] || t == (int)tx[1] || t == (int)tx[2] || t == (int)tx[3]) continue;
        sim_deleteLeaf(g, lf);
    }
    /* redundantCycle!(qnet): :599-612, hasRedundantCycle :544-556 */
    for (int guard = 0; !g.err; guard++) {
        int red = -1;
        for (int i = 0; i < net.n_hybrids && red < 0; i++) {
            const int h = net.hyb_order[i];
            if (!sim_alive(g, h) || !sim_hyb(g, h)) continue;
            int k = 0;
            for (int n = 0; n < N; n++) k += (sim_alive(g, n) && g.ncy[n] == h && g.nne[n] == 3);
            if (k < 2) red = h;
        }
        if (red < 0) break;
        sim_redundantCycleNode(g, red);
        if (guard > 64) { g.err = __LINE__; break; }
    }
    /* cleanExtEdges!: :636-643 */
    for (int i = 0; i < net.n_leaves && !g.err; i++) {
        const int lf = net.leaf_order[i];
        if (!sim_alive(g, lf)) continue;
        if (g.nne[lf] != 1) { g.err = __LINE__; break; }
        sim_deleteIntLeafWhile(g, sim_other(g, g.ne3[lf][0], lf), lf);
    }
    return g.err ? SNAQ_B_BLOCKS : SNAQ_B_OK;
}

#endif /* SNAQ_HESIM_H */
