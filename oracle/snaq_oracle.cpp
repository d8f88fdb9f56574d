/*
 * snaq_oracle.cpp — CPU ORACLE (test infrastructure, NOT the product).
 *
 * A literal C++ restatement of the reference's quartet-pseudolikelihood path:
 * every quartet gets its own copy of the network, the n-4 other leaves are
 * deleted one at a time by pointer surgery, the pruned 4-taxon network is
 * classified, its cycles are eliminated, and expCF / logPseudoLik are computed,
 * exactly in the order the reference does it.  Each function cites the reference
 * lines it follows (paths relative to the reference repo, SNaQ.jl v1.1.2).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (snaq.jl_b200/csrc)
 * never does.  The product computes descriptors *structurally* on the GPU, so
 * agreement between the two is a real cross-check, not a tautology.
 *
 * Parity pinning: this restatement is checked against every golden vector the
 * reference's own tests hold for the path (tests/test_oracle_golden.py; list in
 * SURVEY.md §8c).  The reference itself (Julia + PhyloNetworks.jl, compat "1",
 * unpinned) cannot run in the build container; the PhyloNetworks primitives used
 * here (getOtherNode, hybridEdges, setNode!/setEdge!/removeNode!/removeEdge!,
 * setgamma!, approxEq) are restated from their documented meaning and pinned by
 * those golden vectors.
 */
#include "../include/snaq_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct OErr : std::runtime_error {
    int code;
    OErr(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] void fail(int code, const std::string &m) { throw OErr(code, m); }
#define ERR(msg) fail(SNAQ_ETOPOLOGY, std::string(msg))

// PhyloNetworks approxEq(a,b,absTol=1e-5,relTol=100): |a-b|<absTol if either is
// < eps, else |a-b| < relTol*eps(|a|+|b|).  (PN auxiliary; SURVEY.md App. B.)
inline double eps_of(double x) {
    x = std::fabs(x);
    return std::nextafter(x, INFINITY) - x;
}
inline bool approxEq(double a, double b) {
    const double e = 2.220446049250313e-16;
    if (a < e || b < e) return std::fabs(a - b) < 1e-5;
    return std::fabs(a - b) < 100.0 * eps_of(std::fabs(a) + std::fabs(b));
}

struct Node {
    int number = 0;
    bool leaf = false, hybrid = false, hasHybEdge = false;
    bool bd1 = false, bd2 = false, badtri = false;
    int inCycle = -1, k = -1, typeHyb = -1;
    double gammaz = -1.0;
    int ne = 0;
    int edge[4] = {-1, -1, -1, -1};
    int prev = -1;
    int taxon = -1;
};
struct Edge {
    int number = 0;
    double length = 0, y = 0, z = 0, gamma = 1.0;
    bool hybrid = false, ismajor = true, ident = false, fromBD1 = false;
    int inCycle = -1;
    int nn = 0;
    int node[2] = {-1, -1};
};

// One network object serves as HybridNetwork and QuartetNetwork
// (src/types.jl:169-196): stable storage N/E plus the ordered lists net.node,
// net.edge, net.hybrid, net.leaf holding ids into the storage.
struct Net {
    std::vector<Node> N;
    std::vector<Edge> E;
    std::vector<int> node, edge, hybrid, leaf;
    int numhybrids = 0;
    int numBad = 0;
    // QuartetNetwork extras
    int which = -1;
    std::vector<int> typeHyb;
    double t1 = -1.0;
    int split[4] = {-1, -1, -1, -1};
    int formula[3] = {-1, -1, -1};
    double expCF[3] = {0, 0, 0};
    int qtaxon[4] = {-1, -1, -1, -1};
    std::vector<char> hasEdge;
    std::vector<int> indexht;  // 1-based, like the reference
    std::vector<int> index;    // 0-based positions in q.edge (gamma, t, gammaz edges) or q.node (bad diamond I)
};

// ---------------- PhyloNetworks primitives (restated) -----------------------
inline int getOtherNode(const Net &g, int e, int n) {
    const Edge &ed = g.E[e];
    return ed.node[0] == n ? ed.node[1] : ed.node[0];
}
void refreshHasHybEdge(Net &g, int n) {
    bool any = false;
    for (int i = 0; i < g.N[n].ne; i++) any |= g.E[g.N[n].edge[i]].hybrid;
    g.N[n].hasHybEdge = any;
}
void removeEdge(Net &g, int n, int e) {  // PN removeEdge!(node,edge)
    Node &nd = g.N[n];
    int idx = -1;
    for (int i = 0; i < nd.ne; i++)
        if (nd.edge[i] == e) { idx = i; break; }
    if (idx < 0) ERR("removeEdge!: edge " + std::to_string(g.E[e].number) + " not in node " + std::to_string(nd.number));
    for (int i = idx; i + 1 < nd.ne; i++) nd.edge[i] = nd.edge[i + 1];
    nd.ne--;
    refreshHasHybEdge(g, n);
}
void setEdge(Net &g, int n, int e) {  // PN setEdge!(node,edge)
    Node &nd = g.N[n];
    if (nd.ne >= 4) ERR("setEdge!: node with more than 4 edges");
    nd.edge[nd.ne++] = e;
    refreshHasHybEdge(g, n);
}
void removeNode(Net &g, int n, int e) {  // PN removeNode!(node,edge)
    Edge &ed = g.E[e];
    int idx = -1;
    for (int i = 0; i < ed.nn; i++)
        if (ed.node[i] == n) { idx = i; break; }
    if (idx < 0) ERR("removeNode!: node not in edge");
    for (int i = idx; i + 1 < ed.nn; i++) ed.node[i] = ed.node[i + 1];
    ed.nn--;
}
void setNode(Net &g, int e, int n) {  // PN setNode!(edge,node)
    Edge &ed = g.E[e];
    if (ed.nn >= 2) ERR("setNode!: vector of nodes already has 2 values");
    ed.node[ed.nn++] = n;
    if (g.N[n].leaf) ed.ident = false;  // an edge attached to a leaf is never identifiable
}
template <class V> int findByNumberNode(const Net &g, const V &list, int number) {
    for (size_t i = 0; i < list.size(); i++)
        if (g.N[list[i]].number == number) return (int)i;
    return -1;
}
bool isNodeNumIn(const Net &g, int n, const std::vector<int> &list) {
    return findByNumberNode(g, list, g.N[n].number) >= 0;
}
int getIndexEdgeByNumber(const Net &g, int number) {
    for (size_t i = 0; i < g.edge.size(); i++)
        if (g.E[g.edge[i]].number == number) return (int)i;
    return -1;
}
// deleteNode!(net::QuartetNetwork, n): src/auxiliary.jl:43-58
void deleteNode(Net &g, int n) {
    int idx = findByNumberNode(g, g.node, g.N[n].number);
    if (idx < 0) ERR("Node " + std::to_string(g.N[n].number) + " not in quartet network");
    g.node.erase(g.node.begin() + idx);
    if (g.N[n].hybrid) {  // removeHybrid!(net,n)
        int ih = findByNumberNode(g, g.hybrid, g.N[n].number);
        if (ih < 0) ERR("hybrid node not in net.hybrid");
        g.hybrid.erase(g.hybrid.begin() + ih);
        g.numhybrids--;
    }
    if (g.N[n].leaf) {
        auto it = std::find(g.leaf.begin(), g.leaf.end(), n);
        if (it == g.leaf.end()) ERR("node not in net.leaf");
        g.leaf.erase(it);
    }
}
// deleteEdge!(net::QuartetNetwork, e): src/auxiliary.jl:70-76
void deleteEdge(Net &g, int e) {
    int idx = getIndexEdgeByNumber(g, g.E[e].number);
    if (idx < 0) ERR("edge not in quartet network");
    g.edge.erase(g.edge.begin() + idx);
}
// update_yz!/setLength!: src/auxiliary.jl:108-129
void setLength(Net &g, int e, double len, bool negative = false) {
    if (!(negative || len >= 0))
        fail(SNAQ_ERANGE, "length has to be nonnegative: " + std::to_string(len));
    if (!(len >= -0.4054651081081644))
        fail(SNAQ_ENUMERIC, "length can be negative, but not too negative (greater than -log(1.5))");
    if (len > 10.0) len = 10.0;
    Edge &ed = g.E[e];
    ed.length = len;
    ed.y = std::exp(-len);
    ed.z = 1.0 - ed.y;
}
struct E3 { int a, b, c; };
// PN hybridEdges(node): (major,minor,tree) for a hybrid node; (hybrid, tree edge
// in cycle, tree edge not in cycle) for a tree node with hasHybEdge; else
// node.edge in order with a leaf-attached edge moved last.  -1 = nothing.
E3 hybridEdges(const Net &g, int n) {
    const Node &nd = g.N[n];
    if (nd.ne != 3) ERR("node " + std::to_string(nd.number) + " has " + std::to_string(nd.ne) + " edges instead of 3");
    if (nd.hybrid) {
        int maj = -1, mino = -1, tree = -1;
        for (int i = 0; i < 3; i++) {
            const Edge &e = g.E[nd.edge[i]];
            if (e.hybrid && e.ismajor) maj = nd.edge[i];
            if (e.hybrid && !e.ismajor) mino = nd.edge[i];
            if (!e.hybrid) tree = nd.edge[i];
        }
        return {maj, mino, tree};
    } else if (nd.hasHybEdge) {
        int hyb = -1, tc = -1, tree = -1;
        for (int i = 0; i < 3; i++) {
            const Edge &e = g.E[nd.edge[i]];
            if (e.hybrid) hyb = nd.edge[i];
            if (!e.hybrid && e.inCycle != -1) tc = nd.edge[i];
            if (!e.hybrid && e.inCycle == -1) tree = nd.edge[i];
        }
        return {hyb, tc, tree};
    } else {
        for (int i = 0; i < 3; i++) {
            if (g.N[getOtherNode(g, nd.edge[i], n)].leaf) {
                if (i == 0) return {nd.edge[1], nd.edge[2], nd.edge[0]};
                if (i == 1) return {nd.edge[0], nd.edge[2], nd.edge[1]};
                return {nd.edge[0], nd.edge[1], nd.edge[2]};
            }
        }
        return {nd.edge[0], nd.edge[1], nd.edge[2]};
    }
}
// PN hybridEdges(node, edge): the two other edges, in node.edge order
void hybridEdges2(const Net &g, int n, int e, int &e1, int &e2) {
    const Node &nd = g.N[n];
    if (nd.ne != 3) ERR("node " + std::to_string(nd.number) + " has " + std::to_string(nd.ne) + " edges instead of 3");
    e1 = e2 = -1;
    for (int i = 0; i < 3; i++)
        if (nd.edge[i] != e) { if (e1 < 0) e1 = nd.edge[i]; else e2 = nd.edge[i]; }
}

// isEdgeIdentifiable: src/update.jl:282-310
bool isEdgeIdentifiable(const Net &g, int e) {
    const Edge &ed = g.E[e];
    if (ed.hybrid) {
        int node = g.N[ed.node[0]].hybrid ? ed.node[0] : ed.node[1];
        if (!g.N[node].hybrid) ERR("hybrid edge pointing at tree node");
        E3 h = hybridEdges(g, node);
        return !g.N[getOtherNode(g, h.c, node)].leaf;
    }
    const Node &a = g.N[ed.node[0]], &b = g.N[ed.node[1]];
    if (!a.leaf && !b.leaf) {
        if (!a.hybrid && !b.hybrid && !ed.fromBD1) return true;
        if (a.hybrid || b.hybrid) {
            const Node &h = a.hybrid ? a : b;
            return !h.bd2 && !h.badtri;
        }
        return false;  // Julia falls through returning nothing (falsy use)
    }
    return false;
}

// makeNodeTree!: src/pseudolik.jl:12-24
void makeNodeTree(Net &g, int h) {
    Node &nd = g.N[h];
    if (!nd.hybrid) ERR("cannot make node tree node because it already is");
    nd.gammaz = -1.0;
    nd.bd1 = nd.bd2 = nd.badtri = false;
    nd.k += 1;
    int ih = findByNumberNode(g, g.hybrid, nd.number);  // removeHybrid!(net,hybrid)
    if (ih < 0) ERR("hybrid node not in net.hybrid");
    g.hybrid.erase(g.hybrid.begin() + ih);
    g.numhybrids--;
    nd.hybrid = false;
    nd.hasHybEdge = false;
}
// makeEdgeTree!: src/pseudolik.jl:37-56
void makeEdgeTree(Net &g, int e, int node) {
    Edge &ed = g.E[e];
    if (!ed.hybrid) ERR("cannot make edge tree because it is tree already");
    if (!g.N[node].hybrid) ERR("need the hybrid node at which edge is pointing to");
    ed.hybrid = false;
    ed.ismajor = true;
    ed.gamma = 1.0;
    ed.inCycle = -1;
    int other = getOtherNode(g, e, node);
    refreshHasHybEdge(g, other);
    ed.ident = isEdgeIdentifiable(g, e);
}

// deleteIntLeaf!(net,middle,leaf,negative): src/pseudolik.jl:90-114
int deleteIntLeaf(Net &g, int middle, int leaf, bool negative) {
    Node &m = g.N[middle];
    if (m.ne != 2) ERR("internal node does not have two edges only");
    int leafedge, otheredge;
    if (getOtherNode(g, m.edge[0], middle) == leaf) { leafedge = m.edge[0]; otheredge = m.edge[1]; }
    else if (getOtherNode(g, m.edge[1], middle) == leaf) { leafedge = m.edge[1]; otheredge = m.edge[0]; }
    else ERR("leaf is not attached to internal node by any edge");
    int othernode = getOtherNode(g, otheredge, middle);
    setLength(g, leafedge, g.E[leafedge].length + g.E[otheredge].length, negative);
    removeNode(g, middle, leafedge);
    removeEdge(g, othernode, otheredge);
    setNode(g, leafedge, othernode);
    setEdge(g, othernode, leafedge);
    deleteNode(g, middle);
    deleteEdge(g, otheredge);
    return othernode;
}
// deleteIntLeafWhile!(net,middle,leaf,negative): src/pseudolik.jl:63-68
int deleteIntLeafWhileNode(Net &g, int middle, int leaf, bool negative = false) {
    while (g.N[middle].ne == 2) middle = deleteIntLeaf(g, middle, leaf, negative);
    return middle;
}
// deleteIntLeafWhile!(net,leafedge,leaf,negative): src/pseudolik.jl:76-81
void deleteIntLeafWhileEdge(Net &g, int leafedge, int leaf, bool negative = false) {
    int middle = getOtherNode(g, leafedge, leaf);
    while (g.N[middle].ne == 2) middle = deleteIntLeaf(g, middle, leaf, negative);
}

// removeNoLeaf!/removeNoLeafWhile!: src/pseudolik.jl:615-633
int removeNoLeaf(Net &g, int n) {
    if (g.N[n].leaf) ERR("node is a leaf, so we cannot remove it");
    if (g.N[n].ne != 1) ERR("node does not have 1 edge, so we do not have to remove it");
    int e = g.N[n].edge[0];
    int node = getOtherNode(g, e, n);
    removeEdge(g, node, e);
    deleteNode(g, n);
    deleteEdge(g, e);
    return node;
}
int removeNoLeafWhile(Net &g, int n) {
    while (!g.N[n].leaf && g.N[n].ne == 1) n = removeNoLeaf(g, n);
    return n;
}
// removeLoneHybrid!: src/pseudolik.jl:649-693
std::vector<int> removeLoneHybrid(Net &g, int n) {
    Node &nd = g.N[n];
    if (!nd.hybrid) ERR("cannot remove a lone hybrid if it is not hybrid");
    if (nd.ne != 2) ERR("hybrid node should have only 2 edges to be deleted");
    int ea = nd.edge[0], eb = nd.edge[1];
    if (!(g.E[ea].hybrid && g.E[eb].hybrid)) ERR("both edges have to be hybrid");
    int other1 = getOtherNode(g, ea, n), other2 = getOtherNode(g, eb, n);
    removeEdge(g, other1, ea);
    removeEdge(g, other2, eb);
    deleteEdge(g, ea);
    deleteEdge(g, eb);
    deleteNode(g, n);
    for (int o : {other1, other2}) {
        if (g.N[o].ne == 2 && isNodeNumIn(g, o, g.node)) {
            int leaf1 = getOtherNode(g, g.N[o].edge[0], o);
            int leaf = g.N[leaf1].leaf ? leaf1 : getOtherNode(g, g.N[o].edge[1], o);
            if (g.N[leaf].leaf) deleteIntLeafWhileNode(g, o, leaf);
        }
    }
    std::vector<int> out;
    if (g.N[other1].ne == 1) { if (g.N[other1].leaf) ERR("should not be a leaf"); out.push_back(other1); }
    if (g.N[other2].ne == 1) { if (g.N[other2].leaf) ERR("should not be a leaf"); out.push_back(other2); }
    return out;
}
// removeWeirdNode!/removeWeirdNodes!: src/pseudolik.jl:695-735
std::vector<int> removeWeirdNode(Net &g, int n) {
    std::vector<int> other;
    if (g.N[n].ne == 1 && !g.N[n].leaf) n = removeNoLeafWhile(g, n);
    if (g.N[n].hybrid && g.N[n].ne == 2) {
        if (!(g.E[g.N[n].edge[0]].hybrid && g.E[g.N[n].edge[1]].hybrid)) ERR("two edges for this hybrid node must be hybrid");
        other = removeLoneHybrid(g, n);
    }
    return other;
}
void removeWeirdNodes(Net &g, int n0) {
    std::vector<int> list{n0};
    while (!list.empty()) {
        int n = list.back();
        list.pop_back();
        std::vector<int> res;
        if (g.N[n].ne == 1 && !g.N[n].leaf) res = removeWeirdNode(g, n);
        else if (g.N[n].hybrid && g.N[n].ne == 2) {
            if (!(g.E[g.N[n].edge[0]].hybrid && g.E[g.N[n].edge[1]].hybrid)) ERR("hybrid node only has two edges and they must be hybrids");
            res = removeWeirdNode(g, n);
        }  // else: "@error calling removeWeirdNode on normal node" (log only)
        for (int l : res) list.push_back(l);
    }
}

// the bad-diamond-I rewrite shared by the two mirrored branches of deleteLeaf!
// (src/pseudolik.jl:206-239 and :256-286).  other = tree node whose leaf is being
// removed, ehyb = hybrid edge other--hyb, etree = cycle tree edge at other.
void deleteLeafBadDiamondI(Net &g, int other, int hyb, int ehyb, int etree, int otherT, bool firstBranch) {
    E3 he = hybridEdges(g, hyb);
    int edgemaj = he.a, edgemin = he.b;
    int edge4 = (ehyb == edgemaj) ? edgemin : edgemaj;
    int other3 = getOtherNode(g, edge4, hyb);
    int ind = (getOtherNode(g, edgemaj, hyb) == other3) ? 1 : 2;
    E3 h3 = hybridEdges(g, other3);
    int edge3 = h3.b, edge5 = h3.c;
    int leaf5 = getOtherNode(g, edge5, other3);
    g.E[etree].fromBD1 = true;
    removeNode(g, other, etree);
    removeEdge(g, hyb, ehyb);
    setNode(g, etree, hyb);
    setEdge(g, hyb, etree);
    if (g.N[other3].gammaz == -1) ERR("hybrid node is bad diamond, but gammaz is not well updated");
    setLength(g, etree, -std::log(1 - g.N[other3].gammaz));
    g.E[etree].number = std::stoi(std::to_string(g.N[hyb].number) + std::to_string(ind));
    makeEdgeTree(g, edge4, hyb);
    makeNodeTree(g, hyb);
    removeEdge(g, otherT, edge3);
    removeEdge(g, other3, edge3);
    deleteNode(g, other);
    deleteEdge(g, ehyb);
    deleteEdge(g, edge3);
    removeNode(g, other3, edge4);
    removeEdge(g, leaf5, edge5);
    setNode(g, edge4, leaf5);
    setEdge(g, leaf5, edge4);
    setLength(g, edge4, g.E[edge5].length);
    deleteNode(g, other3);
    deleteEdge(g, edge5);
    if (firstBranch) g.N[hyb].hasHybEdge = false;
}

// deleteLeaf!(net,leaf): src/pseudolik.jl:152-357
void deleteLeaf(Net &g, int leaf) {
    if (!g.N[leaf].leaf) ERR("node is not a leaf, cannot delete it");
    if (!isNodeNumIn(g, leaf, g.leaf)) ERR("node is not in net.leaf, cannot delete it");
    if (g.N[leaf].ne != 1) ERR("strange leaf with not 1 edge");
    const int le = g.N[leaf].edge[0];
    int other = getOtherNode(g, le, leaf);
    if (g.N[other].hybrid) {  // :158-196
        int edge1, edge2;
        hybridEdges2(g, other, le, edge1, edge2);
        if (!(g.E[edge1].hybrid && g.E[edge2].hybrid)) ERR("hybrid node does not have two hybrid edges");
        int other1 = getOtherNode(g, edge1, other), other2 = getOtherNode(g, edge2, other);
        removeEdge(g, other1, edge1);
        removeEdge(g, other2, edge2);
        deleteEdge(g, edge1);
        deleteEdge(g, edge2);
        deleteEdge(g, le);
        deleteNode(g, other);
        deleteNode(g, leaf);
        for (int o : {other1, other2}) {
            if (g.N[o].ne == 2 && isNodeNumIn(g, o, g.node)) {
                int node = getOtherNode(g, g.N[o].edge[0], o);
                int leaf1 = g.N[node].leaf ? node : getOtherNode(g, g.N[o].edge[1], o);
                if (g.N[leaf1].leaf) deleteIntLeafWhileNode(g, o, leaf1);
            }
        }
        for (int o : {other1, other2}) {
            if (g.N[o].ne == 1 && isNodeNumIn(g, o, g.node)) {
                if (g.N[o].leaf) ERR("node was attached no hybrid edge, cannot be a leaf");
                removeWeirdNodes(g, o);
            }
        }
    } else if (g.N[other].hasHybEdge) {  // :198-304
        int edge1, edge2;
        hybridEdges2(g, other, le, edge1, edge2);
        if (!(g.E[edge1].hybrid || g.E[edge2].hybrid)) ERR("node has hybrid edge attribute true, but the edges are not hybrid");
        int other1 = getOtherNode(g, edge1, other), other2 = getOtherNode(g, edge2, other);
        if (g.N[other1].hybrid) {
            if (g.N[other1].bd1) {
                deleteLeafBadDiamondI(g, other, other1, edge1, edge2, other2, true);
            } else {
                removeEdge(g, other, le);
                E3 h = hybridEdges(g, other1);
                if (g.N[getOtherNode(g, h.c, other1)].leaf) {
                    setLength(g, edge1, g.E[edge2].length + g.E[edge1].length);
                    removeEdge(g, other2, edge2);
                    removeNode(g, other, edge1);
                    setNode(g, edge1, other2);
                    setEdge(g, other2, edge1);
                    deleteEdge(g, edge2);
                    deleteNode(g, other);
                }
            }
            deleteNode(g, leaf);
            deleteEdge(g, le);
        } else if (g.N[other2].hybrid) {
            if (g.N[other2].bd1) {
                deleteLeafBadDiamondI(g, other, other2, edge2, edge1, other1, false);
            } else {
                removeEdge(g, other, le);
                E3 h = hybridEdges(g, other2);
                if (g.N[getOtherNode(g, h.c, other2)].leaf) {
                    setLength(g, edge2, g.E[edge2].length + g.E[edge1].length);
                    removeEdge(g, other1, edge1);
                    removeNode(g, other, edge2);
                    setNode(g, edge2, other1);
                    setEdge(g, other1, edge2);
                    deleteEdge(g, edge1);
                    deleteNode(g, other);
                }
            }
            deleteNode(g, leaf);
            deleteEdge(g, le);
        } else {
            ERR("node has hybrid edge, but neither of the other nodes are hybrid");
        }
    } else {  // tree node without hybrid edges, :305-355
        if (g.N[other].ne == 2) {
            int edge1 = (g.N[other].edge[0] == le) ? g.N[other].edge[1] : g.N[other].edge[0];
            int other1 = getOtherNode(g, edge1, other);
            removeEdge(g, other1, edge1);
            removeNode(g, other1, edge1);
            deleteNode(g, other);
            deleteNode(g, leaf);
            deleteEdge(g, le);
            deleteEdge(g, edge1);
            if (g.N[other1].ne == 2 && isNodeNumIn(g, other1, g.node)) {
                int node = getOtherNode(g, g.N[other1].edge[0], other1);
                int leaf1 = g.N[node].leaf ? node : getOtherNode(g, g.N[other1].edge[1], other1);
                if (g.N[leaf1].leaf) other1 = deleteIntLeafWhileNode(g, other1, leaf1);
            }
            if (g.N[other1].hybrid && g.N[other1].ne == 2) removeWeirdNodes(g, other1);
            if (g.N[other1].ne == 1 && isNodeNumIn(g, other1, g.node)) {
                if (g.N[other1].leaf) ERR("node was attached no leaf edge, cannot be a leaf");
                removeWeirdNodes(g, other1);
            }
        } else {
            int edge1, edge2;
            hybridEdges2(g, other, le, edge1, edge2);
            int other1 = getOtherNode(g, edge1, other), other2 = getOtherNode(g, edge2, other);
            removeEdge(g, other, le);
            deleteNode(g, leaf);
            deleteEdge(g, le);
            if (g.N[other1].leaf || g.N[other2].leaf) {
                if (g.N[other1].leaf && g.N[other2].leaf) ERR("just deleted a leaf and its two attached nodes are leaves also");
                int newleaf = g.N[other1].leaf ? other1 : other2;
                int middle = deleteIntLeafWhileNode(g, other, newleaf);
                if (g.N[middle].hybrid) {
                    E3 h = hybridEdges(g, middle);
                    g.E[h.a].ident = false;
                    g.E[h.b].ident = false;
                }
            }
        }
    }
}

// hasRedundantCycle / redundantCycle! / cleanExtEdges!: src/pseudolik.jl:544-643
int hasRedundantCycle(const Net &g) {
    if ((int)g.hybrid.size() != g.numhybrids) ERR("length net.hybrid != net.numhybrids");
    for (int h : g.hybrid) {
        int k = 0;
        for (int n : g.node)
            if (g.N[n].inCycle == g.N[h].number && g.N[n].ne == 3) k++;
        if (k < 2) return h;
    }
    return -1;
}
void redundantCycleNode(Net &g, int n) {
    if (!g.N[n].hybrid) ERR("cannot clean a cycle on a tree node");
    E3 ed = hybridEdges(g, n);
    if (!(ed.a >= 0 && ed.b >= 0 && g.E[ed.a].hybrid && g.E[ed.b].hybrid)) ERR("hybrid node does not have two hybrid edges");
    int other = getOtherNode(g, ed.a, n);
    if (g.N[other].ne == 2) {
        int e = ed.a;
        while (g.N[other].ne == 2) {
            int ind = (e == g.N[other].edge[0]) ? 1 : 0;
            e = g.N[other].edge[ind];
            other = getOtherNode(g, e, other);
        }
        if (other == n) {
            int n1 = getOtherNode(g, ed.a, n);
            deleteIntLeafWhileNode(g, n1, n);
            int edge = g.E[g.N[n].edge[0]].hybrid ? g.N[n].edge[0] : g.N[n].edge[1];
            if (g.E[edge].node[0] == g.E[edge].node[1]) {
                int n3 = getOtherNode(g, ed.c, n);
                removeEdge(g, n3, ed.c);
                deleteNode(g, n);
                deleteEdge(g, edge);
                deleteEdge(g, ed.c);
                if (!g.N[n3].leaf && g.N[n3].ne == 1) removeNoLeafWhile(g, n3);
            }
        }
    }
}
void cleanExtEdges(Net &g) {
    std::vector<int> leaves = g.leaf;  // iteration over net.leaf (not modified inside)
    for (int l : leaves) {
        if (g.N[l].ne != 1) ERR("leaf with not 1 edge");
        deleteIntLeafWhileNode(g, getOtherNode(g, g.N[l].edge[0], l), l);
    }
}
void redundantCycle(Net &g) {
    if (!g.hybrid.empty()) {
        int guard = 0;
        int node = hasRedundantCycle(g);
        while (node >= 0) {
            redundantCycleNode(g, node);
            node = hasRedundantCycle(g);
            if (++guard > 1000) ERR("redundantCycle! does not terminate");
        }
    }
    cleanExtEdges(g);
}

// parameters(net): src/snaq_optimization.jl:54-92
struct Params {
    std::vector<double> x;
    std::vector<int> numht, index;  // index: 0-based position in net.edge / net.node
    int nh = 0, nt = 0, nhz = 0;
};
Params parameters(const Net &g) {
    std::vector<double> t, h, hz;
    std::vector<int> n, n2, hzn, it, ih, ihz;
    for (size_t i = 0; i < g.edge.size(); i++) {
        const Edge &e = g.E[g.edge[i]];
        if (e.ident) { t.push_back(e.length); n.push_back(e.number); it.push_back((int)i); }
        if (e.hybrid && !e.ismajor) {
            int node = g.N[e.node[0]].hybrid ? e.node[0] : e.node[1];
            if (!g.N[node].hybrid) ERR("strange thing, hybrid edge pointing at tree node");
            if (!g.N[node].bd1) { h.push_back(e.gamma); n2.push_back(e.number); ih.push_back((int)i); }
            else {
                E3 he = hybridEdges(g, node);
                int oa = getOtherNode(g, he.a, node), ob = getOtherNode(g, he.b, node);
                hz.push_back(g.N[oa].gammaz);
                hz.push_back(g.N[ob].gammaz);
                hzn.push_back(std::stoi(std::to_string(g.N[node].number) + "1"));
                hzn.push_back(std::stoi(std::to_string(g.N[node].number) + "2"));
                ihz.push_back((int)(std::find(g.node.begin(), g.node.end(), oa) - g.node.begin()));
                ihz.push_back((int)(std::find(g.node.begin(), g.node.end(), ob) - g.node.begin()));
            }
        }
    }
    Params p;
    p.nh = (int)h.size(); p.nt = (int)t.size(); p.nhz = (int)hz.size();
    p.x = h; p.x.insert(p.x.end(), t.begin(), t.end()); p.x.insert(p.x.end(), hz.begin(), hz.end());
    p.numht = n2; p.numht.insert(p.numht.end(), n.begin(), n.end()); p.numht.insert(p.numht.end(), hzn.begin(), hzn.end());
    p.index = ih; p.index.insert(p.index.end(), it.begin(), it.end()); p.index.insert(p.index.end(), ihz.begin(), ihz.end());
    return p;
}

// updateParameters!(net,x): src/snaq_optimization.jl:255-271 (setgamma! with
// changeOther=true gives the partner 1-gamma; ismajor labels are kept because
// every formula downstream is symmetric under the major/minor swap).
void applyParameters(Net &g, const Params &p, const double *x) {
    int np = (int)p.x.size();
    for (int i = 0; i < np; i++) {
        if (i < p.nh) {
            if (!(0 <= x[i] && x[i] <= 1)) fail(SNAQ_ERANGE, "new gamma value should be between 0,1: " + std::to_string(x[i]));
            int e = g.edge[p.index[i]];
            if (!g.E[e].hybrid) ERR("optimizing gamma for tree edge");
            int node = g.N[g.E[e].node[0]].hybrid ? g.E[e].node[0] : g.E[e].node[1];
            E3 he = hybridEdges(g, node);
            int partner = (he.a == e) ? he.b : he.a;
            g.E[e].gamma = x[i];
            g.E[partner].gamma = 1.0 - x[i];
        } else if (i < p.nh + p.nt) {
            setLength(g, g.edge[p.index[i]], x[i]);
        } else {
            if (!(0 <= x[i] && x[i] <= 1)) fail(SNAQ_ERANGE, "new gammaz value should be between 0,1: " + std::to_string(x[i]));
            g.N[g.node[p.index[i]]].gammaz = x[i];
        }
    }
}

// updateHasEdge!(qnet,net): src/pseudolik.jl:392-464
void updateHasEdge(Net &q, const Net &net) {
    std::vector<char> edges, h, hz;
    for (int ei : net.edge) {
        const Edge &e = net.E[ei];
        if (e.ident) {
            int idx = getIndexEdgeByNumber(q, e.number);
            edges.push_back(idx >= 0 && q.E[q.edge[idx]].ident);
        }
        if (e.hybrid && !e.ismajor) {
            int node = net.N[e.node[0]].hybrid ? e.node[0] : e.node[1];
            const Node &nd = net.N[node];
            bool inq = findByNumberNode(q, q.hybrid, nd.number) >= 0;
            if (!nd.bd1) h.push_back(inq);
            else if (inq) { hz.push_back(1); hz.push_back(1); }
            else {
                int ind1 = std::stoi(std::to_string(nd.number) + "1");
                int ind2 = std::stoi(std::to_string(nd.number) + "2");
                int i1 = getIndexEdgeByNumber(q, ind1), i2 = getIndexEdgeByNumber(q, ind2);
                auto internalNonId = [&](int idx) {
                    const Edge &qe = q.E[q.edge[idx]];
                    return !qe.ident && !q.N[qe.node[0]].leaf && !q.N[qe.node[1]].leaf;
                };
                if (i1 < 0 && i2 < 0) { hz.push_back(0); hz.push_back(0); }
                else if (i1 >= 0 && i2 < 0) { hz.push_back(internalNonId(i1)); hz.push_back(0); }
                else if (i2 >= 0 && i1 < 0) { hz.push_back(0); hz.push_back(internalNonId(i2)); }
                else {
                    if (internalNonId(i1) && internalNonId(i2)) ERR("strange qnet: it has two gammaz edges");
                    // reference pushes nothing in this branch (:451-457)
                }
            }
        }
    }
    q.hasEdge = h;
    q.hasEdge.insert(q.hasEdge.end(), edges.begin(), edges.end());
    q.hasEdge.insert(q.hasEdge.end(), hz.begin(), hz.end());
}

// parameters!(qnet,net): src/snaq_optimization.jl:106-180.  indexht = slots of x
// (1-based like the reference); index = positions inside qnet.edge / qnet.node.
void parametersQ(Net &q, const Net &net, const Params &p) {
    int nh = net.numhybrids - net.numBad;
    int ntIdent = p.nt;
    auto findIn = [&](int number, int lo, int hi) {
        for (int i = lo; i < hi; i++) if (p.numht[i] == number) return i - lo;
        return -1;
    };
    auto posNode = [&](int n) { return (int)(std::find(q.node.begin(), q.node.end(), n) - q.node.begin()); };
    std::vector<int> qnh, qnt, qnhz, xh, xt, xhz;
    int np = (int)p.numht.size();
    bool found = false;
    for (int hn : q.hybrid) {
        if (q.N[hn].bd1) {
            int ind1 = std::stoi(std::to_string(q.N[hn].number) + "1");
            int i = findIn(ind1, nh + ntIdent, np);
            if (i < 0) ERR("ind1 not found in nhz");
            E3 he = hybridEdges(q, hn);
            qnhz.push_back(i + 1 + nh + ntIdent);
            qnhz.push_back(i + 2 + nh + ntIdent);
            xhz.push_back(posNode(getOtherNode(q, he.a, hn)));
            xhz.push_back(posNode(getOtherNode(q, he.b, hn)));
            found = true;
            break;
        }
    }
    if (!found) {
        for (size_t pos = 0; pos < q.edge.size(); pos++) {
            const Edge &e = q.E[q.edge[pos]];
            if (e.ident) {
                int i = findIn(e.number, nh, nh + ntIdent);
                if (i < 0) ERR("identifiable edge " + std::to_string(e.number) + " in qnet not found in net");
                qnt.push_back(i + 1 + nh);
                xt.push_back((int)pos);
            }
            if (!e.ident && !q.N[e.node[0]].leaf && !q.N[e.node[1]].leaf && !e.hybrid && e.fromBD1) {
                int i = findIn(e.number, nh + ntIdent, np);
                if (i < 0) ERR("internal edge corresponding to gammaz in qnet not found in ht(net)");
                qnhz.push_back(i + 1 + nh + ntIdent);
                xhz.push_back((int)pos);
            }
            if (e.hybrid && !e.ismajor) {
                int i = findIn(e.number, 0, nh);
                if (i >= 0) { qnh.push_back(i + 1); xh.push_back((int)pos); }
            }
        }
    }
    q.indexht = qnh;
    q.indexht.insert(q.indexht.end(), qnt.begin(), qnt.end());
    q.indexht.insert(q.indexht.end(), qnhz.begin(), qnhz.end());
    q.index = xh;
    q.index.insert(q.index.end(), xt.begin(), xt.end());
    q.index.insert(q.index.end(), xhz.begin(), xhz.end());
}

// update!(qnet,x,net): src/snaq_optimization.jl:198-239, without the `changed`
// gating (stateless: every quartet is treated as changed; the gating only skips
// recomputation, it never alters a recomputed value).
void updateQ(Net &q, const Net &net, const Params &p, const double *x) {
    int nh = net.numhybrids - net.numBad, ntIdent = p.nt;
    bool anyBD1 = false;
    for (int hn : q.hybrid) anyBD1 |= q.N[hn].bd1;
    if (anyBD1) {
        if (q.indexht.size() != 2) ERR("strange qnet from bad diamond I with hybrid node, it should have only 2 elements");
        for (int i = 0; i < 2; i++) {
            double v = x[q.indexht[i] - 1];
            if (!(0 <= v && v <= 1)) fail(SNAQ_ERANGE, "new gammaz value should be between 0,1: " + std::to_string(v));
            q.N[q.node[q.index[i]]].gammaz = v;
        }
        return;
    }
    for (size_t i = 0; i < q.indexht.size(); i++) {
        int slot = q.indexht[i];
        double v = x[slot - 1];
        int e = q.edge[q.index[i]];
        if (slot <= nh) {
            if (!(0 <= v && v <= 1)) fail(SNAQ_ERANGE, "new gamma value should be between 0,1: " + std::to_string(v));
            if (!q.E[e].hybrid) ERR("something odd here, optimizing gamma for tree edge");
            // setgamma!(e, v, true): partner hybrid edge of the same hybrid node gets 1-v
            int node = q.N[q.E[e].node[0]].hybrid ? q.E[e].node[0] : q.E[e].node[1];
            int partner = -1;
            for (int j = 0; j < q.N[node].ne; j++) {
                int ee = q.N[node].edge[j];
                if (ee != e && q.E[ee].hybrid) partner = ee;
            }
            if (partner < 0) ERR("strange hybrid node without 2 hybrid parents");
            q.E[e].gamma = v;
            q.E[partner].gamma = 1.0 - v;
        } else if (slot <= nh + ntIdent) {
            setLength(q, e, v);
        } else {
            if (!(0 <= v && v <= 1)) fail(SNAQ_ERANGE, "new gammaz value should be between 0,1: " + std::to_string(v));
            if (approxEq(v, 1.0)) setLength(q, e, 10.0);
            else setLength(q, e, -std::log(1 - v));
        }
    }
}

// drop dead storage after pruning so that the per-evaluation deepcopy(q.qnet)
// (src/pseudolik.jl:1318) copies a 4-taxon object, as it does in the reference
void compact(Net &q) {
    std::vector<int> nmap(q.N.size(), -1), emap(q.E.size(), -1);
    std::vector<Node> N2; std::vector<Edge> E2;
    for (int n : q.node) { nmap[n] = (int)N2.size(); N2.push_back(q.N[n]); }
    for (int e : q.edge) { emap[e] = (int)E2.size(); E2.push_back(q.E[e]); }
    for (Node &n : N2) { for (int i = 0; i < n.ne; i++) n.edge[i] = emap[n.edge[i]]; n.prev = -1; }
    for (Edge &e : E2) for (int i = 0; i < e.nn; i++) e.node[i] = nmap[e.node[i]];
    for (int &n : q.node) n = nmap[n];
    for (int &e : q.edge) e = emap[e];
    for (int &n : q.hybrid) n = nmap[n];
    for (int &n : q.leaf) n = nmap[n];
    q.N.swap(N2); q.E.swap(E2);
}

// ---------------- expected CFs ------------------------------------------------
// identifyQuartet!(qnet,node): src/pseudolik.jl:746-811
void identifyQuartetNode(Net &q, int node) {
    Node &nd = q.N[node];
    if (!nd.hybrid) ERR("cannot identify the hybridization around a tree node");
    auto inCyc3 = [&](int n) { return q.N[n].inCycle == nd.number && q.N[n].ne == 3; };
    int k = 0;
    for (int n : q.node) if (inCyc3(n)) k++;
    nd.k = k;
    if (k < 2) ERR("strange quartet network with a hybrid node but no cycle");
    if (k == 2) {
        int other = -1;
        for (int n : q.node) if (inCyc3(n) && n != node) { other = n; break; }
        int e1 = hybridEdges(q, node).c, e2 = hybridEdges(q, other).c;
        bool t1 = q.N[getOtherNode(q, e1, node)].leaf || q.N[getOtherNode(q, e2, other)].leaf;
        nd.typeHyb = t1 ? 1 : 3;
        nd.prev = other;
        q.typeHyb.push_back(nd.typeHyb);
    } else if (k == 3) {
        int e3 = hybridEdges(q, node).c;
        if (q.N[getOtherNode(q, e3, node)].leaf) {
            nd.typeHyb = 2;
            q.typeHyb.push_back(2);
            for (int n : q.node) {
                if (inCyc3(n) && n != node) {
                    int c = hybridEdges(q, n).c;
                    if (c >= 0 && q.N[getOtherNode(q, c, n)].leaf) { nd.prev = n; break; }
                }
            }
        } else {
            nd.typeHyb = 4;
            q.typeHyb.push_back(4);
            for (int n : q.node) if (inCyc3(n) && n != node) { nd.prev = n; break; }
        }
    } else if (k == 4) {
        nd.typeHyb = 5;
        q.typeHyb.push_back(5);
        nd.prev = -1;
    } else ERR("strange quartet network with more than 4 nodes in cycle");
}
// cleanUpNode!: src/pseudolik.jl:852-857
void cleanUpNode(Net &q, int node) {
    E3 e = hybridEdges(q, node);
    deleteIntLeafWhileNode(q, getOtherNode(q, e.a, node), node);
    deleteIntLeafWhileNode(q, getOtherNode(q, e.b, node), node);
    deleteIntLeafWhileNode(q, getOtherNode(q, e.c, node), node);
}
// identifyQuartet!(qnet): src/pseudolik.jl:815-848
void identifyQuartet(Net &q) {
    if (q.numhybrids == 0) q.which = 1;
    else if (q.numhybrids == 1) {
        identifyQuartetNode(q, q.hybrid[0]);
        q.which = (q.typeHyb[0] == 5) ? 2 : 1;
    } else {
        std::vector<int> hy = q.hybrid;
        for (int n : hy) { cleanUpNode(q, n); identifyQuartetNode(q, n); }
        bool any5 = false;
        for (int n : q.hybrid) any5 |= (q.N[n].typeHyb == 5);
        q.which = any5 ? 2 : 1;
    }
}
// eliminateLoop!: src/pseudolik.jl:866-888
void eliminateLoop(Net &q, int node, bool internal) {
    E3 e = hybridEdges(q, node);
    deleteIntLeafWhileEdge(q, e.a, node);
    deleteIntLeafWhileEdge(q, e.b, node);
    int other = getOtherNode(q, e.a, node);
    if (getOtherNode(q, e.b, node) != other) ERR("node and other are not the two nodes in a cycle with k=2");
    if (q.N[other].number != q.N[q.N[node].prev].number) ERR("other node should be the same as the stored in node.prev");
    removeEdge(q, node, e.b);
    removeEdge(q, other, e.b);
    deleteEdge(q, e.b);
    if (internal) {
        const Edge &a = q.E[e.a], &b = q.E[e.b];
        setLength(q, e.a, -std::log(1 - a.gamma * a.gamma * a.z - b.gamma * b.gamma * b.z));
    } else {
        int leaf = getOtherNode(q, e.c, node);
        deleteIntLeafWhileNode(q, node, leaf);
    }
}
// eliminateTriangle!: src/pseudolik.jl:909-962
void eliminateTriangle(Net &q, int node, int other, int cas) {
    E3 e = hybridEdges(q, node);
    int edgemaj = e.a, edgemin = e.b, treeedge = e.c;
    if (edgemaj < 0 || edgemin < 0) ERR("edge maj/min is nothing");
    deleteIntLeafWhileEdge(q, edgemaj, node);
    deleteIntLeafWhileEdge(q, edgemin, node);
    int hybedge, otheredge;
    if (getOtherNode(q, edgemaj, node) == other) { hybedge = edgemaj; otheredge = edgemin; }
    else if (getOtherNode(q, edgemin, node) == other) { hybedge = edgemin; otheredge = edgemaj; }
    else ERR("node and other node are not connected by an edge");
    int middle = -1;
    for (int n : q.node)
        if (q.N[n].inCycle == q.N[node].number && q.N[n].ne == 3 && n != other && n != node) { middle = n; break; }
    if (middle < 0) ERR("middle node not found in eliminateTriangle");
    int edge = -1;
    for (int i = 0; i < q.N[middle].ne; i++) {
        int ee = q.N[middle].edge[i];
        if (q.E[ee].inCycle == q.N[node].number && getOtherNode(q, ee, middle) != node) { edge = ee; break; }
    }
    if (edge < 0) ERR("edge number not found in middle edge");
    deleteIntLeafWhileEdge(q, edge, middle);
    if (getOtherNode(q, edge, middle) != other) ERR("middle node and other node are not connected by an edge");
    if (cas == 1) {
        setLength(q, edge, -std::log(1 - q.E[hybedge].gamma * q.E[edge].z));
        removeEdge(q, middle, otheredge);
        removeEdge(q, other, hybedge);
        removeEdge(q, node, treeedge);
        removeNode(q, node, treeedge);
        setEdge(q, other, treeedge);
        setNode(q, treeedge, other);
        deleteEdge(q, otheredge);
        deleteEdge(q, hybedge);
        deleteNode(q, node);
    } else if (cas == 2) {
        const Edge &h = q.E[hybedge], &o = q.E[otheredge];
        if (!(h.hybrid && o.hybrid)) ERR("hybedge and otheredge should by hybrid edges in eliminateTriangle Case 2");
        setLength(q, hybedge, -std::log(o.gamma * o.gamma * o.y + h.gamma * o.gamma * (3 - q.E[edge].y) + h.gamma * h.gamma * h.y), true);
        removeEdge(q, middle, otheredge);
        removeEdge(q, node, otheredge);
        deleteEdge(q, otheredge);
        deleteIntLeafWhileNode(q, node, getOtherNode(q, treeedge, node), true);
        deleteIntLeafWhileEdge(q, edge, other);
    } else ERR("unknown case");
}
// eliminateHybridization!(qnet,node): src/pseudolik.jl:1056-1071
void eliminateHybridizationNode(Net &q, int node) {
    int t = q.N[node].typeHyb;
    if (t == 1) eliminateLoop(q, node, false);
    else if (t == 3) eliminateLoop(q, node, true);
    else if (t == 4) eliminateTriangle(q, node, q.N[node].prev, 2);
    else if (t == 2) eliminateTriangle(q, node, q.N[node].prev, 1);
    else if (t != 5) ERR("node type of hybridization should be 1,2,3,4 or 5");
}
// internalLength!: src/pseudolik.jl:1076-1109
void internalLength(Net &q) {
    if (q.which != 1) return;
    int node = -1, node2 = -1;
    for (int n : q.node) if (q.N[n].ne == 3) { node = n; break; }
    if (node < 0) ERR("not found internal node in qnet with 3 edges");
    for (int n : q.node) if (q.N[n].ne == 3 && n != node) { node2 = n; break; }
    if (node2 < 0) ERR("not found another internal node in qnet with 3 edges");
    for (int i = 0; i < 3; i++) deleteIntLeafWhileEdge(q, q.N[node].edge[i], node, true);
    for (int i = 0; i < 3; i++) deleteIntLeafWhileEdge(q, q.N[node2].edge[i], node2, true);
    int edge = -1;
    for (int i = 0; i < 3; i++)
        if (!q.N[getOtherNode(q, q.N[node].edge[i], node)].leaf) edge = q.N[node].edge[i];
    if (edge < 0) ERR("cannot find internal edge attached to node in qnet");
    if (getOtherNode(q, edge, node) != node2) ERR("strange internal edge found in qnet");
    q.t1 = q.E[edge].length;
}
// eliminateHybridization!(qnet): src/pseudolik.jl:1115-1145
void eliminateHybridization(Net &q) {
    if (q.which == -1) ERR("qnet which has to be updated by now to 1 or 2");
    if (q.numhybrids == 1) eliminateHybridizationNode(q, q.hybrid[0]);
    else if (q.numhybrids > 1) {
        auto anyType1 = [&]() { for (int n : q.hybrid) if (q.N[n].typeHyb == 1) return true; return false; };
        while (q.numhybrids > 0 && anyType1()) {
            std::vector<int> hy = q.hybrid;
            for (int n : hy)
                if (q.N[n].typeHyb == 1) {
                    if (q.N[n].prev < 0) ERR("hybrid node is type 1 hybridization, prev should be automatically set");
                    eliminateHybridizationNode(q, n);
                }
            q.typeHyb.clear();
            if (q.numhybrids > 0) identifyQuartet(q);
        }
        std::vector<int> hy = q.hybrid;
        for (int n : hy) eliminateHybridizationNode(q, n);
    }
    if (q.which == 1) internalLength(q);
}
// whichLeaves: src/pseudolik.jl:1155-1200 (names -> taxon ids)
void whichLeaves(const Net &q, int taxon1, int taxon2, const int lf[4], int &a, int &b) {
    a = b = -1;
    for (int i = 0; i < 4; i++) if (q.N[lf[i]].taxon == taxon1) { a = i + 1; break; }
    if (a < 0) ERR("taxon1 is not one of the 4 leaves in quartet");
    for (int i = 0; i < 4; i++) if (i + 1 != a && q.N[lf[i]].taxon == taxon2) { b = i + 1; break; }
    if (b < 0) ERR("taxon2 is not one of the three remaining leaves");
}
// updateSplit!: src/pseudolik.jl:1207-1224
void updateSplit(Net &q) {
    if (q.which == 1) {
        if (q.split[0] == -1 && q.split[1] == -1 && q.split[2] == -1 && q.split[3] == -1) {
            for (int i = 0; i < 4; i++) q.split[i] = 2;
            int middle = -1;
            for (int n : q.node) if (q.N[n].ne == 3) { middle = n; break; }
            int l1 = -1, l2 = -1;
            for (int i = 0; i < 3; i++) {
                int e = q.N[middle].edge[i];
                if (q.N[getOtherNode(q, e, middle)].leaf) { if (l1 < 0) l1 = e; else if (l2 < 0) l2 = e; }
            }
            if (l1 < 0 || l2 < 0) ERR("updateSplit!: middle node without two leaves");
            int leaf1 = getOtherNode(q, l1, middle), leaf2 = getOtherNode(q, l2, middle);
            for (int i = 0; i < 4; i++) if (q.leaf[i] == leaf1 || q.leaf[i] == leaf2) q.split[i] = 1;
        }
    } else if (q.which == -1) ERR("cannot update split in quartet network if it has not been identified");
}
// updateFormula!: src/pseudolik.jl:1230-1247
void updateFormula(Net &q) {
    if (q.which == 1) {
        if (q.formula[0] == -1 && q.formula[1] == -1 && q.formula[2] == -1) {
            q.formula[0] = q.formula[1] = q.formula[2] = 2;
            if (q.leaf.size() != 4) ERR("strange quartet with not 4 leaves");
            int lf[4] = {q.leaf[0], q.leaf[1], q.leaf[2], q.leaf[3]};
            for (int i = 1; i < 4; i++) {
                int a, b;
                whichLeaves(q, q.qtaxon[0], q.qtaxon[i], lf, a, b);
                if (q.split[a - 1] == q.split[b - 1]) { q.formula[i - 1] = 1; break; }
            }
        }
    } else if (q.which != 2) ERR("qnet.which should be updated already to 1 or 2");
}
// quartetType5!: src/pseudolik.jl:967-1052
void quartetType5(Net &q, int node, int perm[3]) {
    Node &nd = q.N[node];
    if (!(nd.hybrid && nd.typeHyb == 5)) ERR("cannot polish the quartet type 5 hybridization");
    E3 e = hybridEdges(q, node);
    int edge1 = e.a, edge2 = e.b, edge3 = e.c;
    if (!nd.bd1) { deleteIntLeafWhileEdge(q, edge1, node); deleteIntLeafWhileEdge(q, edge2, node); }
    int other1 = getOtherNode(q, edge1, node), other2 = getOtherNode(q, edge2, node);
    E3 h1 = hybridEdges(q, other1), h2 = hybridEdges(q, other2);
    int edge5 = h1.b, edgetree1 = h1.c, edge6 = h2.b, edgetree2 = h2.c;
    if (!nd.bd1) { deleteIntLeafWhileEdge(q, edge5, other1); deleteIntLeafWhileEdge(q, edge6, other2); }
    double cf1, cf2, cf3;
    if (nd.bd1) {
        double g1 = q.N[other1].gammaz, g2 = q.N[other2].gammaz;
        if (g1 == -1 || g2 == -1) ERR("node is bad diamond I but gammaz are -1");
        cf1 = (1 + 2 * g1 - g2) / 3;
        cf2 = (1 + 2 * g2 - g1) / 3;
        cf3 = (1 - g1 - g2) / 3;
    } else {
        const Edge &E1 = q.E[edge1], &E2 = q.E[edge2], &E5 = q.E[edge5], &E6 = q.E[edge6];
        cf1 = E1.gamma * (1 - 2. / 3 * E5.y) + E2.gamma * 1 / 3 * E6.y;
        cf2 = E1.gamma * 1 / 3 * E5.y + E2.gamma * (1 - 2. / 3 * E6.y);
        cf3 = E1.gamma * 1 / 3 * E5.y + E2.gamma * 1 / 3 * E6.y;
    }
    if (edgetree1 < 0 || edgetree2 < 0) ERR("quartetType5!: edgetree is nothing");
    int leaf1 = getOtherNode(q, edge3, node);
    int leaf2 = getOtherNode(q, edgetree1, other1);
    int leaf3 = getOtherNode(q, edgetree2, other2);
    int leaf4 = -1;
    for (int l : q.leaf) if (l != leaf1 && l != leaf2 && l != leaf3) { leaf4 = l; break; }
    int lf[4] = {leaf1, leaf2, leaf3, leaf4};
    auto is12 = [](int a, int b) { return (a == 1 && b == 2) || (a == 2 && b == 1) || (a == 3 && b == 4) || (a == 4 && b == 3); };
    auto is13 = [](int a, int b) { return (a == 1 && b == 3) || (a == 3 && b == 1) || (a == 2 && b == 4) || (a == 4 && b == 2); };
    auto is14 = [](int a, int b) { return (a == 1 && b == 4) || (a == 4 && b == 1) || (a == 3 && b == 2) || (a == 2 && b == 3); };
    double cf[4] = {0, cf1, cf2, cf3};
    int a, b;
    whichLeaves(q, q.qtaxon[0], q.qtaxon[1], lf, a, b);
    if (is12(a, b)) {
        perm[0] = 1;
        whichLeaves(q, q.qtaxon[0], q.qtaxon[2], lf, a, b);
        if (is13(a, b)) { perm[1] = 2; perm[2] = 3; }
        else if (is14(a, b)) { perm[1] = 3; perm[2] = 2; }
        else ERR("strange quartet network, could not find which leaves correspond to taxon1, taxon3");
    } else if (is13(a, b)) {
        perm[0] = 2;
        whichLeaves(q, q.qtaxon[0], q.qtaxon[2], lf, a, b);
        if (is12(a, b)) { perm[1] = 1; perm[2] = 3; }
        else if (is14(a, b)) { perm[1] = 3; perm[2] = 1; }
        else ERR("strange quartet network, could not find which leaves correspond to taxon1, taxon3");
    } else if (is14(a, b)) {
        perm[0] = 3;
        whichLeaves(q, q.qtaxon[0], q.qtaxon[2], lf, a, b);
        if (is13(a, b)) { perm[1] = 2; perm[2] = 1; }
        else if (is12(a, b)) { perm[1] = 1; perm[2] = 2; }
        else ERR("strange quartet network, could not find which leaves correspond to taxon1, taxon3");
    } else ERR("strange quartet network, could not find which leaves correspond to taxon1, taxon2");
    for (int i = 0; i < 3; i++) q.expCF[i] = cf[perm[i]];
    if (!approxEq(q.expCF[0] + q.expCF[1] + q.expCF[2], 1.0))
        fail(SNAQ_ENUMERIC, "strange quartet network with hybridization of type 5: expCF do not add up to 1");
}
// calculateExpCF!: src/pseudolik.jl:1255-1275
void calculateExpCF(Net &q, int perm[3]) {
    if (q.which == 1) {
        if (q.formula[0] != -1 && q.t1 != -1) {
            for (int i = 0; i < 3; i++)
                q.expCF[i] = q.formula[i] == 1 ? 1 - 2. / 3 * std::exp(-q.t1) : 1. / 3 * std::exp(-q.t1);
        } else ERR("quartet network needs to have updated formula and t1 before computing the expCF");
    } else if (q.which == 2) {
        if (q.numhybrids == 1) {
            if (q.N[q.hybrid[0]].typeHyb == 5) quartetType5(q, q.hybrid[0], perm);
            else ERR("strange quartet network type 2 with one hybrid node but it is not type 5");
        } else ERR("quartet network with type 2 but with " + std::to_string(q.numhybrids) + " hybrid nodes");
    }
}
// calculateExpCFAll!(qnet): src/pseudolik.jl:1279-1285
void calculateExpCFAll(Net &q, int perm[3]) {
    identifyQuartet(q);
    eliminateHybridization(q);
    updateSplit(q);
    updateFormula(q);
    calculateExpCF(q, perm);
}

// extractQuartet(net,quartet) + extractQuartet!(net,quartet::Quartet):
// src/pseudolik.jl:472-486, :523-538.  Result is the un-merged q.qnet.
void extractQuartet(const Net &net, const Params &p, const int taxa[4], Net &q) {
    q = net;  // QuartetNetwork(net) = deepcopy(net), src/types.jl:191-193
    int keep[4];
    for (int i = 0; i < 4; i++) {
        keep[i] = -1;
        for (int n : net.node) if (net.N[n].leaf && net.N[n].taxon == taxa[i]) { keep[i] = n; break; }
        if (keep[i] < 0) fail(SNAQ_EINVAL, "taxon " + std::to_string(taxa[i]) + " not in network");
        q.qtaxon[i] = taxa[i];
    }
    std::vector<int> leaves = q.leaf;
    for (int n : leaves) {
        bool in = false;
        for (int i = 0; i < 4; i++) in |= (net.N[keep[i]].number == q.N[n].number);
        if (!in) deleteLeaf(q, n);
    }
    redundantCycle(q);
    updateHasEdge(q, net);
    parametersQ(q, net, p);
}

// logPseudoLik(quartet): src/pseudolik.jl:1388-1410
double logPseudoLikQ(const double exp[3], const double obs[3]) {
    if (exp[0] + exp[1] + exp[2] == 0.0) fail(SNAQ_ENUMERIC, "expCF not updated for quartet");
    double suma = 0.0;
    for (int i = 0; i < 3; i++) {
        if (exp[i] < 0) suma += -1.e15;
        else suma += obs[i] == 0 ? 0.0 : 100 * obs[i] * std::log(exp[i] / obs[i]);
    }
    return suma;
}

// ---------------- flat network -> Net ----------------------------------------
void fromFlat(const snaq_flatnet *f, Net &g) {
    if (!f) fail(SNAQ_EINVAL, "null flat network");
    g.N.assign(f->n_nodes, Node());
    g.E.assign(f->n_edges, Edge());
    for (int i = 0; i < f->n_nodes; i++) {
        Node &n = g.N[i];
        uint32_t fl = f->node_flags[i];
        n.number = f->node_number[i];
        n.leaf = fl & SNAQ_NODE_LEAF; n.hybrid = fl & SNAQ_NODE_HYBRID; n.hasHybEdge = fl & SNAQ_NODE_HASHYBEDGE;
        n.bd1 = fl & SNAQ_NODE_BADDIAMOND1; n.bd2 = fl & SNAQ_NODE_BADDIAMOND2; n.badtri = fl & SNAQ_NODE_BADTRIANGLE;
        n.inCycle = f->node_incycle[i];
        n.gammaz = f->node_gammaz[i];
        n.taxon = f->node_taxon[i];
        n.ne = 0;
        for (int j = 0; j < 3; j++) { int e = f->node_edge[3 * i + j]; if (e >= 0) n.edge[n.ne++] = e; }
        g.node.push_back(i);
    }
    for (int i = 0; i < f->n_edges; i++) {
        Edge &e = g.E[i];
        uint32_t fl = f->edge_flags[i];
        e.number = f->edge_number[i];
        e.hybrid = fl & SNAQ_EDGE_HYBRID; e.ismajor = fl & SNAQ_EDGE_MAJOR;
        e.ident = fl & SNAQ_EDGE_IDENTIFIABLE; e.fromBD1 = fl & SNAQ_EDGE_FROMBD1;
        e.inCycle = f->edge_incycle[i];
        e.length = f->edge_length[i];
        e.y = std::exp(-e.length); e.z = 1.0 - e.y;
        e.gamma = f->edge_gamma[i];
        e.nn = 2; e.node[0] = f->edge_node[2 * i]; e.node[1] = f->edge_node[2 * i + 1];
        g.edge.push_back(i);
    }
    for (int i = 0; i < f->n_hybrids; i++) g.hybrid.push_back(f->hybrid[i]);
    for (int i = 0; i < f->n_leaves; i++) g.leaf.push_back(f->leaf[i]);
    g.numhybrids = f->n_hybrids;
    g.numBad = 0;
    for (int h : g.hybrid) if (g.N[h].bd1) g.numBad++;
}

thread_local std::string g_err;
int run(const std::function<void()> &f) {
    try { f(); return SNAQ_OK; }
    catch (const OErr &e) { g_err = e.what(); return e.code; }
    catch (const std::exception &e) { g_err = e.what(); return SNAQ_EINVAL; }
}
}  // namespace

extern "C" {

const char *oracle_last_error(void) { return g_err.c_str(); }

// parameters!(net): x, numht, counts.  x/numht may be NULL (size query).
int oracle_parameters(const snaq_flatnet *f, int32_t *nparams, int32_t *nh, int32_t *nt, int32_t *nhz,
                      double *x, int32_t *numht, double *upper) {
    return run([&] {
        Net g; fromFlat(f, g);
        Params p = parameters(g);
        if (nparams) *nparams = (int)p.x.size();
        if (nh) *nh = p.nh;
        if (nt) *nt = p.nt;
        if (nhz) *nhz = p.nhz;
        for (size_t i = 0; i < p.x.size(); i++) {
            if (x) x[i] = p.x[i];
            if (numht) numht[i] = p.numht[i];
            if (upper) upper[i] = ((int)i >= p.nh && (int)i < p.nh + p.nt) ? 10.0 : 1.0;  // upper(net)
        }
    });
}


/* debugging aid: dump the pruned quartet network before and after calculateExpCFAll! */
static void dumpNet(const Net &q, const char *title) {
    fprintf(stderr, "--- %s: numhybrids %d which %d t1 %g\n", title, q.numhybrids, q.which, q.t1);
    for (size_t i = 0; i < q.node.size(); i++) {
        const Node &n = q.N[q.node[i]];
        fprintf(stderr, "  node[%zu] num %d leaf %d hyb %d deg %d inCycle %d taxon %d type %d edges:", i, n.number, n.leaf, n.hybrid, n.ne, n.inCycle, n.taxon, n.typeHyb);
        for (int j = 0; j < n.ne; j++) fprintf(stderr, " %d", q.E[n.edge[j]].number);
        fprintf(stderr, "\n");
    }
    for (size_t i = 0; i < q.edge.size(); i++) {
        const Edge &e = q.E[q.edge[i]];
        fprintf(stderr, "  edge %d: %d--%d len %.6f hyb %d maj %d gamma %.4f ident %d inCycle %d\n", e.number, q.N[e.node[0]].number, q.N[e.node[1]].number, e.length, e.hybrid, e.ismajor, e.gamma, e.ident, e.inCycle);
    }
}
int oracle_debug_quartet(const snaq_flatnet *f, const uint16_t *taxa, const double *x) {
    return run([&] {
        Net g; fromFlat(f, g);
        Params p = parameters(g);
        if (x) applyParameters(g, p, x);
        int tx[4] = {taxa[0], taxa[1], taxa[2], taxa[3]};
        Net q; extractQuartet(g, p, tx, q);
        dumpNet(q, "after extractQuartet!");
        int perm[3] = {0, 0, 0};
        identifyQuartet(q);
        dumpNet(q, "after identifyQuartet!");
        eliminateHybridization(q);
        dumpNet(q, "after eliminateHybridization!");
    });
}

/*
 * extractQuartet!(net,d) followed by calculateExpCFAll!(deepcopy(qnet)) per
 * quartet, under parameter vector x (NULL = the network's own values).
 * Outputs (any may be NULL): which[nq]; ntype[nq] + types[nq*max_types] = final
 * qnet.typeHyb; formula[nq*3] (which==1: 1/2; which==2: 11..13 permutation);
 * hasedge[nq*nparams]; nindexht[nq] + indexht[nq*nparams] (1-based slots);
 * expcf[nq*3]; t1[nq].
 */
int oracle_extract(const snaq_flatnet *f, int64_t nq, const uint16_t *taxa, const double *x,
                   int8_t *which, int8_t *ntype, int8_t *types, int32_t max_types, int8_t *formula,
                   uint8_t *hasedge, int32_t *nindexht, int32_t *indexht, double *expcf, double *t1) {
    std::string err; int code = SNAQ_OK;
    int rc = run([&] {
        Net g; fromFlat(f, g);
        Params p = parameters(g);
        if (x) applyParameters(g, p, x);
        const int np = (int)p.x.size();
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t qi = 0; qi < nq; qi++) {
            try {
                int tx[4] = {taxa[4 * qi], taxa[4 * qi + 1], taxa[4 * qi + 2], taxa[4 * qi + 3]};
                Net q; extractQuartet(g, p, tx, q);
                if (hasedge) {
                    if ((int)q.hasEdge.size() != np) ERR("hasEdge length differs from number of parameters");
                    for (int i = 0; i < np; i++) hasedge[qi * np + i] = q.hasEdge[i];
                }
                if (nindexht) nindexht[qi] = (int)q.indexht.size();
                if (indexht) for (int i = 0; i < np; i++) indexht[qi * np + i] = i < (int)q.indexht.size() ? q.indexht[i] : -1;
                Net c = q;  // deepcopy(q.qnet), src/pseudolik.jl:508
                int perm[3] = {0, 0, 0};
                calculateExpCFAll(c, perm);
                if (which) which[qi] = (int8_t)c.which;
                if (ntype) ntype[qi] = (int8_t)c.typeHyb.size();
                if (types) for (int i = 0; i < max_types; i++) types[qi * max_types + i] = i < (int)c.typeHyb.size() ? (int8_t)c.typeHyb[i] : 0;
                if (formula) for (int i = 0; i < 3; i++) formula[qi * 3 + i] = (int8_t)(c.which == 1 ? c.formula[i] : 10 + perm[i]);
                if (expcf) for (int i = 0; i < 3; i++) expcf[qi * 3 + i] = c.expCF[i];
                if (t1) t1[qi] = c.t1;
            } catch (const OErr &e) {
#pragma omp critical
                { if (code == SNAQ_OK) { code = e.code; err = std::string(e.what()) + " (quartet " + std::to_string(qi) + ")"; } }
            }
        }
    });
    if (rc != SNAQ_OK) return rc;
    if (code != SNAQ_OK) { g_err = err; return code; }
    return SNAQ_OK;
}

/* edge numbers of q.qnet.edge after extractQuartet!(net, quartet), in order: what sampleEdgeQuartetWeighted
 * (src/moves_snaq.jl:1451-1453) intersects with its candidate edges.  *n = count (at most cap written). */
int oracle_quartet_edges(const snaq_flatnet *f, const uint16_t *taxa, int32_t *numbers, int32_t cap, int32_t *n) {
    return run([&] {
        Net g; fromFlat(f, g);
        Params p = parameters(g);
        int tx[4] = {taxa[0], taxa[1], taxa[2], taxa[3]};
        Net q; extractQuartet(g, p, tx, q);
        int k = 0;
        for (int ei : q.edge) { if (k < cap) numbers[k] = q.E[ei].number; k++; }
        *n = k;
    });
}

/*
 * out[p] = logPseudoLik(d) (src/pseudolik.jl:1417-1425) after a stateless
 * calculateExpCFAll!(d, X[p,:], net): serial sum in quartet order, unsampled
 * quartets contribute 0.  expcf_last (nq*3, optional) gets the expCF of the last
 * parameter vector.  threads<=0: all available.
 */
int oracle_eval(const snaq_flatnet *f, int64_t nq, const uint16_t *taxa, const double *obscf,
                const uint8_t *sampled, int32_t P, int32_t nparams, const double *X, double *out,
                double *expcf_last, int32_t threads) {
    std::string err; int code = SNAQ_OK;
    int rc = run([&] {
        Net g0; fromFlat(f, g0);
        Params p = parameters(g0);
        if (nparams != (int)p.x.size()) fail(SNAQ_EINVAL, "x has wrong length");
        if (P <= 0) return;
#ifdef _OPENMP
        if (threads > 0) omp_set_num_threads(threads);
#endif
        // extractQuartet!(net,d) once per topology (src/snaq_optimization.jl:354) ...
        std::vector<Net> qnets(nq);
        std::vector<double> lp(nq);
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t qi = 0; qi < nq; qi++) {
            if (sampled && !sampled[qi]) continue;
            try {
                int tx[4] = {taxa[4 * qi], taxa[4 * qi + 1], taxa[4 * qi + 2], taxa[4 * qi + 3]};
                extractQuartet(g0, p, tx, qnets[qi]);
                compact(qnets[qi]);
            } catch (const OErr &e) {
#pragma omp critical
                { if (code == SNAQ_OK) { code = e.code; err = std::string(e.what()) + " (quartet " + std::to_string(qi) + ")"; } }
            }
        }
        if (code != SNAQ_OK) return;
        // ... then the optBL! objective per x (src/snaq_optimization.jl:368-380):
        // update!(q.qnet,x,net); calculateExpCFAll!(deepcopy(q.qnet)); logPseudoLik(d)
        for (int pi = 0; pi < P; pi++) {
            const double *x = X + (size_t)pi * nparams;
#pragma omp parallel for schedule(dynamic, 64)
            for (int64_t qi = 0; qi < nq; qi++) {
                lp[qi] = 0.0;
                if (sampled && !sampled[qi]) continue;
                try {
                    updateQ(qnets[qi], g0, p, x);
                    Net q = qnets[qi];
                    int perm[3];
                    calculateExpCFAll(q, perm);
                    lp[qi] = logPseudoLikQ(q.expCF, obscf + 3 * qi);
                    if (expcf_last && pi == P - 1) for (int i = 0; i < 3; i++) expcf_last[3 * qi + i] = q.expCF[i];
                } catch (const OErr &e) {
#pragma omp critical
                    { if (code == SNAQ_OK) { code = e.code; err = std::string(e.what()) + " (quartet " + std::to_string(qi) + ")"; } }
                }
            }
            double suma = 0;
            for (int64_t qi = 0; qi < nq; qi++) suma += lp[qi];
            out[pi] = -suma;
        }
    });
    if (rc != SNAQ_OK) return rc;
    if (code != SNAQ_OK) { g_err = err; return code; }
    return SNAQ_OK;
}

}  // extern "C"
