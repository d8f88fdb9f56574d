#include "snaq_tables.h"

#include <algorithm>
#include <cstring>
#include <queue>

namespace {
struct Fail { int code; std::string msg; };
[[noreturn]] void fail(int code, const std::string &m) { throw Fail{code, m}; }

template <class T> size_t put(std::vector<unsigned char> &blob, const std::vector<T> &v) {
    size_t off = (blob.size() + 15) & ~size_t(15);
    blob.resize(off + std::max<size_t>(v.size() * sizeof(T), 16), 0);
    if (!v.empty()) std::memcpy(blob.data() + off, v.data(), v.size() * sizeof(T));
    return off;
}
}  // namespace

SnaqTables SnaqHostTables::view(const unsigned char *b) const {
    SnaqTables T;
    T.ntaxa = ntaxa; T.n_tnodes = n_tnodes; T.n_cycles = n_cycles; T.nparams = nparams; T.n_xe = n_xe; T.n_hybrids = n_hybrids;
    T.n_int = n_int; T.n_full = n_full;
    T.leaf_tnode = (const uint16_t *)(b + off_leaf_tnode);
    T.parent = (const uint16_t *)(b + off_parent);
    T.depth = (const uint16_t *)(b + off_depth);
    T.tcycle = (const int16_t *)(b + off_tcycle);
    T.pslot = (const uint16_t *)(b + off_pslot);
    T.blk = (const uint8_t *)(b + off_blk);
    T.cyc_k = (const uint16_t *)(b + off_cyc_k);
    T.cyc_off = (const uint16_t *)(b + off_cyc_off);
    T.cyc_slot = (const uint16_t *)(b + off_cyc_slot);
    T.cyc_gslot = (const uint16_t *)(b + off_cyc_gslot);
    T.cyc_flags = (const uint8_t *)(b + off_cyc_flags);
    T.cyc_hyb = (const uint8_t *)(b + off_cyc_hyb);
    T.torder = (const uint16_t *)(b + off_torder);
    T.cyc_order = (const uint16_t *)(b + off_cyc_order);
    T.dslot = (const uint16_t *)(b + off_dslot);
    T.dnode = (const uint16_t *)(b + off_dnode);
    T.crow = (const uint16_t *)(b + off_crow);
    T.dpath_off = (const uint32_t *)(b + off_dpath_off);
    T.dpath_slot = (const uint16_t *)(b + off_dpath_slot);
    T.blkmax = (const uint16_t *)(b + off_blkmax);
    T.hmax2 = (const uint16_t *)(b + off_hmax2);
    T.hmaxtax = (const uint16_t *)(b + off_hmaxtax);
    T.leaf_lca = nullptr;
    return T;
}

SnaqSimNet SnaqHostTables::sim_view(const unsigned char *b) const {
    SnaqSimNet v = sim;
    v.node_edge = (const int16_t *)(b + so_node_edge);
    v.node_flags = (const uint8_t *)(b + so_node_flags);
    v.node_incycle = (const int16_t *)(b + so_node_incycle);
    v.node_taxon = (const int16_t *)(b + so_node_taxon);
    v.edge_node = (const int16_t *)(b + so_edge_node);
    v.edge_flags = (const uint8_t *)(b + so_edge_flags);
    v.leaf_order = (const int16_t *)(b + so_leaf_order);
    v.hyb_order = (const int16_t *)(b + so_hyb_order);
    v.slot_ref = (const int16_t *)(b + so_slot_ref);
    v.slot_alias = (const int16_t *)(b + so_slot_alias);
    v.edge_number = (const int32_t *)(b + so_edge_number);
    v.node_number = (const int32_t *)(b + so_node_number);
    return v;
}

int snaq_build_tables(const snaq_flatnet *f, int ntaxa, SnaqHostTables &o, std::string &err, int keep_root) {
    try {
        if (!f) fail(SNAQ_EINVAL, "null network");
        const int N = f->n_nodes, E = f->n_edges, H = f->n_hybrids;
        if (N <= 0 || E <= 0 || H < 0) fail(SNAQ_EINVAL, "empty network");
        if (H > SNAQ_MAX_HYB) fail(SNAQ_EINVAL, "more than " + std::to_string(SNAQ_MAX_HYB) + " hybrid nodes");
        if (N > 60000) fail(SNAQ_EINVAL, "network too large for 16-bit node ids");
        auto isLeaf = [&](int n) { return (f->node_flags[n] & SNAQ_NODE_LEAF) != 0; };
        auto isHyb = [&](int n) { return (f->node_flags[n] & SNAQ_NODE_HYBRID) != 0; };
        auto eHyb = [&](int e) { return (f->edge_flags[e] & SNAQ_EDGE_HYBRID) != 0; };
        auto eMaj = [&](int e) { return (f->edge_flags[e] & SNAQ_EDGE_MAJOR) != 0; };
        auto eIdent = [&](int e) { return (f->edge_flags[e] & SNAQ_EDGE_IDENTIFIABLE) != 0; };
        auto other = [&](int e, int n) { return f->edge_node[2 * e] == n ? f->edge_node[2 * e + 1] : f->edge_node[2 * e]; };
        // every index a caller hands over is range-checked before it is used as one
        for (int e = 0; e < E; e++)
            for (int j = 0; j < 2; j++) {
                const int n = f->edge_node[2 * e + j];
                if (n < 0 || n >= N) fail(SNAQ_EINVAL, "edge_node out of range (edge " + std::to_string(f->edge_number[e]) + ")");
            }
        if (f->n_leaves < 0 || f->n_leaves > N) fail(SNAQ_EINVAL, "n_leaves out of range");
        std::vector<int> deg(N, 0);
        for (int n = 0; n < N; n++) {
            for (int j = 0; j < 3; j++) {
                int e = f->node_edge[3 * n + j];
                if (e >= E || e < -1) fail(SNAQ_EINVAL, "node_edge out of range");
                if (e >= 0) {
                    if (f->edge_node[2 * e] != n && f->edge_node[2 * e + 1] != n) fail(SNAQ_EINVAL, "node_edge / edge_node disagree");
                    deg[n]++;
                }
            }
            if (isLeaf(n) ? deg[n] != 1 : deg[n] != 3)
                fail(SNAQ_ETOPOLOGY, "node " + std::to_string(f->node_number[n]) + " has " + std::to_string(deg[n]) + " edges");
        }
        // child (hybrid) end of a hybrid edge
        auto hybChild = [&](int e) {
            int a = f->edge_node[2 * e], b = f->edge_node[2 * e + 1];
            if (isHyb(a) && !isHyb(b)) return a;
            if (isHyb(b) && !isHyb(a)) return b;
            if (isHyb(a) && isHyb(b)) return (f->edge_flags[e] & SNAQ_EDGE_ISCHILD1) ? a : b;
            fail(SNAQ_ETOPOLOGY, "hybrid edge " + std::to_string(f->edge_number[e]) + " does not touch a hybrid node");
        };
        std::vector<int> hybIndex(N, -1);
        for (int h = 0; h < H; h++) {
            int n = f->hybrid[h];
            if (n < 0 || n >= N || !isHyb(n)) fail(SNAQ_EINVAL, "hybrid list entry is not a hybrid node");
            hybIndex[n] = h;
        }
        for (int n = 0; n < N; n++) if (isHyb(n) && hybIndex[n] < 0) fail(SNAQ_EINVAL, "hybrid node missing from the hybrid list");

        // ---- parameters(net): src/snaq_optimization.jl:54-92 --------------------------
        std::vector<int> majE(H, -1), minE(H, -1), childE(H, -1);
        for (int h = 0; h < H; h++) {
            int n = f->hybrid[h];
            for (int j = 0; j < 3; j++) {
                int e = f->node_edge[3 * n + j];
                if (eHyb(e) && hybChild(e) == n) { if (eMaj(e)) majE[h] = e; else minE[h] = e; }
                else if (!eHyb(e)) childE[h] = e;
            }
            if (majE[h] < 0 || minE[h] < 0 || childE[h] < 0)
                fail(SNAQ_ETOPOLOGY, "hybrid node " + std::to_string(f->node_number[n]) + " does not have one major, one minor and one tree edge");
        }
        std::vector<double> hv, tv, zv;
        std::vector<int32_t> hn, tn, zn;
        std::vector<int> gammaSlotOfHyb(H, -1), zSlotOfHyb(H, -1), tIndexOfEdge(E, -1);
        for (int e = 0; e < E; e++) {
            if (eIdent(e)) { tIndexOfEdge[e] = (int)tv.size(); tv.push_back(f->edge_length[e]); tn.push_back(f->edge_number[e]); }
            if (eHyb(e) && !eMaj(e)) {
                int node = hybChild(e);
                int h = hybIndex[node];
                if (!(f->node_flags[node] & SNAQ_NODE_BADDIAMOND1)) {
                    gammaSlotOfHyb[h] = (int)hv.size();
                    hv.push_back(f->edge_gamma[e]); hn.push_back(f->edge_number[e]);
                } else {
                    zSlotOfHyb[h] = (int)zv.size();
                    int oa = other(majE[h], node), ob = other(minE[h], node);
                    zv.push_back(f->node_gammaz[oa]); zv.push_back(f->node_gammaz[ob]);
                    zn.push_back(std::stoi(std::to_string(f->node_number[node]) + "1"));
                    zn.push_back(std::stoi(std::to_string(f->node_number[node]) + "2"));
                }
            }
        }
        o.nh = (int)hv.size(); o.nt = (int)tv.size(); o.nhz = (int)zv.size();
        o.nparams = o.nh + o.nt + o.nhz;
        o.x0 = hv; o.x0.insert(o.x0.end(), tv.begin(), tv.end()); o.x0.insert(o.x0.end(), zv.begin(), zv.end());
        o.numht = hn; o.numht.insert(o.numht.end(), tn.begin(), tn.end()); o.numht.insert(o.numht.end(), zn.begin(), zn.end());
        o.upper.assign(o.nparams, 1.0);
        for (int i = o.nh; i < o.nh + o.nt; i++) o.upper[i] = 10.0;

        // ---- slots of the extended parameter vector ---------------------------------------
        o.xconst.clear();
        o.edge_slot.assign(E, -1);
        for (int e = 0; e < E; e++) {
            if (eIdent(e)) o.edge_slot[e] = o.nh + tIndexOfEdge[e];
            else if (!isLeaf(f->edge_node[2 * e]) && !isLeaf(f->edge_node[2 * e + 1])) {
                o.edge_slot[e] = o.nparams + (int)o.xconst.size();
                double len = f->edge_length[e];
                o.xconst.push_back(len > 10.0 ? 10.0 : len);
            }
        }
        o.n_xe = o.nparams + (int)o.xconst.size();
        if (o.n_xe > SNAQ_MAX_SLOT) fail(SNAQ_EINVAL, "too many parameters for 12-bit slots");
        auto slotOf = [&](int e) { return (uint16_t)(o.edge_slot[e] >= 0 ? o.edge_slot[e] : 0); };

        // ---- cycles: unique path between the two parents in the tree of major edges ---------
        std::vector<int> cycOfNode(N, -1), posOfNode(N, -1);
        std::vector<char> inCycleEdge(E, 0);
        std::vector<std::vector<int>> cycNodes(H), cycEdges(H);
        for (int h = 0; h < H; h++) {
            int Hn = f->hybrid[h];
            int target = other(minE[h], Hn);
            int start = other(majE[h], Hn);
            // BFS from `start` over non-minor edges, not through Hn, not through leaves
            std::vector<int> prevN(N, -2), prevE(N, -1);
            std::queue<int> qu;
            prevN[start] = -1;
            qu.push(start);
            bool found = (start == target);
            while (!qu.empty() && !found) {
                int u = qu.front(); qu.pop();
                for (int j = 0; j < 3 && !found; j++) {
                    int e = f->node_edge[3 * u + j];
                    if (e < 0 || (eHyb(e) && !eMaj(e))) continue;
                    int v = other(e, u);
                    if (v == Hn || isLeaf(v) || prevN[v] != -2) continue;
                    prevN[v] = u; prevE[v] = e;
                    if (v == target) found = true; else qu.push(v);
                }
            }
            if (!found) fail(SNAQ_ETOPOLOGY, "hybrid node " + std::to_string(f->node_number[Hn]) + " does not close a cycle");
            std::vector<int> pn, pe;     // from target back to start
            for (int v = target; v != -1; v = prevN[v]) { pn.push_back(v); if (prevE[v] >= 0) pe.push_back(prevE[v]); }
            std::reverse(pn.begin(), pn.end());
            std::reverse(pe.begin(), pe.end());
            cycNodes[h].push_back(Hn);
            for (int v : pn) cycNodes[h].push_back(v);
            cycEdges[h].push_back(majE[h]);
            for (int e : pe) cycEdges[h].push_back(e);
            cycEdges[h].push_back(minE[h]);
            int k = (int)cycNodes[h].size();
            if (k < 3) fail(SNAQ_ETOPOLOGY, "cycle with only " + std::to_string(k) + " nodes: parallel edges");
            if (k > 255) fail(SNAQ_EINVAL, "cycle with more than 255 nodes");
            for (int j = 0; j < k; j++) {
                int v = cycNodes[h][j];
                if (cycOfNode[v] >= 0) fail(SNAQ_ETOPOLOGY, "cycles of hybrids intersect: not a level-1 network");
                cycOfNode[v] = h; posOfNode[v] = j;
            }
            for (int e : cycEdges[h]) inCycleEdge[e] = 1;
            if ((f->node_flags[Hn] & SNAQ_NODE_BADDIAMOND1) && k != 4) fail(SNAQ_ETOPOLOGY, "bad diamond I flag on a cycle that is not a 4-cycle");
        }

        // ---- blob tree T' -------------------------------------------------------------------
        std::vector<int> tOfNode(N, -1);
        int nt = 0;
        std::vector<int> superOfCycle(H, -1);
        for (int n = 0; n < N; n++) {
            if (cycOfNode[n] < 0) tOfNode[n] = nt++;
        }
        for (int h = 0; h < H; h++) superOfCycle[h] = nt++;
        for (int n = 0; n < N; n++) if (cycOfNode[n] >= 0) tOfNode[n] = superOfCycle[cycOfNode[n]];
        o.n_tnodes = nt;
        std::vector<std::vector<std::pair<int, int>>> adj(nt);   // (neighbour tnode, net edge)
        int ntedges = 0;
        for (int e = 0; e < E; e++) {
            if (inCycleEdge[e]) continue;
            int a = tOfNode[f->edge_node[2 * e]], b = tOfNode[f->edge_node[2 * e + 1]];
            if (a == b) fail(SNAQ_ETOPOLOGY, "edge " + std::to_string(f->edge_number[e]) + " is a chord of a cycle: not level-1");
            if (eHyb(e)) fail(SNAQ_ETOPOLOGY, "hybrid edge outside of its cycle");
            adj[a].push_back({b, e}); adj[b].push_back({a, e});
            ntedges++;
        }
        if (ntedges != nt - 1) fail(SNAQ_ETOPOLOGY, "network is not connected or not level-1 (blob graph is not a tree)");
        auto bfs = [&](int src, std::vector<int> &par, std::vector<int> &pe, std::vector<int> &dep) {
            par.assign(nt, -1); pe.assign(nt, -1); dep.assign(nt, -1);
            std::queue<int> qu; qu.push(src); dep[src] = 0; par[src] = src;
            int last = src;
            while (!qu.empty()) {
                int u = qu.front(); qu.pop(); last = u;
                for (auto &pr : adj[u]) if (dep[pr.first] < 0) { dep[pr.first] = dep[u] + 1; par[pr.first] = u; pe[pr.first] = pr.second; qu.push(pr.first); }
            }
            return last;
        };
        std::vector<int> par, pe, dep;
        int far1 = bfs(0, par, pe, dep);
        for (int v = 0; v < nt; v++) if (dep[v] < 0) fail(SNAQ_ETOPOLOGY, "network is not connected");
        int far2 = bfs(far1, par, pe, dep);
        int root = far2;
        for (int i = 0; i < dep[far2] / 2; i++) root = par[root];     // middle of a diameter: small depths
        if (keep_root >= 0 && keep_root < nt) root = keep_root;
        o.root = root;
        bfs(root, par, pe, dep);
        std::vector<uint16_t> t_parent(nt), t_depth(nt), t_pslot(nt);
        std::vector<int16_t> t_cycle(nt, -1);
        for (int v = 0; v < nt; v++) {
            t_parent[v] = (uint16_t)par[v]; t_depth[v] = (uint16_t)dep[v];
            t_pslot[v] = pe[v] >= 0 ? slotOf(pe[v]) : 0;
        }
        for (int h = 0; h < H; h++) t_cycle[superOfCycle[h]] = (int16_t)h;
        std::vector<uint16_t> t_order(nt, 0);
        for (int n = 0; n < N; n++) if (cycOfNode[n] < 0) t_order[tOfNode[n]] = (uint16_t)n;   // flat nodes are in net.node order

        // ---- taxa ---------------------------------------------------------------------------
        o.ntaxa = ntaxa;
        std::vector<uint16_t> leaf_tnode(std::max(ntaxa, 1), 0xFFFF);
        std::vector<int> nodeOfTaxon(std::max(ntaxa, 1), -1);
        int nleaves = 0;
        for (int n = 0; n < N; n++) {
            if (!isLeaf(n)) continue;
            nleaves++;
            int t = f->node_taxon[n];
            if (t < 0 || t >= ntaxa) fail(SNAQ_EINVAL, "leaf " + std::to_string(f->node_number[n]) + " has taxon id " + std::to_string(t) + " outside the data's taxon list");
            if (nodeOfTaxon[t] >= 0) fail(SNAQ_EINVAL, "two leaves carry the same taxon id");
            nodeOfTaxon[t] = n;
            leaf_tnode[t] = (uint16_t)tOfNode[n];
        }
        if (f->n_leaves != nleaves) fail(SNAQ_EINVAL, "n_leaves does not match the number of leaf nodes");

        // ---- blocks: position of every taxon in every cycle -----------------------------------
        std::vector<uint8_t> blk((size_t)std::max(H, 1) * std::max(ntaxa, 1), 0);
        std::vector<std::vector<char>> blockHasCycle(H);
        for (int h = 0; h < H; h++) {
            int k = (int)cycNodes[h].size();
            blockHasCycle[h].assign(k, 0);
            for (int j = 0; j < k; j++) {
                int cn = cycNodes[h][j];
                int e0 = -1;
                for (int jj = 0; jj < 3; jj++) { int e = f->node_edge[3 * cn + jj]; if (!inCycleEdge[e] || cycOfNode[other(e, cn)] != h) e0 = e; }
                if (e0 < 0) fail(SNAQ_ETOPOLOGY, "cycle node without an edge leaving the cycle");
                // DFS away from the cycle
                std::vector<int> st{other(e0, cn)};
                std::vector<char> seen(N, 0);
                seen[cn] = 1; seen[st[0]] = 1;
                while (!st.empty()) {
                    int u = st.back(); st.pop_back();
                    if (cycOfNode[u] == h) fail(SNAQ_ETOPOLOGY, "block of a cycle node re-enters the cycle: not level-1");
                    if (cycOfNode[u] >= 0) blockHasCycle[h][j] = 1;
                    if (isLeaf(u)) blk[(size_t)h * ntaxa + f->node_taxon[u]] = (uint8_t)j;
                    for (int jj = 0; jj < 3; jj++) {
                        int e = f->node_edge[3 * u + jj];
                        if (e < 0) continue;
                        int v = other(e, u);
                        if (!seen[v]) { seen[v] = 1; st.push_back(v); }
                    }
                }
            }
        }

        // ---- cycle tables -------------------------------------------------------------------------
        o.n_cycles = H; o.n_hybrids = H;
        std::vector<uint16_t> cyc_k(std::max(H, 1), 0), cyc_off(std::max(H, 1), 0), cyc_slot, cyc_order, cyc_gslot(std::max(H, 1), 0);
        std::vector<uint8_t> cyc_flags(std::max(H, 1), 0), cyc_hyb(std::max(H, 1), 0);
        for (int h = 0; h < H; h++) {
            cyc_k[h] = (uint16_t)cycNodes[h].size();
            cyc_off[h] = (uint16_t)cyc_slot.size();
            for (int e : cycEdges[h]) cyc_slot.push_back(slotOf(e));
            for (int v : cycNodes[h]) cyc_order.push_back((uint16_t)v);
            bool bd1 = (f->node_flags[f->hybrid[h]] & SNAQ_NODE_BADDIAMOND1) != 0;
            cyc_flags[h] = bd1 ? 1 : 0;
            cyc_gslot[h] = (uint16_t)(bd1 ? o.nh + o.nt + zSlotOfHyb[h] : gammaSlotOfHyb[h]);
            cyc_hyb[h] = (uint8_t)h;
        }
        if (cyc_slot.empty()) { cyc_slot.push_back(0); cyc_order.push_back(0); }

        o.blob.clear();
        o.off_leaf_tnode = put(o.blob, leaf_tnode);
        o.off_parent = put(o.blob, t_parent);
        o.off_depth = put(o.blob, t_depth);
        o.off_tcycle = put(o.blob, t_cycle);
        o.off_pslot = put(o.blob, t_pslot);
        o.off_blk = put(o.blob, blk);
        o.off_cyc_k = put(o.blob, cyc_k);
        o.off_cyc_off = put(o.blob, cyc_off);
        o.off_cyc_slot = put(o.blob, cyc_slot);
        o.off_cyc_gslot = put(o.blob, cyc_gslot);
        o.off_cyc_flags = put(o.blob, cyc_flags);
        o.off_cyc_hyb = put(o.blob, cyc_hyb);
        o.off_torder = put(o.blob, t_order);
        o.off_cyc_order = put(o.blob, cyc_order);
        // prefix-sum rows: one D row (and one -2D row) per internal blob-tree node
        std::vector<char> isLeafT(nt, 0);
        for (int n = 0; n < N; n++) if (isLeaf(n)) isLeafT[tOfNode[n]] = 1;
        std::vector<uint16_t> dslot(nt, 0xFFFF), dnode;
        for (int v = 0; v < nt; v++) if (!isLeafT[v]) { dslot[v] = (uint16_t)(o.n_xe + 1 + (int)dnode.size()); dnode.push_back((uint16_t)v); }
        o.n_int = (int)dnode.size();
        o.n_full = o.n_xe + 1 + o.n_int;
        std::vector<uint16_t> crow(std::max(H, 1), 0);
        for (int h = 0; h < H; h++) { crow[h] = (uint16_t)o.n_full; o.n_full += SNAQ_CYC_ROWS((int)cycNodes[h].size()); }
        if (o.n_full > 65535) fail(SNAQ_EINVAL, "network too large for 16-bit slots");
        o.off_dslot = put(o.blob, dslot);
        o.off_dnode = put(o.blob, dnode);
        o.off_crow = put(o.blob, crow);
        std::vector<uint32_t> dpath_off(1, 0);
        std::vector<uint16_t> dpath_slot;
        for (uint16_t v0 : dnode) {
            for (int v = v0; t_parent[v] != v; v = t_parent[v]) dpath_slot.push_back(t_pslot[v]);
            dpath_off.push_back((uint32_t)dpath_slot.size());
        }
        o.off_dpath_off = put(o.blob, dpath_off);
        o.off_dpath_slot = put(o.blob, dpath_slot);
        // leaf deletion order statistics per cycle block (hasEdge parity descriptor)
        std::vector<int> leafOrderOfTaxon(std::max(ntaxa, 1), -1);
        for (int i = 0; i < f->n_leaves; i++) {
            int n = f->leaf[i];
            if (n < 0 || n >= N || !isLeaf(n)) fail(SNAQ_EINVAL, "leaf list entry is not a leaf node");
            leafOrderOfTaxon[f->node_taxon[n]] = i;
        }
        std::vector<uint16_t> blkmax(cyc_slot.size(), 0), hmax2(std::max(H, 1), 0xFFFF), hmaxtax(std::max(H, 1), 0);
        for (int h = 0; h < H; h++) {
            int k = (int)cycNodes[h].size();
            std::vector<int> best(k, -1), second(k, -1), besttax(k, -1);
            for (int t = 0; t < ntaxa; t++) {
                if (nodeOfTaxon[t] < 0) continue;
                int j = blk[(size_t)h * ntaxa + t], lo = leafOrderOfTaxon[t];
                if (lo > best[j]) { second[j] = best[j]; best[j] = lo; besttax[j] = t; }
                else if (lo > second[j]) second[j] = lo;
            }
            // a block that contains another cycle is not reduced to a leaf hanging on the cycle node: the
            // merge of src/pseudolik.jl:241-251 cannot fire for it; 0 makes the cascade stop there
            for (int j = 0; j < k; j++) blkmax[cyc_off[h] + j] = blockHasCycle[h][j] ? 0 : (uint16_t)(std::max(best[j], 0) + 1);
            hmax2[h] = second[0] < 0 ? 0xFFFF : (uint16_t)(second[0] + 1);
            hmaxtax[h] = (uint16_t)std::max(besttax[0], 0);
        }
        o.off_blkmax = put(o.blob, blkmax);
        o.off_hmax2 = put(o.blob, hmax2);
        o.off_hmaxtax = put(o.blob, hmaxtax);

        // ---- the network for the hasEdge simulation (snaq_hesim.h): adjacency and flags in net.node / net.edge positions --
        {
            std::vector<int> nodeOfNumber;                      // hybrid node number -> node index (inCycle carries numbers)
            std::vector<int16_t> s_node_edge((size_t)N * 3, -1), s_incyc(N, -1), s_taxon(N, -1), s_edge_node((size_t)E * 2, -1);
            std::vector<uint8_t> s_nfl(N, 0), s_efl(E, 0);
            for (int n = 0; n < N; n++) {
                for (int j = 0; j < 3; j++) s_node_edge[(size_t)3 * n + j] = (int16_t)f->node_edge[3 * n + j];
                const uint32_t fl = f->node_flags[n];
                s_nfl[n] = (uint8_t)(((fl & SNAQ_NODE_LEAF) ? SIM_N_LEAF : 0u) | ((fl & SNAQ_NODE_HYBRID) ? SIM_N_HYB : 0u) |
                                     ((fl & SNAQ_NODE_HASHYBEDGE) ? SIM_N_HASHYB : 0u) | ((fl & SNAQ_NODE_BADDIAMOND1) ? SIM_N_BD1 : 0u) |
                                     ((fl & SNAQ_NODE_BADDIAMOND2) ? SIM_N_BD2 : 0u) | ((fl & SNAQ_NODE_BADTRIANGLE) ? SIM_N_BADTRI : 0u));
                s_taxon[n] = (int16_t)(isLeaf(n) ? f->node_taxon[n] : -1);
                if (f->node_incycle[n] != -1)
                    for (int h = 0; h < H; h++) if (f->node_number[f->hybrid[h]] == f->node_incycle[n]) s_incyc[n] = (int16_t)f->hybrid[h];
            }
            for (int e = 0; e < E; e++) {
                s_edge_node[(size_t)2 * e] = (int16_t)f->edge_node[2 * e];
                s_edge_node[(size_t)2 * e + 1] = (int16_t)f->edge_node[2 * e + 1];
                const uint32_t fl = f->edge_flags[e];
                s_efl[e] = (uint8_t)(((fl & SNAQ_EDGE_HYBRID) ? SIM_E_HYB : 0u) | ((fl & SNAQ_EDGE_MAJOR) ? SIM_E_MAJOR : 0u) |
                                     ((fl & SNAQ_EDGE_IDENTIFIABLE) ? SIM_E_IDENT : 0u) | ((fl & SNAQ_EDGE_FROMBD1) ? SIM_E_FROMBD1 : 0u) |
                                     (f->edge_incycle[e] != -1 ? SIM_E_INCYC : 0u));
            }
            std::vector<int16_t> s_leaf(std::max(f->n_leaves, 1), 0), s_hyb(std::max(H, 1), 0);
            for (int i = 0; i < f->n_leaves; i++) s_leaf[i] = (int16_t)f->leaf[i];
            for (int h = 0; h < H; h++) s_hyb[h] = (int16_t)f->hybrid[h];
            // parameter slots in x order: gamma | t | gammaz (parameters!, src/snaq_optimization.jl:64-91)
            std::vector<int16_t> s_ref(std::max(o.nparams, 1), -1), s_alias(std::max(o.nparams, 1), -1);
            auto pseudo = [&](int node, int ind) { return std::stoi(std::to_string(f->node_number[node]) + std::to_string(ind)); };
            int ig = 0, it = 0, iz = 0;
            for (int e = 0; e < E; e++) {
                if (eIdent(e)) {
                    const int s = o.nh + it++;
                    s_ref[s] = (int16_t)e;
                    for (int h = 0; h < H; h++) {
                        const int hn = f->hybrid[h];
                        if (!(f->node_flags[hn] & SNAQ_NODE_BADDIAMOND1)) continue;
                        for (int ind = 1; ind <= 2; ind++) if (pseudo(hn, ind) == f->edge_number[e]) s_alias[s] = (int16_t)(2 * hn + ind - 1);
                    }
                }
                if (eHyb(e) && !eMaj(e)) {
                    const int node = hybChild(e);
                    if (!(f->node_flags[node] & SNAQ_NODE_BADDIAMOND1)) s_ref[ig++] = (int16_t)node;
                    else {
                        for (int ind = 1; ind <= 2; ind++) {
                            const int s = o.nh + o.nt + iz++;
                            s_ref[s] = (int16_t)node;
                            for (int e2 = 0; e2 < E; e2++) if (f->edge_number[e2] == pseudo(node, ind) && s_alias[s] < 0) s_alias[s] = (int16_t)e2;
                        }
                    }
                }
            }
            o.sim = SnaqSimNet{};
            o.sim.n_nodes = N; o.sim.n_edges = E; o.sim.n_leaves = f->n_leaves; o.sim.n_hybrids = H;
            o.sim.nparams = o.nparams; o.sim.nh = o.nh; o.sim.nt = o.nt;
            o.sim_blob.clear();
            o.so_node_edge = put(o.sim_blob, s_node_edge);
            o.so_node_flags = put(o.sim_blob, s_nfl);
            o.so_node_incycle = put(o.sim_blob, s_incyc);
            o.so_node_taxon = put(o.sim_blob, s_taxon);
            o.so_edge_node = put(o.sim_blob, s_edge_node);
            o.so_edge_flags = put(o.sim_blob, s_efl);
            o.so_leaf_order = put(o.sim_blob, s_leaf);
            o.so_hyb_order = put(o.sim_blob, s_hyb);
            o.so_slot_ref = put(o.sim_blob, s_ref);
            o.so_slot_alias = put(o.sim_blob, s_alias);
            o.so_edge_number = put(o.sim_blob, std::vector<int32_t>(f->edge_number, f->edge_number + E));
            o.so_node_number = put(o.sim_blob, std::vector<int32_t>(f->node_number, f->node_number + N));
        }
        return SNAQ_OK;
    } catch (const Fail &e) {
        err = e.msg;
        return e.code;
    } catch (const std::exception &e) {
        err = e.what();
        return SNAQ_EINVAL;
    }
}

bool snaq_tables_touched(const SnaqHostTables &a, const SnaqHostTables &b, std::vector<uint8_t> &touched) {
    if (a.ntaxa != b.ntaxa || a.n_tnodes != b.n_tnodes || a.n_cycles != b.n_cycles || a.nparams != b.nparams || a.n_xe != b.n_xe ||
        a.n_int != b.n_int || a.n_full != b.n_full || a.root != b.root || a.blob.size() != b.blob.size())
        return false;
    const SnaqTables A = a.view(a.blob.data()), B = b.view(b.blob.data());
    const int nt = a.n_tnodes, nc = a.n_cycles, n = a.ntaxa;
    auto same = [&](const void *x, const void *y, size_t bytes) { return std::memcmp(x, y, bytes) == 0; };
    size_t ncs = 0;
    for (int c = 0; c < nc; c++) ncs = std::max<size_t>(ncs, (size_t)A.cyc_off[c] + A.cyc_k[c]);
    // everything a program depends on that is not a property of one taxon's root path
    if (!same(A.leaf_tnode, B.leaf_tnode, sizeof(uint16_t) * n) || !same(A.tcycle, B.tcycle, sizeof(int16_t) * nt) ||
        !same(A.dslot, B.dslot, sizeof(uint16_t) * nt) || !same(A.torder, B.torder, sizeof(uint16_t) * nt) ||
        !same(A.cyc_k, B.cyc_k, sizeof(uint16_t) * std::max(nc, 1)) || !same(A.cyc_off, B.cyc_off, sizeof(uint16_t) * std::max(nc, 1)) ||
        !same(A.cyc_gslot, B.cyc_gslot, sizeof(uint16_t) * std::max(nc, 1)) || !same(A.cyc_flags, B.cyc_flags, std::max(nc, 1)) ||
        !same(A.cyc_hyb, B.cyc_hyb, std::max(nc, 1)) || !same(A.crow, B.crow, sizeof(uint16_t) * std::max(nc, 1)))
        return false;
    for (int c = 0; c < nc; c++) {
        const size_t o0 = A.cyc_off[c], k = A.cyc_k[c];
        if (!same(A.cyc_slot + o0, B.cyc_slot + o0, sizeof(uint16_t) * k) || !same(A.cyc_order + o0, B.cyc_order + o0, sizeof(uint16_t) * k))
            return false;
    }
    touched.assign((size_t)n, 0);
    for (int t = 0; t < n; t++) {
        int u = A.leaf_tnode[t];
        if (u == 0xFFFF) continue;
        bool ch = false;
        for (int c = 0; c < nc && !ch; c++) ch = A.blk[(size_t)c * n + t] != B.blk[(size_t)c * n + t];
        for (int v = u; !ch; v = A.parent[v]) {
            if (A.parent[v] != B.parent[v] || A.depth[v] != B.depth[v] || A.pslot[v] != B.pslot[v]) ch = true;
            if (A.parent[v] == v) break;
        }
        touched[(size_t)t] = ch ? 1 : 0;
    }
    return true;
}
