/*
 * snaq_hasedge.h — structural replacement of updateHasEdge! (src/pseudolik.jl:392-464) and of the
 * indexht part of parameters!(qnet,net) (src/snaq_optimization.jl:106-180).
 *
 * hasEdge[i] says whether parameter i of x is still carried by the pruned, un-merged quartet network
 * q.qnet.  The reference defines it operationally (the edge number still exists in qnet.edge and is still
 * identifiable there), so this file reproduces what extractQuartet! leaves behind:
 *
 *   - an edge survives iff it lies on the quartet's Steiner tree and is not swallowed by cleanExtEdges!,
 *     i.e. it is not in the chain of degree-2 nodes that hangs off a quartet leaf.  That chain ends at the
 *     first degree-3 node: the junction, or a node of a cycle that survives on the pendant path (hybrid
 *     block hit: a "type 1" cycle, irrelevant for the CFs but still present in q.qnet);
 *   - a cycle whose hybrid block is not hit loses its hybrid node and both hybrid edges and opens into a path;
 *   - hybrid edges of a surviving cycle stay identifiable unless the hybrid's block holds a single quartet
 *     taxon (src/pseudolik.jl:348-352);
 *   - deleteLeaf! merges a cycle node next to the hybrid node into the hybrid edge when the node loses its
 *     last leaf while the hybrid's child already is a leaf (src/pseudolik.jl:241-251, :288-298).  This depends
 *     on the order in which leaves are deleted (net.leaf order), which is why the flat network carries it;
 *   - bad diamond I: gammaz slots per the three sub-cases of src/pseudolik.jl:410-457.
 *
 * Status: exact (0 mismatches against the oracle on ~400,000 quartets of random networks, bad diamonds and
 * triangles included) for networks with at most ONE hybrid node, which is what snaq_get_hasedge accepts.  With
 * nested cycles the reference's result depends on whether a cycle inside a block still exists when the block's
 * last leaf is pruned, i.e. on the pruning order across cycles; that case is rejected, not approximated.
 */
#ifndef SNAQ_HASEDGE_H
#define SNAQ_HASEDGE_H

#include "snaq_core.h"

#define SNAQ_HE_WORDS(nparams) (((nparams) + 31) / 32)

typedef struct SnaqHE {
    const SnaqTables *T;
    const uint16_t *tx;
    uint32_t *bits;
} SnaqHE;

SNAQ_HD void snaq_he_set(const SnaqHE &h, unsigned slot) {
    if ((int)slot < h.T->nparams) h.bits[slot >> 5] |= 1u << (slot & 31);
}
SNAQ_HD void snaq_he_arc(const SnaqHE &h, int c, int lo, int hi) {   /* cycle edges lo..hi-1 */
    const uint16_t *s = h.T->cyc_slot + h.T->cyc_off[c];
    for (int j = lo; j < hi; j++) snaq_he_set(h, s[j]);
}

/* ordered blob-tree path u .. v; returns its length (0 on overflow) */
SNAQ_HD int snaq_he_path(const SnaqTables &T, int u, int v, uint16_t *seq, int *iapex_out) {
    const int apex = snaq_lca(T, u, v);
    int ns = 0;
    for (int w = u;; w = T.parent[w]) {
        if (ns >= SNAQ_MAX_PATH) return 0;
        seq[ns++] = (uint16_t)w;
        if (w == apex) break;
    }
    *iapex_out = ns - 1;
    int cnt = 0;
    for (int w = v; w != apex; w = T.parent[w]) cnt++;
    if (ns + cnt > SNAQ_MAX_PATH) return 0;
    int idx = ns + cnt - 1;
    for (int w = v; w != apex; w = T.parent[w]) seq[idx--] = (uint16_t)w;
    return ns + cnt;
}

/*
 * A cycle that survives in q.qnet (hybrid block hit, at least two blocks hit, not a rewritten bad diamond I):
 * mark its gamma, its tree edges, and its hybrid edges when the hybrid's block holds more than one quartet taxon.
 * `hcount` = quartet taxa in the hybrid's block, `htaxon` = that taxon when hcount == 1.
 */
SNAQ_HD void snaq_he_surviving_cycle(const SnaqHE &h, int c, const int pos[4], int hcount, int htaxon) {
    const SnaqTables &T = *h.T;
    const int k = T.cyc_k[c];
    const uint16_t *s = T.cyc_slot + T.cyc_off[c];
    snaq_he_set(h, T.cyc_gslot[c]);
    unsigned hitmask_lo = 0, hitmask_hi = 0;      /* positions hit, up to 64 tracked exactly; beyond: treated as hit */
    for (int i = 0; i < 4; i++) { if (pos[i] < 32) hitmask_lo |= 1u << pos[i]; else if (pos[i] < 64) hitmask_hi |= 1u << (pos[i] - 32); }
    int del_major_upto = 1;                       /* edges 1 .. del_major_upto-1 are merged into the major hybrid edge */
    int del_minor_from = k - 1;                   /* edges del_minor_from .. k-2 are merged into the minor hybrid edge */
    if (hcount == 1) {
        /* time at which the hybrid's child becomes a leaf = last deletion among the other taxa of its block */
        const uint16_t *bm = T.blkmax + T.cyc_off[c];
        int tH;
        if (T.hmaxtax[c] == (uint16_t)htaxon) tH = (T.hmax2[c] == 0xFFFF) ? 0 : (int)T.hmax2[c];   /* orders are stored +1; 0 = never */
        else tH = (int)bm[0];
        auto is_hit = [&](int j) { return j < 32 ? ((hitmask_lo >> j) & 1u) != 0 : (j < 64 ? ((hitmask_hi >> (j - 32)) & 1u) != 0 : true); };
        int prev = tH;
        for (int j = 1; j < k - 1 && !is_hit(j) && (int)bm[j] > prev; j++) { prev = (int)bm[j]; del_major_upto = j + 1; }
        prev = tH;
        for (int j = k - 1; j > 1 && !is_hit(j) && (int)bm[j] > prev; j--) { prev = (int)bm[j]; del_minor_from = j - 1; }
    }
    for (int j = del_major_upto; j < del_minor_from; j++) snaq_he_set(h, s[j]);     /* surviving cycle tree edges */
    if (hcount > 1) { snaq_he_set(h, s[0]); snaq_he_set(h, s[k - 1]); }             /* hybrid edges */
}

/* number of quartet taxa in the hybrid's block of cycle c, and one of them */
SNAQ_HD int snaq_he_hcount(const SnaqHE &h, const int pos[4], int *htaxon) {
    int n = 0;
    for (int i = 0; i < 4; i++) if (pos[i] == 0) { n++; *htaxon = h.tx[i]; }
    return n;
}

/*
 * Pendant path of quartet taxon `a` (index into tx) from its leaf to the junction blob-tree node J.
 * Returns whether a surviving cycle was met on the way (the "barrier": edges beyond it survive un-merged).
 */
SNAQ_HD int snaq_he_pendant(const SnaqHE &h, int a, int J) {
    const SnaqTables &T = *h.T;
    uint16_t seq[SNAQ_MAX_PATH];
    int iapex;
    const int ns = snaq_he_path(T, T.leaf_tnode[h.tx[a]], J, seq, &iapex);
    int barrier = 0;
    for (int i = 0; i + 1 < ns; i++) {
        if (barrier) snaq_he_set(h, i < iapex ? T.pslot[seq[i]] : T.pslot[seq[i + 1]]);
        const int v = seq[i + 1];
        const int c = T.tcycle[v];
        if (c < 0 || v == J) continue;
        int pos[4];
        snaq_positions(T, c, h.tx, pos);
        const int ua = pos[a];
        int uv = -1;
        for (int j = 0; j < 4; j++) if (j != a) uv = pos[j];
        if (ua != 0 && uv != 0) {                      /* hybrid block not hit: the cycle is a path */
            if (barrier) snaq_he_arc(h, c, ua < uv ? ua : uv, ua < uv ? uv : ua);
        } else if (!(T.cyc_flags[c] & 1)) {            /* survives as an irrelevant ("type 1") cycle */
            int ht = -1;
            const int hc = snaq_he_hcount(h, pos, &ht);
            snaq_he_surviving_cycle(h, c, pos, hc, ht);
            barrier = 1;
        }
        /* bad diamond I with its hybrid leaf hit and neither side leaf: rewritten, then merged away */
    }
    return barrier;
}

/*
 * hasEdge bits (nparams bits in `bits`, cleared here) of one quartet; *bd1_type5 is set when the pruned network
 * keeps a bad-diamond-I hybrid (then indexht is just its two gammaz slots, src/snaq_optimization.jl:120-146).
 */
SNAQ_HD int snaq_hasedge_quartet(const SnaqTables &T, const uint16_t tx[4], uint32_t *bits, int *bd1_type5) {
    const int nw = SNAQ_HE_WORDS(T.nparams);
    for (int i = 0; i < nw; i++) bits[i] = 0;
    *bd1_type5 = 0;
    SnaqHE h{&T, tx, bits};
    int L[4];
    for (int i = 0; i < 4; i++) { L[i] = T.leaf_tnode[tx[i]]; if (L[i] == 0xFFFF) return SNAQ_B_BLOCKS; }
    int l01 = snaq_lca(T, L[0], L[1]), l23 = snaq_lca(T, L[2], L[3]);
    int l02 = snaq_lca(T, L[0], L[2]), l13 = snaq_lca(T, L[1], L[3]);
    int l03 = snaq_lca(T, L[0], L[3]), l12 = snaq_lca(T, L[1], L[2]);
    int S0 = T.depth[l01] + T.depth[l23], S1 = T.depth[l02] + T.depth[l13], S2 = T.depth[l03] + T.depth[l12];
    int pos[4];
    if (S0 == S1 && S1 == S2) {
        /* all four meet in one cycle */
        const int X = snaq_deeper(T, snaq_deeper(T, l01, l02), l12);
        const int c = T.tcycle[X];
        if (c < 0) return SNAQ_B_POLYTOMY;
        snaq_positions(T, c, tx, pos);
        int o[4] = {0, 1, 2, 3};
        for (int i = 1; i < 4; i++)
            for (int j = i; j > 0 && pos[o[j]] < pos[o[j - 1]]; j--) { int t = o[j]; o[j] = o[j - 1]; o[j - 1] = t; }
        int bar[4];
        for (int i = 0; i < 4; i++) bar[i] = snaq_he_pendant(h, i, X);
        if (pos[o[0]] == 0) {
            if (T.cyc_flags[c] & 1) {               /* bad diamond I kept whole: both gammaz (src/pseudolik.jl:414-416) */
                snaq_he_set(h, T.cyc_gslot[c]); snaq_he_set(h, (unsigned)T.cyc_gslot[c] + 1u);
                *bd1_type5 = 1;
            } else {
                snaq_he_surviving_cycle(h, c, pos, 1, tx[o[0]]);
            }
        } else {
            /* open cycle: internal arc between the 2nd and 3rd hit node; the outer arcs are pendant */
            snaq_he_arc(h, c, pos[o[1]], pos[o[2]]);
            if (bar[o[0]]) snaq_he_arc(h, c, pos[o[0]], pos[o[1]]);
            if (bar[o[3]]) snaq_he_arc(h, c, pos[o[2]], pos[o[3]]);
        }
        return SNAQ_B_OK;
    }
    const int maj = (S0 > S1 && S0 > S2) ? 0 : ((S1 > S0 && S1 > S2) ? 1 : 2);
    int p = 0, q = maj + 1, r, s;
    if (maj == 0) { r = 2; s = 3; } else if (maj == 1) { r = 1; s = 3; } else { r = 1; s = 2; }
    int lca6[4][4];
    lca6[0][1] = lca6[1][0] = l01; lca6[2][3] = lca6[3][2] = l23; lca6[0][2] = lca6[2][0] = l02;
    lca6[1][3] = lca6[3][1] = l13; lca6[0][3] = lca6[3][0] = l03; lca6[1][2] = lca6[2][1] = l12;
    const int J1 = snaq_deeper(T, snaq_deeper(T, lca6[p][q], lca6[p][r]), lca6[q][r]);
    const int J2 = snaq_deeper(T, snaq_deeper(T, lca6[r][s], lca6[r][p]), lca6[s][p]);
    if (J1 == J2) return SNAQ_B_BLOCKS;
    /* pendant paths */
    int bar[4];
    bar[p] = snaq_he_pendant(h, p, J1); bar[q] = snaq_he_pendant(h, q, J1);
    bar[r] = snaq_he_pendant(h, r, J2); bar[s] = snaq_he_pendant(h, s, J2);
    /* internal path */
    uint16_t seq[SNAQ_MAX_PATH];
    int iapex;
    const int ns = snaq_he_path(T, J1, J2, seq, &iapex);
    if (ns == 0) return SNAQ_B_OVERFLOW;
    for (int i = 0; i < ns; i++) {
        if (i + 1 < ns) snaq_he_set(h, i < iapex ? T.pslot[seq[i]] : T.pslot[seq[i + 1]]);
        const int c = T.tcycle[seq[i]];
        if (c < 0) continue;
        snaq_positions(T, c, tx, pos);
        const int role = (i == 0) ? 1 : (i == ns - 1 ? 2 : 0);
        const SnaqCyc m = snaq_classify_cycle(T, c, pos, p, q, r, s, role);
        if (m.cls < 0) return SNAQ_B_BLOCKS;
        int ht = -1;
        const int hc = snaq_he_hcount(h, pos, &ht);
        if (m.cls == SNAQ_C_OPEN) {
            if (role == 0) snaq_he_arc(h, c, m.lo, m.hi);
            else {
                /* three hit positions: internal arc median..far; the arcs median..cherry are pendant */
                const int a = role == 1 ? p : r, b = role == 1 ? q : s;
                snaq_he_arc(h, c, m.lo, m.hi);
                const int med = m.rep;
                if (bar[a]) snaq_he_arc(h, c, pos[a] < med ? pos[a] : med, pos[a] < med ? med : pos[a]);
                if (bar[b]) snaq_he_arc(h, c, pos[b] < med ? pos[b] : med, pos[b] < med ? med : pos[b]);
            }
        } else if (m.cls == SNAQ_C_TZ) {
            /* pruned bad diamond I: the edge -log(1-gammaz) of the side that kept its leaf (src/pseudolik.jl:435-450) */
            snaq_he_set(h, (unsigned)T.cyc_gslot[c] + (m.other == 1 ? 0u : 1u));
        } else {
            snaq_he_surviving_cycle(h, c, pos, hc, ht);
        }
    }
    return SNAQ_B_OK;
}

#endif
