/*
 * snaq_core.h — the per-quartet device functions of the engine.
 *
 *   snaq_build_quartet : structural replacement of extractQuartet! +
 *                        identifyQuartet! + eliminateHybridization! +
 *                        updateSplit!/updateFormula!  (reference:
 *                        src/pseudolik.jl:472-540, :746-848, :866-1145, :1207-1247)
 *   snaq_eval_program  : calculateExpCF! / quartetType5! (src/pseudolik.jl:1255-1275,
 *                        :967-1052) for one parameter vector
 *   snaq_quartet_loglik: logPseudoLik(quartet) (src/pseudolik.jl:1388-1410)
 *
 * The reference prunes a deep copy of the network leaf by leaf for every quartet.
 * Here the network is summarised once per topology into small tables (blob tree
 * with parent/depth, per-cycle ordered edge slots, per-(cycle,taxon) block index)
 * and each quartet is classified in O(depth) from them: a level-1 network is a
 * tree of vertex-disjoint cycles, every cycle node owns one "block" of taxa, and
 * the way the four taxa fall into the blocks of the cycles on their connecting
 * paths determines the reference's case.  The result is a short
 * program of u16 words over the slots of the extended parameter vector
 *   xe = [ x (gamma | t | gammaz, parameters!(net)) | constants ]
 * that the evaluation kernels interpret for every candidate x.  Before a batch is evaluated, x is
 * expanded once per candidate (snaq_expand_x) into
 *   xf = [ xe | 0.0 (NULL slot) | D[v] for internal blob-tree nodes | per-cycle rows ]
 * where D[v] is the sum of the edge slots on the path from v to the root, so that the tree part
 * of an internal path J1..J2 is (D[J1] - D[apex]) + (D[J2] - D[apex]); the per-cycle rows are the
 * prefix sums C_j of the cycle's edge lengths plus the quartet-independent type-3 / gammaz terms.
 *
 * Everything here is `SNAQ_HD` so that tests/hostsim can run the exact same code
 * on the CPU against the oracle; the shipped library only runs it on the GPU.
 */
#ifndef SNAQ_CORE_H
#define SNAQ_CORE_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "snaq_math_tables.h"

#if defined(__CUDACC__)
#define SNAQ_HD __host__ __device__ __forceinline__
#define SNAQ_HD_NOINLINE __host__ __device__ __noinline__
#if defined(__CUDA_ARCH__)
#define SNAQ_UNROLL1 _Pragma("unroll 1")
#else
#define SNAQ_UNROLL1
#endif
#else
#define SNAQ_HD inline
#define SNAQ_HD_NOINLINE inline
#define SNAQ_UNROLL1
#endif

/* Lanes that enter a loop with a data-dependent trip count together are made to leave it together: without the explicit
 * barrier the compiler lets the early finishers run ahead and the rest of the builder executes once per group of lanes. */
#if defined(__CUDA_ARCH__)
#define SNAQ_LANES()        __activemask()
#define SNAQ_REJOIN(m)      __syncwarp(m)
#define SNAQ_ANY(m, c)      __any_sync(m, c)      /* loop while any of the lanes has work: the lanes step through it together */
#else
#define SNAQ_LANES()        0u
#define SNAQ_REJOIN(m)      ((void)(m))
#define SNAQ_ANY(m, c)      (c)
#endif

#define SNAQ_MAX_HYB 32          /* max hybrid nodes tracked in the parity descriptor */
#define SNAQ_MAX_SLOT 4095       /* slots are 12 bit inside term headers */
#define SNAQ_MAX_PATH 512        /* max blob-tree nodes on a quartet's internal path */

/*
 * Quartet programs.  Every quartet has one header BYTE (kept in a separate array, only the
 * weight / read-back kernels need it) and a program of u16 words, padded with the NULL slot
 * (index n_xe, value 0.0) to a multiple of 4 words so that programs are 8-byte aligned.
 * Sums are lists of signed PAIRS (p, n) meaning xf[p] - xf[n]: a path of blob-tree edges is
 * (D[J], D[apex]), an arc of an open cycle is (C_hi, C_lo), a single row r is (r, NULL).
 *
 *   class A  tree-like, additive      : [p n]*             t1 = min(10, sum xf[p] - xf[n])
 *            (path edges as prefix-sum rows, open cycle arcs as C rows, type-3 and gammaz terms as rows)
 *   class B  tree-like, type 2/4 terms: [npairs | nterms<<8] [p n]*npairs [term]*nterms
 *            T2 term: [T2|flag|gslot] [C_hi] [C_lo]
 *            T4 term: [T4|far|gslot]  [C_lo] [C_hi] [C_k] [nchainpairs] [p n]*nchainpairs
 *   class C  4-cycle (which==2)       : T5   : [gslot] [C_a] [C_m] [C_b]
 *                                       T5BD1: [zslot1] [zslot2]
 * The device table is sorted by class so that a warp / block only ever runs one of the loops.
 */
#define SNAQ_KIND_TREE   0u      /* which==1: one internal edge of length t1        */
#define SNAQ_KIND_T5     1u      /* which==2: 4-cycle, gamma + two cycle segments   */
#define SNAQ_KIND_T5BD1  2u      /* which==2: bad diamond I, two gammaz             */
/* header byte: bits 0-1 kind | bits 2-4 perm (tree: position of the major CF; type 5:
 * permutation index) | bit 5: has terms (class B)                                   */
#define SNAQ_QHDR(kind, perm, terms) ((uint8_t)((kind) | ((perm) << 2) | ((terms) << 5)))
#define SNAQ_HDR_KIND(h)   ((h) & 3u)
#define SNAQ_HDR_PERM(h)   (((h) >> 2) & 7u)
#define SNAQ_HDR_TERMS(h)  (((h) >> 5) & 1u)
#define SNAQ_HDR_CLASS(h)  (SNAQ_HDR_KIND(h) == SNAQ_KIND_TREE ? SNAQ_HDR_TERMS(h) : 2u)
/* term header: bits 0-1 type | bit 2 flag | bits 4-15 slot */
#define SNAQ_T2   0u   /* 3-cycle, hybrid's child is a leaf  (eliminateTriangle! case 1) */
#define SNAQ_T3   1u   /* 2-cycle on the internal path       (eliminateLoop! internal)   */
#define SNAQ_T4   2u   /* 3-cycle, hybrid's child internal   (eliminateTriangle! case 2) */
#define SNAQ_TZ   3u   /* edge -log(1-gammaz) left by a pruned bad diamond I             */
#define SNAQ_TERM(type, flag, slot) ((uint16_t)((type) | ((flag) << 2) | ((slot) << 4)))

/* build errors */
#define SNAQ_B_OK          0
#define SNAQ_B_POLYTOMY    1   /* four taxa meet at a tree node: not a binary level-1 network */
#define SNAQ_B_BLOCKS      2   /* block pattern inconsistent with the blob tree                */
#define SNAQ_B_OVERFLOW    3   /* more pieces/terms than the header can count                  */

/* evaluation status bits (OR-ed into a device status word) */
#define SNAQ_S_NEGLEN      1u  /* merged length < -log(1.5): src/auxiliary.jl:120 */
#define SNAQ_S_ZEROSUM     2u  /* sum(expCF)==0: src/pseudolik.jl:1391            */
#define SNAQ_S_NAN         4u

#define SNAQ_PENALTY_SLOPE 1.0e6
#define SNAQ_MINLEN (-0.4054651081081644)
#define SNAQ_LOG3   1.0986122886681098

typedef struct SnaqTables {
    int32_t ntaxa, n_tnodes, n_cycles, nparams, n_xe, n_hybrids;
    int32_t n_int, n_full;       /* internal blob-tree nodes; rows of the expanded vector      */
    /* blob tree T': cycles contracted to supernodes, rooted arbitrarily */
    const uint16_t *leaf_tnode;  /* [ntaxa]     T' node of each taxon (0xFFFF: taxon not in net) */
    const uint16_t *parent;      /* [n_tnodes]  parent (root: itself)                     */
    const uint16_t *depth;       /* [n_tnodes]                                            */
    const int16_t  *tcycle;      /* [n_tnodes]  cycle index of a supernode, else -1       */
    const uint16_t *pslot;       /* [n_tnodes]  xe slot of the edge to the parent         */
    const uint16_t *torder;      /* [n_tnodes]  position of a tree node in net.node       */
    const uint16_t *dnode;       /* [n_int]     blob-tree node of the k-th D row           */
    const uint16_t *dslot;       /* [n_tnodes]  slot of D[v] (root-path prefix sum of the edge
                                    slots above v) for internal nodes; 0xFFFF for leaves     */
    /* cycles: position 0 = hybrid node, 1 = its major parent, ..., k-1 = minor parent.
     * edge j joins positions j and (j+1)%k: j=0 major hybrid edge, j=k-1 minor one.      */
    const uint8_t  *blk;         /* [n_cycles*ntaxa] position whose block holds the taxon */
    const uint16_t *cyc_k;       /* [n_cycles]                                            */
    const uint16_t *cyc_off;     /* [n_cycles]  offset of the k edge slots in cyc_slot    */
    const uint16_t *cyc_slot;    /* edge-length slots                                     */
    const uint16_t *cyc_gslot;   /* [n_cycles]  slot of the minor gamma; bad diamond I:
                                    slot of gammaz(major side), minor side is +1         */
    const uint8_t  *cyc_flags;   /* [n_cycles]  bit0 bad diamond I                        */
    const uint8_t  *cyc_hyb;     /* [n_cycles]  index of the hybrid in net.hybrid order   */
    const uint16_t *cyc_order;   /* per cycle position (same offsets as cyc_slot): position of
                                    that cycle node in net.node                            */
    /* rows of the expanded vector that belong to a cycle with k nodes (see snaq_expand_row):
     *   crow[c] + j          j=0..k   C_j  = sum of the lengths of cycle edges 0..j-1
     *   crow[c] + k+1 + o-1  o=1..k-1 type-3 term when the hit node is at position o
     *   crow[c] + 2k + i     i=0,1    -log(1-gammaz_i) capped at 10 (bad diamond I only)          */
    const uint16_t *crow;        /* [n_cycles] */
    /* the D rows as flat slot lists (same order as the walk from the node up to the root): no pointer chasing
     * when a kernel prologue expands x */
    const uint32_t *dpath_off;   /* [n_int+1] */
    const uint16_t *dpath_slot;
    /* leaf deletion order (net.leaf order; only the hasEdge parity descriptor needs it): */
    const uint16_t *blkmax;      /* per cycle position (offsets as cyc_slot): largest net.leaf index in that block */
    const uint16_t *hmax2;       /* [n_cycles] second largest net.leaf index in the hybrid's block, 0xFFFF if single */
    const uint16_t *hmaxtax;     /* [n_cycles] taxon with the largest net.leaf index in the hybrid's block */
    /* optional (NULL: walk the parent pointers): blob-tree LCA of every pair of taxa, [ntaxa*ntaxa], filled on the device
     * once per topology (k_leaf_lca) so that the per-quartet builder does six lookups instead of six walks */
    const uint16_t *leaf_lca;
} SnaqTables;
#define SNAQ_CROW(T, c, j)   ((unsigned)(T).crow[c] + (unsigned)(j))
#define SNAQ_T3ROW(T, c, o)  ((unsigned)(T).crow[c] + (unsigned)(T).cyc_k[c] + (unsigned)(o))
#define SNAQ_TZROW(T, c, i)  ((unsigned)(T).crow[c] + 2u * (unsigned)(T).cyc_k[c] + (unsigned)(i))
#define SNAQ_CYC_ROWS(k)     (2 * (k) + 2)

/* parity descriptor of one quartet */
typedef struct SnaqQuartetInfo {
    int8_t which;                 /* 1 or 2                                          */
    int8_t formula[3];            /* which==1: 1 major/2 minor; which==2: 11,12,13   */
    int8_t type[SNAQ_MAX_HYB];    /* per hybrid (net.hybrid order): 0 or 2..5        */
} SnaqQuartetInfo;

typedef struct SnaqWriter {
    uint16_t *out;   /* NULL: count only */
    int n, cap;
} SnaqWriter;

SNAQ_HD void snaq_put(SnaqWriter &w, unsigned v) {
    if (w.out && w.n < w.cap) w.out[w.n] = (uint16_t)v;
    w.n++;
}

SNAQ_HD int snaq_lca(const SnaqTables &T, int u, int v) {
    while (T.depth[u] > T.depth[v]) u = T.parent[u];
    while (T.depth[v] > T.depth[u]) v = T.parent[v];
    while (u != v) { u = T.parent[u]; v = T.parent[v]; }
    return u;
}

/* LCA of the blob-tree leaves of taxa tx[i], tx[j] */
SNAQ_HD int snaq_lca_taxa(const SnaqTables &T, const uint16_t tx[4], const int L[4], int i, int j) {
    if (T.leaf_lca) return T.leaf_lca[(size_t)tx[i] * (size_t)T.ntaxa + tx[j]];
    return snaq_lca(T, L[i], L[j]);
}

/* emit the slots of cycle edges j = lo .. hi-1 (the arc between positions lo and hi) */
SNAQ_HD int snaq_put_pair(SnaqWriter &w, unsigned plus, unsigned minus) {
    snaq_put(w, plus);
    snaq_put(w, minus);
    return 1;
}
/* the arc between cycle positions lo and hi as the pair (C_hi, C_lo); returns the number of pairs */
SNAQ_HD int snaq_put_arc(const SnaqTables &T, int c, int lo, int hi, SnaqWriter &w) {
    if (hi <= lo) return 0;
    return snaq_put_pair(w, SNAQ_CROW(T, c, hi), SNAQ_CROW(T, c, lo));
}

SNAQ_HD int snaq_deeper(const SnaqTables &T, int a, int b) { return T.depth[a] >= T.depth[b] ? a : b; }

/* positions of the four taxa in cycle c */
SNAQ_HD void snaq_positions(const SnaqTables &T, int c, const uint16_t tx[4], int pos[4]) {
    const uint8_t *b = T.blk + (size_t)c * T.ntaxa;
    for (int i = 0; i < 4; i++) pos[i] = b[tx[i]];
}

/*
 * Classification of a supernode (cycle c) met on the internal path of a tree-like
 * quartet.  (p,q) are the taxa on the J1 side, (r,s) on the J2 side; role 0 =
 * interior, 1 = the supernode is J1 (p,q in different blocks), 2 = it is J2.
 */
#define SNAQ_C_BAD    (-1)
#define SNAQ_C_OPEN   0   /* hybrid block not hit: the cycle is a plain path (arc lo..hi may be empty) */
#define SNAQ_C_T2     2
#define SNAQ_C_T3     3
#define SNAQ_C_T4     4
#define SNAQ_C_TZ     6
typedef struct SnaqCyc {
    int cls;      /* SNAQ_C_*                                                          */
    int lo, hi;   /* OPEN: arc of plain pieces; T2: arc e; T3: hit position in lo; T4: the two hit positions */
    int other;    /* T2/TZ: position of `other`                                         */
    int rep;      /* position of the cycle node that ends up as the degree-3 junction (ends only) */
} SnaqCyc;

SNAQ_HD SnaqCyc snaq_classify_cycle(const SnaqTables &T, int c, const int pos[4], int p, int q, int r, int s, int role) {
    SnaqCyc o;
    o.cls = SNAQ_C_BAD; o.lo = o.hi = o.other = o.rep = 0;
    const int k = T.cyc_k[c];
    const int bd1 = T.cyc_flags[c] & 1;
    if (role == 0) {
        int u = pos[p], v = pos[r];
        if (pos[q] != u || pos[s] != v || u == v) return o;
        if (u != 0 && v != 0) { o.cls = SNAQ_C_OPEN; o.lo = u < v ? u : v; o.hi = u < v ? v : u; return o; }
        if (bd1) return o;                            /* a bad diamond I hybrid has a single leaf below it */
        o.cls = SNAQ_C_T3; o.lo = u ? u : v;          /* eliminateLoop!(internal=true) */
        return o;
    }
    /* end of the path: cherry taxa (a,b) in two different blocks, the far taxa in a third */
    int a = role == 1 ? p : r, b = role == 1 ? q : s, f1 = role == 1 ? r : p, f2 = role == 1 ? s : q;
    int u1 = pos[a], u2 = pos[b], wv = pos[f1];
    if (pos[f2] != wv || u1 == u2 || u1 == wv || u2 == wv) return o;
    int lo = u1 < u2 ? u1 : u2, hi = u1 < u2 ? u2 : u1;
    if (u1 != 0 && u2 != 0 && wv != 0) {              /* hybrid block not hit: path 1..k-1 */
        int m = wv < lo ? lo : (wv > hi ? hi : wv);   /* median of the three positions: the junction */
        o.cls = SNAQ_C_OPEN; o.lo = m < wv ? m : wv; o.hi = m < wv ? wv : m; o.rep = m;
        return o;
    }
    if (wv == 0) {                                    /* type 4: the hybrid block holds the two far taxa */
        if (bd1) return o;
        o.cls = SNAQ_C_T4; o.lo = lo; o.hi = hi;
        /* `other` = first of the two hit nodes in qnet.node order (src/pseudolik.jl:791-796) */
        const uint16_t *ord = T.cyc_order + T.cyc_off[c];
        o.rep = ord[lo] < ord[hi] ? lo : hi;
        return o;
    }
    o.other = u1 ? u1 : u2;                           /* type 2: hybrid block holds one cherry taxon */
    if (bd1) {                                        /* pruned bad diamond I: src/pseudolik.jl:206-239 */
        if (k != 4 || wv != 2) return o;
        o.cls = SNAQ_C_TZ; o.rep = 0;
        return o;
    }
    o.cls = SNAQ_C_T2; o.rep = o.other;
    o.lo = o.other < wv ? o.other : wv; o.hi = o.other < wv ? wv : o.other;
    return o;
}

/* which of cf(12|34), cf(13|24), cf(14|23) a pair of roles (1..4) selects: 1,2,3 */
SNAQ_HD int snaq_pair_class(int r1, int r2) {
    int lo = r1 < r2 ? r1 : r2, hi = r1 < r2 ? r2 : r1;
    if ((lo == 1 && hi == 2) || (lo == 3 && hi == 4)) return 1;
    if ((lo == 1 && hi == 3) || (lo == 2 && hi == 4)) return 2;
    return 3;
}
/* permutation index <-> (p0,p1,p2), a permutation of (1,2,3) */
SNAQ_HD int snaq_perm_index(int p0, int p1) {
    return (p0 - 1) * 2 + ((p1 < ((p0 == 1) ? 3 : (p0 == 2 ? 3 : 2))) ? 0 : 1);
}
SNAQ_HD void snaq_perm_decode(int idx, int perm[3]) {
    /* {1,2,3},{1,3,2},{2,1,3},{2,3,1},{3,1,2},{3,2,1} packed 2 bits per entry (no local array) */
    const unsigned packed[3] = {0xFA5u /*p0: 1,1,2,2,3,3*/, 0x9DEu /*p1: 2,3,1,3,1,2*/, 0x67Bu /*p2: 3,2,3,1,2,1*/};
    perm[0] = (int)((packed[0] >> (2 * idx)) & 3u);
    perm[1] = (int)((packed[1] >> (2 * idx)) & 3u);
    perm[2] = (int)((packed[2] >> (2 * idx)) & 3u);
}

/*
 * Build the evaluation program of one quartet.  tx = the four taxon ids in the
 * data's order.  Returns SNAQ_B_*; w.n is the program length in words.
 */
SNAQ_HD int snaq_build_quartet(const SnaqTables &T, const uint16_t tx[4], SnaqWriter &w, uint8_t *hdr, SnaqQuartetInfo *info) {
    int L[4];
    for (int i = 0; i < 4; i++) {
        L[i] = T.leaf_tnode[tx[i]];
        if (L[i] == 0xFFFF) return SNAQ_B_BLOCKS;
    }
    if (info) {
        info->which = 1;
        info->formula[0] = info->formula[1] = info->formula[2] = 2;
        for (int i = 0; i < SNAQ_MAX_HYB; i++) info->type[i] = 0;
    }
    int l01 = snaq_lca_taxa(T, tx, L, 0, 1), l23 = snaq_lca_taxa(T, tx, L, 2, 3);
    int l02 = snaq_lca_taxa(T, tx, L, 0, 2), l13 = snaq_lca_taxa(T, tx, L, 1, 3);
    int l03 = snaq_lca_taxa(T, tx, L, 0, 3), l12 = snaq_lca_taxa(T, tx, L, 1, 2);
    int S0 = T.depth[l01] + T.depth[l23], S1 = T.depth[l02] + T.depth[l13], S2 = T.depth[l03] + T.depth[l12];
    int pos[4];
    *hdr = 0;
    if (S0 == S1 && S1 == S2) {
        /* the four taxa meet in one blob: it must be a cycle with four hit blocks */
        int X = snaq_deeper(T, snaq_deeper(T, l01, l02), l12);
        int c = T.tcycle[X];
        if (c < 0) return SNAQ_B_POLYTOMY;
        snaq_positions(T, c, tx, pos);
        /* sort the four data indices by position */
        int o[4] = {0, 1, 2, 3};
        for (int i = 1; i < 4; i++)
            for (int j = i; j > 0 && pos[o[j]] < pos[o[j - 1]]; j--) { int t = o[j]; o[j] = o[j - 1]; o[j - 1] = t; }
        if (pos[o[0]] == pos[o[1]] || pos[o[1]] == pos[o[2]] || pos[o[2]] == pos[o[3]]) return SNAQ_B_BLOCKS;
        if (pos[o[0]] == 0) {
            /* type 5 (quartetType5!): leaf1 below the hybrid, leaf2 / leaf3 at the first hit node
             * on the major / minor side, leaf4 in between */
            int role[4];
            role[o[0]] = 1; role[o[1]] = 2; role[o[3]] = 3; role[o[2]] = 4;
            int p0 = snaq_pair_class(role[0], role[1]), p1 = snaq_pair_class(role[0], role[2]);
            int p2 = snaq_pair_class(role[0], role[3]);
            int pidx = snaq_perm_index(p0, p1);
            if (info) {
                info->which = 2;
                info->formula[0] = (int8_t)(10 + p0); info->formula[1] = (int8_t)(10 + p1); info->formula[2] = (int8_t)(10 + p2);
                info->type[T.cyc_hyb[c]] = 5;
            }
            const int g = T.cyc_gslot[c];
            if (T.cyc_flags[c] & 1) {
                snaq_put(w, g);
                snaq_put(w, g + 1);
                *hdr = SNAQ_QHDR(SNAQ_KIND_T5BD1, pidx, 0);
            } else {
                int a = pos[o[1]], m = pos[o[2]], b = pos[o[3]];
                snaq_put(w, g);
                snaq_put(w, SNAQ_CROW(T, c, a));     /* y5 = exp(-(C_m - C_a)), y6 = exp(-(C_b - C_m)) */
                snaq_put(w, SNAQ_CROW(T, c, m));
                snaq_put(w, SNAQ_CROW(T, c, b));
                *hdr = SNAQ_QHDR(SNAQ_KIND_T5, pidx, 0);
            }
            return SNAQ_B_OK;
        }
        /* hybrid block not hit: the cycle is a path, internal edge between the 2nd and 3rd hit node */
        int n0 = snaq_put_arc(T, c, pos[o[1]], pos[o[2]], w);
        int mate = (o[0] == 0) ? o[1] : (o[1] == 0 ? o[0] : (o[2] == 0 ? o[3] : o[2]));   /* partner of data taxon 0 */
        int maj = mate - 1;
        if (n0 > 255) return SNAQ_B_OVERFLOW;
        if (info) info->formula[maj] = 1;
        *hdr = SNAQ_QHDR(SNAQ_KIND_TREE, maj, 0);
        return SNAQ_B_OK;
    }
    /* resolved split: (p,q | r,s) */
    int maj = (S0 > S1 && S0 > S2) ? 0 : ((S1 > S0 && S1 > S2) ? 1 : 2);
    int p = 0, q = maj + 1, r, s;
    if (maj == 0) { r = 2; s = 3; } else if (maj == 1) { r = 1; s = 3; } else { r = 1; s = 2; }
    int lca6[4][4];
    lca6[0][1] = lca6[1][0] = l01; lca6[2][3] = lca6[3][2] = l23; lca6[0][2] = lca6[2][0] = l02;
    lca6[1][3] = lca6[3][1] = l13; lca6[0][3] = lca6[3][0] = l03; lca6[1][2] = lca6[2][1] = l12;
    int J1 = snaq_deeper(T, snaq_deeper(T, lca6[p][q], lca6[p][r]), lca6[q][r]);
    int J2 = snaq_deeper(T, snaq_deeper(T, lca6[r][s], lca6[r][p]), lca6[s][p]);
    if (J1 == J2) return SNAQ_B_BLOCKS;
    int apex;
    if (T.leaf_lca) {
        /* J1 is the LCA of a pair that holds taxon a, J2 of a pair that holds c: lca(J1, J2) is the shallowest of J1, J2 and
         * lca(a, c) (an ancestor of the other is itself the answer, else lca(a, c) lies strictly above both; candidates of equal
         * depth that share a descendant leaf are the same node) */
        const int a = (J1 == lca6[p][q] || J1 == lca6[p][r]) ? p : q;
        const int c = (J2 == lca6[r][s] || J2 == lca6[r][p]) ? r : s;
        const int lac = lca6[a][c];
        apex = T.depth[J1] <= T.depth[J2] ? J1 : J2;
        if (T.depth[lac] < T.depth[apex]) apex = lac;
    } else apex = snaq_lca(T, J1, J2);
    /* the internal path in order J1 .. apex .. J2 */
    /* (no `return` inside a loop with a data-dependent trip count: the lanes of a warp would not reconverge behind it) */
    uint16_t seq[SNAQ_MAX_PATH];
    int ns = 0, err = SNAQ_B_OK;
    const unsigned lanes = SNAQ_LANES();
    for (int v = J1;; v = T.parent[v]) {
        if (ns >= SNAQ_MAX_PATH) { err = SNAQ_B_OVERFLOW; break; }
        seq[ns++] = (uint16_t)v;
        if (v == apex) break;
    }
    SNAQ_REJOIN(lanes);
    const int iapex = ns - 1;
    int cnt = 0;
    for (int v = J2; v != apex; v = T.parent[v]) cnt++;
    SNAQ_REJOIN(lanes);
    if (ns + cnt > SNAQ_MAX_PATH) err = SNAQ_B_OVERFLOW;
    if (!err) {
        int idx = ns + cnt - 1;
        for (int v = J2; v != apex; v = T.parent[v]) seq[idx--] = (uint16_t)v;
        ns += cnt;
    }
    SNAQ_REJOIN(lanes);
    if (err) return err;
    /* classify the two ends; find the type-4 ends (the only terms that can be negative, so the only
     * ones for which the reference's merge order matters once the 10.0 cap is reached) */
    SnaqCyc e1, e2;
    e1.cls = e2.cls = SNAQ_C_OPEN; e1.lo = e1.hi = e2.lo = e2.hi = 0; e1.rep = e2.rep = 0; e1.other = e2.other = 0;
    const int c1 = T.tcycle[J1], c2 = T.tcycle[J2];
    const unsigned lanes1 = SNAQ_LANES();
    if (c1 >= 0) { snaq_positions(T, c1, tx, pos); e1 = snaq_classify_cycle(T, c1, pos, p, q, r, s, 1); if (e1.cls < 0) err = SNAQ_B_BLOCKS; }
    SNAQ_REJOIN(lanes1);
    if (c2 >= 0) { snaq_positions(T, c2, tx, pos); e2 = snaq_classify_cycle(T, c2, pos, p, q, r, s, 2); if (e2.cls < 0) err = SNAQ_B_BLOCKS; }
    SNAQ_REJOIN(lanes1);
    if (err) return err;
    const int t4a = (c1 >= 0 && e1.cls == SNAQ_C_T4), t4b = (c2 >= 0 && e2.cls == SNAQ_C_T4);
    /* chains: edges 0..ja belong to the type-4 term at J1, edges jb..ns-2 to the one at J2 */
    int ja = -1, jb = ns - 1, nearA = 1;
    int aArc = 0, bArc = 0;   /* the chain runs through the open end cycle at the other junction */
    const unsigned lanes3 = SNAQ_LANES();
    if (t4a || t4b) {
        /* cleanUpNode! runs (and pre-merges the path below each hybrid node up to the next degree-3
         * node) iff the pruned quartet network holds more than one hybrid: src/pseudolik.jl:826-830 */
        int nhyb = 0;
        for (int c = 0; c < T.n_cycles; c++) {
            if (T.cyc_flags[c] & 1) continue;
            int ps[4];
            snaq_positions(T, c, tx, ps);
            int hit = (ps[0] == 0) | (ps[1] == 0) | (ps[2] == 0) | (ps[3] == 0);
            int same = (ps[0] == ps[1]) & (ps[1] == ps[2]) & (ps[2] == ps[3]);
            nhyb += (hit && !same);
        }
        const int cleanup = nhyb > 1;
        if (t4a) {
            ja = 0;
            if (cleanup) {
                while (ja < ns - 2) {          /* extend through degree-2 nodes: anything but a type-3 loop */
                    int c = T.tcycle[seq[ja + 1]];
                    if (c >= 0) {
                        snaq_positions(T, c, tx, pos);
                        SnaqCyc m = snaq_classify_cycle(T, c, pos, p, q, r, s, 0);
                        if (m.cls < 0) { err = SNAQ_B_BLOCKS; break; }
                        if (m.cls == SNAQ_C_T3) break;
                    }
                    ja++;
                }
                aArc = (ja == ns - 2);
            }
        }
        if (t4b) {
            jb = ns - 2;
            if (cleanup) {
                while (jb > 0) {
                    int c = T.tcycle[seq[jb]];
                    if (c >= 0) {
                        snaq_positions(T, c, tx, pos);
                        SnaqCyc m = snaq_classify_cycle(T, c, pos, p, q, r, s, 0);
                        if (m.cls < 0) { err = SNAQ_B_BLOCKS; break; }
                        if (m.cls == SNAQ_C_T3) break;
                    }
                    jb--;
                }
                bArc = (jb == 0);
            }
        }
        /* which end is internalLength!'s `node` (first degree-3 node in qnet.node): src/pseudolik.jl:1078 */
        int o1 = c1 >= 0 ? T.cyc_order[T.cyc_off[c1] + e1.rep] : T.torder[J1];
        int o2 = c2 >= 0 ? T.cyc_order[T.cyc_off[c2] + e2.rep] : T.torder[J2];
        nearA = o1 < o2;
        if (t4a && t4b && ja >= jb) {
            /* both ends are type 4 and share one chain: the hybrid that comes first in net.hybrid is
             * eliminated first and takes the whole pre-merged chain */
            if (T.cyc_hyb[c1] < T.cyc_hyb[c2]) { ja = ns - 2; jb = ns - 1; nearA = 1; }
            else { jb = 0; ja = -1; nearA = 0; }
            aArc = bArc = 0;
        }
    }
    SNAQ_REJOIN(lanes3);
    if (err) return err;
    /* owner of a path element: 0 main list, 1 chain of the term at J1, 2 chain of the term at J2.
     * element kinds: edge i (between seq[i] and seq[i+1]); arc of interior node i; end arcs.
     * Terms evaluated on the fly (class B): type 2 and type 4.  Type 3 and gammaz terms are rows. */
    int nfly_pre = (c1 >= 0 && (e1.cls == SNAQ_C_T2 || e1.cls == SNAQ_C_T4)) + (c2 >= 0 && (e2.cls == SNAQ_C_T2 || e2.cls == SNAQ_C_T4));
    const int cnt_word_at = w.n;
    if (nfly_pre > 0) snaq_put(w, 0);
    int n0 = 0, nterms = 0;
    /* additive part */
    const unsigned NUL = (unsigned)T.n_xe;
    if (!t4a && !t4b) {
        /* all blob-tree edges of the path at once, as prefix sums: (D[J1] - D[apex]) + (D[J2] - D[apex]) */
        if (J1 != apex) n0 += snaq_put_pair(w, T.dslot[J1], T.dslot[apex]);
        if (J2 != apex) n0 += snaq_put_pair(w, T.dslot[J2], T.dslot[apex]);
    }
    const unsigned lanes2 = SNAQ_LANES();
    for (int i = 0; SNAQ_ANY(lanes2, i < ns && !err); i++) {
        if (i < ns && !err) do {
            if (i < ns - 1 && (t4a || t4b)) {
                int owner = (t4a && i <= ja) ? 1 : ((t4b && i >= jb) ? 2 : 0);
                if (owner == 0) n0 += snaq_put_pair(w, i < iapex ? T.pslot[seq[i]] : T.pslot[seq[i + 1]], NUL);
            }
            int c = T.tcycle[seq[i]];
            if (c < 0) break;
            SnaqCyc m;
            int owner = 0;
            if (i == 0) { m = e1; owner = (t4b && bArc) ? 2 : 0; }
            else if (i == ns - 1) { m = e2; owner = (t4a && aArc) ? 1 : 0; }
            else {
                snaq_positions(T, c, tx, pos);
                m = snaq_classify_cycle(T, c, pos, p, q, r, s, 0);
                if (m.cls < 0) { err = SNAQ_B_BLOCKS; break; }
                owner = (t4a && i <= ja) ? 1 : ((t4b && i >= jb + 1) ? 2 : 0);
            }
            if (m.cls == SNAQ_C_OPEN) {
                if (owner == 0) n0 += snaq_put_arc(T, c, m.lo, m.hi, w);
            } else if (m.cls == SNAQ_C_T3) {
                n0 += snaq_put_pair(w, SNAQ_T3ROW(T, c, m.lo), NUL);
                if (info) info->type[T.cyc_hyb[c]] = 3;
            } else if (m.cls == SNAQ_C_TZ) {
                /* the edge -log(1-gammaz) of a pruned bad diamond I already exists when cleanUpNode! runs: if the chain of a
                 * type-4 term at the other end reaches this junction, the edge is merged with that chain (owner != 0) */
                if (owner == 0) n0 += snaq_put_pair(w, SNAQ_TZROW(T, c, m.other == 1 ? 0 : 1), NUL);
            }
        } while (0);
        SNAQ_REJOIN(lanes2);
    }
    SNAQ_REJOIN(lanes2);
    if (err) return err;
    /* on-the-fly terms, in path order */
    for (int i = 0; i < ns; i += (ns > 1 ? ns - 1 : 1)) {
        int c = T.tcycle[seq[i]];
        if (c < 0) continue;
        const SnaqCyc m = (i == 0) ? e1 : e2;
        if (m.cls != SNAQ_C_T2 && m.cls != SNAQ_C_T4) continue;
        const int k = T.cyc_k[c];
        const int g = T.cyc_gslot[c];
        nterms++;
        if (m.cls == SNAQ_C_T2) {
            /* gamma of the hybrid edge that lands on `other`: major iff `other` comes first */
            snaq_put(w, SNAQ_TERM(SNAQ_T2, m.other == m.lo ? 1 : 0, g));
            snaq_put(w, SNAQ_CROW(T, c, m.hi)); snaq_put(w, SNAQ_CROW(T, c, m.lo));
            if (info) info->type[T.cyc_hyb[c]] = 2;
        } else {                                   /* type 4 */
            const int isA = (i == 0);
            const int far = isA ? !nearA : nearA;
            snaq_put(w, SNAQ_TERM(SNAQ_T4, far, g));
            /* major side = C_lo (edges 0..lo-1), between the hit nodes = C_hi - C_lo, minor side = C_k - C_hi */
            snaq_put(w, SNAQ_CROW(T, c, m.lo)); snaq_put(w, SNAQ_CROW(T, c, m.hi)); snaq_put(w, SNAQ_CROW(T, c, k));
            const int cnt_at = w.n;
            snaq_put(w, 0);
            /* the chain below the hybrid node that the reference merges with this term first */
            int nchain = 0;
            const int me = isA ? 1 : 2;
            for (int j = 0; j < ns; j++) {
                if (j < ns - 1) {
                    int owner = (t4a && j <= ja) ? 1 : ((t4b && j >= jb) ? 2 : 0);
                    if (owner == me) nchain += snaq_put_pair(w, j < iapex ? T.pslot[seq[j]] : T.pslot[seq[j + 1]], NUL);
                }
                int cc = T.tcycle[seq[j]];
                if (cc < 0) continue;
                if (j == 0) {
                    if (e1.cls == SNAQ_C_OPEN && me == 2 && bArc) nchain += snaq_put_arc(T, cc, e1.lo, e1.hi, w);
                    if (e1.cls == SNAQ_C_TZ && me == 2 && bArc) nchain += snaq_put_pair(w, SNAQ_TZROW(T, cc, e1.other == 1 ? 0 : 1), NUL);
                } else if (j == ns - 1) {
                    if (e2.cls == SNAQ_C_OPEN && me == 1 && aArc) nchain += snaq_put_arc(T, cc, e2.lo, e2.hi, w);
                    if (e2.cls == SNAQ_C_TZ && me == 1 && aArc) nchain += snaq_put_pair(w, SNAQ_TZROW(T, cc, e2.other == 1 ? 0 : 1), NUL);
                } else {
                    int owner = (t4a && j <= ja) ? 1 : ((t4b && j >= jb + 1) ? 2 : 0);
                    if (owner != me) continue;
                    snaq_positions(T, cc, tx, pos);
                    SnaqCyc mm = snaq_classify_cycle(T, cc, pos, p, q, r, s, 0);
                    if (mm.cls == SNAQ_C_OPEN) nchain += snaq_put_arc(T, cc, mm.lo, mm.hi, w);
                }
            }
            if (nchain > 65535) err = SNAQ_B_OVERFLOW;
            if (w.out && cnt_at < w.cap) w.out[cnt_at] = (uint16_t)nchain;
            if (info) info->type[T.cyc_hyb[c]] = 4;
        }
    }
    if (err) return err;
    if (n0 > 255 || nterms > 255 || nterms != nfly_pre) return SNAQ_B_OVERFLOW;
    if (info) info->formula[maj] = 1;
    if (nterms > 0 && w.out && cnt_word_at < w.cap) w.out[cnt_word_at] = (uint16_t)(n0 | (nterms << 8));
    *hdr = SNAQ_QHDR(SNAQ_KIND_TREE, maj, nterms > 0 ? 1 : 0);
    return SNAQ_B_OK;
}

/* ------------------------------------------------------------------------------
 * FP64 exp / log specialised to the ranges of this path (fewer FP64-pipe instructions than the
 * generic library versions; <= 2e-16 relative error for exp on [-12,1], <= 2e-16 absolute or
 * relative error, whichever is larger, for log on normal positive arguments).
 * ------------------------------------------------------------------------------ */
/* polynomial coefficients and other FP64 constants: on the device they live in __constant__ memory */
#define SNAQ_KC_INIT { SNAQ_16_LN2, -SNAQ_LN2_16_HI, -SNAQ_LN2_16_LO, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, \
                       1.0 / 6.0, SNAQ_LN2, SNAQ_LOG3, -2.0 / 3.0, 6755399441055744.0,                                   \
                       -0.1, 1.0 / 9.0, -0.125, 1.0 / 7.0, -1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0 }
#define SNAQ_KC_N 20
#if defined(__CUDACC__)
__constant__ double snaq_kc_dev[SNAQ_KC_N] = SNAQ_KC_INIT;
#endif
static const double snaq_kc_host[SNAQ_KC_N] = SNAQ_KC_INIT;
#if defined(__CUDA_ARCH__)
#define SNAQ_KC(i) snaq_kc_dev[i]
#else
#define SNAQ_KC(i) snaq_kc_host[i]
#endif
#define SNAQ_K_LN2   SNAQ_KC(8)
#define SNAQ_K_LOG3  SNAQ_KC(9)
#define SNAQ_K_M23   SNAQ_KC(10)

typedef struct SnaqMath {
    const double *et;   /* [16] 2^(j/16)                                  */
    const double *li;   /* [16] 1/c_j,        c_j = 1+(j+.5)/16           */
    const double *lc;   /* [16] -log(1/c_j)                               */
} SnaqMath;
#define SNAQ_MATH_DOUBLES 48   /* the three tables, contiguous, 128-byte aligned */

#if defined(__CUDA_ARCH__)
SNAQ_HD int snaq_hi(double v) { return __double2hiint(v); }
SNAQ_HD int snaq_lo(double v) { return __double2loint(v); }
SNAQ_HD double snaq_mk(int hi, int lo) { return __hiloint2double(hi, lo); }
#else
SNAQ_HD int snaq_hi(double v) { int64_t b; memcpy(&b, &v, 8); return (int)(b >> 32); }
SNAQ_HD int snaq_lo(double v) { int64_t b; memcpy(&b, &v, 8); return (int)(b & 0xffffffffLL); }
SNAQ_HD double snaq_mk(int hi, int lo) { int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double v; memcpy(&v, &b, 8); return v; }
#endif

/* exp(u) for u in [-12, 1]: u = k*ln2/16 + r, exp(u) = 2^(k>>4) * 2^((k&15)/16) * exp(r), |r| <= ln2/32 */
SNAQ_HD double snaq_exp(double u, const SnaqMath &M) {
    const double magic = SNAQ_KC(11);                        /* 1.5 * 2^52 */
    const double kd = fma(u, SNAQ_KC(0), magic);
    const int k = snaq_lo(kd);
    const double kf = kd - magic;
    double r = fma(kf, SNAQ_KC(1), u);
    r = fma(kf, SNAQ_KC(2), r);
    double q = fma(r, SNAQ_KC(3), SNAQ_KC(4));
    q = fma(r, q, SNAQ_KC(5));
    q = fma(r, q, SNAQ_KC(6));
    q = fma(r, q, SNAQ_KC(7));
    q = fma(r, q, 0.5);
    q = fma(r, q, 1.0);
    const double tj = M.et[k & 15];
    const double v = fma(tj, r * q, tj);
    return snaq_mk(snaq_hi(v) + ((k >> 4) << 20), snaq_lo(v));
}

/* rare arguments (0, negative, denormal, inf, nan): the library version, kept out of line */
static SNAQ_HD_NOINLINE double snaq_log_slow(double z) { return log(z); }

/* log(z): z = 2^e m, m in [1,2); r = m/c_j - 1 with |r| <= 2^-5; log z = e ln2 - log(1/c_j) + log1p(r).
 * snaq_log_pos: z must be a normal positive number (e.g. the major CF of a tree-like quartet, in [1/3,1]) */
SNAQ_HD double snaq_log_pos(double z, const SnaqMath &M) {
    const int hi = snaq_hi(z);
    const int e = (hi >> 20) - 1023;
    const int j = (hi >> 16) & 15;
    const double m = snaq_mk((hi & 0x000FFFFF) | 0x3FF00000, snaq_lo(z));
    const double r = fma(m, M.li[j], -1.0);                  /* |r| <= 1/32 */
    double q = fma(r, SNAQ_KC(12), SNAQ_KC(13));             /* log1p(r) = r - r^2/2 + ... - r^10/10 */
    q = fma(r, q, SNAQ_KC(14));
    q = fma(r, q, SNAQ_KC(15));
    q = fma(r, q, SNAQ_KC(16));
    q = fma(r, q, SNAQ_KC(17));
    q = fma(r, q, SNAQ_KC(18));
    q = fma(r, q, SNAQ_KC(19));
    q = fma(r, q, -0.5);
    const double p = fma(r * r, q, r);
    return fma((double)e, SNAQ_K_LN2, M.lc[j] + p);
}
SNAQ_HD double snaq_log(double z, const SnaqMath &M) {
    if (!(z >= 2.2250738585072014e-308 && z <= 1.7976931348623157e308)) return snaq_log_slow(z);
    return snaq_log_pos(z, M);
}

/*
 * One row of the expanded parameter vector xf (see the top of this file) for candidate x:
 *   row < nparams            : x[row]
 *   row < n_xe               : constant (length of a non-identifiable internal edge)
 *   row == n_xe              : 0.0 (NULL slot)
 *   row <  n_xe+1+n_int      : D[v]   = sum of x over the edges from v up to the root
 *   else                     : the rows of a cycle (SnaqTables::crow): prefix sums C_j of its edge lengths, its
 *                              type-3 terms, and -log(1-gammaz) for a bad diamond I
 */
SNAQ_HD double snaq_clamp10(double v) { return v > 10.0 ? 10.0 : v; }   /* setLength!: src/auxiliary.jl:122-124 */

/* g / 3.0, correctly rounded (bit for bit the quotient the reference's `g*1/3` produces), without the division sequence:
 * q = RN(g * RN(1/3)) is within one ulp, the residual g - 3q is exact in an FMA, and one correction step gives the
 * correctly rounded quotient (Markstein); three FP64-pipe instructions instead of about twenty.  Checked against g/3.0 on
 * 2e5 values in tests/test_builder_cpu.py. */
SNAQ_HD double snaq_div3(double g) {
    const double rc = 0.33333333333333331483;      /* RN(1/3) */
    const double q = g * rc;
    const double r = fma(-3.0, q, g);
    return fma(r, rc, q);
}

SNAQ_HD double snaq_xe(const SnaqTables &T, const double *x, const double *xconst, int sl) {
    return sl < T.nparams ? x[sl] : xconst[sl - T.nparams];
}
SNAQ_HD double snaq_expand_row(const SnaqTables &T, const SnaqMath &M, const double *x, const double *xconst, int row) {
    if (row < T.n_xe) return snaq_xe(T, x, xconst, row);
    if (row == T.n_xe) return 0.0;
    int k = row - T.n_xe - 1;
    if (k < T.n_int) {
        double d = 0.0;
        const uint32_t i1 = T.dpath_off[k + 1];
        for (uint32_t i = T.dpath_off[k]; i < i1; i++) d += snaq_xe(T, x, xconst, T.dpath_slot[i]);
        return d;
    }
    /* cycle rows: find the cycle (few cycles: linear search) */
    int c = 0;
    while (c + 1 < T.n_cycles && row >= (int)T.crow[c + 1]) c++;
    const int kc = T.cyc_k[c];
    const uint16_t *es = T.cyc_slot + T.cyc_off[c];
    int j = row - (int)T.crow[c];
    if (j <= kc) {                               /* C_j */
        double d = 0.0;
        for (int i = 0; i < j; i++) d += snaq_xe(T, x, xconst, es[i]);
        return d;
    }
    j -= kc + 1;
    if (j < kc - 1) {                            /* type-3 term, hit node at position o = j+1 (src/pseudolik.jl:878) */
        if (T.cyc_flags[c] & 1) return 0.0;
        const int o = j + 1;
        double l1 = 0.0, l2 = 0.0;
        for (int i = 0; i < o; i++) l1 += snaq_xe(T, x, xconst, es[i]);
        for (int i = o; i < kc; i++) l2 += snaq_xe(T, x, xconst, es[i]);
        l1 = snaq_clamp10(l1); l2 = snaq_clamp10(l2);
        const double g2 = x[T.cyc_gslot[c]], g1 = 1.0 - g2;
        return snaq_clamp10(-snaq_log(1.0 - g1 * g1 * (1.0 - snaq_exp(-l1, M)) - g2 * g2 * (1.0 - snaq_exp(-l2, M)), M));
    }
    j -= kc - 1;                                  /* -log(1-gammaz), capped at 10 (src/snaq_optimization.jl:230-234) */
    if (!(T.cyc_flags[c] & 1)) return 0.0;
    return snaq_clamp10(-snaq_log(1.0 - x[T.cyc_gslot[c] + j], M));
}

/* ------------------------------------------------------------------------------
 * evaluation.  XF is a functor slot -> double (the candidate's xe, NULL slot = 0.0).
 * ------------------------------------------------------------------------------ */
/* sum of n signed pairs */
template <class XF>
SNAQ_HD double snaq_sum(const uint16_t *&pc, int n, const XF &x) {
    double a = 0.0;
    SNAQ_UNROLL1
    for (int i = 0; i < n; i++) a += x(pc[2 * i]) - x(pc[2 * i + 1]);
    pc += 2 * n;
    return a;
}

/* class A: t1 = min(10, sum of the slots); nwords counts the NULL padding too */
template <class XF>
SNAQ_HD double snaq_tree_plain_t1(const uint16_t *pc, int nwords, const XF &x) {
    double t = 0.0;
    for (int i = 0; i < nwords; i += 2) t += x(pc[i]) - x(pc[i + 1]);
    return snaq_clamp10(t);
}

/* class B: t1 of a tree-like quartet with type-2 / type-4 terms.
 * t1 = clamp( clamp(Bnear + R) + Bfar ): R = every non-negative piece (any merge order of those gives
 * min(10, sum)); B = a type-4 term merged with its chain, which can be negative and is added first
 * (near end) or last (far end) by internalLength! (src/pseudolik.jl:1076-1109) */
template <class XF>
SNAQ_HD double snaq_tree_terms_t1(const uint16_t *pc, const XF &x, const SnaqMath &M, unsigned &status) {
    const unsigned cw = *pc++;
    double R = snaq_sum(pc, (int)(cw & 255u), x);
    double Bn = 0.0, Bf = 0.0;
    const int nterms = (int)(cw >> 8);
    SNAQ_UNROLL1
    for (int it = 0; it < nterms; it++) {
        const unsigned th = *pc++;
        const double gq = x(th >> 4);          /* minor gamma */
        if ((th & 3u) == SNAQ_T2) {
            const double e = snaq_clamp10(x(pc[0]) - x(pc[1]));
            pc += 2;
            const double gh = (th & 4u) ? 1.0 - gq : gq;
            R += snaq_clamp10(-snaq_log(1.0 - gh * (1.0 - snaq_exp(-e, M)), M));                 /* src/pseudolik.jl:938 */
        } else {
            const double clo = x(pc[0]), chi = x(pc[1]), ck = x(pc[2]);
            const double la = snaq_clamp10(clo);
            const double lb = snaq_clamp10(ck - chi);
            const double le = snaq_clamp10(chi - clo);
            const int nchain = pc[3];
            pc += 4;
            const double ch = snaq_clamp10(snaq_sum(pc, nchain, x));
            const double ga = 1.0 - gq, gb = gq;
            double l4 = -snaq_log(gb * gb * snaq_exp(-lb, M) + ga * gb * (3.0 - snaq_exp(-le, M)) + ga * ga * snaq_exp(-la, M), M);   /* :951 */
            if (l4 < SNAQ_MINLEN) status |= SNAQ_S_NEGLEN;
            l4 = snaq_clamp10(l4);
            const double B = snaq_clamp10(ch + l4);                               /* :956 */
            if (B < SNAQ_MINLEN) status |= SNAQ_S_NEGLEN;
            if (th & 4u) Bf = B; else Bn = B;
        }
    }
    double t = snaq_clamp10(Bn + R);
    if (t < SNAQ_MINLEN) status |= SNAQ_S_NEGLEN;
    t = snaq_clamp10(t + Bf);
    if (t < SNAQ_MINLEN) status |= SNAQ_S_NEGLEN;
    return t;
}

/* tree-like CFs from t1: cf[0] major, cf[1] = cf[2] minor (src/pseudolik.jl:1259) */
SNAQ_HD void snaq_tree_cf(double t, const SnaqMath &M, double cf[3]) {
    const double y = snaq_exp(-t, M);
    cf[0] = fma(SNAQ_K_M23, y, 1.0);
    cf[1] = 1.0 / 3.0 * y;
    cf[2] = cf[1];
}

/* class C: cf1, cf2, cf3 of quartetType5! (src/pseudolik.jl:985-991) */
template <class XF>
SNAQ_HD void snaq_t5_cf(unsigned kind, const uint16_t *pc, const XF &x, const SnaqMath &M, double cf[3]) {
    if (kind == SNAQ_KIND_T5) {
        const double g2 = x(pc[0]), g1 = 1.0 - g2;
        const double ca = x(pc[1]), cm = x(pc[2]), cb = x(pc[3]);
        const double y5 = snaq_exp(-snaq_clamp10(cm - ca), M);
        const double y6 = snaq_exp(-snaq_clamp10(cb - cm), M);
        /* the reference's g*1/3*y = ((g*1)/3)*y: the quotient once per gamma (snaq_div3 = g/3.0 bit for bit) */
        const double g13 = snaq_div3(g1), g23 = snaq_div3(g2);
        cf[0] = g1 * (1.0 - 2.0 / 3.0 * y5) + g23 * y6;
        cf[1] = g13 * y5 + g2 * (1.0 - 2.0 / 3.0 * y6);
        cf[2] = g13 * y5 + g23 * y6;
    } else {
        const double z1 = x(pc[0]), z2 = x(pc[1]);
        cf[0] = snaq_div3(1.0 + 2.0 * z1 - z2);
        cf[1] = snaq_div3(1.0 + 2.0 * z2 - z1);
        cf[2] = snaq_div3(1.0 - z1 - z2);
    }
}

/*
 * Expected CFs of one quartet of any class, in PROGRAM order (tree-like: cf[0] major, cf[1] minor,
 * *t1 = the internal length; type 5: cf1, cf2, cf3 and *t1 = -1).  Generic path used by the
 * per-quartet read-back and by the test harness; the hot loops call the class functions directly.
 */
template <class XF>
SNAQ_HD void snaq_eval_quartet(unsigned hdr, const uint16_t *pc, int nwords, const XF &x, const SnaqMath &M, double cf[3],
                               double *t1out, unsigned &status) {
    const unsigned kind = SNAQ_HDR_KIND(hdr);
    if (kind == SNAQ_KIND_TREE) {
        const double t = SNAQ_HDR_TERMS(hdr) ? snaq_tree_terms_t1(pc, x, M, status) : snaq_tree_plain_t1(pc, nwords, x);
        snaq_tree_cf(t, M, cf);
        *t1out = t;
    } else {
        snaq_t5_cf(kind, pc, x, M, cf);
        *t1out = -1.0;
    }
}

/* program-order CFs -> data order (12|34, 13|24, 14|23) */
SNAQ_HD void snaq_data_order(unsigned hdr, const double cf[3], double out[3]) {
    const unsigned perm = SNAQ_HDR_PERM(hdr);
    if (SNAQ_HDR_KIND(hdr) == SNAQ_KIND_TREE) {
        out[0] = out[1] = out[2] = cf[1];
        out[perm] = cf[0];
    } else {
        int pm[3];
        snaq_perm_decode((int)perm, pm);
        out[0] = cf[pm[0] - 1]; out[1] = cf[pm[1] - 1]; out[2] = cf[pm[2] - 1];
    }
}

/*
 * Per-quartet weights in program order, built once per (topology, obsCF):
 *   W[0..2] = 100*obs of the CF that program slot j produces (tree-like: W[0] major,
 *             W[1] = sum of the two minors, W[2] = 0), W[3] = sum_i 100*obs_i*log(obs_i).
 * Unsampled quartets get all zeros (they contribute 0: src/pseudolik.jl:1390).
 */
SNAQ_HD void snaq_weights(unsigned hdr, const double obs[3], int sampled, double W[4]) {
    W[0] = W[1] = W[2] = W[3] = 0.0;
    if (!sampled) return;
    const unsigned perm = SNAQ_HDR_PERM(hdr);
    double c = 0.0;
    for (int i = 0; i < 3; i++)
        if (obs[i] != 0.0) c += 100.0 * obs[i] * log(obs[i]);
    W[3] = c;
    if (SNAQ_HDR_KIND(hdr) == SNAQ_KIND_TREE) {
        for (int i = 0; i < 3; i++) {
            if ((unsigned)i == perm) W[0] = 100.0 * obs[i];
            else W[1] += 100.0 * obs[i];
        }
    } else {
        int pm[3];
        snaq_perm_decode((int)perm, pm);
        for (int i = 0; i < 3; i++) W[pm[i] - 1] = 100.0 * obs[i];
    }
}

/*
 * logPseudoLik(quartet) (src/pseudolik.jl:1388-1410):
 *   sum_i [ expCF_i<0 ? -1e15 : (obs_i==0 ? 0 : 100 obs_i log(expCF_i/obs_i)) ]
 * written as sum_j W_j log(cf_j) - W[3]; for tree-like quartets log(minor) = -t1 - log 3.
 */
SNAQ_HD double snaq_loglik_tree(double t, double major, const double W[4], const SnaqMath &M) {
    return (W[0] != 0.0 ? W[0] * snaq_log(major, M) : 0.0) - W[1] * (t + SNAQ_K_LOG3) - W[3];
}
/* pen (deduplicated table only, else NULL): {E0, E1, E2, n} of the run this entry stands for, E_j = sum over its sampled
 * quartets of 100 obs_j log(obs_j), n = number of sampled quartets.  The -1e15 of a negative expCF is per QUARTET and
 * replaces that quartet's term including its obs*log(obs) share of W[3], so it is not linear in the weights. */
SNAQ_HD double snaq_loglik_t5(const double cf[3], const double W[4], const SnaqMath &M, unsigned &status, const double *pen = nullptr) {
    if (W[0] == 0.0 && W[1] == 0.0 && W[2] == 0.0) return 0.0;      /* unsampled: contributes 0, penalty included (src/pseudolik.jl:1390) */
    double s = -W[3];
    for (int j = 0; j < 3; j++) {
        if (cf[j] < 0.0) {
            /* -1e15 replaces this term, including its obs*log(obs) share of W[3] */
            if (pen) s += -1.e15 * pen[3] + pen[j];
            else s += -1.e15 + (W[j] != 0.0 ? W[j] * snaq_log_slow(W[j] / 100.0) : 0.0);
        } else if (W[j] != 0.0) {
            s += W[j] * snaq_log(cf[j], M);
        }
    }
    if (cf[0] + cf[1] + cf[2] == 0.0 && (W[0] != 0.0 || W[1] != 0.0 || W[2] != 0.0)) status |= SNAQ_S_ZEROSUM;
    return s;
}
SNAQ_HD double snaq_quartet_loglik(unsigned kind, const double cf[3], double t1, const double W[4], const SnaqMath &M,
                                   unsigned &status, const double *pen = nullptr) {
    if (kind == SNAQ_KIND_TREE) {
        if (cf[0] < 0.0) status |= SNAQ_S_NEGLEN;
        return snaq_loglik_tree(t1, cf[0], W, M);
    }
    return snaq_loglik_t5(cf, W, M, status, pen);
}

/* ------------------------------------------------------------------------------
 * objective + gradient (snaq_eval_grad).  For one quartet: returns its value v (what snaq_quartet_loglik
 * returns) and calls add(slot, dv/d xf[slot]) for every slot of the expanded vector its program reads.
 * The chain rule through the rows of xf that are not plain parameters (prefix sums, type-3 terms,
 * -log(1-gammaz)) is snaq_grad_backprop below.  At a 10.0 cap (setLength!: src/auxiliary.jl:122-124) the
 * derivative of the capped quantity is taken as 0 (the side on which the cap is active).
 * ------------------------------------------------------------------------------ */
template <class XF, class ADD>
SNAQ_HD void snaq_grad_pairs(const uint16_t *&pc, int n, const XF &x, unsigned nul, double gcoef, ADD &add) {
    SNAQ_UNROLL1
    for (int i = 0; i < n; i++) {
        const unsigned p = pc[2 * i], m = pc[2 * i + 1];
        if (p != nul) add(p, gcoef);
        if (m != nul) add(m, -gcoef);
    }
    pc += 2 * n;
}

/* d v / d t1 of a tree-like quartet: v = W0 log(1 - 2/3 e^-t) - W1 (t + log 3) - W3 */
SNAQ_HD double snaq_dv_dt_tree(double t1, const double W[4], const SnaqMath &M, double *major_out, double *curv_out = nullptr) {
    const double a = 2.0 / 3.0 * snaq_exp(-t1, M);
    const double major = 1.0 - a;
    *major_out = major;
    const double r = W[0] * a / major;
    if (curv_out) *curv_out = r / major;      /* -d2 v / d t1^2 = W0 a / (1-a)^2 >= 0 */
    return r - W[1];
}

/* addh(slot, c): curvature bins.  For an additive quartet t1 is a 0/1 combination of edge lengths, so the signed
 * sums of -d2v/dt1^2 over the program's slots transpose (like the gradient) into the exact contribution of these
 * quartets to the DIAGONAL of the Hessian of f with respect to the edge lengths; the other classes add nothing
 * (the diagonal is only used as the initial metric of the quasi-Newton update). */
template <class XF, class ADD, class ADDH>
SNAQ_HD double snaq_grad_quartet(unsigned hdr, const uint16_t *pc, int nwords, const XF &x, const SnaqMath &M, const double W[4],
                                 unsigned nul, ADD &add, ADDH &addh, unsigned &status, const double *pen = nullptr) {
    const unsigned kind = SNAQ_HDR_KIND(hdr);
    if (kind == SNAQ_KIND_TREE && !SNAQ_HDR_TERMS(hdr)) {
        /* class A */
        double t = 0.0;
        for (int i = 0; i < nwords; i += 2) t += x(pc[i]) - x(pc[i + 1]);
        const double t1 = snaq_clamp10(t);
        double major, curv;
        const double gt = snaq_dv_dt_tree(t1, W, M, &major, &curv);
        if (t <= 10.0) {
            const uint16_t *p2 = pc;
            snaq_grad_pairs(pc, nwords / 2, x, nul, gt, add);
            snaq_grad_pairs(p2, nwords / 2, x, nul, curv, addh);
        }
        return snaq_loglik_tree(t1, major, W, M);
    }
    if (kind == SNAQ_KIND_TREE) {
        /* class B: forward pass exactly as snaq_tree_terms_t1, remembering what the reverse pass needs */
        const uint16_t *p0 = pc;
        const double t1 = snaq_tree_terms_t1(pc, x, M, status);
        double major;
        const double gt1 = snaq_dv_dt_tree(t1, W, M, &major);
        if (major < 0.0) status |= SNAQ_S_NEGLEN;
        const double v = snaq_loglik_tree(t1, major, W, M);
        /* recompute the intermediate sums (cheap) to know which caps are active */
        pc = p0;
        const unsigned cw = *pc++;
        const int n0 = (int)(cw & 255u), nterms = (int)(cw >> 8);
        const uint16_t *ppairs = pc;
        double R = snaq_sum(pc, n0, x);
        const uint16_t *pterms = pc;
        double Bn = 0.0, Bf = 0.0;
        for (int it = 0; it < nterms; it++) {
            const unsigned th = *pc++;
            const double gq = x(th >> 4);
            if ((th & 3u) == SNAQ_T2) {
                const double e = snaq_clamp10(x(pc[0]) - x(pc[1]));
                pc += 2;
                const double gh = (th & 4u) ? 1.0 - gq : gq;
                R += snaq_clamp10(-snaq_log(1.0 - gh * (1.0 - snaq_exp(-e, M)), M));
            } else {
                const double clo = x(pc[0]), chi = x(pc[1]), ck = x(pc[2]);
                const double la = snaq_clamp10(clo), lb = snaq_clamp10(ck - chi), le = snaq_clamp10(chi - clo);
                const int nchain = pc[3];
                pc += 4;
                const double ch = snaq_clamp10(snaq_sum(pc, nchain, x));
                const double ga = 1.0 - gq, gb = gq;
                const double l4 = snaq_clamp10(-snaq_log(gb * gb * snaq_exp(-lb, M) + ga * gb * (3.0 - snaq_exp(-le, M)) + ga * ga * snaq_exp(-la, M), M));
                const double B = snaq_clamp10(ch + l4);
                if (th & 4u) Bf = B; else Bn = B;
            }
        }
        const double u = Bn + R, tt = snaq_clamp10(u) + Bf;
        const double gBf = (tt <= 10.0) ? gt1 : 0.0;
        const double gR = (u <= 10.0) ? gBf : 0.0;          /* = d v / d R = d v / d Bn */
        /* reverse pass */
        pc = ppairs;
        snaq_grad_pairs(pc, n0, x, nul, gR, add);
        pc = pterms;
        for (int it = 0; it < nterms; it++) {
            const unsigned th = *pc++;
            const unsigned gs = th >> 4;
            const double gq = x(gs);
            if ((th & 3u) == SNAQ_T2) {
                const unsigned s0 = pc[0], s1 = pc[1];
                pc += 2;
                const double er = x(s0) - x(s1), e = snaq_clamp10(er);
                const double gh = (th & 4u) ? 1.0 - gq : gq;
                const double ye = snaq_exp(-e, M);
                const double A = 1.0 - gh * (1.0 - ye);
                const double term_raw = -snaq_log(A, M);
                if (term_raw <= 10.0 && gR != 0.0) {
                    const double dgh = gR * (1.0 - ye) / A;
                    add(gs, (th & 4u) ? -dgh : dgh);
                    if (er <= 10.0) {
                        const double de = gR * gh * ye / A;
                        if (s0 != nul) add(s0, de);
                        if (s1 != nul) add(s1, -de);
                    }
                }
            } else {
                const unsigned slo = pc[0], shi = pc[1], sk = pc[2];
                const double clo = x(slo), chi = x(shi), ck = x(sk);
                const int nchain = pc[3];
                pc += 4;
                const double la = snaq_clamp10(clo), lb = snaq_clamp10(ck - chi), le = snaq_clamp10(chi - clo);
                const uint16_t *pchain = pc;
                const double ch_raw = snaq_sum(pc, nchain, x);
                const double ga = 1.0 - gq, gb = gq;
                const double yb = snaq_exp(-lb, M), ye = snaq_exp(-le, M), ya = snaq_exp(-la, M);
                const double S = gb * gb * yb + ga * gb * (3.0 - ye) + ga * ga * ya;
                const double l4_raw = -snaq_log(S, M);
                const double B_raw = snaq_clamp10(ch_raw) + snaq_clamp10(l4_raw);
                double gB = (th & 4u) ? gBf : gR;
                if (B_raw > 10.0) gB = 0.0;
                if (gB != 0.0) {
                    if (ch_raw <= 10.0) { const uint16_t *pp = pchain; snaq_grad_pairs(pp, nchain, x, nul, gB, add); }
                    if (l4_raw <= 10.0) {
                        const double gS = -gB / S;
                        const double dlb = gS * (-gb * gb * yb), dle = gS * (ga * gb * ye), dla = gS * (-ga * ga * ya);
                        const double dgb = 2.0 * gb * yb + ga * (3.0 - ye), dga = gb * (3.0 - ye) + 2.0 * ga * ya;
                        add(gs, gS * (dgb - dga));
                        double glo = 0.0, ghi = 0.0, gk = 0.0;
                        if (clo <= 10.0) glo += dla;
                        if (ck - chi <= 10.0) { gk += dlb; ghi -= dlb; }
                        if (chi - clo <= 10.0) { ghi += dle; glo -= dle; }
                        if (slo != nul) add(slo, glo);
                        if (shi != nul) add(shi, ghi);
                        if (sk != nul) add(sk, gk);
                    }
                }
            }
        }
        return v;
    }
    /* class C */
    double cf[3], dz[4] = {0.0, 0.0, 0.0, 0.0};
    snaq_t5_cf(kind, pc, x, M, cf);
    double dcf[3];
    /* a negative expCF costs a flat -1e15 (src/pseudolik.jl:1394-1396): the objective has no slope there, so the
     * reported "gradient" of such a term is the slope of SNAQ_PENALTY_SLOPE * cf, which points out of the region */
    const bool smp = (W[0] != 0.0 || W[1] != 0.0 || W[2] != 0.0);     /* an unsampled quartet contributes nothing, slope included */
    for (int j = 0; j < 3; j++)
        dcf[j] = !smp ? 0.0 : (cf[j] < 0.0 ? SNAQ_PENALTY_SLOPE * (pen ? pen[3] : 1.0) : ((W[j] != 0.0 && cf[j] > 0.0) ? W[j] / cf[j] : 0.0));
    if (kind == SNAQ_KIND_T5) {
        const double g2 = x(pc[0]), g1 = 1.0 - g2;
        const double a5 = x(pc[2]) - x(pc[1]), a6 = x(pc[3]) - x(pc[2]);
        const double y5 = snaq_exp(-snaq_clamp10(a5), M), y6 = snaq_exp(-snaq_clamp10(a6), M);
        const double dy5 = dcf[0] * (-2.0 / 3.0 * g1) + dcf[1] * (g1 / 3.0) + dcf[2] * (g1 / 3.0);
        const double dy6 = dcf[0] * (g2 / 3.0) + dcf[1] * (-2.0 / 3.0 * g2) + dcf[2] * (g2 / 3.0);
        const double dg2 = dcf[0] * (-(1.0 - 2.0 / 3.0 * y5) + y6 / 3.0) + dcf[1] * (-y5 / 3.0 + (1.0 - 2.0 / 3.0 * y6)) +
                           dcf[2] * (-y5 / 3.0 + y6 / 3.0);
        const double da5 = (a5 <= 10.0) ? -y5 * dy5 : 0.0, da6 = (a6 <= 10.0) ? -y6 * dy6 : 0.0;
        dz[0] = dg2; dz[1] = -da5; dz[2] = da5 - da6; dz[3] = da6;
        for (int j = 0; j < 4; j++) if (pc[j] != nul && dz[j] != 0.0) add(pc[j], dz[j]);
    } else {
        dz[0] = (2.0 * dcf[0] - dcf[1] - dcf[2]) / 3.0;
        dz[1] = (-dcf[0] + 2.0 * dcf[1] - dcf[2]) / 3.0;
        add(pc[0], dz[0]);
        add(pc[1], dz[1]);
    }
    return snaq_loglik_t5(cf, W, M, status, pen);
}

/*
 * Chain rule through the expansion x -> xf (snaq_expand_row): gx[i] += sum_rows gxf[row] * d xf[row] / d x[i].
 * gx must hold nparams zeros on entry.  Runs on the host (the table pointers of T are host pointers).
 */
SNAQ_HD void snaq_grad_backprop(const SnaqTables &T, const SnaqMath &M, const double *x, const double *xconst, const double *gxf,
                                double *gx, bool linear_rows_only = false) {
    for (int row = 0; row < T.nparams; row++) gx[row] += gxf[row];
    for (int k = 0; k < T.n_int; k++) {
        const double gr = gxf[T.n_xe + 1 + k];
        if (gr == 0.0) continue;
        for (uint32_t i = T.dpath_off[k]; i < T.dpath_off[k + 1]; i++) { const int sl = T.dpath_slot[i]; if (sl < T.nparams) gx[sl] += gr; }
    }
    for (int c = 0; c < T.n_cycles; c++) {
        const int kc = T.cyc_k[c];
        const uint16_t *es = T.cyc_slot + T.cyc_off[c];
        const int r0 = (int)T.crow[c];
        for (int j = 1; j <= kc; j++) {                         /* C_j */
            const double gr = gxf[r0 + j];
            if (gr == 0.0) continue;
            for (int i = 0; i < j; i++) if (es[i] < T.nparams) gx[es[i]] += gr;
        }
        if (linear_rows_only) continue;
        if (!(T.cyc_flags[c] & 1)) {
            for (int o = 1; o < kc; o++) {                      /* type-3 term rows */
                const double gr = gxf[r0 + kc + 1 + (o - 1)];
                if (gr == 0.0) continue;
                double l1 = 0.0, l2 = 0.0;
                for (int i = 0; i < o; i++) l1 += snaq_xe(T, x, xconst, es[i]);
                for (int i = o; i < kc; i++) l2 += snaq_xe(T, x, xconst, es[i]);
                const double l1c = snaq_clamp10(l1), l2c = snaq_clamp10(l2);
                const double g2 = x[T.cyc_gslot[c]], g1 = 1.0 - g2;
                const double y1 = snaq_exp(-l1c, M), y2 = snaq_exp(-l2c, M);
                const double A = 1.0 - g1 * g1 * (1.0 - y1) - g2 * g2 * (1.0 - y2);
                if (-snaq_log(A, M) > 10.0) continue;
                const double s = gr / A;                        /* d row = -dA / A */
                gx[T.cyc_gslot[c]] += s * (-(2.0 * g1 * (1.0 - y1) - 2.0 * g2 * (1.0 - y2)));
                if (l1 <= 10.0) for (int i = 0; i < o; i++) if (es[i] < T.nparams) gx[es[i]] += s * g1 * g1 * y1;
                if (l2 <= 10.0) for (int i = o; i < kc; i++) if (es[i] < T.nparams) gx[es[i]] += s * g2 * g2 * y2;
            }
        } else {
            for (int j = 0; j < 2; j++) {                       /* -log(1-gammaz) rows */
                const double gr = gxf[r0 + 2 * kc + j];
                if (gr == 0.0) continue;
                const double z = x[T.cyc_gslot[c] + j];
                if (-snaq_log(1.0 - z, M) > 10.0) continue;
                gx[T.cyc_gslot[c] + j] += gr / (1.0 - z);
            }
        }
    }
}

#endif /* SNAQ_CORE_H */
