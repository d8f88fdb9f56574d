/*
 * snaq_tables.h — host side of the descriptor build: turns a flat network
 * (include/snaq_b200.h) into the small structural tables of snaq_core.h and the
 * parameter layout of parameters!(net) (src/snaq_optimization.jl:54-101).
 * Runs once per topology on the host (O(E*h)); everything per quartet happens on
 * the device.
 */
#ifndef SNAQ_TABLES_H
#define SNAQ_TABLES_H

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/snaq_b200.h"
#include "snaq_core.h"
#include "snaq_hesim.h"

struct SnaqHostTables {
    int ntaxa = 0, n_tnodes = 0, n_cycles = 0, nparams = 0, n_xe = 0, n_hybrids = 0;
    int n_int = 0, n_full = 0;                        // internal blob-tree nodes; rows of the expanded vector
    int nh = 0, nt = 0, nhz = 0;                     // x = [gamma | t | gammaz]
    std::vector<double> x0, upper, xconst;           // xconst: values of slots nparams..n_xe-1
    std::vector<int32_t> numht;
    // one contiguous blob holding every table (uploaded to the device as is)
    std::vector<unsigned char> blob;
    size_t off_leaf_tnode = 0, off_parent = 0, off_depth = 0, off_tcycle = 0, off_pslot = 0, off_blk = 0,
           off_cyc_k = 0, off_cyc_off = 0, off_cyc_slot = 0, off_cyc_gslot = 0, off_cyc_flags = 0, off_cyc_hyb = 0, off_torder = 0, off_cyc_order = 0, off_dslot = 0, off_dnode = 0, off_crow = 0, off_dpath_off = 0, off_dpath_slot = 0, off_blkmax = 0, off_hmax2 = 0, off_hmaxtax = 0;
    // extra tables used only by the hasEdge (parity descriptor) kernel
    std::vector<int32_t> edge_slot;                  // per net edge: xe slot or -1
    int root = -1;                                   // blob-tree node the tree is rooted at
    SnaqTables view(const unsigned char *base) const;
    // the network as the hasEdge simulation starts from it (snaq_hesim.h): its own blob, uploaded on demand
    std::vector<unsigned char> sim_blob;
    SnaqSimNet sim{};                                // counts filled in; pointers are set by sim_view()
    size_t so_node_edge = 0, so_node_flags = 0, so_node_incycle = 0, so_node_taxon = 0, so_edge_node = 0, so_edge_flags = 0,
           so_leaf_order = 0, so_hyb_order = 0, so_slot_ref = 0, so_slot_alias = 0, so_edge_number = 0, so_node_number = 0;
    SnaqSimNet sim_view(const unsigned char *base) const;
};

// returns SNAQ_OK or an error code with `err` set.  keep_root >= 0: root the blob tree at that blob-tree node (if it
// exists) instead of the middle of a diameter, so that two successive topologies share root paths (snaq_patch_network).
int snaq_build_tables(const snaq_flatnet *f, int ntaxa, SnaqHostTables &out, std::string &err, int keep_root = -1);

// Which quartets can keep their program when the tables change from `a` to `b`?  A program is a function of the root
// paths of its four taxa (blob-tree nodes, depths, edge slots), of their block positions in every cycle, and of the
// tables that are not per taxon.  Returns false when the latter differ (every quartet must be re-described); else
// touched[t] = 1 for every taxon whose root path or block positions changed.
bool snaq_tables_touched(const SnaqHostTables &a, const SnaqHostTables &b, std::vector<uint8_t> &touched);

#endif
