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
    SnaqTables view(const unsigned char *base) const;
};

// returns SNAQ_OK or an error code with `err` set
int snaq_build_tables(const snaq_flatnet *f, int ntaxa, SnaqHostTables &out, std::string &err);

#endif
