/*
 * snaq_b200.h — C ABI of the B200-native quartet-pseudolikelihood engine.
 *
 * Drop-in boundary for the hot path of JuliaPhylo/SNaQ.jl (reference paths are
 * relative to the reference repository):
 *
 *   extractQuartet!(net,d)              src/pseudolik.jl:503-540
 *   calculateExpCFAll!(d,x,net)         src/pseudolik.jl:1310-1324
 *   logPseudoLik(d)                     src/pseudolik.jl:1388-1425
 *   parameters!(net) / upper(net)       src/snaq_optimization.jl:54-101,275-279
 *   objective closure of optBL!         src/snaq_optimization.jl:368-380
 *   topologyQpseudolik! (score only)    src/pseudolik.jl:1342-1378
 *
 * The reference has no FFI for this path (it is pure Julia); the entry points
 * below are what a Julia `ccall` shim binds.  Plain C
 * types only: the caller owns every host buffer, the engine copies during the
 * call and never retains a host pointer.  Every function returns 0 on success
 * and a nonzero SNAQ_E* code on failure; snaq_last_error() gives the message.
 * There is no CPU fallback: without a CUDA device snaq_create() fails.
 */
#ifndef SNAQ_B200_H
#define SNAQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes -------------------------------------------------------- */
#define SNAQ_OK            0
#define SNAQ_EINVAL        1   /* bad argument / inconsistent flat network     */
#define SNAQ_ECUDA         2   /* CUDA runtime error (message has the detail)  */
#define SNAQ_ENODATA       3   /* set_data / set_network not called yet        */
#define SNAQ_ERANGE        4   /* parameter outside its bounds: gamma or gammaz
                                  outside [0,1] (src/snaq_optimization.jl:213,
                                  221,228) or a length < 0 (src/auxiliary.jl:119) */
#define SNAQ_ENUMERIC      5   /* sum(expCF)==0 (src/pseudolik.jl:1391), type-5
                                  expCF not summing to 1 (:1049), or merged
                                  length < -log(1.5) (src/auxiliary.jl:120)    */
#define SNAQ_ETOPOLOGY     6   /* network is not level-1 / not what the flat
                                  format promises                              */
#define SNAQ_ENOMEM        7

/* ---- node / edge flag bits ---------------------------------------------- */
/* Node flags mirror the PhyloNetworks slots the hot path reads through the
 * accessors of src/types.jl:15-24.                                            */
#define SNAQ_NODE_LEAF          0x01u
#define SNAQ_NODE_HYBRID        0x02u
#define SNAQ_NODE_HASHYBEDGE    0x04u  /* hasHybEdge     src/types.jl:15 */
#define SNAQ_NODE_BADDIAMOND1   0x08u  /* isBadDiamondI  src/types.jl:16 */
#define SNAQ_NODE_BADDIAMOND2   0x10u  /* isBadDiamondII src/types.jl:17 */
#define SNAQ_NODE_BADTRIANGLE   0x20u  /* isBadTriangle  src/types.jl:20 */
/* Edge flags: src/types.jl:5-7 plus the PN Edge fields hybrid/ismajor/ischild1 */
#define SNAQ_EDGE_HYBRID        0x01u
#define SNAQ_EDGE_MAJOR         0x02u
#define SNAQ_EDGE_ISCHILD1      0x04u
#define SNAQ_EDGE_IDENTIFIABLE  0x08u  /* istIdentifiable  src/types.jl:5 */
#define SNAQ_EDGE_FROMBD1       0x10u  /* fromBadDiamondI  src/types.jl:6 */

/*
 * Flattened HybridNetwork: everything the hot path reads from the reference's
 * object graph.  Array order IS information: node[]/edge[] are
 * in net.node / net.edge order (the x layout follows net.edge order,
 * src/snaq_optimization.jl:64-91), node_edge keeps each node.edge order,
 * hybrid[] is net.hybrid order and leaf[] is net.leaf order (leaves are pruned
 * in that order, src/pseudolik.jl:476-483).
 */
typedef struct snaq_flatnet {
    int32_t n_nodes;
    int32_t n_edges;
    int32_t n_hybrids;
    int32_t n_leaves;
    /* nodes */
    const int32_t *node_number;   /* [n_nodes] PN node.number                        */
    const uint32_t *node_flags;   /* [n_nodes] SNAQ_NODE_* bits                       */
    const int32_t *node_incycle;  /* [n_nodes] hybrid node NUMBER of its cycle, or -1 */
    const double  *node_gammaz;   /* [n_nodes] gammaz, -1.0 when unset                */
    const int32_t *node_taxon;    /* [n_nodes] taxon id (0-based index into the data's
                                     taxon list) for leaves, -1 otherwise             */
    const int32_t *node_edge;     /* [n_nodes*3] edge INDICES in node.edge order,
                                     -1 padded                                         */
    /* edges */
    const int32_t *edge_number;   /* [n_edges] PN edge.number                         */
    const uint32_t *edge_flags;   /* [n_edges] SNAQ_EDGE_* bits                       */
    const int32_t *edge_incycle;  /* [n_edges] hybrid node NUMBER, or -1              */
    const double  *edge_length;   /* [n_edges]                                        */
    const double  *edge_gamma;    /* [n_edges] 1.0 for tree edges                     */
    const int32_t *edge_node;     /* [n_edges*2] node INDICES in edge.node order      */
    /* orders */
    const int32_t *hybrid;        /* [n_hybrids] node indices, net.hybrid order       */
    const int32_t *leaf;          /* [n_leaves]  node indices, net.leaf order         */
} snaq_flatnet;

typedef struct snaq_ctx snaq_ctx;   /* opaque: one ctx <=> one CUDA device + stream */

/* Run table (default).  The expected CFs of a quartet depend only on its split and on the pieces of the network between its
 * two junctions, i.e. on its evaluation PROGRAM, and the pseudo-deviance is linear in the weights 100*obsCF: quartets whose
 * programs are identical (they differ in taxa of the same clades) are evaluated ONCE per parameter vector with the SUM of
 * their weights.  The sums are formed when the observed CFs change (snaq_set_data / _set_obscf / _set_sampled /
 * _resample_obscf), per run of identical programs of the class-sorted table, with compensated arithmetic
 * (27,405 quartets -> 1,290 programs at 30 taxa; 3,921,225 -> 9,430 at 100 taxa).  Every (quartet, vector) pair is in the
 * result exactly as before, to rounding (<= 2e-15 relative: the order of the sum changes); the -1e15 penalty of a negative
 * expected CF stays per sampled quartet; per-quartet read-backs (snaq_eval_quartets) use the full table.
 * SNAQ_OPT_PER_QUARTET (env SNAQ_DEDUP=0) selects the per-quartet kernels instead: one weighted contribution computed per
 * (quartet, vector) -- k_eval_runs for batches (the program still evaluated once per run, the weights streamed per
 * quartet), k_eval_lanes for tables with short runs, k_eval_direct / k_eval_grad streaming the per-quartet table for
 * single-vector calls (objective, gradient, optBL!; SNAQ_SV_DEDUP=1 puts only those back on the run table).
 * SNAQ_OPT_DEDUP is kept for source compatibility (it is the default).                                                 */
#define SNAQ_OPT_DEDUP 1u
#define SNAQ_OPT_PER_QUARTET 2u

typedef struct snaq_opts {
    int32_t device;        /* CUDA device ordinal                                  */
    int32_t rank;          /* quartet-sharded mode: this rank (0 if unsharded)     */
    int32_t world_size;    /* quartet-sharded mode: number of ranks (1 if not)     */
    uint32_t flags;        /* SNAQ_OPT_*                                           */
} snaq_opts;

/* ---- life cycle ---------------------------------------------------------- */
int snaq_create(const snaq_opts *opts, snaq_ctx **out);
int snaq_destroy(snaq_ctx *ctx);
const char *snaq_last_error(const snaq_ctx *ctx);   /* ctx may be NULL: last create error */
const char *snaq_version(void);

/* ---- quartet-sharded mode: the exchange step -------------------- */
/* A context created with world_size > 1 keeps the contiguous block [rank*nq/G, (rank+1)*nq/G) of the quartets passed to
 * snaq_set_data; every rank evaluates the SAME parameter vectors and the per-candidate partial sums of logPseudoLik
 * (src/pseudolik.jl:1417-1423 is a plain sum over quartets) are combined over NVLink / NVSwitch.  The collective lives
 * in the library (NCCL is resolved at run time with dlopen; a caller does not bind it):
 *   snaq_comm_unique_id : rank 0 obtains the 128-byte id (ncclGetUniqueId) and hands it to the other ranks by whatever
 *                         means the host program has (Julia: Distributed / a file / MPI);
 *   snaq_comm_init      : every rank, same id: ncclCommInitRank on the context's device, plus the peer mailboxes of the
 *                         fused epilogue (CUDA IPC).  Collective: all ranks must call it.
 * Afterwards snaq_eval, snaq_eval_device, snaq_eval_grad and snaq_optbl return the TOTALS over all ranks, bit-identical
 * on every rank, and every rank must make the same sequence of these calls with the same parameter vectors (they are
 * collectives).  Small batches (P < 16, the optimiser's objective calls) use ONE kernel launch per <= 4 candidates: the
 * last block of every rank stores its partial sums into every peer's mailbox through peer memory and adds the entries
 * in rank order once all have arrived (no second kernel, no NCCL call); larger batches end with an ncclAllReduce of P
 * doubles on the caller's stream.  Without a communicator snaq_eval / snaq_eval_device return this rank's partial
 * sums (the caller all-reduces them), and snaq_eval_grad / snaq_optbl are refused (SNAQ_EINVAL).
 * snaq_comm_info: any out pointer may be NULL.                                                                   */
#define SNAQ_COMM_ID_BYTES 128
int snaq_comm_unique_id(void *id128);
int snaq_comm_init(snaq_ctx *ctx, const void *id128);
int snaq_comm_info(snaq_ctx *ctx, int32_t *rank, int32_t *world_size, int32_t *has_comm, int32_t *fused_epilogue);

/* ---- data: DataCF resident in HBM (src/types.jl:217-273) ----------------- */
/* taxa[q*4+i] = 0-based taxon id of the i-th taxon of quartet q, in the data's
 * own order (which fixes the 12|34, 13|24, 14|23 order of obsCF).             */
int snaq_set_data(snaq_ctx *ctx, int32_t ntaxa, int64_t nq,
                  const uint16_t *taxa, const double *obscf);
/* bootstrap refresh: overwrite obsCF in place (src/readquartetdata.jl:211-217) */
int snaq_set_obscf(snaq_ctx *ctx, const double *obscf);
/* Bootstrap tables on the device: sampleCFfromCI!, src/bootstrap.jl:52-70, followed by
 * readtableCF!, src/readquartetdata.jl:211-217.  snaq_set_ci keeps the credibility intervals lo/hi [nq][3] of the
 * three CFs resident (0 <= lo <= hi <= 1, some hi > 0 in every quartet); snaq_resample_obscf then overwrites obsCF:
 *     c_i = (hi_i - lo_i) * U_i + lo_i,   obsCF_i = c_i / (c_1 + c_2 + c_3)          (:60-67)
 * The reference draws U from Julia's task-local RNG after Random.seed!(seed), which cannot be reproduced outside Julia;
 * here U_i = (splitmix64(seed + (3*q + i) * 0x9E3779B97F4A7C15) >> 11) * 2^-53 with q the quartet's index in the data
 * (the same table on every rank of a quartet-sharded run, reproducible on the host: tests/test_gpu_bootstrap.py).
 * snaq_get_obscf reads the current table back.                                                                      */
int snaq_set_ci(snaq_ctx *ctx, const double *lo, const double *hi);
int snaq_resample_obscf(snaq_ctx *ctx, uint64_t seed);
int snaq_get_obscf(snaq_ctx *ctx, double *obscf);
/* q.sampled mask (src/update.jl:386,406,442); NULL = all sampled             */
int snaq_set_sampled(snaq_ctx *ctx, const uint8_t *mask);

/* ---- topology: extractQuartet!(net,d) ------------------------------------ */
/* Full rebuild of the device descriptor table from a flat network.  Also does
 * parameters!(net): fixes the x layout.                                       */
int snaq_set_network(snaq_ctx *ctx, const snaq_flatnet *net);
/* Incremental variant for the search loop, which edits ONE network in place and re-extracts every quartet at each optBL!
 * (proposedTop!, src/snaq_optimization.jl:1154-1198; extractQuartet! at :354).  The new host tables are rooted where
 * the current ones are; a quartet's program is a function of the root paths and cycle blocks of its four taxa, so the
 * engine compares the two sets of tables itself and classifies again only the quartets that contain a taxon whose root
 * path or block changed (after a nearest-neighbour interchange: the taxa of the two subtrees that changed places).
 * Everything downstream of the classification (class sort, offsets, emission from the staged programs, run table,
 * weights) is redone for the whole table, so the device table is BIT-IDENTICAL to what a build from scratch with the
 * same root gives (tests/test_gpu_patch.py); only the expensive pass (two thirds of a rebuild) shrinks.
 * touched_taxa / n_touched: taxon ids the caller knows to be affected (advisory: added to the engine's own list; may be
 * NULL / 0).  n_touched < 0: re-describe everything (a rebuild that keeps the current root).  Also re-described in
 * full: edits that change the parameter layout or a cycle (hybrid added or removed, bad-diamond status changed), or that
 * touch more than half of the taxa.  snaq_patch_info: taxa re-described by the last call (-1: everything).          */
int snaq_patch_network(snaq_ctx *ctx, const snaq_flatnet *net,
                       const int32_t *touched_taxa, int32_t n_touched);
int snaq_patch_info(snaq_ctx *ctx, int64_t *touched_taxa);
/* order-dependent checksums of the device table: [0] sort permutation, headers, offsets; [1] program words; [2] weights;
 * [3] (runs << 32) | tiles of the run table.  Test aid: a patched table against a rebuilt one.                        */
int snaq_table_checksum(snaq_ctx *ctx, uint64_t out[4]);

/* topologyQpseudolik! (src/pseudolik.jl:1342-1378) for a list of candidate topologies on the resident data:
 * out[c] = pseudo-deviance of nets[c] at its own parameters (each is a descriptor rebuild + one objective call).
 * The context is left holding the LAST network.                                                              */
int snaq_eval_topologies(snaq_ctx *ctx, int32_t C, const snaq_flatnet *nets, double *out);

/* parameters!(net): src/snaq_optimization.jl:54-101.  x = [gamma | t | gammaz].
 * numht: edge numbers (gamma, t) and "<node><1|2>" pseudo numbers (gammaz).
 * Any out pointer may be NULL.                                                */
int snaq_num_params(snaq_ctx *ctx, int32_t *nparams, int32_t *n_gamma,
                    int32_t *n_t, int32_t *n_gammaz);
int snaq_get_params(snaq_ctx *ctx, double *x, int32_t *numht,
                    double *upper /* upper(net), src/snaq_optimization.jl:275 */);

/* ---- the hot path --------------------------------------------------------- */
/* out[p] = logPseudoLik(d) after calculateExpCFAll!(d, X[p,:], net), for P
 * parameter vectors (row-major P x nparams).  Stateless evaluation: every
 * sampled quartet is recomputed (the reference's `changed` gating,
 * src/snaq_optimization.jl:202-206, only skips work, never changes a value).
 * In quartet-sharded mode out[p] is the total over all ranks once snaq_comm_init
 * has been called, else this rank's partial sum.                               */
int snaq_eval(snaq_ctx *ctx, int32_t P, int32_t nparams,
              const double *X, double *out);
/* Same, X and out already in device memory (used by the in-engine optimiser,
 * the benchmark and the NCCL-sharded path).  stream is a cudaStream_t.         */
int snaq_eval_device(snaq_ctx *ctx, int32_t P, int32_t nparams,
                     const double *dX, double *dout, void *stream);

/* snaq_eval_device does not synchronise; this reads (and clears) the status word its kernels
 * OR error bits into, and maps it to SNAQ_ERANGE / SNAQ_ENUMERIC like snaq_eval does.          */
int snaq_check_status(snaq_ctx *ctx);

/* Per-quartet results for one parameter vector x (host buffers):
 * expcf[nq*3] (q.qnet.expCF), logpl[nq] (q.logPseudoLik, src/pseudolik.jl:1407),
 * deltacf[nq] (calculateDeltaCF, src/pseudolik.jl:491-498).  Any may be NULL.  */
int snaq_eval_quartets(snaq_ctx *ctx, int32_t nparams, const double *x,
                       double *expcf, double *logpl, double *deltacf);

/* ---- optBL!: src/snaq_optimization.jl:341-390 --------------------------- */
/* Objective and its gradient in ONE pass over the quartets: *f = logPseudoLik at x, grad[i] = d f / d x[i]
 * (analytic, through the closed forms of src/pseudolik.jl:878,938,951,985-991,1259; one-sided at the
 * 10.0 caps of setLength!, where the reference's objective has a kink; inside the flat -1e15 plateau of a
 * negative expCF, src/pseudolik.jl:1394-1396, the slope of 1e6*expCF so that a descent method can leave it).
 * hdiag[i] (may be NULL; costs extra work) = the contribution of the tree-like quartets without hybrid terms
 * to d2 f / d x[i]^2 for edge lengths, 0 for gammas: the initial metric of the quasi-Newton update.
 * f is the per-quartet sum of snaq_eval (same value to rounding).  The derivative sums are formed in 2^-32 fixed point
 * (integer sums: bit-reproducible whatever the order of arrival), every quartet's contribution rounded once: an
 * absolute error of about 1e-10 * sqrt(number of quartets that touch the slot) per entry (a few 1e-7 at 5e5 quartets).
 * The bins are 96 bits wide (range +-2^63): no table size and no parameter vector inside the bounds can wrap them.
 * Quartet-sharded contexts: with a communicator attached (snaq_comm_init) f and grad are the totals over all ranks
 * (identical on every rank); without one the call is refused (SNAQ_EINVAL), as is snaq_optbl.                      */
int snaq_eval_grad(snaq_ctx *ctx, int32_t nparams, const double *x, double *f,
                   double *grad, double *hdiag);

#define SNAQ_OPTBL_FD       1u   /* flags: finite-difference gradients (2n vectors per launch) instead of snaq_eval_grad */
/* result status, NLopt's numbering (the reference only prints it: src/snaq_optimization.jl:384-386) */
#define SNAQ_OPTBL_FTOL     3
#define SNAQ_OPTBL_XTOL     4
#define SNAQ_OPTBL_MAXEVAL  5
typedef struct snaq_optbl_opts {
    double ftol_rel, ftol_abs, xtol_rel, xtol_abs;   /* optBL!'s ftolRel, ftolAbs, xtolRel, xtolAbs (all > 0, :350) */
    int32_t maxeval;     /* NLopt.maxeval! = 1000 in the reference (:365) bounds BOBYQA's sequential objective calls;
                            here it bounds the engine LAUNCHES (a gradient pass or one ladder of vectors evaluated
                            together is one sequential step); result.nevals counts the vectors evaluated          */
    int32_t ladder;      /* step lengths tried per line-search launch (default 20)                              */
    uint32_t flags;      /* SNAQ_OPTBL_*                                                                         */
    uint32_t reserved;
} snaq_optbl_opts;
typedef struct snaq_optbl_result {
    double fmin;         /* loglik!(net, fmin): :388                                                            */
    double proj_grad;    /* max-norm of the projected gradient at xmin                                          */
    int32_t nevals;      /* parameter vectors evaluated                                                         */
    int32_t nlaunches;   /* engine launches (batches)                                                           */
    int32_t niter;
    int32_t status;      /* SNAQ_OPTBL_FTOL / XTOL / MAXEVAL                                                    */
} snaq_optbl_result;
/* Optimises every parameter of the current network within [0, upper(net)].  x: in = ht(net) (the start, as
 * NLopt.optimize(opt, ht(net)) :383), out = xmin (what updateParameters!(net, xmin) :387 receives).
 * opts may be NULL (the search-time tolerances fRel, fAbs, xRel, xAbs of :36-39).                          */
int snaq_optbl(snaq_ctx *ctx, const snaq_optbl_opts *opts, double *x, snaq_optbl_result *res);

/* Candidate-move batching: optBL! of C candidate networks over the resident data, run concurrently
 * on this context's GPU.  It replaces a loop of `optBL!(proposed net, d)` calls, src/snaq_optimization.jl:1362-1432
 * (one per proposed topology), for callers that can propose several candidates before looking at a result.
 * `workers` (1..32, 0 = 8) worker contexts are kept inside ctx: each has its own stream and table and a device copy of
 * ctx's data (taxa, obsCF, sampling mask), refreshed when ctx's data changes.  Candidate c starts at the parameters its
 * flat network carries (ht(net)); xmin + c*ldx receives its optimum, nparams[c] its parameter count (ldx >= every
 * count), results[c] the summary.  Each result is bit-identical to snaq_set_network + snaq_optbl on one context, whatever
 * runs beside it.  A candidate that fails gets results[c].status = -(error code); the call then returns the first such
 * code with "topology <c>: ..." as message, the other candidates' results stay valid.  ctx's own network is untouched.
 * Not available in quartet-sharded mode.                                                                              */
int snaq_optbl_topologies(snaq_ctx *ctx, int32_t C, const snaq_flatnet *nets, const snaq_optbl_opts *opts,
                          int32_t workers, double *xmin, int32_t ldx, int32_t *nparams, snaq_optbl_result *results);

/* ---- updateBL!: src/readquartetdata.jl:838-861 -------------------------- */
/* The current network must be a tree.  For each edge-length parameter (the t block of x, numht order):
 * nquartets[i] = data quartets that correspond exactly to the edge (one taxon in each of the four subtrees
 * around it: edgesParts :866-889, makeTable :935-960) and edgeL[i] = -log(3/2 (1 - mean obsCF of the tree's
 * resolution over them)) (:846; 0 when nquartets[i]==0).  Not available on a quartet-sharded context (a mean
 * over one rank's quartets is not the mean).  The caller applies the rule of :850-858
 * (only lengths < 0 or == 1.0 are replaced, by max(edgeL,0); edges without quartets get 1.0).               */
int snaq_update_bl(snaq_ctx *ctx, int64_t *nquartets, double *edgeL);

/* ---- weighted quartet sampling of the move proposals ---- */
/* Step 1 of sampleEdgeQuartetWeighted (src/moves_snaq.jl:1443-1449): a weighted sample of quartets with weights
 * q.deltaCF = sum_i |obsCF_i - expCF_i| at the parameters x (0 for unsampled quartets, src/pseudolik.jl:503-517).
 * The reference copies all nq weights to a vector per proposal and calls StatsBase.sample(indices, Weights(w), 1):
 * for one draw that is the inverse-CDF walk `t = rand()*sum(w); i = first index with cumsum(w)[i] >= t`.  Here the
 * weights never leave the device: idx[k] (0-based, the data's quartet order) is that index for the caller's uniform
 * draw u[k] in [0,1).  The prefix sums are formed in fixed chunks of 4096 quartets (chunk totals added serially, a
 * chunk walked 32 quartets at a time in a fixed order), so the result is reproducible and equals the serial walk
 * unless u*sum falls within rounding distance of a boundary.  Not available on a quartet-sharded context.            */
int snaq_sample_quartets(snaq_ctx *ctx, int32_t nparams, const double *x, int32_t n, const double *u, int64_t *idx);

/* ---- descriptors for parity checks ------------------------ */
/* which[nq]       : 1 tree-like, 2 four-cycle (src/types.jl:180)
 * ntype[nq], types[nq*max_types]: typeHyb codes (2..5; type 1 dropped) of the
 *                   hybrids that survive, in net.hybrid order
 * formula[nq*3]   : which==1: 1 major / 2 minor (src/pseudolik.jl:1230-1247);
 *                   which==2: 11,12,13 = which of cf1,cf2,cf3 lands there
 *                   (src/pseudolik.jl:1008-1048)
 * hasedge[nq*nparams] : 0/1, aligned with x (src/pseudolik.jl:392-464); see snaq_get_hasedge
 * Any may be NULL.                                                             */
int snaq_get_descriptors(snaq_ctx *ctx, int8_t *which, int8_t *ntype,
                         int8_t *types, int32_t max_types, int8_t *formula,
                         uint8_t *hasedge);

/* q.qnet.hasEdge and q.qnet.indexht (updateHasEdge!, src/pseudolik.jl:392-464; parameters!(qnet,net),
 * src/snaq_optimization.jl:106-180) for any level-1 network: hasedge[nq*nparams] 0/1 aligned with x; indexht[nq*nparams]
 * the 1-based positions of x the quartet uses in the reference's order, -1 padded, nindexht[nq] their number (either may
 * be NULL).  The reference defines hasEdge operationally ("the edge number still exists in the pruned quartet network").
 * With at most one hybrid node that is a structural property of the quartet; with nested cycles it depends on the order
 * in which extractQuartet! deletes the other leaves (net.leaf order, which the flat network carries), and the engine
 * simulates that pruning per quartet (integer state only, csrc/snaq_hesim.h).  Bit-exact against the oracle for h = 0..5.
 * These arrays only gate recomputation in the reference (src/snaq_optimization.jl:185-206); they never change a value. */
int snaq_get_hasedge(snaq_ctx *ctx, uint8_t *hasedge, int32_t *indexht, int32_t *nindexht);

/* Step 2 of sampleEdgeQuartetWeighted (src/moves_snaq.jl:1451-1463) reads `q.qnet.edge` of the ONE quartet it has drawn:
 * the engine keeps no per-quartet network objects, so this returns the edge numbers of that pruned quartet network, in
 * qnet.edge order (an edge renamed by the bad-diamond-I rewrite carries "<hybrid node number><1|2>", as in the
 * reference).  q: index in the data's quartet order (snaq_sample_quartets returns such indices).  *n = their number; at
 * most cap are written.                                                                                              */
int snaq_get_quartet_edges(snaq_ctx *ctx, int64_t q, int32_t *edge_numbers, int32_t cap, int32_t *n);

/* Bounds-checked build of the library (csrc/Makefile: libsnaqb200_dbg.so, -DSNAQ_BOUNDS): every slot index an evaluation
 * kernel derives from a program word is checked against the expanded parameter vector, every tile record / run / chunk index
 * of the run-length kernel against its stage; *checks and *violations are the process-wide device counters.  The normal
 * build returns -1, -1.  (compute-sanitizer is not available on the GPU pool; profiles/r02_bounds_check.md)            */
int snaq_debug_bounds(snaq_ctx *ctx, int64_t *checks, int64_t *violations);

/* engine counters: [0] launches of the eval kernel, [1] quartet-evals done,
 * [2] H2D bytes, [3] D2H bytes, [4] descriptor words in HBM, [5] launches of
 * the descriptor-build kernels, [6] bytes one single-call evaluation (P <= 4)
 * streams from HBM, [7] distinct programs = runs of the class-sorted table (the run table is built in both table modes)            */
int snaq_get_counters(snaq_ctx *ctx, int64_t out[8]);

/* diagnostics (only when the context was created with SNAQ_TRACE=1 in the environment): phase
 * timestamps (globaltimer, ns) of the last single-call kernel launch, out[block*8 + k],
 * k = 0 entry, 1 prologue done, 2 generic region done, 3 additive region done, 4 exit          */
int snaq_get_trace(snaq_ctx *ctx, uint64_t *out, int32_t max_blocks, int32_t *nblocks);

/* measurement aid for the FP64 roofline: best-of-5 rate of a DFMA microbenchmark kernel
 * (8 independent chains per thread, 8 blocks of 256 threads per SM), in TFLOP/s           */
int snaq_measure_fp64_peak(snaq_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* SNAQ_B200_H */
