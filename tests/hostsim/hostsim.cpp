/*
 * hostsim.cpp — TEST HARNESS ONLY.  Compiles the engine's device functions
 * (snaq.jl_b200/csrc/snaq_core.h) and its host table builder for the CPU so that the
 * structural descriptor build and the specialised exp/log can be fuzzed against the oracle
 * in the CPU-only test tier.  It is not linked into, nor reachable from, the shipped
 * library: the product has no CPU path.
 */
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../snaq.jl_b200/csrc/snaq_hasedge.h"
#include "../../snaq.jl_b200/csrc/snaq_tables.h"

static thread_local std::string g_err;
static const double g_et[16] = SNAQ_EXP_TAB_INIT;
static const double g_li[16] = SNAQ_LOGINV_TAB_INIT;
static const double g_lc[16] = SNAQ_LOGC_TAB_INIT;

struct XRow {
    const double *xf;
    double operator()(unsigned s) const { return xf[s]; }
};

extern "C" {
const char *hostsim_last_error() { return g_err.c_str(); }

void hostsim_math(int n, const double *u, double *e, double *l) {
    SnaqMath M{g_et, g_li, g_lc};
    for (int i = 0; i < n; i++) {
        if (e) e[i] = snaq_exp(u[i], M);
        if (l) l[i] = snaq_log(u[i], M);
    }
}

// the host side of snaq_patch_network: the new tables rooted where the old ones are, and the taxa whose quartets have to be
// re-described.  Returns 1 and fills touched[ntaxa], or 0 when the patch falls back to a full classification; < 0 on error.
int hostsim_touched(const snaq_flatnet *f_old, const snaq_flatnet *f_new, int ntaxa, uint8_t *touched) {
    SnaqHostTables A, B;
    if (snaq_build_tables(f_old, ntaxa, A, g_err)) return -1;
    if (snaq_build_tables(f_new, ntaxa, B, g_err, A.root)) return -2;
    std::vector<uint8_t> t;
    if (!snaq_tables_touched(A, B, t)) return 0;
    for (int i = 0; i < ntaxa; i++) touched[i] = t[(size_t)i];
    return 1;
}

// programs of every quartet on tables rooted at `keep_root` (-1: the default root): first 8 words, length, header
int hostsim_progs_rooted(const snaq_flatnet *f_root, const snaq_flatnet *f, int ntaxa, int64_t nq, const uint16_t *taxa, uint16_t *words8,
                         int32_t *len, uint8_t *hdr) {
    SnaqHostTables R, H;
    if (snaq_build_tables(f_root, ntaxa, R, g_err)) return -1;
    if (snaq_build_tables(f, ntaxa, H, g_err, R.root)) return -2;
    SnaqTables T = H.view(H.blob.data());
    for (int64_t q = 0; q < nq; q++) {
        uint16_t buf[256];
        for (auto &b : buf) b = (uint16_t)H.n_xe;
        SnaqWriter w{buf, 0, 256};
        uint8_t h1 = 0;
        if (snaq_build_quartet(T, taxa + 4 * q, w, &h1, nullptr)) { g_err = "build error"; return -3; }
        for (int k = 0; k < 8; k++) words8[8 * q + k] = buf[k];
        len[q] = w.n;
        hdr[q] = h1;
    }
    return 0;
}

void hostsim_div3(int n, const double *g, double *q) {
    for (int i = 0; i < n; i++) q[i] = snaq_div3(g[i]);
}

int hostsim_params(const snaq_flatnet *f, int ntaxa, int32_t *nparams, int32_t *n_xe, double *x0, int32_t *numht) {
    SnaqHostTables H;
    int rc = snaq_build_tables(f, ntaxa, H, g_err);
    if (rc) return rc;
    *nparams = H.nparams; *n_xe = H.n_xe;
    if (x0) std::memcpy(x0, H.x0.data(), sizeof(double) * H.nparams);
    if (numht) std::memcpy(numht, H.numht.data(), sizeof(int32_t) * H.nparams);
    return 0;
}

int hostsim_counts(const snaq_flatnet *f, int ntaxa, int32_t *nh, int32_t *nt, int32_t *nhz) {
    SnaqHostTables H;
    int rc = snaq_build_tables(f, ntaxa, H, g_err);
    if (rc) return rc;
    *nh = H.nh; *nt = H.nt; *nhz = H.nhz;
    return 0;
}

/* the same through the simulation of extractQuartet! (snaq_hesim.h): any level-1 network */
int hostsim_hasedge_sim(const snaq_flatnet *f, int ntaxa, int64_t nq, const uint16_t *taxa, uint8_t *hasedge, uint8_t *bd1, int16_t *zpos) {
    SnaqHostTables H;
    int rc = snaq_build_tables(f, ntaxa, H, g_err);
    if (rc) return rc;
    const SnaqSimNet net = H.sim_view(H.sim_blob.data());
    std::vector<uint32_t> bits((H.nparams + 31) / 32 + 1);
    std::vector<SnaqSim> scratch(1);
    for (int64_t q = 0; q < nq; q++) {
        int flag = 0;
        const int nz = H.nparams - H.nh - H.nt;
        int e = snaq_hasedge_sim(net, taxa + 4 * q, scratch[0], bits.data(), &flag, (zpos && nz > 0) ? zpos + q * nz : nullptr);
        if (e) { g_err = "hasedge simulation error " + std::to_string(e) + " (line " + std::to_string(scratch[0].err) + ") at quartet " + std::to_string(q); return SNAQ_ETOPOLOGY; }
        for (int i = 0; i < H.nparams; i++) hasedge[q * H.nparams + i] = (bits[i >> 5] >> (i & 31)) & 1u;
        if (bd1) bd1[q] = (uint8_t)flag;
    }
    return 0;
}

/* hasEdge bits (one byte per parameter) and the "bad diamond I kept whole" flag of every quartet */
int hostsim_hasedge(const snaq_flatnet *f, int ntaxa, int64_t nq, const uint16_t *taxa, uint8_t *hasedge, uint8_t *bd1) {
    SnaqHostTables H;
    int rc = snaq_build_tables(f, ntaxa, H, g_err);
    if (rc) return rc;
    SnaqTables T = H.view(H.blob.data());
    std::vector<uint32_t> bits(SNAQ_HE_WORDS(H.nparams) + 1);
    for (int64_t q = 0; q < nq; q++) {
        int flag = 0;
        int e = snaq_hasedge_quartet(T, taxa + 4 * q, bits.data(), &flag);
        if (e) { g_err = "hasedge error " + std::to_string(e) + " at quartet " + std::to_string(q); return SNAQ_ETOPOLOGY; }
        for (int i = 0; i < H.nparams; i++) hasedge[q * H.nparams + i] = (bits[i >> 5] >> (i & 31)) & 1u;
        if (bd1) bd1[q] = (uint8_t)flag;
    }
    return 0;
}

/* builds every program, optionally evaluates P parameter vectors */
int hostsim_run(const snaq_flatnet *f, int ntaxa, int64_t nq, const uint16_t *taxa, const double *obs,
                const uint8_t *sampled, int32_t P, const double *X, double *out_lik, double *expcf_last,
                int8_t *which, int8_t *types, int32_t maxh, int8_t *formula, int32_t *proglen, uint32_t *status_out) {
    SnaqHostTables H;
    int rc = snaq_build_tables(f, ntaxa, H, g_err);
    if (rc) return rc;
    SnaqTables T = H.view(H.blob.data());
    // the pair table the device fills (k_leaf_lca): the counting pass walks the parent pointers, the emitting pass and a
    // third, compared word by word, look the six LCAs up
    std::vector<uint16_t> pair_lca((size_t)ntaxa * ntaxa, 0xFFFF);
    for (int i = 0; i < ntaxa; i++)
        for (int j = 0; j < ntaxa; j++)
            if (T.leaf_tnode[i] != 0xFFFF && T.leaf_tnode[j] != 0xFFFF) pair_lca[(size_t)i * ntaxa + j] = (uint16_t)snaq_lca(T, T.leaf_tnode[i], T.leaf_tnode[j]);
    SnaqTables TL = T;
    TL.leaf_lca = pair_lca.data();
    SnaqMath M{g_et, g_li, g_lc};
    std::vector<std::vector<uint16_t>> progs(nq);
    std::vector<uint8_t> hdrs(nq);
    std::vector<uint16_t> chk;
    for (int64_t q = 0; q < nq; q++) {
        SnaqWriter w{nullptr, 0, 0};
        SnaqQuartetInfo info;
        uint8_t h1 = 0, h2 = 0, h3 = 0;
        int e = snaq_build_quartet(T, taxa + 4 * q, w, &h1, &info);
        if (e) { g_err = "build error " + std::to_string(e) + " at quartet " + std::to_string(q); return SNAQ_ETOPOLOGY; }
        const int padded = (w.n + 3) & ~3;
        progs[q].assign(padded, (uint16_t)H.n_xe);      // NULL-slot padding
        SnaqWriter w2{progs[q].data(), 0, w.n};
        snaq_build_quartet(T, taxa + 4 * q, w2, &h2, nullptr);
        if (w2.n != w.n || h1 != h2) { g_err = "count/emit mismatch"; return SNAQ_EINVAL; }
        chk.assign(padded, (uint16_t)H.n_xe);
        SnaqWriter w3{chk.data(), 0, w.n};
        if (snaq_build_quartet(TL, taxa + 4 * q, w3, &h3, nullptr) || w3.n != w.n || h3 != h1 || chk != progs[q]) {
            g_err = "pair-table LCA path differs from the walks at quartet " + std::to_string(q); return SNAQ_EINVAL;
        }
        hdrs[q] = h1;
        if (proglen) proglen[q] = w.n;
        if (which) which[q] = info.which;
        if (formula) for (int i = 0; i < 3; i++) formula[3 * q + i] = info.formula[i];
        if (types) for (int i = 0; i < maxh; i++) types[q * maxh + i] = i < SNAQ_MAX_HYB ? info.type[i] : 0;
    }
    unsigned status = 0;
    for (int p = 0; p < P; p++) {
        std::vector<double> xf(H.n_full);
        for (int r = 0; r < H.n_full; r++) xf[r] = snaq_expand_row(T, M, X + (size_t)p * H.nparams, H.xconst.data(), r);
        XRow xr{xf.data()};
        double total = 0;
        for (int64_t q = 0; q < nq; q++) {
            double cf[3], t1, W[4];
            snaq_eval_quartet(hdrs[q], progs[q].data(), (int)progs[q].size(), xr, M, cf, &t1, status);
            snaq_weights(hdrs[q], obs + 3 * q, sampled ? sampled[q] : 1, W);
            total += snaq_quartet_loglik(SNAQ_HDR_KIND(hdrs[q]), cf, t1, W, M, status);
            if (expcf_last && p == P - 1) snaq_data_order(hdrs[q], cf, expcf_last + 3 * q);
        }
        out_lik[p] = -total;
    }
    if (status_out) *status_out = status;
    return 0;
}

/* the first 8 words of every program (NULL padded), its length in words and its header byte: layout statistics */
int hostsim_progs(const snaq_flatnet *f, int ntaxa, int64_t nq, const uint16_t *taxa, uint16_t *words8, int32_t *len, uint8_t *hdr) {
    SnaqHostTables H;
    int rc = snaq_build_tables(f, ntaxa, H, g_err);
    if (rc) return rc;
    SnaqTables T = H.view(H.blob.data());
    for (int64_t q = 0; q < nq; q++) {
        uint16_t buf[256];
        for (auto &b : buf) b = (uint16_t)H.n_xe;
        SnaqWriter w{buf, 0, 256};
        uint8_t h1 = 0;
        int e = snaq_build_quartet(T, taxa + 4 * q, w, &h1, nullptr);
        if (e) { g_err = "build error"; return SNAQ_ETOPOLOGY; }
        for (int k = 0; k < 8; k++) words8[8 * q + k] = buf[k];
        len[q] = w.n;
        hdr[q] = h1;
    }
    return 0;
}

/* objective and analytic gradient of one parameter vector (the device functions of snaq_eval_grad, on the CPU) */
int hostsim_grad(const snaq_flatnet *f, int ntaxa, int64_t nq, const uint16_t *taxa, const double *obs,
                 const uint8_t *sampled, const double *x, double *out_lik, double *grad, uint32_t *status_out) {
    SnaqHostTables H;
    int rc = snaq_build_tables(f, ntaxa, H, g_err);
    if (rc) return rc;
    SnaqTables T = H.view(H.blob.data());
    SnaqMath M{g_et, g_li, g_lc};
    std::vector<double> xf(H.n_full), gxf(H.n_full, 0.0);
    for (int r = 0; r < H.n_full; r++) xf[r] = snaq_expand_row(T, M, x, H.xconst.data(), r);
    XRow xr{xf.data()};
    unsigned status = 0;
    double total = 0;
    auto add = [&](unsigned slot, double v) { gxf[slot] += v; };
    auto addh = [&](unsigned, double) {};
    for (int64_t q = 0; q < nq; q++) {
        SnaqWriter w{nullptr, 0, 0};
        uint8_t h1 = 0;
        int e = snaq_build_quartet(T, taxa + 4 * q, w, &h1, nullptr);
        if (e) { g_err = "build error " + std::to_string(e) + " at quartet " + std::to_string(q); return SNAQ_ETOPOLOGY; }
        std::vector<uint16_t> prog((w.n + 3) & ~3, (uint16_t)H.n_xe);
        SnaqWriter w2{prog.data(), 0, w.n};
        snaq_build_quartet(T, taxa + 4 * q, w2, &h1, nullptr);
        double W[4];
        snaq_weights(h1, obs + 3 * q, sampled ? sampled[q] : 1, W);
        total += snaq_grad_quartet(h1, prog.data(), (int)prog.size(), xr, M, W, (unsigned)H.n_xe, add, addh, status);
    }
    for (int r = 0; r < H.n_full; r++) gxf[r] = -gxf[r];            /* f = -sum v */
    for (int i = 0; i < H.nparams; i++) grad[i] = 0.0;
    snaq_grad_backprop(T, M, x, H.xconst.data(), gxf.data(), grad);
    *out_lik = -total;
    if (status_out) *status_out = status;
    return 0;
}

/*
 * CPU stand-in for the engine's context, so that the host-side optimiser (snaq.jl_b200/csrc/snaq_optbl.cpp,
 * which only uses the public C ABI) can be exercised in the CPU-only test tier: the four ABI calls it makes
 * are implemented here on top of the same device functions.  Test harness only.
 */
struct snaq_ctx {
    SnaqHostTables H;
    SnaqTables T;
    int64_t nq = 0;
    std::vector<std::vector<uint16_t>> progs;
    std::vector<uint8_t> hdrs;
    std::vector<double> W;
};

int snaq_comm_info(snaq_ctx *, int32_t *rank, int32_t *world_size, int32_t *has_comm, int32_t *fused) {
    if (rank) *rank = 0;
    if (world_size) *world_size = 1;
    if (has_comm) *has_comm = 0;
    if (fused) *fused = 0;
    return 0;
}
int snaq_num_params(snaq_ctx *c, int32_t *nparams, int32_t *n_gamma, int32_t *n_t, int32_t *n_gammaz) {
    if (nparams) *nparams = c->H.nparams;
    if (n_gamma) *n_gamma = c->H.nh;
    if (n_t) *n_t = c->H.nt;
    if (n_gammaz) *n_gammaz = c->H.nparams - c->H.nh - c->H.nt;
    return 0;
}
int snaq_get_params(snaq_ctx *c, double *x, int32_t *numht, double *upper) {
    for (int i = 0; i < c->H.nparams; i++) {
        if (x) x[i] = c->H.x0[i];
        if (numht) numht[i] = c->H.numht[i];
        if (upper) upper[i] = (i >= c->H.nh && i < c->H.nh + c->H.nt) ? 10.0 : 1.0;
    }
    return 0;
}
int snaq_eval(snaq_ctx *c, int32_t P, int32_t nparams, const double *X, double *out) {
    SnaqMath M{g_et, g_li, g_lc};
    unsigned status = 0;
    std::vector<double> xf(c->H.n_full);
    for (int p = 0; p < P; p++) {
        for (int r = 0; r < c->H.n_full; r++) xf[r] = snaq_expand_row(c->T, M, X + (size_t)p * nparams, c->H.xconst.data(), r);
        XRow xr{xf.data()};
        double total = 0;
        for (int64_t q = 0; q < c->nq; q++) {
            double cf[3], t1;
            snaq_eval_quartet(c->hdrs[q], c->progs[q].data(), (int)c->progs[q].size(), xr, M, cf, &t1, status);
            total += snaq_quartet_loglik(SNAQ_HDR_KIND(c->hdrs[q]), cf, t1, &c->W[4 * q], M, status);
        }
        out[p] = -total;
    }
    return status ? SNAQ_ENUMERIC : 0;
}
int snaq_eval_grad(snaq_ctx *c, int32_t nparams, const double *x, double *f, double *grad, double *hdiag) {
    SnaqMath M{g_et, g_li, g_lc};
    unsigned status = 0;
    std::vector<double> xf(c->H.n_full), gxf(c->H.n_full, 0.0), hxf(c->H.n_full, 0.0);
    for (int r = 0; r < c->H.n_full; r++) xf[r] = snaq_expand_row(c->T, M, x, c->H.xconst.data(), r);
    XRow xr{xf.data()};
    auto add = [&](unsigned slot, double v) { gxf[slot] += v; };
    auto addh = [&](unsigned slot, double v) { hxf[slot] += v; };
    double total = 0;
    for (int64_t q = 0; q < c->nq; q++)
        total += snaq_grad_quartet(c->hdrs[q], c->progs[q].data(), (int)c->progs[q].size(), xr, M, &c->W[4 * q], (unsigned)c->H.n_xe, add, addh, status);
    for (auto &v : gxf) v = -v;
    for (int i = 0; i < nparams; i++) { grad[i] = 0.0; if (hdiag) hdiag[i] = 0.0; }
    snaq_grad_backprop(c->T, M, x, c->H.xconst.data(), gxf.data(), grad);
    if (hdiag && !getenv("HOSTSIM_NO_HDIAG")) snaq_grad_backprop(c->T, M, x, c->H.xconst.data(), hxf.data(), hdiag, true);
    *f = -total;
    return status ? SNAQ_ENUMERIC : 0;
}

/* optBL! on the CPU stand-in: the optimiser code is the product's, the objective is the harness's */
int hostsim_optbl(const snaq_flatnet *f, int ntaxa, int64_t nq, const uint16_t *taxa, const double *obs, const snaq_optbl_opts *opts,
                  double *x, snaq_optbl_result *res) {
    snaq_ctx c;
    int rc = snaq_build_tables(f, ntaxa, c.H, g_err);
    if (rc) return rc;
    c.T = c.H.view(c.H.blob.data());
    c.nq = nq;
    c.progs.resize(nq); c.hdrs.resize(nq); c.W.resize(4 * nq);
    for (int64_t q = 0; q < nq; q++) {
        SnaqWriter w{nullptr, 0, 0};
        uint8_t h1 = 0;
        int e = snaq_build_quartet(c.T, taxa + 4 * q, w, &h1, nullptr);
        if (e) { g_err = "build error " + std::to_string(e) + " at quartet " + std::to_string(q); return SNAQ_ETOPOLOGY; }
        c.progs[q].assign((w.n + 3) & ~3, (uint16_t)c.H.n_xe);
        SnaqWriter w2{c.progs[q].data(), 0, w.n};
        snaq_build_quartet(c.T, taxa + 4 * q, w2, &h1, nullptr);
        c.hdrs[q] = h1;
        snaq_weights(h1, obs + 3 * q, 1, &c.W[4 * q]);
    }
    return snaq_optbl(&c, opts, x, res);
}
}
