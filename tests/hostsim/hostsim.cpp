/*
 * hostsim.cpp — TEST HARNESS ONLY.  Compiles the engine's device functions
 * (snaq.jl_b200/csrc/snaq_core.h) and its host table builder for the CPU so that the
 * structural descriptor build can be fuzzed against the oracle in the CPU-only test
 * tier.  It is not linked into, nor reachable from, the shipped library: the product
 * has no CPU path.
 */
#include <cstring>
#include <string>
#include <vector>

#include "../../snaq.jl_b200/csrc/snaq_tables.h"

static thread_local std::string g_err;

struct XRow {
    const double *x; const double *xc; int np;
    double operator()(unsigned s) const { return (int)s < np ? x[s] : xc[s - np]; }
};

extern "C" {
const char *hostsim_last_error() { return g_err.c_str(); }

int hostsim_params(const snaq_flatnet *f, int ntaxa, int32_t *nparams, int32_t *n_xe, double *x0, int32_t *numht) {
    SnaqHostTables H;
    int rc = snaq_build_tables(f, ntaxa, H, g_err);
    if (rc) return rc;
    *nparams = H.nparams; *n_xe = H.n_xe;
    if (x0) std::memcpy(x0, H.x0.data(), sizeof(double) * H.nparams);
    if (numht) std::memcpy(numht, H.numht.data(), sizeof(int32_t) * H.nparams);
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
    std::vector<std::vector<uint16_t>> progs(nq);
    for (int64_t q = 0; q < nq; q++) {
        SnaqWriter w{nullptr, 0, 0};
        SnaqQuartetInfo info;
        int e = snaq_build_quartet(T, taxa + 4 * q, w, &info);
        if (e) { g_err = "build error " + std::to_string(e) + " at quartet " + std::to_string(q); return SNAQ_ETOPOLOGY; }
        progs[q].assign(w.n, 0);
        SnaqWriter w2{progs[q].data(), 0, w.n};
        snaq_build_quartet(T, taxa + 4 * q, w2, nullptr);
        if (w2.n != w.n) { g_err = "count/emit mismatch"; return SNAQ_EINVAL; }
        if (proglen) proglen[q] = w.n;
        if (which) which[q] = info.which;
        if (formula) for (int i = 0; i < 3; i++) formula[3 * q + i] = info.formula[i];
        if (types) for (int i = 0; i < maxh; i++) types[q * maxh + i] = i < SNAQ_MAX_HYB ? info.type[i] : 0;
    }
    unsigned status = 0;
    for (int p = 0; p < P; p++) {
        XRow xr{X + (size_t)p * H.nparams, H.xconst.data(), H.nparams};
        double total = 0;
        for (int64_t q = 0; q < nq; q++) {
            double cf[3], t1, W[4];
            unsigned kind = snaq_eval_program(progs[q].data(), xr, cf, &t1, status);
            snaq_weights(progs[q][0], obs + 3 * q, sampled ? sampled[q] : 1, W);
            total += snaq_quartet_loglik(kind, cf, t1, W, status);
            if (expcf_last && p == P - 1) snaq_data_order(progs[q][0], cf, expcf_last + 3 * q);
        }
        out_lik[p] = -total;
    }
    if (status_out) *status_out = status;
    return 0;
}
}
