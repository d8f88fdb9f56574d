/*
 * snaq_optbl.cpp — optBL! (src/snaq_optimization.jl:341-390) as a batch-submitting optimiser.
 *
 * The reference hands the objective to NLopt's BOBYQA, which asks for ONE parameter vector at a time
 * (<= 1000 of them).  Here every step of the optimiser is a BATCH of parameter vectors evaluated by one
 * launch of the engine (snaq_eval), or one fused objective+gradient pass (snaq_eval_grad):
 *
 *   gradient   : analytic, one pass over the quartets (snaq_eval_grad), or - SNAQ_OPTBL_FD - central
 *                differences, 2n vectors in one launch (one-sided second-order stencils at a bound), which
 *                also give the diagonal of the Hessian (the initial metric of the quasi-Newton update)
 *   direction  : projected L-BFGS on the free variables (bounds 0 <= gamma, gammaz <= 1, 0 <= t <= 10,
 *                upper(net): src/snaq_optimization.jl:275-279)
 *   line search: the projected points P(x + a_j d) for a geometric ladder of step lengths a_j, ONE launch
 *
 * Stopping rules are NLopt's (ftol_rel, ftol_abs, xtol_rel, xtol_abs, maxeval; the reference passes
 * fRel.. = 1e-6,1e-6,1e-2,1e-3 during the search and fRelBL.. = 1e-12,1e-10,1e-10,1e-10 at the end,
 * src/snaq_optimization.jl:36-47) applied to successive iterates.  Iterates differ from BOBYQA's, so
 * parity with the reference is on the optimum (1e-6 under the tight tolerances), not on
 * the path.  Only the public C ABI is used: this file has no CUDA code.
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "../../include/snaq_b200.h"

namespace {

struct Problem {
    snaq_ctx *ctx;
    int n;
    std::vector<double> lo, hi;
    int64_t nevals = 0, nlaunches = 0;
    int64_t budget = 0;     // what maxeval bounds: objective values the optimiser asked for (a gradient, however obtained, counts as one)
    bool use_fd = false;
};

inline double clampd(double v, double a, double b) { return v < a ? a : (v > b ? b : v); }

// f at a batch of points (rows of X)
int eval_batch(Problem &pb, int P, const std::vector<double> &X, std::vector<double> &f) {
    f.assign(P, 0.0);
    pb.nevals += P;
    pb.budget += P;
    pb.nlaunches += 1;
    return snaq_eval(pb.ctx, P, pb.n, X.data(), f.data());
}

// gradient (and Hessian diagonal) by finite differences: one launch of 2n vectors
int fd_gradient(Problem &pb, const std::vector<double> &x, double f0, std::vector<double> &g, std::vector<double> &hdiag) {
    const int n = pb.n;
    std::vector<double> X((size_t)2 * n * n), f;
    std::vector<double> h(n);
    std::vector<int> mode(n);     // 0 central, +1 forward (x, x+h, x+2h), -1 backward
    for (int i = 0; i < n; i++) {
        const double range = pb.hi[i] - pb.lo[i];
        h[i] = 1e-5 * std::max(1.0, std::min(range, std::fabs(x[i])));
        h[i] = std::min(h[i], 0.25 * range);
        if (x[i] - h[i] < pb.lo[i]) mode[i] = +1;
        else if (x[i] + h[i] > pb.hi[i]) mode[i] = -1;
        else mode[i] = 0;
        double *r0 = &X[(size_t)(2 * i) * n], *r1 = &X[(size_t)(2 * i + 1) * n];
        std::memcpy(r0, x.data(), sizeof(double) * n);
        std::memcpy(r1, x.data(), sizeof(double) * n);
        if (mode[i] == 0) { r0[i] = x[i] + h[i]; r1[i] = x[i] - h[i]; }
        else { r0[i] = x[i] + mode[i] * h[i]; r1[i] = x[i] + 2.0 * mode[i] * h[i]; }
        h[i] = (mode[i] == 0) ? (r0[i] - r1[i]) * 0.5 : std::fabs(r0[i] - x[i]);    // the representable step
    }
    int rc = eval_batch(pb, 2 * n, X, f);
    pb.budget -= 2 * n - 1;            // the stencil of one gradient
    if (rc) return rc;
    g.assign(n, 0.0);
    hdiag.assign(n, 0.0);
    for (int i = 0; i < n; i++) {
        const double a = f[2 * i], b = f[2 * i + 1];
        if (mode[i] == 0) {
            g[i] = (a - b) / (2.0 * h[i]);
            hdiag[i] = (a - 2.0 * f0 + b) / (h[i] * h[i]);
        } else {
            g[i] = mode[i] * (-3.0 * f0 + 4.0 * a - b) / (2.0 * h[i]);
            hdiag[i] = (f0 - 2.0 * a + b) / (h[i] * h[i]);
        }
    }
    return SNAQ_OK;
}

// want_h: also refresh the curvature diagonal (the analytic pass does extra work for it)
int gradient(Problem &pb, const std::vector<double> &x, double &f0, bool have_f, std::vector<double> &g, std::vector<double> &hdiag,
             bool want_h) {
    if (!pb.use_fd) {
        g.assign(pb.n, 0.0);
        if (want_h) hdiag.assign(pb.n, 0.0);
        pb.nevals += 1;
        pb.budget += 1;
        pb.nlaunches += 1;
        return snaq_eval_grad(pb.ctx, pb.n, x.data(), &f0, g.data(), want_h ? hdiag.data() : nullptr);
    }
    if (!have_f) {
        std::vector<double> f;
        int rc = eval_batch(pb, 1, x, f);
        if (rc) return rc;
        f0 = f[0];
    }
    return fd_gradient(pb, x, f0, g, hdiag);
}

struct Pair { std::vector<double> s, y; double rho; };

}  // namespace

extern "C" int snaq_optbl(snaq_ctx *ctx, const snaq_optbl_opts *o, double *x_io, snaq_optbl_result *res) {
    if (!ctx || !x_io || !res) return SNAQ_EINVAL;
    std::memset(res, 0, sizeof(*res));
    int32_t n = 0;
    int rc = snaq_num_params(ctx, &n, nullptr, nullptr, nullptr);
    if (rc) return rc;
    const double ftol_rel = o ? o->ftol_rel : 1e-6, ftol_abs = o ? o->ftol_abs : 1e-6;
    const double xtol_rel = o ? o->xtol_rel : 1e-2, xtol_abs = o ? o->xtol_abs : 1e-3;
    const int maxeval = (o && o->maxeval > 0) ? o->maxeval : 1000;
    const int ladder_full = (o && o->ladder > 0) ? std::min(o->ladder, 64) : 20;
    if (!(ftol_rel > 0 && ftol_abs > 0 && xtol_rel > 0 && xtol_abs > 0)) return SNAQ_EINVAL;   // :350
    Problem pb;
    pb.ctx = ctx;
    pb.n = n;
    pb.use_fd = (o && (o->flags & SNAQ_OPTBL_FD)) || getenv("SNAQ_OPTBL_FD") != nullptr;
    pb.lo.assign(n, 0.0);
    pb.hi.assign(n, 1.0);
    if (n > 0) {
        rc = snaq_get_params(ctx, nullptr, nullptr, pb.hi.data());
        if (rc) return rc;
    }
    std::vector<double> x(x_io, x_io + n);
    for (int i = 0; i < n; i++) x[i] = clampd(x[i], pb.lo[i], pb.hi[i]);
    {   // a quartet-sharded context without a communicator only knows its own partial sums: refuse (the gradient call
        // carries the message); with one, every rank runs this same loop on the same totals
        int32_t world = 1, has_comm = 0;
        if (snaq_comm_info(ctx, nullptr, &world, &has_comm, nullptr) == SNAQ_OK && world > 1 && !has_comm) {
            double f0 = 0.0;
            std::vector<double> g0(std::max(n, 1));
            rc = snaq_eval_grad(ctx, n, x.data(), &f0, g0.data(), nullptr);
            return rc ? rc : SNAQ_EINVAL;
        }
    }
    double f = 0.0;
    std::vector<double> g, hd;
    rc = gradient(pb, x, f, false, g, hd, true);
    if (rc) return rc;
    if (!std::isfinite(f)) {
        // a start on the boundary can have an expected CF of exactly 0 under a positive observed CF (deviance +Inf, no slope):
        // step a hair inside the box and start from there
        for (int i = 0; i < n; i++) { const double m = 1e-4 * (pb.hi[i] - pb.lo[i]); x[i] = clampd(x[i], pb.lo[i] + m, pb.hi[i] - m); }
        rc = gradient(pb, x, f, false, g, hd, true);
        if (rc) return rc;
    }
    if (n == 0) { res->fmin = f; res->nevals = (int32_t)pb.nevals; res->nlaunches = (int32_t)pb.nlaunches; res->status = SNAQ_OPTBL_XTOL; return SNAQ_OK; }
    const int M = std::max(12, std::min(n, 64));   // quasi-Newton pairs kept (full memory for small networks)
    std::deque<Pair> mem;
    int status = SNAQ_OPTBL_MAXEVAL, iter = 0, small_f = 0, small_x = 0;
    std::vector<double> d(n), q(n), xn(n), Xls, fls, gn, hdn;
    const double bnd_eps = 1e-12;
    const bool no_unit_step = getenv("SNAQ_OPTBL_NO_UNIT_STEP") != nullptr;
    const int hd_every = getenv("SNAQ_OPTBL_HD_EVERY") ? std::max(1, atoi(getenv("SNAQ_OPTBL_HD_EVERY"))) : 8;
    const bool verbose = getenv("SNAQ_OPTBL_VERBOSE") != nullptr;      // "inside obj with x ..." of the reference (:369-378)
    // maxeval bounds the number of engine launches: NLopt.maxeval! (:365) bounds BOBYQA's SEQUENTIAL objective calls, and a
    // launch (one gradient pass, or one ladder of up to `ladder` vectors evaluated together) is this optimiser's
    // sequential step; res->nevals reports the parameter vectors evaluated
    while (pb.nlaunches < maxeval) {
        iter++;
        // free variables: not pinned at a bound by the gradient
        std::vector<char> freev(n, 1);
        double pg = 0.0;
        for (int i = 0; i < n; i++) {
            if ((x[i] <= pb.lo[i] + bnd_eps && g[i] > 0.0) || (x[i] >= pb.hi[i] - bnd_eps && g[i] < 0.0)) freev[i] = 0;
            pg = std::max(pg, std::fabs(clampd(x[i] - g[i], pb.lo[i], pb.hi[i]) - x[i]));
        }
        if (pg == 0.0) { status = SNAQ_OPTBL_XTOL; break; }
        // initial metric: Hessian diagonal where it is usable, else the usual gamma scaling
        std::vector<double> h0(n, 1.0);
        double gam = 1.0;
        if (!mem.empty()) {
            double yy = 0.0, sy = 0.0;
            for (int i = 0; i < n; i++) { yy += mem.back().y[i] * mem.back().y[i]; sy += mem.back().s[i] * mem.back().y[i]; }
            if (yy > 0 && sy > 0) gam = sy / yy;
        }
        for (int i = 0; i < n; i++) h0[i] = (hd[i] > 1e-8) ? 1.0 / hd[i] : gam;
        // two-loop recursion on the free subspace
        for (int i = 0; i < n; i++) q[i] = freev[i] ? g[i] : 0.0;
        std::vector<double> al(mem.size());
        for (int k = (int)mem.size() - 1; k >= 0; k--) {
            double a = 0.0;
            for (int i = 0; i < n; i++) if (freev[i]) a += mem[k].s[i] * q[i];
            a *= mem[k].rho;
            al[k] = a;
            for (int i = 0; i < n; i++) if (freev[i]) q[i] -= a * mem[k].y[i];
        }
        if (mem.empty()) for (int i = 0; i < n; i++) q[i] *= h0[i];
        else for (int i = 0; i < n; i++) q[i] *= (hd[i] > 1e-8 ? h0[i] : gam);
        for (size_t k = 0; k < mem.size(); k++) {
            double b = 0.0;
            for (int i = 0; i < n; i++) if (freev[i]) b += mem[k].y[i] * q[i];
            b *= mem[k].rho;
            for (int i = 0; i < n; i++) if (freev[i]) q[i] += (al[k] - b) * mem[k].s[i];
        }
        double gd = 0.0;
        for (int i = 0; i < n; i++) { d[i] = freev[i] ? -q[i] : 0.0; gd += d[i] * g[i]; }
        // inside the -1e15 plateau of a negative expCF (src/pseudolik.jl:1394-1396) the objective is flat: follow the
        // plain (repair) gradient with the range-normalised ladder until a rung leaves the plateau
        const bool plateau = f >= 1e14;
        if (plateau) mem.clear();
        if (plateau || !(gd < 0.0)) {           // not a descent direction: fall back to the scaled projected gradient
            mem.clear();
            gd = 0.0;
            for (int i = 0; i < n; i++) { d[i] = freev[i] ? -g[i] * (plateau ? 1.0 : h0[i]) : 0.0; gd += d[i] * g[i]; }
        }
        if (mem.empty()) {
            // no curvature information yet: scale the step so that a = 1 moves the fastest coordinate by a
            // quarter of its range (the ladder then spans 1 .. 2^-18 of the range)
            double big = 0.0;
            for (int i = 0; i < n; i++) big = std::max(big, std::fabs(d[i]) / (pb.hi[i] - pb.lo[i]));
            if (big > 0.25) { const double sc = 0.25 / big; gd *= sc; for (int i = 0; i < n; i++) d[i] *= sc; }
        }
        const bool want_h = pb.use_fd || (iter % hd_every == 0);      // curvature diagonal: refreshed now and then
        if (!want_h) hdn = hd;
        // with curvature pairs in hand the unit step is usually right: try it (then the minimiser of the parabola through
        // f(0), f'(0), f(1)) with a fused objective+gradient pass and accept on sufficient decrease; the ladder is the fallback
        bool accepted = false;
        double fn = 0.0, a_try = 1.0;
        if (!pb.use_fd && !mem.empty() && !plateau && !no_unit_step) {
            double a = 1.0;
            for (int attempt = 0; attempt < 2 && !accepted; attempt++) {
                double gs = 0.0;
                for (int i = 0; i < n; i++) { xn[i] = clampd(x[i] + a * d[i], pb.lo[i], pb.hi[i]); gs += g[i] * (xn[i] - x[i]); }
                double fr = 0.0;
                rc = gradient(pb, xn, fr, false, gn, hdn, want_h);
                if (rc) return rc;
                if (std::isfinite(fr) && gs < 0.0 && fr <= f + 1e-4 * gs) { accepted = true; fn = fr; a_try = a; break; }
                if (!std::isfinite(fr) || !(gs < 0.0)) break;
                const double denom = 2.0 * (fr - f - gs);            // gs = slope * a for an interior step
                if (!(denom > 0.0)) break;
                const double aq = a * (-gs / denom);
                if (!(aq > 1e-3 * a && aq < 0.9 * a)) break;
                a = aq;
            }
            if (accepted && verbose)
                fprintf(stderr, "OPTBL iter %d: f %.12g -> %.12g (unit-step path, alpha %g, mem %zu, pg %.3g, g.d %.3g)\n", iter, f, fn, a_try, mem.size(), pg, gd);
        }
        if (!accepted) {
        // line search: one launch over a ladder of projected steps
        // with curvature pairs the step length is near 1: eight rungs (4 .. 1/32) in one small launch; the full ladder is
        // for the first iteration, the plateau, and the retry after a failed short ladder
        const int ladder = (mem.empty() || plateau) ? ladder_full : std::min(ladder_full, 8);
        Xls.assign((size_t)ladder * n, 0.0);
        std::vector<double> alphas(ladder);
        for (int j = 0; j < ladder; j++) {
            alphas[j] = std::ldexp(1.0, 2 - j);          // 4, 2, 1, 1/2, ...
            for (int i = 0; i < n; i++) Xls[(size_t)j * n + i] = clampd(x[i] + alphas[j] * d[i], pb.lo[i], pb.hi[i]);
        }
        rc = eval_batch(pb, ladder, Xls, fls);
        if (rc) return rc;
        int best = -1;
        for (int j = 0; j < ladder; j++)
            if (std::isfinite(fls[j]) && fls[j] < f && (best < 0 || fls[j] < fls[best])) best = j;
        if (best < 0) {
            if (verbose) fprintf(stderr, "OPTBL iter %d: no decrease on the ladder (f %.12g, best of ladder %.12g, mem %zu, pg %.3g, g.d %.3g)\n", iter, f,
                                 *std::min_element(fls.begin(), fls.end()), mem.size(), pg, gd);
            if (!mem.empty()) { mem.clear(); continue; }          // retry once along the scaled gradient
            status = SNAQ_OPTBL_FTOL;                              // no decrease along the gradient: stationary to rounding
            break;
        }
        // refine the step with the parabola through the best rung and its neighbours; the refined point is
        // evaluated by the gradient pass itself (no extra launch) and kept only if it is not worse
        a_try = alphas[best];
        if (best > 0 && best + 1 < ladder && std::isfinite(fls[best - 1]) && std::isfinite(fls[best + 1])) {
            const double a0 = alphas[best + 1], a1 = alphas[best], a2 = alphas[best - 1];
            const double f0 = fls[best + 1], f1 = fls[best], f2 = fls[best - 1];
            const double den = (a1 - a0) * (f1 - f2) - (a1 - a2) * (f1 - f0);
            if (den != 0.0) {
                const double a = a1 - 0.5 * ((a1 - a0) * (a1 - a0) * (f1 - f2) - (a1 - a2) * (a1 - a2) * (f1 - f0)) / den;
                if (a > a0 && a < a2) a_try = a;
            }
        }
        for (int i = 0; i < n; i++) xn[i] = clampd(x[i] + a_try * d[i], pb.lo[i], pb.hi[i]);
        fn = fls[best];
        if (a_try != alphas[best]) {
            double fr = 0.0;
            rc = gradient(pb, xn, fr, false, gn, hdn, want_h);
            if (rc) return rc;
            if (std::isfinite(fr) && fr <= fls[best]) fn = fr;
            else { a_try = alphas[best]; for (int i = 0; i < n; i++) xn[i] = Xls[(size_t)best * n + i]; }
        }
        if (verbose) {
            int nfree = 0;
            for (int i = 0; i < n; i++) nfree += freev[i];
            fprintf(stderr, "OPTBL iter %d: f %.12g -> %.12g (alpha %g, mem %zu, free %d/%d, pg %.3g, g.d %.3g)\n", iter, f, fn, a_try,
                    mem.size(), nfree, n, pg, gd);
        }
        if (a_try == alphas[best]) {
            rc = gradient(pb, xn, fn, true, gn, hdn, want_h);
            if (rc) return rc;
        }
        }   // ladder
        // quasi-Newton pair
        Pair p;
        p.s.resize(n); p.y.resize(n);
        double sy = 0.0, ss = 0.0, yy = 0.0;
        for (int i = 0; i < n; i++) { p.s[i] = xn[i] - x[i]; p.y[i] = gn[i] - g[i]; sy += p.s[i] * p.y[i]; ss += p.s[i] * p.s[i]; yy += p.y[i] * p.y[i]; }
        if (sy > 1e-10 * std::sqrt(ss * yy)) { p.rho = 1.0 / sy; mem.push_back(std::move(p)); if ((int)mem.size() > M) mem.pop_front(); }
        // NLopt-style stopping rules on successive iterates
        const double df = f - fn;
        bool xsmall = true;
        for (int i = 0; i < n; i++) {
            const double dx = std::fabs(xn[i] - x[i]);
            if (!(dx <= xtol_abs || dx <= xtol_rel * std::fabs(xn[i]))) { xsmall = false; break; }
        }
        x = xn; f = fn; g = gn; hd = hdn;
        if (df <= ftol_abs || df <= ftol_rel * std::fabs(fn)) small_f++; else small_f = 0;
        if (small_f >= 2) { status = SNAQ_OPTBL_FTOL; break; }
        // XTOL: two consecutive small steps, both taken with curvature information (one short rung of the first
        // scaled-gradient iteration says nothing about the distance to the optimum)
        if (xsmall && !mem.empty()) small_x++; else small_x = 0;
        if (small_x >= 2) { status = SNAQ_OPTBL_XTOL; break; }
    }
    std::memcpy(x_io, x.data(), sizeof(double) * n);
    res->fmin = f;
    res->nevals = (int32_t)std::min<int64_t>(pb.nevals, 2147483647);
    res->nlaunches = (int32_t)pb.nlaunches;
    res->niter = iter;
    res->status = status;
    double pg = 0.0;
    for (int i = 0; i < n; i++) pg = std::max(pg, std::fabs(clampd(x[i] - g[i], pb.lo[i], pb.hi[i]) - x[i]));
    res->proj_grad = pg;
    return SNAQ_OK;
}
