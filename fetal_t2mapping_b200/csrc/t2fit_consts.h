// Host side: validate a t2fit_problem and fold it into the launch constants (FitConsts).
// Restates the parts of fit_voxel that do not depend on the voxel: which box is in force
// (run_t2mapping.py:243-245), the scipy precondition len(bounds)==len(x0) and lb<=ub
// (scipy/optimize/_minimize.py _validate_bounds -> ValueError), and x0 clipped into the box
// (scipy/optimize/_lbfgsb_py.py: x0 = clip(x0, lb, ub)).
#pragma once
#include <math.h>
#include <string>
#include "../../include/t2fit.h"
#include "t2fit_core.cuh"
#include "t2fit_lbfgsb.cuh"

namespace t2fit {

constexpr int kDefaultMaxIterMono = 24;
constexpr int kDefaultMaxIterFloor = 64;
constexpr float kDefaultTolMono = 2e-3f;
constexpr float kDefaultTolFloor = 1e-5f;

inline int n_params(int model) { return model == T2FIT_MODEL_GAUSSIAN ? 2 : 3; }

// box in force for every voxel (lb[0] is replaced per voxel under --no_prior), run_t2mapping.py:243-245
inline int problem_box(const t2fit_problem& p, double (&lb)[3], double (&ub)[3], std::string& err) {
    const int np_ = n_params(p.model);
    for (int i = 0; i < 3; ++i) { lb[i] = i < np_ ? p.lb[i] : 0.0; ub[i] = i < np_ ? p.ub[i] : 0.0; }
    if (p.no_prior) {
        ub[0] = p.no_prior_k_ub;
        lb[1] = p.no_prior_t2_lb;
        ub[1] = p.no_prior_t2_ub;
    }
    for (int i = 0; i < np_; ++i) {
        if (i == 0 && p.no_prior) continue;          // lb[0] is the voxel's first echo
        if (!(lb[i] <= ub[i])) { err = "An upper bound is less than the corresponding lower bound."; return T2FIT_EINVAL; }
    }
    return T2FIT_OK;
}

// constants of the reference-faithful solver (T2FIT_SOLVER_LBFGSB)
inline int make_lb_consts(const t2fit_problem& p, lb::LbConsts& c, std::string& err) {
    if (p.n_echo < 2 || p.n_echo > kMaxEcho) { err = "n_echo must be in [2, 32]"; return T2FIT_EINVAL; }
    if (p.model < T2FIT_MODEL_GAUSSIAN || p.model > T2FIT_MODEL_RICIAN) { err = "unknown model"; return T2FIT_EINVAL; }
    if (!p.te_ms) { err = "te_ms is NULL"; return T2FIT_EINVAL; }
    if (p.n_fit < 0 || p.n_vox < 0) { err = "negative size"; return T2FIT_EINVAL; }
    double lbv[3], ubv[3];
    int rc = problem_box(p, lbv, ubv, err);
    if (rc) return rc;
    const int np_ = n_params(p.model);
    for (int e = 0; e < kMaxEcho; ++e) {
        const double te = e < p.n_echo ? p.te_ms[e] : 0.0;
        if (!isfinite(te)) { err = "non-finite echo time"; return T2FIT_EINVAL; }
        c.te[e] = te;
    }
    c.te_div_safe = 1;                              // every |TE| is 0 or far from the ends of the exponent range (lb::EchoDiv)
    for (int e = 0; e < p.n_echo; ++e)
        if (p.te_ms[e] != 0.0 && !(fabs(p.te_ms[e]) > 1e-100 && fabs(p.te_ms[e]) < 1e100)) c.te_div_safe = 0;
    for (int i = 0; i < 3; ++i) { c.x0[i] = i < np_ ? p.x0[i] : 0.0; c.lb[i] = lbv[i]; c.ub[i] = ubv[i]; }
    c.ftol = p.lbfgsb_ftol > 0.0 ? p.lbfgsb_ftol : 2.220446049250313e-09;
    c.pgtol = p.lbfgsb_gtol > 0.0 ? p.lbfgsb_gtol : 1e-5;
    c.fd_step = 1e-8;
#ifdef T2FIT_HOSTSIM
    if (p.tol < 0.f) c.fd_step = -1.0;              // hostsim test hook: analytic gradient
#endif
    c.maxls = p.lbfgsb_maxls > 0 ? p.lbfgsb_maxls : 20;
    c.maxiter = p.lbfgsb_maxiter > 0 ? p.lbfgsb_maxiter : 15000;
    c.maxfun = p.lbfgsb_maxfun > 0 ? p.lbfgsb_maxfun : 15000;
    c.n_echo = p.n_echo;
    c.no_prior = p.no_prior ? 1 : 0;
    c.norm = p.norm ? 1 : 0;
    c.objective = p.model;
    c.dense = p.solver == T2FIT_SOLVER_LBFGSB_DENSE ? 1 : 0;
    return T2FIT_OK;
}

inline int make_consts(const t2fit_problem& p, FitConsts& c, std::string& err) {
    if (p.n_echo < 2 || p.n_echo > kMaxEcho) { err = "n_echo must be in [2, 32]"; return T2FIT_EINVAL; }
    if (p.model != T2FIT_MODEL_GAUSSIAN && p.model != T2FIT_MODEL_GAUSSIAN_RICIAN) {
        err = "unknown model"; return T2FIT_EINVAL;
    }
    if (!p.te_ms) { err = "te_ms is NULL"; return T2FIT_EINVAL; }
    if (p.n_fit < 0 || p.n_vox < 0) { err = "negative size"; return T2FIT_EINVAL; }
    const int np_ = n_params(p.model);
    double lb[3] = {p.lb[0], p.lb[1], p.lb[2]}, ub[3] = {p.ub[0], p.ub[1], p.ub[2]};
    if (p.no_prior) {                               // run_t2mapping.py:243-245
        ub[0] = p.no_prior_k_ub;
        lb[1] = p.no_prior_t2_lb;
        ub[1] = p.no_prior_t2_ub;
    }
    for (int i = 0; i < np_; ++i) {
        if (i == 0 && p.no_prior) continue;          // lb[0] is the voxel's first echo
        if (!(lb[i] <= ub[i])) { err = "An upper bound is less than the corresponding lower bound."; return T2FIT_EINVAL; }
    }
    if (!(lb[1] > 0.0)) { err = "T2 lower bound must be positive"; return T2FIT_EINVAL; }
    if (np_ == 3) {
        // the fast 3-parameter solver works in sigma^2 (the model depends on sigma only through it): its box is
        // [max(lb,0)^2, ub^2].  A negative / absent (-inf) lower bound means sigma >= 0; an upper bound below 0 leaves
        // no sigma >= 0 in the box and is rejected (the L-BFGS-B solver handles such boxes in sigma itself).
        if (!(lb[2] >= 0.0)) lb[2] = 0.0;
        if (!(ub[2] >= lb[2])) { err = "sigma upper bound must be >= max(sigma lower bound, 0) for the fast solver (use the L-BFGS-B solver)"; return T2FIT_EINVAL; }
    }
    double mean_te = 0;
    for (int e = 0; e < p.n_echo; ++e) {
        if (!isfinite(p.te_ms[e])) { err = "non-finite echo time"; return T2FIT_EINVAL; }
        mean_te += p.te_ms[e];
    }
    mean_te /= p.n_echo;
    for (int e = 0; e < kMaxEcho; ++e) {
        const double te = e < p.n_echo ? p.te_ms[e] : 0.0;
        c.te[e] = (float)te;
        c.nte2[e] = (float)(-te * 1.4426950408889634);
        c.tec[e] = e < p.n_echo ? (float)(te - mean_te) : 0.f;
    }
    for (int i = 0; i < 3; ++i) {
        c.x0[i] = i < np_ ? (float)p.x0[i] : 0.f;
        c.lb[i] = i < np_ ? (float)lb[i] : 0.f;
        c.ub[i] = i < np_ ? (float)ub[i] : 0.f;
    }
    c.r_lo = (float)(1.0 / ub[1]);
    c.r_hi = (float)(1.0 / lb[1]);
    double t2x0 = fmin(fmax(p.x0[1], lb[1]), ub[1]);
    c.r_x0 = fminf(fmaxf((float)(1.0 / t2x0), c.r_lo), c.r_hi);
    const bool mono = p.model == T2FIT_MODEL_GAUSSIAN;
    c.tol = p.tol > 0.f ? p.tol : (mono ? kDefaultTolMono : kDefaultTolFloor);
    c.max_iter = p.max_iter > 0 ? p.max_iter : (mono ? kDefaultMaxIterMono : kDefaultMaxIterFloor);
    c.n_echo = p.n_echo;
    c.no_prior = p.no_prior ? 1 : 0;
    c.norm = p.norm ? 1 : 0;
    c.init_mode = p.init;
    return T2FIT_OK;
}

}  // namespace t2fit
