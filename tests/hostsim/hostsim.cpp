// Host simulation of the device solver core -- TEST TOOL ONLY.
//
// Compiles fetal_t2mapping_b200/csrc/t2fit_core.cuh with g++ (T2FIT_HOSTSIM) so the `-m "not gpu"`
// suite can unit-test the solver LOGIC (bracketing, active set, status codes) on a GPU-less CI box,
// in float (as shipped) and in double (to separate algorithmic from rounding effects).
// It is never linked into libt2fit, never imported by the fetal_t2mapping_b200 package, and is
// not a fallback: the product fails loudly without a CUDA device.
#define T2FIT_HOSTSIM 1
#include "../../fetal_t2mapping_b200/csrc/t2fit_consts.h"
#include "../../fetal_t2mapping_b200/csrc/t2fit_lbfgsb_coop.cuh"
#include "../../fetal_t2mapping_b200/csrc/t2fit_lbfgsb_dense.cuh"

#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>

using namespace t2fit;

template <typename R, int MODEL, int E>
static void run_rows(const float* rows, int64_t m, const FitConsts& c, float* k, float* t2, float* sigma, float* res,
                     float* fun, int32_t* nit, uint8_t* status) {
    for (int64_t i = 0; i < m; ++i) {
        R y[E];
        for (int e = 0; e < E; ++e) y[e] = R(rows[i * E + e]);
        const VoxelFit f = fit_voxel<R, MODEL, E>(y, c, true);
        k[i] = f.k; t2[i] = f.t2; sigma[i] = f.sigma; res[i] = f.res; fun[i] = f.fun;
        nit[i] = f.nit; status[i] = (uint8_t)f.status;
    }
}

template <typename R, int MODEL>
static int dispatch_e(int n_echo, const float* rows, int64_t m, const FitConsts& c, float* k, float* t2, float* sigma,
                      float* res, float* fun, int32_t* nit, uint8_t* status) {
    switch (n_echo) {
#define CASE(E) case E: run_rows<R, MODEL, E>(rows, m, c, k, t2, sigma, res, fun, nit, status); return 0;
        CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
        CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
        default: return T2FIT_EINVAL;
    }
}

static std::string g_err;

extern "C" const char* hostsim_last_error() { return g_err.c_str(); }

// rows: AoS [n_fit, n_echo] taken from p->echoes (mask_idx ignored: pass gathered rows).
extern "C" int hostsim_fit(const t2fit_problem* p, int use_double, float* k, float* t2, float* sigma, float* res,
                           float* fun, int32_t* nit, uint8_t* status) {
    FitConsts c;
    memset(&c, 0, sizeof(c));
    int rc = make_consts(*p, c, g_err);
    if (rc) return rc;
    const bool mono = p->model == T2FIT_MODEL_GAUSSIAN;
    if (use_double) {
        return mono ? dispatch_e<double, kMono2>(p->n_echo, p->echoes, p->n_fit, c, k, t2, sigma, res, fun, nit, status)
                    : dispatch_e<double, kFloor3>(p->n_echo, p->echoes, p->n_fit, c, k, t2, sigma, res, fun, nit, status);
    }
    return mono ? dispatch_e<float, kMono2>(p->n_echo, p->echoes, p->n_fit, c, k, t2, sigma, res, fun, nit, status)
                : dispatch_e<float, kFloor3>(p->n_echo, p->echoes, p->n_fit, c, k, t2, sigma, res, fun, nit, status);
}

// ------------------------------------------------------------------------------------------------
// reference-faithful solver (t2fit_lbfgsb.cuh) on the host: checked against scipy's L-BFGS-B itself
// ------------------------------------------------------------------------------------------------
template <int OBJ, class Run = lb::VoxelRun<OBJ>>
static int dispatch_lb(int n_echo, const float* rows, int64_t m, const lb::LbConsts& c, double* x, double* fun, int32_t* nit,
                       int32_t* nfev, uint8_t* status, int32_t* result, float* trace_f, float* trace_step, int32_t* trace_len,
                       int trace_cap) {
    for (int64_t i = 0; i < m; ++i) {
        Run run;
        memset((void*)&run, 0xA5, sizeof(run));            // stale memory of the previous voxel: nothing may depend on it
        run.start(rows + i * n_echo, c, trace_f ? trace_f + i * trace_cap : nullptr,
                  trace_step ? trace_step + i * trace_cap : nullptr, (trace_f || trace_step) ? trace_cap : 0);
        while (run.active) run.pass(c);
        const lb::LbVoxel v = run.finish();
        x[3 * i] = v.x[0]; x[3 * i + 1] = v.x[1]; x[3 * i + 2] = v.x[2];
        fun[i] = v.fun; nit[i] = v.nit; nfev[i] = v.nfev; status[i] = (uint8_t)v.status;
        // result code in the low byte; instrumentation above it: first breakdown (mask << 8, iteration << 16, pairs << 24)
        // (HOSTSIM_BRK=1 only)
        result[i] = v.result;
        if (getenv("HOSTSIM_BRK")) result[i] |= (run.s.brk_mask << 8) | ((run.s.brk_iter & 0xff) << 16) | ((run.s.brk_col & 0x7f) << 24);
        if (trace_len) trace_len[i] = v.trace_len;
    }
    return 0;
}

extern "C" int hostsim_lbfgsb(const t2fit_problem* p, double* x, double* fun, int32_t* nit, int32_t* nfev, uint8_t* status,
                              int32_t* result, float* trace_f, float* trace_step, int32_t* trace_len, int trace_cap) {
    lb::LbConsts c;
    memset(&c, 0, sizeof(c));
    int rc = make_lb_consts(*p, c, g_err);
    if (rc) return rc;
    switch (p->model) {
        case T2FIT_MODEL_GAUSSIAN: return dispatch_lb<0>(p->n_echo, p->echoes, p->n_fit, c, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
        case T2FIT_MODEL_GAUSSIAN_RICIAN: return dispatch_lb<1>(p->n_echo, p->echoes, p->n_fit, c, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
        default: return dispatch_lb<2>(p->n_echo, p->echoes, p->n_fit, c, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
    }
}

// the dense-matrix form of the same optimiser (csrc/t2fit_lbfgsb_dense.cuh)
extern "C" int hostsim_lbfgsb_dense(const t2fit_problem* p, double* x, double* fun, int32_t* nit, int32_t* nfev, uint8_t* status,
                                    int32_t* result, float* trace_f, float* trace_step, int32_t* trace_len, int trace_cap) {
    lb::LbConsts c;
    memset(&c, 0, sizeof(c));
    int rc = make_lb_consts(*p, c, g_err);
    if (rc) return rc;
    switch (p->model) {
        case T2FIT_MODEL_GAUSSIAN: return dispatch_lb<0, lb::DenseRun<0>>(p->n_echo, p->echoes, p->n_fit, c, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
        case T2FIT_MODEL_GAUSSIAN_RICIAN: return dispatch_lb<1, lb::DenseRun<1>>(p->n_echo, p->echoes, p->n_fit, c, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
        default: return dispatch_lb<2, lb::DenseRun<2>>(p->n_echo, p->echoes, p->n_fit, c, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
    }
}

extern "C" double hostsim_i0e(double x) { return lb::i0e(x); }

// lb::exp_echo (the dense kernel's exponential) on n arguments
extern "C" void hostsim_exp_echo(const double* a, double* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = lb::exp_echo(a[i]);
}

// lb::EchoDiv (the dense kernel's per-echo quotient from one reciprocal per T2) on n operand pairs: q[i] = EchoDiv(b[i])(a[i])
extern "C" void hostsim_echodiv(const double* a, const double* b, double* q, int64_t n, int te_safe) {
    for (int64_t i = 0; i < n; ++i) {
        lb::EchoDiv d;
        d.set(b[i], te_safe != 0);
        q[i] = d(a[i]);
    }
}

// ------------------------------------------------------------------------------------------------
// cooperative (lane-group-per-voxel) form of the same solver on the lane emulator: G fibers per voxel, lanes scheduled
// forwards (reverse = 0) or backwards (reverse = 1) between barriers
// ------------------------------------------------------------------------------------------------
template <int OBJ, int G>
static int dispatch_lb_coop(int n_echo, const float* rows, int64_t m, const lb::LbConsts& c, int reverse, double* x, double* fun,
                            int32_t* nit, int32_t* nfev, uint8_t* status, int32_t* result, float* trace_f, float* trace_step,
                            int32_t* trace_len, int trace_cap) {
    auto* run = new lb::CoopRun<OBJ, G>();
    emu::Lanes em(G, reverse != 0);
    for (int64_t i = 0; i < m; ++i) {
        memset((void*)run, 0xA5, sizeof(*run));          // stale shared memory of the previous voxel: nothing may depend on it
        for (int e = 0; e < n_echo; ++e) run->yraw[e] = rows[i * n_echo + e];
        float* tf = trace_f ? trace_f + i * trace_cap : nullptr;
        float* ts = trace_step ? trace_step + i * trace_cap : nullptr;
        const int cap = (trace_f || trace_step) ? trace_cap : 0;
        em.run([&](int lane) {
            lb::Group<G> grp;
            grp.lane = lane; grp.em = &em;
            run->start(grp, c, tf, ts, cap);
            while (run->active) run->pass(grp, c);
        });
        const lb::LbVoxel v = run->finish();
        x[3 * i] = v.x[0]; x[3 * i + 1] = v.x[1]; x[3 * i + 2] = v.x[2];
        fun[i] = v.fun; nit[i] = v.nit; nfev[i] = v.nfev; status[i] = (uint8_t)v.status; result[i] = v.result;
        if (trace_len) trace_len[i] = v.trace_len;
    }
    delete run;
    return 0;
}

template <int G>
static int dispatch_lb_coop_g(const t2fit_problem* p, const lb::LbConsts& c, int reverse, double* x, double* fun, int32_t* nit,
                              int32_t* nfev, uint8_t* status, int32_t* result, float* trace_f, float* trace_step,
                              int32_t* trace_len, int trace_cap) {
    switch (p->model) {
        case T2FIT_MODEL_GAUSSIAN: return dispatch_lb_coop<0, G>(p->n_echo, p->echoes, p->n_fit, c, reverse, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
        case T2FIT_MODEL_GAUSSIAN_RICIAN: return dispatch_lb_coop<1, G>(p->n_echo, p->echoes, p->n_fit, c, reverse, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
        default: return dispatch_lb_coop<2, G>(p->n_echo, p->echoes, p->n_fit, c, reverse, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
    }
}

extern "C" int hostsim_lbfgsb_coop(const t2fit_problem* p, int lanes, int reverse, double* x, double* fun, int32_t* nit,
                                   int32_t* nfev, uint8_t* status, int32_t* result, float* trace_f, float* trace_step,
                                   int32_t* trace_len, int trace_cap) {
    lb::LbConsts c;
    memset(&c, 0, sizeof(c));
    int rc = make_lb_consts(*p, c, g_err);
    if (rc) return rc;
    if (c.fd_step < 0.0) { g_err = "the cooperative solver has no analytic-gradient hook"; return T2FIT_EINVAL; }
    switch (lanes) {
        case 4: return dispatch_lb_coop_g<4>(p, c, reverse, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
        case 8: return dispatch_lb_coop_g<8>(p, c, reverse, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
        case 16: return dispatch_lb_coop_g<16>(p, c, reverse, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
        case 32: return dispatch_lb_coop_g<32>(p, c, reverse, x, fun, nit, nfev, status, result, trace_f, trace_step, trace_len, trace_cap);
        default: g_err = "lanes must be 4, 8, 16 or 32"; return T2FIT_EINVAL;
    }
}
