// Reference-faithful solver: the optimiser the reference actually runs, one voxel per thread, FP64.
//
// fit_voxel hands its objective to scipy.optimize.minimize(method="L-BFGS-B", jac=False, bounds=...)
// (run_t2mapping.py:260-286).  That is L-BFGS-B 3.0 (Byrd, Lu, Nocedal, Zhu; Morales & Nocedal 2011)
// with m = 10 correction pairs, driven by a 2-point forward-difference gradient with absolute step
// 1e-8 whose sign flips at an upper bound (scipy/optimize/_lbfgsb_py.py, _numdiff.py -- third-party,
// not in the reference tree).  With the reference's loose tolerances (ftol=gtol=1e-2 for the
// 3-parameter fits) the optimiser stops far from the minimiser, so matching the reference point-wise
// means following the same trajectory.  This header restates the published algorithm for n <= 3:
//
//   * generalized Cauchy point along the projected steepest-descent path        (cauchy)
//   * subspace minimisation over the free variables + projection / backtracking  (subsm, v3.0)
//   * More'-Thuente line search (MINPACK-2 dcsrch / dcstep), ftol 1e-3, gtol 0.9, xtol 0.1
//   * limited-memory BFGS pairs, theta = y'y / s'y, update skipped when s'y <= eps * (-g'd)
//   * stopping tests: max |proj g| <= pgtol,  (f_k - f_{k+1}) / max(|f_k|, |f_{k+1}|, 1) <= ftol
//
// The compact representation (S, Y, S'Y, S'S, the Cholesky factor of T and the LEL' factorization of the 2m x 2m K
// matrix, with the incrementally updated WN1 of the published code) is kept as it is although n <= 3 would allow a dense
// n x n matrix: the numerical breakdowns of those factorizations -- and the memory refreshes they trigger -- are part of
// the trajectory the reference follows (a dense BFGS matrix follows scipy on 93 % of the 3-parameter
// voxels, this form on 99 %).
//
// The objective is evaluated in the reference's operation order (numpy semantics: float32 signal,
// float64 arithmetic, numpy's pairwise/8-lane summation order, no FMA contraction), so the only
// arithmetic difference left is the last-ulp behaviour of exp/log, which also differs between the
// hosts the reference runs on.
//
// Plain C++ when T2FIT_HOSTSIM is defined (tests/hostsim, checked against scipy on GPU-less CI).
#pragma once
#include "t2fit_core.cuh"
#include "t2fit_i0e_coeffs.h"
#include <cstring>

// the optimiser core is compiled once per parameter count (not once per echo count): out-of-line on the device
#if T2_DEVICE_BUILD
#define T2_NI __device__ __noinline__
#define T2_ROLLED _Pragma("unroll 1")      // keep the optimiser core compact: it must stay resident in the instruction cache
// innermost dot products (trip count = number of stored pairs, <= 10): rolled for the 3-parameter fits, unrolled x2 for the
// 2-parameter fit.  Measured (profiles/r02_notes.md section 7): the loose 3-parameter presets stop after ~10 iterations, so most
// of these loops run 1-5 times and an unrolled body plus its remainder loop executes MORE instructions than the rolled loop
// (c3 +11 %, c5 +16 % rolled); the 2-parameter fit runs mostly at 10 pairs (c2 +3 % at x2; x4 was round 1's setting, x8 -16 %).
// N is the parameter count of the enclosing Solver<N> / CoopSolver<N, G>.
#ifdef T2_INNER_UNROLL
#define T2_PRAGMA_STR(x) _Pragma(#x)
#define T2_PRAGMA_UNROLL(n) T2_PRAGMA_STR(unroll n)
#define T2_INNER T2_PRAGMA_UNROLL(T2_INNER_UNROLL)
#else
#define T2_INNER _Pragma("unroll (N == 2 ? 2 : 1)")
#endif
#ifndef T2_LINPACK_INLINE
#define T2_LP T2_NI                        // the small LINPACK kernels (dpofa / dtrsl) as calls: less code (+3-6 %)
#else
#define T2_LP T2_HD
#endif
#else
#define T2_NI inline
#define T2_LP inline
#define T2_ROLLED
#define T2_INNER
#endif

namespace t2fit {
namespace lb {

constexpr int kM = 10;                       // maxcor (scipy default)
constexpr double kEpsMch = 2.220446049250313e-16;

enum Result : int { kRunning = 0, kConvPg = 1, kConvF = 2, kAbnormal = 3, kMaxIter = 4, kMaxFun = 5 };

// launch-invariant constants of the faithful solver (kernel parameter, constant bank)
struct LbConsts {
    double te[kMaxEcho];      // TEeffs, float64 as the reference passes them
    double x0[3];             // initial_guess (NOT yet clipped: clipping is per voxel under --no_prior)
    double lb[3], ub[3];      // param_bounds in force (lb[0] per voxel under --no_prior)
    double ftol, pgtol;       // options['ftol'] (default 2.22e-9), options['gtol'] (default 1e-5)
    double fd_step;           // eps = 1e-8
    int maxls;                // options['maxls'] (default 20; the presets say 50)
    int maxiter, maxfun;      // 15000 each
    int n_echo;
    int no_prior;
    int norm;
    int objective;            // 0 gaussian, 1 gaussian_rician, 2 rician
    int dense;                // T2FIT_SOLVER_LBFGSB_DENSE: the dense-matrix form (t2fit_lbfgsb_dense.cuh)
    int te_div_safe;          // all |TE| in {0} or (1e-100, 1e100): lb::EchoDiv may take its short form
};

// ---------------------------------------------------------------------------------------------
// arithmetic without contraction (numpy evaluates every ufunc separately)
// ---------------------------------------------------------------------------------------------
#if T2_DEVICE_BUILD
T2_HD double mul(double a, double b) { return __dmul_rn(a, b); }
T2_HD double add(double a, double b) { return __dadd_rn(a, b); }
T2_HD double sub(double a, double b) { return __dsub_rn(a, b); }
T2_HD float mulf(float a, float b) { return __fmul_rn(a, b); }
T2_HD float addf(float a, float b) { return __fadd_rn(a, b); }
#else
T2_HD double mul(double a, double b) { volatile double r = a * b; return r; }
T2_HD double add(double a, double b) { volatile double r = a + b; return r; }
T2_HD double sub(double a, double b) { volatile double r = a - b; return r; }
T2_HD float mulf(float a, float b) { volatile float r = a * b; return r; }
T2_HD float addf(float a, float b) { volatile float r = a + b; return r; }
#endif

// IEEE division / square root as out-of-line calls on the device: the inline expansions (~25 instructions each)
// would otherwise make up most of the optimiser core's code size
T2_NI double ddiv(double a, double b) { return a / b; }
T2_NI double dsqrt(double a) { return sqrt(a); }

// np.sum over a contiguous float64 vector: numpy's pairwise_sum -- plain loop below 8 elements, else 8
// running lanes combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail (n <= 128)
T2_HD double np_sum(const double* v, int n) {
    if (n < 8) {
        double s = v[0];
        for (int e = 1; e < n; ++e) s = add(s, v[e]);
        return s;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = v[j];
    const int nb = n - (n % 8);
    for (int i = 8; i < nb; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = add(r[j], v[i + j]);
    }
    double s = add(add(add(r[0], r[1]), add(r[2], r[3])), add(add(r[4], r[5]), add(r[6], r[7])));
    for (int i = nb; i < n; ++i) s = add(s, v[i]);
    return s;
}

// ---------------------------------------------------------------------------------------------
// scaled Bessel function  exp(-|x|) I0(x)  (scipy.special.i0e in the reference's rician_obj).
// Chebyshev expansions on the two classical intervals (|x| <= 8 in y = x/2 - 2; |x| > 8 in
// y = 32/x - 2 with a 1/sqrt(x) factor); coefficients fitted to float64 accuracy by
// tools/make_i0e_coeffs.py, evaluated with Clenshaw's recurrence.
// ---------------------------------------------------------------------------------------------
template <int NC>
T2_HD double chebev(double y, const double (&c)[NC]) {
    double b0 = c[0], b1 = 0.0, b2 = 0.0;
#pragma unroll
    for (int i = 1; i < NC; ++i) { b2 = b1; b1 = b0; b0 = y * b1 - b2 + c[i]; }
    return 0.5 * (b0 - b2);
}

T2_NI double i0e(double x) {
    if (x < 0) x = -x;
    if (x <= 8.0) {
        const double a[T2FIT_I0E_NA] = {T2FIT_I0E_A_LIST};
        return chebev<T2FIT_I0E_NA>(0.5 * x - 2.0, a);
    }
    const double b[T2FIT_I0E_NB] = {T2FIT_I0E_B_LIST};
    return chebev<T2FIT_I0E_NB>(32.0 / x - 2.0, b) / sqrt(x);
}

// ---------------------------------------------------------------------------------------------
// objectives (run_t2mapping.py:129-177), numpy operation order
//   y32  : the float32 signal row (normalised in float32 when norm: row / np.max(row))
// ---------------------------------------------------------------------------------------------
// The echo count is a run-time value here (one kernel per objective): the optimiser core, not the
// objective, dominates the cost of this solver.
// One echo's term of the objective at parameters p (shared by the thread-per-voxel and the cooperative kernels, which
// evaluate the echoes of one objective call on different lanes: the per-call quantities k^2, sigma^2, log(sigma^2) are
// recomputed per echo, deterministically the same values).
// the exponential of one echo's term: the only part that depends on T2 (p[1]) alone
template <int OBJ>
T2_HD double objective_expo(double t2, double te) {
    if constexpr (OBJ == 1) return exp(mul(-2.0, te) / t2);     // np.exp(-2 * t / t2)
    else return exp((-te) / t2);                                // np.exp(-t / t2)
}

// The quotient a / t2 of objective_expo for MANY numerators a (one per echo) and ONE denominator (the T2 of an evaluation
// point): the IEEE-754 quotient, bit for bit, from the correctly rounded reciprocal y = RN(1 / t2) taken once --
//     q0 = RN(a y),   r = a - t2 q0  (exact in one FMA),   q = RN(q0 + r y)
// (Markstein 1990: with y correctly rounded and q0 within an ulp or so of a / t2, the second-order correction lands on the
// correctly rounded quotient; a quotient of two doubles is never a rounding midpoint).  Checked against `a / b` on 4e8 pairs
// incl. denominators with an all-ones mantissa: no difference.  Three FP64 instructions per echo instead of the ~30 of a
// division.  Operands near the ends of the exponent range (where a y could overflow / lose bits and a / t2 would not) take
// the division itself: `safe` is decided once per point.
struct EchoDiv {
    double t2, y;
    bool safe;
    T2_HD void set(double t2_, bool te_safe) {
        t2 = t2_;
        y = 1.0 / t2_;
        safe = te_safe && fabs(t2_) > 1e-100 && fabs(t2_) < 1e100;
    }
    T2_HD double operator()(double a) const {
        if (!safe) return ddiv(a, t2);                       // out of line: the echo loop must stay small (instruction cache)
#if T2_DEVICE_BUILD
        const double q0 = __dmul_rn(a, y);
        return __fma_rn(__fma_rn(-t2, q0, a), y, q0);
#else
        const double q0 = mul(a, y);
        return fma(fma(-t2, q0, a), y, q0);
#endif
    }
};

// exp / sqrt of the dense kernel's echo loop: T2_DENSE_CALLS & 1 -> exp as a call, & 2 -> sqrt as a call (one copy of the
// ~60 / ~25 instruction expansions instead of 2 / 4 inlined ones per loop body; A/B in profiles/r02_notes.md section 10)
#ifndef T2_DENSE_CALLS
#define T2_DENSE_CALLS 0
#endif
T2_NI double dexp(double a) { return exp(a); }

// exp(a) for the arguments of the echo loop (a = -TE / T2 or -2 TE / T2, so a <= 0 and far from overflow), T2_DENSE_EXP = 1:
//     k = rint(a log2 e),  r = a - k ln 2 (two FMAs, ln 2 in two parts),  exp(r) by its Taylor polynomial of degree 13
//     (|r| <= 0.347: truncation 4e-18 relative),  2^k by the exponent field
// -- the textbook scheme the CUDA library follows as well (its polynomial is a degree-11 minimax), error <= 1 ulp like the
// library's (tests/test_hostsim_dense.py measures it against mpmath-grade references on the host build of this function).
// The idea: the library's exp, inlined twice per echo, re-materialises its 11 coefficients through uniform-register moves on
// every call (44 UMOV of the ~250 instructions of one echo); here they come from the constant bank.  MEASURED SLOWER (same
// box, c3 9.95e7 -> 9.3e7, c5 1.30e8 -> 1.26e8 fits/s: the compiler fetches the coefficients with LDCU.128 / LDC.64 in front of
// every polynomial, whose latency the four chains do not hide, and the polynomial is two FMAs longer), so it is OFF by
// default and kept as the record of the experiment.  Anything outside (-700, 700) goes to the library as a call.
#ifndef T2_DENSE_EXP
#define T2_DENSE_EXP 0
#endif
#if T2_DEVICE_BUILD
__device__ __constant__ double kExpTaylor[12] = {   // 1/13! ... 1/2!
    1.60590438368216133409e-10, 2.08767569878681001866e-09, 2.50521083854417202239e-08, 2.75573192239858882758e-07,
    2.75573192239858925110e-06, 2.48015873015873015658e-05, 1.98412698412698412526e-04, 1.38888888888888894189e-03,
    8.33333333333333321769e-03, 4.16666666666666643537e-02, 1.66666666666666657415e-01, 5.00000000000000000000e-01};
#else
static const double kExpTaylor[12] = {
    1.60590438368216133409e-10, 2.08767569878681001866e-09, 2.50521083854417202239e-08, 2.75573192239858882758e-07,
    2.75573192239858925110e-06, 2.48015873015873015658e-05, 1.98412698412698412526e-04, 1.38888888888888894189e-03,
    8.33333333333333321769e-03, 4.16666666666666643537e-02, 1.66666666666666657415e-01, 5.00000000000000000000e-01};
#endif
T2_HD double exp_echo(double a) {
    if (!(fabs(a) < 700.0)) return dexp(a);
    const double kMagic = 6755399441055744.0;                    // 1.5 * 2^52: adding it rounds to an integer (to nearest even)
#if T2_DEVICE_BUILD
    const double t = __fma_rn(a, 1.44269504088896338700e+00, kMagic);
    const int k = __double2loint(t);
    const double kd = __dsub_rn(t, kMagic);
    double r = __fma_rn(kd, -6.93147180559945286227e-01, a);
    r = __fma_rn(kd, -2.31904681384629955842e-17, r);
    double p = kExpTaylor[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) p = __fma_rn(p, r, kExpTaylor[i]);
    p = __fma_rn(p, r, 1.0);
    p = __fma_rn(p, r, 1.0);
    return __dmul_rn(p, __hiloint2double((k << 20) + 0x3ff00000, 0));
#else
    volatile double tv = fma(a, 1.44269504088896338700e+00, kMagic);
    const double t = tv;
    long long bits;
    memcpy(&bits, &t, 8);
    const int k = (int)(bits & 0xffffffffll);
    volatile double kdv = t - kMagic;
    const double kd = kdv;
    double r = fma(kd, -6.93147180559945286227e-01, a);
    r = fma(kd, -2.31904681384629955842e-17, r);
    double p = kExpTaylor[0];
    for (int i = 1; i < 12; ++i) p = fma(p, r, kExpTaylor[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const long long sb = ((long long)(k + 1023)) << 52;
    double sc;
    memcpy(&sc, &sb, 8);
    return mul(p, sc);
#endif
}
// the host build keeps libm's exp in the solver itself (the CPU parity tests are pinned to it); exp_echo is exported for its
// own accuracy test
#if T2_DEVICE_BUILD
T2_HD double loop_exp(double a) { return (T2_DENSE_CALLS & 1) ? dexp(a) : (T2_DENSE_EXP ? exp_echo(a) : exp(a)); }
#else
T2_HD double loop_exp(double a) { return exp(a); }
#endif
T2_HD double loop_sqrt(double a) { return (T2_DENSE_CALLS & 2) ? dsqrt(a) : sqrt(a); }

template <int OBJ>
T2_HD double objective_expo(const EchoDiv& d, double te) {
    if constexpr (OBJ == 1) return loop_exp(d(mul(-2.0, te)));  // np.exp(-2 * t / t2)
    else return loop_exp(d(-te));                               // np.exp(-t / t2)
}

// the term given that exponential (u): the forward differences in k and sigma reuse the exponentials of the base point
template <int OBJ>
T2_HD double objective_term_u(const double* p, float y, double u) {
    if constexpr (OBJ == 0) {                                   // gauss_obj :141-147
        const double m = mul(p[0], u);                               // k * np.exp(-t / t2)
        const double r = sub((double)y, m);
        return mul(r, r);
    } else if constexpr (OBJ == 1) {                            // gauss_rician_obj :149-155
        const double k2 = mul(p[0], p[0]), s2 = mul(p[2], p[2]);
        const double m = sqrt(add(mul(k2, u), s2));
        const double r = sub((double)y, m);
        return mul(r, r);
    } else {                                                    // rician_obj :157-177 (negative log-likelihood)
        const double k = p[0], s2 = mul(p[2], p[2]);
        const double ls2 = log(s2), ts2 = mul(2.0, s2);
        const double m = mul(k, u);
        const double x = mul(m, (double)y) / s2;
        const float lg = logf(y);                                   // np.log(float32) stays float32
        const float y2 = mulf(y, y);                                // signal**2 stays float32
        const double a = sub((double)lg, ls2);
        const double b = add((double)y2, mul(m, m)) / ts2;
        const double cc = add(fabs(x), log(i0e(x)));
        return add(sub(a, b), cc);
    }
}

template <int OBJ>
T2_HD double objective_term(const double* p, float y, double te) {
    return objective_term_u<OBJ>(p, y, objective_expo<OBJ>(p[1], te));
}

// np.sum of the E terms and the final scaling (mean for the two least-squares objectives, negated sum for the NLL)
template <int OBJ>
T2_HD double objective_reduce(const double* v, int E) {
    if constexpr (OBJ == 2) return -np_sum(v, E);
    else return np_sum(v, E) / (double)E;
}

template <int OBJ>
T2_NI double objective(const double* p, const float* y32, const LbConsts& c) {
    const int E = c.n_echo;
    double v[kMaxEcho];
    for (int e = 0; e < E; ++e) v[e] = objective_term<OBJ>(p, y32[e], c.te[e]);
    return objective_reduce<OBJ>(v, E);
}

// ---------------------------------------------------------------------------------------------
// More'-Thuente safeguarded step (MINPACK-2 dcstep)
// ---------------------------------------------------------------------------------------------
T2_HD void dcstep_body(double& stx_, double& fx_, double& dx_, double& sty_, double& fy_, double& dy_, double& stp_,
                         double fp, double dp, bool& brackt_, double stpmin, double stpmax) {
    const double sgnd = dp * ddiv(dx_, fabs(dx_));
    double stpf;
    if (fp > fx_) {
        const double th = ddiv(3.0 * (fx_ - fp), stp_ - stx_) + dx_ + dp;
        const double s = rmax(fabs(th), rmax(fabs(dx_), fabs(dp)));
        const double ths = ddiv(th, s);
        double gamma = s * dsqrt(ths * ths - ddiv(dx_, s) * ddiv(dp, s));
        if (stp_ < stx_) gamma = -gamma;
        const double p = (gamma - dx_) + th, q = ((gamma - dx_) + gamma) + dp, rr = ddiv(p, q);
        const double stpc = stx_ + rr * (stp_ - stx_);
        const double stpq = stx_ + (ddiv(dx_, ddiv(fx_ - fp, stp_ - stx_) + dx_) * 0.5) * (stp_ - stx_);
        stpf = (fabs(stpc - stx_) < fabs(stpq - stx_)) ? stpc : stpc + (stpq - stpc) * 0.5;
        brackt_ = true;
    } else if (sgnd < 0.0) {
        const double th = ddiv(3.0 * (fx_ - fp), stp_ - stx_) + dx_ + dp;
        const double s = rmax(fabs(th), rmax(fabs(dx_), fabs(dp)));
        const double ths = ddiv(th, s);
        double gamma = s * dsqrt(ths * ths - ddiv(dx_, s) * ddiv(dp, s));
        if (stp_ > stx_) gamma = -gamma;
        const double p = (gamma - dp) + th, q = ((gamma - dp) + gamma) + dx_, rr = ddiv(p, q);
        const double stpc = stp_ + rr * (stx_ - stp_);
        const double stpq = stp_ + ddiv(dp, dp - dx_) * (stx_ - stp_);
        stpf = (fabs(stpc - stp_) > fabs(stpq - stp_)) ? stpc : stpq;
        brackt_ = true;
    } else if (fabs(dp) < fabs(dx_)) {
        const double th = ddiv(3.0 * (fx_ - fp), stp_ - stx_) + dx_ + dp;
        const double s = rmax(fabs(th), rmax(fabs(dx_), fabs(dp)));
        const double ths = ddiv(th, s);
        double gamma = s * dsqrt(rmax(0.0, ths * ths - ddiv(dx_, s) * ddiv(dp, s)));
        if (stp_ > stx_) gamma = -gamma;
        const double p = (gamma - dp) + th, q = (gamma + (dx_ - dp)) + gamma, rr = ddiv(p, q);
        double stpc;
        if (rr < 0.0 && gamma != 0.0) stpc = stp_ + rr * (stx_ - stp_);
        else if (stp_ > stx_) stpc = stpmax;
        else stpc = stpmin;
        const double stpq = stp_ + ddiv(dp, dp - dx_) * (stx_ - stp_);
        if (brackt_) {
            stpf = (fabs(stpc - stp_) < fabs(stpq - stp_)) ? stpc : stpq;
            if (stp_ > stx_) stpf = rmin(stp_ + 0.66 * (sty_ - stp_), stpf);
            else stpf = rmax(stp_ + 0.66 * (sty_ - stp_), stpf);
        } else {
            stpf = (fabs(stpc - stp_) > fabs(stpq - stp_)) ? stpc : stpq;
            stpf = rmin(stpmax, stpf);
            stpf = rmax(stpmin, stpf);
        }
    } else {
        if (brackt_) {
            const double th = ddiv(3.0 * (fp - fy_), sty_ - stp_) + dy_ + dp;
            const double s = rmax(fabs(th), rmax(fabs(dy_), fabs(dp)));
            const double ths = ddiv(th, s);
            double gamma = s * dsqrt(ths * ths - ddiv(dy_, s) * ddiv(dp, s));
            if (stp_ > sty_) gamma = -gamma;
            const double p = (gamma - dp) + th, q = ((gamma - dp) + gamma) + dy_, rr = ddiv(p, q);
            stpf = stp_ + rr * (sty_ - stp_);
        } else if (stp_ > stx_) stpf = stpmax;
        else stpf = stpmin;
    }
    if (fp > fx_) {
        sty_ = stp_; fy_ = fp; dy_ = dp;
    } else {
        if (sgnd < 0.0) { sty_ = stx_; fy_ = fx_; dy_ = dx_; }
        stx_ = stp_; fx_ = fp; dx_ = dp;
    }
    stp_ = stpf;
}


// ---------------------------------------------------------------------------------------------
// the optimiser state of one voxel
// ---------------------------------------------------------------------------------------------
template <int N>
struct Solver {
    static constexpr int M2 = 2 * kM;
    // problem
    double l[N], u[N];
    int nbd[N];                 // 0 unbounded, 1 lower, 2 both, 3 upper
    double ftol, pgtol;
    int maxls;
    bool cnstnd, boxed;
    // iterate
    double x[N], g[N], f;
    double t[N], r[N], d[N], z[N];
    int iwhere[N];
    // limited-memory matrices of the compact representation (names follow the published code):
    //   ws, wy : S and Y, one correction pair per slot (circular, `head` = oldest)
    //   sy, ss : S'Y (lower triangle + diagonal used) and S'S (upper triangle used)
    //   wt     : Cholesky factor J' of theta S'S + L D^-1 L'   (upper triangle)
    //   wn1    : [Y'ZZ'Y, L_a'+R_z'; L_a+R_z, S'AA'S] kept incrementally (lower triangle)
    //   wn     : LEL' factorization of the indefinite K matrix of the subspace problem (upper triangle)
    double ws[kM][N], wy[kM][N];
    double sy[kM][kM], ss[kM][kM], wt[kM][kM];
    double wn[M2][M2], wn1[M2][M2];
    double rd[kM], rsd[kM];     // 1 / D_kk and 1 / sqrt(D_kk), D = diag(S'Y): the divisions of bmv / formt as multiplications
    double pc[M2], cc[M2];      // p and c of the Cauchy search (c = W'(xcp - x) feeds the reduced gradient)
    double theta;
    int col, head, itail, iupdat;
    bool updatd;
    int index[N], indx2[N], nfree, nenter, ileave;
    // line search / bookkeeping
    double fold, dnorm, dtd, gd, gdold, stp, stpmx, sbgnrm;
    int iter, ifun, iback, nfgv;
    // dcsrch state
    bool brackt;
    int stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
    int result;

    // ---- projected gradient norm -------------------------------------------------------
    T2_HD void projgr() {
        double s = 0.0;
        T2_ROLLED for (int i = 0; i < N; ++i) {
            double gi = g[i];
            if (nbd[i] != 0) {
                if (gi < 0.0) { if (nbd[i] >= 2) gi = rmax(x[i] - u[i], gi); }
                else { if (nbd[i] <= 2) gi = rmin(x[i] - l[i], gi); }
            }
            s = rmax(s, fabs(gi));
        }
        sbgnrm = s;
    }

    // "refresh the lbfgs memory and restart the iteration"
#ifdef T2FIT_HOSTSIM
    void reset_memory() { col = 0; head = 0; theta = 1.0; iupdat = 0; updatd = false; stale_rows = 0; }
#else
    T2_HD void reset_memory() { col = 0; head = 0; theta = 1.0; iupdat = 0; updatd = false; }
#endif
#ifdef T2FIT_HOSTSIM
    int brk_mask = 0, brk_iter = -1, brk_col = -1;            // test instrumentation: which factorization broke down first, where
    int stale_rows = 0;
    void note_break(int bit) { if (!brk_mask) { brk_iter = iter; brk_col = col; } brk_mask |= bit; }
#else
    T2_HD void note_break(int) {}
#endif

    // ---- LINPACK-style kernels on the small dense matrices ----------------------------------
    // Cholesky A = R'R of the leading nn x nn block starting at (o, o), upper triangle in place; false = not SPD
    template <int LD>
    T2_LP static bool dpofa(double (&a)[LD][LD], int o, int nn) {
        T2_ROLLED for (int j = 0; j < nn; ++j) {
            double s = 0.0;
            T2_ROLLED for (int k = 0; k < j; ++k) {
                double tt = a[o + k][o + j];
                T2_INNER for (int q = 0; q < k; ++q) tt -= a[o + q][o + k] * a[o + q][o + j];
                tt = ddiv(tt, a[o + k][o + k]);
                a[o + k][o + j] = tt;
                s += tt * tt;
            }
            s = a[o + j][o + j] - s;
            if (!(s > 0.0)) return false;
            a[o + j][o + j] = dsqrt(s);
        }
        return true;
    }
    // solve T' x = b (job 11) / T x = b (job 01), T = upper triangle of the leading nn x nn block; false = zero pivot
    template <int LD>
    T2_LP static bool dtrsl_t(const double (&a)[LD][LD], int nn, double* b) {
        T2_ROLLED for (int j = 0; j < nn; ++j) if (a[j][j] == 0.0) return false;
        T2_ROLLED for (int j = 0; j < nn; ++j) {
            double s = b[j];
            T2_INNER for (int q = 0; q < j; ++q) s -= a[q][j] * b[q];
            b[j] = ddiv(s, a[j][j]);
        }
        return true;
    }
    template <int LD>
    T2_LP static bool dtrsl_n(const double (&a)[LD][LD], int nn, double* b) {
        T2_ROLLED for (int j = 0; j < nn; ++j) if (a[j][j] == 0.0) return false;
        T2_ROLLED for (int j = nn - 1; j >= 0; --j) {
            double s = b[j];
            T2_INNER for (int q = j + 1; q < nn; ++q) s -= a[j][q] * b[q];
            b[j] = ddiv(s, a[j][j]);
        }
        return true;
    }

    // ---- product of the 2col x 2col middle matrix with v (bmv) -------------------------------
    T2_NI bool bmv(const double* v, double* p) const {
        if (col == 0) return true;
        p[col] = v[col];
        T2_ROLLED for (int i = 1; i < col; ++i) {
            double sum = 0.0;
            T2_INNER for (int k = 0; k < i; ++k) sum += sy[i][k] * v[k] * rd[k];
            p[col + i] = v[col + i] + sum;
        }
        if (!dtrsl_t<kM>(wt, col, p + col)) return false;
        T2_ROLLED for (int i = 0; i < col; ++i) p[i] = v[i] * rsd[i];
        if (!dtrsl_n<kM>(wt, col, p + col)) return false;
        T2_ROLLED for (int i = 0; i < col; ++i) p[i] = -p[i] * rsd[i];
        T2_ROLLED for (int i = 0; i < col; ++i) {
            double sum = 0.0;
            T2_INNER for (int k = i + 1; k < col; ++k) sum += sy[k][i] * p[col + k];
            p[i] += sum * rd[i];
        }
        return true;
    }

    // ---- generalized Cauchy point: z (= xcp), iwhere, c = W'(xcp - x); false = singular middle matrix ----
    T2_NI bool cauchy() {
        T2_ROLLED for (int i = 0; i < N; ++i) z[i] = x[i];
        if (sbgnrm <= 0.0) return true;
        bool bnded = true, any_unbounded = false;
        int nbreak = 0, ibkmin = 0;
        const int col2 = 2 * col;
        double bkmin = 0.0, f1 = 0.0;
        double dd[N], tt[N], wbp[M2], v[M2];
        int iorder[N];
        T2_ROLLED for (int i = 0; i < col2; ++i) pc[i] = 0.0;
        T2_ROLLED for (int i = 0; i < N; ++i) {
            const double neggi = -g[i];
            double tl = 0.0, tu = 0.0;
            if (iwhere[i] != 3 && iwhere[i] != -1) {
                if (nbd[i] <= 2) tl = x[i] - l[i];
                if (nbd[i] >= 2) tu = u[i] - x[i];
                const bool xlower = nbd[i] <= 2 && tl <= 0.0;
                const bool xupper = nbd[i] >= 2 && tu <= 0.0;
                iwhere[i] = 0;
                if (xlower) { if (neggi <= 0.0) iwhere[i] = 1; }
                else if (xupper) { if (neggi >= 0.0) iwhere[i] = 2; }
                else if (fabs(neggi) <= 0.0) iwhere[i] = -3;
            }
            if (iwhere[i] != 0 && iwhere[i] != -1) {
                dd[i] = 0.0;
            } else {
                dd[i] = neggi;
                f1 -= neggi * neggi;
                T2_ROLLED for (int j = 0; j < col; ++j) {                 // p := p - W'e_i g_i
                    const int pt = (head + j) % kM;
                    pc[j] += wy[pt][i] * neggi;
                    pc[col + j] += ws[pt][i] * neggi;
                }
                if (nbd[i] <= 2 && nbd[i] != 0 && neggi < 0.0) {
                    iorder[nbreak] = i; tt[nbreak] = ddiv(tl, -neggi);
                    if (nbreak == 0 || tt[nbreak] < bkmin) { bkmin = tt[nbreak]; ibkmin = nbreak; }
                    ++nbreak;
                } else if (nbd[i] >= 2 && neggi > 0.0) {
                    iorder[nbreak] = i; tt[nbreak] = ddiv(tu, neggi);
                    if (nbreak == 0 || tt[nbreak] < bkmin) { bkmin = tt[nbreak]; ibkmin = nbreak; }
                    ++nbreak;
                } else {
                    any_unbounded = true;
                    if (fabs(neggi) > 0.0) bnded = false;
                }
            }
        }
        if (theta != 1.0) T2_ROLLED for (int j = 0; j < col; ++j) pc[col + j] *= theta;
        if (nbreak == 0 && !any_unbounded) return true;        // d is the zero vector
        T2_ROLLED for (int i = 0; i < col2; ++i) cc[i] = 0.0;
        double f2 = -theta * f1;
        const double f2_org = f2;
        if (col > 0) {
            if (!bmv(pc, v)) return false;
            double dot = 0.0;
            T2_ROLLED for (int i = 0; i < col2; ++i) dot += v[i] * pc[i];
            f2 -= dot;
        }
        double dtm = ddiv(-f1, f2), tsum = 0.0;
        bool all_fixed = false;
        if (nbreak > 0) {
            int nleft = nbreak, it = 1;
            double tj = 0.0;
            for (;;) {
                const double tj0 = tj;
                int ibp;
                if (it == 1) {
                    tj = bkmin; ibp = iorder[ibkmin];
                } else {
                    if (it == 2 && ibkmin != nbreak - 1) { tt[ibkmin] = tt[nbreak - 1]; iorder[ibkmin] = iorder[nbreak - 1]; }
                    int jm = 0;                               // least of the remaining breakpoints -> slot nleft-1
                    T2_ROLLED for (int j = 1; j < nleft; ++j) if (tt[j] < tt[jm]) jm = j;
                    const double tv = tt[jm]; const int iv = iorder[jm];
                    tt[jm] = tt[nleft - 1]; iorder[jm] = iorder[nleft - 1];
                    tt[nleft - 1] = tv; iorder[nleft - 1] = iv;
                    tj = tv; ibp = iv;
                }
                const double dt = tj - tj0;
                if (dtm < dt) break;                          // the minimiser lies within this segment
                tsum += dt; --nleft; ++it;
                const double dibp = dd[ibp];
                dd[ibp] = 0.0;
                double zibp;
                if (dibp > 0.0) { zibp = u[ibp] - x[ibp]; z[ibp] = u[ibp]; iwhere[ibp] = 2; }
                else { zibp = l[ibp] - x[ibp]; z[ibp] = l[ibp]; iwhere[ibp] = 1; }
                if (nleft == 0 && nbreak == N) { dtm = dt; all_fixed = true; break; }
                const double dibp2 = dibp * dibp;
                f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
                f2 = f2 - theta * dibp2;
                if (col > 0) {
                    T2_ROLLED for (int i = 0; i < col2; ++i) cc[i] += dt * pc[i];
                    T2_ROLLED for (int j = 0; j < col; ++j) {
                        const int pt = (head + j) % kM;
                        wbp[j] = wy[pt][ibp];
                        wbp[col + j] = theta * ws[pt][ibp];
                    }
                    if (!bmv(wbp, v)) return false;
                    double wmc = 0.0, wmp = 0.0, wmw = 0.0;
                    T2_ROLLED for (int i = 0; i < col2; ++i) { wmc += cc[i] * v[i]; wmp += pc[i] * v[i]; wmw += wbp[i] * v[i]; }
                    T2_ROLLED for (int i = 0; i < col2; ++i) pc[i] -= dibp * wbp[i];
                    f1 += dibp * wmc;
                    f2 += 2.0 * dibp * wmp - dibp2 * wmw;
                }
                f2 = rmax(kEpsMch * f2_org, f2);
                if (nleft > 0) { dtm = ddiv(-f1, f2); continue; }
                if (bnded) { f1 = 0.0; f2 = 0.0; dtm = 0.0; }
                else dtm = ddiv(-f1, f2);
                break;
            }
        }
        if (!all_fixed) {
            if (dtm <= 0.0) dtm = 0.0;
            tsum += dtm;
            T2_ROLLED for (int i = 0; i < N; ++i) z[i] += tsum * dd[i];
        }
        if (col > 0) T2_ROLLED for (int i = 0; i < col2; ++i) cc[i] += dtm * pc[i];
        return true;
    }

    // ---- free / active index sets at the Cauchy point, entering and leaving variables (freev) ----
    T2_NI bool freev() {
        nenter = 0; ileave = N;
        if (iter > 0 && cnstnd) {
            T2_ROLLED for (int i = 0; i < nfree; ++i) { const int k = index[i]; if (iwhere[k] > 0) indx2[--ileave] = k; }
            T2_ROLLED for (int i = nfree; i < N; ++i) { const int k = index[i]; if (iwhere[k] <= 0) indx2[nenter++] = k; }
        }
        const bool wrk = (ileave < N) || (nenter > 0) || updatd;
        nfree = 0;
        int iact = N;
        T2_ROLLED for (int i = 0; i < N; ++i) {
            if (iwhere[i] <= 0) index[nfree++] = i;
            else index[--iact] = i;
        }
        return wrk;
    }

    // ---- LEL' factorization of the K matrix of the subspace problem (formk); false = not SPD ----
    T2_NI bool formk() {
        int upcl;
        if (updatd) {
            if (iupdat > kM) {                                  // shift the old part of WN1
                T2_ROLLED for (int jy = 0; jy < kM - 1; ++jy) {
                    const int js = kM + jy;
                    T2_ROLLED for (int q = 0; q < kM - 1 - jy; ++q) {
                        wn1[jy + q][jy] = wn1[jy + 1 + q][jy + 1];
                        wn1[js + q][js] = wn1[js + 1 + q][js + 1];
                    }
                    T2_ROLLED for (int q = 0; q < kM - 1; ++q) wn1[kM + q][jy] = wn1[kM + 1 + q][jy + 1];
                }
            }
            // new rows in blocks (1,1), (2,1) and (2,2)
            const int ipntr = (head + col - 1) % kM;
            const int iy = col - 1, is = kM + col - 1;
            T2_ROLLED for (int jy = 0; jy < col; ++jy) {
                const int js = kM + jy, jpntr = (head + jy) % kM;
                double t1 = 0.0, t2 = 0.0, t3 = 0.0;
                T2_ROLLED for (int k = 0; k < nfree; ++k) { const int k1 = index[k]; t1 += wy[ipntr][k1] * wy[jpntr][k1]; }
                T2_ROLLED for (int k = nfree; k < N; ++k) {
                    const int k1 = index[k];
                    t2 += ws[ipntr][k1] * ws[jpntr][k1];
                    t3 += ws[ipntr][k1] * wy[jpntr][k1];
                }
                wn1[iy][jy] = t1; wn1[is][js] = t2; wn1[is][jy] = t3;
            }
            // new column in block (2,1)
            const int jy = col - 1, jpntr = (head + col - 1) % kM;
            T2_ROLLED for (int i = 0; i < col; ++i) {
                const int is2 = kM + i, ip = (head + i) % kM;
                double t3 = 0.0;
                T2_ROLLED for (int k = 0; k < nfree; ++k) { const int k1 = index[k]; t3 += ws[ip][k1] * wy[jpntr][k1]; }
                wn1[is2][jy] = t3;
            }
            upcl = col - 1;
        } else {
            upcl = col;
        }
        // old parts of blocks (1,1) and (2,2): variables that entered / left the free set.  With an unchanged free set
        // (the usual case) every correction below is +0 - 0: the two sweeps over WN1 are skipped.
        const int upcl_old = (nenter > 0 || ileave < N) ? upcl : 0;
        T2_ROLLED for (int iy = 0; iy < upcl_old; ++iy) {
            const int is = kM + iy, ipntr = (head + iy) % kM;
            T2_ROLLED for (int jy = 0; jy <= iy; ++jy) {
                const int js = kM + jy, jpntr = (head + jy) % kM;
                double t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0;
                T2_ROLLED for (int k = 0; k < nenter; ++k) {
                    const int k1 = indx2[k];
                    t1 += wy[ipntr][k1] * wy[jpntr][k1];
                    t2 += ws[ipntr][k1] * ws[jpntr][k1];
                }
                T2_ROLLED for (int k = ileave; k < N; ++k) {
                    const int k1 = indx2[k];
                    t3 += wy[ipntr][k1] * wy[jpntr][k1];
                    t4 += ws[ipntr][k1] * ws[jpntr][k1];
                }
                wn1[iy][jy] = wn1[iy][jy] + t1 - t3;
                wn1[is][js] = wn1[is][js] - t2 + t4;
            }
        }
        // old part of block (2,1)
        T2_ROLLED for (int is0 = 0; is0 < upcl_old; ++is0) {
            const int is = kM + is0, ipntr = (head + is0) % kM;
            T2_ROLLED for (int jy = 0; jy < upcl; ++jy) {
                const int jpntr = (head + jy) % kM;
                double t1 = 0.0, t3 = 0.0;
                T2_ROLLED for (int k = 0; k < nenter; ++k) { const int k1 = indx2[k]; t1 += ws[ipntr][k1] * wy[jpntr][k1]; }
                T2_ROLLED for (int k = ileave; k < N; ++k) { const int k1 = indx2[k]; t3 += ws[ipntr][k1] * wy[jpntr][k1]; }
                if (is0 <= jy) wn1[is][jy] = wn1[is][jy] + t1 - t3;
                else wn1[is][jy] = wn1[is][jy] - t1 + t3;
            }
        }
        // upper triangle of WN = [D + Y'ZZ'Y/theta, -L_a' + R_z'; -L_a + R_z, S'AA'S theta]
#ifdef T2FIT_LB_DIRECT_WN
        T2_ROLLED for (int iy = 0; iy < col; ++iy) {
            const int is = col + iy, ip = (head + iy) % kM;
            T2_ROLLED for (int jy = 0; jy < col; ++jy) {
                const int js = col + jy, jp = (head + jy) % kM;
                double yzy = 0.0, sas = 0.0, say = 0.0, szy = 0.0;
                T2_ROLLED for (int k = 0; k < nfree; ++k) { const int k1 = index[k]; yzy += wy[ip][k1] * wy[jp][k1]; szy += ws[ip][k1] * wy[jp][k1]; }
                T2_ROLLED for (int k = nfree; k < N; ++k) { const int k1 = index[k]; sas += ws[ip][k1] * ws[jp][k1]; say += ws[ip][k1] * wy[jp][k1]; }
                if (jy <= iy) { wn[jy][iy] = ddiv(yzy, theta); wn[js][is] = sas * theta; }
                wn[jy][is] = (jy < iy) ? -say : szy;
            }
            wn[iy][iy] += sy[iy][iy];
        }
#else
        T2_ROLLED for (int iy = 0; iy < col; ++iy) {
            const int is = col + iy, is1 = kM + iy;
            T2_ROLLED for (int jy = 0; jy <= iy; ++jy) {
                const int js = col + jy, js1 = kM + jy;
                wn[jy][iy] = ddiv(wn1[iy][jy], theta);
                wn[js][is] = wn1[is1][js1] * theta;
            }
            T2_ROLLED for (int jy = 0; jy < iy; ++jy) wn[jy][is] = -wn1[is1][jy];
            T2_ROLLED for (int jy = iy; jy < col; ++jy) wn[jy][is] = wn1[is1][jy];
            wn[iy][iy] += sy[iy][iy];
        }
#endif
        // Cholesky of the (1,1) block, L^-1 (-L_a' + R_z') in the (1,2) block, then the (2,2) block
        if (!dpofa<M2>(wn, 0, col)) return false;
        const int col2 = 2 * col;
        T2_ROLLED for (int js = col; js < col2; ++js) {
            T2_ROLLED for (int j = 0; j < col; ++j) {                   // solve L x = wn(1:col, js), L' stored in the upper triangle
                double s = wn[j][js];
                T2_INNER for (int q = 0; q < j; ++q) s -= wn[q][j] * wn[q][js];
                wn[j][js] = ddiv(s, wn[j][j]);
            }
        }
        T2_ROLLED for (int is = col; is < col2; ++is)
            T2_ROLLED for (int js = is; js < col2; ++js) {
                double dot = 0.0;
                T2_INNER for (int q = 0; q < col; ++q) dot += wn[q][is] * wn[q][js];
                wn[is][js] += dot;
            }
        return dpofa<M2>(wn, col, col);
    }

    // ---- subspace minimisation over the free variables at the Cauchy point (cmprlb + subsm) ----
    T2_NI bool subsm() {
        const int col2 = 2 * col;
        double rr_[N], wv[M2];
        // reduced gradient r = -Z'(B (xcp - x) + g)
        if (!cnstnd && col > 0) {
            T2_ROLLED for (int i = 0; i < N; ++i) rr_[i] = -g[i];
        } else {
            T2_ROLLED for (int i = 0; i < nfree; ++i) { const int k = index[i]; rr_[i] = -theta * (z[k] - x[k]) - g[k]; }
            if (!bmv(cc, wv)) return false;
            T2_ROLLED for (int j = 0; j < col; ++j) {
                const int pt = (head + j) % kM;
                const double a1 = wv[j], a2 = theta * wv[col + j];
                T2_ROLLED for (int i = 0; i < nfree; ++i) { const int k = index[i]; rr_[i] += wy[pt][k] * a1 + ws[pt][k] * a2; }
            }
        }
        // wv = W'Z d, then K^-1 wv through the LEL' factors
        T2_ROLLED for (int i = 0; i < col; ++i) {
            const int pt = (head + i) % kM;
            double t1 = 0.0, t2 = 0.0;
            T2_ROLLED for (int j = 0; j < nfree; ++j) { const int k = index[j]; t1 += wy[pt][k] * rr_[j]; t2 += ws[pt][k] * rr_[j]; }
            wv[i] = t1; wv[col + i] = theta * t2;
        }
        if (!dtrsl_t<M2>(wn, col2, wv)) return false;
        T2_ROLLED for (int i = 0; i < col; ++i) wv[i] = -wv[i];
        if (!dtrsl_n<M2>(wn, col2, wv)) return false;
        // d = (1/theta) r + (1/theta^2) Z'W wv
        const double inv_theta = ddiv(1.0, theta);
        T2_ROLLED for (int jy = 0; jy < col; ++jy) {
            const int js = col + jy, pt = (head + jy) % kM;
            T2_ROLLED for (int i = 0; i < nfree; ++i) { const int k = index[i]; rr_[i] += ddiv(wy[pt][k] * wv[jy], theta) + ws[pt][k] * wv[js]; }
        }
        T2_ROLLED for (int i = 0; i < nfree; ++i) rr_[i] *= inv_theta;
        // projection of the Newton point onto the box (v3.0), else backtrack along the Newton direction
        double xp[N];
        T2_ROLLED for (int i = 0; i < N; ++i) xp[i] = z[i];
        bool iword = false;
        T2_ROLLED for (int a = 0; a < nfree; ++a) {
            const int k = index[a];
            const double dk = rr_[a], xk = z[k];
            if (nbd[k] == 0) z[k] = xk + dk;
            else if (nbd[k] == 1) { z[k] = rmax(l[k], xk + dk); if (z[k] == l[k]) iword = true; }
            else if (nbd[k] == 2) { z[k] = rmin(u[k], rmax(l[k], xk + dk)); if (z[k] == l[k] || z[k] == u[k]) iword = true; }
            else { z[k] = rmin(u[k], xk + dk); if (z[k] == u[k]) iword = true; }
        }
        if (!iword) return true;
        double dd_p = 0.0;
        T2_ROLLED for (int i = 0; i < N; ++i) dd_p += (z[i] - x[i]) * g[i];
        if (dd_p > 0.0) {
            T2_ROLLED for (int i = 0; i < N; ++i) z[i] = xp[i];
            double alpha = 1.0, temp1 = 1.0;
            int ibd = -1;
            T2_ROLLED for (int a = 0; a < nfree; ++a) {
                const int k = index[a];
                const double dk = rr_[a];
                if (nbd[k] != 0) {
                    if (dk < 0.0 && nbd[k] <= 2) {
                        const double temp2 = l[k] - z[k];
                        if (temp2 >= 0.0) temp1 = 0.0;
                        else if (dk * alpha < temp2) temp1 = ddiv(temp2, dk);
                    } else if (dk > 0.0 && nbd[k] >= 2) {
                        const double temp2 = u[k] - z[k];
                        if (temp2 <= 0.0) temp1 = 0.0;
                        else if (dk * alpha > temp2) temp1 = ddiv(temp2, dk);
                    }
                    if (temp1 < alpha) { alpha = temp1; ibd = a; }
                }
            }
            if (alpha < 1.0 && ibd >= 0) {
                const double dk = rr_[ibd];
                const int k = index[ibd];
                if (dk > 0.0) { z[k] = u[k]; rr_[ibd] = 0.0; }
                else if (dk < 0.0) { z[k] = l[k]; rr_[ibd] = 0.0; }
            }
            T2_ROLLED for (int a = 0; a < nfree; ++a) z[index[a]] += alpha * rr_[a];
        }
        return true;
    }

    // ---- new correction pair into WS, WY, S'S, S'Y (matupd) and the factor of T (formt); false = T not SPD ----
    T2_NI bool update_pairs(double rr, double dr) {
        updatd = true;
        ++iupdat;
        if (iupdat <= kM) { col = iupdat; itail = (head + iupdat - 1) % kM; }
        else { itail = (itail + 1) % kM; head = (head + 1) % kM; }
        T2_ROLLED for (int i = 0; i < N; ++i) { ws[itail][i] = d[i]; wy[itail][i] = r[i]; }
        theta = ddiv(rr, dr);
        if (iupdat > kM) {                                      // move the old information up and left
            T2_ROLLED for (int j = 0; j < col - 1; ++j) {
                T2_ROLLED for (int q = 0; q <= j; ++q) ss[q][j] = ss[q + 1][j + 1];
                T2_ROLLED for (int q = 0; q < col - 1 - j; ++q) sy[j + q][j] = sy[j + 1 + q][j + 1];
            }
        }
        T2_ROLLED for (int j = 0; j < col - 1; ++j) {                     // last row of S'Y, last column of S'S
            const int pt = (head + j) % kM;
            double a = 0.0, b = 0.0;
            T2_ROLLED for (int i = 0; i < N; ++i) { a += d[i] * wy[pt][i]; b += ws[pt][i] * d[i]; }
            sy[col - 1][j] = a;
            ss[j][col - 1] = b;
        }
        ss[col - 1][col - 1] = (stp == 1.0) ? dtd : stp * stp * dtd;
        sy[col - 1][col - 1] = dr;
        T2_ROLLED for (int k = 0; k < col; ++k) { rd[k] = ddiv(1.0, sy[k][k]); rsd[k] = ddiv(1.0, dsqrt(sy[k][k])); }
        // T = theta S'S + L D^-1 L' (upper triangle), then its Cholesky factor
        T2_ROLLED for (int j = 0; j < col; ++j) wt[0][j] = theta * ss[0][j];
        T2_ROLLED for (int i = 1; i < col; ++i)
            T2_ROLLED for (int j = i; j < col; ++j) {
                const int k1 = (i < j ? i : j);
                double ddum = 0.0;
                T2_INNER for (int k = 0; k < k1; ++k) ddum += sy[i][k] * sy[j][k] * rd[k];
                wt[i][j] = ddum + theta * ss[i][j];
            }
        return dpofa<kM>(wt, 0, col);
    }

    // ---- More'-Thuente safeguarded step (dcstep_body above), one out-of-line copy per parameter count ----
    T2_NI static void dcstep(double& stx_, double& fx_, double& dx_, double& sty_, double& fy_, double& dy_, double& stp_,
                             double fp, double dp, bool& brackt_, double stpmin, double stpmax) {
        dcstep_body(stx_, fx_, dx_, sty_, fy_, dy_, stp_, fp, dp, brackt_, stpmin, stpmax);
    }

    // dcsrch after the first call: 0 = evaluate at the new stp, 1 = line search finished (CONVERGENCE or WARNING)
    T2_NI int dcsrch_next(double fv, double gv) {
        const double ls_ftol = 1e-3, ls_gtol = 0.9, ls_xtol = 0.1, stpmin = 0.0, stpmax = stpmx;
        const double ftest = finit + stp * gtest;
        if (stage == 1 && fv <= ftest && gv >= 0.0) stage = 2;
        bool fin = false;
        if (brackt && (stp <= stmin || stp >= stmax)) fin = true;                 // rounding errors prevent progress
        if (brackt && stmax - stmin <= ls_xtol * stmax) fin = true;               // xtol test satisfied
        if (stp == stpmax && fv <= ftest && gv <= gtest) fin = true;              // stp = stpmax
        if (stp == stpmin && (fv > ftest || gv >= gtest)) fin = true;             // stp = stpmin
        if (fv <= ftest && fabs(gv) <= ls_gtol * (-ginit)) fin = true;            // strong Wolfe conditions hold
        (void)ls_ftol;
        if (fin) return 1;
        if (stage == 1 && fv <= fx && fv > ftest) {
            const double fm = fv - stp * gtest;
            double fxm = fx - stx * gtest, fym = fy - sty * gtest;
            const double gm = gv - gtest;
            double gxm = gx - gtest, gym = gy - gtest;
            dcstep(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax);
            fx = fxm + stx * gtest; fy = fym + sty * gtest;
            gx = gxm + gtest; gy = gym + gtest;
        } else {
            dcstep(stx, fx, gx, sty, fy, gy, stp, fv, gv, brackt, stmin, stmax);
        }
        if (brackt) {
            if (fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
            width1 = width;
            width = fabs(sty - stx);
        }
        if (brackt) { stmin = rmin(stx, sty); stmax = rmax(stx, sty); }
        else { stmin = stp + 1.1 * (stp - stx); stmax = stp + 4.0 * (stp - stx); }
        stp = rmax(stp, stpmin);
        stp = rmin(stp, stpmax);
        if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= ls_xtol * stmax)) stp = stx;
        return 0;
    }

    // ---- start: bounds, projected start point ------------------------------------------
    T2_NI void setup(const double* x0, const double* lo, const double* hi, double ftol_, double pgtol_, int maxls_) {
        ftol = ftol_; pgtol = pgtol_; maxls = maxls_;
        cnstnd = false; boxed = true;
        T2_ROLLED for (int i = 0; i < N; ++i) {
            const bool hl = lo[i] > -INFINITY, hu = hi[i] < INFINITY;
            nbd[i] = hl ? (hu ? 2 : 1) : (hu ? 3 : 0);
            l[i] = hl ? lo[i] : 0.0; u[i] = hu ? hi[i] : 0.0;
            double xi = x0[i];
            if (hl) xi = rmax(xi, l[i]);                              // x0 = np.clip(x0, lb, ub)
            if (hu) xi = rmin(xi, u[i]);
            x[i] = xi;
            if (nbd[i] != 2) boxed = false;
            if (nbd[i] == 0) iwhere[i] = -1;
            else { cnstnd = true; iwhere[i] = (nbd[i] == 2 && u[i] - l[i] <= 0.0) ? 3 : 0; }
        }
        reset_memory();
#ifdef T2FIT_HOSTSIM
        brk_mask = 0; brk_iter = -1; brk_col = -1; stale_rows = 0;
#endif
        itail = 0; nfree = N; nenter = 0; ileave = N;
        T2_ROLLED for (int i = 0; i < N; ++i) { index[i] = i; indx2[i] = i; }
        fold = dnorm = dtd = gd = gdold = stp = stpmx = sbgnrm = 0.0;
        iter = ifun = iback = nfgv = 0;
        result = kRunning;
        // scipy hands setulb a zero-initialised workspace for every minimize() call (wa = zeros(...)), and the algorithm
        // does read entries it never computed: formk adds only the NEWEST pair's row / column to WN1, so a pair stored
        // while formk was skipped (no free variable at the Cauchy point) leaves its row unset, and the Cauchy search
        // returns before clearing c when no variable moves.  Without this the result of a voxel depended on the voxel
        // the thread had fitted before (1 of 1500 voxels of the c3 fixture).
        // (WN1 and the Cauchy vectors are the arrays with such reads; every used entry of WS, WY, S'Y, S'S, WT and WN is
        // computed during the run before it is read)
        T2_ROLLED for (int i = 0; i < M2; ++i) { T2_ROLLED for (int j = 0; j <= i; ++j) wn1[i][j] = 0.0; pc[i] = 0.0; cc[i] = 0.0; }
    }

    // f, g at the start point have been evaluated
    T2_NI void begin(double f0, const double* g0) {
        f = f0;
        T2_ROLLED for (int i = 0; i < N; ++i) g[i] = g0[i];
        nfgv = 1;
        projgr();
        if (sbgnrm <= pgtol) { result = kConvPg; return; }
        start_iteration();
    }

    // new search direction and the first trial point of its line search (label 222 ... 666)
    T2_NI void start_iteration() {
        for (;;) {
            bool wrk;
            if (!cnstnd && col > 0) {
                T2_ROLLED for (int i = 0; i < N; ++i) z[i] = x[i];
                wrk = updatd;
            } else {
                if (!cauchy()) { note_break(1); reset_memory(); continue; }
                wrk = freev();
            }
#ifdef T2FIT_HOSTSIM
            if (nfree == 0 && col != 0 && updatd) ++stale_rows;
#endif
            if (nfree != 0 && col != 0) {
                if (wrk && !formk()) { note_break(2); reset_memory(); continue; }
                if (!subsm()) { note_break(4); reset_memory(); continue; }
            }
            T2_ROLLED for (int i = 0; i < N; ++i) d[i] = z[i] - x[i];
            // lnsrlb, first entry
            dtd = 0.0;
            T2_ROLLED for (int i = 0; i < N; ++i) dtd += d[i] * d[i];
            dnorm = dsqrt(dtd);
            stpmx = 1e10;
            if (cnstnd) {
                if (iter == 0) stpmx = 1.0;
                else {
                    T2_ROLLED for (int i = 0; i < N; ++i) {
                        const double a1 = d[i];
                        if (nbd[i] != 0) {
                            if (a1 < 0.0 && nbd[i] <= 2) {
                                const double a2 = l[i] - x[i];
                                if (a2 >= 0.0) stpmx = 0.0;
                                else if (a1 * stpmx < a2) stpmx = ddiv(a2, a1);
                            } else if (a1 > 0.0 && nbd[i] >= 2) {
                                const double a2 = u[i] - x[i];
                                if (a2 <= 0.0) stpmx = 0.0;
                                else if (a1 * stpmx > a2) stpmx = ddiv(a2, a1);
                            }
                        }
                    }
                }
            }
            stp = (iter == 0 && !boxed) ? rmin(ddiv(1.0, dnorm), stpmx) : 1.0;
            T2_ROLLED for (int i = 0; i < N; ++i) { t[i] = x[i]; r[i] = g[i]; }
            fold = f; ifun = 0; iback = 0;
            gd = 0.0;
            T2_ROLLED for (int i = 0; i < N; ++i) gd += g[i] * d[i];
            gdold = gd;
            if (gd >= 0.0) {                                  // not a descent direction
                if (col == 0) { result = kAbnormal; return; }
                note_break(16);
                reset_memory();
                continue;
            }
            // dcsrch 'START'
            brackt = false; stage = 1; finit = f; ginit = gd; gtest = 1e-3 * ginit;
            width = stpmx - 0.0; width1 = width * 2.0;
            stx = 0.0; fx = finit; gx = ginit; sty = 0.0; fy = finit; gy = ginit;
            stmin = 0.0; stmax = stp + 4.0 * stp;
            ifun = 1; ++nfgv; iback = 0;
            trial_point();
            return;
        }
    }

    T2_HD void trial_point() {
        if (stp == 1.0) { T2_ROLLED for (int i = 0; i < N; ++i) x[i] = z[i]; }
        else { T2_ROLLED for (int i = 0; i < N; ++i) x[i] = stp * d[i] + t[i]; }
    }

    // f, g at the trial point x have been evaluated.  Returns true when a NEW ITERATE was accepted
    // (scipy's callback / nit point); `result` != kRunning means the run has ended.
    T2_NI bool advance(double fv, const double* gv) {
        f = fv;
        T2_ROLLED for (int i = 0; i < N; ++i) g[i] = gv[i];
        gd = 0.0;
        T2_ROLLED for (int i = 0; i < N; ++i) gd += g[i] * d[i];
#ifdef T2FIT_LB_TRACE
        printf("  ls: iter %d ifun %d stp %.10g f %.10g gd %.6g (finit %.10g ginit %.6g) x %.8f %.8f\n", iter, ifun, stp, f, gd, finit, ginit, x[0], x[1]);
#endif
        if (dcsrch_next(f, gd) == 0) {
            ++ifun; ++nfgv; iback = ifun - 1;
            if (iback >= maxls) {                             // line search gave up: back to the start of it
                T2_ROLLED for (int i = 0; i < N; ++i) { x[i] = t[i]; g[i] = r[i]; }
                f = fold;
                if (col == 0) { result = kAbnormal; return false; }
                note_break(32);
                reset_memory();
                start_iteration();
                return false;
            }
            trial_point();
            return false;
        }
        ++iter;
        projgr();
        return true;
    }

    // after the driver has counted the iterate (nit, callback, maxiter): tests, pair update, next direction
    T2_NI void continue_after_iterate() {
        if (sbgnrm <= pgtol) { result = kConvPg; return; }
        const double ddum0 = rmax(fabs(fold), rmax(fabs(f), 1.0));
        if ((fold - f) <= ftol * ddum0) { result = kConvF; return; }
        double rr = 0.0;
        T2_ROLLED for (int i = 0; i < N; ++i) { r[i] = g[i] - r[i]; rr += r[i] * r[i]; }
        double dr, ddum;
        if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
        else {
            dr = (gd - gdold) * stp;
            T2_ROLLED for (int i = 0; i < N; ++i) d[i] *= stp;
            ddum = -gdold * stp;
        }
        if (dr <= kEpsMch * ddum) {
            updatd = false;                                   // skip the update (curvature condition fails)
        } else if (!update_pairs(rr, dr)) {
            note_break(8);
            reset_memory();                                   // T not positive definite: refresh the memory
        }
        start_iteration();
    }
};

// 2-point forward difference of scipy's approx_derivative with bounds (abs step, sign flip at a bound)
template <int N>
T2_HD double fd_step(const double* x, const double* lo, const double* hi, int i, double h) {
    const double lower = x[i] - lo[i], upper = hi[i] - x[i];
    const double xt = x[i] + h;
    const bool violated = (xt < lo[i]) || (xt > hi[i]);
    const bool fitting = fabs(h) <= rmax(lower, upper);
    if (violated && fitting) h = -h;
    if (!fitting) h = (upper >= lower) ? upper : -lower;
    return h;
}

// ---------------------------------------------------------------------------------------------
// One voxel through the reference's optimiser.  Mirrors fit_voxel (run_t2mapping.py:237-312):
// normalisation, per-voxel bounds, minimize(...), and what it returns (x, success, nit, fun).
// trace_f / trace_step (may be null): the callback's f_val / step_size per iteration (:180-234), up to
// trace_cap entries.
// ---------------------------------------------------------------------------------------------
struct LbVoxel {
    double x[3];
    double fun;
    int nit, nfev, status, result, trace_len;
};

// One voxel as a resumable run: start() = fit_voxel's preamble (:237-245), pass() = one fun_and_grad(x) of scipy's
// ScalarFunction (f plus N forward differences) followed by the optimiser's reaction to it, finish() = what
// fit_voxel returns.  The device kernel drives many runs per thread from a work queue; the host simulation one.
template <int OBJ>
struct VoxelRun {
    static constexpr int N = (OBJ == 0) ? 2 : 3;
    Solver<N> s;
    float y[kMaxEcho];
    double lo[3], hi[3];
    double xprev[N];
    float* trace_f;
    float* trace_step;
    int trace_cap, tl;
    int nit, nfev, status;
    bool have_prev, started, active;

    T2_NI void start(const float* yraw, const LbConsts& c, float* tf, float* ts, int tcap) {
        const int E = c.n_echo;
        bool finite = true;
        float ymax = yraw[0];
        T2_ROLLED for (int e = 0; e < E; ++e) {
            y[e] = yraw[e];
            finite = finite && ((yraw[e] - yraw[e]) == 0.0f);
            ymax = yraw[e] > ymax ? yraw[e] : ymax;
        }
        if (c.norm) {                                         // float32 / float32 (:237-238)
            T2_ROLLED for (int e = 0; e < E; ++e) { y[e] = yraw[e] / ymax; finite = finite && ((y[e] - y[e]) == 0.0f); }
        }
        for (int i = 0; i < 3; ++i) { lo[i] = c.lb[i]; hi[i] = c.ub[i]; }
        if (c.no_prior) lo[0] = (double)yraw[0];              // :243-245 (upper bound and T2 box are in c)
        status = kOk;
        if (c.no_prior && (yraw[0] > (float)hi[0])) status = kBadBounds;   // scipy: "An upper bound is less than ..."
        else if (!finite) status = kNonFinite;
        if (OBJ == 2 && status == kOk) {                      // rician: log(signal) needs signal > 0
            T2_ROLLED for (int e = 0; e < E; ++e) if (!(y[e] > 0.0f)) status = kNonFinite;
        }
        s.setup(c.x0, lo, hi, c.ftol, c.pgtol, c.maxls);
        nit = 0; nfev = 0; tl = 0;
        have_prev = false; started = false;
        trace_f = tf; trace_step = ts; trace_cap = tcap;
        active = status == kOk;
    }

    T2_NI void pass(const LbConsts& c) {
        const int E = c.n_echo;
        double fv = 0.0, gv[N];
#ifdef T2FIT_HOSTSIM
        if (c.fd_step < 0.0) {                                // test hook: analytic gradient (validates the optimiser core
            fv = objective<OBJ>(s.x, y, c);                   // against scipy with jac=True, free of finite-difference noise)
            for (int i = 0; i < N; ++i) gv[i] = 0.0;
            for (int e = 0; e < E; ++e) {
                if (OBJ == 0) {
                    const double u = exp(-c.te[e] / s.x[1]), m = s.x[0] * u, rr = (double)y[e] - m;
                    gv[0] += -2.0 * rr * u / E;
                    gv[1] += -2.0 * rr * m * c.te[e] / (s.x[1] * s.x[1]) / E;
                } else if (OBJ == 1) {
                    const double u2 = exp(-2.0 * c.te[e] / s.x[1]), m = sqrt(s.x[0] * s.x[0] * u2 + s.x[2] * s.x[2]);
                    const double rr = (double)y[e] - m;
                    gv[0] += -2.0 * rr * (s.x[0] * u2 / m) / E;
                    gv[1] += -2.0 * rr * (s.x[0] * s.x[0] * u2 * c.te[e] / (s.x[1] * s.x[1]) / m) / E;
                    gv[2 % N] += -2.0 * rr * (s.x[2 % N] / m) / E;
                }
            }
        } else
#endif
        {
            (void)E;
            fv = objective<OBJ>(s.x, y, c);
            double xt[N];
            T2_ROLLED for (int i = 0; i < N; ++i) {
#pragma unroll
                for (int j = 0; j < N; ++j) xt[j] = s.x[j];
                const double h = fd_step<N>(s.x, lo, hi, i, c.fd_step);
                xt[i] = s.x[i] + h;
                const double dx = xt[i] - s.x[i];
                gv[i] = ddiv(objective<OBJ>(xt, y, c) - fv, dx);
            }
        }
        nfev += N + 1;
        if (!started && !(fv - fv == 0.0)) {                  // objective not finite at the start point: scipy ends ABNORMAL there
            s.f = fv; s.result = kAbnormal; active = false;
            return;
        }
        if (!started) {
            started = true;
            s.begin(fv, gv);
        } else if (s.advance(fv, gv)) {
            ++nit;                                            // scipy: n_iterations += 1; callback(x)
            if (tl < trace_cap) {
                double st = NAN;
                if (have_prev) {
                    st = 0.0;
                    for (int i = 0; i < N; ++i) st += (s.x[i] - xprev[i]) * (s.x[i] - xprev[i]);
                    st = sqrt(st);
                }
                if (trace_f) trace_f[tl] = (float)s.f;
                if (trace_step) trace_step[tl] = (float)st;
            }
            ++tl;
            for (int i = 0; i < N; ++i) xprev[i] = s.x[i];
            have_prev = true;
            if (nit >= c.maxiter) s.result = kMaxIter;
            else if (nfev > c.maxfun) s.result = kMaxFun;
            else s.continue_after_iterate();
        }
        if (s.result != kRunning) active = false;
    }

    T2_HD LbVoxel finish() const {
        LbVoxel out;
        for (int i = 0; i < N; ++i) out.x[i] = s.x[i];
        if (N < 3) out.x[2] = 0.0;
        out.fun = s.f;
        out.nit = nit;
        out.nfev = nfev;
        out.result = s.result;
        out.trace_len = tl < trace_cap ? tl : trace_cap;
        int st = status;
        if (st == kOk) {
            // success False: ABNORMAL (x = start of the failed line search), maxiter or maxfun
            if (s.result == kAbnormal || s.result == kMaxIter || s.result == kMaxFun) st = kNotConverged;
        } else {
            out.fun = NAN; out.nit = 0;
            if (st == kBadBounds) { out.x[0] = out.x[1] = out.x[2] = NAN; }
        }
        out.status = st;
        return out;
    }
};

}  // namespace lb
}  // namespace t2fit
