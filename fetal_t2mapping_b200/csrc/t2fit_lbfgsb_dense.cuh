// L-BFGS-B for n <= 3 with the limited-memory matrix held as a DENSE n x n matrix: the same algorithm as t2fit_lbfgsb.cuh
// (scipy.optimize.minimize(method="L-BFGS-B") as fit_voxel calls it, run_t2mapping.py:260-286), the same objectives in numpy
// operation order, the same forward differences, line search and stopping tests -- but every product with
//
//     B = theta I - W M W'                      (W = [Y, theta S], M the 2m x 2m middle matrix of Byrd, Lu, Nocedal, Zhu 1995)
//
// is taken with the n x n matrix itself.  Byrd, Nocedal & Schnabel (1994, theorem 2.3) show that the compact form above IS
// the result of the BFGS updates with the stored pairs (oldest first) applied to theta I, so with n <= 3:
//
//   * B is rebuilt from the <= 10 stored pairs after every accepted pair (theta changes with every pair): <= 10 rank-2
//     updates of a 3 x 3 matrix, ~ 400 flops instead of the Cholesky factorizations of a 10 x 10 and a 20 x 20 matrix;
//   * generalized Cauchy point: f' = (g + B (z - x))' d and f'' = d' B d on every segment of the projected path, directly;
//   * subspace minimisation: B_FF d_F = -(g + B (xcp - x))_F over the free set F, a <= 3 x 3 Cholesky solve, then the
//     v3.0 projection / backtracking step unchanged.
//
// State per voxel: the pairs (480 B) + ~40 scalars, against 10.9 KB of compact matrices; per iteration the optimiser core
// costs a few hundred flops, so the run time is the objective evaluations (4 per gradient).
//
// What it is NOT: bit-for-bit the trajectory of scipy's arithmetic.  In exact arithmetic both forms walk the same path; in
// floating point the compact matrices of an n <= 3 problem become numerically singular once more than n pairs are stored,
// scipy's Cholesky factorizations then break down now and then and the memory is refreshed (a steepest-descent restart),
// and with the reference's loose presets (ftol = gtol = 1e-2) the stopping point depends on the path.  The dense matrix
// never breaks down.  Parity figures of both forms: DESIGN.md section 3b, tests/test_hostsim_dense.py.
//
// Plain C++ when T2FIT_HOSTSIM is defined (tests/hostsim).
#pragma once
#include "t2fit_lbfgsb.cuh"

namespace t2fit {
namespace lb {

// The bracket of the More'-Thuente search (10 doubles, touched only while a line search goes past its first trial point) has
// a store of its own: members on the host; on the device a column of shared memory [slot][thread], so that 20 registers'
// worth of cold state neither occupies registers during the echo loop nor comes back from local memory (L2 latency) when
// dcsrch needs it.
enum LsSlot : int { kGx, kGy, kFx, kFy, kStx, kSty, kStmin, kStmax, kWidth, kWidth1, kLsSlots };
template <int STRIDE>
struct LsStore {
    double* p;
    T2_HD double& operator[](int i) { return p[i * STRIDE]; }
};
template <>
struct LsStore<0> {
    double v[kLsSlots];
    T2_HD double& operator[](int i) { return v[i]; }
};

template <int N, int LS_STRIDE = 0>
struct DenseSolver {
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
    // limited memory: the pairs (circular, `head` = oldest) and the matrix they define
    // The pairs are the ONLY dynamically indexed data of the solver, and they are deliberately NOT members: an array indexed
    // at run time inside this struct keeps the WHOLE struct in local memory (the compiler's scalar replacement gives up on an
    // aggregate as soon as one access into it has an unknown offset) -- measured: every x, g, B, ... access an LDL / STL,
    // 20 KB of DRAM traffic per voxel.  `pw` points at kPairDoubles doubles owned by the caller: s[kM][N], y[kM][N] and
    // 1 / (y's) [kM] (rebuild() needs that quotient at every iteration; it does not change while the pair is stored).
    double* pw;
    static constexpr int kPairDoubles = kM * (2 * N + 1);
    T2_HD double& ws(int pt, int i) { return pw[pt * N + i]; }
    T2_HD double& wy(int pt, int i) { return pw[kM * N + pt * N + i]; }
    T2_HD double& wiys(int pt) { return pw[2 * kM * N + pt]; }
    double B[N][N];
    double theta;
    int col, head, itail, iupdat;
    bool want_dir;              // the driver calls start_iteration() (ONE call site: the routine is inlined)
    bool updatd, stale;         // a pair was stored after the last direction; a stored pair never reached the K matrix (below)
    // line search / bookkeeping
    double fold, dnorm, dtd, gd, gdold, stp, stpmx, sbgnrm;
    int iter, ifun, iback, nfgv;
    bool brackt;
    int stage;
    LsStore<LS_STRIDE> ls;      // gx, gy, fx, fy, stx, sty, stmin, stmax, width, width1 (finit = fold, ginit = gdold, gtest = 1e-3 gdold)
    int result;

    T2_HD void projgr() {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double gi = g[i];
            if (nbd[i] != 0) {
                if (gi < 0.0) { if (nbd[i] >= 2) gi = rmax(x[i] - u[i], gi); }
                else { if (nbd[i] <= 2) gi = rmin(x[i] - l[i], gi); }
            }
            s = rmax(s, fabs(gi));
        }
        sbgnrm = s;
    }

    T2_HD void reset_memory() {
        col = 0; head = 0; theta = 1.0; iupdat = 0; updatd = false; stale = false;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) B[i][j] = (i == j) ? 1.0 : 0.0;
    }

#ifdef T2FIT_HOSTSIM
    int brk_mask = 0, brk_iter = -1, brk_col = -1;            // test instrumentation (as in Solver)
    void note_break(int bit) { if (!brk_mask) { brk_iter = iter; brk_col = col; } brk_mask |= bit; }
#else
    T2_HD void note_break(int) {}
#endif

    // B = theta I, then the BFGS update of every stored pair, oldest first
    T2_HD void rebuild() {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) B[i][j] = (i == j) ? theta : 0.0;
        // the pairs sit in local memory: the loads of pair q + 1 are issued before the update with pair q (every update needs
        // the matrix the previous one left, so nothing else hides their latency)
        double sn[N], yn[N], iysn;
        {
            const int pt = head % kM;
#pragma unroll
            for (int i = 0; i < N; ++i) { sn[i] = ws(pt, i); yn[i] = wy(pt, i); }
            iysn = wiys(pt);
        }
        T2_ROLLED for (int q = 0; q < col; ++q) {
            double s[N], y[N], bs[N];
            const double iys = iysn;
#pragma unroll
            for (int i = 0; i < N; ++i) { s[i] = sn[i]; y[i] = yn[i]; }
            if (q + 1 < col) {
                const int pt = (head + q + 1) % kM;
#pragma unroll
                for (int i = 0; i < N; ++i) { sn[i] = ws(pt, i); yn[i] = wy(pt, i); }
                iysn = wiys(pt);
            }
            double sbs = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < N; ++j) a += B[i][j] * s[j];
                bs[i] = a; sbs += a * s[i];
            }
            const double isbs = 1.0 / sbs;
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = i; j < N; ++j) {
                    const double v = B[i][j] - bs[i] * bs[j] * isbs + y[i] * y[j] * iys;
                    B[i][j] = v; B[j][i] = v;
                }
        }
    }

    // ---- generalized Cauchy point: z (= xcp), iwhere ------------------------------------------------------------
    T2_HD void cauchy() {
#pragma unroll
        for (int i = 0; i < N; ++i) z[i] = x[i];
        if (sbgnrm <= 0.0) return;
        bool bnded = true, any_unbounded = false;
        int nbreak = 0;
        double dd[N], tt[N], zx[N];                               // search direction, breakpoints, z - x of the fixed variables
        bool pending[N];
        double f1 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double neggi = -g[i];
            double tl = 0.0, tu = 0.0;
            pending[i] = false; tt[i] = 0.0; zx[i] = 0.0;
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
                // (one division site for both directions: the lanes of a warp that need either call it together)
                const bool to_lower = nbd[i] <= 2 && nbd[i] != 0 && neggi < 0.0, to_upper = !to_lower && nbd[i] >= 2 && neggi > 0.0;
                if (to_lower || to_upper) { tt[i] = ddiv(to_lower ? tl : tu, to_lower ? -neggi : neggi); pending[i] = true; ++nbreak; }
                else { any_unbounded = true; if (fabs(neggi) > 0.0) bnded = false; }
            }
        }
        if (nbreak == 0 && !any_unbounded) return;              // d is the zero vector
        const double f2_org = -theta * f1;
        double f2 = quad(dd);
        double dtm = ddiv(-f1, f2), tsum = 0.0, tj = 0.0;
        bool all_fixed = false;
        int nleft = nbreak;
        while (nleft > 0) {
            int ibp = -1;                                         // least of the remaining breakpoints (first one on ties)
            double tbest = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) if (pending[i] && (ibp < 0 || tt[i] < tbest)) { ibp = i; tbest = tt[i]; }
            const double tj0 = tj;
            tj = tbest;
            const double dt = tj - tj0;
            if (dtm < dt) break;                                  // the minimiser lies within this segment
            tsum += dt; --nleft;
#pragma unroll
            for (int i = 0; i < N; ++i) if (i == ibp) {           // (static indices: the state stays in registers)
                pending[i] = false;
                const double dibp = dd[i];
                dd[i] = 0.0;
                if (dibp > 0.0) { zx[i] = u[i] - x[i]; z[i] = u[i]; iwhere[i] = 2; }
                else { zx[i] = l[i] - x[i]; z[i] = l[i]; iwhere[i] = 1; }
            }
            if (nleft == 0 && nbreak == N) { dtm = dt; all_fixed = true; break; }
            // f' and f'' of the quadratic model on the next segment: (g + B (z - x))' d and d' B d
            double zc[N];
#pragma unroll
            for (int i = 0; i < N; ++i) zc[i] = zx[i] + tsum * dd[i];
            f1 = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double a = g[i];
#pragma unroll
                for (int j = 0; j < N; ++j) a += B[i][j] * zc[j];
                f1 += a * dd[i];
            }
            f2 = rmax(kEpsMch * f2_org, quad(dd));
            if (nleft > 0) { dtm = ddiv(-f1, f2); continue; }
            if (bnded) { f1 = 0.0; f2 = 0.0; dtm = 0.0; }
            else dtm = ddiv(-f1, f2);
            break;
        }
        if (!all_fixed) {
            if (dtm <= 0.0) dtm = 0.0;
            tsum += dtm;
#pragma unroll
            for (int i = 0; i < N; ++i) z[i] += tsum * dd[i];
        }
    }

    T2_HD double quad(const double* v) const {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double a = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j) a += B[i][j] * v[j];
            s += a * v[i];
        }
        return s;
    }

    // a / d for a pivot d > 0 (finite or +inf): a zero numerator -- every entry of the identity padding below -- is its own
    // quotient, sign included; the IEEE division would take its slow path for it (~60 instructions) to say the same
    T2_HD static double div0(double a, double d) { return a == 0.0 ? a : ddiv(a, d); }

    // ---- subspace minimisation over the free variables at the Cauchy point; false = B_FF not positive definite ----
    T2_HD bool subsm() {
        // The free set F is NOT compacted: rows / columns of the other variables are replaced by those of the identity and
        // their right-hand side by 0.  Every operation that touches them is x - 0 * y, x / 1 or sqrt(1), exact, so the
        // factorization and the solves produce bit for bit what the compacted |F| x |F| system gives, with indices known at
        // compile time (the state stays in registers).
        bool fr[N];
#pragma unroll
        for (int i = 0; i < N; ++i) fr[i] = iwhere[i] <= 0;
        // reduced gradient r = -Z'(g + B (xcp - x)) and the reduced matrix Z'BZ
        double rr_[N], A[N][N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double sgn = g[i];
#pragma unroll
            for (int j = 0; j < N; ++j) sgn += B[i][j] * (z[j] - x[j]);
            rr_[i] = fr[i] ? -sgn : 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j) A[i][j] = (fr[i] && fr[j]) ? B[i][j] : (i == j ? 1.0 : 0.0);
        }
        // Cholesky A = R'R (upper triangle), R'R d = r
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double sq = 0.0;
#pragma unroll
            for (int k = 0; k < j; ++k) {
                double tt = A[k][j];
#pragma unroll
                for (int q = 0; q < k; ++q) tt -= A[q][k] * A[q][j];
                tt = div0(tt, A[k][k]);
                A[k][j] = tt;
                sq += tt * tt;
            }
            sq = A[j][j] - sq;
            if (!(sq > 0.0)) return false;
            A[j][j] = dsqrt(sq);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double sq = rr_[j];
#pragma unroll
            for (int q = 0; q < j; ++q) sq -= A[q][j] * rr_[q];
            rr_[j] = div0(sq, A[j][j]);
        }
#pragma unroll
        for (int j = N - 1; j >= 0; --j) {
            double sq = rr_[j];
#pragma unroll
            for (int q = j + 1; q < N; ++q) sq -= A[j][q] * rr_[q];
            rr_[j] = div0(sq, A[j][j]);
        }
        // projection of the Newton point onto the box (v3.0), else backtrack along the Newton direction
        double xp[N];
#pragma unroll
        for (int i = 0; i < N; ++i) xp[i] = z[i];
        bool iword = false;
#pragma unroll
        for (int k = 0; k < N; ++k) if (fr[k]) {
            const double dk = rr_[k], xk = z[k];
            if (nbd[k] == 0) z[k] = xk + dk;
            else if (nbd[k] == 1) { z[k] = rmax(l[k], xk + dk); if (z[k] == l[k]) iword = true; }
            else if (nbd[k] == 2) { z[k] = rmin(u[k], rmax(l[k], xk + dk)); if (z[k] == l[k] || z[k] == u[k]) iword = true; }
            else { z[k] = rmin(u[k], xk + dk); if (z[k] == u[k]) iword = true; }
        }
        if (!iword) return true;
        double dd_p = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) dd_p += (z[i] - x[i]) * g[i];
        if (dd_p > 0.0) {
#pragma unroll
            for (int i = 0; i < N; ++i) z[i] = xp[i];
            double alpha = 1.0, temp1 = 1.0;
            int ibd = -1;
#pragma unroll
            for (int k = 0; k < N; ++k) if (fr[k]) {
                const double dk = rr_[k];
                if (nbd[k] != 0) {
                    const bool dn = dk < 0.0 && nbd[k] <= 2, up = !dn && dk > 0.0 && nbd[k] >= 2;
                    if (dn || up) {
                        const double temp2 = (dn ? l[k] : u[k]) - z[k];
                        if (dn ? temp2 >= 0.0 : temp2 <= 0.0) temp1 = 0.0;
                        else if (dn ? dk * alpha < temp2 : dk * alpha > temp2) temp1 = ddiv(temp2, dk);
                    }
                    if (temp1 < alpha) { alpha = temp1; ibd = k; }
                }
            }
#pragma unroll
            for (int k = 0; k < N; ++k) if (alpha < 1.0 && k == ibd) {
                const double dk = rr_[k];
                if (dk > 0.0) { z[k] = u[k]; rr_[k] = 0.0; }
                else if (dk < 0.0) { z[k] = l[k]; rr_[k] = 0.0; }
            }
#pragma unroll
            for (int k = 0; k < N; ++k) if (fr[k]) z[k] += alpha * rr_[k];
        }
        return true;
    }

    // ---- new correction pair (matupd), theta, and the matrix they define ----
    T2_HD void update_pairs(double rr, double dr) {
        updatd = true;
        ++iupdat;
        if (iupdat <= kM) { col = iupdat; itail = (head + iupdat - 1) % kM; }
        else { itail = (itail + 1) % kM; head = (head + 1) % kM; }
#pragma unroll
        double ys = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) { ws(itail, i) = d[i]; wy(itail, i) = r[i]; ys += r[i] * d[i]; }
        wiys(itail) = 1.0 / ys;
        theta = ddiv(rr, dr);
        rebuild();
    }

    // dcsrch after the first call: 0 = evaluate at the new stp, 1 = line search finished (CONVERGENCE or WARNING)
    T2_HD int dcsrch_next(double fv, double gv) {
        const double ls_gtol = 0.9, ls_xtol = 0.1, stpmin = 0.0, stpmax = stpmx;
        const double finit = fold, ginit = gdold, gtest = 1e-3 * ginit;             // as set when the search started
        const double ftest = finit + stp * gtest;
        if (stage == 1 && fv <= ftest && gv >= 0.0) stage = 2;
        double stmin = ls[kStmin], stmax = ls[kStmax];
        bool fin = false;
        if (brackt && (stp <= stmin || stp >= stmax)) fin = true;                 // rounding errors prevent progress
        if (brackt && stmax - stmin <= ls_xtol * stmax) fin = true;               // xtol test satisfied
        if (stp == stpmax && fv <= ftest && gv <= gtest) fin = true;              // stp = stpmax
        if (stp == stpmin && (fv > ftest || gv >= gtest)) fin = true;             // stp = stpmin
        if (fv <= ftest && fabs(gv) <= ls_gtol * (-ginit)) fin = true;            // strong Wolfe conditions hold
        if (fin) return 1;
        double stx = ls[kStx], sty = ls[kSty], fx = ls[kFx], fy = ls[kFy], gx = ls[kGx], gy = ls[kGy];
        // the modified function of stage 1 (psi = f - gtest stp) or f itself: ONE call of the safeguarded step
        const bool mod = stage == 1 && fv <= fx && fv > ftest;
        const double sh = mod ? gtest : 0.0;
        double fxm = mod ? fx - stx * gtest : fx, fym = mod ? fy - sty * gtest : fy;
        double gxm = gx - sh, gym = gy - sh;
        const double fm = mod ? fv - stp * gtest : fv, gm = gv - sh;
        dcstep_body(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax);
        fx = mod ? fxm + stx * gtest : fxm; fy = mod ? fym + sty * gtest : fym;
        gx = gxm + sh; gy = gym + sh;
        if (brackt) {
            const double width = ls[kWidth];
            if (fabs(sty - stx) >= 0.66 * ls[kWidth1]) stp = stx + 0.5 * (sty - stx);
            ls[kWidth1] = width;
            ls[kWidth] = fabs(sty - stx);
        }
        if (brackt) { stmin = rmin(stx, sty); stmax = rmax(stx, sty); }
        else { stmin = stp + 1.1 * (stp - stx); stmax = stp + 4.0 * (stp - stx); }
        stp = rmax(stp, stpmin);
        stp = rmin(stp, stpmax);
        if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= ls_xtol * stmax)) stp = stx;
        ls[kStx] = stx; ls[kSty] = sty; ls[kFx] = fx; ls[kFy] = fy; ls[kGx] = gx; ls[kGy] = gy;
        ls[kStmin] = stmin; ls[kStmax] = stmax;
        return 0;
    }

    T2_HD void setup(const double* x0, const double* lo, const double* hi, double ftol_, double pgtol_, int maxls_) {
        ftol = ftol_; pgtol = pgtol_; maxls = maxls_;
        cnstnd = false; boxed = true;
#pragma unroll
        for (int i = 0; i < N; ++i) {
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
        brk_mask = 0; brk_iter = -1; brk_col = -1;
#endif
        itail = 0;
        fold = dnorm = dtd = gd = gdold = stp = stpmx = sbgnrm = 0.0;
        iter = ifun = iback = nfgv = 0;
        result = kRunning; want_dir = false;
    }

    T2_HD void begin(double f0, const double* g0) {
        f = f0;
#pragma unroll
        for (int i = 0; i < N; ++i) g[i] = g0[i];
        nfgv = 1;
        projgr();
        if (sbgnrm <= pgtol) { result = kConvPg; return; }
        want_dir = true;
    }

    // new search direction and the first trial point of its line search
    T2_HD void start_iteration() {
        for (;;) {
            bool any_free = true;
            if (!cnstnd && col > 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) z[i] = x[i];
            } else {
                cauchy();
                any_free = false;
#pragma unroll
                for (int i = 0; i < N; ++i) any_free = any_free || (iwhere[i] <= 0);
            }
            // The one breakdown of the published code that is structural, not rounding: formk adds only the NEWEST pair's row
            // to its incrementally kept K matrix, and it is skipped while no variable is free at the Cauchy point.  A pair
            // stored during such an iteration never reaches K; when variables become free again the factorization of the
            // incomplete matrix fails and the memory is refreshed (a steepest-descent restart).  89 of 89 such events on the
            // golden fixtures end that way in scipy, and they are ALL the breakdowns seen there, so this form restarts too.
            if (!any_free && col != 0 && updatd) stale = true;
            if (any_free && col != 0) {
                if (stale) { note_break(2); reset_memory(); continue; }
                if (!subsm()) { note_break(4); reset_memory(); continue; }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) d[i] = z[i] - x[i];
            dtd = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) dtd += d[i] * d[i];
            dnorm = dsqrt(dtd);
            stpmx = 1e10;
            if (cnstnd) {
                if (iter == 0) stpmx = 1.0;
                else {
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const double a1 = d[i];
                        if (nbd[i] != 0) {
                            const bool dn = a1 < 0.0 && nbd[i] <= 2, up = !dn && a1 > 0.0 && nbd[i] >= 2;
                            if (dn || up) {
                                const double a2 = (dn ? l[i] : u[i]) - x[i];
                                if (dn ? a2 >= 0.0 : a2 <= 0.0) stpmx = 0.0;
                                else if (dn ? a1 * stpmx < a2 : a1 * stpmx > a2) stpmx = ddiv(a2, a1);
                            }
                        }
                    }
                }
            }
            stp = (iter == 0 && !boxed) ? rmin(ddiv(1.0, dnorm), stpmx) : 1.0;
#pragma unroll
            for (int i = 0; i < N; ++i) { t[i] = x[i]; r[i] = g[i]; }
            fold = f; ifun = 0; iback = 0;
            gd = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) gd += g[i] * d[i];
            gdold = gd;
            if (gd >= 0.0) {                                  // not a descent direction
                if (col == 0) { result = kAbnormal; return; }
                note_break(16);
                reset_memory();
                continue;
            }
            brackt = false; stage = 1;                        // dcsrch, first call (finit = fold = f, ginit = gdold = gd)
            ls[kWidth] = stpmx - 0.0; ls[kWidth1] = (stpmx - 0.0) * 2.0;
            ls[kStx] = 0.0; ls[kFx] = f; ls[kGx] = gd; ls[kSty] = 0.0; ls[kFy] = f; ls[kGy] = gd;
            ls[kStmin] = 0.0; ls[kStmax] = stp + 4.0 * stp;
            ifun = 1; ++nfgv; iback = 0;
            trial_point();
            return;
        }
    }

    T2_HD void trial_point() {
        if (stp == 1.0) {
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = z[i];
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = stp * d[i] + t[i];
        }
    }

    // f, g at the trial point x have been evaluated.  Returns true when a NEW ITERATE was accepted.
    T2_HD bool advance(double fv, const double* gv) {
        f = fv;
        gd = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) { g[i] = gv[i]; gd += g[i] * d[i]; }
        if (dcsrch_next(f, gd) == 0) {
            ++ifun; ++nfgv; iback = ifun - 1;
            if (iback >= maxls) {                             // line search gave up: back to the start of it
#pragma unroll
                for (int i = 0; i < N; ++i) { x[i] = t[i]; g[i] = r[i]; }
                f = fold;
                if (col == 0) { result = kAbnormal; return false; }
                note_break(32);
                reset_memory();
                want_dir = true;
                return false;
            }
            trial_point();
            return false;
        }
        ++iter;
        projgr();
        return true;
    }

    T2_HD void continue_after_iterate() {
        if (sbgnrm <= pgtol) { result = kConvPg; return; }
        const double ddum0 = rmax(fabs(fold), rmax(fabs(f), 1.0));
        if ((fold - f) <= ftol * ddum0) { result = kConvF; return; }
        double rr = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) { r[i] = g[i] - r[i]; rr += r[i] * r[i]; }
        double dr, ddum;
        if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
        else {
            dr = (gd - gdold) * stp;
#pragma unroll
            for (int i = 0; i < N; ++i) d[i] *= stp;
            ddum = -gdold * stp;
        }
        if (dr <= kEpsMch * ddum) updatd = false;             // skip the update (curvature condition fails)
        else update_pairs(rr, dr);
        want_dir = true;
    }
};

// ---------------------------------------------------------------------------------------------
// Where a voxel keeps the two arrays its objective loop walks: the signal row y[E] and, for each of the N + 1 points of one
// fun_and_grad call, the 8 running sums of numpy's pairwise np.sum.  Host build and generic callers: plain arrays inside the
// run.  lbfgsb_dense_kernel: shared memory, [slot][thread] (the slot index is uniform across a warp: no bank conflicts), so
// that the loop touches no local memory at all.
// ---------------------------------------------------------------------------------------------
struct DenseLocalMem {
    float y_[kMaxEcho];
    double acc_[4 * 8];
    double pairs_[kM * 7];
    static constexpr int kLsStride = 0;                      // the line-search bracket is a member of the solver
    T2_HD double* pairs() { return pairs_; }
    T2_HD double* ls_column() { return nullptr; }
    T2_HD float& y(int e) { return y_[e]; }
    T2_HD const float& y(int e) const { return y_[e]; }
    T2_HD double& acc(int p, int j) { return acc_[p * 8 + j]; }
};

template <int STRIDE>
struct DenseStridedMem {                                     // y_ / acc_ point at this thread's column
    float* y_;
    double* acc_;
    double* pairs_;                                          // DenseSolver::kPairDoubles doubles of the thread's own (local memory)
    double* ls_;                                             // [kLsSlots][STRIDE]: this thread's column of the line-search bracket
    static constexpr int kLsStride = STRIDE;
    T2_HD double* pairs() { return pairs_; }
    T2_HD double* ls_column() { return ls_; }
    T2_HD float& y(int e) { return y_[e * STRIDE]; }
    T2_HD const float& y(int e) const { return y_[e * STRIDE]; }
    T2_HD double& acc(int p, int j) { return acc_[(p * 8 + j) * STRIDE]; }
};

// what one evaluation point contributes to every echo's term (the per-call quantities of objective_term_u, hoisted: the same
// operations on the same operands, so the same bits)
template <int OBJ>
struct PointPre {
    double a, s2, ls2, ts2;
    T2_HD void set(double k, double sigma) {
        if constexpr (OBJ == 0) a = k;
        else if constexpr (OBJ == 1) { a = mul(k, k); s2 = mul(sigma, sigma); }
        else { a = k; s2 = mul(sigma, sigma); ls2 = log(s2); ts2 = mul(2.0, s2); }
    }
    // objective_term_u (t2fit_lbfgsb.cuh) with yd = (double)y, lg = (double)logf(y), y2 = (double)(y * y in float32)
    T2_HD double term(double yd, double lg, double y2, double u) const {
        if constexpr (OBJ == 0) {
            const double r = sub(yd, mul(a, u));
            return mul(r, r);
        } else if constexpr (OBJ == 1) {
            const double r = sub(yd, loop_sqrt(add(mul(a, u), s2)));
            return mul(r, r);
        } else {
            const double m = mul(a, u);
            const double x = mul(m, yd) / s2;
            const double aa = sub(lg, ls2);
            const double b = add(y2, mul(m, m)) / ts2;
            const double cc = add(fabs(x), log(i0e(x)));
            return add(sub(aa, b), cc);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// One voxel through the dense form: fit_voxel's preamble (run_t2mapping.py:237-245), scipy's fun_and_grad per pass, what
// fit_voxel returns -- the same contract as VoxelRun (t2fit_lbfgsb.cuh), force-inlined so that the state stays in registers.
// Every larger routine of the solver has exactly one call site in pass().
// ---------------------------------------------------------------------------------------------
template <int OBJ, class Mem = DenseLocalMem>
struct DenseRun {
    static constexpr int N = (OBJ == 0) ? 2 : 3;
    DenseSolver<N, Mem::kLsStride> s;
    Mem m;
    double xprev[N];
    float* trace_f;
    float* trace_step;
    int trace_cap, tl;
    int nit, nfev, status;
    bool have_prev, started, active;

    // the signal row is in m.y(0 .. E-1) already (the caller loaded it there)
    T2_HD void start(const LbConsts& c, float* tf, float* ts, int tcap) {
        const int E = c.n_echo;
        bool finite = true;
        const float y0 = m.y(0);
        float ymax = y0;
        T2_ROLLED for (int e = 0; e < E; ++e) {
            const float v = m.y(e);
            finite = finite && ((v - v) == 0.0f);
            ymax = v > ymax ? v : ymax;
        }
        if (c.norm) {                                         // float32 / float32 (:237-238)
            T2_ROLLED for (int e = 0; e < E; ++e) { const float v = m.y(e) / ymax; m.y(e) = v; finite = finite && ((v - v) == 0.0f); }
        }
        double lo[3], hi[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { lo[i] = c.lb[i]; hi[i] = c.ub[i]; }
        if (c.no_prior) lo[0] = (double)y0;                   // :243-245 (upper bound and T2 box are in c)
        status = kOk;
        if (c.no_prior && (y0 > (float)hi[0])) status = kBadBounds;   // scipy: "An upper bound is less than ..."
        else if (!finite) status = kNonFinite;
        if (OBJ == 2 && status == kOk) {                      // rician: log(signal) needs signal > 0
            T2_ROLLED for (int e = 0; e < E; ++e) if (!(m.y(e) > 0.0f)) status = kNonFinite;
        }
        s.pw = m.pairs();
        if constexpr (Mem::kLsStride != 0) s.ls.p = m.ls_column();
        s.setup(c.x0, lo, hi, c.ftol, c.pgtol, c.maxls);
        nit = 0; nfev = 0; tl = 0;
        have_prev = false; started = false;
        trace_f = tf; trace_step = ts; trace_cap = tcap;
        active = status == kOk;
    }

    T2_HD void start(const float* yraw, const LbConsts& c, float* tf, float* ts, int tcap) {
        T2_ROLLED for (int e = 0; e < c.n_echo; ++e) m.y(e) = yraw[e];
        start(c, tf, ts, tcap);
    }

    // scipy's approx_derivative step for variable i (absolute step, sign flip at a bound); the box as the solver holds it
    T2_HD double fd_h(int i, double h) const {
        const double lo = (s.nbd[i] == 1 || s.nbd[i] == 2) ? s.l[i] : -INFINITY, hi = s.nbd[i] >= 2 ? s.u[i] : INFINITY;
        const double lower = s.x[i] - lo, upper = hi - s.x[i];
        const double xt = s.x[i] + h;
        const bool violated = (xt < lo) || (xt > hi);
        const bool fitting = fabs(h) <= rmax(lower, upper);
        if (violated && fitting) h = -h;
        if (!fitting) h = (upper >= lower) ? upper : -lower;
        return h;
    }

    // f(x) and the N forward differences of one fun_and_grad call in ONE walk over the echoes.  The N + 1 points differ in
    // one coordinate each; the points x + h e_k and x + h e_sigma have the exponentials of x (same T2), so an echo costs 2
    // exponentials and N + 1 terms, and those N + 1 dependent chains (exp / sqrt / i0e) are independent of each other: the
    // loop body offers the scheduler four-way instruction-level parallelism where four separate loops offered none.
    // np.sum of every point's terms is accumulated on the fly in numpy's pairwise order (np_sum above: below 8 elements a
    // plain running sum; else 8 running sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the E % 8 last terms one
    // by one), so every value is bit for bit what objective<OBJ>() returns at that point.
    T2_HD void fun_and_grad(const LbConsts& c, double& fv, double* gv) {
        const int E = c.n_echo;
        double xt[N];
        PointPre<OBJ> pre[N + 1];
        pre[0].set(s.x[0], N == 3 ? s.x[N - 1] : 0.0);
#pragma unroll
        for (int i = 0; i < N; ++i) xt[i] = s.x[i] + fd_h(i, c.fd_step);
        pre[1].set(xt[0], N == 3 ? s.x[N - 1] : 0.0);
        pre[2] = pre[0];
        if constexpr (N == 3) pre[3].set(s.x[0], xt[N - 1]);
        EchoDiv da, db;                                       // 1 / T2 of the two T2 values of this call, once
        da.set(s.x[1], c.te_div_safe != 0);
        db.set(xt[1], c.te_div_safe != 0);
        const int nb = E >= 8 ? E - (E % 8) : 0;
        double sum[N + 1];
#pragma unroll
        for (int p = 0; p <= N; ++p) sum[p] = 0.0;
        T2_ROLLED for (int phase = 0; phase < 2; ++phase) {
            const int e1 = phase ? E : nb;
            T2_ROLLED for (int e = phase ? nb : 0; e < e1; ++e) {
                const float yf = m.y(e);
                const double yd = (double)yf, te = c.te[e];
                double lg = 0.0, y2 = 0.0;
                if constexpr (OBJ == 2) { lg = (double)logf(yf); y2 = (double)mulf(yf, yf); }
                const double ua = objective_expo<OBJ>(da, te), ub = objective_expo<OBJ>(db, te);
                double v[N + 1];
#pragma unroll
                for (int p = 0; p <= N; ++p) v[p] = pre[p].term(yd, lg, y2, p == 2 ? ub : ua);
                if (phase == 0) {
                    const int j = e & 7;
                    if (e < 8) {
#pragma unroll
                        for (int p = 0; p <= N; ++p) m.acc(p, j) = v[p];
                    } else {
#pragma unroll
                        for (int p = 0; p <= N; ++p) m.acc(p, j) = add(m.acc(p, j), v[p]);
                    }
                } else {
                    const bool first = e == 0;                // E < 8: np.sum starts from the first term itself
#pragma unroll
                    for (int p = 0; p <= N; ++p) sum[p] = first ? v[p] : add(sum[p], v[p]);
                }
            }
            if (phase == 0 && nb) {
#pragma unroll
                for (int p = 0; p <= N; ++p)
                    sum[p] = add(add(add(m.acc(p, 0), m.acc(p, 1)), add(m.acc(p, 2), m.acc(p, 3))),
                                 add(add(m.acc(p, 4), m.acc(p, 5)), add(m.acc(p, 6), m.acc(p, 7))));
            }
        }
        fv = (OBJ == 2) ? -sum[0] : sum[0] / (double)E;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double fi = (OBJ == 2) ? -sum[i + 1] : sum[i + 1] / (double)E;
            gv[i] = ddiv(fi - fv, xt[i] - s.x[i]);
        }
    }

    T2_HD void pass(const LbConsts& c) {
        double fv, gv[N];
#ifdef T2FIT_HOSTSIM
        if (c.fd_step < 0.0) {                                // test hook: analytic gradient (as VoxelRun)
            const int E = c.n_echo;
            float y[kMaxEcho];
            for (int e = 0; e < E; ++e) y[e] = m.y(e);
            fv = objective<OBJ>(s.x, y, c);
            for (int i = 0; i < N; ++i) gv[i] = 0.0;
            for (int e = 0; e < E; ++e) {
                if (OBJ == 0) {
                    const double u = exp(-c.te[e] / s.x[1]), mm = s.x[0] * u, rr = (double)y[e] - mm;
                    gv[0] += -2.0 * rr * u / E;
                    gv[1] += -2.0 * rr * mm * c.te[e] / (s.x[1] * s.x[1]) / E;
                } else if (OBJ == 1) {
                    const double u2 = exp(-2.0 * c.te[e] / s.x[1]), mm = sqrt(s.x[0] * s.x[0] * u2 + s.x[2 % N] * s.x[2 % N]);
                    const double rr = (double)y[e] - mm;
                    gv[0] += -2.0 * rr * (s.x[0] * u2 / mm) / E;
                    gv[1] += -2.0 * rr * (s.x[0] * s.x[0] * u2 * c.te[e] / (s.x[1] * s.x[1]) / mm) / E;
                    gv[2 % N] += -2.0 * rr * (s.x[2 % N] / mm) / E;
                }
            }
        } else
#endif
        fun_and_grad(c, fv, gv);
        nfev += N + 1;
        if (!started) {
            if (!(fv - fv == 0.0)) {                          // objective not finite at the start point: scipy ends ABNORMAL there
                s.f = fv; s.result = kAbnormal; active = false;
                return;
            }
            started = true;
            s.begin(fv, gv);
        } else if (s.advance(fv, gv)) {
            ++nit;                                            // scipy: n_iterations += 1; callback(x)
            if (tl < trace_cap) {
                double st = NAN;
                if (have_prev) {
                    st = 0.0;
#pragma unroll
                    for (int i = 0; i < N; ++i) st += (s.x[i] - xprev[i]) * (s.x[i] - xprev[i]);
                    st = sqrt(st);
                }
                if (trace_f) trace_f[tl] = (float)s.f;
                if (trace_step) trace_step[tl] = (float)st;
            }
            ++tl;
            if (trace_cap > 0) {                              // (the previous iterate is the trace's step size only)
#pragma unroll
                for (int i = 0; i < N; ++i) xprev[i] = s.x[i];
                have_prev = true;
            }
            if (nit >= c.maxiter) s.result = kMaxIter;
            else if (nfev > c.maxfun) s.result = kMaxFun;
            else s.continue_after_iterate();
        }
        if (s.result == kRunning && s.want_dir) { s.want_dir = false; s.start_iteration(); }
        if (s.result != kRunning) active = false;
    }

    T2_HD LbVoxel finish() const {
        LbVoxel out;
#pragma unroll
        for (int i = 0; i < N; ++i) out.x[i] = s.x[i];
        if (N < 3) out.x[2] = 0.0;
        out.fun = s.f;
        out.nit = nit;
        out.nfev = nfev;
        out.result = s.result;
        out.trace_len = tl < trace_cap ? tl : trace_cap;
        int st = status;
        if (st == kOk) {
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
