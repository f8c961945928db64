// Reference-faithful solver, cooperative form: ONE VOXEL PER LANE GROUP (G = 8 / 16 / 32 lanes), optimiser state in
// SHARED MEMORY.
//
// Same algorithm and the same floating-point operations in the same order as t2fit_lbfgsb.cuh (the thread-per-voxel
// restatement of scipy's L-BFGS-B that is validated against scipy itself): every expression below is the expression of
// the serial solver, only WHO evaluates it differs.  The thread-per-voxel kernel keeps ~7-11 KB of compact L-BFGS
// matrices per thread in local memory; with ~900 resident threads per SM that state lives in DRAM (profiles/r01: 980 KB
// of DRAM traffic per voxel, 19 % issue-active, 10 of 32 lanes busy).  Here the state of a voxel is ~7.6 KB of shared
// memory owned by a group of G lanes:
//   * scalar control (line search, Cauchy breakpoints, convergence tests, bookkeeping) runs on the group's MASTER lane
//     exactly as written in the serial solver;
//   * the dense kernels -- the three Cholesky factorizations (T, and the two diagonal blocks of the LEL' factorization of
//     K), the triangular solves, the assembly of T / WN, the inner products with the correction pairs, the objective's
//     echo terms -- run across the lanes: rows of a factorization are processed one after the other (their pivots are a
//     sequential chain), the elements of a row in parallel; forward substitution keeps the running right-hand side in
//     registers and hands each solved unknown round with a shuffle (the accumulation order per element is the serial
//     one); back substitution needs the unknowns in DESCENDING order but the serial code sums them in ASCENDING order, so
//     it stays a sequential chain on the master lane.
//   * lanes agree on control flow through one control word in shared memory (written by the master, read after a group
//     barrier); all barriers are __syncwarp(group mask) -- no block-level synchronisation anywhere.
// Matrices are packed: WN (upper triangle) and WN1 (lower triangle) share one 20 x 21 array, S'Y (lower) and S'S (upper)
// one 10 x 11 array.
//
// Host build (T2FIT_HOSTSIM): the same code runs on a fiber-based lane emulator (tests/hostsim/lane_emu.h), forwards and
// backwards over the lanes, and is compared bit for bit with the serial solver (tests/test_hostsim_coop.py).
#pragma once
#include <stddef.h>
#include "t2fit_lbfgsb.cuh"

#if !T2_DEVICE_BUILD
#include "../../tests/hostsim/lane_emu.h"
#endif

namespace t2fit {
namespace lb {

// ---------------------------------------------------------------------------------------------
// the lanes of one voxel
// ---------------------------------------------------------------------------------------------
template <int G>
struct Group {
    int lane;                 // 0 .. G-1; lane 0 is the master
#if T2_DEVICE_BUILD
    unsigned mask;            // the G lanes of this group within their warp
    T2_HD void sync() const { __syncwarp(mask); }
    T2_HD double shfl(double v, int src) const { return __shfl_sync(mask, v, src, G); }
    T2_HD int shfl(int v, int src) const { return __shfl_sync(mask, v, src, G); }
    T2_HD long long shfl(long long v, int src) const { return __shfl_sync(mask, v, src, G); }
    T2_HD bool any(bool p) const { return __any_sync(mask, p) != 0; }
#else
    emu::Lanes* em;
    void sync() const { em->barrier(lane); }
    double shfl(double v, int src) const { return em->shfl(lane, v, src); }
    int shfl(int v, int src) const { return (int)em->shfl(lane, (double)v, src); }
    long long shfl(long long v, int src) const { return (long long)em->shfl(lane, (double)v, src); }
    bool any(bool p) const { return em->any(lane, p); }
#endif
    T2_HD bool master() const { return lane == 0; }
};

// rank t of the column-major enumeration of an upper triangle -> (a, b), a <= b:  t = b (b + 1) / 2 + a
T2_HD void tri_unrank(int t, int& a, int& b) {
    int bb = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    if ((bb + 1) * (bb + 2) / 2 <= t) ++bb;
    if (bb * (bb + 1) / 2 > t) --bb;
    b = bb;
    a = t - bb * (bb + 1) / 2;
}

// control words (master -> group)
enum Ctl : int { kCtlNone = 0, kCtlDone, kCtlStartIter, kCtlTrial, kCtlRestart, kCtlAccepted, kCtlUpdate, kCtlSkipUpdate,
                 kCtlContinue, kCtlReturn, kCtlBreak, kCtlBmv };

// ---------------------------------------------------------------------------------------------
// optimiser state of one voxel (shared memory) + the cooperative solver
// ---------------------------------------------------------------------------------------------
template <int N, int G>
struct CoopSolver {
    static constexpr int M2 = 2 * kM;
    static constexpr int R2 = (M2 + G - 1) / G;      // elements of a 2m-vector per lane
    static constexpr int R1 = (kM + G - 1) / G;      // elements of an m-vector per lane
    // packed matrices
    double wnn[M2][M2 + 1];        // WN(i,j), i <= j -> wnn[i][j+1];  WN1(i,j), i >= j -> wnn[i][j]
    double sys[kM][kM + 1];        // SY(i,j), i >= j -> sys[i][j];    SS(i,j), i <= j -> sys[i][j+1]
    double wt[kM][kM];             // upper triangle
    double ws[kM][N], wy[kM][N];
    double rd[kM], rsd[kM];
    // scratch vectors of the direction phase; the objective's echo terms of the evaluation phase share the space
    union Scratch {
        struct Vec { double pc[M2], cc[M2], v[M2], wbp[M2], wv[M2]; } vec;
        double fterm[(N + 1) * kMaxEcho];
    } scr;
    // problem
    double l[N], u[N];
    int nbd[N];
    double ftol, pgtol;
    int maxls;
    bool cnstnd, boxed;
    // iterate
    double x[N], g[N], f;
    double t[N], r[N], d[N], z[N], xp[N], rr_[N];
    int iwhere[N];
    double theta;
    int col, head, itail, iupdat;
    bool updatd;
    int index[N], indx2[N], nfree, nenter, ileave;
    double fold, dnorm, dtd, gd, gdold, stp, stpmx, sbgnrm;
    int iter, ifun, iback, nfgv;
    bool brackt;
    int stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
    int result;
    // master -> group.  Discipline: a block of master-lane code is preceded by a group barrier (every lane has finished
    // reading what the master is about to change) and followed by one (its writes are visible); a control word is read by
    // all lanes between two barriers (bcast_ctl) so the master cannot overwrite it before a slower lane has looked.
    int ctl, ctl2;
    bool wrk;
    double upd_rr, upd_dr;
    double hand_a;                // a scalar handed from the master to the lanes (cauchy)
    int hand_i;
    double dd_c[N];               // cauchy: search direction components (0 for fixed variables)
    int contrib_c[N];             // cauchy: variable i enters p = W'd

    T2_HD int bcast_ctl(const Group<G> grp) { grp.sync(); const int c = ctl; grp.sync(); return c; }

    T2_HD double& WN(int i, int j) { return wnn[i][j + 1]; }
    T2_HD double& WN1(int i, int j) { return wnn[i][j]; }
    T2_HD double& SY(int i, int j) { return sys[i][j]; }
    T2_HD double& SS(int i, int j) { return sys[i][j + 1]; }
    T2_HD double SYc(int i, int j) const { return sys[i][j]; }

    // =========================================================================================
    // master-lane routines: the serial solver's code, verbatim
    // =========================================================================================
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

    T2_HD void reset_memory() { col = 0; head = 0; theta = 1.0; iupdat = 0; updatd = false; }

    T2_NI bool freev() {
        nenter = 0; ileave = N;
        if (iter > 0 && cnstnd) {
            T2_ROLLED for (int i = 0; i < nfree; ++i) { const int k = index[i]; if (iwhere[k] > 0) indx2[--ileave] = k; }
            T2_ROLLED for (int i = nfree; i < N; ++i) { const int k = index[i]; if (iwhere[k] <= 0) indx2[nenter++] = k; }
        }
        const bool w = (ileave < N) || (nenter > 0) || updatd;
        nfree = 0;
        int iact = N;
        T2_ROLLED for (int i = 0; i < N; ++i) {
            if (iwhere[i] <= 0) index[nfree++] = i;
            else index[--iact] = i;
        }
        return w;
    }

    // back substitution T x = b on the master lane (the serial code sums the solved unknowns in ascending order)
    template <bool WNM>
    T2_HD double tel(int i, int j) const { return WNM ? wnn[i][j + 1] : wt[i][j]; }
    template <bool WNM>
    T2_HD void trsl_n_master(int nn, double* b) {
        T2_ROLLED for (int j = nn - 1; j >= 0; --j) {
            double s = b[j];
            T2_INNER for (int q = j + 1; q < nn; ++q) s -= tel<WNM>(j, q) * b[q];
            b[j] = ddiv(s, tel<WNM>(j, j));
        }
    }

    // =========================================================================================
    // cooperative dense kernels
    // =========================================================================================
    // any zero pivot on the diagonal of the triangular factor (LINPACK dtrsl's info)
    template <bool WNM>
    T2_HD bool zero_pivot(const Group<G> grp, int nn) const {
        bool zp = false;
        for (int j = grp.lane; j < nn; j += G) zp = zp || (tel<WNM>(j, j) == 0.0);
        return grp.any(zp);
    }

    // forward substitution T' x = b (T upper): x_j = (b_j - sum_{q<j} T[q][j] x_q) / T[j][j].  Lane l keeps the running
    // right-hand sides of the unknowns l, l + G, ... in registers; unknown q is solved by its owner and handed round.
    // b must be complete and synchronised; the result is visible to the group on return.
    template <bool WNM, int R>
    T2_HD void trsl_t(const Group<G> grp, int nn, double* b) {
        double s[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) { const int j = grp.lane + rr * G; s[rr] = j < nn ? b[j] : 0.0; }
        T2_ROLLED for (int q = 0; q < nn; ++q) {
            const int rq = q / G, owner = q % G;
            double sq = s[0];
#pragma unroll
            for (int rr = 1; rr < R; ++rr) if (rq == rr) sq = s[rr];
            double xq = ddiv(sq, tel<WNM>(q, q));
            xq = grp.shfl(xq, owner);
            if (grp.lane == owner) b[q] = xq;
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                const int j = grp.lane + rr * G;
                if (j > q && j < nn) s[rr] -= tel<WNM>(q, j) * xq;
            }
        }
        grp.sync();
    }

    // rows o .. o+nn-1 of an upper Cholesky-type factorization in place: row k = columns k+1 .. jend-1 (jend may extend
    // past the diagonal block: the (1,2) block of WN is the same recurrence), inner products over rows o .. k-1 only.
    // The pivot of a row is evaluated by every lane (same loads as its own element).  false = pivot not positive.
    template <bool WNM>
    T2_HD double& tref(int i, int j) { return WNM ? wnn[i][j + 1] : wt[i][j]; }
    template <bool WNM, int R>
    T2_HD bool chol_rows(const Group<G> grp, int o, int nn, int jend) {
        T2_ROLLED for (int k = o; k < o + nn; ++k) {
            double tt[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { const int j = k + 1 + grp.lane + rr * G; tt[rr] = j < jend ? tel<WNM>(k, j) : 0.0; }
            double s = 0.0;
            T2_ROLLED for (int q = o; q < k; ++q) {
                const double aqk = tel<WNM>(q, k);
                s += aqk * aqk;
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const int j = k + 1 + grp.lane + rr * G;
                    if (j < jend) tt[rr] -= aqk * tel<WNM>(q, j);
                }
            }
            s = tel<WNM>(k, k) - s;
            if (!(s > 0.0)) return false;            // same value on every lane
            const double piv = dsqrt(s);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                const int j = k + 1 + grp.lane + rr * G;
                if (j < jend) tref<WNM>(k, j) = ddiv(tt[rr], piv);
            }
            grp.sync();                              // row k complete (and every lane has read the old diagonal)
            if (grp.master()) tref<WNM>(k, k) = piv;
        }
        grp.sync();
        return true;
    }

    // ---- product of the 2col x 2col middle matrix with v (bmv); v complete and synchronised --------------
    T2_NI bool bmv(const Group<G> grp, const double* v, double* p) {
        const int cl = col;
        if (cl == 0) return true;
        if (zero_pivot<false>(grp, cl)) return false;
        for (int i = grp.lane; i < cl; i += G) {
            double sum = 0.0;
            T2_INNER for (int k = 0; k < i; ++k) sum += SY(i, k) * v[k] * rd[k];
            p[cl + i] = (i == 0) ? v[cl] : v[cl + i] + sum;
        }
        grp.sync();
        trsl_t<false, R1>(grp, cl, p + cl);
        for (int i = grp.lane; i < cl; i += G) p[i] = v[i] * rsd[i];
        if (grp.master()) trsl_n_master<false>(cl, p + cl);
        grp.sync();
        for (int i = grp.lane; i < cl; i += G) {
            const double pi = -p[i] * rsd[i];
            double sum = 0.0;
            T2_INNER for (int k = i + 1; k < cl; ++k) sum += SY(k, i) * p[cl + k];
            p[i] = pi + sum * rd[i];
        }
        grp.sync();
        return true;
    }

    // ---- generalized Cauchy point ------------------------------------------------------------------
    // Scalar logic on the master lane (its locals live across the cooperative calls); p = W'd, the vector updates and
    // bmv across the lanes.  Entered with the group synchronised.
    T2_NI bool cauchy(const Group<G> grp) {
        double* pc = scr.vec.pc; double* cc = scr.vec.cc; double* v = scr.vec.v; double* wbp = scr.vec.wbp;
        bool bnded = true, any_unbounded = false, all_fixed = false;
        int nbreak = 0, ibkmin = 0, nleft = 0, it = 1, ibp = 0;
        double bkmin = 0.0, f1 = 0.0, f2 = 0.0, f2_org = 0.0, dtm = 0.0, tsum = 0.0, tj = 0.0, dibp = 0.0, dibp2 = 0.0, dt = 0.0;
        double tt[N];
        int iorder[N];
        const int cl = col, col2 = 2 * cl, hd = head;           // stable during the call
        const double th = theta;
        grp.sync();
        if (grp.master()) {
            T2_ROLLED for (int i = 0; i < N; ++i) z[i] = x[i];
            ctl = (sbgnrm <= 0.0) ? kCtlReturn : kCtlContinue;
            ctl2 = kCtlContinue;
            if (ctl == kCtlContinue) {
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
                        dd_c[i] = 0.0;
                        contrib_c[i] = 0;
                    } else {
                        dd_c[i] = neggi;
                        contrib_c[i] = 1;
                        f1 -= neggi * neggi;
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
                if (nbreak == 0 && !any_unbounded) ctl2 = kCtlReturn;          // d is the zero vector
            }
        }
        grp.sync();
        const int c1 = ctl, c2 = ctl2;        // (not written again before the next barrier-separated master block)
        if (c1 == kCtlReturn) return true;
        // p = W'd: the serial code adds variable after variable, the same order per element here
        for (int j = grp.lane; j < cl; j += G) {
            const int pt = (hd + j) % kM;
            double a = 0.0, b = 0.0;
            T2_ROLLED for (int i = 0; i < N; ++i) {
                if (contrib_c[i]) { a += wy[pt][i] * dd_c[i]; b += ws[pt][i] * dd_c[i]; }
            }
            if (th != 1.0) b *= th;
            pc[j] = a; pc[cl + j] = b;
        }
        if (c2 == kCtlReturn) { grp.sync(); return true; }
        for (int i = grp.lane; i < col2; i += G) cc[i] = 0.0;
        grp.sync();
        if (grp.master()) { f2 = -th * f1; f2_org = f2; }
        if (cl > 0) {
            if (!bmv(grp, pc, v)) return false;
            if (grp.master()) {
                double dot = 0.0;
                T2_ROLLED for (int i = 0; i < col2; ++i) dot += v[i] * pc[i];
                f2 -= dot;
            }
        }
        if (grp.master()) { dtm = ddiv(-f1, f2); nleft = nbreak; }
        bool first = true;
        for (;;) {
            grp.sync();
            if (grp.master()) {
                if (first && nbreak == 0) ctl = kCtlBreak;
                else {
                    const double tj0 = tj;
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
                    dt = tj - tj0;
                    if (dtm < dt) ctl = kCtlBreak;                // the minimiser lies within this segment
                    else {
                        tsum += dt; --nleft; ++it;
                        dibp = dd_c[ibp];
                        dd_c[ibp] = 0.0;
                        double zibp;
                        if (dibp > 0.0) { zibp = u[ibp] - x[ibp]; z[ibp] = u[ibp]; iwhere[ibp] = 2; }
                        else { zibp = l[ibp] - x[ibp]; z[ibp] = l[ibp]; iwhere[ibp] = 1; }
                        if (nleft == 0 && nbreak == N) { dtm = dt; all_fixed = true; ctl = kCtlBreak; }
                        else {
                            dibp2 = dibp * dibp;
                            f1 = f1 + dt * f2 + dibp2 - th * dibp * zibp;
                            f2 = f2 - th * dibp2;
                            hand_a = dt; hand_i = ibp;
                            ctl = cl > 0 ? kCtlBmv : kCtlContinue;
                        }
                    }
                }
            }
            first = false;
            const int c = bcast_ctl(grp);
            if (c == kCtlBreak) break;
            if (c == kCtlBmv) {
                const double dtl = hand_a;
                const int ib = hand_i;
                for (int i = grp.lane; i < col2; i += G) cc[i] += dtl * pc[i];
                for (int j = grp.lane; j < cl; j += G) {
                    const int pt = (hd + j) % kM;
                    wbp[j] = wy[pt][ib];
                    wbp[cl + j] = th * ws[pt][ib];
                }
                grp.sync();
                if (!bmv(grp, wbp, v)) return false;
                if (grp.master()) {
                    double wmc = 0.0, wmp = 0.0, wmw = 0.0;
                    T2_ROLLED for (int i = 0; i < col2; ++i) { wmc += cc[i] * v[i]; wmp += pc[i] * v[i]; wmw += wbp[i] * v[i]; }
                    f1 += dibp * wmc;
                    f2 += 2.0 * dibp * wmp - dibp2 * wmw;
                    hand_a = dibp;
                }
                grp.sync();
                const double db = hand_a;
                for (int i = grp.lane; i < col2; i += G) pc[i] -= db * wbp[i];
            }
            grp.sync();
            if (grp.master()) {
                f2 = rmax(kEpsMch * f2_org, f2);
                if (nleft > 0) { dtm = ddiv(-f1, f2); ctl = kCtlContinue; }
                else {
                    if (bnded) { f1 = 0.0; f2 = 0.0; dtm = 0.0; }
                    else dtm = ddiv(-f1, f2);
                    ctl = kCtlBreak;
                }
            }
            if (bcast_ctl(grp) == kCtlBreak) break;
        }
        grp.sync();
        if (grp.master()) {
            if (!all_fixed) {
                if (dtm <= 0.0) dtm = 0.0;
                tsum += dtm;
                T2_ROLLED for (int i = 0; i < N; ++i) z[i] += tsum * dd_c[i];
            }
            hand_a = dtm;
        }
        grp.sync();
        if (cl > 0) {
            const double dtl = hand_a;
            for (int i = grp.lane; i < col2; i += G) cc[i] += dtl * pc[i];
        }
        grp.sync();
        return true;
    }

    // ---- LEL' factorization of the K matrix of the subspace problem (formk); false = not SPD -----------
    T2_NI bool formk(const Group<G> grp) {
        const int cl = col, hd = head, nf = nfree, nen = nenter, ilv = ileave;      // stable during the call
        const bool upd = updatd;
        int upcl;
        if (upd) {
            if (iupdat > kM) {                                  // shift the old part of WN1 one up and one left
                constexpr int RV = (3 * kM + G - 1) / G;
                T2_ROLLED for (int jy = 0; jy < kM - 1; ++jy) {
                    const int js = kM + jy;
                    const int n1 = kM - 1 - jy, ntot = 2 * n1 + (kM - 1);
                    double val[RV];
#pragma unroll
                    for (int rr = 0; rr < RV; ++rr) {
                        const int e = grp.lane + rr * G;
                        val[rr] = 0.0;
                        if (e < n1) val[rr] = WN1(jy + 1 + e, jy + 1);
                        else if (e < 2 * n1) val[rr] = WN1(js + 1 + (e - n1), js + 1);
                        else if (e < ntot) val[rr] = WN1(kM + 1 + (e - 2 * n1), jy + 1);
                    }
                    grp.sync();
#pragma unroll
                    for (int rr = 0; rr < RV; ++rr) {
                        const int e = grp.lane + rr * G;
                        if (e < n1) WN1(jy + e, jy) = val[rr];
                        else if (e < 2 * n1) WN1(js + (e - n1), js) = val[rr];
                        else if (e < ntot) WN1(kM + (e - 2 * n1), jy) = val[rr];
                    }
                }
                grp.sync();
            }
            // new rows in blocks (1,1), (2,1) and (2,2)
            const int ipntr = (hd + cl - 1) % kM;
            const int iy = cl - 1, is = kM + cl - 1;
            for (int jy = grp.lane; jy < cl; jy += G) {
                const int js = kM + jy, jpntr = (hd + jy) % kM;
                double t1 = 0.0, t2 = 0.0, t3 = 0.0;
                T2_ROLLED for (int k = 0; k < nf; ++k) { const int k1 = index[k]; t1 += wy[ipntr][k1] * wy[jpntr][k1]; }
                T2_ROLLED for (int k = nf; k < N; ++k) {
                    const int k1 = index[k];
                    t2 += ws[ipntr][k1] * ws[jpntr][k1];
                    t3 += ws[ipntr][k1] * wy[jpntr][k1];
                }
                WN1(iy, jy) = t1; WN1(is, js) = t2; WN1(is, jy) = t3;
            }
            grp.sync();
            // new column in block (2,1)
            const int jyn = cl - 1, jpn = (hd + cl - 1) % kM;
            for (int i = grp.lane; i < cl; i += G) {
                const int is2 = kM + i, ip = (hd + i) % kM;
                double t3 = 0.0;
                T2_ROLLED for (int k = 0; k < nf; ++k) { const int k1 = index[k]; t3 += ws[ip][k1] * wy[jpn][k1]; }
                WN1(is2, jyn) = t3;
            }
            grp.sync();
            upcl = cl - 1;
        } else {
            upcl = cl;
        }
        // old parts of blocks (1,1) and (2,2): variables that entered / left the free set
        const int upcl_old = (nen > 0 || ilv < N) ? upcl : 0;
        T2_ROLLED for (int iy = 0; iy < upcl_old; ++iy) {
            const int is = kM + iy, ipntr = (hd + iy) % kM;
            for (int jy = grp.lane; jy <= iy; jy += G) {
                const int js = kM + jy, jpntr = (hd + jy) % kM;
                double t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0;
                T2_ROLLED for (int k = 0; k < nen; ++k) {
                    const int k1 = indx2[k];
                    t1 += wy[ipntr][k1] * wy[jpntr][k1];
                    t2 += ws[ipntr][k1] * ws[jpntr][k1];
                }
                T2_ROLLED for (int k = ilv; k < N; ++k) {
                    const int k1 = indx2[k];
                    t3 += wy[ipntr][k1] * wy[jpntr][k1];
                    t4 += ws[ipntr][k1] * ws[jpntr][k1];
                }
                WN1(iy, jy) = WN1(iy, jy) + t1 - t3;
                WN1(is, js) = WN1(is, js) - t2 + t4;
            }
        }
        // old part of block (2,1)
        T2_ROLLED for (int is0 = 0; is0 < upcl_old; ++is0) {
            const int is = kM + is0, ipntr = (hd + is0) % kM;
            for (int jy = grp.lane; jy < upcl; jy += G) {
                const int jpntr = (hd + jy) % kM;
                double t1 = 0.0, t3 = 0.0;
                T2_ROLLED for (int k = 0; k < nen; ++k) { const int k1 = indx2[k]; t1 += ws[ipntr][k1] * wy[jpntr][k1]; }
                T2_ROLLED for (int k = ilv; k < N; ++k) { const int k1 = indx2[k]; t3 += ws[ipntr][k1] * wy[jpntr][k1]; }
                if (is0 <= jy) WN1(is, jy) = WN1(is, jy) + t1 - t3;
                else WN1(is, jy) = WN1(is, jy) - t1 + t3;
            }
        }
        grp.sync();
        // upper triangle of WN = [D + Y'ZZ'Y/theta, -L_a' + R_z'; -L_a + R_z, S'AA'S theta]
        const double th = theta;
        T2_ROLLED for (int iy = 0; iy < cl; ++iy) {
            const int is = cl + iy, is1 = kM + iy;
            for (int jy = grp.lane; jy < cl; jy += G) {
                const int js = cl + jy, js1 = kM + jy;
                if (jy <= iy) {
                    double w11 = ddiv(WN1(iy, jy), th);
                    if (jy == iy) w11 += SY(iy, iy);
                    WN(jy, iy) = w11;
                    WN(js, is) = WN1(is1, js1) * th;
                }
                WN(jy, is) = (jy < iy) ? -WN1(is1, jy) : WN1(is1, jy);
            }
        }
        grp.sync();
        const int col2 = 2 * cl;
        // Cholesky of the (1,1) block and L^-1 (-L_a' + R_z') in the (1,2) block: the same row recurrence
        if (!chol_rows<true, R2>(grp, 0, cl, col2)) return false;
        // (2,2) block += X'X
        const int npair = cl * (cl + 1) / 2;
        for (int tix = grp.lane; tix < npair; tix += G) {
            int a, b;
            tri_unrank(tix, a, b);
            const int is = cl + a, js = cl + b;
            double dot = 0.0;
            T2_INNER for (int q = 0; q < cl; ++q) dot += WN(q, is) * WN(q, js);
            WN(is, js) += dot;
        }
        grp.sync();
        return chol_rows<true, R1>(grp, cl, cl, col2);
    }

    // ---- subspace minimisation over the free variables at the Cauchy point (cmprlb + subsm) -------------
    T2_NI bool subsm(const Group<G> grp) {
        double* cc = scr.vec.cc; double* wv = scr.vec.wv;
        const int cl = col, col2 = 2 * cl, nf = nfree, hd = head;
        const double th = theta;
        // reduced gradient r = -Z'(B (xcp - x) + g)
        if (!cnstnd && cl > 0) {
            if (grp.master()) T2_ROLLED for (int i = 0; i < N; ++i) rr_[i] = -g[i];
        } else {
            if (grp.master()) T2_ROLLED for (int i = 0; i < nf; ++i) { const int k = index[i]; rr_[i] = -th * (z[k] - x[k]) - g[k]; }
            if (!bmv(grp, cc, wv)) return false;
            if (grp.master()) {
                T2_ROLLED for (int j = 0; j < cl; ++j) {
                    const int pt = (hd + j) % kM;
                    const double a1 = wv[j], a2 = th * wv[cl + j];
                    T2_ROLLED for (int i = 0; i < nf; ++i) { const int k = index[i]; rr_[i] += wy[pt][k] * a1 + ws[pt][k] * a2; }
                }
            }
        }
        grp.sync();
        // wv = W'Z d, then K^-1 wv through the LEL' factors
        for (int i = grp.lane; i < cl; i += G) {
            const int pt = (hd + i) % kM;
            double t1 = 0.0, t2 = 0.0;
            T2_ROLLED for (int j = 0; j < nf; ++j) { const int k = index[j]; t1 += wy[pt][k] * rr_[j]; t2 += ws[pt][k] * rr_[j]; }
            wv[i] = t1; wv[cl + i] = th * t2;
        }
        grp.sync();
        if (zero_pivot<true>(grp, col2)) return false;
        trsl_t<true, R2>(grp, col2, wv);
        for (int i = grp.lane; i < cl; i += G) wv[i] = -wv[i];
        grp.sync();
        if (grp.master()) trsl_n_master<true>(col2, wv);
        grp.sync();
        // d = (1/theta) r + (1/theta^2) Z'W wv: one lane per free variable, pairs in the serial order
        for (int i = grp.lane; i < nf; i += G) {
            const int k = index[i];
            double acc = rr_[i];
            T2_ROLLED for (int jy = 0; jy < cl; ++jy) {
                const int js = cl + jy, pt = (hd + jy) % kM;
                acc += ddiv(wy[pt][k] * wv[jy], th) + ws[pt][k] * wv[js];
            }
            rr_[i] = acc;
        }
        grp.sync();
        if (grp.master()) {
            const double inv_theta = ddiv(1.0, th);
            T2_ROLLED for (int i = 0; i < nf; ++i) rr_[i] *= inv_theta;
            // projection of the Newton point onto the box (v3.0), else backtrack along the Newton direction
            T2_ROLLED for (int i = 0; i < N; ++i) xp[i] = z[i];
            bool iword = false;
            T2_ROLLED for (int a = 0; a < nf; ++a) {
                const int k = index[a];
                const double dk = rr_[a], xk = z[k];
                if (nbd[k] == 0) z[k] = xk + dk;
                else if (nbd[k] == 1) { z[k] = rmax(l[k], xk + dk); if (z[k] == l[k]) iword = true; }
                else if (nbd[k] == 2) { z[k] = rmin(u[k], rmax(l[k], xk + dk)); if (z[k] == l[k] || z[k] == u[k]) iword = true; }
                else { z[k] = rmin(u[k], xk + dk); if (z[k] == u[k]) iword = true; }
            }
            if (iword) {
                double dd_p = 0.0;
                T2_ROLLED for (int i = 0; i < N; ++i) dd_p += (z[i] - x[i]) * g[i];
                if (dd_p > 0.0) {
                    T2_ROLLED for (int i = 0; i < N; ++i) z[i] = xp[i];
                    double alpha = 1.0, temp1 = 1.0;
                    int ibd = -1;
                    T2_ROLLED for (int a = 0; a < nf; ++a) {
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
                    T2_ROLLED for (int a = 0; a < nf; ++a) z[index[a]] += alpha * rr_[a];
                }
            }
        }
        grp.sync();
        return true;
    }

    // ---- new correction pair into WS, WY, S'S, S'Y (matupd) and the factor of T (formt); false = T not SPD ----
    // upd_rr / upd_dr hold rr and dr (written by the master, synchronised)
    T2_NI bool update_pairs(const Group<G> grp) {
        const double rrv = upd_rr, drv = upd_dr;
        if (grp.master()) {
            updatd = true;
            ++iupdat;
            if (iupdat <= kM) { col = iupdat; itail = (head + iupdat - 1) % kM; }
            else { itail = (itail + 1) % kM; head = (head + 1) % kM; }
            T2_ROLLED for (int i = 0; i < N; ++i) { ws[itail][i] = d[i]; wy[itail][i] = r[i]; }
            theta = ddiv(rrv, drv);
        }
        grp.sync();
        const int cl = col, hd = head, itl = itail;
        const double th = theta;
        if (iupdat > kM) {                                      // move the old information up and left
            constexpr int RV = (kM + G - 1) / G;
            T2_ROLLED for (int j = 0; j < cl - 1; ++j) {
                const int n1 = j + 1, ntot = n1 + (cl - 1 - j);
                double val[RV];
#pragma unroll
                for (int rr = 0; rr < RV; ++rr) {
                    const int e = grp.lane + rr * G;
                    val[rr] = 0.0;
                    if (e < n1) val[rr] = SS(e + 1, j + 1);
                    else if (e < ntot) val[rr] = SY(j + 1 + (e - n1), j + 1);
                }
                grp.sync();
#pragma unroll
                for (int rr = 0; rr < RV; ++rr) {
                    const int e = grp.lane + rr * G;
                    if (e < n1) SS(e, j) = val[rr];
                    else if (e < ntot) SY(j + (e - n1), j) = val[rr];
                }
            }
            grp.sync();
        }
        for (int j = grp.lane; j < cl - 1; j += G) {            // last row of S'Y, last column of S'S
            const int pt = (hd + j) % kM;
            double a = 0.0, b = 0.0;
            T2_ROLLED for (int i = 0; i < N; ++i) { a += ws[itl][i] * wy[pt][i]; b += ws[pt][i] * ws[itl][i]; }
            SY(cl - 1, j) = a;
            SS(j, cl - 1) = b;
        }
        if (grp.master()) {
            SS(cl - 1, cl - 1) = (stp == 1.0) ? dtd : stp * stp * dtd;
            SY(cl - 1, cl - 1) = drv;
        }
        grp.sync();
        for (int k = grp.lane; k < cl; k += G) { rd[k] = ddiv(1.0, SY(k, k)); rsd[k] = ddiv(1.0, dsqrt(SY(k, k))); }
        grp.sync();
        // T = theta S'S + L D^-1 L' (upper triangle), then its Cholesky factor
        const int npair = cl * (cl + 1) / 2;
        for (int tix = grp.lane; tix < npair; tix += G) {
            int i, j;
            tri_unrank(tix, i, j);                              // i <= j
            if (i == 0) wt[0][j] = th * SS(0, j);
            else {
                double ddum = 0.0;
                T2_INNER for (int k = 0; k < i; ++k) ddum += SY(i, k) * SY(j, k) * rd[k];
                wt[i][j] = ddum + th * SS(i, j);
            }
        }
        grp.sync();
        return chol_rows<false, R1>(grp, 0, cl, cl);
    }

    // ---- More'-Thuente safeguarded step (MINPACK-2 dcstep), master lane ------------------------------
    T2_NI static void dcstep(double& stx_, double& fx_, double& dx_, double& sty_, double& fy_, double& dy_, double& stp_,
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

    // dcsrch after the first call (master lane): 0 = evaluate at the new stp, 1 = line search finished
    T2_NI int dcsrch_next(double fv, double gv) {
        const double ls_gtol = 0.9, ls_xtol = 0.1, stpmin = 0.0, stpmax = stpmx;
        const double ftest = finit + stp * gtest;
        if (stage == 1 && fv <= ftest && gv >= 0.0) stage = 2;
        bool fin = false;
        if (brackt && (stp <= stmin || stp >= stmax)) fin = true;
        if (brackt && stmax - stmin <= ls_xtol * stmax) fin = true;
        if (stp == stpmax && fv <= ftest && gv <= gtest) fin = true;
        if (stp == stpmin && (fv > ftest || gv >= gtest)) fin = true;
        if (fv <= ftest && fabs(gv) <= ls_gtol * (-ginit)) fin = true;
        if (fin) return 1;
        // dcstep works on locals (the state lives in shared memory: no references into it across the call)
        double stx_ = stx, sty_ = sty, stp_ = stp;
        bool br = brackt;
        if (stage == 1 && fv <= fx && fv > ftest) {
            const double fm = fv - stp * gtest;
            double fxm = fx - stx * gtest, fym = fy - sty * gtest;
            const double gm = gv - gtest;
            double gxm = gx - gtest, gym = gy - gtest;
            dcstep(stx_, fxm, gxm, sty_, fym, gym, stp_, fm, gm, br, stmin, stmax);
            fx = fxm + stx_ * gtest; fy = fym + sty_ * gtest;
            gx = gxm + gtest; gy = gym + gtest;
        } else {
            double fx_ = fx, gx_ = gx, fy_ = fy, gy_ = gy;
            dcstep(stx_, fx_, gx_, sty_, fy_, gy_, stp_, fv, gv, br, stmin, stmax);
            fx = fx_; gx = gx_; fy = fy_; gy = gy_;
        }
        stx = stx_; sty = sty_; stp = stp_; brackt = br;
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

    // ---- start: bounds, projected start point (master lane) ----------------------------------------
    T2_NI void setup(const double* x0, const double* lo, const double* hi, double ftol_, double pgtol_, int maxls_) {
        ftol = ftol_; pgtol = pgtol_; maxls = maxls_;
        cnstnd = false; boxed = true;
        T2_ROLLED for (int i = 0; i < N; ++i) {
            const bool hl = lo[i] > -INFINITY, hu = hi[i] < INFINITY;
            nbd[i] = hl ? (hu ? 2 : 1) : (hu ? 3 : 0);
            l[i] = hl ? lo[i] : 0.0; u[i] = hu ? hi[i] : 0.0;
            double xi = x0[i];
            if (hl) xi = rmax(xi, l[i]);
            if (hu) xi = rmin(xi, u[i]);
            x[i] = xi;
            if (nbd[i] != 2) boxed = false;
            if (nbd[i] == 0) iwhere[i] = -1;
            else { cnstnd = true; iwhere[i] = (nbd[i] == 2 && u[i] - l[i] <= 0.0) ? 3 : 0; }
        }
        reset_memory();
        itail = 0; nfree = N; nenter = 0; ileave = N;
        T2_ROLLED for (int i = 0; i < N; ++i) { index[i] = i; indx2[i] = i; }
        fold = dnorm = dtd = gd = gdold = stp = stpmx = sbgnrm = 0.0;
        iter = ifun = iback = nfgv = 0;
        result = kRunning;
    }

    T2_HD void trial_point() {
        if (stp == 1.0) { T2_ROLLED for (int i = 0; i < N; ++i) x[i] = z[i]; }
        else { T2_ROLLED for (int i = 0; i < N; ++i) x[i] = stp * d[i] + t[i]; }
    }

    // new search direction and the first trial point of its line search (label 222 ... 666); all lanes
    T2_NI void start_iteration(const Group<G> grp) {
        for (;;) {
            grp.sync();
            const bool use_cauchy = !(!cnstnd && col > 0);
            if (!use_cauchy) {
                if (grp.master()) { T2_ROLLED for (int i = 0; i < N; ++i) z[i] = x[i]; wrk = updatd; }
            } else {
                if (!cauchy(grp)) { if (grp.master()) reset_memory(); continue; }
                if (grp.master()) wrk = freev();
            }
            grp.sync();
            if (nfree != 0 && col != 0) {
                if (wrk && !formk(grp)) { grp.sync(); if (grp.master()) reset_memory(); continue; }
                if (!subsm(grp)) { grp.sync(); if (grp.master()) reset_memory(); continue; }
            }
            if (grp.master()) {
                T2_ROLLED for (int i = 0; i < N; ++i) d[i] = z[i] - x[i];
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
                    if (col == 0) { result = kAbnormal; ctl = kCtlReturn; }
                    else { reset_memory(); ctl = kCtlContinue; }
                } else {
                    brackt = false; stage = 1; finit = f; ginit = gd; gtest = 1e-3 * ginit;
                    width = stpmx - 0.0; width1 = width * 2.0;
                    stx = 0.0; fx = finit; gx = ginit; sty = 0.0; fy = finit; gy = ginit;
                    stmin = 0.0; stmax = stp + 4.0 * stp;
                    ifun = 1; ++nfgv; iback = 0;
                    trial_point();
                    ctl = kCtlReturn;
                }
            }
            grp.sync();
            if (ctl == kCtlReturn) return;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// One voxel as a resumable cooperative run (mirror of VoxelRun): start / pass / finish
// ---------------------------------------------------------------------------------------------
template <int OBJ, int G>
struct CoopRun {
    static constexpr int N = (OBJ == 0) ? 2 : 3;
    CoopSolver<N, G> s;
    float y[kMaxEcho];
    float yraw[kMaxEcho];          // the row as loaded (the kernel's lanes write it before start())
    double lo[3], hi[3];
    double xprev[N];
    double fg[N + 1];              // f and the forward differences' f values of the current evaluation
    float* trace_f;
    float* trace_step;
    int trace_cap, tl;
    int nit, nfev, status;
    bool have_prev, started, active;

    // master lane does fit_voxel's preamble (:237-245); yraw complete and synchronised
    T2_NI void start(const Group<G> grp, const LbConsts& c, float* tf, float* ts, int tcap) {
        if (grp.master()) {
            const int E = c.n_echo;
            bool finite = true;
            float ymax = yraw[0];
            T2_ROLLED for (int e = 0; e < E; ++e) {
                y[e] = yraw[e];
                finite = finite && ((yraw[e] - yraw[e]) == 0.0f);
                ymax = yraw[e] > ymax ? yraw[e] : ymax;
            }
            if (c.norm) {
                T2_ROLLED for (int e = 0; e < E; ++e) { y[e] = yraw[e] / ymax; finite = finite && ((y[e] - y[e]) == 0.0f); }
            }
            for (int i = 0; i < 3; ++i) { lo[i] = c.lb[i]; hi[i] = c.ub[i]; }
            if (c.no_prior) lo[0] = (double)yraw[0];
            status = kOk;
            if (c.no_prior && (yraw[0] > (float)hi[0])) status = kBadBounds;
            else if (!finite) status = kNonFinite;
            if (OBJ == 2 && status == kOk) {
                T2_ROLLED for (int e = 0; e < E; ++e) if (!(y[e] > 0.0f)) status = kNonFinite;
            }
            s.setup(c.x0, lo, hi, c.ftol, c.pgtol, c.maxls);
            nit = 0; nfev = 0; tl = 0;
            have_prev = false; started = false;
            trace_f = tf; trace_step = ts; trace_cap = tcap;
            active = status == kOk;
        }
        // zero-initialised workspace, as scipy hands to setulb for every minimize() call (see Solver::setup)
        {
            using SolverT = CoopSolver<N, G>;
            static_assert(offsetof(SolverT, scr) + sizeof(s.scr) - offsetof(SolverT, wnn) ==
                          sizeof(s.wnn) + sizeof(s.sys) + sizeof(s.wt) + sizeof(s.ws) + sizeof(s.wy) + sizeof(s.rd) + sizeof(s.rsd) + sizeof(s.scr),
                          "matrix workspace must be contiguous");
            double* m0 = &s.wnn[0][0];
            constexpr int nm = (int)((sizeof(s.wnn) + sizeof(s.sys) + sizeof(s.wt) + sizeof(s.ws) + sizeof(s.wy) + sizeof(s.rd) + sizeof(s.rsd) + sizeof(s.scr)) / sizeof(double));
            for (int i = grp.lane; i < nm; i += G) m0[i] = 0.0;
        }
        grp.sync();
    }

    // f at x and at the N forward-difference points: the (N + 1) * E echo terms across the lanes, one lane per point for
    // numpy's pairwise sum, then the gradient on every lane.  Returns f; gv on every lane.
    T2_NI double fun_and_grad(const Group<G> grp, const LbConsts& c, double* gv) {
        const int E = c.n_echo;
        double xb[N], dx[N];
#pragma unroll
        for (int i = 0; i < N; ++i) xb[i] = s.x[i];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double h = fd_step<N>(xb, lo, hi, i, c.fd_step);
            const double xt = xb[i] + h;
            dx[i] = xt - xb[i];
        }
        double* ft = s.scr.fterm;
        for (int it = grp.lane; it < (N + 1) * E; it += G) {
            const int pnt = it / E, e = it - pnt * E;
            double xt[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                xt[i] = xb[i];
                if (pnt == i + 1) { const double h = fd_step<N>(xb, lo, hi, i, c.fd_step); xt[i] = xb[i] + h; }
            }
            ft[pnt * kMaxEcho + e] = objective_term<OBJ>(xt, y[e], c.te[e]);
        }
        grp.sync();
        for (int pnt = grp.lane; pnt < N + 1; pnt += G) fg[pnt] = objective_reduce<OBJ>(ft + pnt * kMaxEcho, E);
        grp.sync();
        const double fv = fg[0];
#pragma unroll
        for (int i = 0; i < N; ++i) gv[i] = ddiv(fg[i + 1] - fv, dx[i]);
        return fv;
    }

    T2_NI void pass(const Group<G> grp, const LbConsts& c) {
        double gv[N];
        const double fv = fun_and_grad(grp, c, gv);
        grp.sync();                                      // everyone has read fg / x before the master moves on
        if (grp.master()) {
            nfev += N + 1;
            s.ctl = kCtlNone;
            if (!started && !(fv - fv == 0.0)) {
                s.f = fv; s.result = kAbnormal; active = false;
                s.ctl = kCtlDone;
            } else if (!started) {
                started = true;
                s.f = fv;
                T2_ROLLED for (int i = 0; i < N; ++i) s.g[i] = gv[i];
                s.nfgv = 1;
                s.projgr();
                if (s.sbgnrm <= s.pgtol) { s.result = kConvPg; s.ctl = kCtlDone; }
                else s.ctl = kCtlStartIter;
            } else {
                // advance(): the line search's reaction to f, g at the trial point
                s.f = fv;
                T2_ROLLED for (int i = 0; i < N; ++i) s.g[i] = gv[i];
                s.gd = 0.0;
                T2_ROLLED for (int i = 0; i < N; ++i) s.gd += s.g[i] * s.d[i];
                if (s.dcsrch_next(s.f, s.gd) == 0) {
                    ++s.ifun; ++s.nfgv; s.iback = s.ifun - 1;
                    if (s.iback >= s.maxls) {
                        T2_ROLLED for (int i = 0; i < N; ++i) { s.x[i] = s.t[i]; s.g[i] = s.r[i]; }
                        s.f = s.fold;
                        if (s.col == 0) { s.result = kAbnormal; s.ctl = kCtlDone; }
                        else { s.reset_memory(); s.ctl = kCtlStartIter; }
                    } else {
                        s.trial_point();
                        s.ctl = kCtlTrial;
                    }
                } else {
                    ++s.iter;
                    s.projgr();
                    // scipy: n_iterations += 1; callback(x)
                    ++nit;
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
                    if (nit >= c.maxiter) { s.result = kMaxIter; s.ctl = kCtlDone; }
                    else if (nfev > c.maxfun) { s.result = kMaxFun; s.ctl = kCtlDone; }
                    else {
                        // continue_after_iterate(): tests, then the pair update
                        if (s.sbgnrm <= s.pgtol) { s.result = kConvPg; s.ctl = kCtlDone; }
                        else {
                            const double ddum0 = rmax(fabs(s.fold), rmax(fabs(s.f), 1.0));
                            if ((s.fold - s.f) <= s.ftol * ddum0) { s.result = kConvF; s.ctl = kCtlDone; }
                            else {
                                double rr = 0.0;
                                T2_ROLLED for (int i = 0; i < N; ++i) { s.r[i] = s.g[i] - s.r[i]; rr += s.r[i] * s.r[i]; }
                                double dr, ddum;
                                if (s.stp == 1.0) { dr = s.gd - s.gdold; ddum = -s.gdold; }
                                else {
                                    dr = (s.gd - s.gdold) * s.stp;
                                    T2_ROLLED for (int i = 0; i < N; ++i) s.d[i] *= s.stp;
                                    ddum = -s.gdold * s.stp;
                                }
                                if (dr <= kEpsMch * ddum) { s.updatd = false; s.ctl = kCtlSkipUpdate; }
                                else { s.upd_rr = rr; s.upd_dr = dr; s.ctl = kCtlUpdate; }
                            }
                        }
                    }
                }
            }
        }
        grp.sync();
        const int ctl = s.ctl;
        if (ctl == kCtlUpdate) {
            const bool ok = s.update_pairs(grp);
            if (!ok) { grp.sync(); if (grp.master()) s.reset_memory(); }
        }
        if (ctl == kCtlUpdate || ctl == kCtlSkipUpdate || ctl == kCtlStartIter) s.start_iteration(grp);
        grp.sync();
        if (grp.master() && s.result != kRunning) active = false;
        grp.sync();
    }

    // master lane
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
