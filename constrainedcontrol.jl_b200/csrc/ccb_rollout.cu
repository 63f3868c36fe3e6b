// csrc/ccb_rollout.cu -- K1: fused closed-loop rollout kernel (sm_100a, FP64).
//
// One group of G lanes (G = 8, 16 or 32; a warp holds 32/G systems) owns one system for the whole rollout:
//   for k: [record] -> control!(k) -> applyFτ! -> newton!(ε, newtonIter, lineIter) -> updatestate!
// which is the body of ConstrainedDynamics.jl's simulate! loop with the reference's
// control laws fused in (reference src/control/lqr.jl:89-139, lqr_tracking.jl:46-71, pid.jl:43-88).
//
// Work decomposition inside a group: lanes act as *body lanes* (one body each) for the integrator residual and
// the 6x6 body blocks, and as *row lanes* (one scalar constraint row = one λ component each) for the joint
// residuals / Jacobian rows and the rows of the constraint Schur complement S = Gv D^-1 G^T.  The KKT system is
// solved by eliminating the (block-diagonal) bodies first and sweeping S along the mechanism tree in leaf-first
// order with the rhs carried along (block Gauss-Jordan, fill-in only inside kinematic loops), then
// back-substituting root-first.  All per-system blocks live in shared memory; HBM is touched only for the initial
// and final state, the optional recording and the gain rows.
#include "ccb_dev_math.cuh"
#include "ccb_internal.h"

namespace ccb {

namespace {

struct SysMem {
    double *z, *s, *s0, *p3, *dinv, *tw, *d, *fext, *lam, *lam0, *rhs, *gk, *S, *u, *dz, *ds, *dl;
};

__device__ __forceinline__ SysMem carve(double* base, const DevMech& M) {
    SysMem m;
    m.z = base + M.o_z; m.s = base + M.o_s; m.s0 = base + M.o_s0; m.p3 = base + M.o_p3; m.dinv = base + M.o_dinv; m.tw = base + M.o_tw;
    m.d = base + M.o_d; m.fext = base + M.o_fext; m.lam = base + M.o_lam; m.lam0 = base + M.o_lam0; m.rhs = base + M.o_rhs; m.gk = base + M.o_gk;
    m.S = base + M.o_S; m.u = base + M.o_u; m.dz = base + M.o_dz; m.ds = base + M.o_ds; m.dl = base + M.o_dl;
    return m;
}

// Residual of the KKT system at the current iterate (s, lam); with JAC also the per-body blocks, the Jacobian
// rows, the Schur rhs and the rows of S.  Returns the sum of squares of the residual over this lane's items.
template <int G, bool JAC>
__device__ __forceinline__ double eval_system(const DevMech& M, const SysMem& m, int lane, bool active) {
    const double dt = M.dt, idt = 1.0 / M.dt, c4 = 4.0 / (M.dt * M.dt);
    double ss = 0.0;
    if (active) {
        for (int b = lane; b < M.nb; b += G) {
            const double* zb = m.z + 13 * b;
            const double* sb = m.s + 6 * b;
            V3 x2 = ld3(zb), v1 = ld3(zb + 7), w1 = ld3(zb + 10), v2 = ld3(sb), w2 = ld3(sb + 3);
            Q4 q2 = ld4(zb + 3);
            double sq2 = sqrt(c4 - dot(w2, w2));
            V3 x3 = x2 + dt * v2;
            Q4 q3 = qmul(q2, Q4{0.5 * dt * sq2, 0.5 * dt * w2.x, 0.5 * dt * w2.y, 0.5 * dt * w2.z});
            st3(m.p3 + 7 * b, x3);
            st4(m.p3 + 7 * b + 3, q3);
            const double mass = __ldg(M.mass + b);
            double J[9];
#pragma unroll
            for (int i = 0; i < 9; i++) J[i] = __ldg(M.inertia + 9 * b + i);
            V3 F = ld3(m.fext + 6 * b), tau = ld3(m.fext + 6 * b + 3);
            double sq1 = sqrt(c4 - dot(w1, w1));
            V3 Jw1 = mat3_mul(J, w1), Jw2 = mat3_mul(J, w2);
            V3 dT = {mass * ((v2.x - v1.x) * idt) - F.x, mass * ((v2.y - v1.y) * idt) - F.y, mass * ((v2.z - v1.z) * idt - M.g) - F.z};
            V3 dR = sq2 * Jw2 + cross(w2, Jw2) - sq1 * Jw1 + cross(w1, Jw1) - 2.0 * tau;
            double d[6] = {dT.x, dT.y, dT.z, dR.x, dR.y, dR.z};
            for (int e = __ldg(M.binc_ptr + b); e < __ldg(M.binc_ptr + b + 1); e++) {   // d -= G^T λ (Jacobian at the knot)
                const int r = __ldg(M.binc_row + e);
                const double* g = m.gk + 12 * r + 6 * __ldg(M.binc_side + e);
                const double l = m.lam[r];
#pragma unroll
                for (int i = 0; i < 6; i++) d[i] = fma(-g[i], l, d[i]);
            }
#pragma unroll
            for (int i = 0; i < 6; i++) { m.d[6 * b + i] = d[i]; ss = fma(d[i], d[i], ss); }
            if (JAC) {
                // rotational block of D_b = sq2 J + [w2]x J - [J w2]x - (J w2) w2^T / sq2, then its inverse
                double Dr[9];
                const double w[3] = {w2.x, w2.y, w2.z}, jw[3] = {Jw2.x, Jw2.y, Jw2.z};
                const double isq = 1.0 / sq2;
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
                        double wxJ = w[i1] * J[i2 * 3 + j] - w[i2] * J[i1 * 3 + j];   // ([w]x J)_ij
                        double jwx = (j == i1) ? -jw[i2] : ((j == i2) ? jw[i1] : 0.0);  // ([Jw]x)_ij
                        Dr[i * 3 + j] = sq2 * J[i * 3 + j] + wxJ - jwx - jw[i] * w[j] * isq;
                    }
                inv3(Dr, m.dinv + 9 * b);
                const double c = 0.25 * dt * dt;
                double* T = m.tw + 9 * b;   // T_w = (dt^2/4)(w w^T / sq2 + sq2 I - [w]x)
                T[0] = c * (w[0] * w[0] * isq + sq2); T[1] = c * (w[0] * w[1] * isq + w[2]); T[2] = c * (w[0] * w[2] * isq - w[1]);
                T[3] = c * (w[1] * w[0] * isq - w[2]); T[4] = c * (w[1] * w[1] * isq + sq2); T[5] = c * (w[1] * w[2] * isq + w[0]);
                T[6] = c * (w[2] * w[0] * isq + w[1]); T[7] = c * (w[2] * w[1] * isq - w[0]); T[8] = c * (w[2] * w[2] * isq + sq2);
            }
        }
    }
    __syncwarp();
    if (active) {
        for (int r = lane; r < M.ml; r += G) {
            const int j = __ldg(M.row_joint + r);
            const int a = __ldg(M.j_parent + j), b = __ldg(M.j_child + j);
            const int kind = __ldg(M.row_kind + r);
            V3 pr = ldg3(M.row_proj + 3 * r), p1 = ldg3(M.j_p1 + 3 * j), p2 = ldg3(M.j_p2 + 3 * j);
            Q4 qoff = ldg4(M.j_qoff + 4 * j);
            V3 xa = {0.0, 0.0, 0.0};
            Q4 qa = {1.0, 0.0, 0.0, 0.0};
            if (a >= 0) { xa = ld3(m.p3 + 7 * a); qa = ld4(m.p3 + 7 * a + 3); }
            V3 xb = ld3(m.p3 + 7 * b);
            Q4 qb = ld4(m.p3 + 7 * b + 3);
            double g;
            V3 Gax, Gap, Gbx, Gbp;
            row_eval(kind, pr, p1, p2, qoff, xa, qa, xb, qb, JAC, g, Gax, Gap, Gbx, Gbp);
            ss = fma(g, g, ss);
            if (!JAC) {
                m.rhs[r] = g;
            } else {
                // Gv = [G_x dt, G_phi T_w];  w = Gv D^-1  (D = blkdiag(m/dt I, Drot))
                V3 wbx = (dt * dt / __ldg(M.mass + b)) * Gbx;
                V3 wbp = vec_mat3(vec_mat3(Gbp, m.tw + 9 * b), m.dinv + 9 * b);
                double rhs = g - dot(wbx, ld3(m.d + 6 * b)) - dot(wbp, ld3(m.d + 6 * b + 3));
                V3 wax = {0.0, 0.0, 0.0}, wap = {0.0, 0.0, 0.0};
                if (a >= 0) {
                    wax = (dt * dt / __ldg(M.mass + a)) * Gax;
                    wap = vec_mat3(vec_mat3(Gap, m.tw + 9 * a), m.dinv + 9 * a);
                    rhs -= dot(wax, ld3(m.d + 6 * a)) + dot(wap, ld3(m.d + 6 * a + 3));
                }
                m.rhs[r] = rhs;
                // row of S = Gv D^-1 G^T against every coupled joint (G at the knot), zero for pure fill-in blocks
                const int ncol = __ldg(M.j_ncol + j), lr = r - __ldg(M.j_row0 + j);
                double* Srow = m.S + __ldg(M.j_soff + j) + lr * ncol;
                for (int c = 0; c < ncol; c++) Srow[c] = 0.0;
                for (int e = __ldg(M.nbr_ptr + j); e < __ldg(M.nbr_ptr + j + 1); e++) {
                    const int q = __ldg(M.nbr_joint + e), cp = __ldg(M.nbr_cpos + e), mask = __ldg(M.nbr_pairs + e);
                    const int rq0 = __ldg(M.j_row0 + q), dq = __ldg(M.j_dim + q);
                    for (int t = 0; t < dq; t++) {
                        const double* gq = m.gk + 12 * (rq0 + t);
                        double v = 0.0;
                        if (mask & 1) v += dot(wax, ld3(gq)) + dot(wap, ld3(gq + 3));
                        if (mask & 2) v += dot(wax, ld3(gq + 6)) + dot(wap, ld3(gq + 9));
                        if (mask & 4) v += dot(wbx, ld3(gq)) + dot(wbp, ld3(gq + 3));
                        if (mask & 8) v += dot(wbx, ld3(gq + 6)) + dot(wbp, ld3(gq + 9));
                        Srow[cp + t] = v;
                    }
                }
                Srow[lr] += M.reg;
            }
        }
    }
    __syncwarp();
    return ss;
}

// Elimination of one joint block of compile-time size DC.  Every lane factorises S_cc redundantly in registers (LU
// without pivoting: S_cc is the Schur complement of a quasi-definite KKT matrix) and applies the solve to the column(s)
// it owns of [S_c,later | rhs_c]; afterwards lanes own (row, column) items of the later neighbours' Schur update.
template <int G, int DC>
__device__ __forceinline__ void eliminate_block(const DevMech& M, const SysMem& m, int c, int lane, bool active) {
    const int r0 = __ldg(M.j_row0 + c), ncol = __ldg(M.j_ncol + c);
    const int a0 = __ldg(M.acol_ptr + c), nac = __ldg(M.acol_ptr + c + 1) - a0;
    double* Sc = m.S + __ldg(M.j_soff + c);
    if (active && lane <= nac) {
        double A[DC][DC], invd[DC];
#pragma unroll
        for (int i = 0; i < DC; i++)
#pragma unroll
            for (int j = 0; j < DC; j++) A[i][j] = Sc[i * ncol + j];
#pragma unroll
        for (int k = 0; k < DC; k++) {
            invd[k] = 1.0 / A[k][k];
#pragma unroll
            for (int i = k + 1; i < DC; i++) {
                A[i][k] *= invd[k];
#pragma unroll
                for (int j = k + 1; j < DC; j++) A[i][j] = fma(-A[i][k], A[k][j], A[i][j]);
            }
        }
        for (int t = lane; t <= nac; t += G) {
            double x[DC];
            const int cc = (t < nac) ? __ldg(M.acol + a0 + t) : 0;
#pragma unroll
            for (int i = 0; i < DC; i++) x[i] = (t < nac) ? Sc[i * ncol + cc] : m.rhs[r0 + i];
#pragma unroll
            for (int k = 0; k < DC; k++)
#pragma unroll
                for (int i = k + 1; i < DC; i++) x[i] = fma(-A[i][k], x[k], x[i]);
#pragma unroll
            for (int k = DC - 1; k >= 0; k--) {
#pragma unroll
                for (int j = k + 1; j < DC; j++) x[k] = fma(-A[k][j], x[j], x[k]);
                x[k] *= invd[k];
            }
#pragma unroll
            for (int i = 0; i < DC; i++) {
                if (t < nac) Sc[i * ncol + cc] = x[i];
                else m.rhs[r0 + i] = x[i];
            }
        }
    }
    __syncwarp();
    // Schur update of every later neighbour p:  S_p,* -= S_pc U_c,* ; rhs_p -= S_pc y_c
    for (int e = __ldg(M.lat_ptr + c); e < __ldg(M.lat_ptr + c + 1); e++) {
        const int p = __ldg(M.lat_p + e), pc0 = __ldg(M.lat_pc0 + e), mo = __ldg(M.lat_map + e);
        const int dp = __ldg(M.j_dim + p), ncp = __ldg(M.j_ncol + p), rp0 = __ldg(M.j_row0 + p);
        double* Sp = m.S + __ldg(M.j_soff + p);
        const int items = dp * (nac + 1);
        if (active) {
            for (int it = lane; it < items; it += G) {
                const int row = it / (nac + 1), i = it - row * (nac + 1);
                const double* t = Sp + row * ncp + pc0;
                double acc = 0.0;
                if (i < nac) {
                    const int cc = __ldg(M.acol + a0 + i);
#pragma unroll
                    for (int k = 0; k < DC; k++) acc = fma(t[k], Sc[k * ncol + cc], acc);
                    Sp[row * ncp + __ldg(M.map + mo + i)] -= acc;
                } else {
#pragma unroll
                    for (int k = 0; k < DC; k++) acc = fma(t[k], m.rhs[r0 + k], acc);
                    m.rhs[rp0 + row] -= acc;
                }
            }
        }
        __syncwarp();
    }
}

// Solve S dλ = rhs by the leaf-first block sweep (rhs carried along), back-substitute, then recover the body steps
// ds_b = D_b^-1 (d_b + G_b^T dλ).  On exit m.rhs holds dλ and m.ds holds ds.
template <int G> __device__ __forceinline__ void solve_system(const DevMech& M, const SysMem& m, int lane, bool active) {
    for (int idx = 0; idx < M.nj; idx++) {
        const int c = __ldg(M.elim + idx);
        switch (__ldg(M.j_dim + c)) {
            case 1: eliminate_block<G, 1>(M, m, c, lane, active); break;
            case 2: eliminate_block<G, 2>(M, m, c, lane, active); break;
            case 3: eliminate_block<G, 3>(M, m, c, lane, active); break;
            case 4: eliminate_block<G, 4>(M, m, c, lane, active); break;
            case 5: eliminate_block<G, 5>(M, m, c, lane, active); break;
            default: eliminate_block<G, 6>(M, m, c, lane, active); break;
        }
    }
    // back substitution, root first
    for (int idx = M.nj - 1; idx >= 0; idx--) {
        const int c = __ldg(M.elim + idx);
        const int dc = __ldg(M.j_dim + c), r0 = __ldg(M.j_row0 + c), ncol = __ldg(M.j_ncol + c);
        const int a0 = __ldg(M.acol_ptr + c), nac = __ldg(M.acol_ptr + c + 1) - a0;
        if (active && lane < dc && nac > 0) {
            const double* row = m.S + __ldg(M.j_soff + c) + lane * ncol;
            double x = m.rhs[r0 + lane];
            for (int t = 0; t < nac; t++) x = fma(-row[__ldg(M.acol + a0 + t)], m.rhs[__ldg(M.acol_row + a0 + t)], x);
            m.rhs[r0 + lane] = x;
        }
        __syncwarp();
    }
    // body steps
    if (active) {
        for (int b = lane; b < M.nb; b += G) {
            double t[6];
#pragma unroll
            for (int i = 0; i < 6; i++) t[i] = m.d[6 * b + i];
            for (int e = __ldg(M.binc_ptr + b); e < __ldg(M.binc_ptr + b + 1); e++) {
                const int r = __ldg(M.binc_row + e);
                const double* g = m.gk + 12 * r + 6 * __ldg(M.binc_side + e);
                const double l = m.rhs[r];
#pragma unroll
                for (int i = 0; i < 6; i++) t[i] = fma(g[i], l, t[i]);
            }
            const double k = M.dt / __ldg(M.mass + b);
            V3 dw = mat3_mul(m.dinv + 9 * b, v3(t[3], t[4], t[5]));
            m.ds[6 * b + 0] = k * t[0]; m.ds[6 * b + 1] = k * t[1]; m.ds[6 * b + 2] = k * t[2];
            m.ds[6 * b + 3] = dw.x; m.ds[6 * b + 4] = dw.y; m.ds[6 * b + 5] = dw.z;
        }
    }
    __syncwarp();
}

// control!(mechanism, k): fills m.u[e] (scalar input per eqc) and then the body wrenches m.fext.
// Called with m.z holding the CONTINUOUS state (xc,qc,vc,wc) when `pre` (PID reads it), and again
// with the knot state (x2,q2,v1,w1) for the laws that read body.state.*sol[2] (lqr.jl:98-103).
template <int G>
__device__ __forceinline__ void control_pre(const DevMech& M, const DevCtrl& C, const SysMem& m, int lane, bool active, int sys, int k) {
    if (active) for (int e = lane; e < M.ne; e += G) m.u[e] = 0.0;
    __syncwarp();
    if (C.kind == CCB_CTRL_PID && active) {
        // reference src/control/pid.jl:69-88; one lane walks the controlled eqcs in order (N <= a handful)
        if (lane == 0) {
            for (int i = 0; i < C.nu; i++) {
                const int e = __ldg(C.eqc + i), j = __ldg(M.eqc_act_joint + e), kind = __ldg(M.eqc_act_kind + e);
                const int a = __ldg(M.j_parent + j), b = __ldg(M.j_child + j);
                V3 xa = {0.0, 0.0, 0.0};
                Q4 qa = {1.0, 0.0, 0.0, 0.0};
                if (a >= 0) { xa = ld3(m.z + 13 * a); qa = ld4(m.z + 13 * a + 3); }
                V3 xb = ld3(m.z + 13 * b);
                Q4 qb = ld4(m.z + 13 * b + 3);
                V3 ax = ldg3(M.j_axis + 3 * j);
                double coord;
                bool revolute = false;
                if (kind == 1) {   // minimalCoordinates of a translational DoF: axis . g_T
                    V3 w = xb + qrot(qb, ldg3(M.j_p2 + 3 * j)) - xa;
                    coord = dot(ax, qrotT(qa, w) - ldg3(M.j_p1 + 3 * j));
                } else {           // rotational DoF: axis . rotation_vector(qa^-1 qb qoff^-1)
                    Q4 qe = qmul(qmul(qconj(qa), qb), qconj(ldg4(M.j_qoff + 4 * j)));
                    double nv = sqrt(qe.x * qe.x + qe.y * qe.y + qe.z * qe.z);
                    coord = (nv == 0.0) ? 0.0 : 2.0 * atan2(nv, qe.s) * (ax.x * qe.x + ax.y * qe.y + ax.z * qe.z) / nv;
                    revolute = true;   // Translational3 + Rotational2 (pid.jl:44)
                }
                double err = __ldg(C.goal + (size_t)sys * C.goal_bs + i) - coord;
                if (revolute) {       // pid.jl:46-53
                    const double pi = 3.14159265358979323846;
                    if (err > pi) err -= 2.0 * pi;
                    else if (err < -pi) err += 2.0 * pi;
                }
                double* integ = C.pid_int + (size_t)sys * C.nu + i;
                double* last = C.pid_last + (size_t)sys * C.nu + i;
                if (k == 1) *last = err;                 // pid.jl:73
                const double in = *integ + err * M.dt;   // pid.jl:76
                *integ = in;
                const double der = (err - *last) / M.dt; // pid.jl:77
                m.u[e] += __ldg(C.P + i) * err + __ldg(C.I + i) * in + __ldg(C.D + i) * der;
                *last = err;
            }
        }
    } else if (C.kind == CCB_CTRL_TABLE && active) {
        if (lane == 0 && k >= 1 && k <= C.N)
            for (int i = 0; i < C.nu; i++) m.u[__ldg(C.eqc + i)] += __ldg(C.U + (size_t)sys * C.U_bs + (size_t)(k - 1) * C.nu + i);
    }
    __syncwarp();
}

template <int G>
__device__ __forceinline__ void control_post(const DevMech& M, const DevCtrl& C, const SysMem& m, int lane, bool active, int sys, int k) {
    const int n = 12 * M.nb;
    bool lqr = false;
    const int csys = C.sys_div > 1 ? sys / C.sys_div : sys;   // several systems may share one table (ccb_ctrl_set_replication)
    const double *K = nullptr, *zd = nullptr, *Fd = nullptr;
    if (C.kind == CCB_CTRL_LQR_INF) {            // lqr.jl:116-139
        lqr = true;
        K = C.K + (size_t)csys * C.K_bs; zd = C.zd + (size_t)csys * C.zd_bs; Fd = C.Fd + (size_t)csys * C.Fd_bs;
    } else if (C.kind == CCB_CTRL_LQR_FIN) {     // lqr.jl:89-114, gated by k < N (:106)
        lqr = k < C.N;
        K = C.K + (size_t)csys * C.K_bs + (size_t)(k - 1) * C.nu * n; zd = C.zd + (size_t)csys * C.zd_bs; Fd = C.Fd + (size_t)csys * C.Fd_bs;
    } else if (C.kind == CCB_CTRL_TRACKING) {    // lqr_tracking.jl:46-71, gated by k < N (:63)
        lqr = k < C.N;
        K = C.K + (size_t)csys * C.K_bs + (size_t)(k - 1) * C.nu * n; zd = C.zd + (size_t)csys * C.zd_bs + (size_t)(k - 1) * 13 * M.nb;
        Fd = C.Fd + (size_t)csys * C.Fd_bs + (size_t)(k - 1) * C.nu;
    }
    if (lqr) {
        if (active) {
            for (int b = lane; b < M.nb; b += G) {   // Δz = [x2 - xd; v1 - vd; imag(qd \ q2); w1 - wd]   lqr.jl:98-103
                const double* zb = m.z + 13 * b;
                const double* d = zd + 13 * b;
                double* o = m.dz + 12 * b;
                Q4 qe = qmul(qconj(ldg4(d + 3)), ld4(zb + 3));
                for (int i = 0; i < 3; i++) { o[i] = zb[i] - __ldg(d + i); o[3 + i] = zb[7 + i] - __ldg(d + 7 + i); o[9 + i] = zb[10 + i] - __ldg(d + 10 + i); }
                o[6] = qe.x; o[7] = qe.y; o[8] = qe.z;
            }
        }
        __syncwarp();
        for (int i = 0; i < C.nu; i++) {            // u_i = Fτd[i] - K[k][i] * Δz   lqr.jl:108
            double acc = 0.0;
            if (active) for (int c = lane; c < n; c += G) acc = fma(__ldg(K + (size_t)i * n + c), m.dz[c], acc);
            acc = group_sum<G>(acc);
            if (active && lane == 0) m.u[__ldg(C.eqc + i)] += __ldg(Fd + i) - acc;
        }
        __syncwarp();
    }
    // applyFτ!: joint input -> body wrenches, gathered per body so that the summation order is fixed
    if (active) {
        for (int b = lane; b < M.nb; b += G) {
            V3 F = {0.0, 0.0, 0.0}, T = {0.0, 0.0, 0.0};
            if (C.kind != CCB_CTRL_NONE) {
                for (int i = 0; i < C.nu; i++) {
                    const int e = __ldg(C.eqc + i), j = __ldg(M.eqc_act_joint + e);
                    const int pa = __ldg(M.j_parent + j), ch = __ldg(M.j_child + j);
                    if (pa != b && ch != b) continue;
                    const double u = m.u[e];
                    V3 ax = u * ldg3(M.j_axis + 3 * j);
                    Q4 qa = {1.0, 0.0, 0.0, 0.0};
                    if (pa >= 0) qa = ld4(m.z + 13 * pa + 3);
                    Q4 qb = ld4(m.z + 13 * ch + 3);
                    if (__ldg(M.eqc_act_kind + e) == 1) {
                        V3 Fb = qrot(qa, ax);
                        if (pa == b) { F = F - Fb; T = T - cross(ldg3(M.j_p1 + 3 * j), ax); }
                        if (ch == b) { F = F + Fb; T = T + cross(ldg3(M.j_p2 + 3 * j), qrotT(qb, Fb)); }
                    } else {
                        if (pa == b) T = T - ax;
                        if (ch == b) T = T + qrotT(qb, qrot(qa, ax));
                    }
                }
            }
            st3(m.fext + 6 * b, F);
            st3(m.fext + 6 * b + 3, T);
        }
    }
    __syncwarp();
}

template <int G> __global__ void __launch_bounds__(32) rollout_kernel(const DevMech M, const DevCtrl C, const SimArgs A) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x % G, grp = threadIdx.x / G;
    const int sys = blockIdx.x * (32 / G) + grp;
    const bool valid = sys < A.B;
    const SysMem m = carve(smem + (size_t)grp * M.smem_doubles, M);
    const int nst = 13 * M.nb;
    const double dt = M.dt, c4 = 4.0 / (M.dt * M.dt);

    if (valid) {
        for (int i = lane; i < nst; i += G) m.z[i] = A.state[(size_t)sys * nst + i];
        for (int i = lane; i < M.ml; i += G) m.lam[i] = A.lambda[(size_t)sys * M.ml + i];
    }
    __syncwarp();
    int total_iters = 0;
    bool failed = false;

    for (int step = 0; step < A.steps; step++) {
        const int k = A.k0 + step;
        if (valid && A.storage) {   // saveToStorage!: (xc,qc,vc,wc) at the start of step k
            double* out = A.storage + ((size_t)step * A.B + sys) * nst;
            for (int i = lane; i < nst; i += G) out[i] = m.z[i];
        }
        control_pre<G>(M, C, m, lane, valid, sys, k);
        if (valid) {                // discretizestate!/updatestate!: knot x2 = xc + vc dt, q2 = qc (x) w-bar(wc)
            for (int b = lane; b < M.nb; b += G) {
                double* zb = m.z + 13 * b;
                V3 v = ld3(zb + 7), w = ld3(zb + 10);
                if (!A.knot_input) {
                    st3(zb, ld3(zb) + dt * v);
                    double sq = sqrt(c4 - dot(w, w));
                    st4(zb + 3, qmul(ld4(zb + 3), Q4{0.5 * dt * sq, 0.5 * dt * w.x, 0.5 * dt * w.y, 0.5 * dt * w.z}));
                }
                for (int i = 0; i < 6; i++) m.s[6 * b + i] = zb[7 + i];   // Newton starts from (v1, w1)
            }
        }
        __syncwarp();
        control_post<G>(M, C, m, lane, valid, sys, k);
        if (valid) {                // Jacobian rows at the knot (x2,q2): they carry the constraint forces, fixed for the step
            for (int r = lane; r < M.ml; r += G) {
                const int j = __ldg(M.row_joint + r);
                const int a = __ldg(M.j_parent + j), b = __ldg(M.j_child + j);
                V3 xa = {0.0, 0.0, 0.0};
                Q4 qa = {1.0, 0.0, 0.0, 0.0};
                if (a >= 0) { xa = ld3(m.z + 13 * a); qa = ld4(m.z + 13 * a + 3); }
                double g;
                V3 Gax, Gap, Gbx, Gbp;
                row_eval(__ldg(M.row_kind + r), ldg3(M.row_proj + 3 * r), ldg3(M.j_p1 + 3 * j), ldg3(M.j_p2 + 3 * j), ldg4(M.j_qoff + 4 * j), xa, qa,
                         ld3(m.z + 13 * b), ld4(m.z + 13 * b + 3), true, g, Gax, Gap, Gbx, Gbp);
                double* o = m.gk + 12 * r;
                if (a < 0) { Gax = v3(0.0, 0.0, 0.0); Gap = Gax; }
                st3(o, Gax); st3(o + 3, Gap); st3(o + 6, Gbx); st3(o + 9, Gbp);
            }
        }
        __syncwarp();

        // ---- newton! with lineSearch!
        bool done = !valid;
        double normf0 = 0.0;
        int it_used = 0;
        for (int n = 1; n <= A.newton_iter; n++) {
            if (__all_sync(0xffffffffu, done)) break;
            const bool act = !done;
            double nf = sqrt(group_sum<G>(eval_system<G, true>(M, m, lane, act)));
            if (n == 1) normf0 = nf;
            solve_system<G>(M, m, lane, act);
            // full-step size; keep the step in ds / dl (the residual evaluations reuse d / rhs)
            double sd = 0.0;
            if (act) {
                for (int i = lane; i < 6 * M.nb; i += G) { double v = m.ds[i]; sd = fma(v, v, sd); m.s0[i] = m.s[i]; }
                for (int i = lane; i < M.ml; i += G) { double v = m.rhs[i]; sd = fma(v, v, sd); m.dl[i] = v; m.lam0[i] = m.lam[i]; }
            }
            const double full = sqrt(group_sum<G>(sd));
            // a step below eps taken from a residual below eps ends the solve whatever scale the line search would pick
            const bool tiny = act && full < A.eps && normf0 < A.eps;
            double alpha = 1.0, normf1 = normf0;
            bool ls_done = !act || tiny;
            if (tiny) {
                for (int i = lane; i < 6 * M.nb; i += G) m.s[i] = m.s0[i] - m.ds[i];
                for (int i = lane; i < M.ml; i += G) m.lam[i] = m.lam0[i] - m.dl[i];
            }
            __syncwarp();
            for (int ls = 0; ls <= A.line_iter; ls++) {    // lineSearch!: s = s_prev - 2^-j step, first j with |f| decreased
                if (__all_sync(0xffffffffu, ls_done)) break;
                const bool la = !ls_done;
                if (la) {
                    for (int i = lane; i < 6 * M.nb; i += G) m.s[i] = fma(-alpha, m.ds[i], m.s0[i]);
                    for (int i = lane; i < M.ml; i += G) m.lam[i] = fma(-alpha, m.dl[i], m.lam0[i]);
                }
                __syncwarp();
                const double nf1 = sqrt(group_sum<G>(eval_system<G, false>(M, m, lane, la)));
                if (la) {
                    normf1 = nf1;
                    if (nf1 >= normf0 && ls < A.line_iter) alpha *= 0.5;
                    else ls_done = true;
                }
            }
            if (act) {
                it_used = n;
                const double nd = alpha * full;
                if (tiny || (nd < A.eps && normf1 < A.eps)) done = true;
                if (!(normf1 == normf1)) { done = true; failed = true; }   // NaN residual: no iterate can recover, the step has failed
                normf0 = normf1;
            }
        }
        if (valid) {
            total_iters += it_used;
            if (!done) failed = true;
            for (int b = lane; b < M.nb; b += G)    // updatestate!: (xc,qc,vc,wc) <- (x2,q2,v2,w2)
                for (int i = 0; i < 6; i++) m.z[13 * b + 7 + i] = m.s[6 * b + i];
        }
        __syncwarp();
    }
    if (valid && A.knot_input) {
        for (int i = lane; i < 6 * M.nb; i += G) A.s_out[(size_t)sys * 6 * M.nb + i] = m.s[i];
        for (int i = lane; i < M.ml; i += G) A.lambda[(size_t)sys * M.ml + i] = m.lam[i];
        if (A.status && lane == 0) write_status(A, sys, total_iters, failed);
    } else if (valid) {
        for (int i = lane; i < nst; i += G) A.state[(size_t)sys * nst + i] = m.z[i];
        for (int i = lane; i < M.ml; i += G) A.lambda[(size_t)sys * M.ml + i] = m.lam[i];
        if (A.status && lane == 0) write_status(A, sys, total_iters, failed);
    }
}

}  // namespace

int rollout_smem_bytes(const DevMech& M, int lanes) { return (32 / lanes) * M.smem_doubles * (int)sizeof(double); }

int launch_rollout(const DevMech& M, const DevCtrl& C, const SimArgs& a, int lanes, int default_lanes, cudaStream_t stream) {
    if ((lanes == 0 || lanes == 5) && M.rows_ok) return launch_rollout_rows(M, C, a, stream);
    if (lanes == 5) { set_error("lanes = 5 (row-per-lane kernel) needs a single serial chain of <= 6 joints of dimension <= 5"); return CCB_E_INVALID; }
    if ((lanes == 0 || lanes == 6) && mw_kernel_eligible(M)) return launch_rollout_mw(M, C, a, stream);
    if (lanes == 6) { set_error("lanes = 6 (multi-warp rows kernel) needs a single serial chain with joints of dimension <= 5"); return CCB_E_INVALID; }
    if ((lanes == 0 || lanes == 1) && M.is_chain && M.nj <= 32)
        return launch_rollout_chain(M, C, a, M.chain_parent, M.chain_child, M.chain_depth, M.chain_max_depth, M.chain_max_dim, stream);
    if (lanes == 1) { set_error("lanes = 1 (chain kernel) requested for a mechanism that is not a serial chain"); return CCB_E_INVALID; }
    if (lanes == 0) lanes = default_lanes;
    if (lanes != 8 && lanes != 16 && lanes != 32) { set_error("lanes per system must be 0 (auto), 5 (row-per-lane kernel), 6 (multi-warp rows kernel), 1 (chain kernel), 8, 16 or 32"); return CCB_E_INVALID; }
    const int per_cta = 32 / lanes;
    const int grid = (a.B + per_cta - 1) / per_cta;
    const size_t smem = (size_t)rollout_smem_bytes(M, lanes);
    if (smem > 227 * 1024) { set_error("mechanism too large for the shared-memory rollout kernel"); return CCB_E_UNSUPPORTED; }
#define CCB_LAUNCH(GL)                                                                                                   \
    do {                                                                                                                 \
        CCB_CUDA(cudaFuncSetAttribute(rollout_kernel<GL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        rollout_kernel<GL><<<grid, 32, smem, stream>>>(M, C, a);                                                         \
    } while (0)
    if (lanes == 8) CCB_LAUNCH(8);
    else if (lanes == 16) CCB_LAUNCH(16);
    else CCB_LAUNCH(32);
#undef CCB_LAUNCH
    g_launches++;
    CCB_CUDA(cudaGetLastError());
    return CCB_OK;
}

}  // namespace ccb
