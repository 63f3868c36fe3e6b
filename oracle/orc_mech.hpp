// oracle/orc_mech.hpp -- TEST INFRASTRUCTURE ONLY (CPU oracle). Not part of the shipped product path.
//
// CPU restatement of the parts of the hot path that live in the un-vendored dependency
// ConstrainedDynamics.jl ^0.9.1 (reference Project.toml:7,12) / GraphBasedSystems.jl, following the
// published algorithm (README.md:14,16 -> arXiv 2002.11245, 2010.05886) as re-specified in
// SURVEY.md Appendix A.  PARITY UNPINNED by the reference's own tests (all `@test true`,
// test/examples/*.jl); pinned instead by the known-answer tests in tests/test_oracle_*.py.
//
// Reference call sites that this file serves:
//   simulate!        examples/lqr_pendulum.jl:43, examples/trackingLQR_triple_cartpole.jl:53,149
//   setForce!        src/control/lqr.jl:109,135  src/control/lqr_tracking.jl:66  src/control/pid.jl:84
//   linearsystem     src/control/lqr.jl:63       src/control/lqr_tracking.jl:88
//   minimalCoordinates src/control/pid.jl:45,55
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>
#include "orc_math.hpp"

namespace orc {

struct Joint {
    int eqc;          // owning equality constraint (0-based)
    int parent;       // body index, -1 = origin
    int child;        // body index
    int nT, nR;       // constrained translational / rotational rows (0..3)
    int dim;          // nT + nR
    int row0;         // offset of this joint's rows in the global lambda vector
    double PT[3][3];  // row projector, first nT rows valid (I, V12 or V3)
    double PR[3][3];
    double axis[3];   // normalised free axis (nullspacemat, pid.jl:44 / setForce!)
    double p1[3], p2[3], qoff[4];
};

struct Mech {
    int nb = 0, ne = 0, nj = 0, ml = 0;
    double dt = 0.01, g = -9.81, reg = 0.0;
    std::vector<double> m;        // [nb]
    std::vector<double> J;        // [nb][9] row-major
    std::vector<Joint> joints;    // [nj]
    std::vector<int> eqc_parent;  // [ne]
    std::vector<int> eqc_ptr;     // [ne+1]
    std::vector<int> eqc_dim;     // [ne]
    std::vector<int> eqc_row0;    // [ne]
    std::vector<int> act_joint;   // [ne] joint index carrying the 1-DoF input, or -1
    std::vector<int> act_kind;    // [ne] 0 none, 1 translational force, 2 rotational torque
    // graph (restatement-defined tie-breaking: ascending ids, post-order)
    int nnodes = 0;
    std::vector<int> eqc_id, body_id;     // 1-based node ids (eqcs first)
    std::vector<int> dfs_list, parents;   // dfs_list: children before parents; parents[id-1] = id or 0
    std::vector<int> loop_edges;          // pairs (id_a,id_b) of non-tree edges
    // symbolic block pattern for the mixed-node LDU (with fill for loops)
    std::vector<int> node_dim, node_off;  // per node id-1: dimension and offset in the KKT vector
    std::vector<int> blk;                 // [nnodes*nnodes] offset into block pool or -1
    int pool = 0;
};

// orthonormal completion of an axis; restatement-defined (upstream takes it from an SVD, SURVEY A.3;
// every quantity the parity contract names is invariant to this choice)
inline void axis_basis(const double a_in[3], double V12[2][3], double V3_[3]) {
    double n = std::sqrt(a_in[0] * a_in[0] + a_in[1] * a_in[1] + a_in[2] * a_in[2]);
    double a[3] = {1, 0, 0};
    if (n > 0) { a[0] = a_in[0] / n; a[1] = a_in[1] / n; a[2] = a_in[2] / n; }
    V3_[0] = a[0]; V3_[1] = a[1]; V3_[2] = a[2];
    double h[3] = {0, 0, 0};
    if (std::fabs(a[0]) < 0.9) h[0] = 1; else h[1] = 1;
    double hd = h[0] * a[0] + h[1] * a[1] + h[2] * a[2];
    double v1[3] = {h[0] - hd * a[0], h[1] - hd * a[1], h[2] - hd * a[2]};
    double n1 = std::sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
    for (int i = 0; i < 3; i++) v1[i] /= n1;
    double v2[3] = {a[1] * v1[2] - a[2] * v1[1], a[2] * v1[0] - a[0] * v1[2], a[0] * v1[1] - a[1] * v1[0]};
    for (int i = 0; i < 3; i++) { V12[0][i] = v1[i]; V12[1][i] = v2[i]; }
}

inline void set_projector(int n, const double axis[3], double P[3][3]) {
    double V12[2][3], V3_[3];
    axis_basis(axis, V12, V3_);
    std::memset(P, 0, sizeof(double) * 9);
    if (n == 3) { P[0][0] = P[1][1] = P[2][2] = 1; }
    else if (n == 2) { for (int i = 0; i < 3; i++) { P[0][i] = V12[0][i]; P[1][i] = V12[1][i]; } }
    else if (n == 1) { for (int i = 0; i < 3; i++) P[0][i] = V3_[i]; }
}

template <class T> struct Pose { V3<T> x; Q4<T> q; };
template <class T> inline Pose<T> origin_pose() { return {{T(0.0), T(0.0), T(0.0)}, {T(1.0), T(0.0), T(0.0), T(0.0)}}; }

// ---------------------------------------------------------------- joints
// g_T = R_a^T (x_b + R_b p2 - x_a) - p1 ; g_R = vec(q_a^-1 q_b q_off^-1); rows projected by PT / PR.
template <class T> void joint_g(const Joint& J, const Pose<T>& a, const Pose<T>& b, T* out) {
    if (J.nT > 0) {
        V3<T> w = b.x + qrot(b.q, lift3<T>(J.p2)) - a.x;
        V3<T> gT = qrotT(a.q, w) - lift3<T>(J.p1);
        for (int r = 0; r < J.nT; r++) out[r] = T(J.PT[r][0]) * gT.x + T(J.PT[r][1]) * gT.y + T(J.PT[r][2]) * gT.z;
    }
    if (J.nR > 0) {
        Q4<T> qe = qmul(qmul(qconj(a.q), b.q), qconj(lift4<T>(J.qoff)));
        for (int r = 0; r < J.nR; r++) out[J.nT + r] = T(J.PR[r][0]) * qe.x + T(J.PR[r][1]) * qe.y + T(J.PR[r][2]) * qe.z;
    }
}

// Tangent ("rotational") Jacobians G_a, G_b = d g / d(x, phi), q <- q (x) [1; phi]   (dim x 6 each, row-major)
template <class T> void joint_G(const Joint& J, const Pose<T>& a, const Pose<T>& b, T* Ga, T* Gb) {
    T Ta[3][6], Tb[3][6], Ra[3][6], Rb[3][6];
    if (J.nT > 0) {
        V3<T> p2 = lift3<T>(J.p2);
        V3<T> w = b.x + qrot(b.q, p2) - a.x;
        V3<T> raw = qrotT(a.q, w);
        for (int j = 0; j < 3; j++) {
            V3<T> e = {T(j == 0 ? 1.0 : 0.0), T(j == 1 ? 1.0 : 0.0), T(j == 2 ? 1.0 : 0.0)};
            V3<T> c = qrotT(a.q, e);                                       // column j of R_a^T
            Tb[0][j] = c.x; Tb[1][j] = c.y; Tb[2][j] = c.z;
            Ta[0][j] = -c.x; Ta[1][j] = -c.y; Ta[2][j] = -c.z;
            V3<T> cb = qrotT(a.q, qrot(b.q, T(-2.0) * cross(p2, e)));      // -2 R_a^T R_b [p2]x e_j
            Tb[0][3 + j] = cb.x; Tb[1][3 + j] = cb.y; Tb[2][3 + j] = cb.z;
            V3<T> ca = T(2.0) * cross(raw, e);                             // 2 [R_a^T w]x e_j
            Ta[0][3 + j] = ca.x; Ta[1][3 + j] = ca.y; Ta[2][3 + j] = ca.z;
        }
    }
    if (J.nR > 0) {
        Q4<T> qoi = qconj(lift4<T>(J.qoff));
        Q4<T> qab = qmul(qconj(a.q), b.q);
        Q4<T> qe = qmul(qab, qoi);
        for (int j = 0; j < 3; j++) {
            Q4<T> e = {T(0.0), T(j == 0 ? 1.0 : 0.0), T(j == 1 ? 1.0 : 0.0), T(j == 2 ? 1.0 : 0.0)};
            Q4<T> cb = qmul(qmul(qab, e), qoi);   // V L(qab) R(qoff^-1) V^T e_j
            Q4<T> ca = qmul(e, qe);               // V R(qe) V^T e_j  (negated below)
            Rb[0][j] = Rb[1][j] = Rb[2][j] = T(0.0);
            Ra[0][j] = Ra[1][j] = Ra[2][j] = T(0.0);
            Rb[0][3 + j] = cb.x; Rb[1][3 + j] = cb.y; Rb[2][3 + j] = cb.z;
            Ra[0][3 + j] = -ca.x; Ra[1][3 + j] = -ca.y; Ra[2][3 + j] = -ca.z;
        }
    }
    for (int r = 0; r < J.nT; r++)
        for (int c = 0; c < 6; c++) {
            Ga[r * 6 + c] = T(J.PT[r][0]) * Ta[0][c] + T(J.PT[r][1]) * Ta[1][c] + T(J.PT[r][2]) * Ta[2][c];
            Gb[r * 6 + c] = T(J.PT[r][0]) * Tb[0][c] + T(J.PT[r][1]) * Tb[1][c] + T(J.PT[r][2]) * Tb[2][c];
        }
    for (int r = 0; r < J.nR; r++)
        for (int c = 0; c < 6; c++) {
            Ga[(J.nT + r) * 6 + c] = T(J.PR[r][0]) * Ra[0][c] + T(J.PR[r][1]) * Ra[1][c] + T(J.PR[r][2]) * Ra[2][c];
            Gb[(J.nT + r) * 6 + c] = T(J.PR[r][0]) * Rb[0][c] + T(J.PR[r][1]) * Rb[1][c] + T(J.PR[r][2]) * Rb[2][c];
        }
}

// ---------------------------------------------------------------- integrator
template <class T> inline void next_pose(const Pose<T>& p2, V3<T> v2, V3<T> w2, double dt, Pose<T>& p3) {
    p3.x = p2.x + T(dt) * v2;
    T sq = sqrt(T(4.0 / (dt * dt)) - dot(w2, w2));
    Q4<T> om = {T(dt / 2) * sq, T(dt / 2) * w2.x, T(dt / 2) * w2.y, T(dt / 2) * w2.z};
    p3.q = qmul(p2.q, om);
}

template <class T> inline V3<T> matvec3(const double* J, V3<T> w) {
    return {T(J[0]) * w.x + T(J[1]) * w.y + T(J[2]) * w.z, T(J[3]) * w.x + T(J[4]) * w.y + T(J[5]) * w.z,
            T(J[6]) * w.x + T(J[7]) * w.y + T(J[8]) * w.z};
}

// d = [d_T; d_R] without constraint forces; F world-frame force, tau body-frame torque
template <class T>
void body_dyn(double m, const double* J, double dt, double g, V3<T> v1, V3<T> w1, V3<T> v2, V3<T> w2, V3<T> F, V3<T> tau, T* d) {
    T idt = T(1.0 / dt);
    d[0] = T(m) * ((v2.x - v1.x) * idt) - F.x;
    d[1] = T(m) * ((v2.y - v1.y) * idt) - F.y;
    d[2] = T(m) * ((v2.z - v1.z) * idt - T(g)) - F.z;
    T sq1 = sqrt(T(4.0 / (dt * dt)) - dot(w1, w1));
    T sq2 = sqrt(T(4.0 / (dt * dt)) - dot(w2, w2));
    V3<T> Jw1 = matvec3(J, w1), Jw2 = matvec3(J, w2);
    V3<T> r = sq2 * Jw2 + cross(w2, Jw2) - sq1 * Jw1 + cross(w1, Jw1) - T(2.0) * tau;
    d[3] = r.x; d[4] = r.y; d[5] = r.z;
}

// 6x6 Jacobian d d / d(v2, w2) (row-major)
template <class T> inline void body_D(double m, const double* J, double dt, const T* w2, T* D) {
    for (int i = 0; i < 36; i++) D[i] = T(0.0);
    D[0] = D[7] = D[14] = T(m / dt);
    T sq2 = sqrt(T(4.0 / (dt * dt)) - (w2[0] * w2[0] + w2[1] * w2[1] + w2[2] * w2[2]));
    T Jw[3] = {T(J[0]) * w2[0] + T(J[1]) * w2[1] + T(J[2]) * w2[2], T(J[3]) * w2[0] + T(J[4]) * w2[1] + T(J[5]) * w2[2],
               T(J[6]) * w2[0] + T(J[7]) * w2[1] + T(J[8]) * w2[2]};
    T W[9] = {T(0.0), -w2[2], w2[1], w2[2], T(0.0), -w2[0], -w2[1], w2[0], T(0.0)};
    T JW[9] = {T(0.0), -Jw[2], Jw[1], Jw[2], T(0.0), -Jw[0], -Jw[1], Jw[0], T(0.0)};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            T wj = T(0.0);
            for (int k = 0; k < 3; k++) wj += W[i * 3 + k] * T(J[k * 3 + j]);
            D[(3 + i) * 6 + 3 + j] = sq2 * T(J[i * 3 + j]) + wj - JW[i * 3 + j] - Jw[i] * w2[j] / sq2;
        }
}

// T_w = V L(Omega)^T dOmega/dw2 = (dt^2/4) (w w^T / sq + sq I - [w]x)   (3x3)
template <class T> inline void T_omega(const T* w, double dt, T* Tm) {
    T sq = sqrt(T(4.0 / (dt * dt)) - (w[0] * w[0] + w[1] * w[1] + w[2] * w[2]));
    T c = T(dt * dt / 4);
    T W[9] = {T(0.0), -w[2], w[1], w[2], T(0.0), -w[0], -w[1], w[0], T(0.0)};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Tm[i * 3 + j] = c * (w[i] * w[j] / sq + (i == j ? sq : T(0.0)) - W[i * 3 + j]);
}

// 1-DoF joint input u -> body wrenches (SURVEY A.3 `setForce!`/`applyFtau!`): F world frame, tau body frame
template <class T>
void joint_input(const Joint& J, int kind, T u, const Pose<T>& a, const Pose<T>& b, V3<T>& Fa, V3<T>& ta, V3<T>& Fb, V3<T>& tb) {
    V3<T> ax = {T(J.axis[0]) * u, T(J.axis[1]) * u, T(J.axis[2]) * u};
    V3<T> z = {T(0.0), T(0.0), T(0.0)};
    if (kind == 1) {            // translational: F given in the parent frame
        Fa = qrot(a.q, -ax);
        Fb = -Fa;
        ta = cross(lift3<T>(J.p1), -ax);
        tb = cross(lift3<T>(J.p2), qrotT(b.q, Fb));
    } else {                    // rotational: tau given in the parent frame
        Fa = z; Fb = z;
        ta = -ax;
        tb = qrotT(b.q, qrot(a.q, ax));
    }
}

template <class T> inline Pose<T> body_pose2(const T* z, int b) {
    const T* p = z + 13 * b;
    return {{p[0], p[1], p[2]}, {p[3], p[4], p[5], p[6]}};
}

// Full KKT residual f = [d_b (6 nb); g_j (ml)].
// z: per body 13 = (x2, q2, v1, w1)   s: per body 6 = (v2, w2)   lam: ml   u: per eqc scalar input
template <class T> void residual(const Mech& M, const T* z, const T* s, const T* lam, const T* u, T* f) {
    const int nb = M.nb;
    std::vector<V3<T>> F(nb, V3<T>{T(0.0), T(0.0), T(0.0)}), tau(nb, V3<T>{T(0.0), T(0.0), T(0.0)});
    for (int e = 0; e < M.ne; e++) {
        if (M.act_kind[e] == 0) continue;
        const Joint& J = M.joints[M.act_joint[e]];
        Pose<T> a = J.parent >= 0 ? body_pose2(z, J.parent) : origin_pose<T>();
        Pose<T> b = body_pose2(z, J.child);
        V3<T> Fa, ta, Fb, tb;
        joint_input(J, M.act_kind[e], u[e], a, b, Fa, ta, Fb, tb);
        if (J.parent >= 0) { F[J.parent] = F[J.parent] + Fa; tau[J.parent] = tau[J.parent] + ta; }
        F[J.child] = F[J.child] + Fb; tau[J.child] = tau[J.child] + tb;
    }
    for (int b = 0; b < nb; b++) {
        const T* p = z + 13 * b;
        const T* sb = s + 6 * b;
        body_dyn<T>(M.m[b], &M.J[9 * b], M.dt, M.g, {p[7], p[8], p[9]}, {p[10], p[11], p[12]}, {sb[0], sb[1], sb[2]},
                    {sb[3], sb[4], sb[5]}, F[b], tau[b], f + 6 * b);
    }
    T Ga[36], Gb[36];
    for (const Joint& J : M.joints) {
        Pose<T> a = J.parent >= 0 ? body_pose2(z, J.parent) : origin_pose<T>();
        Pose<T> b = body_pose2(z, J.child);
        joint_G(J, a, b, Ga, Gb);
        for (int r = 0; r < J.dim; r++)
            for (int c = 0; c < 6; c++) {
                if (J.parent >= 0) f[6 * J.parent + c] -= Ga[r * 6 + c] * lam[J.row0 + r];
                f[6 * J.child + c] -= Gb[r * 6 + c] * lam[J.row0 + r];
            }
    }
    for (const Joint& J : M.joints) {
        Pose<T> a3 = origin_pose<T>(), b3;
        if (J.parent >= 0) {
            const T* sa = s + 6 * J.parent;
            next_pose<T>(body_pose2(z, J.parent), {sa[0], sa[1], sa[2]}, {sa[3], sa[4], sa[5]}, M.dt, a3);
        }
        const T* sb = s + 6 * J.child;
        next_pose<T>(body_pose2(z, J.child), {sb[0], sb[1], sb[2]}, {sb[3], sb[4], sb[5]}, M.dt, b3);
        joint_g(J, a3, b3, f + 6 * nb + J.row0);
    }
}

}  // namespace orc
